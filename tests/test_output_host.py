"""The particle dump (sphexample_b200/output.py) round-trips through the reference-semantics loader."""
import numpy as np

import util
from sphexample_b200 import output
from sphexample_b200.config import Fluid
from sphexample_b200.preprocess import LoadSpecificCSV


def _state(case):
    p = case.particles
    return {"Position": p.Position, "Velocity": p.Velocity + 0.25, "Density": p.Density, "Pressure": p.Density * 0 + 7.5,
            "ID": p.ID, "Type": p.Type, "GroupMarker": p.GroupMarker}


def test_dump_round_trips_through_the_loader(tmp_path):
    for case, dim in ((util.case_c1("float64"), 2), (util.case_3d_small("float32"), 3)):
        st = _state(case)
        path = str(tmp_path / f"dump{dim}.csv")
        fluid = st["Type"] == int(Fluid)
        n = output.write_particles_csv(path, st, select=fluid)
        assert n == int(fluid.sum())
        pts, rho, typ, mk, ids = LoadSpecificCSV(dim, np.float64, Fluid, 2, path)
        assert np.array_equal(ids, st["ID"][fluid])                       # Idp + 1 == ID
        assert np.array_equal(pts, st["Position"][fluid].astype(np.float64))   # repr-exact, 2D in columns 0 and 2
        assert np.array_equal(rho, st["Density"][fluid].astype(np.float64))
        assert typ.tolist() == [int(Fluid)] * n and mk.tolist() == [2] * n


def test_header_is_the_dualsphysics_layout(tmp_path):
    st = _state(util.case_c1("float64"))
    path = str(tmp_path / "d.csv")
    output.write_particles_csv(path, st)
    head = open(path).readline().strip()
    assert head == '"Idp","Vel:0","Vel:1","Vel:2","Rhop","Press","Type","Mk","Points:0","Points:1","Points:2"'
