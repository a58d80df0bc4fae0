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


def test_vtp_dump_is_well_formed_and_round_trips(tmp_path):
    """the .vtp writer: header parses as XML, every appended block sits where its offset says and holds the data"""
    import re
    import struct
    case = util.case_3d_small("float32")
    st = _state(case)
    st["Acceleration"] = case.particles.Position * 0 + 1.5
    path = str(tmp_path / "p.vtp")
    n = output.write_particles_vtp(path, st)
    raw = open(path, "rb").read()
    head, tail = raw.split(b'<AppendedData encoding="raw">\n_', 1)
    assert n == len(case.particles) and f'NumberOfPoints="{n}"'.encode() in head and tail.endswith(b"\n</AppendedData>\n</VTKFile>\n")
    import xml.etree.ElementTree as ET
    ET.fromstring(head + b'<AppendedData encoding="raw"></AppendedData></VTKFile>')     # well-formed
    offs = {m.group(1).decode(): int(m.group(2)) for m in re.finditer(rb'Name="(\w+)"[^>]*offset="(\d+)"', head)}

    def block(name, dtype):
        o = offs[name]
        size = struct.unpack("<Q", tail[o:o + 8])[0]
        return np.frombuffer(tail[o + 8:o + 8 + size], dtype)
    assert np.array_equal(block("Points", "<f4").reshape(n, 3), st["Position"].astype(np.float32))
    assert np.array_equal(block("Density", "<f4"), st["Density"].astype(np.float32))
    assert np.array_equal(block("ID", "<i8"), st["ID"])
    assert np.array_equal(block("Velocity", "<f4").reshape(n, 3), (st["Velocity"]).astype(np.float32))
    assert np.array_equal(block("offsets", "<i8"), np.arange(1, n + 1))
    # 2D states land in the x and z columns
    c2 = util.case_c1("float64")
    n2 = output.write_particles_vtp(str(tmp_path / "q.vtp"), _state(c2))
    assert n2 == len(c2.particles)
