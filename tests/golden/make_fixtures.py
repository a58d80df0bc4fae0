"""Regenerates the input fixtures under tests/golden/ from the reference's shipped CSV exports.

Run in the build container only (needs /root/reference):  python tests/golden/make_fixtures.py
The fixtures are particle LAYOUTS (positions, densities, ids, types, ghost nodes) read through
sphexample_b200.preprocess with the reference loader's semantics (src/PreProcess.jl:12-43,
:217-243); no reference outputs exist to capture because the Julia reference cannot run here.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from sphexample_b200.config import Fixed, Fluid, Geometry  # noqa: E402
from sphexample_b200.preprocess import AllocateDataStructures, LoadMDBCNormals  # noqa: E402

REF = os.environ.get("SPH_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, parts):
    np.savez_compressed(os.path.join(OUT, name), Position=parts.Position, Density=parts.Density,
                        ID=parts.ID, Type=parts.Type, GroupMarker=parts.GroupMarker,
                        GhostPoints=parts.GhostPoints, GhostNormals=parts.GhostNormals)
    print(name, len(parts))


def geom(bound, fluid):
    return [Geometry(os.path.join(REF, "input", bound), 1, Fixed), Geometry(os.path.join(REF, "input", fluid), 2, Fluid)]


if __name__ == "__main__":
    save("dam_break_2d_dp0.02.npz", AllocateDataStructures(
        geom("dam_break_2d/DamBreak2d_Dp0.02_Bound.csv", "dam_break_2d/DamBreak2d_Dp0.02_Fluid.csv"), 2))
    sw = AllocateDataStructures(
        geom("still_wedge/StillWedge_Dp0.02_Bound.csv", "still_wedge/StillWedge_Dp0.02_Fluid.csv"), 2)
    LoadMDBCNormals(sw, os.path.join(REF, "input/still_wedge_mdbc/StillWedge_Dp0.02_GhostNodes_Correct.csv"))
    save("still_wedge_mdbc_dp0.02.npz", sw)
    save("dam_break_3d_dp0.02.npz", AllocateDataStructures(
        geom("dam_break_3d/DamBreak3d_Dp0.02_Bound.csv", "dam_break_3d/DamBreak3d_Dp0.02_Fluid.csv"), 3))
    # the exact upstream 3D case (example/Dambreak3d.jl: dx = 0.0085, N = 171 496)
    save("dam_break_3d_dp0.0085.npz", AllocateDataStructures(
        geom("dam_break_3d/DamBreak3d_Dp0.0085_Bound.csv", "dam_break_3d/DamBreak3d_Dp0.0085_Fluid.csv"), 3))
