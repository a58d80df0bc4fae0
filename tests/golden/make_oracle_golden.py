"""Regression pins of the ORACLE itself: its outputs on the reference's shipped layouts at a fixed
step, committed so that a later edit of oracle/sph_oracle.cpp (or of the fixtures / constants) cannot
drift unnoticed.  These are NOT reference-produced vectors (Julia cannot run here);
they freeze the restatement that the GPU is compared with.
  python tests/golden/make_oracle_golden.py      (rewrites tests/golden/oracle_*.npz)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import util  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = {"c1_2d": (lambda: util.case_c1("float64"), 50), "3d_small": (lambda: util.case_3d_small("float64"), 20),
         "c5_mdbc": (lambda: util.case_c5("float64"), 50)}


def run(name):
    mk, steps = CASES[name]
    case = mk()                                   # unperturbed: exactly the shipped layout, from rest
    p = util.params_of(case)
    o = orc.Oracle(p, case.particles, nthreads=1)
    o.step(steps, True)
    ids = o.ids
    order = np.argsort(ids, kind="stable")
    pick = order[:: max(1, len(ids) // 256)]      # every k-th particle by ID
    rep = o.report()
    return dict(steps=steps, ids=ids[pick], rho=o.get("rho")[pick], vel=o.get("vel")[pick], pos=o.get("pos")[pick],
                press=o.get("press")[pick], total_time=rep["total_time"], n_rebuilds=rep["n_rebuilds"],
                sum_rho=o.get("rho").sum(), sum_v2=(o.get("vel") ** 2).sum())


if __name__ == "__main__":
    orc.build()
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, f"oracle_{name}.npz"), **run(name))
        print("wrote", name)
