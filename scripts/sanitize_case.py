"""A short run of the small parity cases for compute-sanitizer (memcheck / racecheck / synccheck):
every kernel of a step — cell rebuild, list build, list reorder, ring kernel, cull kernel (lists=0),
fused half/full updates, mDBC — on C1 (2D fp64), the 19 k 3D case (fp32) and C5 (mDBC).
  compute-sanitizer --tool memcheck python scripts/sanitize_case.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from sphexample_b200.simulation import Simulation  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for name, mk, opts in (("c1_2d_f64 lists", lambda: util.perturb(util.case_c1("float64"), vel_scale=3.0), {"lists": 1}),
                       ("3d_f32 lists", lambda: util.perturb(util.case_3d_small("float32"), vel_scale=3.0), {"lists": 1}),
                       ("3d_f32 cull", lambda: util.perturb(util.case_3d_small("float32"), vel_scale=3.0), {"lists": 0}),
                       ("c5_mdbc_f64", lambda: util.case_c5("float64"), {})):
    case = mk()
    sim = Simulation(util.params_of(case))
    for k, v in opts.items():
        sim.set_option(k, v)
    sim.set_option("graph", 0)
    sim.upload(case.particles)
    rep = sim.step(steps, reset_delta_x=True)
    st = sim.download(order="id", fields=("Density",))
    print(name, "steps", rep["iteration"], "rebuilds", rep["n_rebuilds"], "list builds", sim.stat("list_builds"),
          "rho mean", float(st["Density"].mean()), flush=True)
    sim.close()
print("sanitize_case done")
