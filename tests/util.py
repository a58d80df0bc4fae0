"""Shared helpers of the parity tests: fixtures -> cases, oracle runs, comparisons."""
import os

import numpy as np

from sphexample_b200 import _abi, cases, config, make_params
from sphexample_b200.preprocess import SimParticles, make_particles

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name, dtype=np.float64) -> SimParticles:
    z = np.load(os.path.join(GOLDEN, name))
    p = make_particles(z["Position"], z["Density"], z["Type"], z["GroupMarker"], z["ID"], dtype=dtype)
    # make_particles sorts by ID; ghost arrays in the fixture are already in that order
    p.GhostPoints[:] = z["GhostPoints"]
    p.GhostNormals[:] = z["GhostNormals"]
    return p


def case_c1(float_type="float64"):
    """C1: shipped 2D dam break, N = 6 881"""
    dt = np.float64 if float_type == "float64" else np.float32
    return cases.case_dam_break_2d(0.02, float_type, particles=load_fixture("dam_break_2d_dp0.02.npz", dt))


def case_3d_small(float_type="float64"):
    """shipped 3D dam break at Dp 0.02 (N ~ 19 k): the 3D parity case the oracle finishes in seconds"""
    dt = np.float64 if float_type == "float64" else np.float32
    return cases.case_dam_break_3d(0.02, float_type, particles=load_fixture("dam_break_3d_dp0.02.npz", dt))


def case_c5(float_type="float64"):
    dt = np.float64 if float_type == "float64" else np.float32
    return cases.case_still_wedge_mdbc(load_fixture("still_wedge_mdbc_dp0.02.npz", dt), float_type)


def params_of(case, geometry=()):
    return make_params(case.meta, case.consts, case.kernel, case.viscosity, case.diffusion, geometry)


def perturb(case, seed=0, vel_scale=0.5, rho_scale=2.0, jitter=0.1):
    """Deterministic non-trivial state: jittered fluid positions, smooth velocity field, density
    noise — so that every pair term (viscosity branch, diffusion, continuity) is exercised."""
    rng = np.random.default_rng(seed)
    p = case.particles
    dt = p.Position.dtype
    fluid = p.Type == 1
    dp = case.consts.dx
    pos = p.Position.astype(np.float64).copy()
    pos[fluid] += rng.uniform(-jitter, jitter, pos[fluid].shape) * dp
    x = pos
    vel = np.zeros_like(x)
    vel[:, 0] = vel_scale * np.sin(3.0 * x[:, -1] + 1.0) * fluid
    vel[:, -1] = -vel_scale * np.cos(2.0 * x[:, 0]) * fluid
    if x.shape[1] == 3:
        vel[:, 1] = 0.3 * vel_scale * np.sin(5.0 * x[:, 0] + 2.0 * x[:, 2]) * fluid
    rho = p.Density.astype(np.float64) + rho_scale * rng.uniform(0, 1, len(p))
    p.Position[:] = pos.astype(dt)
    p.Velocity[:] = vel.astype(dt)
    p.Density[:] = rho.astype(dt)
    return case


def by_id(ids, arr):
    out = np.empty_like(arr)
    order = np.argsort(ids, kind="stable")
    return arr[order]


def relerr(a, b):
    """max |a - b| / max |b|  (the SURVEY's 'x * scale' tolerance)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))


def check(err, tol):
    """assert err < tol, and log the measured value (SPH_PARITY_LOG=<file>) so that one GPU run
    reports every margin, not just the first failure"""
    import inspect
    import json
    fr = inspect.stack()[1]
    test = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
    log = os.environ.get("SPH_PARITY_LOG")
    if log:
        with open(log, "a") as fh:
            fh.write(json.dumps({"test": test, "line": fr.lineno, "err": err, "tol": tol, "ok": bool(err < tol)}) + "\n")
    assert err < tol, f"{test}:{fr.lineno}: measured {err:.3e} >= tolerance {tol:.1e}"


def case_3d_shipped(float_type="float32"):
    """the exact upstream 3D case: input/dam_break_3d Dp 0.0085 (N = 171 496), example/Dambreak3d.jl constants"""
    dt = np.float64 if float_type == "float64" else np.float32
    return cases.case_dam_break_3d(0.0085, float_type, particles=load_fixture("dam_break_3d_dp0.0085.npz", dt))


def case_c2(float_type="float64"):
    """C2: the 2D dam break regenerated at dp = 0.0058 (N = 59 909)"""
    return cases.case_dam_break_2d(0.0058, float_type)


def set_cubic_spline(p):
    """switch a parameter block to CubicSpline (+ its αD, src/SPHKernels.jl:24-26, and the tensile correction's ε)"""
    p.kernel = _abi.KERNEL_CUBICSPLINE
    p.alphaD = {2: 10.0 / (7.0 * np.pi * p.h ** 2), 3: 1.0 / (np.pi * p.h ** 3)}[int(p.dim)]
    p.cubic_eps = 0.2
