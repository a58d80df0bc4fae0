"""CPU check of the PRODUCT's pair/integrator formulas (sphexample_b200/csrc/sph_physics.cuh,
compiled as host C++ through tests/physics_shim.cpp) against the oracle — no GPU involved.
The traversal here is all-pairs; the device traversal is covered by the -m gpu parity tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util
from sphexample_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("shim") / "physics_shim.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                           os.path.join(HERE, "physics_shim.cpp"), "-o", out])
    L = C.CDLL(out)
    vp = C.c_void_p
    L.shim_pair_sums.argtypes = [C.POINTER(_abi.Params), C.c_int] + [vp] * 8 + [C.c_int, C.c_int, vp, vp, vp]
    L.shim_half_full.argtypes = [C.POINTER(_abi.Params), C.c_int] + [vp] * 7 + [C.c_double, vp, vp, vp, C.c_int]
    L.shim_eos.restype = C.c_double
    L.shim_eos.argtypes = [C.POINTER(_abi.Params), C.c_double]
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _roles(cells, n):
    idx = np.arange(n)
    same = np.all(cells[:, None, :] == cells[None, :, :], axis=2)
    return np.ascontiguousarray(np.where(same, idx[:, None] < idx[None, :], idx[:, None] > idx[None, :]).astype(np.uint8))


def _subset(case, n_max, lo, hi):
    p = case.particles
    sel = np.all((p.Position >= lo) & (p.Position <= hi), axis=1)
    return p.permuted(np.nonzero(sel)[0][:n_max])


def _run_case(oracle_lib, shim, dim, tweak, generic, use_float=False, pass2=False):
    if dim == 2:
        case = util.perturb(util.case_c1())
        parts = _subset(case, 900, np.array([0.0, 0.0]), np.array([0.6, 0.45]))
    else:
        case = util.perturb(util.case_3d_small())
        parts = _subset(case, 900, np.array([0.0, 0.0, 0.0]), np.array([0.2, 0.2, 0.16]))
    p = util.params_of(case)
    tweak(p)
    o = oracle_lib.Oracle(p, parts)
    o.update_neighbors()
    o.pressure(0)
    o.neighbor_loop(0)
    n = len(parts)
    if pass2:
        o.half_time_step(2e-5)
        o.pressure(1)
        o.neighbor_loop(1)
        pos, vel, rho = o.get("pos_h"), o.get("vel_h"), o.get("rho_h")
    else:
        pos, vel, rho = o.get("pos"), o.get("vel"), o.get("rho")
    press, rho_n, vel_n = o.get("press"), o.get("rho"), o.get("vel")
    ml = (o.types == 1).astype(np.float64)
    roles = _roles(o.cells, n)
    drho, acc = np.zeros(n), np.zeros((n, dim))
    aux = np.zeros((n, 2 + 2 * dim))
    shim.shim_pair_sums(C.byref(p), n, _ptr(pos), _ptr(vel), _ptr(rho), _ptr(press), _ptr(rho_n), _ptr(vel_n),
                        _ptr(ml), _ptr(roles), int(generic), int(use_float), _ptr(drho), _ptr(acc), _ptr(aux))
    return o, drho, acc, aux


def _set(**kw):
    def f(p):
        for k, v in kw.items():
            setattr(p, k, v)
    return f


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("pass2", [False, True])
def test_fast_pair_body_matches_oracle(oracle_lib, shim, dim, pass2):
    o, drho, acc, _ = _run_case(oracle_lib, shim, dim, _set(), generic=False, pass2=pass2)
    assert util.relerr(drho, o.get("drhodt")) < 1e-12
    assert util.relerr(acc, o.get("acc")) < 1e-12


@pytest.mark.parametrize("use_float,tol", [(False, 1e-12), (True, 2e-4)])
def test_masked_pair_body_equals_the_guarded_one(oracle_lib, shim, use_float, tol):
    """the list kernel's branch-free body (kernel-gradient factor zeroed outside H, self pair and
    far pairs included) sums to the same result as the cut-off-guarded body"""
    case = util.perturb(util.case_3d_small())
    parts = _subset(case, 700, np.array([0.0, 0.0, 0.0]), np.array([0.2, 0.2, 0.16]))
    p = util.params_of(case)
    o = oracle_lib.Oracle(p, parts)
    o.update_neighbors()
    o.pressure(0)
    o.neighbor_loop(0)
    n = len(parts)
    pos, vel, rho, press = o.get("pos"), o.get("vel"), o.get("rho"), o.get("press")
    ml = (o.types == 1).astype(np.float64)
    roles = _roles(o.cells, n)
    drho, acc = np.zeros(n), np.zeros((n, 3))
    shim.shim_pair_sums_masked.argtypes = [C.POINTER(_abi.Params), C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_void_p]
    rc = shim.shim_pair_sums_masked(C.byref(p), n, _ptr(pos), _ptr(vel), _ptr(rho), _ptr(press), _ptr(rho), _ptr(ml),
                                    _ptr(roles), int(use_float), _ptr(drho), _ptr(acc))
    assert rc == 0 and np.all(np.isfinite(drho)) and np.all(np.isfinite(acc))
    assert util.relerr(drho, o.get("drhodt")) < tol
    assert util.relerr(acc, o.get("acc")) < tol


@pytest.mark.parametrize("dim", [2, 3])
def test_fast_pair_body_fp32_tolerance(oracle_lib, shim, dim):
    o, drho, acc, _ = _run_case(oracle_lib, shim, dim, _set(), generic=False, use_float=True)
    assert util.relerr(drho, o.get("drhodt")) < 2e-4   # fp32: x_a - x_b cancellation dominates
    assert util.relerr(acc, o.get("acc")) < 2e-4


MODELS = [
    dict(),                                                   # Artificial + Linear through the generic body
    dict(viscosity=_abi.VISC_LAMINAR, nu0=1e-3),
    dict(viscosity=_abi.VISC_LAMINAR_SPS, nu0=1e-3),
    dict(viscosity=_abi.VISC_ZERO, diffusion=_abi.DDT_ZERO),
    dict(diffusion=_abi.DDT_ZERO_GRAVITY_LINEAR),
    dict(diffusion=_abi.DDT_COMPLEX),
    dict(kernel=_abi.KERNEL_CUBICSPLINE),
    dict(shifting=1, kernel_output=1),
]


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("pass2", [False, True])
def test_generic_pair_body_matches_oracle(oracle_lib, shim, dim, model, pass2):
    def tweak(p):
        _set(**model)(p)
        if p.kernel == _abi.KERNEL_CUBICSPLINE:   # alphaD of the cubic spline, src/SPHKernels.jl:24-26
            p.alphaD = {2: 10 / (7 * np.pi * p.h ** 2), 3: 1 / (np.pi * p.h ** 3)}[dim]
    o, drho, acc, aux = _run_case(oracle_lib, shim, dim, tweak, generic=True, pass2=pass2)
    assert util.relerr(drho, o.get("drhodt")) < 1e-11
    assert util.relerr(acc, o.get("acc")) < 1e-11
    if model.get("shifting"):
        assert util.relerr(aux[:, 0], o.get("divr")) < 1e-11
        assert util.relerr(aux[:, 2:2 + dim], o.get("gradC")) < 1e-11
        assert util.relerr(aux[:, 1], o.get("kern")) < 1e-11
        assert util.relerr(aux[:, 2 + dim:], o.get("kgrad")) < 1e-11


@pytest.mark.parametrize("dim", [2, 3])
def test_integrators_match_oracle(oracle_lib, shim, dim):
    case = util.perturb(util.case_c1() if dim == 2 else util.case_3d_small())
    parts = case.particles.permuted(np.arange(0, len(case.particles), 7))
    p = util.params_of(case)
    o = oracle_lib.Oracle(p, parts)
    o.update_neighbors()
    o.pressure(0)
    o.neighbor_loop(0)
    n = len(parts)
    ty = o.types
    gf = np.where(ty == 1, -1.0, np.where(ty == 3, 1.0, 0.0))
    ml = (ty == 1).astype(np.float64)
    pos, vel, acc, rho, drho = o.get("pos"), o.get("vel"), o.get("acc"), o.get("rho"), o.get("drhodt")
    pos_h, vel_h, rho_h = np.zeros_like(pos), np.zeros_like(vel), np.zeros_like(rho)
    dt = 3e-5
    shim.shim_half_full(C.byref(p), n, _ptr(pos), _ptr(vel), _ptr(acc), _ptr(rho), _ptr(drho), _ptr(gf), _ptr(ml),
                        dt / 2, _ptr(pos_h), _ptr(vel_h), _ptr(rho_h), 0)
    o.half_time_step(dt / 2)
    assert np.array_equal(pos_h, o.get("pos_h")) and np.array_equal(vel_h, o.get("vel_h"))
    assert np.array_equal(rho_h, o.get("rho_h"))
    o.pressure(1)
    o.neighbor_loop(1)
    acc2, drho2 = o.get("acc"), o.get("drhodt")
    shim.shim_half_full(C.byref(p), n, _ptr(pos), _ptr(vel), _ptr(acc2), _ptr(rho), _ptr(drho2), _ptr(gf), _ptr(ml),
                        dt, _ptr(pos_h), _ptr(vel_h), _ptr(rho_h), 1)
    o.full_time_step(dt)
    assert np.array_equal(pos, o.get("pos")) and np.array_equal(vel, o.get("vel"))
    assert np.array_equal(rho, o.get("rho")) and np.array_equal(acc2, o.get("acc"))
    o.pressure(0)
    pr = np.array([shim.shim_eos(C.byref(p), r) for r in rho[:50]])
    assert np.array_equal(pr, o.get("press")[:50])
