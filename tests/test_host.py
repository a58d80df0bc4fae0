"""CPU tests of the host-side mirror of the reference's configuration / pre-processing API."""
import math

import numpy as np
import pytest

from sphexample_b200 import _abi, cases, config, make_params, next_output_time
from sphexample_b200.preprocess import gravity_factor_and_motion_limiter, make_particles

import util


def test_simulation_constants_defaults_and_aliases():
    """src/SimulationConstantsConfiguration.jl:36-52 (note the 2D default m0 = rho0*dx^2)"""
    c = config.SimulationConstants()
    assert c.rho0 == 1000 and c.dx == 0.02 and c.m0 == pytest.approx(0.4)
    assert c.c0 == pytest.approx(math.sqrt(9.81 * 2) * 20)
    assert c.Cb == pytest.approx(c.c0 ** 2 * 1000 / 7) and c.gamma_inv == pytest.approx(1 / 7)
    c2 = config.SimulationConstants(**{"ρ₀": 998.0, "α": 0.1, "δᵩ": 0.2, "c₀": 30.0})
    assert (c2.rho0, c2.alpha, c2.delta_phi, c2.c0) == (998.0, 0.1, 0.2, 30.0)
    with pytest.raises(AssertionError):
        config.SimulationConstants(dx=-1.0)
    with pytest.raises(TypeError):
        config.SimulationConstants(bogus=1)


def test_kernel_instance():
    """src/SPHKernels.jl:20-27,42-72"""
    k = config.SPHKernelInstance(3, config.WendlandC2(), dx=0.0085, k=math.sqrt(3))
    assert k.h == pytest.approx(math.sqrt(3) * 0.0085) and k.H == pytest.approx(math.sqrt(3) * k.h)
    k2 = config.SPHKernelInstance(2, config.WendlandC2(), dx=0.02)
    assert k2.h == pytest.approx(0.04) and k2.H == pytest.approx(0.08) and k2.H2 == pytest.approx(0.0064)
    assert k2.alphaD == pytest.approx(7 / (4 * math.pi * 0.04 ** 2)) and k2.eta2 == pytest.approx((0.01 * 0.04) ** 2)
    with pytest.raises(ValueError):
        config.SPHKernelInstance(2, config.WendlandC2())


def test_gravity_factor_motion_limiter_rules():
    """src/PreProcess.jl:78-98 (Q6): Fluid -1/1, Moving +1/0, Fixed 0/0"""
    gf, ml = gravity_factor_and_motion_limiter(np.array([1, 2, 3], np.uint8))
    assert gf.tolist() == [-1.0, 0.0, 1.0] and ml.tolist() == [1.0, 0.0, 0.0]


def test_generators_reproduce_shipped_counts():
    """the dp = 0.02 2D generator matches the shipped files' counts exactly"""
    p = cases.dam_break_2d(0.02)
    assert len(p) == 6881 and int((p.Type == 1).sum()) == 4416
    shipped = util.load_fixture("dam_break_2d_dp0.02.npz")
    assert len(shipped) == 6881 and int((shipped.Type == 1).sum()) == 4416
    assert abs(cases.dam_break_3d_count(0.0085) - 171496) / 171496 < 0.005
    assert len(cases.dam_break_3d(0.03)) == cases.dam_break_3d_count(0.03)
    dp = cases.dp_for_count_3d(1_000_000)
    assert abs(cases.dam_break_3d_count(dp) - 1_000_000) < 30_000


def test_next_output_time():
    """src/SPHCellList.jl:687-698"""
    m = config.SimulationMetaData(Dimensions=2, OutputTimes=0.01, SimulationTime=1.0)
    assert next_output_time(m) == 0.0
    m.OutputIterationCounter = 3
    assert next_output_time(m) == pytest.approx(0.03)
    m2 = config.SimulationMetaData(Dimensions=2, OutputTimes=[0.1, 0.5, 0.9], SimulationTime=1.0)
    m2.OutputIterationCounter = 2
    assert next_output_time(m2) == 0.5
    m2.OutputIterationCounter = 3
    assert next_output_time(m2) == 1.0


def test_make_params_packs_motion_and_modes():
    case = util.case_c1()
    geo = [config.Geometry("a.csv", 7, config.Moving, config.MotionDetails(2.8, 0.0, 3.0, (1.0, 0.0)))]
    case.meta.ShiftingMode = config.PlanarShifting
    p = make_params(case.meta, case.consts, case.kernel, config.LaminarSPS(), config.LinearDensityDiffusion(), geo)
    assert p.n_motions == 1 and p.motions[0].group_marker == 7 and p.motions[0].velocity == 2.8
    assert list(p.motions[0].direction) == [1.0, 0.0, 0.0]
    assert p.shifting == 1 and p.viscosity == _abi.VISC_LAMINAR_SPS and p.dim == 2 and p.real_bytes == 8
