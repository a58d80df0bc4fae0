"""The product's step-control logic (csrc/sph_control.h, the body of k_step_control / k_step_end),
driven on the CPU through the shim: Δt, the rebuild trigger (src/SPHCellList.jl:739-762,
src/TimeStepping.jl:24-46), the neighbour-list state machine and the slab-mode pause."""
import ctypes as C

import numpy as np
import pytest

from test_physics_host import shim  # noqa: F401  (fixture)

H, C0, CFL = 0.0077, 33.14, 0.2


def trace(shim, disp, visc, acc2, vel, fails=None, skin=0.0, pause=0, use_float=0, delta_x0=0.0, motion_vmax=0.0):  # noqa: F811
    n = len(disp)
    f = lambda a: np.ascontiguousarray(a, np.float64)
    d2, vi, a2, v2 = f(np.asarray(disp) ** 2), f(visc), f(acc2), f(np.asarray(vel) ** 2)
    fl = np.ascontiguousarray(np.zeros(n) if fails is None else fails, np.uint8)
    out = np.zeros((n, 10))
    shim.shim_control_trace.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_double] * 5 + [C.c_int, C.c_int, C.c_double, C.c_void_p]
    shim.shim_control_trace(n, d2.ctypes.data, vi.ctypes.data, a2.ctypes.data, v2.ctypes.data, fl.ctypes.data, H, C0, CFL,
                            skin, motion_vmax, pause, use_float, delta_x0, out.ctypes.data)
    keys = ("dt", "do_rebuild", "list_build", "mode0", "mode1", "paused", "list_move", "delta_x", "total_time", "list_off")
    return {k: out[:, i] for i, k in enumerate(keys)}


def test_dt_follows_the_reference_formula(shim):  # noqa: F811
    rng = np.random.default_rng(0)
    n = 50
    visc, acc2 = rng.uniform(0, 5, n), rng.uniform(0, 400, n)
    acc2[0] = 0.0                                           # first step: every acceleration is zero -> dt₁ = +inf (Q3)
    t = trace(shim, np.zeros(n), visc, acc2, np.zeros(n))
    dt1 = np.sqrt(H / np.sqrt(acc2, where=acc2 > 0, out=np.full(n, np.nan)))
    dt1[acc2 == 0] = np.inf
    ref = CFL * np.minimum(dt1, H / (C0 + visc))            # src/TimeStepping.jl:40-45
    assert np.allclose(t["dt"], ref, rtol=1e-15)
    assert t["dt"][0] == pytest.approx(CFL * H / (C0 + visc[0]))
    assert np.allclose(t["total_time"], np.cumsum(ref), rtol=1e-14)


def test_rebuild_when_four_times_the_displacement_reaches_h(shim):  # noqa: F811
    n = 40
    disp = np.full(n, 0.03 * H)
    t = trace(shim, disp, np.zeros(n), np.ones(n), np.zeros(n), delta_x0=1.0 + H)   # SimulationLoop entry: Δx = 1 + h (:739)
    assert t["do_rebuild"][0] == 1                          # forced rebuild on entry (Q4)
    # Δx restarts at 0 and grows by 4·disp per step (:723,744): rebuild at the first step where it reaches h
    acc, expect = 0.0, []
    for s in range(1, n):
        acc += 4 * disp[s]
        hit = acc >= H
        expect.append(int(hit))
        if hit:
            acc = 0.0
    assert t["do_rebuild"][1:].astype(int).tolist() == expect
    assert sum(expect) >= 3


@pytest.mark.parametrize("use_float", [0, 1])
def test_lists_are_only_used_inside_their_displacement_bound(shim, use_float):  # noqa: F811
    rng = np.random.default_rng(1)
    n = 400
    vel = np.abs(rng.normal(1.5, 0.8, n)) + 0.1
    vel[150:170] = 12.0                                      # a burst: the bound must force builds
    disp = 0.5 * CFL * H / C0 * vel                          # half-step displacement ~ dt/2 · v
    skin = 0.1 * 2 * H
    t = trace(shim, disp, np.zeros(n), np.ones(n), vel, skin=skin, use_float=use_float)
    builds = t["list_build"].astype(int)
    assert builds[0] == 1 and 3 < builds.sum() < n // 2      # reused, but not forever
    moved = 0.0                                              # independent bookkeeping of the displacement bound
    for s in range(n):
        dt = t["dt"][s]
        if s > 0:
            moved += t["dt"][s - 1] * max(vel[s - 1], vel[s])
        if builds[s] or t["do_rebuild"][s]:
            assert builds[s] == 1                            # a cell rebuild always rebuilds the lists
            moved = 0.0
        half = 0.5 * dt * vel[s]
        if t["mode0"][s] == 2:
            assert 2 * moved <= skin * (1 + 1e-12)
        if t["mode1"][s] == 2:
            assert 2 * (moved + half) <= skin * (1 + 1e-12)
        assert t["mode0"][s] == 2                            # pass 1 always has fresh-enough lists (built at xₙ if need be)


def test_a_failed_build_switches_lists_off_until_the_cells_change(shim):  # noqa: F811
    n = 60
    vel = np.full(n, 1.0)
    disp = np.full(n, 0.02 * H)                              # rebuild every 13 steps
    fails = np.zeros(n, np.uint8)
    fails[0] = 1
    t = trace(shim, disp, np.zeros(n), np.ones(n), vel, fails=fails, skin=0.2 * H, delta_x0=1.0 + H)
    first_rebuild = int(np.nonzero(t["do_rebuild"][1:])[0][0]) + 1
    assert t["list_build"][0] == 1
    assert np.all(t["list_off"][1:first_rebuild] == 1)
    assert np.all(t["mode0"][1:first_rebuild] == 0) and np.all(t["list_build"][1:first_rebuild] == 0)
    assert t["list_build"][first_rebuild] == 1 and t["list_off"][first_rebuild] == 0   # the new cells get a new chance


def test_slab_mode_pauses_the_step_that_needs_a_rebuild(shim):  # noqa: F811
    n = 30
    disp = np.full(n, 0.05 * H)
    a = trace(shim, disp, np.zeros(n), np.ones(n), np.zeros(n), pause=0, delta_x0=1.0 + H)
    b = trace(shim, disp, np.zeros(n), np.ones(n), np.zeros(n), pause=1, delta_x0=1.0 + H)
    assert np.array_equal(a["do_rebuild"], b["do_rebuild"]) and np.array_equal(a["dt"], b["dt"])
    assert np.array_equal(b["paused"], b["do_rebuild"])      # paused exactly on rebuild steps ...
    assert not a["paused"].any()                             # ... and never on one GPU
    assert np.allclose(a["total_time"], b["total_time"])


def test_lean_sequence_pauses_whenever_it_lacks_a_kernel_the_step_needs(shim):  # noqa: F811
    """pause bits 1 | 2 (single-GPU lean sequence: no UpdateNeighbors! chain, no cull kernels): a step pauses
    exactly when it rebuilds or when one of its passes is not served by the lists; the decisions themselves
    (dt, rebuilds, list builds, modes) are those of the full sequence"""
    n = 120
    rng = np.random.default_rng(5)
    disp = (0.02 + 0.02 * rng.random(n)) * H
    vel = np.full(n, 1.0)
    vel[40:60] = 400.0            # a burst of speed: the half-step displacement outgrows the skin, pass 2 goes to the cull kernel
    fails = np.zeros(n, np.uint8)
    fails[75:] = 1                # the next build overflows: lists off until the next cell rebuild
    kw = dict(fails=fails, skin=0.1 * H, delta_x0=1.0 + H)
    a = trace(shim, disp ** 2, np.zeros(n), np.ones(n), vel ** 2, pause=0, **kw)
    b = trace(shim, disp ** 2, np.zeros(n), np.ones(n), vel ** 2, pause=3, **kw)
    for k in ("dt", "do_rebuild", "list_build", "mode0", "mode1", "delta_x", "list_off"):
        assert np.array_equal(a[k], b[k]), k
    assert np.allclose(a["total_time"], b["total_time"])
    need_full = (b["do_rebuild"] == 1) | (b["mode0"] != 2) | (b["mode1"] != 2)
    assert np.array_equal(b["paused"] == 1, need_full)
    assert need_full.any() and (~need_full).any() and not a["paused"].any()
    assert np.any((b["mode1"] != 2) & (b["mode0"] == 2)) and np.any(b["list_off"] == 1)    # both reasons occurred


def test_moving_bodies_count_towards_the_bound(shim):  # noqa: F811
    n = 100
    t0 = trace(shim, np.zeros(n), np.zeros(n), np.ones(n), np.zeros(n), skin=0.2 * H)
    t1 = trace(shim, np.zeros(n), np.zeros(n), np.ones(n), np.zeros(n), skin=0.2 * H, motion_vmax=2.8)
    assert t0["list_build"].sum() == 1 and t1["list_build"].sum() > 3


def test_list_bound_holds_for_actual_symplectic_motion(shim):  # noqa: F811
    """End-to-end check of the skin logic against real particle motion: particles move by the loop's
    own update formulas (x½ = x + v·dt/2, v' = v + a·dt, x' = x + (v + v')/2·dt) under a random,
    time-varying acceleration field, with dt, rebuilds and list decisions taken by the product's
    step_control.  Whenever a pass is allowed to use the lists, every pair within H at that pass's
    positions must already be in the list built (within H + skin) at the last build."""
    rng = np.random.default_rng(4)
    n, Hc, skin = 300, 2 * H, 0.1 * 2 * H
    box = 6 * Hc
    x = rng.uniform(0, box, (n, 3))
    v = rng.normal(0, 2.0, (n, 3))
    shim.shim_ctl_new.restype = C.c_void_p
    shim.shim_ctl_new.argtypes = [C.c_double]
    shim.shim_ctl_head.argtypes = [C.c_void_p] + [C.c_double] * 9 + [C.c_int, C.c_void_p]
    shim.shim_ctl_body.argtypes = [C.c_void_p]
    shim.shim_ctl_free.argtypes = [C.c_void_p]
    ctl = shim.shim_ctl_new(1.0 + H)
    out = np.zeros(8)

    def within(p, r):
        d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
        return d2 <= r * r

    listed, x_half_prev, uses, builds = None, None, 0, 0
    try:
        for step in range(260):
            a = rng.normal(0, 3.0e3, (n, 3)) * (1.0 + np.sin(0.05 * step))       # strong, varying accelerations
            disp2 = 0.0 if x_half_prev is None else float(((x_half_prev - x) ** 2).sum(1).max())
            shim.shim_ctl_head(ctl, disp2, 0.0, float((a ** 2).sum(1).max()), float((v ** 2).sum(1).max()), H, C0, CFL, skin, 0.0, 0,
                               out.ctypes.data)
            dt, build, m0, m1 = out[0], int(out[2]), int(out[3]), int(out[4])
            if build:
                listed = within(x, Hc + skin)                                      # k_list_build at xₙ
                builds += 1
            if m0 == 2:
                assert listed is not None and not np.any(within(x, Hc) & ~listed), f"pass 1 of step {step}"
                uses += 1
            x_half = x + v * (dt / 2)
            if m1 == 2:
                assert not np.any(within(x_half, Hc) & ~listed), f"pass 2 of step {step}"
                uses += 1
            v_new = v + a * dt
            x = x + (v + v_new) / 2 * dt
            v = v_new
            x_half_prev = x_half                                                   # xₙ⁺ of this step, compared with xₙ₊₁ at the next head
            shim.shim_ctl_body(ctl)
    finally:
        shim.shim_ctl_free(ctl)
    assert builds >= 3 and uses > 2 * builds                                       # the lists were reused, and rebuilt as particles moved


def test_brick_bound_holds_for_relative_motion(shim):  # noqa: F811
    """Per-brick list validity (brick_list_decision, csrc/sph_control.h): a cloud that moves FAST as a
    whole but deforms slowly.  The decision sees only D = the diagonal of the velocity bounding box at
    the two step heads; whenever it keeps the lists, every pair within H at the pass's positions must
    have been within H + skin at the last build.  The global rule (2 dt max|v|) would rebuild ~20x
    more often here."""
    rng = np.random.default_rng(7)
    n, Hc = 250, 2 * H
    skin = 0.1 * Hc
    x = rng.uniform(0, 5 * Hc, (n, 3))
    v = np.array([25.0, -10.0, 5.0]) + rng.normal(0, 0.3, (n, 3))       # bulk speed ~27 m/s, relative ~1 m/s
    shim.shim_brick_decision.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_double, C.c_double, C.c_double]
    move = C.c_float(0.0)

    def within(p, r):
        d2 = ((p[:, None, :] - p[None, :, :]) ** 2).sum(-1)
        return d2 <= r * r

    def box_diag(*vs):
        allv = np.concatenate(vs)
        return float(np.linalg.norm(allv.max(0) - allv.min(0)))

    listed, v_prev, dt_prev, builds, global_builds, gmove = within(x, Hc + skin), v.copy(), 0.0, 1, 1, 0.0
    for step in range(400):
        dt = CFL * H / (C0 + 5.0)
        a = rng.normal(0, 1.5e2, (n, 3)) * (1.0 + np.sin(0.07 * step)) + np.array([0.0, 0.0, -9.81])
        D = box_diag(v, v_prev) * 1.0001
        if shim.shim_brick_decision(C.byref(move), D, dt_prev, dt / 2, skin):
            listed = within(x, Hc + skin)
            builds += 1
        # what the global rule would have done
        vmax = max(np.linalg.norm(v, axis=1).max(), np.linalg.norm(v_prev, axis=1).max())
        gmove += dt_prev * vmax
        if gmove + dt / 2 * vmax > 0.49 * skin:
            gmove, global_builds = 0.0, global_builds + 1
        assert not np.any(within(x, Hc) & ~listed), f"pass 1 of step {step}"
        x_half = x + v * (dt / 2)
        assert not np.any(within(x_half, Hc) & ~listed), f"pass 2 of step {step}"
        v_new = v + a * dt
        x = x + (v + v_new) / 2 * dt
        v_prev, v, dt_prev = v, v_new, dt
    assert builds >= 3 and global_builds > 4 * builds, (builds, global_builds)


def test_lean_batches_predicted_from_the_control_block(shim):  # noqa: F811
    """the host loop of run_steps (sphb200.cu) replayed on the CPU with the product's own step_control /
    lean_steps_ahead: lean batches sized from the control block, a full step when the estimate says the next
    step rebuilds, pause + resume when a lean step rebuilds after all.  With a steady flow the estimate never
    misses; with an accelerating one a miss costs a pause, never a wrong decision."""
    import ctypes as C
    shim.shim_ctl_new.restype = C.c_void_p
    shim.shim_ctl_new.argtypes = [C.c_double]
    shim.shim_ctl_head.argtypes = [C.c_void_p] + [C.c_double] * 9 + [C.c_int, C.c_void_p]
    shim.shim_ctl_body.argtypes = [C.c_void_p]
    shim.shim_ctl_free.argtypes = [C.c_void_p]
    shim.shim_ctl_lean_ahead.restype = C.c_longlong
    shim.shim_ctl_lean_ahead.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_longlong]
    shim.shim_ctl_paused.argtypes = [C.c_void_p]
    h = H / 2.0
    out = np.zeros(8)

    def run(disp_of_step, nsteps, batch):
        ctl = shim.shim_ctl_new(1.0 + h)
        stats = {"lean": 0, "full": 0, "pauses": 0, "rebuilds": 0, "wasted": 0}
        trace = []

        def head(step, pause):
            d = disp_of_step(step)
            shim.shim_ctl_head(ctl, d * d, 0.0, 1.0, 0.0, h, C0, CFL, 0.0, 0.0, pause, out.ctypes.data)
            return int(out[1]), int(out[5])
        try:
            step = 0
            while step < nsteps:
                ahead = shim.shim_ctl_lean_ahead(ctl, h, 0.0, batch)
                if ahead >= 1:
                    n = min(ahead, nsteps - step)
                    for k in range(n):                       # the batch, enqueued blind
                        rebuild, done = head(step, 1)
                        if done:                             # paused: the rest of the batch runs empty
                            assert rebuild == 1 and shim.shim_ctl_paused(ctl) == 1
                            stats["pauses"] += 1
                            stats["wasted"] += n - k - 1
                            shim.shim_ctl_body(ctl)          # host: resume, full body with UpdateNeighbors!
                            stats["rebuilds"] += 1
                            trace.append(step)
                            step += 1
                            break
                        shim.shim_ctl_body(ctl)
                        stats["lean"] += 1
                        step += 1
                else:
                    rebuild, done = head(step, 0)            # a full step: may rebuild in place
                    assert not done
                    shim.shim_ctl_body(ctl)
                    stats["full"] += 1
                    stats["rebuilds"] += rebuild
                    if rebuild:
                        trace.append(step)
                    step += 1
        finally:
            shim.shim_ctl_free(ctl)
        return stats, trace

    def reference_rebuild_steps(disp_of_step, nsteps):      # the plain sequence: every step a full step
        ctl = shim.shim_ctl_new(1.0 + h)
        steps = []
        for s in range(nsteps):
            d = disp_of_step(s)
            shim.shim_ctl_head(ctl, d * d, 0.0, 1.0, 0.0, h, C0, CFL, 0.0, 0.0, 0, out.ctypes.data)
            if int(out[1]):
                steps.append(s)
            shim.shim_ctl_body(ctl)
        shim.shim_ctl_free(ctl)
        return steps

    steady = lambda s: 0.006 * h                             # 4 * disp = 0.024 h per step: a rebuild every ~42 steps
    st, tr = run(steady, 600, 64)
    assert tr == reference_rebuild_steps(steady, 600) and len(tr) >= 10
    assert st["pauses"] == 0 and st["lean"] >= 0.85 * 600 and st["full"] <= 8 * len(tr)
    speeding = lambda s: 0.002 * h * (1.0 + 0.02 * s)        # the flow accelerates: the last increment underestimates the next ones
    st, tr = run(speeding, 600, 64)
    assert tr == reference_rebuild_steps(speeding, 600) and len(tr) >= 10
    assert st["lean"] >= 0.75 * 600 and st["pauses"] <= len(tr) and st["wasted"] <= 3 * len(tr)
    bursts = lambda s: (0.05 if s % 37 == 36 else 0.001) * h  # a sudden jump the estimate cannot see: the pause catches it
    st, tr = run(bursts, 400, 64)
    assert tr == reference_rebuild_steps(bursts, 400) and st["pauses"] >= 1
