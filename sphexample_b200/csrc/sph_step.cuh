// sph_step.cuh — the streaming kernels of one SimulationLoop iteration
// (src/SPHCellList.jl:742-802): Δt / Δx reductions and the step control block, ProgressMotion,
// Pressure!, the stand-alone half/full updates (the fused versions live in the interaction
// kernel's epilogue) and the mDBC ghost-node correction.
#pragma once

#include "sph_device.cuh"

namespace sph {

// ---------------------------------------------------------------------------------------------
// S0 + S1: device-wide reductions for update_delta_x! (src/SPHCellList.jl:706-724) and Δt
// (src/TimeStepping.jl:24-46, Q3: absolute positions, every particle type).
//   red_disp2 = max ‖xₙ⁺ - x‖²      red_visc = max |h v·x / (x·x + η²)|      red_acc2 = max ‖a‖²
// min over i of sqrt(h/‖aᵢ‖) == sqrt(h / max ‖aᵢ‖) (monotone), so one max serves.
// warp shuffle -> one atomicMax per warp on the bit pattern (all three are non-negative).
// ---------------------------------------------------------------------------------------------
template <class T>
struct StepRed {
    T disp2, visc, acc2, vel2;
    bool nan;
};
template <class T>
__device__ __forceinline__ void step_red_zero(StepRed<T> &r) {
    r.disp2 = r.visc = r.acc2 = r.vel2 = T(0);
    r.nan = false;
}
// one particle's terms: x, v, a of the (new) state n, xh = the half-step position (have_half)
template <class T, int D>
__device__ __forceinline__ void step_red_particle(StepRed<T> &r, const T *x, const T *v, const T *a, const T *xh, bool have_half,
                                                  T h, T eta2) {
    T vx = T(0), xx = T(0), aa = T(0), dd = T(0), vv = T(0);
#pragma unroll
    for (int k = 0; k < D; ++k) {
        vx += v[k] * x[k];
        xx += x[k] * x[k];
        aa += a[k] * a[k];
        vv += v[k] * v[k];
    }
    if (have_half) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            T d = xh[k] - x[k];
            dd += d * d;
        }
    }
    T visc = sph_abs(h * vx / (xx + eta2));
    r.nan |= !(visc == visc) || !(aa == aa) || !(dd == dd);
    r.disp2 = sph_max(r.disp2, dd);
    r.visc = sph_max(r.visc, visc);
    r.acc2 = sph_max(r.acc2, aa);
    r.vel2 = sph_max(r.vel2, vv);
}
// whole warp: shuffle reduction, then one atomicMax per quantity on the bit pattern (all non-negative)
template <class T>
__device__ __forceinline__ void step_red_commit(Ctl *ctl, StepRed<T> r) {
    r.disp2 = warp_max(r.disp2);
    r.visc = warp_max(r.visc);
    r.acc2 = warp_max(r.acc2);
    r.vel2 = warp_max(r.vel2);
    const bool nan = __any_sync(0xffffffffu, r.nan);
    if ((threadIdx.x & 31) == 0) {
        if (nan) {
            atomicCAS(&ctl->error, 0, SPH_ERR_ENUMERIC);
        } else {
            if (r.disp2 > T(0)) atomic_max_nonneg(&ctl->red_disp2, (double)r.disp2);
            if (r.visc > T(0)) atomic_max_nonneg(&ctl->red_visc, (double)r.visc);
            if (r.acc2 > T(0)) atomic_max_nonneg(&ctl->red_acc2, (double)r.acc2);
            if (r.vel2 > T(0)) atomic_max_nonneg(&ctl->red_vel2, (double)r.vel2);
        }
    }
}

// Stand-alone form (first step after an upload, stage-level calls): in the step loop the fused
// corrector of pass 2 accumulates the same terms in its epilogue (ctl->red_ready) and this kernel
// returns at once.
template <class T, int D>
__global__ void k_reduce_dt_dx(const typename Lay<T, D>::TA *__restrict__ A, const typename Lay<T, D>::TB *__restrict__ B,
                               const typename Lay<T, D>::TA *__restrict__ Ah,
                               const typename Lay<T, D>::TV *__restrict__ acc, int p0, int p1, T h, T eta2,
                               int have_half, Ctl *ctl) {
    using L = Lay<T, D>;
    if (ctl->error || ctl->done || ctl->red_ready) return;
    StepRed<T> red;
    step_red_zero(red);
    for (int i = p0 + blockIdx.x * blockDim.x + threadIdx.x; i < p1; i += gridDim.x * blockDim.x) {
        T x[D], v[D], rs, P, a[D], xh[D];
        L::unpack(A[i], B[i], x, v, rs, P);
        L::getv(acc[i], a);
#pragma unroll
        for (int k = 0; k < D; ++k) xh[k] = x[k];
        if (have_half) L::pos(Ah[i], xh);
        step_red_particle<T, D>(red, x, v, a, xh, have_half != 0, h, eta2);
    }
    step_red_commit(ctl, red);
}

// One thread: S0/S1 completion, the S2 decision, the while-condition and the neighbour-list state
// (sph_control.h holds the logic, shared with the CPU tests).
// h_rebuild / h_lists: handles of the CUDA-graph conditional nodes that hold the UpdateNeighbors!
// chain and the list maintenance of a captured step (0 outside a conditional step graph): the
// kernels behind them run only in the steps that need them instead of starting up empty.
template <class T>
__global__ void k_step_control(Ctl *ctl, GridInfo *grid, T h, T c0, T cfl, double list_skin, double motion_vmax,
                               int pause_on_rebuild, int list_local, cudaGraphConditionalHandle h_rebuild,
                               cudaGraphConditionalHandle h_lists) {
    step_control<T>(ctl, grid, h, c0, cfl, list_skin, motion_vmax, pause_on_rebuild, list_local);
    const bool live = !ctl->error && !ctl->done;
    if (h_rebuild) cudaGraphSetConditional(h_rebuild, (live && ctl->do_rebuild) ? 1u : 0u);
    if (h_lists) cudaGraphSetConditional(h_lists, (live && ctl->list_build != 0) ? 1u : 0u);
}

// UpdateMetaData!, src/SPHCellList.jl:679-685 (S19) (+ the list-maintenance accounting of the step)
__device__ __forceinline__ void step_end_full(Ctl *ctl, const GridInfo *grid) {
    if (!(ctl->error || ctl->done || !ctl->step_open) && ctl->bricks_flagged > 0) {
        ctl->list_build_equiv += (double)ctl->bricks_flagged / (double)max(1, grid->nbricks);
        ctl->n_list_builds += 1;
    }
    ctl->bricks_flagged = 0;
    ctl->bricks_urgent = 0;
    step_end(ctl);
}
__global__ void k_step_end(Ctl *ctl, const GridInfo *grid) { step_end_full(ctl, grid); }

// lean sequence: S19 of the previous step (if it is still open) and the head of the next one in ONE
// one-thread kernel — between two lean steps nothing else happens (the batch ends with a k_step_end)
template <class T>
__global__ void k_step_end_control(Ctl *ctl, GridInfo *grid, T h, T c0, T cfl, double list_skin, double motion_vmax,
                                   int pause_bits, int list_local) {
    step_end_full(ctl, grid);
    step_control<T>(ctl, grid, h, c0, cfl, list_skin, motion_vmax, pause_bits, list_local);
}

// test hook (option test_fail_list_build_at): pretend this step's list build overflowed
__global__ void k_test_fail_list_build(Ctl *ctl) {
    if (!ctl->error && !ctl->done) ctl->list_fail |= 2;
}

// any change of positions or cells outside the step sequence voids the neighbour lists
__global__ void k_invalidate_lists(Ctl *ctl) {
    // (also voids the Δt / Δx reductions a fused pass 2 left behind: the state is about to change)
    ctl->red_ready = 0;
    ctl->red_disp2 = ctl->red_visc = ctl->red_acc2 = ctl->red_vel2 = 0ull;
    ctl->list_valid = 0;
    ctl->list_build = 0;
    ctl->list_mode[0] = ctl->list_mode[1] = 0;
}

__global__ void k_reset_counters(Ctl *ctl) {
#pragma unroll
    for (int k = 0; k < 8; ++k) ctl->work_counter[k] = 0;
}

// snapshot of ρₙ for the pass-2 diffusion / viscosity terms (Q2)
template <class T, int D>
__global__ void k_snapshot_rho(const typename Lay<T, D>::TA *__restrict__ A, T *__restrict__ RN, int n, const Ctl *ctl) {
    if (ctl->error || ctl->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        RN[i] = sph_abs(Lay<T, D>::rhos(A[i]));
}

// Pressure!(P, ρ), src/SimulationEquations.jl:18-24
template <class T, int D>
__global__ void k_pressure(typename Lay<T, D>::TA *A, typename Lay<T, D>::TB *B, int n, Phys<T> ph, const Ctl *ctl) {
    if (ctl->error || ctl->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        typename Lay<T, D>::TA a = A[i];
        typename Lay<T, D>::TB b = B[i];
        Lay<T, D>::set_P(a, b, eos_gamma7(ph, sph_abs(Lay<T, D>::rhos(a))));
        A[i] = a;
        B[i] = b;
    }
}

// ProgressMotion, src/SPHCellList.jl:575-596 (Q8): Moving particles only, uses TotalTime of the
// step start.  dt2 < 0 means "read ctl->dt2".
struct MotionTable {
    int n;
    unsigned long long group[16];
    double velocity[16], start[16], duration[16], dir[16][3];
};

template <class T, int D>
__global__ void k_progress_motion(typename Lay<T, D>::TA *A, typename Lay<T, D>::TB *B, const uint8_t *__restrict__ type,
                                  const unsigned long long *__restrict__ group, int n, MotionTable mt, double dt2_arg,
                                  const Ctl *ctl) {
    using L = Lay<T, D>;
    if (ctl->error || ctl->done) return;
    const T dt2 = (T)(dt2_arg < 0.0 ? ctl->dt2 : dt2_arg);
    const T tnow = (T)ctl->total_time;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (type[i] != 3) continue;
        int m = -1;
        for (int k = 0; k < mt.n; ++k)
            if (mt.group[k] == group[i]) m = k;
        if (m < 0) continue;
        T st = (T)mt.start[m], du = (T)mt.duration[m];
        T should = (st <= tnow && tnow <= (st + du)) ? T(1) : T(0);
        T x[D], v[D], rs, P;
        typename L::TA a = A[i];
        typename L::TB b = B[i];
        L::unpack(a, b, x, v, rs, P);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            v[k] = (T)mt.velocity[m] * (T)mt.dir[m][k] * should;
            x[k] += v[k] * dt2;
        }
        L::pack(a, b, x, v, rs, P);
        A[i] = a;
        B[i] = b;
    }
}

// stand-alone HalfTimeStep + LimitDensityAtBoundary!(ρₙ⁺) (+ Pressure!(ρₙ⁺)), S9/S10/S13
template <class T, int D>
__global__ void k_half_step(const typename Lay<T, D>::TA *__restrict__ A, const typename Lay<T, D>::TB *__restrict__ B,
                            typename Lay<T, D>::TV *acc, const T *__restrict__ drhodt, const uint8_t *__restrict__ type,
                            typename Lay<T, D>::TA *Ah, typename Lay<T, D>::TB *Bh, int p0, int p1, Phys<T> ph,
                            double dt2_arg, const Ctl *ctl) {
    using L = Lay<T, D>;
    if (ctl->error || ctl->done) return;
    const T dt2 = (T)(dt2_arg < 0.0 ? ctl->dt2 : dt2_arg);
    for (int i = p0 + blockIdx.x * blockDim.x + threadIdx.x; i < p1; i += gridDim.x * blockDim.x) {
        T x[D], v[D], rs, P, a[D];
        L::unpack(A[i], B[i], x, v, rs, P);
        L::getv(acc[i], a);
        const uint8_t ty = type[i];
        const T gf = (T)type_gf(ty), ml = (T)type_ml(ty);
        T xh[D], vh[D], rhoh;
        half_step<T, D>(ph, x, v, a, sph_abs(rs), drhodt[i], gf, ml, dt2, xh, vh, rhoh);
        acc[i] = L::mkv(a);
        typename L::TA oa;
        typename L::TB ob;
        L::pack(oa, ob, xh, vh, ml > T(0) ? rhoh : -rhoh, eos_gamma7(ph, rhoh));
        Ah[i] = oa;
        Bh[i] = ob;
    }
}

// stand-alone LimitDensityAtBoundary!(ρ) + DensityEpsi! + FullTimeStep (+ Pressure!(ρ)), S16-S18
template <class T, int D>
__global__ void k_full_step(typename Lay<T, D>::TA *A, typename Lay<T, D>::TB *B, typename Lay<T, D>::TV *acc,
                            const T *__restrict__ drhodt, const typename Lay<T, D>::TA *__restrict__ Ah,
                            const uint8_t *__restrict__ type, const typename Lay<T, D>::TV *__restrict__ gradC,
                            const T *__restrict__ divr, int p0, int p1, Phys<T> ph, double dt_arg, const Ctl *ctl) {
    using L = Lay<T, D>;
    if (ctl->error || ctl->done) return;
    const T dt = (T)(dt_arg < 0.0 ? ctl->dt : dt_arg);
    for (int i = p0 + blockIdx.x * blockDim.x + threadIdx.x; i < p1; i += gridDim.x * blockDim.x) {
        T x[D], v[D], rs, P, a[D], gc[D];
        L::unpack(A[i], B[i], x, v, rs, P);
        L::getv(acc[i], a);
        const uint8_t ty = type[i];
        const T gf = (T)type_gf(ty), ml = (T)type_ml(ty);
        T rho = sph_abs(rs);
        T rhoh = sph_abs(L::rhos(Ah[i]));
        T dv = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) gc[k] = T(0);
        if (ph.shifting) {
            L::getv(gradC[i], gc);
            dv = divr[i];
        }
        full_step<T, D>(ph, x, v, a, rho, drhodt[i], rhoh, gf, ml, dt, gc, dv);
        typename L::TA oa;
        typename L::TB ob;
        L::pack(oa, ob, x, v, ml > T(0) ? rho : -rho, eos_gamma7(ph, rho));
        A[i] = oa;
        B[i] = ob;
        acc[i] = L::mkv(a);
    }
}

// ---------------------------------------------------------------------------------------------
// ApplyMDBCBeforeHalf!, src/SPHCellList.jl:219-266,319-365,491-505,598-622 (S6).  One thread per
// boundary particle that has a ghost node: gather fluid neighbours of the ghost node over the
// full 3^D stencil of the ghost node's cell, build b (D+1) and A (D+1)², solve in registers.
// Two phases like the reference: new densities go to `rho_new` first (k_mdbc_gather), then are
// applied (k_mdbc_apply) so that no thread reads a density another thread has just corrected.
// ---------------------------------------------------------------------------------------------
// The ghost node's sums and the solve (NeighborLoopMDBC! + the branches of ApplyMDBCCorrection, :598-622).
// Returns 0: no new density; 1: ρ_new = sol[0] + Σ sol[k+1]·(x_i − x_ghost)[k] — the low-order branch
// (|det A| < 1e-3, A₀₀ > 0) comes back as sol = {b₀/A₀₀, 0, …}, which that expression reproduces exactly.
// WARP-cooperative: all 32 lanes call it for the SAME node (a node has ~120 candidates, one thread per node
// made the kernel a 65 us latency chain on C5); lane l takes candidates l, l + 32, … of every row of the
// stencil, the partial sums are combined by a butterfly (same order on every lane: all lanes hold the
// same sums and take the same branch).
template <class T, int D>
__device__ __forceinline__ int mdbc_node_solve(const typename Lay<T, D>::TA *__restrict__ A, const uint8_t *__restrict__ type,
                                               const int *__restrict__ cell_start, const GridInfo *grid, const AxisMap &am,
                                               const Phys<T> &ph, const int (&gc)[3], const T (&gp)[D], double (&sol)[D + 1]) {
    using L = Lay<T, D>;
    constexpr int E = D + 1;
    const int lane = threadIdx.x & 31;
    const int nx = grid->nx, nm = grid->nm, ns = grid->ns;
    int cx = gc[am.ax_f] - grid->cmin[am.ax_f];
    int cm = (D == 3) ? gc[am.ax_m] - grid->cmin[am.ax_m] : 0;
    int cs = gc[am.ax_s] - grid->cmin[am.ax_s];
    double bv[E], Am[E][E];
    for (int r = 0; r < E; ++r) {
        bv[r] = 0.0;
        for (int c = 0; c < E; ++c) Am[r][c] = 0.0;
    }
    for (int ds = -1; ds <= 1; ++ds)
        for (int dm = (D == 3 ? -1 : 0); dm <= (D == 3 ? 1 : 0); ++dm) {
            int rs_ = cs + ds, rm = cm + dm;
            if (rs_ < 0 || rs_ >= ns || rm < 0 || rm >= nm) continue;
            int x0 = max(cx - 1, 0), x1 = min(cx + 1, nx - 1);
            if (x0 > x1) continue;
            int rk = (rs_ * nm + rm) * nx;
            int jb = cell_start[rk + x0], je = cell_start[rk + x1 + 1];
            for (int j = jb + lane; j < je; j += 32) {
                if (type[j] != 1) continue;
                typename L::TA aj = A[j];
                T xj[D];
                L::pos(aj, xj);
                T xij[D], r2 = T(0);
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    xij[k] = gp[k] - xj[k];
                    r2 += xij[k] * xij[k];
                }
                if (!(r2 <= ph.H2)) continue;
                T d = sph_sqrt(sph_abs(r2));
                T q = sph_min(sph_max(d * ph.h_inv, T(0)), T(2));
                T W = kernel_w(ph, q);
                T gW[D];
                if (ph.kernel == K_WENDLAND) {
                    T qm2 = q - T(2);
                    T fac = ph.gradw_c * (qm2 * qm2 * qm2);
#pragma unroll
                    for (int k = 0; k < D; ++k) gW[k] = fac * xij[k];
                } else {
                    T dwdq = (q <= T(1)) ? ph.alphaD * (T(-3) * q + T(2.25) * (q * q))
                                         : ph.alphaD * T(-0.75) * ((T(2) - q) * (T(2) - q));
                    T sc = dwdq * ph.h_inv;
#pragma unroll
                    for (int k = 0; k < D; ++k) gW[k] = sc * xij[k] / (d + ph.eta2);
                }
                T Vj = ph.m0 / sph_abs(L::rhos(aj));
                double col[E];
                col[0] = (double)(Vj * W);
                bv[0] += (double)(ph.m0 * W);
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    col[k + 1] = (double)(Vj * gW[k]);
                    bv[k + 1] += (double)(ph.m0 * gW[k]);
                }
#pragma unroll
                for (int r = 0; r < E; ++r) {
                    Am[r][0] += col[r];
#pragma unroll
                    for (int c = 1; c < E; ++c) Am[r][c] += (double)(-xij[c - 1]) * col[r];
                }
            }
        }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int r = 0; r < E; ++r) {
            bv[r] += __shfl_xor_sync(0xffffffffu, bv[r], off);
#pragma unroll
            for (int c = 0; c < E; ++c) Am[r][c] += __shfl_xor_sync(0xffffffffu, Am[r][c], off);
        }
    }
    double detA = det_lu<E>(Am);
    if (fabs(detA) >= 1e-3) {
        solve_lu<E>(Am, bv, sol);
        return 1;
    }
    if (Am[0][0] > 0.0) {
        sol[0] = bv[0] / Am[0][0];
#pragma unroll
        for (int k = 0; k < D; ++k) sol[k + 1] = 0.0;
        return 1;
    }
    return 0;
}

// ρ_new from the solve, in the reference's expression order; NaN -> ρ₀ (the `isnan` guards of :614-620)
template <class T, int D>
__device__ __forceinline__ T mdbc_extrapolate(const double *sol, const T (&xi)[D], const T (&gp)[D], T rho0) {
    double v1 = sol[0];
#pragma unroll
    for (int k = 0; k < D; ++k) v1 = fma(sol[k + 1], (double)xi[k] - (double)gp[k], v1);
    return (v1 == v1) ? (T)v1 : rho0;
}

template <class T, int D>
__global__ void k_mdbc_gather(const typename Lay<T, D>::TA *__restrict__ A, const typename Lay<T, D>::TV *__restrict__ ghost,
                              const uint8_t *__restrict__ type, const int *__restrict__ cell_start, const GridInfo *grid,
                              AxisMap am, int n, Phys<T> ph, double inv_cutoff, T *__restrict__ rho_new,
                              uint8_t *__restrict__ has_new, const Ctl *ctl) {
    using L = Lay<T, D>;
    constexpr int E = D + 1;
    if (ctl->error || ctl->done) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    // one warp per particle (every lane reads the same ghost point: the branches below are warp-uniform)
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += nwarps) {
        T gp[D];
        L::getv(ghost[i], gp);
        bool zero = true;
#pragma unroll
        for (int k = 0; k < D; ++k) zero &= (gp[k] == T(0));
        if (zero) {   // Q10 sentinel: "no ghost node"
            if (lane == 0) has_new[i] = 0;
            continue;
        }
        int bad = 0;
        int gc[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < D; ++k) gc[k] = map_floor_dev((double)gp[k], inv_cutoff, bad);
        double sol[E];
        const int ok = mdbc_node_solve<T, D>(A, type, cell_start, grid, am, ph, gc, gp, sol);
        if (lane != 0) continue;
        has_new[i] = (uint8_t)ok;
        if (!ok) continue;
        T xi[D];
        L::pos(A[i], xi);
        rho_new[i] = mdbc_extrapolate<T, D>(sol, xi, gp, ph.rho0);
    }
}

// ---- slab mode ---------------------------------------------------------------------------------
// A ghost node lies up to a few cells from its boundary particle, possibly across a slab face and
// beyond the one-layer halo, so the particle's rank cannot always evaluate it.  Ghost nodes are
// static (the reference never moves GhostPoints, Q10): every rank holds the global node table
// (point, particle ID; ascending ID) and evaluates the nodes whose CELL lies in its owned layers —
// the node's whole 3^D stencil is then local (owned + halo), traversed in the same order as on one
// GPU.  out[g] = {1, sol[0..D]} or zeros; an all-reduce (sum: exactly one rank writes a node) gives
// every rank every node's solve, and each rank extrapolates to the particles it holds (owned AND
// halo copies, which keeps the copies identical to their originals) by looking their ID up.
template <class T, int D>
__global__ void k_mdbc_nodes(const typename Lay<T, D>::TA *__restrict__ A, const typename Lay<T, D>::TV *__restrict__ g_point,
                             int ng, const uint8_t *__restrict__ type, const int *__restrict__ cell_start, const GridInfo *grid,
                             AxisMap am, Phys<T> ph, double inv_cutoff, int own_lo, int own_hi, double *__restrict__ out,
                             const Ctl *ctl) {
    using L = Lay<T, D>;
    constexpr int E = D + 1;
    if (ctl->error || ctl->done) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ng; g += nwarps) {
        double *o = out + (size_t)g * (E + 1);
        if (lane <= E) o[lane] = 0.0;
        __syncwarp();
        T gp[D];
        L::getv(g_point[g], gp);
        int bad = 0;
        int gc[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < D; ++k) gc[k] = map_floor_dev((double)gp[k], inv_cutoff, bad);
        if (gc[am.ax_s] < own_lo || gc[am.ax_s] >= own_hi) continue;   // another rank's node
        double sol[E];
        if (!mdbc_node_solve<T, D>(A, type, cell_start, grid, am, ph, gc, gp, sol)) continue;
        if (lane == 0) {
            o[0] = 1.0;
#pragma unroll
            for (int k = 0; k < E; ++k) o[k + 1] = sol[k];
        }
    }
}

template <class T, int D>
__global__ void k_mdbc_apply_nodes(typename Lay<T, D>::TA *A, T *RN, const uint8_t *__restrict__ type,
                                   const long long *__restrict__ id, int n, const typename Lay<T, D>::TV *__restrict__ g_point,
                                   const long long *__restrict__ g_id, int ng, const double *__restrict__ sols, T rho0, Ctl *ctl) {
    using L = Lay<T, D>;
    constexpr int E = D + 1;
    if (ctl->error || ctl->done || ng < 1) return;
    const long long id_lo = g_id[0], id_hi = g_id[ng - 1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const long long pid = id[i];
        if (pid < id_lo || pid > id_hi) continue;
        int lo = 0, hi = ng - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (g_id[mid] < pid) lo = mid + 1;
            else hi = mid;
        }
        if (g_id[lo] != pid) continue;
        const double *o = sols + (size_t)lo * (E + 1);
        if (o[0] == 0.0) continue;
        typename L::TA a = A[i];
        T xi[D], gp[D];
        L::pos(a, xi);
        L::getv(g_point[lo], gp);
        T r = mdbc_extrapolate<T, D>(o + 1, xi, gp, rho0);
        if (!(r > T(0))) {   // see k_mdbc_apply
            atomicCAS(&ctl->error, 0, SPH_ERR_ENUMERIC);
            continue;
        }
        L::set_rhos(a, type[i] == 1 ? r : -r);
        A[i] = a;
        RN[i] = r;
    }
}

template <class T, int D>
__global__ void k_mdbc_apply(typename Lay<T, D>::TA *A, T *RN, const uint8_t *__restrict__ type,
                             const T *__restrict__ rho_new, const uint8_t *__restrict__ has_new, int n, Ctl *ctl) {
    using L = Lay<T, D>;
    if (ctl->error || ctl->done) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!has_new[i]) continue;
        typename L::TA a = A[i];
        T r = rho_new[i];
        // The table carries the MotionLimiter in the SIGN of the stored density, so a density must be
        // positive.  The reference would write a zero or negative extrapolation as it is (and go on with
        // it); that cannot be represented here — and is not physical — so it stops the run instead.
        if (!(r > T(0))) {
            atomicCAS(&ctl->error, 0, SPH_ERR_ENUMERIC);
            continue;
        }
        L::set_rhos(a, type[i] == 1 ? r : -r);
        A[i] = a;
        RN[i] = r;
    }
}

}  // namespace sph
