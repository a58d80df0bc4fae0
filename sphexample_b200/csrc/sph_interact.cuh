// sph_interact.cuh — the pair traversal: NeighborLoop! ∘ ComputeInteractions!
// (src/SPHCellList.jl:168-217,268-317) as a GATHER over the full 3^D stencil, with the
// symplectic half/full updates (src/SPHCellList.jl:624-677) fused into the epilogue.
// No atomics, no per-thread accumulator copies (ResetStep!/ReductionStep!, :367-484, vanish).
//
// Work unit = one "brick": up to BT consecutive (cell-sorted) particles of one row of cells.
// Because x is the fastest key component, the candidates of a brick are, for each of the
// 3^(D-1) neighbouring rows, ONE contiguous span of the sorted arrays (cells cx0-1 .. cx1+1);
// one elected thread stages those 9 (3 in 2D) spans of each packed array into shared memory with
// cp.async.bulk (1-D TMA) completing on an mbarrier.  Persistent CTAs fetch bricks from an atomic
// work counter.  Three kernels share this frame:
//
//   k_interact       the CULL kernel: each thread owns one target particle; a warp walks the union
//                    of its lanes' windows with broadcast shared-memory reads, tests the cut-off
//                    (~18 % hit rate in 3D) and runs the pair body in place, or (COMPACT) from
//                    per-thread lists in shared memory so that the body executes on dense warps.
//   k_list_build     the same walk without physics: records, per particle, every candidate inside
//                    the window and within H + skin as (window index | role) in global memory;
//   k_list_reorder   gives every list the bank-aware entry order of sph_listorder.h;
//   k_interact_ring  stages whole brick windows through a ring of shared-memory slots and runs the
//                    pair body over the recorded entries only (~190 instead of ~770 candidates),
//                    branch-free (these three live in sph_ring.cuh).
//
// ctl->list_mode[pass] (k_step_control) says which of k_interact / k_interact_ring serves a pass;
// the other one returns at once.
//
// Pair-set fidelity: a candidate b is evaluated for target a iff b's (stale) cell is within the
// 3^D stencil of a's (stale) cell AND |x_a - x_b|² <= H² now — exactly the reference's set,
// including its misses between rebuilds — hence the per-lane window test (applied by the cull walk
// and by the list build; cells do not change while a list lives).
#pragma once

#include <type_traits>

#include "sph_device.cuh"
#include "sph_listorder.h"
#include "sph_step.cuh"

namespace sph {

constexpr int LIST_CAP = 64;        // per-thread accepted-neighbour list entries
constexpr int ROLE_BIT = 0x8000;

enum { EPI_STORE = 0, EPI_FUSED = 1 };

template <class T, int D>
struct InteractArgs {
    using L = Lay<T, D>;
    // pass inputs: state n (pass 0) or state n+½ (pass 1)
    const typename L::TA *A;
    const typename L::TB *B;
    const T *RN;                       // ρₙ (pass 1: state-n density, Q2); unused in pass 0
    T *rn_out;                         // pass 0, fused: where the epilogue leaves ρₙ of its particle (null: k_snapshot_rho did it)
    const typename L::TB *Bn;          // vₙ of neighbours (pass 1 + LaminarSPS only)
    // own state n, read and overwritten by the fused corrector (pass 1)
    typename L::TA *An_rw;
    typename L::TB *Bn_rw;
    // fused predictor output (pass 0)
    typename L::TA *Ah_out;
    typename L::TB *Bh_out;
    // plain outputs
    T *drhodt;                         // EPI_STORE
    typename L::TV *acc;               // EPI_STORE and fused corrector
    typename L::TV *gradC;             // PlanarShifting
    T *divr;
    T *ksum;                           // StoreKernelOutput
    typename L::TV *kgrad;
    // cell list
    const int *cell_start;
    const int *ckey;
    const uint8_t *type;
    const Brick *bricks;
    const GridInfo *grid;
    Ctl *ctl;
    Phys<T> phys;
    int cap;                           // staged candidates per stage (multiple of 4)
    int epilogue;                      // EPI_STORE / EPI_FUSED
    int use_tma;                       // 1: cp.async.bulk staging, 0: cooperative ld/st staging
    AxisMap am;                        // which component is the fast / middle / slab axis of the cell key (roles: row_role)
    int counter_slot;                  // which ctl->work_counter this launch consumes
    int brick_part;                    // 0: all bricks, 1: slab-boundary bricks, 2: interior bricks
    // slab mode, single launch per pass: the boundary-layer bricks [0, nbricks_bnd) are taken first;
    // the CTA that finishes the last of them raises *bnd_flag to bnd_epoch, which releases the halo
    // exchange waiting on another stream (cuStreamWaitValue32) while the interior bricks go on
    unsigned *bnd_flag;
    unsigned bnd_epoch;
    // per-particle neighbour lists (sph_interact.cuh, "lists"): entry k of particle i is the u16
    // ((k & 7)-th half-word of) nl[(k >> 3) * nl_stride + i]
    uint4 *nl;
    int *nl_cnt;
    size_t nl_stride;
    int lcap;                          // list capacity per particle (multiple of 8)
    int list_cap_cand;                 // candidates one ring slot of the list kernel holds; its last 8 records are the sentinels
    int list_reorder;                  // 1: k_list_reorder runs after a build (bank-aware entry order, sph_listorder.h)
    const int *brick_flag;             // per brick: 2 rebuild now, 1 due soon (rebuilt if the step builds at all), 0 good (k_brick_bounds)
    float *brick_move;                 // per brick: accumulated relative-displacement bound (reset by the build)
    T Hs2;                             // (H + skin)^2: acceptance radius of a list build
    int force_cull;                    // 1: ignore ctl->list_mode (stage-level entry points)
    int lean_guard;                    // 1: no cull kernel stands by in this sequence — a failed list build pauses the step
};

// ctl->list_mode[pass]: which kernel serves the pass
enum { LM_CULL = 0, LM_USE = 2 };
constexpr int LIST_IDX_MASK = 0x7fff;   // entry = window index | role << 15

template <class T, int D, int PASS, bool GENERIC>
struct StageSizes {
    using L = Lay<T, D>;
    static constexpr int esA = sizeof(typename L::TA);
    static constexpr int esB = sizeof(typename L::TB);
    static constexpr int esR = PASS ? sizeof(T) : 0;
    static constexpr int esBn = (PASS && GENERIC) ? sizeof(typename L::TB) : 0;
    static constexpr int per_candidate = esA + esB + esR + esBn;
};

// Called by every thread after the end-of-brick barrier (all of the brick's results are written).
template <class T, int D>
__device__ __forceinline__ void signal_boundary_brick(const InteractArgs<T, D> &g, int bidx, int pass) {
    if (g.bnd_flag == nullptr || threadIdx.x != 0) return;
    const int nbnd = g.grid->nbricks_bnd;
    if (bidx >= nbnd) return;
    __threadfence();
    if (atomicAdd(&g.ctl->bnd_done[pass], 1) + 1 == nbnd) {
        __threadfence_system();
        atomicMax(g.bnd_flag, g.bnd_epoch);
    }
}

// What the fused epilogue reads of its own particle, loaded at the START of the particle's work by
// the list kernel (at the end it would be an exposed memory round trip per 32-target sub-brick).
template <class T, int D>
struct EpiloguePrefetch {
    typename Lay<T, D>::TA an;   // own state n (pass 2 only)
    typename Lay<T, D>::TB bn;
    uint8_t type;
};

// Per-particle tail of a pass: plain stores (stage-level entry points) or the fused symplectic
// half / full update, shared by the cull kernel and the list kernel.
template <class T, int D, int PASS, bool GENERIC>
__device__ __forceinline__ void interact_epilogue(const InteractArgs<T, D> &g, int i, const T *xa, const T *va, T rho_a,
                                                  const PairAccum<T, D> &sacc, T drho, T *acc, StepRed<T> &red,
                                                  const EpiloguePrefetch<T, D> *pre = nullptr) {
    using L = Lay<T, D>;
    using TA = typename L::TA;
    using TB = typename L::TB;
    const Phys<T> &ph = g.phys;
    if (GENERIC) {
        drho = sacc.drho;
#pragma unroll
        for (int k = 0; k < D; ++k) acc[k] = sacc.acc[k];
        if (ph.shifting) {
            g.gradC[i] = L::mkv(sacc.gradC);
            g.divr[i] = sacc.divr;
        }
        if (ph.kernel_output) {
            g.ksum[i] = sacc.ksum;
            g.kgrad[i] = L::mkv(sacc.kgrad);
        }
    }
    if (g.epilogue == EPI_STORE) {
        g.drhodt[i] = drho;
        g.acc[i] = L::mkv(acc);
    } else {
        const uint8_t ty = pre ? pre->type : g.type[i];
        const T gf = (T)type_gf(ty), ml = (T)type_ml(ty);
        if (PASS == 0) {
            // HalfTimeStep + LimitDensityAtBoundary!(ρₙ⁺) + Pressure!(ρₙ⁺)  (S9, S10, S13)
            const T dt2 = (T)g.ctl->dt2;
            T xh[D], vh[D], rhoh;
            half_step<T, D>(ph, xa, va, acc, rho_a, drho, gf, ml, dt2, xh, vh, rhoh);
            TA oa;
            TB ob;
            L::pack(oa, ob, xh, vh, ml > T(0) ? rhoh : -rhoh, eos_gamma7(ph, rhoh));
            g.Ah_out[i] = oa;
            g.Bh_out[i] = ob;
            if (g.rn_out) g.rn_out[i] = rho_a;   // ρₙ for the pass-2 diffusion / viscosity terms (Q2): no separate snapshot sweep
        } else {
            // LimitDensityAtBoundary!(ρ) + DensityEpsi! + FullTimeStep + Pressure!  (S16-S18, S5)
            const T dt = (T)g.ctl->dt;
            T xn[D], vn[D], rs, Pn;
            if (pre) L::unpack(pre->an, pre->bn, xn, vn, rs, Pn);
            else L::unpack(g.An_rw[i], g.Bn_rw[i], xn, vn, rs, Pn);
            T rho = sph_abs(rs);
            T gc[D];
#pragma unroll
            for (int k = 0; k < D; ++k) gc[k] = GENERIC ? sacc.gradC[k] : T(0);
            full_step<T, D>(ph, xn, vn, acc, rho, drho, rho_a, gf, ml, dt, gc, GENERIC ? sacc.divr : T(0));
            // S0 / S1 of the NEXT step (update_delta_x!, Δt: src/SPHCellList.jl:706-724, src/TimeStepping.jl:24-46)
            // on the state just produced; xa is this step's half-step position
            step_red_particle<T, D>(red, xn, vn, acc, xa, true, ph.h, ph.eta2);
            TA oa;
            TB ob;
            L::pack(oa, ob, xn, vn, ml > T(0) ? rho : -rho, eos_gamma7(ph, rho));
            g.An_rw[i] = oa;
            g.Bn_rw[i] = ob;
            g.acc[i] = L::mkv(acc);
        }
    }
}

template <class T, int D, int PASS, bool GENERIC, bool COMPACT, int BT>
__global__ void __launch_bounds__(BT) k_interact(const InteractArgs<T, D> g) {
    using L = Lay<T, D>;
    using TA = typename L::TA;
    using TB = typename L::TB;
    using TV = typename L::TV;
    constexpr int NR = (D == 3) ? 9 : 3;
    using SS = StageSizes<T, D, PASS, GENERIC>;

    if (g.ctl->error || g.ctl->done) return;
    // which kernel serves this pass: the list kernel when the lists are valid, this one otherwise
    const int lmode = g.force_cull ? LM_CULL : g.ctl->list_mode[PASS];
    if (lmode == LM_USE && !g.ctl->list_fail) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap = g.cap;
    TA *sA = reinterpret_cast<TA *>(smem_raw);
    TB *sB = reinterpret_cast<TB *>(smem_raw + (size_t)cap * SS::esA);
    T *sR = reinterpret_cast<T *>(smem_raw + (size_t)cap * (SS::esA + SS::esB));
    TB *sBn = reinterpret_cast<TB *>(smem_raw + (size_t)cap * (SS::esA + SS::esB + SS::esR));
    unsigned short *slist = reinterpret_cast<unsigned short *>(smem_raw + (size_t)cap * SS::per_candidate);
    // (slist is LIST_CAP * BT entries when COMPACT, otherwise unused / zero-sized)

    __shared__ uint64_t s_bar;
    __shared__ int s_brick;
    __shared__ int s_w0a[NR], s_len[NR], s_off[NR + 1];

    const int tid = threadIdx.x;
    const Phys<T> &ph = g.phys;
    const bool use_sps = GENERIC && PASS && (ph.viscosity == V_SPS);

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;

    const int nx = g.grid->nx, nm = g.grid->nm;
    const int nbricks = g.grid->nbricks;
    const int brick_first = g.brick_part == 2 ? g.grid->nbricks_bnd : 0;
    const int brick_end = g.brick_part == 1 ? g.grid->nbricks_bnd : nbricks;
    const int npad = (g.grid->n_total + 3) & ~3;
    StepRed<T> red;   // fused corrector (pass 2): Δt / Δx reductions of the new state
    step_red_zero(red);

    for (;;) {
        if (tid == 0) s_brick = brick_first + atomicAdd(&g.ctl->work_counter[g.counter_slot], 1);
        __syncthreads();
        const int bidx = s_brick;
        if (bidx >= brick_end) break;
        const Brick br = g.bricks[bidx];
        const int key0 = g.ckey[br.t0], key1 = g.ckey[br.t1 - 1];
        const int cx0 = key0 % nx, cx1 = key1 % nx;
        const int rowbase = key0 - cx0;

        // ---- candidate spans of the 3^(D-1) neighbouring rows (aligned to 4 elements) ----
        if (tid < NR) {
            int dm = (D == 3) ? (tid % 3 - 1) : 0;
            int ds = (D == 3) ? (tid / 3 - 1) : (tid - 1);
            int rk = rowbase + (ds * nm + dm) * nx;
            int w0 = g.cell_start[rk + cx0 - 1];
            int w1 = g.cell_start[rk + cx1 + 2];
            int w0a = w0 & ~3;
            int w1a = min((w1 + 3) & ~3, npad);
            if (w1 <= w0) w1a = w0a;   // empty row: stage nothing
            s_w0a[tid] = w0a;
            s_len[tid] = w1a - w0a;
        }
        __syncthreads();
        if (tid == 0) {
            int o = 0;
#pragma unroll
            for (int r = 0; r < NR; ++r) {
                s_off[r] = o;
                o += s_len[r];
            }
            s_off[NR] = o;
        }
        __syncthreads();
        const int total = s_off[NR];

        // ---- this thread's target particle ------------------------------------------------
        const int i = br.t0 + tid;
        const bool valid = i < br.t1;
        const int warp_first = br.t0 + (tid & ~31);
        const bool warp_has_work = warp_first < br.t1;
        const int last_lane = min(31, br.t1 - warp_first - 1) & 31;
        T xa[D], va[D], rho_a = T(1), P_a = T(0), rhon_a = T(1), ml_a = T(0);
        int cxi = cx0, cs_a = 0, ce_a = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) xa[k] = va[k] = T(0);
        if (valid) {
            T rs;
            L::unpack(g.A[i], g.B[i], xa, va, rs, P_a);
            rho_a = sph_abs(rs);
            ml_a = rs > T(0) ? T(1) : T(0);
            rhon_a = PASS ? g.RN[i] : rho_a;
            int ki = g.ckey[i];
            cxi = ki - rowbase;
            cs_a = g.cell_start[ki];
            ce_a = g.cell_start[ki + 1];
        }
        PairSide<T, D> sa;   // GENERIC only
        PairAccum<T, D> sacc;
        accum_zero(sacc);
        if (GENERIC) {
#pragma unroll
            for (int k = 0; k < D; ++k) {
                sa.x[k] = xa[k];
                sa.v[k] = va[k];
                sa.vn[k] = va[k];
            }
            sa.rho = rho_a;
            sa.P = P_a;
            sa.rho_n = rhon_a;
            sa.ml = ml_a;
            if (use_sps && valid) {
                T dummy_x[D], rs, Pd;
                L::unpack(g.An_rw[i], g.Bn_rw[i], dummy_x, sa.vn, rs, Pd);
            }
        }
        T drho = T(0), acc[D];
#pragma unroll
        for (int k = 0; k < D; ++k) acc[k] = T(0);
        const FastTarget<T> ft = make_fast_target<T>(ph, rho_a, P_a, rhon_a, ml_a, PASS == 0);   // !GENERIC only
        FastSums<T, D> fs;
        fast_zero(fs);

        // pair body for one staged candidate (smem slot sj), shared by both phases
        auto pair_body = [&](int sj, bool a_is_i) {
            T xb[D], vb[D], rsb, P_b;
            L::unpack(sA[sj], sB[sj], xb, vb, rsb, P_b);
            T rho_b = sph_abs(rsb);
            T ml_b = rsb > T(0) ? T(1) : T(0);
            T rhon_b = PASS ? sR[sj] : rho_b;
            T xab[D], r2 = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                xab[k] = xa[k] - xb[k];
                r2 += xab[k] * xab[k];
            }
            if (!GENERIC) {
                pair_fast<T, D, PASS == 0>(ph, ft, xab, r2, va, vb, rho_b, P_b, rhon_b, rsb > T(0), a_is_i, fs);
            } else {
                PairSide<T, D> sb;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    sb.x[k] = xb[k];
                    sb.v[k] = vb[k];
                    sb.vn[k] = vb[k];
                }
                sb.rho = rho_b;
                sb.P = P_b;
                sb.rho_n = rhon_b;
                sb.ml = ml_b;
                if (use_sps) {
                    L::vel(sBn[sj], sb.vn);
                }
                pair_generic<T, D>(ph, sa, sb, xab, r2, a_is_i, sacc);
            }
        };
        // per-thread list: slist[k * BT + tid]; waddr = shared address of the next free entry
        const uint32_t waddr0 = smem_u32(slist + tid);
        const uint32_t waddr_full = waddr0 + (uint32_t)((LIST_CAP - 4) * BT * 2);   // > : fewer than 4 free
        uint32_t waddr = waddr0;
        auto flush = [&]() {
            if (COMPACT) {
                const int cnt = (int)((waddr - waddr0) / (uint32_t)(BT * 2));
                int m = warp_max(cnt);
                for (int k = 0; k < m; ++k) {
                    if (k < cnt) {
                        unsigned e = slist[k * BT + tid];
                        pair_body((int)(e & (ROLE_BIT - 1)), (e & ROLE_BIT) != 0);
                    }
                }
                waddr = waddr0;
            }
        };

        // ---- stages: the concatenated candidate sequence in pieces of <= cap -----------------
        for (int s0 = 0; s0 < total; s0 += cap) {
            const int s1 = min(s0 + cap, total);
            if (g.use_tma) {
                if (tid == 0) {
                    uint32_t bytes = 0;
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        int lo = max(s_off[r], s0), hi = min(s_off[r + 1], s1);
                        if (lo < hi) bytes += (uint32_t)(hi - lo) * (uint32_t)SS::per_candidate;
                    }
                    if (!use_sps) bytes -= (uint32_t)(s1 - s0) * (uint32_t)SS::esBn;
                    fence_proxy_async();
                    mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
                    for (int r = 0; r < NR; ++r) {
                        int lo = max(s_off[r], s0), hi = min(s_off[r + 1], s1);
                        if (lo < hi) {
                            int n = hi - lo;
                            size_t src = (size_t)s_w0a[r] + (size_t)(lo - s_off[r]);
                            int dst = lo - s0;
                            tma_load_1d(sA + dst, g.A + src, (uint32_t)n * SS::esA, &s_bar);
                            tma_load_1d(sB + dst, g.B + src, (uint32_t)n * SS::esB, &s_bar);
                            if (PASS) tma_load_1d(sR + dst, g.RN + src, (uint32_t)n * SS::esR, &s_bar);
                            if (use_sps) tma_load_1d(sBn + dst, g.Bn + src, (uint32_t)n * SS::esBn, &s_bar);
                        }
                    }
                }
                mbar_wait(&s_bar, phase);
                phase ^= 1u;
            } else {
                for (int r = 0; r < NR; ++r) {
                    int lo = max(s_off[r], s0), hi = min(s_off[r + 1], s1);
                    size_t src = (size_t)s_w0a[r] + (size_t)(lo - s_off[r]);
                    int dst = lo - s0;
                    for (int k = tid; k < hi - lo; k += BT) {
                        sA[dst + k] = g.A[src + k];
                        sB[dst + k] = g.B[src + k];
                        if (PASS) sR[dst + k] = g.RN[src + k];
                        if (use_sps) sBn[dst + k] = g.Bn[src + k];
                    }
                }
                __syncthreads();
            }

            // ---- phase 1 (+ phase 2 when lists fill): walk the rows of this stage ------------
            for (int r = 0; r < NR; ++r) {
                const int lo_s = max(s_off[r], s0), hi_s = min(s_off[r + 1], s1);
                if (lo_s >= hi_s) continue;
                const int dm = (D == 3) ? (r % 3 - 1) : 0;
                const int ds = (D == 3) ? (r / 3 - 1) : (r - 1);
                const int rk = rowbase + (ds * nm + dm) * nx;
                // this lane's window in row r: cells cx-1 .. cx+1 (stale cells, exact pair set)
                int lo = 0, hi = 0;
                if (valid) {
                    lo = g.cell_start[rk + cxi - 1];
                    hi = g.cell_start[rk + cxi + 2];
                }
                // role of the row (row_role): +1 every cell of the row is lower in the reference's
                // cell order (a is "i"), -1 higher (a is "j"), 0: decided per candidate —
                //   role = j < r_lim  and  unsigned(j - r_base) > r_thr
                // the target's own row: b's cell is lower, or same cell and a < b (r_lim = ce_a, r_base = cs_a,
                // r_thr = i - cs_a); a row the fast axis outranks: b's fast-axis cell is lower, or the same
                // and the row's remaining offset says so (r_lim = r_base = start or end of that cell)
                int rowrole, rowpost;
                row_role(g.am, D, dm, ds, &rowrole, &rowpost);
                int r_lim = ce_a, r_base = cs_a;
                unsigned r_thr = (unsigned)(i - cs_a);
                if (rowrole == 0 && (dm != 0 || ds != 0)) {
                    r_lim = r_base = valid ? g.cell_start[rk + cxi + (rowpost > 0 ? 1 : 0)] : 0;
                    r_thr = 0x7fffffffu;
                }
                // staged global index range of this row in this stage
                const int jbase = s_w0a[r] - s_off[r];   // global j = staged position + jbase
                int jb = lo_s + jbase, je = hi_s + jbase;
                // clip to the union of the warp's lane windows (windows are monotone in i), then
                // widen to multiples of 4: the staged piece is 4-aligned at both ends and the
                // per-lane window test rejects the extra candidates, so the unrolled body needs
                // no bounds checks
                // (lane windows are monotone in i and the valid lanes are a prefix of the warp:
                //  the union is [lo of lane 0, hi of the last valid lane))
                int ulo = __shfl_sync(0xffffffffu, lo, 0);
                int uhi = __shfl_sync(0xffffffffu, hi, last_lane);
                if (!warp_has_work) {
                    ulo = INT_MAX;
                    uhi = INT_MIN;
                }
                jb = max(jb, ulo) & ~3;
                je = (min(je, uhi) + 3) & ~3;
                const int sbase = -jbase - s0;          // smem slot = j + sbase
                const unsigned wlen = (unsigned)(hi - lo);
                const T H2 = ph.H2;
                auto cull = [&](auto same_row) {
                    constexpr bool SAME_ROW = decltype(same_row)::value;
                    const unsigned role_const = rowrole > 0 ? (unsigned)ROLE_BIT : 0u;
                    for (int j4 = jb; j4 < je; j4 += 4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = j4 + u;
                            const int sj = j + sbase;
                            T xb[D];
                            L::pos(sA[sj], xb);
                            T r2 = T(0);
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                T dlt = xa[k] - xb[k];
                                r2 += dlt * dlt;
                            }
                            // lane window [lo, hi) in one unsigned compare; the self pair is let
                            // through on the fast path (it contributes exact zeros) and rejected
                            // where a term is non-zero at r = 0 (kernel sums of the generic path)
                            bool ok = (r2 <= H2) & ((unsigned)(j - lo) < wlen);
                            if (GENERIC) ok &= (j != i);
                            // role bit: constant per row, or per candidate (see above)
                            unsigned code = (unsigned)(j + (sbase + (int)role_const));
                            if (SAME_ROW)
                                code = (unsigned)sj | (((j < r_lim) & ((unsigned)(j - r_base) > r_thr)) ? (unsigned)ROLE_BIT : 0u);
                            if (COMPACT) {
                                // predicated append (no branch): store + pointer bump under `ok`
                                asm volatile(
                                    "{\n\t.reg .pred p;\n\t"
                                    "setp.ne.u32 p, %2, 0;\n\t"
                                    "@p st.shared.u16 [%0], %1;\n\t"
                                    "@p add.u32 %0, %0, %3;\n\t}"
                                    : "+r"(waddr)
                                    : "h"((unsigned short)code), "r"((unsigned)ok), "n"(BT * 2)
                                    : "memory");
                            } else if (ok) {
                                pair_body(sj, (code & ROLE_BIT) != 0);
                            }
                        }
                        if (COMPACT) {
                            if (__any_sync(0xffffffffu, waddr > waddr_full)) flush();
                        }
                    }
                };
                if (rowrole == 0) cull(std::true_type{});
                else cull(std::false_type{});
            }
            flush();
            __syncthreads();   // everyone is done with this stage's shared memory
        }

        // ---- epilogue ---------------------------------------------------------------------
        if (!GENERIC) fast_finish<T, D>(ft, fs, drho, acc);
        if (valid) interact_epilogue<T, D, PASS, GENERIC>(g, i, xa, va, rho_a, sacc, drho, acc, red);
        __syncthreads();   // s_brick / s_off reuse
        signal_boundary_brick(g, bidx, PASS);
    }
    if (PASS == 1 && g.epilogue == EPI_FUSED) step_red_commit(g.ctl, red);
}

}  // namespace sph
