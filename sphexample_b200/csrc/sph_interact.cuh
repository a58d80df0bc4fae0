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
//                    the window and within H + skin as (window index | role) in global memory.
//   k_interact_list  stages the brick's whole window and runs the pair body over the recorded
//                    entries only (~190 instead of ~770 candidates), branch-free.
//
// ctl->list_mode[pass] (k_step_control) says which of k_interact / k_interact_list serves a pass;
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
    int ref_major_is_s;                // 1: slab axis is the reference's most significant axis
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
    int list_cap_cand;                 // candidates the list kernel can stage per brick
    T Hs2;                             // (H + skin)^2: acceptance radius of a list build
    int force_cull;                    // 1: ignore ctl->list_mode (stage-level entry points)
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

// Per-particle tail of a pass: plain stores (stage-level entry points) or the fused symplectic
// half / full update, shared by the cull kernel and the list kernel.
template <class T, int D, int PASS, bool GENERIC>
__device__ __forceinline__ void interact_epilogue(const InteractArgs<T, D> &g, int i, const T *xa, const T *va, T rho_a,
                                                  const PairAccum<T, D> &sacc, T drho, T *acc) {
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
        const uint8_t ty = g.type[i];
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
        } else {
            // LimitDensityAtBoundary!(ρ) + DensityEpsi! + FullTimeStep + Pressure!  (S16-S18, S5)
            const T dt = (T)g.ctl->dt;
            T xn[D], vn[D], rs, Pn;
            L::unpack(g.An_rw[i], g.Bn_rw[i], xn, vn, rs, Pn);
            T rho = sph_abs(rs);
            T gc[D];
#pragma unroll
            for (int k = 0; k < D; ++k) gc[k] = GENERIC ? sacc.gradC[k] : T(0);
            full_step<T, D>(ph, xn, vn, acc, rho, drho, rho_a, gf, ml, dt, gc, GENERIC ? sacc.divr : T(0));
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
    const int lane = tid & 31;
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
                pair_fast<T, D, PASS == 0>(ph, ft, xab, r2, va, vb, rho_b, P_b, rhon_b, rsb > T(0), a_is_i, drho, acc);
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
                // role of the row: +1 b's row is lower in the reference's cell order
                // (a is "i"), -1 higher (a is "j"), 0 same row (decided per candidate)
                int major = g.ref_major_is_s ? ds : dm;
                int minor = g.ref_major_is_s ? dm : ds;
                int rowrole = major != 0 ? -major : -minor;
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
                    const unsigned self_off = (unsigned)(i - cs_a);
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
                            // role bit: constant per row, except in the target's own
                            // row where a is "i" iff b's cell is lower, or same cell and a < b:
                            //   j in [lo, cs_a) or (i, ce_a)  <=>  j < ce_a and not cs_a <= j <= i
                            unsigned code = (unsigned)(j + (sbase + (int)role_const));
                            if (SAME_ROW)
                                code = (unsigned)sj | (((j < ce_a) & ((unsigned)(j - cs_a) > self_off)) ? (unsigned)ROLE_BIT : 0u);
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
        if (valid) interact_epilogue<T, D, PASS, GENERIC>(g, i, xa, va, rho_a, sacc, drho, acc);
        __syncthreads();   // s_brick / s_off reuse
        signal_boundary_brick(g, bidx, PASS);
    }
}

// =================================================================================================
// List build: the cull walk of k_interact without any physics.  Stages POSITIONS only, applies the
// reference's stale-cell window test and r² <= (H + skin)², and appends (window index | role) to
// the particle's list in global memory.  Runs when k_step_control raises ctl->list_build, on the
// state-n positions, before pass 1 of that step.
// =================================================================================================
template <class T, int D, bool GENERIC, int BT, bool ORDER = false>
__global__ void __launch_bounds__(BT) k_list_build(const InteractArgs<T, D> g) {
    using L = Lay<T, D>;
    using TA = typename L::TA;
    constexpr int NR = (D == 3) ? 9 : 3;
    constexpr int esA = sizeof(TA);

    if (g.ctl->error || g.ctl->done || !g.ctl->list_build) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap = g.cap;
    TA *sA = reinterpret_cast<TA *>(smem_raw);
    // per-thread append buffer slist[k * BT + tid] (LIST_CAP entries): accepted candidates are
    // appended branch-free and leave for global memory 8 at a time as 16-byte stores
    unsigned short *slist = reinterpret_cast<unsigned short *>(smem_raw + (size_t)cap * esA);
    unsigned short *slist2 = slist + LIST_CAP * BT;   // scratch column of the bank-aware ordering (list_order only)

    __shared__ uint64_t s_bar;
    __shared__ int s_brick;
    __shared__ int s_w0a[NR], s_len[NR], s_off[NR + 1];

    const int tid = threadIdx.x;
    const Phys<T> &ph = g.phys;
    (void)ph;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t phase = 0;

    const int nx = g.grid->nx, nm = g.grid->nm;
    const int nbricks = g.grid->nbricks;
    const int npad = (g.grid->n_total + 3) & ~3;
    const T Hs2 = g.Hs2;

    for (;;) {
        if (tid == 0) s_brick = atomicAdd(&g.ctl->work_counter[6], 1);
        __syncthreads();
        const int bidx = s_brick;
        if (bidx >= nbricks) break;
        const Brick br = g.bricks[bidx];
        const int key0 = g.ckey[br.t0], key1 = g.ckey[br.t1 - 1];
        const int cx0 = key0 % nx, cx1 = key1 % nx;
        const int rowbase = key0 - cx0;
        if (tid < NR) {
            int dm = (D == 3) ? (tid % 3 - 1) : 0;
            int ds = (D == 3) ? (tid / 3 - 1) : (tid - 1);
            int rk = rowbase + (ds * nm + dm) * nx;
            int w0 = g.cell_start[rk + cx0 - 1];
            int w1 = g.cell_start[rk + cx1 + 2];
            int w0a = w0 & ~3;
            int w1a = min((w1 + 3) & ~3, npad);
            if (w1 <= w0) w1a = w0a;
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
            if (o + 1 > g.list_cap_cand || o >= LIST_IDX_MASK) atomicOr(&g.ctl->list_fail, 1);
        }
        __syncthreads();
        const int total = s_off[NR];

        const int i = br.t0 + tid;
        const bool valid = i < br.t1;
        const int warp_first = br.t0 + (tid & ~31);
        const bool warp_has_work = warp_first < br.t1;
        const int last_lane = min(31, br.t1 - warp_first - 1) & 31;
        T xa[D];
        int cxi = cx0, cs_a = 0, ce_a = 0;
#pragma unroll
        for (int k = 0; k < D; ++k) xa[k] = T(0);
        if (valid) {
            L::pos(g.A[i], xa);
            int ki = g.ckey[i];
            cxi = ki - rowbase;
            cs_a = g.cell_start[ki];
            ce_a = g.cell_start[ki + 1];
        }
        int lcount = 0;                       // list slots already in global memory (multiple of 8)
        uint4 *const gl = g.nl + i;
        const int lcap = g.lcap;
        const uint32_t waddr0 = smem_u32(slist + tid);
        const uint32_t waddr_full = waddr0 + (uint32_t)((LIST_CAP - 4) * BT * 2);   // > : fewer than 4 free
        uint32_t waddr = waddr0;
        // move whole chunks of 8 buffered entries to the global list; `final` pads the tail with the
        // sentinel (window index `total`), otherwise up to 7 entries stay buffered
        auto flush = [&](bool final) {
            const int cnt = (int)((waddr - waddr0) / (uint32_t)(BT * 2));
            const int nchunks = final ? ((cnt + 7) >> 3) : (cnt >> 3);
            if constexpr (!ORDER) {
                // (kept in exactly this form: the instruction stream validated on hardware in round 1)
                for (int c = 0; c < nchunks; ++c) {
                    unsigned e[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) e[u] = (c * 8 + u < cnt) ? (unsigned)slist[(c * 8 + u) * BT + tid] : (unsigned)total;
                    if (lcount + 8 <= lcap && valid)
                        gl[(size_t)(lcount >> 3) * g.nl_stride] =
                            make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
                    lcount += 8;
                }
            } else {
                const int m = final ? cnt : nchunks * 8;   // entries that leave now
                BankRotator rot;
                auto in = [&](int k) -> unsigned { return (unsigned)slist[k * BT + tid]; };
                auto tmp = [&](int p) -> unsigned short & { return slist2[p * BT + tid]; };
                rot.prepare(m, tid, in, tmp);
                for (int c = 0; c < nchunks; ++c) {
                    unsigned e[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int k = c * 8 + u;
                        e[u] = (k < m) ? rot.pull(k, tmp) : (unsigned)total;
                    }
                    if (lcount + 8 <= lcap && valid)
                        gl[(size_t)(lcount >> 3) * g.nl_stride] =
                            make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
                    lcount += 8;
                }
            }
            const int rem = final ? 0 : (cnt & 7);
            for (int u = 0; u < rem; ++u) slist[u * BT + tid] = slist[(nchunks * 8 + u) * BT + tid];
            waddr = waddr0 + (uint32_t)(rem * BT * 2);
        };

        for (int s0 = 0; s0 < total; s0 += cap) {
            const int s1 = min(s0 + cap, total);
            if (tid == 0) {
                uint32_t bytes = 0;
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    int lo = max(s_off[r], s0), hi = min(s_off[r + 1], s1);
                    if (lo < hi) bytes += (uint32_t)(hi - lo) * (uint32_t)esA;
                }
                fence_proxy_async();
                mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    int lo = max(s_off[r], s0), hi = min(s_off[r + 1], s1);
                    if (lo < hi)
                        tma_load_1d(sA + (lo - s0), g.A + ((size_t)s_w0a[r] + (size_t)(lo - s_off[r])), (uint32_t)(hi - lo) * esA, &s_bar);
                }
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1u;

            for (int r = 0; r < NR; ++r) {
                const int lo_s = max(s_off[r], s0), hi_s = min(s_off[r + 1], s1);
                if (lo_s >= hi_s) continue;
                const int dm = (D == 3) ? (r % 3 - 1) : 0;
                const int ds = (D == 3) ? (r / 3 - 1) : (r - 1);
                const int rk = rowbase + (ds * nm + dm) * nx;
                int lo = 0, hi = 0;
                if (valid) {
                    lo = g.cell_start[rk + cxi - 1];
                    hi = g.cell_start[rk + cxi + 2];
                }
                int major = g.ref_major_is_s ? ds : dm;
                int minor = g.ref_major_is_s ? dm : ds;
                int rowrole = major != 0 ? -major : -minor;
                const int jbase = s_w0a[r] - s_off[r];       // global j = window index + jbase
                int jb = lo_s + jbase, je = hi_s + jbase;
                int ulo = __shfl_sync(0xffffffffu, lo, 0);
                int uhi = __shfl_sync(0xffffffffu, hi, last_lane);
                if (!warp_has_work) {
                    ulo = INT_MAX;
                    uhi = INT_MIN;
                }
                jb = max(jb, ulo) & ~3;
                je = (min(je, uhi) + 3) & ~3;
                const int sbase = -jbase - s0;               // smem slot = j + sbase
                const unsigned wlen = (unsigned)(hi - lo);
                const unsigned role_const = rowrole > 0 ? (unsigned)ROLE_BIT : 0u;
                const unsigned self_off = (unsigned)(i - cs_a);
                const bool same_row = rowrole == 0;
                for (int j4 = jb; j4 < je; j4 += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = j4 + u;
                        T xb[D];
                        L::pos(sA[j + sbase], xb);
                        T r2 = T(0);
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            T dlt = xa[k] - xb[k];
                            r2 += dlt * dlt;
                        }
                        bool ok = (r2 <= Hs2) & ((unsigned)(j - lo) < wlen);
                        if (GENERIC) ok &= (j != i);
                        unsigned code = (unsigned)(j + (int)(role_const - (unsigned)jbase));
                        if (same_row)
                            code = (unsigned)(j - jbase) | (((j < ce_a) & ((unsigned)(j - cs_a) > self_off)) ? (unsigned)ROLE_BIT : 0u);
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t"
                            "setp.ne.u32 p, %2, 0;\n\t"
                            "@p st.shared.u16 [%0], %1;\n\t"
                            "@p add.u32 %0, %0, %3;\n\t}"
                            : "+r"(waddr)
                            : "h"((unsigned short)code), "r"((unsigned)ok), "n"(BT * 2)
                            : "memory");
                    }
                    if (__any_sync(0xffffffffu, waddr > waddr_full)) flush(false);
                }
            }
            __syncthreads();   // everyone is done with this stage's shared memory
        }
        flush(true);
        if (valid) {
            if (lcount > lcap) {
                atomicOr(&g.ctl->list_fail, 2);
                lcount = lcap;
            }
            g.nl_cnt[i] = lcount;   // slots incl. the sentinel padding of the last chunk
        }
        __syncthreads();   // s_brick / s_off reuse
    }
}

// =================================================================================================
// The LIST kernel.  Between two list builds the accepted neighbours of a particle barely change,
// so re-testing the ~770 (3D) stencil candidates in every pass is wasted issue slots.  A build
// (k_list_build) records, per particle, the window indices of all candidates that
// pass the reference's stale-cell window test and lie within H + skin; until some pair could have
// closed a gap of `skin` (k_step_control keeps the bound: 2 x accumulated max displacement), this
// kernel evaluates the pair body over those ~180 entries only, ~75 % of which are inside H.
//   * the brick's whole candidate window is staged into shared memory by the same TMA spans as in
//     the cull kernel, so a list entry is a 15-bit window index + the role bit;
//   * list chunks of 8 entries are 16-byte loads, coalesced across the warp ([chunk][particle]);
//   * the accepted set is exactly the cull kernel's: r² <= H² is re-tested on current positions,
//     and the window test was applied at build time (cells do not change between rebuilds).
// =================================================================================================
template <class T, int D, int PASS, bool GENERIC, int BT>
__global__ void __launch_bounds__(BT) k_interact_list(const InteractArgs<T, D> g) {
    using L = Lay<T, D>;
    using TA = typename L::TA;
    using TB = typename L::TB;
    constexpr int NR = (D == 3) ? 9 : 3;
    using SS = StageSizes<T, D, PASS, GENERIC>;

    if (g.ctl->error || g.ctl->done) return;
    if (g.ctl->list_mode[PASS] != LM_USE || g.ctl->list_fail) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int cap = g.list_cap_cand;
    TA *sA = reinterpret_cast<TA *>(smem_raw);
    TB *sB = reinterpret_cast<TB *>(smem_raw + (size_t)cap * SS::esA);
    T *sR = reinterpret_cast<T *>(smem_raw + (size_t)cap * (SS::esA + SS::esB));
    TB *sBn = reinterpret_cast<TB *>(smem_raw + (size_t)cap * (SS::esA + SS::esB + SS::esR));

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

    for (;;) {
        if (tid == 0) s_brick = brick_first + atomicAdd(&g.ctl->work_counter[g.counter_slot], 1);
        __syncthreads();
        const int bidx = s_brick;
        if (bidx >= brick_end) break;
        const Brick br = g.bricks[bidx];
        const int key0 = g.ckey[br.t0], key1 = g.ckey[br.t1 - 1];
        const int cx0 = key0 % nx, cx1 = key1 % nx;
        const int rowbase = key0 - cx0;
        // candidate spans: identical arithmetic to the cull kernel (the list indexes this window)
        if (tid < NR) {
            int dm = (D == 3) ? (tid % 3 - 1) : 0;
            int ds = (D == 3) ? (tid / 3 - 1) : (tid - 1);
            int rk = rowbase + (ds * nm + dm) * nx;
            int w0 = g.cell_start[rk + cx0 - 1];
            int w1 = g.cell_start[rk + cx1 + 2];
            int w0a = w0 & ~3;
            int w1a = min((w1 + 3) & ~3, npad);
            if (w1 <= w0) w1a = w0a;
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
            // stage the whole window
            uint32_t bytes = (uint32_t)o * (uint32_t)(SS::per_candidate - (use_sps ? 0 : SS::esBn));
            if (bytes) {
                fence_proxy_async();
                mbar_arrive_expect_tx(&s_bar, bytes);
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const int len = s_len[r];
                    if (len > 0) {
                        const size_t src = (size_t)s_w0a[r];
                        const int dst = s_off[r];
                        tma_load_1d(sA + dst, g.A + src, (uint32_t)len * SS::esA, &s_bar);
                        tma_load_1d(sB + dst, g.B + src, (uint32_t)len * SS::esB, &s_bar);
                        if (PASS) tma_load_1d(sR + dst, g.RN + src, (uint32_t)len * SS::esR, &s_bar);
                        if (use_sps) tma_load_1d(sBn + dst, g.Bn + src, (uint32_t)len * SS::esBn, &s_bar);
                    }
                }
            }
        }
        __syncthreads();
        const int total = s_off[NR];

        // ---- this thread's target particle (overlaps the TMA flight) -----------------------
        const int i = br.t0 + tid;
        const bool valid = i < br.t1;
        T xa[D], va[D], rho_a = T(1), P_a = T(0), rhon_a = T(1), ml_a = T(0);
#pragma unroll
        for (int k = 0; k < D; ++k) xa[k] = va[k] = T(0);
        int nchunk = 0;
        if (valid) {
            T rs;
            L::unpack(g.A[i], g.B[i], xa, va, rs, P_a);
            rho_a = sph_abs(rs);
            ml_a = rs > T(0) ? T(1) : T(0);
            rhon_a = PASS ? g.RN[i] : rho_a;
            nchunk = (g.nl_cnt[i] + 7) >> 3;
        }
        const uint4 *lp = g.nl + i;
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (nchunk > 0) nxt = lp[0];
        PairSide<T, D> sa;
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
        const FastTarget<T> ft = make_fast_target<T>(ph, rho_a, P_a, rhon_a, ml_a, PASS == 0);

        if (total > 0) {
            mbar_wait(&s_bar, phase);
            phase ^= 1u;
        }
        // sentinel candidate (window index `total`): infinitely far away, so r² <= H² fails
        if (tid == 0) {
            T far[D], zero[D];
#pragma unroll
            for (int k = 0; k < D; ++k) {
                far[k] = T(1e15);   // finite: the masked pair body must not meet inf * 0
                zero[k] = T(0);
            }
            TA fa;
            TB fb;
            L::pack(fa, fb, far, zero, T(1), T(0));
            sA[total] = fa;
            sB[total] = fb;
            if (PASS) sR[total] = T(1);
        }
        __syncthreads();

        const T H2 = ph.H2;
        auto pair_entry = [&](unsigned e) {
            const int sj = (int)(e & LIST_IDX_MASK);
            const bool a_is_i = (e >> 15) != 0;
            T xb[D], vb[D], rsb, P_b;
            L::unpack(sA[sj], sB[sj], xb, vb, rsb, P_b);
            T xab[D], r2 = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                xab[k] = xa[k] - xb[k];
                r2 += xab[k] * xab[k];
            }
            T rho_b = sph_abs(rsb);
            T rhon_b = PASS ? sR[sj] : rho_b;
            if (!GENERIC) {
                // branch-free: the 8 entries of a chunk interleave (the kernel is latency-bound at the
                // occupancy its shared-memory window allows); entries outside H contribute exact zeros
                pair_fast<T, D, PASS == 0, true>(ph, ft, xab, r2, va, vb, rho_b, P_b, rhon_b, rsb > T(0), a_is_i, drho, acc);
            } else if (r2 <= H2) {
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
                sb.ml = rsb > T(0) ? T(1) : T(0);
                if (use_sps) L::vel(sBn[sj], sb.vn);
                pair_generic<T, D>(ph, sa, sb, xab, r2, a_is_i, sacc);
            }
        };
        const int mchunk = warp_max(nchunk);
        for (int c = 0; c < mchunk; ++c) {
            const uint4 cur = nxt;
            if (c + 1 < nchunk) nxt = lp[(size_t)(c + 1) * g.nl_stride];
            if (c < nchunk) {
                const unsigned e8[8] = {cur.x & 0xffffu, cur.x >> 16, cur.y & 0xffffu, cur.y >> 16,
                                        cur.z & 0xffffu, cur.z >> 16, cur.w & 0xffffu, cur.w >> 16};
                if (!GENERIC) {
                    // all 16-24 shared-memory gathers of the chunk first, then the 8 masked pair
                    // bodies: the loads' latency is paid once per chunk, not once per entry
                    TA a8[8];
                    TB b8[8];
                    T r8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int sj = (int)(e8[u] & LIST_IDX_MASK);
                        a8[u] = sA[sj];
                        b8[u] = sB[sj];
                        if (PASS) r8[u] = sR[sj];
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        T xb[D], vb[D], rsb, P_b;
                        L::unpack(a8[u], b8[u], xb, vb, rsb, P_b);
                        T xab[D], r2 = T(0);
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            xab[k] = xa[k] - xb[k];
                            r2 += xab[k] * xab[k];
                        }
                        const T rho_b = sph_abs(rsb);
                        pair_fast<T, D, PASS == 0, true>(ph, ft, xab, r2, va, vb, rho_b, P_b, PASS ? r8[u] : rho_b, rsb > T(0),
                                                         (e8[u] >> 15) != 0, drho, acc);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) pair_entry(e8[u]);
                }
            }
        }
        if (valid) interact_epilogue<T, D, PASS, GENERIC>(g, i, xa, va, rho_a, sacc, drho, acc);
        __syncthreads();   // shared window / s_brick reuse
        signal_boundary_brick(g, bidx, PASS);
    }
}

}  // namespace sph
