// sph_ring.cuh — the neighbour-LIST path of the pair traversal (NeighborLoop! ∘ ComputeInteractions!,
// src/SPHCellList.jl:168-217,268-317): k_list_build, k_list_reorder, k_interact_ring.
//
// Between two list builds the accepted neighbours of a particle barely change, so re-testing the
// ~770 (3D) stencil candidates in every pass (k_interact, sph_interact.cuh) is wasted issue slots.
//   k_list_build     does the cull walk once, without physics: per particle, the window index
//                    (| role bit) of every candidate that passes the reference's stale-cell
//                    window test and lies within H + skin, in global memory, [chunk of 8][particle].
//   k_list_reorder   rewrites every list in the bank-aware "rainbow" order of sph_listorder.h.
//   k_interact_ring  evaluates the pair body over the listed entries only, branch-free (entries
//                    beyond H contribute exact zeros), until k_step_control decides that some pair
//                    could have closed a gap of `skin` (2 x accumulated max displacement) and orders
//                    the next build.  The accepted set is exactly the cull kernel's: r² <= H² is
//                    re-tested on current positions, the window test was applied at build time
//                    (cells do not change between rebuilds).
//
// k_interact_ring is warp-specialised and free of CTA-wide barriers: one producer warp walks the
// brick list (atomic work counter), computes the 3^(D-1) row spans of a brick's candidate window and
// stages them with cp.async.bulk (1-D TMA) into one of NSLOT shared-memory slots, completing on the
// slot's `full` mbarrier; NCW consumer warps pull 32-target sub-bricks of the staged bricks from a
// shared-memory ticket, gather the listed records with LDS.128 and run the pair body, and release a
// slot through its `empty` mbarrier.  The window of brick k+2 is in flight while bricks k and k+1
// are computed, and a warp whose targets have short lists simply takes the next sub-brick — the
// round-1 kernel lost 23 % of its warp time at brick barriers (profiles/r2a_*).
#pragma once

#include "sph_interact.cuh"

// tuning constants of the ring (overridable at compile time for sweeps: scripts/build_variants.sh)
#ifndef SPH_RING_NSLOT
#define SPH_RING_NSLOT 3
#endif
#ifndef SPH_RING_SLOT_KB
#define SPH_RING_SLOT_KB 72
#endif
#ifndef SPH_RING_NCW
#define SPH_RING_NCW 8
#endif
#ifndef SPH_RING_PACKED
#define SPH_RING_PACKED 1
#endif
#ifndef SPH_RING_LIST_PF
#define SPH_RING_LIST_PF 3
#endif

namespace sph {

// Shared-memory geometry of the ring: NSLOT slots of SLOT_BYTES; a slot holds CAP candidate records
// of every staged array, the last 8 of which are the sentinel records of the 8 bank groups.
// CAP is the same for both passes (pass 2 stages ρₙ too: the tighter one decides), because the
// bricks — and the window indices in the lists — are shared by the passes.
template <class T, int D, bool GENERIC>
struct RingGeom {
    using S0 = StageSizes<T, D, 0, GENERIC>;
    using S1 = StageSizes<T, D, 1, GENERIC>;
    static constexpr int NSLOT = SPH_RING_NSLOT;
    static constexpr int SLOT_BYTES = SPH_RING_SLOT_KB * 1024;
    static constexpr int CAP_RAW = (SLOT_BYTES / S1::per_candidate) & ~7;
    static constexpr int CAP = CAP_RAW < 32760 ? CAP_RAW : 32760;   // 15-bit window index
    static constexpr int SMEM = NSLOT * SLOT_BYTES;
    static constexpr int NCW = SPH_RING_NCW;                        // consumer warps
    static constexpr int THREADS = (NCW + 1) * 32;
    // a brick's window (4-aligned row spans) + the 8 sentinels must fit
    static constexpr int WINDOW_LIMIT = CAP - 8 - 6 * ((D == 3) ? 9 : 3);
};

// =================================================================================================
// Per-brick list maintenance (brick_list_decision, sph_control.h).
//   k_cell_vbox     bounding box of the velocities of every cell's particles at this step head
//                   (owned and halo particles alike; 6 floats per cell, rounded outward), into the
//                   buffer ctl->vbox_cur — the other buffer still holds the previous head's boxes;
//   k_brick_bounds  per brick: the diagonal D of the union of both heads' boxes over the brick's
//                   window cells bounds every relative velocity in the window; the brick's
//                   accumulated relative-displacement bound decides whether its lists are rebuilt now.
// Both run every step while lists are in use (~15 us at 1 M particles); k_list_build / k_list_reorder
// then skip the bricks that are not flagged.
// =================================================================================================
// (one thread per cell; a warp-per-cell variant with coalesced reads measured 49 us against 21 us:
//  most of the ~130 k cells of the dense grid are empty and cost it a round of shuffles each)
template <class T, int D>
__global__ void k_cell_vbox(const typename Lay<T, D>::TB *__restrict__ B, const int *__restrict__ cell_start, const GridInfo *grid,
                            const Ctl *ctl, float *__restrict__ vbox, size_t buf_stride) {
    if (ctl->error || ctl->done || ctl->list_build == LIST_BUILD_NONE) return;
    float *const out = vbox + (size_t)ctl->vbox_cur * buf_stride;
    const int ncell = grid->ncell;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x) {
        const int s = cell_start[c], e = cell_start[c + 1];
        float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int j = s; j < e; ++j) {
            T v[D];
            Lay<T, D>::vel(B[j], v);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float vd = sizeof(T) == 8 ? __double2float_rd((double)v[k]) : (float)v[k];
                const float vu = sizeof(T) == 8 ? __double2float_ru((double)v[k]) : (float)v[k];
                lo[k] = fminf(lo[k], vd);
                hi[k] = fmaxf(hi[k], vu);
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            out[(size_t)c * 6 + k] = lo[k];
            out[(size_t)c * 6 + 3 + k] = hi[k];
        }
    }
}

// one WARP per brick: the lanes share the window's cells (a thread per brick spent 0.4 ms in ~750 dependent loads)
template <int D>
__global__ void k_brick_bounds(Ctl *ctl, const GridInfo *grid, const Brick *__restrict__ bricks, const int *__restrict__ ckey,
                               const float *__restrict__ vbox, size_t buf_stride, float *__restrict__ brick_move,
                               int *__restrict__ brick_flag, double skin, double lookahead) {
    if (ctl->error || ctl->done || ctl->list_build == LIST_BUILD_NONE) return;
    constexpr int NR = (D == 3) ? 9 : 3;
    const int nbricks = grid->nbricks, nx = grid->nx, nm = grid->nm;
    const bool all = ctl->list_build == LIST_BUILD_ALL;
    const float vcap = 2.0f * (float)ctl->vmax_now * 1.0001f;   // no two particles differ by more than 2 max|v|
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    int flagged = 0;
    for (int b = warp; b < nbricks; b += nwarps) {
        float move = 0.f;
        int flag = 2;
        if (!all) {
            const Brick br = bricks[b];
            const int key0 = ckey[br.t0], key1 = ckey[br.t1 - 1];
            const int cx0 = key0 % nx, cx1 = key1 % nx;
            const int rowbase = key0 - cx0;
            const int ncx = cx1 - cx0 + 3;               // cells cx0-1 .. cx1+1 of every row
            float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
            for (int q = lane; q < NR * ncx; q += 32) {
                const int r = q / ncx, cx = cx0 - 1 + (q - r * ncx);
                const int dm = (D == 3) ? (r % 3 - 1) : 0;
                const int ds = (D == 3) ? (r / 3 - 1) : (r - 1);
                const size_t cell = (size_t)(rowbase + (ds * nm + dm) * nx + cx);
#pragma unroll
                for (int buf = 0; buf < 2; ++buf) {
                    const float *bx = vbox + (size_t)buf * buf_stride + cell * 6;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        lo[k] = fminf(lo[k], bx[k]);
                        hi[k] = fmaxf(hi[k], bx[3 + k]);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < D; ++k) {
                lo[k] = warp_min(lo[k]);
                hi[k] = warp_max(hi[k]);
            }
            float d2 = 0.f;
#pragma unroll
            for (int k = 0; k < D; ++k) {
                const float w = hi[k] >= lo[k] ? (hi[k] - lo[k]) : 0.f;
                d2 += w * w;
            }
            const float Dd = fminf(sqrtf(d2) * 1.0001f, vcap);
            move = brick_move[b];
            flag = brick_list_decision(&move, Dd, ctl->current_dt, ctl->dt2, skin, lookahead);
        }
        if (lane == 0) {
            brick_move[b] = move;
            brick_flag[b] = flag;            // 2 urgent, 1 due soon, 0 good
            flagged += flag == 2 ? 1 : 0;
        }
    }
    if (lane == 0 && flagged) atomicAdd(&ctl->bricks_urgent, flagged);
}

// =================================================================================================
// List build: the cull walk of k_interact without any physics.  Stages POSITIONS only, applies the
// reference's stale-cell window test and r² <= (H + skin)², and appends (window index | role) to
// the particle's list in global memory.  Runs when k_step_control raises ctl->list_build, on the
// state-n positions, before pass 1 of that step.
// =================================================================================================
template <class T, int D, bool GENERIC, int BT>
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

    __shared__ uint64_t s_bar;
    __shared__ int s_brick;
    __shared__ int s_w0a[NR], s_len[NR], s_off[NR + 1];

    const int tid = threadIdx.x;
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
    const bool flagged_only = g.ctl->list_build == LIST_BUILD_FLAGGED;
    if (flagged_only && g.ctl->bricks_urgent == 0) return;   // every brick's lists are still good: no build in this step

    for (;;) {
        if (tid == 0) s_brick = atomicAdd(&g.ctl->work_counter[6], 1);
        __syncthreads();
        const int bidx = s_brick;
        if (bidx >= nbricks) break;
        if (flagged_only && !g.brick_flag[bidx]) {   // this brick's lists are still good (k_brick_bounds)
            __syncthreads();                         // (s_brick is rewritten at the top of the loop)
            continue;
        }
        if (tid == 0 && g.brick_move) {              // rebuilt at the positions of this step head: the bound restarts
            g.brick_move[bidx] = 0.f;
            atomicAdd(&g.ctl->bricks_flagged, 1);
        }
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
            // the list kernel stages the whole window in one ring slot, whose last 8 records are the sentinels
            if (o > g.list_cap_cand - 8) atomicOr(&g.ctl->list_fail, 1);
        }
        __syncthreads();
        const int total = s_off[NR];
        const unsigned total8 = (unsigned)(g.list_cap_cand - 8);   // padding entry = sentinel of bank group 0 (list_sentinel_base)

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
        // sentinel, otherwise up to 7 entries stay buffered
        auto flush = [&](bool final) {
            const int cnt = (int)((waddr - waddr0) / (uint32_t)(BT * 2));
            const int nchunks = final ? ((cnt + 7) >> 3) : (cnt >> 3);
            for (int c = 0; c < nchunks; ++c) {
                unsigned e[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) e[u] = (c * 8 + u < cnt) ? (unsigned)slist[(c * 8 + u) * BT + tid] : total8;
                if (lcount + 8 <= lcap && valid)
                    gl[(size_t)(lcount >> 3) * g.nl_stride] =
                        make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
                lcount += 8;
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
                int rowrole, rowpost;                        // roles: see k_interact / row_role
                row_role(g.am, D, dm, ds, &rowrole, &rowpost);
                int r_lim = ce_a, r_base = cs_a;
                unsigned r_thr = (unsigned)(i - cs_a);
                if (rowrole == 0 && (dm != 0 || ds != 0)) {
                    r_lim = r_base = valid ? g.cell_start[rk + cxi + (rowpost > 0 ? 1 : 0)] : 0;
                    r_thr = 0x7fffffffu;
                }
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
                // (two instantiations of the walk: the per-candidate role logic of the target's own row
                //  costs 6 issue slots per candidate even when predicated off)
                auto walk = [&](auto same_row_tag) {
                    constexpr bool SAME_ROW = decltype(same_row_tag)::value;
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
                            if (SAME_ROW)
                                code = (unsigned)(j - jbase) | (((j < r_lim) & ((unsigned)(j - r_base) > r_thr)) ? (unsigned)ROLE_BIT : 0u);
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
                };
                if (rowrole == 0) walk(std::true_type{});
                else walk(std::false_type{});
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
// List reorder: every particle's list, as built (window order), is rewritten in place in the
// bank-aware order of sph_listorder.h (rainbow_order is the readable statement of the algorithm and
// what the CPU tests drive; this kernel is the same algorithm with the chunk loop unrolled so that
// every half-word extraction is static, and the global loads prefetched two chunks ahead).
// One thread per particle, a CTA per brick (the lane phase of a particle is its index in the brick,
// mod 8); the reordered list is assembled in shared memory (out[chunk][slot][thread]) and leaves as
// 16-byte stores.  Lists longer than max_slots (option reorder_slots), or with more than REORDER_OVF_CAP entries in
// over-full bank groups, stay as built (valid, only slower).  Runs right after k_list_build.
// =================================================================================================
constexpr int REORDER_OVF_CAP = 32;

template <int BT>
__global__ void __launch_bounds__(BT) k_list_reorder(Ctl *ctl, const GridInfo *grid, const Brick *__restrict__ bricks,
                                                     const int *__restrict__ brick_flag, unsigned sentinel_base, uint4 *nl,
                                                     const int *__restrict__ nl_cnt, size_t nl_stride, int lcap, int max_slots) {
    if (ctl->error || ctl->done || !ctl->list_build || ctl->list_fail) return;
    const bool flagged_only = ctl->list_build == LIST_BUILD_FLAGGED;
    if (flagged_only && ctl->bricks_urgent == 0) return;
    extern __shared__ __align__(16) unsigned short s_out[];   // [max_slots][BT], then the overflow scratch [REORDER_OVF_CAP][BT]
    unsigned short *const s_ovf = s_out + (size_t)max_slots * BT;
    __shared__ int s_brick;
    const int tid = threadIdx.x;
    const int nbricks = grid->nbricks;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_brick = atomicAdd(&ctl->work_counter[7], 1);
        __syncthreads();
        const int bidx = s_brick;
        if (bidx >= nbricks) break;
        if (flagged_only && !brick_flag[bidx]) continue;   // (the barrier at the top of the loop protects s_brick)
        const Brick br = bricks[bidx];
        const unsigned total8 = sentinel_base;   // window indices >= this are sentinels / padding
        for (int i = br.t0 + tid; i < br.t1; i += BT) {
            const int n_slots = min(nl_cnt[i], lcap);
            if (n_slots > max_slots) continue;
            const int nc = n_slots >> 3;
            const unsigned q = (unsigned)(i - br.t0) & 7u;
            uint4 *const lp = nl + i;
            unsigned short *const my_out = s_out + tid;
            unsigned cnt_lo = 0u, cnt_hi = 0u;   // 8 x 8-bit entry counts per bank group
            int novf = 0;
            bool ok = true;
            auto place = [&](unsigned e) {
                const unsigned idx = e & LIST_INDEX_MASK;
                if (idx < total8) {   // (else: padding of the build)
                    const unsigned r = idx & 7u;
                    const unsigned sh = (r & 3u) * 8u;
                    const bool hi = (r & 4u) != 0u;
                    const unsigned c = ((hi ? cnt_hi : cnt_lo) >> sh) & 0xffu;
                    const unsigned inc = 1u << sh;
                    cnt_lo += hi ? 0u : inc;
                    cnt_hi += hi ? inc : 0u;
                    if ((int)c < nc) {
                        my_out[(size_t)(c * 8u + ((r - q) & 7u)) * BT] = (unsigned short)e;
                    } else if (novf < REORDER_OVF_CAP) {
                        s_ovf[(size_t)novf * BT + tid] = (unsigned short)e;
                        ++novf;
                    } else {
                        ok = false;
                    }
                }
            };
            auto place8 = [&](const uint4 &v) {
                place(v.x & 0xffffu); place(v.x >> 16); place(v.y & 0xffffu); place(v.y >> 16);
                place(v.z & 0xffffu); place(v.z >> 16); place(v.w & 0xffffu); place(v.w >> 16);
            };
            uint4 b0 = make_uint4(0, 0, 0, 0), b1 = b0;
            if (nc > 0) b0 = lp[0];
            if (nc > 1) b1 = lp[nl_stride];
            for (int c = 0; c < nc; c += 2) {
                place8(b0);
                if (c + 2 < nc) b0 = lp[(size_t)(c + 2) * nl_stride];
                if (c + 1 < nc) {
                    place8(b1);
                    if (c + 3 < nc) b1 = lp[(size_t)(c + 3) * nl_stride];
                }
            }
            if (!ok) continue;
            // holes: chunks [count of group r, nc) of slot (r - q) & 7 take the over-full groups' entries
            // first, then the group's own sentinel (window index total8 + r)
            int po = 0;
#pragma unroll
            for (unsigned r = 0; r < 8u; ++r) {
                const unsigned c0 = (((r & 4u) ? cnt_hi : cnt_lo) >> ((r & 3u) * 8u)) & 0xffu;
                const unsigned u = (r - q) & 7u;
                for (int c = (int)c0; c < nc; ++c) {
                    unsigned short e = (unsigned short)(total8 + r);
                    if (po < novf) e = s_ovf[(size_t)(po++) * BT + tid];
                    my_out[(size_t)((unsigned)c * 8u + u) * BT] = e;
                }
            }
            for (int c = 0; c < nc; ++c) {
                unsigned e[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) e[u] = my_out[(size_t)(c * 8 + u) * BT];
                lp[(size_t)c * nl_stride] =
                    make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
            }
        }
    }
}

// Test hook (option verify_lists): for every particle of every brick, every candidate of its own
// (stale-cell) window that lies within H at the positions of the pass about to run must be in the
// particle's list.  Brute force over global memory; counts the misses into ctl->list_missing.  This
// is the on-device proof that the per-brick displacement bounds never let a list go stale.
template <class T, int D, bool GENERIC, int BT>
__global__ void __launch_bounds__(BT) k_list_verify(const InteractArgs<T, D> g, int pass) {
    using L = Lay<T, D>;
    constexpr int NR = (D == 3) ? 9 : 3;
    if (g.ctl->error || g.ctl->done || g.ctl->list_fail || g.ctl->list_mode[pass] != LM_USE) return;
    __shared__ int s_w0a[NR], s_len[NR], s_off[NR + 1];
    const int tid = threadIdx.x;
    const int nx = g.grid->nx, nm = g.grid->nm, nbricks = g.grid->nbricks;
    const int npad = (g.grid->n_total + 3) & ~3;
    long long missing = 0;
    for (int bidx = blockIdx.x; bidx < nbricks; bidx += gridDim.x) {
        const Brick br = g.bricks[bidx];
        const int key0 = g.ckey[br.t0], key1 = g.ckey[br.t1 - 1];
        const int cx0 = key0 % nx, cx1 = key1 % nx;
        const int rowbase = key0 - cx0;
        __syncthreads();
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
            for (int r = 0; r < NR; ++r) {
                s_off[r] = o;
                o += s_len[r];
            }
            s_off[NR] = o;
        }
        __syncthreads();
        for (int i = br.t0 + tid; i < br.t1; i += BT) {
            T xa[D];
            L::pos(g.A[i], xa);
            const int ki = g.ckey[i];
            const int cxi = ki - rowbase;
            const int nslots = g.nl_cnt[i];
            for (int r = 0; r < NR; ++r) {
                const int dm = (D == 3) ? (r % 3 - 1) : 0;
                const int ds = (D == 3) ? (r / 3 - 1) : (r - 1);
                const int rk = rowbase + (ds * nm + dm) * nx;
                const int lo = g.cell_start[rk + cxi - 1], hi = g.cell_start[rk + cxi + 2];
                const int jbase = s_w0a[r] - s_off[r];
                for (int j = lo; j < hi; ++j) {
                    if (j == i) continue;
                    T xb[D], r2 = T(0);
                    L::pos(g.A[j], xb);
#pragma unroll
                    for (int k = 0; k < D; ++k) r2 += (xa[k] - xb[k]) * (xa[k] - xb[k]);
                    if (!(r2 <= g.phys.H2)) continue;
                    const unsigned want = (unsigned)(j - jbase);
                    bool found = false;
                    for (int k = 0; k < nslots && !found; ++k) {
                        const uint4 v = g.nl[(size_t)(k >> 3) * g.nl_stride + (size_t)i];
                        const unsigned w = (k & 4) ? ((k & 2) ? v.w : v.z) : ((k & 2) ? v.y : v.x);
                        found = ((((k & 1) ? (w >> 16) : w) & LIST_INDEX_MASK) == want);
                    }
                    missing += found ? 0 : 1;
                }
            }
        }
    }
    if (missing) atomicAdd((unsigned long long *)&g.ctl->list_missing, (unsigned long long)missing);
}

// Diagnostic (sphb200_get_stat "list_wavefronts"): the shared-memory cost model of sph_listorder.h
// evaluated on the lists as they are in memory.  One thread per quarter warp (8 consecutive targets of
// a 32-target sub-brick); for every list position the 8 lanes' window indices are compared: a
// 16-byte gather costs as many wavefronts as the largest number of DISTINCT indices in one bank
// group (index mod 8).  out[0] += wavefronts, out[1] += quarter-warp loads, out[2] += real entries.
__global__ void k_list_diag(const GridInfo *grid, const Brick *__restrict__ bricks, unsigned sentinel_base,
                            const uint4 *__restrict__ nl, const int *__restrict__ nl_cnt, size_t nl_stride,
                            unsigned long long *out) {
    const int nbricks = grid->nbricks;
    unsigned long long wf = 0, loads = 0, real = 0;
    for (int b = blockIdx.x; b < nbricks; b += gridDim.x) {
        const Brick br = bricks[b];
        const unsigned total8 = sentinel_base;
        const int nq = (br.t1 - br.t0 + 7) >> 3;
        for (int qd = threadIdx.x; qd < nq; qd += blockDim.x) {
            const int i0 = br.t0 + qd * 8;
            int cnt[8], m = 0;
            for (int l = 0; l < 8; ++l) {
                cnt[l] = (i0 + l < br.t1) ? nl_cnt[i0 + l] : 0;
                m = max(m, cnt[l]);
            }
            for (int k = 0; k < m; ++k) {
                unsigned idx[8];
                for (int l = 0; l < 8; ++l) {
                    if (k < cnt[l]) {
                        const uint4 v = nl[(size_t)(k >> 3) * nl_stride + (size_t)(i0 + l)];
                        const unsigned w = (k & 4) ? ((k & 2) ? v.w : v.z) : ((k & 2) ? v.y : v.x);
                        idx[l] = ((k & 1) ? (w >> 16) : w) & LIST_INDEX_MASK;
                        real += idx[l] < total8;
                    } else {
                        idx[l] = total8 + (unsigned)((l + k) & 7);   // the padding chunk of the list kernel
                    }
                }
                int worst = 1;
                for (int l = 0; l < 8; ++l) {
                    int distinct = 1;   // distinct indices in idx[l]'s bank group, counted at their first occurrence
                    bool first = true;
                    for (int j = 0; j < l; ++j) first &= idx[j] != idx[l];
                    if (!first) continue;
                    for (int j = l + 1; j < 8; ++j) {
                        bool isnew = (idx[j] & 7u) == (idx[l] & 7u) && idx[j] != idx[l];
                        for (int t = l + 1; t < j; ++t) isnew &= idx[t] != idx[j];
                        distinct += isnew;
                    }
                    worst = max(worst, distinct);
                }
                wf += (unsigned)worst;
                ++loads;
            }
        }
    }
    atomicAdd(&out[0], wf);
    atomicAdd(&out[1], loads);
    atomicAdd(&out[2], real);
}

// =================================================================================================
// pair_fast (sph_physics.cuh) for TWO list entries at once in packed fp32 (sm_100 FFMA2 / FMUL2 /
// FADD2: two IEEE fp32 operations per issue slot).  The list kernel is bound by issue slots, not by
// the fp32 pipe (profiles/r2e: 70 % issue-active, 47 % fma pipe), and FFMA2 runs at half the FFMA
// rate (scripts/ubench: 1.98 vs 3.81 warp-instructions per clock and SM) — same flops, half the
// slots.  Lane-wise the operations, their order and their rounding are exactly those of pair_fast:
// the results are bit-identical to the scalar body.  The two entries' first-level differences and
// the MUFU results are produced by scalar instructions straight into the halves of register pairs.
// S3 is accumulated with the opposite sign (fast_finish2 folds the sign into its scale).
// =================================================================================================
template <int D>
struct FastSums2 {
    float2 s1, s2, s3n[D];
};
template <int D>
__device__ __forceinline__ void fast_zero2(FastSums2<D> &s) {
    s.s1 = s.s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < D; ++k) s.s3n[k] = make_float2(0.f, 0.f);
}
template <int D>
__device__ __forceinline__ void fast_finish2(const FastTarget<float> &a, const FastSums2<D> &s, float &drho, float *acc) {
    drho = a.k_cont * (s.s1.x + s.s1.y) + a.k_ddt * (s.s2.x + s.s2.y);
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = -a.k_acc * (s.s3n[k].x + s.s3n[k].y);
}
struct FastConst2 {   // broadcast pairs of the per-target / per-run constants
    float2 h_inv, m2, eta2, rhon, m_rhon, m_ddt_lin, P, c_visc_n;
};
__device__ __forceinline__ FastConst2 make_fast_const2(const Phys<float> &p, const FastTarget<float> &a) {
    FastConst2 c;
    c.h_inv = make_float2(p.h_inv, p.h_inv);
    c.m2 = make_float2(-2.f, -2.f);
    c.eta2 = make_float2(p.eta2, p.eta2);
    c.rhon = make_float2(a.rhon, a.rhon);
    c.m_rhon = make_float2(-a.rhon, -a.rhon);
    c.m_ddt_lin = make_float2(-p.ddt_lin, -p.ddt_lin);
    c.P = make_float2(a.P, a.P);
    c.c_visc_n = make_float2(-a.c_visc, -a.c_visc);
    return c;
}
// xab / vab: x_a − x_b and v_a − v_b of the two entries; rs_b: their signed densities (sign = fluid);
// rhon_b: their state-n densities (pass 2; ignored when SAME_RHO)
template <int D, bool SAME_RHO>
__device__ __forceinline__ void pair_fast_x2(const FastTarget<float> &a, const FastConst2 &c, const float2 *xab, const float2 *vab,
                                             float2 rs_b, float2 P_b, float2 rhon_b, bool role0, bool role1, FastSums2<D> &s) {
    float2 r2 = __fmul2_rn(xab[0], xab[0]);
    float2 vdotx = __fmul2_rn(vab[0], xab[0]);
#pragma unroll
    for (int k = 1; k < D; ++k) {
        r2 = __ffma2_rn(xab[k], xab[k], r2);
        vdotx = __ffma2_rn(vab[k], xab[k], vdotx);
    }
    const float2 d = make_float2(sph_sqrt_fast(r2.x), sph_sqrt_fast(r2.y));
    float2 qm2 = __ffma2_rn(d, c.h_inv, c.m2);
    qm2.x = fminf(qm2.x, 0.f);
    qm2.y = fminf(qm2.y, 0.f);
    const float2 fac = __fmul2_rn(qm2, __fmul2_rn(qm2, qm2));
    const float2 irb = make_float2(sph_rcp(fabsf(rs_b.x)), sph_rcp(fabsf(rs_b.y)));
    float2 irnb = irb, sum_rhon, dr;
    if (SAME_RHO) {
        sum_rhon = make_float2(a.rhon + fabsf(rs_b.x), a.rhon + fabsf(rs_b.y));
        dr = make_float2(fabsf(rs_b.x) - a.rhon, fabsf(rs_b.y) - a.rhon);
    } else {
        irnb = make_float2(sph_rcp(rhon_b.x), sph_rcp(rhon_b.y));
        sum_rhon = __fadd2_rn(c.rhon, rhon_b);
        dr = __fadd2_rn(rhon_b, c.m_rhon);
    }
    s.s1 = __ffma2_rn(__fmul2_rn(irb, fac), vdotx, s.s1);
    const float2 den = __fmul2_rn(__fadd2_rn(r2, c.eta2), sum_rhon);
    const float2 ip = make_float2(sph_rcp(den.x), sph_rcp(den.y));   // 1 / ((r²+η²)(ρₙ_a+ρₙ_b))
    const float2 inv = __fmul2_rn(ip, sum_rhon);                     // 1 / (r²+η²)
    const float2 diff = __ffma2_rn(c.m_ddt_lin, xab[D - 1], dr);
    float2 V;
    V.x = role0 ? irnb.x : a.inv_rhon;
    V.y = role1 ? irnb.y : a.inv_rhon;
    V.x = rs_b.x > 0.f ? V.x : 0.f;
    V.y = rs_b.y > 0.f ? V.y : 0.f;
    s.s2 = __ffma2_rn(V, __fmul2_rn(diff, __fmul2_rn(__fmul2_rn(fac, r2), inv)), s.s2);
    const float2 mn = make_float2(fminf(vdotx.x, 0.f), fminf(vdotx.y, 0.f));
    const float2 pq = __fmul2_rn(__fadd2_rn(c.P, P_b), irb);
    const float2 c1n = __ffma2_rn(c.c_visc_n, __fmul2_rn(mn, ip), pq);   // −(c_visc min(v·x,0) ip − (P_a+P_b)/ρ_b)
    const float2 cfn = __fmul2_rn(c1n, fac);
#pragma unroll
    for (int k = 0; k < D; ++k) s.s3n[k] = __ffma2_rn(cfn, xab[k], s.s3n[k]);
}

// =================================================================================================
// The LIST kernel (see the file header).
// =================================================================================================
// SPLIT (1 or 4): lanes per target.  A lane normally walks the whole list of its own target; with a few
// thousand fp64 particles that is one long chain per lane on a handful of warps (C1: 27 us per pass on 4
// warps of 54 SMs, each warp alone on its fp64 pipe) while most of the GPU idles.  With SPLIT = 4 a
// sub-brick is 8 targets: lane l serves target l / 4 and takes every 4th list chunk, starting at chunk
// l % 4; the partial sums (linear: FastSums) meet in two shuffles and lane 0 of the quad runs the epilogue.
// Same lists, same accepted pairs; only the order of the summation differs.  Host picks it by size.
template <class T, int D, int PASS, bool GENERIC, int SPLIT = 1>
__global__ void __launch_bounds__(RingGeom<T, D, GENERIC>::THREADS, 1) k_interact_ring(const InteractArgs<T, D> g) {
    using L = Lay<T, D>;
    using TA = typename L::TA;
    using TB = typename L::TB;
    using RG = RingGeom<T, D, GENERIC>;
    using SS = StageSizes<T, D, PASS, GENERIC>;
    constexpr int NR = (D == 3) ? 9 : 3;
    constexpr int NSLOT = RG::NSLOT, CAP = RG::CAP, NCW = RG::NCW;
    constexpr int LIST_PF = SPH_RING_LIST_PF;   // list chunks in flight per lane
    constexpr bool PACKED = !GENERIC && std::is_same<T, float>::value && SPH_RING_PACKED;   // pair_fast_x2
    static_assert(SPLIT == 1 || (SPLIT == 4 && !GENERIC && !PACKED), "split lists: the scalar fast pair body only");
    constexpr int TPS = 32 / SPLIT;             // targets per sub-brick
    constexpr int OFF_B = CAP * SS::esA, OFF_R = OFF_B + CAP * SS::esB, OFF_BN = OFF_R + CAP * SS::esR;
    static_assert(OFF_BN + CAP * SS::esBn <= RG::SLOT_BYTES, "ring slot too small");

    if (g.ctl->error || g.ctl->done) return;
    if (g.ctl->list_mode[PASS] != LM_USE || g.ctl->list_fail) {
        // lean sequence (no cull kernel behind this one): this step's list build overflowed — stop the step here;
        // the host finishes it with the full sequence.  (Every CTA returns whether or not it sees `done` yet.)
        if (PASS == 0 && g.lean_guard && g.ctl->list_fail && blockIdx.x == 0 && threadIdx.x == 0) {
            g.ctl->paused = 2;
            g.ctl->done = 1;
        }
        return;
    }

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_full[NSLOT], s_empty[NSLOT];
    __shared__ int s_meta[NSLOT][4];            // t0, t1, window length (< 0: no more bricks), brick index
    __shared__ int s_sub[NSLOT], s_done[NSLOT];  // sub-brick ticket / finished sub-bricks of the staged brick

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const Phys<T> &ph = g.phys;
    const bool use_sps = GENERIC && PASS && (ph.viscosity == V_SPS);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], NCW);
            s_sub[s] = 0;
            s_done[s] = 0;
        }
        mbar_fence_init();
    }
    // The 8 sentinel records (one per bank group) are the last 8 records of every slot, written once:
    // infinitely far away, so that the clamped kernel factor is an exact zero.  A staged window never
    // reaches them (k_list_build checks), and the lists' padding refers to them by a fixed index.
    const int sent_base = g.list_cap_cand - 8;
    if (tid < 8 * NSLOT) {
        unsigned char *const sbs = smem_raw + (size_t)(tid >> 3) * RG::SLOT_BYTES;
        T far[D], zero[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            far[k] = T(1e15);   // finite: the branch-free pair body must not meet inf * 0
            zero[k] = T(0);
        }
        TA fa;
        TB fb;
        L::pack(fa, fb, far, zero, T(1), T(0));
        const int sj = sent_base + (tid & 7);
        reinterpret_cast<TA *>(sbs)[sj] = fa;
        reinterpret_cast<TB *>(sbs + OFF_B)[sj] = fb;
        if (PASS) reinterpret_cast<T *>(sbs + OFF_R)[sj] = T(1);
        if (use_sps) reinterpret_cast<TB *>(sbs + OFF_BN)[sj] = fb;
    }
    __syncthreads();

    if (warp == NCW) {
        // ======================= producer warp ==================================================
        const int nx = g.grid->nx, nm = g.grid->nm;
        const int brick_first = g.brick_part == 2 ? g.grid->nbricks_bnd : 0;
        const int brick_end = g.brick_part == 1 ? g.grid->nbricks_bnd : g.grid->nbricks;
        const int npad = (g.grid->n_total + 3) & ~3;
        for (int seq = 0;; ++seq) {
            const int slot = seq % NSLOT;
            const uint32_t use = (uint32_t)(seq / NSLOT);
            unsigned char *const sb = smem_raw + (size_t)slot * RG::SLOT_BYTES;
            // next brick and its row spans (while the consumers may still be using the slot)
            int bidx = 0;
            if (lane == 0) bidx = brick_first + atomicAdd(&g.ctl->work_counter[g.counter_slot], 1);
            bidx = __shfl_sync(0xffffffffu, bidx, 0);
            const bool more = bidx < brick_end;
            Brick br = Brick{0, 0};
            int w0a = 0, len = 0;
            if (more) {
                br = g.bricks[bidx];
                const int key0 = g.ckey[br.t0], key1 = g.ckey[br.t1 - 1];
                const int cx0 = key0 % nx, cx1 = key1 % nx;
                const int rowbase = key0 - cx0;
                if (lane < NR) {
                    const int dm = (D == 3) ? (lane % 3 - 1) : 0;
                    const int ds = (D == 3) ? (lane / 3 - 1) : (lane - 1);
                    const int rk = rowbase + (ds * nm + dm) * nx;
                    const int w0 = g.cell_start[rk + cx0 - 1];
                    const int w1 = g.cell_start[rk + cx1 + 2];
                    w0a = w0 & ~3;
                    int w1a = min((w1 + 3) & ~3, npad);
                    if (w1 <= w0) w1a = w0a;   // empty row: stage nothing
                    len = w1a - w0a;
                }
            }
            int off = len;   // inclusive scan over the NR span lengths (identical arithmetic to the list build)
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, off, o);
                if (lane >= o) off += v;
            }
            const int total = __shfl_sync(0xffffffffu, off, NR - 1);
            off -= len;
            if (use > 0) mbar_wait(&s_empty[slot], (use - 1) & 1u);
            if (!more) {
                if (lane == 0) {
                    s_meta[slot][2] = -1;
                    mbar_arrive(&s_full[slot]);
                }
                break;
            }
            TA *const sA = reinterpret_cast<TA *>(sb);
            TB *const sB = reinterpret_cast<TB *>(sb + OFF_B);
            T *const sR = reinterpret_cast<T *>(sb + OFF_R);
            TB *const sBn = reinterpret_cast<TB *>(sb + OFF_BN);
            if (lane == 0) {
                s_meta[slot][0] = br.t0;
                s_meta[slot][1] = br.t1;
                s_meta[slot][2] = total;
                s_meta[slot][3] = bidx;
                s_sub[slot] = 0;
                s_done[slot] = 0;
            }
            __syncwarp();
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)total * (uint32_t)(SS::per_candidate - (use_sps ? 0 : SS::esBn));
                fence_proxy_async();
                mbar_arrive_expect_tx(&s_full[slot], bytes);
            }
            __syncwarp();
            if (lane < NR && len > 0) {
                const size_t src = (size_t)w0a;
                tma_load_1d(sA + off, g.A + src, (uint32_t)len * SS::esA, &s_full[slot]);
                tma_load_1d(sB + off, g.B + src, (uint32_t)len * SS::esB, &s_full[slot]);
                if (PASS) tma_load_1d(sR + off, g.RN + src, (uint32_t)len * SS::esR, &s_full[slot]);
                if (use_sps) tma_load_1d(sBn + off, g.Bn + src, (uint32_t)len * SS::esBn, &s_full[slot]);
            }
        }
        return;
    }

    // =========================== consumer warps =================================================
    const int nbnd = g.bnd_flag ? g.grid->nbricks_bnd : 0;
    const T H2 = ph.H2;
    StepRed<T> red;   // fused corrector (pass 2): Δt / Δx reductions of the new state, one commit per warp
    step_red_zero(red);
    for (int seq = 0;; ++seq) {
        const int slot = seq % NSLOT;
        const uint32_t use = (uint32_t)(seq / NSLOT);
        mbar_wait(&s_full[slot], use & 1u);
        if (s_meta[slot][2] < 0) break;
        const int t0 = s_meta[slot][0], t1 = s_meta[slot][1], bidx = s_meta[slot][3];
        const int nsub = (t1 - t0 + TPS - 1) / TPS;
        const unsigned char *const sb = smem_raw + (size_t)slot * RG::SLOT_BYTES;
        const TA *const sA = reinterpret_cast<const TA *>(sb);
        const TB *const sB = reinterpret_cast<const TB *>(sb + OFF_B);
        const T *const sR = reinterpret_cast<const T *>(sb + OFF_R);
        const TB *const sBn = reinterpret_cast<const TB *>(sb + OFF_BN);
        // this lane's padding chunk: the sentinels of the bank groups its rainbow order expects
        uint4 pad;
        {
            unsigned e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) e[u] = (unsigned)sent_base + (unsigned)((lane + u) & 7);
            pad = make_uint4(e[0] | (e[1] << 16), e[2] | (e[3] << 16), e[4] | (e[5] << 16), e[6] | (e[7] << 16));
        }
        for (;;) {
            int sub = 0;
            if (lane == 0) sub = atomicAdd(&s_sub[slot], 1);
            sub = __shfl_sync(0xffffffffu, sub, 0);
            if (sub >= nsub) break;

            // ---- this lane's target particle -------------------------------------------------
            const int i = t0 + sub * TPS + lane / SPLIT;
            const bool valid = i < t1;
            [[maybe_unused]] const int sq = lane % SPLIT;   // which chunks of the target's list this lane takes
            T xa[D], va[D], rho_a = T(1), P_a = T(0), rhon_a = T(1), ml_a = T(0);
#pragma unroll
            for (int k = 0; k < D; ++k) xa[k] = va[k] = T(0);
            int nchunk = 0;
            EpiloguePrefetch<T, D> epf;
            epf.type = 0;
            if (valid) {
                T rs;
                L::unpack(g.A[i], g.B[i], xa, va, rs, P_a);
                rho_a = sph_abs(rs);
                ml_a = rs > T(0) ? T(1) : T(0);
                rhon_a = PASS ? g.RN[i] : rho_a;
                nchunk = g.nl_cnt[i] >> 3;
                if (g.epilogue == EPI_FUSED) {
                    epf.type = g.type[i];
                    if (PASS) {
                        epf.an = g.An_rw[i];
                        epf.bn = g.Bn_rw[i];
                    }
                }
            }
            // List chunks are prefetched LIST_PF chunks ahead: LIST_PF register buffers, each with its
            // own loop-carried pointer, and the chunk loop unrolled LIST_PF times so that no buffer is
            // ever copied (a copy would wait for the load).  The loads are unconditional — the chunk
            // index is only clamped to the allocation, so a load depends on nothing but the loop; what a
            // lane reads beyond its own list is replaced by the padding chunk when it is consumed.  A
            // pointer advances right BEFORE its next load, i.e. LIST_PF chunk bodies after the previous
            // one: its registers must not be rewritten while that load is still queued behind the
            // shared-memory gathers (a write-after-read stall that cost 40 % of the kernel, profiles/r2c).
            const int last_chunk = (g.lcap >> 3) - 1;
            const size_t lstride = g.nl_stride;
            [[maybe_unused]] const uint4 *pp[LIST_PF];
            [[maybe_unused]] uint4 pf[LIST_PF];
            if constexpr (SPLIT == 1) {
#pragma unroll
                for (int d = 0; d < LIST_PF; ++d) {
                    pp[d] = g.nl + (valid ? i : t0) + (size_t)min(d, last_chunk) * lstride;
                    pf[d] = ld_nc_v4(pp[d]);
                }
            }
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
            const FastTarget<T> ft = make_fast_target<T>(ph, rho_a, P_a, rhon_a, ml_a, PASS == 0);
            FastSums<T, D> fs;
            fast_zero(fs);
            [[maybe_unused]] FastSums2<D> fs2;
            [[maybe_unused]] FastConst2 fc2;
            if constexpr (PACKED) {
                fast_zero2(fs2);
                fc2 = make_fast_const2(ph, ft);
            }

            const int mchunk = warp_max(nchunk);
            // one chunk of 8 entries; lanes whose list has ended run the padding chunk: the warp stays converged
            auto chunk_body = [&](const int c, const uint4 &raw) {
                const uint4 cur = (c < nchunk) ? raw : pad;
                const unsigned w4[4] = {cur.x, cur.y, cur.z, cur.w};
                if (!GENERIC) {
                    // all 16-24 shared-memory gathers of the chunk first, then the 8 branch-free pair
                    // bodies: the loads' latency is paid once per chunk, not once per entry
                    TA a8[8];
                    TB b8[8];
                    T r8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const unsigned w = w4[u >> 1];
                        const int sj = (int)((u & 1) ? ((w >> 16) & LIST_INDEX_MASK) : (w & LIST_INDEX_MASK));
                        a8[u] = sA[sj];
                        b8[u] = sB[sj];
                        if (PASS) r8[u] = sR[sj];
                    }
                    if constexpr (PACKED) {
                        // two entries per packed body (one 32-bit list word = the two entries)
#pragma unroll
                        for (int u = 0; u < 8; u += 2) {
                            const unsigned w = w4[u >> 1];
                            float x0[D], v0[D], x1[D], v1[D], rs0, rs1, P0, P1;
                            L::unpack(a8[u], b8[u], x0, v0, rs0, P0);
                            L::unpack(a8[u + 1], b8[u + 1], x1, v1, rs1, P1);
                            float2 xab2[D], vab2[D];
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                xab2[k] = make_float2(xa[k] - x0[k], xa[k] - x1[k]);
                                vab2[k] = make_float2(va[k] - v0[k], va[k] - v1[k]);
                            }
                            pair_fast_x2<D, PASS == 0>(ft, fc2, xab2, vab2, make_float2(rs0, rs1), make_float2(P0, P1),
                                                       PASS ? make_float2(r8[u], r8[u + 1]) : make_float2(0.f, 0.f),
                                                       (w & 0x8000u) != 0u, (int)w < 0, fs2);
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const unsigned w = w4[u >> 1];
                            const bool a_is_i = (u & 1) ? ((int)w < 0) : ((w & 0x8000u) != 0u);
                            T xb[D], vb[D], rsb, P_b;
                            L::unpack(a8[u], b8[u], xb, vb, rsb, P_b);
                            T xab[D], r2 = T(0);
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                xab[k] = xa[k] - xb[k];
                                r2 += xab[k] * xab[k];
                            }
                            const T rho_b = sph_abs(rsb);
                            pair_fast<T, D, PASS == 0>(ph, ft, xab, r2, va, vb, rho_b, P_b, PASS ? r8[u] : rho_b, rsb > T(0), a_is_i, fs);
                        }
                    }
                } else {
#pragma unroll 2
                    for (int u = 0; u < 8; ++u) {
                        const unsigned w = w4[u >> 1];
                        const unsigned e = (u & 1) ? (w >> 16) : (w & 0xffffu);
                        const int sj = (int)(e & LIST_INDEX_MASK);
                        const bool a_is_i = (e >> 15) != 0;
                        T xb[D], vb[D], rsb, P_b;
                        L::unpack(sA[sj], sB[sj], xb, vb, rsb, P_b);
                        T xab[D], r2 = T(0);
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            xab[k] = xa[k] - xb[k];
                            r2 += xab[k] * xab[k];
                        }
                        if (r2 <= H2) {
                            PairSide<T, D> sbd;
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                sbd.x[k] = xb[k];
                                sbd.v[k] = vb[k];
                                sbd.vn[k] = vb[k];
                            }
                            sbd.rho = sph_abs(rsb);
                            sbd.P = P_b;
                            sbd.rho_n = PASS ? sR[sj] : sbd.rho;
                            sbd.ml = rsb > T(0) ? T(1) : T(0);
                            if (use_sps) L::vel(sBn[sj], sbd.vn);
                            pair_generic<T, D>(ph, sa, sbd, xab, r2, a_is_i, sacc);
                        }
                    }
                }
            };
            if constexpr (SPLIT == 1) {
                const int pf_last = min(last_chunk, mchunk - 1);   // nothing beyond the warp's longest list is fetched from DRAM
                for (int c0 = 0; c0 < mchunk; c0 += LIST_PF) {
#pragma unroll
                    for (int d = 0; d < LIST_PF; ++d) {
                        if (d == 0 || c0 + d < mchunk) {   // (warp-uniform)
                            chunk_body(c0 + d, pf[d]);
                            if (c0 + d + LIST_PF <= pf_last) pp[d] += (size_t)LIST_PF * lstride;   // (else: re-read, an L2 hit)
                            pf[d] = ld_nc_v4(pp[d]);
                        }
                    }
                }
            } else {
                // chunks sq, sq + SPLIT, ... of the target's list; one chunk prefetched ahead
                const uint4 *const lp = g.nl + (valid ? i : t0);
                const int mloop = (mchunk + SPLIT - 1) / SPLIT;     // (mchunk: the warp's longest list)
                uint4 nxt = ld_nc_v4(lp + (size_t)min(sq, last_chunk) * lstride);
                for (int k = 0; k < mloop; ++k) {
                    const uint4 raw = nxt;
                    nxt = ld_nc_v4(lp + (size_t)min(sq + (k + 1) * SPLIT, last_chunk) * lstride);
                    chunk_body(sq + k * SPLIT, raw);               // beyond the list's end: the padding chunk
                }
#pragma unroll
                for (int off = 1; off < SPLIT; off <<= 1) {
                    fs.s1 += __shfl_xor_sync(0xffffffffu, fs.s1, off);
                    fs.s2 += __shfl_xor_sync(0xffffffffu, fs.s2, off);
#pragma unroll
                    for (int k = 0; k < D; ++k) fs.s3[k] += __shfl_xor_sync(0xffffffffu, fs.s3[k], off);
                }
            }
            T drho = T(0), acc[D];
#pragma unroll
            for (int k = 0; k < D; ++k) acc[k] = T(0);
            if constexpr (PACKED) fast_finish2<D>(ft, fs2, drho, acc);
            else if (!GENERIC) fast_finish<T, D>(ft, fs, drho, acc);
            if (valid && (SPLIT == 1 || sq == 0))
                interact_epilogue<T, D, PASS, GENERIC>(g, i, xa, va, rho_a, sacc, drho, acc, red, g.epilogue == EPI_FUSED ? &epf : nullptr);

            if (bidx < nbnd) {
                // slab mode: the warp that retires the last sub-brick of the last boundary brick
                // releases the halo exchange waiting on another stream
                __threadfence();
                __syncwarp();
                if (lane == 0 && atomicAdd(&s_done[slot], 1) + 1 == nsub) {
                    if (atomicAdd(&g.ctl->bnd_done[PASS], 1) + 1 == nbnd) {
                        __threadfence_system();
                        atomicMax(g.bnd_flag, g.bnd_epoch);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[slot]);
    }
    if (PASS == 1 && g.epilogue == EPI_FUSED) step_red_commit(g.ctl, red);
}

}  // namespace sph
