// sph_cells.cuh — UpdateNeighbors! on the device (src/SPHCellList.jl:56-61,118-163):
// particle -> cell hash, stable counting-sort reorder, dense cell start table, brick list.
//
// All kernels are predicated on ctl->do_rebuild so that the whole step sequence can be
// enqueued (or graph-captured) without the host knowing whether this step rebuilds.
//
// The reference sorts the particle table by CartesianIndex (last dimension most significant,
// x fastest) with a stable sort.  Here the key is
//     key = ((c_s - cmin_s) * nm + (c_m - cmin_m)) * nx + (c_x - cmin_x)
// on a dense grid over the bounding box of occupied cells padded by one cell on every side
// (so the 3^D stencil never leaves the table).  s is the slab axis (most significant so that
// slab halos are contiguous index ranges); with s = last dimension this is exactly the
// reference's ordering.  Stability: an atomic counting sort places particles of one cell in
// arbitrary order, then k_stable_rank re-ranks each cell's members by their previous index,
// which is what a stable sort would have produced — deterministic run to run.
#pragma once

#include "sph_bricks.h"
#include "sph_device.cuh"

namespace sph {

// Which table entries take part in a rebuild (slab mode, sph_slab_impl.cuh).  Disabled: all of
// [0, n).  Enabled: only [p0, p1) and [q0, n) are live, and a live particle whose slab coordinate
// lies outside [keep_lo, keep_hi) is dropped (it has been handed to a neighbour rank).  Dropped
// entries sort into a trash bucket behind the last cell and fall off the table.
struct SlabFilter {
    int enabled;
    int p0, p1, q0;
    int keep_lo, keep_hi;
};
constexpr int DEAD_COORD = INT_MIN;

// map_floor, src/SPHCellList.jl:56-61: sign(x) * trunc(muladd(|x|, H⁻¹, 0.5)), evaluated in
// double for both storage precisions so that an fp32 run bins exactly like the fp64 oracle fed
// the same (fp32-representable) positions.
__device__ __forceinline__ int map_floor_dev(double x, double inv_cutoff, int &bad) {
    double t = trunc(fma(fabs(x), inv_cutoff, 0.5));
    if (!(t < 1.0e9)) {   // also catches NaN
        bad = 1;
        t = 0.0;
    }
    int s = (x > 0.0) - (x < 0.0);
    return s * (int)t;
}

template <class T, int D>
__global__ void k_cell_bbox(const typename Lay<T, D>::TA *__restrict__ A, int n, double inv_cutoff,
                            int *__restrict__ ccoord, Ctl *ctl, GridInfo *grid, AxisMap am, SlabFilter flt) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    int lo[D], hi[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        lo[k] = INT_MAX;
        hi[k] = INT_MIN;
    }
    int bad = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        bool live = !flt.enabled || (i >= flt.p0 && i < flt.p1) || i >= flt.q0;
        int c[D];
        if (live) {
            T x[D];
            Lay<T, D>::pos(A[i], x);
#pragma unroll
            for (int k = 0; k < D; ++k) c[k] = map_floor_dev((double)x[k], inv_cutoff, bad);
            if (flt.enabled && (c[am.ax_s] < flt.keep_lo || c[am.ax_s] >= flt.keep_hi)) live = false;
        }
        if (!live) {
            ccoord[(size_t)i * D] = DEAD_COORD;
            continue;
        }
#pragma unroll
        for (int k = 0; k < D; ++k) {
            ccoord[(size_t)i * D + k] = c[k];
            lo[k] = min(lo[k], c[k]);
            hi[k] = max(hi[k], c[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        lo[k] = warp_min(lo[k]);
        hi[k] = warp_max(hi[k]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (lo[k] <= hi[k]) {
                atomicMin(&grid->bb_min[k], lo[k]);
                atomicMax(&grid->bb_max[k], hi[k]);
            }
        }
    }
    if (bad) atomicCAS(&ctl->error, 0, SPH_ERR_ENUMERIC);
}

// One thread: dense grid extents from the bounding box.  own_lo/own_hi: owned cell-coordinate
// range [lo, hi) along the slab axis (INT_MIN/INT_MAX when not decomposed).
template <int D>
__global__ void k_grid_setup(Ctl *ctl, GridInfo *grid, AxisMap am, long long cell_cap, long long row_cap,
                             int own_lo, int own_hi) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    int ext[3] = {1, 1, 1};
    for (int k = 0; k < D; ++k) {
        long long e = (long long)grid->bb_max[k] - (long long)grid->bb_min[k] + 3;
        if (e < 3 || e > (1ll << 30)) {
            ctl->error = SPH_ERR_ECAPACITY;
            return;
        }
        ext[k] = (int)e;
        grid->cmin[k] = grid->bb_min[k] - 1;
    }
    int nx = ext[am.ax_f];
    int nm = (D == 3) ? ext[am.ax_m] : 1;
    int ns = ext[am.ax_s];
    long long ncell = (long long)nx * nm * ns;
    long long nrows = (long long)nm * ns;
    if (ncell + 1 > cell_cap || nrows > row_cap) {
        ctl->error = SPH_ERR_ECAPACITY;
        return;
    }
    grid->nx = nx;
    grid->nm = nm;
    grid->ns = ns;
    grid->ncell = (int)ncell;
    grid->nrows = (int)nrows;
    grid->nbricks = 0;   // k_build_bricks appends
    grid->nbricks_bnd = 0;
    // owned rows: slab coordinate c_s in [own_lo, own_hi)
    long long s0 = (long long)own_lo - grid->cmin[am.ax_s];
    long long s1 = (long long)own_hi - grid->cmin[am.ax_s];
    if (own_lo == INT_MIN) s0 = 0;
    if (own_hi == INT_MAX) s1 = ns;
    s0 = s0 < 0 ? 0 : (s0 > ns ? ns : s0);
    s1 = s1 < s0 ? s0 : (s1 > ns ? ns : s1);
    grid->own_row0 = (int)(s0 * nm);
    grid->own_row1 = (int)(s1 * nm);
}

__global__ void k_zero_counts(const Ctl *ctl, const GridInfo *grid, int *__restrict__ cell_count) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    int n = grid->ncell + 1;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) cell_count[c] = 0;
}

template <int D>
__global__ void k_cell_count(const int *__restrict__ ccoord, int n, AxisMap am, const Ctl *ctl,
                             const GridInfo *grid, int *__restrict__ key_out, int *__restrict__ slot_out,
                             int *__restrict__ cell_count) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    const int nx = grid->nx, nm = grid->nm;
    const int cx0 = grid->cmin[am.ax_f], cm0 = (D == 3) ? grid->cmin[am.ax_m] : 0, cs0 = grid->cmin[am.ax_s];
    const int trash = grid->ncell;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (ccoord[(size_t)i * D] == DEAD_COORD) {
            key_out[i] = trash;
            slot_out[i] = atomicAdd(&cell_count[trash], 1);
            continue;
        }
        int cx = ccoord[(size_t)i * D + am.ax_f] - cx0;
        int cm = (D == 3) ? ccoord[(size_t)i * D + am.ax_m] - cm0 : 0;
        int cs = ccoord[(size_t)i * D + am.ax_s] - cs0;
        int key = (cs * nm + cm) * nx + cx;
        key_out[i] = key;
        slot_out[i] = atomicAdd(&cell_count[key], 1);
    }
}

// ---- exclusive scan of cell_count[0 .. ncell] into cell_start (three small kernels) ----------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *smem_warp, int &total) {
    // inclusive warp scan
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) smem_warp[w] = inc;
    __syncthreads();
    int nw = blockDim.x >> 5;
    if (w == 0) {
        int s = lane < nw ? smem_warp[lane] : 0;
        int si = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += u;
        }
        if (lane < nw) smem_warp[lane] = si - s;   // exclusive warp offsets
        if (lane == 31) smem_warp[32] = si;        // block total
    }
    __syncthreads();
    int res = smem_warp[w] + inc - v;
    total = smem_warp[32];
    __syncthreads();
    return res;
}

__global__ void k_scan_partials(const Ctl *ctl, const GridInfo *grid, const int *__restrict__ cell_count,
                                int *__restrict__ partial) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    __shared__ int sw[33];
    int n = grid->ncell + 1;
    int base = blockIdx.x * SCAN_CHUNK;
    if (base >= n) return;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int c = base + k * SCAN_THREADS + threadIdx.x;
        if (c < n) s += cell_count[c];
    }
    int total;
    block_exclusive_scan(s, sw, total);
    if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void k_scan_top(const Ctl *ctl, const GridInfo *grid, int *__restrict__ partial) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    __shared__ int sw[33];
    int n = grid->ncell + 1;
    int nblk = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    int carry = 0;
    for (int b0 = 0; b0 < nblk; b0 += blockDim.x) {
        int b = b0 + threadIdx.x;
        int v = b < nblk ? partial[b] : 0;
        int total;
        int ex = block_exclusive_scan(v, sw, total);
        if (b < nblk) partial[b] = carry + ex;
        carry += total;
    }
}

__global__ void k_scan_final(const Ctl *ctl, const GridInfo *grid, const int *__restrict__ cell_count,
                             const int *__restrict__ partial, int *__restrict__ cell_start) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    __shared__ int sw[33];
    int n = grid->ncell + 1;
    int base = blockIdx.x * SCAN_CHUNK;
    if (base >= n) return;
    // thread t owns SCAN_ITEMS consecutive cells
    int c0 = base + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (c0 + k < n) ? cell_count[c0 + k] : 0;
        s += v[k];
    }
    int total;
    int ex = block_exclusive_scan(s, sw, total) + partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (c0 + k < n) cell_start[c0 + k] = ex;
        ex += v[k];
    }
}

__global__ void k_scatter_unstable(const Ctl *ctl, const int *__restrict__ key, const int *__restrict__ slot, int n,
                                   const int *__restrict__ cell_start, int *__restrict__ tmp_idx) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        tmp_idx[cell_start[key[i]] + slot[i]] = i;
}

// Stable order inside each cell.  The reference's stable sort keeps, inside a cell, the order of
// the table before the sort, and that table was ordered by (previous cell in the reference's
// column-major cell order, order inside that cell).  The per-particle ORDER KEY carries exactly
// this pair — (reference cell key of the previous rebuild) << 10 | (rank inside that cell) — so
// the within-cell order (which decides the density-diffusion roles of same-cell pairs)
// is reproduced even when this table is sorted with another major axis or split over ranks.
// On one GPU with the default axes it coincides with "ascending previous index".
constexpr int OKEY_RANK_BITS = 10;
constexpr int OKEY_AXIS_BITS = 18;
template <int D>
__device__ __forceinline__ unsigned long long order_key(const int *c, int rank) {
    unsigned long long k = 0;
    const int off = 1 << (OKEY_AXIS_BITS - 1), lim = (1 << OKEY_AXIS_BITS) - 1;
#pragma unroll
    for (int a = D - 1; a >= 0; --a) k = (k << OKEY_AXIS_BITS) | (unsigned long long)min(max(c[a] + off, 0), lim);
    return (k << OKEY_RANK_BITS) | (unsigned long long)min(rank, (1 << OKEY_RANK_BITS) - 1);
}

__global__ void k_stable_rank(const Ctl *ctl, const GridInfo *grid, const int *__restrict__ key,
                              const int *__restrict__ tmp_idx, int n, const int *__restrict__ cell_start,
                              const unsigned long long *__restrict__ okey, int *__restrict__ perm) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    const int nlive = cell_start[grid->ncell];   // entries behind it are the trash bucket
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nlive; p += gridDim.x * blockDim.x) {
        int v = tmp_idx[p];
        int c = key[v];
        int s = cell_start[c], e = cell_start[c + 1];
        const unsigned long long kv = okey[v];
        int rank = 0;
        for (int q = s; q < e; ++q) {
            int u = tmp_idx[q];
            unsigned long long ku = okey[u];
            rank += (ku < kv) | ((ku == kv) & (u < v));
        }
        perm[s + rank] = v;
    }
}

// gather the particle table into scratch (new order), then copy back
template <class T, int D>
struct Table {
    typename Lay<T, D>::TA *A;
    typename Lay<T, D>::TB *B;
    typename Lay<T, D>::TV *acc;
    typename Lay<T, D>::TV *ghost;   // may be null
    long long *id;
    unsigned long long *group;
    unsigned long long *okey;        // within-cell order key, see k_stable_rank
    uint8_t *type;
    int *ckey;
};

template <class T, int D>
__global__ void k_gather_table(const Ctl *ctl, const GridInfo *grid, const int *__restrict__ cell_start,
                               const int *__restrict__ perm, Table<T, D> src, Table<T, D> dst,
                               const int *__restrict__ key_prev_order, const int *__restrict__ ccoord_prev_order) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    const int n = cell_start[grid->ncell];
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        int s = perm[p];
        dst.A[p] = src.A[s];
        dst.B[p] = src.B[s];
        dst.acc[p] = src.acc[s];
        if (src.ghost) dst.ghost[p] = src.ghost[s];
        dst.id[p] = src.id[s];
        dst.group[p] = src.group[s];
        dst.type[p] = src.type[s];
        const int key = key_prev_order[s];
        dst.ckey[p] = key;
        int c[D];
#pragma unroll
        for (int k = 0; k < D; ++k) c[k] = ccoord_prev_order[(size_t)s * D + k];
        dst.okey[p] = order_key<D>(c, p - cell_start[key]);
    }
}

template <class T, int D>
__global__ void k_copy_table(const Ctl *ctl, const GridInfo *grid, const int *__restrict__ cell_start,
                             Table<T, D> src, Table<T, D> dst) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    const int n = cell_start[grid->ncell];
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        dst.A[p] = src.A[p];
        dst.B[p] = src.B[p];
        dst.acc[p] = src.acc[p];
        if (src.ghost) dst.ghost[p] = src.ghost[p];
        dst.id[p] = src.id[p];
        dst.group[p] = src.group[p];
        dst.type[p] = src.type[p];
        dst.okey[p] = src.okey[p];
        dst.ckey[p] = src.ckey[p];
    }
}

// Brick list: every owned row (c_m, c_s) of cells is cut into segments of consecutive particles;
// one brick is the unit of work of the interaction kernels.  A brick closes when it holds `bt`
// particles (possibly in the middle of a cell) or when taking in the next cell would push its
// candidate window — the particles of cells [first-1, last+1] of the 3^(D-1) neighbouring rows,
// the thing the interaction kernels stage into shared memory — beyond `wlimit`.  The second rule
// keeps sparse rows (two tank walls 100 cells apart in one row) from producing bricks whose
// window spans the whole row.  One thread per row, two passes (count, write); bricks are appended
// through an atomic counter, so their order (not their content) varies from run to run.
// `part` selects the rows: 0 = all owned rows; 1 = the first and last owned slab layer (their
// particles are what the neighbour ranks hold as halo, so these bricks are computed first and their
// results travel while the interior is computed); 2 = the layers in between.
template <int D>
__global__ void k_build_bricks(Ctl *ctl, GridInfo *grid, const int *__restrict__ cell_start, int bt, int wlimit,
                               Brick *__restrict__ bricks, int brick_cap, int part) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    constexpr int NR = (D == 3) ? 9 : 3;
    const int nx = grid->nx, nm = grid->nm;
    const int o0 = grid->own_row0, o1 = grid->own_row1;
    const int i0 = min(o0 + nm, o1), i1 = max(o1 - nm, i0);     // interior rows [i0, i1)
    int r0 = o0, r1 = o1, skip0 = 0, skip1 = 0;                 // rows [r0, r1) minus [skip0, skip1)
    if (part == 1) { skip0 = i0; skip1 = i1; }
    if (part == 2) { r0 = i0; r1 = i1; }
    for (int r = r0 + blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += gridDim.x * blockDim.x) {
        if (r >= skip0 && r < skip1) continue;
        const int rowbase = r * nx;
        const int p0 = cell_start[rowbase], p1 = cell_start[rowbase + nx];
        if (p1 <= p0) continue;
        int roff[NR];
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            int dm = (D == 3) ? (q % 3 - 1) : 0;
            int ds = (D == 3) ? (q / 3 - 1) : (q - 1);
            roff[q] = rowbase + (ds * nm + dm) * nx;
        }
        // two walks of the same deterministic sequence: count, reserve, write
        int out = 0, nb = 0;
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1) {
                out = atomicAdd(&grid->nbricks, nb);
                if (out + nb > brick_cap) {
                    atomicCAS(&ctl->error, 0, SPH_ERR_ECAPACITY);
                    break;
                }
            }
            walk_row_bricks<NR>(cell_start, rowbase, nx, roff, bt, wlimit, [&](int t0, int t1) {
                if (pass) bricks[out++] = Brick{t0, t1};
                else ++nb;
            });
        }
    }
}

__global__ void k_mark_boundary_bricks(const Ctl *ctl, GridInfo *grid) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    grid->nbricks_bnd = grid->nbricks;
}

// last kernel of the rebuild sequence: table layout for the slab exchange, and a successful
// rebuild clears the request
__global__ void k_finish_rebuild(Ctl *ctl, GridInfo *grid, const int *__restrict__ cell_start, int count_rebuild) {
    if (ctl->error || ctl->done || !ctl->do_rebuild) return;
    const int nx = grid->nx, nm = grid->nm;
    const int r0 = grid->own_row0, r1 = grid->own_row1;
    grid->own_p0 = cell_start[(size_t)r0 * nx];
    grid->own_p1 = cell_start[(size_t)r1 * nx];
    grid->own_l1 = cell_start[(size_t)min(r0 + nm, r1) * nx];
    grid->own_l2 = cell_start[(size_t)max(r1 - nm, r0) * nx];
    grid->n_total = cell_start[grid->ncell];
    ctl->n_rebuilds += count_rebuild;
    ctl->do_rebuild = 0;
}

}  // namespace sph
