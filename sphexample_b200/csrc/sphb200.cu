// sphb200.cu — host side of libsphb200.so: the C-ABI of include/sphb200.h on top of the sm_100a
// kernels in sph_cells.cuh / sph_interact.cuh / sph_step.cuh.
//
// One handle = one CUDA device + one stream.  The step sequence of SimulationLoop
// (src/SPHCellList.jl:742-802) is enqueued without host round trips: Δt, Δx, the rebuild
// decision and the loop condition live in a device-resident control block (sph::Ctl), rebuild
// kernels are predicated on it, and the host only synchronises once per batch of steps.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <chrono>
#include <numeric>
#include <thread>
#include <string>
#include <vector>

#include "../../include/sphb200.h"
#include "sph_cells.cuh"
#include "sph_interact.cuh"
#include "sph_ring.cuh"
#include "sph_slab.cuh"
#include "sph_step.cuh"

using namespace sph;

static thread_local std::string g_create_error;

struct sphb200_sim {
    std::string err;
    virtual ~sphb200_sim() {}
    virtual int upload(int64_t n, const void *pos, const void *vel, const void *acc, const void *rho, const uint8_t *type,
                       const uint64_t *group, const int64_t *id, const void *ghost, const void *gnorm) = 0;
    virtual int download(int order, void *pos, void *vel, void *acc, void *rho, void *press, int64_t *id, uint8_t *type,
                         uint64_t *group, int64_t *cells) = 0;
    virtual int64_t num_particles() const = 0;
    virtual int set_time(double t, int64_t it) = 0;
    virtual int simulation_loop(double t_next, sphb200_report *rep) = 0;
    virtual int step(int64_t n, int reset_dx, sphb200_report *rep) = 0;
    virtual int get_report(sphb200_report *rep) = 0;
    virtual int64_t launch_count() const = 0;
    virtual int update_neighbors(int64_t *index_counter) = 0;
    virtual int get_cell_list(int64_t *n_cells, int64_t *cells, int64_t *start) = 0;
    virtual int pressure(int half) = 0;
    virtual int neighbor_loop(int pass, void *drhodt_out, void *acc_out) = 0;
    virtual int delta_t(double *dt) = 0;
    virtual int progress_motion(double dt2) = 0;
    virtual int apply_mdbc() = 0;
    virtual int half_time_step(double dt2) = 0;
    virtual int full_time_step(double dt) = 0;
    virtual int download_half(void *ph, void *vh, void *rh, void *prh) = 0;
    virtual int download_aux(void *gradc, void *divr, void *ksum, void *kgrad) = 0;
    virtual int set_stream(void *stream) = 0;
    virtual int set_option(const char *name, double value) = 0;
    virtual int get_stat(const char *name, double *value) = 0;
    virtual int comm_init(const uint8_t *id, int rank, int world, int axis) = 0;
    virtual int set_slab(int64_t lo, int64_t hi) = 0;
    virtual int set_ghost_nodes(int64_t ng, const void *points, const int64_t *ids) = 0;
    virtual int column_histogram(int axis, int64_t *cell_min, int64_t *n_columns, int64_t *counts, int64_t cap) = 0;
    virtual int stage_times(double *ms_out, int n) = 0;
};

namespace {

static int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    // enlarge to `count` elements keeping the first `keep` (stream-ordered copy, then free)
    cudaError_t grow(size_t count, size_t keep, cudaStream_t st) {
        if (count <= n && p) return cudaSuccess;
        T *q = nullptr;
        cudaError_t e = cudaMalloc((void **)&q, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        if (p && keep) {
            e = cudaMemcpyAsync(q, p, std::min(keep, n) * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        if (p) cudaFree(p);
        p = q;
        n = count;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
};

template <class T, int D>
__global__ void k_pack_upload(int n, const T *__restrict__ pos, const T *__restrict__ vel, const T *__restrict__ acc,
                              const T *__restrict__ rho, const T *__restrict__ ghost, const uint8_t *__restrict__ type,
                              Phys<T> ph, typename Lay<T, D>::TA *A, typename Lay<T, D>::TB *B,
                              typename Lay<T, D>::TV *accv, typename Lay<T, D>::TV *ghostv,
                              const long long *__restrict__ ids, unsigned long long *okey) {
    using L = Lay<T, D>;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T x[D], v[D], a[D], gp[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            x[k] = pos[(size_t)i * D + k];
            v[k] = vel ? vel[(size_t)i * D + k] : T(0);
            a[k] = acc ? acc[(size_t)i * D + k] : T(0);
            gp[k] = ghost ? ghost[(size_t)i * D + k] : T(0);
        }
        T r = rho[i];
        typename L::TA oa;
        typename L::TB ob;
        L::pack(oa, ob, x, v, type[i] == 1 ? r : -r, eos_gamma7(ph, r));   // Pressure! of RunSimulation, :835
        A[i] = oa;
        B[i] = ob;
        accv[i] = L::mkv(a);
        if (ghostv) ghostv[i] = L::mkv(gp);
        // initial within-cell order = table order (single GPU) / ascending ID (slab mode, where
        // each rank holds a subset of the reference's ID-sorted table, src/PreProcess.jl:116)
        okey[i] = ids ? (unsigned long long)ids[i] : (unsigned long long)i;
    }
}

// upload helpers: defaults for the optional columns and the bounding box, without host passes over the table
__global__ void k_fill_defaults(int n, unsigned long long *group, long long *id) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (group) group[i] = 1ull;
        if (id) id[i] = (long long)i + 1;
    }
}
// min / max of every coordinate -> out[0..2] = min, out[3..5] = max, out[6] != 0: a non-finite coordinate
template <class T, int D>
__global__ void k_upload_bbox(int n, const T *__restrict__ pos, double *out) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    bool bad = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const double x = (double)pos[(size_t)i * D + k];
            bad |= !(x == x) || fabs(x) > 1e300;
            lo[k] = fmin(lo[k], x);
            hi[k] = fmax(hi[k], x);
        }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = warp_min(lo[k]);
        hi[k] = warp_max(hi[k]);
    }
    bad = __any_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < D; ++k) {
            // atomic min / max on doubles by compare-and-swap
            unsigned long long *plo = (unsigned long long *)&out[k], *phi = (unsigned long long *)&out[3 + k];
            unsigned long long old = *plo, assumed;
            do {
                assumed = old;
                if (!(lo[k] < __longlong_as_double((long long)assumed))) break;
                old = atomicCAS(plo, assumed, (unsigned long long)__double_as_longlong(lo[k]));
            } while (old != assumed);
            old = *phi;
            do {
                assumed = old;
                if (!(hi[k] > __longlong_as_double((long long)assumed))) break;
                old = atomicCAS(phi, assumed, (unsigned long long)__double_as_longlong(hi[k]));
            } while (old != assumed);
        }
        if (bad) out[6] = 1.0;
    }
}

template <class T, int D>
__global__ void k_unpack_download(int n, const typename Lay<T, D>::TA *__restrict__ A, const typename Lay<T, D>::TB *__restrict__ B,
                                  const typename Lay<T, D>::TV *__restrict__ accv, T *pos, T *vel, T *acc, T *rho, T *press) {
    using L = Lay<T, D>;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T x[D], v[D], a[D], rs, P;
        L::unpack(A[i], B[i], x, v, rs, P);
#pragma unroll
        for (int k = 0; k < D; ++k) a[k] = T(0);
        if (accv) L::getv(accv[i], a);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            if (pos) pos[(size_t)i * D + k] = x[k];
            if (vel) vel[(size_t)i * D + k] = v[k];
            if (acc) acc[(size_t)i * D + k] = a[k];
        }
        if (rho) rho[i] = sph_abs(rs);
        if (press) press[i] = P;
    }
}

template <class T, int D>
__global__ void k_unpack_vec(int n, const typename Lay<T, D>::TV *__restrict__ src, T *dst) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        T a[D];
        Lay<T, D>::getv(src[i], a);
#pragma unroll
        for (int k = 0; k < D; ++k) dst[(size_t)i * D + k] = a[k];
    }
}

template <class T, int D>
class Sim final : public sphb200_sim {
    using L = Lay<T, D>;
    using TA = typename L::TA;
    using TB = typename L::TB;
    using TV = typename L::TV;
    static constexpr int BT = 128;

  public:
    sphb200_params prm;
    Phys<T> ph;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    int64_t n = 0;        // particles held (owned + halo)
    size_t n_alloc = 0;
    bool have_cells = false, have_half = false, uploaded = false;
    int64_t launches = 0;
    // options
    int opt_compact, opt_tma, opt_smem_kb, opt_batch;
    int opt_lists, opt_lcap, opt_list_smem_kb;   // per-particle neighbour lists (sph_ring.cuh)
    int opt_list_reorder;                        // bank-aware entry order (sph_listorder.h, k_list_reorder)
    int opt_list_local;                          // per-brick list validity (brick_list_decision) instead of one global bound
    int opt_list_local_auto = 1;                 // ... chosen at upload from the particle count unless set explicitly
    int opt_verify_lists = 0;                    // test hook: count listed-pair misses before every pass (k_list_verify)
    int opt_list_lookahead = 3;                  // bricks due within this many steps are rebuilt along with the urgent ones
    int opt_build_smem_kb = 10;                  // staged positions per stage of k_list_build (r3b: 8-14 KB -> 2.76-2.80 ms first step, 24 KB 2.99)
    int opt_reorder_slots = 224;                 // longest list k_list_reorder handles (its shared memory: (slots + 32) * 256 B per CTA)
    DevBuf<float> vbox, brick_move;              // per-cell velocity boxes (2 buffers x 6 floats), per-brick displacement bounds
    DevBuf<int> brick_flag;
    double opt_skin;                             // list skin as a fraction of H
    DevBuf<uint4> nl;
    DevBuf<int> nl_cnt;
    size_t nl_stride = 0;
    int cull_force = 1;   // the cull kernel ignores ctl->list_mode (lists off / stage-level calls)
    bool snapshot_in_epilogue = false;   // this step's ρₙ snapshot is taken by the pass-1 epilogue
    int brick_part = 0;   // which bricks the next interaction launches take: 0 all, 1 slab boundary, 2 interior
    unsigned *bnd_flag = nullptr;   // slab mode: flag the next interaction launches raise after their boundary bricks
    unsigned bnd_epoch = 0;
    bool generic = false;
    AxisMap am;
    int own_lo = INT_MIN, own_hi = INT_MAX;
    // particle table (cell-sorted) and scratch copy for the reorder
    DevBuf<TA> A, A2, Ah;
    DevBuf<TB> B, B2, Bh;
    DevBuf<TV> acc, acc2, ghost, ghost2, gradC, kgrad;
    DevBuf<T> RN, drhodt, divr, ksum, rho_new;
    DevBuf<long long> id, id2;
    DevBuf<unsigned long long> group, group2, okey, okey2;
    DevBuf<uint8_t> type, type2, has_new;
    // slab-mode mDBC: the global ghost-node table (static) and the per-node solves that are all-reduced
    DevBuf<TV> g_point;
    DevBuf<long long> g_id;
    DevBuf<double> g_sol;
    int n_ghost_nodes = -1;   // -1: no table handed over yet
    DevBuf<int> ckey, ckey2, ccoord, key_tmp, slot_tmp, tmp_idx, perm;
    // cell structure
    DevBuf<int> cell_count, cell_start, scan_partial;
    DevBuf<Brick> bricks;
    // cudaFuncSetAttribute / occupancy results are per device and per handle: kernel -> {dynamic smem, CTAs per SM}
    std::map<const void *, std::pair<int, int>> kernel_cfg;
    long long cell_cap = 0, row_cap = 0;
    int brick_cap = 0;
    DevBuf<Ctl> d_ctl;
    DevBuf<GridInfo> d_grid;
    Ctl *h_ctl = nullptr;        // pinned mirrors
    GridInfo *h_grid = nullptr;
    DevBuf<unsigned char> stage;   // raw staging for upload / download
    MotionTable motions;
    SlabComm slab;

    Sim(const sphb200_params &p, int dev) : prm(p), device(dev) {
        // defaults from the r1 sweeps on B200 (profiles/): the kernel is issue-bound and wants
        // >= 6 CTAs/SM, i.e. a small staged window; with the MUFU-based fp32 pair body the
        // divergent single-phase walk beats the two-phase lists, in fp64 the lists win
        opt_compact = env_int("SPHB200_COMPACT", sizeof(T) == 8 ? 1 : 0);
        opt_tma = env_int("SPHB200_TMA", 1);
        opt_smem_kb = env_int("SPHB200_SMEM_KB", sizeof(T) == 8 ? 40 : 24);
        opt_batch = env_int("SPHB200_BATCH", 64);
        opt_graph = env_int("SPHB200_GRAPH", 1);
        // measured on B200 (profiles/r2r_configs.jsonl): the conditional nodes cost as much as the ~17 empty
        // kernels they replace (C1 57 vs 60, C2 304 vs 317, C5 19.6 vs 20.4 Mpu/s) — off by default
        opt_graph_cond = env_int("SPHB200_GRAPH_COND", 0);
        // measured on B200 (profiles/r4b_small_launches.txt): an empty predicated kernel costs 2.4 us, the 13 of the
        // UpdateNeighbors! chain a third of a C1 step -> steps that will not rebuild replay a graph without the chain
        opt_lean = env_int("SPHB200_LEAN", 1);
        // lists: fp32 only by default (an fp64 3D window does not fit shared memory), never with
        // PlanarShifting (the shifting displacement is not covered by the |v| dt bound)
        // (2D fp64 windows fit too: C2 186 -> 325 Mpu/s, profiles/r1m_configs.jsonl)
        opt_lists = env_int("SPHB200_LISTS", (sizeof(T) == 4 || D == 2) ? 1 : 0);
        opt_lcap = env_int("SPHB200_LCAP", D == 3 ? 320 : 96);
        opt_list_smem_kb = env_int("SPHB200_LIST_SMEM_KB", 0);    // > 0: pretend a ring slot holds only this much (tests of the overflow fallback)
        opt_skin = env_int("SPHB200_SKIN_PCT", 4) * 0.01;       // r2q sweep with per-brick rebuilds: 3 % .. 10 % -> 871 / 874 / 867 / 858 / 848 / 818 Mpu/s
        opt_list_reorder = env_int("SPHB200_LIST_REORDER", 1);
        // per-brick maintenance costs two latency-bound kernels per step (~17 us) and saves most of the list
        // builds: worth it at 1 M particles (815 -> 874 Mpu/s, profiles/r2n), a loss on the small cases (C1 80 vs
        // 92, C2 369 vs 405, C5 36 vs 40 Mpu/s, profiles/r4c_small_profile*.jsonl) -> chosen at upload by size
        opt_list_local = env_int("SPHB200_LIST_LOCAL", -1);
        opt_list_local_auto = opt_list_local < 0;
        if (opt_list_local_auto) opt_list_local = 1;
        opt_list_lookahead = env_int("SPHB200_LIST_LOOKAHEAD", 3);
        am.ax_f = 0;       // default: the reference's own cell order (x fastest, last component most significant)
        am.ax_s = D - 1;
        am.ax_m = (D == 3) ? 1 : 0;
        build_phys();
    }
    ~Sim() override {
        drop_step_graph();
        if (slab.comm) {
            if (!slab.dead) nccl::api().CommDestroy(slab.comm);
            else if (nccl::api().CommAbort) nccl::api().CommAbort(slab.comm);
        }
        if (slab.xstream) cudaStreamDestroy(slab.xstream);
        if (slab.ev_bnd) cudaEventDestroy(slab.ev_bnd);
        if (slab.ev_x) cudaEventDestroy(slab.ev_x);
        if (slab.d_flag) cudaFree(slab.d_flag);
        if (slab.d_counts) cudaFree(slab.d_counts);
        if (slab.h_counts) cudaFreeHost(slab.h_counts);
        if (h_ctl) cudaFreeHost(h_ctl);
        if (h_grid) cudaFreeHost(h_grid);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }

    int fail(int code, const char *fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(SPHB200_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

    void build_phys() {
        const sphb200_params &p = prm;
        ph = phys_from_params<T>(p);
        generic = !(p.kernel == SPHB200_KERNEL_WENDLANDC2 && p.viscosity == SPHB200_VISC_ARTIFICIAL &&
                    p.diffusion == SPHB200_DDT_LINEAR && !p.shifting && !p.kernel_output);
        memset(&motions, 0, sizeof motions);
        motions.n = p.n_motions;
        for (int k = 0; k < p.n_motions && k < SPHB200_MAX_MOTIONS; ++k) {
            motions.group[k] = (unsigned long long)p.motions[k].group_marker;
            motions.velocity[k] = p.motions[k].velocity;
            motions.start[k] = p.motions[k].start_time;
            motions.duration[k] = p.motions[k].duration;
            for (int d = 0; d < 3; ++d) motions.dir[k][d] = p.motions[k].direction[d];
        }
    }

    int init() {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(SPHB200_ECUDA, "device %d is sm_%d%d; libsphb200 is built for sm_100a only", device, prop.major, prop.minor);
        num_sms = prop.multiProcessorCount;
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
        CK(d_ctl.alloc(1));
        CK(d_grid.alloc(1));
        CK(cudaMallocHost((void **)&h_ctl, sizeof(Ctl)));
        CK(cudaMallocHost((void **)&h_grid, sizeof(GridInfo)));
        memset(h_ctl, 0, sizeof(Ctl));
        memset(h_grid, 0, sizeof(GridInfo));
        CK(cudaMemcpyAsync(d_ctl.p, h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_grid.p, h_grid, sizeof(GridInfo), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        return SPHB200_OK;
    }

    int set_stream(void *s) override {
        CK(cudaSetDevice(device));
        drop_step_graph();
        CK(cudaStreamSynchronize(stream));
        if (own_stream && stream) cudaStreamDestroy(stream);
        stream = (cudaStream_t)s;
        own_stream = false;
        return SPHB200_OK;
    }
    int set_option(const char *name, double value) override {
        std::string k(name ? name : "");
        drop_step_graph();
        if (k == "graph") { opt_graph = (int)value; return SPHB200_OK; }
        if (k == "graph_cond") { opt_graph_cond = (int)value; return SPHB200_OK; }
        if (k == "lean") { opt_lean = (int)value; return SPHB200_OK; }
        if (k == "split") { opt_split = (int)value; return SPHB200_OK; }
        if (k == "brick_targets") { opt_brick_targets = ((int)value + 31) & ~31; return SPHB200_OK; }   // 0 = by particle count (takes effect at the next UpdateNeighbors!)
        if (k == "test_fail_list_build_at") { opt_test_fail_at = (int64_t)value; return SPHB200_OK; }   // n-th lean step (graph off)
        if (k == "compact") opt_compact = (int)value;
        else if (k == "tma") opt_tma = (int)value;
        else if (k == "smem_kb") opt_smem_kb = (int)value;
        else if (k == "batch") opt_batch = std::max(1, (int)value);
        else if (k == "generic") generic = generic || value != 0.0;
        else if (k == "lists") opt_lists = (int)value;
        else if (k == "skin") opt_skin = value;
        else if (k == "lcap") opt_lcap = std::max(8, ((int)value + 7) & ~7);
        else if (k == "list_smem_kb") opt_list_smem_kb = (int)value;
        else if (k == "list_reorder") opt_list_reorder = (int)value;
        else if (k == "list_local") { opt_list_local = (int)value; opt_list_local_auto = 0; }
        else if (k == "verify_lists") opt_verify_lists = (int)value;
        else if (k == "list_lookahead") opt_list_lookahead = std::max(0, (int)value);
        else if (k == "build_smem_kb") opt_build_smem_kb = std::max(8, (int)value);
        else if (k == "reorder_slots") opt_reorder_slots = std::max(64, ((int)value + 7) & ~7);
        else return fail(SPHB200_EINVAL, "unknown option '%s'", k.c_str());
        return SPHB200_OK;
    }

    int get_stat(const char *name, double *value) override {
        std::string k(name ? name : "");
        if (!value) return fail(SPHB200_EINVAL, "get_stat: null output");
        CK(cudaSetDevice(device));
        int rc = sync_ctl();
        if (rc) return rc;
        if (k == "lean_steps") *value = (double)n_lean_steps;       // steps replayed without the UpdateNeighbors! chain
        else if (k == "lean_pauses") *value = (double)n_lean_pauses;   // ... of which had to be finished with it after all
        else if (k == "list_builds") *value = h_ctl->list_build_equiv;        // in units of "every brick once"
        else if (k == "list_build_steps") *value = h_ctl->n_list_builds;   // steps in which at least one brick was rebuilt
        else if (k == "list_missing") *value = (double)h_ctl->list_missing;
        else if (k == "list_off") *value = h_ctl->list_off || h_ctl->list_fail;
        else if (k == "list_fail_reason") *value = h_ctl->list_fail ? h_ctl->list_fail : h_ctl->list_fail_last;
        else if (k == "halo_bytes_per_step") *value = (double)slab.halo_bytes_per_step;
        else if (k == "migrated") *value = (double)slab.n_migrated;
        else if (k == "n_total") *value = (double)n;
        else if (k == "list_wavefronts" || k == "list_entries") {
            // model cost of the lists in memory: wavefronts per quarter-warp 16-byte gather (1.0 = conflict-free)
            if (!lists_on() || !h_ctl->list_valid) { *value = 0.0; return SPHB200_OK; }
            DevBuf<unsigned long long> acc3;
            CK(acc3.alloc(3));
            CK(cudaMemsetAsync(acc3.p, 0, 24, stream));
            k_list_diag<<<num_sms * 4, 128, 0, stream>>>(d_grid.p, bricks.p, (unsigned)(list_cap_cand() - 8), nl.p, nl_cnt.p, nl_stride, acc3.p);
            ++launches;
            unsigned long long h3[3];
            CK(cudaMemcpyAsync(h3, acc3.p, 24, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            *value = (k == "list_entries") ? (double)h3[2] / std::max<double>(1.0, (double)num_particles())
                                           : (double)h3[0] / std::max<double>(1.0, (double)h3[1]);
        }
        else return fail(SPHB200_EINVAL, "unknown stat '%s'", k.c_str());
        return SPHB200_OK;
    }

    int grid_for(int64_t count, int threads = 256) const {
        int64_t b = (count + threads - 1) / threads;
        return (int)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)num_sms * 16));
    }

    // ------------------------------------------------------------------ allocation
    int alloc_particles(int64_t count) {
        size_t na = (size_t)((count + 3) & ~3ll) + 8;
        if (na <= n_alloc) return SPHB200_OK;
        drop_step_graph();
        na = na + na / 8;   // headroom for slab migration
        CK(A.alloc(na)); CK(A2.alloc(na)); CK(Ah.alloc(na));
        CK(B.alloc(na)); CK(B2.alloc(na)); CK(Bh.alloc(na));
        CK(acc.alloc(na)); CK(acc2.alloc(na));
        CK(RN.alloc(na)); CK(drhodt.alloc(na));
        CK(id.alloc(na)); CK(id2.alloc(na)); CK(group.alloc(na)); CK(group2.alloc(na));
        CK(okey.alloc(na)); CK(okey2.alloc(na));
        CK(type.alloc(na)); CK(type2.alloc(na));
        CK(ckey.alloc(na)); CK(ckey2.alloc(na)); CK(ccoord.alloc(na * D));
        CK(key_tmp.alloc(na)); CK(slot_tmp.alloc(na)); CK(tmp_idx.alloc(na)); CK(perm.alloc(na));
        if (prm.mdbc) { CK(ghost.alloc(na)); CK(ghost2.alloc(na)); CK(rho_new.alloc(na)); CK(has_new.alloc(na)); }
        if (prm.shifting) { CK(gradC.alloc(na)); CK(divr.alloc(na)); }
        if (prm.kernel_output) { CK(ksum.alloc(na)); CK(kgrad.alloc(na)); }
        CK(stage.alloc(na * (size_t)(sizeof(T) * (4 * D + 2) + 8) + 256));
        // zero everything the TMA staging may touch beyond n (finite padding)
        CK(cudaMemsetAsync(A.p, 0, na * sizeof(TA), stream)); CK(cudaMemsetAsync(Ah.p, 0, na * sizeof(TA), stream));
        CK(cudaMemsetAsync(B.p, 0, na * sizeof(TB), stream)); CK(cudaMemsetAsync(Bh.p, 0, na * sizeof(TB), stream));
        CK(cudaMemsetAsync(RN.p, 0, na * sizeof(T), stream));
        CK(cudaMemsetAsync(acc.p, 0, na * sizeof(TV), stream));
        n_alloc = na;
        return SPHB200_OK;
    }
    // slab mode: make room for `count` table entries, keeping what the table and the send scratch hold
    int grow_particles(int64_t count) {
        size_t na = (size_t)((count + 3) & ~3ll) + 8;
        if (na <= n_alloc) return SPHB200_OK;
        na = na + na / 4;
        const size_t keep = n_alloc;
        CK(A.grow(na, keep, stream)); CK(A2.grow(na, keep, stream)); CK(Ah.grow(na, keep, stream));
        CK(B.grow(na, keep, stream)); CK(B2.grow(na, keep, stream)); CK(Bh.grow(na, keep, stream));
        CK(acc.grow(na, keep, stream)); CK(acc2.grow(na, keep, stream));
        CK(RN.grow(na, keep, stream)); CK(drhodt.grow(na, keep, stream));
        CK(id.grow(na, keep, stream)); CK(id2.grow(na, keep, stream));
        CK(group.grow(na, keep, stream)); CK(group2.grow(na, keep, stream));
        CK(okey.grow(na, keep, stream)); CK(okey2.grow(na, keep, stream));
        CK(type.grow(na, keep, stream)); CK(type2.grow(na, keep, stream));
        CK(ckey.grow(na, keep, stream)); CK(ckey2.grow(na, 0, stream)); CK(ccoord.grow(na * D, 0, stream));
        CK(key_tmp.grow(na, 0, stream)); CK(slot_tmp.grow(na, 0, stream)); CK(tmp_idx.grow(na, 0, stream)); CK(perm.grow(na, 0, stream));
        if (prm.shifting) { CK(gradC.grow(na, keep, stream)); CK(divr.grow(na, keep, stream)); }
        if (prm.kernel_output) { CK(ksum.grow(na, keep, stream)); CK(kgrad.grow(na, keep, stream)); }
        CK(stage.grow(na * (size_t)(sizeof(T) * (4 * D + 2) + 8) + 256, 0, stream));
        n_alloc = na;
        {
            int rc = ensure_lists();
            if (rc) return rc;
        }
        brick_cap = 0;   // re-sized with the cell tables
        long long cc = cell_cap;
        cell_cap = 0;
        return alloc_cells(cc);
    }
    int alloc_cells(long long cells_needed) {
        long long cap = std::max<long long>(cells_needed, 4096);
        if (cap > cell_cap) drop_step_graph();
        if (cap > (1ll << 28)) return fail(SPHB200_ECAPACITY, "cell grid of %lld cells exceeds the dense-grid limit", cap);
        if (cap <= cell_cap) return SPHB200_OK;
        CK(cell_count.alloc((size_t)cap + 8));
        CK(cell_start.alloc((size_t)cap + 8));
        CK(scan_partial.alloc((size_t)(cap / SCAN_CHUNK + 2)));
        cell_cap = cap;
        row_cap = cap / 3 + 1;
        brick_cap = (int)std::min<long long>((long long)(n_alloc / 8) + row_cap + 16, INT_MAX);
        CK(bricks.alloc((size_t)brick_cap));
        CK(brick_move.alloc((size_t)brick_cap));
        CK(brick_flag.alloc((size_t)brick_cap));
        CK(vbox.alloc((size_t)(cap + 8) * 12));
        return SPHB200_OK;
    }

    // ------------------------------------------------------------------ state transfer
    int upload(int64_t count, const void *pos, const void *vel, const void *accel, const void *rho, const uint8_t *ty,
               const uint64_t *grp, const int64_t *ids, const void *gp, const void *gn) override {
        (void)gn;
        if (count < 1 || count > (int64_t)INT_MAX / 8 || !pos || !rho || !ty)
            return fail(SPHB200_EINVAL, "upload: need n >= 1, position, density and type");
        CK(cudaSetDevice(device));
        if (count != n) drop_step_graph();   // n is a kernel argument (a re-upload of the same size keeps the captured steps:
                                             // buffers, counts and options are what they were; reallocations drop them themselves)
        if (opt_list_local_auto) opt_list_local = count >= 131072 ? 1 : 0;
        int rc = alloc_particles(count);
        if (rc) return rc;
        if ((rc = ensure_lists())) return rc;
        n = count;
        // raw arrays -> staging -> packed layout on the device
        unsigned char *sp = stage.p;
        size_t vb = (size_t)count * D * sizeof(T), sb = (size_t)count * sizeof(T);
        T *d_pos = (T *)sp; sp += vb;
        T *d_vel = (T *)sp; sp += vb;
        T *d_acc = (T *)sp; sp += vb;
        T *d_gp = (T *)sp; sp += vb;
        T *d_rho = (T *)sp; sp += sb;
        CK(cudaMemcpyAsync(d_pos, pos, vb, cudaMemcpyHostToDevice, stream));
        if (vel) CK(cudaMemcpyAsync(d_vel, vel, vb, cudaMemcpyHostToDevice, stream));
        if (accel) CK(cudaMemcpyAsync(d_acc, accel, vb, cudaMemcpyHostToDevice, stream));
        if (gp && prm.mdbc) CK(cudaMemcpyAsync(d_gp, gp, vb, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(d_rho, rho, sb, cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(type.p, ty, (size_t)count, cudaMemcpyHostToDevice, stream));
        if (grp) CK(cudaMemcpyAsync(group.p, grp, (size_t)count * 8, cudaMemcpyHostToDevice, stream));
        if (ids) CK(cudaMemcpyAsync(id.p, ids, (size_t)count * 8, cudaMemcpyHostToDevice, stream));
        if (!grp || !ids) {   // defaults: GroupMarker 1, ID = 1 .. N
            k_fill_defaults<<<grid_for(count), 256, 0, stream>>>((int)count, grp ? nullptr : group.p, ids ? nullptr : id.p);
            ++launches;
        }
        k_pack_upload<T, D><<<grid_for(count), 256, 0, stream>>>((int)count, d_pos, vel ? d_vel : nullptr,
                                                                accel ? d_acc : nullptr, d_rho,
                                                                (gp && prm.mdbc) ? d_gp : nullptr, type.p, ph, A.p, B.p,
                                                                acc.p, (prm.mdbc && !slab.active) ? ghost.p : nullptr,
                                                                slab.active ? id.p : nullptr, okey.p);
        ++launches;
        CK(cudaGetLastError());
        // size the dense cell grid from the bounding box of the uploaded positions (grown on demand later)
        {
            double *d_bb = (double *)(stage.p + (((4 * vb + sb + 63) / 64) + 1) * 64);   // behind the staged columns (the stage buffer has spare room)
            double h_bb[7] = {1e300, 1e300, 1e300, -1e300, -1e300, -1e300, 0.0};
            CK(cudaMemcpyAsync(d_bb, h_bb, sizeof h_bb, cudaMemcpyHostToDevice, stream));
            k_upload_bbox<T, D><<<grid_for(count), 256, 0, stream>>>((int)count, d_pos, d_bb);
            ++launches;
            CK(cudaMemcpyAsync(h_bb, d_bb, sizeof h_bb, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            long long cells = 1;
            for (int k = 0; k < D; ++k) {
                double e = (h_bb[3 + k] - h_bb[k]) * prm.H_inv + 12.0;
                if (h_bb[6] != 0.0 || !(e < 1e9)) return fail(SPHB200_EINVAL, "upload: non-finite or absurd position range");
                cells *= (long long)e;
                if (cells > (1ll << 40)) break;
            }
            rc = alloc_cells(2 * cells + 1024);
            if (rc) return rc;
        }
        // reset the loop state (a fresh SimParticles table)
        memset(h_ctl, 0, sizeof(Ctl));
        CK(cudaMemcpyAsync(d_ctl.p, h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        have_cells = false;
        have_half = false;
        uploaded = true;
        slab.own_p0 = slab.l1 = 0;
        slab.own_p1 = slab.l2 = (int)count;
        return SPHB200_OK;
    }

    int download(int order, void *pos, void *vel, void *accel, void *rho, void *press, int64_t *ids, uint8_t *ty,
                 uint64_t *grp, int64_t *cells) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "download before upload");
        CK(cudaSetDevice(device));
        // slab mode: only the owned range [own_p0, own_p1) of the table is this rank's to report
        const int64_t off = slab.active ? slab.own_p0 : 0;
        const int64_t cnt = slab.active ? slab.own_p1 - slab.own_p0 : n;
        unsigned char *sp = stage.p;
        size_t vb = (size_t)cnt * D * sizeof(T), sb = (size_t)cnt * sizeof(T);
        T *d_pos = (T *)sp; sp += vb;
        T *d_vel = (T *)sp; sp += vb;
        T *d_acc = (T *)sp; sp += vb;
        sp += vb;
        T *d_rho = (T *)sp; sp += sb;
        T *d_pr = (T *)sp; sp += sb;
        k_unpack_download<T, D><<<grid_for(cnt), 256, 0, stream>>>((int)cnt, A.p + off, B.p + off, acc.p + off, d_pos, d_vel, d_acc,
                                                                  d_rho, d_pr);
        ++launches;
        CK(cudaGetLastError());
        std::vector<long long> hid;
        std::vector<int64_t> order_idx;
        if (order == 1) {
            hid.resize((size_t)cnt);
            CK(cudaMemcpyAsync(hid.data(), id.p + off, (size_t)cnt * 8, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            order_idx.resize((size_t)cnt);
            std::iota(order_idx.begin(), order_idx.end(), 0);
            std::stable_sort(order_idx.begin(), order_idx.end(), [&](int64_t a, int64_t b) { return hid[a] < hid[b]; });
        }
        auto fetch = [&](void *dst, const void *dsrc, size_t elem, int comps) -> int {
            if (!dst) return 0;
            size_t bytes = (size_t)cnt * elem * comps;
            if (order != 1) {
                CK(cudaMemcpyAsync(dst, dsrc, bytes, cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                return 0;
            }
            std::vector<unsigned char> tmp(bytes);
            CK(cudaMemcpyAsync(tmp.data(), dsrc, bytes, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            size_t rb = elem * comps;
            for (int64_t k = 0; k < cnt; ++k) memcpy((unsigned char *)dst + k * rb, tmp.data() + order_idx[k] * rb, rb);
            return 0;
        };
        int rc;
        if ((rc = fetch(pos, d_pos, sizeof(T), D))) return rc;
        if ((rc = fetch(vel, d_vel, sizeof(T), D))) return rc;
        if ((rc = fetch(accel, d_acc, sizeof(T), D))) return rc;
        if ((rc = fetch(rho, d_rho, sizeof(T), 1))) return rc;
        if ((rc = fetch(press, d_pr, sizeof(T), 1))) return rc;
        if ((rc = fetch(ids, id.p + off, 8, 1))) return rc;
        if ((rc = fetch(ty, type.p + off, 1, 1))) return rc;
        if ((rc = fetch(grp, group.p + off, 8, 1))) return rc;
        if (cells) {
            // Cells field = CartesianIndex assigned at the last UpdateNeighbors! (stale in between)
            std::vector<int> hk((size_t)cnt);
            if (!have_cells) {
                memset(cells, 0, (size_t)cnt * D * 8);
            } else {
                CK(cudaMemcpyAsync(hk.data(), ckey.p + off, (size_t)cnt * 4, cudaMemcpyDeviceToHost, stream));
                CK(cudaMemcpyAsync(h_grid, d_grid.p, sizeof(GridInfo), cudaMemcpyDeviceToHost, stream));
                CK(cudaStreamSynchronize(stream));
                for (int64_t k = 0; k < cnt; ++k) {
                    int64_t src = order == 1 ? order_idx[k] : k;
                    int c[3];
                    key_to_cell(hk[src], c);
                    for (int d = 0; d < D; ++d) cells[k * D + d] = c[d];
                }
            }
        }
        return SPHB200_OK;
    }
    void key_to_cell(int key, int *c) const {
        int cx = key % h_grid->nx;
        int r = key / h_grid->nx;
        int cm = r % h_grid->nm, cs = r / h_grid->nm;
        c[am.ax_f] = cx + h_grid->cmin[am.ax_f];
        if (D == 3) c[am.ax_m] = cm + h_grid->cmin[am.ax_m];
        c[am.ax_s] = cs + h_grid->cmin[am.ax_s];
    }
    int64_t num_particles() const override { return slab.active ? slab.own_p1 - slab.own_p0 : n; }
    int64_t launch_count() const override { return launches; }

    // Wait for the stream.  In slab mode the stream carries NCCL operations whose completion depends
    // on the other ranks: instead of blocking forever behind a dead or diverged peer, poll with a
    // deadline (SPHB200_SLAB_TIMEOUT_S, default 600) and fail loudly.
    int wait_stream() {
        if (!slab.active) {
            CK(cudaStreamSynchronize(stream));
            return SPHB200_OK;
        }
        const double limit_s = (double)env_int("SPHB200_SLAB_TIMEOUT_S", 600);
        const auto t0 = std::chrono::steady_clock::now();
        for (long spins = 0;; ++spins) {
            cudaError_t e = cudaStreamQuery(stream);
            if (e == cudaSuccess) return SPHB200_OK;
            if (e != cudaErrorNotReady)
                return fail(SPHB200_ECUDA, "cudaStreamQuery failed: %s (%s:%d)", cudaGetErrorString(e), __FILE__, __LINE__);
            if (spins > 4000) {   // the common case completes within the busy spins
                const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                if (el > limit_s) slab.dead = true;
                if (el > limit_s)
                    return fail(SPHB200_ENCCL, "rank %d: no progress for %.0f s in a slab-mode wait (a peer rank died or the ranks diverged)",
                                slab.rank, el);
                std::this_thread::sleep_for(std::chrono::microseconds(20));
            }
        }
    }
    int sync_ctl() {
        CK(cudaMemcpyAsync(h_ctl, d_ctl.p, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        CK(cudaMemcpyAsync(h_grid, d_grid.p, sizeof(GridInfo), cudaMemcpyDeviceToHost, stream));
        return wait_stream();
    }
    int push_ctl() {
        CK(cudaMemcpyAsync(d_ctl.p, h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, stream));
        return SPHB200_OK;
    }
    int set_time(double t, int64_t it) override {
        CK(cudaSetDevice(device));
        int rc = sync_ctl();
        if (rc) return rc;
        h_ctl->total_time = t;
        h_ctl->iteration = it;
        rc = push_ctl();
        if (rc) return rc;
        CK(cudaStreamSynchronize(stream));
        return SPHB200_OK;
    }
    void fill_report(sphb200_report *rep) {
        if (!rep) return;
        rep->iteration = h_ctl->iteration;
        rep->index_counter = 0;
        rep->n_rebuilds = h_ctl->n_rebuilds;
        rep->n_particles = slab.active ? (int64_t)(slab.own_p1 - slab.own_p0) : n;
        rep->n_halo = slab.active ? n - rep->n_particles : 0;
        rep->total_time = h_ctl->total_time;
        rep->current_dt = h_ctl->current_dt;
        rep->delta_x = h_ctl->delta_x;
    }
    int get_report(sphb200_report *rep) override {
        CK(cudaSetDevice(device));
        int rc = sync_ctl();
        if (rc) return rc;
        fill_report(rep);
        if (rep && have_cells) {
            int64_t nc = 0;
            rc = count_occupied(&nc);
            if (rc) return rc;
            rep->index_counter = nc + 1;   // IndexCounter counts the dummy first entry too, :145-160
        }
        return SPHB200_OK;
    }

    // ------------------------------------------------------------------ kernel launch helpers
    Table<T, D> table(bool scratch) {
        Table<T, D> t;
        t.A = scratch ? A2.p : A.p;
        t.B = scratch ? B2.p : B.p;
        t.acc = scratch ? acc2.p : acc.p;
        t.ghost = (prm.mdbc && !slab.active) ? (scratch ? ghost2.p : ghost.p) : nullptr;   // slab mode: node table instead (set_ghost_nodes)
        t.id = scratch ? id2.p : id.p;
        t.group = scratch ? group2.p : group.p;
        t.okey = scratch ? okey2.p : okey.p;
        t.type = scratch ? type2.p : type.p;
        t.ckey = scratch ? ckey2.p : ckey.p;
        return t;
    }

    // UpdateNeighbors! — every kernel is predicated on ctl->do_rebuild
    int enqueue_rebuild(const SlabFilter &flt = SlabFilter{0, 0, 0, 0, 0, 0}, int count_rebuild = 1) {
        const int nn = (int)n;
        const int gp = grid_for(nn);
        const int gc = grid_for(cell_cap);
        const int nscan = (int)(cell_cap / SCAN_CHUNK + 1);
        k_cell_bbox<T, D><<<gp, 256, 0, stream>>>(A.p, nn, prm.H_inv, ccoord.p, d_ctl.p, d_grid.p, am, flt);
        k_grid_setup<D><<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p, am, cell_cap, row_cap, own_lo, own_hi);
        k_zero_counts<<<gc, 256, 0, stream>>>(d_ctl.p, d_grid.p, cell_count.p);
        k_cell_count<D><<<gp, 256, 0, stream>>>(ccoord.p, nn, am, d_ctl.p, d_grid.p, key_tmp.p, slot_tmp.p, cell_count.p);
        k_scan_partials<<<nscan, SCAN_THREADS, 0, stream>>>(d_ctl.p, d_grid.p, cell_count.p, scan_partial.p);
        k_scan_top<<<1, 1024, 0, stream>>>(d_ctl.p, d_grid.p, scan_partial.p);
        k_scan_final<<<nscan, SCAN_THREADS, 0, stream>>>(d_ctl.p, d_grid.p, cell_count.p, scan_partial.p, cell_start.p);
        k_scatter_unstable<<<gp, 256, 0, stream>>>(d_ctl.p, key_tmp.p, slot_tmp.p, nn, cell_start.p, tmp_idx.p);
        k_stable_rank<<<gp, 256, 0, stream>>>(d_ctl.p, d_grid.p, key_tmp.p, tmp_idx.p, nn, cell_start.p, okey.p, perm.p);
        k_gather_table<T, D><<<gp, 256, 0, stream>>>(d_ctl.p, d_grid.p, cell_start.p, perm.p, table(false), table(true), key_tmp.p,
                                                     ccoord.p);
        k_copy_table<T, D><<<gp, 256, 0, stream>>>(d_ctl.p, d_grid.p, cell_start.p, table(true), table(false));
        if (slab.active) {   // boundary-layer bricks first: a pass takes them first and ships their results early
            k_build_bricks<D><<<grid_for(row_cap, 128), 128, 0, stream>>>(d_ctl.p, d_grid.p, cell_start.p, brick_targets(), brick_window_limit(),
                                                                           bricks.p, brick_cap, 1);
            k_mark_boundary_bricks<<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p);
            k_build_bricks<D><<<grid_for(row_cap, 128), 128, 0, stream>>>(d_ctl.p, d_grid.p, cell_start.p, brick_targets(), brick_window_limit(),
                                                                           bricks.p, brick_cap, 2);
            launches += 2;
        } else {
            k_build_bricks<D><<<grid_for(row_cap, 128), 128, 0, stream>>>(d_ctl.p, d_grid.p, cell_start.p, brick_targets(), brick_window_limit(),
                                                                           bricks.p, brick_cap, 0);
        }
        k_finish_rebuild<<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p, cell_start.p, count_rebuild);
        launches += 13;
        CK(cudaGetLastError());
        return SPHB200_OK;
    }

    bool lists_on() const { return opt_lists && !prm.shifting && opt_skin > 0.0; }
    // candidates one ring slot of the list kernel holds (sentinels included)
    int list_cap_cand() const {
        int cap = generic ? RingGeom<T, D, true>::CAP : RingGeom<T, D, false>::CAP;
        if (opt_list_smem_kb > 0) {
            const int per = generic ? RingGeom<T, D, true>::S1::per_candidate : RingGeom<T, D, false>::S1::per_candidate;
            cap = std::min(cap, ((opt_list_smem_kb * 1024 - 64) / per) & ~7);
        }
        return std::max(cap, 16);   // (its last 8 records are the sentinels)
    }
    // candidates a brick's window may hold: what the list kernel can stage, else a bound that merely
    // keeps sparse rows from producing row-long windows
    int brick_window_limit() {
        if (!lists_on()) return 8192;
        const int lim = list_cap_cand() - 8 - 6 * ((D == 3) ? 9 : 3);   // RingGeom::WINDOW_LIMIT for the effective cap
        return lim >= 64 ? lim : 64;
    }
    // Particles per brick (<= BT).  A brick is served by ONE CTA of the list kernel, i.e. one SM: with few particles
    // the bricks must be small enough to reach every SM several times over — C1 (6 881 particles) is 54 bricks of 128,
    // so 94 of the 148 SMs idle while the others grind through a brick each on their fp64 pipes
    // (profiles/r4h_prof_ring_c1_ncu.txt: SMs active 39 % of the kernel).  Halved until there are two bricks per SM —
    // and no further: once every SM is busy, smaller bricks only add staging (C2, 60 k particles: 425 Mpu/s with
    // bricks of 128, 332 with 64; C1: 110 -> 148, C5: 48 -> 67 with 32; profiles/r4i_small_profile*.jsonl).
    int brick_targets() const {
        if (opt_brick_targets > 0) return std::min(opt_brick_targets, BT);
        int bt = BT;
        while (bt > 32 && n / bt < 2 * (int64_t)num_sms) bt >>= 1;
        return bt;
    }
    double motion_vmax() const {
        double v = 0.0;
        for (int k = 0; k < motions.n; ++k) {
            double d2 = 0.0;
            for (int c = 0; c < D; ++c) d2 += motions.dir[k][c] * motions.dir[k][c];
            v = std::max(v, fabs(motions.velocity[k]) * sqrt(d2));
        }
        return v;
    }
    int ensure_lists() {
        if (!lists_on()) return SPHB200_OK;
        const size_t need = (size_t)(opt_lcap / 8) * n_alloc;
        if (nl.n < need || nl_stride != n_alloc) {
            nl.release();
            CK(nl.alloc(need));
            CK(nl_cnt.alloc(n_alloc));
            nl_stride = n_alloc;
            k_invalidate_lists<<<1, 1, 0, stream>>>(d_ctl.p);
            ++launches;
        }
        return SPHB200_OK;
    }
    void fill_args(InteractArgs<T, D> &g, int pass, int epilogue) {
        memset(&g, 0, sizeof g);
        g.A = pass ? Ah.p : A.p;
        g.B = pass ? Bh.p : B.p;
        g.RN = RN.p;
        g.rn_out = (pass == 0 && epilogue == EPI_FUSED && snapshot_in_epilogue) ? RN.p : nullptr;
        g.Bn = B2.p;   // vₙ snapshot (LaminarSPS pass 2, Q2); B itself is rewritten by the fused corrector
        g.An_rw = A.p;
        g.Bn_rw = B.p;
        g.Ah_out = Ah.p;
        g.Bh_out = Bh.p;
        g.drhodt = drhodt.p;
        g.acc = acc.p;
        g.gradC = gradC.p;
        g.divr = divr.p;
        g.ksum = ksum.p;
        g.kgrad = kgrad.p;
        g.cell_start = cell_start.p;
        g.ckey = ckey.p;
        g.type = type.p;
        g.bricks = bricks.p;
        g.grid = d_grid.p;
        g.ctl = d_ctl.p;
        g.phys = ph;
        g.epilogue = epilogue;
        g.use_tma = opt_tma;
        g.am = am;
        g.counter_slot = pass * 3 + brick_part;
        g.brick_part = brick_part;
        g.bnd_flag = bnd_flag;
        g.bnd_epoch = bnd_epoch;
        g.nl = nl.p;
        g.nl_cnt = nl_cnt.p;
        g.nl_stride = nl_stride;
        g.lcap = opt_lcap;
        g.list_cap_cand = list_cap_cand();
        g.list_reorder = opt_list_reorder;
        g.brick_flag = brick_flag.p;
        g.brick_move = opt_list_local ? brick_move.p : nullptr;
        const double Hs = prm.H * (1.0 + opt_skin);
        g.Hs2 = (T)(Hs * Hs);
        g.force_cull = cull_force;
        g.lean_guard = lean_enqueue ? 1 : 0;
    }
    // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) + occupancy, once per kernel, smem size and HANDLE
    // (the attribute is per device; a handle is bound to one device)
    template <class K>
    int configure_kernel(K kern, int threads, int smem, int *ctas_per_sm) {
        auto it = kernel_cfg.find((const void *)kern);
        if (it == kernel_cfg.end() || it->second.first != smem) {
            int ctas = 0;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, threads, smem));
            if (ctas < 1) return fail(SPHB200_ECUDA, "kernel does not fit on an SM (%d threads, %d B shared memory)", threads, smem);
            kernel_cfg[(const void *)kern] = std::make_pair(smem, ctas);
            *ctas_per_sm = ctas;
        } else {
            *ctas_per_sm = it->second.second;
        }
        return SPHB200_OK;
    }
    int persistent_blocks(int ctas_per_sm) const {
        int blocks = num_sms * ctas_per_sm;
        int64_t maxb = (n + BT - 1) / BT + (int64_t)row_cap;
        if ((int64_t)blocks > maxb) blocks = (int)std::max<int64_t>(1, maxb);
        return blocks;
    }
    // the list kernel: one persistent CTA per SM (producer warp + consumer warps over a ring of windows)
    // fp64, few particles: 4 lanes per target (k_interact_ring, SPLIT) — below ~32 k particles there are fewer
    // 32-target sub-bricks than consumer warps on the GPU, and a pass is as long as one lane's walk of its list
    bool ring_split() const {
        if (sizeof(T) != 8 || generic) return false;
        return opt_split < 0 ? n <= 32768 : opt_split != 0;
    }
    template <int PASS, bool GEN>
    int launch_ring_t(int epilogue) {
        using RG = RingGeom<T, D, GEN>;
        auto kern = k_interact_ring<T, D, PASS, GEN>;
        if constexpr (std::is_same<T, double>::value && !GEN && D == 2) {
            if (ring_split()) kern = k_interact_ring<T, D, PASS, GEN, 4>;
        }
        int ctas = 0, rc;
        if ((rc = configure_kernel(kern, RG::THREADS, RG::SMEM, &ctas))) return rc;
        InteractArgs<T, D> g;
        fill_args(g, PASS, epilogue);
        kern<<<persistent_blocks(1), RG::THREADS, RG::SMEM, stream>>>(g);
        ++launches;
        CK(cudaGetLastError());
        return SPHB200_OK;
    }

    // the physics-free list build + the bank-aware reorder (both run only when k_step_control raised ctl->list_build)
    template <bool GEN>
    int launch_list_build() {
        if (opt_list_local) {   // which bricks need new lists (everything else in this function skips the others)
            const size_t bs = (size_t)(cell_cap + 8) * 6;
            k_cell_vbox<T, D><<<grid_for(cell_cap), 256, 0, stream>>>(B.p, cell_start.p, d_grid.p, d_ctl.p, vbox.p, bs);
            k_brick_bounds<D><<<num_sms * 8, 256, 0, stream>>>(d_ctl.p, d_grid.p, bricks.p, ckey.p, vbox.p, bs, brick_move.p,
                                                                           brick_flag.p, opt_skin * prm.H, (double)opt_list_lookahead);
            launches += 2;
            CK(cudaGetLastError());
        }
        auto kern = k_list_build<T, D, GEN, BT>;
        const int list_bytes = LIST_CAP * BT * 2;   // append buffer
        int smem = std::min(opt_build_smem_kb, 200) * 1024;
        int cap = std::min(((smem - 64) / (int)sizeof(TA)) & ~3, 32764);
        smem = cap * (int)sizeof(TA) + list_bytes;
        int ctas = 0, rc;
        if ((rc = configure_kernel(kern, BT, smem, &ctas))) return rc;
        InteractArgs<T, D> g;
        fill_args(g, 0, EPI_FUSED);
        g.cap = cap;
        kern<<<persistent_blocks(ctas), BT, smem, stream>>>(g);
        ++launches;
        CK(cudaGetLastError());
        if (opt_list_reorder) {
            auto rk = k_list_reorder<BT>;
            const int rsmem = (opt_reorder_slots + REORDER_OVF_CAP) * BT * 2;
            if ((rc = configure_kernel(rk, BT, rsmem, &ctas))) return rc;
            rk<<<persistent_blocks(ctas), BT, rsmem, stream>>>(d_ctl.p, d_grid.p, bricks.p, brick_flag.p, (unsigned)(list_cap_cand() - 8), nl.p,
                                                             nl_cnt.p, nl_stride, opt_lcap, opt_reorder_slots);
            ++launches;
            CK(cudaGetLastError());
        }
        return SPHB200_OK;
    }

    template <int PASS, bool GEN, bool COMPACT>
    int launch_interact_t(int epilogue) {
        using SS = StageSizes<T, D, PASS, GEN>;
        auto kern = k_interact<T, D, PASS, GEN, COMPACT, BT>;
        const int list_bytes = COMPACT ? LIST_CAP * BT * 2 : 0;
        int smem = std::min(opt_smem_kb, 200) * 1024;
        int cap = ((smem - list_bytes - 64) / SS::per_candidate) & ~3;
        cap = std::min(cap, 32764);
        if (cap < 64) return fail(SPHB200_EINVAL, "shared-memory budget too small");
        smem = cap * SS::per_candidate + list_bytes;
        int ctas = 0, rc;
        if ((rc = configure_kernel(kern, BT, smem, &ctas))) return rc;
        InteractArgs<T, D> g;
        fill_args(g, PASS, epilogue);
        g.cap = cap;
        kern<<<persistent_blocks(ctas), BT, smem, stream>>>(g);
        ++launches;
        CK(cudaGetLastError());
        return SPHB200_OK;
    }
    // the (predicated) list build + reorder of a step; runs before pass 1
    int enqueue_list_build() {
        if (!lists_on()) return SPHB200_OK;
        int rc = ensure_lists();
        if (rc) return rc;
        return generic ? launch_list_build<true>() : launch_list_build<false>();
    }
    int launch_interact(int pass, int epilogue) {
        if (lists_on() && epilogue == EPI_FUSED) {
            // per pass two launches, exactly one of which does the work (ctl->list_mode[pass]): the
            // cull kernel or the list kernel
            int rc = ensure_lists();
            if (rc) return rc;
            if (opt_verify_lists && brick_part != 2) {
                InteractArgs<T, D> g;
                fill_args(g, pass, epilogue);
                if (generic) k_list_verify<T, D, true, BT><<<num_sms * 2, BT, 0, stream>>>(g, pass);
                else k_list_verify<T, D, false, BT><<<num_sms * 2, BT, 0, stream>>>(g, pass);
                ++launches;
            }
            if (!lean_enqueue && (rc = launch_cull(pass, epilogue, 0))) return rc;   // (lean: see enqueue_lean_sequence)
            if (generic) return pass ? launch_ring_t<1, true>(epilogue) : launch_ring_t<0, true>(epilogue);
            return pass ? launch_ring_t<1, false>(epilogue) : launch_ring_t<0, false>(epilogue);
        }
        return launch_cull(pass, epilogue, 1);
    }
    int launch_cull(int pass, int epilogue, int force_cull) {
        cull_force = force_cull;
        if (generic) return pass ? launch_interact_t<1, true, false>(epilogue) : launch_interact_t<0, true, false>(epilogue);
        if (opt_compact) return pass ? launch_interact_t<1, false, true>(epilogue) : launch_interact_t<0, false, true>(epilogue);
        return pass ? launch_interact_t<1, false, false>(epilogue) : launch_interact_t<0, false, false>(epilogue);
    }

    // state-n snapshots the pass-2 pair terms read (Q2): ρₙ always, vₙ for LaminarSPS
    // fused: the pass-1 epilogue writes ρₙ of every particle it owns, so the sweep is only needed for
    // the stage-level entry points and for the halo copies of slab mode
    int enqueue_snapshots(bool fused_step = false) {
        snapshot_in_epilogue = fused_step && !slab.active;
        if (!snapshot_in_epilogue) {
            k_snapshot_rho<T, D><<<grid_for(n), 256, 0, stream>>>(A.p, RN.p, (int)n, d_ctl.p);
            ++launches;
            CK(cudaGetLastError());
        }
        if (prm.viscosity == SPHB200_VISC_LAMINAR_SPS)
            CK(cudaMemcpyAsync(B2.p, B.p, (size_t)n * sizeof(TB), cudaMemcpyDeviceToDevice, stream));
        return SPHB200_OK;
    }
    int enqueue_mdbc() {
        if (slab.active) return slab_enqueue_mdbc();
        const int nn = (int)n;
        k_mdbc_gather<T, D><<<grid_for((int64_t)nn * 32, 128), 128, 0, stream>>>(A.p, ghost.p, type.p, cell_start.p, d_grid.p, am, nn, ph,
                                                                   prm.H_inv, rho_new.p, has_new.p, d_ctl.p);
        k_mdbc_apply<T, D><<<grid_for(nn), 256, 0, stream>>>(A.p, RN.p, type.p, rho_new.p, has_new.p, nn, d_ctl.p);
        launches += 2;
        CK(cudaGetLastError());
        return SPHB200_OK;
    }
    int enqueue_motion(double dt2) {
        if (motions.n == 0) return SPHB200_OK;
        k_progress_motion<T, D><<<grid_for(n), 256, 0, stream>>>(A.p, B.p, type.p, group.p, (int)n, motions, dt2, d_ctl.p);
        ++launches;
        CK(cudaGetLastError());
        return SPHB200_OK;
    }

    // phase A of one step: S0, S1 and the S2 decision
    int enqueue_step_head() {
        const int p0 = slab.active ? slab.own_p0 : 0, p1 = slab.active ? slab.own_p1 : (int)n;
        k_reduce_dt_dx<T, D><<<grid_for(p1 - p0), 256, 0, stream>>>(A.p, B.p, Ah.p, acc.p, p0, p1, ph.h, ph.eta2, have_half ? 1 : 0,
                                                                    d_ctl.p);
        int rc;
        if (slab.active && (rc = slab_allreduce_ctl())) return rc;
        k_step_control<T><<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p, ph.h, ph.c0, (T)prm.cfl,
                                               lists_on() ? opt_skin * prm.H : 0.0, motion_vmax(), slab.active ? 1 : 0, opt_list_local,
                                               capturing_cond ? cond_rebuild : 0ull, capturing_cond ? cond_lists : 0ull);
        launches += 2;
        CK(cudaGetLastError());
        return SPHB200_OK;
    }
    // phase B: S2 .. S19, in the pieces a conditional step graph is assembled from.
    // ev (optional, 10 events): stage boundaries for stage_times()
    int enqueue_body_rebuild() { return enqueue_rebuild(); }                  // S2  "02 Calculate IndexCounter"
    int enqueue_body_pre(cudaEvent_t *ev = nullptr) {
        int rc;
        if ((rc = enqueue_motion(-1.0))) return rc;                           // S3  "Motion"
        if ((rc = enqueue_snapshots(true))) return rc;
        if (ev) CK(cudaEventRecord(ev[2], stream));
        if (prm.mdbc && (rc = enqueue_mdbc())) return rc;                     // S6  "04 Apply MDBC before Half TimeStep"
        return SPHB200_OK;
    }
    int enqueue_body_passes(cudaEvent_t *ev = nullptr, bool with_end = true) {
        int rc;
        if ((rc = launch_interact(0, EPI_FUSED))) return rc;                  // S4-S10, S13  "05", "03", "06", "07"
        if (ev) CK(cudaEventRecord(ev[5], stream));
        if ((rc = enqueue_motion(-1.0))) return rc;                           // S12 "Motion"
        if (ev) CK(cudaEventRecord(ev[6], stream));
        if ((rc = launch_interact(1, EPI_FUSED))) return rc;                  // S11, S14-S18  "08", "03", "09", "10", "11"
        if (ev) CK(cudaEventRecord(ev[7], stream));
        if (with_end) {
            k_step_end<<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p);               // S19 "12 Update MetaData"
            ++launches;
        }
        CK(cudaGetLastError());
        have_half = true;
        have_cells = true;
        return SPHB200_OK;
    }
    int enqueue_step_body(cudaEvent_t *ev = nullptr) {
        int rc;
#define EV(k) if (ev) CK(cudaEventRecord(ev[k], stream))
        EV(0);
        if ((rc = enqueue_body_rebuild())) return rc;
        EV(1);
        if ((rc = enqueue_body_pre(ev))) return rc;
        EV(3);
        if ((rc = enqueue_list_build())) return rc;                           //     neighbour-list maintenance (no reference stage)
        EV(4);
        if ((rc = enqueue_body_passes(ev))) return rc;
        EV(8);
#undef EV
        return SPHB200_OK;
    }

    // grow the dense cell grid after an ECAPACITY stop; returns OK if the step can be resumed
    int recover_capacity() {
        long long need = 1;
        for (int k = 0; k < D; ++k) need *= ((long long)h_grid->bb_max[k] - h_grid->bb_min[k] + 3);
        if (need <= 0 || need > (1ll << 28)) return fail(SPHB200_ECAPACITY, "particles left the tractable domain (dense cell grid would need %lld cells)", need);
        int rc = alloc_cells(2 * need + 1024);
        if (rc) return rc;
        h_ctl->error = 0;
        return push_ctl();
    }

    // ---- one step = one CUDA graph launch ---------------------------------------------------------
    // Every decision of a step lives on the device (Ctl) and every kernel is predicated on it, so the
    // launch sequence of a step is static: it is captured once (≈ 25 kernels) and replayed.  That takes
    // the host's per-launch cost and the inter-kernel gaps out of small cases (2D, 10^4-10^5 particles),
    // where a step is shorter than its launch overhead.  Re-captured when anything a kernel argument
    // depends on changes (upload, reallocation, options, stream).
    cudaGraph_t step_graph = nullptr;
    cudaGraphExec_t step_exec = nullptr;
    int64_t step_graph_launches = 0;
    int opt_graph = 1;
    int opt_graph_cond = 0;   // UpdateNeighbors! and the list maintenance behind CUDA-graph conditional (IF) nodes
    cudaGraphConditionalHandle cond_rebuild = 0, cond_lists = 0;   // non-zero only while a conditional step graph is captured / alive
    bool capturing_cond = false;   // the handles are kernel arguments ONLY inside that graph (plain launches must not touch them)
    // the lean step (no reductions sweep, no UpdateNeighbors! chain; k_step_control pauses a step that needs the chain)
    cudaGraph_t lean_graph = nullptr;
    cudaGraphExec_t lean_exec = nullptr;
    int64_t lean_graph_launches = 0;
    int opt_lean = 1;
    int opt_brick_targets = env_int("SPHB200_BRICK_TARGETS", 0);
    int opt_split = env_int("SPHB200_SPLIT", -1);   // list kernel, fp64 2D: 4 lanes per target (-1: by particle count)
    int64_t n_lean_steps = 0, n_lean_pauses = 0, opt_test_fail_at = 0;
    void drop_step_graph() {
        if (step_exec) cudaGraphExecDestroy(step_exec);
        if (step_graph) cudaGraphDestroy(step_graph);
        step_exec = nullptr;
        step_graph = nullptr;
        if (lean_exec) cudaGraphExecDestroy(lean_exec);
        if (lean_graph) cudaGraphDestroy(lean_graph);
        lean_exec = nullptr;
        lean_graph = nullptr;
        cond_rebuild = cond_lists = 0;
    }
    // The step as ONE graph whose rarely needed parts sit behind conditional nodes:
    //   head (reductions, control: sets the two conditions) -> IF(rebuild){UpdateNeighbors! chain}
    //   -> motion, snapshots, mDBC -> IF(lists){velocity boxes, brick bounds, list build, reorder} -> passes, step end.
    // A step that neither rebuilds cells nor lists is ~10 kernel nodes instead of ~30, which is what a
    // 10^3-10^5 particle case (shorter than its own launch overhead) is made of.  Returns false (and
    // leaves no graph behind) if the driver refuses any part: the caller then captures the flat graph.
    bool capture_conditional_step_graph() {
        const cudaStreamCaptureMode mode = cudaStreamCaptureModeThreadLocal;
        cudaGraph_t g = nullptr;
        bool ok = cudaGraphCreate(&g, 0) == cudaSuccess;
        ok = ok && cudaGraphConditionalHandleCreate(&cond_rebuild, g, 0, cudaGraphCondAssignDefault) == cudaSuccess;
        ok = ok && cudaGraphConditionalHandleCreate(&cond_lists, g, 0, cudaGraphCondAssignDefault) == cudaSuccess;
        std::vector<cudaGraphNode_t> tail;
        // one captured segment appended to graph `dst` behind `tail`; the new tail is returned in `tail`
        auto segment = [&](cudaGraph_t dst, bool track_tail, auto &&body) -> bool {
            if (cudaStreamBeginCaptureToGraph(stream, dst, tail.empty() || dst != g ? nullptr : tail.data(), nullptr,
                                              dst != g ? 0 : tail.size(), mode) != cudaSuccess)
                return false;
            const int rc = body();
            bool good = rc == 0;
            if (good && track_tail) {
                cudaStreamCaptureStatus st;
                const cudaGraphNode_t *deps = nullptr;
                size_t ndeps = 0;
                good = cudaStreamGetCaptureInfo(stream, &st, nullptr, nullptr, &deps, &ndeps) == cudaSuccess;
                if (good) tail.assign(deps, deps + ndeps);
            }
            cudaGraph_t out = nullptr;
            good = (cudaStreamEndCapture(stream, &out) == cudaSuccess) && good;
            return good;
        };
        auto conditional = [&](cudaGraphConditionalHandle h, auto &&body) -> bool {
            cudaGraphNodeParams p = {};
            p.type = cudaGraphNodeTypeConditional;
            p.conditional.handle = h;
            p.conditional.type = cudaGraphCondTypeIf;
            p.conditional.size = 1;
            cudaGraphNode_t node = nullptr;
            if (cudaGraphAddNode(&node, g, tail.data(), tail.size(), &p) != cudaSuccess) return false;
            cudaGraph_t inner = p.conditional.phGraph_out[0];
            std::vector<cudaGraphNode_t> saved;
            saved.swap(tail);
            const bool good = segment(inner, false, body);
            tail.assign(1, node);
            return good;
        };
        capturing_cond = true;
        ok = ok && segment(g, true, [&] { return enqueue_step_head(); });
        capturing_cond = false;
        ok = ok && conditional(cond_rebuild, [&] { return enqueue_body_rebuild(); });
        ok = ok && segment(g, true, [&] { return enqueue_body_pre(); });
        if (lists_on()) ok = ok && conditional(cond_lists, [&] { return enqueue_list_build(); });
        ok = ok && segment(g, true, [&] { return enqueue_body_passes(); });
        ok = ok && cudaGraphInstantiate(&step_exec, g, 0) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
                cudaGraph_t junk = nullptr;
                cudaStreamEndCapture(stream, &junk);
            }
            cudaGetLastError();
            if (step_exec) cudaGraphExecDestroy(step_exec);
            step_exec = nullptr;
            if (g) cudaGraphDestroy(g);
            cond_rebuild = cond_lists = 0;
            return false;
        }
        step_graph = g;
        return true;
    }
    int enqueue_step() {
        int rc;
        if ((rc = ensure_lists())) return rc;   // any (re)allocation happens outside a stream capture
        // (the legacy default stream cannot be captured: callers that hand it over via set_stream get plain launches)
        if (!opt_graph || !have_half || !have_cells || stream == nullptr) {   // first steps: arguments still change, attributes get set
            if ((rc = enqueue_step_head())) return rc;
            return enqueue_step_body();
        }
        if (!step_exec && opt_graph_cond) {
            const int64_t l0 = launches;
            if (capture_conditional_step_graph()) {
                step_graph_launches = launches - l0;
            } else {
                opt_graph_cond = 0;   // flat capture below
            }
            launches = l0;
        }
        if (!step_exec) {
            const int64_t l0 = launches;
            if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                opt_graph = 0;   // not capturable here: plain launches from now on
                if ((rc = enqueue_step_head())) return rc;
                return enqueue_step_body();
            }
            rc = enqueue_step_head();
            if (!rc) rc = enqueue_step_body();
            cudaError_t e = cudaStreamEndCapture(stream, &step_graph);
            if (rc) {
                drop_step_graph();
                return rc;
            }
            if (e != cudaSuccess) {
                drop_step_graph();
                return fail(SPHB200_ECUDA, "step graph capture failed: %s", cudaGetErrorString(e));
            }
            CK(cudaGraphInstantiate(&step_exec, step_graph, 0));
            step_graph_launches = launches - l0;
            launches = l0;
        }
        CK(cudaGraphLaunch(step_exec, stream));
        launches += step_graph_launches;
        return SPHB200_OK;
    }

    // One lean step: S0/S1 come from the previous pass 2 (ctl->red_ready, checked by the caller); no
    // UpdateNeighbors! chain and, with lists, no cull kernels standing by — the step control pauses a step that
    // turns out to need either (ctl->paused = 1), the list kernel one whose list build overflowed (paused = 2).
    // S19 of a lean step rides in the head kernel of the next one; the caller closes the batch with k_step_end.
    bool lean_enqueue = false;
    int enqueue_lean_sequence() {
        int rc;
        k_step_end_control<T><<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p, ph.h, ph.c0, (T)prm.cfl, lists_on() ? opt_skin * prm.H : 0.0,
                                                   motion_vmax(), lists_on() ? 3 : 1, opt_list_local);
        ++launches;
        CK(cudaGetLastError());
        if ((rc = enqueue_body_pre())) return rc;
        if ((rc = enqueue_list_build())) return rc;
        if (opt_test_fail_at > 0 && n_lean_steps == opt_test_fail_at && lists_on()) {   // test hook (plain launches only)
            k_test_fail_list_build<<<1, 1, 0, stream>>>(d_ctl.p);
            ++launches;
        }
        lean_enqueue = lists_on();
        rc = enqueue_body_passes(nullptr, false);
        lean_enqueue = false;
        return rc;
    }
    int enqueue_step_lean() {
        int rc;
        if ((rc = ensure_lists())) return rc;
        ++n_lean_steps;
        if (!opt_graph || stream == nullptr) return enqueue_lean_sequence();
        if (!lean_exec) {
            const int64_t l0 = launches;
            if (cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
                cudaGetLastError();
                opt_graph = 0;
                return enqueue_lean_sequence();
            }
            rc = enqueue_lean_sequence();
            cudaError_t e = cudaStreamEndCapture(stream, &lean_graph);
            if (rc || e != cudaSuccess) {
                drop_step_graph();
                return rc ? rc : fail(SPHB200_ECUDA, "lean step graph capture failed: %s", cudaGetErrorString(e));
            }
            CK(cudaGraphInstantiate(&lean_exec, lean_graph, 0));
            lean_graph_launches = launches - l0;
            launches = l0;
        }
        CK(cudaGraphLaunch(lean_exec, stream));
        launches += lean_graph_launches;
        return SPHB200_OK;
    }
    // how many of the next steps can go without UpdateNeighbors!, from the host's copy of the control block:
    // delta_x grows by about last_disp4 per step and triggers at h (step_control); 0 = the next step rebuilds
    int64_t lean_steps_ahead() const {
        if (!opt_lean || !have_cells || !have_half) return 0;
        return sph::lean_steps_ahead(*h_ctl, (double)ph.h, lists_on() ? opt_skin * prm.H : 0.0, opt_batch);
    }

    int run_steps(int64_t nsteps, bool until_target) {
        if (!uploaded) return fail(SPHB200_ESTATE, "step before upload");
        CK(cudaSetDevice(device));
        if (slab.active) return run_steps_slab(nsteps, until_target);
        int64_t done_steps = 0;
        int rc = sync_ctl();
        if (rc) return rc;
        const int64_t it0 = h_ctl->iteration;
        while (until_target || done_steps < nsteps) {
            int64_t batch = opt_batch;
            if (!until_target) batch = std::min<int64_t>(batch, nsteps - done_steps);
            else if (h_ctl->current_dt > 0.0) {
                double rem = (h_ctl->target_time - h_ctl->total_time) / h_ctl->current_dt;
                batch = std::max<int64_t>(1, std::min<int64_t>(batch, (int64_t)(rem * 1.02) + 2));
            } else {
                batch = 1;
            }
            if (opt_lean) {
                const int64_t ahead = lean_steps_ahead();
                if (ahead >= 1) {
                    batch = std::min(batch, ahead);
                    for (int64_t s = 0; s < batch; ++s)
                        if ((rc = enqueue_step_lean())) return rc;
                    k_step_end<<<1, 1, 0, stream>>>(d_ctl.p, d_grid.p);   // S19 of the batch's last step
                    ++launches;
                } else {
                    if ((rc = enqueue_step())) return rc;      // the full sequence (it may rebuild), one step, then look again
                }
                if ((rc = sync_ctl())) return rc;
                if (h_ctl->paused && !h_ctl->error) {
                    // a lean step found that it has to rebuild after all: its control part is done, the rest of the
                    // batch ran empty behind it; finish it with the full body (UpdateNeighbors! included)
                    ++n_lean_pauses;
                    const int where = h_ctl->paused;
                    h_ctl->paused = 0;
                    h_ctl->done = 0;
                    if ((rc = push_ctl())) return rc;
                    if (where == 2) rc = enqueue_body_passes();   // motion, mDBC and the (failed) list build have run
                    else rc = enqueue_step_body();
                    if (rc) return rc;
                    if ((rc = sync_ctl())) return rc;
                }
            } else {
                for (int64_t s = 0; s < batch; ++s)
                    if ((rc = enqueue_step())) return rc;
                if ((rc = sync_ctl())) return rc;
            }
            for (int attempt = 0; h_ctl->error == SPHB200_ECAPACITY; ++attempt) {
                if (attempt >= 3) return fail(SPHB200_ECAPACITY, "cell grid / brick list capacity exceeded");
                if ((rc = recover_capacity())) return rc;
                if (h_ctl->step_open && (rc = enqueue_step_body())) return rc;
                if ((rc = sync_ctl())) return rc;
            }
            if (h_ctl->error == SPHB200_ENUMERIC)
                return fail(SPHB200_ENUMERIC, "non-finite state or time step at iteration %lld (t = %g)", h_ctl->iteration, h_ctl->total_time);
            if (h_ctl->error) return fail(h_ctl->error, "device reported error %d", h_ctl->error);
            done_steps = h_ctl->iteration - it0;
            if (until_target && h_ctl->done) break;
        }
        return SPHB200_OK;
    }
    int run_steps_slab(int64_t nsteps, bool until_target);
    int slab_exchange_counts(int to_left, int to_right, int *from_left, int *from_right);
    int slab_exchange_records(Table<T, D> from, int sl0, int nl, int sr0, int nr, int dst0, int rl, int rr);
    int slab_exchange_halo(TA *a, TB *b, cudaStream_t st);
    int slab_pass(int pass, TA *xa, TB *xb, cudaEvent_t *xev = nullptr);
    int slab_allreduce_ctl();
    int slab_enqueue_mdbc();
    int slab_sort(const SlabFilter &flt, int count_rebuild);
    int slab_rebuild();
    int slab_step_body(cudaEvent_t *ev, bool host_synced = true, cudaEvent_t *xev = nullptr);
    int slab_resume_after_pause();
    int slab_check_head(bool *stop, bool until_target);
    int slab_stage_times(double *ms_out, int cnt);

    int step(int64_t nsteps, int reset_dx, sphb200_report *rep) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "step before upload");
        CK(cudaSetDevice(device));
        if (nsteps < 0) return fail(SPHB200_EINVAL, "negative step count");
        if (reset_dx || !have_cells) {
            int rc = sync_ctl();
            if (rc) return rc;
            h_ctl->delta_x = (double)(T(1) + ph.h);   // src/SPHCellList.jl:739
            h_ctl->use_target = 0;
            h_ctl->done = 0;
            if ((rc = push_ctl())) return rc;
        }
        int rc = run_steps(nsteps, false);
        if (rc) return rc;
        fill_report(rep);
        return SPHB200_OK;
    }
    int simulation_loop(double t_next, sphb200_report *rep) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "simulation_loop before upload");
        CK(cudaSetDevice(device));
        int rc = sync_ctl();
        if (rc) return rc;
        h_ctl->delta_x = (double)(T(1) + ph.h);
        h_ctl->use_target = 1;
        h_ctl->target_time = t_next;
        h_ctl->done = 0;
        if ((rc = push_ctl())) return rc;
        rc = run_steps(0, true);
        h_ctl->use_target = 0;
        h_ctl->done = 0;
        int rc2 = push_ctl();
        cudaStreamSynchronize(stream);
        if (rc) return rc;
        if (rc2) return rc2;
        fill_report(rep);
        return SPHB200_OK;
    }

    // ------------------------------------------------------------------ stage-level entry points
    int force_flag_rebuild() {
        int rc = sync_ctl();
        if (rc) return rc;
        h_ctl->do_rebuild = 1;
        h_ctl->done = 0;
        if ((rc = push_ctl())) return rc;
        for (int k = 0; k < 3; ++k) {
            h_grid->bb_min[k] = INT_MAX;
            h_grid->bb_max[k] = INT_MIN;
        }
        CK(cudaMemcpyAsync(d_grid.p, h_grid, sizeof(GridInfo), cudaMemcpyHostToDevice, stream));
        return SPHB200_OK;
    }
    int update_neighbors(int64_t *index_counter) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "update_neighbors before upload");
        CK(cudaSetDevice(device));
        if (slab.active) return fail(SPHB200_ESTATE, "stage-level calls are single-GPU only");
        invalidate_lists();
        int rc;
        for (int attempt = 0; attempt < 3; ++attempt) {
            if ((rc = force_flag_rebuild())) return rc;
            if ((rc = enqueue_rebuild())) return rc;
            if ((rc = sync_ctl())) return rc;
            if (h_ctl->error != SPHB200_ECAPACITY) break;
            if ((rc = recover_capacity())) return rc;
        }
        if (h_ctl->error) return fail(h_ctl->error, "update_neighbors: device reported error %d", h_ctl->error);
        have_cells = true;
        if (index_counter) {
            int64_t nc = 0;
            if ((rc = count_occupied(&nc))) return rc;
            *index_counter = nc + 1;
        }
        return SPHB200_OK;
    }
    int fetch_cell_start(std::vector<int> &cs) {
        cs.resize((size_t)h_grid->ncell + 1);
        CK(cudaMemcpyAsync(cs.data(), cell_start.p, cs.size() * 4, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return SPHB200_OK;
    }
    int count_occupied(int64_t *out) {
        std::vector<int> cs;
        int rc = sync_ctl();
        if (rc) return rc;
        if ((rc = fetch_cell_start(cs))) return rc;
        int64_t nc = 0;
        for (int c = 0; c < h_grid->ncell; ++c) nc += cs[c + 1] > cs[c];
        *out = nc;
        return SPHB200_OK;
    }
    int get_cell_list(int64_t *n_cells, int64_t *cells, int64_t *start) override {
        if (!have_cells) return fail(SPHB200_ESTATE, "get_cell_list before update_neighbors");
        CK(cudaSetDevice(device));
        std::vector<int> cs;
        int rc = sync_ctl();
        if (rc) return rc;
        if ((rc = fetch_cell_start(cs))) return rc;
        // occupied cells in the REFERENCE's order (last dimension most significant)
        struct Occ { int c[3]; int s, e; };
        std::vector<Occ> occ;
        for (int key = 0; key < h_grid->ncell; ++key)
            if (cs[key + 1] > cs[key]) {
                Occ o;
                o.c[0] = o.c[1] = o.c[2] = 0;
                key_to_cell(key, o.c);
                o.s = cs[key];
                o.e = cs[key + 1];
                occ.push_back(o);
            }
        if (!(am.ax_f == 0 && am.ax_s == D - 1))   // the device key order is not the reference's
            std::stable_sort(occ.begin(), occ.end(), [](const Occ &a, const Occ &b) {
                for (int k = D - 1; k >= 0; --k)
                    if (a.c[k] != b.c[k]) return a.c[k] < b.c[k];
                return false;
            });
        if (n_cells) *n_cells = (int64_t)occ.size();
        if (cells)
            for (size_t k = 0; k < occ.size(); ++k)
                for (int d = 0; d < D; ++d) cells[k * D + d] = occ[k].c[d];
        if (start) {
            for (size_t k = 0; k < occ.size(); ++k) start[k] = occ[k].s;
            start[occ.size()] = occ.empty() ? 0 : occ.back().e;
        }
        return SPHB200_OK;
    }
    int pressure(int half) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "pressure before upload");
        CK(cudaSetDevice(device));
        k_pressure<T, D><<<grid_for(n), 256, 0, stream>>>(half ? Ah.p : A.p, half ? Bh.p : B.p, (int)n, ph, d_ctl.p);
        ++launches;
        CK(cudaGetLastError());
        return SPHB200_OK;
    }
    int neighbor_loop(int pass, void *drhodt_out, void *acc_out) override {
        if (!have_cells) return fail(SPHB200_ESTATE, "neighbor_loop before update_neighbors");
        if (pass && !have_half) return fail(SPHB200_ESTATE, "neighbor_loop(pass 1) before half_time_step");
        CK(cudaSetDevice(device));
        k_reset_counters<<<1, 1, 0, stream>>>(d_ctl.p);
        ++launches;
        int rc;
        if (!pass) {
            if ((rc = enqueue_snapshots())) return rc;
        }
        if ((rc = launch_interact(pass ? 1 : 0, EPI_STORE))) return rc;
        if (drhodt_out) CK(cudaMemcpyAsync(drhodt_out, drhodt.p, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, stream));
        if (acc_out) {
            T *d_tmp = (T *)stage.p;
            k_unpack_vec<T, D><<<grid_for(n), 256, 0, stream>>>((int)n, acc.p, d_tmp);
            ++launches;
            CK(cudaMemcpyAsync(acc_out, d_tmp, (size_t)n * D * sizeof(T), cudaMemcpyDeviceToHost, stream));
        }
        CK(cudaStreamSynchronize(stream));
        CK(cudaGetLastError());
        return SPHB200_OK;
    }
    int delta_t(double *dt) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "delta_t before upload");
        CK(cudaSetDevice(device));
        int rc = sync_ctl();
        if (rc) return rc;
        h_ctl->red_disp2 = h_ctl->red_visc = h_ctl->red_acc2 = h_ctl->red_vel2 = 0ull;
        h_ctl->red_ready = 0;   // whatever a fused pass 2 left behind is recomputed by the next step head
        if ((rc = push_ctl())) return rc;
        k_reduce_dt_dx<T, D><<<grid_for(n), 256, 0, stream>>>(A.p, B.p, Ah.p, acc.p, 0, (int)n, ph.h, ph.eta2, 0, d_ctl.p);
        ++launches;
        if ((rc = sync_ctl())) return rc;
        T visc = (T)bits_to_double_host(h_ctl->red_visc);
        T acc2 = (T)bits_to_double_host(h_ctl->red_acc2);
        T dt1 = sph_sqrt(ph.h / sph_sqrt(acc2));
        T dt2 = ph.h / (ph.c0 + visc);
        if (dt) *dt = (double)((T)prm.cfl * std::min(dt1, dt2));
        h_ctl->red_disp2 = h_ctl->red_visc = h_ctl->red_acc2 = h_ctl->red_vel2 = 0ull;
        if ((rc = push_ctl())) return rc;
        CK(cudaStreamSynchronize(stream));
        return SPHB200_OK;
    }
    static double bits_to_double_host(unsigned long long b) {
        double d;
        memcpy(&d, &b, 8);
        return d;
    }
    void invalidate_lists() {
        k_invalidate_lists<<<1, 1, 0, stream>>>(d_ctl.p);
        ++launches;
    }
    int progress_motion(double dt2) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "progress_motion before upload");
        CK(cudaSetDevice(device));
        invalidate_lists();
        return enqueue_motion(dt2);
    }
    int apply_mdbc() override {
        if (!prm.mdbc) return SPHB200_OK;   // NoMDBC: no-op method, src/SPHCellList.jl:486-489
        if (!have_cells) return fail(SPHB200_ESTATE, "apply_mdbc before update_neighbors");
        CK(cudaSetDevice(device));
        int rc;
        if ((rc = enqueue_snapshots())) return rc;
        return enqueue_mdbc();
    }
    int half_time_step(double dt2) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "half_time_step before upload");
        CK(cudaSetDevice(device));
        invalidate_lists();
        k_half_step<T, D><<<grid_for(n), 256, 0, stream>>>(A.p, B.p, acc.p, drhodt.p, type.p, Ah.p, Bh.p, 0, (int)n, ph, dt2, d_ctl.p);
        ++launches;
        CK(cudaGetLastError());
        have_half = true;
        return SPHB200_OK;
    }
    int full_time_step(double dt) override {
        if (!have_half) return fail(SPHB200_ESTATE, "full_time_step before half_time_step");
        CK(cudaSetDevice(device));
        invalidate_lists();
        k_full_step<T, D><<<grid_for(n), 256, 0, stream>>>(A.p, B.p, acc.p, drhodt.p, Ah.p, type.p, gradC.p, divr.p, 0, (int)n, ph,
                                                           dt, d_ctl.p);
        ++launches;
        CK(cudaGetLastError());
        return SPHB200_OK;
    }
    int download_half(void *pos_h, void *vel_h, void *rho_h, void *press_h) override {
        if (!have_half) return fail(SPHB200_ESTATE, "download_half before half_time_step");
        CK(cudaSetDevice(device));
        unsigned char *sp = stage.p;
        size_t vb = (size_t)n * D * sizeof(T), sb = (size_t)n * sizeof(T);
        T *d_pos = (T *)sp; sp += vb;
        T *d_vel = (T *)sp; sp += vb;
        T *d_rho = (T *)sp; sp += sb;
        T *d_pr = (T *)sp; sp += sb;
        k_unpack_download<T, D><<<grid_for(n), 256, 0, stream>>>((int)n, Ah.p, Bh.p, nullptr, d_pos, d_vel, nullptr, d_rho, d_pr);
        ++launches;
        if (pos_h) CK(cudaMemcpyAsync(pos_h, d_pos, vb, cudaMemcpyDeviceToHost, stream));
        if (vel_h) CK(cudaMemcpyAsync(vel_h, d_vel, vb, cudaMemcpyDeviceToHost, stream));
        if (rho_h) CK(cudaMemcpyAsync(rho_h, d_rho, sb, cudaMemcpyDeviceToHost, stream));
        if (press_h) CK(cudaMemcpyAsync(press_h, d_pr, sb, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        return SPHB200_OK;
    }
    int download_aux(void *gradc_out, void *divr_out, void *ksum_out, void *kgrad_out) override {
        if (!uploaded) return fail(SPHB200_ESTATE, "download_aux before upload");
        CK(cudaSetDevice(device));
        T *d_tmp = (T *)stage.p;
        size_t vb = (size_t)n * D * sizeof(T), sb = (size_t)n * sizeof(T);
        if (gradc_out) {
            if (!prm.shifting) return fail(SPHB200_EINVAL, "gradC requires PlanarShifting");
            k_unpack_vec<T, D><<<grid_for(n), 256, 0, stream>>>((int)n, gradC.p, d_tmp);
            ++launches;
            CK(cudaMemcpyAsync(gradc_out, d_tmp, vb, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
        }
        if (divr_out) {
            if (!prm.shifting) return fail(SPHB200_EINVAL, "div r requires PlanarShifting");
            CK(cudaMemcpyAsync(divr_out, divr.p, sb, cudaMemcpyDeviceToHost, stream));
        }
        if (ksum_out) {
            if (!prm.kernel_output) return fail(SPHB200_EINVAL, "kernel sums require StoreKernelOutput");
            CK(cudaMemcpyAsync(ksum_out, ksum.p, sb, cudaMemcpyDeviceToHost, stream));
        }
        if (kgrad_out) {
            if (!prm.kernel_output) return fail(SPHB200_EINVAL, "kernel gradient sums require StoreKernelOutput");
            k_unpack_vec<T, D><<<grid_for(n), 256, 0, stream>>>((int)n, kgrad.p, d_tmp);
            ++launches;
            CK(cudaMemcpyAsync(kgrad_out, d_tmp, vb, cudaMemcpyDeviceToHost, stream));
        }
        CK(cudaStreamSynchronize(stream));
        return SPHB200_OK;
    }

    // per-stage device times of ONE extra step (ms), see SPHB200_STAGE_* in include/sphb200.h: the
    // reference's TimerOutputs labels "01" .. "12" (src/SPHCellList.jl:748-800) grouped the way the fused
    // kernels group them.  Advances the simulation by one step.
    int stage_times(double *ms_out, int cnt) override {
        if (!uploaded || !have_cells) return fail(SPHB200_ESTATE, "stage_times needs a running simulation");
        if (!ms_out || cnt < 1) return fail(SPHB200_EINVAL, "stage_times: no output array");
        CK(cudaSetDevice(device));
        for (int k = 0; k < cnt; ++k) ms_out[k] = 0.0;
        if (slab.active) return slab_stage_times(ms_out, cnt);
        cudaEvent_t ev[10];
        for (auto &e : ev) CK(cudaEventCreate(&e));
        int rc = 0;
        CK(cudaEventRecord(ev[9], stream));
        if ((rc = enqueue_step_head())) return rc;
        if ((rc = enqueue_step_body(ev))) return rc;
        CK(cudaStreamSynchronize(stream));
        double st[SPHB200_N_STAGES] = {0};
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, ev[9], ev[0]));
        st[SPHB200_STAGE_TIMESTEP] = ms;
        for (int k = 0; k < 8; ++k) {
            CK(cudaEventElapsedTime(&ms, ev[k], ev[k + 1]));
            st[SPHB200_STAGE_REBUILD + k] = ms;
        }
        for (int k = 0; k < cnt && k < SPHB200_N_STAGES; ++k) ms_out[k] = st[k];
        for (auto &e : ev) cudaEventDestroy(e);
        if ((rc = sync_ctl())) return rc;
        if (h_ctl->error) return fail(h_ctl->error, "device reported error %d", h_ctl->error);
        return SPHB200_OK;
    }

    // ------------------------------------------------------------------ slabs (sph_slab.cuh)
    int comm_init(const uint8_t *uid, int rank, int world, int axis) override;
    int set_slab(int64_t lo, int64_t hi) override;
    int set_ghost_nodes(int64_t ng, const void *points, const int64_t *ids) override;
    int column_histogram(int axis, int64_t *cell_min, int64_t *n_columns, int64_t *counts, int64_t cap) override;
#undef CK
};

}  // namespace

#include "sph_slab_impl.cuh"

// =================================================================================================
// C-ABI
// =================================================================================================
extern "C" {

int sphb200_abi_version(void) { return SPHB200_ABI_VERSION; }

int sphb200_create(const sphb200_params *p, int device, sphb200_sim **out) {
    if (!p || !out) {
        g_create_error = "sphb200_create: null argument";
        return SPHB200_EINVAL;
    }
    *out = nullptr;
    if (p->abi_version != SPHB200_ABI_VERSION) {
        g_create_error = "sphb200_create: ABI version mismatch";
        return SPHB200_EINVAL;
    }
    if ((p->dim != 2 && p->dim != 3) || (p->real_bytes != 4 && p->real_bytes != 8)) {
        g_create_error = "sphb200_create: dim must be 2 or 3 and real_bytes 4 or 8";
        return SPHB200_EINVAL;
    }
    if (!(p->h > 0.0) || !(p->H > 0.0) || !(p->rho0 > 0.0) || !(p->m0 > 0.0) || !(p->c0 > 0.0) || !(p->cfl > 0.0) ||
        p->n_motions < 0 || p->n_motions > SPHB200_MAX_MOTIONS || p->kernel < 0 || p->kernel > 1 || p->viscosity < 0 ||
        p->viscosity > 3 || p->diffusion < 0 || p->diffusion > 3) {
        g_create_error = "sphb200_create: invalid constants or model selectors";
        return SPHB200_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1) {
        g_create_error = std::string("sphb200_create: no CUDA device available (") + cudaGetErrorString(e) +
                         "); libsphb200 has no CPU fallback";
        return SPHB200_ECUDA;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "sphb200_create: device index out of range";
        return SPHB200_EINVAL;
    }
    sphb200_sim *s = nullptr;
    int rc = SPHB200_OK;
    if (p->dim == 2 && p->real_bytes == 8) { auto *q = new Sim<double, 2>(*p, device); rc = q->init(); s = q; }
    else if (p->dim == 2) { auto *q = new Sim<float, 2>(*p, device); rc = q->init(); s = q; }
    else if (p->real_bytes == 8) { auto *q = new Sim<double, 3>(*p, device); rc = q->init(); s = q; }
    else { auto *q = new Sim<float, 3>(*p, device); rc = q->init(); s = q; }
    if (rc != SPHB200_OK) {
        g_create_error = s->err;
        delete s;
        return rc;
    }
    *out = s;
    return SPHB200_OK;
}

int sphb200_destroy(sphb200_sim *sim) {
    delete sim;
    return SPHB200_OK;
}
const char *sphb200_last_error(const sphb200_sim *sim) { return sim ? sim->err.c_str() : g_create_error.c_str(); }

#define NEED(s)  \
    if (!(s)) return SPHB200_EINVAL

int sphb200_upload(sphb200_sim *s, int64_t n, const void *pos, const void *vel, const void *acc, const void *rho,
                   const uint8_t *type, const uint64_t *group, const int64_t *id, const void *gp, const void *gn) {
    NEED(s);
    return s->upload(n, pos, vel, acc, rho, type, group, id, gp, gn);
}
int sphb200_download(sphb200_sim *s, int order, void *pos, void *vel, void *acc, void *rho, void *press, int64_t *id,
                     uint8_t *type, uint64_t *group, int64_t *cells) {
    NEED(s);
    return s->download(order, pos, vel, acc, rho, press, id, type, group, cells);
}
int64_t sphb200_num_particles(const sphb200_sim *s) { return s ? s->num_particles() : 0; }
int sphb200_set_time(sphb200_sim *s, double t, int64_t it) { NEED(s); return s->set_time(t, it); }
int sphb200_simulation_loop(sphb200_sim *s, double t, sphb200_report *r) { NEED(s); return s->simulation_loop(t, r); }
int sphb200_step(sphb200_sim *s, int64_t n, int reset, sphb200_report *r) { NEED(s); return s->step(n, reset, r); }
int sphb200_get_report(sphb200_sim *s, sphb200_report *r) { NEED(s); return s->get_report(r); }
int64_t sphb200_launch_count(const sphb200_sim *s) { return s ? s->launch_count() : 0; }
int sphb200_update_neighbors(sphb200_sim *s, int64_t *ic) { NEED(s); return s->update_neighbors(ic); }
int sphb200_get_cell_list(sphb200_sim *s, int64_t *nc, int64_t *cells, int64_t *start) { NEED(s); return s->get_cell_list(nc, cells, start); }
int sphb200_pressure(sphb200_sim *s, int half) { NEED(s); return s->pressure(half); }
int sphb200_neighbor_loop(sphb200_sim *s, int pass, void *d, void *a) { NEED(s); return s->neighbor_loop(pass, d, a); }
int sphb200_delta_t(sphb200_sim *s, double *dt) { NEED(s); return s->delta_t(dt); }
int sphb200_progress_motion(sphb200_sim *s, double dt2) { NEED(s); return s->progress_motion(dt2); }
int sphb200_apply_mdbc(sphb200_sim *s) { NEED(s); return s->apply_mdbc(); }
int sphb200_half_time_step(sphb200_sim *s, double dt2) { NEED(s); return s->half_time_step(dt2); }
int sphb200_full_time_step(sphb200_sim *s, double dt) { NEED(s); return s->full_time_step(dt); }
int sphb200_download_half(sphb200_sim *s, void *p, void *v, void *r, void *pr) { NEED(s); return s->download_half(p, v, r, pr); }
int sphb200_download_aux(sphb200_sim *s, void *g, void *d, void *k, void *kg) { NEED(s); return s->download_aux(g, d, k, kg); }
int sphb200_set_stream(sphb200_sim *s, void *stream) { NEED(s); return s->set_stream(stream); }
int sphb200_set_option(sphb200_sim *s, const char *name, double value) { NEED(s); return s->set_option(name, value); }
int sphb200_get_stat(sphb200_sim *s, const char *name, double *value) { NEED(s); return s->get_stat(name, value); }
int sphb200_stage_times(sphb200_sim *s, double *ms, int n) { NEED(s); return s->stage_times(ms, n); }
int sphb200_comm_unique_id(uint8_t id_out[128]) { return slab_unique_id(id_out); }
int sphb200_comm_init(sphb200_sim *s, const uint8_t id[128], int rank, int world, int axis) { NEED(s); return s->comm_init(id, rank, world, axis); }
int sphb200_set_slab(sphb200_sim *s, int64_t lo, int64_t hi) { NEED(s); return s->set_slab(lo, hi); }
int sphb200_set_ghost_nodes(sphb200_sim *s, int64_t ng, const void *points, const int64_t *ids) { NEED(s); return s->set_ghost_nodes(ng, points, ids); }
int sphb200_column_histogram(sphb200_sim *s, int axis, int64_t *cmin, int64_t *ncol, int64_t *counts, int64_t cap) {
    NEED(s);
    return s->column_histogram(axis, cmin, ncol, counts, cap);
}
}
