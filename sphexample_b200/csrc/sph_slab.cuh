// sph_slab.cuh — multi-GPU slab decomposition: NCCL binding and per-rank exchange state.
//
// One process per GPU.  The domain is cut into slabs of whole cell layers along the slab axis
// `s` (the most significant component of the cell key, so a layer is ONE contiguous index range of
// the cell-sorted table).  Each rank owns the particles whose slab coordinate lies in
// [own_lo, own_hi) and holds a read-only copy of the neighbouring ranks' adjacent layer (the halo,
// one cell = one interaction radius wide).  Traffic, all over NCCL send/recv on the handle's stream
// (NVLink 5 / NVSwitch between the GPUs of a node):
//   * every step:      one 5-word all-reduce (max) of the Δt / Δx / |v| reductions + error flag;
//   * every half step: the two boundary layers' packed state (A, B arrays) to the two neighbours
//                      — plain contiguous ranges, no pack kernel, because sender and receiver hold
//                      the layer in the same order (both sort by the same order keys).  The bricks
//                      of the two boundary layers are computed first; their exchange then runs on
//                      a second (high-priority) stream while the interior bricks are computed.
//                      (Experimental, SPHB200_SLAB_WAITVALUE=1: one launch per pass, the CTA that
//                      retires the last boundary brick raises a flag on which the exchange stream
//                      waits with cuStreamWaitValue32.)
//   * every rebuild:   migration of the particles that left the slab, then the boundary layers'
//                      full records.
// NCCL is dlopen()ed (the copy already loaded in the process, e.g. torch's, else libnccl.so.2) so
// that the single-GPU library has no link-time dependency on it.
#pragma once

#include <dlfcn.h>

#include <string>

#include "sph_device.cuh"

namespace sph {

namespace nccl {
struct Comm;
typedef Comm *comm_t;
struct UniqueId { char internal[128]; };
enum { Success = 0 };
enum { Int8 = 0, Int32 = 2, Uint64 = 5, Float64 = 8 };   // ncclDataType_t values used here
enum { Sum = 0, Max = 2 };                  // ncclRedOp_t

struct Api {
    void *lib = nullptr;
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(comm_t *, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*CommAbort)(comm_t) = nullptr;   // optional
    int (*Send)(const void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;

    bool load(std::string &err) {
        if (lib) return true;
        const char *env = getenv("SPHB200_NCCL_LIB");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        // prefer a copy that is already mapped into the process (torch ships its own)
        for (const char *nm : names)
            if (nm && *nm && (lib = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL))) break;
        for (const char *nm : names) {
            if (lib) break;
            if (nm && *nm) lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!lib) {
            err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
            return false;
        }
#define SPH_NCCL_SYM(field, name)                                   \
    *(void **)(&field) = dlsym(lib, name);                          \
    if (!field) {                                                   \
        err = std::string("NCCL symbol missing: ") + name;          \
        lib = nullptr;                                              \
        return false;                                               \
    }
        SPH_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        SPH_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        SPH_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        SPH_NCCL_SYM(Send, "ncclSend")
        SPH_NCCL_SYM(Recv, "ncclRecv")
        SPH_NCCL_SYM(AllReduce, "ncclAllReduce")
        SPH_NCCL_SYM(GroupStart, "ncclGroupStart")
        SPH_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        SPH_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef SPH_NCCL_SYM
        *(void **)(&CommAbort) = dlsym(lib, "ncclCommAbort");
        return true;
    }
};

inline Api &api() {
    static Api a;
    return a;
}
}  // namespace nccl

struct SlabComm {
    bool active = false;
    bool dead = false;              // a wait timed out: the communicator is aborted, not destroyed, on teardown
    int rank = 0, world = 1;
    int left = -1, right = -1;      // neighbour ranks along the slab axis (-1: domain end)
    nccl::comm_t comm = nullptr;
    cudaStream_t xstream = nullptr; // the half-step halo exchanges run here, beside the interior bricks
    cudaEvent_t ev_bnd = nullptr, ev_x = nullptr;
    // single-launch passes: the kernel raises *d_flag to `epoch` when its boundary bricks are done,
    // the exchange stream waits for that value (driver API cuStreamWaitValue32)
    unsigned *d_flag = nullptr;
    unsigned epoch = 0;
    int (*wait_value32)(cudaStream_t, unsigned long long, unsigned, unsigned) = nullptr;
    int *d_counts = nullptr;        // 8 ints on the device: migration / halo count exchange
    int *h_counts = nullptr;        // pinned mirror
    // host mirrors of the table layout after the last rebuild (sorted order):
    //   [0, own_p0) left halo | [own_p0, l1) first owned layer | ... | [l2, own_p1) last owned layer | [own_p1, n) right halo
    int own_p0 = 0, own_p1 = 0, l1 = 0, l2 = 0;
    long long halo_bytes_per_step = 0;   // bytes this rank sends per step in the two half-step exchanges
    long long n_migrated = 0;            // particles handed to neighbours so far
};

int slab_unique_id(uint8_t *id_out);

// ---------------------------------------------------------------------------------------------
// kernels of the rebuild-time exchange
// ---------------------------------------------------------------------------------------------
// particles of the owned range that have left [own_lo, own_hi) along the slab axis -> index lists
template <class T, int D>
__global__ void k_slab_classify(const typename Lay<T, D>::TA *__restrict__ A, int p0, int p1, double inv_cutoff, int ax_s,
                                int own_lo, int own_hi, int *__restrict__ list_l, int *__restrict__ list_r,
                                int *__restrict__ counts, Ctl *ctl) {
    if (ctl->error) return;
    for (int i = p0 + blockIdx.x * blockDim.x + threadIdx.x; i < p1; i += gridDim.x * blockDim.x) {
        T x[D];
        Lay<T, D>::pos(A[i], x);
        double xs = (double)x[ax_s];
        double t = trunc(fma(fabs(xs), inv_cutoff, 0.5));
        if (!(t < 1.0e9)) {
            atomicCAS(&ctl->error, 0, SPH_ERR_ENUMERIC);
            continue;
        }
        int c = ((xs > 0.0) - (xs < 0.0)) * (int)t;
        if (c < own_lo) list_l[atomicAdd(&counts[0], 1)] = i;
        else if (c >= own_hi) list_r[atomicAdd(&counts[1], 1)] = i;
    }
}

// a received migrant must land inside the owned range: a particle cannot cross a whole slab
// between two rebuilds (it moves less than one cell), so anything else is a decomposition error
template <class T, int D>
__global__ void k_slab_check_arrivals(const typename Lay<T, D>::TA *__restrict__ A, int q0, int q1, double inv_cutoff,
                                      int ax_s, int own_lo, int own_hi, Ctl *ctl) {
    if (ctl->error) return;
    for (int i = q0 + blockIdx.x * blockDim.x + threadIdx.x; i < q1; i += gridDim.x * blockDim.x) {
        T x[D];
        Lay<T, D>::pos(A[i], x);
        double xs = (double)x[ax_s];
        int c = ((xs > 0.0) - (xs < 0.0)) * (int)trunc(fma(fabs(xs), inv_cutoff, 0.5));
        if (c < own_lo || c >= own_hi) atomicCAS(&ctl->error, 0, SPH_ERR_ESTATE);
    }
}

// end of a pass: raise the flag unconditionally — it is already up unless the pass ran empty
// (error / pause), in which case the exchange stream must not be left waiting
__global__ void k_slab_signal(unsigned *flag, unsigned epoch) { atomicMax(flag, epoch); }

__global__ void k_slab_pre_allreduce(Ctl *ctl) {
    ctl->red_err = ctl->error ? (unsigned long long)(-ctl->error) : 0ull;
}

// per-column particle histogram along one axis (for choosing balanced slab edges)
template <class T, int D>
__global__ void k_column_histogram(const typename Lay<T, D>::TA *__restrict__ A, int p0, int p1, double inv_cutoff, int axis,
                                   int cmin, int ncol, int *__restrict__ hist) {
    for (int i = p0 + blockIdx.x * blockDim.x + threadIdx.x; i < p1; i += gridDim.x * blockDim.x) {
        T x[D];
        Lay<T, D>::pos(A[i], x);
        double xs = (double)x[axis];
        int c = ((xs > 0.0) - (xs < 0.0)) * (int)trunc(fma(fabs(xs), inv_cutoff, 0.5));
        c = min(max(c - cmin, 0), ncol - 1);
        atomicAdd(&hist[c], 1);
    }
}

}  // namespace sph
