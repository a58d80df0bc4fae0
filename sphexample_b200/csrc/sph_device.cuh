// sph_device.cuh — device data layout, control block and small PTX helpers of libsphb200.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sph_bricks.h"
#include "sph_control.h"
#include "sph_physics.cuh"

namespace sph {

// ---------------------------------------------------------------------------------------------
// Packed particle state in HBM (cell-sorted order).  One 16-byte (fp32) / 32-byte (fp64)
// vector per particle and array, so that a neighbour row is one contiguous, 16-byte aligned
// byte span that a single cp.async.bulk (TMA) can stage into shared memory.
//   3D:  A = {x, y, z, ρ±}   B = {vx, vy, vz, P}
//   2D:  A = {x, z, ρ±, P}   B = {vx, vz}
// ρ± carries the MotionLimiter in its sign: +ρ for fluid (ML = 1), −ρ for boundary (ML = 0)
// (src/PreProcess.jl:89-98); densities are strictly positive so the encoding is exact.
// ---------------------------------------------------------------------------------------------
template <class T> struct alignas(sizeof(T) * 4) V4 { T x, y, z, w; };
template <class T> struct alignas(sizeof(T) * 2) V2 { T x, y; };

template <class T, int D> struct Lay;
template <class T> struct Lay<T, 3> {
    using TA = V4<T>;
    using TB = V4<T>;
    using TV = V4<T>;   // plain D-vector storage (acceleration, ghost points …)
    static __device__ __forceinline__ void unpack(const TA &a, const TB &b, T *x, T *v, T &rs, T &P) {
        x[0] = a.x; x[1] = a.y; x[2] = a.z; rs = a.w;
        v[0] = b.x; v[1] = b.y; v[2] = b.z; P = b.w;
    }
    static __device__ __forceinline__ void pack(TA &a, TB &b, const T *x, const T *v, T rs, T P) {
        a.x = x[0]; a.y = x[1]; a.z = x[2]; a.w = rs;
        b.x = v[0]; b.y = v[1]; b.z = v[2]; b.w = P;
    }
    static __device__ __forceinline__ void pos(const TA &a, T *x) { x[0] = a.x; x[1] = a.y; x[2] = a.z; }
    static __device__ __forceinline__ void vel(const TB &b, T *v) { v[0] = b.x; v[1] = b.y; v[2] = b.z; }
    static __device__ __forceinline__ T rhos(const TA &a) { return a.w; }
    static __device__ __forceinline__ void set_rhos(TA &a, T rs) { a.w = rs; }
    static __device__ __forceinline__ void set_P(TA &, TB &b, T P) { b.w = P; }
    static __device__ __forceinline__ void getv(const TV &a, T *x) { x[0] = a.x; x[1] = a.y; x[2] = a.z; }
    static __device__ __forceinline__ TV mkv(const T *x) { TV r; r.x = x[0]; r.y = x[1]; r.z = x[2]; r.w = T(0); return r; }
};
template <class T> struct Lay<T, 2> {
    using TA = V4<T>;
    using TB = V2<T>;
    using TV = V2<T>;
    static __device__ __forceinline__ void unpack(const TA &a, const TB &b, T *x, T *v, T &rs, T &P) {
        x[0] = a.x; x[1] = a.y; rs = a.z; P = a.w;
        v[0] = b.x; v[1] = b.y;
    }
    static __device__ __forceinline__ void pack(TA &a, TB &b, const T *x, const T *v, T rs, T P) {
        a.x = x[0]; a.y = x[1]; a.z = rs; a.w = P;
        b.x = v[0]; b.y = v[1];
    }
    static __device__ __forceinline__ void pos(const TA &a, T *x) { x[0] = a.x; x[1] = a.y; }
    static __device__ __forceinline__ void vel(const TB &b, T *v) { v[0] = b.x; v[1] = b.y; }
    static __device__ __forceinline__ T rhos(const TA &a) { return a.z; }
    static __device__ __forceinline__ void set_rhos(TA &a, T rs) { a.z = rs; }
    static __device__ __forceinline__ void set_P(TA &a, TB &, T P) { a.w = P; }
    static __device__ __forceinline__ void getv(const TV &a, T *x) { x[0] = a.x; x[1] = a.y; }
    static __device__ __forceinline__ TV mkv(const T *x) { TV r; r.x = x[0]; r.y = x[1]; return r; }
};

// GravityFactor / MotionLimiter from ParticleType, src/PreProcess.jl:78-98 (Q6)
__host__ __device__ __forceinline__ float type_gf(uint8_t t) { return t == 1 ? -1.f : (t == 3 ? 1.f : 0.f); }
__host__ __device__ __forceinline__ float type_ml(uint8_t t) { return t == 1 ? 1.f : 0.f; }


// ---------------------------------------------------------------------------------------------
// atomics on non-negative reals via their bit patterns
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_nonneg(unsigned long long *addr, double v) {
    if (!(v >= 0.0)) v = __longlong_as_double(0x7ff8000000000000ll);  // NaN sorts above +inf
    if (v == 0.0) v = 0.0;                                             // -0.0 -> +0.0
    atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}
__device__ __forceinline__ double bits_to_double(unsigned long long b) { return __longlong_as_double((long long)b); }

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D TMA (cp.async.bulk) helpers; SASS: SYNCS.* / UBLKCP
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    uint32_t addr = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a non-zero multiple of 16
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// streaming 16-byte load through the read-only path, pinned where it is written (volatile): the
// neighbour lists are read once per pass and never written while a pass runs
__device__ __forceinline__ uint4 ld_nc_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <class T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u > v ? u : v;
    }
    return v;
}
template <class T> __device__ __forceinline__ T warp_min(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(0xffffffffu, v, o);
        v = u < v ? u : v;
    }
    return v;
}

}  // namespace sph
