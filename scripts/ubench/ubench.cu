// Micro-benchmarks that decide design questions of the list kernel on B200 (sm_100a):
//   issue rate of FFMA vs packed FFMA2 (fma.rn.f32x2), MUFU rate, LDS.128 gather cost vs bank-group conflicts.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/ubench.bin scripts/ubench/ubench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__global__ void k_ffma(float *out, int iters) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float b = 1.0001f, c = 0.5f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__global__ void k_ffma2(float *out, int iters) {
    unsigned long long a[8];
    for (int u = 0; u < 8; ++u) a[u] = ((unsigned long long)__float_as_uint((float)threadIdx.x + u) << 32) | __float_as_uint((float)u);
    const unsigned long long b = ((unsigned long long)__float_as_uint(1.0001f) << 32) | __float_as_uint(1.0002f);
    const unsigned long long c = ((unsigned long long)__float_as_uint(0.5f) << 32) | __float_as_uint(0.25f);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int v = 0; v < 8; ++v) a[v] = ffma2(a[v], b, c);
        }
    }
    float s = 0;
    for (int u = 0; u < 8; ++u) s += __uint_as_float((unsigned)a[u]) + __uint_as_float((unsigned)(a[u] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu(float *out, int iters) {
    float a[8];
    for (int u = 0; u < 8; ++u) a[u] = 1.5f + threadIdx.x + u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int v = 0; v < 8; ++v) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[v]));
        }
    }
    float s = 0;
    for (int u = 0; u < 8; ++u) s += a[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: per 8 FFMA `nm` MUFU
template <int NM>
__global__ void k_mix(float *out, int iters) {
    float a[8], m[8];
    for (int u = 0; u < 8; ++u) { a[u] = threadIdx.x + u; m[u] = 1.5f + u; }
    const float b = 1.0001f, c = 0.5f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int v = 0; v < 8; ++v) a[v] = fmaf(a[v], b, c);
#pragma unroll
            for (int v = 0; v < NM; ++v) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(m[v]));
        }
    }
    float s = 0;
    for (int u = 0; u < 8; ++u) s += a[u] + m[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// LDS.128 gathers: lane reads record idx = table[step][lane]; mode decides the conflict pattern
__global__ void k_lds(float *out, int iters, int mode) {
    extern __shared__ float4 sm[];
    const int n = 2048;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned idx = 0;
    // mode 0: conflict-free rotation (lane q reads group (q+k)%8); 1: all lanes same group, distinct records (8-way);
    // 2: pseudo-random records; 3: broadcast
    unsigned seed = threadIdx.x * 2654435761u + 12345u;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = seed * 1664525u + 1013904223u;
            const unsigned r = (seed >> 12) % 250u;
            if (mode == 0) idx = r * 8u + ((lane + u) & 7);
            else if (mode == 1) idx = ((r + lane * 7u) % 250u) * 8u + 3u;
            else if (mode == 2) idx = (seed >> 10) & 2047u;
            else idx = (unsigned)(i + u) & 2047u;
            float4 v = sm[idx];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <class F>
static float timeit(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount; const double ghz = p.clockRate * 1e-6;
    float *out; cudaMalloc(&out, sizeof(float) * sms * 1024 * 4);
    const int iters = 4000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int th = warps * 32;
        const double winst = (double)iters * 64 * warps * sms;   // warp-instructions of the measured kind
        float t1 = timeit([&] { k_ffma<<<sms, th>>>(out, iters); });
        float t2 = timeit([&] { k_ffma2<<<sms, th>>>(out, iters); });
        float t3 = timeit([&] { k_mufu<<<sms, th>>>(out, iters); });
        float t4 = timeit([&] { k_mix<2><<<sms, th>>>(out, iters); });
        float t5 = timeit([&] { k_mix<4><<<sms, th>>>(out, iters); });
        printf("{\"warps_per_sm\": %d, \"ffma_winst_per_clk_sm\": %.3f, \"ffma2_winst_per_clk_sm\": %.3f, \"mufu_winst_per_clk_sm\": %.3f, "
               "\"mix8f2m_ffma_per_clk_sm\": %.3f, \"mix8f4m_ffma_per_clk_sm\": %.3f, \"clock_ghz_nominal\": %.3f}\n",
               warps, winst / (t1 * 1e-3) / (ghz * 1e9) / sms, winst / (t2 * 1e-3) / (ghz * 1e9) / sms,
               winst / (t3 * 1e-3) / (ghz * 1e9) / sms, winst / (t4 * 1e-3) / (ghz * 1e9) / sms, winst / (t5 * 1e-3) / (ghz * 1e9) / sms, ghz);
    }
    cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 16);
    for (int mode = 0; mode < 4; ++mode) {
        const int warps = 8, th = warps * 32, it2 = 2000;
        float t = timeit([&] { k_lds<<<sms, th, 2048 * 16>>>(out, it2, mode); });
        const double lds = (double)it2 * 8 * warps;   // per SM
        printf("{\"lds128_mode\": %d, \"clk_per_warp_lds128\": %.2f}\n", mode, (t * 1e-3) * ghz * 1e9 / lds);
    }
    return 0;
}
