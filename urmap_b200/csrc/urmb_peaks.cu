// Measured denominators for the two rooflines SURVEY.md §8d names (bench.py reports against them):
//   urmb_peak_gather : random 32-byte-sector reads over a large device buffer (the UFI blob), the bound of the
//                      slot-probe / list-hop / candidate-window gathers;
//   urmb_peak_alu    : 32-bit integer ALU issue rate, the bound of the gapless-extension and DP-cell work.
// Measurement helpers only: no mapping result depends on them.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/urmb.h"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}

// Every thread issues `per_thread` independent loads of VEC 32-bit words at random sector-aligned addresses, UNROLL in
// flight at a time; the xor of everything read is stored so that nothing is optimised away.
template <int VEC>
__global__ void __launch_bounds__(256) gather_peak_kernel(const uint32_t *buf, uint64_t n_sectors, uint32_t per_thread,
                                                          uint64_t seed, uint32_t *sink) {
    constexpr int UNROLL = 8;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < per_thread; i += UNROLL) {
        uint32_t v[UNROLL][VEC];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint64_t s = __umul64hi(mix64(seed + tid * per_thread + i + u), n_sectors);   // uniform in [0, n_sectors)
            const uint32_t *p = buf + s * 8;
            if constexpr (VEC == 1) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v[u][0]) : "l"(p));
            else if constexpr (VEC == 2)
                asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v[u][0]), "=r"(v[u][VEC > 1 ? 1 : 0]) : "l"(p));
            else
                asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[u][0]), "=r"(v[u][VEC > 1 ? 1 : 0]), "=r"(v[u][VEC > 2 ? 2 : 0]), "=r"(v[u][VEC > 3 ? 3 : 0])
                             : "l"(p));
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc ^= v[u][k];
    }
    if (acc == 0x9e3779b9u) sink[0] = acc;   // practically never; keeps the loads alive
}

// Four independent chains of LOP3 / IADD3 / SHF per thread, 32 ALU ops per trip in the SASS (IADD+LOP3 pairs merge partly).
__global__ void __launch_bounds__(256) alu_peak_kernel(uint32_t trips, uint32_t *sink) {
    uint32_t a = threadIdx.x, b = blockIdx.x, c = a * 2654435761u, d = b ^ 0x5bd1e995u;
    for (uint32_t t = 0; t < trips; ++t) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a = (a + b) ^ c;  b = (b ^ d) + a;  c = (c + a) ^ d;  d = (d ^ b) + c;   // 8 ops over 4 chains
        }
        a = __funnelshift_l(a, b, 7); c = __funnelshift_l(c, d, 9);
    }
    if ((a ^ b ^ c ^ d) == 0x12345u) sink[0] = a;
}

std::string g_peak_err;

}  // namespace

extern "C" {

int urmb_peak_gather(const void *d_buf, uint64_t n_bytes, uint32_t access_bytes, uint64_t n_access, float *ms) {
    if (!d_buf || !ms || n_bytes < 4096 || (access_bytes != 4 && access_bytes != 8 && access_bytes != 16)) return URMB_E_ARG;
    if (((uintptr_t)d_buf & 31u) != 0) return URMB_E_ARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return URMB_E_NODEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256, blocks = sms * 8;
    uint32_t per_thread = (uint32_t)((n_access + (uint64_t)threads * blocks - 1) / ((uint64_t)threads * blocks));
    per_thread = (per_thread + 7u) & ~7u;
    uint32_t *sink = nullptr;
    cudaEvent_t e0, e1;
    if (cudaMalloc(&sink, 4) != cudaSuccess) return URMB_E_CUDA;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const uint64_t n_sectors = n_bytes / 32;
    float best = 0.f;
    for (int rep = 0; rep < 3; ++rep) {   // rep 0 warms up; best of the next two
        cudaEventRecord(e0);
        const uint64_t seed = 0x1234567ull + (uint64_t)rep * 0x9e3779b97f4a7c15ull;
        if (access_bytes == 4) gather_peak_kernel<1><<<blocks, threads>>>((const uint32_t *)d_buf, n_sectors, per_thread, seed, sink);
        else if (access_bytes == 8) gather_peak_kernel<2><<<blocks, threads>>>((const uint32_t *)d_buf, n_sectors, per_thread, seed, sink);
        else gather_peak_kernel<4><<<blocks, threads>>>((const uint32_t *)d_buf, n_sectors, per_thread, seed, sink);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { g_peak_err = cudaGetErrorString(e); cudaFree(sink); return URMB_E_CUDA; }
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0 && (best == 0.f || t < best)) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    // time is for per_thread * threads * blocks accesses; scale to the count the caller asked for
    *ms = best * (float)((double)n_access / ((double)per_thread * threads * blocks));
    return URMB_OK;
}

int urmb_peak_alu(uint64_t ops_per_thread, float *ms, double *total_ops) {
    if (!ms || !total_ops || ops_per_thread < 32) return URMB_E_ARG;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return URMB_E_NODEVICE;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256, blocks = sms * 8;
    const uint32_t trips = (uint32_t)(ops_per_thread / 32);
    uint32_t *sink = nullptr;
    cudaEvent_t e0, e1;
    if (cudaMalloc(&sink, 4) != cudaSuccess) return URMB_E_CUDA;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        alu_peak_kernel<<<blocks, threads>>>(trips, sink);
        cudaEventRecord(e1);
        cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) { g_peak_err = cudaGetErrorString(e); cudaFree(sink); return URMB_E_CUDA; }
        float t = 0.f;
        cudaEventElapsedTime(&t, e0, e1);
        if (rep > 0 && (best == 0.f || t < best)) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *ms = best;
    *total_ops = 32.0 * trips * (double)threads * blocks;
    return URMB_OK;
}

}  // extern "C"
