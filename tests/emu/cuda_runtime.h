// tests/emu/cuda_runtime.h -- DEBUGGING HARNESS, NOT PRODUCT CODE.
//
// A minimal lock-step warp emulator that lets urmap_b200/csrc/urmb_kernels.cu be compiled by g++ and
// stepped through on a machine without a GPU (this build container has none).  Each warp runs as 32
// ucontext fibers; every warp collective (__shfl*, __ballot, __any, __syncwarp) is a rendez-vous that
// also CHECKS that all 32 lanes reached the same collective -- i.e. it catches divergence bugs that
// would hang or corrupt on real hardware.  It is only ever built by tests/ (see tests/emu/build.sh);
// the package never loads it and the shipped library has no CPU path.
#pragma once
#define URMB_EMU 1
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n)

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
typedef void *cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaGetLastError() { return 0; }
template <class F> inline int cudaFuncSetAttribute(F, int, int) { return 0; }
template <class F> inline int cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return 0; }

namespace emu {
enum Kind { K_NONE, K_SHFL, K_SHFL_UP, K_SHFL_DOWN, K_SHFL_XOR, K_BALLOT, K_ANY, K_SYNC };
uint64_t collective(Kind kind, uint64_t value, int arg, int site);
uint8_t *smem_base();
void launch(const std::function<void()> &body, int grid, int block, size_t smem);
}  // namespace emu

#define URMB_DYN_SMEM(name) uint8_t *name = emu::smem_base()
#define URMB_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch([&]() { kern(__VA_ARGS__); }, (grid), (block), (smem))

template <class T> inline T __ldg(const T *p) { return *p; }

template <class T> inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <class T> inline T emu_shfl(T v, int src, int line) {
    return emu_unbits<T>(emu::collective(emu::K_SHFL, emu_bits(v), src & 31, line));
}
template <class T> inline T emu_shfl_up(T v, unsigned delta, int line) {
    return emu_unbits<T>(emu::collective(emu::K_SHFL_UP, emu_bits(v), (int)delta, line));
}
template <class T> inline T emu_shfl_down(T v, unsigned delta, int line) {
    return emu_unbits<T>(emu::collective(emu::K_SHFL_DOWN, emu_bits(v), (int)delta, line));
}
template <class T> inline T emu_shfl_xor(T v, int mask, int line) {
    return emu_unbits<T>(emu::collective(emu::K_SHFL_XOR, emu_bits(v), mask, line));
}
inline unsigned emu_ballot(int pred, int line) { return (unsigned)emu::collective(emu::K_BALLOT, pred ? 1 : 0, 0, line); }
inline int emu_any(int pred, int line) { return (int)emu::collective(emu::K_ANY, pred ? 1 : 0, 0, line); }
inline void emu_syncwarp(int line) { emu::collective(emu::K_SYNC, 0, 0, line); }
#define __shfl_sync(m, v, l) emu_shfl((v), (l), __LINE__)
#define __shfl_up_sync(m, v, d) emu_shfl_up((v), (d), __LINE__)
#define __shfl_down_sync(m, v, d) emu_shfl_down((v), (d), __LINE__)
#define __shfl_xor_sync(m, v, d) emu_shfl_xor((v), (d), __LINE__)
#define __ballot_sync(m, p) emu_ballot((p), __LINE__)
#define __any_sync(m, p) emu_any((p), __LINE__)
#define __syncwarp() emu_syncwarp(__LINE__)

inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
inline unsigned atomicMax(unsigned *p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
inline unsigned atomicSub(unsigned *p, unsigned v) { unsigned o = *p; *p = o - v; return o; }
inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) { unsigned o = *p; if (o == cmp) *p = v; return o; }
inline void __syncthreads() {}
#define __shared__ static
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __ffsll(long long v) { return v ? __builtin_ctzll((unsigned long long)v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t shift) {
    shift &= 31;
    return shift ? (hi << shift) | (lo >> (32 - shift)) : hi;
}
using std::max;
using std::min;
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline int min(int a, unsigned b) { return a < (int)b ? a : (int)b; }
