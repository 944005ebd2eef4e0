// cuda_hostemu.h — TEST INFRASTRUCTURE (never linked into the product).
//
// Lets g++ compile the *unmodified kernel bodies* of the HBM-bound .cu files (elementwise.cu, transformer.cu,
// loss_optim.cu, augment_metrics.cu, runtime.cu) for the host, so that their index arithmetic, launch geometry,
// shared-memory reductions, warp shuffles and argument validation run in the build container (which has no GPU) through
// the very same C ABI (include/gdl_b200.h).  tests/hostemu/build.py rewrites only the CUDA-specific syntax:
//   kernel<<<grid, block, smem, stream>>>(args)   ->  hostemu::launch(grid, block, smem, [=]() { kernel(args); })
//   extern __shared__ T name[];                   ->  T* name = (T*)hostemu::dyn_smem();
//   __shared__                                    ->  static   (blocks run one after the other)
//   asm volatile(...)                             ->  hostemu::unsupported_asm()   (tcgen05 / TMA helpers: never reached)
//
// Execution model: the blocks of a grid run sequentially; the threads of a block are fibers (ucontext) scheduled round
// robin on one OS thread.  __syncthreads() and the warp shuffles are barriers over the live fibers of the block / warp
// (a fiber that returned no longer takes part, as on the GPU).  Deterministic; data races of the CUDA code are NOT
// detected, tensor-core / TMA kernels are NOT covered — this checks the scalar CUDA code, not the hardware.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>

#include "hostemu_tc.h"

#define GDL_HOSTEMU 1  // common.cuh: GDL_LAUNCH runs the kernel body on the host (hostemu::launch)

#ifndef __grid_constant__
#define __grid_constant__
#endif

// ---- built-in variables ------------------------------------------------------------------------------------------
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;
constexpr int warpSize = 32;

namespace hostemu {
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void* dyn_smem();
size_t dyn_smem_offset();
void sync_block();
void sync_warp();
uint64_t* warp_slots();  // 32 x 8-byte exchange slots of the calling fiber's warp
void sync_named(int id, int nthreads);
void yield_spin();      // a failed mbarrier poll: stay runnable, let the other fibers run
void note_progress();   // deadlock detection: something observable changed
int lane();
int warp_index();
int thread_linear();
[[noreturn]] void unsupported_asm();
}  // namespace hostemu

// ---- qualifiers --------------------------------------------------------------------------------------------------
#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)

// cuda_runtime.h offers the kernel-pointer overload only under nvcc
template <typename K>
inline cudaError_t cudaFuncSetAttribute(K* kernel, cudaFuncAttribute, int) {
  return kernel != nullptr ? cudaSuccess : cudaErrorInvalidDeviceFunction;
}

inline void __syncthreads() { hostemu::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { hostemu::sync_warp(); }
[[noreturn]] inline void __trap() { abort(); }
inline long long clock64() { return 0; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)(uintptr_t)p; }

template <typename T>
inline T hostemu_shfl(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of up to 8 bytes");
  uint64_t* s = hostemu::warp_slots();
  memcpy(&s[hostemu::lane()], &v, sizeof(T));
  hostemu::sync_warp();
  T r;
  memcpy(&r, &s[src_lane & 31], sizeof(T));
  hostemu::sync_warp();
  return r;
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return hostemu_shfl(v, hostemu::lane() ^ m); }
template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) { return hostemu_shfl(v, src); }
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  const int l = hostemu::lane();
  return hostemu_shfl(v, l + (int)d < 32 ? l + (int)d : l);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  const int l = hostemu::lane();
  return hostemu_shfl(v, l - (int)d >= 0 ? l - (int)d : l);
}

// ---- atomics (fibers are cooperative: plain read-modify-write is atomic here) ------------------------------------------
template <typename T>
inline T atomicAdd(T* p, T v) {
  T o = *p;
  *p = o + v;
  return o;
}
inline unsigned long long atomicAdd(unsigned long long* p, unsigned int v) { return atomicAdd(p, (unsigned long long)v); }
template <typename T>
inline T atomicMax(T* p, T v) {
  T o = *p;
  if (v > o) *p = v;
  return o;
}
template <typename T>
inline T atomicMin(T* p, T v) {
  T o = *p;
  if (v < o) *p = v;
  return o;
}
template <typename T>
inline T atomicExch(T* p, T v) {
  T o = *p;
  *p = v;
  return o;
}
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) {
  unsigned o = *p;
  if (o == cmp) *p = v;
  return o;
}

// ---- device math / intrinsics ----------------------------------------------------------------------------------------
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __frcp_rn(float a) { return 1.f / a; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __expf(float a) { return expf(a); }
inline float __logf(float a) { return logf(a); }
inline float __fdividef(float a, float b) { return a / b; }
inline float rsqrtf(float a) { return 1.f / sqrtf(a); }
inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline int __float_as_int(float f) { return (int)__float_as_uint(f); }
inline float __int_as_float(int i) { return __uint_as_float((unsigned)i); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline T __ldcs(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
inline void __threadfence() {}
inline void __threadfence_system() {}
template <typename T>
inline void __stcs(T* p, T v) { *p = v; }

// CUDA's mixed-type min / max overloads
using std::max;
using std::min;
inline long long min(long long a, int b) { return a < b ? a : (long long)b; }
inline long long min(int a, long long b) { return a < b ? (long long)a : b; }
inline long long max(long long a, int b) { return a > b ? a : (long long)b; }
inline long long max(int a, long long b) { return a > b ? (long long)a : b; }
inline unsigned min(unsigned a, int b) { return a < (unsigned)b ? a : (unsigned)b; }
inline unsigned min(int a, unsigned b) { return (unsigned)a < b ? (unsigned)a : b; }
inline float min(float a, double b) { return a < (float)b ? a : (float)b; }
inline float max(float a, double b) { return a > (float)b ? a : (float)b; }
