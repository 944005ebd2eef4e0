// common.cuh — sm_100a PTX wrappers shared by every kernel in libgdlb200.so.
//
// Everything here is hand-written inline PTX for Blackwell (tcgen05 / TMEM / TMA /
// mbarrier). No CUTLASS, no Triton. Compile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define GDL_DEVINL __device__ __forceinline__

namespace gdl {

constexpr int kNumSMsB200 = 148;

// ----------------------------------------------------------------------------------------
// error reporting across the C ABI (see include/gdl_b200.h)
// ----------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define GDL_CHECK_CUDA(expr)                                   \
  do {                                                         \
    int _st = ::gdl::check_cuda((expr), #expr);                \
    if (_st != 0) return _st;                                  \
  } while (0)

#define GDL_REQUIRE(cond, code, ...)                           \
  do {                                                         \
    if (!(cond)) {                                             \
      ::gdl::set_last_error(__VA_ARGS__);                      \
      return (code);                                           \
    }                                                          \
  } while (0)

// ----------------------------------------------------------------------------------------
// host: one-time, PER-DEVICE opt-in to > 48 KB of dynamic shared memory (the attribute belongs to the
// device's context, not to the process: a process that drives several GPUs must set it on each)
// ----------------------------------------------------------------------------------------
struct PerDeviceOnce {
  unsigned long long mask = 0;  // bit d = done on device d (benign race: setting the attribute twice is harmless)
};
template <typename K>
inline cudaError_t set_max_dyn_smem_once(PerDeviceOnce& once, K kernel, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = 1ull << (dev & 63);
  if (__atomic_load_n(&once.mask, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) __atomic_fetch_or(&once.mask, bit, __ATOMIC_RELEASE);
  return e;
}

// SM count of the CURRENT device (persistent kernels launch min(work, #SM) CTAs), cached per device
inline int device_sm_count() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return kNumSMsB200;
  int& n = cache[dev & 63];
  if (n == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = kNumSMsB200;
    n = v;
  }
  return n;
}

// ----------------------------------------------------------------------------------------
// kernel launches: programmatic dependent launch (PDL).
// Every kernel of the library is launched through GDL_LAUNCH and begins with GDL_PDL_ENTRY().  With the "pdl" option on
// (gdl_set_option("pdl", 1); the Python host sets it from GDL_PDL) a launch carries cudaLaunchAttributeProgrammaticStreamSerialization: the grid may
// become resident while its predecessor in the stream (or in the captured graph) is still running, and every thread then
// blocks in griddepcontrol.wait until that predecessor has COMPLETED and its memory is visible — nothing is read or
// written before it, so the stream's semantics are unchanged; only the launch latency between two dependent kernels
// (a model step is 700..6000 of them) is hidden.  griddepcontrol.launch_dependents right after the wait lets the NEXT
// kernel do the same.  Both instructions are no-ops in a grid launched without the attribute.  The rule "first
// statement of every __global__ function" is enforced by tests/test_boundary_cpu.py.
// ----------------------------------------------------------------------------------------
extern int g_opt_pdl;  // runtime.cu

GDL_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
GDL_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define GDL_PDL_ENTRY()    \
  do {                     \
    ::gdl::pdl_wait();     \
    ::gdl::pdl_trigger();  \
  } while (0)

#ifdef GDL_HOSTEMU
// tests/hostemu: the kernel body runs thread by thread on the host
#define GDL_LAUNCH(kern, grid, block, smem, stream, ...) \
  hostemu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); })
#else
template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_opt_pdl ? 1u : 0u;
  // a failed launch is picked up by the cudaGetLastError() that follows every launch site
  (void)cudaLaunchKernelEx(&cfg, kern, static_cast<Args&&>(args)...);
}
#define GDL_LAUNCH(kern, grid, block, smem, stream, ...) \
  ::gdl::launch_kernel(kern, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream), __VA_ARGS__)
#endif

// ----------------------------------------------------------------------------------------
// shared-memory address helpers
// ----------------------------------------------------------------------------------------
GDL_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

GDL_DEVINL uint32_t lane_id() { return threadIdx.x & 31; }

GDL_DEVINL bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
GDL_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
GDL_DEVINL void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
GDL_DEVINL void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
GDL_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
GDL_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
GDL_DEVINL uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must never hang the GPU box (that is a "strike");
// after ~4e9 SM cycles (≈2 s) the kernel traps instead.
GDL_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0) {
      if (clock64() - t0 > 4000000000ll) {
        printf("gdl: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x,
               threadIdx.x, smem_u32(bar), parity);
        __trap();
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ----------------------------------------------------------------------------------------
GDL_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
GDL_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}
GDL_DEVINL void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store (smem tile -> global, bulk async group completion); out-of-bounds box elements are clipped
GDL_DEVINL void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
GDL_DEVINL void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
GDL_DEVINL void bulk_wait_group_read() {  // <= N groups still READING their smem source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
GDL_DEVINL void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
GDL_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, TMEM loads
// ----------------------------------------------------------------------------------------
GDL_DEVINL void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp, .sync.aligned
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
GDL_DEVINL void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
GDL_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
GDL_DEVINL void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
GDL_DEVINL void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 and fp16 inputs, fp32 accum.
GDL_DEVINL void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed.
// (implies tcgen05.fence::before_thread_sync)
GDL_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive 32-bit columns: thread t of the warp gets lane (base_lane + t).
GDL_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
GDL_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: PTX ISA "tcgen05 matrix descriptor"/"instruction descriptor")
// ----------------------------------------------------------------------------------------
// swizzle_bytes in {32, 64, 128}
GDL_DEVINL uint32_t umma_layout_type(int swizzle_bytes) {
  return swizzle_bytes == 128 ? 2u : (swizzle_bytes == 64 ? 4u : 6u);
}
// Shared-memory matrix descriptor.
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride-dim byte offset>>4 bits [46,48) version = 1 (sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B)
GDL_DEVINL uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}
// Instruction descriptor for kind::f16: D=f32, A/B = bf16 (fmt 1) or f16 (fmt 0).
//   [4,6) c_format(1=f32) [7,10) a_format [10,13) b_format [15] a_major [16] b_major
//   [17,23) N>>3 [24,29) M>>4.  major: 0 = K-major, 1 = MN-major.
GDL_DEVINL uint32_t umma_idesc(int m, int n, int ab_fmt, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(ab_fmt & 7) << 7;
  d |= (uint32_t)(ab_fmt & 7) << 10;
  d |= (uint32_t)(a_mn_major & 1) << 15;
  d |= (uint32_t)(b_mn_major & 1) << 16;
  d |= (uint32_t)((n >> 3) & 0x3f) << 17;
  d |= (uint32_t)((m >> 4) & 0x1f) << 24;
  return d;
}

// ----------------------------------------------------------------------------------------
// small numeric helpers
// ----------------------------------------------------------------------------------------
GDL_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
GDL_DEVINL uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
GDL_DEVINL float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
GDL_DEVINL float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

GDL_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------------------
// epilogue helpers shared by the tensor-core conv kernels
// ----------------------------------------------------------------------------------------
GDL_DEVINL uint8_t* align_smem_1024(uint8_t* raw) {
  uint32_t a = smem_u32(raw);
  uint32_t pad = (1024u - (a & 1023u)) & 1023u;
  return raw + pad;
}

template <typename T>
GDL_DEVINL void store_row16(T* dst, const float (&f)[16], int nvalid, int vec_ok);

template <>
GDL_DEVINL void store_row16<float>(float* dst, const float (&f)[16], int nvalid, int vec_ok) {
  if (nvalid == 16 && vec_ok) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) d4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) dst[i] = f[i];
  }
}
template <>
GDL_DEVINL void store_row16<__nv_bfloat16>(__nv_bfloat16* dst, const float (&f)[16], int nvalid,
                                           int vec_ok) {
  if (nvalid == 16 && vec_ok) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                       pack_bf16x2(f[6], f[7]));
    d4[1] = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]),
                       pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) dst[i] = __float2bfloat16_rn(f[i]);
  }
}
template <>
GDL_DEVINL void store_row16<__half>(__half* dst, const float (&f)[16], int nvalid, int vec_ok) {
  if (nvalid == 16 && vec_ok) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]),
                       pack_f16x2(f[6], f[7]));
    d4[1] = make_uint4(pack_f16x2(f[8], f[9]), pack_f16x2(f[10], f[11]), pack_f16x2(f[12], f[13]),
                       pack_f16x2(f[14], f[15]));
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < nvalid) dst[i] = __float2half_rn(f[i]);
  }
}

}  // namespace gdl
