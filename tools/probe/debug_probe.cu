// debug_probe.cu — development probe, built by tools/probe_shift.py into its own library (NOT part of libgdlb200.so): does a UMMA shared-memory descriptor whose
// start address is NOT aligned to the 1024-byte swizzle repeat address the rows one expects?  If it does,
// a conv tile loaded once with its halo can serve several filter taps by shifting the descriptor start.
//   mode 0: K-major SW128 A operand [rows = pixels][64 ch]; start = base + shift*128 B
//   mode 1: MN-major SW128 A operand [K rows = pixels][2 atoms x 64 ch]; start = base + shift*128 B
// bo_mode 1 sets the descriptor's base_offset field (bits 49..51) to (start >> 7) & 7.
#include "../../include/gdl_b200.h"
#include "../../geo-deep-learning_b200/csrc/tmap.cuh"

extern "C" int gdl_debug_shift_probe(const void* a, const void* b, float* out, int mode, int shift, int bo_mode, void* stream);

namespace gdl {

struct ProbeParams {
  CUtensorMap tmA;  // mode 0: [144 rows][64] box (64,144); mode 1: [80 rows][128] box (64,80) x2
  CUtensorMap tmB;  // mode 0: [64 rows][64] box (64,64) K-major; mode 1: [80 rows][64] box (64,80) MN-major
  int mode, shift, bo_mode;
  float* out;  // [128][64]
};

__global__ void __launch_bounds__(128, 1) probe_shift_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + 32768;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(&tmem_base_s, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    if (p.mode == 0) {
      mbar_expect_tx(&bar_load, 144 * 128 + 64 * 128);
      tma_load_2d(a_s, &p.tmA, &bar_load, 0, 0);
      tma_load_2d(b_s, &p.tmB, &bar_load, 0, 0);
    } else {
      mbar_expect_tx(&bar_load, 2 * 80 * 128 + 80 * 128);
      tma_load_2d(a_s, &p.tmA, &bar_load, 0, 0);
      tma_load_2d(a_s + 80 * 128, &p.tmA, &bar_load, 64, 0);
      tma_load_2d(b_s, &p.tmB, &bar_load, 0, 0);
    }
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(a_s) + p.shift * 128;
    const uint32_t b0 = smem_u32(b_s);
    const uint64_t bo = p.bo_mode ? (uint64_t)((a0 >> 7) & 7) << 49 : 0;
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t da, db;
      uint32_t idesc;
      if (p.mode == 0) {
        da = umma_smem_desc(a0 + kk * 32, 16, 1024, 2) | bo;
        db = umma_smem_desc(b0 + kk * 32, 16, 1024, 2);
        idesc = umma_idesc(128, 64, 1, 0, 0);
      } else {
        da = umma_smem_desc(a0 + kk * 2048, 80 * 128, 1024, 2) | bo;
        db = umma_smem_desc(b0 + kk * 2048, 16, 1024, 2);
        idesc = umma_idesc(128, 64, 1, 1, 1);
      }
      umma_f16(tmem, da, db, idesc, kk != 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int j = 0; j < 4; ++j) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + ((uint32_t)(warp * 32) << 16) + j * 16, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) p.out[row * 64 + j * 16 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace gdl

using namespace gdl;

// a: mode 0 bf16 [144][64], mode 1 bf16 [80][128]; b: mode 0 bf16 [64][64] (rows = N, K-major), mode 1 bf16 [80][64]
extern "C" int gdl_debug_shift_probe(const void* a, const void* b, float* out, int mode, int shift, int bo_mode,
                                     void* stream) {
  GDL_REQUIRE(a && b && out && (mode == 0 || mode == 1) && shift >= 0 && shift <= 16, GDL_ERR_INVALID, "probe: bad args");
  ProbeParams p;
  p.mode = mode;
  p.shift = shift;
  p.bo_mode = bo_mode;
  p.out = out;
  int st;
  if (mode == 0) {
    st = make_tmap_2d(&p.tmA, a, kDtBF16, 64, 144, 64, 64, 144, 128);
    if (st) return st;
    st = make_tmap_2d(&p.tmB, b, kDtBF16, 64, 64, 64, 64, 64, 128);
  } else {
    st = make_tmap_2d(&p.tmA, a, kDtBF16, 128, 80, 128, 64, 80, 128);
    if (st) return st;
    st = make_tmap_2d(&p.tmB, b, kDtBF16, 64, 80, 64, 64, 80, 128);
  }
  if (st) return st;
  static PerDeviceOnce attr_once;
  GDL_CHECK_CUDA(set_max_dyn_smem_once(attr_once, probe_shift_kernel, 65536));
  probe_shift_kernel<<<1, 128, 49152 + 1024, (cudaStream_t)stream>>>(p);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
