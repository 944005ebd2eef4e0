// igemm_conv.cu — tensor-core implicit-GEMM convolution for sm_100a (forward / dgrad and wgrad).
//
// Forward (gdl_conv2d_nhwc_fwd), per CTA tile of 128 output pixels x BN output channels:
//   D[128 x BN] (fp32, TMEM) = sum over taps (r,s), sources, channel chunks of
//       A[128 pixels x BK channels]  (TMA 4-D box of the NHWC input shifted by the tap; the
//                                     padding halo is TMA out-of-bounds zero fill)
//     x B[BN x BK]                   (TMA 2-D box of the packed weights [Cout][R][S][Ctot])
//   Both operands are K-major in shared memory with the hardware swizzle that matches BK
//   (128B / 64B / 32B for BK = 64 / 32 / 16 channels); tcgen05.mma.kind::f16, M = 128, N = BN.
//   Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM alloc), warps 2..5 = epilogue
//   (tcgen05.ld -> bias/ReLU -> NHWC store).  Persistent grid, double-buffered accumulators.
//
// Wgrad (gdl_conv2d_nhwc_wgrad): D[128 (Cout) x BN (Cin)] += dYᵀ[64 px x 128] · X_tap[64 px x BN]
//   with BOTH operands MN-major (pixels are the contraction dim, channels are contiguous in
//   HBM and in the TMA box); split over pixel ranges, partial sums added with red.global.add.
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "../../include/gdl_b200.h"
#include "tmap.cuh"
#include "det_reduce.cuh"

namespace gdl {

// runtime options (gdl_set_option); env vars GDL_CONV_HALO / GDL_WGRAD_HALO / GDL_WGRAD_L2_MB seed them
static int g_opt_conv_halo = -1, g_opt_wgrad_halo = -1, g_opt_conv_epilogue = -1, g_opt_conv_rows = -1;
int conv3x3_rows_try(const gdl_conv_fwd_t* d, cudaStream_t stream, int* status, int* dry);  // conv3x3_rows.cu
int wgrad3x3_rows_try(const gdl_conv_wgrad_t* d, cudaStream_t stream, int* status);  // wgrad3x3_rows.cu
static int g_opt_wgrad_rows = -1, g_opt_wgrad_sched = -1;
static long long g_opt_wgrad_l2_mb = -1;
static int opt_int(int& slot, const char* env, int dflt) {
  if (slot < 0) {
    const char* e = getenv(env);
    slot = e ? atoi(e) : dflt;
  }
  return slot;
}

constexpr int kMaxStages = 8;
constexpr int kConvThreads = 192;  // 6 warps
constexpr int kSmemBudget = 196 * 1024;  // + 16.5 KB static epilogue staging + barriers < 227 KB
constexpr int kMinSmemRequest = 120 * 1024;  // forces 1 CTA / SM (TMEM is allocated per CTA)

struct ConvFwdKParams {
  CUtensorMap tmA[GDL_MAX_SRC];
  CUtensorMap tmB;
  int num_src;
  int src_chunks[GDL_MAX_SRC];
  int src_coff[GDL_MAX_SRC];
  int Ctot;
  int R, S, pad_h, pad_w;
  int Nimg, Ho, Wo;
  int TH, TW, tiles_w, tiles_h;
  int Cout, BN, n_tiles, num_tiles;
  int BK, stages, a_bytes, b_bytes, stage_bytes, tmem_cols;
  int ab_fmt;
  void* out;
  int out_dtype;
  long long ldo;
  int vec_ok;
  int pair_ok;  // (unused)
  int res_vec_ok;  // fp32 residual rows are 16-byte aligned -> float4 loads
  int epi_mode;    // 0 = direct row stores, 1 = smem-transposed coalesced stores, 2 = smem-staged TMA tile stores
  CUtensorMap tmO; // epi_mode 2: output map, box = (o_slab channels, TW, TH)
  int o_slab;      // channels per staged slab (16 / 32 / 64)
  int o_stage_bytes, o_swz_mask, o_smem_off;  // staging tile bytes, swizzle chunk mask, offset in dynamic smem
  const float* bias;
  int relu;
  const float* oscale;   // optional per-channel multiplier applied to (acc + bias) before the residual (LayerScale)
  const void* residual;  // optional, added in the epilogue (fp32 residual stream or 16-bit)
  int res_dtype;
  long long ldr;
  int groups, tiles_per_group;  // grouped launch (all attention heads): tile = group * tiles_per_group + tile-in-group
  int g_a, g_b, g_o;     // per-group offsets: A channel, B contiguous-dimension element, output channel
  int w_rows_per_img;    // batched B operand: weight row offset per image (attention GEMMs), 0 = shared
  int w_mn_major;        // B operand stored [K rows][N cols] (N contiguous): P.V and dS.K of the attention
  // halo mode (3x3, pad 1, 128x1-pixel tiles, BK = 64): one k-iteration loads ONE input row segment with
  // its 2 halo pixels (130 px) and the weights of the 3 horizontal taps; the 3 taps are 3 MMAs whose A
  // descriptor start is shifted by 0/1/2 rows (128 B) inside the swizzled tile -> A traffic / 3.
  int halo;
  int halo_bo;           // set the descriptor base_offset field for the shifted start (probe-determined)
  int b_slot;            // bytes between the 3 per-tap weight tiles of a stage
  // BatchNorm statistics fused into the TMA-store epilogue (epi_mode 2, 64-channel 16-bit slabs, one n-tile): per-channel
  // sum(x - pivot) / sum((x - pivot)^2) of the ROUNDED output, read back from the staged slab; per-CTA partials -> det_finish
  int bn_on, log2_tw;
  float* bn_sums;
  const float* bn_pivot;
  DetCtx bn_det;
};

__global__ void __launch_bounds__(kConvThreads, 1)
conv_fwd_kernel(const __grid_constant__ ConvFwdKParams p) {
  GDL_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);

  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float stage_buf[4][32 * 36];  // per-epilogue-warp transpose tile (stride 36: 128-bit conflict free)
  __shared__ long long stage_pix[4][32];                  // pixel index of each accumulator row of the current tile

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.R * p.S;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.num_src; ++s) tma_prefetch_desc(&p.tmA[s]);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_smem, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(p.a_bytes + p.b_bytes);
      for (int tile_g = blockIdx.x; tile_g < p.num_tiles; tile_g += gridDim.x) {
        const int grp = tile_g / p.tiles_per_group;
        const int tile = tile_g - grp * p.tiles_per_group;
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int img = m_tile / tiles_per_img;
        const int t_in = m_tile - img * tiles_per_img;
        const int h0 = (t_in / p.tiles_w) * p.TH;
        const int w0 = (t_in % p.tiles_w) * p.TW;
        const int n0 = n_tile * p.BN;
        const int ga = grp * p.g_a, gb = grp * p.g_b;
        if (p.halo) {
          const uint32_t txh = (uint32_t)(130 * p.BK * 2 + 3 * p.b_bytes);
          for (int r = 0; r < 3; ++r) {
            for (int src = 0; src < p.num_src; ++src) {
              for (int ch = 0; ch < p.src_chunks[src]; ++ch) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* a_dst = smem + (size_t)stage * p.stage_bytes;
                uint8_t* b_dst = a_dst + p.a_bytes;
                mbar_expect_tx(&full_bar[stage], txh);
                tma_load_4d(a_dst, &p.tmA[src], &full_bar[stage], ch * p.BK, w0 - 1, h0 + r - 1, img);
                for (int s = 0; s < 3; ++s)
                  tma_load_2d(b_dst + s * p.b_slot, &p.tmB, &full_bar[stage],
                              (r * 3 + s) * p.Ctot + p.src_coff[src] + ch * p.BK, n0);
                if (++stage == p.stages) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
          }
          continue;
        }
        for (int tap = 0; tap < taps; ++tap) {
          const int r = tap / p.S, s = tap - r * p.S;
          for (int src = 0; src < p.num_src; ++src) {
            const int kbase = tap * p.Ctot + p.src_coff[src];
            for (int ch = 0; ch < p.src_chunks[src]; ++ch) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              uint8_t* a_dst = smem + (size_t)stage * p.stage_bytes;
              uint8_t* b_dst = a_dst + p.a_bytes;
              mbar_expect_tx(&full_bar[stage], tx);
              tma_load_4d(a_dst, &p.tmA[src], &full_bar[stage], ch * p.BK + ga, w0 + s - p.pad_w,
                          h0 + r - p.pad_h, img);
              if (p.w_mn_major)
                tma_load_2d(b_dst, &p.tmB, &full_bar[stage], n0 + gb, kbase + ch * p.BK + img * p.w_rows_per_img);
              else
                tma_load_2d(b_dst, &p.tmB, &full_bar[stage], kbase + ch * p.BK + gb, n0 + img * p.w_rows_per_img);
              if (++stage == p.stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(128, p.BN, p.ab_fmt, 0, p.w_mn_major);
      const uint32_t lt = umma_layout_type(p.BK * 2);
      const uint32_t sbo = 8u * p.BK * 2u;
      const int ksteps = p.BK / 16;
      // MN-major B: one atom of BN (<= 64) contiguous channels per K row; 8 K-rows per swizzle group
      const uint32_t ltB = p.w_mn_major ? umma_layout_type(p.BN * 2) : lt;
      const uint32_t sboB = p.w_mn_major ? 8u * p.BN * 2u : sbo;
      const uint32_t kstepB = p.w_mn_major ? 16u * p.BN * 2u : 32u;
      int k_iters = 0;
      for (int src = 0; src < p.num_src; ++src) k_iters += p.src_chunks[src];
      k_iters *= p.halo ? 3 : taps;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint32_t b_addr = a_addr + p.a_bytes;
          if (p.halo) {
            for (int s3 = 0; s3 < 3; ++s3) {
              const uint32_t a_s = a_addr + s3 * p.BK * 2;  // shift by one pixel row (BK*2 bytes)
              const uint64_t bo = p.halo_bo ? ((uint64_t)((a_s >> 7) & 7) << 49) : 0;
              for (int kk = 0; kk < ksteps; ++kk) {
                const uint64_t da = umma_smem_desc(a_s + kk * 32, 16, sbo, lt) | bo;
                const uint64_t db = umma_smem_desc(b_addr + s3 * p.b_slot + kk * 32, 16, sbo, lt);
                umma_f16(d_tmem, da, db, idesc, (uint32_t)((k | kk | s3) != 0));
              }
            }
          } else {
            for (int kk = 0; kk < ksteps; ++kk) {
              const uint64_t da = umma_smem_desc(a_addr + kk * 32, 16, sbo, lt);
              const uint64_t db = umma_smem_desc(b_addr + kk * kstepB, 16, sboB, ltB);
              umma_f16(d_tmem, da, db, idesc, (uint32_t)((k | kk) != 0));
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> smem transpose -> coalesced HBM =====================
    // tcgen05.ld hands every thread one accumulator ROW; storing rows from registers makes each warp store hit
    // 32 different lines (measured: ~4 cycles per 16-byte store, the bound of every tile with K < ~2300).
    // Each warp transposes 32x32 blocks through a private smem tile (row stride 36 floats: 128-bit writes and
    // reads are bank-conflict free) and writes 16-byte vectors such that 4 (16-bit) / 8 (fp32) consecutive lanes
    // cover one pixel's 64 / 128 contiguous bytes; bias / residual / activation are applied on the way out.
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;
    const int th = row / p.TW, tw = row - th * p.TW;
    float* stg = &stage_buf[q][0];
    long long* spix = &stage_pix[q][0];
    const bool out16 = p.out_dtype != GDL_F32;
    const int vw = out16 ? 8 : 4;           // channels per 16-byte output vector
    const int lpr = 32 / vw;                // lanes per pixel row (4 or 8)
    const int rpi = 32 / lpr;               // pixel rows per iteration (8 or 4)
    const int my_r = lane / lpr, my_seg = lane % lpr;
    if (p.epi_mode == 2) {
      // TMA-store variant: each thread writes its pixel row (16-bit) into a swizzled smem slab of 128 px x o_slab
      // channels; one thread hands the slab to the TMA store engine (full-line writes, asynchronous).  Row stores from
      // registers cost ~4 SM-cycles per 16-byte store (32 lines per warp store) and bound every tile with K < ~2300.
      const bool issuer = (warp == 2 && lane == 0);
      uint8_t* smem_o = smem + p.o_smem_off;
      const int esz = p.out_dtype == GDL_F32 ? 4 : 2;
      const int pitch = p.o_slab * esz;
      const int th = row / p.TW, tw = row - th * p.TW;
      int st = 0;
      int it = 0;
      float* bn_acc = &stage_buf[q][0];  // [2][BN] running sums of this warp's rows (the transpose tile is unused in this mode)
      if (p.bn_on) {
        for (int i = lane; i < 2 * p.BN; i += 32) bn_acc[i] = 0.f;
        __syncwarp();
      }
      for (int tile_g = blockIdx.x; tile_g < p.num_tiles; tile_g += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int grp = tile_g / p.tiles_per_group;
        const int tile = tile_g - grp * p.tiles_per_group;
        const int go = grp * p.g_o;
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int img = m_tile / tiles_per_img;
        const int t_in = m_tile - img * tiles_per_img;
        const int h0 = (t_in / p.tiles_w) * p.TH;
        const int w0 = (t_in % p.tiles_w) * p.TW;
        const int n0 = n_tile * p.BN;
        const bool valid = (h0 + th < p.Ho) && (w0 + tw < p.Wo);  // residual rows outside the image are not read
        const long long pix = ((long long)img * p.Ho + h0 + th) * p.Wo + w0 + tw;
        mbar_wait(&tfull_bar[acc], aphase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
        for (int cb0 = 0; cb0 < p.BN; cb0 += p.o_slab) {
          if (issuer) bulk_wait_group_read<1>();  // the store issued two slabs ago no longer reads staging[st]
          named_bar_sync(1, 128);
          uint8_t* stg = smem_o + (size_t)st * p.o_stage_bytes;
          for (int cb = cb0; cb < cb0 + p.o_slab; cb += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_addr + cb, v);
            tmem_ld_wait();
            const int c0 = n0 + cb;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < p.Cout) f[i] += __ldg(p.bias + c0 + i);
            }
            if (p.oscale != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < p.Cout) f[i] *= __ldg(p.oscale + c0 + i);
            }
            if (p.residual != nullptr && valid && c0 < p.Cout) {
              // every thread reads its own (contiguous) residual row segment; Cout % 4 == 0 and 16-byte aligned rows
              const long long roff = pix * p.ldr + c0;
              const int nv = min(16, p.Cout - c0);
              if (p.res_dtype == GDL_F32) {
                const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + roff);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (4 * i < nv) {
                    const float4 r4 = rp[i];
                    f[4 * i] += r4.x; f[4 * i + 1] += r4.y; f[4 * i + 2] += r4.z; f[4 * i + 3] += r4.w;
                  }
              } else if (p.res_dtype == GDL_BF16) {
                const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nv) f[i] += __bfloat162float(rp[i]);
              } else {
                const __half* rp = reinterpret_cast<const __half*>(p.residual) + roff;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nv) f[i] += __half2float(rp[i]);
              }
            }
            if (p.relu == 1) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            } else if (p.relu == 2) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = 0.5f * f[i] * (1.f + erff(f[i] * 0.70710678118654752f));
            }
            uint32_t off = (uint32_t)(row * pitch + (cb - cb0) * esz);
            if (p.out_dtype == GDL_F32) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t o = off ^ (((off >> 7) & (uint32_t)p.o_swz_mask) << 4);
                *reinterpret_cast<float4*>(stg + o) = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                off += 16;
              }
            } else {
              uint4 lo, hi;
              if (p.out_dtype == GDL_BF16) {
                lo = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                pack_bf16x2(f[6], f[7]));
                hi = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]),
                                pack_bf16x2(f[14], f[15]));
              } else {
                lo = make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]), pack_f16x2(f[4], f[5]),
                                pack_f16x2(f[6], f[7]));
                hi = make_uint4(pack_f16x2(f[8], f[9]), pack_f16x2(f[10], f[11]), pack_f16x2(f[12], f[13]),
                                pack_f16x2(f[14], f[15]));
              }
              const uint32_t o0 = off ^ (((off >> 7) & (uint32_t)p.o_swz_mask) << 4);
              off += 16;
              const uint32_t o1 = off ^ (((off >> 7) & (uint32_t)p.o_swz_mask) << 4);
              *reinterpret_cast<uint4*>(stg + o0) = lo;
              *reinterpret_cast<uint4*>(stg + o1) = hi;
            }
          }
          if (cb0 + p.o_slab >= p.BN) {  // all TMEM reads of this tile done: release the accumulator
            tc_fence_before();
            mbar_arrive(&tempty_bar[acc]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (issuer && n0 + cb0 < p.Cout) {
            tma_store_4d(&p.tmO, stg, n0 + cb0 + go, w0, h0, img);
            bulk_commit_group();
          }
          if (p.bn_on && cb0 + 2 * lane < p.Cout) {
            // the slab (128 pixels x 64 channels, 16-bit, swizzled) is complete in shared memory: lane = one channel pair
            // (one 32-bit word per row: conflict free), warp q = rows q, q+4, ...; rows outside the image are skipped
            const int c = cb0 + 2 * lane;
            const float pv0 = p.bn_pivot ? __ldg(p.bn_pivot + c) : 0.f;
            const float pv1 = (p.bn_pivot && c + 1 < p.Cout) ? __ldg(p.bn_pivot + c + 1) : 0.f;
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
            // slab rows are 128 bytes (64 channels x 2 B): row r starts at r * 128 and its 16-byte chunks are XOR-ed with
            // (r & 7); this warp's rows r = q + 4k alternate between two chunk masks
            const bool full = (h0 + p.TH <= p.Ho) && (w0 + p.TW <= p.Wo);
            const uint32_t xo[2] = {(uint32_t)(lane * 4) ^ ((uint32_t)q << 4), (uint32_t)(lane * 4) ^ ((uint32_t)(q + 4) << 4)};
            const bool is_bf16 = p.out_dtype == GDL_BF16;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
              const int r = q + 4 * k;
              const uint32_t u = *reinterpret_cast<const uint32_t*>(stg + r * 128 + xo[k & 1]);
              float x0, x1;
              if (is_bf16) {
                x0 = bf16_lo(u);
                x1 = bf16_hi(u);
              } else {
                const __half2 h2 = *reinterpret_cast<const __half2*>(&u);
                x0 = __low2float(h2);
                x1 = __high2float(h2);
              }
              if (full || ((h0 + (r >> p.log2_tw) < p.Ho) && (w0 + (r & (p.TW - 1)) < p.Wo))) {
                x0 -= pv0;
                x1 -= pv1;
                s1a += x0;
                s1b += x1;
                s2a = fmaf(x0, x0, s2a);
                s2b = fmaf(x1, x1, s2b);
              }
            }
            bn_acc[c] += s1a;
            bn_acc[c + 1] += s1b;
            bn_acc[p.BN + c] += s2a;
            bn_acc[p.BN + c + 1] += s2b;
          }
          st ^= 1;
        }
      }
      if (issuer) bulk_wait_group<0>();
      if (p.bn_on) {
        // this CTA's partial = the 4 warps' sums in warp order -> its slot; det_finish (below) adds the CTAs in order
        named_bar_sync(1, 128);
        for (int i = (int)threadIdx.x - 64; i < 2 * p.BN; i += 128) {
          const int c = i < p.BN ? i : i - p.BN;
          if (c < p.Cout) {
            const float v = ((stage_buf[0][i] + stage_buf[1][i]) + stage_buf[2][i]) + stage_buf[3][i];
            det_put(p.bn_det, 2 * p.Cout, i < p.BN ? c : p.Cout + c, v);
          }
        }
      }
    } else if (p.epi_mode == 0) {
      // direct variant: every thread stores its own accumulator row (16-byte vectors, 32 lines per warp store)
      int it = 0;
      for (int tile_g = blockIdx.x; tile_g < p.num_tiles; tile_g += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        const int grp = tile_g / p.tiles_per_group;
        const int tile = tile_g - grp * p.tiles_per_group;
        const int go = grp * p.g_o;
        const int n_tile = tile % p.n_tiles;
        const int m_tile = tile / p.n_tiles;
        const int img = m_tile / tiles_per_img;
        const int t_in = m_tile - img * tiles_per_img;
        const int h = (t_in / p.tiles_w) * p.TH + th;
        const int w = (t_in % p.tiles_w) * p.TW + tw;
        const int n0 = n_tile * p.BN;
        const bool valid = (h < p.Ho) && (w < p.Wo);
        const long long pix = ((long long)img * p.Ho + h) * p.Wo + w;
        mbar_wait(&tfull_bar[acc], aphase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
        for (int j = 0; j < p.BN / 16; ++j) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_addr + j * 16, v);
          tmem_ld_wait();
          const int c0 = n0 + j * 16;
          const int nvalid = min(16, p.Cout - c0);
          if (valid && nvalid > 0) {
            float f[16];
  #pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (p.bias != nullptr) {
  #pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nvalid) f[i] += __ldg(p.bias + c0 + i);
            }
            if (p.oscale != nullptr) {
  #pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < nvalid) f[i] *= __ldg(p.oscale + c0 + i);
            }
            if (p.residual != nullptr) {
              const long long roff = pix * p.ldr + c0;
              if (p.res_dtype == GDL_F32) {
                const float* r = reinterpret_cast<const float*>(p.residual) + roff;
  #pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nvalid) f[i] += r[i];
              } else if (p.res_dtype == GDL_BF16) {
                const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff;
  #pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nvalid) f[i] += __bfloat162float(r[i]);
              } else {
                const __half* r = reinterpret_cast<const __half*>(p.residual) + roff;
  #pragma unroll
                for (int i = 0; i < 16; ++i)
                  if (i < nvalid) f[i] += __half2float(r[i]);
              }
            }
            if (p.relu == 1) {
  #pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.f);
            } else if (p.relu == 2) {  // exact (erf) GELU, nn.GELU default
  #pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = 0.5f * f[i] * (1.f + erff(f[i] * 0.70710678118654752f));
            }
            const long long off = pix * p.ldo + c0 + go;
            if (p.out_dtype == GDL_F32)
              store_row16<float>(reinterpret_cast<float*>(p.out) + off, f, nvalid, p.vec_ok);
            else if (p.out_dtype == GDL_BF16)
              store_row16<__nv_bfloat16>(reinterpret_cast<__nv_bfloat16*>(p.out) + off, f, nvalid,
                                         p.vec_ok);
            else
              store_row16<__half>(reinterpret_cast<__half*>(p.out) + off, f, nvalid, p.vec_ok);
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
      }
    } else {
    int it = 0;
    for (int tile_g = blockIdx.x; tile_g < p.num_tiles; tile_g += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int grp = tile_g / p.tiles_per_group;
      const int tile = tile_g - grp * p.tiles_per_group;
      const int go = grp * p.g_o;
      const int n_tile = tile % p.n_tiles;
      const int m_tile = tile / p.n_tiles;
      const int img = m_tile / tiles_per_img;
      const int t_in = m_tile - img * tiles_per_img;
      const int h = (t_in / p.tiles_w) * p.TH + th;
      const int w = (t_in % p.tiles_w) * p.TW + tw;
      const int n0 = n_tile * p.BN;
      const bool valid = (h < p.Ho) && (w < p.Wo);
      __syncwarp();
      spix[lane] = valid ? ((long long)img * p.Ho + h) * p.Wo + w : -1;  // pixel index of row `lane`, -1 = outside
      mbar_wait(&tfull_bar[acc], aphase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.BN);
      for (int cb = 0; cb < p.BN; cb += 32) {
        const int ccols = min(32, p.BN - cb);  // 16 or 32
        uint32_t v0[16], v1[16];
        tmem_ld_32x32b_x16(t_addr + cb, v0);
        if (ccols > 16) tmem_ld_32x32b_x16(t_addr + cb + 16, v1);
        tmem_ld_wait();
        __syncwarp();
        float4* srow = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          srow[i] = make_float4(__uint_as_float(v0[4 * i]), __uint_as_float(v0[4 * i + 1]),
                                __uint_as_float(v0[4 * i + 2]), __uint_as_float(v0[4 * i + 3]));
        if (ccols > 16) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            srow[4 + i] = make_float4(__uint_as_float(v1[4 * i]), __uint_as_float(v1[4 * i + 1]),
                                      __uint_as_float(v1[4 * i + 2]), __uint_as_float(v1[4 * i + 3]));
        }
        __syncwarp();
        const int c0 = n0 + cb;
        const int cseg = my_seg * vw;            // first channel of this lane's vector inside the chunk
        const int col = c0 + cseg;
        const bool seg_in = cseg < ccols;
        const int nval = seg_in ? min(vw, p.Cout - col) : 0;  // valid channels of this lane's vector
        float bv[8], sv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bv[j] = (p.bias != nullptr && j < nval) ? __ldg(p.bias + col + j) : 0.f;
          sv[j] = (p.oscale != nullptr && j < nval) ? __ldg(p.oscale + col + j) : 1.f;
        }
#pragma unroll 4
        for (int r0 = 0; r0 < 32; r0 += rpi) {
          const int rr = r0 + my_r;
          const long long pix = spix[rr];
          if (pix < 0 || nval <= 0) continue;
          float f[8];
          const float4* sp = reinterpret_cast<const float4*>(stg + rr * 36 + cseg);
          const float4 a4 = sp[0];
          f[0] = a4.x; f[1] = a4.y; f[2] = a4.z; f[3] = a4.w;
          if (out16) {
            const float4 b4 = sp[1];
            f[4] = b4.x; f[5] = b4.y; f[6] = b4.z; f[7] = b4.w;
          } else {
            f[4] = f[5] = f[6] = f[7] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = (f[j] + bv[j]) * sv[j];
          if (p.residual != nullptr) {
            const long long ro = pix * p.ldr + col;
            if (p.res_dtype == GDL_F32) {
              const float* rp = reinterpret_cast<const float*>(p.residual) + ro;
              if (p.res_vec_ok && nval == vw) {
                const float4 r4 = *reinterpret_cast<const float4*>(rp);
                f[0] += r4.x; f[1] += r4.y; f[2] += r4.z; f[3] += r4.w;
                if (out16) {
                  const float4 s4 = *reinterpret_cast<const float4*>(rp + 4);
                  f[4] += s4.x; f[5] += s4.y; f[6] += s4.z; f[7] += s4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (j < nval) f[j] += rp[j];
              }
            } else if (p.res_dtype == GDL_BF16) {
              const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + ro;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nval) f[j] += __bfloat162float(rp[j]);
            } else {
              const __half* rp = reinterpret_cast<const __half*>(p.residual) + ro;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nval) f[j] += __half2float(rp[j]);
            }
          }
          if (p.relu == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
          } else if (p.relu == 2) {  // exact (erf) GELU, nn.GELU default
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.5f * f[j] * (1.f + erff(f[j] * 0.70710678118654752f));
          }
          const long long off = pix * p.ldo + col + go;
          if (p.out_dtype == GDL_F32) {
            float* o = reinterpret_cast<float*>(p.out) + off;
            if (p.vec_ok && nval == 4) *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
            else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < nval) o[j] = f[j];
            }
          } else if (p.out_dtype == GDL_BF16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
            if (p.vec_ok && nval == 8)
              *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                        pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nval) o[j] = __float2bfloat16_rn(f[j]);
            }
          } else {
            __half* o = reinterpret_cast<__half*>(p.out) + off;
            if (p.vec_ok && nval == 8)
              *reinterpret_cast<uint4*>(o) = make_uint4(pack_f16x2(f[0], f[1]), pack_f16x2(f[2], f[3]),
                                                        pack_f16x2(f[4], f[5]), pack_f16x2(f[6], f[7]));
            else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nval) o[j] = __float2half_rn(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
    }  // epi_mode
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (p.bn_on) det_finish(p.bn_det, 2 * p.Cout, p.bn_sums);
}

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
static int pow2_ge(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

// choose TH x TW = npix (power of two) minimising the padded area of an H x W image
static void choose_tile(int H, int W, int npix, int* TH, int* TW) {
  long long best = -1;
  int bth = 1, btw = npix;
  for (int tw = npix; tw >= 1; tw >>= 1) {
    int th = npix / tw;
    long long cost = (long long)((W + tw - 1) / tw) * tw * ((H + th - 1) / th) * th;
    if (best < 0 || cost < best) {
      best = cost;
      bth = th;
      btw = tw;
    }
  }
  *TH = bth;
  *TW = btw;
}

static int chunk_width(const gdl_src_t* src, int n) {
  int bk = 64;
  for (int i = 0; i < n; ++i) {
    while (bk > 16 && (src[i].channels % bk) != 0) bk >>= 1;
  }
  return bk;
}

static int sm_count() { return device_sm_count(); }

static int validate_srcs(int num_src, const gdl_src_t* src, int* ctot) {
  GDL_REQUIRE(num_src >= 1 && num_src <= GDL_MAX_SRC, GDL_ERR_INVALID,
              "num_src must be in [1,%d] (got %d)", GDL_MAX_SRC, num_src);
  int c = 0;
  for (int i = 0; i < num_src; ++i) {
    GDL_REQUIRE(src[i].ptr != nullptr, GDL_ERR_INVALID, "source %d: null pointer", i);
    GDL_REQUIRE(src[i].channels > 0 && src[i].channels % 16 == 0, GDL_ERR_INVALID,
                "source %d: channels must be a positive multiple of 16 (got %d)", i, src[i].channels);
    GDL_REQUIRE(src[i].ld >= src[i].channels && src[i].ld % 8 == 0, GDL_ERR_INVALID,
                "source %d: pixel stride %d invalid for %d channels", i, src[i].ld, src[i].channels);
    c += src[i].channels;
  }
  *ctot = c;
  return 0;
}

extern int g_opt_sra_max_ctas;  // sra_attention.cu
extern int g_opt_deterministic;  // runtime.cu
}  // namespace gdl

using namespace gdl;

extern "C" int gdl_set_option(const char* name, long long value) {
  GDL_REQUIRE(name != nullptr, GDL_ERR_INVALID, "set_option: null name");
  if (!strcmp(name, "conv_halo")) g_opt_conv_halo = (int)value;
  else if (!strcmp(name, "wgrad_halo")) g_opt_wgrad_halo = (int)value;
  else if (!strcmp(name, "conv_epilogue")) g_opt_conv_epilogue = (int)value;
  else if (!strcmp(name, "wgrad_l2_mb")) g_opt_wgrad_l2_mb = value;
  else if (!strcmp(name, "conv_rows")) g_opt_conv_rows = (int)value;
  else if (!strcmp(name, "wgrad_rows")) g_opt_wgrad_rows = (int)value;
  else if (!strcmp(name, "wgrad_sched")) g_opt_wgrad_sched = (int)value;  // 1: few long units (one epilogue per CTA); 0: round-1 rule
  else if (!strcmp(name, "deterministic")) g_opt_deterministic = (int)value;  // 1 (default): ordered reductions when a workspace is registered
  else if (!strcmp(name, "sra_max_ctas")) g_opt_sra_max_ctas = (int)value;  // 0 = one CTA per SM (tests: fewer, longer CTAs)
  else if (!strcmp(name, "pdl")) g_opt_pdl = value != 0;  // programmatic dependent launch of every kernel (common.cuh)
  else {
    set_last_error("set_option: unknown option '%s'", name);
    return GDL_ERR_INVALID;
  }
  return 0;
}

// dry != nullptr: validate and plan only (nothing is launched); *dry = 1 when the BatchNorm statistics (d->bn_sums) would be
// produced by the conv's own epilogue, 0 when the statistics kernel would run after it
static int conv_fwd_impl(const gdl_conv_fwd_t* d, void* stream_, int* dry) {
  GDL_REQUIRE(d != nullptr, GDL_ERR_INVALID, "null descriptor");
  cudaStream_t stream = (cudaStream_t)stream_;
  int Ctot = 0;
  int st = validate_srcs(d->num_src, d->src, &Ctot);
  if (st) return st;
  GDL_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, GDL_ERR_INVALID,
              "bad conv shape N=%d H=%d W=%d Cout=%d", d->N, d->H, d->W, d->Cout);
  GDL_REQUIRE(d->R >= 1 && d->S >= 1 && d->R <= 15 && d->S <= 15 && d->pad_h >= 0 && d->pad_w >= 0,
              GDL_ERR_INVALID, "bad filter R=%d S=%d pad=%d,%d", d->R, d->S, d->pad_h, d->pad_w);
  GDL_REQUIRE(d->dtype == GDL_BF16 || d->dtype == GDL_F16, GDL_ERR_INVALID, "operand dtype must be bf16/f16");
  GDL_REQUIRE(d->out_dtype == GDL_BF16 || d->out_dtype == GDL_F16 || d->out_dtype == GDL_F32,
              GDL_ERR_INVALID, "bad out_dtype %d", d->out_dtype);
  GDL_REQUIRE(d->weight && d->out, GDL_ERR_INVALID, "null weight/out");
  const int Ho = d->H + 2 * d->pad_h - d->R + 1;
  const int Wo = d->W + 2 * d->pad_w - d->S + 1;
  GDL_REQUIRE(Ho > 0 && Wo > 0, GDL_ERR_INVALID, "empty output %dx%d", Ho, Wo);
  GDL_REQUIRE(d->ldo >= d->Cout, GDL_ERR_INVALID, "ldo %d < Cout %d", d->ldo, d->Cout);

  GDL_REQUIRE(d->residual == nullptr || d->ldr >= d->Cout, GDL_ERR_INVALID, "residual stride %d < Cout", d->ldr);
  if (opt_int(g_opt_conv_rows, "GDL_CONV_ROWS", 1) > 0) {
    // narrow 3x3 convs: weight-stationary row-rolling kernel (conv3x3_rows.cu)
    int st_rows = 0;
    if (conv3x3_rows_try(d, stream, &st_rows, dry)) return st_rows;
  }

  ConvFwdKParams p;
  memset(&p, 0, sizeof(p));
  p.num_src = d->num_src;
  p.Ctot = Ctot;
  p.R = d->R;
  p.S = d->S;
  p.pad_h = d->pad_h;
  p.pad_w = d->pad_w;
  p.BK = chunk_width(d->src, d->num_src);
  // geometry: a pointwise conv over NHWC is a flat GEMM over N*H*W pixels
  int N = d->N, H = d->H, W = d->W, oH = Ho, oW = Wo;
  if (d->R == 1 && d->S == 1 && d->pad_h == 0 && d->pad_w == 0 && d->w_rows_per_img == 0) {
    long long M = (long long)N * H * W;
    GDL_REQUIRE(M < (1ll << 31), GDL_ERR_UNSUPPORTED, "too many pixels");
    W = (int)M;
    H = 1;
    N = 1;
    oH = 1;
    oW = W;
  }
  if (d->w_mn_major) {
    GDL_REQUIRE(d->R == 1 && d->S == 1 && d->num_src == 1, GDL_ERR_INVALID, "w_mn_major: pointwise, single source only");
    GDL_REQUIRE((d->Cout == 16 || d->Cout == 32 || d->Cout == 64), GDL_ERR_UNSUPPORTED,
                "w_mn_major: Cout must be 16, 32 or 64 (got %d)", d->Cout);
  }
  p.Nimg = N;
  p.Ho = oH;
  p.Wo = oW;
  choose_tile(oH, oW, 128, &p.TH, &p.TW);
  p.tiles_w = (oW + p.TW - 1) / p.TW;
  p.tiles_h = (oH + p.TH - 1) / p.TH;
  p.Cout = d->Cout;
  p.n_tiles = (d->Cout + 255) / 256;
  p.BN = (((d->Cout + p.n_tiles - 1) / p.n_tiles) + 15) / 16 * 16;
  long long num_tiles = (long long)N * p.tiles_w * p.tiles_h * p.n_tiles;
  const int G = d->groups > 1 ? d->groups : 1;
  if (G > 1) {
    GDL_REQUIRE(d->R == 1 && d->S == 1 && d->num_src == 1 && !d->bias && !d->oscale && !d->residual, GDL_ERR_INVALID,
                "grouped launch: pointwise, one source, no bias / oscale / residual");
    GDL_REQUIRE(d->g_src_stride % 8 == 0 && d->g_w_stride % 8 == 0 && d->g_out_stride > 0, GDL_ERR_INVALID,
                "grouped launch: strides must be multiples of 8 elements");
  }
  p.groups = G;
  p.tiles_per_group = (int)num_tiles;
  p.g_a = G > 1 ? d->g_src_stride : 0;
  p.g_b = G > 1 ? d->g_w_stride : 0;
  p.g_o = G > 1 ? d->g_out_stride : 0;
  num_tiles *= G;
  GDL_REQUIRE(num_tiles < (1ll << 31), GDL_ERR_UNSUPPORTED, "too many tiles");
  p.num_tiles = (int)num_tiles;
  p.a_bytes = 128 * p.BK * 2;
  p.b_bytes = p.BN * p.BK * 2;
  p.stage_bytes = p.a_bytes + ((p.b_bytes + 1023) / 1024) * 1024;
  {
    // conv_halo: 0 = off, 1 = on.  (The tools/probe_shift.py probe showed that a descriptor start shifted
    // by whole rows inside the 1024-byte swizzle repeat addresses the expected rows with base_offset = 0.)
    const int halo_cfg = opt_int(g_opt_conv_halo, "GDL_CONV_HALO", 1);
    if (halo_cfg > 0 && d->R == 3 && d->S == 3 && d->pad_h == 1 && d->pad_w == 1 && p.TH == 1 && p.TW == 128 &&
        p.BN <= 128 && !d->w_mn_major && d->w_rows_per_img == 0) {
      p.halo = 1;
      p.halo_bo = 0;
      p.a_bytes = ((130 * p.BK * 2 + 1023) / 1024) * 1024;
      p.b_slot = ((p.b_bytes + 1023) / 1024) * 1024;
      p.stage_bytes = p.a_bytes + 3 * p.b_slot;
    }
  }
  p.epi_mode = opt_int(g_opt_conv_epilogue, "GDL_CONV_EPILOGUE", 2);  // 0 direct, 1 smem transpose, 2 TMA store (default; falls back to 0 when not applicable)
  int smem_budget = kSmemBudget;
  if (p.epi_mode == 2) {
    const int esz_o = d->out_dtype == GDL_F32 ? 4 : 2;
    const bool aligned = ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) && ((d->ldo * esz_o) % 16 == 0);
    bool res_ok = true;
    if (d->residual != nullptr) {  // rows are read as float4 / 16-bit scalars by the thread that owns the pixel
      const int esz_r = d->res_dtype == GDL_F32 ? 4 : 2;
      res_ok = ((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0) && ((d->ldr * esz_r) % 16 == 0);
    }
    // staged slab rows are 128 / 64 / 32 bytes (hardware swizzle span): 64 / 32 / 16 channels at 2 B, 32 / 16 at 4 B
    int slab = p.BN % 64 == 0 ? 64 : (p.BN % 32 == 0 ? 32 : 16);
    if (esz_o == 4 && slab == 64) slab = 32;
    const int row_bytes = slab * esz_o;
    // grouped: the output map spans all groups, so a slab must not overhang its group's Cout channels
    const bool grp_ok = G == 1 || ((d->g_out_stride * esz_o) % 16 == 0 && d->Cout % slab == 0);
    if (aligned && res_ok && grp_ok && d->Cout % (16 / esz_o) == 0 && d->Cout % 4 == 0) {
      p.o_slab = slab;
      p.o_stage_bytes = 128 * row_bytes;
      p.o_swz_mask = row_bytes == 128 ? 7 : (row_bytes == 64 ? 3 : 1);
      smem_budget -= 2 * p.o_stage_bytes;
    } else {
      p.epi_mode = 0;
    }
  }
  p.stages = smem_budget / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  GDL_REQUIRE(p.stages >= 2, GDL_ERR_UNSUPPORTED, "tile does not fit shared memory");
  p.o_smem_off = p.stages * p.stage_bytes;  // 1024-aligned: stage_bytes is a multiple of 1024
  p.tmem_cols = pow2_ge(2 * p.BN);
  p.ab_fmt = d->dtype == GDL_BF16 ? 1 : 0;
  p.out = d->out;
  p.out_dtype = d->out_dtype;
  p.ldo = d->ldo;
  const int esz = d->out_dtype == GDL_F32 ? 4 : 2;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) && ((d->ldo * esz) % 16 == 0) &&
             (G == 1 || (d->g_out_stride * esz) % 16 == 0);
  p.pair_ok = 0;
  p.res_vec_ok = d->residual != nullptr && d->res_dtype == GDL_F32 &&
                 ((reinterpret_cast<uintptr_t>(d->residual) & 15) == 0) && (d->ldr % 4 == 0);
  p.bias = d->bias;
  p.relu = d->relu;
  p.oscale = d->oscale;
  p.residual = d->residual;
  p.res_dtype = d->res_dtype;
  p.ldr = d->ldr;
  p.w_rows_per_img = d->w_rows_per_img;
  p.w_mn_major = d->w_mn_major;
  GDL_REQUIRE(d->residual == nullptr || d->ldr >= d->Cout, GDL_ERR_INVALID, "residual stride %d < Cout", d->ldr);

  int coff = 0;
  for (int i = 0; i < d->num_src; ++i) {
    p.src_chunks[i] = d->src[i].channels / p.BK;
    p.src_coff[i] = coff;
    coff += d->src[i].channels;
    st = make_tmap_nhwc(&p.tmA[i], d->src[i].ptr, d->dtype, d->src[i].channels + (long long)(G - 1) * p.g_a, W, H, N,
                        d->src[i].ld, p.BK, p.halo ? 130 : p.TW, p.TH, p.BK * 2);
    if (st) return st;
  }
  const long long Ktot = (long long)d->R * d->S * Ctot;
  const long long w_ld = d->w_ld > 0 ? d->w_ld : (d->w_mn_major ? d->Cout : Ktot);
  if (d->w_mn_major) {
    // weight matrix [w_rows (K index, all images)][Cout] with row stride w_ld: box = (Cout cols, BK rows)
    const long long w_rows = d->w_rows > 0 ? d->w_rows : Ktot;
    st = make_tmap_2d(&p.tmB, d->weight, d->dtype, d->Cout + (long long)(G - 1) * p.g_b, w_rows, w_ld, p.BN, p.BK,
                      p.BN * 2);
  } else {
    const long long w_rows = d->w_rows > 0 ? d->w_rows : d->Cout;
    st = make_tmap_2d(&p.tmB, d->weight, d->dtype, Ktot + (long long)(G - 1) * p.g_b, w_rows, w_ld, p.BK, p.BN,
                      p.BK * 2);
  }
  if (st) return st;

  if (p.epi_mode == 2) {
    st = make_tmap_nhwc(&p.tmO, d->out, d->out_dtype, d->Cout + (long long)(G - 1) * p.g_o, oW, oH, N, d->ldo, p.o_slab,
                        p.TW, p.TH,
                        p.o_slab * (d->out_dtype == GDL_F32 ? 4 : 2));
    if (st) return st;
  }
  int smem = p.stages * p.stage_bytes + 2 * p.o_stage_bytes + 1024;
  if (smem < kMinSmemRequest) smem = kMinSmemRequest;
  static PerDeviceOnce attr_once;
  GDL_CHECK_CUDA(set_max_dyn_smem_once(attr_once, conv_fwd_kernel, kSmemBudget + 4096));
  int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  bool bn_fused = false;
  // Fuse only where the epilogue hides behind the MMAs of the next tile: the statistics triple the epilogue's instruction
  // count, and on short-K GEMMs (1x1 convs with K = 64..256) that cost as much as the separate pass (measured, run 8:
  // 64 -> 256 pointwise at 128^2: +0.13 ms per launch either way; 256 -> 256 3x3: +0.02 ms fused vs 0.13 ms separate).
  static int bn_min_k = -1;
  if (bn_min_k < 0) {
    const char* e = getenv("GDL_BN_FUSE_MIN_K");
    bn_min_k = e ? atoi(e) : 512;
  }
  if (d->bn_sums != nullptr && p.epi_mode == 2 && p.o_slab == 64 && d->out_dtype != GDL_F32 && p.n_tiles == 1 && G == 1 &&
      (long long)d->R * d->S * Ctot >= bn_min_k) {
    const DetWs ws = det_workspace();
    if (det_grid(ws, grid, 2 * d->Cout) == grid) {
      bn_fused = true;
      p.bn_on = 1;
      p.log2_tw = 0;
      while ((1 << p.log2_tw) < p.TW) ++p.log2_tw;
      p.bn_sums = d->bn_sums;
      p.bn_pivot = d->bn_pivot;
      p.bn_det = det_ctx(ws, grid, 2 * d->Cout);
      if (dry == nullptr) GDL_CHECK_CUDA(cudaMemsetAsync(d->bn_sums, 0, 2 * (size_t)d->Cout * sizeof(float), stream));
    }
  }
  if (dry != nullptr) {
    *dry = bn_fused ? 1 : 0;
    return 0;
  }
  GDL_LAUNCH(conv_fwd_kernel, grid, kConvThreads, smem, stream, p);
  GDL_CHECK_CUDA(cudaGetLastError());
  if (d->bn_sums != nullptr && !bn_fused)  // shapes the epilogue cannot cover: the statistics kernel on the stored output
    return gdl_bn_stats(d->out, d->out_dtype, (long long)N * oH * oW, d->Cout, d->ldo, d->bn_sums, d->bn_pivot, stream_);
  return 0;
}

extern "C" int gdl_conv2d_nhwc_fwd(const gdl_conv_fwd_t* d, void* stream_) { return conv_fwd_impl(d, stream_, nullptr); }

extern "C" int gdl_conv2d_bn_fusable(const gdl_conv_fwd_t* d, int* fused) {
  GDL_REQUIRE(fused != nullptr, GDL_ERR_INVALID, "conv2d_bn_fusable: null result pointer");
  *fused = 0;
  if (d == nullptr || d->bn_sums == nullptr) return 0;
  return conv_fwd_impl(d, nullptr, fused);
}

// ==========================================================================================
// wgrad
// ==========================================================================================
namespace gdl {

constexpr int kMaxNTiles = 48;
constexpr int kWgPix = 64;  // pixels (contraction length) per pipeline stage

struct ConvWgradKParams {
  CUtensorMap tmDY;
  CUtensorMap tmX[GDL_MAX_SRC];
  int Ctot, Cout;
  int R, S, pad_h, pad_w;
  int TH, TW, tiles_w, tiles_h;
  int pix_blocks;  // N * tiles_h * tiles_w
  int ksplit, pb_per_split;
  int m_tiles, n_ntiles, num_units;
  int caA, caB;  // channel atom widths (elements) of dY and X boxes: 16/32/64
  int nt_src[kMaxNTiles], nt_c0[kMaxNTiles], nt_w[kMaxNTiles], nt_coff[kMaxNTiles];
  int stages, a_bytes, stage_bytes, tmem_cols, bn_max;
  int ab_fmt;
  float* dw;
  long long dw_ld;          // row stride of dw (elements)
  long long dw_img_stride;  // batched: dw of image i starts at dw + i * dw_img_stride
  int batched;              // one independent dW per image (attention: dV = P^T dO, dK = dS^T Q)
  int pb_per_img;
  // grouped (all attention heads in one launch): unit = group * upg + unit-in-group; group g reads dY / X channels shifted by
  // g * g_dy / g * g_x and adds into dw + g * g_dw
  int upg, g_x, g_dy;
  long long g_dw;
  // halo mode (3x3, pad 1, 64x1-pixel blocks, 64-channel X atoms): a unit owns one filter ROW r and keeps
  // 3 accumulators (s = 0,1,2); each k-iteration loads the X row segment once with its 2 halo pixels
  // (66 px) and the 3 horizontal taps are MMAs whose B descriptor start is shifted by s rows (128 B).
  int halo;
  int nsub;        // accumulators per unit (1, or 3 in halo mode)
  int nacc;        // TMEM accumulator sets in flight (2 = double buffered, 1 = single)
  int unit_taps;   // units along the tap axis (R*S, or 3 filter rows in halo mode)
  int b_atom_bytes;  // bytes of one X atom in a stage (64 px, or 66 px padded to 9 KiB in halo mode)
  // ordered pixel-split accumulation (det_reduce.cuh): with `partials` every unit stores its accumulator tile
  // ([nsub][128 rows][bn_max] floats at partials + unit * tile_floats) instead of adding it to dw with fp32 atomics;
  // wgrad_reduce_kernel then adds the ksplit partial tiles of each output tile in split order.
  float* partials;      // null: red.global.add in arrival order
  long long tile_floats;
  int Nimg;             // batched: images
};

__global__ void __launch_bounds__(kConvThreads, 1)
conv_wgrad_kernel(const __grid_constant__ ConvWgradKParams p) {
  GDL_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);

  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.unit_taps;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  // A-operand atoms beyond Cout are never loaded: they must read as zero.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = p.stages * p.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDY);
    tma_prefetch_desc(&p.tmX[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < p.stages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_smem, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int atomA_bytes = kWgPix * p.caA * 2;  // one MN atom (caA channels) x 64 pixels
  const int atomB_bytes = p.b_atom_bytes;
  const uint32_t atomB_tx = (uint32_t)((p.halo ? kWgPix + 2 : kWgPix) * p.caB * 2);
  const uint32_t acc_stride = (uint32_t)(p.nsub * p.bn_max);

  // unit -> (tap, n-tile, m-tile, k-split); tap fastest so co-resident CTAs share dY / X in L2
  auto decode = [&](int u, int& tap, int& nt, int& mt, int& ks, int& grp) {
    grp = u / p.upg;
    u -= grp * p.upg;
    tap = u % taps;
    u /= taps;
    nt = u % p.n_ntiles;
    u /= p.n_ntiles;
    mt = u % p.m_tiles;
    ks = u / p.m_tiles;
  };
  // pixel-block range of split `ks`; in batched mode ks = img * ksplit + split-within-image
  auto pb_range = [&](int ks, int& pb0, int& pb1, int& img) {
    if (p.batched) {
      img = ks / p.ksplit;
      const int sp = ks - img * p.ksplit;
      pb0 = img * p.pb_per_img + sp * p.pb_per_split;
      pb1 = min((img + 1) * p.pb_per_img, pb0 + p.pb_per_split);
    } else {
      img = 0;
      pb0 = ks * p.pb_per_split;
      pb1 = min(p.pix_blocks, pb0 + p.pb_per_split);
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x) {
        int tap, nt, mt, ks, grp;
        decode(u, tap, nt, mt, ks, grp);
        const int r = p.halo ? tap : tap / p.S;
        const int s = p.halo ? 0 : tap - r * p.S;  // halo: the box starts one pixel left (s = 0) and is 66 wide
        const int m0 = mt * 128;
        const int a_atoms = min(128, p.Cout - m0) / p.caA;  // Cout % caA == 0
        const int b_atoms = p.nt_w[nt] / p.caB;
        const int src = p.nt_src[nt];
        const int c0 = p.nt_c0[nt];
        const uint32_t tx = (uint32_t)(a_atoms * atomA_bytes) + (uint32_t)b_atoms * atomB_tx;
        int pb0, pb1, uimg;
        pb_range(ks, pb0, pb1, uimg);
        for (int pb = pb0; pb < pb1; ++pb) {
          const int img = pb / tiles_per_img;
          const int t_in = pb - img * tiles_per_img;
          const int h0 = (t_in / p.tiles_w) * p.TH;
          const int w0 = (t_in % p.tiles_w) * p.TW;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = smem + (size_t)stage * p.stage_bytes;
          uint8_t* b_dst = a_dst + p.a_bytes;
          mbar_expect_tx(&full_bar[stage], tx);
          for (int a = 0; a < a_atoms; ++a)
            tma_load_4d(a_dst + a * atomA_bytes, &p.tmDY, &full_bar[stage], m0 + a * p.caA + grp * p.g_dy, w0, h0, img);
          for (int b = 0; b < b_atoms; ++b)
            tma_load_4d(b_dst + b * atomB_bytes, &p.tmX[src], &full_bar[stage], c0 + b * p.caB + grp * p.g_x,
                        w0 + s - p.pad_w, h0 + r - p.pad_h, img);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t ltA = umma_layout_type(p.caA * 2), ltB = umma_layout_type(p.caB * 2);
      const uint32_t sboA = 8u * p.caA * 2u, sboB = 8u * p.caB * 2u;  // 8 pixel rows
      const uint32_t kstepA = 16u * p.caA * 2u, kstepB = 16u * p.caB * 2u;  // 16 pixel rows
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
        int tap, nt, mt, ks, grp;
        decode(u, tap, nt, mt, ks, grp);
        const int bn = p.nt_w[nt];
        const uint32_t idesc = umma_idesc(128, bn, p.ab_fmt, 1, 1);
        const int acc = it % p.nacc;
        const uint32_t aphase = (it / p.nacc) & 1;
        mbar_wait(&tempty_bar[acc], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        int pb0, pb1, uimg;
        pb_range(ks, pb0, pb1, uimg);
        for (int pb = pb0; pb < pb1; ++pb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint32_t b_addr = a_addr + p.a_bytes;
          for (int s3 = 0; s3 < p.nsub; ++s3) {
            for (int kk = 0; kk < kWgPix / 16; ++kk) {
              const uint64_t da = umma_smem_desc(a_addr + kk * kstepA, atomA_bytes, sboA, ltA);
              const uint64_t db = umma_smem_desc(b_addr + s3 * p.caB * 2 + kk * kstepB, atomB_bytes, sboB, ltB);
              umma_f16(d_tmem + (uint32_t)(s3 * p.bn_max), da, db, idesc, (uint32_t)((pb > pb0) | (kk != 0)));
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int it = 0;
    for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++it) {
      int tap, nt, mt, ks, grp;
      decode(u, tap, nt, mt, ks, grp);
      const int acc = it % p.nacc;
      const uint32_t aphase = (it / p.nacc) & 1;
      const int m = mt * 128 + row;
      const int bn = p.nt_w[nt];
      int pb0, pb1, uimg;
      pb_range(ks, pb0, pb1, uimg);
      const bool nonempty = pb0 < pb1;
      mbar_wait(&tfull_bar[acc], aphase);
      tc_fence_after();
      for (int s3 = 0; s3 < p.nsub; ++s3) {
        const int tap_idx = p.halo ? tap * 3 + s3 : tap;
        float* dst = p.dw + (long long)uimg * p.dw_img_stride + (long long)m * p.dw_ld + (long long)tap_idx * p.Ctot +
                     p.nt_coff[nt] + grp * p.g_dw;
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * acc_stride +
                                (uint32_t)(s3 * p.bn_max);
        for (int j = 0; j < bn / 16; ++j) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_addr + j * 16, v);
          tmem_ld_wait();
          if (p.partials != nullptr) {
            // ordered mode: plain 128-bit stores of this unit's partial tile (every row, empty units store zeros)
            float4* d4 = reinterpret_cast<float4*>(p.partials + (long long)u * p.tile_floats +
                                                   ((long long)s3 * 128 + row) * p.bn_max + j * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              d4[i] = nonempty ? make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                             __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
          } else if (m < p.Cout && nonempty) {
            // 4 x red.global.add.v4.f32 (16-byte aligned: Ctot, channel offsets and j*16 are multiples of 16)
            float* d = dst + j * 16;
#pragma unroll
            for (int i = 0; i < 4; ++i)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4 * i),
                           "f"(__uint_as_float(v[4 * i])), "f"(__uint_as_float(v[4 * i + 1])),
                           "f"(__uint_as_float(v[4 * i + 2])), "f"(__uint_as_float(v[4 * i + 3]))
                           : "memory");
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// dw[tile] += partial(split 0) + partial(split 1) + ... (fixed order): one thread per float4 of an output tile.
struct WgradReduceParams {
  const float* partials;
  long long tile_floats;
  float* dw;
  long long dw_ld, dw_img_stride, g_dw;
  int Ctot, Cout, bn_max, nsub, halo;
  int taps, n_ntiles, m_tiles, ksplit, Nimg, groups, upg;
  int nt_w[kMaxNTiles], nt_coff[kMaxNTiles];
};

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const __grid_constant__ WgradReduceParams p) {
  GDL_PDL_ENTRY();
  const int vec_per_row = p.bn_max / 4;
  const long long per_tile = (long long)p.nsub * 128 * vec_per_row;
  const long long base_tiles = (long long)p.taps * p.n_ntiles * p.m_tiles;
  const long long total = (long long)p.groups * p.Nimg * base_tiles * per_tile;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int v4 = (int)(t % vec_per_row);
    t /= vec_per_row;
    const int row = (int)(t % 128);
    t /= 128;
    const int s3 = (int)(t % p.nsub);
    t /= p.nsub;
    const int tap = (int)(t % p.taps);
    t /= p.taps;
    const int nt = (int)(t % p.n_ntiles);
    t /= p.n_ntiles;
    const int mt = (int)(t % p.m_tiles);
    t /= p.m_tiles;
    const int img = (int)(t % p.Nimg);
    const int grp = (int)(t / p.Nimg);
    const int m = mt * 128 + row;
    const int col = v4 * 4;
    if (m >= p.Cout || col >= p.nt_w[nt]) continue;
    // unit index of (tile, split): tap fastest, then n-tile, m-tile, split (decode() of conv_wgrad_kernel)
    const long long u0 = (long long)grp * p.upg + (((long long)img * p.ksplit * p.m_tiles + mt) * p.n_ntiles + nt) * p.taps + tap;
    const long long ustride = base_tiles;
    const float* src = p.partials + u0 * p.tile_floats + ((long long)s3 * 128 + row) * p.bn_max + col;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int sp = 0; sp < p.ksplit; ++sp) {
      const float4 x = __ldcg(reinterpret_cast<const float4*>(src + (long long)sp * ustride * p.tile_floats));
      a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w;
    }
    const int tap_idx = p.halo ? tap * 3 + s3 : tap;
    float4* d = reinterpret_cast<float4*>(p.dw + (long long)img * p.dw_img_stride + (long long)m * p.dw_ld +
                                          (long long)tap_idx * p.Ctot + p.nt_coff[nt] + col + (long long)grp * p.g_dw);
    float4 o = *d;
    o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
    *d = o;
  }
}

}  // namespace gdl

extern "C" int gdl_conv2d_nhwc_wgrad(const gdl_conv_wgrad_t* d, void* stream_) {
  GDL_REQUIRE(d != nullptr, GDL_ERR_INVALID, "null descriptor");
  cudaStream_t stream = (cudaStream_t)stream_;
  int Ctot = 0;
  int st = validate_srcs(d->num_src, d->src, &Ctot);
  if (st) return st;
  GDL_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && d->Cout > 0 && d->Cout % 16 == 0, GDL_ERR_INVALID,
              "wgrad: bad shape N=%d H=%d W=%d Cout=%d (Cout must be a multiple of 16)", d->N, d->H,
              d->W, d->Cout);
  GDL_REQUIRE(d->R >= 1 && d->S >= 1 && d->R <= 15 && d->S <= 15, GDL_ERR_INVALID, "bad filter");
  GDL_REQUIRE(d->dtype == GDL_BF16 || d->dtype == GDL_F16, GDL_ERR_INVALID, "operand dtype must be bf16/f16");
  GDL_REQUIRE(d->dy && d->dw, GDL_ERR_INVALID, "null dy/dw");
  GDL_REQUIRE(d->ld_dy >= d->Cout && d->ld_dy % 8 == 0, GDL_ERR_INVALID, "bad ld_dy %d", d->ld_dy);
  const int Ho = d->H + 2 * d->pad_h - d->R + 1;
  const int Wo = d->W + 2 * d->pad_w - d->S + 1;
  GDL_REQUIRE(Ho > 0 && Wo > 0, GDL_ERR_INVALID, "empty output");

  if (opt_int(g_opt_wgrad_rows, "GDL_WGRAD_ROWS", 1) > 0) {
    // narrow outputs (Cout 16/32/64), 3x3: paired-tap row-streaming kernel (wgrad3x3_rows.cu)
    int st_rows = 0;
    if (wgrad3x3_rows_try(d, stream, &st_rows)) return st_rows;
  }

  ConvWgradKParams p;
  memset(&p, 0, sizeof(p));
  p.Ctot = Ctot;
  p.Cout = d->Cout;
  p.R = d->R;
  p.S = d->S;
  p.pad_h = d->pad_h;
  p.pad_w = d->pad_w;
  const int G = d->groups > 1 ? d->groups : 1;
  if (G > 1) {
    // all heads of dV = P^T.dO / dK = dS^T.q in one launch
    GDL_REQUIRE(d->batched && d->num_src == 1 && d->R == 1 && d->S == 1 && d->pad_h == 0 && d->pad_w == 0, GDL_ERR_INVALID,
                "grouped wgrad: batched 1x1 products of one source only");
    GDL_REQUIRE(d->g_src_stride % 8 == 0 && d->g_dy_stride % 8 == 0 && d->g_dw_stride % 4 == 0 && d->g_src_stride > 0 &&
                    d->g_dy_stride > 0,
                GDL_ERR_INVALID, "grouped wgrad: group strides must keep 16-byte alignment");
    GDL_REQUIRE(d->ld_dy >= d->Cout + (G - 1) * d->g_dy_stride && d->src[0].ld >= d->src[0].channels + (G - 1) * d->g_src_stride,
                GDL_ERR_INVALID, "grouped wgrad: groups exceed the leading dimensions");
  }
  p.g_x = G > 1 ? d->g_src_stride : 0;
  p.g_dy = G > 1 ? d->g_dy_stride : 0;
  p.g_dw = G > 1 ? d->g_dw_stride : 0;
  p.caB = chunk_width(d->src, d->num_src);
  p.caA = 64;
  while (p.caA > 16 && (d->Cout % p.caA) != 0) p.caA >>= 1;

  int N = d->N, H = d->H, W = d->W, oH = Ho, oW = Wo;
  if (d->R == 1 && d->S == 1 && d->pad_h == 0 && d->pad_w == 0 && !d->batched) {
    long long M = (long long)N * H * W;
    GDL_REQUIRE(M < (1ll << 31), GDL_ERR_UNSUPPORTED, "too many pixels");
    W = (int)M;
    H = 1;
    N = 1;
    oH = 1;
    oW = W;
  }
  choose_tile(oH, oW, kWgPix, &p.TH, &p.TW);
  {
    const int halo_cfg = opt_int(g_opt_wgrad_halo, "GDL_WGRAD_HALO", 1);
    p.halo = halo_cfg > 0 && d->R == 3 && d->S == 3 && d->pad_h == 1 && d->pad_w == 1 && p.TH == 1 &&
             p.TW == kWgPix && !d->batched;
  }
  p.nsub = p.halo ? 3 : 1;
  p.unit_taps = p.halo ? 3 : d->R * d->S;
  p.b_atom_bytes = p.halo ? (((kWgPix + 2) * p.caB * 2 + 1023) / 1024) * 1024 : kWgPix * p.caB * 2;
  p.tiles_w = (oW + p.TW - 1) / p.TW;
  p.tiles_h = (oH + p.TH - 1) / p.TH;
  long long pbs = (long long)N * p.tiles_w * p.tiles_h;
  GDL_REQUIRE(pbs < (1ll << 31), GDL_ERR_UNSUPPORTED, "too many pixel blocks");
  p.pix_blocks = (int)pbs;
  p.m_tiles = (d->Cout + 127) / 128;

  // n-tiles: chunks of <= 256 channels that never straddle two sources
  int nn = 0, coff = 0, bn_max = 16;
  for (int i = 0; i < d->num_src; ++i) {
    int c = d->src[i].channels;
    const int cap = p.halo ? 128 : 256;  // halo: 3 accumulators of <= 128 columns fit TMEM (single-buffered)
    int parts = (c + cap - 1) / cap;
    int w = ((c + parts - 1) / parts + p.caB - 1) / p.caB * p.caB;
    for (int c0 = 0; c0 < c; c0 += w) {
      GDL_REQUIRE(nn < kMaxNTiles, GDL_ERR_UNSUPPORTED, "too many channel tiles");
      p.nt_src[nn] = i;
      p.nt_c0[nn] = c0;
      p.nt_w[nn] = (c - c0) < w ? (c - c0) : w;
      p.nt_coff[nn] = coff + c0;
      if (p.nt_w[nn] > bn_max) bn_max = p.nt_w[nn];
      ++nn;
    }
    coff += c;
    st = make_tmap_nhwc(&p.tmX[i], d->src[i].ptr, d->dtype, c + (long long)(G - 1) * d->g_src_stride, W, H, N, d->src[i].ld, p.caB,
                        p.halo ? p.TW + 2 : p.TW, p.TH, p.caB * 2);
    if (st) return st;
  }
  p.n_ntiles = nn;
  p.bn_max = bn_max;
  st = make_tmap_nhwc(&p.tmDY, d->dy, d->dtype, d->Cout + (long long)(G - 1) * d->g_dy_stride, oW, oH, N, d->ld_dy, p.caA, p.TW, p.TH,
                      p.caA * 2);
  if (st) return st;

  const int taps = p.unit_taps;
  p.batched = d->batched;
  p.pb_per_img = p.tiles_w * p.tiles_h;
  p.dw_ld = d->dw_ld > 0 ? d->dw_ld : (long long)d->R * d->S * Ctot;
  p.dw_img_stride = d->dw_img_stride;
  const DetWs ws = det_workspace();
  p.tile_floats = (long long)p.nsub * 128 * bn_max;  // one partial accumulator tile (ordered mode)
  const long long max_units_ws = ws.ok() ? ws.slot_floats / p.tile_floats : (1ll << 40);
  if (d->batched) {
    // one independent product per image: split each image's pixel blocks on its own
    GDL_REQUIRE(!(d->R == 1 && d->S == 1 && d->pad_h == 0 && d->pad_w == 0) || N == d->N, GDL_ERR_INVALID, "batched wgrad");
    long long bu = (long long)p.m_tiles * nn * taps * N;
    int ks = (int)((2ll * sm_count() + bu - 1) / bu);
    int max_ks = p.pb_per_img / 8;
    if (max_ks < 1) max_ks = 1;
    if (ks > max_ks) ks = max_ks;
    while (ks > 1 && bu * ks * G > max_units_ws) --ks;  // ordered mode: the partial tiles must fit the workspace
    if (ks < 1) ks = 1;
    p.pb_per_split = (p.pb_per_img + ks - 1) / ks;
    p.ksplit = (p.pb_per_img + p.pb_per_split - 1) / p.pb_per_split;
    long long units = bu * p.ksplit;
    GDL_REQUIRE(units < (1ll << 31), GDL_ERR_UNSUPPORTED, "too many work units");
    p.num_units = (int)units;
  }
  long long base_units = (long long)p.m_tiles * nn * taps;
  int ks = 1;
  if (opt_int(g_opt_wgrad_sched, "GDL_WGRAD_SCHED", 1) > 0) {
    // Round 2: few LONG units.  Split the pixel range into as few pieces as keep every SM busy: with base_units * ks <= #SM
    // each CTA owns one (tile, pixel range) for the whole launch and accumulates it in TMEM — ONE epilogue per CTA instead of
    // one per ~64-pixel-block unit (the round-1 rule below produced up to ~4600 units per launch: as many bytes of fp32
    // reductions as operand bytes, and, in ordered mode, that many partial tiles).  CTAs of one pixel range run in lockstep
    // (equal work per pixel block), so the range is still read from HBM once and shared through L2.  When base_units
    // exceeds the SM count, ks is the split with the smallest makespan  ceil(base_units * ks / #SM) / ks.
    const int sms = sm_count();
    int max_ks = p.pix_blocks / 16;
    if (max_ks > 64) max_ks = 64;
    if (max_ks < 1) max_ks = 1;
    double best = 1e30;
    for (int g = 1; g <= max_ks; ++g) {
      const long long units = base_units * g;
      if (g > 1 && units > max_units_ws) break;
      const double cost = (double)((units + sms - 1) / sms) / g;
      if (cost < best * 0.97) {
        best = cost;
        ks = g;
      }
    }
  } else {
    // Round-1 rule: (a) the grid should see >= ~3 waves; (b) the pixels one wave of co-resident units streams (dY + X rows
    // of one split) should stay L2-resident (budget GDL_WGRAD_L2_MB, default 8 MB); (c) keep >= 32 stages of MMA work per unit.
    ks = (int)((3ll * sm_count() + base_units - 1) / base_units);
    if (g_opt_wgrad_l2_mb < 0) {
      const char* e = getenv("GDL_WGRAD_L2_MB");
      g_opt_wgrad_l2_mb = e ? atoll(e) : 8;
    }
    const long long budget = g_opt_wgrad_l2_mb * (1ll << 20);
    const long long bytes_per_pb = (long long)kWgPix * (Ctot + d->Cout) * 2;
    long long pb_budget = budget / bytes_per_pb;
    if (pb_budget < 32) pb_budget = 32;
    int ks_l2 = (int)((p.pix_blocks + pb_budget - 1) / pb_budget);
    if (ks_l2 > ks) ks = ks_l2;
    int max_ks = p.pix_blocks / 32;
    if (max_ks < 1) max_ks = 1;
    if (ks > max_ks) ks = max_ks;
    while (ks > 1 && base_units * ks > max_units_ws) --ks;  // ordered mode: the partial tiles must fit the workspace
  }
  if (ks < 1) ks = 1;
  if (!d->batched) {
    p.pb_per_split = (p.pix_blocks + ks - 1) / ks;
    p.ksplit = (p.pix_blocks + p.pb_per_split - 1) / p.pb_per_split;
    long long units = base_units * p.ksplit;
    GDL_REQUIRE(units < (1ll << 31), GDL_ERR_UNSUPPORTED, "too many work units");
    p.num_units = (int)units;
  }

  p.upg = p.num_units;
  GDL_REQUIRE((long long)p.num_units * G < (1ll << 31), GDL_ERR_UNSUPPORTED, "too many work units");
  p.num_units *= G;
  p.Nimg = d->batched ? N : 1;

  p.a_bytes = kWgPix * 128 * 2;
  const int b_bytes = p.halo ? ((bn_max + p.caB - 1) / p.caB) * p.b_atom_bytes : kWgPix * bn_max * 2;
  p.stage_bytes = p.a_bytes + ((b_bytes + 1023) / 1024) * 1024;
  p.stages = kSmemBudget / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  GDL_REQUIRE(p.stages >= 2, GDL_ERR_UNSUPPORTED, "tile does not fit shared memory");
  p.nacc = (2 * p.nsub * bn_max <= 512) ? 2 : 1;
  p.tmem_cols = pow2_ge(p.nacc * p.nsub * bn_max);
  GDL_REQUIRE(p.tmem_cols <= 512, GDL_ERR_UNSUPPORTED, "accumulators do not fit TMEM");
  p.ab_fmt = d->dtype == GDL_BF16 ? 1 : 0;
  p.dw = d->dw;

  if (p.ksplit > 1 && ws.ok()) {
    // several units add into the same dW tile: give every unit a partial tile and add them in split order afterwards
    GDL_REQUIRE((long long)p.num_units * p.tile_floats <= ws.slot_floats, GDL_ERR_UNSUPPORTED,
                "wgrad: %d partial tiles of %lld floats exceed the registered workspace (gdl_query_workspace_bytes)",
                p.num_units, p.tile_floats);
    p.partials = ws.slots;
  }

  int smem = p.stages * p.stage_bytes + 1024;
  if (smem < kMinSmemRequest) smem = kMinSmemRequest;
  static PerDeviceOnce attr_once;
  GDL_CHECK_CUDA(set_max_dyn_smem_once(attr_once, conv_wgrad_kernel, kSmemBudget + 4096));
  int grid = p.num_units < sm_count() ? p.num_units : sm_count();
  GDL_LAUNCH(conv_wgrad_kernel, grid, kConvThreads, smem, stream, p);
  GDL_CHECK_CUDA(cudaGetLastError());
  if (p.partials != nullptr) {
    WgradReduceParams r;
    memset(&r, 0, sizeof(r));
    r.partials = p.partials;
    r.tile_floats = p.tile_floats;
    r.dw = p.dw;
    r.dw_ld = p.dw_ld;
    r.dw_img_stride = p.dw_img_stride;
    r.g_dw = p.g_dw;
    r.Ctot = p.Ctot;
    r.Cout = p.Cout;
    r.bn_max = bn_max;
    r.nsub = p.nsub;
    r.halo = p.halo;
    r.taps = p.unit_taps;
    r.n_ntiles = p.n_ntiles;
    r.m_tiles = p.m_tiles;
    r.ksplit = p.ksplit;
    r.Nimg = p.Nimg;
    r.groups = G;
    r.upg = p.upg;
    for (int i = 0; i < p.n_ntiles; ++i) {
      r.nt_w[i] = p.nt_w[i];
      r.nt_coff[i] = p.nt_coff[i];
    }
    const long long total = (long long)G * p.Nimg * p.unit_taps * p.n_ntiles * p.m_tiles * p.nsub * 128 * (bn_max / 4);
    long long rb = (total + 255) / 256;
    if (rb > 8ll * sm_count()) rb = 8ll * sm_count();
    GDL_LAUNCH(wgrad_reduce_kernel, (int)rb, 256, 0, stream, r);
    GDL_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// ==========================================================================================
// weight layout transforms (tiny, HBM-bound; run once per optimizer step)
// ==========================================================================================
namespace gdl {

// dst row r (of `rows` rows, `cols` valid columns, row stride dst_ld, zero padded):
//   mode 0: rows = Cout, cols = (r,s,c)        dst[k][(r,s,c)]           = src[k][c][r][s]
//   mode 1: rows = Cin,  cols = (r',s',k)      dst[c][(R-1-r,S-1-s,k)]   = src[k][c][r][s]   (dgrad)
//   mode 2: rows = (r,s,c), cols = k           dst[(r,s,c)][k]           = src[k][c][r][s]   (im2col dgrad)
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ src, T* __restrict__ dst, int Cout, int Cin,
                                   int R, int S, int mode, int rows, int cols, int dst_ld) {
  GDL_PDL_ENTRY();
  const long long total = (long long)rows * dst_ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / dst_ld);
    const int col = (int)(i - (long long)row * dst_ld);
    float v = 0.f;
    if (col < cols) {
      int k, c, r, s;
      if (mode == 0) {
        k = row;
        c = col % Cin;
        const int t = col / Cin;
        s = t % S;
        r = t / S;
      } else if (mode == 1) {
        c = row;
        k = col % Cout;
        const int t = col / Cout;
        s = S - 1 - (t % S);
        r = R - 1 - (t / S);
      } else {
        k = col;
        c = row % Cin;
        const int t = row / Cin;
        s = t % S;
        r = t / S;
      }
      v = src[(((long long)k * Cin + c) * R + r) * S + s];
    }
    if constexpr (std::is_same<T, __nv_bfloat16>::value)
      dst[i] = __float2bfloat16_rn(v);
    else
      dst[i] = __float2half_rn(v);
  }
}

// All packed operands of a model in ONE launch (after the optimizer step of the fused trainer: the fp32 masters changed,
// the cached 16-bit operands are refreshed in place).  Entry e covers chunks [chunk0[e], chunk0[e+1]) of kRepackChunk
// elements; a block finds its entry by binary search.
constexpr int kRepackChunk = 4096;
template <typename T>
__global__ void __launch_bounds__(256) repack_weights_kernel(const gdl_repack_t* __restrict__ table, const int* __restrict__ chunk0,
                                                              int n_entries) {
  GDL_PDL_ENTRY();
  int lo = 0, hi = n_entries;  // last entry with chunk0[e] <= blockIdx.x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (chunk0[mid] <= (int)blockIdx.x) lo = mid;
    else hi = mid;
  }
  const gdl_repack_t e = table[lo];
  const int Cout = e.Cout, Cin = e.Cin, R = e.R, S = e.S, mode = e.mode, dst_ld = e.dst_ld;
  const int rows = mode == 0 ? Cout : (mode == 1 ? Cin : R * S * Cin);
  const int cols = mode == 0 ? R * S * Cin : (mode == 1 ? R * S * Cout : Cout);
  const long long total = (long long)rows * dst_ld;
  const long long base = (long long)((int)blockIdx.x - chunk0[lo]) * kRepackChunk;
  const float* __restrict__ src = e.src;
  T* __restrict__ dst = reinterpret_cast<T*>(e.dst);
  for (long long i = base + threadIdx.x; i < base + kRepackChunk && i < total; i += blockDim.x) {
    const int row = (int)(i / dst_ld);
    const int col = (int)(i - (long long)row * dst_ld);
    float v = 0.f;
    if (col < cols) {
      int k, c, r, s2;
      if (mode == 0) {
        k = row;
        c = col % Cin;
        const int t = col / Cin;
        s2 = t % S;
        r = t / S;
      } else if (mode == 1) {
        c = row;
        k = col % Cout;
        const int t = col / Cout;
        s2 = S - 1 - (t % S);
        r = R - 1 - (t / S);
      } else {
        k = col;
        c = row % Cin;
        const int t = row / Cin;
        s2 = t % S;
        r = t / S;
      }
      v = src[(((long long)k * Cin + c) * R + r) * S + s2];
    }
    if constexpr (std::is_same<T, __nv_bfloat16>::value)
      dst[i] = __float2bfloat16_rn(v);
    else
      dst[i] = __float2half_rn(v);
  }
}

// grad fp32 [Cout][src_ld] with columns (r,s,c) -> fp32 OIHW
__global__ void unpack_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout,
                                    int Cin, int R, int S, int src_ld, int accumulate) {
  GDL_PDL_ENTRY();
  const long long total = (long long)Cout * Cin * R * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int s = t % S;
    t /= S;
    const int r = t % R;
    t /= R;
    const int c = t % Cin;
    const int k = t / Cin;
    const float v = src[(long long)k * src_ld + ((long long)r * S + s) * Cin + c];
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

// ---- pixel-packed form of a narrow 3-wide conv ---------------------------------------------------------------
// A 3x3 / pad-1 conv over Ci in {16, 32} channels wastes the tensor-core tile (32/64-byte TMA rows, N = 16).  Viewing
// f = 64/Ci horizontally adjacent pixels as ONE pixel of f*Ci channels turns it into a 3x3 conv (N,H,W/f,f*Ci) ->
// (N,H,W/f,f*Co) with the block-Toeplitz weights below: same memory, 128-byte rows, N = f*Co.
//   dst[(j,o)][ky][sx][(j',c)] = src[o][ky][kx][c],  kx = f*(sx-1) + j' - j + 1  (zero when kx is outside 0..2)
template <typename T>
__global__ void widen_weight_kernel(const T* __restrict__ src, T* __restrict__ dst, int Co, int Ci, int R, int f) {
  GDL_PDL_ENTRY();
  const long long total = (long long)f * Co * R * 3 * f * Ci;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int cc = (int)(t % (f * Ci));
    t /= f * Ci;
    const int sx = (int)(t % 3);
    t /= 3;
    const int ky = (int)(t % R);
    const int row = (int)(t / R);
    const int jp = cc / Ci, c = cc - jp * Ci;
    const int j = row / Co, o = row - j * Co;
    const int kx = f * (sx - 1) + jp - j + 1;
    T v = T(0.f);
    if (kx >= 0 && kx < 3) v = src[(((long long)o * R + ky) * 3 + kx) * Ci + c];
    dst[i] = v;
  }
}

// fp32 gradient of the widened weights [f*Co][src_ld] (columns (ky,sx,(j',c))) -> fp32 OIHW of the real conv
__global__ void fold_widened_wgrad_kernel(const float* __restrict__ src, int src_ld, int src_co,
                                          float* __restrict__ dst, int Co, int Ci, int R, int f, int accumulate) {
  GDL_PDL_ENTRY();
  const long long total = (long long)Co * Ci * R * 3;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    const int kx = (int)(t % 3);
    t /= 3;
    const int ky = (int)(t % R);
    t /= R;
    const int c = (int)(t % Ci);
    const int o = (int)(t / Ci);
    float v = 0.f;
    for (int j = 0; j < f; ++j) {
      const int d = j + kx - 1;  // input pixel relative to the packed pixel's first column
      const int sx = d < 0 ? 0 : (d >= f ? 2 : 1);
      const int jp = d - f * (sx - 1);
      v += src[(long long)(j * src_co + o) * src_ld + ((long long)(ky * 3 + sx) * f + jp) * Ci + c];
    }
    dst[i] = accumulate ? dst[i] + v : v;
  }
}

}  // namespace gdl

extern "C" int gdl_widen_conv_weight(const void* src, void* dst, int Co, int Ci, int R, int f, int dtype,
                                     void* stream) {
  GDL_REQUIRE(src && dst && Co > 0 && Ci > 0 && R > 0 && f >= 2, GDL_ERR_INVALID, "widen_conv_weight: bad args");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_INVALID, "widen_conv_weight: dtype");
  const long long total = (long long)f * Co * R * 3 * f * Ci;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == GDL_BF16)
    GDL_LAUNCH(gdl::widen_weight_kernel<__nv_bfloat16>, blocks, 256, 0, (cudaStream_t)stream, 
        (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, Co, Ci, R, f);
  else
    GDL_LAUNCH(gdl::widen_weight_kernel<__half>, blocks, 256, 0, (cudaStream_t)stream, (const __half*)src, (__half*)dst, Co, Ci,
                                                                            R, f);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_fold_widened_wgrad(const float* src, int src_ld, int src_co, float* dst, int Co, int Ci, int R,
                                      int f, int accumulate, void* stream) {
  GDL_REQUIRE(src && dst && Co > 0 && Ci > 0 && R > 0 && f >= 2, GDL_ERR_INVALID, "fold_widened_wgrad: bad args");
  if (src_co <= 0) src_co = Co;
  GDL_REQUIRE(src_co >= Co, GDL_ERR_INVALID, "fold_widened_wgrad: src_co < Co");
  if (src_ld <= 0) src_ld = R * 3 * f * Ci;
  GDL_REQUIRE(src_ld >= R * 3 * f * Ci, GDL_ERR_INVALID, "fold_widened_wgrad: src_ld too small");
  const long long total = (long long)Co * Ci * R * 3;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  GDL_LAUNCH(gdl::fold_widened_wgrad_kernel, blocks, 256, 0, (cudaStream_t)stream, src, src_ld, src_co, dst, Co, Ci, R, f,
                                                                           accumulate);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_pack_conv_weight(const float* src, void* dst, int Cout, int Cin, int R, int S,
                                    int mode, int dst_ld, int dtype, void* stream) {
  GDL_REQUIRE(src && dst && Cout > 0 && Cin > 0 && R > 0 && S > 0, GDL_ERR_INVALID, "pack_weight: bad args");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_INVALID, "pack_weight: dtype");
  GDL_REQUIRE(mode >= 0 && mode <= 2, GDL_ERR_INVALID, "pack_weight: mode %d", mode);
  const int rows = mode == 0 ? Cout : (mode == 1 ? Cin : R * S * Cin);
  const int cols = mode == 0 ? R * S * Cin : (mode == 1 ? R * S * Cout : Cout);
  if (dst_ld <= 0) dst_ld = cols;
  GDL_REQUIRE(dst_ld >= cols, GDL_ERR_INVALID, "pack_weight: dst_ld %d < %d", dst_ld, cols);
  const long long total = (long long)rows * dst_ld;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == GDL_BF16)
    GDL_LAUNCH(pack_weight_kernel<__nv_bfloat16>, blocks, 256, 0, (cudaStream_t)stream, 
        src, (__nv_bfloat16*)dst, Cout, Cin, R, S, mode, rows, cols, dst_ld);
  else
    GDL_LAUNCH(pack_weight_kernel<__half>, blocks, 256, 0, (cudaStream_t)stream, src, (__half*)dst, Cout, Cin, R, S,
                                                                        mode, rows, cols, dst_ld);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_repack_weights(const gdl_repack_t* table_dev, const int* chunk0_dev, int n_entries, int total_chunks,
                                  int dtype, void* stream) {
  GDL_REQUIRE(table_dev && chunk0_dev && n_entries > 0 && total_chunks > 0, GDL_ERR_INVALID, "repack_weights: bad args");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_INVALID, "repack_weights: dtype");
  if (dtype == GDL_BF16)
    GDL_LAUNCH(repack_weights_kernel<__nv_bfloat16>, total_chunks, 256, 0, (cudaStream_t)stream, table_dev, chunk0_dev, n_entries);
  else
    GDL_LAUNCH(repack_weights_kernel<__half>, total_chunks, 256, 0, (cudaStream_t)stream, table_dev, chunk0_dev, n_entries);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int gdl_unpack_conv_wgrad(const float* src, float* dst, int Cout, int Cin, int R, int S,
                                     int src_ld, int accumulate, void* stream) {
  GDL_REQUIRE(src && dst && Cout > 0 && Cin > 0 && R > 0 && S > 0, GDL_ERR_INVALID, "unpack_wgrad: bad args");
  if (src_ld <= 0) src_ld = R * S * Cin;
  GDL_REQUIRE(src_ld >= R * S * Cin, GDL_ERR_INVALID, "unpack_wgrad: src_ld too small");
  const long long total = (long long)Cout * Cin * R * S;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  GDL_LAUNCH(unpack_wgrad_kernel, blocks, 256, 0, (cudaStream_t)stream, src, dst, Cout, Cin, R, S, src_ld, accumulate);
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}
