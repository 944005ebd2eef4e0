// sra_attention.cu — fused attention forward, head dim 64:  o = softmax(scale * q.k^T) . v  per (image, head), without the score
// tensor ever reaching HBM.  Two users:
//   * SegFormer's spatial-reduction attention (Attention.forward, mix_transformer.py:131-159; MiT-B1 .. B5): at most 256 reduced
//     keys — a 512x512 tile has exactly 256 at every stage (N / sr^2) — so the whole key range is ONE block: K and V stay in
//     shared memory across the query tiles of a head, and for training the normalised probabilities are written once (the backward's
//     dV = P^T.dO and softmax_bwd read them);
//   * DOFA's ViT blocks (timm Attention inside dofa_v2.py:445-487): ~1000-1300 keys, streamed in blocks of 128 through a two-stage
//     ring with the online-softmax recurrence (running max m, running sum l, O <- O * 2^(m_old - m_new) + P~.V_block), forward only.
//
// One CTA works on 128-query tiles of one (image, head) at a time:
//   warp 0      TMA producer: K / V blocks, Q tiles (128 x 64, double buffered)
//   warp 1      MMA issuer:   S = Q.K_blk^T (M 128, N = block keys, K 64) -> TMEM columns [0, 256)
//                             O_blk = P~.V_blk (M 128, N 64, K = block keys) -> TMEM columns [256, 320); V is the MN-major B operand
//   warps 2-5   one query row per thread: block max and exp2 straight from tcgen05.ld, un-normalised P~ (16-bit) into the swizzled
//               shared-memory A operand of the second MMA, O accumulated in registers across key blocks, scaled by 1 / l on the way
//               out (TMA store).
// The 128 x 256 fp32 score tile is exactly 256 TMEM columns.  Per tile the exponentials bound the time (MUFU), not the tensor pipe —
// the point of the kernel is the HBM traffic it removes: scores written / read / re-written / read (8 B per score) become 0
// (inference) or one 2-byte write (SRA training).
#include <stdint.h>

#include "../../include/gdl_b200.h"
#include "tmap.cuh"

namespace gdl {

constexpr int kSraThreads = 192;
constexpr int kSraD = 64;           // head dim of MiT-B1 .. B5 and of ViT-B / L (C / heads = 64; MiT-B0 has 32: three-kernel path)
constexpr int kSraMaxKeys = 256;    // keys of one block (one S tile in TMEM)
constexpr int kSraStreamKeys = 128; // block size when the keys are streamed (two-stage ring inside the same 64 KB)
constexpr int kSraTileBytes = 128 * 128;                 // 128 rows x 64 16-bit values, SWIZZLE_128B
constexpr int kSraOffK = 0, kSraOffV = 32768, kSraOffQ = 65536, kSraOffP = 98304, kSraOffO = 163840;
constexpr int kSraSmem = kSraOffO + kSraTileBytes;       // 176 KB (+1 KB alignment slack)
constexpr uint32_t kSraColO = 256;                        // TMEM column of the O accumulator

int g_opt_sra_max_ctas = 0;  // gdl_set_option("sra_max_ctas", n): cap the persistent grid (0 = SM count)

struct SraParams {
  CUtensorMap tmQ, tmK, tmV, tmO, tmP;
  CUtensorMap tmPin;                // backward: the saved probabilities (load side)
  float scale;                      // backward: dS = scale * P * (dP - delta)
  int B, heads, qtiles, nk;
  int kb, nblocks;                  // keys per block, blocks per query tile (1: K / V resident per (image, head))
  int lp;                           // columns per head of the saved probabilities
  int save_p;
  float scale_log2e;                // scale * log2(e): p = exp2(scale_log2e * (s - max))
  int items, per_cta;
};

template <int FMT>
GDL_DEVINL uint32_t sra_pack2(float a, float b) {
  return FMT == 1 ? pack_bf16x2(a, b) : pack_f16x2(a, b);
}
template <int FMT>
GDL_DEVINL float2 sra_unpack2(uint32_t u) {
  if (FMT == 1) return make_float2(bf16_lo(u), bf16_hi(u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// BWD = false: forward (above).  BWD = true: the query-tile-local half of the backward with the same pipeline — the roles are
//   Q -> dO,  K -> V (first MMA: dP = dO.V^T),  softmax -> dS = scale * P * (dP - rowsum(P * dP)) written IN PLACE over the P tile that
//   TMA loaded,  V -> K (second MMA: dQ = dS.K),  O -> dQ,  saved P -> dS (stored for the dK = dS^T.q wgrad).  One key block only.
template <int FMT, bool BWD>
__global__ void __launch_bounds__(kSraThreads, 1) sra_attention_kernel(const __grid_constant__ SraParams p) {
  GDL_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);

  __shared__ __align__(8) uint64_t kv_full[2], kv_empty[2];
  __shared__ __align__(8) uint64_t q_full[2], q_empty[2];
  __shared__ __align__(8) uint64_t s_full, p_full, o_full;
  __shared__ __align__(8) uint64_t pin_full, p_free;  // backward: P tile landed / P tile may be overwritten
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int item0 = blockIdx.x * p.per_cta;
  const int item1 = min(p.items, item0 + p.per_cta);
  const bool resident = p.nblocks == 1;  // one key block: K / V are loaded once per (image, head) and reused by its query tiles
  const int kv_stage_bytes = p.kb * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    tma_prefetch_desc(&p.tmO);
    if (p.save_p) tma_prefetch_desc(&p.tmP);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&kv_full[i], 1);
        mbar_init(&kv_empty[i], 1);
        mbar_init(&q_full[i], 1);
        mbar_init(&q_empty[i], 1);
      }
      mbar_init(&s_full, 1);
      mbar_init(&p_full, 128);
      mbar_init(&o_full, 1);
      mbar_init(&pin_full, 1);
      mbar_init(&p_free, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int cur_bg = -1;
      uint32_t kc = 0;  // K / V loads issued so far: stage kc & 1 (resident mode: always stage 0), phase (kc >> 1) & 1
      for (int item = item0, it = 0; item < item1; ++item, ++it) {
        const int bg = item / p.qtiles, qt = item - bg * p.qtiles;
        const int b = bg / p.heads, g = bg - b * p.heads;
        const int s = it & 1;
        mbar_wait(&q_empty[s], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[s], (uint32_t)kSraTileBytes);
        tma_load_4d(smem + kSraOffQ + s * kSraTileBytes, &p.tmQ, &q_full[s], g * kSraD, qt * 128, 0, b);
        if (BWD) {
          // the P tile is single buffered: wait until the previous tile's dS stores have read it and its second MMA is done
          mbar_wait(&p_free, it & 1);
          mbar_expect_tx(&pin_full, (uint32_t)(p.nk * 256));
          for (int ch = 0; ch < p.nk; ch += 64)
            tma_load_4d(smem + kSraOffP + (ch >> 6) * kSraTileBytes, &p.tmPin, &pin_full, g * p.lp + ch, qt * 128, 0, b);
        }
        if (resident) {
          if (bg != cur_bg) {
            mbar_wait(&kv_empty[0], (kc & 1) ^ 1);  // every MMA that read the previous K / V has completed
            mbar_expect_tx(&kv_full[0], (uint32_t)(2 * kv_stage_bytes));
            tma_load_4d(smem + kSraOffK, &p.tmK, &kv_full[0], g * kSraD, 0, 0, b);
            tma_load_4d(smem + kSraOffV, &p.tmV, &kv_full[0], g * kSraD, 0, 0, b);
            ++kc;
            cur_bg = bg;
          }
        } else {
          for (int j = 0; j < p.nblocks; ++j, ++kc) {
            const int st = kc & 1;
            mbar_wait(&kv_empty[st], ((kc >> 1) & 1) ^ 1);
            mbar_expect_tx(&kv_full[st], (uint32_t)(2 * kv_stage_bytes));
            // rows past the image's last key are zero-filled by the TMA unit (per-image tensor map)
            tma_load_4d(smem + kSraOffK + st * kv_stage_bytes, &p.tmK, &kv_full[st], g * kSraD, j * p.kb, 0, b);
            tma_load_4d(smem + kSraOffV + st * kv_stage_bytes, &p.tmV, &kv_full[st], g * kSraD, j * p.kb, 0, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, p.kb, FMT, 0, 0);
      const uint32_t idesc_o = umma_idesc(128, kSraD, FMT, 0, 1);  // B = V: [key][d], d contiguous (MN-major)
      const uint32_t lt = umma_layout_type(128);
      const uint32_t k_base = smem_u32(smem + kSraOffK), v_base = smem_u32(smem + kSraOffV);
      const uint32_t p_addr = smem_u32(smem + kSraOffP);
      int cur_bg = -1;
      uint32_t kc = 0;  // K / V blocks consumed (same numbering as the producer)
      uint32_t sb = 0;  // (tile, block) steps: phase of s_full / p_full / o_full
      for (int item = item0, it = 0; item < item1; ++item, ++it) {
        const int bg = item / p.qtiles;
        const int s = it & 1;
        mbar_wait(&q_full[s], (it >> 1) & 1);
        const uint32_t q_addr = smem_u32(smem + kSraOffQ + s * kSraTileBytes);
        for (int j = 0; j < p.nblocks; ++j, ++sb) {
          int st = 0;
          if (resident) {
            if (bg != cur_bg) {
              mbar_wait(&kv_full[0], kc & 1);
              ++kc;
              cur_bg = bg;
            }
          } else {
            st = kc & 1;
            mbar_wait(&kv_full[st], (kc >> 1) & 1);
            ++kc;
          }
          tc_fence_after();
          // S may be overwritten: p_full of the previous step (awaited below) was signalled after its last read of S
          const uint32_t k_addr = k_base + st * kv_stage_bytes, v_addr = v_base + st * kv_stage_bytes;
          for (int kk = 0; kk < kSraD / 16; ++kk)
            umma_f16(tmem_base, umma_smem_desc(q_addr + kk * 32, 16, 1024, lt), umma_smem_desc(k_addr + kk * 32, 16, 1024, lt),
                     idesc_s, (uint32_t)(kk != 0));
          if (j == p.nblocks - 1) umma_commit(&q_empty[s]);
          umma_commit(&s_full);
          mbar_wait(&p_full, sb & 1);  // P~ of this step is in shared memory (and O_blk of the previous step has been read)
          tc_fence_after();
          for (int ks = 0; ks < p.kb / 16; ++ks)
            umma_f16(tmem_base + kSraColO,
                     umma_smem_desc(p_addr + (ks >> 2) * kSraTileBytes + (ks & 3) * 32, 16, 1024, lt),
                     umma_smem_desc(v_addr + ks * 2048, 16, 1024, lt), idesc_o, (uint32_t)(ks != 0));
          umma_commit(&o_full);
          if (resident) {
            const bool last_of_bg = item + 1 == item1 || (item + 1) / p.qtiles != bg;
            if (last_of_bg) umma_commit(&kv_empty[0]);
          } else {
            umma_commit(&kv_empty[st]);
          }
        }
      }
    }
  } else {
    // ===================== softmax + epilogue: one query row per thread =====================
    const int q = warp & 3;           // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;
    const bool issuer = (warp == 2 && lane == 0);
    const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* p_smem = smem + kSraOffP;
    uint8_t* o_smem = smem + kSraOffO;
    const uint32_t rsw = (uint32_t)(row & 7);
    uint32_t sb = 0;
    for (int item = item0; item < item1; ++item) {
      const int bg = item / p.qtiles, qt = item - bg * p.qtiles;
      const int b = bg / p.heads, g = bg - b * p.heads;
      float m_run = -INFINITY, l_run = 0.f;
      float o_run[kSraD];
      if (BWD && issuer) {
        bulk_wait_group_read<0>();  // the dS / dQ stores of the previous tile have finished reading shared memory
        mbar_arrive(&p_free);
      }
      for (int j = 0; j < p.nblocks; ++j, ++sb) {
        const int valid = min(p.kb, p.nk - j * p.kb);  // keys of this block that exist (the rest is TMA zero fill)
        mbar_wait(&s_full, sb & 1);
        tc_fence_after();
        if (BWD) {
          mbar_wait(&pin_full, sb & 1);  // this tile's P (TMA writes are visible after the wait)
          // pass 1: delta = sum_j P_j * dP_j  (own row: P from the swizzled tile, dP from TMEM)
          float delta = 0.f;
          for (int cb = 0; cb < p.nk; cb += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_row + cb, v);
            tmem_ld_wait();
            const uint8_t* chunk = p_smem + (cb >> 6) * kSraTileBytes + row * 128;
            const uint32_t u0 = (uint32_t)((cb & 63) >> 3);
            const uint4 w0 = *reinterpret_cast<const uint4*>(chunk + ((u0 ^ rsw) << 4));
            const uint4 w1 = *reinterpret_cast<const uint4*>(chunk + (((u0 + 1) ^ rsw) << 4));
            const uint32_t pw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 pp = sra_unpack2<FMT>(pw[i]);
              delta = fmaf(pp.x, __uint_as_float(v[2 * i]), delta);
              delta = fmaf(pp.y, __uint_as_float(v[2 * i + 1]), delta);
            }
          }
          // pass 2: dS in place
          for (int cb = 0; cb < p.nk; cb += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_row + cb, v);
            tmem_ld_wait();
            uint8_t* chunk = p_smem + (cb >> 6) * kSraTileBytes + row * 128;
            const uint32_t u0 = (uint32_t)((cb & 63) >> 3);
            uint4* q0 = reinterpret_cast<uint4*>(chunk + ((u0 ^ rsw) << 4));
            uint4* q1 = reinterpret_cast<uint4*>(chunk + (((u0 + 1) ^ rsw) << 4));
            const uint4 w0 = *q0, w1 = *q1;
            const uint32_t pw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            uint32_t dw[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 pp = sra_unpack2<FMT>(pw[i]);
              dw[i] = sra_pack2<FMT>(p.scale * pp.x * (__uint_as_float(v[2 * i]) - delta),
                                     p.scale * pp.y * (__uint_as_float(v[2 * i + 1]) - delta));
            }
            *q0 = make_uint4(dw[0], dw[1], dw[2], dw[3]);
            *q1 = make_uint4(dw[4], dw[5], dw[6], dw[7]);
          }
          l_run = 1.f;
          tc_fence_before();
          fence_proxy_async_smem();
          mbar_arrive(&p_full);
          mbar_wait(&o_full, sb & 1);
          tc_fence_after();
#pragma unroll
          for (int cb = 0; cb < kSraD; cb += 16) {
            uint32_t v2[16];
            tmem_ld_32x32b_x16(t_row + kSraColO + cb, v2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o_run[cb + i] = __uint_as_float(v2[i]);
          }
          tc_fence_before();
          continue;
        }
        // pass 1: block maximum
        float mx = m_run;
        for (int cb = 0; cb < valid; cb += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_row + cb, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (cb + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float alpha = exp2f((m_run - mx) * p.scale_log2e);  // 0 for the first block (m_run = -inf)
        m_run = mx;
        if (j == 0) {
          // the TMA stores of the previous tile must have finished reading the P / O staging tiles before they are rewritten
          if (issuer) bulk_wait_group_read<0>();
          named_bar_sync(1, 128);
        }
        // pass 2: p~ = exp2(scale_log2e * (s - max)), block sum, 16-bit P~ into the swizzled A operand of the second MMA
        float sum = 0.f;
        const float moff = mx * p.scale_log2e;
        for (int cb = 0; cb < p.kb; cb += 16) {
          float e[16];
          if (cb < valid) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(t_row + cb, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              e[i] = cb + i < valid ? exp2f(fmaf(__uint_as_float(v[i]), p.scale_log2e, -moff)) : 0.f;
              sum += e[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = 0.f;
          }
          uint8_t* chunk = p_smem + (cb >> 6) * kSraTileBytes + row * 128;
          const uint32_t u0 = (uint32_t)((cb & 63) >> 3);
          *reinterpret_cast<uint4*>(chunk + ((u0 ^ rsw) << 4)) =
              make_uint4(sra_pack2<FMT>(e[0], e[1]), sra_pack2<FMT>(e[2], e[3]), sra_pack2<FMT>(e[4], e[5]), sra_pack2<FMT>(e[6], e[7]));
          *reinterpret_cast<uint4*>(chunk + (((u0 + 1) ^ rsw) << 4)) =
              make_uint4(sra_pack2<FMT>(e[8], e[9]), sra_pack2<FMT>(e[10], e[11]), sra_pack2<FMT>(e[12], e[13]), sra_pack2<FMT>(e[14], e[15]));
        }
        l_run = l_run * alpha + sum;
        tc_fence_before();          // S reads of this thread are complete before the arrive
        fence_proxy_async_smem();   // generic-proxy writes of P~ visible to the tensor core (async proxy)
        mbar_arrive(&p_full);
        // O_blk = P~.V_blk: accumulate in registers with the running rescale
        mbar_wait(&o_full, sb & 1);
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < kSraD; cb += 16) {
          uint32_t v[16];
          tmem_ld_32x32b_x16(t_row + kSraColO + cb, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o_run[cb + i] = j == 0 ? __uint_as_float(v[i]) : fmaf(o_run[cb + i], alpha, __uint_as_float(v[i]));
        }
        tc_fence_before();  // O_blk has been read: the next second MMA (issued after the next p_full) may overwrite it
      }
      const float inv = 1.f / l_run;  // backward: l_run = 1
      uint8_t* orow = o_smem + row * 128;
#pragma unroll
      for (int cb = 0; cb < kSraD; cb += 16) {
        const uint32_t u0 = (uint32_t)(cb >> 3);
        *reinterpret_cast<uint4*>(orow + ((u0 ^ rsw) << 4)) =
            make_uint4(sra_pack2<FMT>(o_run[cb] * inv, o_run[cb + 1] * inv), sra_pack2<FMT>(o_run[cb + 2] * inv, o_run[cb + 3] * inv),
                       sra_pack2<FMT>(o_run[cb + 4] * inv, o_run[cb + 5] * inv), sra_pack2<FMT>(o_run[cb + 6] * inv, o_run[cb + 7] * inv));
        *reinterpret_cast<uint4*>(orow + (((u0 + 1) ^ rsw) << 4)) =
            make_uint4(sra_pack2<FMT>(o_run[cb + 8] * inv, o_run[cb + 9] * inv), sra_pack2<FMT>(o_run[cb + 10] * inv, o_run[cb + 11] * inv),
                       sra_pack2<FMT>(o_run[cb + 12] * inv, o_run[cb + 13] * inv), sra_pack2<FMT>(o_run[cb + 14] * inv, o_run[cb + 15] * inv));
      }
      if (p.save_p && !BWD) {
        // one key block: normalise this thread's own row of P~ in place (the second MMA has completed: o_full) — the saved
        // probabilities of the backward
        for (int ch = 0; ch < p.nk; ch += 64) {
          uint8_t* chunk = p_smem + (ch >> 6) * kSraTileBytes + row * 128;
#pragma unroll
          for (uint32_t u = 0; u < 8; ++u) {
            uint4* ptr = reinterpret_cast<uint4*>(chunk + ((u ^ rsw) << 4));
            uint4 w = *ptr;
            float2 a = sra_unpack2<FMT>(w.x), bb = sra_unpack2<FMT>(w.y), cc = sra_unpack2<FMT>(w.z), dd = sra_unpack2<FMT>(w.w);
            w.x = sra_pack2<FMT>(a.x * inv, a.y * inv);
            w.y = sra_pack2<FMT>(bb.x * inv, bb.y * inv);
            w.z = sra_pack2<FMT>(cc.x * inv, cc.y * inv);
            w.w = sra_pack2<FMT>(dd.x * inv, dd.y * inv);
            *ptr = w;
          }
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (issuer) {
        tma_store_4d(&p.tmO, o_smem, g * kSraD, qt * 128, 0, b);  // rows past the image's last query are clipped
        if (p.save_p || BWD)  // forward: normalised P; backward: dS
          for (int ch = 0; ch < p.nk; ch += 64)
            tma_store_4d(&p.tmP, p_smem + (ch >> 6) * kSraTileBytes, g * p.lp + ch, qt * 128, 0, b);
        bulk_commit_group();
      }
    }
    if (issuer) bulk_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// q / k / v / o: pointers to head 0's 64 columns of row 0 of image 0; rows of one image are ld* elements apart, images nq (nk) rows
// apart; head g sits 64*g columns to the right.  cols_* = columns of the tensor from that pointer on (the TMA bound of the maps).
static int sra_launch(const void* q, long long ldq, int cols_q, const void* k, long long ldk, const void* v, long long ldv, int cols_kv,
                      void* o, long long ldo, void* p_out, long long ldp, int B, int nq, int nk, int heads, float scale, int dtype,
                      cudaStream_t s, const void* pin = nullptr, long long ldpin = 0) {
  SraParams p;
  memset(&p, 0, sizeof(p));
  p.B = B;
  p.heads = heads;
  p.qtiles = (nq + 127) / 128;
  p.nk = nk;
  if (nk <= kSraMaxKeys && nk % 16 == 0) {
    p.kb = nk;
    p.nblocks = 1;
  } else {
    p.kb = kSraStreamKeys;
    p.nblocks = (nk + kSraStreamKeys - 1) / kSraStreamKeys;
  }
  p.lp = nk;
  p.save_p = p_out != nullptr;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.items = B * heads * p.qtiles;
  const int sms = g_opt_sra_max_ctas > 0 ? g_opt_sra_max_ctas : device_sm_count();
  const int grid = p.items < sms ? p.items : sms;
  p.per_cta = (p.items + grid - 1) / grid;
  int st = make_tmap_nhwc(&p.tmQ, q, dtype, cols_q, nq, 1, B, ldq, kSraD, 128, 1, 128);
  if (st) return st;
  st = make_tmap_nhwc(&p.tmK, k, dtype, cols_kv, nk, 1, B, ldk, kSraD, p.kb, 1, 128);
  if (st) return st;
  st = make_tmap_nhwc(&p.tmV, v, dtype, cols_kv, nk, 1, B, ldv, kSraD, p.kb, 1, 128);
  if (st) return st;
  st = make_tmap_nhwc(&p.tmO, o, dtype, 64 * heads, nq, 1, B, ldo, kSraD, 128, 1, 128);
  if (st) return st;
  if (p.save_p) {
    st = make_tmap_nhwc(&p.tmP, p_out, dtype, (long long)heads * nk, nq, 1, B, ldp, 64, 128, 1, 128);
    if (st) return st;
  }
  if (pin != nullptr) {
    st = make_tmap_nhwc(&p.tmPin, pin, dtype, (long long)heads * nk, nq, 1, B, ldpin, 64, 128, 1, 128);
    if (st) return st;
    p.scale = scale;
  }
  const int smem = kSraSmem + 1024;
  // trailing CTAs may have no item (per_cta rounding): they only allocate and free their TMEM columns
#define GDL_SRA_LAUNCH(FMT, BWD)                                                            \
  do {                                                                                      \
    static PerDeviceOnce once;                                                              \
    GDL_CHECK_CUDA(set_max_dyn_smem_once(once, sra_attention_kernel<FMT, BWD>, smem));      \
    GDL_LAUNCH((sra_attention_kernel<FMT, BWD>), grid, kSraThreads, smem, s, p);                      \
  } while (0)
  if (pin != nullptr) {
    if (dtype == GDL_BF16) GDL_SRA_LAUNCH(1, true); else GDL_SRA_LAUNCH(0, true);
  } else {
    if (dtype == GDL_BF16) GDL_SRA_LAUNCH(1, false); else GDL_SRA_LAUNCH(0, false);
  }
#undef GDL_SRA_LAUNCH
  GDL_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace gdl

using namespace gdl;

extern "C" int gdl_sra_attention_fwd(const void* q, long long ldq, const void* kv, long long ldkv, void* o, long long ldo,
                                     void* p_out, long long ldp, int B, int N, int heads, int nk, int c, float scale,
                                     int dtype, void* stream) {
  GDL_REQUIRE(q && kv && o && B > 0 && N > 0 && heads > 0 && nk > 0, GDL_ERR_INVALID, "sra_attention: bad args");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_INVALID, "sra_attention: 16-bit dtype expected");
  GDL_REQUIRE(c == heads * kSraD, GDL_ERR_UNSUPPORTED, "sra_attention: head dim %d (64 expected)", heads ? c / heads : 0);
  // saving the probabilities needs the whole key range in one block and whole 64-key store boxes (a box must not spill into the
  // next head's columns); ragged query tiles are zero-filled / clipped by the TMA unit.  Without P (inference) any shape goes: more
  // than 256 keys are streamed
  GDL_REQUIRE(p_out == nullptr || (nk % 64 == 0 && nk <= kSraMaxKeys), GDL_ERR_UNSUPPORTED,
              "sra_attention: %d keys (saving P needs a multiple of 64, at most %d: use the three-kernel path otherwise)", nk, kSraMaxKeys);

  GDL_REQUIRE(ldq >= c && ldo >= c && ldkv >= 2 * c && (p_out == nullptr || ldp >= (long long)heads * nk), GDL_ERR_INVALID,
              "sra_attention: leading dimensions too small");
  const char* kvp = reinterpret_cast<const char*>(kv);
  return sra_launch(q, ldq, c, kvp, ldkv, kvp + (size_t)c * 2, ldkv, c, o, ldo, p_out, ldp, B, N, nk, heads, scale, dtype,
                    (cudaStream_t)stream);
}

extern "C" int gdl_mha_flash_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                                 long long ldo, int B, int N, int heads, float scale, int dtype, void* stream) {
  GDL_REQUIRE(q && k && v && o && B > 0 && N > 0 && heads > 0, GDL_ERR_INVALID, "mha_flash: bad args");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_INVALID, "mha_flash: 16-bit dtype expected");
  const int c = heads * kSraD;
  GDL_REQUIRE(ldq >= c && ldk >= c && ldv >= c && ldo >= c, GDL_ERR_INVALID, "mha_flash: leading dimensions too small (head dim 64)");
  return sra_launch(q, ldq, c, k, ldk, v, ldv, c, o, ldo, nullptr, 0, B, N, N, heads, scale, dtype, (cudaStream_t)stream);
}

extern "C" int gdl_sra_attention_bwd(const void* d_o, long long lddo, const void* kv, long long ldkv, const void* p_saved, long long ldp,
                                     void* dq, long long lddq, void* ds, long long ldds, int B, int N, int heads, int nk, int c,
                                     float scale, int dtype, void* stream) {
  GDL_REQUIRE(d_o && kv && p_saved && dq && ds && B > 0 && N > 0 && heads > 0 && nk > 0, GDL_ERR_INVALID, "sra_attention_bwd: bad args");
  GDL_REQUIRE(dtype == GDL_BF16 || dtype == GDL_F16, GDL_ERR_INVALID, "sra_attention_bwd: 16-bit dtype expected");
  GDL_REQUIRE(c == heads * kSraD, GDL_ERR_UNSUPPORTED, "sra_attention_bwd: head dim %d (64 expected)", heads ? c / heads : 0);
  GDL_REQUIRE(nk % 64 == 0 && nk <= kSraMaxKeys, GDL_ERR_UNSUPPORTED, "sra_attention_bwd: %d keys (a multiple of 64 up to %d)", nk,
              kSraMaxKeys);
  GDL_REQUIRE(lddo >= c && lddq >= c && ldkv >= 2 * c && ldp >= (long long)heads * nk && ldds >= (long long)heads * nk, GDL_ERR_INVALID,
              "sra_attention_bwd: leading dimensions too small");
  GDL_REQUIRE(p_saved != ds, GDL_ERR_INVALID, "sra_attention_bwd: dS must not alias P (dV = P^T.dO still reads P)");
  const char* kvp = reinterpret_cast<const char*>(kv);
  // roles: Q <- dO, first-MMA operand <- V, second-MMA operand <- K, O <- dQ, stored tile <- dS
  return sra_launch(d_o, lddo, c, kvp + (size_t)c * 2, ldkv, kvp, ldkv, c, dq, lddq, ds, ldds, B, N, nk, heads, scale, dtype,
                    (cudaStream_t)stream, p_saved, ldp);
}
