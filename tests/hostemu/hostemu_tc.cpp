// hostemu_tc.cpp — TEST INFRASTRUCTURE: functional model of TMA / mbarrier / TMEM / tcgen05.mma (see hostemu_tc.h).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <deque>
#include <functional>
#include <memory>
#include <random>
#include <vector>

#include "cuda_hostemu.h"

namespace hostemu {
namespace tc {

#define TC_FAIL(...)                          \
  do {                                        \
    fprintf(stderr, "hostemu/tc: " __VA_ARGS__); \
    fprintf(stderr, "\n");                    \
    abort();                                  \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// shared-memory window: 32-bit shared addresses are offsets into the 1024-byte aligned arena that holds the dynamic shared
// memory (the arena base stands for shared address 0; dyn_smem() starts a little above it, as after static allocations)
// ---------------------------------------------------------------------------------------------------------------------
static uint8_t* arena() { return reinterpret_cast<uint8_t*>(hostemu::dyn_smem()) - hostemu::dyn_smem_offset(); }
constexpr uint32_t kArenaBytes = 256 * 1024;

uint32_t smem_u32(const void* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p), b = reinterpret_cast<uintptr_t>(arena());
  if (a >= b && a < b + kArenaBytes) return (uint32_t)(a - b);
  return 0xf0000000u | (uint32_t)(a & 0x0ffffff8u);  // a static __shared__ object (mbarrier, flag): only ever printed
}
static uint8_t* smem_ptr(uint32_t addr, uint32_t bytes, const char* what) {
  if (addr + bytes > kArenaBytes) TC_FAIL("%s: shared address %u (+%u) outside the dynamic shared-memory window", what, addr, bytes);
  return arena() + addr;
}
static inline uint32_t swizzle(uint32_t addr, int span) {  // Swizzle<B,4,3> on the absolute shared address
  switch (span) {
    case 128: return addr ^ (((addr >> 7) & 7u) << 4);
    case 64: return addr ^ (((addr >> 7) & 3u) << 4);
    case 32: return addr ^ (((addr >> 7) & 1u) << 4);
    default: return addr;
  }
}

bool elect_one() { return hostemu::lane() == 0; }

// ---------------------------------------------------------------------------------------------------------------------
// static counters of the modelled units (tools/hostemu_counters.py): what one launch asks of TMA, the tensor core and TMEM
// ---------------------------------------------------------------------------------------------------------------------
enum { kCtrTmaLoadBytes, kCtrTmaStoreBytes, kCtrTmaLoads, kCtrMmaIssued, kCtrMmaMacs, kCtrTmemLd, kCtrMbarWaitsFailed, kCtrRedAdds, kNumCtr };
static unsigned long long g_ctr[kNumCtr];

// ---------------------------------------------------------------------------------------------------------------------
// asynchronous mode (GDL_HOSTEMU_ASYNC=<seed>): TMA copies, tensor-core operations and commits do NOT complete at issue.  They
// are queued and completed later, at random scheduler passes: TMA loads in any order, tcgen05 operations in issue order
// (commit after the MMAs before it), TMA stores in order.  Operands are read and results written AT COMPLETION, so a missing
// mbarrier wait / fence-less reuse of a shared-memory stage, a TMEM accumulator or a staging tile computes with stale or
// clobbered data — the class of bug the synchronous model cannot see.
// ---------------------------------------------------------------------------------------------------------------------
static int g_async = -1;
static bool g_early_read = false;  // even seeds: MMAs / TMA stores read shared memory at issue; odd seeds: at completion
static std::mt19937 g_rng;
static std::vector<std::function<void()>> g_q_tma;   // completes in any order
static std::deque<std::function<void()>> g_q_tc;     // in order
static std::deque<std::function<void()>> g_q_store;  // in order; entries of one bulk group end with a marker (empty function)
static int g_store_groups = 0;                        // committed bulk groups still pending
static size_t g_max_tma = 0, g_max_tc = 0, g_max_store = 0, g_deferred = 0;  // GDL_HOSTEMU_ASYNC_STATS=1: printed at exit
struct AsyncStats {
  ~AsyncStats() {
    if (getenv("GDL_HOSTEMU_ASYNC_STATS"))
      fprintf(stderr, "hostemu async: %zu deferred operations, deepest queues: %zu TMA loads, %zu tensor-core ops, %zu store entries\n",
              g_deferred, g_max_tma, g_max_tc, g_max_store);
  }
} g_async_stats;
static bool async_on() {
  if (g_async < 0) {
    const char* e = getenv("GDL_HOSTEMU_ASYNC");
    g_async = (e && *e && atoi(e) >= 0) ? 1 : 0;
    g_rng.seed(e ? (unsigned)atoi(e) : 0u);
    g_early_read = e && (atoi(e) % 2 == 0);
  }
  return g_async == 1;
}
static void run_store_front() {
  auto fn = std::move(g_q_store.front());
  g_q_store.pop_front();
  if (fn) fn(); else --g_store_groups;
}
// called by the fiber scheduler after every pass; `stuck`: no fiber could make progress on its own
bool async_tick(bool stuck) {
  if (!async_on()) return false;
  bool did = false;
  g_max_tma = std::max(g_max_tma, g_q_tma.size());
  g_max_tc = std::max(g_max_tc, g_q_tc.size());
  g_max_store = std::max(g_max_store, g_q_store.size());
  g_deferred += g_q_tma.size() + g_q_tc.size() + g_q_store.size();
  auto run_tma = [&](size_t i) {
    auto fn = std::move(g_q_tma[i]);
    g_q_tma[i] = std::move(g_q_tma.back());
    g_q_tma.pop_back();
    fn();
    did = true;
  };
  auto run_tc = [&]() {
    auto fn = std::move(g_q_tc.front());
    g_q_tc.pop_front();
    fn();
    did = true;
  };
  for (size_t i = 0; i < g_q_tma.size();) {
    if ((g_rng() & 7) == 0) run_tma(i); else ++i;
  }
  while (!g_q_tc.empty() && (g_rng() & 3) == 0) run_tc();
  while (!g_q_store.empty() && (g_rng() & 31) == 0) {  // stores linger: only wait_group[.read] (or luck) completes them early
    run_store_front();
    did = true;
  }
  if (stuck && !did) {
    // nothing can move on its own: complete ONE pending operation, chosen at random (not the oldest: completion order is free)
    const int nq = (!g_q_tma.empty()) + (!g_q_tc.empty()) + (!g_q_store.empty());
    if (nq > 0) {
      int pick = (int)(g_rng() % (unsigned)nq);
      if (!g_q_tma.empty() && pick-- == 0) {
        run_tma(g_rng() % g_q_tma.size());
      } else if (!g_q_tc.empty() && pick-- == 0) {
        run_tc();
      } else {
        run_store_front();
        did = true;
      }
    }
  }
  if (did) hostemu::note_progress();
  return did || !g_q_tma.empty() || !g_q_tc.empty() || !g_q_store.empty();
}
static void async_drain_all() {
  while (!g_q_tma.empty() || !g_q_tc.empty() || !g_q_store.empty()) {
    for (auto& fn : g_q_tma) fn();
    g_q_tma.clear();
    while (!g_q_tc.empty()) {
      auto fn = std::move(g_q_tc.front());
      g_q_tc.pop_front();
      fn();
    }
    while (!g_q_store.empty()) run_store_front();
  }
}
// cp.async.bulk.commit_group / wait_group[.read] N (executed by the thread that issued the stores)
void bulk_commit_group() {
  if (!async_on()) return;
  g_q_store.emplace_back();  // group marker
  ++g_store_groups;
}
void bulk_wait_group(int n) {
  if (!async_on()) return;
  while (g_store_groups > n) run_store_front();
}

// ---------------------------------------------------------------------------------------------------------------------
// mbarrier: the 8-byte object itself holds the state
// ---------------------------------------------------------------------------------------------------------------------
struct MBar {
  int32_t tx;        // outstanding transaction bytes (may go negative transiently)
  uint16_t pending;  // arrivals still expected in the current phase
  uint16_t init_phase;  // bit 15 = phase parity, bits 0..14 = arrival count of a phase
};
static_assert(sizeof(MBar) == 8, "mbarrier state fits the 64-bit object");
static MBar* mb(uint64_t* bar) { return reinterpret_cast<MBar*>(bar); }
static void mbar_check(MBar* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->init_phase ^= 0x8000u;
    b->pending = b->init_phase & 0x7fffu;
  }
  hostemu::note_progress();
}
void mbar_init(uint64_t* bar, uint32_t count) {
  if (count == 0 || count > 0x7fff) TC_FAIL("mbarrier.init: count %u", count);
  MBar* b = mb(bar);
  b->tx = 0;
  b->pending = (uint16_t)count;
  b->init_phase = (uint16_t)count;
  hostemu::note_progress();
}
static void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  MBar* b = mb(bar);
  b->tx -= (int32_t)bytes;
  mbar_check(b);
}
void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {  // mbarrier.arrive.expect_tx
  MBar* b = mb(bar);
  if (b->pending == 0) TC_FAIL("mbarrier.arrive.expect_tx on a barrier with no pending arrival (thread %d)", hostemu::thread_linear());
  b->tx += (int32_t)bytes;
  b->pending -= 1;
  mbar_check(b);
}
void mbar_arrive(uint64_t* bar) {
  MBar* b = mb(bar);
  if (b->pending == 0) TC_FAIL("mbarrier.arrive on a barrier with no pending arrival (thread %d)", hostemu::thread_linear());
  b->pending -= 1;
  mbar_check(b);
}
uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  MBar* b = mb(bar);
  const uint32_t phase = (b->init_phase >> 15) & 1u;
  if (phase != (parity & 1u)) return 1;  // the phase with this parity has completed
  ++g_ctr[kCtrMbarWaitsFailed];
  hostemu::yield_spin();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// tensor maps + TMA tile copies
// ---------------------------------------------------------------------------------------------------------------------
struct TMap {
  uint32_t magic, rank;
  uint8_t* base;
  uint32_t esz, swz;
  uint32_t dims[5], box[5];
  uint64_t strides[4];  // bytes, dims 1..rank-1
};
static_assert(sizeof(TMap) <= sizeof(CUtensorMap), "model fits the opaque tensor map");
constexpr uint32_t kMagic = 0x544d4150u;

static CUresult encode_tiled(CUtensorMap* out, CUtensorMapDataType dt, cuuint32_t rank, void* base, const cuuint64_t* dims,
                             const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave il,
                             CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
  TMap t;
  memset(&t, 0, sizeof(t));
  t.magic = kMagic;
  t.rank = rank;
  t.base = reinterpret_cast<uint8_t*>(base);
  t.esz = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4 : 2;
  t.swz = sw == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : sw == CU_TENSOR_MAP_SWIZZLE_64B ? 64 : sw == CU_TENSOR_MAP_SWIZZLE_32B ? 32 : 0;
  // the driver's documented requirements
  if (rank < 1 || rank > 5 || il != CU_TENSOR_MAP_INTERLEAVE_NONE) return CUDA_ERROR_INVALID_VALUE;
  if (reinterpret_cast<uintptr_t>(base) & 15) return CUDA_ERROR_INVALID_VALUE;
  for (uint32_t i = 0; i < rank; ++i) {
    if (dims[i] == 0 || dims[i] > (1ull << 32) || box[i] == 0 || box[i] > 256 || estr[i] != 1) return CUDA_ERROR_INVALID_VALUE;
    t.dims[i] = (uint32_t)dims[i];
    t.box[i] = box[i];
    if (i + 1 < rank) {
      if (strides[i] % 16 != 0 || strides[i] >= (1ull << 40)) return CUDA_ERROR_INVALID_VALUE;
      t.strides[i] = strides[i];
    }
  }
  const uint32_t inner = t.box[0] * t.esz;
  if (inner % 16 != 0) return CUDA_ERROR_INVALID_VALUE;
  if (t.swz != 0 && inner > t.swz) return CUDA_ERROR_INVALID_VALUE;
  memset(out, 0, sizeof(*out));
  memcpy(out, &t, sizeof(t));
  return CUDA_SUCCESS;
}
void* encode_entry_point() { return reinterpret_cast<void*>(&encode_tiled); }

static const TMap& tmap(const CUtensorMap* m) {
  const TMap& t = *reinterpret_cast<const TMap*>(m);
  if (t.magic != kMagic) TC_FAIL("TMA with an un-encoded tensor map");
  return t;
}

// global byte offset of box element (i0..i4) or -1 when out of bounds
static inline long long g_off(const TMap& t, const int* c, const int* i) {
  long long off = 0;
  for (uint32_t d = 0; d < t.rank; ++d) {
    const long long x = (long long)c[d] + i[d];
    if (x < 0 || x >= (long long)t.dims[d]) return -1;
    off += d == 0 ? x * t.esz : x * (long long)t.strides[d - 1];
  }
  return off;
}

static void tma_copy(const TMap& t, uint32_t saddr, const int* c, bool load) {
  if (saddr & 127) TC_FAIL("TMA: shared address %u is not 128-byte aligned", saddr);
  if (t.swz == 128 && false) {}
  uint32_t nbox = 1;
  for (uint32_t d = 0; d < t.rank; ++d) nbox *= t.box[d];
  smem_ptr(saddr, nbox * t.esz, "TMA");
  const uint32_t inner_bytes = t.box[0] * t.esz;
  int i[5] = {0, 0, 0, 0, 0};
  uint32_t lin = 0;  // dense byte offset inside the box
  const uint32_t b1 = t.rank > 1 ? t.box[1] : 1, b2 = t.rank > 2 ? t.box[2] : 1, b3 = t.rank > 3 ? t.box[3] : 1, b4 = t.rank > 4 ? t.box[4] : 1;
  for (i[4] = 0; i[4] < (int)b4; ++i[4])
    for (i[3] = 0; i[3] < (int)b3; ++i[3])
      for (i[2] = 0; i[2] < (int)b2; ++i[2])
        for (i[1] = 0; i[1] < (int)b1; ++i[1], lin += inner_bytes)
          for (uint32_t ch = 0; ch < inner_bytes; ch += 16) {  // the swizzle permutes 16-byte chunks
            uint8_t* s = arena() + swizzle(saddr + lin + ch, (int)t.swz);
            for (uint32_t e = 0; e < 16; e += t.esz) {
              i[0] = (int)((ch + e) / t.esz);
              const long long go = g_off(t, c, i);
              if (load) {
                if (go < 0)
                  memset(s + e, 0, t.esz);
                else
                  memcpy(s + e, t.base + go, t.esz);
              } else if (go >= 0) {
                memcpy(t.base + go, s + e, t.esz);
              }
            }
          }
}

void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  const TMap& t = tmap(m);
  if (t.rank != 2) TC_FAIL("tma_load_2d with a rank-%u map", t.rank);
  g_ctr[kCtrTmaLoadBytes] += t.box[0] * t.box[1] * t.esz;
  ++g_ctr[kCtrTmaLoads];
  const uint32_t saddr = smem_u32(smem_dst), bytes = t.box[0] * t.box[1] * t.esz;
  auto fn = [t, saddr, c0, c1, bar, bytes]() {
    const int c[5] = {c0, c1, 0, 0, 0};
    tma_copy(t, saddr, c, true);
    mbar_complete_tx(bar, bytes);
  };
  if (async_on()) g_q_tma.emplace_back(fn); else fn();
}
void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  const TMap& t = tmap(m);
  if (t.rank != 4) TC_FAIL("tma_load_4d with a rank-%u map", t.rank);
  g_ctr[kCtrTmaLoadBytes] += t.box[0] * t.box[1] * t.box[2] * t.box[3] * t.esz;
  ++g_ctr[kCtrTmaLoads];
  const uint32_t saddr = smem_u32(smem_dst), bytes = t.box[0] * t.box[1] * t.box[2] * t.box[3] * t.esz;
  auto fn = [t, saddr, c0, c1, c2, c3, bar, bytes]() {
    const int c[5] = {c0, c1, c2, c3, 0};
    tma_copy(t, saddr, c, true);
    mbar_complete_tx(bar, bytes);
  };
  if (async_on()) g_q_tma.emplace_back(fn); else fn();
}
void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  const TMap& t = tmap(m);
  if (t.rank != 4) TC_FAIL("tma_store_4d with a rank-%u map", t.rank);
  g_ctr[kCtrTmaStoreBytes] += t.box[0] * t.box[1] * t.box[2] * t.box[3] * t.esz;
  const uint32_t saddr = smem_u32(smem_src);
  auto fn = [t, saddr, c0, c1, c2, c3]() {
    const int c[5] = {c0, c1, c2, c3, 0};
    tma_copy(t, saddr, c, false);
  };
  // asynchronous mode, odd seeds: the tile is read at completion (a staging tile rewritten before wait_group.read shows)
  if (async_on() && !g_early_read) g_q_store.emplace_back(fn); else fn();
  hostemu::note_progress();
}

void named_bar_sync(int id, int nthreads) {
  if (id < 1 || id > 15 || nthreads % 32 != 0) TC_FAIL("bar.sync %d, %d", id, nthreads);
  hostemu::sync_named(id, nthreads);
}

// ---------------------------------------------------------------------------------------------------------------------
// TMEM
// ---------------------------------------------------------------------------------------------------------------------
static float g_tmem[128][512];
static uint32_t g_tmem_used = 0;      // bump allocator (the kernels allocate once per block)
static uint32_t g_tmem_live = 0;      // columns currently allocated

void block_begin() {
  g_tmem_used = g_tmem_live = 0;
  for (auto& row : g_tmem)
    for (float& v : row) v = __builtin_nanf("");  // reading an accumulator nobody wrote is a bug
}
void block_end(unsigned bx, unsigned by, unsigned bz) {
  async_drain_all();  // a CTA's outstanding bulk operations complete before its resources are released
  if (g_tmem_live != 0) TC_FAIL("block (%u,%u,%u) exits with %u TMEM columns still allocated", bx, by, bz, g_tmem_live);
}

void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // .sync.aligned: the whole warp executes it, one allocation
  if (hostemu::lane() != 0) return;
  if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) TC_FAIL("tcgen05.alloc: %u columns (power of two in 32..512)", ncols);
  if (g_tmem_used + ncols > 512) TC_FAIL("tcgen05.alloc: %u + %u columns exceed TMEM", g_tmem_used, ncols);
  *smem_result = g_tmem_used;  // lane 0 in bits 31..16, column in bits 15..0
  g_tmem_used += ncols;
  g_tmem_live += ncols;
  hostemu::note_progress();
}
void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (hostemu::lane() != 0) return;
  if ((taddr >> 16) != 0 || (taddr & 0xffff) + ncols > 512 || ncols > g_tmem_live) TC_FAIL("tcgen05.dealloc(%#x, %u)", taddr, ncols);
  g_tmem_live -= ncols;
}

void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  const uint32_t lane_base = taddr >> 16, col = taddr & 0xffff;
  // a warp may only touch the TMEM lane quarter (warp index % 4)
  if (lane_base != 32u * (uint32_t)(hostemu::warp_index() & 3))
    TC_FAIL("tcgen05.ld: warp %d addresses TMEM lanes %u.. (allowed: %d..)", hostemu::warp_index(), lane_base, 32 * (hostemu::warp_index() & 3));
  if (col + 16 > 512) TC_FAIL("tcgen05.ld: columns %u..%u", col, col + 15);
  if (hostemu::lane() == 0) ++g_ctr[kCtrTmemLd];
  const float* row = g_tmem[lane_base + (uint32_t)hostemu::lane()];
  for (int i = 0; i < 16; ++i) memcpy(&v[i], &row[col + i], 4);
}

// ---------------------------------------------------------------------------------------------------------------------
// tcgen05.mma kind::f16, cta_group::1: D[M x N] (+)= A[M x 16] . B[N x 16]^T, operands from shared-memory descriptors
// ---------------------------------------------------------------------------------------------------------------------
struct Desc {
  uint32_t start, lbo, sbo;
  int span;  // swizzle span in bytes
};
static Desc decode_desc(uint64_t d, const char* which) {
  Desc r;
  r.start = (uint32_t)(d & 0x3fff) << 4;
  r.lbo = (uint32_t)((d >> 16) & 0x3fff) << 4;
  r.sbo = (uint32_t)((d >> 32) & 0x3fff) << 4;
  if (((d >> 46) & 3) != 1) TC_FAIL("%s descriptor: version field %llu (sm_100 expects 1)", which, (unsigned long long)((d >> 46) & 3));
  const uint32_t lt = (uint32_t)(d >> 61) & 7;
  r.span = lt == 2 ? 128 : lt == 4 ? 64 : lt == 6 ? 32 : -1;
  if (r.span < 0) TC_FAIL("%s descriptor: layout type %u not modelled (swizzled layouts only)", which, lt);
  return r;
}
static inline float load16(uint32_t addr, int fmt) {
  const uint8_t* p = arena() + addr;
  if (fmt == 1) {
    __nv_bfloat16 h;
    memcpy(&h, p, 2);
    return __bfloat162float(h);
  }
  __half h;
  memcpy(&h, p, 2);
  return __half2float(h);
}
// element (r, k) of an operand tile with `rows` rows (M or N) and 16 k-elements
static inline uint32_t elem_addr(const Desc& d, int mn_major, int r, int k) {
  const uint32_t span = (uint32_t)d.span, per = span / 2;  // 16-bit elements per swizzle row
  uint32_t off;
  if (!mn_major)  // K-major: 8-row groups SBO apart, rows `span` bytes apart, the 16 k-elements contiguous
    off = (uint32_t)(r >> 3) * d.sbo + (uint32_t)(r & 7) * span + (uint32_t)k * 2;
  else  // MN-major: `per` contiguous MN elements, MN blocks LBO apart; 8 k-rows `span` bytes apart, k groups SBO apart
    off = ((uint32_t)r / per) * d.lbo + ((uint32_t)r % per) * 2 + (uint32_t)(k >> 3) * d.sbo + (uint32_t)(k & 7) * span;
  return swizzle(d.start + off, d.span);
}

// operands gathered from shared memory (through the descriptors) into dense tiles; the product accumulated into TMEM
struct MmaTiles {
  int M, N;
  uint32_t col0, accumulate;
  std::vector<float> A, B;  // [M][16], [N][16]
};
static void umma_gather(MmaTiles& t, uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  const int N = (int)((idesc >> 17) & 0x3f) << 3, M = (int)((idesc >> 24) & 0x1f) << 4;
  const int afmt = (idesc >> 7) & 7, bfmt = (idesc >> 10) & 7, a_mn = (idesc >> 15) & 1, b_mn = (idesc >> 16) & 1;
  if (((idesc >> 4) & 3) != 1) TC_FAIL("tcgen05.mma: accumulator format %u (f32 expected)", (idesc >> 4) & 3);
  if (afmt > 1 || bfmt > 1) TC_FAIL("tcgen05.mma kind::f16: operand formats %d / %d", afmt, bfmt);
  if (!(M == 128 && N % 16 == 0 && N >= 16 && N <= 256) && !(M == 64 && N % 8 == 0 && N >= 8 && N <= 256))
    TC_FAIL("tcgen05.mma: illegal shape M %d N %d", M, N);
  if (M != 128) TC_FAIL("tcgen05.mma: M = %d accumulator layout not modelled", M);
  const uint32_t lane0 = tmem_d >> 16, col0 = tmem_d & 0xffff;
  if (lane0 != 0 || col0 + (uint32_t)N > 512) TC_FAIL("tcgen05.mma: accumulator at lane %u, columns %u..%u", lane0, col0, col0 + N - 1);
  const Desc da = decode_desc(desc_a, "A"), db = decode_desc(desc_b, "B");
  t.M = M;
  t.N = N;
  t.col0 = col0;
  t.accumulate = accumulate;
  t.A.resize((size_t)M * 16);
  t.B.resize((size_t)N * 16);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < 16; ++k) {
      const uint32_t ad = elem_addr(da, a_mn, m, k);
      if (ad + 2 > kArenaBytes) TC_FAIL("tcgen05.mma: A operand reads shared address %u", ad);
      t.A[(size_t)m * 16 + k] = load16(ad, afmt);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < 16; ++k) {
      const uint32_t ad = elem_addr(db, b_mn, n, k);
      if (ad + 2 > kArenaBytes) TC_FAIL("tcgen05.mma: B operand reads shared address %u", ad);
      t.B[(size_t)n * 16 + k] = load16(ad, bfmt);
    }
}
static void umma_accumulate(const MmaTiles& t) {
  for (int m = 0; m < t.M; ++m) {
    float* drow = &g_tmem[m][t.col0];
    const float* a = &t.A[(size_t)m * 16];
    for (int n = 0; n < t.N; ++n) {
      const float* b = &t.B[(size_t)n * 16];
      float acc = 0.f;
      for (int k = 0; k < 16; ++k) acc += a[k] * b[k];
      drow[n] = t.accumulate ? drow[n] + acc : acc;
    }
  }
  hostemu::note_progress();
}
void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  ++g_ctr[kCtrMmaIssued];
  g_ctr[kCtrMmaMacs] += (unsigned long long)(((idesc >> 24) & 0x1f) << 4) * (((idesc >> 17) & 0x3f) << 3) * 16;
  if (!async_on()) {
    static MmaTiles t;
    umma_gather(t, tmem_d, desc_a, desc_b, idesc, accumulate);
    umma_accumulate(t);
  } else if (g_early_read) {
    // operands read AT ISSUE (a missing wait for the TMA load / the P~ writes shows), accumulator written at completion
    auto t = std::make_shared<MmaTiles>();
    umma_gather(*t, tmem_d, desc_a, desc_b, idesc, accumulate);
    g_q_tc.emplace_back([t]() { umma_accumulate(*t); });
  } else {
    // operands read AT COMPLETION (a stage / tile reused before the commit was observed shows)
    g_q_tc.emplace_back([=]() {
      MmaTiles t;
      umma_gather(t, tmem_d, desc_a, desc_b, idesc, accumulate);
      umma_accumulate(t);
    });
  }
}
// arrives when every tensor-core operation issued before it has completed (in-order queue; synchronous mode: at once)
void umma_commit(uint64_t* bar) {
  if (async_on())
    g_q_tc.emplace_back([bar]() { mbar_arrive(bar); });
  else
    mbar_arrive(bar);
}

void red_add_f32(float* dst, float v) {
  ++g_ctr[kCtrRedAdds];
  *dst += v;
}
void check_align(const void* p, unsigned bytes, const char* what) {
  if (reinterpret_cast<uintptr_t>(p) % bytes) TC_FAIL("%s: address %p is not %u-byte aligned (misaligned address fault on the GPU)", what, p, bytes);
}

}  // namespace tc
}  // namespace hostemu

extern "C" void hostemu_counters_reset(void) { memset(hostemu::tc::g_ctr, 0, sizeof(hostemu::tc::g_ctr)); }
// out[8]: TMA load bytes, TMA store bytes, TMA load instructions, tcgen05.mma issued, MACs, tcgen05.ld (per warp), failed mbarrier polls,
// fp32 global reductions
extern "C" void hostemu_counters_get(unsigned long long* out) { memcpy(out, hostemu::tc::g_ctr, sizeof(hostemu::tc::g_ctr)); }
// seed < 0: synchronous completion; even seed: asynchronous, operands read at issue; odd seed: asynchronous, operands read at completion
extern "C" void hostemu_set_async(int seed) {
  hostemu::tc::g_async = seed >= 0 ? 1 : 0;
  hostemu::tc::g_rng.seed((unsigned)(seed >= 0 ? seed : 0));
  hostemu::tc::g_early_read = seed >= 0 && seed % 2 == 0;
}
