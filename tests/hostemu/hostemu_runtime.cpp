// hostemu_runtime.cpp — TEST INFRASTRUCTURE: the fiber scheduler behind cuda_hostemu.h and host stand-ins for the few
// CUDA runtime calls the C-ABI wrappers make (memset / memcpy act on host memory; there is no device).
#include <ucontext.h>

#include <vector>

#include "cuda_hostemu.h"

namespace hostemu {

enum { kRun = 0, kWaitBlock = 1, kWaitWarp = 2, kDone = 3, kWaitNamed = 4 };
constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  int state;
  uint3 tid;
  int warp, lane;
  int bar_id, bar_n;  // named barrier (bar.sync id, n) the fiber waits at
  bool spun;          // last yield was a failed mbarrier poll
};

static std::vector<Fiber> g_fibers;
static std::vector<char*> g_stacks;
static std::vector<uint64_t> g_slots;  // 32 per warp
static ucontext_t g_main;
static Fiber* g_cur = nullptr;
static const std::function<void()>* g_body = nullptr;
alignas(1024) static unsigned char g_smem[256 * 1024];

// dynamic shared memory starts 16-byte (not 1024-byte) aligned, a little above shared address 0, as behind static allocations
constexpr size_t kDynOffset = 1040;
void* dyn_smem() { return g_smem + kDynOffset; }
size_t dyn_smem_offset() { return kDynOffset; }
int lane() { return g_cur->lane; }
uint64_t* warp_slots() { return g_slots.data() + (size_t)g_cur->warp * 32; }

static void yield_as(int state) {
  Fiber* f = g_cur;
  f->state = state;
  swapcontext(&f->ctx, &g_main);
}
void sync_block() { yield_as(kWaitBlock); }
void sync_warp() { yield_as(kWaitWarp); }
void sync_named(int id, int n) {
  g_cur->bar_id = id;
  g_cur->bar_n = n;
  yield_as(kWaitNamed);
}
static unsigned long long g_progress = 0;
void note_progress() { ++g_progress; }
void yield_spin() {
  g_cur->spun = true;
  yield_as(kRun);
}
int warp_index() { return g_cur->warp; }
int thread_linear() { return g_cur->warp * 32 + g_cur->lane; }
namespace tc {
bool async_tick(bool stuck);
void* encode_entry_point();
void block_begin();
void block_end(unsigned bx, unsigned by, unsigned bz);
}

[[noreturn]] void unsupported_asm() {
  fprintf(stderr, "hostemu: inline PTX reached (tensor-core / TMA code is not emulated)\n");
  abort();
}

static void fiber_entry() {
  (*g_body)();
  yield_as(kDone);
  abort();  // a finished fiber is never resumed
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads <= 0 || nthreads > 1024 || smem + kDynOffset > sizeof(g_smem)) {
    fprintf(stderr, "hostemu: bad launch (%d threads, %zu B shared)\n", nthreads, smem);
    abort();
  }
  if ((int)g_fibers.size() < nthreads) g_fibers.resize(nthreads);
  while ((int)g_stacks.size() < nthreads) g_stacks.push_back((char*)malloc(kStackBytes));
  const int nwarps = (nthreads + 31) / 32;
  g_slots.assign((size_t)nwarps * 32, 0);
  g_body = &body;
  gridDim = grid;
  blockDim = block;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        blockIdx = make_uint3(bx, by, bz);
        for (int t = 0; t < nthreads; ++t) {
          Fiber& f = g_fibers[t];
          f.state = kRun;
          f.spun = false;
          f.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
          f.warp = t / 32;
          f.lane = t % 32;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = g_stacks[t];
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, fiber_entry, 0);
        }
        tc::block_begin();
        int idle_passes = 0;
        for (;;) {
          bool ran = false;
          const unsigned long long progress0 = g_progress;
          int spinning = 0;
          for (int t = 0; t < nthreads; ++t) {
            Fiber& f = g_fibers[t];
            if (f.state != kRun) continue;
            g_cur = &f;
            threadIdx = f.tid;
            f.spun = false;
            swapcontext(&g_main, &f.ctx);
            ran = true;
            if (f.state == kDone) ++g_progress;
            if (f.state == kRun && f.spun) ++spinning;
          }
          // every fiber is now waiting or done: open the barriers whose live participants have all arrived
          int live = 0, at_block = 0;
          for (int t = 0; t < nthreads; ++t) {
            live += g_fibers[t].state != kDone;
            at_block += g_fibers[t].state == kWaitBlock;
          }
          if (live == 0) break;
          bool released = false;
          if (at_block == live) {
            for (int t = 0; t < nthreads; ++t)
              if (g_fibers[t].state == kWaitBlock) g_fibers[t].state = kRun;
            released = true;
          }
          for (int w = 0; w < nwarps && at_block != live; ++w) {
            int wl = 0, ww = 0;
            for (int t = w * 32; t < nthreads && t < w * 32 + 32; ++t) {
              wl += g_fibers[t].state != kDone;
              ww += g_fibers[t].state == kWaitWarp;
            }
            if (ww > 0 && ww == wl) {
              for (int t = w * 32; t < nthreads && t < w * 32 + 32; ++t) g_fibers[t].state = g_fibers[t].state == kWaitWarp ? kRun : g_fibers[t].state;
              released = true;
            }
          }
          for (int id = 0; id < 16; ++id) {  // bar.sync id, n
            int cnt = 0, need = 0;
            for (int t = 0; t < nthreads; ++t)
              if (g_fibers[t].state == kWaitNamed && g_fibers[t].bar_id == id) {
                ++cnt;
                need = g_fibers[t].bar_n;
              }
            if (cnt > 0 && cnt >= need) {
              for (int t = 0; t < nthreads; ++t)
                if (g_fibers[t].state == kWaitNamed && g_fibers[t].bar_id == id) g_fibers[t].state = kRun;
              released = true;
            }
          }
          if (released) ++g_progress;
          // asynchronous mode: complete some of the queued TMA / tensor-core operations (all fibers are blocked or polling now)
          const bool pending = tc::async_tick(g_progress == progress0);
          idle_passes = (g_progress == progress0) ? idle_passes + 1 : 0;
          if (pending && idle_passes <= 64 && (released || spinning > 0 || g_progress != progress0)) continue;
          if ((!released && spinning == 0 && g_progress == progress0) || idle_passes > 4) {
            fprintf(stderr, "hostemu: block (%u,%u,%u) stuck (%s): %d live fibers, %d at __syncthreads, %d polling an mbarrier\n",
                    bx, by, bz, spinning ? "mbarrier deadlock" : (ran ? "divergent barrier" : "nothing runnable"), live, at_block, spinning);
            for (int t = 0; t < nthreads; ++t)
              if (g_fibers[t].state != kDone)
                fprintf(stderr, "  thread %d: %s\n", t, g_fibers[t].state == kRun ? "polling" : g_fibers[t].state == kWaitBlock ? "__syncthreads"
                        : g_fibers[t].state == kWaitWarp ? "warp barrier" : "named barrier");
            abort();
          }
        }
        tc::block_end(bx, by, bz);
      }
  g_body = nullptr;
  g_cur = nullptr;
}

}  // namespace hostemu

// ---- CUDA runtime stand-ins (host memory, no device) -----------------------------------------------------------------
extern "C" {
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaPeekAtLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "hostemu"; }
const char* cudaGetErrorName(cudaError_t) { return "hostemu"; }
cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, enum cudaDeviceAttr a, int) {
  *v = a == cudaDevAttrMultiProcessorCount ? 148 : 0;
  return cudaSuccess;
}
cudaError_t cudaFuncSetAttribute(const void*, enum cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {
  memset(p, v, n);
  return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, enum cudaMemcpyKind, cudaStream_t) {
  memmove(d, s, n);
  return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, unsigned long long, enum cudaDriverEntryPointQueryResult* q) {
  *fn = hostemu::tc::encode_entry_point();  // cuTensorMapEncodeTiled stand-in of the functional model
  if (q) *q = cudaDriverEntryPointSuccess;
  return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp* p, int) {  // the header maps the name to its _v2 symbol
  memset(p, 0, sizeof(*p));
  p->multiProcessorCount = 148;
  p->major = 10;
  return cudaSuccess;
}
}
