// hostemu_runtime.cpp — TEST INFRASTRUCTURE: the fiber scheduler behind cuda_hostemu.h and host stand-ins for the few
// CUDA runtime calls the C-ABI wrappers make (memset / memcpy act on host memory; there is no device).
#include <ucontext.h>

#include <vector>

#include "cuda_hostemu.h"

namespace hostemu {

enum { kRun = 0, kWaitBlock = 1, kWaitWarp = 2, kDone = 3 };
constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  int state;
  uint3 tid;
  int warp, lane;
};

static std::vector<Fiber> g_fibers;
static std::vector<char*> g_stacks;
static std::vector<uint64_t> g_slots;  // 32 per warp
static ucontext_t g_main;
static Fiber* g_cur = nullptr;
static const std::function<void()>* g_body = nullptr;
alignas(1024) static unsigned char g_smem[256 * 1024];

void* dyn_smem() { return g_smem; }
int lane() { return g_cur->lane; }
uint64_t* warp_slots() { return g_slots.data() + (size_t)g_cur->warp * 32; }

static void yield_as(int state) {
  Fiber* f = g_cur;
  f->state = state;
  swapcontext(&f->ctx, &g_main);
}
void sync_block() { yield_as(kWaitBlock); }
void sync_warp() { yield_as(kWaitWarp); }

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
  if (nthreads <= 0 || nthreads > 1024 || smem > sizeof(g_smem)) {
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
          f.tid = make_uint3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
          f.warp = t / 32;
          f.lane = t % 32;
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = g_stacks[t];
          f.ctx.uc_stack.ss_size = kStackBytes;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, fiber_entry, 0);
        }
        for (;;) {
          bool ran = false;
          for (int t = 0; t < nthreads; ++t) {
            Fiber& f = g_fibers[t];
            if (f.state != kRun) continue;
            g_cur = &f;
            threadIdx = f.tid;
            swapcontext(&g_main, &f.ctx);
            ran = true;
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
          if (!released) {
            fprintf(stderr, "hostemu: block (%u,%u,%u) stuck (%s): %d live fibers, %d at __syncthreads, the rest in a partial warp barrier\n",
                    bx, by, bz, ran ? "divergent barrier" : "nothing runnable", live, at_block);
            abort();
          }
        }
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
  *fn = nullptr;  // no driver: tensor maps (TMA kernels) are not emulated
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
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
