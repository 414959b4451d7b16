// TEST INFRASTRUCTURE — a small CUDA emulator: enough of the device language and of the runtime API to
// build ALL of pysparselp_b200/csrc (kernels, setup, stats, graphs) with g++ and run it on the CPU.
// tests/emul/make_emul.py rewrites the `kernel<<<grid, block, smem, stream>>>(args)` launches into
// emul::launch(grid, block, [=] { kernel(args); }) and compiles the result against this header.
//
// Execution model: blocks run one after the other; the threads of a block are fibers that are resumed
// round-robin, so __syncthreads() and the warp shuffles are real rendezvous between the 256 "threads" of a
// CTA.  On x86-64 a fiber switch is a dozen instructions (emul_switch in emul_nccl.cpp: callee-saved
// registers and the stack pointer; glibc's swapcontext costs a sigprocmask system call per switch, which
// dominated the run time of the emulated tests); elsewhere it falls back to ucontext.  Streams are synchronous, events are wall-clock, a captured graph is the recorded
// list of launches.  All emulator state is thread_local: one OS thread plays one rank ("GPU"), peer memory
// is the shared address space, and tests/emul/emul_nccl.cpp supplies an in-process NCCL.
//
// Nothing in the product includes or loads this; the product has no CPU path.
#pragma once
#include <chrono>
#include <sched.h>
#include <ucontext.h>

#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

// ---------------------------------------------------------------------------------- device language
struct EmulDim3 { unsigned x = 1, y = 1, z = 1; };
namespace emul {
inline thread_local EmulDim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
}
#define threadIdx emul::g_threadIdx
#define blockIdx emul::g_blockIdx
#define blockDim emul::g_blockDim
#define gridDim emul::g_gridDim

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static thread_local

namespace emul {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 64 * 1024;

#if defined(__x86_64__) && !defined(CPPPD_EMUL_UCONTEXT)
#define CPPPD_EMUL_FAST_SWITCH 1
// saves the callee-saved state on the current stack, stores the stack pointer in *save_sp, continues on load_sp
extern "C" void emul_switch(void **save_sp, void *load_sp);
struct Context {
  void *sp = nullptr;
};
inline void switch_context(Context &from, Context &to) { emul_switch(&from.sp, to.sp); }
// a context that starts in `entry` (which never returns) on the given stack
inline void make_context(Context &c, char *stack, size_t bytes, void (*entry)()) {
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~(uintptr_t)15;
  void **sp = reinterpret_cast<void **>(top - 72);  // [mxcsr|fpcw] r15 r14 r13 r12 rbx rbp <return address> <pad>
  memset(sp, 0, 72);
  const uint32_t mxcsr = 0x1F80;
  const uint16_t fpcw = 0x037F;
  memcpy(reinterpret_cast<char *>(sp), &mxcsr, 4);
  memcpy(reinterpret_cast<char *>(sp) + 4, &fpcw, 2);
  sp[7] = reinterpret_cast<void *>(entry);  // `ret` of emul_switch jumps here with rsp = top - 8 (as after a call)
  c.sp = sp;
}
#else
struct Context {
  ucontext_t uc;
};
inline void switch_context(Context &from, Context &to) { swapcontext(&from.uc, &to.uc); }
inline void make_context(Context &c, char *stack, size_t bytes, void (*entry)()) {
  getcontext(&c.uc);
  c.uc.uc_stack.ss_sp = stack;
  c.uc.uc_stack.ss_size = bytes;
  c.uc.uc_link = nullptr;
  makecontext(&c.uc, entry, 0);
}
#endif

struct Fiber {
  Context ctx;
  bool done = true;
};

struct Rendezvous {  // all live participants must arrive before anyone leaves
  unsigned long long gen = 0;
  int arrived = 0;
};

struct BlockState {
  int nthreads = 0, live = 0;
  int warp_live[kMaxThreads / 32];
  Rendezvous bar, warp_bar[kMaxThreads / 32];
  unsigned char slot[kMaxThreads][8];
  Fiber fibers[kMaxThreads];
  Context main_ctx;
  const std::function<void()> *body = nullptr;
  int current = -1;
  char *stacks = nullptr;
  bool progress = false;
};
inline thread_local BlockState g_block;

inline void yield() { switch_context(g_block.fibers[g_block.current].ctx, g_block.main_ctx); }

inline void release_if_complete(Rendezvous &r, int live) {
  if (live > 0 && r.arrived >= live) {
    r.arrived = 0;
    r.gen++;
    g_block.progress = true;
  }
}

inline void arrive_and_wait(Rendezvous &r, int &live) {
  const unsigned long long my_gen = r.gen;
  r.arrived++;
  release_if_complete(r, live);
  while (r.gen == my_gen) yield();
}

inline void fiber_entry() {
  BlockState &b = g_block;
  (*b.body)();
  const int t = b.current;
  b.fibers[t].done = true;
  b.live--;
  b.warp_live[t >> 5]--;
  b.progress = true;
  release_if_complete(b.bar, b.live);  // threads that exited do not take part in later barriers
  release_if_complete(b.warp_bar[t >> 5], b.warp_live[t >> 5]);
  switch_context(b.fibers[t].ctx, b.main_ctx);
  abort();  // a finished fiber is never resumed
}

inline void run_block(int nthreads, const std::function<void()> &body) {
  BlockState &b = g_block;
  if (!b.stacks) b.stacks = static_cast<char *>(malloc(kStackBytes * kMaxThreads));
  b.nthreads = b.live = nthreads;
  b.body = &body;
  b.bar = Rendezvous();
  for (int w = 0; w < kMaxThreads / 32; ++w) {
    b.warp_bar[w] = Rendezvous();
    b.warp_live[w] = 0;
  }
  for (int t = 0; t < nthreads; ++t) {
    b.warp_live[t >> 5]++;
    Fiber &f = b.fibers[t];
    f.done = false;
    make_context(f.ctx, b.stacks + kStackBytes * t, kStackBytes, fiber_entry);
  }
  while (b.live > 0) {
    b.progress = false;
    for (int t = 0; t < nthreads; ++t) {
      if (b.fibers[t].done) continue;
      b.current = t;
      g_threadIdx.x = (unsigned)t;
      switch_context(b.main_ctx, b.fibers[t].ctx);
    }
    if (!b.progress && b.live > 0) {
      fprintf(stderr, "emul: dead-lock inside a block (%d threads waiting)\n", b.live);
      abort();
    }
  }
}

// ---- graphs: a capture records the launches instead of running them
struct Graph { std::vector<std::function<void()>> nodes; };
inline thread_local Graph *g_capturing = nullptr;

inline void run_grid(unsigned grid, unsigned block, const std::function<void()> &body) {
  g_gridDim.x = grid;
  g_blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b) {
    g_blockIdx.x = b;
    run_block((int)block, body);
  }
}

template <typename F>
inline void launch(unsigned grid, unsigned block, F body) {
  std::function<void()> fn = body;
  if (g_capturing) g_capturing->nodes.push_back([=] { run_grid(grid, block, fn); });
  else run_grid(grid, block, fn);
}

}  // namespace emul

inline void __syncthreads() { emul::g_block.progress = true; emul::arrive_and_wait(emul::g_block.bar, emul::g_block.live); }
inline void __syncwarp(unsigned = 0xffffffffu) {
  const int w = emul::g_block.current >> 5;
  emul::arrive_and_wait(emul::g_block.warp_bar[w], emul::g_block.warp_live[w]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emul::BlockState &b = emul::g_block;
  const int t = b.current, w = t >> 5;
  memcpy(b.slot[t], &v, sizeof(T));
  emul::arrive_and_wait(b.warp_bar[w], b.warp_live[w]);
  T out = v;
  const int src = (t & ~31) | ((t ^ lane_mask) & 31);
  if (src < b.nthreads && !b.fibers[src].done) memcpy(&out, b.slot[src], sizeof(T));
  emul::arrive_and_wait(b.warp_bar[w], b.warp_live[w]);
  return out;
}
// value of lane `src_lane` / of the lane `delta` below (own value when there is none)
template <typename T>
inline T emul_shfl_from(T v, int src_lane_or_delta, bool up) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emul::BlockState &b = emul::g_block;
  const int t = b.current, w = t >> 5, lane = t & 31;
  memcpy(b.slot[t], &v, sizeof(T));
  emul::arrive_and_wait(b.warp_bar[w], b.warp_live[w]);
  T out = v;
  const int sl = up ? lane - src_lane_or_delta : (src_lane_or_delta & 31);
  if (sl >= 0) {
    const int src = (t & ~31) | sl;
    if (src < b.nthreads && !b.fibers[src].done) memcpy(&out, b.slot[src], sizeof(T));
  }
  emul::arrive_and_wait(b.warp_bar[w], b.warp_live[w]);
  return out;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src_lane) { return emul_shfl_from(v, src_lane, false); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) { return emul_shfl_from(v, (int)delta, true); }
inline unsigned __ballot_sync(unsigned, int pred) {
  emul::BlockState &b = emul::g_block;
  const int t = b.current, w = t >> 5;
  const int v = pred ? 1 : 0;
  memcpy(b.slot[t], &v, sizeof v);
  emul::arrive_and_wait(b.warp_bar[w], b.warp_live[w]);
  unsigned out = 0;
  for (int l = 0; l < 32; ++l) {
    const int src = (t & ~31) | l;
    int sv = 0;
    if (src < b.nthreads && !b.fibers[src].done) memcpy(&sv, b.slot[src], sizeof sv);
    if (sv) out |= 1u << l;
  }
  emul::arrive_and_wait(b.warp_bar[w], b.warp_live[w]);
  return out;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __any_sync(unsigned, int pred) {
  int v = pred ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
// a spinning thread waits for another *rank* (another OS thread): not a dead-lock of this block
inline void __nanosleep(unsigned) {
  emul::g_block.progress = true;
  sched_yield();
  emul::yield();
}

struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
template <typename T> inline T __ldg(const T *p) { return *p; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldca(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
template <typename T> inline void __stcs(T *p, T v) { *p = v; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline long long __double_as_longlong(double v) { long long o; memcpy(&o, &v, 8); return o; }

template <typename T> inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }
inline unsigned atomicOr(unsigned *p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p = o | v; return o; }
using std::max;
using std::min;
inline long long max(long long a, long b) { return a > b ? a : (long long)b; }

// ---------------------------------------------------------------------------------- runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmul = 1 };
typedef struct EmulStream *cudaStream_t;
struct EmulEvent { std::chrono::steady_clock::time_point t; };
typedef EmulEvent *cudaEvent_t;
typedef emul::Graph *cudaGraph_t;
typedef emul::Graph *cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyHostToHost };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaDevAttrMultiProcessorCount = 16,
       cudaIpcMemLazyEnablePeerAccess = 1 };
struct cudaIpcMemHandle_t { char reserved[64]; };

inline const char *cudaGetErrorString(cudaError_t e) { return e ? "emulated CUDA error" : "no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 4; return cudaSuccess; }  // 4 "SMs"
enum { cudaLimitMaxL2FetchGranularity = 5 };
inline cudaError_t cudaDeviceSetLimit(int, size_t) { return cudaSuccess; }
// Device memory is NOT zero-initialised on a GPU (and torch's caching allocator hands out recycled blocks), while
// a fresh malloc usually is: every allocation is filled with a poison byte so that a kernel relying on zeros fails
// here too (0x5A: doubles become 5.6e129, int32 indices 1.5e9).  CPPPD_EMUL_POISON=<byte> changes it, -1 disables.
inline int poison_byte() {
  static const int v = [] { const char *e = getenv("CPPPD_EMUL_POISON"); return e ? atoi(e) : 0x5A; }();
  return v;
}
inline cudaError_t cudaMalloc(void **p, size_t bytes) {
  *p = malloc(bytes ? bytes : 1);
  if (*p && poison_byte() >= 0) memset(*p, poison_byte(), bytes ? bytes : 1);
  return *p ? cudaSuccess : cudaErrorEmul;
}
template <typename T> inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <typename T> inline cudaError_t cudaMallocHost(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmulEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { emul::g_capturing = new emul::Graph(); return cudaSuccess; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t *g) { *g = emul::g_capturing; emul::g_capturing = nullptr; return cudaSuccess; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t *e, cudaGraph_t g, int) { *e = new emul::Graph(*g); return cudaSuccess; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t g) { delete g; return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t g, cudaStream_t) { for (auto &n : g->nodes) n(); return cudaSuccess; }
// all emulated ranks are threads of one process: an IPC handle is the pointer itself
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

// device-wide nanosecond clock (%globaltimer on the GPU)
inline unsigned long long global_timer_ns() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}
