// TEST INFRASTRUCTURE -- a tiny host-side CUDA emulator (not a fallback, never loaded by the product path).
//
// Lets the SIMT translation units of pixel-nerf_b200/csrc (stage kernels, SIMT field engine, field backward and the
// C-ABI glue) be compiled with g++ and executed on the CPU so that kernel indexing / orchestration bugs show up in the
// `-m "not gpu"` suite.  A kernel launch runs the blocks one after another; the threads of a block are fibers on one
// OS thread, so __syncthreads / __syncwarp / __shfl_* have their CUDA meaning (barriers + an exchange buffer) and a
// run is deterministic.
// tests/cuda_emu/build_emu.py rewrites `k<<<grid, block, smem, stream>>>(args);` into emu::launch(...) calls.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <type_traits>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
using std::max;
using std::min;

typedef void* cudaStream_t;
typedef struct EmuEvent* cudaEvent_t;
enum cudaError_t { cudaSuccess = 0, cudaErrorUnknown = 1 };
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulator"; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
// ---- what csrc/pnr_mgpu.cu uses: ONE emulated device, every "peer" is the host heap, all work is synchronous -------
enum { cudaMemcpyDefault = 4, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaErrorPeerAccessAlreadyEnabled = 704 };
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d >= 0 && d < 64 ? cudaSuccess : cudaErrorUnknown; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int a, int b) { *can = (a != b); return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void* d, int, const void* s, int, size_t n, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, int, cudaStream_t) {
  for (size_t r = 0; r < h; ++r) memcpy(static_cast<char*>(d) + r * dp, static_cast<const char*>(s) + r * sp, w);
  return cudaSuccess;
}

namespace emu {
// ---- cooperative threads as fibers on ONE OS thread (x86-64): deterministic, no futex storms ----------------------
extern "C" void emu_switch(void** save_sp, void* load_sp);
#if defined(__x86_64__)
asm(R"(
.text
.weak emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");
#else
#error "tests/cuda_emu needs x86-64"
#endif

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;            // from the process-wide pool below (allocated once, never zeroed)
  bool done = false;
  uint3 tid{0, 0, 0};
  int lin = 0;
};

struct Block {
  std::vector<Fiber> fibers;
  std::vector<uint32_t> xbuf;      // [warp][32] shuffle exchange
  std::vector<double> dyn;         // dynamic shared memory (extern __shared__), 8-byte aligned
  std::vector<int> warp_live, warp_arrived;
  std::vector<unsigned> warp_gen;
  int live = 0, arrived = 0;
  unsigned gen = 0;
  void* sched_sp = nullptr;
  Fiber* cur = nullptr;
};

inline uint3 tIdx, bIdx;
inline dim3 bDim, gDim;
inline Block* blk = nullptr;
inline int lin_tid = 0;
inline const void* body_ptr = nullptr;
inline void (*body_call)(const void*) = nullptr;

inline void yield() {
  Fiber* me = blk->cur;
  emu_switch(&me->sp, blk->sched_sp);
  tIdx = me->tid;                   // resumed: this fiber's ids are the current ones again
  lin_tid = me->lin;
}

inline void fiber_main() {
  body_call(body_ptr);
  Block* b = blk;
  Fiber* me = b->cur;
  me->done = true;
  // an exited thread no longer takes part in barriers: release waiters it was the last one missing for
  const int w = me->lin / 32;
  b->live--;
  b->warp_live[w]--;
  if (b->live > 0 && b->arrived == b->live) { b->arrived = 0; b->gen++; }
  if (b->warp_live[w] > 0 && b->warp_arrived[w] == b->warp_live[w]) { b->warp_arrived[w] = 0; b->warp_gen[w]++; }
  emu_switch(&me->sp, b->sched_sp);
  __builtin_unreachable();
}

inline void block_barrier() {
  Block* b = blk;
  if (++b->arrived == b->live) { b->arrived = 0; b->gen++; return; }
  const unsigned g = b->gen;
  while (b->gen == g) yield();
}

inline void warp_barrier() {
  Block* b = blk;
  const int w = lin_tid / 32;
  if (++b->warp_arrived[w] == b->warp_live[w]) { b->warp_arrived[w] = 0; b->warp_gen[w]++; return; }
  const unsigned g = b->warp_gen[w];
  while (b->warp_gen[w] == g) yield();
}

enum Mode { SEQ = 0, WARP = 1, BLOCK = 2 };   // what the kernel uses (build_emu.py reads it off the source)

template <typename T, typename = std::enable_if_t<std::is_arithmetic<T>::value>>
inline dim3 to_dim3(T v) { return dim3((unsigned)v); }
inline dim3 to_dim3(dim3 d) { return d; }

// Blocks run one after another (static __shared__ arrays are per-kernel statics).  Kernels without barriers or
// shuffles (SEQ) are a plain loop over the threads; the others run their threads as fibers, round-robin, each until
// its next barrier.
template <int MODE, class F>
void launch(dim3 g, dim3 b, size_t smem, F f) {
  const int nt = (int)(b.x * b.y * b.z);
  const size_t nblocks = (size_t)g.x * g.y * g.z;
  if (nt == 0 || nblocks == 0) return;
  bDim = b;
  gDim = g;
  Block block;
  block.dyn.assign((smem + 7) / 8 + 1, 0.0);
  blk = &block;
  auto ids = [&](size_t bi) {
    bIdx = uint3{(unsigned)(bi % g.x), (unsigned)((bi / g.x) % g.y), (unsigned)(bi / ((size_t)g.x * g.y))};
  };
  if (MODE == SEQ) {
    for (size_t bi = 0; bi < nblocks; ++bi) {
      ids(bi);
      for (int t = 0; t < nt; ++t) {
        tIdx = uint3{(unsigned)t % b.x, ((unsigned)t / b.x) % b.y, (unsigned)t / (b.x * b.y)};
        lin_tid = t;
        f();
      }
    }
    return;
  }
  const int nw = (nt + 31) / 32;
  const size_t STACK = 256 * 1024;
  static std::vector<std::unique_ptr<char[]>> pool;
  while ((int)pool.size() < nt) pool.emplace_back(new char[STACK]);
  block.fibers.resize(nt);
  for (int t = 0; t < nt; ++t) block.fibers[t].stack = pool[t].get();
  block.xbuf.assign((size_t)nw * 32, 0u);
  body_ptr = &f;
  body_call = [](const void* p) { (*static_cast<const F*>(p))(); };
  for (size_t bi = 0; bi < nblocks; ++bi) {
    ids(bi);
    block.live = nt;
    block.arrived = 0;
    block.warp_live.assign(nw, 0);
    block.warp_arrived.assign(nw, 0);
    block.warp_gen.assign(nw, 0u);
    for (int t = 0; t < nt; ++t) {
      Fiber& fb = block.fibers[t];
      fb.done = false;
      fb.tid = uint3{(unsigned)t % b.x, ((unsigned)t / b.x) % b.y, (unsigned)t / (b.x * b.y)};
      fb.lin = t;
      block.warp_live[t / 32]++;
      // initial frame: six callee-saved registers, then the entry address; 16-byte aligned at function entry
      uintptr_t top = (reinterpret_cast<uintptr_t>(fb.stack) + STACK) & ~uintptr_t(15);
      void** sp = reinterpret_cast<void**>(top - 8);     // after `ret`, rsp % 16 == 8 like after a call
      *--sp = reinterpret_cast<void*>(&fiber_main);
      for (int i = 0; i < 6; ++i) *--sp = nullptr;
      fb.sp = sp;
    }
    int remaining = nt;
    while (remaining > 0) {
      for (int t = 0; t < nt; ++t) {
        Fiber& fb = block.fibers[t];
        if (fb.done) continue;
        block.cur = &fb;
        tIdx = fb.tid;
        lin_tid = fb.lin;
        emu_switch(&block.sched_sp, fb.sp);
        if (fb.done) --remaining;
      }
    }
  }
}
}  // namespace emu

#define threadIdx emu::tIdx
#define blockIdx emu::bIdx
#define blockDim emu::bDim
#define gridDim emu::gDim

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) {
  static_assert(sizeof(T) == 4, "4-byte shuffles only");
  const int w = emu::lin_tid / 32, lane = emu::lin_tid % 32;
  uint32_t* buf = emu::blk->xbuf.data() + (size_t)w * 32;
  memcpy(&buf[lane], &v, 4);
  emu::warp_barrier();
  T r;
  memcpy(&r, &buf[src & 31], 4);
  emu::warp_barrier();
  return r;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned m, T v, int lane_mask, int = 32) {
  return __shfl_sync(m, v, (emu::lin_tid % 32) ^ lane_mask);
}

static inline float atomicAdd(float* p, float v) {   // one OS thread: plain read-modify-write
  const float old = *p;
  *p = old + v;
  return old;
}

template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline int __float2int_rz(float x) { return (int)x; }
