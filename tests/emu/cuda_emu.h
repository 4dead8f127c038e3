// cuda_emu.h — lock-step CPU emulation of the CUDA constructs dm_control_b200/csrc/b200mj.cu uses.
//
// TEST INFRASTRUCTURE ONLY. It lets `pytest -m "not gpu"` compile the *kernel source itself* with g++
// (-DB200MJ_CPU_EMU) and run it for a handful of environments, so that logic errors in the CUDA path (indexing, layout
// aliasing, launch orchestration, the C ABI) surface on a machine without a GPU. It says nothing about performance or
// about data races, it is never built by __graft_entry__.build(), and nothing under dm_control_b200/ can load it.
//
// Model: every CUDA thread of a block is a fiber (own stack, hand-rolled x86-64 context switch) and the fibers of one
// block run cooperatively on one OS thread; blocks are distributed over a few OS threads, each with its own
// "shared memory". A warp collective (__shfl*_sync, __ballot_sync, __any_sync, __syncwarp) or a CTA barrier
// (__syncthreads*) is a rendezvous: a fiber that arrives yields round-robin until all expected participants have
// arrived, so every lane observes exactly the values its peers published at that call — the semantics the kernels rely
// on (full-mask collectives in warp-uniform control flow only). Fibers that return from the kernel leave the
// barriers, as exited CUDA threads do. Scheduling is deterministic.
#pragma once
#if !defined(__x86_64__)
#error "the emulation's context switch is written for x86-64"
#endif
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ thread_local

using std::max;
using std::min;

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct alignas(16) double2 { double x, y; };

// ---- fibers -------------------------------------------------------------------------------------------------------
asm(R"(
.text
.p2align 4
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
extern "C" __attribute__((visibility("hidden"))) void emu_switch(void** save_sp, void* load_sp);

struct emu_bar { unsigned expected = 0, arrived = 0, gen = 0; };
struct emu_warp { emu_bar bar; uint64_t slot[32]; };
struct emu_fiber { void* sp = nullptr; emu_dim3 tid, bid; int lane = 0; emu_warp* warp = nullptr; bool done = false; };
struct emu_block {
  std::vector<emu_fiber> f; std::vector<emu_warp> warps; emu_bar bar; int vote = 0; int cur = 0; int live = 0;
  void* sched_sp = nullptr; const std::function<void()>* body = nullptr;
};
enum { EMU_STACK = 256 * 1024 };

inline thread_local emu_block* emu_blk = nullptr;
inline thread_local emu_fiber* emu_cur = nullptr;
inline emu_dim3 blockDim, gridDim;
#define threadIdx (emu_cur->tid)
#define blockIdx (emu_cur->bid)

alignas(16) inline thread_local double smem[232448 / 8];     // dynamic shared memory of the block this OS thread runs

// Scheduling order of the fibers of a block: ascending thread index by default, descending with B200MJ_EMU_ORDER=reverse.
// Code that is correct under CUDA's memory model (every cross-lane dependency through shared memory separated by a
// collective or a barrier) gives the same results under both; a missing __syncwarp() shows up as a difference,
// because only one of the two orders happens to run the producer before the consumer.
inline int emu_dir() { static const int d = [] { const char* e = std::getenv("B200MJ_EMU_ORDER"); return (e && e[0] == 'r') ? -1 : 1; }(); return d; }
inline int emu_next(const emu_block* b, int i) {
  const int n = (int)b->f.size(), d = emu_dir();
  do { i += d; if (i == n) i = 0; else if (i < 0) i = n - 1; } while (b->f[i].done);
  return i;
}
inline void emu_yield() {
  emu_block* b = emu_blk;
  int i = emu_next(b, b->cur);
  if (i == b->cur) return;
  emu_fiber* from = &b->f[b->cur];
  b->cur = i; emu_cur = &b->f[i];
  emu_switch(&from->sp, b->f[i].sp);
}
inline void emu_wait(emu_bar& bar) {
  const unsigned gen = bar.gen;
  if (++bar.arrived >= bar.expected) { bar.arrived = 0; bar.gen++; return; }
  while (bar.gen == gen) emu_yield();
}
inline void emu_drop(emu_bar& bar) {       // a participant leaves for good
  bar.expected--;
  if (bar.expected > 0 && bar.arrived >= bar.expected) { bar.arrived = 0; bar.gen++; }
}
inline void emu_fiber_entry() {
  (*emu_blk->body)();
  emu_block* b = emu_blk;
  emu_fiber* me = emu_cur;
  me->done = true;
  emu_drop(me->warp->bar); emu_drop(b->bar);
  void* dummy;
  if (--b->live == 0) emu_switch(&dummy, b->sched_sp);       // last one out returns to the block runner
  int i = emu_next(b, b->cur);
  b->cur = i; emu_cur = &b->f[i];
  emu_switch(&dummy, b->f[i].sp);
  std::abort();                                                // a finished fiber is never resumed
}

// ---- warp collectives -----------------------------------------------------------------------------------------
template <class T> inline T emu_exchange(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu_warp* w = emu_cur->warp;
  uint64_t bits = 0; std::memcpy(&bits, &v, sizeof(T));
  w->slot[emu_cur->lane] = bits;
  emu_wait(w->bar);
  uint64_t r = w->slot[src & 31];
  emu_wait(w->bar);
  T out; std::memcpy(&out, &r, sizeof(T)); return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lanemask) { return emu_exchange(v, emu_cur->lane ^ lanemask); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int delta) { int l = emu_cur->lane; return emu_exchange(v, l - delta >= 0 ? l - delta : l); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) { int l = emu_cur->lane; return emu_exchange(v, l + delta < 32 ? l + delta : l); }
inline unsigned __ballot_sync(unsigned, int pred) {
  emu_warp* w = emu_cur->warp;
  w->slot[emu_cur->lane] = pred ? 1u : 0u;
  emu_wait(w->bar);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(w->slot[i] & 1u) << i;
  emu_wait(w->bar);
  return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_wait(emu_cur->warp->bar); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }

// ---- CTA barriers ---------------------------------------------------------------------------------------------
inline void __syncthreads() { emu_wait(emu_blk->bar); }
inline int __syncthreads_or(int pred) {
  emu_block* b = emu_blk;
  if (pred) b->vote = 1;
  emu_wait(b->bar);
  int r = b->vote;
  emu_wait(b->bar);
  b->vote = 0;
  emu_wait(b->bar);
  return r;
}

inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// ---- launch ---------------------------------------------------------------------------------------------------
inline void emu_run_block(unsigned bidx, unsigned block, const std::function<void()>& body, char* stacks) {
  emu_block blk;
  blk.f.resize(block); blk.warps.resize(block / 32);
  blk.bar.expected = block; blk.live = (int)block; blk.body = &body;
  for (unsigned w = 0; w < block / 32; w++) { blk.warps[w].bar.expected = 32; std::memset(blk.warps[w].slot, 0, sizeof(blk.warps[w].slot)); }
  for (unsigned t = 0; t < block; t++) {
    emu_fiber& f = blk.f[t];
    f.tid.x = t; f.bid.x = bidx; f.lane = (int)(t & 31); f.warp = &blk.warps[t >> 5];
    char* top = stacks + (size_t)(t + 1) * EMU_STACK;
    void** a = (void**)(((uintptr_t)top & ~(uintptr_t)15) - 16);      // a % 16 == 0: after `ret`, rsp % 16 == 8 as at a call
    a[0] = (void*)&emu_fiber_entry; a[1] = nullptr;
    void** sp = a - 6;
    for (int k = 0; k < 6; k++) sp[k] = nullptr;
    f.sp = sp;
  }
  const int first = emu_dir() > 0 ? 0 : (int)block - 1;
  emu_blk = &blk; blk.cur = first; emu_cur = &blk.f[first];
  emu_switch(&blk.sched_sp, blk.f[first].sp);
  emu_blk = nullptr; emu_cur = nullptr;
}
template <class K, class... A> inline void emu_launch(K kernel, unsigned grid, unsigned block, size_t smem_bytes, A... args) {
  if (smem_bytes > sizeof(smem) || block % 32 != 0 || block == 0) std::abort();
  blockDim.x = block; gridDim.x = grid;
  const std::function<void()> body = [&] { kernel(args...); };
  std::atomic<unsigned> next{0};
  auto worker = [&] {
    char* stacks = (char*)std::malloc((size_t)block * EMU_STACK);
    for (unsigned b = next.fetch_add(1); b < grid; b = next.fetch_add(1)) emu_run_block(b, block, body, stacks);
    std::free(stacks);
  };
  unsigned nw = std::min(grid, std::max(1u, std::min(8u, std::thread::hardware_concurrency())));
  std::vector<std::thread> ts;
  for (unsigned w = 0; w < nw; w++) ts.emplace_back(worker);
  for (auto& t : ts) t.join();
}
#define B200MJ_LAUNCH(kernel, grid, block, smem, stream, ...) emu_launch(kernel, (unsigned)(grid), (unsigned)(block), (size_t)(smem), __VA_ARGS__)

// ---- the handful of runtime calls the host side makes: "device" memory is host memory, streams are synchronous ------
typedef int cudaError_t;
enum { cudaSuccess = 0 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)std::malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
// blocks run on several host threads: a real atomic (the fibers of one block are cooperative, but global counters are shared)
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
