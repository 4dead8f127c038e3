// cuda_emu.h — lock-step CPU emulation of the CUDA constructs dm_control_b200/csrc/b200mj.cu uses.
//
// TEST INFRASTRUCTURE ONLY. It lets `pytest -m "not gpu"` compile the *kernel source itself* with g++
// (-DB200MJ_CPU_EMU) and run it for a handful of environments, so that logic errors in the CUDA path (indexing, layout
// aliasing, launch orchestration, the C ABI) surface on a machine without a GPU. It says nothing about performance or
// about data races, it is never built by __graft_entry__.build(), and nothing under dm_control_b200/ can load it.
//
// Model: one OS thread per CUDA thread of the running block, blocks one after another. Warp collectives
// (__shfl*_sync, __ballot_sync, __any_sync, __syncwarp) and CTA barriers (__syncthreads*) are rendezvous points on
// std::barrier, so every lane observes exactly the values its peers published at that call — the semantics the kernels
// rely on (they only use full-mask collectives in warp-uniform control flow). Threads that return from the kernel
// drop out of the barriers, as exited CUDA threads do.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __device__
#define __global__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__

using std::max;
using std::min;

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
struct alignas(16) double2 { double x, y; };
struct emu_warp { std::barrier<> bar{32}; uint64_t slot[32]; };
struct emu_block {
  explicit emu_block(unsigned n) : bar((std::ptrdiff_t)n) {}
  std::barrier<> bar; std::atomic<int> vote{0};
};

inline thread_local emu_dim3 threadIdx, blockIdx;
inline emu_dim3 blockDim, gridDim;
inline thread_local emu_warp* emu_w = nullptr;
inline thread_local emu_block* emu_b = nullptr;
inline thread_local int emu_lane = 0;

alignas(16) inline double smem[232448 / 8];     // dynamic shared memory of the (single) running block

// ---- warp collectives -------------------------------------------------------------------------------------------
template <class T> inline T emu_exchange(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t bits = 0; std::memcpy(&bits, &v, sizeof(T));
  emu_w->slot[emu_lane] = bits;
  emu_w->bar.arrive_and_wait();
  uint64_t r = emu_w->slot[src & 31];
  emu_w->bar.arrive_and_wait();
  T out; std::memcpy(&out, &r, sizeof(T)); return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lanemask) { return emu_exchange(v, emu_lane ^ lanemask); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int delta) { return emu_exchange(v, emu_lane - delta >= 0 ? emu_lane - delta : emu_lane); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int delta) { return emu_exchange(v, emu_lane + delta < 32 ? emu_lane + delta : emu_lane); }
inline unsigned __ballot_sync(unsigned, int pred) {
  emu_w->slot[emu_lane] = pred ? 1u : 0u;
  emu_w->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (unsigned)(emu_w->slot[i] & 1u) << i;
  emu_w->bar.arrive_and_wait();
  return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_w->bar.arrive_and_wait(); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }

// ---- CTA barriers -----------------------------------------------------------------------------------------------
inline void __syncthreads() { emu_b->bar.arrive_and_wait(); }
inline int __syncthreads_or(int pred) {
  if (pred) emu_b->vote.store(1);
  emu_b->bar.arrive_and_wait();
  int r = emu_b->vote.load();
  emu_b->bar.arrive_and_wait();
  if (threadIdx.x == 0) emu_b->vote.store(0);
  emu_b->bar.arrive_and_wait();
  return r;
}

inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// ---- launch -----------------------------------------------------------------------------------------------------
template <class K, class... A> inline void emu_launch(K kernel, unsigned grid, unsigned block, size_t smem_bytes, A... args) {
  if (smem_bytes > sizeof(smem) || block % 32 != 0) std::abort();
  blockDim.x = block; gridDim.x = grid;
  for (unsigned b = 0; b < grid; b++) {
    emu_block blk(block);
    std::vector<std::unique_ptr<emu_warp>> warps;
    for (unsigned w = 0; w < block / 32; w++) warps.emplace_back(new emu_warp);
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < block; t++)
      ts.emplace_back([&, t] {
        threadIdx.x = t; blockIdx.x = b; emu_lane = (int)(t & 31); emu_w = warps[t >> 5].get(); emu_b = &blk;
        kernel(args...);
        emu_w->bar.arrive_and_drop();       // exited threads no longer take part in barriers
        blk.bar.arrive_and_drop();
      });
    for (auto& th : ts) th.join();
  }
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
