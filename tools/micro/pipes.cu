// Micro-benchmark: per-SM throughput of DFMA / SHFL / IMAD / LDS / DDIV-free chains on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double* out, int iters, double seed) {
  double a = seed + threadIdx.x, b = 1.0000001, c = 0.5, d = a + 1, e = a + 2, f = a + 3;
  int ia = threadIdx.x, ib = 3, ic = 7, id = 11;
  __shared__ double sm[1024];
  sm[threadIdx.x & 1023] = a;
  __syncthreads();
  for (int i = 0; i < iters; i++) {
    if (OP == 0) { a = fma(a, b, c); d = fma(d, b, c); e = fma(e, b, c); f = fma(f, b, c); }
    if (OP == 1) { a = __shfl_xor_sync(0xffffffffu, a, 1); d = __shfl_xor_sync(0xffffffffu, d, 2); e = __shfl_xor_sync(0xffffffffu, e, 4); f = __shfl_xor_sync(0xffffffffu, f, 8); }
    if (OP == 2) { ia = ia * ib + ic; ib = ib * ic + id; ic = ic * id + ia; id = id * ia + ib; }
    if (OP == 3) { a += sm[(ia + i) & 1023]; d += sm[(ia + 2 * i + 1) & 1023]; e += sm[(ia + 3 * i + 2) & 1023]; f += sm[(ia + 5 * i + 3) & 1023]; }
    if (OP == 4) { a = a * b; d = d + c; e = e * b; f = f + c; }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f + ia + ib + ic + id;
}
template <int OP> void run(const char* name, double ops_per_iter_per_thread) {
  double* out; cudaMalloc(&out, 148 * 32 * 1024 * 8);
  for (int warps = 1; warps <= 32; warps *= 2) {
    int iters = 20000;
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    k<OP><<<148, 32 * warps>>>(out, 100, 1.0);
    cudaEventRecord(s);
    k<OP><<<148, 32 * warps>>>(out, iters, 1.0);
    cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    double inst = (double)iters * ops_per_iter_per_thread * warps;   // warp-instructions per SM
    printf("%-6s warps/SM %2d: %.3f ms  -> %.3f warp-instr/cycle/SM (at 1.9 GHz)\n", name, warps, ms, inst / (ms * 1e-3 * 1.9e9));
  }
}
int main() {
  run<0>("DFMA", 4); run<1>("SHFL64", 8); run<2>("IMAD", 4); run<3>("LDS64", 4); run<4>("DMUL/DADD", 4);
  return 0;
}
