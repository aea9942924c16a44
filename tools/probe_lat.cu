// Dependent-issue latencies of the instructions on the cell chain (round 2 design work): one warp per block,
// a serial chain of N operations, cycles per operation from clock64().
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
template <int WHICH>
__global__ void lat_kernel(unsigned* out, long long* cyc, unsigned a0, unsigned b, unsigned c, int n) {
  unsigned a = a0 + threadIdx.x, d = a0 * 3 + 1, e = a0 ^ 0x5555;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (WHICH == 0) a = (unsigned)__viaddmax_s32((int)a, (int)b, (int)c);
      if (WHICH == 1) a = __viaddmin_s16x2(a, b, c);
      if (WHICH == 2) a = (a | b) ^ c;
      if (WHICH == 3) { unsigned m = __viaddmin_s16x2(d, b, a); unsigned v = __viaddmin_s16x2(e, c, m); a = v | 0x00030003u; }  // the DIRS cell chain
      if (WHICH == 4) { unsigned m = __viaddmin_s16x2(d, b, a); a = __viaddmin_s16x2(e, c, m); }                              // the score cell chain
      if (WHICH == 5) a = (unsigned)__shfl_up_sync(0xffffffffu, (int)a, 1, 8) + b;
      if (WHICH == 6) asm("prmt.b32 %0, %1, %2, %3;" : "=r"(a) : "r"(a), "r"(b), "r"(c));
      if (WHICH == 7) a = a * b + c;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (a == 0x12345678u) out[0] = a + d + e;
}
template <int W> double run(unsigned* o, long long* c) {
  const int n = 4096;
  lat_kernel<W><<<1, 32>>>(o, c, 7, 3, 5, n);
  lat_kernel<W><<<1, 32>>>(o, c, 7, 3, 5, n);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  return (double)h / (n * 16.0);
}
int main() {
  unsigned* o; long long* c; cudaMalloc(&o, 64); cudaMalloc(&c, 64);
  printf("{\"cycles_per_dependent_op\": {\"viaddmax_s32\": %.2f, \"viaddmin_s16x2\": %.2f, \"lop3\": %.2f, \"dirs_cell_chain(3 ops)\": %.2f, \"score_cell_chain(2 ops)\": %.2f, \"shfl_up+iadd\": %.2f, \"prmt\": %.2f, \"imad\": %.2f}}\n",
         run<0>(o, c), run<1>(o, c), run<2>(o, c), run<3>(o, c), run<4>(o, c), run<5>(o, c), run<6>(o, c), run<7>(o, c));
  return 0;
}
