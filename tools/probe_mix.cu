// Standalone probe (design work, not part of the library): throughput of the instruction MIX of a 16x2 cell
// pair when nothing depends on anything (NI independent cells per thread, full occupancy) - the pipe bound the
// dependent step loop of bsw_warp16.h is measured against.  cycles = SMSP cycles per warp-level cell pair.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_probe_mix tools/probe_mix.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define VMN(d, a, b, c) asm volatile("{.reg .b32 t; add.s16x2 t, %1, %2; min.s16x2 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define PRMT(d, a, b, c) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define ORI(d, a) asm volatile("or.b32 %0, %1, 0x00030003;" : "=r"(d) : "r"(a))
#define MAD(d, a, b, c) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define MAD4(d, a, c) asm volatile("mad.lo.u32 %0, %1, 4, %2;" : "=r"(d) : "r"(a), "r"(c))
#define ADD(d, a, b) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define SUB(d, a, b) asm volatile("sub.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define SHLADD(d, a, b) asm volatile("{.reg .b32 t; shl.b32 t, %1, 2; add.u32 %0, t, %2;}" : "=r"(d) : "r"(a), "r"(b))

constexpr int NI = 12;
// MIX: 0 = 2 VMN; 1 = +PRMT; 2 = +OR (ALU part of a direction cell); 3 = +2 IMAD (the direction cell);
//      4 = 2 VMN + 2 IMAD; 5 = 2 IMAD; 6 = direction cell with LEA + IADD instead of the IMADs;
//      7 = direction cell with one IMAD (acc*4+v) + one IADD(-hc) ; 8 = PRMT + 2 VMN + OR + 1 IMAD
template <int MIX>
__global__ void __launch_bounds__(128) mix_kernel(unsigned* out, int iters, unsigned seed) {
  unsigned h[NI], a[NI], m[NI], v[NI], cd[NI];
#pragma unroll
  for (int u = 0; u < NI; u++) { h[u] = seed * (u + 3) + threadIdx.x; a[u] = seed + u; m[u] = v[u] = cd[u] = u; }
  const unsigned b = seed | 1u, c = seed ^ 0x12341234u, neg1 = 0u - (seed != 77u), sel = seed & 0x3333;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int u = 0; u < NI; u++) {
        if (MIX == 1 || MIX == 2 || MIX == 3 || MIX == 6 || MIX == 7 || MIX == 8) PRMT(cd[u], b, c, sel); else cd[u] = c;
        if (MIX != 5) { VMN(m[u], h[u], b, c); VMN(v[u], h[u], cd[u], m[u]); }
        if (MIX == 2 || MIX == 3 || MIX == 6 || MIX == 7 || MIX == 8) ORI(h[u], v[u]); else if (MIX != 5) h[u] = v[u];
        if (MIX == 3 || MIX == 4) { unsigned t; MAD4(t, a[u], v[u]); MAD(a[u], neg1, h[u], t); }
        if (MIX == 5) { unsigned t; MAD4(t, a[u], h[u]); MAD(a[u], neg1, h[u], t); }
        if (MIX == 6) { unsigned t; SHLADD(t, a[u], v[u]); SUB(a[u], t, h[u]); }
        if (MIX == 7) { unsigned t; MAD4(t, a[u], v[u]); SUB(a[u], t, h[u]); }
        if (MIX == 8) { MAD4(a[u], a[u], v[u]); }
      }
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int u = 0; u < NI; u++) r ^= h[u] ^ a[u];
  if (r == 0x7fffffffu) out[0] = r;
}

template <int MIX>
void run(unsigned* d_out, const char* name) {
  const int iters = 1024, threads = 128, blocks = 148 * 16;
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mix_kernel<MIX>, threads, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    mix_kernel<MIX><<<blocks, threads>>>(d_out, iters, rep + 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double warp_cells = (double)blocks * (threads / 32) * iters * 4.0 * NI;
  // SMSP cycles per warp-level cell pair at 1.965 GHz (the clock the bench sees), 592 SMSPs
  const double cyc = best * 1e-3 * 1.965e9 * 592.0 / warp_cells;
  printf("{\"mix\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"cycles_per_warp_cell_pair\": %.2f, \"slot_gcups\": %.0f}\n", name, nb, best, cyc,
         warp_cells * 64.0 / (best * 1e-3) / 1e9);
}

int main() {
  unsigned* d_out;
  cudaMalloc(&d_out, 4096);
  run<0>(d_out, "2 VMN");
  run<1>(d_out, "PRMT + 2 VMN (score cell)");
  run<2>(d_out, "PRMT + 2 VMN + OR");
  run<3>(d_out, "PRMT + 2 VMN + OR + 2 IMAD (direction cell)");
  run<4>(d_out, "2 VMN + 2 IMAD");
  run<5>(d_out, "2 IMAD");
  run<6>(d_out, "PRMT + 2 VMN + OR + LEA + IADD");
  run<7>(d_out, "PRMT + 2 VMN + OR + IMAD + IADD");
  run<8>(d_out, "PRMT + 2 VMN + OR + 1 IMAD");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
