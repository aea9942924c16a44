// Standalone probe (design work, not part of the library): how much would the 16x2 step loop gain from more
// resident warps (fewer registers: no per-slot "up" addends) and from 17 instead of 18 slots per lane?
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_probe_occ tools/probe_occ.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int LG = 8, G = 32 / LG;
template <int C, int TILE>
struct GroupSm {
  uint64_t btab[TILE + LG + 8];
  uint16_t asel[TILE + LG * C + 16 + 2];
};
__device__ __forceinline__ unsigned prmt(unsigned lo, unsigned hi, unsigned sel) {
  unsigned r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return r;
}
// DIRS: direction accumulators; UREG: per-slot U registers (else one common register + 2 selects per step, the
// non-uniform stripe scheme); NB: resident blocks per SM asked of the compiler
template <bool DIRS, int C, int NB, bool UREG, int TILE>
__global__ void __launch_bounds__(128, NB) loop_kernel(unsigned* out, int steps, unsigned seed, int reps) {
  __shared__ GroupSm<C, TILE> sm[4][G];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LG, gl = lane % LG;
  GroupSm<C, TILE>& s = sm[warp][grp];
  for (int i = gl; i < TILE + LG + 8; i += LG) s.btab[i] = 0x0102030405060708ull * (seed + i);
  for (int i = gl; i < TILE + LG * C + 16; i += LG) s.asel[i] = (uint16_t)((seed * 7 + i * 13) & 0x3333);
  __syncwarp();
  unsigned H[C], acc[C], U[UREG ? C : 1];
#pragma unroll
  for (int k = 0; k < C; k++) { H[k] = seed + k * 4; acc[k] = 0; }
#pragma unroll
  for (int k = 0; k < (UREG ? C : 1); k++) U[k] = out[k + 64];
  const unsigned neg1 = 0u - (unsigned)(seed != 0x7ffffffeu);
  const unsigned lneg = gl == 0 ? 0x80008000u : 0u;
  const unsigned shortm = (gl >= (int)(seed & 7)) ? 0xffffffffu : 0u;
  unsigned* fp = out + (blockIdx.x * 128 + threadIdx.x);
  for (int rep = 0; rep < reps; rep++) {
    const uint16_t* pa = s.asel + gl * (C - 1);
    const uint64_t* pb = s.btab + (LG - 1 - gl);
    for (int t = 0; t < steps; t += 2) {
      if ((t & (TILE - 1)) == 0) { pa = s.asel + gl * (C - 1); pb = s.btab + (LG - 1 - gl); }
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const uint64_t tb = pb[u];
        const unsigned tlo = (unsigned)tb, thi = (unsigned)(tb >> 32);
        const uint16_t* pw = pa + u;
        unsigned outl = H[C - 1];
        if (!UREG) outl = (H[C - 2] & shortm) | (H[C - 1] & ~shortm);
        unsigned left = (unsigned)__shfl_up_sync(0xffffffffu, outl, 1, LG) | lneg;
        unsigned right = 0;
#pragma unroll
        for (int k = 0; k < C; k++) {
          const unsigned cd = prmt(tlo, thi, pw[k]);
          unsigned up = (k == C - 1) ? right : H[(k + 1) % C];
          if (!UREG && k == C - 2) up = (right & shortm) | (up & ~shortm);
          const unsigned m = __viaddmin_s16x2(up, UREG ? U[k] : U[0], left);
          const unsigned v = __viaddmin_s16x2(H[k], cd, m);
          unsigned hc = v;
          if (DIRS) { hc = v | 0x00030003u; acc[k] = (acc[k] * 4u + v) + neg1 * hc; }
          H[k] = hc; left = hc;
          if (k == 0) right = (unsigned)__shfl_down_sync(0xffffffffu, H[0], 1, LG);
        }
      }
      pa += 2; pb += 2;
      if (DIRS && ((t + 2) & 7) == 0) {
#pragma unroll
        for (int k = 0; k < C; k++) fp[k * 128 * 592] = acc[k];
      }
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int k = 0; k < C; k++) r ^= H[k] ^ acc[k];
  if (r == 0x7fffffffu) out[0] = r;
}

template <bool DIRS, int C, int NB, bool UREG, int TILE>
void run_loop(unsigned* d_out, const char* name, int limit_blocks = 0) {
  const int steps = 1024, reps = 40;
  int nb = 0;
  size_t dyn = 0;  // dynamic shared memory that limits the resident blocks to limit_blocks
  if (limit_blocks) {
    dyn = (size_t)(227 * 1024 / limit_blocks) - sizeof(GroupSm<C, TILE>) * 4 * G - 1024;
    cudaFuncSetAttribute(loop_kernel<DIRS, C, NB, UREG, TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
  }
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, loop_kernel<DIRS, C, NB, UREG, TILE>, 128, dyn);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, loop_kernel<DIRS, C, NB, UREG, TILE>);
  const int blocks = 148 * nb;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    loop_kernel<DIRS, C, NB, UREG, TILE><<<blocks, 128, dyn>>>(d_out, steps, rep + 1, reps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double cells = (double)blocks * 128 * (double)steps * reps * C * 2.0;
    if (rep) best = cells / (ms * 1e-3) > best ? cells / (ms * 1e-3) : best;
  }
  // "real" = band-64 cells (129 columns of the C*LG computed)
  printf("{\"name\": \"%s\", \"dirs\": %d, \"C\": %d, \"tile\": %d, \"regs\": %d, \"smem\": %d, \"blocks_per_sm\": %d, \"slot_gcups\": %.0f, \"band64_gcups\": %.0f}\n",
         name, (int)DIRS, C, TILE, fa.numRegs, (int)fa.sharedSizeBytes, nb, best / 1e9, best / 1e9 * 129.0 / (C * LG));
}

int main(int argc, char** argv) {
  unsigned* d_out;
  cudaMalloc(&d_out, (size_t)18 * 128 * 148 * 9 * 4 + 4096);
  if (argc > 1 && argv[1][0] == 's') {  // throughput against resident warps per scheduler
    for (int nb = 1; nb <= 4; nb++) run_loop<true, 18, 4, true, 128>(d_out, "dirs C18 U[] limited", nb);
    for (int nb = 1; nb <= 7; nb++) run_loop<false, 18, 4, true, 128>(d_out, "score C18 U[] limited", nb);
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'd') { run_loop<true, 18, 4, true, 128>(d_out, "dirs C18 U[] nb4"); return 0; }
  if (argc > 1 && argv[1][0] == 'c') { run_loop<false, 18, 4, true, 128>(d_out, "score C18 U[] nb4"); return 0; }
  run_loop<true, 18, 4, true, 128>(d_out, "dirs C18 U[] nb4");
  run_loop<true, 18, 4, false, 128>(d_out, "dirs C18 noU nb4");
  run_loop<true, 17, 4, false, 128>(d_out, "dirs C17 noU nb4");
  run_loop<true, 17, 5, false, 128>(d_out, "dirs C17 noU nb5");
  run_loop<true, 17, 5, false, 64>(d_out, "dirs C17 noU nb5 tile64");
  run_loop<true, 17, 6, false, 64>(d_out, "dirs C17 noU nb6 tile64");
  run_loop<false, 18, 4, true, 128>(d_out, "score C18 U[] nb4");
  run_loop<false, 17, 4, false, 128>(d_out, "score C17 noU nb4");
  run_loop<false, 17, 5, false, 128>(d_out, "score C17 noU nb5");
  run_loop<false, 17, 6, false, 64>(d_out, "score C17 noU nb6 tile64");
  run_loop<false, 17, 8, false, 64>(d_out, "score C17 noU nb8 tile64");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
