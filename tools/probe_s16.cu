// Standalone issue-rate probe for the 16x2 SIMD cell update (round 2 design work, not part of the
// library): single-instruction rates of the DPX 16x2 family and steady-state step loops that mimic
// the K1 stripe (C slots per lane, two shuffles per step, selector loads from shared memory) in the
// candidate instruction mixes.  Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_probe_s16 tools/probe_s16.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int WHICH>
__global__ void __launch_bounds__(256) rate_kernel(unsigned* out, int iters, unsigned seed) {
  unsigned a[8];
#pragma unroll
  for (int u = 0; u < 8; u++) a[u] = seed * (u + 3) + threadIdx.x;
  const unsigned b = seed | 1u, c = seed ^ 0x12341234u;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        if (WHICH == 0) a[u] = (unsigned)__viaddmax_s32((int)a[u], (int)b, (int)c);
        if (WHICH == 1) a[u] = __viaddmax_s16x2(a[u], b, c);
        if (WHICH == 2) a[u] = __vmaxs2(a[u], c ^ a[(u + 1) & 7]);
        if (WHICH == 3) a[u] = __vimax3_s16x2(a[u], b, a[(u + 1) & 7]);
        if (WHICH == 4) a[u] = __dp2a_lo(b, c, a[u]);
        if (WHICH == 5) a[u] = __vadd2(a[u], b);
        if (WHICH == 6) a[u] = __viaddmax_u16x2(a[u], b, c);
        if (WHICH == 7) asm("prmt.b32 %0, %1, %2, %3;" : "=r"(a[u]) : "r"(a[u]), "r"(b), "r"(c));
        if (WHICH == 8) a[u] = (a[u] & b) ^ c;
        if (WHICH == 9) a[u] = a[u] * b + c;
        if (WHICH == 10) a[u] = __dp4a(b, c, a[u]);
      }
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) r ^= a[u];
  if (r == 0x7fffffffu) out[0] = r;
}

// ---- step loops -------------------------------------------------------------------------------------
constexpr int C = 18, LG = 8, G = 32 / LG, TILE = 128;
struct GroupSm {
  uint64_t btab[TILE + LG + 8];
  uint16_t asel[TILE + LG * C + 16 + 2];  // +2: odd group stride in words
};

__device__ __forceinline__ unsigned prmt(unsigned lo, unsigned hi, unsigned sel) {
  unsigned r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return r;
}

// MODE 0: s32 cells as in round 1 (PRMT per 4 cells, IDP.4A, VIADDMNMX, VIMNMX, LOP3, 2 IMAD), DIRS
// MODE 1: s16x2, PRMT per cell pair, 2 VIADDMNMX.S16x2, LOP3, 2 IMAD (DIRS)
// MODE 2: s16x2, PRMT per two cell pairs + IDP.2A, VIADDMNMX + VIMNMX, LOP3, 2 IMAD (DIRS)
// MODE 3: MODE 1 without directions (score only)
// MODE 4: MODE 2 without directions
// MODE 5: s32 score only (round 1)
template <int MODE>
__global__ void __launch_bounds__(128, 4) loop_kernel(unsigned* out, int steps, unsigned seed, int reps) {
  __shared__ GroupSm sm[4][G];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane / LG, gl = lane % LG;
  GroupSm& s = sm[warp][grp];
  for (int i = gl; i < TILE + LG + 8; i += LG) s.btab[i] = 0x0102030405060708ull * (seed + i);
  for (int i = gl; i < TILE + LG * C + 16; i += LG) s.asel[i] = (uint16_t)((seed * 7 + i * 13) & 0x3333);
  __syncwarp();
  unsigned H[C], acc[C], U[C];
#pragma unroll
  for (int k = 0; k < C; k++) { H[k] = seed + k * 4; acc[k] = 0; U[k] = out[k + 64]; }
  const unsigned neg1 = 0u - (unsigned)(seed != 0x7ffffffeu);
  const unsigned lneg = gl == 0 ? 0x80008000u : 0u;
  unsigned* fp = out + (blockIdx.x * 128 + threadIdx.x);
  constexpr bool DIRS = (MODE <= 2);
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
        unsigned left = (unsigned)__shfl_up_sync(0xffffffffu, H[C - 1], 1, LG) | lneg;
        unsigned right = 0;
        if (MODE == 0 || MODE == 5) {
          unsigned cd4[(C + 3) / 4];
#pragma unroll
          for (int q = 0; q < (C + 3) / 4; q++) cd4[q] = prmt(tlo, thi, pw[4 * q]);
#pragma unroll
          for (int k = 0; k < C; k++) {
            const int up = (k == C - 1) ? (int)right : (int)H[(k + 1) % C];
            const int d = (int)__dp4a(cd4[k / 4], 1u << (8 * (k % 4)), H[k]);
            const int m = __viaddmax_s32(up, (int)U[k], (int)left);
            const int v = max(d, m);
            int hc = v;
            if (MODE == 0) { hc = v & ~3; acc[k] = (acc[k] * 4u + (unsigned)v) + neg1 * (unsigned)hc; }
            H[k] = hc; left = hc;
            if (k == 0) right = (unsigned)__shfl_down_sync(0xffffffffu, H[0], 1, LG);
          }
        } else if (MODE == 1 || MODE == 3) {
#pragma unroll
          for (int k = 0; k < C; k++) {
            const unsigned cd = prmt(tlo, thi, pw[k]);
            const unsigned up = (k == C - 1) ? right : H[(k + 1) % C];
            const unsigned m = __viaddmax_s16x2(up, U[k], left);
            const unsigned v = __viaddmax_s16x2(H[k], cd, m);
            unsigned hc = v;
            if (MODE == 1) { hc = v & 0xfffcfffcu; acc[k] = (acc[k] * 4u + v) + neg1 * hc; }
            H[k] = hc; left = hc;
            if (k == 0) right = (unsigned)__shfl_down_sync(0xffffffffu, H[0], 1, LG);
          }
        } else {
          unsigned cd2[C / 2];
#pragma unroll
          for (int q = 0; q < C / 2; q++) cd2[q] = prmt(tlo, thi, pw[2 * q]);
#pragma unroll
          for (int k = 0; k < C; k++) {
            const unsigned up = (k == C - 1) ? right : H[(k + 1) % C];
            const unsigned d = (k & 1) ? __dp2a_hi(0x80000001u, cd2[k / 2], H[k]) : __dp2a_lo(0x80000001u, cd2[k / 2], H[k]);
            const unsigned m = __viaddmax_u16x2(up, U[k], left);
            const unsigned v = __vmaxu2(d, m);
            unsigned hc = v;
            if (MODE == 2) { hc = v & 0xfffcfffcu; acc[k] = (acc[k] * 4u + v) + neg1 * hc; }
            H[k] = hc; left = hc;
            if (k == 0) right = (unsigned)__shfl_down_sync(0xffffffffu, H[0], 1, LG);
          }
        }
      }
      pa += 2; pb += 2;
      if (DIRS && ((t + 2) & (MODE == 0 ? 15 : 7)) == 0) {
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

template <int WHICH>
double run_rate(unsigned* d_out) {
  const int iters = 2048, threads = 256, blocks = 148 * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    rate_kernel<WHICH><<<blocks, threads>>>(d_out, iters, rep + 1);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * threads * iters * 64.0;
    if (rep) best = ops / (ms * 1e-3) > best ? ops / (ms * 1e-3) : best;
  }
  return best;
}

template <int MODE>
double run_loop(unsigned* d_out) {
  const int steps = 1024, reps = 40, blocks = 148 * 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    loop_kernel<MODE><<<blocks, 128>>>(d_out, steps, rep + 1, reps);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double jobs_per_lane = (MODE == 0 || MODE == 5) ? 1.0 : 2.0;
    const double cells = (double)blocks * 128 * (double)steps * reps * C * jobs_per_lane;
    if (rep) best = cells / (ms * 1e-3) > best ? cells / (ms * 1e-3) : best;
  }
  return best;
}

int main() {
  unsigned* d_out;
  cudaMalloc(&d_out, (size_t)18 * 128 * 592 * 4 + 4096);
  const char* names[] = {"viaddmax_s32", "viaddmax_s16x2", "vmaxs2", "vimax3_s16x2", "dp2a_lo", "vadd2", "viaddmax_u16x2", "prmt", "lop3", "imad", "dp4a"};
  printf("{\"rates_T_lane_ops_per_s\": {");
  double r[11];
  r[0] = run_rate<0>(d_out); r[1] = run_rate<1>(d_out); r[2] = run_rate<2>(d_out); r[3] = run_rate<3>(d_out);
  r[4] = run_rate<4>(d_out); r[5] = run_rate<5>(d_out); r[6] = run_rate<6>(d_out); r[7] = run_rate<7>(d_out);
  r[8] = run_rate<8>(d_out); r[9] = run_rate<9>(d_out); r[10] = run_rate<10>(d_out);
  for (int i = 0; i < 11; i++) printf("%s\"%s\": %.2f", i ? ", " : "", names[i], r[i] / 1e12);
  printf("}}\n");
  const char* ln[] = {"s32_dirs_r1", "s16x2_dirs_prmt", "s16x2_dirs_dp2a", "s16x2_score_prmt", "s16x2_score_dp2a", "s32_score_r1"};
  double l[6];
  l[0] = run_loop<0>(d_out); l[1] = run_loop<1>(d_out); l[2] = run_loop<2>(d_out); l[3] = run_loop<3>(d_out);
  l[4] = run_loop<4>(d_out); l[5] = run_loop<5>(d_out);
  printf("{\"loop_gcups\": {");
  for (int i = 0; i < 6; i++) printf("%s\"%s\": %.0f", i ? ", " : "", ln[i], l[i] / 1e9);
  printf("}}\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
