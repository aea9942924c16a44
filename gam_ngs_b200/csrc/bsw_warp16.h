// K1s - the warp-level banded overlap DP of bsw_warp.h with 16x2 SIMD cells: every 32-bit register
// holds one cell of TWO jobs (job A in the low half-word, job B in the high one), so a group of LG
// lanes advances a PAIR of jobs per step with the DPX 16x2 instructions
//      m = VIADDMNMX.S16x2(up, U, left)        v = VIADDMNMX.S16x2(diag, Cd, m)
// (two instructions per two cells instead of three per cell), one PRMT per cell pair for the
// substitution bytes and - with the direction store - one LOP3 and two IMAD.
//
// Computes what BandedSmithWaterman::find_alignment computes
// (/root/reference/lib/src/alignment/banded_smith_waterman.cc:69-323); same band coordinates, lane
// stripes, skewed wavefront and tie-breaking as bsw_warp.h (read that header first).  What differs:
//
// Values.  A half-word holds  V - base  with  V = ((H + beta*j) << SH) | tag,  beta = -gap (no per-row
// offset: alpha = 0), so  diag: V + Cd,  Cd = (S << SH) | tag;  up: V + U,  U = ((2*gap) << SH) | 1;
// left: V.  H itself grows without bound (|H| <= 5 rows), but neighbouring cells
// cannot differ by much: for cells of one row  -8 <= H(i,j) - H(i,j-1) <= 13  and for cells of one
// column  -4 <= H(i,j) - H(i-1,j) <= 5  (induction over the recurrence .cc:160-164 with S in [-4,5],
// gap <= -5; the never-written region pos < 0 is identically 0 and obeys the same bounds).  So the
// C cells of a lane stripe span at most 21*(C-1) score units, and every LANE keeps its own base:
// every kRebaseSteps steps a lane subtracts (slot 0 - kT0) from its registers and adds it to its
// 32-bit base; the two values a lane exchanges with its neighbours per step are translated by the
// difference of the two bases (dL, dR; one VIADD.16x2 each).  With kT0 = 6144 and 256 steps between
// rebases every live half-word stays inside [1700, 18000] for any band, row count and input
// (bsw_warp16.h: "range budget"), so the 16-bit arithmetic never wraps on a value that is used.
//
// Substitution bytes.  PRMT sees 8 table bytes: 4 for job A's row (indexed by the a-base A,T,C,G), 4
// for job B's.  The selector of a cell pair is one 16-bit shared-memory entry
//      [ selA | 8+selA | 4+selB | 12+selB ]   (nibble 8+x replicates the sign of byte x)
// so the result is the two sign-extended Cd half-words.  There is no room for N or the padding symbol:
//   * jobs whose windows hold an N are not run here (the kernel checks the N masks of both windows of
//     both jobs first and falls back to the 32-bit body),
//   * the never-written region pos < 0 is kept by FREEZING those cells: while a lane
//     still has cells with pos < 0 (steps t < -p0) the PAD step variant does not update them,
//   * positions >= |a| are a closed region (nothing flows back into filled cells), computed with an
//     arbitrary base; the end-cell search treats them as the never-filled zeros they are (.cc:183).
// Band column 2B has no "up" neighbour: its U is kUpBlock16; the padding columns to its right are
// reset at every rebase so that they cannot run away from the filled ones.
//
// Directions.  A 32-bit word holds the tags of 8 consecutive steps of one band column for both jobs
// (A: low half-word, B: high half-word, oldest tag in the top bit pair of its half); word
// ((t>>3)*C + k)*LG + l of the PAIR's region.  A pair's region is the two per-job regions of the
// 32-bit layout side by side, so the fallback can use them as they are.
#pragma once
#include "bsw_common.h"
#include "bsw_warp.h"

namespace gamx {

constexpr int kT0 = 6144;             // a lane's slot 0 after a rebase (multiple of 4)
constexpr int kRebaseSteps = 256;     // steps between rebases (power of two, multiple of 8)
constexpr int kUpBlock16 = -32768 + 4096;  // "up" addend of band column 2B: below every live value, no wrap
constexpr uint32_t kNeg16x2 = 0x80008000u;
// range budget (DIRS, the wider case; units of a quarter score): after a rebase slot 0 = kT0 and slot
// k <= kT0 + 84*k; in 256 steps a cell's column drifts by [-16, +20] per step; incoming left/right
// neighbours add 84, the intermediate up + U subtracts at most 232:
//   live min >= 6144 - 4096 - 84 - 232 - 20 = 1712      live max <= 6144 + 1428 + 5120 + 107 = 12799
//   padding columns (reset to kT0 at a rebase, fed by column 2B) <= 12799 + 5120 = 17919
//   padding + kUpBlock16 in [-26960, -10753]: no wrap, below every live value.

// ---- 16x2 helpers (DPX on the device; the host forms wrap exactly like the hardware) -------------
GAMX_HD int half_lo(uint32_t v) { return (int)(int16_t)(uint16_t)(v & 0xffffu); }
GAMX_HD int half_hi(uint32_t v) { return (int)(int16_t)(uint16_t)(v >> 16); }
GAMX_HD uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
GAMX_HD uint32_t vadd2w(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __vadd2(a, b);
#else
  return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16);
#endif
}
GAMX_HD uint32_t vsub2w(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __vsub2(a, b);
#else
  return ((a - b) & 0xffffu) | (((a >> 16) - (b >> 16)) << 16);
#endif
}
// per half-word max(a + b, c), signed, the sum wraps
GAMX_HD uint32_t viaddmax2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __viaddmax_s16x2(a, b, c);
#else
  const uint32_t s = vadd2w(a, b);
  const int lo = half_lo(s) > half_lo(c) ? half_lo(s) : half_lo(c);
  const int hi = half_hi(s) > half_hi(c) ? half_hi(s) : half_hi(c);
  return pack2(lo, hi);
#endif
}
// byte permute with the sign-replicating selector nibbles (8+x): the full PTX default mode
GAMX_HD uint32_t prmt_sx(uint32_t lo, uint32_t hi, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return r;
#else
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  uint32_t r = 0;
  for (int n = 0; n < 4; n++) {
    const uint32_t s = (sel >> (4 * n)) & 15u;
    uint32_t byte = (uint32_t)((v >> (8 * (s & 7u))) & 0xffu);
    if (s & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
    r |= byte << (8 * n);
  }
  return r;
#endif
}
// bitwise select: (a & m) | (b & ~m), one LOP3
GAMX_HD uint32_t bitsel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }

template <int C, int LG>
struct alignas(16) GroupSmem16 {
  uint64_t btab[(kTileSteps + LG + 7) / 8 * 8];  // rows of the tile: low word = job A's 4-byte Cd table, high word = job B's
  uint16_t asel[kTileSteps + LG * C + 16];       // one PRMT selector per a-position of the tile (both jobs)
  int cap[2][C * LG];                            // latched "last column" cells (true V), [job][slot][lane]
};
template <int C, int LG>
struct WarpSmem16 {
  GroupSmem16<C, LG> g[LG >= 32 ? 1 : 32 / LG];
};

// words of a PAIR's direction region (8 steps per word; an even number of 8-step blocks so that the
// traceback can always read two consecutive blocks as one 16-step word)
GAMX_HD uint64_t k1_dir_words16(int x, int c, int lg) {
  const uint64_t steps = (uint64_t)x + lg + 16;
  return 2 * ((steps + 15) / 16) * (uint64_t)c * lg;
}

// Does the view range [from, to] (view positions, from <= to, inside the view) hold an N?  The lanes of
// a group share the mask words; every lane returns its partial answer (OR-reduce over the group).
GAMX_HD uint32_t range_n_bits(const SeqStore& s, const SeqView& v, int64_t from, int64_t to, int gl, int lg) {
  if (to < from) return 0u;
  const int64_t lo = v.dir > 0 ? v.origin + from : v.origin - to;
  const int64_t hi = v.dir > 0 ? v.origin + to : v.origin - from;
  const int64_t m0 = lo >> 5, m1 = hi >> 5;
  uint32_t any = 0;
  for (int64_t m = m0 + gl; m <= m1; m += lg) {
    uint32_t wv = s.nmask[m];
    if (m == m0) wv &= 0xffffffffu << (lo & 31);
    if (m == m1) wv &= 0xffffffffu >> (31 - (hi & 31));
    any |= wv;
  }
  return any;
}
// N bits of both windows of a job (a: every position a cell of the band can read; b: the DP rows)
GAMX_HD uint32_t job_n_bits(const SeqStore& s, const DevJob* J, int gl, int lg) {
  if (!J) return 0u;
  int64_t a_lo = J->p0 < 0 ? 0 : J->p0;
  int64_t a_hi = (int64_t)J->p0 + J->x - 1 + 2 * (int64_t)J->band;
  if (a_hi > (int64_t)J->la - 1) a_hi = (int64_t)J->la - 1;
  return range_n_bits(s, J->a, a_lo, a_hi, gl, lg) | range_n_bits(s, J->b, 0, (int64_t)J->x - 1, gl, lg);
}

// One warp: 32/LG PAIRS.  JA / JB / outA / outB are per-lane arguments, uniform within a group of LG
// lanes: the group's two jobs (JB null: the high half idles; both null: idle group) and their result
// slots.  Preconditions (the caller checks them, see k1s_kernel): both jobs are "regular" jobs of the
// geometry (C, LG) with the same band and gap, and no N in any window.  pair_dirs: the group's
// direction region.  The end cell, score and layout tag (has_match = 1 + half) go to the result
// records; the traceback kernel completes them.
template <int C, int LG, bool DIRS, class W>
GAMX_HD void warp_align16(W& w, const DevJob* JA, const DevJob* JB, const SeqStore& store, WarpSmem16<C, LG>& wsm,
                          uint32_t* pair_dirs, DevResult* outA, DevResult* outB) {
  static_assert(stripe_supported(C), "lane stripe width");
  static_assert(LG == 4 || LG == 8 || LG == 16 || LG == 32, "lanes per pair");
  constexpr int SH = DIRS ? 2 : 0;
  constexpr int UF = unroll_of(C);
  constexpr uint32_t CLEAN = DIRS ? 0xfffcfffcu : 0xffffffffu;
  const int lane = w.lane();
  const int grp = lane / LG, gl = lane % LG;
  GroupSmem16<C, LG>& sm = wsm.g[grp];
  const DevJob* Jh[2] = {JA, JB};
  const bool liveh[2] = {JA != nullptr, JB != nullptr};
  const DevJob* J0 = JA ? JA : JB;   // band and gap are common to the pair
  const bool live = J0 != nullptr;
  uint32_t* fp = DIRS ? pair_dirs + gl : nullptr;  // running flush pointer

  const int B = live ? J0->band : 0, Y = 2 * B + 1;
  const int ld = (Y - 1) / C, kd = (Y - 1) - ld * C;  // lane/slot of band column 2B
  const int gap = live ? J0->gap : -8;
  const int beta = -gap;
  const int j0 = gl * C;
  int X[2], la[2], p0[2], kc[2], tcap0[2];
  SeqView va[2], vb[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    X[h] = liveh[h] ? Jh[h]->x : 0;
    la[h] = liveh[h] ? Jh[h]->la : 0;
    p0[h] = liveh[h] ? Jh[h]->p0 : 0;
    kc[h] = liveh[h] ? Jh[h]->kc : -1;
    tcap0[h] = kc[h] - gl * (C - 1);  // step at which slot 0 holds a "last column" cell (slot k: tcap0 - k)
    va[h].origin = vb[h].origin = 0; va[h].dir = vb[h].dir = 1; va[h].comp = vb[h].comp = 0;
    if (liveh[h]) { va[h] = Jh[h]->a; vb[h] = Jh[h]->b; }
  }

  // warp-uniform extents
  int t_end = 0, x_min = 0x7fffffff, win_lo = 0x7fffffff, win_hi = -1, pad_end = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    if (!liveh[h]) continue;
    t_end = imax(t_end, X[h] + ld);
    x_min = imin(x_min, X[h]);
    if (kc[h] >= 0) { win_lo = imin(win_lo, kc[h] - (ld + 1) * (C - 1)); win_hi = imax(win_hi, kc[h]); }
    pad_end = imax(pad_end, -p0[h]);  // steps t < -p0 still meet cells with pos < 0
  }
#pragma unroll
  for (int d = LG; d < 32; d <<= 1) {
    t_end = imax(t_end, w.shfl_xor(t_end, d, 32));
    x_min = imin(x_min, w.shfl_xor(x_min, d, 32));
    win_lo = imin(win_lo, w.shfl_xor(win_lo, d, 32));
    win_hi = imax(win_hi, w.shfl_xor(win_hi, d, 32));
    pad_end = imax(pad_end, w.shfl_xor(pad_end, d, 32));
  }
  const int T_total = t_end;

  // Cd bytes (signed): match, mismatch
  const uint32_t cdMb = (uint32_t)((kScoreMatch << SH) | (DIRS ? kTagDiagMatch : 0)) & 0xffu;
  const uint32_t cdXb = (uint32_t)((kScoreMismatch * (1 << SH)) | (DIRS ? kTagDiagMis : 0)) & 0xffu;

  uint32_t H[C], acc[C], U[C];
  const int u_up = ((2 * gap) * (1 << SH)) | (DIRS ? kTagUp : 0);
#pragma unroll
  for (int k = 0; k < C; k++) {
    H[k] = 0; acc[k] = 0;
    U[k] = (gl == ld && k == kd) ? pack2(kUpBlock16, kUpBlock16) : pack2(u_up, u_up);
  }
  const uint32_t neg1 = (uint32_t)(gap >> 31);  // -1 in a register the compiler cannot fold (FMA-pipe accumulate, see bsw_warp.h)
  // neighbour exchange: lane 0 has no left neighbour, lanes from ld on take no "up" from the right
  const uint32_t lkeep = gl == 0 ? 0u : 0xffffffffu, lor = gl == 0 ? kNeg16x2 : 0u;
  const uint32_t rkeep = gl >= ld ? 0u : 0xffffffffu;
  const int kdl = gl < ld ? C - 1 : (gl == ld ? kd : -1);  // last slot of this lane that is a band column
  int base[2] = {0, 0};     // true V = half-word + base
  uint32_t dL = 0, dR = 0;  // base of the left / right neighbour lane minus this lane's, per half

  // ---- tile staging: selectors of the a-positions, Cd tables of the b-rows ----------------------------
#define GAMX16_STAGE_TILE(T0)                                                                          \
  {                                                                                                    \
    w.sync();                                                                                          \
    const int na = kTileSteps + LG * C - (LG - 1);                                                     \
    for (int c0 = gl * 8; c0 < na; c0 += LG * 8) {                                                     \
      uint32_t byt[2][2];  /* [job][positions 0-3 / 4-7]: selector byte per position */               \
      _Pragma("unroll") for (int h = 0; h < 2; h++) {                                                  \
        const int pos0 = p0[h] + (T0) + c0;                                                            \
        uint32_t codes = 0, nfl = 0;                                                                   \
        if (liveh[h] && pos0 + 7 >= 0 && pos0 < la[h]) load_codes16(store, va[h], pos0, la[h], &codes, &nfl); \
        codes = (codes ^ (va[h].comp * 0x5555u)) & 0xffffu;                                            \
        const uint32_t nib = spread2to4(codes);            /* 8 nibbles, values 0..3 */                \
        const uint32_t tag = h ? 0xc4c4c4c4u : 0x80808080u;                                            \
        _Pragma("unroll") for (int q = 0; q < 2; q++) {                                                \
          uint32_t x = (nib >> (16 * q)) & 0xffffu;                                                    \
          x = (x | (x << 8)) & 0x00ff00ffu;                                                            \
          x = (x | (x << 4)) & 0x0f0f0f0fu;                /* one nibble value per byte */             \
          byt[h][q] = (x * 0x11u) | tag;                   /* [sel | 8+sel] resp. [4+sel | 12+sel] */  \
        }                                                                                              \
      }                                                                                                \
      Quad qv;                                                                                         \
      qv.v[0] = prmt(byt[0][0], byt[1][0], 0x5140u); qv.v[1] = prmt(byt[0][0], byt[1][0], 0x7362u);    \
      qv.v[2] = prmt(byt[0][1], byt[1][1], 0x5140u); qv.v[3] = prmt(byt[0][1], byt[1][1], 0x7362u);    \
      *reinterpret_cast<Quad*>(sm.asel + c0) = qv;    /* c0 % 8 == 0: 16-byte aligned */               \
    }                                                                                                  \
    const int nb = kTileSteps + LG - 1;                                                                \
    for (int r0 = gl * 8; r0 < nb; r0 += LG * 8) {                                                     \
      const int i0 = (T0) - (LG - 1) + r0;  /* table r0+q is row i0+q */                               \
      uint32_t cds[2];                                                                                 \
      _Pragma("unroll") for (int h = 0; h < 2; h++) {                                                  \
        uint32_t codes = 0, nfl = 0;                                                                   \
        if (liveh[h] && i0 + 7 >= 0 && i0 < X[h]) load_codes16(store, vb[h], i0, X[h], &codes, &nfl);  \
        cds[h] = codes ^ (vb[h].comp * 0x5555u);                                                       \
      }                                                                                                \
      _Pragma("unroll") for (int q = 0; q < 8; q++) {                                                  \
        const uint32_t ta = (cdXb * 0x01010101u) ^ ((cdXb ^ cdMb) << (8 * ((cds[0] >> (2 * q)) & 3u))); \
        const uint32_t tb = (cdXb * 0x01010101u) ^ ((cdXb ^ cdMb) << (8 * ((cds[1] >> (2 * q)) & 3u))); \
        sm.btab[r0 + q] = ((uint64_t)tb << 32) | ta;                                                   \
      }                                                                                                \
    }                                                                                                  \
    w.sync();                                                                                          \
  }

  // Rebase: slot 0 -> kT0; the padding columns of the lane are reset (see header); the neighbours'
  // rebase amounts update the base differences.
#define GAMX16_REBASE()                                                                                \
  {                                                                                                    \
    const uint32_t R = vsub2w(H[0], pack2(kT0, kT0));                                                  \
    _Pragma("unroll") for (int k = 0; k < C; k++) {                                                    \
      H[k] = vsub2w(H[k], R);                                                                          \
      if (k > kdl) H[k] = pack2(kT0, kT0);                                                             \
    }                                                                                                  \
    base[0] += half_lo(R); base[1] += half_hi(R);                                                      \
    const uint32_t Rl = (uint32_t)w.shfl_up((int)R, 1, LG), Rr = (uint32_t)w.shfl_down((int)R, 1, LG); \
    dL = vsub2w(vadd2w(dL, Rl), R);                                                                    \
    dR = vsub2w(vadd2w(dR, Rr), R);                                                                    \
  }

  // ---- first row, banded_smith_waterman.cc:112-132 (see bsw_warp.h), one job after the other ---------
  GAMX16_STAGE_TILE(0)
  {
    const uint64_t tb = sm.btab[LG - 1];  // row 0
    const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);
    const int kNone = -(1 << 28);
    int vt[2][C];  // true V of row 0 (with tag)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int sc[C], hloc[C];
      int run = kNone;
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int j = j0 + k, pos = p0[h] + j;
        const uint32_t cdp = prmt_sx(tlo, thi, sm.asel[gl * C + k]);
        const int cd = h ? half_hi(cdp) : half_lo(cdp);
        const bool valid = liveh[h] && pos >= 0 && pos < la[h] && j < Y;
        sc[k] = valid ? (cd & 0xff) : -1;                 // Cd byte of the cell, -1: never written
        const int s = cd >> SH;
        if (valid) run = (pos > 0 && j > 0) ? imax(run, s) : s;
        hloc[k] = run;
      }
      const bool restarts = liveh[h] && (-p0[h] >= j0) && (-p0[h] < j0 + C);  // the cell with pos == 0 is mine
      int incl = run;
      int cut = restarts;
#pragma unroll
      for (int d = 1; d < LG; d <<= 1) {
        const int o = w.shfl_up(incl, d, LG), oc = w.shfl_up(cut, d, LG);
        if (gl >= d) { if (!cut) incl = imax(incl, o); cut |= oc; }
      }
      int in = w.shfl_up(incl, 1, LG);
      if (gl == 0) in = kNone;
      bool open = true;
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int j = j0 + k, pos = p0[h] + j;
        int v;
        if (sc[k] < 0) {
          v = (beta * j) * (1 << SH);  // never written: 0
        } else {
          if (!(pos > 0 && j > 0)) open = false;
          const int cds = (int)(int8_t)(uint8_t)sc[k];
          const int s = cds >> SH;
          const int hh = open ? imax(hloc[k], in) : hloc[k];
          v = ((hh + beta * j) * (1 << SH)) | ((DIRS && hh == s) ? (cds & 3) : 0);
        }
        vt[h][k] = v;
      }
      base[h] = (vt[h][0] & ~(DIRS ? 3 : 0)) - kT0;
    }
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int ca = vt[0][k] & ~(DIRS ? 3 : 0), cb = vt[1][k] & ~(DIRS ? 3 : 0);
      H[k] = pack2(ca - base[0], cb - base[1]);
      if (DIRS) acc[k] = pack2(vt[0][k] & 3, vt[1][k] & 3);
      if (k > kdl) H[k] = pack2(kT0, kT0);
      if (tcap0[0] - k == gl) sm.cap[0][k * LG + gl] = ca;  // "last column" cell in row 0
      if (tcap0[1] - k == gl) sm.cap[1][k * LG + gl] = cb;
    }
    {
      const int bl0 = w.shfl_up(base[0], 1, LG), bl1 = w.shfl_up(base[1], 1, LG);
      const int br0 = w.shfl_down(base[0], 1, LG), br1 = w.shfl_down(base[1], 1, LG);
      dL = pack2(bl0 - base[0], bl1 - base[1]);
      dR = pack2(br0 - base[0], br1 - base[1]);
    }
    if (DIRS && T_total == 1) {  // a single row on a single lane: no step will flush its directions
#pragma unroll
      for (int k = 0; k < C; k++) if (live) fp[k * LG] = (acc[k] & 0x00030003u) << 14;
    }
  }

  // One step tt; PA points at the selector of this lane's slot 0, PB at its row table.
  // KIND 0 (FAST): every lane is on a row in [1, X-1] of both jobs, no cell with pos < 0, nothing to
  //                latch; the caller flushes.
  // KIND 1 (SLOW): lanes outside their row range keep their registers (per job), "last column" cells
  //                are latched, the direction flush is decided per step.
  // KIND 2 (PAD):  SLOW, and cells with pos < 0 are frozen (they hold the never-written zeros).
#define GAMX16_STEP(KIND, TT, PA, PB)                                                                  \
  {                                                                                                    \
    const int tt = (TT);                                                                               \
    const uint64_t tb = *(PB);                                                                         \
    const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);                                     \
    uint32_t left = (vadd2w((uint32_t)w.shfl_up((int)H[C - 1], 1, LG), dL) & lkeep) | lor;             \
    const bool started = (KIND) == 0 || (tt - gl >= 1);                                                \
    uint32_t keep = 0xffffffffu;                                                                       \
    int q0 = 0, q1 = 0;                                                                                \
    if ((KIND) != 0) {                                                                                 \
      keep = ((started && tt - gl < X[0]) ? 0xffffu : 0u) | ((started && tt - gl < X[1]) ? 0xffff0000u : 0u); \
      q0 = -p0[0] - gl * (C - 1) - tt;  /* slots k < q are cells with pos < 0 */                       \
      q1 = -p0[1] - gl * (C - 1) - tt;                                                                 \
    }                                                                                                  \
    const int dcap0 = tcap0[0] - tt, dcap1 = tcap0[1] - tt;                                            \
    uint32_t capv0 = 0, capv1 = 0;                                                                     \
    uint32_t right = 0;                                                                                \
    _Pragma("unroll") for (int k = 0; k < C; k++) {                                                    \
      const uint32_t cd = prmt_sx(tlo, thi, (PA)[k]);                                                  \
      const uint32_t up = (k == C - 1) ? right : H[(k + 1) % C];                                       \
      const uint32_t m = viaddmax2(up, U[k], left);                                                    \
      const uint32_t v = viaddmax2(H[k], cd, m);                                                       \
      const uint32_t hc = v & CLEAN;                                                                   \
      if (DIRS && started) acc[k] = (acc[k] * 4u + v) + neg1 * hc;                                     \
      if ((KIND) == 0) {                                                                               \
        H[k] = hc;                                                                                     \
      } else {                                                                                         \
        uint32_t kk = keep;                                                                            \
        if ((KIND) == 2) kk &= (k >= q0 ? 0xffffu : 0u) | (k >= q1 ? 0xffff0000u : 0u);                \
        H[k] = bitsel(kk, hc, H[k]);                                                                   \
        capv0 = (dcap0 == k) ? H[k] : capv0;                                                           \
        capv1 = (dcap1 == k) ? H[k] : capv1;                                                           \
      }                                                                                                \
      left = H[k];                                                                                     \
      if (k == 0) right = vadd2w((uint32_t)w.shfl_down((int)H[0], 1, LG), dR) & rkeep;                 \
    }                                                                                                  \
    if ((KIND) != 0) {                                                                                 \
      if ((unsigned)dcap0 < (unsigned)C) sm.cap[0][dcap0 * LG + gl] = half_lo(capv0) + base[0];        \
      if ((unsigned)dcap1 < (unsigned)C) sm.cap[1][dcap1 * LG + gl] = half_hi(capv1) + base[1];        \
      if (DIRS && ((tt & 7) == 7 || tt == T_total - 1)) {                                              \
        const int sh = 2 * (7 - (tt & 7));                                                             \
        _Pragma("unroll") for (int k = 0; k < C; k++) {                                                \
          if (live) fp[k * LG] = ((acc[k] << sh) & 0xffffu) | (((acc[k] >> 16) << sh) << 16);          \
          if (tt - gl >= 0) acc[k] = 0;  /* (a lane still before its row 0 keeps that row's tags) */    \
        }                                                                                              \
        fp += C * LG;                                                                                  \
      }                                                                                                \
    }                                                                                                  \
  }

  int t = 1;  // row 0 is done; lane gl starts its row 1 at step gl + 1
  int t0 = 0;
  while (t < T_total) {
    if (t >= t0 + kTileSteps) { t0 += kTileSteps; GAMX16_STAGE_TILE(t0) }
    const uint16_t* pa = sm.asel + (gl * (C - 1) - t0);    // pa[t]: selector of this lane's slot 0 at step t
    const uint64_t* pb = sm.btab + ((LG - 1) - gl - t0);   // pb[t]: table of row t - gl
    const int stop = imin(t0 + kTileSteps, T_total);       // first step this tile does not cover
    while (t < stop) {
      // steps [t, fast_hi) are steady state: every lane has started (t >= LG) and is on a row < X of
      // both jobs, no cell with pos < 0 is left (t >= pad_end), the last step (partial flush) is
      // excluded, no lane meets its "last column" cells
      int fast_hi = imin(stop, imin(x_min, T_total - 1));
      if (t <= win_hi) fast_hi = imin(fast_hi, win_lo);
      int nf = (t >= LG && t >= pad_end && (t & (UF - 1)) == 0) ? (fast_hi - t) / UF : 0;
      if (nf > 0) {
        const uint16_t* pa_t = pa + t;
        const uint64_t* pb_t = pb + t;
        do {
#pragma unroll
          for (int u = 0; u < UF; u++) GAMX16_STEP(0, t + u, pa_t + u, pb_t + u)
          t += UF; pa_t += UF; pb_t += UF;
          if (DIRS && (t & 7) == 0) {
#pragma unroll
            for (int k = 0; k < C; k++) { if (live) fp[k * LG] = acc[k]; acc[k] = 0; }  // (an idle group owns no scratch)
            fp += C * LG;
          }
          if ((t & (kRebaseSteps - 1)) == 0) GAMX16_REBASE()
        } while (--nf > 0);
      } else {
        const int slow_stop = imin(stop, (t & ~(UF - 1)) + UF);  // up to the next group boundary
        do {
          if (t < pad_end) GAMX16_STEP(2, t, pa + t, pb + t)
          else GAMX16_STEP(1, t, pa + t, pb + t)
          t++;
          if ((t & (kRebaseSteps - 1)) == 0) GAMX16_REBASE()
        } while (t < slow_stop);
      }
    }
  }
#undef GAMX16_STEP
#undef GAMX16_REBASE
#undef GAMX16_STAGE_TILE

  // ---- end-cell selection, .cc:174-212: last row (columns ascending) before last column, per job ----
#pragma unroll
  for (int h = 0; h < 2; h++) {
    EndBest best;
    best.found = 0; best.val = 0; best.ord = 0;
    const DevJob* Jp = Jh[h];
    if (liveh[h]) {
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int j = j0 + k;
        if (j >= Jp->jlo && j <= Jp->jhi) {
          const int vtrue = (h ? half_hi(H[k]) : half_lo(H[k])) + base[h];
          const int val = (j < Jp->jfill) ? ((vtrue >> SH) - beta * j) : 0;
          best.consider(val, j);
        }
      }
      if (kc[h] >= 0) {
#pragma unroll
        for (int k = 0; k < C; k++) {
          const int j = j0 + k, i = tcap0[h] - k - gl;  // the row this slot was on when it met pos == end_a
          if (i >= 0 && i < X[h] && j <= 2 * B && i >= Jp->col_imin) {
            const int val = Jp->col_zero ? 0 : ((sm.cap[h][k * LG + gl] >> SH) - beta * j);
            best.consider(val, Y + i);
          }
        }
      }
    }
#pragma unroll
    for (int d = LG / 2; d >= 1; d >>= 1) {
      const int of = w.shfl_xor(best.found, d, 32), ov = w.shfl_xor(best.val, d, 32), oo = w.shfl_xor(best.ord, d, 32);
      if (of) best.consider(ov, oo);
    }
    if (liveh[h] && gl == 0) {
      DevResult R;
      R.status = kStatusOk; R.score = 0; R.end_i = 0; R.end_j = 0;
      R.has_match = 1 + h;  // layout tag for the traceback kernel: half h of a 16x2 pair region
      R.n_ops = R.n_match = R.n_mismatch = R.n_gap_a = R.n_gap_b = 0;
      R.tail_gap_a = R.tail_gap_b = 0;
      R.begin_a = R.begin_bx = 0;
      R.first_match_a = R.first_match_x = R.last_match_a = R.last_match_x = 0;
      R.ops_start = 0;
      if (!best.found) {
        R.status = kStatusEmpty;  // .cc:215
      } else {
        const int ei = best.ord < Y ? X[h] - 1 : best.ord - Y;
        const int ej = best.ord < Y ? best.ord : kc[h] - ei;
        R.score = best.val; R.end_i = ei; R.end_j = ej;
        if (p0[h] + ei + ej >= la[h]) R.status = kStatusOutOfRange;  // first traceback step reads a.at(pos), .cc:231/:265
      }
      *(h ? outB : outA) = R;
    }
  }
  w.sync();  // the shared-memory tiles are free for the next pair
}

// Traceback fetcher for half `half` of a pair region: two consecutive 8-step blocks make the 16-step
// word k1_traceback_t expects (oldest tag on top).
struct PairFetch {
  const uint32_t* dirs;
  int C, LG, half;
  GAMX_HD uint32_t operator()(int blk, int k, int l) const {
    const uint32_t w0 = dirs[((uint32_t)(2 * blk) * (uint32_t)C + (uint32_t)k) * (uint32_t)LG + (uint32_t)l];
    const uint32_t w1 = dirs[((uint32_t)(2 * blk + 1) * (uint32_t)C + (uint32_t)k) * (uint32_t)LG + (uint32_t)l];
    return half ? ((w0 & 0xffff0000u) | (w1 >> 16)) : ((w0 << 16) | (w1 & 0xffffu));
  }
};

}  // namespace gamx
