// K1 - warp-level banded overlap DP with inter-task parallelism: every group of LG lanes
// (LG = 32, 16, 8 or 4; 1, 2, 4 or 8 pairs per warp) owns one pair.
// K2 - the same skewed anti-diagonal wavefront with LG = 64, 128 or 256 lanes: one pair per CTA
// (wide bands, or few long pairs); only the exchange policy W differs.
//
// Computes what BandedSmithWaterman::find_alignment computes
// (/root/reference/lib/src/alignment/banded_smith_waterman.cc:69-323) for the jobs the host
// classifies as "regular" (DESIGN.md 4): -29 <= gap <= -5, 2*band+1 <= LG*C, windows that
// start inside both contigs.  Everything else goes to the generic kernel (bsw_generic.h).
//
// Layout.  Band coordinates (i, j): row i <-> b[begin_b+i], column j <-> a[pos],
// pos = begin_a - band + i + j.  Lane gl of a group owns the C consecutive band columns
// j = gl*C .. gl*C+C-1 ("slots"); at step t it processes row i = t - gl, so the three
// dependencies of a cell
//      diag (i-1, j)    -> the lane's own register of the previous step
//      up   (i-1, j+1)  -> own register, or lane gl+1's slot 0 of THIS step      (1 shuffle)
//      left (i,   j-1)  -> own register, or lane gl-1's last slot of the PREVIOUS step (1 shuffle)
// cost two shuffles per C cells.  H lives in registers only; nothing but the 2-bit
// directions ever goes to memory.
//
// Cell update.  Stored value V = ((H + alpha*i + beta*j) << 2) | tag with beta = -gap,
// alpha = -2*gap, which makes both gap moves free:
//      V = max( diag + Cd , up + 1 , left )        Cd = ((S + alpha) << 2) | (2 + is_match)
// The low two bits of the max are the direction with exactly the reference's priority
// diag > up > left on ties (.cc:272-307), and tag^1 is the edit op.
//
// Instruction mix per cell (the ALU pipe is the binding resource):
//   ALU pipe : VIADDMNMX  m = max(up + U, left)          FMA pipe : IDP.4A  d = diag + Cd
//              VIMNMX     v = max(d, m)                             IMAD    acc = 4*acc + v
//              LOP3       h = v & ~3                                IMAD    acc = acc + (-1)*h
//              1/4 PRMT   four Cd bytes at once
// Cd comes from an 8-byte per-row table (indexed by the a-base; N and padding need no branch):
// the a-bases of the tile sit in shared memory as overlapping 4-nibble windows, so one LDS.U16 is
// the PRMT selector of four consecutive slots and one PRMT yields their four Cd bytes; IDP.4A with
// a one-hot multiplier adds byte k to the diagonal.  acc is the lane's direction word (16 tags, the
// oldest in the top bit pair).  Score-only (DIRS=false) keeps VIADDMNMX + VIMNMX + 1/4 PRMT on the ALU pipe.
//
// The first row (.cc:112-132: a running maximum that takes its left neighbour without the gap
// penalty) is computed for all lanes at once by a prefix-max scan before the step loop.  Steps
// then run in two variants (a small code footprint matters: the kernel is issue-bound and the
// instruction cache is shared by warps in different phases):
//   FAST  steady state, unrolled groups of UF steps starting at multiples of UF: every lane on a row
//         in [1, X-1], nothing to latch; the only per-group overhead is two shared-memory pointer
//         bumps, the group counter and the flush test
//   SLOW  one step at a time: pipeline fill / drain, tile and job tails: lanes outside their row range
//         keep their registers; also latches the "last column" cells (pos == end_a, .cc:197-212)
#pragma once
#include "bsw_common.h"
#include "bsw_traceback.h"

namespace gamx {

constexpr int kTileSteps = 128;  // steps per shared-memory sequence tile
constexpr int kMaxC = 18;        // widest lane stripe: band <= (32*18-1)/2 = 287 (19 and 20 were tried: no gain)

GAMX_HD constexpr bool stripe_supported(int c) { return c >= 2 && c <= kMaxC; }
// steps per unrolled steady-state group (divides 16, so that a direction flush falls on a group end):
// keeps the loop body around 150-350 instructions
#ifndef GAMX_UF4_MAX_C
#define GAMX_UF4_MAX_C 12
#endif
GAMX_HD constexpr int unroll_of(int c) { return c <= GAMX_UF4_MAX_C ? 4 : 2; }

struct alignas(16) Quad { uint32_t v[4]; };

// Selector windows are 16-bit and lane gl reads window gl*(C-1) + t + 4q, so the lane stride is
// (C-1)/2 words: stripe widths with C-1 = 8, 12 or 16 read them with 4-, 2- and 8-way bank conflicts,
// and the LSU pipe (one wavefront per cycle per SM for all four schedulers) then limits the kernel.
// The host avoids those widths (bsw_host.h: 13 -> 14, 17 -> 18, and bands that would take 16 lanes x 9
// slots take 8 lanes x 18 instead).  (A padded layout for C = 9 was tried and measured no gain.)
template <int C>
GAMX_HD constexpr int asel_phys(int w) { return w; }

template <int C, int LG>
struct alignas(16) GroupSmem {
  uint64_t btab[(kTileSteps + LG + 7) / 8 * 8];  // b-rows of the tile as 8-byte Cd tables (staged in runs of 8)
  uint16_t asel[kTileSteps + LG * C + 16];   // a-bases as 4-nibble windows: codes of positions p..p+3
  int cap[C * LG];                           // latched "last column" cells, [slot][lane]
};
template <int C, int LG>
struct WarpSmem {
  GroupSmem<C, LG> g[LG >= 32 ? 1 : 32 / LG];
};

// number of direction words one job needs in the K1 layout (a few extra step blocks because the
// groups of a warp share one step loop: the host sizes the scratch for the largest job of a launch)
GAMX_HD uint64_t k1_dir_words(int x, int band, int c, int lg) {
  (void)band;
  const uint64_t steps = (uint64_t)x + lg + 16;
  return ((steps + 15) / 16) * (uint64_t)c * lg;
}

struct EndBest {
  int found, val, ord;
  GAMX_HD void consider(int v, int o) {
    if (!found || v > val || (v == val && o < ord)) { found = 1; val = v; ord = o; }
  }
};

// One warp: 32/LG jobs.  Jp / out are per-lane arguments, uniform within a group of LG lanes:
// the group's job (null: idle group) and its result slot.  dirs: the warp's direction scratch,
// group g uses dirs + g*group_stride.
// TB: walk the stored directions right here (group leader lane).  The product kernels pass false: they
// only record the end cell, and the traceback kernel (one job per thread) finishes the results.
template <int C, int LG, bool DIRS, bool TB = true, class W>
GAMX_HD void warp_align(W& w, const DevJob* Jp, const SeqStore& store, WarpSmem<C, LG>& wsm, uint32_t* dirs,
                        uint64_t group_stride, uint32_t* ops_buf, DevResult* out) {
  static_assert(stripe_supported(C), "lane stripe width");
  // LG <= 32: groups inside one warp (K1).  LG = 64..256: one pair per CTA (K2); the policy W then
  // implements the neighbour exchanges through shared memory and a block barrier.
  static_assert(LG == 4 || LG == 8 || LG == 16 || LG == 32 || LG == 64 || LG == 128 || LG == 256, "lanes per pair");
  constexpr int SH = DIRS ? 2 : 0;
  constexpr int UF = unroll_of(C);
  constexpr int NQ = (C + 3) / 4;  // 4-slot selector groups
  const int lane = w.lane();
  const int grp = lane / LG, gl = lane % LG;
  const bool live = Jp != nullptr;
  GroupSmem<C, LG>& sm = wsm.g[grp];
  uint32_t* gdirs = DIRS ? dirs + (uint64_t)grp * group_stride : nullptr;
  uint32_t* fp = DIRS ? gdirs + gl : nullptr;  // running flush pointer: step blocks are written in order

  const int X = live ? Jp->x : 0, B = live ? Jp->band : 0, Y = 2 * B + 1;
  const int ld = (Y - 1) / C, kd = (Y - 1) - ld * C;  // lane/slot of band column 2B
  const int gap = live ? Jp->gap : -8;
  const int alpha = -2 * gap, beta = -gap;
  const int la = live ? Jp->la : 0, p0 = live ? Jp->p0 : 0;
  const int kc = live ? Jp->kc : -1;
  const int j0 = gl * C;
  SeqView va, vb;
  va.origin = vb.origin = 0; va.dir = vb.dir = 1; va.comp = vb.comp = 0;
  if (live) { va = Jp->a; vb = Jp->b; }

  // warp-uniform extents
  int t_end = live ? X + ld : 0;     // this group's last step + 1
  int x_min = live ? X : 0x7fffffff;
  int win_lo = (live && kc >= 0) ? kc - (ld + 1) * (C - 1) : 0x7fffffff;
  int win_hi = (live && kc >= 0) ? kc : -1;
#pragma unroll
  for (int d = LG; d < 32; d <<= 1) {
    t_end = imax(t_end, w.shfl_xor(t_end, d, 32));
    x_min = imin(x_min, w.shfl_xor(x_min, d, 32));
    win_lo = imin(win_lo, w.shfl_xor(win_lo, d, 32));
    win_hi = imax(win_hi, w.shfl_xor(win_hi, d, 32));
  }
  const int T_total = t_end;

  // Cd bytes (see header comment)
  const uint32_t tagM = DIRS ? (uint32_t)kTagDiagMatch : 0u, tagX = DIRS ? (uint32_t)kTagDiagMis : 0u;
  const uint32_t cdM = ((uint32_t)(kScoreMatch + alpha) << SH) | tagM;
  const uint32_t cdX = ((uint32_t)(kScoreMismatch + alpha) << SH) | tagX;
  const uint32_t cdZ = ((uint32_t)alpha << SH) | tagM;  // N against anything: score 0, MATCH op
  const uint32_t cdP = ((uint32_t)alpha << SH) | tagX;  // padding: score 0

  // Direction words are accumulated with two multiply-adds (both on the FMA pipe, leaving the ALU
  // pipe to the DP): acc = 4*acc + v, acc += neg1 * (v & ~3), i.e. acc = 4*acc + tag (mod 2^32): the
  // last 16 tags, oldest in the top bit pair.  neg1 is -1 in a register the compiler cannot see
  // through, so the second one stays a multiply-add instead of becoming an ALU-pipe subtraction.
  int H[C];
  uint32_t acc[C];
  int U[C];
  const uint32_t neg1 = (uint32_t)(gap >> 31);  // gap < 0 for every job of this kernel
#pragma unroll
  for (int k = 0; k < C; k++) {
    H[k] = 0; acc[k] = 0;
    U[k] = (gl == ld && k == kd) ? kBlock : (DIRS ? 1 : 0);
  }
  const int lneg = gl == 0 ? kNegInf : 0;  // OR-ed into the shuffled left neighbour: band column 0 has none
  const int tcap0 = kc - gl * (C - 1);  // step at which slot 0 holds a "last column" cell (slot k: tcap0 - k)

  // Stages the sequence tile for steps [t0, t0 + kTileSteps): a-bases as 4-nibble windows, b-rows as
  // 8-byte Cd tables.  A lane fetches runs of 16 view positions with four word loads (load_codes16)
  // and expands them with bit operations; N and out-of-range positions (pad symbol) are patched in a
  // rare slow path.  Eight windows = one 16-byte store.
#define GAMX_STAGE_TILE(T0)                                                                           \
  {                                                                                                   \
    w.sync();                                                                                         \
    const int na = kTileSteps + LG * C - (LG - 1) + 3;                                                \
    const int pa0 = p0 + (T0);                                                                        \
    for (int c0 = gl * 8; c0 < na; c0 += LG * 8) {                                                    \
      const int pos0 = pa0 + c0;           /* windows c0..c0+7 need positions pos0 .. pos0+10 */      \
      uint32_t codes = 0, nfl = 0, valid = 0;                                                         \
      if (live && pos0 + 10 >= 0 && pos0 < la) {                                                      \
        load_codes16(store, va, pos0, la, &codes, &nfl);                                              \
        const int vlo = pos0 < 0 ? -pos0 : 0, vhi = imin(11, la - pos0);                              \
        valid = ((1u << vhi) - 1u) & ~((1u << vlo) - 1u);                                             \
      }                                                                                               \
      const uint32_t cx = va.comp * 0x11111111u;                                                      \
      uint32_t nlo = spread2to4(codes & 0xffffu) ^ cx;                                                \
      uint32_t nhi = (spread2to4((codes >> 16) & 0x3fu) ^ cx) & 0xfffu;                               \
      const uint32_t special = (nfl | ~valid) & 0x7ffu;                                               \
      if (special) {                                                                                  \
        uint64_t nib = ((uint64_t)nhi << 32) | nlo;                                                   \
        _Pragma("unroll") for (int q = 0; q < 11; q++) {                                              \
          if ((special >> q) & 1u) {                                                                  \
            const uint64_t code = ((valid >> q) & 1u) ? (uint64_t)kCodeN : (uint64_t)kCodePad;        \
            nib = (nib & ~((uint64_t)0xf << (4 * q))) | (code << (4 * q));                            \
          }                                                                                           \
        }                                                                                             \
        nlo = (uint32_t)nib; nhi = (uint32_t)(nib >> 32);                                             \
      }                                                                                               \
      uint32_t ow[4];                                                                                 \
      _Pragma("unroll") for (int j = 0; j < 4; j++) {                                                 \
        const uint32_t x = j == 0 ? nlo : funnel_r(nlo, nhi, 8u * j);   /* nibbles 2j .. 2j+7 */      \
        ow[j] = (x & 0xffffu) | ((x << 12) & 0xffff0000u);               /* windows 2j, 2j+1 */        \
      }                                                                                               \
      Quad qv;                                                                                        \
      qv.v[0] = ow[0]; qv.v[1] = ow[1]; qv.v[2] = ow[2]; qv.v[3] = ow[3];                             \
      *reinterpret_cast<Quad*>(sm.asel + c0) = qv;    /* c0 % 8 == 0: 16-byte aligned */              \
    }                                                                                                 \
    const int nb = kTileSteps + LG - 1;                                                               \
    for (int r0 = gl * 8; r0 < nb; r0 += LG * 8) {                                                    \
      const int i0 = (T0) - (LG - 1) + r0;  /* table r0+q is row i0+q */                              \
      uint32_t codes = 0, nfl = 0, valid = 0;                                                         \
      if (live && i0 + 7 >= 0 && i0 < X) {                                                            \
        load_codes16(store, vb, i0, X, &codes, &nfl);                                                 \
        const int vlo = i0 < 0 ? -i0 : 0, vhi = imin(8, X - i0);                                      \
        valid = ((1u << vhi) - 1u) & ~((1u << vlo) - 1u);                                             \
      }                                                                                               \
      codes ^= vb.comp * 0x5555u;                                                                     \
      _Pragma("unroll") for (int q = 0; q < 8; q++) {                                                 \
        const uint32_t bc = (codes >> (2 * q)) & 3u;                                                  \
        uint32_t lo = (cdX * 0x01010101u) ^ ((cdX ^ cdM) << (8 * bc)), hi = cdZ | (cdP << 8);         \
        if (!((valid >> q) & 1u)) { lo = cdP * 0x01010101u; hi = cdP | (cdP << 8); }                  \
        else if ((nfl >> q) & 1u) { lo = cdZ * 0x01010101u; hi = cdM | (cdP << 8); }                  \
        sm.btab[r0 + q] = ((uint64_t)hi << 32) | lo;                                                  \
      }                                                                                               \
    }                                                                                                 \
    w.sync();                                                                                         \
  }

  // ---- first row, banded_smith_waterman.cc:112-132 --------------------------------------------------
  // h(0,j) = (pos > 0 && j > 0) ? max(S, h(0,j-1)) : S for filled cells (gap < every substitution
  // score, so the gap terms of .cc:122/.cc:130 never win and force_start changes nothing here):
  // a prefix maximum of S that restarts at the cell with pos == 0.  Never-written cells are 0.
  GAMX_STAGE_TILE(0)
  {
    const uint64_t tb = sm.btab[LG - 1];  // row 0
    const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);
    const int kNone = -(1 << 28);
    int sc[C];
    int run = kNone;  // running maximum inside the lane
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int j = j0 + k, pos = p0 + j;
      const uint32_t cd4 = prmt(tlo, thi, sm.asel[asel_phys<C>(gl * C + 4 * (k / 4))]);
      const int cd = (int)((cd4 >> (8 * (k % 4))) & 0xffu);
      const bool valid = live && pos >= 0 && pos < la && j < Y;
      sc[k] = valid ? cd : -1;                        // Cd byte of the cell, -1: never written
      const int s = (cd >> SH) - alpha;
      if (valid) run = (pos > 0 && j > 0) ? imax(run, s) : s;
      H[k] = run;                                      // local prefix maximum (provisional)
    }
    // exclusive prefix maximum of the lanes' totals; a lane holding the pos == 0 cell restarts it
    const bool restarts = live && (-p0 >= j0) && (-p0 < j0 + C);  // the cell with pos == 0 is mine
    int incl = run;       // inclusive scan value
    int cut = restarts;   // 1: nothing from lower lanes may pass through this lane
#pragma unroll
    for (int d = 1; d < LG; d <<= 1) {
      const int o = w.shfl_up(incl, d, LG), oc = w.shfl_up(cut, d, LG);
      if (gl >= d) { if (!cut) incl = imax(incl, o); cut |= oc; }
    }
    int in = w.shfl_up(incl, 1, LG);
    if (gl == 0) in = kNone;
    bool open = true;  // the incoming maximum applies until the lane's own restart cell
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int j = j0 + k, pos = p0 + j;
      int v;
      if (sc[k] < 0) {
        v = (beta * j) << SH;  // never written: 0
      } else {
        if (!(pos > 0 && j > 0)) open = false;
        const int s = (sc[k] >> SH) - alpha;
        const int h = open ? imax(H[k], in) : H[k];
        v = ((h + beta * j) << SH) | ((DIRS && h == s) ? (sc[k] & 3) : 0);
      }
      if (DIRS) acc[k] = (uint32_t)(v & 3);
      H[k] = DIRS ? (v & ~3) : v;
      if (tcap0 - k == gl) sm.cap[k * LG + gl] = H[k];  // "last column" cell in row 0
    }
    if (DIRS && T_total == 1) {  // a single row on a single lane: no step will flush its directions
#pragma unroll
      for (int k = 0; k < C; k++) if (live) fp[k * LG] = acc[k] << 30;
    }
  }

  // The Cd bytes of one step: PA / PB point at the selector windows / row table of that step.
  // PA + PHYS(J0 + 4q) is the selector window of slots 4q.. of that step (PHYS: asel_phys relative
  // to a window index that is a multiple of 4, so that it is a compile-time constant in the FAST path)
#define GAMX_LOAD_CD(CD, PA, J0, PB)                                                                  \
  {                                                                                                   \
    const uint64_t tb = *(PB);                                                                        \
    const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);                                    \
    _Pragma("unroll") for (int q = 0; q < NQ; q++) (CD)[q] = prmt(tlo, thi, (PA)[asel_phys<C>((J0) + 4 * q)]); \
  }
  // One step tt with its Cd bytes in CD.  SLOW: lanes whose row is outside [1, X-1] keep their
  // registers (direction words keep shifting once a lane has started), "last column" cells are
  // latched and the direction flush is decided per step; the FAST variant leaves the flush to its caller.
#define GAMX_STEP(SLOW, TT, CD)                                                                       \
  {                                                                                                   \
    const int tt = (TT);                                                                              \
    int left = w.shfl_up(H[C - 1], 1, LG) | lneg;                                                     \
    const bool started = !(SLOW) || (tt - gl >= 1);                                                   \
    const bool act = !(SLOW) || (started && tt - gl < X);                                             \
    const int dcap = tcap0 - tt;  /* the slot that is on a "last column" cell in this step */         \
    int capv = 0;                                                                                     \
    int right = 0;                                                                                    \
    _Pragma("unroll") for (int k = 0; k < C; k++) {                                                   \
      const int up = (k == C - 1) ? right : H[(k + 1) % C];                                           \
      const int d = add_byte((CD)[k / 4], k % 4, H[k]);                                               \
      const int m = viaddmax(up, U[k], left);                                                         \
      const int v = imax(d, m);                                                                       \
      int hc = v;                                                                                     \
      if (DIRS) {                                                                                     \
        hc = v & ~3;                                                                                  \
        if (started) acc[k] = (acc[k] * 4u + (uint32_t)v) + neg1 * (uint32_t)hc;                      \
      }                                                                                               \
      if (act) H[k] = hc;                                                                             \
      if (SLOW) capv = (dcap == k) ? H[k] : capv;                                                     \
      left = H[k];                                                                                    \
      if (k == 0) right = w.shfl_down(H[0], 1, LG);                                                   \
    }                                                                                                 \
    if ((SLOW) && (unsigned)dcap < (unsigned)C) sm.cap[dcap * LG + gl] = capv;                         \
    if ((SLOW) && DIRS && ((tt & 15) == 15 || tt == T_total - 1)) {                                   \
      const int sh = 2 * (15 - (tt & 15));                                                            \
      _Pragma("unroll") for (int k = 0; k < C; k++) if (live) fp[k * LG] = acc[k] << sh;               \
      fp += C * LG;                                                                                   \
    }                                                                                                 \
  }

  int t = 1;  // row 0 is done; lane gl starts its row 1 at step gl + 1
  int t0 = 0;
  while (t < T_total) {
    if (t >= t0 + kTileSteps) { t0 += kTileSteps; GAMX_STAGE_TILE(t0) }
    const int wl = gl * (C - 1);                           // window of this lane's slot 0 at step t0
    const uint64_t* pb = sm.btab + ((LG - 1) - gl - t0);   // pb[t]: table of row t - gl
    const int stop = imin(t0 + kTileSteps, T_total);       // first step this tile does not cover
    while (t < stop) {
      // steps [t, fast_hi) are steady state: every lane has started (t >= LG) and is on a row < X,
      // the job's last step (partial flush) is excluded, no lane meets its "last column" cells
      int fast_hi = imin(stop, imin(x_min, T_total - 1));
      if (t <= win_hi) fast_hi = imin(fast_hi, win_lo);
      int nf = (t >= LG && (t & (UF - 1)) == 0) ? (fast_hi - t) / UF : 0;
      if (nf > 0) {
        // software pipeline: the shared-memory loads and byte permutes of step t+1 are issued before
        // the cells of step t, so their latency never sits in front of a step's dependent chain
        // (the look-ahead of a group's last step reads one entry past the group: inside the arrays,
        // and overwritten before use when it belongs to the next tile)
        const uint16_t* pa_t = sm.asel + asel_phys<C>(wl + (t - t0));
        const uint64_t* pb_t = pb + t;
        uint32_t cdn[NQ];
        GAMX_LOAD_CD(cdn, pa_t, 0, pb_t)
        do {
#pragma unroll
          for (int u = 0; u < UF; u++) {
            uint32_t cd4[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++) cd4[q] = cdn[q];
            GAMX_LOAD_CD(cdn, pa_t, u + 1, pb_t + u + 1)
            GAMX_STEP(false, t + u, cd4)
          }
          t += UF; pa_t += asel_phys<C>(UF); pb_t += UF;
          if (DIRS && (t & 15) == 0) {
#pragma unroll
            for (int k = 0; k < C; k++) if (live) fp[k * LG] = acc[k];  // (an idle group owns no scratch)
            fp += C * LG;
          }
        } while (--nf > 0);
      } else {
        const int slow_stop = imin(stop, (t & ~(UF - 1)) + UF);  // up to the next group boundary
        do {
          uint32_t cd4[NQ];
          {  // any t: the physical index is computed at run time
            const uint64_t tb = pb[t];
            const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);
            const int w0 = wl + (t - t0);
#pragma unroll
            for (int q = 0; q < NQ; q++) cd4[q] = prmt(tlo, thi, sm.asel[asel_phys<C>(w0 + 4 * q)]);
          }
          GAMX_STEP(true, t, cd4)
          t++;
        } while (t < slow_stop);
      }
    }
  }
#undef GAMX_STEP
#undef GAMX_LOAD_CD
#undef GAMX_STAGE_TILE

  // ---- end-cell selection, .cc:174-212: last row (columns ascending) before last column ----
  EndBest best;
  best.found = 0; best.val = 0; best.ord = 0;
  if (live) {
#pragma unroll
    for (int k = 0; k < C; k++) {
      const int j = j0 + k;
      if (j >= Jp->jlo && j <= Jp->jhi) {
        const int val = (j < Jp->jfill) ? ((H[k] >> SH) - alpha * (X - 1) - beta * j) : 0;
        best.consider(val, j);
      }
    }
    if (kc >= 0) {
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int j = j0 + k, i = tcap0 - k - gl;  // the row this slot was on when it met pos == end_a
        if (i >= 0 && i < X && j <= 2 * B && i >= Jp->col_imin) {
          const int val = Jp->col_zero ? 0 : ((sm.cap[k * LG + gl] >> SH) - alpha * i - beta * j);
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
  w.sync();  // direction words of all lanes are visible to the group leaders

  if (live && gl == 0) {
    DevResult R;
    R.status = kStatusOk; R.score = 0; R.end_i = 0; R.end_j = 0; R.has_match = 0;
    R.n_ops = R.n_match = R.n_mismatch = R.n_gap_a = R.n_gap_b = 0;
    R.tail_gap_a = R.tail_gap_b = 0;
    R.begin_a = R.begin_bx = 0;
    R.first_match_a = R.first_match_x = R.last_match_a = R.last_match_x = 0;
    R.ops_start = 0;
    if (!best.found) {
      R.status = kStatusEmpty;  // .cc:215
    } else {
      const int ei = best.ord < Y ? X - 1 : best.ord - Y;
      const int ej = best.ord < Y ? best.ord : kc - ei;
      R.score = best.val; R.end_i = ei; R.end_j = ej;
      if (p0 + ei + ej >= la) {
        R.status = kStatusOutOfRange;  // first traceback step reads a.at(pos), .cc:231/:265
      } else if (DIRS && TB && !(Jp->mode & 0x100)) {  // 0x100: debug knob GAMX_DEBUG_SKIP_TRACEBACK
        k1_traceback(gdirs, C, LG, ei, ej, p0, (Jp->mode & 0xff) == kModeFull, ops_buf + Jp->ops_word, Jp->ops_cap, R);
        R.ops_start = Jp->ops_word * 16 + Jp->ops_cap - R.n_ops;
      }
    }
    *out = R;
  }
}

}  // namespace gamx
