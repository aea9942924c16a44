// K1 - warp-level banded overlap DP with inter-task parallelism: every group of LG lanes
// (LG = 32, 16 or 8; 1, 2 or 4 pairs per warp) owns one pair.
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
// diag > up > left on ties (.cc:272-307), and tag^1 is the edit op.  Per cell: one PRMT (Cd
// from an 8-byte per-row table indexed by the a-base), two VIADDMNMX, one LOP3 (strip the
// tag) on the ALU pipe, plus two IMADs on the FMA pipe that append the tag to the lane's
// direction word.  The score-only variant (DIRS=false) drops the tag handling: 3 ALU ops/cell.
//
// Step variants.  Steps run in unrolled groups of C (register renaming for the sliding a-window):
//   <CAPTURE=false, MASKED=false>  steady state
//   <CAPTURE=true,  *>             additionally latches the "last column" cells (pos == end_a,
//                                  .cc:197-212) of the anti-diagonals that contain them
//   <*, MASKED=true>               pipeline drain: lanes past their last row keep their registers
// and a general single step handles the pipeline fill (first row, .cc:112-132).
#pragma once
#include "bsw_common.h"
#include "bsw_traceback.h"

namespace gamx {

constexpr int kTileSteps = 128;  // steps per shared-memory sequence tile
constexpr int kMaxC = 18;        // widest lane stripe: band <= (32*18-1)/2 = 287

// A lane's stripe of C slots is split into C/S sub-blocks of S slots.  The a-bases of a sub-block
// live in S registers that are renamed (not moved) across an unrolled group of S steps; each
// sub-block fetches one new base per step from shared memory.  Unrolling by S instead of C keeps
// the steady-state loop body around 200-350 instructions for every C (an unroll by C grows as
// C^2 and overflows the instruction cache: measured "no instruction" stalls at C = 9).
// S = the largest divisor of C whose unrolled group stays within the budget (about 8 instructions
// per cell with directions, 3.3 without, plus per-step overhead).
GAMX_HD constexpr int sub_block_of(int c, bool dirs) {
  int best = 1;
  for (int sb = 1; sb <= c; sb++) {
    if (c % sb) continue;
    const int body = sb * ((dirs ? 80 : 33) * c / 10 + c / sb + 12);
    if (body <= (dirs ? 400 : 460)) best = sb;
  }
  return best;
}
// stripes whose sub-block would degenerate to S = 1 (one shared-memory load per cell and step:
// measured 2-3x slower) are rounded up to the next supported width by the host
GAMX_HD constexpr bool stripe_supported(int c) {
  return c >= 2 && c <= 18 && sub_block_of(c, true) >= 2 && sub_block_of(c, false) >= 2;
}

template <int C, int LG>
struct GroupSmem {
  // a-bases of the tile as PRMT selectors (0x7770 | code), b-rows as 8-byte Cd tables
  uint64_t btab[kTileSteps + LG];
  uint16_t asel[kTileSteps + LG * C + 2];
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
template <int C, int LG, bool DIRS, class W>
GAMX_HD void warp_align(W& w, const DevJob* Jp, const SeqStore& store, WarpSmem<C, LG>& wsm, uint32_t* dirs,
                        uint64_t group_stride, uint32_t* ops_buf, DevResult* out) {
  static_assert(stripe_supported(C), "lane stripe width");
  constexpr int S = sub_block_of(C, DIRS), NB = C / S;
  static_assert(NB * S == C, "stripe = whole sub-blocks");
  // LG <= 32: groups inside one warp (K1).  LG = 64..256: one pair per CTA (K2); the policy W then
  // implements the neighbour exchanges through shared memory and a block barrier.
  static_assert(LG == 8 || LG == 16 || LG == 32 || LG == 64 || LG == 128 || LG == 256, "lanes per pair");
  constexpr int SH = DIRS ? 2 : 0;
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

  // Direction words are accumulated as the difference of two multiply-add chains (both on the
  // FMA pipe, leaving the ALU pipe to the DP): accV = 4*accV + v, accH = 4*accH + (v & ~3);
  // accV - accH (mod 2^32) = the last 16 tags, oldest in the top bit pair.
  int H[C], capv[C];
  uint32_t A[C], accV[C], accH[C];
  int U[C];
#pragma unroll
  for (int k = 0; k < C; k++) {
    H[k] = 0; capv[k] = 0; A[k] = 0x7775u; accV[k] = 0; accH[k] = 0;
    U[k] = (gl == ld && k == kd) ? kBlock : (DIRS ? 1 : 0);
  }
  const int tcap0 = kc - gl * (C - 1);  // step at which slot 0 holds a "last column" cell (slot k: tcap0 - k)

  int t = 0;
  while (t < T_total) {
    // ---- stage the sequence tile for steps [t0, t0 + kTileSteps) -------------------------
    const int t0 = t;
    w.sync();
    {
      const int na = kTileSteps + LG * C - (LG - 1);
      const int pa0 = p0 + t0;
      for (int idx = gl; idx < na; idx += LG) {
        const int pos = pa0 + idx;
        const uint32_t code = (live && pos >= 0 && pos < la) ? load_code(store, va, pos) : (uint32_t)kCodePad;
        sm.asel[idx] = (uint16_t)(0x7770u | code);
      }
      const int nb = kTileSteps + LG - 1;
      for (int idx = gl; idx < nb; idx += LG) {
        const int i = t0 - (LG - 1) + idx;
        const uint32_t bc = (live && i >= 0 && i < X) ? load_code(store, vb, i) : (uint32_t)kCodePad;
        uint32_t lo, hi;
        if (bc < 4u) { lo = (cdX * 0x01010101u) ^ ((cdX ^ cdM) << (8 * bc)); hi = cdZ | (cdP << 8); }
        else if (bc == (uint32_t)kCodeN) { lo = cdZ * 0x01010101u; hi = cdM | (cdP << 8); }
        else { lo = cdP * 0x01010101u; hi = cdP | (cdP << 8); }
        sm.btab[idx] = ((uint64_t)hi << 32) | lo;
      }
    }
    w.sync();
    const uint16_t* pa = sm.asel + (gl * (C - 1) - t0);    // pa[t + k]: slot k's base at step t
    const uint64_t* pb = sm.btab + ((LG - 1) - gl - t0);   // pb[t]: table of row t - gl
    if (t0 == 0) {
      // bases of step 0 (the last slot of every sub-block is (re)loaded by the step itself)
#pragma unroll
      for (int k = 0; k < C; k++) A[k] = pa[k];
    }
    const int tile_end = imin(t0 + kTileSteps, T_total);

    // S unrolled steps starting at step t (every lane's row >= 1).
    // CAPTURE: latch "last column" cells; MASKED: lanes whose row is past X-1 keep their registers.
#define GAMX_STEP_GROUP(CAPTURE, MASKED)                                                              \
  {                                                                                                   \
    const int dcap = tcap0 - t;                                                                       \
    _Pragma("unroll") for (int u = 0; u < S; u++) {                                                   \
      const int tt = t + u;                                                                           \
      _Pragma("unroll") for (int sb = 0; sb < NB; sb++)                                               \
          A[sb * S + (u + S - 1) % S] = pa[tt + sb * S + S - 1];                                      \
      const uint64_t tb = pb[tt];                                                                     \
      const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);                                  \
      int left = w.shfl_up(H[C - 1], 1, LG);                                                          \
      if (gl == 0) left = kNegInf;                                                                    \
      const bool act = !(MASKED) || (tt - gl < X);                                                    \
      int right = 0;                                                                                  \
      _Pragma("unroll") for (int k = 0; k < C; k++) {                                                 \
        const int up = (k == C - 1) ? right : H[(k + 1) % C];                                         \
        const int cd = (int)prmt(tlo, thi, A[(k / S) * S + (u + k % S) % S]);                         \
        const int m = viaddmax(up, U[k], left);                                                       \
        const int v = viaddmax(H[k], cd, m);                                                          \
        int hc = v;                                                                                   \
        if (DIRS) {                                                                                   \
          hc = v & ~3;                                                                                \
          accV[k] = accV[k] * 4u + (uint32_t)v;                                                       \
          accH[k] = accH[k] * 4u + (uint32_t)hc;                                                      \
        }                                                                                             \
        if (MASKED) { if (act) H[k] = hc; } else H[k] = hc;                                           \
        if (CAPTURE) { if (dcap == u + k) capv[k] = H[k]; }                                           \
        left = H[k];                                                                                  \
        if (k == 0) right = w.shfl_down(H[0], 1, LG);                                                 \
      }                                                                                               \
      if (DIRS && (tt & 15) == 15) {                                                                  \
        _Pragma("unroll") for (int k = 0; k < C; k++) fp[k * LG] = accV[k] - accH[k];                  \
        fp += C * LG;                                                                                 \
      }                                                                                               \
    }                                                                                                 \
    t += S;                                                                                           \
  }

    while (t < tile_end) {
      // the warp's final step always takes the general path (it flushes the partial direction words)
      const bool grouped = (t >= LG) && (t + S <= tile_end) && (t + S - 1 <= T_total - 2);
      if (grouped) {
        const bool masked = t + S - 1 > x_min - 1;
        const bool capture = !(t + S - 1 < win_lo || t > win_hi);
        if (!masked && !capture) GAMX_STEP_GROUP(false, false)
        else if (!masked) GAMX_STEP_GROUP(true, false)
        else GAMX_STEP_GROUP(true, true)
        continue;
      }

      // ---- one general step: pipeline fill (first row), tile/tail remainders ---------------
      {
        const int i = t - gl;
        const bool act = live && (i >= 0) && (i < X);
#pragma unroll
        for (int sb = 0; sb < NB; sb++) A[sb * S + S - 1] = pa[t + sb * S + S - 1];
        const uint64_t tb = pb[t];
        const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);
        int left = w.shfl_up(H[C - 1], 1, LG);
        if (gl == 0) left = kNegInf;
        const bool row0 = act && i == 0;
        if (row0) {
          // first row, banded_smith_waterman.cc:112-132 (gap < every substitution score, so the
          // gap terms of .cc:122/.cc:130 never win and force_start changes nothing here)
          int lt = (left >> SH) - beta * (j0 - 1);  // true score of (0, j0-1)
#pragma unroll
          for (int k = 0; k < C; k++) {
            const int j = j0 + k, pos = p0 + j;
            int v;
            if (pos < 0 || pos >= la || j >= Y) {
              v = (beta * j) << SH;  // never-written cell: 0
            } else {
              const int cd = (int)prmt(tlo, thi, A[k]);
              const int s = (cd >> SH) - alpha;
              const int h = (pos > 0 && j > 0) ? imax(s, lt) : s;
              v = ((h + beta * j) << SH) | ((DIRS && h == s) ? (cd & 3) : 0);
              lt = h;
            }
            if (DIRS) { accV[k] = accV[k] * 4u + (uint32_t)v; accH[k] = accH[k] * 4u + (uint32_t)(v & ~3); H[k] = v & ~3; }
            else H[k] = v;
          }
        } else {
          const int cd = (int)prmt(tlo, thi, A[0]);
          const int m = viaddmax(H[1 % C], U[0], left);
          const int v = viaddmax(H[0], cd, m);
          const int hc = DIRS ? (v & ~3) : v;
          if (DIRS) { accV[0] = accV[0] * 4u + (uint32_t)v; accH[0] = accH[0] * 4u + (uint32_t)hc; }
          if (act) H[0] = hc;
        }
        const int right = w.shfl_down(H[0], 1, LG);
        if (!row0) {
#pragma unroll
          for (int k = 1; k < C; k++) {
            const int up = (k == C - 1) ? right : H[(k + 1) % C];
            const int cd = (int)prmt(tlo, thi, A[k]);
            const int m = viaddmax(up, U[k], H[k - 1]);
            const int v = viaddmax(H[k], cd, m);
            const int hc = DIRS ? (v & ~3) : v;
            if (DIRS) { accV[k] = accV[k] * 4u + (uint32_t)v; accH[k] = accH[k] * 4u + (uint32_t)hc; }
            if (act) H[k] = hc;
          }
        }
        // latch "last column" cells, .cc:197-212: slot k at step tcap0 - k
        const int dc = tcap0 - t;
#pragma unroll
        for (int k = 0; k < C; k++)
          if (dc == k) capv[k] = H[k];
#pragma unroll
        for (int k = 0; k < C; k++)
          if (k % S != S - 1) A[k] = A[k + 1];
        if (DIRS && ((t & 15) == 15 || t == T_total - 1)) {
          const int sh = 2 * (15 - (t & 15));
#pragma unroll
          for (int k = 0; k < C; k++) fp[k * LG] = (accV[k] - accH[k]) << sh;
          fp += C * LG;
        }
        t++;
      }
    }
#undef GAMX_STEP_GROUP
  }

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
          const int val = Jp->col_zero ? 0 : ((capv[k] >> SH) - alpha * i - beta * j);
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
      } else if (DIRS && !(Jp->mode & 0x100)) {  // 0x100: debug knob GAMX_DEBUG_SKIP_TRACEBACK
        k1_traceback<C, LG>(gdirs, ei, ej, p0, (Jp->mode & 0xff) == kModeFull, ops_buf + Jp->ops_word, Jp->ops_cap, R);
        R.ops_start = Jp->ops_word * 16 + Jp->ops_cap - R.n_ops;
      }
    }
    *out = R;
  }
}

}  // namespace gamx
