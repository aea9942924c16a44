// K1 - warp-per-pair banded overlap DP (inter-task parallelism: every warp owns one pair).
//
// Computes what BandedSmithWaterman::find_alignment computes
// (/root/reference/lib/src/alignment/banded_smith_waterman.cc:69-323) for the jobs the host
// classifies as "regular" (DESIGN.md 4.1): -29 <= gap <= -5, 2*band+1 <= 32*C, windows that
// start inside both contigs.  Everything else goes to the generic kernel (bsw_generic.h).
//
// Layout.  Band coordinates (i, j): row i <-> b[begin_b+i], column j <->
// a[pos], pos = begin_a - band + i + j.  Lane l owns the C consecutive band columns
// j = l*C .. l*C+C-1 ("slots"); at step t it processes row i = t - l, so the three
// dependencies of a cell
//      diag (i-1, j)    -> the lane's own register of the previous step
//      up   (i-1, j+1)  -> own register, or lane l+1's slot 0 of THIS step   (1 shuffle)
//      left (i,   j-1)  -> own register, or lane l-1's last slot of the PREVIOUS step (1 shuffle)
// cost two shuffles per C cells.  H lives in registers only; nothing but the 2-bit
// directions ever goes to memory.
//
// Cell update.  Stored value V = ((H + alpha*i + beta*j) << 2) | tag with
// beta = -gap, alpha = -2*gap, which makes both gap moves free:
//      V = max( diag + Cd , up + 1 , left )        Cd = ((S + alpha) << 2) | (2 + is_match)
// The low two bits of the max are the direction with exactly the reference's priority
// diag > up > left on ties (.cc:272-307), and tag^1 is the edit op.  Per cell: one PRMT (Cd
// from an 8-byte per-row table indexed by the a-base), two VIADDMNMX, one LOP3 (strip the
// tag) on the ALU pipe, plus two IMAD-class ops that append the tag to the lane's direction
// word.  The score-only variant (DIRS=false) drops the tag handling: 3 ALU ops per cell.
#pragma once
#include "bsw_common.h"
#include "bsw_traceback.h"

namespace gamx {

constexpr int kTileSteps = 256;  // steps per shared-memory sequence tile
constexpr int kMaxC = 17;        // widest lane stripe: band <= (32*17-1)/2 = 271

template <int C>
struct WarpSmem {
  // a-bases of the tile as PRMT selectors (0x7770 | code), b-rows as 8-byte Cd tables
  uint64_t btab[kTileSteps + 32];
  uint16_t asel[kTileSteps + 32 * C];
};

template <int C>
struct K1DirAt {
  const uint32_t* dirs;
  GAMX_HD int operator()(int x, int y) const {
    const int l = y / C, k = y - l * C, t = x + l;
    const uint32_t w = dirs[((size_t)(t >> 4) * C + k) * 32 + l];
    return (int)((w >> (2 * (15 - (t & 15)))) & 3u);
  }
};

// number of direction words one job needs in the K1 layout
GAMX_HD uint64_t k1_dir_words(int x, int band, int c) {
  const int ld = (2 * band) / c;
  const uint64_t steps = (uint64_t)x + ld;
  return ((steps + 15) / 16) * (uint64_t)c * 32;
}

struct EndBest {
  int found, val, ord;
  GAMX_HD void consider(int v, int o) {
    if (!found || v > val || (v == val && o < ord)) { found = 1; val = v; ord = o; }
  }
};

template <int C, bool DIRS, class W>
GAMX_HD void warp_align(W& w, const DevJob& J, const SeqStore& S, WarpSmem<C>& sm, uint32_t* dirs,
                        uint32_t* ops_buf, DevResult* out) {
  static_assert(C >= 2 && C <= kMaxC, "lane stripe width");
  constexpr int SH = DIRS ? 2 : 0;
  const int lane = w.lane();
  const int X = J.x, B = J.band, Y = 2 * B + 1;
  const int ld = (Y - 1) / C, kd = (Y - 1) - ld * C;  // lane/slot of band column 2B
  const int alpha = -2 * J.gap, beta = -J.gap;
  const int la = J.la, p0 = J.p0;
  const int T_total = X + ld;
  const int j0 = lane * C;

  // Cd bytes (see header comment)
  const uint32_t tagM = DIRS ? (uint32_t)kTagDiagMatch : 0u, tagX = DIRS ? (uint32_t)kTagDiagMis : 0u;
  const uint32_t cdM = ((uint32_t)(kScoreMatch + alpha) << SH) | tagM;
  const uint32_t cdX = ((uint32_t)(kScoreMismatch + alpha) << SH) | tagX;
  const uint32_t cdZ = ((uint32_t)alpha << SH) | tagM;  // N against anything: score 0, MATCH op
  const uint32_t cdP = ((uint32_t)alpha << SH) | tagX;  // padding: score 0

  int H[C];
  // Direction words are accumulated as the difference of two multiply-add chains (both on the
  // FMA pipe, leaving the ALU pipe to the DP): accV = 4*accV + v, accH = 4*accH + (v & ~3);
  // accV - accH (mod 2^32) = the last 16 tags, oldest in the top bit pair.
  uint32_t A[C], accV[C], accH[C];
  int U[C];
#pragma unroll
  for (int k = 0; k < C; k++) {
    H[k] = 0; A[k] = 0x7775u; accV[k] = 0; accH[k] = 0;
    U[k] = (lane == ld && k == kd) ? kBlock : (DIRS ? 1 : 0);
  }

  EndBest best;
  best.found = 0; best.val = 0; best.ord = 0;

  // "last column" cells (pos == end_a) are met in steps [win_lo, win_hi]
  const int kc = J.kc;
  const int win_lo = kc >= 0 ? kc - (ld + 1) * (C - 1) : 1, win_hi = kc >= 0 ? kc : 0;

  int t = 0;
  while (t < T_total) {
    // ---- stage the sequence tile for steps [t0, t0 + kTileSteps) -------------------------
    const int t0 = t;
    w.sync();
    {
      const int na = kTileSteps + 32 * C - 31;
      const int pa0 = p0 + t0;
      for (int idx = lane; idx < na; idx += 32) {
        const int pos = pa0 + idx;
        const uint32_t code = (pos >= 0 && pos < la) ? load_code(S, J.a, pos) : (uint32_t)kCodePad;
        sm.asel[idx] = (uint16_t)(0x7770u | code);
      }
      const int nb = kTileSteps + 31;
      for (int idx = lane; idx < nb; idx += 32) {
        const int i = t0 - 31 + idx;
        const uint32_t bc = (i >= 0 && i < X) ? load_code(S, J.b, i) : (uint32_t)kCodePad;
        uint32_t lo, hi;
        if (bc < 4u) { lo = (cdX * 0x01010101u) ^ ((cdX ^ cdM) << (8 * bc)); hi = cdZ | (cdP << 8); }
        else if (bc == (uint32_t)kCodeN) { lo = cdZ * 0x01010101u; hi = cdM | (cdP << 8); }
        else { lo = cdP * 0x01010101u; hi = cdP | (cdP << 8); }
        sm.btab[idx] = ((uint64_t)hi << 32) | lo;
      }
    }
    w.sync();
    const uint16_t* pa = sm.asel + (lane * (C - 1) + (C - 1) - t0);  // pa[t]: slot C-1's base at step t
    const uint64_t* pb = sm.btab + (31 - lane - t0);                 // pb[t]: table of row t - lane
    if (t0 == 0) {
      // slots 0..C-2 of step 0
#pragma unroll
      for (int k = 0; k < C - 1; k++) A[k] = sm.asel[lane * (C - 1) + k];
    }
    const int tile_end = imin(t0 + kTileSteps, T_total);

    while (t < tile_end) {
      // the final step always takes the general path (it flushes the partial direction words)
      const bool fast = (t >= 32) && (t + C - 1 <= X - 1) && (t + C - 1 <= T_total - 2) &&
                        (t + C <= tile_end) && (t + C - 1 < win_lo || t > win_hi);
      if (fast) {
        // ---- C steps, every lane on a row in [1, X-1], no candidate capture ----------------
#pragma unroll
        for (int u = 0; u < C; u++) {
          const int tt = t + u;
          A[(u + C - 1) % C] = pa[tt];
          const uint64_t tb = pb[tt];
          const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);
          int left = w.shfl_up(H[C - 1], 1);
          if (lane == 0) left = kNegInf;
          int right = 0;
#pragma unroll
          for (int k = 0; k < C; k++) {
            const int up = (k == C - 1) ? right : H[(k + 1) % C];
            const int cd = (int)prmt(tlo, thi, A[(u + k) % C]);
            const int m = viaddmax(up, U[k], left);
            const int v = viaddmax(H[k], cd, m);
            if (DIRS) {
              const int hc = v & ~3;
              accV[k] = accV[k] * 4u + (uint32_t)v;
              accH[k] = accH[k] * 4u + (uint32_t)hc;
              H[k] = hc;
            } else {
              H[k] = v;
            }
            left = H[k];
            if (k == 0) right = w.shfl_down(H[0], 1);
          }
          if (DIRS && (tt & 15) == 15) {
#pragma unroll
            for (int k = 0; k < C; k++) dirs[((uint32_t)(tt >> 4) * C + k) * 32 + lane] = accV[k] - accH[k];
          }
        }
        t += C;
        continue;
      }

      // ---- one general step: pipeline fill/drain, first row, candidate capture -------------
      {
        const int i = t - lane;
        const bool act = (i >= 0) && (i < X);
        A[C - 1] = pa[t];
        const uint64_t tb = pb[t];
        const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);
        int left = w.shfl_up(H[C - 1], 1);
        if (lane == 0) left = kNegInf;
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
        const int right = w.shfl_down(H[0], 1);
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
        // "last column" candidates, .cc:197-212: the cell of this row with i + j == kc
        if (act && kc >= 0 && i >= J.col_imin) {
#pragma unroll
          for (int k = 0; k < C; k++) {
            const int j = j0 + k;
            if (i + j == kc && j <= 2 * B) {
              const int val = J.col_zero ? 0 : ((H[k] >> SH) - alpha * i - beta * j);
              best.consider(val, Y + i);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < C - 1; k++) A[k] = A[k + 1];
        if (DIRS && ((t & 15) == 15 || t == T_total - 1)) {
          const int sh = 2 * (15 - (t & 15));
#pragma unroll
          for (int k = 0; k < C; k++) dirs[((uint32_t)(t >> 4) * C + k) * 32 + lane] = (accV[k] - accH[k]) << sh;
        }
        t++;
      }
    }
  }

  // ---- end-cell selection, .cc:174-212: last row (columns ascending) before last column ----
#pragma unroll
  for (int k = 0; k < C; k++) {
    const int j = j0 + k;
    if (j >= J.jlo && j <= J.jhi) {
      const int val = (j < J.jfill) ? ((H[k] >> SH) - alpha * (X - 1) - beta * j) : 0;
      best.consider(val, j);
    }
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const int of = w.shfl_xor(best.found, d), ov = w.shfl_xor(best.val, d), oo = w.shfl_xor(best.ord, d);
    if (of) best.consider(ov, oo);
  }
  w.sync();  // direction words of all lanes are visible to lane 0

  if (lane == 0) {
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
      } else if (DIRS) {
        k1_traceback<C>(dirs, ei, ej, p0, J.mode == kModeFull, ops_buf + J.ops_word, J.ops_cap, R);
        R.ops_start = J.ops_word * 16 + J.ops_cap - R.n_ops;
      }
    }
    *out = R;
  }
}

}  // namespace gamx
