// Device traceback: walks the stored 2-bit directions from the selected end cell and
// produces everything the reference derives from its edit string.
//
// Restates the loop of banded_smith_waterman.cc:227-311 (moves) and the reductions of
// my_alignment.cc:167-296 (first/last match, gap counts).  Because the direction of every
// cell was decided at fill time with the reference's tie-breaking (diag > up > left and the
// band-edge overrides), the walk needs neither scores nor sequence.
//
// Two walkers:
//   traceback_walk     one cell per iteration, any direction layout (generic kernel).
//   k1_traceback       K1 layout: a direction word holds 16 consecutive rows of one band
//                      column, so a run of DIAG moves (the common case: ~98% of the path at
//                      2% divergence) is consumed 16 cells per iteration with bit tricks and
//                      its ops are emitted as one bulk insert.
#pragma once
#include "bsw_common.h"

namespace gamx {

GAMX_HD uint32_t brev32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __brev(v);
#else
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
  return (v >> 16) | (v << 16);
#endif
}
GAMX_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return v ? __builtin_ctz(v) : -1;
#endif
}
// reverses the order of the 16 bit-pairs of v
GAMX_HD uint32_t pairrev16(uint32_t v) {
  const uint32_t b = brev32(v);
  return ((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1);
}

// Packs ops back-to-front into a region of `cap` ops (cap % 16 == 0): the r-th emitted op
// (r = 0 is the LAST op of the edit string) lands at region position cap-1-r, 2 bits each,
// position g at bits [2*(g&15), +2) of word g>>4.
struct OpsWriter {
  uint32_t* words;
  int64_t next;   // next position to fill (decreasing); -1 when the region is exhausted
  uint32_t cur;   // bits already placed in word next>>4
  bool on;
  GAMX_HD void init(uint32_t* w, uint32_t cap, bool enable) { words = w; next = (int64_t)cap - 1; cur = 0; on = enable; }
  // bits: r ops (1 <= r <= 16), pair p = the p-th op emitted (p = 0 first, i.e. latest in the
  // edit string); they occupy positions next, next-1, ..., next-r+1
  GAMX_HD void push_run(uint32_t bits, int r) {
    if (!on) return;
    if (next - r + 1 < 0) { next = -1; on = false; return; }  // cannot happen: cap bounds the path length
    const uint32_t fld = pairrev16(bits) >> (2 * (16 - r));   // pair q = op at position low+q
    const int64_t low = next - r + 1;
    const uint64_t v64 = (uint64_t)fld << (2 * (low & 15));
    const uint32_t lo = (uint32_t)v64, hi = (uint32_t)(v64 >> 32);
    if ((low >> 4) != (next >> 4)) {
      words[next >> 4] = cur | hi;
      cur = lo;
    } else {
      cur |= lo;
    }
    next = low - 1;
    if ((low & 15) == 0) { words[low >> 4] = cur; cur = 0; }
  }
  GAMX_HD void push(uint32_t op) { push_run(op, 1); }
  GAMX_HD void finish() {
    if (!on || next < 0) return;
    if ((next & 15) != 15) words[next >> 4] = cur;
  }
};

struct WalkStats {
  uint32_t n_match = 0, n_mis = 0, n_ga = 0, n_gb = 0, tail_ga = 0, tail_gb = 0;
  int has_match = 0;
  int64_t fm_a = 0, fm_x = 0, lm_a = 0, lm_x = 0;
  GAMX_HD void store(DevResult& R, int64_t pos, int x) const {
    R.n_match = n_match;
    R.n_mismatch = n_mis;
    R.n_gap_a = n_ga;
    R.n_gap_b = n_gb;
    R.n_ops = n_match + n_mis + n_ga + n_gb;
    R.tail_gap_a = tail_ga;
    R.tail_gap_b = tail_gb;
    R.has_match = has_match;
    R.begin_a = pos + 1;          // .cc:321
    R.begin_bx = (int64_t)x + 1;
    R.first_match_a = fm_a; R.first_match_x = fm_x;
    R.last_match_a = lm_a; R.last_match_x = lm_x;
  }
};

// DirAt: int operator()(int x, int y) -> tag (kTagLeft/kTagUp/kTagDiagMis/kTagDiagMatch)
template <class DirAt>
GAMX_HD void traceback_walk(const DirAt& dir_at, int end_i, int end_j, int64_t p0, bool want_ops,
                            uint32_t* ops_words, uint32_t ops_cap, DevResult& R) {
  OpsWriter ow;
  ow.init(ops_words, ops_cap, want_ops);
  WalkStats s;
  int x = end_i, y = end_j;
  int64_t pos = p0 + x + y;
  while (x >= 0 && y >= 0 && pos >= 0) {
    const int tag = dir_at(x, y);
    ow.push((uint32_t)(tag ^ 1));
    if (tag >= kTagDiagMis) {
      if (tag == kTagDiagMatch) {
        if (!s.has_match) { s.has_match = 1; s.lm_a = pos; s.lm_x = x; s.tail_ga = s.n_ga; s.tail_gb = s.n_gb; }
        s.fm_a = pos; s.fm_x = x;
        s.n_match++;
      } else {
        s.n_mis++;
      }
      x--; pos--;
    } else if (tag == kTagUp) {
      s.n_ga++; x--; y++;
    } else {
      s.n_gb++; y--; pos--;
    }
  }
  ow.finish();
  s.store(R, pos, x);
}

// K1 layout: word ((t>>4)*C + k)*LG + l holds the tags of band column j = l*C+k for the 16
// steps t = x + l of one step block, the tag of step offset o = t&15 at bits [2*(15-o), +2).
// C (stripe width) and LG (lanes per pair) are the geometry of the kernel that stored the words.
// The walk is a chain of dependent loads, so it is run one job per THREAD by the traceback kernel
// (thousands of independent walks in flight) rather than by a lane of the warp that filled the band.
//
// Fetch: uint32_t operator()(int blk, int k, int l) -> the word of step block blk of band column
// (lane l, slot k).  DirectFetch loads it; the warp-per-job traceback kernel passes a fetcher whose
// lanes load 32 consecutive step blocks of the column at once (one round trip per 512 rows).
struct DirectFetch {
  const uint32_t* dirs;
  int C, LG;
  GAMX_HD uint32_t operator()(int blk, int k, int l) const {
    return dirs[((uint32_t)blk * (uint32_t)C + (uint32_t)k) * (uint32_t)LG + (uint32_t)l];
  }
};

template <class Fetch>
GAMX_HD void k1_traceback_t(Fetch& fetch, int C, int end_i, int end_j, int p0, bool want_ops,
                            uint32_t* ops_words, uint32_t ops_cap, DevResult& R) {
  OpsWriter ow;
  ow.init(ops_words, ops_cap, want_ops);
  WalkStats s;
  int x = end_i, y = end_j;
  int pos = p0 + x + y;
  int l = y / C, k = y - l * C;
  while (x >= 0 && y >= 0 && pos >= 0) {
    const int t = x + l, o = t & 15;
    const uint32_t w = fetch(t >> 4, k, l);
    const uint32_t ws = w >> (2 * (15 - o));  // pair p = tag of row x-p (p <= o)
    const uint32_t tag = ws & 3u;
    if (tag >= (uint32_t)kTagDiagMis) {
      const uint32_t d = (ws >> 1) & 0x55555555u;       // DIAG flags
      const uint32_t nd = ~d & 0x55555555u;
      int r = nd ? (ctz32(nd) >> 1) : 16;                // length of the DIAG run in this word
      r = imin(r, imin(o + 1, imin(x + 1, pos + 1)));
      const uint32_t maskr = r >= 16 ? 0xffffffffu : ((1u << (2 * r)) - 1u);
      const uint32_t mm = ws & d & maskr;                // MATCH flags (tag == 3)
      if (mm) {
        if (!s.has_match) {
          const int pf = ctz32(mm) >> 1;
          s.has_match = 1; s.lm_a = pos - pf; s.lm_x = x - pf; s.tail_ga = s.n_ga; s.tail_gb = s.n_gb;
        }
        const int pl = (31 - clz32(mm)) >> 1;
        s.fm_a = pos - pl; s.fm_x = x - pl;
      }
      const int m = popc32(mm);
      s.n_match += (uint32_t)m;
      s.n_mis += (uint32_t)(r - m);
      ow.push_run((ws ^ 0x55555555u) & maskr, r);        // op = tag ^ 1
      x -= r; pos -= r;
    } else if (tag == (uint32_t)kTagUp) {
      ow.push(0u);  // GAP_A
      s.n_ga++; x--; y++;
      if (++k == C) { k = 0; l++; }
    } else {
      ow.push(1u);  // GAP_B
      s.n_gb++; y--; pos--;
      if (--k < 0) { k = C - 1; l--; }
    }
  }
  ow.finish();
  s.store(R, (int64_t)pos, x);
}

GAMX_HD void k1_traceback(const uint32_t* dirs, int C, int LG, int end_i, int end_j, int p0, bool want_ops,
                          uint32_t* ops_words, uint32_t ops_cap, DevResult& R) {
  DirectFetch f{dirs, C, LG};
  k1_traceback_t(f, C, end_i, end_j, p0, want_ops, ops_words, ops_cap, R);
}

}  // namespace gamx
