// Device traceback: walks the stored 2-bit directions from the selected end cell and
// produces everything the reference derives from its edit string.
//
// Restates the loop of banded_smith_waterman.cc:227-311 (moves) and the reductions of
// my_alignment.cc:167-296 (first/last match, gap counts).  Because the direction of every
// cell was decided at fill time with the reference's tie-breaking (diag > up > left and the
// band-edge overrides), the walk needs neither scores nor sequence.
#pragma once
#include "bsw_common.h"

namespace gamx {

// Packs ops back-to-front into a region of `cap` ops (cap % 16 == 0): the r-th emitted op
// (r = 0 is the LAST op of the edit string) lands at region position cap-1-r, 2 bits each,
// position g at bits [2*(g&15), +2) of word g>>4.
struct OpsWriter {
  uint32_t* words;
  uint32_t cap;
  uint32_t r;
  uint32_t cur;
  bool on;
  GAMX_HD void init(uint32_t* w, uint32_t c, bool enable) { words = w; cap = c; r = 0; cur = 0; on = enable; }
  GAMX_HD void push(uint32_t op) {
    if (!on) return;
    if (r >= cap) { r++; return; }  // cannot happen: cap >= x_size + |a| window bound
    const uint32_t g = cap - 1 - r;
    cur |= op << (2 * (g & 15));
    if ((g & 15) == 0) { words[g >> 4] = cur; cur = 0; }
    r++;
  }
  // push the same op n times
  GAMX_HD void finish() {
    if (!on || r == 0 || r > cap) return;
    const uint32_t g = cap - r;  // position of the first op
    if ((g & 15) != 0) words[g >> 4] = cur;
  }
};

// DirAt: int operator()(int x, int y) -> tag (kTagLeft/kTagUp/kTagDiagMis/kTagDiagMatch)
template <class DirAt>
GAMX_HD void traceback_walk(const DirAt& dir_at, int end_i, int end_j, int64_t p0, bool want_ops,
                            uint32_t* ops_words, uint32_t ops_cap, DevResult& R) {
  OpsWriter ow;
  ow.init(ops_words, ops_cap, want_ops);
  int x = end_i, y = end_j;
  int64_t pos = p0 + x + y;
  uint32_t n_match = 0, n_mis = 0, n_ga = 0, n_gb = 0;
  uint32_t tail_ga = 0, tail_gb = 0;
  int has_match = 0;
  int64_t fm_a = 0, fm_x = 0, lm_a = 0, lm_x = 0;
  while (x >= 0 && y >= 0 && pos >= 0) {
    const int tag = dir_at(x, y);
    ow.push((uint32_t)(tag ^ 1));
    if (tag >= kTagDiagMis) {
      if (tag == kTagDiagMatch) {
        if (!has_match) { has_match = 1; lm_a = pos; lm_x = x; tail_ga = n_ga; tail_gb = n_gb; }
        fm_a = pos; fm_x = x;
        n_match++;
      } else {
        n_mis++;
      }
      x--; pos--;
    } else if (tag == kTagUp) {
      n_ga++; x--; y++;
    } else {
      n_gb++; y--; pos--;
    }
  }
  ow.finish();
  R.n_match = n_match;
  R.n_mismatch = n_mis;
  R.n_gap_a = n_ga;
  R.n_gap_b = n_gb;
  R.n_ops = n_match + n_mis + n_ga + n_gb;
  R.tail_gap_a = tail_ga;
  R.tail_gap_b = tail_gb;
  R.has_match = has_match;
  R.begin_a = pos + 1;
  R.begin_bx = (int64_t)x + 1;
  R.first_match_a = fm_a; R.first_match_x = fm_x;
  R.last_match_a = lm_a; R.last_match_x = lm_x;
}

}  // namespace gamx
