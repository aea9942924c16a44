// Generic kernel body: one thread aligns one job, following the reference statement by
// statement (banded_smith_waterman.cc:69-323) for ANY argument combination - arbitrary gap
// score, force flags, windows that run off either contig, the out_of_range cases.  It is the
// GPU path for the rare jobs the fast warp kernel (bsw_warp.h) does not accept; it is not a
// CPU fallback (it runs on the device) and it is not fast.
//
// Differences from the reference are storage only: two rolling int64 rows instead of the
// full matrix, and the traceback direction of every filled cell decided at fill time
//  and stored 2 bits per cell, row-major.
#pragma once
#include "bsw_common.h"
#include "bsw_traceback.h"

namespace gamx {

struct GenDirAt {
  const uint32_t* dirs;
  uint64_t y_size;
  GAMX_HD int operator()(int x, int y) const {
    const uint64_t c = (uint64_t)x * y_size + (uint64_t)y;
    return (int)((dirs[c >> 4] >> (2 * (c & 15))) & 3u);
  }
};

GAMX_HD void gen_set_dir(uint32_t* dirs, uint64_t c, uint32_t tag) {
  const uint32_t sh = 2 * (uint32_t)(c & 15);
  dirs[c >> 4] = (dirs[c >> 4] & ~(3u << sh)) | (tag << sh);
}

GAMX_HD int64_t max64(int64_t a, int64_t b) { return a > b ? a : b; }

// rows: 2*y_size int64 scratch; dirs: ceil(x_size*y_size/16) words scratch; ops: device ops buffer
GAMX_HD void generic_align(const GenJob& J, const SeqStore& S, int64_t* rows, uint32_t* dirs,
                           uint32_t* ops_buf, DevResult& R) {
  const int64_t FM = kForceMaxGap;
  const bool fs = J.force_start != 0, fe = J.force_end != 0;
  const uint64_t la = J.la, lb = J.lb, band = J.band, begin_a = J.begin_a, begin_b = J.begin_b;
  const uint64_t end_a = J.end_a;
  const uint64_t x_size = J.x_size, y_size = 2 * band + 1;
  const int64_t gap = J.gap;
  R.status = kStatusOk;
  R.score = 0; R.end_i = 0; R.end_j = 0; R.has_match = 0;
  R.n_ops = R.n_match = R.n_mismatch = R.n_gap_a = R.n_gap_b = 0;
  R.tail_gap_a = R.tail_gap_b = 0;
  R.begin_a = R.begin_bx = 0;
  R.first_match_a = R.first_match_x = R.last_match_a = R.last_match_x = 0;
  R.ops_start = 0;

  int64_t* prev = rows;
  int64_t* cur = rows + y_size;
  // best "last column" candidate (cells with pos == end_a, .cc:197-212), earliest row wins ties
  bool col_found = false;
  int64_t col_best = 0, col_i = 0, col_j = 0;
  const int64_t kc = (int64_t)(end_a - begin_a + band);  // i + j of those cells (may wrap: checked below)
  // (end_a < 2^40 first, like the fast path's guard in bsw_host.h: with end_a near 2^64 - e.g. UINT64_MAX from
  //  m_at + mlen - 1 with mlen = 0 - the sums below wrap into the valid range, but the reference has no
  //  last-column cell then: int_type(end_a) is negative, .cc:197)
  const bool kc_valid = end_a < ((uint64_t)1 << 40) && end_a + band >= begin_a && (end_a - begin_a + band) < (uint64_t)1 << 40;

  for (uint64_t i = 0; i < x_size; i++) {
    // columns whose pos = begin_a + i + j - band lies in [0, la)
    const int64_t base = (int64_t)(begin_a + i - band);  // pos of column 0
    for (uint64_t j = 0; j < y_size; j++) cur[j] = 0;
    int64_t jb = base < 0 ? -base : 0;
    int64_t pos_excl = (int64_t)la;
    // with force_start the first row also touches pos <= FORCE_MAXGAP_LEN even beyond |a| (.cc:116)
    if (i == 0 && fs && pos_excl < FM + 1) pos_excl = FM + 1;
    int64_t je = pos_excl - base;  // exclusive
    if (je > (int64_t)y_size) je = (int64_t)y_size;
    for (int64_t j = jb; j < je; j++) {
      const int64_t pos = base + j;
      const uint64_t c = i * y_size + (uint64_t)j;
      int64_t h;
      uint32_t tag;
      if (i == 0) {
        // first row, .cc:112-132
        const bool in1 = (!fs && pos >= 0 && (uint64_t)pos < la) || (fs && pos >= 0 && pos <= FM);
        const bool in2 = fs && pos > FM && (uint64_t)pos < la;
        if (!in1 && !in2) continue;
        if ((uint64_t)pos >= la || begin_b >= lb) { R.status = kStatusOutOfRange; return; }
        const uint32_t ca = load_code(S, J.a, pos), cb = load_code(S, J.b, (int64_t)begin_b);
        const int64_t s = subst_score(ca, cb);
        const bool cnd = pos > 0 && j > 0;
        if (in1) {
          const int64_t left = cnd ? cur[j - 1] : gap;
          h = cnd ? max64(max64(s, gap), left) : max64(gap, s);
        } else {
          const int64_t left = cnd ? cur[j - 1] : gap;
          h = cnd ? max64(s, left) : s;
        }
        // direction as the traceback would derive it (.cc:229-308 with x == 0)
        const uint32_t dtag = is_match_op(ca, cb) ? kTagDiagMatch : kTagDiagMis;
        if (pos == 0) {
          if (h == s) tag = dtag;
          else if (j == (int64_t)y_size - 1 || h == gap) tag = kTagLeft;  // x == 0 <= FM: left allowed
          else tag = kTagUp;
        } else {
          const bool up_ok = !(fs && pos > FM);
          if (h == s) tag = dtag;  // diag = 0 + s
          else if (j < (int64_t)y_size - 1 && j > 0 && up_ok && h == gap) tag = kTagUp;
          else if (j < (int64_t)y_size - 1 && j > 0) tag = kTagLeft;
          else if (j < (int64_t)y_size - 1) tag = kTagUp;
          else tag = kTagLeft;
        }
      } else {
        // fill, .cc:135-171
        if (begin_b + i >= lb) { R.status = kStatusOutOfRange; return; }
        const uint32_t ca = load_code(S, J.a, pos), cb = load_code(S, J.b, (int64_t)(begin_b + i));
        const int64_t s = subst_score(ca, cb);
        const uint32_t dtag = is_match_op(ca, cb) ? kTagDiagMatch : kTagDiagMis;
        const bool not_last = j < (int64_t)y_size - 1;
        const int64_t up = not_last ? prev[j + 1] + gap : gap;
        if (pos == 0) {
          if (!fs || (int64_t)i <= FM) h = not_last ? max64(max64(s, up), gap) : max64(s, gap);
          else h = not_last ? max64(s, up) : s;
          const bool left_ok = !(fs && (int64_t)i > FM);
          if (h == s) tag = dtag;
          else if (!not_last || (left_ok && h == gap)) tag = kTagLeft;
          else tag = kTagUp;
        } else {
          const int64_t diag = prev[j] + s;
          const int64_t left = (j > 0) ? cur[j - 1] + gap : gap;
          if (not_last && j > 0) h = max64(max64(diag, up), left);
          else if (not_last) h = max64(diag, up);
          else if (j > 0) h = max64(diag, left);
          else h = diag;
          if (h == diag) tag = dtag;
          else if (not_last && j > 0 && h == up) tag = kTagUp;
          else if (not_last && j > 0) tag = kTagLeft;
          else if (not_last) tag = kTagUp;
          else tag = kTagLeft;
        }
      }
      cur[j] = h;
      gen_set_dir(dirs, c, tag);
    }
    // "last column" candidate of this row: the cell with i + j == kc (filled or not)
    if (kc_valid && kc - (int64_t)i >= 0 && kc - (int64_t)i <= (int64_t)(2 * band)) {
      const int64_t j = kc - (int64_t)i;
      const bool ok = !fe || (i >= x_size - 1 - (uint64_t)FM);  // unsigned wrap for x_size < 11, as .cc:201
      if (ok) {
        const int64_t v = cur[j];
        if (!col_found || v > col_best) { col_found = true; col_best = v; col_i = (int64_t)i; col_j = j; }
      }
    }
    int64_t* t = prev; prev = cur; cur = t;
  }
  // prev now holds the last row.  End-cell selection, .cc:174-212.
  bool found = false;
  int64_t max_i = 0, max_j = 0, max_score = 0;
  if (!fe) {
    for (uint64_t j = 0; j < y_size; j++) {
      const int64_t pos = (int64_t)(begin_a + (x_size - 1) + j - band);
      if (pos >= 0 && (uint64_t)pos <= end_a) {
        if (!found || prev[j] > max_score) { found = true; max_i = (int64_t)x_size - 1; max_j = (int64_t)j; max_score = prev[j]; }
      }
    }
  }
  if (col_found && (!found || col_best > max_score)) { found = true; max_i = col_i; max_j = col_j; max_score = col_best; }
  if (!found) { R.status = kStatusEmpty; return; }
  R.score = (int32_t)max_score;
  R.end_i = (int32_t)max_i;
  R.end_j = (int32_t)max_j;
  const int64_t p0 = (int64_t)(begin_a - band);
  const int64_t pos_end = p0 + max_i + max_j;
  // first traceback iteration reads a.at(pos), b.at(begin_b+x) (.cc:231,:265)
  if (pos_end >= 0 && ((uint64_t)pos_end >= la || begin_b + (uint64_t)max_i >= lb)) { R.status = kStatusOutOfRange; return; }
  if (J.mode == kModeScore) return;
  GenDirAt da{dirs, y_size};
  traceback_walk(da, (int)max_i, (int)max_j, p0, J.mode == kModeFull, ops_buf + J.ops_word, J.ops_cap, R);
  R.ops_start = J.ops_word * 16 + J.ops_cap - R.n_ops;
}

}  // namespace gamx
