// Host-side job preparation shared by the C-ABI implementation (gamx.cu) and the CPU
// lane simulator (tests/sim/warp_sim.cc): 2-bit packing, the guards and sizes of
// banded_smith_waterman.cc:90-97, classification fast/generic, and the conversion of a
// device result into the fields of MyAlignment + its reductions (my_alignment.cc:167-296).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/gamx.h"
#include "bsw_common.h"
#include "bsw_warp.h"

namespace gamx {

// ---- packed contig store (host mirror) -----------------------------------------------------
// Every contig starts at a multiple of 32 bases so that both arrays are word-aligned per contig.
struct HostStore {
  std::vector<uint32_t> packed;  // 2 bits per base
  std::vector<uint32_t> nmask;   // 1 bit per base
  std::vector<uint64_t> start;   // first base index of contig id
  std::vector<uint64_t> length;
  uint64_t n_bases = 0;          // padded total

  int64_t add(const uint8_t* codes, uint64_t len) {
    const uint64_t s = n_bases;
    const uint64_t padded = (len + 31) & ~uint64_t(31);
    packed.resize((s + padded) / 16 + 2, 0u);
    nmask.resize((s + padded) / 32 + 2, 0u);
    for (uint64_t i = 0; i < len; i++) {
      const uint64_t idx = s + i;
      const uint32_t c = codes[i];
      if (c >= 4) nmask[idx >> 5] |= 1u << (idx & 31);
      else packed[idx >> 4] |= c << (2 * (idx & 15));
    }
    start.push_back(s);
    length.push_back(len);
    n_bases = s + padded;
    return (int64_t)start.size() - 1;
  }
  void clear() { packed.clear(); nmask.clear(); start.clear(); length.clear(); n_bases = 0; }
};

inline uint8_t ascii_to_code(char ch) {  // nucleotide.code.hpp:47-75
  switch (ch) {
    case 'A': case 'a': return 0;
    case 'T': case 't': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    default: return 4;
  }
}

// view = (rc ? reverse_complement(contig) : contig)[off, off+len)
inline SeqView make_view(uint64_t cstart, uint64_t clen, bool rc, uint64_t off) {
  SeqView v;
  if (!rc) { v.origin = (int64_t)(cstart + off); v.dir = 1; v.comp = 0; }
  else { v.origin = (int64_t)(cstart + clen - 1 - off); v.dir = -1; v.comp = 1; }
  return v;
}

// ---- classification -------------------------------------------------------------------------
enum { kClassEarly = 0, kClassWarp = 1, kClassGeneric = 2, kClassCta = 3 };

struct Prepared {
  int cls;             // kClass*
  int early_status;    // for kClassEarly
  int c;               // lane stripe width for kClassWarp
  int lg;              // lanes per pair for kClassWarp (32, 16 or 8)
  uint64_t x_size;
  uint64_t cells;      // x_size * (2*band+1), the unit of the GCUPS metric
  uint64_t la, lb;
  uint64_t begin_b;
  uint64_t dir_words;  // direction scratch this job needs (fast: K1 layout, generic: row-major)
  uint64_t gen_rows;   // generic: int64 row scratch (2*y_size)
  uint32_t ops_cap;    // ops (multiple of 16), FULL mode only
  uint32_t gen_idx;    // kClassGeneric: index of the job's GenJob in the caller's side array
  DevJob dj;           // kClassWarp
};

// Lanes per pair (LG) and stripe width (C) for a band: the combination with the best lane
// utilisation (2*band+1)/(LG*C), discounted by 6 % for stripes wider than 12 slots (their selector
// loads conflict in shared-memory banks; measured on B200 for band 64: LG=16,C=9 beats LG=8,C=17,
// profiles/r1c_geometry_probe.txt); ties go to fewer lanes per pair.  LG < 32 only from C >= 5.
// GAMX_FORCE_LG / GAMX_FORCE_C (environment) override the choice for experiments.
inline void geometry_for_band(uint64_t band, bool dirs, int* c_out, int* lg_out) {
  (void)dirs;
  static const int forced = [] { const char* e = getenv("GAMX_FORCE_LG"); return e ? atoi(e) : 0; }();
  static const int forced_c = [] { const char* e = getenv("GAMX_FORCE_C"); return e ? atoi(e) : 0; }();
  const uint64_t y = 2 * band + 1;
  int best_c = 0, best_lg = 0;
  double best = -1.0;
  const int lgs[4] = {4, 8, 16, 32};
  for (int n = 0; n < 4; n++) {
    const int lg = lgs[n];
    int c = (int)((y + lg - 1) / lg);
    if (c < 2) c = 2;
    // stripes of 13 or 17 slots read their selector windows with a lane stride of 6 or 8 words: 2- and
    // 8-way shared-memory bank conflicts that make the LSU the bottleneck (band 256, score only: 2963
    // GCUPS with C = 17, 4735 with C = 18) - take one slot more
    static const bool no_even = getenv("GAMX_NO_EVEN_C") != nullptr;  // experiments only
    if (!no_even && (c == 13 || c == 17) && c + 1 <= kMaxC) c++;
    // (9-slot stripes of several groups per warp keep 4-way conflicts of the selector loads - a lane stride of 4
    //  words leaves 8 banks whatever the distance between the groups' arrays, bsw_warp16.h group_pad_bytes - but
    //  10 slots measured no better overall: uniform 1 kb pairs at band 16 +1..2 %, mixed lengths -9 %)
    if (forced_c > c && forced_c <= kMaxC) c = forced_c;
    if (c > kMaxC) continue;
    if (forced && lg != forced && (y + forced - 1) / forced <= (uint64_t)kMaxC) continue;
    if (!forced && lg < 32 && c < 5 && y > 64) continue;
    const double score = (double)y / (double)(lg * c);
    if (score > best + 1e-9) { best = score; best_c = c; best_lg = lg; }
  }
  *c_out = best_c; *lg_out = best_lg;
}

// K2 geometry (one pair per CTA): the fewest threads whose stripes cover the band, C <= kMaxC.
inline bool geometry_cta(uint64_t band, int* c_out, int* lg_out) {
  const uint64_t y = 2 * band + 1;
  const int lgs[3] = {64, 128, 256};
  for (int n = 0; n < 3; n++) {
    const int lg = lgs[n];
    int c = (int)((y + lg - 1) / lg);
    if (c < 2) c = 2;
    if ((c == 13 || c == 17) && c + 1 <= kMaxC) c++;  // (bank conflicts of the selector loads, see geometry_for_band)
    if (c > kMaxC) continue;
    *c_out = c; *lg_out = lg;
    return true;
  }
  return false;
}
constexpr uint64_t kMaxBandCta = (256ull * kMaxC - 1) / 2;  // 2303

// Latency mode: a band that fits one warp, spread over a 64-lane CTA instead (half the cells per
// lane and step) for batches too small to fill the device with one warp per pair.
inline void cta_geometry_for_latency(uint64_t band, int* c_out, int* lg_out) {
  const uint64_t y = 2 * band + 1;
  int c = (int)((y + 63) / 64);
  if (c < 2) c = 2;
  while (c <= kMaxC && !stripe_supported(c)) c++;
  *c_out = c; *lg_out = 64;
}

// a_view / b_view: views of position 0 of a and b; la / lb: view lengths (a.size(), b.size()).
// Fills P; for jobs classified kClassGeneric the raw arguments are written to *gj (when not null).
inline void prepare_job(Prepared& P, GenJob* gj, const SeqView& a_view, uint64_t la, const SeqView& b_view,
                        uint64_t lb, uint64_t begin_a, uint64_t end_a, uint64_t begin_b, uint64_t end_b,
                        uint64_t band, int64_t gap, bool fs, bool fe, int mode) {
  memset(&P, 0, sizeof(P));
  P.la = la; P.lb = lb; P.begin_b = begin_b;
  // banded_smith_waterman.cc:90-95 with the reference's unsigned arithmetic
  if (end_b < begin_b) { P.cls = kClassEarly; P.early_status = kStatusEmpty; return; }
  uint64_t eb = end_b;
  if (eb >= lb) eb = lb - 1;
  uint64_t x = eb - begin_b + 1;
  const uint64_t lim = la + band - begin_a;
  if (lim < x) x = lim;
  if (x > (uint64_t)kMaxAlignment) x = kMaxAlignment;
  P.x_size = x;
  const uint64_t y = 2 * band + 1;
  P.cells = x * y;
  if (x == 0) { P.cls = kClassEarly; P.early_status = kStatusUndefined; return; }

  // upper bound on the edit-string length: (#DIAG+#UP) <= x, (#DIAG+#LEFT) <= |a|,
  // #LEFT <= #UP + 2*band + 1
  uint64_t ops = x + (la < x + y ? la : x + y);
  P.ops_cap = (mode == kModeFull) ? (uint32_t)((ops + 15) & ~uint64_t(15)) : 0u;

  const bool regular = la >= 1 && lb >= 1 && begin_b <= eb && eb < lb && begin_a < la + band &&
                       gap >= -29 && gap <= -5 && band <= kMaxBandCta && !(fs && la <= (uint64_t)kForceMaxGap) &&
                       la < (1ull << 30) && lb < (1ull << 30);
  if (!regular) {
    P.cls = kClassGeneric;
    if (gj) {
      GenJob& g = *gj;
      memset(&g, 0, sizeof(g));
      g.a = a_view; g.b = b_view; g.la = la; g.lb = lb;
      g.begin_a = begin_a; g.end_a = end_a; g.begin_b = begin_b; g.end_b = end_b;
      g.band = band; g.gap = gap; g.force_start = fs; g.force_end = fe; g.mode = mode;
      g.ops_cap = P.ops_cap; g.x_size = x;
    }
    P.dir_words = (x * y + 15) / 16;
    P.gen_rows = 2 * y;
    return;
  }

  if (y <= 32ull * kMaxC) {
    P.cls = kClassWarp;
    geometry_for_band(band, mode != kModeScore, &P.c, &P.lg);
  } else {
    P.cls = kClassCta;  // band too wide for one warp: CTA-per-pair kernel
    geometry_cta(band, &P.c, &P.lg);
  }
  DevJob& d = P.dj;
  d.a = a_view;
  d.b = b_view;
  d.b.origin = b_view.origin + (int64_t)b_view.dir * (int64_t)begin_b;  // DP row 0
  d.la = (int32_t)la;
  const int64_t p0 = (int64_t)begin_a - (int64_t)band;
  d.p0 = (int32_t)p0;
  d.x = (int32_t)x;
  d.band = (int32_t)band;
  const int64_t last_row_pos0 = p0 + (int64_t)x - 1;  // pos of (x-1, 0)
  // last column, .cc:197-212: cells with i + j == end_a - begin_a + band
  d.kc = -1;
  if (end_a + band >= begin_a && end_a < (1ull << 40)) {
    const uint64_t k = end_a + band - begin_a;
    if (k <= (x - 1) + 2 * band) d.kc = (int32_t)k;
  }
  // last row, .cc:179-192
  if (fe) { d.jlo = 1; d.jhi = 0; }
  else {
    int64_t jlo = last_row_pos0 < 0 ? -last_row_pos0 : 0;
    int64_t jhi = (int64_t)y - 1;
    if (end_a < (1ull << 40)) {
      const int64_t lim_j = (int64_t)end_a - last_row_pos0;
      if (lim_j < jhi) jhi = lim_j;
    }
    if (jhi < 0) { jlo = 1; jhi = 0; }
    d.jlo = (int32_t)jlo; d.jhi = (int32_t)jhi;
  }
  int64_t jfill = (int64_t)la - last_row_pos0;
  if (jfill < 0) jfill = 0;
  if (jfill > (int64_t)y) jfill = (int64_t)y;
  d.jfill = (int32_t)jfill;
  d.col_imin = fe ? (x >= 11 ? (int32_t)(x - 1 - kForceMaxGap) : 0x7fffffff) : 0;  // .cc:201
  d.col_zero = end_a >= la ? 1 : 0;
  d.gap = (int32_t)gap;
  static const bool skip_tb = getenv("GAMX_DEBUG_SKIP_TRACEBACK") != nullptr;  // profiling experiments only
  d.mode = mode | ((skip_tb && mode != kModeScore) ? 0x100 : 0);
  d.ops_cap = P.ops_cap;
  P.dir_words = (mode == kModeScore) ? 0 : k1_dir_words((int)x, (int)band, P.c, P.lg);
}

// ---- device result -> gamx_result -----------------------------------------------------------
inline void finalize_result(const Prepared& P, const DevResult* dr, int mode, gamx_result* out) {
  memset(out, 0, sizeof(*out));
  out->x_size = P.x_size;
  if (P.cls == kClassEarly) { out->status = P.early_status; return; }
  out->status = dr->status;
  if (dr->status != kStatusOk) return;
  out->score = dr->score;
  out->end_i = dr->end_i;
  out->end_j = dr->end_j;
  out->a_size = P.la;
  out->b_size = P.lb;
  if (mode == kModeScore) return;
  const uint64_t ba = (uint64_t)dr->begin_a, bb = P.begin_b + (uint64_t)dr->begin_bx;
  out->begin_a = ba;
  out->begin_b = bb;
  out->n_ops = dr->n_ops;
  out->n_match = dr->n_match;
  out->n_mismatch = dr->n_mismatch;
  out->n_gap_a = dr->n_gap_a;
  out->n_gap_b = dr->n_gap_b;
  out->has_match = dr->has_match;
  // banded_smith_waterman.cc:319
  out->homology = dr->n_ops == 0 ? 0.0 : (double)((uint64_t)dr->n_match * 100) / (double)dr->n_ops;
  if (dr->has_match) {
    out->first_match_a = (uint64_t)dr->first_match_a;                 // my_alignment.cc:167-193
    out->first_match_b = P.begin_b + (uint64_t)dr->first_match_x;
    out->last_match_a = (uint64_t)dr->last_match_a;                   // my_alignment.cc:228-262
    out->last_match_b = P.begin_b + (uint64_t)dr->last_match_x;
    out->gaps_a = dr->n_gap_a - dr->tail_gap_a;                       // my_alignment.cc:265-296
    out->gaps_b = dr->n_gap_b - dr->tail_gap_b;
  } else {
    out->first_match_a = ba + dr->n_gap_b + dr->n_mismatch;
    out->first_match_b = bb + dr->n_gap_a + dr->n_mismatch;
    out->last_match_a = ba;
    out->last_match_b = bb;
  }
  out->last_pos_a = ba + dr->n_match + dr->n_mismatch + dr->n_gap_b;  // my_alignment.cc:196-226
  out->last_pos_b = bb + dr->n_match + dr->n_mismatch + dr->n_gap_a;
  out->ops_offset = dr->ops_start;
}

}  // namespace gamx
