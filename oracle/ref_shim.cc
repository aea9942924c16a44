// TEST INFRASTRUCTURE ONLY - never linked into the product library.
//
// C-ABI wrapper around the UNMODIFIED reference aligner, compiled from the
// sources where they lie under /root/reference (see oracle/Makefile).  It is
// the ground truth the C restatement (oracle/bsw_oracle.c) and the CUDA path
// are pinned against, and the "reference" CPU arm of bench.py.
//
// Wrapped reference entry points:
//   BandedSmithWaterman::find_alignment   lib/src/alignment/banded_smith_waterman.cc:69-323
//   first_match_pos / last_match_pos      lib/src/alignment/my_alignment.cc:167-193, 228-262
//   last_pos / gaps_before_last_match     lib/src/alignment/my_alignment.cc:196-226, 265-296
//   ABlast::findHits                      lib/src/alignment/ablast.cc:41-76
//   reverse_complement / chop_begin       lib/include/assembly/contig.code.hpp:225-229, 253-257
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <list>
#include <stdexcept>
#include <thread>
#include <vector>

#include "alignment/ablast.hpp"
#include "alignment/banded_smith_waterman.hpp"
#include "alignment/my_alignment.hpp"
#include "assembly/contig.hpp"

namespace {

Contig make_contig(const uint8_t* codes, uint64_t len) {
  Contig c("c", size_t(len));
  for (uint64_t i = 0; i < len; i++) {
    uint8_t v = codes[i] > 4 ? 4 : codes[i];
    c.at(i) = Nucleotide(BaseType(v));
  }
  return c;
}

}  // namespace

extern "C" {

// Mirrors the fields of MyAlignment (lib/include/alignment/my_alignment.hpp:65-126)
// plus the four helper reductions gam-merge reads from it.
struct gamref_result {
  int32_t status;  // 0 = alignment returned, 2 = std::out_of_range thrown
  int32_t has_first_match, has_last_match, has_last_pos;
  int64_t score;
  uint64_t begin_a, begin_b, a_size, b_size;
  uint64_t n_ops;
  double homology;
  uint64_t first_match_a, first_match_b;
  uint64_t last_match_a, last_match_b;
  uint64_t last_pos_a, last_pos_b;
  uint64_t gaps_a, gaps_b;
  int32_t has_gaps;
  int32_t pad_;
};

// Opaque contig handles so repeated alignments do not rebuild the vectors.
void* gamref_contig_new(const uint8_t* codes, uint64_t len) {
  return new Contig(make_contig(codes, len));
}
void gamref_contig_free(void* c) { delete static_cast<Contig*>(c); }
uint64_t gamref_contig_size(void* c) { return static_cast<Contig*>(c)->size(); }
void gamref_contig_codes(void* c, uint8_t* out) {
  Contig* p = static_cast<Contig*>(c);
  for (size_t i = 0; i < p->size(); i++) out[i] = uint8_t(p->at(i).base());
}
void* gamref_contig_revcomp(void* c) {
  Contig* n = new Contig(*static_cast<Contig*>(c));
  reverse_complement(*n);
  return n;
}
void* gamref_contig_chop_begin(void* c, uint64_t from) {
  try {
    return new Contig(chop_begin(*static_cast<Contig*>(c), size_t(from)));
  } catch (...) {
    return nullptr;
  }
}

static void fill_result(const MyAlignment& al, gamref_result* r, uint8_t* ops, uint64_t ops_cap) {
  r->status = 0;
  r->score = al.score();
  r->begin_a = al.begin_a();
  r->begin_b = al.begin_b();
  r->a_size = al.a_size();
  r->b_size = al.b_size();
  r->n_ops = al.length();
  r->homology = al.homology();
  std::pair<MyAlignment::size_type, MyAlignment::size_type> p;
  r->has_first_match = first_match_pos(al, p) ? 1 : 0;
  r->first_match_a = p.first;
  r->first_match_b = p.second;
  r->has_last_match = last_match_pos(al, p) ? 1 : 0;
  r->last_match_a = p.first;
  r->last_match_b = p.second;
  r->has_last_pos = last_pos(al, p) ? 1 : 0;
  r->last_pos_a = p.first;
  r->last_pos_b = p.second;
  r->has_gaps = gaps_before_last_match(al, p) ? 1 : 0;
  r->gaps_a = p.first;
  r->gaps_b = p.second;
  if (ops) {
    const MyAlignment::SeqType& s = al.sequence();
    uint64_t n = s.size() < ops_cap ? s.size() : ops_cap;
    for (uint64_t i = 0; i < n; i++) ops[i] = uint8_t(s[i]);
  }
}

// gap == INT64_MIN selects the 1-argument ctor (band only); otherwise the
// 5-argument ctor is used (only gap and band take effect in the reference).
int gamref_align(void* a, uint64_t begin_a, uint64_t end_a, void* b, uint64_t begin_b,
                 uint64_t end_b, uint64_t band, int64_t gap, int force_start, int force_end,
                 gamref_result* r, uint8_t* ops, uint64_t ops_cap) {
  memset(r, 0, sizeof(*r));
  try {
    if (gap == INT64_MIN) {
      BandedSmithWaterman sw((BandedSmithWaterman::size_type)band);
      MyAlignment al = sw.find_alignment(*static_cast<Contig*>(a), begin_a, end_a,
                                         *static_cast<Contig*>(b), begin_b, end_b,
                                         force_start != 0, force_end != 0);
      fill_result(al, r, ops, ops_cap);
    } else {
      BandedSmithWaterman sw(MATCH_SCORE, MISMATCH_SCORE, gap, GAP_EXT_SCORE,
                             (BandedSmithWaterman::size_type)band);
      MyAlignment al = sw.find_alignment(*static_cast<Contig*>(a), begin_a, end_a,
                                         *static_cast<Contig*>(b), begin_b, end_b,
                                         force_start != 0, force_end != 0);
      fill_result(al, r, ops, ops_cap);
    }
  } catch (const std::out_of_range&) {
    r->status = 2;
  }
  return r->status;
}

// One-shot convenience for byte-code inputs.
int gamref_align_codes(const uint8_t* a, uint64_t la, uint64_t begin_a, uint64_t end_a,
                       const uint8_t* b, uint64_t lb, uint64_t begin_b, uint64_t end_b,
                       uint64_t band, int64_t gap, int force_start, int force_end,
                       gamref_result* r, uint8_t* ops, uint64_t ops_cap) {
  Contig ca = make_contig(a, la), cb = make_contig(b, lb);
  return gamref_align(&ca, begin_a, end_a, &cb, begin_b, end_b, band, gap, force_start,
                      force_end, r, ops, ops_cap);
}

// ABlast::findHits (lib/src/alignment/ablast.cc:41-76); returns number of hits,
// writes up to cap of them (ascending, as the std::list order).
uint64_t gamref_find_hits(void* a, uint64_t a_start, uint64_t a_end, void* b, uint64_t b_start,
                          uint64_t b_end, uint32_t* hits, uint64_t cap) {
  ABlast ab;  // word size 20, lib/include/alignment/ablast.hpp:33
  std::list<uint32_t> h = ab.findHits(*static_cast<Contig*>(a), a_start, a_end,
                                      *static_cast<Contig*>(b), b_start, b_end);
  uint64_t n = 0;
  for (std::list<uint32_t>::const_iterator it = h.begin(); it != h.end(); ++it, ++n)
    if (n < cap) hits[n] = *it;
  return n;
}

// Multithreaded CPU baseline: the same parallel shape as
// lib/src/pctg/ThreadedBuildPctg.cc:159-169 without the graph layer: n_threads
// workers, each constructing its own BandedSmithWaterman per call (as
// PctgBuilder.cc:1628 does) and pulling job indices from a shared counter.
// Jobs are full-window alignments (a, 0, la-1, b, 0, lb-1).  Returns seconds.
// cells_out = sum over jobs of x_size * (2*band+1)  (banded_smith_waterman.cc:93-97).
double gamref_bench(const uint8_t* const* a_seqs, const uint64_t* a_lens,
                    const uint8_t* const* b_seqs, const uint64_t* b_lens, uint64_t n_jobs,
                    uint64_t band, int n_threads, uint64_t* cells_out, int64_t* score_sum_out) {
  std::vector<Contig> A, B;
  A.reserve(n_jobs);
  B.reserve(n_jobs);
  for (uint64_t i = 0; i < n_jobs; i++) {
    A.push_back(make_contig(a_seqs[i], a_lens[i]));
    B.push_back(make_contig(b_seqs[i], b_lens[i]));
  }
  std::atomic<uint64_t> next(0);
  std::atomic<uint64_t> cells(0);
  std::atomic<int64_t> ssum(0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; t++) {
    th.emplace_back([&]() {
      uint64_t my_cells = 0;
      int64_t my_sum = 0;
      for (;;) {
        uint64_t i = next.fetch_add(1);
        if (i >= n_jobs) break;
        try {
          BandedSmithWaterman sw((BandedSmithWaterman::size_type)band);
          MyAlignment al =
              sw.find_alignment(A[i], 0, A[i].size() - 1, B[i], 0, B[i].size() - 1, false, false);
          my_sum += al.score();
        } catch (const std::out_of_range&) {
        }
        uint64_t x = B[i].size();
        uint64_t lim = A[i].size() + band;
        if (lim < x) x = lim;
        if (x > BSW_MAX_ALIGNMENT) x = BSW_MAX_ALIGNMENT;
        my_cells += x * (2 * band + 1);
      }
      cells += my_cells;
      ssum += my_sum;
    });
  }
  for (auto& t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (cells_out) *cells_out = cells.load();
  if (score_sum_out) *score_sum_out = ssum.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
