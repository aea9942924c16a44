// C++ drop-in for the reference's aligner interface, implemented over the C ABI (include/gamx.h).
//
// Same class names, constructors, method signatures, defaults and error behaviour as
//   BandedSmithWaterman   /root/reference/lib/include/alignment/banded_smith_waterman.hpp:41-72
//   MyAlignment           /root/reference/lib/include/alignment/my_alignment.hpp:65-126
//   first_match_pos / last_match_pos / last_pos / gaps_before_last_match
//                         /root/reference/lib/src/alignment/my_alignment.cc:167-296
// so that lib/src/pctg/PctgBuilder.cc:1410,1544-1607,1628,1669,1698 compile unchanged against
// it.  Everything lives in namespace gamx_dropin; a reference build would
// `using` these names in place of its own headers.  The sequence type is a template parameter
// of find_alignment: anything with size() and operator[] yielding a value convertible to a
// base code 0..4 (the reference's Contig/Nucleotide qualify through Nucleotide::base()).
//
// A single find_alignment call is a batch of one: correct, thread-safe, not fast.  The batch
// collector (AlignBatch below) is what turns gam-merge's per-thread calls into GPU-sized rounds.
#pragma once
#include <stdint.h>

#include <list>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gamx.h"

namespace gamx_dropin {

typedef int64_t ScoreType;                                   // my_alignment.hpp:55
enum AlignmentAlphabet : uint8_t { GAP_A, GAP_B, MATCH, MISMATCH };  // my_alignment.hpp:57-62

class MyAlignment {
 public:
  typedef int64_t int_type;
  typedef uint64_t size_type;
  typedef std::vector<AlignmentAlphabet> SeqType;

  MyAlignment() : _begin_a(0), _begin_b(0), _a_size(0), _b_size(0), _score(0), _homology(0), _has_red(false) {}
  explicit MyAlignment(double homology) : MyAlignment() { _homology = homology; }

  size_type begin_a() const { return _begin_a; }
  size_type begin_b() const { return _begin_b; }
  size_type a_size() const { return _a_size; }
  size_type b_size() const { return _b_size; }
  const SeqType& sequence() const { return _sequence; }
  size_type length() const { return _has_red ? _n_ops : _sequence.size(); }
  ScoreType score() const { return _score; }
  double homology() const { return _homology; }
  void set_begin_a(size_type v) { _begin_a = v; }
  void set_begin_b(size_type v) { _begin_b = v; }
  void set_homology(double h) { _homology = h; }

  // filled from a gamx_result; ops may be null (ENDPOINTS mode: reductions only)
  static MyAlignment from_result(const gamx_result& r, const uint8_t* ops_buf) {
    if (r.status == GAMX_JOB_EMPTY) return MyAlignment();  // banded_smith_waterman.cc:90, :215
    if (r.status == GAMX_JOB_OUT_OF_RANGE) throw std::out_of_range("vector::_M_range_check (Contig::at)");
    if (r.status != GAMX_JOB_OK) throw std::logic_error("find_alignment: reference behaviour undefined (x_size == 0)");
    MyAlignment al;
    al._begin_a = r.begin_a; al._begin_b = r.begin_b; al._a_size = r.a_size; al._b_size = r.b_size;
    al._score = r.score; al._homology = r.homology;
    al._has_red = true; al._n_ops = r.n_ops; al._has_match = r.has_match != 0;
    al._fm = std::make_pair(r.first_match_a, r.first_match_b);
    al._lm = std::make_pair(r.last_match_a, r.last_match_b);
    al._lp = std::make_pair(r.last_pos_a, r.last_pos_b);
    al._gaps = std::make_pair(r.gaps_a, r.gaps_b);
    if (ops_buf) {
      std::vector<uint8_t> tmp(r.n_ops);
      gamx_unpack_ops(ops_buf, r.ops_offset, r.n_ops, tmp.data());
      al._sequence.resize(r.n_ops);
      for (uint64_t k = 0; k < r.n_ops; k++) al._sequence[k] = AlignmentAlphabet(tmp[k]);
    }
    return al;
  }

 private:
  size_type _begin_a, _begin_b, _a_size, _b_size;
  SeqType _sequence;
  ScoreType _score;
  double _homology;
  // device-side reductions (my_alignment.cc:167-296), so callers never need the edit string
  bool _has_red;
  size_type _n_ops = 0;
  bool _has_match = false;
  std::pair<size_type, size_type> _fm, _lm, _lp, _gaps;
  friend bool first_match_pos(const MyAlignment&, std::pair<size_type, size_type>&);
  friend bool last_match_pos(const MyAlignment&, std::pair<size_type, size_type>&);
  friend bool last_pos(const MyAlignment&, std::pair<size_type, size_type>&);
  friend bool gaps_before_last_match(const MyAlignment&, std::pair<size_type, size_type>&);
};

inline bool first_match_pos(const MyAlignment& A, std::pair<MyAlignment::size_type, MyAlignment::size_type>& pos) {
  if (A._has_red) { pos = A._fm; return A._has_match; }
  pos.first = A.begin_a(); pos.second = A.begin_b(); return false;  // default MyAlignment()
}
inline bool last_match_pos(const MyAlignment& A, std::pair<MyAlignment::size_type, MyAlignment::size_type>& pos) {
  if (A._has_red) { pos = A._lm; return A._has_match; }
  pos.first = A.begin_a(); pos.second = A.begin_b(); return false;
}
inline bool last_pos(const MyAlignment& A, std::pair<MyAlignment::size_type, MyAlignment::size_type>& pos) {
  if (A._has_red) { pos = A._lp; return A._has_match; }
  pos.first = A.begin_a(); pos.second = A.begin_b(); return false;
}
inline bool gaps_before_last_match(const MyAlignment& A, std::pair<MyAlignment::size_type, MyAlignment::size_type>& g) {
  if (A._has_red) { g = A._gaps; return A._has_match; }
  g.first = 0; g.second = 0; return false;
}

// One process-wide context (all visible GPUs), created on first use.
inline gamx_ctx* default_context() {
  static gamx_ctx* ctx = [] {
    gamx_ctx* c = nullptr;
    if (gamx_create(&c, nullptr, 0) != GAMX_OK)
      throw std::runtime_error("gamx: no usable CUDA device (this aligner has no CPU fallback)");
    return c;
  }();
  return ctx;
}

// base code 0..4 of a sequence element; overload it (same namespace as your element type, found
// by ADL) for the reference's Nucleotide:  inline int base_code(const Nucleotide& n) { return n.base(); }
inline int base_code(uint8_t v) { return v > 4 ? 4 : v; }

template <class SeqT>
inline std::vector<uint8_t> codes_of(const SeqT& s) {
  std::vector<uint8_t> c(s.size());
  for (size_t i = 0; i < c.size(); i++) c[i] = (uint8_t)base_code(s[i]);
  return c;
}

// Collects find_alignment calls and runs them as one GPU batch: the "batch collector" that
// replaces gam-merge's per-thread synchronous calls (PctgBuilder.cc:1669 inside
// ThreadedBuildPctg.cc:305-339).  Contigs are uploaded once and referenced by id.
class AlignBatch {
 public:
  explicit AlignBatch(gamx_ctx* ctx = nullptr) : _ctx(ctx ? ctx : default_context()) {}

  template <class SeqT>
  uint32_t add_contig(const SeqT& s) {
    std::vector<uint8_t> c = codes_of(s);
    int64_t id = gamx_add_contig(_ctx, c.data(), c.size());
    if (id < 0) throw std::runtime_error(gamx_last_error(_ctx));
    return (uint32_t)id;
  }
  // view = (rc ? reverse_complement(contig) : contig)[off, off+len)
  size_t add(uint32_t a_id, bool a_rc, uint64_t a_off, uint64_t a_len, uint64_t begin_a, uint64_t end_a,
             uint32_t b_id, bool b_rc, uint64_t b_off, uint64_t b_len, uint64_t begin_b, uint64_t end_b,
             uint32_t band = GAMX_DEFAULT_BAND, int32_t gap = GAMX_DEFAULT_GAP, bool force_start = false,
             bool force_end = false, int mode = GAMX_MODE_ENDPOINTS) {
    gamx_job j = {};
    j.a_id = a_id; j.b_id = b_id; j.a_rc = a_rc; j.b_rc = b_rc; j.a_off = a_off; j.a_len = a_len;
    j.b_off = b_off; j.b_len = b_len; j.begin_a = begin_a; j.end_a = end_a; j.begin_b = begin_b; j.end_b = end_b;
    j.band = band; j.gap = gap; j.force_start = force_start; j.force_end = force_end; j.mode = (uint8_t)mode;
    _jobs.push_back(j);
    return _jobs.size() - 1;
  }
  // runs every collected job; results()[k] belongs to the k-th add()
  void run() {
    _results.assign(_jobs.size(), gamx_result());
    const uint64_t cap = gamx_ops_capacity(_ctx, _jobs.data(), _jobs.size());
    _ops.assign((cap + 3) / 4 + 8, 0);
    if (gamx_align_batch(_ctx, _jobs.data(), _jobs.size(), _results.data(), _ops.data(), cap) != GAMX_OK)
      throw std::runtime_error(gamx_last_error(_ctx));
  }
  const std::vector<gamx_result>& results() const { return _results; }
  MyAlignment alignment(size_t k) const {
    return MyAlignment::from_result(_results[k], _jobs[k].mode == GAMX_MODE_FULL ? _ops.data() : nullptr);
  }
  void clear() { _jobs.clear(); _results.clear(); }
  size_t size() const { return _jobs.size(); }

 private:
  gamx_ctx* _ctx;
  std::vector<gamx_job> _jobs;
  std::vector<gamx_result> _results;
  std::vector<uint8_t> _ops;
};

class BandedSmithWaterman {
 public:
  typedef long int int_type;
  typedef unsigned long int size_type;

  BandedSmithWaterman() : _gap_score(GAMX_DEFAULT_GAP), _band_size(GAMX_DEFAULT_BAND) {}
  // only gap_score and band_size take effect, as in banded_smith_waterman.cc:48-59 vs :119-162
  BandedSmithWaterman(const ScoreType& /*match*/, const ScoreType& /*mismatch*/, const ScoreType& gap_score,
                      const ScoreType& /*gap_ext*/, const size_type& band_size)
      : _gap_score(gap_score), _band_size(band_size) {}
  explicit BandedSmithWaterman(const size_type& band_size) : _gap_score(GAMX_DEFAULT_GAP), _band_size(band_size) {}

  // banded_smith_waterman.hpp:68-71.  Uploads both sequences, aligns on the GPU, returns the
  // full MyAlignment (edit string included).
  template <class SeqT>
  MyAlignment find_alignment(const SeqT& a, size_type begin_a, size_type end_a, const SeqT& b, size_type begin_b,
                             size_type end_b, bool force_start = false, bool force_end = false) const {
    if (_band_size > 0xffffffffull || _gap_score < INT32_MIN || _gap_score > INT32_MAX)
      throw std::invalid_argument("gamx: band or gap score outside the 32-bit range of gamx_job");
    // One context per calling thread, created on first use and reused: the legacy call pattern is N host threads
    // each issuing synchronous calls (ThreadedBuildPctg.cc:159-169) - a call costs two small uploads and one
    // batch, not the creation of streams and buffers.  The context holds just the two sequences of the call.
    gamx_ctx* ctx = call_context();
    if (gamx_clear_contigs(ctx) != GAMX_OK) throw std::runtime_error(gamx_last_error(ctx));
    AlignBatch batch(ctx);
    const uint32_t ia = batch.add_contig(a), ib = batch.add_contig(b);
    batch.add(ia, false, 0, UINT64_MAX, begin_a, end_a, ib, false, 0, UINT64_MAX, begin_b, end_b,
              (uint32_t)_band_size, (int32_t)_gap_score, force_start, force_end, GAMX_MODE_FULL);
    batch.run();
    return batch.alignment(0);
  }

  // the calling thread's context (first visible GPU), destroyed when the thread ends
  static gamx_ctx* call_context() {
    struct Holder {
      gamx_ctx* c = nullptr;
      ~Holder() { if (c) gamx_destroy(c); }
    };
    static thread_local Holder h;
    if (!h.c && gamx_create(&h.c, nullptr, 1) != GAMX_OK)
      throw std::runtime_error("gamx: no usable CUDA device (this aligner has no CPU fallback)");
    return h.c;
  }

 private:
  const ScoreType _gap_score;
  const size_type _band_size;
};

}  // namespace gamx_dropin
