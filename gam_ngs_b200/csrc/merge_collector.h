// f1 - batch collector for gam-merge's alignment stage (host side, part of libgamx.so).
//
// Restates, as a per-merge-block state machine that advances in ROUNDS,
//     PctgBuilder::alignMergeBlock     /root/reference/lib/src/pctg/PctgBuilder.cc:726-844
//     PctgBuilder::findBestAlignment   /root/reference/lib/src/pctg/PctgBuilder.cc:1361-1614
//     PctgBuilder::alignBlocks         /root/reference/lib/src/pctg/PctgBuilder.cc:1617-1708
//     PctgBuilder::is_good             /root/reference/lib/src/pctg/PctgBuilder.cc:1711-1730
// The reference runs these synchronously, one find_alignment call at a time per pthread
// (ThreadedBuildPctg.cc:305-339).  Alignments inside a merge block are chained (the next window
// starts at last_match_pos of the previous one, .cc:1662-1666) but merge blocks are independent
// (BuildPctgFunctions.cc:82-84), so every round gathers the next pending alignment (or findHits
// call) of every live merge block and runs them as ONE GPU batch.
//
// Nothing here computes an alignment: all DP and k-mer voting happen in the CUDA kernels via
// gamx_align_batch / gamx_find_hits_batch.  Orientation changes and tails are store views
// (rc flag, offset), never copies.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/gamx.h"

namespace gamx {

constexpr double kMinHomology = 95.0;  // MIN_HOMOLOGY, PctgBuilder.hpp:63-65

struct AlnLite {  // what gam-merge reads from a MyAlignment
  double homology = 0.0;
  uint64_t length = 0;
  uint64_t first_a = 0, first_b = 0, last_a = 0, last_b = 0;
};

struct MergeState {
  enum Phase { kChain, kTailHits, kTailAlign, kDone };
  Phase phase = kChain;
  int status = 0;          // 0 ok, 2 exception (std::out_of_range / std::domain_error in the reference)
  // inputs
  const gamx_merge_block* mb = nullptr;
  const gamx_block* blocks = nullptr;
  uint64_t msz = 0, ssz = 0;
  bool reversed_order = false;   // blocks processed back to front (.cc:1650,1679)
  // orientation state (findBestAlignment)
  double con_prob = 0.0;
  int attempts = 0;              // orientations tried so far (max 2)
  bool rev = false;              // slave currently reverse-complemented
  int64_t master_start = 0, slave_start = 0, slave_end = 0;
  int64_t s_start_fwd = 0, s_end_fwd = 0;   // slave window in the stored (forward) orientation
  int64_t align_threshold = 0, threshold = 0;
  // chain state (alignBlocks)
  uint32_t k = 0;
  int64_t m_at = 0, s_at = 0;
  uint64_t lm_a = 0, lm_b = 0;   // last_match_pos of the previous alignment
  std::vector<AlnLite> aligns;
  // tails
  bool want_left = false, want_right = false, left_rev = false, right_rev = false;
  bool have_left = false, have_right = false;
  uint64_t as_a = 0, as_b = 0, ae_a = 0, ae_b = 0;   // alignStart / alignEnd
  gamx_hits_result left_hits = {}, right_hits = {};
  AlnLite left, right;
  size_t job_main = 0, job_left = 0, job_right = 0;  // indices into the round's batches
  // Orientation speculation: while the first orientation's chain runs, the chain of the OTHER
  // orientation - what findBestAlignment would start only after the first one has failed (.cc:1435-1459,
  // 1485-1509) - advances in the same rounds.  It is consulted only if the first chain turns out bad; a merge
  // block that needs the retry then finishes in max(k) instead of 2k rounds.  Its results are exactly the
  // retry's (the chain depends on the orientation only), so nothing observable changes.
  struct Shadow {
    bool active = false, threw = false;
    bool rev = false;
    uint32_t k = 0;
    int64_t slave_start = 0, slave_end = 0, m_at = 0, s_at = 0;
    uint64_t lm_a = 0, lm_b = 0;
    std::vector<AlnLite> aligns;
    uint64_t cells = 0;      // DP cells of its alignments so far (committed to the statistics only on a switch)
    size_t job = 0;
    bool issued = false;     // a job of this chain is in the current round
  } sh;
};

inline const gamx_block& block_at(const MergeState& st, uint32_t k) {
  const uint32_t n = st.mb->n_blocks;
  return st.blocks[st.mb->first_block + (st.reversed_order ? n - 1 - k : k)];
}
inline int32_t frame_len(int32_t b, int32_t e) { return e < b ? 0 : e - b + 1; }  // Frame.cc:124-127

inline AlnLite lite_from(const gamx_result& r) {
  AlnLite a;
  if (r.status != GAMX_JOB_OK) return a;  // default MyAlignment()
  a.homology = r.homology; a.length = r.n_ops;
  a.first_a = r.first_match_a; a.first_b = r.first_match_b;
  a.last_a = r.last_match_a; a.last_b = r.last_match_b;
  return a;
}

inline bool is_good_list(const std::vector<AlnLite>& v, uint64_t min_len) {  // .cc:1711-1723
  uint64_t total = 0;
  for (const AlnLite& a : v) { if (a.homology < kMinHomology) return false; total += a.length; }
  return total >= min_len;
}
inline bool is_good_one(const AlnLite& a, uint64_t min_len) { return a.homology >= kMinHomology && a.length >= min_len; }

// sets up the chain in orientation `rev` (reverse_complement + coordinate flip, .cc:1443-1448)
inline void start_chain(MergeState& st, bool rev, int64_t s_start_fwd, int64_t s_end_fwd) {
  st.rev = rev;
  if (rev) { st.slave_start = (int64_t)st.ssz - s_end_fwd - 1; st.slave_end = (int64_t)st.ssz - s_start_fwd - 1; }
  else { st.slave_start = s_start_fwd; st.slave_end = s_end_fwd; }
  st.k = 0; st.m_at = st.master_start; st.s_at = st.slave_start; st.lm_a = st.lm_b = 0;
  st.aligns.clear();
  st.phase = MergeState::kChain;
  st.attempts++;
}

// starts the speculative chain of the orientation opposite to the running first attempt
inline void start_shadow(MergeState& st) {
  MergeState::Shadow& h = st.sh;
  h = MergeState::Shadow();
  h.active = true;
  h.rev = !st.rev;
  if (h.rev) { h.slave_start = (int64_t)st.ssz - st.s_end_fwd - 1; h.slave_end = (int64_t)st.ssz - st.s_start_fwd - 1; }
  else { h.slave_start = st.s_start_fwd; h.slave_end = st.s_end_fwd; }
  h.m_at = st.master_start; h.s_at = h.slave_start;
}

// the first orientation failed: the speculative chain becomes the running (second) attempt
inline void adopt_shadow(MergeState& st) {
  MergeState::Shadow& h = st.sh;
  st.rev = h.rev; st.slave_start = h.slave_start; st.slave_end = h.slave_end;
  st.k = h.k; st.m_at = h.m_at; st.s_at = h.s_at; st.lm_a = h.lm_a; st.lm_b = h.lm_b;
  st.aligns.swap(h.aligns);
  st.phase = MergeState::kChain;
  st.attempts++;
  h.active = false;
}

}  // namespace gamx
