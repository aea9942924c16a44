// K1s - the warp-level banded overlap DP of bsw_warp.h with 16x2 SIMD cells: every 32-bit register
// holds one cell of TWO jobs (job A in the low half-word, job B in the high one), so a group of LG
// lanes advances a PAIR of jobs per step with the DPX 16x2 instructions
//      m = VIADDMNMX.S16x2(up, U, left)        v = VIADDMNMX.S16x2(diag, Cd, m)
// (two instructions per two cells instead of three per cell), one PRMT per cell pair for the
// substitution bytes and - with the direction store - one LOP3 and two IMAD.
//
// Computes what BandedSmithWaterman::find_alignment computes
// (/root/reference/lib/src/alignment/banded_smith_waterman.cc:69-323); same band coordinates, lane
// stripes, skewed wavefront and tie-breaking as bsw_warp.h (read that header first).  What differs:
//
// Values.  Scores are stored NEGATED and the cell update is a minimum.  A half-word holds  W - base  with
//      W = CL - ((H + beta*j) << SH)      ("clean": low SH bits all set, CL = 2^SH - 1),    beta = -gap,
// SH = 2 with the direction store (the low bits of a fresh minimum are the direction), SH = 1 without:
//      diag: W + Cd,  Cd = -(S << SH) - CL + tag     up: W + U,  U = -((2*gap) << SH) - CL + tagUp     left: W
// with the tags  DIAG match 0 < DIAG mismatch 1 < UP 2 < LEFT 3 = CL: the smallest candidate wins and, on
// equal scores, the smallest tag - exactly the reference's priority diag > up > left (.cc:272-307).  A
// minimum is cleaned by OR-ing CL (one LOP3, needed anyway to strip the tag).  Why negated: the never-written
// region pos < 0 must stay identically 0, i.e. those cells must reproduce themselves, which
// takes Cd in [-CL, 0] - and PRMT can make the constants 0x0000 and 0xffff out of ANY table byte by sign
// replication, so the selector of a position < 0 simply replicates a sign: (W + {0,-1}) | CL = W.  (With
// maxima and AND-cleaning the constant would have to be in [0, CL]; -1 breaks it.)
//
// H itself grows without bound (|H| <= 5 rows), but neighbouring cells cannot differ by much:
// for cells of one row  -8 <= H(i,j) - H(i,j-1) <= 13  and for cells of one column  -4 <= H(i,j) - H(i-1,j) <= 5
// (induction over the recurrence .cc:160-164 with S in [-4,5], gap <= -5; the region pos < 0 obeys the same
// bounds).  So the C cells of a lane stripe span at most 21*(C-1) score units, and every LANE keeps its own
// base: every kRebaseSteps steps a lane subtracts (slot 0 - kT0) from its registers and adds it to its 32-bit
// base; the two values a lane exchanges with its neighbours per step are translated by the difference of the
// two bases ("left": one VIADD.16x2 with dL; "up" from the right: folded into the addend of the last slot, no
// instruction).  With kT0 = -6144 and 256 steps between rebases every live
// half-word stays inside [-18000, -1700] for any band, row count and input ("range budget" below), so the
// 16-bit arithmetic never wraps on a value that is used.
//
// Substitution bytes.  PRMT sees 8 table bytes: 4 for job A's row (indexed by the a-base A,T,C,G), 4 for job
// B's.  The selector of a cell pair is one 16-bit shared-memory entry
//      [ selA | 8+selA | 4+selB | 12+selB ]   (nibble 8+x replicates the sign of byte x)
// so the result is the two sign-extended Cd half-words; positions < 0 take [8 | 8 | 12 | 12].  There is no
// room for N: jobs whose windows hold an N are not run here (the kernel checks the N masks of both windows of
// both jobs first and leaves such jobs to the 32-bit kernel, gamx.cu: k1s_kernel's retry list).  Positions >= |a| are a closed region (nothing flows
// back into filled cells), computed with an arbitrary base; the end-cell search treats them as the
// never-filled zeros they are (.cc:183).  Band column 2B has no "up" neighbour: its U is kUpBlock16; the
// padding columns to its right are reset at every rebase so that they cannot run away from the filled ones.
//
// "Last column" cells (pos == end_a, .cc:197-212) sit in a different slot of a different lane every step.
// Selecting them out of the registers would cost two compare/select pairs per cell pair on the ALU pipe, the
// binding one; instead a lane that holds such a cell dumps its stripe to shared memory and reads the slot
// back by index (LSU pipe, idle otherwise), only while the warp is inside the capture window.
//
// Directions.  A 32-bit word holds the tags of 8 consecutive steps of one band column for both jobs
// (A: low half-word, B: high half-word, oldest tag in the top bit pair of its half); it is stored as the step
// loop accumulates it and converted back to the encoding of bsw_warp.h (LEFT 0, UP 1, DIAG 2/3 = 3 - tag) by
// the traceback (pair_word16); word ((t>>3)*C + k)*LG + l of the PAIR's region.  A pair's region is the two per-job regions of the 32-bit layout side by side, so
// the fallback can use them as they are.
#pragma once
#include "bsw_common.h"
#include "bsw_warp.h"

namespace gamx {

constexpr int kT0 = -6144;            // a lane's slot 0 after a rebase is kT0 + CL
constexpr int kRebaseSteps = 256;     // steps between rebases (power of two, multiple of 8)
constexpr int kUpBlock16 = 32768 - 4096;   // "up" addend of band column 2B: above every live value, no wrap
// range budget (SH = 2, the wider case; units of a quarter score, W falls as H + beta*j rises): after a
// rebase slot 0 = kT0 + 3 and slot k >= kT0 - 84*k; in 256 steps a column drifts by [-20, +16] per step;
// incoming left/right neighbours differ by at most 84, the intermediate up + U adds at most 232:
//   live max <= -6141 + 4096 + 84 + 232 + 20 = -1709      live min >= -6141 - 1428 - 5120 - 107 = -12796
//   padding columns (reset at a rebase, fed by column 2B) >= -12796 - 5120 = -17916
//   padding + kUpBlock16 in [10756, 26963]: no wrap, above every live value.

// ---- 16x2 helpers (DPX on the device; the host forms wrap exactly like the hardware) -------------
GAMX_HD int half_lo(uint32_t v) { return (int)(int16_t)(uint16_t)(v & 0xffffu); }
GAMX_HD int half_hi(uint32_t v) { return (int)(int16_t)(uint16_t)(v >> 16); }
GAMX_HD uint32_t pack2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
GAMX_HD uint32_t vadd2w(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __vadd2(a, b);
#else
  return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16);
#endif
}
GAMX_HD uint32_t vsub2w(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __vsub2(a, b);
#else
  return ((a - b) & 0xffffu) | (((a >> 16) - (b >> 16)) << 16);
#endif
}
// per half-word min(a + b, c), signed, the sum wraps: one VIADDMNMX.S16x2
GAMX_HD uint32_t viaddmin2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __viaddmin_s16x2(a, b, c);
#else
  const uint32_t s = vadd2w(a, b);
  const int lo = half_lo(s) < half_lo(c) ? half_lo(s) : half_lo(c);
  const int hi = half_hi(s) < half_hi(c) ? half_hi(s) : half_hi(c);
  return pack2(lo, hi);
#endif
}
// byte permute with the sign-replicating selector nibbles (8+x): the full PTX default mode
GAMX_HD uint32_t prmt_sx(uint32_t lo, uint32_t hi, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return r;
#else
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  uint32_t r = 0;
  for (int n = 0; n < 4; n++) {
    const uint32_t s = (sel >> (4 * n)) & 15u;
    uint32_t byte = (uint32_t)((v >> (8 * (s & 7u))) & 0xffu);
    if (s & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
    r |= byte << (8 * n);
  }
  return r;
#endif
}
// A value the compiler must keep in a register: the per-slot "up" addends differ in one slot of one lane only,
// and recomputing them (a compare and two selects per cell, as ptxas prefers under register pressure) costs
// as much as the cell update itself.
GAMX_HD uint32_t opaque(uint32_t v) {
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(v));
#endif
  return v;
}
// bitwise select: (a & m) | (b & ~m), one LOP3
GAMX_HD uint32_t bitsel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }

template <int C, int LG>
struct alignas(16) GroupSmem16Raw {
  uint64_t btab[(kTileSteps + LG + 15) / 16 * 16];  // rows of the tile: low word = job A's 4-byte Cd table, high word = job B's (staged in runs of 16)
  uint16_t asel[(kTileSteps + LG * C + 15) / 16 * 16 + 16];  // one PRMT selector per a-position of the tile (both jobs; runs of 16)
  uint16_t cap[2][C * LG];                       // latched "last column" cells (half-words of the lane's frame), [job][slot][lane]
  uint32_t dump[C * LG];                         // a lane's stripe, to read one slot back by index, [slot][lane]
};

// Bank conflicts between the lane groups of a warp.  Every step a lane loads its selectors (16 bits, half-word index
// gl * (C - 1) + s with s common to the warp) and its row table (64 bits, index LG - 1 - gl + s) from ITS GROUP's
// arrays, so which banks a warp instruction touches depends on the distance between two groups' arrays.  With the
// natural struct size that distance is a multiple of 32 words for C = 18 / LG = 4 (band 29..35) and for a few other
// geometries: eight groups on the same banks, 8-way conflicts of every selector load (ncu: 4.6 short-scoreboard
// stall cycles per issue, band 32 at half the rate of band 64).  group_pad_bytes() picks, at compile time, the
// padding (a multiple of 16 bytes) after each group's arrays that minimises the modelled conflict degree.
constexpr int smem_conflict_degree(const int* words, int n) {
  int worst = 0;
  for (int i = 0; i < n; i++) {
    int distinct = 0;  // distinct words on lane i's bank, counted at the first lane that holds each word
    for (int j = 0; j < n; j++) {
      if (((words[j] ^ words[i]) & 31) != 0) continue;
      bool first = true;
      for (int k = 0; k < j; k++) first = first && words[k] != words[j];
      distinct += first ? 1 : 0;
    }
    worst = distinct > worst ? distinct : worst;
  }
  return worst;
}
constexpr int smem_group_cost(int c, int lg, int stride_words) {
  int cost = 0;
  for (int par = 0; par < 2; par++) {  // selector loads: both alignments of the half-word index
    int words[32] = {};
    for (int lane = 0; lane < 32; lane++) words[lane] = (lane / lg) * stride_words + (((lane % lg) * (c - 1) + par) >> 1);
    cost += 8 * smem_conflict_degree(words, 32);  // (about C / 2 of them per step against one table load)
  }
  for (int half = 0; half < 2; half++) {  // the 64-bit table load: two half-warp transactions
    int words[32] = {};
    for (int q = 0; q < 16; q++) {
      const int lane = half * 16 + q, e = lg - 1 - lane % lg;
      words[2 * q] = (lane / lg) * stride_words + 2 * e;
      words[2 * q + 1] = words[2 * q] + 1;
    }
    cost += smem_conflict_degree(words, 32);
  }
  return cost;
}
template <int C, int LG>
constexpr int group_pad_bytes() {
  if (LG >= 32) return 0;  // one group per warp
  const int base = (int)sizeof(GroupSmem16Raw<C, LG>);
  int best = 0, best_cost = smem_group_cost(C, LG, base / 4);
  for (int pad = 16; pad < 128; pad += 16) {
    const int cost = smem_group_cost(C, LG, (base + pad) / 4);
    if (cost < best_cost) { best_cost = cost; best = pad; }
  }
  return best;
}
template <int N> struct SmemPad { uint8_t pad_[N]; };
template <> struct SmemPad<0> {};
template <int C, int LG>
struct alignas(16) GroupSmem16 : GroupSmem16Raw<C, LG>, SmemPad<group_pad_bytes<C, LG>()> {};
template <int C, int LG>
struct WarpSmem16 {
  GroupSmem16<C, LG> g[LG >= 32 ? 1 : 32 / LG];
};

// words of a PAIR's direction region (8 steps per word, the stored form is the step loop's accumulator; an even number of 8-step blocks so that the
// traceback can always read two consecutive blocks as one 16-step word)
GAMX_HD uint64_t k1_dir_words16(int x, int c, int lg) {
  const uint64_t steps = (uint64_t)x + lg + 16;
  return 2 * ((steps + 15) / 16) * (uint64_t)c * lg;
}

// Tile staging reads runs of 16 view positions.  A StageView is a view prepared once per tile: the
// pointer to the packed word that holds view position 0, so that everything after it is 32-bit
// arithmetic on small numbers (bsw_common.h's load_codes16 works on 64-bit store indices).
struct StageView {
  const uint32_t* wp;  // packed word of view position 0
  int r;               // position of view position 0 inside that word (0..15)
  int dir;             // +1 / -1
  uint32_t comp;       // complement: every 2-bit code ^ 1
};
GAMX_HD StageView stage_view(const SeqStore& s, const SeqView& v) {
  StageView sv;
  sv.wp = s.packed + (v.origin >> 4);
  sv.r = (int)(v.origin & 15);
  sv.dir = v.dir;
  sv.comp = v.comp;
  return sv;
}
// 2-bit codes of the view positions p .. p+15 (field q = position p+q, complement applied; garbage where
// the position is outside [0, len)).  Only words that hold a position of [0, len) are touched.
GAMX_HD uint32_t stage_codes16(const StageView& sv, int p, int len) {
  const bool fwd = sv.dir > 0;
  const int i0 = fwd ? sv.r + p : sv.r - p - 15;               // lowest index of the run, relative to word 0
  const int wmin = fwd ? 0 : (sv.r - (len - 1)) >> 4, wmax = fwd ? (sv.r + len - 1) >> 4 : 0;
  int w0 = i0 >> 4, w1 = w0 + 1;
  w0 = imin(imax(w0, wmin), wmax);
  w1 = imin(imax(w1, wmin), wmax);
  const uint32_t pa = sv.wp[w0], pb = sv.wp[w1];
  uint32_t c = funnel_r(pa, pb, 2u * (uint32_t)(i0 & 15));
  if (!fwd) {  // the run was fetched in ascending store order = descending view order
    const uint32_t b = brev32(c);
    c = ((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1);  // bit-reversed pairs back in bit order
  }
  return c ^ (sv.comp * 0x55555555u);
}

// Does the view range [from, to] (view positions, from <= to, inside the view) hold an N?  The lanes of
// a group share the mask words; every lane returns its partial answer (OR-reduce over the group).
GAMX_HD uint32_t range_n_bits(const SeqStore& s, const SeqView& v, int from, int to, int gl, int lg) {
  if (to < from) return 0u;
  const uint32_t* mp = s.nmask + (v.origin >> 5);  // mask word of view position 0
  const int r = (int)(v.origin & 31);
  const int lo = v.dir > 0 ? r + from : r - to, hi = v.dir > 0 ? r + to : r - from;  // bit range relative to that word
  const int m0 = lo >> 5, m1 = hi >> 5;
  uint32_t any = 0;
  for (int m = m0 + gl; m <= m1; m += lg) {
    uint32_t wv = mp[m];
    if (m == m0) wv &= 0xffffffffu << (lo & 31);
    if (m == m1) wv &= 0xffffffffu >> (31 - (hi & 31));
    any |= wv;
  }
  return any;
}
// N bits of both windows of a job (a: every position a cell of the band can read; b: the DP rows)
GAMX_HD uint32_t job_n_bits(const SeqStore& s, const DevJob* J, int gl, int lg) {
  if (!J) return 0u;
  const int a_lo = J->p0 < 0 ? 0 : J->p0;
  const int64_t a_end = (int64_t)J->p0 + J->x - 1 + 2 * (int64_t)J->band;
  const int a_hi = a_end > (int64_t)J->la - 1 ? J->la - 1 : (int)a_end;
  return range_n_bits(s, J->a, a_lo, a_hi, gl, lg) | range_n_bits(s, J->b, 0, J->x - 1, gl, lg);
}

// One warp: 32/LG PAIRS.  JA / JB / outA / outB are per-lane arguments, uniform within a group of LG
// lanes: the group's two jobs (JB null: the high half idles; both null: idle group) and their result
// slots.  Preconditions (the caller checks them, see k1s_kernel): both jobs are "regular" jobs of the
// geometry (C, LG) with the same band and gap, and no N in any window.  pair_dirs: the group's
// direction region.  The end cell, score and layout tag (has_match = 1 + half) go to the result
// records; the traceback kernel completes them.
template <int C, int LG, bool DIRS, class W>
GAMX_HD void warp_align16(W& w, const DevJob* JA, const DevJob* JB, const SeqStore& store, WarpSmem16<C, LG>& wsm,
                          uint32_t* pair_dirs, DevResult* outA, DevResult* outB) {
  static_assert(stripe_supported(C), "lane stripe width");
  static_assert(LG == 4 || LG == 8 || LG == 16 || LG == 32, "lanes per pair");
  constexpr int SH = DIRS ? 2 : 1;
  constexpr int CL = (1 << SH) - 1;                         // low bits of a clean value
  constexpr uint32_t CL2 = (uint32_t)CL * 0x00010001u;
  constexpr uint32_t ACC0 = 0x00010001u;                    // bias of the direction accumulator (see GAMX16_STEP)
  constexpr int UF = unroll_of(C);
  const int lane = w.lane();
  const int grp = lane / LG, gl = lane % LG;
  GroupSmem16<C, LG>& sm = wsm.g[grp];
  const DevJob* Jh[2] = {JA, JB};
  const bool liveh[2] = {JA != nullptr, JB != nullptr};
  const DevJob* J0 = JA ? JA : JB;   // band and gap are common to the pair
  const bool live = J0 != nullptr;
  // running flush pointer.  An idle group flushes like the others (no predicate in the step loop): its
  // pair_dirs is a sink of C * LG words that it does not advance in.
  uint32_t* fp = DIRS ? pair_dirs + gl : nullptr;
  const int fp_step = live ? C * LG : 0;

  const int B = live ? J0->band : 0, Y = 2 * B + 1;
  const int ld = (Y - 1) / C, kd = (Y - 1) - ld * C;  // lane/slot of band column 2B
  const int gap = live ? J0->gap : -8;
  const int beta = -gap;
  const int j0 = gl * C;
  int X[2], la[2], p0[2], kc[2], tcap0[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    X[h] = liveh[h] ? Jh[h]->x : 0;
    la[h] = liveh[h] ? Jh[h]->la : 0;
    p0[h] = liveh[h] ? Jh[h]->p0 : 0;
    kc[h] = liveh[h] ? Jh[h]->kc : -1;
    tcap0[h] = kc[h] - gl * (C - 1);  // step at which slot 0 holds a "last column" cell (slot k: tcap0 - k)
  }

  // warp-uniform extents
  int t_end = 0, x_min = 0x7fffffff, win_lo = 0x7fffffff, win_hi = -1, pad_end = 0;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    if (!liveh[h]) continue;
    t_end = imax(t_end, X[h] + ld);
    x_min = imin(x_min, X[h]);
    if (kc[h] >= 0) { win_lo = imin(win_lo, kc[h] - (ld + 1) * (C - 1)); win_hi = imax(win_hi, kc[h]); }
    pad_end = imax(pad_end, -p0[h]);  // steps t < -p0 still meet cells with pos < 0
  }
#pragma unroll
  for (int d = LG; d < 32; d <<= 1) {
    t_end = imax(t_end, w.shfl_xor(t_end, d, 32));
    x_min = imin(x_min, w.shfl_xor(x_min, d, 32));
    win_lo = imin(win_lo, w.shfl_xor(win_lo, d, 32));
    win_hi = imax(win_hi, w.shfl_xor(win_hi, d, 32));
    pad_end = imax(pad_end, w.shfl_xor(pad_end, d, 32));
  }
  const int T_total = t_end;

  // Cd bytes (signed): match, mismatch - see the header for the encoding
  const int cdM = -(kScoreMatch * (1 << SH)) - (DIRS ? CL : 0) + (DIRS ? 0 : 0);   // tag DIAG match = 0
  const int cdX = -(kScoreMismatch * (1 << SH)) - (DIRS ? CL : 0) + (DIRS ? 1 : 0); // tag DIAG mismatch = 1
  const uint32_t cdMb = (uint32_t)cdM & 0xffu, cdXb = (uint32_t)cdX & 0xffu;

  uint32_t H[C], acc[C], U[C];
  const int u_up = -((2 * gap) * (1 << SH)) - (DIRS ? CL : 0) + (DIRS ? 2 : 0);    // tag UP = 2
#pragma unroll
  for (int k = 0; k < C; k++) {
    H[k] = 0; acc[k] = ACC0;
    U[k] = opaque((gl == ld && k == kd) ? pack2(kUpBlock16, kUpBlock16) : pack2(u_up, u_up));
  }
  const uint32_t neg1 = (uint32_t)(gap >> 31);  // -1 in a register the compiler cannot fold (FMA-pipe accumulate, see bsw_warp.h)
  // Neighbour exchange.  The value a lane receives is in the SENDER's frame.  "left" (from lane gl-1) is
  // translated by dL, one VIADD.16x2 per step; lane 0 has no left neighbour: its dL is the blocking constant
  // (the shuffle hands it its own last slot, + kUpBlock16 is above every live value, no wrap).  "up" of the
  // last slot (from lane gl+1) needs no instruction of its own: the translation is folded into that slot's
  // "up" addend U[C-1]; lanes from ld on take no "up" from the right: their U[C-1] is the blocking addend.
  const bool has_left = gl != 0, has_right = gl < ld;
  const int kdl = gl < ld ? C - 1 : (gl == ld ? kd : -1);  // last slot of this lane that is a band column
  const uint32_t T02 = pack2(kT0 + CL, kT0 + CL);
  uint32_t* const dumpp = sm.dump + gl;                      // [slot * LG]
  uint16_t* const capp0 = sm.cap[0] + gl;
  uint16_t* const capp1 = sm.cap[1] + gl;
  int base[2] = {0, 0};     // true W = half-word + base
  uint32_t dL = 0;          // base of the left neighbour lane minus this lane's, per half (lane 0: the block)

  // ---- tile staging: selectors of the a-positions, Cd tables of the b-rows ----------------------------
  // (the views are re-read from the job records here: they are needed once per 128 steps only)
#define GAMX16_STAGE_TILE(T0)                                                                          \
  {                                                                                                    \
    w.sync();                                                                                          \
    StageView sa[2], sb[2];                                                                            \
    _Pragma("unroll") for (int h = 0; h < 2; h++) {                                                    \
      sa[h].wp = sb[h].wp = store.packed; sa[h].r = sb[h].r = 0; sa[h].dir = sb[h].dir = 1; sa[h].comp = sb[h].comp = 0; \
      if (liveh[h]) { sa[h] = stage_view(store, Jh[h]->a); sb[h] = stage_view(store, Jh[h]->b); }      \
    }                                                                                                  \
    const int na = kTileSteps + LG * C - (LG - 1);                                                     \
    for (int c0 = gl * 16; c0 < na; c0 += LG * 16) {                                                   \
      uint32_t byt[2][4];  /* [job][positions 4q .. 4q+3]: selector byte per position */              \
      _Pragma("unroll") for (int h = 0; h < 2; h++) {                                                  \
        const int pos0 = p0[h] + (T0) + c0;                                                            \
        uint32_t codes = 0;                                                                            \
        if (liveh[h] && pos0 + 15 >= 0 && pos0 < la[h]) codes = stage_codes16(sa[h], pos0, la[h]);     \
        const uint32_t tag = h ? 0xc4c4c4c4u : 0x80808080u;                                            \
        const uint32_t padb = h ? 0xccccccccu : 0x88888888u;  /* pos < 0: both nibbles replicate a sign */ \
        _Pragma("unroll") for (int q = 0; q < 4; q++) {                                                \
          uint32_t x = (codes >> (8 * q)) & 0xffu;          /* four 2-bit codes */                     \
          x = (x | (x << 12)) & 0x000f000fu;                                                           \
          x = (x | (x << 6)) & 0x03030303u;                 /* one code per byte */                    \
          x = (x * 0x11u) | tag;                            /* [sel | 8+sel] resp. [4+sel | 12+sel] */ \
          const int nneg = -(pos0 + 4 * q);                 /* positions of this quad that are < 0 */  \
          if (nneg > 0) {                                                                              \
            const uint32_t mk = nneg >= 4 ? 0xffffffffu : ((1u << (8 * nneg)) - 1u);                   \
            x = (x & ~mk) | (padb & mk);                                                               \
          }                                                                                            \
          byt[h][q] = x;                                                                               \
        }                                                                                              \
      }                                                                                                \
      Quad q0, q1;                                                                                     \
      q0.v[0] = prmt(byt[0][0], byt[1][0], 0x5140u); q0.v[1] = prmt(byt[0][0], byt[1][0], 0x7362u);    \
      q0.v[2] = prmt(byt[0][1], byt[1][1], 0x5140u); q0.v[3] = prmt(byt[0][1], byt[1][1], 0x7362u);    \
      q1.v[0] = prmt(byt[0][2], byt[1][2], 0x5140u); q1.v[1] = prmt(byt[0][2], byt[1][2], 0x7362u);    \
      q1.v[2] = prmt(byt[0][3], byt[1][3], 0x5140u); q1.v[3] = prmt(byt[0][3], byt[1][3], 0x7362u);    \
      *reinterpret_cast<Quad*>(sm.asel + c0) = q0;    /* c0 % 16 == 0: 32-byte aligned */              \
      *reinterpret_cast<Quad*>(sm.asel + c0 + 8) = q1;                                                 \
    }                                                                                                  \
    const int nb = kTileSteps + LG - 1;                                                                \
    for (int r0 = gl * 16; r0 < nb; r0 += LG * 16) {                                                   \
      const int i0 = (T0) - (LG - 1) + r0;  /* table r0+q is row i0+q */                               \
      uint32_t cds[2];                                                                                 \
      _Pragma("unroll") for (int h = 0; h < 2; h++) {                                                  \
        cds[h] = 0;                                                                                    \
        if (liveh[h] && i0 + 15 >= 0 && i0 < X[h]) cds[h] = stage_codes16(sb[h], i0, X[h]);            \
      }                                                                                                \
      _Pragma("unroll") for (int q = 0; q < 16; q++) {                                                 \
        const uint32_t ta = (cdXb * 0x01010101u) ^ ((cdXb ^ cdMb) << (8 * ((cds[0] >> (2 * q)) & 3u))); \
        const uint32_t tb = (cdXb * 0x01010101u) ^ ((cdXb ^ cdMb) << (8 * ((cds[1] >> (2 * q)) & 3u))); \
        sm.btab[r0 + q] = ((uint64_t)tb << 32) | ta;                                                   \
      }                                                                                                \
    }                                                                                                  \
    w.sync();                                                                                          \
  }

  // Rebase: slot 0 -> kT0 + CL; the padding columns of the lane are reset (see header); the latched
  // "last column" cells follow the frame; the neighbours' rebase amounts update the base differences.
#define GAMX16_REBASE()                                                                                \
  {                                                                                                    \
    const uint32_t R = vsub2w(H[0], T02);                                                              \
    _Pragma("unroll") for (int k = 0; k < C; k++) {                                                    \
      H[k] = vsub2w(H[k], R);                                                                          \
      if (k > kdl) H[k] = T02;                                                                         \
      capp0[k * LG] = (uint16_t)(capp0[k * LG] - (R & 0xffffu));                                       \
      capp1[k * LG] = (uint16_t)(capp1[k * LG] - (R >> 16));                                           \
    }                                                                                                  \
    base[0] += half_lo(R); base[1] += half_hi(R);                                                      \
    const uint32_t Rl = (uint32_t)w.shfl_up((int)R, 1, LG), Rr = (uint32_t)w.shfl_down((int)R, 1, LG); \
    if (has_left) dL = vsub2w(vadd2w(dL, Rl), R);                                                      \
    if (has_right) U[C - 1] = vsub2w(vadd2w(U[C - 1], Rr), R);                                         \
  }

  // ---- first row, banded_smith_waterman.cc:112-132 (see bsw_warp.h), one job after the other ---------
  GAMX16_STAGE_TILE(0)
  {
    const uint64_t tb = sm.btab[LG - 1];  // row 0
    const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);
    const int kNone = -(1 << 28);
    int wt[2][C];   // clean true W of row 0
    int tg[2][C];   // its tag
#pragma unroll
    for (int h = 0; h < 2; h++) {
      int sc[C], hloc[C];  // substitution score of the cell (kNone: never written), running maximum
      int run = kNone;
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int j = j0 + k, pos = p0[h] + j;
        const uint32_t cdp = prmt_sx(tlo, thi, sm.asel[gl * C + k]);
        const int cd = h ? half_hi(cdp) : half_lo(cdp);
        const bool valid = liveh[h] && pos >= 0 && pos < la[h] && j < Y;
        const int s = (cd == cdM) ? kScoreMatch : kScoreMismatch;  // (no N here: the table holds two values)
        sc[k] = valid ? s : kNone;
        if (valid) run = (pos > 0 && j > 0) ? imax(run, s) : s;
        hloc[k] = run;
      }
      const bool restarts = liveh[h] && (-p0[h] >= j0) && (-p0[h] < j0 + C);  // the cell with pos == 0 is mine
      int incl = run;
      int cut = restarts;
#pragma unroll
      for (int d = 1; d < LG; d <<= 1) {
        const int o = w.shfl_up(incl, d, LG), oc = w.shfl_up(cut, d, LG);
        if (gl >= d) { if (!cut) incl = imax(incl, o); cut |= oc; }
      }
      int in = w.shfl_up(incl, 1, LG);
      if (gl == 0) in = kNone;
      bool open = true;
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int j = j0 + k, pos = p0[h] + j;
        int hh = 0, tag = CL;  // never written: 0
        if (sc[k] != kNone) {
          if (!(pos > 0 && j > 0)) open = false;
          hh = open ? imax(hloc[k], in) : hloc[k];
          // direction of a row-0 cell: DIAG when the cell holds its own substitution score
          if (DIRS && hh == sc[k]) tag = sc[k] == kScoreMatch ? 0 : 1;
        }
        wt[h][k] = CL - (hh + beta * j) * (1 << SH);
        tg[h][k] = tag;
      }
      base[h] = wt[h][0] - (kT0 + CL);
    }
#pragma unroll
    for (int k = 0; k < C; k++) {
      H[k] = pack2(wt[0][k] - base[0], wt[1][k] - base[1]);
      if (DIRS) acc[k] = pack2(tg[0][k], tg[1][k]) + ACC0;
      if (k > kdl) H[k] = T02;
      if (tcap0[0] - k == gl) sm.cap[0][k * LG + gl] = (uint16_t)(H[k] & 0xffffu);  // "last column" cell in row 0
      if (tcap0[1] - k == gl) sm.cap[1][k * LG + gl] = (uint16_t)(H[k] >> 16);
    }
    {
      const int bl0 = w.shfl_up(base[0], 1, LG), bl1 = w.shfl_up(base[1], 1, LG);
      const int br0 = w.shfl_down(base[0], 1, LG), br1 = w.shfl_down(base[1], 1, LG);
      dL = has_left ? pack2(bl0 - base[0], bl1 - base[1]) : pack2(kUpBlock16, kUpBlock16);
      U[C - 1] = has_right ? pack2(u_up + br0 - base[0], u_up + br1 - base[1]) : pack2(kUpBlock16, kUpBlock16);
    }
    if (DIRS && T_total == 1) {  // a single row on a single lane: no step will flush its directions
#pragma unroll
      for (int k = 0; k < C; k++) if (live) fp[k * LG] = 0x00010000u - (((0x00010000u - acc[k]) & 0x00030003u) << 14);
    }
  }

  // One step tt; PA points at the selector of this lane's slot 0, PB at its row table.
  // KIND 0 (FAST): every lane is on a row in [1, X-1] of both jobs, nothing to latch; the caller flushes.
  //                Without the direction store minima need no cleaning, except while cells with pos < 0
  //                are around (their Cd is 0 or -1): KIND 3 is FAST with the cleaning OR.
  // KIND 4       : FAST inside the capture window: lanes that hold a "last column" cell latch it.
  // KIND 1 (SLOW): lanes outside their row range keep their registers (per job; pipeline drain, jobs of
  //                different length), capture, the direction flush is decided per step.
  // KIND 2       : SLOW while lanes are still waiting for their first row (pipeline fill, t < LG).
  // Direction accumulator: acc holds (tags so far) + ACC0, because a step adds  v - (v | CL)  = tag - CL
  // per half instead of the tag: the constant keeps the recurrence  acc = 4*acc + v - hc  exact (two IMAD on
  // the FMA pipe).  The accumulator is stored as it is; the word of tags is ~(acc - ACC0) = 0x10000 - acc,
  // which also turns the tags into the LEFT 0 / UP 1 / DIAG 2,3 encoding of bsw_warp.h - the traceback does
  // that subtraction on the few words it reads (pair_word16).  acc restarts at ACC0 after every 8-step word
  // (the low half must not shift into the high one).
#define GAMX16_CAPTURE(TT)                                                                             \
  {                                                                                                    \
    const int dcap0 = tcap0[0] - (TT), dcap1 = tcap0[1] - (TT);  /* the slot that is on a "last column" cell */ \
    if ((unsigned)dcap0 < (unsigned)C || (unsigned)dcap1 < (unsigned)C) {                              \
      _Pragma("unroll") for (int k = 0; k < C; k++) dumpp[k * LG] = H[k];                              \
      if ((unsigned)dcap0 < (unsigned)C) capp0[dcap0 * LG] = (uint16_t)(dumpp[dcap0 * LG] & 0xffffu);  \
      if ((unsigned)dcap1 < (unsigned)C) capp1[dcap1 * LG] = (uint16_t)(dumpp[dcap1 * LG] >> 16);      \
    }                                                                                                  \
  }
#define GAMX16_STEP(KIND, TT, PA, PB)                                                                  \
  {                                                                                                    \
    constexpr bool kSlow = (KIND) == 1 || (KIND) == 2;                                                 \
    const int tt = (TT);                                                                               \
    const uint64_t tb = *(PB);                                                                         \
    const uint32_t tlo = (uint32_t)tb, thi = (uint32_t)(tb >> 32);                                     \
    uint32_t left = vadd2w((uint32_t)w.shfl_up((int)H[C - 1], 1, LG), dL);                             \
    const bool started = (KIND) != 2 || (tt - gl >= 1);                                                \
    uint32_t keep = 0xffffffffu;                                                                       \
    if (kSlow)                                                                                         \
      keep = ((started && tt - gl < X[0]) ? 0xffffu : 0u) | ((started && tt - gl < X[1]) ? 0xffff0000u : 0u); \
    uint32_t right = 0;                                                                                \
    _Pragma("unroll") for (int k = 0; k < C; k++) {                                                    \
      const uint32_t cd = prmt_sx(tlo, thi, (PA)[k]);                                                  \
      const uint32_t up = (k == C - 1) ? right : H[(k + 1) % C];                                       \
      const uint32_t m = viaddmin2(up, U[k], left);                                                    \
      const uint32_t v = viaddmin2(H[k], cd, m);                                                       \
      const uint32_t hc = (DIRS || (KIND) != 0) ? (v | CL2) : v;                                       \
      if (DIRS && started) acc[k] = (acc[k] * 4u + v) + neg1 * hc;                                     \
      H[k] = kSlow ? bitsel(keep, hc, H[k]) : hc;                                                      \
      left = H[k];                                                                                     \
      if (k == 0) right = (uint32_t)w.shfl_down((int)H[0], 1, LG);                                     \
    }                                                                                                  \
    if ((KIND) == 4) GAMX16_CAPTURE(tt)                                                                \
    if (kSlow) {                                                                                       \
      if (tt >= win_lo && tt <= win_hi) GAMX16_CAPTURE(tt)                                             \
      if (DIRS && ((tt & 7) == 7 || tt == T_total - 1)) {                                              \
        if ((tt & 7) == 7) {                                                                           \
          _Pragma("unroll") for (int k = 0; k < C; k++) fp[k * LG] = acc[k];                           \
        } else {  /* the last word of the job, partly filled: the tags move to the top of their half */ \
          const int sh = 2 * (7 - (tt & 7));                                                           \
          _Pragma("unroll") for (int k = 0; k < C; k++) {                                              \
            const uint32_t wd = 0x00010000u - acc[k];                                                  \
            fp[k * LG] = 0x00010000u - (((wd << sh) & 0xffffu) | (((wd >> 16) << sh) << 16));          \
          }                                                                                            \
        }                                                                                              \
        if (tt - gl >= 0) {  /* (a lane still before its row 0 keeps that row's tags) */               \
          _Pragma("unroll") for (int k = 0; k < C; k++) acc[k] = ACC0;                                 \
        }                                                                                              \
        fp += fp_step;                                                                                 \
      }                                                                                                \
    }                                                                                                  \
  }

  // Steady state runs in BLOCKS of 8 steps that start at multiples of 8 (= one direction word per slot, and
  // tile and rebase boundaries are block boundaries): UF steps unrolled, 8 / UF times, then the flush -
  // no per-step tests.  KIND is fixed per run of blocks.
#define GAMX16_BLOCKS(KIND)                                                                            \
  do {                                                                                                 \
    const uint16_t* const pa_end = pa_t + 8;                                                           \
    int tq = t;                                                                                        \
    _Pragma("unroll 1") do {                                                                           \
      _Pragma("unroll") for (int u = 0; u < UF; u++) GAMX16_STEP(KIND, tq + u, pa_t + u, pb_t + u)     \
      pa_t += UF; pb_t += UF; tq += UF;                                                                \
    } while (pa_t != pa_end);                                                                          \
    t += 8;                                                                                            \
    if (DIRS) {                                                                                        \
      _Pragma("unroll") for (int k = 0; k < C; k++) { fp[k * LG] = acc[k]; acc[k] = ACC0; }            \
      fp += fp_step;                                                                                   \
    }                                                                                                  \
    if ((t & (kRebaseSteps - 1)) == 0) GAMX16_REBASE()                                                 \
  } while (--nb > 0 && ((KIND) != 4 || t <= win_hi));

  int t = 1;  // row 0 is done; lane gl starts its row 1 at step gl + 1
  int t0 = 0;
  while (t < T_total) {
    if (t >= t0 + kTileSteps) { t0 += kTileSteps; GAMX16_STAGE_TILE(t0) }
    const uint16_t* pa = sm.asel + (gl * (C - 1) - t0);    // pa[t]: selector of this lane's slot 0 at step t
    const uint64_t* pb = sm.btab + ((LG - 1) - gl - t0);   // pb[t]: table of row t - gl
    const int stop = imin(t0 + kTileSteps, T_total);       // first step this tile does not cover
    while (t < stop) {
      // steps [t, fast_hi) are steady state: every lane has started (t >= LG) and is on a row < X of
      // both jobs, the last step (partial flush) is excluded.  The variant is fixed per run of blocks:
      // capture window or not, cells with pos < 0 around or not (score only).
      int fast_hi = imin(stop, imin(x_min, T_total - 1));
      const bool capture = t + 8 > win_lo && t <= win_hi;  // some step of the next block may meet a "last column" cell
      if (!capture && t <= win_hi) fast_hi = imin(fast_hi, win_lo & ~7);
      const bool padding = !DIRS && t < pad_end;  // cells with pos < 0 around: minima need the cleaning OR
      if (padding && !capture) fast_hi = imin(fast_hi, (pad_end + 7) & ~7);
      int nb = (t >= LG && (t & 7) == 0) ? (fast_hi - t) >> 3 : 0;
      if (nb > 0) {
        const uint16_t* pa_t = pa + t;
        const uint64_t* pb_t = pb + t;
        if (capture) GAMX16_BLOCKS(4)
        else if (padding) GAMX16_BLOCKS(3)
        else GAMX16_BLOCKS(0)
      } else {
        const int slow_stop = imin(stop, (t & ~7) + 8);  // up to the next block boundary
        do {
          if (t < LG) GAMX16_STEP(2, t, pa + t, pb + t)
          else GAMX16_STEP(1, t, pa + t, pb + t)
          t++;
          if ((t & (kRebaseSteps - 1)) == 0) GAMX16_REBASE()
        } while (t < slow_stop);
      }
    }
  }
#undef GAMX16_BLOCKS
#undef GAMX16_CAPTURE
#undef GAMX16_STEP
#undef GAMX16_REBASE
#undef GAMX16_STAGE_TILE

  // ---- end-cell selection, .cc:174-212: last row (columns ascending) before last column, per job ----
#pragma unroll
  for (int h = 0; h < 2; h++) {
    EndBest best;
    best.found = 0; best.val = 0; best.ord = 0;
    const DevJob* Jp = Jh[h];
    if (liveh[h]) {
#pragma unroll
      for (int k = 0; k < C; k++) {
        const int j = j0 + k;
        if (j >= Jp->jlo && j <= Jp->jhi) {
          const int wtrue = (h ? half_hi(H[k]) : half_lo(H[k])) + base[h];
          const int val = (j < Jp->jfill) ? (((CL - wtrue) >> SH) - beta * j) : 0;
          best.consider(val, j);
        }
      }
      if (kc[h] >= 0) {
#pragma unroll
        for (int k = 0; k < C; k++) {
          const int j = j0 + k, i = tcap0[h] - k - gl;  // the row this slot was on when it met pos == end_a
          if (i >= 0 && i < X[h] && j <= 2 * B && i >= Jp->col_imin) {
            const int wtrue = (int)(int16_t)sm.cap[h][k * LG + gl] + base[h];
            const int val = Jp->col_zero ? 0 : (((CL - wtrue) >> SH) - beta * j);
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
    if (liveh[h] && gl == 0) {
      DevResult R;
      R.status = kStatusOk; R.score = 0; R.end_i = 0; R.end_j = 0;
      R.has_match = 1 + h;  // layout tag for the traceback kernel: half h of a 16x2 pair region
      R.n_ops = R.n_match = R.n_mismatch = R.n_gap_a = R.n_gap_b = 0;
      R.tail_gap_a = R.tail_gap_b = 0;
      R.begin_a = R.begin_bx = 0;
      R.first_match_a = R.first_match_x = R.last_match_a = R.last_match_x = 0;
      R.ops_start = 0;
      if (!best.found) {
        R.status = kStatusEmpty;  // .cc:215
      } else {
        const int ei = best.ord < Y ? X[h] - 1 : best.ord - Y;
        const int ej = best.ord < Y ? best.ord : kc[h] - ei;
        R.score = best.val; R.end_i = ei; R.end_j = ej;
        if (p0[h] + ei + ej >= la[h]) R.status = kStatusOutOfRange;  // first traceback step reads a.at(pos), .cc:231/:265
      }
#ifdef GAMX_DEBUG_PATH2
      printf("  half %d: found %d val %d ord %d status %d X %d la %d p0 %d kc %d jlo %d jhi %d jfill %d base %d H0 %d\n", h, best.found, best.val, best.ord, R.status, X[h], la[h], p0[h], kc[h], Jp->jlo, Jp->jhi, Jp->jfill, base[h], h ? half_hi(H[0]) : half_lo(H[0]));
#endif
      *(h ? outB : outA) = R;
    }
  }
  w.sync();  // the shared-memory tiles are free for the next pair
}

// Traceback fetcher for half `half` of a pair region: two consecutive 8-step blocks make the 16-step
// word k1_traceback_t expects (oldest tag on top).
// a0, a1: the stored accumulators of the two blocks (see GAMX16_STEP: word of tags = 0x10000 - accumulator)
GAMX_HD uint32_t pair_word16(uint32_t a0, uint32_t a1, int half) {
  const uint32_t w0 = 0x00010000u - a0, w1 = 0x00010000u - a1;
  return half ? ((w0 & 0xffff0000u) | (w1 >> 16)) : ((w0 << 16) | (w1 & 0xffffu));
}
struct PairFetch {
  const uint32_t* dirs;
  int C, LG, half;
  GAMX_HD uint32_t operator()(int blk, int k, int l) const {
    const uint32_t a0 = dirs[((uint32_t)(2 * blk) * (uint32_t)C + (uint32_t)k) * (uint32_t)LG + (uint32_t)l];
    const uint32_t a1 = dirs[((uint32_t)(2 * blk + 1) * (uint32_t)C + (uint32_t)k) * (uint32_t)LG + (uint32_t)l];
    return pair_word16(a0, a1, half);
  }
};

}  // namespace gamx
