// Shared definitions for the banded overlap aligner kernels.
//
// Everything here is host+device: the kernel bodies are written against a small "warp
// policy" so the very same code can be executed by the CPU lane simulator in tests/sim
// (test infrastructure) and by the sm_100a kernels in gamx.cu (the product).
//
// Semantics follow the reference's BandedSmithWaterman::find_alignment
// (/root/reference/lib/src/alignment/banded_smith_waterman.cc:69-323); the exact
// contract is restated in DESIGN.md section 3.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GAMX_HD __host__ __device__ __forceinline__
#define GAMX_D __device__ __forceinline__
#else
#define GAMX_HD inline
#define GAMX_D inline
#endif

namespace gamx {

// ---- constants of the reference -------------------------------------------------------
constexpr int kForceMaxGap = 10;        // FORCE_MAXGAP_LEN, banded_smith_waterman.hpp:37
constexpr int kMaxAlignment = 500000;   // BSW_MAX_ALIGNMENT, banded_smith_waterman.hpp:39
constexpr int kScoreMatch = 5;          // SCORING_MATRIX, banded_smith_waterman.cc:80-88
constexpr int kScoreMismatch = -4;

// base codes (nucleotide.hpp:35-43) + a private padding symbol that scores 0 against
// everything; it stands for "a[pos] with pos outside [0, |a|)" (see DESIGN.md 3.3).
constexpr int kCodeN = 4;
constexpr int kCodePad = 5;

// direction tags stored 2 bits per cell; op = tag ^ 1 gives the reference's
// AlignmentAlphabet (my_alignment.hpp:57-62): LEFT->GAP_B(1), UP->GAP_A(0),
// DIAG mismatch->MISMATCH(3), DIAG match->MATCH(2).
constexpr int kTagLeft = 0, kTagUp = 1, kTagDiagMis = 2, kTagDiagMatch = 3;

constexpr int kStatusOk = 0, kStatusEmpty = 1, kStatusOutOfRange = 2, kStatusUndefined = 3;

constexpr int kModeScore = 0, kModeEndpoints = 1, kModeFull = 2;

constexpr int kNegInf = -(1 << 30);   // "no left neighbour" at band column 0
constexpr int kBlock = -(1 << 29);    // addend that removes the "up" candidate at column 2B

// ---- sequence store: 2 bits per base + N bitmask ----------------------------------------
struct SeqStore {
  const uint32_t* packed;  // 16 bases per word, base i at bits [2*(i&15), +2) of word i>>4
  const uint32_t* nmask;   // 32 bases per word, bit (i&31) of word i>>5 set when base i is N
};

// A view (contig, orientation, offset) resolved by the host to "store index of view
// position p" = origin + dir * p; comp = 1 complements (A<->T, C<->G is code ^ 1,
// nucleotide.code.hpp:129-144).
struct SeqView {
  int64_t origin;
  int32_t dir;   // +1 or -1
  uint32_t comp; // 0 or 1
};

GAMX_HD uint32_t load_code(const SeqStore& s, const SeqView& v, int64_t p) {
  const int64_t idx = v.origin + (int64_t)v.dir * p;
  const uint32_t w = s.packed[idx >> 4];
  const uint32_t n = (s.nmask[idx >> 5] >> (idx & 31)) & 1u;
  const uint32_t c = ((w >> (2 * (idx & 15))) & 3u) ^ v.comp;
  return n ? (uint32_t)kCodeN : c;
}

// 64-bit funnel shift right: the low 32 bits of ((hi:lo) >> s), 0 <= s < 32
GAMX_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31u));
#endif
}

// Bulk form of load_code for tile staging: the 16 view positions p .. p+15 with four word loads
// instead of 32.  *codes: 2-bit field q = base code of position p+q (NOT yet complemented; garbage
// where the position is N or outside [0, len)), *nflags: bit q set when position p+q is N.  Only words
// that hold a position of [0, len) are touched (len >= 1), so nothing is read outside the contig.
GAMX_HD void load_codes16(const SeqStore& s, const SeqView& v, int64_t p, int64_t len, uint32_t* codes, uint32_t* nflags) {
  const bool fwd = v.dir > 0;
  const int64_t idx0 = fwd ? v.origin + p : v.origin - p - 15;               // lowest store index of the run
  const int64_t imin = fwd ? v.origin : v.origin - (len - 1);
  const int64_t imax = fwd ? v.origin + (len - 1) : v.origin;
  const int64_t wmin = imin >> 4, wmax = imax >> 4, mmin = imin >> 5, mmax = imax >> 5;
  int64_t w0 = idx0 >> 4, w1 = w0 + 1, m0 = idx0 >> 5, m1 = m0 + 1;
  w0 = w0 < wmin ? wmin : (w0 > wmax ? wmax : w0);
  w1 = w1 < wmin ? wmin : (w1 > wmax ? wmax : w1);
  m0 = m0 < mmin ? mmin : (m0 > mmax ? mmax : m0);
  m1 = m1 < mmin ? mmin : (m1 > mmax ? mmax : m1);
  const uint32_t pa = s.packed[w0], pb = s.packed[w1], na = s.nmask[m0], nb = s.nmask[m1];
  uint32_t c = funnel_r(pa, pb, 2u * (uint32_t)(idx0 & 15));
  uint32_t n = funnel_r(na, nb, (uint32_t)(idx0 & 31)) & 0xffffu;
  if (!fwd) {  // the run was fetched in ascending store order = descending view order
#if defined(__CUDA_ARCH__)
    const uint32_t b = __brev(c);
    n = __brev(n) >> 16;
#else
    uint32_t b = c;
    b = ((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1);
    b = ((b >> 2) & 0x33333333u) | ((b & 0x33333333u) << 2);
    b = ((b >> 4) & 0x0f0f0f0fu) | ((b & 0x0f0f0f0fu) << 4);
    b = ((b >> 8) & 0x00ff00ffu) | ((b & 0x00ff00ffu) << 8);
    b = (b >> 16) | (b << 16);
    uint32_t m = n;
    m = ((m >> 1) & 0x5555u) | ((m & 0x5555u) << 1);
    m = ((m >> 2) & 0x3333u) | ((m & 0x3333u) << 2);
    m = ((m >> 4) & 0x0f0fu) | ((m & 0x0f0fu) << 4);
    n = ((m >> 8) & 0x00ffu) | ((m & 0x00ffu) << 8);
#endif
    c = ((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1);  // bit-reversed pairs back in bit order
  }
  *codes = c;
  *nflags = n;
}

// spreads eight 2-bit fields (16 bits) to eight nibbles (32 bits)
GAMX_HD uint32_t spread2to4(uint32_t x) {
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  return x;
}

// Substitution score of the reference's 5x5 matrix extended with the pad symbol.
GAMX_HD int subst_score(uint32_t a, uint32_t b) {
  if (a == (uint32_t)kCodePad || b == (uint32_t)kCodePad) return 0;
  if (a == b) return kScoreMatch;
  if (a == (uint32_t)kCodeN || b == (uint32_t)kCodeN) return 0;
  return kScoreMismatch;
}
// MATCH op rule of the traceback, banded_smith_waterman.cc:239,274.
GAMX_HD int is_match_op(uint32_t a, uint32_t b) {
  return (a == b) || a == (uint32_t)kCodeN || b == (uint32_t)kCodeN;
}

// ---- job as the kernels see it (prepared by the host in gamx.cu) --------------------
struct DevJob {
  SeqView a;        // view position 0 of a
  SeqView b;        // b.origin already points at view position begin_b (DP row 0)
  int32_t la;       // a.size()
  int32_t p0;       // begin_a - band: a-position of cell (row 0, column 0); may be negative
  int32_t x;        // x_size, DP rows (banded_smith_waterman.cc:93-95)
  int32_t band;     // _band_size; y_size = 2*band+1
  int32_t kc;       // anti-diagonal i+j of the "last column" cells (pos == end_a), -1 if none
  int32_t jlo, jhi; // last-row candidate columns (jhi < jlo: none, e.g. force_end)
  int32_t jfill;    // last-row columns j < jfill are filled cells (pos < |a|)
  int32_t col_imin; // smallest row of a last-column candidate (force_end rule, .cc:201)
  int32_t col_zero; // 1 when end_a >= |a|: last-column cells are never-filled zeros
  int32_t gap;      // _gap_score (negative)
  int32_t mode;     // kMode*
  uint32_t ops_cap; // capacity (in ops, multiple of 16) of this job's ops region
  uint32_t pad_;
  uint64_t ops_word; // first word of this job's ops region in the device ops buffer
};

// raw job for the generic (literal) kernel: every reference argument as it was passed
struct GenJob {
  SeqView a, b;       // view position 0 of a and of b
  uint64_t la, lb;
  uint64_t begin_a, end_a, begin_b, end_b;
  uint64_t band;
  int64_t gap;
  int32_t force_start, force_end;
  int32_t mode;
  uint32_t ops_cap;
  uint64_t ops_word;
  uint64_t x_size;     // host-computed (.cc:93-95)
  uint64_t rows_off;   // int64 offset into the generic row scratch (2*y_size values)
  uint64_t dirs_off;   // word offset into the generic direction scratch
};

struct DevResult {
  int32_t status;
  int32_t score;
  int32_t end_i, end_j;
  int32_t has_match;
  uint32_t n_ops, n_match, n_mismatch, n_gap_a, n_gap_b;
  uint32_t tail_gap_a, tail_gap_b;  // gaps after the last MATCH (my_alignment.cc:265-296)
  int64_t begin_a;                  // pos+1 after the traceback loop (.cc:321)
  int64_t begin_bx;                 // x+1 (host adds begin_b)
  int64_t first_match_a, first_match_x;
  int64_t last_match_a, last_match_x;
  uint64_t ops_start;               // op index of op 0 relative to the device ops buffer
};

// ---- integer helpers (DPX on the device, plain C++ on the host) ---------------------------
GAMX_HD int imax(int a, int b) { return a > b ? a : b; }
GAMX_HD int imin(int a, int b) { return a < b ? a : b; }

// max(a + b, c): one VIADDMNMX on sm_90+/sm_100a
GAMX_HD int viaddmax(int a, int b, int c) {
#if defined(__CUDA_ARCH__)
  return __viaddmax_s32(a, b, c);
#else
  const int s = a + b;
  return s > c ? s : c;
#endif
}

// byte permute: result byte n = byte (sel nibble n & 7) of the 8-byte value {hi,lo}
GAMX_HD uint32_t prmt(uint32_t lo, uint32_t hi, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  // raw PRMT: only the low 16 selector bits are read and all our selector nibbles are <= 7, so
  // the masking __byte_perm adds (one extra LOP3 per selector) is not needed
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
  return r;
#else
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  uint32_t r = 0;
  for (int n = 0; n < 4; n++) {
    const uint32_t s = (sel >> (4 * n)) & 7u;
    r |= (uint32_t)((v >> (8 * s)) & 0xffu) << (8 * n);
  }
  return r;
#endif
}

// c + (byte n of a): IDP.4A with a one-hot multiplier, runs on the FMA pipe at full rate
// (measured 18.5 T lane-ops/s, the same as IMAD and VIADDMNMX)
GAMX_HD int add_byte(uint32_t a, int n, int c) {
#if defined(__CUDA_ARCH__)
  return (int)__dp4a(a, 1u << (8 * n), (uint32_t)c);
#else
  return (int)((uint32_t)c + ((a >> (8 * n)) & 0xffu));
#endif
}

GAMX_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
GAMX_HD int clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __clz((int)v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}

}  // namespace gamx
