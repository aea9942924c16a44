/*
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * CPU-baseline legs may load this file; the product library never links it.
 *
 * Plain-C restatement of the reference's banded, linear-gap overlap aligner
 *   BandedSmithWaterman::find_alignment
 *     /root/reference/lib/src/alignment/banded_smith_waterman.cc:69-323
 * and of the MyAlignment reductions gam-merge reads
 *     /root/reference/lib/src/alignment/my_alignment.cc:167-296.
 *
 * Parity status: PINNED.  The reference publishes no golden vectors
 * (SURVEY.md section 4), so this restatement is pinned by executing the
 * unmodified reference itself (oracle/_ref/libgamref.so, built by
 * oracle/Makefile from /root/reference) on seeded randomized and edge-case
 * inputs (tests/test_oracle.py) and by the committed fixtures in
 * tests/golden/ that were generated from that reference build
 * (tests/golden/make_golden.py).
 *
 * The arithmetic below keeps the reference's types: size_type = uint64_t,
 * int_type = int64_t, ScoreType = int64_t, including the places where the
 * reference compares signed with unsigned values.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BSWO_FORCE_MAXGAP_LEN 10     /* banded_smith_waterman.hpp:37 */
#define BSWO_MAX_ALIGNMENT 500000u   /* banded_smith_waterman.hpp:39 */

/* my_alignment.hpp:57-62 */
enum { BSWO_GAP_A = 0, BSWO_GAP_B = 1, BSWO_MATCH = 2, BSWO_MISMATCH = 3 };

/* status values */
enum {
  BSWO_OK = 0,            /* an alignment was produced                        */
  BSWO_EMPTY = 1,         /* reference returns MyAlignment() (.cc:90, :215)    */
  BSWO_OUT_OF_RANGE = 2,  /* reference throws std::out_of_range (Contig::at)   */
  BSWO_UNDEFINED = 3,     /* reference has undefined behaviour (x_size == 0)   */
  BSWO_NOMEM = 4
};

typedef struct {
  int32_t status;
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
  /* extras the reference does not expose; used to check the CUDA kernels */
  uint64_t n_match;
  uint64_t x_size;
  int64_t end_i, end_j;
} bswo_result;

/* banded_smith_waterman.cc:80-88, index order A,T,C,G,N (nucleotide.hpp:35-43) */
static const int BSWO_S[5][5] = {
    {5, -4, -4, -4, 0}, {-4, 5, -4, -4, 0}, {-4, -4, 5, -4, 0}, {-4, -4, -4, 5, 0}, {0, 0, 0, 0, 5}};

static int64_t max2(int64_t a, int64_t b) { return a > b ? a : b; }

/* my_alignment.cc:167-193 */
int bswo_first_match_pos(const uint8_t* ops, uint64_t n, uint64_t begin_a, uint64_t begin_b,
                         uint64_t* pa, uint64_t* pb) {
  *pa = begin_a;
  *pb = begin_b;
  for (uint64_t i = 0; i < n; i++) {
    switch (ops[i]) {
      case BSWO_MATCH: return 1;
      case BSWO_GAP_A: (*pb)++; break;
      case BSWO_GAP_B: (*pa)++; break;
      default: (*pa)++; (*pb)++; break;
    }
  }
  return 0;
}

/* my_alignment.cc:196-226 */
int bswo_last_pos(const uint8_t* ops, uint64_t n, uint64_t begin_a, uint64_t begin_b,
                  uint64_t* pa, uint64_t* pb) {
  int seen = 0;
  *pa = begin_a;
  *pb = begin_b;
  for (uint64_t i = 0; i < n; i++) {
    switch (ops[i]) {
      case BSWO_MATCH: seen = 1; (*pa)++; (*pb)++; break;
      case BSWO_GAP_A: (*pb)++; break;
      case BSWO_GAP_B: (*pa)++; break;
      default: (*pa)++; (*pb)++; break;
    }
  }
  return seen;
}

/* my_alignment.cc:228-262 */
int bswo_last_match_pos(const uint8_t* ops, uint64_t n, uint64_t begin_a, uint64_t begin_b,
                        uint64_t* pa, uint64_t* pb) {
  uint64_t a = begin_a, b = begin_b;
  int seen = 0;
  *pa = begin_a;
  *pb = begin_b;
  for (uint64_t i = 0; i < n; i++) {
    switch (ops[i]) {
      case BSWO_MATCH: seen = 1; *pa = a; *pb = b; a++; b++; break;
      case BSWO_GAP_A: b++; break;
      case BSWO_GAP_B: a++; break;
      default: a++; b++; break;
    }
  }
  return seen;
}

/* my_alignment.cc:265-296 */
int bswo_gaps_before_last_match(const uint8_t* ops, uint64_t n, uint64_t* ga, uint64_t* gb) {
  uint64_t a = 0, b = 0, la = 0, lb = 0;
  int seen = 0;
  for (uint64_t i = 0; i < n; i++) {
    switch (ops[i]) {
      case BSWO_MATCH: seen = 1; la = a; lb = b; break;
      case BSWO_GAP_A: a++; break;
      case BSWO_GAP_B: b++; break;
      default: break;
    }
  }
  *ga = la;
  *gb = lb;
  return seen;
}

/*
 * a, b: base codes 0..4 (A,T,C,G,N).  ops (may be NULL): receives the edit
 * string front-to-back, at most ops_cap entries (n_ops is always the full
 * length).  gap is the reference's _gap_score (-8 unless the 5-arg ctor is
 * used, banded_smith_waterman.cc:48-59).
 */
int bswo_align(const uint8_t* a, uint64_t la, uint64_t begin_a, uint64_t end_a,
               const uint8_t* b, uint64_t lb, uint64_t begin_b, uint64_t end_b,
               uint64_t band, int64_t gap, int force_start, int force_end,
               bswo_result* r, uint8_t* ops, uint64_t ops_cap) {
  const int64_t FM = BSWO_FORCE_MAXGAP_LEN;
  const int fs = force_start != 0, fe = force_end != 0;
  memset(r, 0, sizeof(*r));

  /* .cc:90-91 */
  if (end_b < begin_b) { r->status = BSWO_EMPTY; return r->status; }
  if (end_b >= lb) end_b = lb - 1; /* wraps when lb == 0, as in the reference */

  /* .cc:93-97 (unsigned arithmetic, wraps like the reference) */
  uint64_t x_size = end_b - begin_b + 1;
  uint64_t lim = la + band - begin_a;
  if (lim < x_size) x_size = lim;
  if (x_size > BSWO_MAX_ALIGNMENT) x_size = BSWO_MAX_ALIGNMENT;
  const uint64_t y_size = 2 * band + 1;
  r->x_size = x_size;
  if (x_size == 0) { r->status = BSWO_UNDEFINED; return r->status; }

  /* .cc:102-107: zero-initialised x_size * y_size matrix */
  int64_t* sw = (int64_t*)calloc((size_t)(x_size * y_size), sizeof(int64_t));
  if (!sw) { r->status = BSWO_NOMEM; return r->status; }
#define SW(i, j) sw[(uint64_t)(i) * y_size + (uint64_t)(j)]
#define THROW() do { free(sw); r->status = BSWO_OUT_OF_RANGE; return r->status; } while (0)
#define A_AT(p) do { if ((uint64_t)(p) >= la) THROW(); } while (0)
#define B_AT(p) do { if ((uint64_t)(p) >= lb) THROW(); } while (0)

  /* .cc:112-132: first row */
  for (uint64_t j = 0; j < y_size; j++) {
    int64_t pos = (int64_t)(begin_a - band + j);
    if ((!fs && pos >= 0 && (uint64_t)pos < la) || (fs && pos >= 0 && pos <= FM)) {
      A_AT(pos); B_AT(begin_b);
      int64_t diag = BSWO_S[a[pos]][b[begin_b]];
      int64_t up = gap;
      int c = (pos > 0 && j > 0);
      int64_t left = c ? SW(0, j - 1) : gap; /* note: no gap penalty added (.cc:120) */
      SW(0, j) = c ? max2(max2(diag, up), left) : max2(up, diag);
    }
    if (fs && pos > FM && (uint64_t)pos < la) {
      A_AT(pos); B_AT(begin_b);
      int64_t diag = BSWO_S[a[pos]][b[begin_b]];
      int c = (pos > 0 && j > 0);
      int64_t left = c ? SW(0, j - 1) : gap;
      SW(0, j) = c ? max2(diag, left) : diag;
    }
  }

  /* .cc:135-171: fill */
  for (uint64_t i = 1; i < x_size; i++) {
    for (uint64_t j = 0; j < y_size; j++) {
      int64_t pos = (int64_t)(begin_a + i + j - band);
      if (pos >= 0 && (uint64_t)pos < la) {
        B_AT(begin_b + i);
        int64_t s = BSWO_S[a[pos]][b[begin_b + i]];
        if ((!fs && pos == 0) || (fs && pos == 0 && (int64_t)i <= FM)) {
          int64_t up = (j < y_size - 1) ? SW(i - 1, j + 1) + gap : gap;
          int64_t left = gap;
          SW(i, j) = (j < y_size - 1) ? max2(max2(s, up), left) : max2(s, left);
        } else if (fs && pos == 0 && (int64_t)i > FM) {
          int64_t up = (j < y_size - 1) ? SW(i - 1, j + 1) + gap : gap;
          SW(i, j) = (j < y_size - 1) ? max2(s, up) : s;
        } else {
          int64_t diag = SW(i - 1, j) + s;
          int64_t up = (j < y_size - 1) ? SW(i - 1, j + 1) + gap : gap;
          int64_t left = (j > 0) ? SW(i, j - 1) + gap : gap;
          if (j < y_size - 1 && j > 0) SW(i, j) = max2(max2(diag, up), left);
          else if (j < y_size - 1) SW(i, j) = max2(diag, up);
          else if (j > 0) SW(i, j) = max2(diag, left);
          else SW(i, j) = diag;
        }
      }
    }
  }

  /* .cc:174-212: end-cell selection */
  int found = 0;
  int64_t max_i = 0, max_j = 0, max_score = 0;
  for (uint64_t j = 0; !fe && j < y_size; j++) {
    int64_t pos = (int64_t)(begin_a + (x_size - 1) + j - band);
    if (pos >= 0 && (uint64_t)pos <= end_a) {
      if (!found || SW(x_size - 1, j) > max_score) {
        found = 1; max_i = (int64_t)(x_size - 1); max_j = (int64_t)j;
        max_score = SW(x_size - 1, j);
      }
    }
  }
  {
    int ge = (uint64_t)(int64_t)end_a >= (begin_a + band);
    int64_t i = ge ? (int64_t)end_a - (int64_t)(begin_a + band) : 0;
    int64_t j = ge ? (int64_t)(2 * band) : (int64_t)(2 * band - (begin_a + band - end_a));
    for (; (uint64_t)i < x_size && j >= 0; i++) {
      if (!fe || (fe && (uint64_t)i >= x_size - 1 - (uint64_t)FM && (uint64_t)i < x_size)) {
        if (!found || SW(i, j) > max_score) {
          found = 1; max_i = i; max_j = j; max_score = SW(i, j);
        }
      }
      j--;
    }
  }
  if (!found) { free(sw); r->status = BSWO_EMPTY; return r->status; } /* .cc:215 */

  /* .cc:220-311: traceback; ops are produced back-to-front */
  uint64_t cap = x_size + y_size + 16, n = 0, n_match = 0;
  uint8_t* rev = (uint8_t*)malloc((size_t)cap);
  if (!rev) { free(sw); r->status = BSWO_NOMEM; return r->status; }
#undef THROW
#define THROW() do { free(sw); free(rev); r->status = BSWO_OUT_OF_RANGE; return r->status; } while (0)
#define PUSH(op) do { if (n == cap) { cap *= 2; rev = (uint8_t*)realloc(rev, (size_t)cap); } rev[n++] = (op); } while (0)
  int64_t x = max_i, y = max_j;
  int64_t pos = (int64_t)(begin_a + (uint64_t)x + (uint64_t)y - band);
  while (x >= 0 && y >= 0 && pos >= 0) {
    A_AT(pos); B_AT(begin_b + (uint64_t)x);
    uint8_t ca = a[pos], cb = b[begin_b + (uint64_t)x];
    int64_t s = BSWO_S[ca][cb];
    int is_match = (ca == cb) || ca == 4 || cb == 4; /* .cc:239, :274 */
    if (pos == 0) {
      int64_t diag = s;
      int64_t left = gap;
      int left_ok = !(fs && x > FM); /* .cc:235: left = INT64_MIN */
      if (SW(x, y) == diag) {
        PUSH(is_match ? BSWO_MATCH : BSWO_MISMATCH); n_match += is_match; x--;
      } else if (y == (int64_t)y_size - 1 || (left_ok && SW(x, y) == left)) {
        PUSH(BSWO_GAP_B); y--;
      } else {
        PUSH(BSWO_GAP_A); x--; y++;
      }
    } else {
      int64_t diag = (x > 0 ? SW(x - 1, y) : 0) + s;
      int64_t up = (x > 0 && y < (int64_t)y_size - 1) ? SW(x - 1, y + 1) + gap : gap;
      int up_ok = 1;
      if (fs && x == 0 && pos >= 0 && pos <= FM) up = gap;   /* .cc:269 */
      else if (fs && x == 0) up_ok = 0;                      /* .cc:270: INT64_MIN */
      if (SW(x, y) == diag) {
        PUSH(is_match ? BSWO_MATCH : BSWO_MISMATCH); n_match += is_match; x--;
      } else if (y < (int64_t)y_size - 1 && y > 0 && up_ok && SW(x, y) == up) {
        PUSH(BSWO_GAP_A); x--; y++;
      } else if (y < (int64_t)y_size - 1 && y > 0) {
        PUSH(BSWO_GAP_B); y--;
      } else if (y < (int64_t)y_size - 1) {
        PUSH(BSWO_GAP_A); x--; y++;
      } else {
        PUSH(BSWO_GAP_B); y--;
      }
    }
    pos = (int64_t)(begin_a + (uint64_t)x + (uint64_t)y - band);
  }

  /* .cc:319-321 */
  r->status = BSWO_OK;
  r->score = max_score;
  r->begin_a = (uint64_t)(pos + 1);
  r->begin_b = begin_b + (uint64_t)x + 1;
  r->a_size = la;
  r->b_size = lb;
  r->n_ops = n;
  r->n_match = n_match;
  r->homology = (n == 0) ? 0.0 : (double)(n_match * 100) / (double)n;
  r->end_i = max_i;
  r->end_j = max_j;

  /* reverse into front-to-back order for the reductions */
  for (uint64_t k = 0; k < n / 2; k++) { uint8_t t = rev[k]; rev[k] = rev[n - 1 - k]; rev[n - 1 - k] = t; }
  r->has_first_match = bswo_first_match_pos(rev, n, r->begin_a, r->begin_b, &r->first_match_a, &r->first_match_b);
  r->has_last_match = bswo_last_match_pos(rev, n, r->begin_a, r->begin_b, &r->last_match_a, &r->last_match_b);
  r->has_last_pos = bswo_last_pos(rev, n, r->begin_a, r->begin_b, &r->last_pos_a, &r->last_pos_b);
  r->has_gaps = bswo_gaps_before_last_match(rev, n, &r->gaps_a, &r->gaps_b);
  if (ops) memcpy(ops, rev, (size_t)(n < ops_cap ? n : ops_cap));
  free(rev);
  free(sw);
  return r->status;
#undef SW
#undef THROW
#undef A_AT
#undef B_AT
#undef PUSH
}

/* Cell count of one job: x_size * (2*band+1), banded_smith_waterman.cc:93-97. */
uint64_t bswo_cells(uint64_t la, uint64_t begin_a, uint64_t lb, uint64_t begin_b, uint64_t end_b,
                    uint64_t band) {
  if (end_b < begin_b) return 0;
  if (end_b >= lb) end_b = lb - 1;
  uint64_t x = end_b - begin_b + 1, lim = la + band - begin_a;
  if (lim < x) x = lim;
  if (x > BSWO_MAX_ALIGNMENT) x = BSWO_MAX_ALIGNMENT;
  return x * (2 * band + 1);
}

/* ------------------------------------------------------------------------------------------------
 * ABlast::findHits restated:
 *   /root/reference/lib/src/alignment/ablast.cc:41-76, helpers lib/include/alignment/ablast.hpp:52-107.
 * k-mer (w = 20) diagonal voting.  Codes are radix-4 numbers whose digits are the base codes 0..4
 * (N = 4 aliases with a carry, ablast.hpp:53-59); only diagonals idx_a >= idx_b are counted
 * (ablast.hpp:71-76).  Returns the number of hits (diagonals with the maximal non-zero count, in
 * ascending order) and writes up to cap of them as a_start + d truncated to 32 bits (ablast.cc:58-73).
 * Pinned against the compiled reference (gamref_find_hits) in tests/test_oracle.py.
 * ---------------------------------------------------------------------------------------------- */
#define BSWO_WORD 20 /* ABLAST_DEFAULT_WORD_SIZE, ablast.hpp:33 */

typedef struct { uint64_t code; uint64_t pos; } bswo_kmer;
static int bswo_kmer_cmp(const void* x, const void* y) {
  const bswo_kmer* a = (const bswo_kmer*)x; const bswo_kmer* b = (const bswo_kmer*)y;
  if (a->code != b->code) return a->code < b->code ? -1 : 1;
  return a->pos < b->pos ? -1 : (a->pos > b->pos);
}
static uint64_t bswo_code(const uint8_t* s, uint64_t p) {
  uint64_t c = 0;
  for (uint64_t i = p; i < p + BSWO_WORD; i++) c = 4 * c + s[i]; /* (LAST_BASE-1)*code + base */
  return c;
}

uint64_t bswo_find_hits(const uint8_t* a, uint64_t la, uint64_t a_start, uint64_t a_end,
                        const uint8_t* b, uint64_t lb, uint64_t b_start, uint64_t b_end,
                        uint32_t* hits, uint64_t cap, uint64_t* max_count_out) {
  if (max_count_out) *max_count_out = 0;
  if (la == 0 || lb == 0) return 0;                                   /* ablast.cc:47 */
  if (a_end >= la) a_end = la - 1;                                    /* :49-50 */
  if (b_end >= lb) b_end = lb - 1;
  if (a_start > a_end || b_start > b_end) return 0;                   /* :52 */
  if (a_end + 1 < BSWO_WORD + a_start || b_end + 1 < BSWO_WORD + b_start) return 0; /* :53 */
  const uint64_t na = a_end - BSWO_WORD + 1 - a_start + 1, nb = b_end - BSWO_WORD + 1 - b_start + 1;
  const uint64_t nf = a_end - a_start + 1;
  bswo_kmer* ka = (bswo_kmer*)malloc((size_t)na * sizeof(bswo_kmer));
  uint64_t* f = (uint64_t*)calloc((size_t)nf, sizeof(uint64_t));
  if (!ka || !f) { free(ka); free(f); return 0; }
  for (uint64_t i = 0; i < na; i++) { ka[i].code = bswo_code(a, a_start + i); ka[i].pos = a_start + i; }
  qsort(ka, (size_t)na, sizeof(bswo_kmer), bswo_kmer_cmp);
  for (uint64_t j = 0; j < nb; j++) {
    const uint64_t code = bswo_code(b, b_start + j);
    uint64_t lo = 0, hi = na; /* lower bound */
    while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (ka[mid].code < code) lo = mid + 1; else hi = mid; }
    for (uint64_t k = lo; k < na && ka[k].code == code; k++) {
      const uint64_t idx_a = ka[k].pos - a_start, idx_b = j;
      if (idx_a >= idx_b) f[idx_a - idx_b] += 1;                      /* mark_found, ablast.hpp:71-76 */
    }
  }
  uint64_t max_score = 0, n = 0;
  for (uint64_t i = 0; i < nf; i++) {                                 /* ablast.cc:58-73 */
    if (f[i] == 0) continue;
    if (f[i] > max_score) { max_score = f[i]; n = 0; if (n < cap) hits[n] = (uint32_t)(a_start + i); n = 1; }
    else if (f[i] == max_score) { if (n < cap) hits[n] = (uint32_t)(a_start + i); n++; }
  }
  if (max_count_out) *max_count_out = max_score;
  free(ka); free(f);
  return n;
}
