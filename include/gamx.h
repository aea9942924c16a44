/*
 * gamx.h - C ABI of the B200-native banded overlap aligner (the drop-in boundary).
 *
 * What it replaces in the reference (vice87/gam-ngs):
 *   BandedSmithWaterman::find_alignment
 *       lib/include/alignment/banded_smith_waterman.hpp:68-71   (declaration)
 *       lib/src/alignment/banded_smith_waterman.cc:69-323       (definition)
 *   and the MyAlignment reductions gam-merge reads from its result
 *       first_match_pos / last_match_pos / last_pos / gaps_before_last_match
 *       lib/src/alignment/my_alignment.cc:167-296
 *   as called from PctgBuilder::alignBlocks / findBestAlignment
 *       lib/src/pctg/PctgBuilder.cc:1544-1607, 1669, 1698.
 *
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.  The C++
 * drop-in classes (gam_ngs_b200/cpp/gamx_dropin.hpp) and the Python binding
 * (gam_ngs_b200/capi.py) are thin layers over these entry points.  INTEGRATION.md shows
 * the reference-side change a maintainer would make.
 *
 * There is no CPU fallback: every alignment is computed by the CUDA kernels in
 * gam_ngs_b200/csrc/.  If no CUDA device is usable, gamx_create() fails.
 *
 * Thread safety: a gamx_ctx serialises its own entry points with an internal mutex, so the
 * legacy one-call-per-pthread pattern (lib/src/pctg/ThreadedBuildPctg.cc:159-169) is safe;
 * use one batch per call (or one ctx per thread) for throughput.
 */
#ifndef GAMX_H_
#define GAMX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAMX_ABI_VERSION 1

#if defined(__GNUC__)
#define GAMX_API __attribute__((visibility("default")))
#else
#define GAMX_API
#endif

/* Base codes = the reference's BaseType, lib/include/assembly/nucleotide.hpp:35-43 */
enum { GAMX_BASE_A = 0, GAMX_BASE_T = 1, GAMX_BASE_C = 2, GAMX_BASE_G = 3, GAMX_BASE_N = 4 };

/* Edit operations = the reference's AlignmentAlphabet, lib/include/alignment/my_alignment.hpp:57-62 */
enum { GAMX_OP_GAP_A = 0, GAMX_OP_GAP_B = 1, GAMX_OP_MATCH = 2, GAMX_OP_MISMATCH = 3 };

/* Reference defaults: banded_smith_waterman.hpp:37-39, my_alignment.hpp:46 */
#define GAMX_DEFAULT_BAND 150
#define GAMX_DEFAULT_GAP (-8)
#define GAMX_FORCE_MAXGAP_LEN 10
#define GAMX_MAX_ALIGNMENT 500000

/* Per-job behavioural outcome, so a C++ shim can reproduce the reference exactly. */
enum {
  GAMX_JOB_OK = 0,           /* an alignment was produced                                        */
  GAMX_JOB_EMPTY = 1,        /* reference returns a default MyAlignment() (.cc:90, .cc:215)       */
  GAMX_JOB_OUT_OF_RANGE = 2, /* reference throws std::out_of_range from Contig::at (.cc:231,:265) */
  GAMX_JOB_UNDEFINED = 3     /* reference behaviour is undefined (x_size == 0, .cc:102-122)       */
};

/* What to compute for a job. */
enum {
  GAMX_MODE_SCORE = 0,     /* score + end cell only (no traceback; begin/stat fields are zero)   */
  GAMX_MODE_ENDPOINTS = 1, /* + begin_a/begin_b, n_ops, n_match, first/last match (device traceback,
                              no edit string copied back) - everything gam-merge reads            */
  GAMX_MODE_FULL = 2       /* + the full edit string (2 bits per op) and run-length CIGAR         */
};

/* Infrastructure errors (negative return values). */
enum {
  GAMX_OK = 0,
  GAMX_ERR_CUDA = -1,
  GAMX_ERR_INVALID = -2,
  GAMX_ERR_NOMEM = -3,
  GAMX_ERR_OPS_CAPACITY = -4, /* ops buffer too small; gamx_ops_capacity() tells how much */
  GAMX_ERR_NO_DEVICE = -5
};

typedef struct gamx_ctx gamx_ctx;

/*
 * One alignment = one find_alignment call (banded_smith_waterman.hpp:68-71).
 * `a` and `b` are views into contigs of the context's store:
 *   view = (rc ? reverse_complement(contig) : contig)[off, off+len)
 * which covers reverse_complement (PctgBuilder.cc:1443) and chop_begin tails
 * (PctgBuilder.cc:1577,1595) without re-uploading sequence.  len == UINT64_MAX means
 * "to the end of the contig".  begin/end are in view coordinates, exactly the
 * arguments the reference call would receive; a.size() is the view length.
 */
typedef struct {
  uint32_t a_id, b_id;
  uint8_t a_rc, b_rc;
  uint8_t force_start, force_end;
  uint8_t mode; /* GAMX_MODE_* */
  uint8_t reserved_[3];
  uint64_t a_off, a_len;
  uint64_t b_off, b_len;
  uint64_t begin_a, end_a, begin_b, end_b;
  uint32_t band; /* _band_size (ctor argument, banded_smith_waterman.cc:61-67) */
  int32_t gap;   /* _gap_score; GAMX_DEFAULT_GAP unless the 5-argument ctor is used (.cc:48-59) */
} gamx_job;

/*
 * Result of one job.  Fields mirror MyAlignment (my_alignment.hpp:65-126) plus the
 * reductions of my_alignment.cc:167-296 computed on the device.
 *   homology = n_ops ? (double)(n_match*100)/(double)n_ops : 0   (banded_smith_waterman.cc:319)
 */
typedef struct {
  int32_t status; /* GAMX_JOB_* */
  int32_t has_match;
  int64_t score;
  uint64_t begin_a, begin_b, a_size, b_size;
  uint64_t n_ops, n_match; /* length() and number of MATCH ops */
  uint64_t n_mismatch, n_gap_a, n_gap_b;
  double homology;
  uint64_t first_match_a, first_match_b; /* first_match_pos(); valid output even if !has_match */
  uint64_t last_match_a, last_match_b;   /* last_match_pos()                                  */
  uint64_t last_pos_a, last_pos_b;       /* last_pos()                                        */
  uint64_t gaps_a, gaps_b;               /* gaps_before_last_match()                          */
  int64_t end_i, end_j;                  /* selected end cell (row, band column), .cc:174-212  */
  uint64_t x_size;                       /* DP rows, .cc:93-95; cells = x_size*(2*band+1)      */
  uint64_t ops_offset;                   /* FULL mode: index (in ops) of op 0 inside ops_buf   */
} gamx_result;

/* ---- lifecycle ------------------------------------------------------------------- */

/* Creates a context driving the given CUDA devices (device_ids == NULL: devices 0..n-1;
 * n_devices == 0: all visible devices).  Jobs of a batch are sharded over the devices by
 * DP cost; results are gathered on the host; no collectives are involved. */
GAMX_API int gamx_create(gamx_ctx** out, const int* device_ids, int n_devices);
GAMX_API void gamx_destroy(gamx_ctx* ctx);
GAMX_API int gamx_device_count(const gamx_ctx* ctx);
GAMX_API const char* gamx_last_error(const gamx_ctx* ctx);
GAMX_API int gamx_abi_version(void);

/* ---- contig store (replaces the vector<Nucleotide> copies of PctgBuilder.cc:747-748) -- */

/* Adds a contig given as base codes 0..4 (values > 4 are treated as N, like
 * nucleotide.code.hpp:47-75 does for unknown characters).  At the next batch the raw codes are
 * staged to every device with pinned async copies and packed there (kernel K0) to 2 bits per
 * base plus an N bitmask.  Returns the contig id (>= 0) or a negative error. */
GAMX_API int64_t gamx_add_contig(gamx_ctx* ctx, const uint8_t* codes, uint64_t len);
/* Same, from FASTA characters (ACGTacgt, everything else -> N). */
GAMX_API int64_t gamx_add_contig_ascii(gamx_ctx* ctx, const char* seq, uint64_t len);
/* Loads every record of a FASTA file as a contig (the record structure and the character table of the
 * reference's reader: io_contig.code.hpp:540-565 readNextSequence, nucleotide.code.hpp:47-75).  Returns the id
 * of the first contig (the records get consecutive ids in file order) or a negative error; *n_contigs (may be
 * NULL) receives the record count.  gamx_contig_name returns the first word of a record's header (a per-thread
 * copy, valid until the thread's next call; "" for contigs that were not loaded from FASTA). */
GAMX_API int64_t gamx_add_fasta(gamx_ctx* ctx, const char* path, uint64_t* n_contigs);
GAMX_API const char* gamx_contig_name(const gamx_ctx* ctx, uint32_t id);
/* Bulk form: n contigs whose codes are concatenated in `codes` (lengths[n] bases each).  The raw
 * bytes are copied to every device (directly when `codes` is pinned host memory, through pinned
 * staging otherwise) and packed there by a kernel; the call returns when `codes` may be reused.
 * Returns the id of the first contig (ids are consecutive) or a negative error. */
GAMX_API int64_t gamx_add_contigs(gamx_ctx* ctx, const uint8_t* codes, const uint64_t* lengths, uint64_t n);
/* Same, but only enqueues the copies and the pack kernel on the devices' streams and returns: the
 * upload then overlaps the host-side planning of the next gamx_align_batch, which is stream-ordered
 * behind it.  The copy proceeds in pieces of ~128 MB (shorter ones first and last), all enqueued by this
 * call, each followed by its pack launch on a second stream; a pipelined gamx_align_batch waits per chunk for
 * the pieces its jobs refer to, so it computes on the first contigs while later ones still cross PCIe.
 * (A context drives ~20 streams per device: gamx_create sets CUDA_DEVICE_MAX_CONNECTIONS=32 if it is unset,
 * which takes effect only if the process has not created its CUDA context yet - see INTEGRATION.md 3a.)
 * `codes` must be PINNED host memory and must stay valid and unchanged until the next gamx_align_batch,
 * gamx_align_batch_cigar or gamx_plan_create call on this context has returned: those calls finish the upload
 * before they return on every path - errors, empty batches and batches that never refer to the last contigs
 * included (gamx_add_contigs*, gamx_clear_contigs and gamx_destroy do so too). */
GAMX_API int64_t gamx_add_contigs_async(gamx_ctx* ctx, const uint8_t* codes, const uint64_t* lengths, uint64_t n);
GAMX_API uint64_t gamx_contig_length(const gamx_ctx* ctx, uint32_t id);
GAMX_API int gamx_clear_contigs(gamx_ctx* ctx);

/* ---- alignment ---------------------------------------------------------------------- */

/* Upper bound on the number of ops the batch can emit in FULL mode (0 for other modes). */
GAMX_API uint64_t gamx_ops_capacity(const gamx_ctx* ctx, const gamx_job* jobs, uint64_t n);

/* Aligns a batch.  On GAMX_OK every results[i] is filled; on a negative return the records are undefined
 * (on GAMX_ERR_INVALID and on errors of the pipelined path they are left untouched).  ops_buf (may be NULL when no job is in FULL
 * mode) receives the edit strings packed 2 bits per op, op k of a job at bits
 * [2*((ops_offset+k)%4), +2) of byte (ops_offset+k)/4; ops_cap is its capacity in ops. */
GAMX_API int gamx_align_batch(gamx_ctx* ctx, const gamx_job* jobs, uint64_t n, gamx_result* results,
                     uint8_t* ops_buf, uint64_t ops_cap);

/* The same batch with the edit strings returned as run-length CIGARs that are built ON THE DEVICE (the packed
 * ops never cross PCIe): job i's runs are runs[run_offsets[i] .. run_offsets[i+1]) (run_offsets has n + 1
 * entries), each run = length << 2 | op (GAMX_OP_*), in edit-string order - what gamx_cigar_rle produces from the
 * packed ops of gamx_align_batch.  Jobs that are not in FULL mode, or return no alignment, have no runs.
 * *runs_needed (may be NULL) receives the total; GAMX_ERR_OPS_CAPACITY when runs_cap is too small (results and
 * runs are then undefined; call again with a buffer of *runs_needed entries). */
GAMX_API int gamx_align_batch_cigar(gamx_ctx* ctx, const gamx_job* jobs, uint64_t n, gamx_result* results,
                                    uint64_t* run_offsets, uint32_t* runs, uint64_t runs_cap, uint64_t* runs_needed);

/* Expands n_ops packed ops starting at ops_offset to one byte per op (GAMX_OP_*), the layout
 * of the reference's std::vector<AlignmentAlphabet>. */
GAMX_API void gamx_unpack_ops(const uint8_t* ops_buf, uint64_t ops_offset, uint64_t n_ops, uint8_t* out);

/* Run-length CIGAR of a packed edit string: writes up to cap (op,len) pairs
 * (op in the low 2 bits, length in the upper 30 of each uint32), returns the number of runs. */
GAMX_API uint64_t gamx_cigar_rle(const uint8_t* ops_buf, uint64_t ops_offset, uint64_t n_ops,
                        uint32_t* runs, uint64_t cap);

/* gamx_align_batch pipelines large batches that return no edit strings (no FULL-mode job): the batch is
 * cut into chunks of `jobs_per_chunk` jobs in the caller's order; while chunk c computes, chunk c+1 is
 * prepared and uploaded and chunk c-1 is read back, and every chunk waits only for the contig upload
 * pieces (gamx_add_contigs_async) its own jobs refer to.  Batches shorter than two chunks take the
 * single-launch path.  0 disables pipelining.  Default 65536 (environment GAMX_PIPELINE_CHUNK); the
 * first two chunks are a quarter and a half of that so that the device starts early. */
GAMX_API int gamx_set_pipeline_chunk(gamx_ctx* ctx, uint64_t jobs_per_chunk);

/* ---- device-resident (pre-staged) batches: the kernel-only timing path -------------- */

typedef struct gamx_plan gamx_plan;
/* Validates, sorts, shards and uploads a batch once; gamx_plan_run() then only launches the
 * kernels (inputs already resident in HBM) and gamx_plan_fetch() copies the results back.
 * A plan owns no device memory of its own: its descriptors, results and scratch live in the context's
 * buffers, and the next gamx_align_batch*, gamx_merge_align or gamx_plan_create on the context takes them
 * over.  gamx_plan_run / gamx_plan_fetch on a plan that has lost its buffers return GAMX_ERR_INVALID (they
 * never run on, or return, another batch's data): keep one live plan per context. */
GAMX_API int gamx_plan_create(gamx_ctx* ctx, const gamx_job* jobs, uint64_t n, gamx_plan** out);
GAMX_API int gamx_plan_run(gamx_plan* plan);    /* asynchronous; enqueues on each device's stream */
GAMX_API int gamx_plan_sync(gamx_plan* plan);   /* waits for all devices */
GAMX_API int gamx_plan_fetch(gamx_plan* plan, gamx_result* results, uint8_t* ops_buf, uint64_t ops_cap);
/* Device time of the last run in milliseconds (CUDA events on the launching streams, max
 * over devices); kernel_ms[i] receives per-kernel-family times when not NULL (see
 * gamx_plan_kernel_names). */
GAMX_API float gamx_plan_last_ms(gamx_plan* plan);
GAMX_API uint64_t gamx_plan_cells(const gamx_plan* plan);        /* sum of x_size*(2*band+1)           */
GAMX_API uint64_t gamx_plan_kernel_launches(const gamx_plan* plan); /* kernels launched per run        */
GAMX_API void gamx_plan_destroy(gamx_plan* plan);

/* ---- seed finder: ABlast::findHits (lib/src/alignment/ablast.cc:41-76) ----------------- */

/* One findHits(a, a_start, a_end, b, b_start, b_end) call on views of stored contigs: k-mer
 * (w = 20) diagonal voting, only diagonals idx_a >= idx_b (ablast.hpp:71-76), code aliasing of N
 * as in ablast.hpp:53-59.  The reference returns every diagonal with the maximal count in
 * ascending order and its callers use front() or back() (PctgBuilder.cc:1544,1560,1584,1602): the
 * result carries both plus the list length. */
typedef struct {
  uint32_t a_id, b_id;
  uint8_t a_rc, b_rc;
  uint8_t reserved_[6];
  uint64_t a_off, a_len, b_off, b_len; /* views, as in gamx_job */
  uint64_t a_start, a_end, b_start, b_end;
} gamx_hits_job;

typedef struct {
  uint32_t n_hits;     /* hits.size() */
  uint32_t max_count;  /* votes of the best diagonal(s) */
  uint64_t first_hit;  /* hits.front(): a_start + smallest best diagonal (32-bit truncated like the reference) */
  uint64_t last_hit;   /* hits.back() */
} gamx_hits_result;

GAMX_API int gamx_find_hits_batch(gamx_ctx* ctx, const gamx_hits_job* jobs, uint64_t n, gamx_hits_result* results);

/* ---- gam-merge alignment stage as batched rounds (the "batch collector") ---------------- */

/* One block of a merge block: the two frames (0-based inclusive contig coordinates, Frame.hpp:53-60)
 * and the read count (Block.hpp:68-71).  strand: 0 '+', 1 '-'. */
typedef struct {
  int32_t num_reads;
  uint8_t m_strand, s_strand;
  uint8_t reserved_[2];
  int32_t m_begin, m_end, s_begin, s_end;
} gamx_block;

/* One merge block = one CompactAssemblyGraph vertex (MergeDescriptor.hpp:40-69): a master contig,
 * a slave contig (ids in the context's store), its blocks [first_block, first_block + n_blocks) and
 * the tail flags the graph code sets. */
typedef struct {
  uint32_t m_id, s_id;
  uint32_t first_block, n_blocks;
  uint8_t m_ltail, m_rtail, s_ltail, s_rtail;
} gamx_merge_block;

/* What PctgBuilder::alignMergeBlock (PctgBuilder.cc:726-844) writes into the MergeBlock. */
typedef struct {
  int32_t status;      /* 0 ok; 2: the reference would throw (std::out_of_range / std::domain_error) and
                          drop the graph (ThreadedBuildPctg.cc:322-329) */
  int32_t align_ok, align_rev;
  int32_t coords_set;  /* 0 when alignMergeBlock returns before assigning the coordinates (.cc:825-829) */
  int32_t m_start, m_end, s_start, s_end;
  uint32_t n_alignments, n_hits_calls; /* find_alignment / findHits calls this merge block needed */
} gamx_merge_result;

typedef struct {
  uint64_t rounds, alignments, hits_calls, cells; /* GPU batches issued, jobs in them, DP cells */
} gamx_merge_stats;

/* Runs the alignment stage of gam-merge for n merge blocks: per merge block the chained block
 * alignments in the orientation the strand evidence suggests, the retry in the other orientation,
 * then up to two tail alignments seeded by findHits - as rounds of GPU batches over all merge
 * blocks.  Band 150 and gap -8 as at every reference call site (PctgBuilder.cc:1410,1628). */
GAMX_API int gamx_merge_align(gamx_ctx* ctx, const gamx_merge_block* mbs, uint64_t n, const gamx_block* blocks,
                              uint64_t n_blocks, gamx_merge_result* results, gamx_merge_stats* stats);

/* ---- sharding ------------------------------------------------------------------------ */

/* The cost-balanced split gamx_align_batch applies over a context's devices (longest-processing-
 * time greedy on DP cells), exposed so callers that run one process per GPU can shard a batch the
 * same way: shard_out[i] in [0, n_shards) for job i of cost cost[i].  Pure host code. */
GAMX_API int gamx_shard_by_cost(const uint64_t* cost, uint64_t n, int n_shards, int32_t* shard_out);

/* Which fill kernel a band width maps to: *stripe_width = C (band columns per lane) and
 * *lanes_per_pair = LG (8, 16, 32: warp-level kernel K1, 32/LG pairs per warp; 64, 128, 256: CTA-per-pair
 * kernel K2).  Returns 0, or GAMX_ERR_INVALID when the band needs the generic kernel.  Pure host code. */
GAMX_API int gamx_band_geometry(uint64_t band, int* stripe_width, int* lanes_per_pair);

/* ---- diagnostics ------------------------------------------------------------------------- */

/* Exercises the host worker pool the batch preparation runs on (pure host code, no device): `callers`
 * threads submit parallel passes over `items` elements at the same time - the way the producer and the
 * consumer of a pipelined batch, or several contexts, do - and check every pass's result.  Returns 0, or
 * the number of passes that came out wrong. */
GAMX_API int gamx_host_selftest(int callers, uint64_t items);

/* ---- microbenchmarks used for the roofline denominators (bench.py) ------------------ */

/* Measures the integer/DPX issue peak of device `dev_index` of the context with a
 * register-only kernel: which = 0 VIADDMNMX(s32), 1 VIMNMX3(s32), 2 VIADDMNMX(s16x2),
 * 3 LOP3, 4 PRMT, 5 IMAD.  Returns lane-ops per second (1 op = 1 instruction lane). */
GAMX_API double gamx_measure_int_peak(gamx_ctx* ctx, int dev_index, int which);

#ifdef __cplusplus
}
#endif
#endif /* GAMX_H_ */
