"""ctypes binding of the C ABI in include/gamx.h (libgamx.so).

This is plumbing only: every alignment is computed by the CUDA kernels behind the C ABI.
There is no CPU fallback; creating a Context without a usable CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

MODE_SCORE, MODE_ENDPOINTS, MODE_FULL = 0, 1, 2
JOB_OK, JOB_EMPTY, JOB_OUT_OF_RANGE, JOB_UNDEFINED = 0, 1, 2, 3
DEFAULT_BAND, DEFAULT_GAP = 150, -8
U64_MAX = 2**64 - 1
ERR_OPS_CAPACITY = -4


class GamxJob(C.Structure):
    _fields_ = [
        ("a_id", C.c_uint32), ("b_id", C.c_uint32),
        ("a_rc", C.c_uint8), ("b_rc", C.c_uint8),
        ("force_start", C.c_uint8), ("force_end", C.c_uint8),
        ("mode", C.c_uint8), ("reserved_", C.c_uint8 * 3),
        ("a_off", C.c_uint64), ("a_len", C.c_uint64),
        ("b_off", C.c_uint64), ("b_len", C.c_uint64),
        ("begin_a", C.c_uint64), ("end_a", C.c_uint64),
        ("begin_b", C.c_uint64), ("end_b", C.c_uint64),
        ("band", C.c_uint32), ("gap", C.c_int32),
    ]


class GamxResult(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("has_match", C.c_int32), ("score", C.c_int64),
        ("begin_a", C.c_uint64), ("begin_b", C.c_uint64), ("a_size", C.c_uint64), ("b_size", C.c_uint64),
        ("n_ops", C.c_uint64), ("n_match", C.c_uint64),
        ("n_mismatch", C.c_uint64), ("n_gap_a", C.c_uint64), ("n_gap_b", C.c_uint64),
        ("homology", C.c_double),
        ("first_match_a", C.c_uint64), ("first_match_b", C.c_uint64),
        ("last_match_a", C.c_uint64), ("last_match_b", C.c_uint64),
        ("last_pos_a", C.c_uint64), ("last_pos_b", C.c_uint64),
        ("gaps_a", C.c_uint64), ("gaps_b", C.c_uint64),
        ("end_i", C.c_int64), ("end_j", C.c_int64),
        ("x_size", C.c_uint64), ("ops_offset", C.c_uint64),
    ]


JOB_DTYPE = np.dtype([
    ("a_id", "<u4"), ("b_id", "<u4"), ("a_rc", "u1"), ("b_rc", "u1"), ("force_start", "u1"),
    ("force_end", "u1"), ("mode", "u1"), ("reserved_", "u1", (3,)),
    ("a_off", "<u8"), ("a_len", "<u8"), ("b_off", "<u8"), ("b_len", "<u8"),
    ("begin_a", "<u8"), ("end_a", "<u8"), ("begin_b", "<u8"), ("end_b", "<u8"),
    ("band", "<u4"), ("gap", "<i4")], align=True)

HITS_JOB_DTYPE = np.dtype([
    ("a_id", "<u4"), ("b_id", "<u4"), ("a_rc", "u1"), ("b_rc", "u1"), ("reserved_", "u1", (6,)),
    ("a_off", "<u8"), ("a_len", "<u8"), ("b_off", "<u8"), ("b_len", "<u8"),
    ("a_start", "<u8"), ("a_end", "<u8"), ("b_start", "<u8"), ("b_end", "<u8")], align=True)
HITS_RESULT_DTYPE = np.dtype([("n_hits", "<u4"), ("max_count", "<u4"), ("first_hit", "<u8"), ("last_hit", "<u8")],
                             align=True)

BLOCK_DTYPE = np.dtype([("num_reads", "<i4"), ("m_strand", "u1"), ("s_strand", "u1"), ("reserved_", "u1", (2,)),
                        ("m_begin", "<i4"), ("m_end", "<i4"), ("s_begin", "<i4"), ("s_end", "<i4")], align=True)
MERGE_BLOCK_DTYPE = np.dtype([("m_id", "<u4"), ("s_id", "<u4"), ("first_block", "<u4"), ("n_blocks", "<u4"),
                              ("m_ltail", "u1"), ("m_rtail", "u1"), ("s_ltail", "u1"), ("s_rtail", "u1")], align=True)
MERGE_RESULT_DTYPE = np.dtype([("status", "<i4"), ("align_ok", "<i4"), ("align_rev", "<i4"), ("coords_set", "<i4"),
                               ("m_start", "<i4"), ("m_end", "<i4"), ("s_start", "<i4"), ("s_end", "<i4"),
                               ("n_alignments", "<u4"), ("n_hits_calls", "<u4")], align=True)
MERGE_STATS_DTYPE = np.dtype([("rounds", "<u8"), ("alignments", "<u8"), ("hits_calls", "<u8"), ("cells", "<u8")])

RESULT_DTYPE = np.dtype([
    ("status", "<i4"), ("has_match", "<i4"), ("score", "<i8"),
    ("begin_a", "<u8"), ("begin_b", "<u8"), ("a_size", "<u8"), ("b_size", "<u8"),
    ("n_ops", "<u8"), ("n_match", "<u8"), ("n_mismatch", "<u8"), ("n_gap_a", "<u8"), ("n_gap_b", "<u8"),
    ("homology", "<f8"),
    ("first_match_a", "<u8"), ("first_match_b", "<u8"), ("last_match_a", "<u8"), ("last_match_b", "<u8"),
    ("last_pos_a", "<u8"), ("last_pos_b", "<u8"), ("gaps_a", "<u8"), ("gaps_b", "<u8"),
    ("end_i", "<i8"), ("end_j", "<i8"), ("x_size", "<u8"), ("ops_offset", "<u8")], align=True)

assert JOB_DTYPE.itemsize == C.sizeof(GamxJob), (JOB_DTYPE.itemsize, C.sizeof(GamxJob))
assert RESULT_DTYPE.itemsize == C.sizeof(GamxResult), (RESULT_DTYPE.itemsize, C.sizeof(GamxResult))

EXPORTS = [
    "gamx_abi_version", "gamx_create", "gamx_destroy", "gamx_device_count", "gamx_last_error",
    "gamx_add_contig", "gamx_add_contig_ascii", "gamx_add_contigs", "gamx_add_contigs_async", "gamx_contig_length", "gamx_clear_contigs",
    "gamx_add_fasta", "gamx_contig_name",
    "gamx_ops_capacity", "gamx_set_pipeline_chunk", "gamx_align_batch", "gamx_align_batch_cigar", "gamx_unpack_ops", "gamx_cigar_rle",
    "gamx_plan_create", "gamx_plan_run", "gamx_plan_sync", "gamx_plan_fetch", "gamx_plan_last_ms",
    "gamx_plan_cells", "gamx_plan_kernel_launches", "gamx_plan_destroy", "gamx_measure_int_peak",
    "gamx_shard_by_cost", "gamx_find_hits_batch", "gamx_merge_align", "gamx_band_geometry",
    "gamx_host_selftest",
]

_lib = None


class GamxError(RuntimeError):
    pass


def load_library(build_if_missing: bool = True):
    """Loads gam_ngs_b200/libgamx.so (building it in-tree with nvcc when absent)."""
    global _lib
    if _lib is not None:
        return _lib
    lib_path = os.environ.get("GAMX_LIB", _build.LIB)  # GAMX_LIB: alternative build, experiments only
    if lib_path == _build.LIB and build_if_missing and _build.needs_build():
        _build.build()
    if not os.path.exists(lib_path):
        raise GamxError("libgamx.so is missing: run `python -m gam_ngs_b200.build`")
    L = C.CDLL(lib_path)
    vp, u64, u8p = C.c_void_p, C.c_uint64, C.POINTER(C.c_uint8)
    L.gamx_abi_version.restype = C.c_int
    L.gamx_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), C.c_int]
    L.gamx_create.restype = C.c_int
    L.gamx_destroy.argtypes = [vp]
    L.gamx_destroy.restype = None
    L.gamx_device_count.argtypes = [vp]
    L.gamx_device_count.restype = C.c_int
    L.gamx_last_error.argtypes = [vp]
    L.gamx_last_error.restype = C.c_char_p
    L.gamx_add_contig.argtypes = [vp, u8p, u64]
    L.gamx_add_contig.restype = C.c_int64
    L.gamx_add_contig_ascii.argtypes = [vp, C.c_char_p, u64]
    L.gamx_add_contig_ascii.restype = C.c_int64
    L.gamx_add_contigs.argtypes = [vp, vp, vp, u64]
    L.gamx_add_contigs.restype = C.c_int64
    L.gamx_add_contigs_async.argtypes = [vp, vp, vp, u64]
    L.gamx_add_contigs_async.restype = C.c_int64
    L.gamx_contig_length.argtypes = [vp, C.c_uint32]
    L.gamx_contig_length.restype = u64
    L.gamx_clear_contigs.argtypes = [vp]
    L.gamx_clear_contigs.restype = C.c_int
    L.gamx_ops_capacity.argtypes = [vp, vp, u64]
    L.gamx_ops_capacity.restype = u64
    L.gamx_set_pipeline_chunk.argtypes = [vp, u64]
    L.gamx_set_pipeline_chunk.restype = C.c_int
    L.gamx_align_batch.argtypes = [vp, vp, u64, vp, vp, u64]
    L.gamx_align_batch.restype = C.c_int
    L.gamx_align_batch_cigar.argtypes = [vp, vp, u64, vp, vp, vp, u64, vp]
    L.gamx_align_batch_cigar.restype = C.c_int
    L.gamx_add_fasta.argtypes = [vp, C.c_char_p, vp]
    L.gamx_add_fasta.restype = C.c_int64
    L.gamx_contig_name.argtypes = [vp, C.c_uint32]
    L.gamx_contig_name.restype = C.c_char_p
    L.gamx_unpack_ops.argtypes = [vp, u64, u64, vp]
    L.gamx_unpack_ops.restype = None
    L.gamx_cigar_rle.argtypes = [vp, u64, u64, vp, u64]
    L.gamx_cigar_rle.restype = u64
    L.gamx_plan_create.argtypes = [vp, vp, u64, C.POINTER(vp)]
    L.gamx_plan_create.restype = C.c_int
    for f in ("gamx_plan_run", "gamx_plan_sync"):
        getattr(L, f).argtypes = [vp]
        getattr(L, f).restype = C.c_int
    L.gamx_plan_fetch.argtypes = [vp, vp, vp, u64]
    L.gamx_plan_fetch.restype = C.c_int
    L.gamx_plan_last_ms.argtypes = [vp]
    L.gamx_plan_last_ms.restype = C.c_float
    L.gamx_plan_cells.argtypes = [vp]
    L.gamx_plan_cells.restype = u64
    L.gamx_plan_kernel_launches.argtypes = [vp]
    L.gamx_plan_kernel_launches.restype = u64
    L.gamx_plan_destroy.argtypes = [vp]
    L.gamx_plan_destroy.restype = None
    L.gamx_find_hits_batch.argtypes = [vp, vp, u64, vp]
    L.gamx_find_hits_batch.restype = C.c_int
    L.gamx_merge_align.argtypes = [vp, vp, u64, vp, u64, vp, vp]
    L.gamx_merge_align.restype = C.c_int
    L.gamx_band_geometry.argtypes = [u64, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.gamx_band_geometry.restype = C.c_int
    L.gamx_host_selftest.argtypes = [C.c_int, u64]
    L.gamx_host_selftest.restype = C.c_int
    L.gamx_measure_int_peak.argtypes = [vp, C.c_int, C.c_int]
    L.gamx_measure_int_peak.restype = C.c_double
    _lib = L
    return L


def band_geometry(band: int):
    """(stripe width C, lanes per pair LG) of the fill kernel a band width maps to."""
    c, lg = C.c_int(0), C.c_int(0)
    if load_library().gamx_band_geometry(int(band), C.byref(c), C.byref(lg)) != 0:
        return None
    return c.value, lg.value


def shard_by_cost(cost, n_shards: int) -> np.ndarray:
    """The cost-balanced split the library applies over a context's devices (longest processing time first),
    for callers that run one process per GPU: shard index per item.  Pure host code."""
    cost = np.ascontiguousarray(cost, dtype=np.uint64)
    out = np.zeros(len(cost), dtype=np.int32)
    L = load_library()
    L.gamx_shard_by_cost.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    L.gamx_shard_by_cost.restype = C.c_int
    if L.gamx_shard_by_cost(cost.ctypes.data, len(cost), int(n_shards), out.ctypes.data) != 0:
        raise GamxError("gamx_shard_by_cost failed")
    return out


def make_hits_jobs(n: int) -> np.ndarray:
    jobs = np.zeros(n, dtype=HITS_JOB_DTYPE)
    jobs["a_len"] = U64_MAX
    jobs["b_len"] = U64_MAX
    return jobs


def make_jobs(n: int) -> np.ndarray:
    """Zeroed job array with the reference's defaults (band 150, gap -8, whole contigs)."""
    jobs = np.zeros(n, dtype=JOB_DTYPE)
    jobs["a_len"] = U64_MAX
    jobs["b_len"] = U64_MAX
    jobs["band"] = DEFAULT_BAND
    jobs["gap"] = DEFAULT_GAP
    jobs["mode"] = MODE_ENDPOINTS
    return jobs


class Plan:
    """A validated, sharded batch whose descriptors are resident on the devices.  The plan lives in the context's
    buffers: any later batch, merge round or plan on the same context takes them over, after which run() / fetch()
    raise (GAMX_ERR_INVALID) instead of touching another batch's data.  Keep one live plan per context."""

    def __init__(self, ctx: "Context", jobs: np.ndarray):
        self.ctx = ctx
        self.n = len(jobs)
        self.jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        self._h = C.c_void_p()
        ctx._check(ctx.lib.gamx_plan_create(ctx._h, self.jobs.ctypes.data, self.n, C.byref(self._h)))
        self.ops_capacity = int(ctx.lib.gamx_ops_capacity(ctx._h, self.jobs.ctypes.data, self.n))

    def run(self):
        self.ctx._check(self.ctx.lib.gamx_plan_run(self._h))

    def sync(self):
        self.ctx._check(self.ctx.lib.gamx_plan_sync(self._h))

    def fetch(self):
        results = np.zeros(self.n, dtype=RESULT_DTYPE)
        ops = np.zeros((self.ops_capacity + 3) // 4 + 8, dtype=np.uint8)
        self.ctx._check(self.ctx.lib.gamx_plan_fetch(self._h, results.ctypes.data, ops.ctypes.data,
                                                     self.ops_capacity))
        return results, ops

    @property
    def last_ms(self) -> float:
        return float(self.ctx.lib.gamx_plan_last_ms(self._h))

    @property
    def cells(self) -> int:
        return int(self.ctx.lib.gamx_plan_cells(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(self.ctx.lib.gamx_plan_kernel_launches(self._h))

    def close(self):
        if self._h:
            self.ctx.lib.gamx_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    def __init__(self, devices=None):
        self.lib = load_library()
        self._h = C.c_void_p()
        if devices is None:
            rc = self.lib.gamx_create(C.byref(self._h), None, 0)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.gamx_create(C.byref(self._h), arr, len(devices))
        if rc != 0:
            raise GamxError(f"gamx_create failed with {rc}: no usable CUDA device "
                            "(this package has no CPU fallback)")

    def _check(self, rc: int):
        if rc != 0:
            raise GamxError(f"gamx error {rc}: {self.lib.gamx_last_error(self._h).decode()}")

    @property
    def device_count(self) -> int:
        return self.lib.gamx_device_count(self._h)

    def add_contig(self, codes) -> int:
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        cid = self.lib.gamx_add_contig(self._h, codes.ctypes.data_as(C.POINTER(C.c_uint8)), len(codes))
        if cid < 0:
            self._check(int(cid))
        return int(cid)

    def add_contig_ascii(self, seq: str | bytes) -> int:
        if isinstance(seq, str):
            seq = seq.encode()
        cid = self.lib.gamx_add_contig_ascii(self._h, seq, len(seq))
        if cid < 0:
            self._check(int(cid))
        return int(cid)

    def add_contigs(self, codes, lengths, async_upload: bool = False) -> int:
        """Bulk upload: `codes` = concatenated base codes (numpy uint8 array or a raw pointer to
        pinned host memory), `lengths` = bases per contig.  Returns the first contig id.
        async_upload=True (pinned pointer only) lets the copy overlap the planning of the next batch."""
        lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
        ptr = codes if isinstance(codes, int) else np.ascontiguousarray(codes, dtype=np.uint8).ctypes.data
        fn = self.lib.gamx_add_contigs_async if (async_upload and isinstance(codes, int)) else self.lib.gamx_add_contigs
        first = fn(self._h, ptr, lengths.ctypes.data, len(lengths))
        if first < 0:
            self._check(int(first))
        return int(first)

    def clear_contigs(self):
        self._check(self.lib.gamx_clear_contigs(self._h))

    def contig_length(self, cid: int) -> int:
        return int(self.lib.gamx_contig_length(self._h, cid))

    def ops_capacity(self, jobs: np.ndarray) -> int:
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        return int(self.lib.gamx_ops_capacity(self._h, jobs.ctypes.data, len(jobs)))

    def set_pipeline_chunk(self, jobs_per_chunk: int):
        """Chunk size of the pipelined gamx_align_batch (0 disables pipelining)."""
        self._check(self.lib.gamx_set_pipeline_chunk(self._h, int(jobs_per_chunk)))

    def align_batch(self, jobs: np.ndarray, out: np.ndarray | None = None):
        """Host buffers in, host buffers out (the end-to-end path): returns (results, ops).
        out: optional caller-owned RESULT_DTYPE array of len(jobs) records to write into."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        n = len(jobs)
        if out is not None:
            assert out.dtype == RESULT_DTYPE and len(out) == n and out.flags["C_CONTIGUOUS"]
        results = out if out is not None else np.empty(n, dtype=RESULT_DTYPE)  # every record is written by the library
        # large batches are first tried without an ops buffer (no scan of the modes on the Python side);
        # the library answers ERR_OPS_CAPACITY when some job wants its edit string
        cap = 0
        if n < 100000 and n and jobs["mode"].max() == MODE_FULL:
            cap = self.ops_capacity(jobs)
        ops = np.zeros((cap + 3) // 4 + 8, dtype=np.uint8)
        rc = self.lib.gamx_align_batch(self._h, jobs.ctypes.data, n, results.ctypes.data, ops.ctypes.data, cap)
        if rc == ERR_OPS_CAPACITY and cap == 0:
            cap = self.ops_capacity(jobs)
            ops = np.zeros((cap + 3) // 4 + 8, dtype=np.uint8)
            rc = self.lib.gamx_align_batch(self._h, jobs.ctypes.data, n, results.ctypes.data, ops.ctypes.data, cap)
        self._check(rc)
        return results, ops

    def align_batch_cigar(self, jobs: np.ndarray):
        """FULL-mode batch with the edit strings returned as run-length CIGARs built on the device:
        (results, run_offsets[n + 1], runs) with runs[k] = length << 2 | op."""
        jobs = np.ascontiguousarray(jobs, dtype=JOB_DTYPE)
        n = len(jobs)
        results = np.empty(n, dtype=RESULT_DTYPE)
        offs = np.zeros(n + 1, dtype=np.uint64)
        need = np.zeros(1, dtype=np.uint64)
        runs = np.zeros(max(1024, 64 * n), dtype=np.uint32)
        rc = self.lib.gamx_align_batch_cigar(self._h, jobs.ctypes.data, n, results.ctypes.data, offs.ctypes.data,
                                             runs.ctypes.data, len(runs), need.ctypes.data)
        if rc == ERR_OPS_CAPACITY:
            runs = np.zeros(int(need[0]), dtype=np.uint32)
            rc = self.lib.gamx_align_batch_cigar(self._h, jobs.ctypes.data, n, results.ctypes.data, offs.ctypes.data,
                                                 runs.ctypes.data, len(runs), need.ctypes.data)
        self._check(rc)
        return results, offs, runs[: int(offs[n])]

    def add_fasta(self, path: str):
        """Loads every record of a FASTA file; returns (first contig id, number of contigs)."""
        n = np.zeros(1, dtype=np.uint64)
        first = int(self.lib.gamx_add_fasta(self._h, path.encode(), n.ctypes.data))
        self._check(first if first < 0 else 0)
        return first, int(n[0])

    def contig_name(self, cid: int) -> str:
        return self.lib.gamx_contig_name(self._h, cid).decode()

    def find_hits_batch(self, jobs: np.ndarray) -> np.ndarray:
        """ABlast::findHits for a batch of (a window, b window) jobs (HITS_JOB_DTYPE)."""
        jobs = np.ascontiguousarray(jobs, dtype=HITS_JOB_DTYPE)
        res = np.zeros(len(jobs), dtype=HITS_RESULT_DTYPE)
        self._check(self.lib.gamx_find_hits_batch(self._h, jobs.ctypes.data, len(jobs), res.ctypes.data))
        return res

    def merge_align(self, merge_blocks: np.ndarray, blocks: np.ndarray):
        """gam-merge's alignment stage (PctgBuilder::alignMergeBlock) for a set of merge blocks, as
        rounds of GPU batches.  Returns (results[MERGE_RESULT_DTYPE], stats dict)."""
        mbs = np.ascontiguousarray(merge_blocks, dtype=MERGE_BLOCK_DTYPE)
        blk = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
        res = np.zeros(len(mbs), dtype=MERGE_RESULT_DTYPE)
        stats = np.zeros(1, dtype=MERGE_STATS_DTYPE)
        self._check(self.lib.gamx_merge_align(self._h, mbs.ctypes.data, len(mbs), blk.ctypes.data, len(blk),
                                              res.ctypes.data, stats.ctypes.data))
        return res, {k: int(stats[0][k]) for k in stats.dtype.names}

    def plan(self, jobs: np.ndarray) -> Plan:
        return Plan(self, jobs)

    def unpack_ops(self, ops: np.ndarray, offset: int, n_ops: int) -> np.ndarray:
        out = np.zeros(n_ops, dtype=np.uint8)
        self.lib.gamx_unpack_ops(ops.ctypes.data, offset, n_ops, out.ctypes.data)
        return out

    def cigar_rle(self, ops: np.ndarray, offset: int, n_ops: int):
        runs = np.zeros(max(n_ops, 1), dtype=np.uint32)
        n = self.lib.gamx_cigar_rle(ops.ctypes.data, offset, n_ops, runs.ctypes.data, len(runs))
        runs = runs[:n]
        return [(int(r & 3), int(r >> 2)) for r in runs]

    def measure_int_peak(self, which: int = 0, dev_index: int = 0) -> float:
        return float(self.lib.gamx_measure_int_peak(self._h, dev_index, which))

    def close(self):
        if self._h:
            self.lib.gamx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
