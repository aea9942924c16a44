"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/gamx.h declares; without a CUDA device the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import gam_ngs_b200
from gam_ngs_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gamx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gamx_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/gamx.h but not exported"
    assert sorted(capi.EXPORTS) == declared
    assert lib.gamx_abi_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(capi.GamxJob) == capi.JOB_DTYPE.itemsize == 88
    assert C.sizeof(capi.GamxResult) == capi.RESULT_DTYPE.itemsize == 192
    for name, _ in capi.GamxJob._fields_:
        if name != "reserved_":
            assert getattr(capi.GamxJob, name).offset == capi.JOB_DTYPE.fields[name][1]
    for name, _ in capi.GamxResult._fields_:
        assert getattr(capi.GamxResult, name).offset == capi.RESULT_DTYPE.fields[name][1]


def test_host_side_helpers_without_gpu():
    """unpack / CIGAR helpers are pure host code in the C ABI."""
    lib = capi.load_library()
    ops = np.array([2, 2, 3, 0, 0, 1, 2, 2, 2], dtype=np.uint8)
    packed = np.zeros(8, dtype=np.uint8)
    off = 5
    for k, op in enumerate(ops):
        g = off + k
        packed[g >> 2] |= op << (2 * (g & 3))
    out = np.zeros(len(ops), dtype=np.uint8)
    lib.gamx_unpack_ops(packed.ctypes.data, off, len(ops), out.ctypes.data)
    assert (out == ops).all()
    runs = np.zeros(16, dtype=np.uint32)
    n = lib.gamx_cigar_rle(packed.ctypes.data, off, len(ops), runs.ctypes.data, 16)
    assert [(int(r & 3), int(r >> 2)) for r in runs[:n]] == [(2, 2), (3, 1), (0, 2), (1, 1), (2, 3)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(gam_ngs_b200.GamxError):
        gam_ngs_b200.Context()


def _build_dropin(tmp_path):
    import subprocess
    exe = str(tmp_path / "test_dropin")
    libdir = os.path.join(ROOT, "gam_ngs_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_dropin.cc"),
                    "-L" + libdir, "-lgamx", "-pthread", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_cpp_dropin_compiles_links_and_fails_loudly_without_gpu(tmp_path):
    """The C++ drop-in (same signatures as banded_smith_waterman.hpp:41-72) builds against
    libgamx.so; without a CUDA device it reports so (exit 77), it never computes on the CPU."""
    import subprocess
    import torch
    capi.load_library()
    exe = _build_dropin(tmp_path)
    rc = subprocess.run([exe], capture_output=True).returncode
    assert rc == (0 if torch.cuda.is_available() else 77)


def test_band_geometry_is_pure_host_code():
    """gamx_band_geometry needs no device: the kernel family a band width maps to (stripe width C,
    lanes per pair LG).  Even stripe widths instead of 13/17, fewest lanes per pair on ties."""
    from gam_ngs_b200 import capi
    assert capi.band_geometry(16) == (9, 4)      # 8 pairs per warp
    assert capi.band_geometry(32) == (18, 4)
    assert capi.band_geometry(64) == (18, 8)     # BASELINE config 2: 4 pairs per warp
    assert capi.band_geometry(150) == (10, 32)   # the reference's default band
    assert capi.band_geometry(256) == (18, 32)   # BASELINE config 3
    assert capi.band_geometry(512) == (18, 64)   # CTA-per-pair kernel
    assert capi.band_geometry(1024) == (18, 128)
    assert capi.band_geometry(2303) == (18, 256)
    assert capi.band_geometry(2304) is None      # generic kernel
    for band in range(0, 2304, 7):
        c, lg = capi.band_geometry(band)
        assert 2 <= c <= 18 and lg in (4, 8, 16, 32, 64, 128, 256) and c * lg >= 2 * band + 1 and c not in (13, 17)


def test_host_worker_pool_under_concurrent_callers():
    """The pool behind batch preparation (parallel index / descriptor / result passes): several threads submit
    passes at once - producer and consumer of a pipelined batch, concurrent contexts - sizes below and above the
    one-slice threshold, results checked inside the library.  Pure host code."""
    lib = capi.load_library()
    for callers, items in ((1, 0), (1, 1000), (1, 200_000), (2, 65_536), (6, 50_000), (8, 3_000)):
        assert lib.gamx_host_selftest(callers, items) == 0, (callers, items)
