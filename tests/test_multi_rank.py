"""CPU tests of the N>1 path: (1) bench.py's rank plumbing (barrier, max-over-ranks time,
whole-job aggregate) with world_size 2 over gloo; (2) the library's cost-balanced sharding of a
batch over devices (host logic, exported for this test)."""
import ctypes as C
import os
import subprocess
import sys
import textwrap

import numpy as np

from gam_ngs_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gloo_world_size_2_aggregation(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        from gam_ngs_b200.dist import Ranks, whole_job_rate, shard_seed
        r = Ranks(backend="gloo")
        assert r.world == 2
        r.barrier()
        # rank 0 processes 100 units/step in 2 s, rank 1 processes 300 units/step in 4 s (3 steps)
        units, secs = (100.0, 2.0) if r.rank == 0 else (300.0, 4.0)
        rate = whole_job_rate(units, secs, 3, r)
        assert abs(rate - (400.0 * 3 / 4.0)) < 1e-9, rate
        assert r.max(float(r.rank)) == 1.0 and r.sum(1.0) == 2.0
        assert shard_seed(1000, r.rank) == 1000 + r.rank
        r.barrier()
        r.close()
        print("rank", r.rank, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("ok") == 2


def test_library_sharding_is_cost_balanced():
    lib = capi.load_library()
    lib.gamx_shard_by_cost.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
    lib.gamx_shard_by_cost.restype = C.c_int
    rng = np.random.default_rng(8)
    cost = (rng.integers(10_000, 50_000, size=5000) * 513).astype(np.uint64)  # config-3 like: 10-50 kb, band 256
    for nd in (1, 2, 4, 8):
        out = np.zeros(len(cost), dtype=np.int32)
        assert lib.gamx_shard_by_cost(cost.ctypes.data, len(cost), nd, out.ctypes.data) == 0
        assert out.min() == 0 and out.max() == nd - 1
        loads = np.array([cost[out == d].sum() for d in range(nd)], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.01, loads  # LPT: within 1 % of perfect balance
    # every job lands on exactly one device (disjoint slices of the caller's arrays: host-side gather)
    assert np.bincount(out, minlength=8).sum() == len(cost)
