"""Debug: pipelined gamx_align_batch vs the resident plan on a config-2 shaped batch."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gen
import gam_ngs_b200 as g
from gam_ngs_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
rng = np.random.default_rng(1)
a, al, b, bl = gen.bulk_pairs(rng, n, 1000)
host = torch.empty(len(a) + len(b), dtype=torch.uint8, pin_memory=True)
hv = host.numpy()
ao = np.concatenate([[0], np.cumsum(al)]).astype(np.int64); bo = np.concatenate([[0], np.cumsum(bl)]).astype(np.int64)
lengths = np.empty(2 * n, dtype=np.uint64); a_id = np.empty(n, np.uint32); b_id = np.empty(n, np.uint32)
pos = cid = 0
for lo in range(0, n, 16384):
    hi = min(n, lo + 16384)
    for src, off, ln, ids in ((a, ao, al, a_id), (b, bo, bl, b_id)):
        seg = src[off[lo]:off[hi]]; hv[pos:pos + len(seg)] = seg; pos += len(seg)
        lengths[cid:cid + hi - lo] = ln[lo:hi]; ids[lo:hi] = np.arange(cid, cid + hi - lo); cid += hi - lo
jobs = g.make_jobs(n)
jobs["a_id"] = a_id; jobs["b_id"] = b_id
jobs["end_a"] = al - 1; jobs["end_b"] = bl - 1; jobs["band"] = 64; jobs["mode"] = 1
ctx = g.Context(devices=[0])
ctx.add_contigs(host.data_ptr(), lengths)
plan = ctx.plan(jobs); plan.run(); plan.sync(); ref, _ = plan.fetch(); plan.close()
out = np.empty(n, dtype=capi.RESULT_DTYPE)
for name, chunk, asyn in (("single,async", 0, True), ("pipe,sync", 65536, False), ("pipe,async", 65536, True), ("pipe,async", 65536, True)):
    ctx.set_pipeline_chunk(chunk)
    ctx.clear_contigs()
    ctx.add_contigs(host.data_ptr(), lengths, async_upload=asyn)
    t = time.perf_counter()
    got, _ = ctx.align_batch(jobs, out=out)
    dt = time.perf_counter() - t
    bad = np.nonzero([got[k].tobytes() != ref[k].tobytes() for k in range(0, n)])[0] if got.tobytes() != ref.tobytes() else []
    print(name, f"{dt*1e3:.1f} ms", "mismatches:", len(bad), list(bad[:8]), flush=True)
    if len(bad):
        k = int(bad[0]); print("  got", got[k]); print("  ref", ref[k])
