"""Host-side breakdown of the end-to-end path (run with GAMX_TIMING=1)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import gen
import gam_ngs_b200 as g

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
rng = np.random.default_rng(1)
cache = f"/tmp/e2e_probe_{n}.npz"  # (a sweep runs this script once per setting: generate once)
if os.path.exists(cache):
    z = np.load(cache); a, al, b, bl = z["a"], z["al"], z["b"], z["bl"]
else:
    a, al, b, bl = gen.bulk_pairs(rng, n, 1000)
    np.savez(cache, a=a, al=al, b=b, bl=bl)
host = torch.empty(len(a) + len(b), dtype=torch.uint8, pin_memory=True)
hv = host.numpy()
ao = np.concatenate([[0], np.cumsum(al)]).astype(np.int64); bo = np.concatenate([[0], np.cumsum(bl)]).astype(np.int64)
lengths = np.empty(2 * n, dtype=np.uint64); a_id = np.empty(n, np.uint32); b_id = np.empty(n, np.uint32)
pos = cid = 0
for lo in range(0, n, 16384):  # blocks of pairs: [a-contigs][b-contigs], like bench.py
    hi = min(n, lo + 16384)
    for src, off, ln, ids in ((a, ao, al, a_id), (b, bo, bl, b_id)):
        seg = src[off[lo]:off[hi]]; hv[pos:pos + len(seg)] = seg; pos += len(seg)
        lengths[cid:cid + hi - lo] = ln[lo:hi]; ids[lo:hi] = np.arange(cid, cid + hi - lo); cid += hi - lo
jobs = g.make_jobs(n)
jobs["a_id"] = a_id; jobs["b_id"] = b_id
jobs["end_a"] = al - 1; jobs["end_b"] = bl - 1; jobs["band"] = 64; jobs["mode"] = 1
ctx = g.Context(devices=[0])
tot, resid = [], []
for it in range(5):
    t0 = time.perf_counter(); ctx.clear_contigs()
    t1 = time.perf_counter(); ctx.add_contigs(host.data_ptr(), lengths, async_upload=True)
    t2 = time.perf_counter(); res, ops = ctx.align_batch(jobs)
    t3 = time.perf_counter()
    tot.append(1e3 * (t3 - t0))
    print(f"iter {it}: clear {1e3*(t1-t0):.1f} add_contigs {1e3*(t2-t1):.1f} align_batch {1e3*(t3-t2):.1f} total {1e3*(t3-t0):.1f} ms", flush=True)
for it in range(3):  # contigs stay resident: the pipelined batch alone
    t2 = time.perf_counter(); res, ops = ctx.align_batch(jobs)
    t3 = time.perf_counter()
    resid.append(1e3 * (t3 - t2))
    print(f"resident store, iter {it}: align_batch {1e3*(t3-t2):.1f} ms", flush=True)
env = {k: v for k, v in os.environ.items() if k.startswith("GAMX_") and k != "GAMX_TIMING"}
print(f"SUMMARY {env}: end to end min {min(tot[1:]):.1f} ms, resident store min {min(resid):.1f} ms", flush=True)
