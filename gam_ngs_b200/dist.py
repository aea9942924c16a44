"""Multi-GPU plumbing for bench.py: one process per GPU, torch.distributed only for the barrier
and for reducing the timing/work counters.  The alignment path itself needs no collective: pair
batches shard independently, results are gathered on the host."""
from __future__ import annotations

import os


class Ranks:
    """Thin wrapper so the same code runs with NCCL on GPUs and with gloo in the CPU tests."""

    def __init__(self, backend: str | None = None, device=None):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        self.device = device
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29531")
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend, rank=self.rank, world_size=self.world, **kw)
            self.dist = dist
            if device is None:
                self.device = torch.device("cuda", self.local_rank) if backend == "nccl" else torch.device("cpu")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        if self.torch.cuda.is_available():
            self.torch.cuda.synchronize()

    def _reduce(self, x: float, op) -> float:
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX if self.dist else None)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM if self.dist else None)

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()
            self.dist = None


def shard_seed(base_seed: int, rank: int) -> int:
    """Every rank aligns its own independent pair batch (weak scaling)."""
    return base_seed + rank


def whole_job_rate(units_per_rank: float, seconds_this_rank: float, steps: int, ranks: Ranks) -> float:
    """Whole-job throughput: units all ranks processed / slowest rank's time."""
    total = ranks.sum(units_per_rank) * steps
    return total / ranks.max(seconds_this_rank)
