"""Multi-process plumbing for `bench.py --gpus N` (one process per GPU, torch.distributed).

The path shards by independent clusters: every rank diffs its own snapshot and nothing crosses ranks on the data
path.  torch.distributed is used for the barrier around the timed region, the max-over-ranks
of the elapsed time and gathering per-rank counters — NCCL on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

import os


def rank_env() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def rank_seed(base_seed: int, rank: int) -> int:
    """Every rank generates a different cluster (same distributions)."""
    return int(base_seed) + 1000 * int(rank)


class Ranks:
    def __init__(self, backend: str | None = None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local_rank = rank_env()
        self.device = device
        self.owns_group = False
        if self.world > 1 and not dist.is_initialized():
            kw = {}
            if backend == "nccl" and device is not None:
                kw["device_id"] = device
            dist.init_process_group(backend or "gloo", **kw)
            self.owns_group = True

    def barrier(self):
        if self.device is not None and self.device.type == "cuda":
            self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            if self.device is not None and self.device.type == "cuda":
                self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_counts(self, values: list[int]) -> list[list[int]]:
        """all_gather of a small list of per-rank integers (ops, objects ...)."""
        if self.world == 1:
            return [list(values)]
        t = self.torch.tensor(values, dtype=self.torch.int64, device=self.device if self.device is not None else "cpu")
        out = [self.torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [[int(v) for v in o.tolist()] for o in out]

    def close(self):
        if self.owns_group and self.dist.is_initialized():
            self.dist.barrier()
            self.dist.destroy_process_group()
