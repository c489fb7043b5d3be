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


def _gpu_pci_path(local_rank: int):
    import torch
    pr = torch.cuda.get_device_properties(local_rank)
    if all(hasattr(pr, k) for k in ("pci_bus_id", "pci_device_id", "pci_domain_id")):
        return f"/sys/bus/pci/devices/{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    import subprocess
    out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)], capture_output=True, text=True, timeout=20).stdout.strip()
    dom, rest = out.lower().split(":", 1)  # 00000000:1B:00.0
    return f"/sys/bus/pci/devices/{dom[-4:]}:{rest}"


class NumaBinding:
    """Pin this process (and the threads it starts) to the CPUs of the NUMA node the GPU hangs off while host tables are
    allocated and first-touched, so that the pages the engine later copies from are local to the GPU's PCIe root complex; with
    8 ranks copying 1.7 GB each per step, remote-node reads held the end-to-end replicas efficiency at 0.65.  `release()`
    restores the original affinity (CPU-side work such as the oracle then uses every core again).  `node` is None when the
    topology cannot be read (containers without /sys topology): nothing is changed then."""

    def __init__(self, local_rank: int):
        self.node, self.saved = None, None
        try:
            path = _gpu_pci_path(local_rank)
            node = int(open(path + "/numa_node").read().strip())
            ids = set()
            for part in open(path + "/local_cpulist").read().strip().split(","):
                if "-" in part:
                    lo, hi = part.split("-")
                    ids.update(range(int(lo), int(hi) + 1))
                elif part:
                    ids.add(int(part))
            allowed = os.sched_getaffinity(0)
            ids &= allowed
            if node >= 0:
                self.node = node
            if node >= 0 and ids and ids != allowed:
                os.sched_setaffinity(0, ids)
                self.saved = allowed
        except Exception:  # noqa: BLE001 - leave the affinity alone
            pass

    def release(self):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:  # noqa: BLE001
                pass
            self.saved = None


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
