"""torch.distributed plumbing for multi-process runs (one process per GPU, launched by torchrun).

The pairwise path needs no data-path collective: ranks own disjoint panels (dg_run_part).  The only
communication is the timing protocol of bench.py: a barrier on both sides of the timed region, MAX of
the per-rank times, SUM of the per-rank pair counts.  Backend: nccl when CUDA is available, else gloo
(the CPU tests run world_size 2 over gloo)."""
from __future__ import annotations

import os


class Dist:
    def __init__(self, backend: str | None = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.pg = None
        self.device = None
        if self.world > 1:
            import torch
            import torch.distributed as td
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                self.device = torch.device("cuda", self.local_rank)
            else:
                self.device = torch.device("cpu")
            # NCCL logs (its version banner included) go to stdout by default: keep stdout to bench.py's one JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            import datetime
            # a rank that dies (or a mismatched collective) must not hold the GPUs for NCCL's default 10 minutes
            td.init_process_group(backend=backend, rank=self.rank, world_size=self.world, timeout=datetime.timedelta(seconds=240))
            self.pg = td

    def barrier(self):
        if self.pg is not None:
            self.pg.barrier()

    def _reduce(self, value: float, op: str) -> float:
        if self.pg is None:
            return float(value)
        import torch
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device)
        self.pg.all_reduce(t, op=getattr(self.pg.ReduceOp, op))
        return float(t.item())

    def all_gather_into(self, out, inp):
        """NCCL all-gather of equal-sized device tensors (the input alignment: each rank uploads 1/N of the code
        bytes over PCIe and receives the rest over NVLink)."""
        if self.pg is None:
            out.copy_(inp)
        else:
            self.pg.all_gather_into_tensor(out, inp)

    def max(self, value: float) -> float:
        return self._reduce(value, "MAX")

    def sum(self, value: float) -> float:
        return self._reduce(value, "SUM")

    def close(self):
        if self.pg is not None:
            self.pg.destroy_process_group()
            self.pg = None


def my_panels(plan, rank: int, world: int):
    """The panels dg_run_part(part=rank, n_parts=world) processes: those dg_plan_parts assigns to `rank` (largest panels
    first, each to the least-loaded part)."""
    if world <= 1:
        return list(plan)
    from . import api
    part_of = api.plan_parts(plan, world)
    return [p for k, p in enumerate(plan) if part_of[k] == rank]
