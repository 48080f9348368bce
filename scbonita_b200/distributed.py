"""Population sharding across GPUs: one process per GPU, the packed cell matrix replicated, individuals
split in contiguous blocks, one all-gather of the int64 fitness vector per generation.

The reference has no distributed layer at all (one SLURM job per network, pipeline.py:24-61); the GA itself
is sequential Python.  After the all-gather every rank holds the complete fitness vector, so the unchanged,
deterministic selection code (ruleMaker.py:1024-1075) runs identically on every rank.

``torch.distributed`` is plumbing only: NCCL over NVLink on the GPU box (the device buffer of the library is
handed to the collective without a host round trip), gloo in the CPU tests.
"""
import numpy as np

from .simulator import MODE_SUM


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of rank; block sizes differ by at most one."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _DevicePointer:
    """Minimal __cuda_array_interface__ view so that torch can wrap the library's device buffer."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}


class ShardedEvaluator:
    """Scores a population with every rank of ``group`` working on its block of individuals."""

    def __init__(self, sim, group=None):
        import torch.distributed as dist
        self.sim, self.dist, self.group = sim, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)

    def evaluate(self, individuals, knockouts=None, knockins=None, and_nodes=None, and_inv=None):
        """Returns int64[P] -- the full fitness vector (sum of cellwise errors), identical on every rank."""
        import torch
        inds = np.ascontiguousarray(individuals, dtype=np.int32)
        P = inds.shape[0]
        lo, hi = shard_bounds(P, self.rank, self.world)
        width = -(-P // self.world)                                  # every rank contributes `width` slots
        per_ind = and_nodes is not None and np.ndim(and_nodes) == 4
        local = np.zeros(width, dtype=np.int64)
        if hi > lo:
            an = and_nodes[lo:hi] if per_ind else and_nodes
            ai = and_inv[lo:hi] if per_ind else and_inv
            self.sim.upload(inds[lo:hi], knockouts, knockins, an, ai)
            self.sim.run(MODE_SUM)
        if self.backend == "nccl" and hi - lo == width:
            # all-gather straight from the library's device buffer, on the library's stream
            self.sim.sync()
            ptr, _ = self.sim.result_device_pointer(MODE_SUM)
            dev = torch.device("cuda", torch.cuda.current_device())
            src = torch.as_tensor(_DevicePointer(ptr, width), device=dev)
            out = torch.empty(self.world * width, dtype=torch.int64, device=dev)
            self.dist.all_gather_into_tensor(out, src, group=self.group)
            gathered = out.cpu().numpy()
        else:
            if hi > lo:
                local[:hi - lo] = self.sim.fetch(MODE_SUM)
            src = torch.from_numpy(local)
            if self.backend == "nccl":
                src = src.cuda()
            outs = [torch.empty_like(src) for _ in range(self.world)]
            self.dist.all_gather(outs, src, group=self.group)
            gathered = torch.cat(outs).cpu().numpy()
        full = np.zeros(P, dtype=np.int64)
        for r in range(self.world):
            a, b = shard_bounds(P, r, self.world)
            full[a:b] = gathered[r * width:r * width + (b - a)]
        return full
