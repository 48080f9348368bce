"""Population sharding across GPUs: one process per GPU, the packed cell matrix replicated, individuals
split in contiguous blocks, one all-gather of the int64 fitness vector per generation;
and independent pathways spread over the GPUs with no communication until the end (config C4).

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


def pathway_cost(n_nodes, n_cells, evaluations=1):
    """Work of one pathway in cell-rule evaluations: nodes x cells x 99 steps x scored rule sets."""
    return float(n_nodes) * float(n_cells) * 99.0 * float(evaluations)


def assign_pathways(costs, world):
    """Independent KEGG pathways over the GPUs of a box (BASELINE config C4, SURVEY 8e ii): longest-processing-time
    first, every pathway to the currently least loaded rank, ties to the lower rank.  The reference runs one SLURM
    job per network (pipeline.py:24-61); here a pathway is one context on one GPU and nothing is exchanged until the
    results are collected.  Returns ``world`` lists of pathway indices; deterministic, so every rank can compute it."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * int(world)
    out = [[] for _ in range(int(world))]
    for i in order:
        r = min(range(int(world)), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(costs[i])
    return out


def run_pathways(pathways, evaluate, costs=None, group=None):
    """Every rank evaluates the pathways :func:`assign_pathways` gives it (``evaluate(index, pathway)`` builds the
    context on this rank's GPU and returns a picklable result) and all ranks collect all results at the end with ONE
    ``all_gather_object`` -- the only communication of config C4.  Returns the results in pathway order."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if costs is None:
        costs = [1.0] * len(pathways)
    mine = assign_pathways(costs, world)[rank]
    local = {i: evaluate(i, pathways[i]) for i in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, local, group=group)
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(pathways))]


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
            # fetch() also reads the compile diagnostics (bad index, > 7 terms, ...) and raises on them: garbage
            # fitness must not be gathered
            local[:hi - lo] = self.sim.fetch(MODE_SUM)
        # the collective is chosen from (P, world, backend) only, so every rank issues the same one
        src = torch.from_numpy(local)
        if self.backend == "nccl":
            src = src.cuda()
        out = torch.empty(self.world * width, dtype=torch.int64, device=src.device)
        self.dist.all_gather_into_tensor(out, src, group=self.group)
        gathered = out.cpu().numpy()
        full = np.zeros(P, dtype=np.int64)
        for r in range(self.world):
            a, b = shard_bounds(P, r, self.world)
            full[a:b] = gathered[r * width:r * width + (b - a)]
        return full
