"""The N > 1 path on CPU: world_size 2 over gloo, every rank drives the emulator build of the kernels on its
block of individuals, one all-gather of the fitness vector.  (On the GPU box the same code runs over NCCL.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from common import EMU_LIB, ROOT, Problem
from scbonita_b200.distributed import assign_pathways, pathway_cost, shard_bounds


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 500, 501):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, pop, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scbonita_b200.distributed import ShardedEvaluator
    pr = Problem.synthetic(30, 700, pop)
    with pr.simulator(EMU_LIB, max_ctas=2) as sim:
        ev = ShardedEvaluator(sim)
        full = ev.evaluate(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai)
        shared = ev.evaluate(pr.individuals, pr.ko, pr.ki)                      # shared tables path
    q.put((rank, full.tolist(), shared.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("pop", [7, 8])
def test_two_ranks_gloo_match_oracle(emu_lib, pop):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, pop, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pr = Problem.synthetic(30, 700, pop)
    want = pr.oracle()["fitness"].tolist()
    pr2 = Problem.synthetic(30, 700, pop)
    pr2.pop_an = pr2.pop_ai = None
    want_shared = pr2.oracle()["fitness"].tolist()
    for rank, full, shared in got:
        assert full == want, rank
        assert shared == want_shared, rank


def test_pathway_assignment_is_balanced_and_deterministic():
    sizes = [47, 60, 100, 159, 200, 300] * 8 + [341, 47]                    # ~50 KEGG pathways (config C4)
    costs = [pathway_cost(n, 60_000, 500) for n in sizes]
    parts = assign_pathways(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    assert parts == assign_pathways(costs, 8)
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) <= sum(costs) / 8 + max(costs)                        # the LPT guarantee
    assert max(loads) / (sum(costs) / 8) < 1.05
    assert assign_pathways(costs, 1) == [sorted(range(len(sizes)), key=lambda i: (-costs[i], i))]


def _pathway_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from scbonita_b200.distributed import pathway_cost, run_pathways
    sizes = [12, 30, 21, 9, 17]
    seen = []

    def evaluate(i, n):
        seen.append(i)
        pr = Problem.synthetic(n, 400, 3)
        with pr.simulator(EMU_LIB, max_ctas=2) as sim:
            return sim.evaluate_population(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai).tolist()
    res = run_pathways(sizes, evaluate, costs=[pathway_cost(n, 400, 3) for n in sizes])
    q.put((rank, sorted(seen), res))
    dist.destroy_process_group()


def test_independent_pathways_two_ranks_gloo(emu_lib):
    """Config C4 in miniature: five pathways over two ranks, no communication until the final gather."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pathway_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sizes = [12, 30, 21, 9, 17]
    want = [Problem.synthetic(n, 400, 3).oracle()["fitness"].tolist() for n in sizes]
    seen_all = sorted(i for _, seen, _ in got for i in seen)
    assert seen_all == list(range(5))                                       # every pathway evaluated exactly once
    for rank, seen, res in got:
        assert res == want, rank
