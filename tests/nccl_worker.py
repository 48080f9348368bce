"""Worker of tests/test_fullsize.py::test_two_gpu_sharding_nccl (launched with torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from common import Problem  # noqa: E402
from scbonita_b200 import simulator  # noqa: E402
from scbonita_b200.distributed import ShardedEvaluator  # noqa: E402

r = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(r)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
pr = Problem.synthetic(60, 5000, 16)
model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pr.node_positions,
                              pr.net.ind_len)
with simulator.Simulator(pr.bm, model, device=r) as sim:
    ev = ShardedEvaluator(sim)
    full = ev.evaluate(pr.individuals, pr.ko, pr.ki, pr.pop_an, pr.pop_ai)
    odd = ev.evaluate(pr.individuals[:15], pr.ko, pr.ki, pr.pop_an[:15], pr.pop_ai[:15])      # uneven shards
want = pr.oracle()["fitness"]
assert np.array_equal(full, want), (full, want)
assert np.array_equal(odd, want[:15]), (odd, want[:15])

# the same exchange over peer memory, as the last kernel of each generation's graph (scb_gather_*): three generations,
# so that both halves of the double-buffered gather buffers are used, with a different shard every time
world = dist.get_world_size()
width = 8
with simulator.Simulator(pr.bm, model, device=r) as sim:
    mine = torch.from_numpy(sim.gather_init(world, r, width)).cuda()
    handles = torch.empty(world * 128, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(handles, mine)
    sim.gather_connect(handles.cpu().numpy())
    for g in range(3):
        order = np.roll(np.arange(16), 5 * g)
        sl = order[r * width:(r + 1) * width]
        sim.upload(pr.individuals[sl], pr.ko, pr.ki, pr.pop_an[sl], pr.pop_ai[sl])
        sim.run()
        got = sim.gather_result()
        assert np.array_equal(got, want[order]), (g, got, want[order])
    dist.barrier()
    sim.gather_close()
dist.destroy_process_group()
print("rank", r, "ok")
