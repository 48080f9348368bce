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
dist.destroy_process_group()
print("rank", r, "ok")
