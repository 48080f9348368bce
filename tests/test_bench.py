"""bench.py's own arm runs end to end (on the emulator build, tiny shape) and prints the contracted JSON keys."""
import argparse
import json

import bench
from scbonita_b200 import _lib


def test_bench_line_has_contract_keys(emu_lib, monkeypatch, capsys):
    monkeypatch.setattr(_lib, "DEFAULT_PATH", emu_lib)
    args = argparse.Namespace(gpus=1, steps=2, warmup=1, impl="ours", workload="T0", opt=["max_ctas=2"],
                              no_cpu_baseline=False, scaling="strong", gather="peer", no_order=False, no_cases=True)
    bench.run_ours(args, bench.WORKLOADS["T0"], 0, 1, 0)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert key in line, key
    assert line["config"]["workload"].startswith("T0")
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["gpu_launches"] == 3 * 2
    # extra evidence next to the contract keys: compact hand-over, binding (shared-memory) and integer-logic rooflines
    assert line["e2e_compact"]["h2d_bytes_per_step"] < line["e2e"]["h2d_bytes_per_step"]
    assert set(("bound", "achieved", "peak", "frac")) <= set(line["roofline_binding"])
    assert "achieved_algorithmic" in line["roofline_lop"] and line["stats"]["state_rows_moved"] > 0
    assert line["cpu_baseline_1thread"]["cores"] == 1 and set(("min", "median", "max")) <= set(line["per_step_ms"]["max_over_ranks"])


def test_lpt_shards_are_balanced_and_rectangular():
    import numpy as np
    rng = np.random.default_rng(3)
    costs = rng.integers(100, 5000, 500)
    shards, width = bench.lpt_shards(costs, 8)
    assert width == 63 and sorted(i for sh in shards for i in sh) == list(range(500))
    assert all(len(sh) <= width for sh in shards)
    loads = [int(costs[sh].sum()) for sh in shards]
    assert max(loads) / (sum(loads) / 8) < 1.02


def test_reference_arm_line(capsys):
    args = argparse.Namespace(gpus=1, steps=1, warmup=0, impl="reference", workload="T0", opt=[], no_cpu_baseline=False, scaling="strong")
    bench.run_reference(args, bench.WORKLOADS["T0"], 0)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
