#!/usr/bin/env python
"""bench.py -- GA-generation throughput of the rule-inference fitness hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C3|C2]

A "step" is one GA generation: every individual of the population scored against every cell
(rule compilation K0 + simulator K2 + resolver K3, plus the fitness all-gather when N > 1).
Workload C3 = BASELINE.json configs[2], the configuration the north-star target is quoted on:
synthetic 50 000 cells x 200-node KEGG-shaped network, population 500 (per GPU: the population is
sharded across ranks, every rank holds the packed matrix; weak scaling).  Unit: cell-rule
evaluations (CRE) = cells x nodes x 99 synchronous steps per individual.

One JSON line on stdout (rank 0).  `value` times device-resident inputs with CUDA events on the
library's stream, L2 flushed between steps; `e2e` times the public host-buffer API (pinned host
tables -> H2D -> kernels -> D2H fitness).  `--impl reference` times the UNMODIFIED reference C
(oracle/_ref, built from /root/reference in the build container) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from scbonita_b200 import synth  # noqa: E402

WORKLOADS = {
    "C2": dict(nodes=60, cells=10_000, population=100, name="C2: synthetic 10k cells x 60-node network, population 100"),
    "C3": dict(nodes=200, cells=50_000, population=500, name="C3: synthetic 50k cells x 200-node network, population 500"),
    "T0": dict(nodes=24, cells=2_500, population=6, name="T0: tiny self-test shape (not a benchmark)"),
    "N90": dict(nodes=90, cells=50_000, population=500, name="N90: 50k cells x 90-node network, population 500 (occupancy experiment, not a benchmark)"),
}
STEPS_PER_CELL = 99            # simulator.c:375 with simSteps = 100
METRIC = "cell_rule_evals_per_sec"
UNIT = "CRE/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured"
    except Exception:
        return 6650.0, 1965.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_problem(wl, seed_shift=0):
    pr = synth.make_problem(wl["nodes"], wl["cells"], wl["population"], pop_seed=91011 + seed_shift)
    bm = np.zeros((wl["nodes"], wl["cells"]), np.uint8)
    bm[pr.node_positions] = pr.bin_rows
    return pr, bm


# ---------------------------------------------------------------------------------------------
# reference arm: the unmodified reference C on the host cores
# ---------------------------------------------------------------------------------------------
def reference_sample(pr, cells, n_ind, threads):
    """Times n_ind individuals x `cells` cells of the workload on the host cores.  Returns (seconds, CRE, kind).

    kind "reference": oracle/_ref, the UNMODIFIED reference scSyncBool driven with its own ctypes convention,
    `threads` concurrent calls (ctypes releases the GIL; this is how the reference parallelises local search,
    singleCell.py:490-492).  If the prebuilt reference library is not on this box, kind "port": the scalar C
    restatement oracle/scb_oracle.c (same loops, runtime strides), also one call per thread."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    m = oracle.Model(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pr.node_positions,
                     pr.net.ind_len, pr.knockouts, pr.knockins)
    try:
        ref = oracle.RefSimulator(cells)
        bm = ref.dense(pr.bin_rows[:, :cells].astype(np.int32), pr.node_positions)
        kind = "reference"

        def one(p):
            p = p % pr.individuals.shape[0]
            return ref.sync_bool(m, pr.individuals[p], bm, cells, 0, and_nodes=pr.pop_and_nodes[p],
                                 and_inv=pr.pop_and_inv[p]).sum()
    except (FileNotFoundError, OSError):
        bm = np.zeros((pr.net.n, cells), np.int32)
        bm[pr.node_positions] = pr.bin_rows[:, :cells]
        kind = "port"

        def one(p):
            p = p % pr.individuals.shape[0]
            return oracle.sync_bool(m, pr.individuals[p], bm, cells, 0, and_nodes=pr.pop_and_nodes[p],
                                    and_inv=pr.pop_and_inv[p])[0].sum()

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(n_ind)))
    dt = time.perf_counter() - t0
    return dt, float(n_ind) * cells * pr.net.n * STEPS_PER_CELL, kind


def run_reference(args, wl, rank):
    if rank != 0:
        return
    pr, _ = build_problem(wl)
    cores = len(os.sched_getaffinity(0))
    threads = min(cores, 64)
    cells = min(wl["cells"], 12_500)
    for _ in range(args.warmup):
        reference_sample(pr, cells, threads, threads)
    tot_t, tot_u, kind = 0.0, 0.0, "reference"
    for _ in range(args.steps):
        dt, u, kind = reference_sample(pr, cells, threads, threads)
        tot_t += dt; tot_u += u
    sample = "%d individuals x %d cells x %d nodes per step, %d concurrent scSyncBool calls (%s, gcc -O3)" % (
        threads, cells, wl["nodes"], threads,
        "oracle/_ref = the unmodified reference C" if kind == "reference" else "oracle/scb_oracle.c restatement")
    value = tot_u / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args, wl, rank, world, local_rank):
    from scbonita_b200 import _lib, simulator
    lib = _lib.load()
    if lib.scb_device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists)")
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # weak scaling: every rank scores a population of the same size drawn with the same seeds, so that the
    # per-rank work is identical and the measured efficiency reflects the collective, not sampling noise
    pr, bm = build_problem(wl)
    if args.scaling == "strong" and world > 1:
        # fixed global population: every rank scores its contiguous block (scbonita_b200.distributed.shard_bounds)
        from scbonita_b200.distributed import shard_bounds
        lo, hi = shard_bounds(wl["population"], rank, world)
        width = -(-wl["population"] // world)
        lo = min(lo, wl["population"] - width)          # equal block sizes for the all-gather
        pr.individuals = pr.individuals[lo:lo + width]
        pr.pop_and_nodes = pr.pop_and_nodes[lo:lo + width]
        pr.pop_and_inv = pr.pop_and_inv[lo:lo + width]
        wl = dict(wl, population=width)
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv,
                                  pr.node_positions, pr.net.ind_len)
    sim = simulator.Simulator(bm, model, device=local_rank)
    for kv in args.opt:
        k, v = kv.split("=")
        sim.set_option(getattr(_lib, "OPT_" + k.upper()), int(v))
    P, N, C = wl["population"], wl["nodes"], wl["cells"]
    units_per_step = float(P) * C * N * STEPS_PER_CELL * world

    # pinned host copies of the reference-format tables (what a GA generation hands over)
    pin = []

    def pinned(a):
        h = simulator.PinnedArray(a.shape, a.dtype)
        h.array[...] = a
        pin.append(h)
        return h.array
    h_ind = pinned(np.ascontiguousarray(pr.individuals, np.int32))
    h_an = pinned(np.ascontiguousarray(pr.pop_and_nodes, np.int32))
    h_ai = pinned(np.ascontiguousarray(pr.pop_and_inv, np.int32))
    h_out = pinned(np.zeros(P, np.int64))
    # the same population in the compact form a GA driver can keep: upstream triples, not term tables
    up, ui = simulator.upstream_from_tables(pr.net.and_len, pr.pop_and_nodes, pr.pop_and_inv)
    h_up, h_ui = pinned(up), pinned(ui)
    h2d_compact = h_ind.nbytes + h_up.nbytes + h_ui.nbytes + 4 * N * 4
    h2d = h_ind.nbytes + h_an.nbytes + h_ai.nbytes + 4 * N * 4
    d2h = h_out.nbytes

    gather = None
    if world > 1:
        import torch
        ptr, nbytes = None, None
        sim.upload(h_ind, pr.knockouts, pr.knockins, h_an, h_ai)
        sim.run(simulator.MODE_SUM)
        sim.sync()
        ptr, nbytes = sim.result_device_pointer(simulator.MODE_SUM)

        class _Wrap:
            __cuda_array_interface__ = {"shape": (P,), "typestr": "<i8", "data": (ptr, False), "version": 3}
        local = torch.as_tensor(_Wrap(), device=torch.device("cuda", local_rank))
        allfit = torch.empty(world * P, dtype=torch.int64, device=local.device)
        ext = torch.cuda.ExternalStream(sim.stream_pointer(), device=local.device)

        def gather():
            with torch.cuda.stream(ext):
                dist.all_gather_into_tensor(allfit, local)

    def barrier():
        sim.sync()
        if world > 1:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def step_resident():
        sim.run(simulator.MODE_SUM)
        if gather:
            gather()

    def step_e2e():
        sim.upload(h_ind, pr.knockouts, pr.knockins, h_an, h_ai)
        sim.run(simulator.MODE_SUM)
        if gather:
            gather()
        sim.fetch(simulator.MODE_SUM, out=h_out)

    def step_e2e_compact():
        sim.upload_compact(h_ind, h_up, h_ui, pr.knockouts, pr.knockins)
        sim.run(simulator.MODE_SUM)
        if gather:
            gather()
        sim.fetch(simulator.MODE_SUM, out=h_out)

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        barrier()
        total_ms, sim_ms, stats = 0.0, 0.0, None
        for _ in range(steps):
            sim.flush_l2()
            sim.event_record(0)
            step()
            sim.event_record(1)
            total_ms += sim.event_elapsed_ms(0, 1)
            if step is step_resident:
                sim.fetch(simulator.MODE_SUM, out=h_out)          # untimed: reads the stats of this step
                stats = sim.last_stats
                sim_ms += stats["sim_ms"]
        barrier()
        if world > 1:
            import torch
            t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms, sim_ms, stats

    sim.upload(h_ind, pr.knockouts, pr.knockins, h_an, h_ai)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, sim_ms, stats = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms, _, _ = timed(step_e2e, args.steps, max(args.warmup, 3))
    fit = h_out.copy()
    e2e_compact_ms, _, _ = timed(step_e2e_compact, args.steps, max(args.warmup, 3))
    if not np.array_equal(fit, h_out):
        raise RuntimeError("compact and table uploads disagree")
    sim.upload(h_ind, pr.knockouts, pr.knockins, h_an, h_ai)
    # roofline pass: the simulator kernel K2 timed on its own.  In the default configuration the same launch also
    # resolves queued groups in its tail (K3 work), so K2's own duration is taken with that fusion switched off.
    fused_default = "fuse=0" not in args.opt
    if fused_default:
        sim.set_option(_lib.OPT_FUSE, 0)
    _, sim_ms, stats = timed(step_resident, min(args.steps, 5), 2)
    sim_steps = min(args.steps, 5)
    if fused_default:
        sim.set_option(_lib.OPT_FUSE, 1)

    if rank == 0:
        hbm_peak, sm_max, peak_kind = load_peaks()
        ms_per_step = total_ms / args.steps
        value = units_per_step / (ms_per_step * 1e-3)
        info = sim.info()
        k2_ms = sim_ms / sim_steps
        per_gpu_units = float(P) * C * N * STEPS_PER_CELL
        # algorithmic HBM bytes of one K2 launch: packed matrix N*C/8, the full rule program (16 B/node) and the five
        # reduced programs K0 emits per individual ((N + 1) x 16 B each), fitness 8 B
        alg_bytes = N * C / 8.0 + P * N * 16.0 + P * 5 * (N + 1) * 16.0 + P * 8.0
        hbm_ach = alg_bytes / (k2_ms * 1e-3) / 1e9
        traffic = None
        try:        # dram__bytes_read.sum + dram__bytes_write.sum of K2 from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "r1", "k2_dram_traffic.json")) as f:
                t = json.load(f)
            if t.get("workload") == args.workload:
                traffic = t["dram_bytes_per_launch"]
        except Exception:
            pass
        # binding resource: shared-memory bandwidth, 128 B/clk/SM.  Executed traffic of one launch =
        # executed (tile, step) pairs x nodes x 128*K bytes x (inputs + 1 store); inputs/node from the network
        sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
        # state loads per node-update: 1, 1, 2, 3 for 0, 1, 3, 7 AND-terms (upper bound: K0 drops dead inputs)
        arity = float(np.mean([{0: 1, 1: 1, 3: 2, 7: 3}.get(int(a), 3) for a in pr.net.and_len]))
        smem_bytes_exec = stats["steps_executed"] * N * 128.0 * info["words_per_lane"] * (arity + 1.0)
        if stats.get("state_rows_moved"):
            # exact: K2 counts the rows its steps load and store (the steady program moves far fewer than the bound above)
            smem_bytes_exec = stats["state_rows_moved"] * 128.0 * info["words_per_lane"]
        smem_peak = 148 * 128.0 * sm_mhz * 1e6 / 1e9
        smem_ach = smem_bytes_exec / (k2_ms * 1e-3) / 1e9
        # measured on-box peaks of the on-chip resources (scb_microbench: every SM busy, independent chains)
        try:
            mb = {k: sim.microbench(k) for k in ("lop3", "popc", "lds")}
        except Exception as exc:
            mb = {"error": str(exc)}
        smem_peak_nominal = smem_peak
        if mb.get("lds"):
            smem_peak = mb["lds"] / 1e9
        # integer-logic roofline: algorithmic 7 LOP3 per (node, 32-cell word, step) = 0.21875 LOP3 per
        # CRE (a runtime 3-input truth table as a 7-mux tree); executed: K2 issues ONE LOP3 per node-word-step (the
        # table is an immediate reached through a jump table) and only for the steps it does not skip
        lop_alg = per_gpu_units * 0.21875 / (k2_ms * 1e-3)
        lop_exec = stats["steps_executed"] * N * 32.0 * info["words_per_lane"] / (k2_ms * 1e-3)
        cpu = None
        if not args.no_cpu_baseline:
            try:
                cores = min(len(os.sched_getaffinity(0)), 64)
                cells = min(C, 12_500)
                dt, u, kind = reference_sample(pr, cells, cores, cores)
                cpu = {"value": u / dt, "unit": UNIT, "cores": cores, "kind": kind,
                       "sample": "%d individuals x %d cells x %d nodes, %d concurrent scSyncBool calls of %s, %.1f s" % (
                           cores, cells, N, cores, "the unmodified reference C (oracle/_ref)" if kind == "reference"
                           else "the C restatement (oracle/scb_oracle.c)", dt)}
            except Exception as exc:           # oracle/_ref absent on this box
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % exc}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl["name"], "population_per_gpu": P, "cells": C, "nodes": N,
                       "population_seeds": "identical on every rank",
                       "sharding": "individuals sharded over %d rank(s), packed matrix replicated, all-gather of int64 fitness" % world,
                       "l2": "flushed (256 MiB memset) between timed steps", "geometry": info,
                       "options": args.opt},
            "ga_generations_per_sec": 1e3 / ms_per_step,
            "e2e": {"value": units_per_step / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "e2e_compact": {"value": units_per_step / (e2e_compact_ms / args.steps * 1e-3), "unit": UNIT,
                            "ms_per_step": e2e_compact_ms / args.steps, "h2d_bytes_per_step": int(h2d_compact),
                            "d2h_bytes_per_step": int(d2h),
                            "note": "same generation, population handed over as bit-strings + upstream triples "
                                    "(scb_pop_upload_compact) instead of the reference's [N][7][3] tables"},
            "gpu_launches": (3 if fused_default else 4) * args.steps,   # K0 compile, K2 simulate (+ in-kernel K3), K3 fill
            "roofline": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                         "frac": hbm_ach / hbm_peak, "traffic": traffic, "algorithmic_bytes": alg_bytes,
                         "peak_source": peak_kind,
                         "kernel": "scb_k_sim<%d>" % info["words_per_lane"], "kernel_ms": k2_ms,
                         "kernel_timing": "CUDA events around K2 on the library's stream, K2 run without the in-kernel "
                                          "resolver tail (option fuse=0) so that the duration is K2's own",
                         "note": "the path is shared-memory/issue bound, not HBM bound (SURVEY 8d): see roofline_binding"},
            "roofline_binding": {"bound": "smem", "achieved": smem_ach, "peak": smem_peak, "unit": "GB/s",
                                 "frac": smem_ach / smem_peak, "sm_mhz": sm_mhz,
                                 "peak_source": "measured: scb_microbench LDS.128, all SMs" if mb.get("lds") else
                                                "nominal 148 SMs x 128 B/clk x SM clock",
                                 "peak_nominal": smem_peak_nominal, "frac_of_nominal": smem_ach / smem_peak_nominal,
                                 "executed_step_fraction": stats["steps_executed"] / (stats["tiles"] * 99.0),
                                 "algorithmic_cre_per_sec_k2": per_gpu_units / (k2_ms * 1e-3)},
            "roofline_lop": {"bound": "int-logic", "unit": "LOP3/s", "peak": mb.get("lop3"),
                             "peak_source": "measured: scb_microbench LOP3, all SMs",
                             "achieved_algorithmic": lop_alg, "frac_algorithmic": lop_alg / mb["lop3"] if mb.get("lop3") else None,
                             "achieved_executed": lop_exec, "frac_executed": lop_exec / mb["lop3"] if mb.get("lop3") else None,
                             "popc_peak": mb.get("popc"),
                             "note": "algorithmic = CRE/s x 7/32 (SURVEY 8d mux-tree figure); executed = one LOP3 per "
                                     "node-word-step actually simulated (upper bound: counts all N nodes, the reduced "
                                     "programs compute fewer); POPC is used once per (node, word, tile)"},
            "stats": stats, "clocks": clocks, "cpu_baseline": cpu,
            "fitness_checksum": int(fit.sum()),
        }
        print(json.dumps(line), flush=True)
    sim.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    # stdout carries exactly one JSON line: everything else that libraries print there (NCCL's version banner,
    # ...) is sent to stderr by pointing fd 1 at fd 2 until the line is written
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--opt", action="append", default=[], help="library option, e.g. early_exit=0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default): 500 individuals per GPU; strong: the 500 individuals split over the GPUs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
    else:
        run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
