#!/usr/bin/env python
"""bench.py -- GA-generation throughput of the rule-inference fitness hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C3|C2|C4|C5|L0]
                    [--scaling strong|weak] [--gather peer|nccl] [--no-order] [--no-cases] [--opt key=value ...]

A "step" is one GA generation: every individual of the population scored against every cell (one CUDA graph per rank:
rule compilation K0 + simulator K2 with the resolver inside + fill-forward, plus the fitness exchange when N > 1).
Workload C3 = BASELINE.json configs[2], the configuration the north-star target is quoted on: synthetic 50 000 cells x
200-node KEGG-shaped network, population 500.  With N > 1 the population stays 500 (strong scaling, the default: the
individuals are dealt to the ranks by measured cost, every rank holds the packed matrix, the gathered fitness vector is
checked against a single-GPU evaluation); `--scaling weak` gives every rank its own 500.  Unit: cell-rule evaluations
(CRE) = cells x nodes x 99 synchronous steps per individual.

One JSON line on stdout (rank 0).  `value` times device-resident inputs with CUDA events on the library's stream, L2
flushed between steps, max over ranks per step; `e2e` times the public host-buffer API (pinned host tables -> H2D ->
kernels -> D2H fitness); `e2e_compact` / `e2e_resident` the same generation handed over as upstream triples / as an
in-place replacement of half the population.  `robustness` repeats the generation on other network seeds, without
knock-outs, with a fresh population every step and on an adversarial network.  `--impl reference` times the UNMODIFIED
reference C (oracle/_ref, built from /root/reference in the build container) on all host cores.
Other workloads: C2 (small configuration), C4 (independent pathways over the GPUs), C5 (attractor analysis),
L0 (per-call cost of the exported legacy symbol).
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
    "C3s8": dict(nodes=200, cells=50_000, population=63, name="C3s8: one rank's share of C3 at 8 GPUs (63 individuals; tuning aid, not a benchmark)"),
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            # nvidia-smi's own start-up (NVML initialisation takes driver locks) must not fall into the timed region:
            # wait for its first sample
            t0 = time.time()
            while not self.lines and time.time() - t0 < 3.0 and self.proc.poll() is None:
                time.sleep(0.01)
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
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl["name"], "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
class Generation:
    """One (dataset, network, population shard) on this rank's GPU with pinned host copies of what a GA generation
    hands over, and the step functions the timed loops call."""

    def __init__(self, pr, bm, device, opts):
        from scbonita_b200 import _lib, simulator
        self.simulator, self.pr = simulator, pr
        model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv,
                                      pr.node_positions, pr.net.ind_len)
        self.sim = simulator.Simulator(bm, model, device=device)
        for kv in opts:
            k, v = kv.split("=")
            self.sim.set_option(getattr(_lib, "OPT_" + k.upper()), int(v))
        self.pin = []
        self.set_population(pr.individuals, pr.pop_and_nodes, pr.pop_and_inv)

    def pinned(self, a):
        h = self.simulator.PinnedArray(a.shape, a.dtype)
        h.array[...] = a
        self.pin.append(h)
        return h.array

    def set_population(self, inds, an, ai):
        sim_mod = self.simulator
        self.P = inds.shape[0]
        self.h_ind = self.pinned(np.ascontiguousarray(inds, np.int32))
        self.h_an = self.pinned(np.ascontiguousarray(an, np.int32))
        self.h_ai = self.pinned(np.ascontiguousarray(ai, np.int32))
        self.h_out = self.pinned(np.zeros(self.P, np.int64))
        up, ui = sim_mod.upstream_from_tables(self.pr.net.and_len, an, ai)
        self.h_up, self.h_ui = self.pinned(up), self.pinned(ui)
        n = self.pr.net.n
        self.h2d = self.h_ind.nbytes + self.h_an.nbytes + self.h_ai.nbytes + 4 * n * 4
        self.h2d_compact = self.h_ind.nbytes + self.h_up.nbytes + self.h_ui.nbytes + 4 * n * 4
        self.d2h = self.h_out.nbytes

    def upload(self):
        self.sim.upload(self.h_ind, self.pr.knockouts, self.pr.knockins, self.h_an, self.h_ai)

    def step_resident(self):
        self.sim.run(self.simulator.MODE_SUM)

    def step_e2e(self):
        self.upload()
        self.sim.run(self.simulator.MODE_SUM)
        self.sim.fetch(self.simulator.MODE_SUM, out=self.h_out)

    def step_e2e_resident(self):
        """(mu + lambda) generation on a resident population: the second half of the slots are this generation's
        offspring (uploaded as upstream triples), the first half are the parents, whose sums stay on the device."""
        half = self.P // 2
        sl = slice(self.P - half, self.P)
        self.sim.update(self.slots_half, self.h_ind[sl], self.h_up[sl], self.h_ui[sl])
        self.sim.run(self.simulator.MODE_SUM)
        self.sim.fetch(self.simulator.MODE_SUM, out=self.h_out)

    def step_e2e_compact(self):
        self.sim.upload_compact(self.h_ind, self.h_up, self.h_ui, self.pr.knockouts, self.pr.knockins)
        self.sim.run(self.simulator.MODE_SUM)
        self.sim.fetch(self.simulator.MODE_SUM, out=self.h_out)


def timed_steps(gen, step, steps, warmup, barrier, after=None, stats_each=False):
    """`steps` timed calls of `step` after `warmup` untimed ones; every step bracketed by CUDA events on the library's
    stream, L2 flushed before it.  Returns (per-step ms list, stats of the last step, K2-only ms list)."""
    sim = gen.sim
    for _ in range(warmup):
        step()
        if after:
            after()
    barrier()
    ms, k2, stats = [], [], None
    for _ in range(steps):
        sim.flush_l2()
        sim.event_record(0)
        step()
        sim.event_record(1)
        ms.append(sim.event_elapsed_ms(0, 1))
        if after:
            after()
        if stats_each:
            sim.fetch(gen.simulator.MODE_SUM, out=gen.h_out)       # untimed: reads the stats of this step
            stats = sim.last_stats
            k2.append(stats["sim_ms"])
    barrier()
    return ms, stats, k2


def spread(ms):
    a = np.asarray(ms, np.float64)
    return {"min": float(a.min()), "median": float(np.median(a)), "max": float(a.max())}


def lpt_shards(costs, world):
    """Individuals over the ranks, longest processing time first (every individual to the least loaded rank, ties to
    the lower rank): the shards of a strong-scaling generation are balanced by predicted cost, not by count.  Every
    rank gets the same number of slots (the exchange buffer is rectangular)."""
    P = len(costs)
    width = -(-P // world)
    order = sorted(range(P), key=lambda i: (-int(costs[i]), i))
    load, out = [0] * world, [[] for _ in range(world)]
    for i in order:
        r = min((k for k in range(world) if len(out[k]) < width), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(costs[i])
    return out, width


def robustness_cases(args, wl, device):
    """The 10 ms target off the happy path: other network seeds, no knock-outs, a fresh population in
    every step, an adversarial long-period network of the same size.  A few steps each."""
    cases = []
    P, N, C = wl["population"], wl["nodes"], wl["cells"]
    units = float(P) * C * N * STEPS_PER_CELL

    def run_case(name, pr, bm, fresh=None):
        gen = Generation(pr, bm, device, args.opt)
        gen.upload()
        pops = fresh or [None]
        ms_all, fracs = [], []
        for k, pop in enumerate(pops):
            if pop is not None:
                gen.set_population(*pop)
                gen.upload()
            ms, stats, _ = timed_steps(gen, gen.step_resident, 3 if fresh is None else 1, 2 if k == 0 else 1, gen.sim.sync, stats_each=True)
            ms_all += ms
            fracs.append(stats["steps_executed"] / (stats["tiles"] * 99.0))
        e2e, _, _ = timed_steps(gen, gen.step_e2e, 2, 1, gen.sim.sync)
        cases.append({"case": name, "ms_per_generation": spread(ms_all), "e2e_ms": float(np.median(e2e)),
                      "executed_step_fraction": float(np.mean(fracs)), "not_found_cells": int(stats["not_found_cells"]),
                      "resolver_chunks": int(stats["resolver_chunks"]), "cre_per_sec": units / (np.median(ms_all) * 1e-3),
                      "fitness_checksum": int(gen.h_out.sum())})
        gen.sim.close()

    for seed in (4321, 99, 2024, 7):
        pr = synth.make_problem(N, C, P, net_seed=seed, cell_seed=5678 + seed, pop_seed=91011 + seed)
        bm = np.zeros((N, C), np.uint8); bm[pr.node_positions] = pr.bin_rows
        run_case("network seed %d" % seed, pr, bm)
    pr, bm = build_problem(wl)
    pr.knockouts = np.zeros(N, np.int32); pr.knockins = np.zeros(N, np.int32)
    run_case("default network, KO/KI all zero", pr, bm)
    pr, bm = build_problem(wl)
    fresh = [synth.make_population(pr.net, P, seed=555 + k) for k in range(3)]
    run_case("default network, fresh population every step", pr, bm, fresh=fresh)
    rings = (45, 50, 57) if N >= 180 else (N // 4, N // 3)
    chain = N - 1 - sum(rings) - 4
    if chain >= 2:
        pr = synth.make_gated_rings(N, C, P, ring_lens=rings, chain_len=chain)
        bm = np.zeros((N, C), np.uint8); bm[pr.node_positions] = pr.bin_rows
        run_case("adversarial: gated rings of period %s (every tile runs 99 steps and goes to the exact resolver)" %
                 "/".join(str(2 * r) for r in rings), pr, bm)
    return cases


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
    pr, bm = build_problem(wl)
    Pglobal, N, C = wl["population"], wl["nodes"], wl["cells"]
    full_inds, full_an, full_ai = pr.individuals, pr.pop_and_nodes, pr.pop_and_inv
    strong = args.scaling == "strong"
    shard, width, costs_note = None, Pglobal, None
    if world > 1 and strong:
        # fixed global population (BASELINE config 3: "population 500, GA sweep at 1/2/4/8"): the individuals are
        # dealt to the ranks by the cost a first, untimed generation measured on rank 0 (a (mu + lambda) GA scored
        # most of them in the previous generation; scb_pop_costs is that predictor)
        import torch
        costs = np.zeros(Pglobal, np.int64)
        if rank == 0:
            g0 = Generation(pr, bm, local_rank, args.opt)
            g0.upload(); g0.step_resident(); g0.sim.sync()
            costs[:] = g0.sim.costs()
            g0.sim.close()
        t = torch.from_numpy(costs).cuda()
        dist.broadcast(t, 0)
        costs = t.cpu().numpy()
        shards, width = lpt_shards(costs, world)
        shard = np.asarray(shards[rank], np.int64)
        loads = [int(costs[sh].sum()) for sh in shards]
        costs_note = {"predicted_load_max_over_mean": max(loads) / (sum(loads) / world)}
        pr.individuals, pr.pop_and_nodes, pr.pop_and_inv = full_inds[shard], full_an[shard], full_ai[shard]
    gen = Generation(pr, bm, local_rank, args.opt)
    sim, P = gen.sim, gen.P
    units_per_step = float(P if not (world > 1 and strong) else Pglobal) * C * N * STEPS_PER_CELL * (world if not strong else 1)

    def barrier():
        sim.sync()
        if world > 1:
            import torch
            torch.cuda.synchronize()
            dist.barrier()
            # The ranks leave an NCCL barrier up to a few hundred microseconds apart (host wake-ups), which the first
            # timed generation's exchange would then wait out.  All ranks of a box share CLOCK_MONOTONIC: leave together.
            t = torch.tensor([time.monotonic() + 0.02], dtype=torch.float64, device="cuda")        # generous: the first broadcast sets up channels
            dist.broadcast(t, 0)
            deadline = float(t.item())
            while time.monotonic() < deadline:
                pass

    # ---- the generation's exchange: fitness shards over peer memory (one kernel at the end of the graph), or NCCL
    gather_after, gather_kind, allfit = None, "none", None
    gen.upload()
    if world > 1:
        import torch
        if args.gather == "peer":
            handles = sim.gather_init(world, rank, width)
            hs = [None] * world
            dist.all_gather_object(hs, handles.tobytes())
            sim.gather_connect(np.frombuffer(b"".join(hs), np.uint8))
            gather_kind = "peer memory: scb_k_gather, last kernel of the generation's CUDA graph (P2P stores over NVLink + flags)"
        else:
            sim.run(simulator.MODE_SUM); sim.sync()
            ptr, _ = sim.result_device_pointer(simulator.MODE_SUM)

            class _Wrap:
                __cuda_array_interface__ = {"shape": (P,), "typestr": "<i8", "data": (ptr, False), "version": 3}
            local = torch.as_tensor(_Wrap(), device=torch.device("cuda", local_rank))
            pad = torch.zeros(width, dtype=torch.int64, device=local.device)
            allfit = torch.empty(world * width, dtype=torch.int64, device=local.device)
            ext = torch.cuda.ExternalStream(sim.stream_pointer(), device=local.device)

            def gather_after():
                with torch.cuda.stream(ext):
                    pad[:P].copy_(local)
                    dist.all_gather_into_tensor(allfit, pad)
            gather_kind = "NCCL all_gather_into_tensor issued on the library's stream after the graph"

    def step_resident():
        gen.step_resident()
        if gather_after:
            gather_after()

    def step_e2e():
        gen.upload()
        gen.sim.run(simulator.MODE_SUM)
        if gather_after:
            gather_after()
        gen.sim.fetch(simulator.MODE_SUM, out=gen.h_out)

    def reduce_max(ms_list):
        tot = float(np.sum(ms_list))
        if world > 1:
            import torch
            t = torch.tensor(ms_list + [tot], dtype=torch.float64, device="cuda")
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
            return tmax.cpu().numpy(), tmin.cpu().numpy()
        a = np.asarray(ms_list + [tot])
        return a, a

    # cost-ordered work list from the warm-up generation
    gen.step_resident(); sim.sync()
    if not args.no_order:
        sim.set_order(sim.cost_order())
    else:
        sim.set_option(_lib.OPT_AUTO_ORDER, 0)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res_ms, _, _ = timed_steps(gen, step_resident, args.steps, args.warmup, barrier)
    clocks = sampler.stop() if rank == 0 else None
    res_max, res_min = reduce_max(res_ms)
    total_ms = float(res_max[-1])                       # max over ranks of the timed region's device time
    e2e_ms_list, _, _ = timed_steps(gen, step_e2e, args.steps, max(args.warmup, 3), barrier)
    e2e_max, _ = reduce_max(e2e_ms_list)
    e2e_ms = float(e2e_max[-1])
    fit = gen.h_out.copy()
    e2e_compact_ms, e2e_resident_ms = None, None
    if world == 1:
        c_ms, _, _ = timed_steps(gen, gen.step_e2e_compact, args.steps, max(args.warmup, 3), barrier)
        e2e_compact_ms = float(np.sum(c_ms))
        if not np.array_equal(fit, gen.h_out):
            raise RuntimeError("compact and table uploads disagree")
        # resident population: half of the slots replaced per generation (compact upload still in place)
        half = P // 2
        gen.slots_half = np.arange(P - half, P, dtype=np.int32)
        r_ms, _, _ = timed_steps(gen, gen.step_e2e_resident, args.steps, max(args.warmup, 3), barrier)
        e2e_resident_ms = float(np.sum(r_ms))
        if not np.array_equal(fit, gen.h_out):
            raise RuntimeError("resident-population update and full upload disagree")

    # ---- two contexts from two host threads: the upload of one population overlaps the kernels of the other (every
    # context has its own stream; a GA generation cannot overlap with its own successor, two GA runs can)
    e2e_overlap = None
    if world == 1 and args.workload in ("C2", "C3"):
        try:
            g2 = Generation(pr, bm, local_rank, args.opt)
            pair = (gen, g2)
            for g in pair:
                g.upload(); g.step_resident(); g.sim.sync()
                g.sim.set_order(g.sim.cost_order())
            reps = max(args.steps, 5)

            def drive(g):
                for _ in range(reps):
                    g.step_e2e()

            for g in pair:
                g.step_e2e()
            threads = [threading.Thread(target=drive, args=(g,)) for g in pair]
            t0 = time.perf_counter()
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            wall = time.perf_counter() - t0
            ok = bool(np.array_equal(g2.h_out, fit) and np.array_equal(gen.h_out, fit))
            g2.sim.close()
            e2e_overlap = {"ms_per_generation": wall / (2 * reps) * 1e3, "contexts": 2, "generations": 2 * reps, "results_identical": ok,
                           "value": units_per_step / (wall / (2 * reps)), "unit": UNIT,
                           "note": "wall clock, host buffers in the reference's table format: two contexts driven by two host threads, "
                                   "each generation = upload + kernels + fetch; the copies of one overlap the kernels of the other"}
        except Exception as exc:
            e2e_overlap = {"error": str(exc)}

    # ---- N > 1: the gathered vector must be the single-GPU answer
    verify = None
    if world > 1:
        import torch
        gen.step_resident(); 
        if gather_after:
            gather_after()
        sim.sync(); torch.cuda.synchronize()
        gathered = sim.gather_result() if args.gather == "peer" else allfit.cpu().numpy()
        local_fit = sim.fetch(simulator.MODE_SUM).astype(np.int64)
        ok_local = bool(np.array_equal(gathered[rank * width:rank * width + P], local_fit))
        full = np.zeros(Pglobal if strong else world * P, np.int64)
        if strong:
            for r, sh in enumerate(lpt_shards(costs, world)[0]):
                full[np.asarray(sh, np.int64)] = gathered[r * width:r * width + len(sh)]
        else:
            full[:] = gathered[:world * P]
        ok_single = None
        if rank == 0:
            # rank 0 scores the WHOLE population on its own GPU (untimed): the exchanged vector must be that answer
            prf, bmf = build_problem(wl)
            gf = Generation(prf, bmf, local_rank, args.opt)
            ref = gf.sim.evaluate_population(prf.individuals, prf.knockouts, prf.knockins, prf.pop_and_nodes, prf.pop_and_inv)
            gf.sim.close()
            ok_single = bool(np.array_equal(full, ref if strong else np.tile(ref, world)))
        flags = [None] * world
        dist.all_gather_object(flags, ok_local)
        verify = {"every_rank_finds_its_shard_in_the_gathered_vector": all(flags),
                  "gathered_vector_equals_single_gpu_evaluation": ok_single,
                  "global_fitness_checksum": int(full.sum())}
        if rank == 0 and not (all(flags) and ok_single):
            raise RuntimeError("multi-GPU fitness vector differs from the single-GPU evaluation: %s" % verify)

    # ---- N > 1, strong scaling: the weak figure beside it (every rank scores its own copy of the whole population,
    # no exchange: the reference's "one job per network" at full size on every GPU)
    weak = None
    if world > 1 and strong:
        try:
            prw, bmw = build_problem(wl)
            gw = Generation(prw, bmw, local_rank, args.opt)
            gw.upload(); gw.step_resident(); gw.sim.sync()
            gw.sim.set_order(gw.sim.cost_order())
            w_ms, _, _ = timed_steps(gw, gw.step_resident, args.steps, args.warmup, barrier)
            w_max, _ = reduce_max(w_ms)
            gw.sim.close()
            weak = {"ms_per_step": float(w_max[-1]) / args.steps, "population_per_gpu": Pglobal,
                    "value": float(world) * Pglobal * C * N * STEPS_PER_CELL / (float(w_max[-1]) / args.steps * 1e-3), "unit": UNIT,
                    "note": "weak scaling: the full population on every GPU, no exchange; max over ranks"}
        except Exception as exc:                      # an extra, never a reason to lose the main line
            weak = {"error": str(exc)}

    # ---- roofline pass: K2 on its own (plain launches, resolver not fused, so that the duration is K2's own)
    gen.upload()
    sim.set_option(_lib.OPT_GRAPH, 0)
    fused_default = "fuse=0" not in args.opt
    if fused_default:
        sim.set_option(_lib.OPT_FUSE, 0)
    gsave, gather_after = gather_after, None
    _, stats, k2_list = timed_steps(gen, gen.step_resident, min(args.steps, 5), 2, sim.sync, stats_each=True)
    stages = None
    if fused_default:
        sim.set_option(_lib.OPT_FUSE, 1)
        # the same generation as plain launches with the resolver inside K2: where the time of the graph goes
        _, sf, _ = timed_steps(gen, gen.step_resident, 3, 1, sim.sync, stats_each=True)
        stages = {k: round(float(sf[k]), 4) for k in ("kernel_ms", "clear_ms", "compile_ms", "sim_ms", "resolve_ms", "fill_ms", "gather_ms")}
        stages["note"] = "plain launches (SCB_OPT_GRAPH = 0), resolver inside K2, events between the launches; last of 3 steps"
    sim.set_option(_lib.OPT_GRAPH, 1)
    gather_after = gsave

    if rank == 0:
        hbm_peak, sm_max, peak_kind = load_peaks()
        ms_per_step = total_ms / args.steps
        value = units_per_step / (ms_per_step * 1e-3)
        info = sim.info()
        k2_ms = float(np.mean(k2_list))
        per_gpu_units = float(P) * C * N * STEPS_PER_CELL
        # algorithmic HBM bytes of one K2 launch: the packed matrix N*C/8, one compiled program per individual
        # (16 B per node) and the fitness; the reduced programs / program images K0 also emits are this implementation's
        # own traffic and show up in `traffic`, not here
        alg_bytes = N * C / 8.0 + P * N * 16.0 + P * 8.0
        hbm_ach = alg_bytes / (k2_ms * 1e-3) / 1e9
        traffic = None
        try:        # dram__bytes_read.sum + dram__bytes_write.sum of K2 from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "r2", "k2_dram_traffic.json")) as f:
                t = json.load(f)
            if t.get("workload") == args.workload:
                traffic = t["dram_bytes_per_launch"]
        except Exception:
            pass
        sm_mhz = (clocks or {}).get("sm_mhz") or sm_max
        smem_bytes_exec = stats["state_rows_moved"] * 128.0 * info["words_per_lane"]
        smem_peak = 148 * 128.0 * sm_mhz * 1e6 / 1e9
        smem_ach = smem_bytes_exec / (k2_ms * 1e-3) / 1e9
        try:
            mb = {k: sim.microbench(k) for k in ("lop3", "popc", "lds")}
        except Exception as exc:
            mb = {"error": str(exc)}
        smem_peak_nominal = smem_peak
        if mb.get("lds"):
            smem_peak = mb["lds"] / 1e9
        lop_alg = per_gpu_units * 0.21875 / (k2_ms * 1e-3)
        lop_exec = stats["steps_executed"] * N * 32.0 * info["words_per_lane"] / (k2_ms * 1e-3)
        cpu, cpu1 = None, None
        if not args.no_cpu_baseline and world == 1:      # the reported CPU baseline is an N = 1 item (the other ranks would wait ~4 s in the final barrier)
            try:
                prc, _ = build_problem(wl)
                cores = min(len(os.sched_getaffinity(0)), 64)
                cells = min(C, 12_500)
                dt, u, kind = reference_sample(prc, cells, cores, cores)
                what = "the unmodified reference C (oracle/_ref)" if kind == "reference" else "the C restatement (oracle/scb_oracle.c)"
                cpu = {"value": u / dt, "unit": UNIT, "cores": cores, "kind": kind,
                       "sample": "%d individuals x %d cells x %d nodes, %d concurrent scSyncBool calls of %s, %.1f s" % (
                           cores, cells, N, cores, what, dt)}
                # the reference as shipped is single-threaded (the OpenMP pragma is commented out, simulator.c:366-368)
                c1 = min(C, 6_000)
                dt1, u1, kind1 = reference_sample(prc, c1, 1, 1)
                cpu1 = {"value": u1 / dt1, "unit": UNIT, "cores": 1, "kind": kind1,
                        "sample": "1 individual x %d cells x %d nodes, one scSyncBool call of %s, %.1f s" % (c1, N, what, dt1)}
            except Exception as exc:           # oracle/_ref absent on this box
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % exc}
        cases = None
        if world == 1 and not args.no_cases and args.workload in ("C2", "C3"):
            cases = robustness_cases(args, wl, local_rank)
        kernels_per_step = (3 if fused_default else 4) + (1 if (world > 1 and args.gather == "peer") else 0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling,
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": wl["name"], "population_global": Pglobal if (strong or world == 1) else world * P,
                       "population_per_gpu": P, "cells": C, "nodes": N,
                       "sharding": ("one GPU" if world == 1 else
                                    ("the %d individuals dealt to %d ranks by predicted cost (LPT), " % (Pglobal, world) if strong else
                                     "%d individuals per rank (same seeds on every rank), " % P) +
                                    "packed matrix replicated, fitness exchange: " + gather_kind),
                       "launch": "one CUDA graph per generation and rank (clears, K0, K2 + in-kernel resolver, fill-forward%s)" % (
                           ", peer exchange" if (world > 1 and args.gather == "peer") else ""),
                       "work_order": "tiles handed out costliest individual first (costs of the previous generation; the robustness cases: the cost K0 predicts)" if not args.no_order else "index order",
                       "l2": "flushed (256 MiB memset) between timed steps", "geometry": info,
                       "clock_sampling": "nvidia-smi every 100 ms on rank 0 during the resident timed region",
                       "options": args.opt},
            "ga_generations_per_sec": 1e3 / ms_per_step,
            "per_step_ms": {"max_over_ranks": spread(res_max[:-1]), "min_over_ranks": spread(res_min[:-1]),
                            "every_step_max_over_ranks": [round(float(x), 4) for x in res_max[:-1]],
                            "note": "device time of every timed generation incl. the exchange; max - min over ranks at a step = wait for the slowest rank"},
            "e2e": {"value": units_per_step / (e2e_ms / args.steps * 1e-3), "unit": UNIT,
                    "ms_per_step": e2e_ms / args.steps, "h2d_bytes_per_step": int(gen.h2d), "d2h_bytes_per_step": int(gen.d2h)},
            "gpu_launches": kernels_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                         "frac": hbm_ach / hbm_peak, "traffic": traffic, "algorithmic_bytes": alg_bytes,
                         "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                         "peak_source": peak_kind,
                         "kernel": "scb_k_sim<%d>" % info["words_per_lane"], "kernel_ms": k2_ms,
                         "kernel_timing": "CUDA events around K2 on the library's stream (plain launches, resolver not fused, so that the duration is K2's own)",
                         "note": "the path is shared-memory / latency bound, not HBM bound (SURVEY 8d): see roofline_binding"},
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
            "stats": stats, "stages_ms": stages, "clocks": clocks, "cpu_baseline": cpu, "cpu_baseline_1thread": cpu1,
            "fitness_checksum": int(fit.sum()),
        }
        if e2e_compact_ms is not None:
            line["e2e_compact"] = {"value": units_per_step / (e2e_compact_ms / args.steps * 1e-3), "unit": UNIT,
                                   "ms_per_step": e2e_compact_ms / args.steps, "h2d_bytes_per_step": int(gen.h2d_compact),
                                   "d2h_bytes_per_step": int(gen.d2h),
                                   "note": "same generation, population handed over as bit-strings + upstream triples "
                                           "(scb_pop_upload_compact) instead of the reference's [N][7][3] tables"}
        if e2e_resident_ms is not None:
            half = P // 2
            line["e2e_resident"] = {"value": float(half) * C * N * STEPS_PER_CELL / (e2e_resident_ms / args.steps * 1e-3), "unit": UNIT,
                                    "ms_per_step": e2e_resident_ms / args.steps, "individuals_scored_per_step": half,
                                    "h2d_bytes_per_step": int(gen.h2d_compact * half // P), "d2h_bytes_per_step": int(gen.d2h),
                                    "note": "(mu + lambda) generation on a resident population (scb_pop_update): only the "
                                            "offspring half is uploaded, compiled and simulated; value counts only those"}
        if e2e_overlap is not None:
            line["e2e_two_contexts"] = e2e_overlap
        if weak is not None:
            line["weak_scaling"] = weak
        if costs_note:
            line["shard_balance"] = costs_note
        if verify:
            line["multi_gpu_check"] = verify
        if cases:
            worst = max(c["ms_per_generation"]["max"] for c in cases)
            line["robustness"] = {"cases": cases, "worst_ms_per_generation": worst,
                                  "target_ms": 10.0, "worst_case_meets_target": bool(worst < 10.0)}
        print(json.dumps(line), flush=True)
    if world > 1 and args.gather == "peer":
        barrier()
        sim.gather_close()
    sim.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# C4: ~50 independent KEGG-sized pathways over the GPUs of a box, no collective until the end
# ---------------------------------------------------------------------------------------------
C4_SIZES = (47, 60, 100, 159, 200, 300)          # packaged hsa00010 (47), C2 (60), hsa04066 (159), C3 (200), hsa04010 cut to 300
C4_PATHWAYS, C4_CELLS, C4_POP = 48, 60_000, 500


def run_c4(args, rank, world, local_rank):
    """BASELINE config 4 (GSE198339-shaped: ~60k cells, ~50 pathways): what the reference runs as one SLURM job per
    network (pipeline.py:12-75).  Every pathway is its own context (own packed matrix, own stream) on the GPU the
    longest-processing-time assignment gives it (scbonita_b200.distributed.assign_pathways); a step scores one GA
    generation of 500 individuals for every pathway; nothing is exchanged until the fitness vectors are collected at
    the end."""
    from scbonita_b200 import _lib, simulator
    from scbonita_b200.distributed import assign_pathways, pathway_cost
    lib = _lib.load()
    if lib.scb_device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists)")
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sizes = [C4_SIZES[i % len(C4_SIZES)] for i in range(C4_PATHWAYS)]
    costs = [pathway_cost(n, C4_CELLS, C4_POP) for n in sizes]
    mine = assign_pathways(costs, world)[rank]
    base = {}                                        # one network + population per size, a different cell matrix per pathway
    gens = []
    for i in mine:
        n = sizes[i]
        if n not in base:
            base[n] = synth.make_problem(n, 8, C4_POP)
        b = base[n]
        rows, pos = synth.make_cells(n, C4_CELLS, seed=5678 + i)
        pr = synth.SynthProblem(b.net, rows, pos, b.individuals, b.pop_and_nodes, b.pop_and_inv, b.knockouts, b.knockins)
        bm = np.zeros((n, C4_CELLS), np.uint8); bm[pos] = rows
        g = Generation(pr, bm, local_rank, args.opt)
        g.upload()
        gens.append(g)
    units = float(sum(costs))                        # cell-rule evaluations of one step (all pathways, all ranks)

    def barrier():
        for g in gens:
            g.sim.sync()
        if world > 1:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    def step_resident():
        for g in gens:
            g.step_resident()                        # asynchronous: the pathways' launches overlap on the GPU

    def step_e2e():
        for g in gens:
            g.upload(); g.sim.run(simulator.MODE_SUM)
        for g in gens:
            g.sim.fetch(simulator.MODE_SUM, out=g.h_out)

    def timed(step, steps, warmup):
        import time as _t
        for _ in range(warmup):
            step()
        barrier()
        # wall clock around device-synchronised steps: the pathways run on their own streams, one CUDA event pair cannot
        # bracket them all
        t0 = _t.perf_counter()
        for _ in range(steps):
            step()
            for g in gens:
                g.sim.sync()
        dt = (_t.perf_counter() - t0) * 1e3
        barrier()
        if world > 1:
            import torch
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = timed(step_e2e, args.steps, max(args.warmup, 3))
    local = {i: int(g.h_out.sum()) for i, g in zip(mine, gens)}
    sums = [local]
    if world > 1:
        sums = [None] * world
        dist.all_gather_object(sums, local)          # the only communication of C4: the results, at the end
    if rank == 0:
        merged = {}
        for part in sums:
            merged.update(part)
        ms = total_ms / args.steps
        cpu = None
        if not args.no_cpu_baseline:
            try:
                cores = min(len(os.sched_getaffinity(0)), 64)
                prc = synth.make_problem(100, 12_500, cores)
                dt, u, kind = reference_sample(prc, 12_500, cores, cores)
                cpu = {"value": u / dt, "unit": UNIT, "cores": cores, "kind": kind,
                       "sample": "%d individuals x 12500 cells x 100 nodes (one of the six pathway sizes), %d concurrent scSyncBool calls, %.1f s" % (cores, cores, dt)}
            except Exception as exc:
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % exc}
        h2d = sum(g.h2d for g in gens); d2h = sum(g.d2h for g in gens)
        line = {"metric": METRIC, "value": units / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": {"workload": "C4: %d pathways (%s nodes) x %d cells x population %d, pathways spread over %d GPU(s) longest first, "
                                       "no collective until the final gather of the results" % (C4_PATHWAYS, "/".join(map(str, C4_SIZES)), C4_CELLS, C4_POP, world),
                           "pathways_on_rank0": len(mine), "timing": "wall clock around device-synchronised steps, max over ranks",
                           "options": args.opt},
                "e2e": {"value": units / (e2e_ms / args.steps * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "note": "bytes of rank 0's pathways"},
                "gpu_launches": 3 * len(mine) * args.steps,
                "roofline": {"bound": "hbm", "achieved": None, "peak": load_peaks()[0], "unit": "GB/s", "frac": None, "traffic": None,
                             "note": "per-kernel roofline: see the C3 line (same kernels); C4 measures the pathway-level sharding"},
                "clocks": clocks, "cpu_baseline": cpu, "fitness_checksum": int(sum(merged.values())), "pathways_scored": len(merged)}
        print(json.dumps(line), flush=True)
    for g in gens:
        g.sim.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# L0: the level-0 drop-in: per-call cost of the exported scSyncBool symbol
# ---------------------------------------------------------------------------------------------
def run_l0(args, rank, world, local_rank):
    """What a user gets who only swaps ./simulator.so: one scSyncBool call per individual, 17 x c_void_p, dense int32
    matrix with the reference's row stride handed over on every call (ruleMaker.py:1109-1219).  Times the exported
    symbol of this library (context cache, upload, K0 + K2 + K3, result copy) and, on the same arguments, the unmodified
    reference C (oracle/_ref) on one host core -- the reference's own per-call cost without its Python marshalling."""
    import ctypes
    import time as _t
    from scbonita_b200 import _lib
    lib = _lib.load()
    if lib.scb_device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists)")
    if rank != 0:
        return
    V = ctypes.c_void_p
    out = []
    for name, n, cells, stride in (("C1 shape: 47 nodes x 100 cells", 47, 100, 15000), ("200 nodes x 10000 cells", 200, 10000, 15000)):
        pr = synth.make_problem(n, cells, 4)
        bm = np.zeros((int(pr.node_positions.max()) + 1, stride), np.int32)
        bm[pr.node_positions, :cells] = pr.bin_rows
        vals = np.zeros((100, 20000), np.int32)                  # simData[100][NODE] of the reference build
        lib.scb_set_legacy_geometry(20000, stride)
        fn = lib.scSyncBool
        fn.restype, fn.argtypes = None, None
        al, parse, ko, ki, pos = (np.ascontiguousarray(a, np.int32) for a in (pr.net.and_len, pr.net.parse, pr.knockouts, pr.knockins, pr.node_positions))

        def call(f, i, errors):
            ind = np.ascontiguousarray(pr.individuals[i], np.int32)
            an = np.ascontiguousarray(pr.pop_and_nodes[i], np.int32)
            ai = np.ascontiguousarray(pr.pop_and_inv[i], np.int32)
            f(V(vals.ctypes.data), V(ind.ctypes.data), V(len(ind)), V(n), V(al.ctypes.data), V(parse.ctypes.data), V(an.ctypes.data),
              V(ai.ctypes.data), V(100), V(ko.ctypes.data), V(ki.ctypes.data), V(cells), V(bm.ctypes.data), V(pos.ctypes.data),
              V(errors.ctypes.data), V(0), V(0))

        e_ours, e_ref = np.zeros(cells, np.int32), np.zeros(cells, np.int32)
        for i in range(4):
            call(fn, i, e_ours)                                  # warm-up: context built and cached on the first call
        reps = max(args.steps, 5) * 4
        t0 = _t.perf_counter()
        for r in range(reps):
            call(fn, r % 4, e_ours)
        ours_ms = (_t.perf_counter() - t0) / reps * 1e3
        ref_ms = None
        try:
            import oracle                                        # the checker / CPU baseline, not the product
            path = os.path.join(oracle.REF_DIR, oracle.REF_CELL_VARIANTS[stride])
            ref = ctypes.cdll.LoadLibrary(path) if os.path.exists(path) else None
            if ref is not None:
                rf = ref.scSyncBool
                rf.restype, rf.argtypes = None, None
                call(rf, 3, e_ref)
                same = bool(np.array_equal(e_ref, e_ours))
                t0 = _t.perf_counter()
                nref = 2 if cells > 1000 else 8
                for r in range(nref):
                    call(rf, r % 4, e_ref)
                ref_ms = (_t.perf_counter() - t0) / nref * 1e3
            else:
                same = None
        except Exception as exc:                                 # oracle/_ref not built for this stride
            same, ref_ms = None, None
            sys.stderr.write("reference build unavailable for the level-0 comparison: %s\n" % exc)
        out.append({"shape": name, "row_stride": stride, "ours_ms_per_call": ours_ms, "reference_c_ms_per_call_one_core": ref_ms,
                    "identical_errors": same, "cre_per_call": float(cells) * n * STEPS_PER_CELL})
    lib.scb_set_legacy_geometry(20000, 15000)
    big = out[-1]
    line = {"metric": METRIC, "value": big["cre_per_call"] / (big["ours_ms_per_call"] * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": big["ours_ms_per_call"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": "L0: level-0 drop-in, one exported scSyncBool call per individual (wall clock per call, host buffers in and out)"},
            "calls": out, "e2e": {"value": big["cre_per_call"] / (big["ours_ms_per_call"] * 1e-3), "unit": UNIT, "ms_per_step": big["ours_ms_per_call"],
                                  "h2d_bytes_per_step": 200 * 21 * 4 * 2, "d2h_bytes_per_step": 10000 * 4},
            "gpu_launches": 3, "roofline": {"bound": "hbm", "achieved": None, "peak": load_peaks()[0], "unit": "GB/s", "frac": None, "traffic": None},
            "cpu_baseline": None}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# C5: attractor analysis, 100k cells x 300-node network
# ---------------------------------------------------------------------------------------------
def run_c5(args, rank, world, local_rank):
    """BASELINE config 5 (singleCell.assignAttractors, singleCell.py:1162-1244): per-cell trajectories + attractor
    windows in the reference's Python semantics (K4), the unique attractor states, then the cells x attractors Hamming
    distance with first-minimum decider (K5).  One GPU (the path is one pathway's post-processing); replicas at N > 1."""
    from scbonita_b200 import _lib, simulator
    import time as _t
    lib = _lib.load()
    if lib.scb_device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback exists)")
    N, C = 300, 100_000
    pr = synth.make_problem(N, C, 1, knock=False)
    bm = np.zeros((N, C), np.uint8); bm[pr.node_positions] = pr.bin_rows
    model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.pop_and_nodes[0], pr.pop_and_inv[0], pr.node_positions, pr.net.ind_len)
    sim = simulator.Simulator(bm, model, device=local_rank)
    ind = pr.individuals[0]
    attrs = sim.attractor_set(ind)
    A = int(attrs.shape[0])

    def step():
        a = sim.attractor_set(ind, packed=True)
        sim.hamming_argmin(a, want_dist=False)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_traj = t_set = t_ham = 0.0
    for _ in range(args.steps):
        t0 = _t.perf_counter(); sim._attractors_packed(ind, None, None, "python", 10)
        t1 = _t.perf_counter(); a = sim.attractor_set(ind, packed=True)
        t2 = _t.perf_counter(); sim.hamming_argmin(a, want_dist=False)
        t3 = _t.perf_counter()
        t_traj += t1 - t0; t_set += t2 - t1; t_ham += t3 - t2
    clocks = sampler.stop()
    K = args.steps
    cre = float(C) * N * 9                           # Python semantics: 10 states = 9 synchronous steps per cell
    ham = float(C) * A * N
    try:
        mb = {k: sim.microbench(k) for k in ("popc",)}
    except Exception:
        mb = {}
    popc_peak = mb.get("popc")
    ms = (t_set + t_ham) / K * 1e3
    line = {"metric": METRIC, "value": cre / (t_set / K), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": "C5: attractor analysis, 100k cells x 300-node network (trajectories + windows + unique attractor states + Hamming decider)",
                       "attractors": A, "timing": "wall clock through the host API (host buffers in, results out): every stage is an e2e number"},
            "stages_ms": {"trajectories_windows_states_to_host": t_traj / K * 1e3, "attractor_set (trajectories + windows + device dedupe, unique states to host)": t_set / K * 1e3,
                          "hamming_decider": t_ham / K * 1e3},
            "e2e": {"value": cre / (t_set / K), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": int(A * ((N + 31) // 32) * 4),
                    "d2h_bytes_per_step": int(A * ((N + 31) // 32) * 4 + C * 8)},
            "hamming": {"bit_compares_per_sec": ham / (t_ham / K), "popc_peak_ops_per_sec": popc_peak,
                        "frac_of_popc_roof": (ham / 32.0 / (t_ham / K)) / popc_peak if popc_peak else None,
                        "note": "one POPC per 32 node bits; wall clock of the call: H2D of the packed attractors, kernel, D2H of decider + distance"},
            "gpu_launches": 4 * K, "roofline": {"bound": "hbm", "achieved": None, "peak": load_peaks()[0], "unit": "GB/s", "frac": None, "traffic": None},
            "clocks": clocks, "cpu_baseline": None}
    if rank == 0:
        print(json.dumps(line), flush=True)
    sim.close()


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
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS) + ["C4", "C5", "L0"])
    ap.add_argument("--opt", action="append", default=[], help="library option, e.g. early_exit=0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = the workload's population split over the GPUs (BASELINE config 3: "
                         "'population 500, GA sweep at 1/2/4/8'); weak = the whole population on every GPU")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: the generation's fitness exchange -- peer memory kernel (default) or NCCL all-gather")
    ap.add_argument("--no-order", action="store_true", help="hand tiles out in index order instead of costliest first")
    ap.add_argument("--no-cases", action="store_true", help="skip the robustness cases (other seeds, adversarial network)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "C4" and args.impl == "ours":
        return run_c4(args, rank, world, local_rank)
    if args.workload == "C5" and args.impl == "ours":
        return run_c5(args, rank, world, local_rank)
    if args.workload == "L0" and args.impl == "ours":
        return run_l0(args, rank, world, local_rank)
    wl = WORKLOADS["C3" if args.workload in ("C4", "C5", "L0") else args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
    else:
        run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
