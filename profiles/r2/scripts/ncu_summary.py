"""Summaries of the round-2 ncu captures (gpurun_out/<tag>_*) into profiles/r2/: the launch list of the bench command
(per kernel: launches, total and mean duration, share), the headline metrics of the full captures of K2 and K0, and
k2_dram_traffic.json (what bench.py reports as roofline.traffic)."""
import csv, json, subprocess, sys, os, collections

tag = sys.argv[1]
out_dir = "profiles/r2"
rows = list(csv.reader(l for l in open("gpurun_out/%s_launches.csv" % tag) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * (1.0 if r[ui] in ("us", "usecond") else 1e3)
    a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(out_dir, "launch_list_summary.md"), "w") as f:
    f.write("Launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-cases` under\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold caches, serialised: shares, not absolutes).\n\n")
    f.write("| kernel | launches | total us | mean us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f | %.1f %% |\n" % (k, n, t, t / n, 100 * t / tot))
print(open(os.path.join(out_dir, "launch_list_summary.md")).read())

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio"]
traffic = {}
for name in sys.argv[2:]:
    rep = "gpurun_out/%s.ncu-rep" % name
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units, vals = rr[0], rr[1], rr[2]
    with open(os.path.join(out_dir, "ncu_%s.csv" % name), "w") as f:
        f.write("metric,unit,value\n")
        f.write("Kernel Name,,\"%s\"\n" % vals[h.index("Kernel Name")])
        for w in WANT:
            if w in h:
                f.write("%s,%s,%s\n" % (w, units[h.index(w)], vals[h.index(w)]))
    d = {w: (vals[h.index(w)], units[h.index(w)]) for w in WANT if w in h}
    print(name, vals[h.index("Kernel Name")], {k: v for k, v in d.items() if "stalled" not in k})
    def to_bytes(v, u):
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    traffic[name] = {"kernel": vals[h.index("Kernel Name")], "dram_bytes_read": to_bytes(*d["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(*d["dram__bytes_write.sum"]),
                     "duration_under_ncu": " ".join(d["gpu__time_duration.sum"])}
if traffic:
    k2 = traffic.get("%s_k2" % tag)
    tot = (k2["dram_bytes_read"] + k2["dram_bytes_write"]) if k2 else None
    json.dump({"source": "ncu --set full --clock-control none, one launch of K2 of the C3 generation (plain launch, resolver not fused)",
               "workload": "C3", "dram_bytes_per_launch": tot, "k2_dram_bytes_per_launch": tot,
               "note": "ncu flushes the caches before the profiled launch, so the 22.6 MB of program images K0 has just written "
                       "(6 images x 7.5 KB x 500 individuals, L2-resident in a live run) are counted as DRAM reads; the algorithmic "
                       "figure is 2.85 MB (matrix + one 16-byte descriptor per node and individual + results)",
               "captures": traffic},
              open(os.path.join(out_dir, "k2_dram_traffic.json"), "w"), indent=1)
    # before / after table of K2
    def load(name):
        out = {}
        path = os.path.join(out_dir, "ncu_%s.csv" % name)
        if os.path.exists(path):
            for r in list(csv.reader(open(path)))[1:]:
                if len(r) >= 3:
                    out[r[0]] = (r[2], r[1])
        return out
    mid, head = load("r1kernel_k2"), load("%s_k2" % tag)
    r1 = {"gpu__time_duration.sum": ("2.758208", "ms"), "dram__bytes_read.sum": ("10.355712", "Mbyte"), "smsp__inst_executed.sum": ("1621452466", "inst"),
          "smsp__issue_active.avg.pct_of_peak_sustained_active": ("54.130978", "%"), "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": ("459698450", ""),
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": ("57.915651", "%"), "launch__registers_per_thread": ("89", "register/thread"),
          "sm__warps_active.avg.pct_of_peak_sustained_active": ("31.245146", "%"), "sm__throughput.avg.pct_of_peak_sustained_elapsed": ("51.094344", "%")}
    with open(os.path.join(out_dir, "k2_before_after.md"), "w") as f:
        f.write("K2 (`scb_k_sim<4, true>`) on the C3 generation, one launch, `ncu --set full --clock-control none`, resolver not fused.\n"
                "Columns: the kernel in the middle of round 1 (round-1 sources rebuilt and captured on a round-2 box), the kernel at the end of\n"
                "round 1 (`profiles/r1/ncu_raw_scb_k_sim4_round1_final.txt`, the metrics kept there), HEAD of round 2.\n"
                "Durations under ncu are cold-cache, serialised replays; the live numbers are in bench.py's line.\n\n")
        f.write("| metric | mid round 1 | end of round 1 | round 2 (HEAD) |\n|---|---:|---:|---:|\n")
        for w in WANT:
            if w in mid or w in head or w in r1:
                f.write("| %s | %s | %s | %s |\n" % (w, " ".join(mid.get(w, ("", ""))), " ".join(r1.get(w, ("", ""))), " ".join(head.get(w, ("", "")))))
    print(open(os.path.join(out_dir, "k2_dram_traffic.json")).read())
