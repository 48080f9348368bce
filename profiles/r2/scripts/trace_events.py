"""Timeline of one K2 launch (measurement build, see trace_steps.py): every tile and every resolver item with the CTA
that ran it and its start / end on the global timer.  Answers: what ends the launch -- tiles, items, or waiting?"""
import ctypes, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from scbonita_b200 import _lib, simulator
import bench

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "C3"]
pr, bm = bench.build_problem(wl)
model = simulator.ModelTables(pr.net.and_len, pr.net.parse, pr.net.and_nodes, pr.net.and_inv, pr.node_positions, pr.net.ind_len)
sim = simulator.Simulator(bm, model)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    sim.set_option(getattr(_lib, "OPT_" + k.upper()), int(v))
sim.evaluate_population(pr.individuals, pr.knockouts, pr.knockins, pr.pop_and_nodes, pr.pop_and_inv)
sim.set_order(sim.cost_order())
for _ in range(2):
    sim.evaluate_population(pr.individuals, pr.knockouts, pr.knockins, pr.pop_and_nodes, pr.pop_and_inv)
ev = np.zeros(16384 * 4, np.uint32)
sim.lib.scb_debug_events.argtypes = [ctypes.c_void_p] * 2
n = sim.lib.scb_debug_events(sim._ctx, ctypes.c_void_p(ev.ctypes.data))
ev = ev.reshape(-1, 4)[:n].astype(np.int64)
kind, cta = ev[:, 0] & 0xff, ev[:, 0] >> 8
t0 = ev[:, 1].min()
a, b = ((ev[:, 1] - t0) & 0xffffffff) / 1e3, ((ev[:, 2] - t0) & 0xffffffff) / 1e3     # microseconds
tiles, items = kind == 0, kind != 0
steps = ev[:, 3] & 0xff
unproven = (ev[:, 3] >> 31) & 1
print("events", n, "tiles", int(tiles.sum()), "resolver items", int(items.sum()), "(in the tail: %d)" % int((kind == 2).sum()))
print("launch span %.1f us; last tile ends %.1f us; last item ends %.1f us" % (b.max(), b[tiles].max(), b[items].max() if items.any() else 0))
d = b - a
for lo, hi in ((0, 10), (10, 30), (30, 60), (60, 98), (98, 100)):
    m = tiles & (steps > lo) & (steps <= hi)
    if m.any():
        print("tiles ending at steps (%d, %d]: %d, %.1f us mean, %.1f max" % (lo, hi, int(m.sum()), d[m].mean(), d[m].max()))
print("tiles left unproven (-> resolver):", int((tiles & (unproven == 1)).sum()), " first ends %.1f us, last ends %.1f us" %
      ((b[tiles & (unproven == 1)].min(), b[tiles & (unproven == 1)].max()) if (tiles & (unproven == 1)).any() else (0, 0)))
if items.any():
    print("items: %.1f us mean, %.1f min, %.1f max; first starts %.1f us, last starts %.1f us" % (d[items].mean(), d[items].min(), d[items].max(), a[items].min(), a[items].max()))
busy = np.zeros(int(cta.max()) + 1)
np.add.at(busy, cta, d)
end = np.zeros_like(busy)
np.maximum.at(end, cta, b)
print("CTAs %d: busy %.1f us mean (%.1f min, %.1f max) of %.1f; CTA end times: %.1f min, %.1f median" % (len(busy), busy.mean(), busy.min(), busy.max(), b.max(), end.min(), np.median(end)))
order = np.argsort(b)[-12:]
print("last events:", [(("tile", "item", "tail-item")[int(kind[i])], int(cta[i]), round(float(a[i]), 1), round(float(b[i]), 1), int(steps[i])) for i in order])
print(sim.last_stats)
