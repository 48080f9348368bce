"""Race detection without a GPU: the kernel sources run under the host-thread emulator built with
-fsanitize=thread (every CUDA thread is a real thread, __syncthreads a real barrier), so a missing barrier or an
unsynchronised shared-memory access shows up as a ThreadSanitizer report.  (compute-sanitizer is closed on the
GPU pool.)"""
import os
import subprocess
import sys

import pytest

from common import ROOT

CASE = r'''
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from common import Problem, ring_problem, tables
from scbonita_b200 import _lib, simulator
LIB = %(lib)r
for pr in (Problem.synthetic(20, 1500, 3), ring_problem([7, 11], 6, 1500)):
    sim = simulator.Simulator(pr.bm, tables(pr.net, pr.node_positions), lib_path=LIB)
    sim.set_option(_lib.OPT_MAX_CTAS, 2); sim.set_option(_lib.OPT_WARPS, 4)
    pr.check(sim, modes=(0, 1, 2))
    sim.close()
print("tsan-case-done")
'''


def test_kernels_are_race_free_under_threadsanitizer():
    tsan_rt = subprocess.run(["/usr/bin/gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(tsan_rt) or not os.path.exists(tsan_rt):
        pytest.skip("libtsan not available")
    emu = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-C", emu, "tsan"], stdout=subprocess.DEVNULL)
    lib = os.path.join(emu, "_build", "libscbonita_emu_tsan.so")
    env = dict(os.environ, LD_PRELOAD=tsan_rt, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4")
    out = subprocess.run([sys.executable, "-c", CASE % dict(root=ROOT, lib=lib)], env=env, capture_output=True,
                         text=True, timeout=1500)
    assert "tsan-case-done" in out.stdout, out.stderr[-3000:]
    assert "WARNING: ThreadSanitizer" not in out.stderr, out.stderr[:6000]
