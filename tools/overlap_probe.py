#!/usr/bin/env python
"""Probe: do two batches on two slots (two streams) finish sooner side by side than one after the other? Tells whether the HBM-bound
traceback of one batch hides under the issue-bound forward pass of another.   python tools/overlap_probe.py [cfg4|cfg2]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ubu_b200 as U
from ubu_b200 import workloads as WL

which = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
mk = {"cfg4": lambda s: WL.cfg4(n, tie_stress=0, seed=s), "cfg2": lambda s: WL.uniform("cfg2", n, 76, 500, seed=s, paired=True)}[which]
w0, w1 = mk(11), mk(12)
with U.SwAligner(device=0, slots=2) as al:
    al.stage(0, w0.reads, w0.windows, w0.pair_ref_idx); al.stage(1, w1.reads, w1.windows, w1.pair_ref_idx); al.sync(0); al.sync(1)
    for _ in range(2):
        al.run(0); al.run(1); al.sync(0); al.sync(1)
    def timed(f, reps=5):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        return 1e3 * float(np.median(ts))
    def serial():
        al.run(0); al.sync(0); al.run(1); al.sync(1)
    def side():
        al.run(0); al.run(1); al.sync(0); al.sync(1)
    s = timed(serial); p = timed(side)
    print(f"{which} 2 x {n}: one after the other {s:.2f} ms, side by side {p:.2f} ms, ratio {p / s:.3f}; last_run_ms slot0 {al.last_run_ms(0)}")
