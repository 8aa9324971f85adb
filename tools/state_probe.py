#!/usr/bin/env python
"""Probe: the traceback time of one binary differs between handles of one process (which hardware queues the handle's streams land on).
Samples `reps` fresh handles: python tools/state_probe.py [reps] [cfg4|cfg2|cfg2n|cfg3]; the environment selects the variant
(UBU_TB_SCHED, UBU_TB_PHASES ...)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ubu_b200 as U
from ubu_b200 import workloads as WL

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
which = sys.argv[2] if len(sys.argv) > 2 else "cfg4"
w = {"cfg4": lambda: WL.cfg4(1_000_000, tie_stress=20_000), "cfg4s": lambda: WL.cfg4(100_000, tie_stress=2_000), "cfg2": lambda: WL.cfg2(1_000_000),
     "cfg2n": lambda: WL.with_n(WL.uniform("cfg2", 1_000_000, 76, 500, seed=5, paired=True, keep_codes=True), 0.005),
     "cfg3": lambda: WL.cfg3_sharded(5_000_000)}[which]()
tb, tot = [], []
for r in range(reps):
    with U.SwAligner(device=0, slots=2) as al:
        al.stage(0, w.reads, w.windows, w.pair_ref_idx); al.sync(0)
        al.run(0); al.sync(0)
        ms = []
        for _ in range(3):
            al.run(0); al.sync(0); ms.append(al.last_run_ms(0))
        m = np.mean(np.array(ms), axis=0)
        tb.append(float(m[2])); tot.append(float(m[0]))
env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("UBU_TB"))
print(f"{which} [{env}] per fresh handle: total " + " ".join(f"{x:.2f}" for x in tot) + " | traceback " + " ".join(f"{x:.2f}" for x in tb)
      + f" | mean total {np.mean(tot):.2f} max {np.max(tot):.2f}", flush=True)
