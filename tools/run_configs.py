#!/usr/bin/env python
"""Runs the BASELINE.json configs (scaled by --scale) on cuda:0: throughput of the device-resident path per config
and a parity check of a sample against the CPU oracle. Not the bench contract (that is bench.py); a survey tool.

    python tools/run_configs.py --scale 0.2 --check 2000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import ubu_b200 as U  # noqa: E402
from ubu_b200 import workloads as WL  # noqa: E402


def check(al, w, res, n_check, threads):
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(w.n_tasks, size=min(n_check, w.n_tasks), replace=False))
    q = [w.reads.codes(int(k)) for k in pick]
    idx = w.pair_ref_idx
    wi = [int(idx[k]) if idx is not None else int(k) for k in pick]
    r = [w.windows.codes(k) for k in wi]
    ores, ocig = O.align_batch(q, r, None, nthreads=threads)
    bad = 0
    for n, k in enumerate(pick):
        g = res.results[int(k)]
        ok = all(int(g[f]) == int(ores[n][f]) for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance"))
        ok = ok and res.cigar(int(k)) == ocig[n]
        bad += 0 if ok else 1
    return len(pick), bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--check", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--only", default="")
    ap.add_argument("--flush", action="store_true", help="write a 512 MB buffer before every timed run (cold L2, as bench.py does)")
    a = ap.parse_args()
    cfgs = {
        "cfg1": lambda: WL.cfg1(10_000),
        "cfg2": lambda: WL.cfg2(int(1_000_000 * a.scale)),
        "cfg3": lambda: WL.cfg3_sharded(int(5_000_000 * a.scale)),
        "cfg4": lambda: WL.cfg4(int(1_000_000 * a.scale), tie_stress=int(20_000 * a.scale)),
        "cfg2+N": lambda: WL.with_n(WL.uniform("cfg2", int(1_000_000 * a.scale), 76, 500, seed=5, paired=True, keep_codes=True), 0.005),
    }
    threads = os.cpu_count() or 1
    out = []
    if a.flush:
        import torch
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda:0")
    with U.SwAligner(device=0, slots=2) as al:
        for name, make in cfgs.items():
            if a.only and name not in a.only.split(","):
                continue
            t0 = time.time(); w = make(); tgen = time.time() - t0
            al.stage(0, w.reads, w.windows, w.pair_ref_idx); al.sync(0)
            al.run(0); al.sync(0)
            ms = []
            for _ in range(a.steps):
                if a.flush:
                    flush.zero_(); torch.cuda.synchronize()
                al.run(0); al.sync(0); ms.append(al.last_run_ms(0))
            m = np.mean(np.array(ms), axis=0)
            res = al.fetch(0, w.n_tasks)
            nchk, bad = check(al, w, res, a.check, threads)
            row = {"config": name, "tasks": w.n_tasks, "cells": w.cells, "ms": float(m[0]), "fwd_ms": float(m[1]), "tb_ms": float(m[2]),
                   "gcups": w.cells / (m[0] * 1e-3) / 1e9, "reads_per_s": w.n_tasks / (m[0] * 1e-3), "launches": al.last_run_launches(0),
                   "unaligned": int((res.results["flags"] & 1).sum()), "tb_classes": [int(x) for x in al.debug_counters(0)[:10]], "checked": nchk, "mismatching": bad, "gen_s": round(tgen, 1)}
            out.append(row)
            print(json.dumps(row), flush=True)
    return 1 if any(r["mismatching"] for r in out) else 0


if __name__ == "__main__":
    sys.exit(main())
