#!/usr/bin/env python
"""Strong scaling of the multi-GPU driver (ubu_sw_multi_*) in ONE process: the same cfg5-shaped input on 1, 2, 4, ... visible GPUs.

    python tools/multi_probe.py --reads 10000000 --gpus 1,2 [--chunk 65536] [--slots 4] [--sam]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ubu_b200 import multi as M  # noqa: E402
from ubu_b200 import workloads as WL  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=10_000_000)
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--slots", type=int, default=4)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--sam", action="store_true")
    a = ap.parse_args()
    w = WL.cfg3_sharded(a.reads, seed=WL.SEEDS["cfg5"], threads=16)
    base = None
    for n in [int(x) for x in a.gpus.split(",")]:
        with M.MultiAligner(devices=list(range(n)), chunk=a.chunk, slots=a.slots) as ma:
            b = ma.pin_workload(w)
            ma.align_pinned(b)
            ts = []
            for _ in range(a.steps):
                t0 = time.perf_counter(); ma.align_pinned(b); ts.append(time.perf_counter() - t0)
            dt = float(np.median(ts))
            row = {"gpus": n, "reads": a.reads, "chunk": a.chunk, "slots": a.slots, "ms": dt * 1e3, "gcups": w.cells / dt / 1e9, "reads_per_s": a.reads / dt}
            if base is None:
                base = dt
            row["speedup_vs_first"] = base / dt
            if a.sam:
                fd = os.open(os.devnull, os.O_WRONLY)
                ts = []
                for _ in range(max(2, a.steps // 2)):
                    t0 = time.perf_counter(); ma.align_pinned(b, sam_fd=fd); ts.append(time.perf_counter() - t0)
                os.close(fd)
                row["to_sam_ms"] = float(np.median(ts)) * 1e3; row["to_sam_reads_per_s"] = a.reads / float(np.median(ts)); row["sam_bytes"] = ma.last_sam_bytes
            print(json.dumps(row), flush=True)
            del b


if __name__ == "__main__":
    main()
