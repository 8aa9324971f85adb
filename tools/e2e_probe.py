#!/usr/bin/env python
"""End-to-end (host buffers in / host results out) timing probe for ubu_sw_submit / ubu_sw_wait: sweeps chunk size and
slot count, with the pipeline drained after every pass over the batch ("drain") or kept full across passes ("stream").
Development tool; bench.py holds the contract measurement.

    python tools/e2e_probe.py --chunks 250000,125000 --slots 3,4 --passes 6
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ubu_b200 as U  # noqa: E402
from ubu_b200 import sharding as SH  # noqa: E402
from ubu_b200 import workloads as WL  # noqa: E402


def pin(a):
    p = U.pinned_empty(a.shape, a.dtype); p[...] = a; return p


def pin_set(sq):
    return U.SeqSet(pin(sq.packed), pin(sq.off), pin(sq.len), None, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--chunks", default="250000,125000,62500")
    ap.add_argument("--slots", default="3,4")
    ap.add_argument("--passes", type=int, default=6)
    ap.add_argument("--detail", action="store_true")
    ap.add_argument("--ramp", action="store_true", help="quarter- and half-size chunks at both ends of a pass (shorter pipeline fill and drain)")
    a = ap.parse_args()
    n = a.n
    w = WL.uniform("cfg2", n, 76, 500, seed=1002, paired=True)
    cells = float(w.cells)
    for slots in [int(x) for x in a.slots.split(",")]:
        al = U.SwAligner(device=0, slots=slots)
        for chunk in [int(x) for x in a.chunks.split(",")]:
            bounds = [(lo, min(lo + chunk, n)) for lo in range(0, n, chunk)]
            if a.ramp and n >= 4 * chunk:
                sizes = [chunk // 4, chunk // 4, chunk // 2]
                body = n - 2 * sum(sizes)
                sizes = sizes + [chunk] * (body // chunk) + ([body % chunk] if body % chunk else []) + sizes[::-1]
                edges = np.concatenate([[0], np.cumsum(sizes)])
                bounds = [(int(edges[k]), int(edges[k + 1])) for k in range(len(sizes))]
            chunks = []
            for lo, hi in bounds:
                wsl, ix = SH.window_slice(w.windows, w.pair_ref_idx, lo, hi)
                chunks.append((pin_set(w.reads.slice(lo, hi)), pin_set(wsl), pin(ix)))
            res = U.pinned_empty((n,), U.RESULT_DTYPE)
            pools = [U.pinned_empty((6 * (hi - lo) + 4096,), np.uint32) for lo, hi in bounds]
            # drain after every pass
            best = 1e9
            for rep in range(a.passes):
                ts, tw = [], []
                t0 = time.perf_counter(); infl = []
                for ci, (lo, hi) in enumerate(bounds):
                    if len(infl) == al.slots:
                        x = time.perf_counter(); al.wait(infl.pop(0)); tw.append(time.perf_counter() - x)
                    x = time.perf_counter(); infl.append(al.submit(*chunks[ci], res[lo:hi], pools[ci])); ts.append(time.perf_counter() - x)
                for t in infl:
                    x = time.perf_counter(); al.wait(t); tw.append(time.perf_counter() - x)
                tot = time.perf_counter() - t0
                best = min(best, tot)
                if a.detail and rep == a.passes - 1:
                    print(f"   submit ms {[round(x * 1e3, 2) for x in ts]}\n   wait ms {[round(x * 1e3, 2) for x in tw]}")
            # pipeline kept full across passes
            infl = []
            t0 = None
            for rep in range(a.passes + 1):
                if rep == 1:
                    t0 = time.perf_counter()          # first pass = warm-up, its tail overlaps the second like every later one
                for ci, (lo, hi) in enumerate(bounds):
                    if len(infl) == al.slots:
                        al.wait(infl.pop(0))
                    infl.append(al.submit(*chunks[ci], res[lo:hi], pools[ci]))
            for t in infl:
                al.wait(t)
            stream = (time.perf_counter() - t0) / a.passes
            print(f"slots {slots} chunk {chunk:7d}: drain {best * 1e3:6.2f} ms/pass = {cells / best / 1e9:6.0f} GCUPS | "
                  f"stream {stream * 1e3:6.2f} ms/pass = {cells / stream / 1e9:6.0f} GCUPS", flush=True)
        al.close()


if __name__ == "__main__":
    main()
