#!/usr/bin/env python
"""Throughput of the long-read path (sw_long_kernel, reads of 257-4096 bp: the contig-sized queries of Aligner.align) on cuda:0:
contigs of 300-2500 bp against 3 kb windows, every record checked against the oracle.

    python tools/bench_long.py [--contigs 4000]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=4000)
    a = ap.parse_args()
    rng = np.random.default_rng(17)
    W = 3000
    wins = rng.integers(0, 4, size=(a.contigs, W), dtype=np.uint8)
    reads = []
    for k in range(a.contigs):
        L = int(rng.integers(300, 2501))
        s = int(rng.integers(0, W - L + 1))
        r = wins[k, s:s + L].copy()
        sub = rng.random(L) < 0.01
        r[sub] = (r[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
        if rng.random() < 0.5 and L > 400:                       # a deletion of 5-60 bp in the sample, as the assembler finds them
            cut = int(rng.integers(100, L - 100)); d = int(rng.integers(5, 61))
            r = np.concatenate([r[:cut], r[cut + d:]])
        reads.append(r)
    rs, ws = U.pack_codes(reads), U.pack_uniform(wins)
    cells = float(sum(len(r) for r in reads)) * W
    with U.SwAligner(device=0, slots=2) as al:
        al.stage(0, rs, ws, None); al.sync(0)
        al.run(0); al.sync(0)
        ms = []
        for _ in range(3):
            al.run(0); al.sync(0); ms.append(al.last_run_ms(0)[0])
        got = al.fetch(0, len(reads))
    t0 = time.time(); ores, ocig = O.align_batch(reads, list(wins), None, nthreads=os.cpu_count() or 1); t_cpu = time.time() - t0
    bad = sum(1 for k in range(len(reads)) if any(int(got.results[k][f]) != int(ores[k][f]) for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance"))
              or got.cigar(k) != ocig[k])
    m = float(np.mean(ms))
    print(json.dumps({"contigs": a.contigs, "mean_len": float(np.mean([len(r) for r in reads])), "window": W, "cells": cells, "ms": m,
                      "gcups": cells / (m * 1e-3) / 1e9, "contigs_per_s": a.contigs / (m * 1e-3), "records_differing": bad,
                      "oracle_s": round(t_cpu, 1), "oracle_gcups": round(cells / t_cpu / 1e9, 2)}))


if __name__ == "__main__":
    main()
