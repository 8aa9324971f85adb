#!/usr/bin/env python
"""BASELINE.json configs[4] at a chosen scale: 150 bp reads against 2 kb windows, generated shard by shard from (seed, shard),
sharded over the GPUs of one process (ubu_b200.sharding.MultiGpuAligner: a host thread + a handle per GPU, every handle's own
upload / compute / download pipeline, chunks merged in input order on the host), then written as SAM records in input order.

Reports, per SURVEY.md section 8e: reads/s to merged binary results, the SAM writer's own rate, and reads/s to SAM text when
the two run one after the other per shard. The full config is 100M reads (3.0e13 cells); --shards x --reads-per-shard picks
how much of it to run. Not the bench contract (bench.py); prints one JSON line.

    python tools/run_cfg5.py --gpus 2 --shards 10 --reads-per-shard 1000000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ubu_b200 as U  # noqa: E402
from ubu_b200 import sharding as SH  # noqa: E402
from ubu_b200 import workloads as WL  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--shards", type=int, default=5)
    ap.add_argument("--reads-per-shard", type=int, default=1_000_000)
    ap.add_argument("--windows-per-shard", type=int, default=0, help="window pool of a shard (0 = reads / 5)")
    ap.add_argument("--chunk", type=int, default=125_000)
    ap.add_argument("--sam", default="/dev/null", help="where the SAM text goes")
    ap.add_argument("--check", type=int, default=500, help="reads per shard checked against the CPU oracle")
    a = ap.parse_args()
    L, W = 150, 2000
    pool = a.windows_per_shard or max(a.reads_per_shard // 5, 1)
    fd = os.open(a.sam, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    t_align = t_sam = t_gen = 0.0
    sam_bytes = 0
    cells = 0
    bad = checked = 0
    with SH.MultiGpuAligner(list(range(a.gpus)), chunk=a.chunk, aligner_factory=lambda dev: U.SwAligner(device=dev, slots=4)) as mga:
        for s in range(a.shards + 1):                      # shard 0 runs twice: the first pass warms the handles up (allocations)
            shard = max(s - 1, 0)
            t0 = time.perf_counter()
            w = WL.uniform("cfg5", a.reads_per_shard, L, W, n_windows=pool, seed=WL.SEEDS["cfg5"] + 7919 * shard, keep_codes=a.check > 0)
            # reads arrive grouped by window, as coordinate-sorted input does in the reference (ReAligner.java:237): a chunk then
            # needs only its own stretch of windows
            order = np.argsort(w.pair_ref_idx, kind="stable")
            nb = (L + 3) // 4
            w.reads = U.SeqSet(np.ascontiguousarray(w.reads.packed[: w.n_tasks * nb].reshape(w.n_tasks, nb)[order]).reshape(-1), w.reads.off, w.reads.len, None, None)
            w.pair_ref_idx = np.ascontiguousarray(w.pair_ref_idx[order])
            if w.read_codes is not None:
                w.read_codes = w.read_codes[order]
            t1 = time.perf_counter()
            al = mga.align(w.reads, w.windows, w.pair_ref_idx)
            t2 = time.perf_counter()
            nbytes = U.write_sam(fd if s else -1, w.reads, al, w.pair_ref_idx, first_index=shard * a.reads_per_shard, qname_prefix="read")
            t3 = time.perf_counter()
            if s == 0:
                continue
            t_gen += t1 - t0; t_align += t2 - t1; t_sam += t3 - t2; sam_bytes += nbytes; cells += w.cells
            if a.check:
                import oracle_lib as O

                pick = np.sort(np.random.default_rng(shard).choice(w.n_tasks, size=min(a.check, w.n_tasks), replace=False))
                ores, ocig = O.align_batch([w.read_codes[k] for k in pick], [w.window_codes[int(w.pair_ref_idx[k])] for k in pick], None,
                                           nthreads=os.cpu_count() or 1)
                for j, k in enumerate(pick):
                    ok = all(int(al.results[int(k)][f]) == int(ores[j][f]) for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance"))
                    bad += 0 if ok and al.cigar(int(k)) == ocig[j] else 1
                checked += len(pick)
    os.close(fd)
    n = a.shards * a.reads_per_shard
    line = {"config": "cfg5 (150 bp vs 2 kb windows, sharded, ordered merge, SAM)", "gpus": a.gpus, "reads": n, "cells": cells,
            "fraction_of_full_config": n / 100_000_000,
            "to_merged_binary_results": {"reads_per_s": n / t_align, "gcups": cells / t_align / 1e9, "seconds": t_align},
            "sam_writer": {"reads_per_s": n / t_sam, "gb_per_s": sam_bytes / t_sam / 1e9, "bytes": sam_bytes, "seconds": t_sam, "threads": os.cpu_count()},
            "to_sam_text_serial": {"reads_per_s": n / (t_align + t_sam), "gcups": cells / (t_align + t_sam) / 1e9},
            "generation_seconds_not_counted": t_gen, "oracle_checked": checked, "oracle_mismatches": bad}
    print(json.dumps(line))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
