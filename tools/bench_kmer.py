#!/usr/bin/env python
"""Throughput of the k-mer graph build on cuda:0: k-mer occurrences per second with the reads
resident in HBM, the share of each pass, the HBM traffic the hash-table kernels imply, and the CPU oracle beside it on a
bounded sample. Not the bench contract (bench.py measures the alignment path); prints one JSON line.

    python tools/bench_kmer.py --regions 5000 --reads-per-region 200 --read-len 100 --k 63
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ubu_b200 as U  # noqa: E402
from ubu_b200 import kmer as K  # noqa: E402


def make(n_regions, per, L, ref_len, seed, sub=0.005):
    rng = np.random.default_rng(seed)
    refs = rng.integers(0, 4, size=(n_regions, ref_len), dtype=np.uint8)
    starts = rng.integers(0, ref_len - L + 1, size=(n_regions, per))
    codes = refs[np.arange(n_regions)[:, None, None], starts[:, :, None] + np.arange(L)[None, None, :]].reshape(-1, L)
    m = rng.random(codes.shape) < sub
    codes[m] = (codes[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
    return codes, np.repeat(np.arange(n_regions, dtype=np.uint32), per)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--regions", type=int, default=5000)
    ap.add_argument("--reads-per-region", type=int, default=200)
    ap.add_argument("--read-len", type=int, default=100)
    ap.add_argument("--ref-len", type=int, default=400)
    ap.add_argument("--k", type=int, default=63)
    ap.add_argument("--min-node-frequency", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--shuffle", action="store_true", help="interleave the regions' reads in the batch (default: region by region)")
    ap.add_argument("--cpu-regions", type=int, default=20, help="regions in the bounded CPU-oracle sample (0 = skip)")
    a = ap.parse_args()
    codes, regions = make(a.regions, a.reads_per_region, a.read_len, a.ref_len, 4242)
    if a.shuffle:
        perm = np.random.default_rng(1).permutation(codes.shape[0])
        codes, regions = codes[perm], regions[perm]
    reads = U.pack_uniform(codes)
    with K.KmerGraphBuilder(a.k, a.min_node_frequency) as b:
        b.stage(reads, regions, max_nodes=0)
        b.run()
        nodes, _ = b.fetch(int(b.last_run_stats()[0]))
        n_nodes = nodes.shape[0]
        b.stage(reads, regions, max_nodes=n_nodes + 1024)            # a realistic bound, as a caller that knows its coverage gives
        for _ in range(2):
            b.run()
        ms = []
        for _ in range(a.steps):
            b.run(); ms.append(b.last_run_ms())
        occ, table_bytes, launches = b.last_run_stats()
        nodes2, _ = b.fetch(n_nodes + 1024)
        assert nodes2.tobytes() == nodes.tobytes()
    m = np.mean(np.array(ms), axis=0)
    line = {"metric": "k-mer graph build, occurrences/s", "value": occ / (m[0] * 1e-3), "unit": "k-mer occurrences/s",
            "ms": {"total": float(m[0]), "clear_and_read_classes": float(m[1]), "insert_and_edges": float(m[2]), "order_and_emit": float(m[3])},
            "config": {"reads": int(codes.shape[0]), "read_len": a.read_len, "regions": a.regions, "shuffled": bool(a.shuffle), "k": a.k, "occurrences": occ,
                       "nodes": int(n_nodes), "table_bytes": table_bytes, "launches": launches},
            # algorithmic traffic of the hash-table kernel: the node's 128-byte line comes in from HBM and goes back once per
            # occurrence (the table is far larger than L2 and concurrent warps work on distant reads); peak = MEASURED_PEAKS hbm
            "roofline": {"bound": "hbm", "kernel": "kmer_insert_kernel", "bytes_per_occurrence": 256,
                         "achieved": occ * 256 / (m[2] * 1e-3) / 1e9, "peak": 6551.4, "unit": "GB/s",
                         "frac": occ * 256 / (m[2] * 1e-3) / 1e9 / 6551.4}}
    if a.cpu_regions:
        import kmer_oracle as KO

        nr = min(a.cpu_regions, a.regions)
        sel = np.nonzero(regions < nr)[0]
        texts = ["".join("ACTG"[c] for c in codes[i]) for i in sel]
        t0 = time.perf_counter()
        rows, _ = KO.build_graphs(texts, regions[sel], a.k, a.min_node_frequency)
        dt = time.perf_counter() - t0
        o = len(sel) * (a.read_len - a.k + 1)
        line["cpu_baseline"] = {"value": o / dt, "unit": "k-mer occurrences/s", "cores": 1, "kind": "port",
                                "sample": f"{nr} of {a.regions} regions ({len(sel)} reads), oracle/kmer_oracle.py (pure Python), {dt:.2f} s"}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
