#!/usr/bin/env python
"""EVERY record and CIGAR of a BASELINE config at its full size against the CPU oracle (all host cores), not a sample.

    python tools/check_full.py [--resident] cfg2 cfg4 cfg3 > profiles/full_check_rNN.jsonl
--resident: through ubu_sw_stage / run / fetch (one resident batch: the path that runs a large mixed-length batch's traceback in
phases) instead of one ubu_sw_align_batch call (which streams the batch through the slots in chunks).
cfg3 (5M reads, 1.5e12 cells) costs the oracle a few minutes on 16-32 cores; the others seconds."""
import ctypes as C
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

FIELDS = ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance")


def flat(sq):
    ln = sq.len.astype(np.int64)
    if ln.size and np.all(ln == ln[0]) and sq.nmask is None:          # uniform: unpack row-wise
        L = int(ln[0]); nb = (L + 3) // 4
        b = sq.packed[: sq.count * nb].reshape(sq.count, nb)
        codes = ((b[:, :, None] >> np.array([0, 2, 4, 6], dtype=np.uint8)) & 3).reshape(sq.count, nb * 4)[:, :L]
        return np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1), (np.arange(sq.count, dtype=np.uint64) * np.uint64(L))
    seq = np.repeat(np.arange(sq.count), ln)
    start = np.concatenate([[0], np.cumsum(ln)[:-1]])
    pos = np.arange(int(ln.sum())) - start[seq]
    codes = (sq.packed[sq.off.astype(np.int64)[seq] + pos // 4] >> ((pos % 4) * 2).astype(np.uint8)) & 3
    if sq.nmask is not None:
        codes = np.where((sq.nmask[sq.nmask_off.astype(np.int64)[seq] + pos // 8] >> (pos % 8).astype(np.uint8)) & 1, 4, codes)
    return np.ascontiguousarray(codes, dtype=np.uint8), start.astype(np.uint64)


def main():
    makers = {"cfg1": lambda: WL.cfg1(10_000), "cfg2": lambda: WL.cfg2(1_000_000), "cfg3": lambda: WL.cfg3_sharded(5_000_000),
              "cfg4": lambda: WL.cfg4(1_000_000, tie_stress=20_000),
              "cfg2+N": lambda: WL.with_n(WL.uniform("cfg2", 1_000_000, 76, 500, seed=5, paired=True, keep_codes=True), 0.005)}
    threads = os.cpu_count() or 1
    with U.SwAligner(device=0, slots=4) as al:
        resident = "--resident" in sys.argv[1:]
        for name in [a for a in sys.argv[1:] if not a.startswith("--")]:
            w = makers[name]()
            n = w.n_tasks
            t0 = time.time()
            if resident:
                al.stage(0, w.reads, w.windows, w.pair_ref_idx); al.run(0); al.sync(0); got = al.fetch(0, n)
            else:
                got = al.align(w.reads, w.windows, w.pair_ref_idx)
            t_gpu = time.time() - t0
            q, qoff = flat(w.reads); r, roff = flat(w.windows)
            rl, wl = w.reads.len.astype(np.uint32), w.windows.len.astype(np.uint32)
            idx = None if w.pair_ref_idx is None else np.ascontiguousarray(w.pair_ref_idx, dtype=np.uint32)
            res = np.zeros(n, dtype=O.SWO_RESULT)
            stride = int(rl.max()) + 8
            cig = np.zeros(n * stride, dtype=np.uint32)
            p = O.SwoParams(1, -3, 5, 2)
            t0 = time.time()
            st = O.lib().swo_align_batch(q.ctypes.data, qoff.ctypes.data, rl.ctypes.data, r.ctypes.data, roff.ctypes.data, wl.ctypes.data,
                                         None if idx is None else idx.ctypes.data, n, C.byref(p), res.ctypes.data, cig.ctypes.data, stride, threads)
            t_cpu = time.time() - t0
            assert st == 0
            g = got.results
            bad = np.zeros(n, dtype=bool)
            for f in FIELDS:
                bad |= g[f] != res[f]
            bad |= (g["flags"] & 1) != (res["flags"] & 1).astype(g["flags"].dtype)
            bad |= g["cigar_n"].astype(np.int64) != res["cigar_n"].astype(np.int64)
            cn = np.where(bad, 0, res["cigar_n"]).astype(np.int64)
            task = np.repeat(np.arange(n), cn)
            k = np.arange(int(cn.sum())) - np.concatenate([[0], np.cumsum(cn)[:-1]])[task]
            diff = got.cigar_pool[g["cigar_off"].astype(np.int64)[task] + k] != cig[task * stride + k]
            bad[np.unique(task[diff])] = True
            print(json.dumps({"config": name, "path": "stage/run/fetch" if resident else "align_batch", "launches": al.last_run_launches(0) if resident else None, "records": n, "cells": int(w.cells), "records_differing": int(bad.sum()),
                              "cigar_ops_compared": int(cn.sum()), "gapped_records": int((res["cigar_n"] > 3).sum()),
                              "unaligned": int((res["flags"] & 1).sum()), "gpu_call_s": round(t_gpu, 3), "oracle_s": round(t_cpu, 1),
                              "oracle_threads": threads, "oracle_gcups": round(w.cells / t_cpu / 1e9, 2)}), flush=True)
            del w, got, q, r, res, cig


if __name__ == "__main__":
    main()
