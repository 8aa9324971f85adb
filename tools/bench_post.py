#!/usr/bin/env python
"""Throughput of the post-alignment entry points on cuda:0, through the C ABI with host
buffers (these calls are synchronous: H2D + kernel + D2H are inside the timed region). HBM-bound integer work; the
kernel-only figure is printed from CUDA-event-free wall time minus nothing -- treat it as end to end.

    python tools/bench_post.py --n 1000000
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ubu_b200 as U  # noqa: E402
from ubu_b200 import post as POST  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    a = ap.parse_args()
    rng = np.random.default_rng(3)
    n, L = a.n, 100
    # recount: reads placed on a 50 Mbp reference, CIGAR 40M2I58M / 100M / 30M3D70M
    G = 50_000_000
    ref = rng.integers(0, 4, size=G, dtype=np.uint8)
    reftext = U.text_from_codes(ref)
    pos = rng.integers(1, G - 200, size=n)
    kinds = rng.integers(0, 3, size=n)
    cig_of = {0: [(100, 0)], 1: [(40, 0), (2, 1), (58, 0)], 2: [(30, 0), (3, 2), (70, 0)]}
    cigars = [cig_of[int(k)] for k in kinds]
    reads = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
    rs = U.pack_uniform(reads)
    with U.SwAligner() as al:
        t0 = time.perf_counter(); POST.load_refs(al, [reftext]); t_load = time.perf_counter() - t0
        pool, off, cn = POST.encode_cigars(cigars)
        placed = np.zeros(n, dtype=POST.PLACED_DTYPE)
        placed["ref_idx"], placed["pos"], placed["cigar_off"], placed["cigar_n"] = 0, pos, off, cn
        placed["mapq"], placed["yq"], placed["ym"] = 30, 60, 1
        out = np.zeros(n, dtype=POST.RECOUNT_DTYPE)
        import ctypes as C
        st = rs.as_struct()
        def recount_once():
            al._check(al._lib.ubu_sw_recount_batch(al._h, C.byref(st), placed.ctypes.data, pool.ctypes.data, int(pool.shape[0]), out.ctypes.data))
        recount_once()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            recount_once()
        dt = (time.perf_counter() - t0) / a.steps
        bytes_in = rs.packed.nbytes + rs.off.nbytes + rs.len.nbytes + placed.nbytes + pool.nbytes
        print(json.dumps({"op": "recount (XM/NM/MAPQ)", "reads": n, "ms": dt * 1e3, "reads_per_s": n / dt, "h2d_bytes": int(bytes_in),
                          "d2h_bytes": int(out.nbytes), "ref_upload_s": round(t_load, 3), "mean_xm": float(out["xm"].mean())}))
        # projection: contigs 50M10D50M40M.. ; reads placed uniformly
        ncont = 50_000
        ccig = [[(60, 0), (int(rng.integers(1, 30)), 2), (80, 0), (int(rng.integers(1, 6)), 1), (120, 0)] for _ in range(ncont)]
        cpos = rng.integers(1, G, size=ncont)
        cidx = rng.integers(0, ncont, size=n); position = rng.integers(0, 170, size=n); rlen = np.full(n, L)
        POST.project(al, cpos, ccig, cidx[:1000], position[:1000], rlen[:1000])
        t0 = time.perf_counter()
        for _ in range(a.steps):
            o, _ = POST.project(al, cpos, ccig, cidx, position, rlen) if False else (None, None)
        # time the C call only (the Python list building in POST.project is not the path)
        cpool, coff, cn2 = POST.encode_cigars(ccig)
        contigs = np.zeros(ncont, dtype=POST.CONTIG_DTYPE); contigs["pos"], contigs["cigar_off"], contigs["cigar_n"] = cpos, coff, cn2
        pin = np.zeros(n, dtype=POST.PROJ_IN_DTYPE); pin["contig_idx"], pin["position"], pin["read_len"] = cidx, position, rlen
        pout = np.zeros(n, dtype=POST.PROJ_OUT_DTYPE); cap = 8 * n + 64; opool = np.zeros(cap, dtype=np.uint32); used = C.c_size_t(0)
        def project_once():
            al._check(al._lib.ubu_sw_project_batch(al._h, contigs.ctypes.data, ncont, cpool.ctypes.data, int(cpool.shape[0]), pin.ctypes.data, n,
                                                   pout.ctypes.data, opool.ctypes.data, cap, C.byref(used)))
        project_once()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            project_once()
        dt = (time.perf_counter() - t0) / a.steps
        print(json.dumps({"op": "project (contig CIGAR -> genome POS + CIGAR)", "reads": n, "ms": dt * 1e3, "reads_per_s": n / dt,
                          "h2d_bytes": int(pin.nbytes + contigs.nbytes + cpool.nbytes), "d2h_bytes": int(pout.nbytes + used.value * 4),
                          "ok_fraction": float((pout["flags"] == 0).mean())}))


if __name__ == "__main__":
    main()
