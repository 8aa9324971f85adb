#!/usr/bin/env python
"""Probe: a mixed-length batch (cfg4) through ONE ubu_sw_align_batch call (the library streams it through its slots in chunks, host
buffers pinned) next to the same batch resident on the device (stage / run).  python tools/stream_probe.py [n_reads] [slots]"""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ubu_b200 as U
from ubu_b200 import _native as N, workloads as WL

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = WL.cfg4(n, tie_stress=n // 50)
def pin(a):
    p = U.pinned_empty(a.shape, a.dtype); p[...] = a; return p
def pin_set(sq):
    return U.SeqSet(pin(sq.packed), pin(sq.off), pin(sq.len), None, None)
with U.SwAligner(device=0, slots=slots) as al:
    al.stage(0, w.reads, w.windows, w.pair_ref_idx); al.sync(0)
    al.run(0); al.sync(0)
    ms = []
    for _ in range(3):
        al.run(0); al.sync(0); ms.append(al.last_run_ms(0)[0])
    res0 = al.fetch(0, n)
    r, ww, ix = pin_set(w.reads), pin_set(w.windows), None if w.pair_ref_idx is None else pin(w.pair_ref_idx)
    res_p = U.pinned_empty((n,), U.RESULT_DTYPE); pool = U.pinned_empty((8 * n + 4096,), np.uint32)
    rs_, ws_, used_ = r.as_struct(), ww.as_struct(), C.c_size_t(0)
    def one_call():
        rc = N.lib().ubu_sw_align_batch(al._h, C.byref(rs_), C.byref(ws_), None if ix is None else ix.ctypes.data, res_p.ctypes.data,
                                        pool.ctypes.data, pool.shape[0], C.byref(used_))
        if rc: raise N.UbuSwError(rc)
    one_call(); one_call()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); one_call(); ts.append(time.perf_counter() - t0)
    same = all(np.array_equal(res_p[f], res0.results[f]) for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance", "cigar_n"))
    env = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("UBU_"))
    print(f"cfg4 {n} reads [{env}]: resident {np.mean(ms):.2f} ms = {w.cells / np.mean(ms) / 1e6:.0f} GCUPS; one align_batch call from pinned host "
          f"buffers {1e3 * np.mean(ts):.2f} ms = {w.cells / np.mean(ts) / 1e9:.0f} GCUPS (min {1e3 * min(ts):.2f}); same results: {same}", flush=True)
