#!/usr/bin/env python
"""Times the one public call, ubu_sw_align_batch, on the cfg2 batch with pinned host buffers: the streamed form (chunks through
the handle's slots) against the one-shot form (UBU_NO_STREAM=1), for a few pipeline depths. Development tool."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402

import ubu_b200 as U  # noqa: E402
from ubu_b200 import _native as N  # noqa: E402
from ubu_b200 import workloads as WL  # noqa: E402


def pin(a):
    p = U.pinned_empty(a.shape, a.dtype); p[...] = a; return p


def main():
    w = WL.cfg2(1_000_000)
    reads = U.SeqSet(pin(w.reads.packed), pin(w.reads.off), pin(w.reads.len), None, None)
    wins = U.SeqSet(pin(w.windows.packed), pin(w.windows.off), pin(w.windows.len), None, None)
    idx = pin(w.pair_ref_idx)
    n = w.n_tasks
    res = U.pinned_empty((n,), U.RESULT_DTYPE); pool = U.pinned_empty((6 * n + 4096,), np.uint32)
    rs, ws = reads.as_struct(), wins.as_struct()
    used = C.c_size_t(0)
    for slots in (2, 4, 6):
        with U.SwAligner(slots=slots) as a:
            for mode in ("stream", "noramp", "oneshot"):
                os.environ.pop("UBU_NO_STREAM", None); os.environ.pop("UBU_STREAM_NORAMP", None)
                if mode == "oneshot":
                    os.environ["UBU_NO_STREAM"] = "1"
                if mode == "noramp":
                    os.environ["UBU_STREAM_NORAMP"] = "1"
                def call():
                    rc = N.lib().ubu_sw_align_batch(a._h, C.byref(rs), C.byref(ws), idx.ctypes.data, res.ctypes.data, pool.ctypes.data,
                                                    pool.shape[0], C.byref(used))
                    assert rc == 0, rc
                call(); call()
                t0 = time.perf_counter()
                for _ in range(8):
                    call()
                dt = (time.perf_counter() - t0) / 8
                print(f"slots {slots} {mode:8s} {dt * 1e3:7.2f} ms  {w.cells / dt / 1e9:6.0f} GCUPS  {n / dt / 1e6:6.1f} M reads/s", flush=True)
    os.environ.pop("UBU_NO_STREAM", None); os.environ.pop("UBU_STREAM_NORAMP", None)


if __name__ == "__main__":
    main()
