"""Python face of the C ABI (include/ubu_sw.h): one SwAligner per GPU.

Mirrors the role of the reference's `Aligner` object (src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:13-16):
constructed once, then handed batches. The alignment itself always runs in libubu_sw.so on the GPU;
this module only marshals numpy buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from . import _native as N
from .seqs import SeqSet

RESULT_DTYPE = np.dtype(
    [("score", "<i4"), ("ref_start", "<i4"), ("ref_end", "<i4"), ("q_start", "<i4"), ("q_end", "<i4"),
     ("edit_distance", "<i4"), ("cigar_off", "<u4"), ("cigar_n", "<u2"), ("flags", "<u2")]
)
assert RESULT_DTYPE.itemsize == C.sizeof(N.Result) == 32

OP_CHARS = "MIDNSHP=X"


@dataclass
class Alignments:
    """Results in input order (SPEC.md section 6)."""

    results: np.ndarray      # RESULT_DTYPE [n]
    cigar_pool: np.ndarray   # uint32, BAM-encoded ops

    def __len__(self) -> int:
        return int(self.results.shape[0])

    def cigar(self, k: int) -> List[Tuple[int, int]]:
        r = self.results[k]
        ops = self.cigar_pool[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["cigar_n"])]
        return [(int(o) >> 4, int(o) & 15) for o in ops]

    def cigar_string(self, k: int) -> str:
        c = self.cigar(k)
        return "".join(f"{n}{OP_CHARS[o]}" for n, o in c) if c else "*"

    def record(self, k: int) -> dict:
        r = self.results[k]
        d = {f: int(r[f]) for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance", "flags")}
        d["cigar"] = self.cigar(k)
        return d


class SwAligner:
    """One handle = one GPU (ubu_sw_create). Not thread-safe; use one per host thread / GPU."""

    def __init__(self, match: int = 1, mismatch: int = -3, gap_open: int = 5, gap_ext: int = 2,
                 device: int = 0, slots: int = 4):
        self._lib = N.lib()
        self._h = C.c_void_p()
        p = N.Params(match, mismatch, gap_open, gap_ext)
        rc = self._lib.ubu_sw_create(C.byref(p), device, slots, C.byref(self._h))
        if rc != N.OK:
            self._h = C.c_void_p()
            raise N.UbuSwError(rc)
        self.device, self.slots = device, slots
        self._keep = {}

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ubu_sw_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != N.OK:
            raise N.UbuSwError(rc, self._lib.ubu_sw_last_error(self._h).decode())

    @staticmethod
    def _idx(pair_ref_idx: Optional[np.ndarray], n: int):
        if pair_ref_idx is None:
            return None, None
        a = np.ascontiguousarray(pair_ref_idx, dtype=np.uint32)
        if a.shape[0] != n:
            raise ValueError("pair_ref_idx must have one entry per read")
        return a, a.ctypes.data

    # -- synchronous call (ubu_sw_align_batch) --------------------------------------------------
    def align(self, reads: SeqSet, windows: SeqSet, pair_ref_idx: Optional[np.ndarray] = None,
              cigar_cap: Optional[int] = None) -> Alignments:
        n = reads.count
        idx, idx_p = self._idx(pair_ref_idx, n)
        rs, ws = reads.as_struct(), windows.as_struct()
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cap = int(cigar_cap if cigar_cap is not None else 6 * n + 4096)
        while True:
            pool = np.zeros(cap, dtype=np.uint32)
            used = C.c_size_t(0)
            rc = self._lib.ubu_sw_align_batch(self._h, C.byref(rs), C.byref(ws), idx_p, res.ctypes.data,
                                              pool.ctypes.data, cap, C.byref(used))
            if rc == N.ERR_CIGAR_POOL and cigar_cap is None:
                cap = int(used.value) + 16
                continue
            self._check(rc)
            return Alignments(res, pool[: used.value])

    # -- staged calls (device-resident timing) ---------------------------------------------------
    def stage(self, slot: int, reads: SeqSet, windows: SeqSet, pair_ref_idx: Optional[np.ndarray] = None) -> None:
        idx, idx_p = self._idx(pair_ref_idx, reads.count)
        rs, ws = reads.as_struct(), windows.as_struct()
        self._keep[slot] = (reads, windows, idx, rs, ws)      # buffers must outlive the async copies
        self._check(self._lib.ubu_sw_stage(self._h, slot, C.byref(rs), C.byref(ws), idx_p))

    def run(self, slot: int) -> None:
        self._check(self._lib.ubu_sw_run(self._h, slot))

    def sync(self, slot: int) -> None:
        self._check(self._lib.ubu_sw_sync(self._h, slot))

    def fetch(self, slot: int, n: int, cigar_cap: Optional[int] = None) -> Alignments:
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cap = int(cigar_cap if cigar_cap is not None else 6 * n + 4096)
        while True:
            pool = np.zeros(cap, dtype=np.uint32)
            used = C.c_size_t(0)
            rc = self._lib.ubu_sw_fetch(self._h, slot, res.ctypes.data, pool.ctypes.data, cap, C.byref(used))
            if rc == N.ERR_CIGAR_POOL and cigar_cap is None:
                cap = int(used.value) + 16
                continue
            self._check(rc)
            return Alignments(res, pool[: used.value])

    def last_run_ms(self, slot: int) -> Tuple[float, float, float]:
        ms = (C.c_float * 3)()
        self._check(self._lib.ubu_sw_last_run_ms(self._h, slot, ms))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def last_run_launches(self, slot: int) -> int:
        return int(self._lib.ubu_sw_last_run_launches(self._h, slot))

    def debug_counters(self, slot: int) -> np.ndarray:
        """Tasks per traceback class [0..9], CIGAR ops [10], error flags [11], strip work units [12] of the last run()."""
        out = np.zeros(16, dtype=np.uint32)
        self._check(self._lib.ubu_sw_debug_counters(self._h, slot, out.ctypes.data))
        return out

    # -- pipelined calls (ubu_sw_submit / ubu_sw_wait) -------------------------------------------
    def submit(self, reads: SeqSet, windows: SeqSet, pair_ref_idx: Optional[np.ndarray], res: np.ndarray,
               pool: np.ndarray) -> int:
        idx, idx_p = self._idx(pair_ref_idx, reads.count)
        rs, ws = reads.as_struct(), windows.as_struct()
        ticket = C.c_int(-1)
        rc = self._lib.ubu_sw_submit(self._h, C.byref(rs), C.byref(ws), idx_p, res.ctypes.data, pool.ctypes.data,
                                     pool.shape[0], C.byref(ticket))
        self._check(rc)
        self._keep[("t", ticket.value)] = (reads, windows, idx, rs, ws, res, pool)
        return ticket.value

    def wait(self, ticket: int) -> int:
        used = C.c_size_t(0)
        rc = self._lib.ubu_sw_wait(self._h, ticket, C.byref(used))
        self._keep.pop(("t", ticket), None)
        self._check(rc)
        return int(used.value)


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over cudaHostAlloc'd memory (ubu_sw_host_alloc); freed when the array is collected."""
    import weakref

    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    lib = N.lib()
    p = lib.ubu_sw_host_alloc(max(n, 1))
    if not p:
        raise MemoryError("ubu_sw_host_alloc failed")
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.ubu_sw_host_free, p)
    return arr


def write_sam(fd: int, reads: SeqSet, al: Alignments, pair_ref_idx: Optional[np.ndarray] = None, first_index: int = 0,
              qname_prefix: str = "r", window_pos0: Optional[np.ndarray] = None, threads: int = 0) -> int:
    """ubu_sw_sam_write: SAM lines for a batch of results in input order (QNAME FLAG RNAME POS MAPQ CIGAR * 0 0 SEQ * NM:i AS:i),
    formatted on the host's cores and written to file descriptor fd (fd < 0: format only). Returns the bytes produced."""
    n = len(al)
    idx = None if pair_ref_idx is None else np.ascontiguousarray(pair_ref_idx, dtype=np.uint32)
    pos0 = None if window_pos0 is None else np.ascontiguousarray(window_pos0, dtype=np.int64)
    pool = np.ascontiguousarray(al.cigar_pool, dtype=np.uint32)
    names = N.SamNames(qname_prefix.encode("ascii"), first_index, None if idx is None else idx.ctypes.data, None,
                       None if pos0 is None else pos0.ctypes.data, int(pool.shape[0]))
    rs = reads.as_struct()
    res = np.ascontiguousarray(al.results)
    out = C.c_uint64(0)
    rc = N.lib().ubu_sw_sam_write(fd, C.byref(rs), res.ctypes.data, pool.ctypes.data if pool.size else None, n, C.byref(names), threads,
                                  C.byref(out))
    if rc != N.OK:
        raise N.UbuSwError(rc)
    return int(out.value)
