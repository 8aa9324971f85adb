"""Python face of the multi-GPU driver (include/ubu_sw.h, ubu_sw_multi_*): one process, a host thread and a handle per GPU
inside the library, chunk c -> GPU c mod N, records and CIGAR ops landing in input order in the caller's arrays.

The reference hands parallelism to `bwa -t N` (src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:19,49) and treats its
inputs as independent work (ReAligner.java:180-184,199-200); reads shard trivially, so there is no collective here."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _native as N
from .aligner import RESULT_DTYPE, Alignments, pinned_empty
from .seqs import SeqSet


@dataclass
class PinnedBatch:
    """A batch copied once into pinned host memory, with pinned output arrays: what a caller that reuses its buffers holds."""

    reads: SeqSet
    windows: SeqSet
    idx: Optional[np.ndarray]
    results: np.ndarray
    pool: np.ndarray
    used: int = 0

    def alignments(self) -> Alignments:
        return Alignments(self.results, self.pool[: self.used])


def _pin(a: np.ndarray) -> np.ndarray:
    p = pinned_empty(a.shape, a.dtype)
    p[...] = a
    return p


def _pin_set(s: SeqSet) -> SeqSet:
    return SeqSet(_pin(s.packed), _pin(s.off), _pin(s.len), None if s.nmask is None else _pin(s.nmask),
                  None if s.nmask is None else _pin(s.nmask_off))


class MultiAligner:
    def __init__(self, devices: Optional[Sequence[int]] = None, chunk: int = 0, slots: int = 4,
                 match: int = 1, mismatch: int = -3, gap_open: int = 5, gap_ext: int = 2):
        self._lib = N.lib()
        self._m = C.c_void_p()
        p = N.Params(match, mismatch, gap_open, gap_ext)
        dev = None if devices is None else (C.c_int * len(devices))(*devices)
        rc = self._lib.ubu_sw_multi_create(C.byref(p), dev, 0 if devices is None else len(devices), slots, chunk, C.byref(self._m))
        if rc != N.OK:
            self._m = C.c_void_p()
            raise N.UbuSwError(rc)
        self.n_devices = int(self._lib.ubu_sw_multi_device_count(self._m))
        self.last_sam_bytes = 0

    def close(self) -> None:
        if getattr(self, "_m", None) is not None and self._m:
            self._lib.ubu_sw_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, reads: SeqSet, windows: SeqSet, idx, res: np.ndarray, pool: np.ndarray, sam_fd: Optional[int], sam_threads: int,
              qname_prefix: str, first_index: int):
        rs, ws = reads.as_struct(), windows.as_struct()
        used = C.c_size_t(0)
        idx_p = None if idx is None else idx.ctypes.data
        if sam_fd is None:
            rc = self._lib.ubu_sw_multi_align(self._m, C.byref(rs), C.byref(ws), idx_p, res.ctypes.data, pool.ctypes.data, pool.shape[0], C.byref(used))
        else:
            names = N.SamNames(qname_prefix.encode("ascii"), first_index, idx_p, None, None, int(pool.shape[0]))
            nbytes = C.c_uint64(0)
            rc = self._lib.ubu_sw_multi_align_sam(self._m, C.byref(rs), C.byref(ws), idx_p, res.ctypes.data, pool.ctypes.data, pool.shape[0],
                                                  C.byref(used), sam_fd, C.byref(names), sam_threads, C.byref(nbytes))
            self.last_sam_bytes = int(nbytes.value)
        return rc, int(used.value)

    def align(self, reads: SeqSet, windows: SeqSet, idx: Optional[np.ndarray] = None, sam_fd: Optional[int] = None, sam_threads: int = 0,
              qname_prefix: str = "r", first_index: int = 0) -> Alignments:
        """Records in input order; the CIGAR pool is sliced per chunk (not dense), offsets in the records are final."""
        n = reads.count
        idx = None if idx is None else np.ascontiguousarray(idx, dtype=np.uint32)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cap = 6 * n + 4096
        while True:
            pool = np.zeros(cap, dtype=np.uint32)
            rc, used = self._call(reads, windows, idx, res, pool, sam_fd, sam_threads, qname_prefix, first_index)
            if rc == N.ERR_CIGAR_POOL and used > cap:
                cap = used + 16
                continue
            if rc != N.OK:
                raise N.UbuSwError(rc, self._lib.ubu_sw_multi_last_error(self._m).decode())
            return Alignments(res, pool[:used])

    def pin_workload(self, w) -> PinnedBatch:
        n = w.n_tasks
        return PinnedBatch(_pin_set(w.reads), _pin_set(w.windows), None if w.pair_ref_idx is None else _pin(np.ascontiguousarray(w.pair_ref_idx, dtype=np.uint32)),
                           pinned_empty((n,), RESULT_DTYPE), pinned_empty((6 * n + 4096,), np.uint32))

    def align_pinned(self, b: PinnedBatch, sam_fd: Optional[int] = None, sam_threads: int = 0) -> int:
        rc, used = self._call(b.reads, b.windows, b.idx, b.results, b.pool, sam_fd, sam_threads, "r", 0)
        if rc != N.OK:
            raise N.UbuSwError(rc, self._lib.ubu_sw_multi_last_error(self._m).decode())
        b.used = used
        return used
