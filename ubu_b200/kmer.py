"""Python face of include/ubu_kmer.h: the assembler's k-mer graph build on the GPU.

Mirrors the graph-building half of the reference's `Assembler` (src/main/java/edu/unc/bioinf/ubu/assembly/Assembler.java):
`addToGraph` (:452-474), `filterLowFrequencyNodes` (:201-227), `identifyRootNodes` (:270-276), with the reference's setter
names as constructor arguments (`setKmerSize` :165, `setMinNodeFrequncy` :181). The contig walk (`buildContigs`, :278-367) stays
on the host and reads the node table this returns. The work always runs in libubu_sw.so on the GPU; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from . import _native as N
from .seqs import SeqSet

NODE_DTYPE = np.dtype([("kmer", "<u8", (2,)), ("region", "<u4"), ("count", "<u4"), ("first_read", "<u4"), ("first_pos", "<u2"),
                       ("flags", "<u2"), ("to_node", "<u4", (4,)), ("from_node", "<u4", (4,))])
assert NODE_DTYPE.itemsize == 64

_BASES = "ACTG"        # Sequence.java:10-13


def kmer_string(node, k: int) -> str:
    v = int(node["kmer"][0]) | (int(node["kmer"][1]) << 64)
    return "".join(_BASES[(v >> (2 * j)) & 3] for j in range(k))


class KmerGraphBuilder:
    """One handle = one GPU (ubu_kmer_create)."""

    def __init__(self, kmer_size: int, min_node_frequency: int, device: int = 0):
        self._lib = N.lib()
        self._h = C.c_void_p()
        rc = self._lib.ubu_kmer_create(device, C.byref(self._h))
        if rc != N.OK:
            self._h = C.c_void_p()
            raise N.UbuSwError(rc)
        self.params = N.KmerParams(kmer_size, min_node_frequency)
        self.kmer_size = kmer_size
        self._keep = None

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.ubu_kmer_destroy(self._h)
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
            raise N.UbuSwError(rc, self._lib.ubu_kmer_last_error(self._h).decode())

    @staticmethod
    def _regions(regions, n):
        if regions is None:
            return None, None
        a = np.ascontiguousarray(regions, dtype=np.uint32)
        if a.shape[0] != n:
            raise ValueError("regions must have one entry per read")
        return a, a.ctypes.data

    def build(self, reads: SeqSet, regions: Optional[np.ndarray] = None, max_nodes: int = 0) -> Tuple[np.ndarray, int]:
        """ubu_kmer_build: returns (node table, reads skipped for N)."""
        reg, reg_p = self._regions(regions, reads.count)
        rs = reads.as_struct()
        cap = max(int(max_nodes), 1 << 16)
        while True:
            out = np.zeros(cap, dtype=NODE_DTYPE)
            n_nodes, skipped = C.c_size_t(0), C.c_uint32(0)
            rc = self._lib.ubu_kmer_build(self._h, C.byref(self.params), C.byref(rs), reg_p, max_nodes, out.ctypes.data, cap,
                                          C.byref(n_nodes), C.byref(skipped))
            if rc == N.ERR_CAPACITY and not max_nodes and n_nodes.value > cap:
                cap = int(n_nodes.value)
                continue
            self._check(rc)
            return out[: n_nodes.value], int(skipped.value)

    # staged calls (device-resident timing)
    def stage(self, reads: SeqSet, regions: Optional[np.ndarray] = None, max_nodes: int = 0) -> None:
        reg, reg_p = self._regions(regions, reads.count)
        rs = reads.as_struct()
        self._keep = (reads, reg, rs)
        self._check(self._lib.ubu_kmer_stage(self._h, C.byref(self.params), C.byref(rs), reg_p, max_nodes))

    def run(self) -> None:
        self._check(self._lib.ubu_kmer_run(self._h))

    def fetch(self, cap: int) -> Tuple[np.ndarray, int]:
        out = np.zeros(max(cap, 1), dtype=NODE_DTYPE)
        n_nodes, skipped = C.c_size_t(0), C.c_uint32(0)
        self._check(self._lib.ubu_kmer_fetch(self._h, out.ctypes.data, out.shape[0], C.byref(n_nodes), C.byref(skipped)))
        return out[: n_nodes.value], int(skipped.value)

    def last_run_ms(self) -> Tuple[float, float, float, float]:
        ms = (C.c_float * 4)()
        self._check(self._lib.ubu_kmer_last_run_ms(self._h, ms))
        return tuple(float(x) for x in ms)

    def last_run_stats(self) -> Tuple[int, int, int]:
        occ, tb, ln = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        self._check(self._lib.ubu_kmer_last_run_stats(self._h, C.byref(occ), C.byref(tb), C.byref(ln)))
        return int(occ.value), int(tb.value), int(ln.value)


def contigs_from_table(nodes: np.ndarray, k: int, min_contig_length: int = 0, max_depth: int = 10000,
                       filtered: Optional[np.ndarray] = None) -> List[str]:
    """Host-side contig walk over a returned node table (Assembler.buildContigs / buildContig, Assembler.java:278-367, without
    the region-ratio test and the potential-contig cap): roots in table order, successors in Node.toNodes order with filtered
    nodes dropped. Recursive like the reference; meant for region-sized graphs. `filtered` overrides the table's filter flags
    (roots are then re-derived: unfiltered nodes without an unfiltered predecessor)."""
    flags = nodes["flags"].copy()
    if filtered is not None:
        flags = np.where(filtered, N.KMER_FLAG_FILTERED, 0).astype(np.uint16)
        for j in range(nodes.shape[0]):
            live = [int(t) for t in nodes[j]["from_node"] if t != N.KMER_NONE and not filtered[int(t)]]
            if not filtered[j] and not live:
                flags[j] |= N.KMER_FLAG_ROOT
    seqs = [kmer_string(nd, k) for nd in nodes]
    out: List[str] = []

    def walk(j: int, visited: frozenset, contig: str, depth: int) -> None:
        if depth > max_depth:
            raise RecursionError("DepthExceededException")
        if j in visited:
            if len(contig) >= min_contig_length:
                out.append(contig)
            return
        visited = visited | {j}
        succ = [int(t) for t in nodes[j]["to_node"] if t != N.KMER_NONE and not (flags[int(t)] & N.KMER_FLAG_FILTERED)]
        if not succ:
            contig = contig + seqs[j]
            if len(contig) >= min_contig_length:
                out.append(contig)
            return
        contig = contig + seqs[j][0]
        for t in succ:
            walk(t, visited, contig, depth + 1)

    for j in range(nodes.shape[0]):
        if flags[j] & N.KMER_FLAG_ROOT:
            walk(j, frozenset(), "", 1)
    return out
