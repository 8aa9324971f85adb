"""Python face of the post-alignment entry points of include/ubu_sw.h:

    project()  <- ReAligner.updateReadAlignment + ReadBlock CIGAR algebra   (ReAligner.java:1293-1368, ReadBlock.java:74-175)
    recount()  <- updateMismatchAndEditDistance's per-read body             (ReAligner.java:279-287, CompareToReference.java:83-123)

Both run on the GPU in libubu_sw.so; this module only marshals numpy arrays."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .aligner import SwAligner
from .seqs import SeqSet, codes_from_text

PLACED_DTYPE = np.dtype([("ref_idx", "<u4"), ("pos", "<i4"), ("cigar_off", "<u4"), ("cigar_n", "<u2"), ("mapq", "<i2"), ("yq", "<i2"), ("ym", "<i2")])
RECOUNT_DTYPE = np.dtype([("xm", "<i4"), ("nm", "<i4"), ("mapq", "<i4"), ("flags", "<u4")])
CONTIG_DTYPE = np.dtype([("pos", "<i4"), ("cigar_off", "<u4"), ("cigar_n", "<u4")])
PROJ_IN_DTYPE = np.dtype([("contig_idx", "<u4"), ("position", "<i4"), ("read_len", "<u4")])
PROJ_OUT_DTYPE = np.dtype([("pos", "<i4"), ("cigar_off", "<u4"), ("cigar_n", "<u2"), ("flags", "<u2")])
assert PLACED_DTYPE.itemsize == 20 and RECOUNT_DTYPE.itemsize == 16 and CONTIG_DTYPE.itemsize == 12
assert PROJ_IN_DTYPE.itemsize == 12 and PROJ_OUT_DTYPE.itemsize == 12

OPS = {"M": 0, "I": 1, "D": 2, "N": 3, "S": 4}


def encode_cigars(cigars: Sequence[Sequence[Tuple[int, int]]]):
    """[(len, op), ...] per record -> (pool uint32, off uint32[], n uint32[])."""
    n = np.array([len(c) for c in cigars], dtype=np.uint32)
    off = np.zeros(len(cigars), dtype=np.uint32)
    if len(cigars):
        off[1:] = np.cumsum(n[:-1])
    pool = np.array([(ln << 4) | op for c in cigars for ln, op in c] or [0], dtype=np.uint32)
    return pool, off, n


def pack_refs(texts: Sequence[str]):
    """Long reference sequences in the 2-bit code + N plane, 64-bit offsets (include/ubu_sw.h ubu_sw_refs)."""
    codes = [codes_from_text(t) for t in texts]
    lens = np.array([c.shape[0] for c in codes], dtype=np.uint32)
    nb = (lens.astype(np.int64) + 3) // 4
    mb = (lens.astype(np.int64) + 7) // 8
    off = np.zeros(len(texts), dtype=np.uint64); moff = np.zeros(len(texts), dtype=np.uint64)
    if len(texts):
        off[1:] = np.cumsum(nb[:-1]); moff[1:] = np.cumsum(mb[:-1])
    packed = np.zeros(int(nb.sum()) + 16, dtype=np.uint8); nmask = np.zeros(int(mb.sum()) + 16, dtype=np.uint8)
    for k, c in enumerate(codes):
        idx = np.arange(c.shape[0])
        isn = c >= 4
        np.bitwise_or.at(packed, int(off[k]) + idx // 4, (np.where(isn, 0, c).astype(np.uint8) << ((idx % 4) * 2).astype(np.uint8)).astype(np.uint8))
        if isn.any():
            sel = idx[isn]
            np.bitwise_or.at(nmask, int(moff[k]) + sel // 8, (1 << (sel % 8)).astype(np.uint8))
    return packed, off, lens, nmask, moff


def load_refs(al: SwAligner, texts: Sequence[str]) -> None:
    packed, off, lens, nmask, moff = pack_refs(texts)
    r = N.Refs()
    r.packed, r.off, r.len, r.nmask, r.nmask_off = packed.ctypes.data, off.ctypes.data, lens.ctypes.data, nmask.ctypes.data, moff.ctypes.data
    r.count, r.packed_bytes, r.nmask_bytes = len(texts), packed.size - 16, nmask.size - 16
    al._check(al._lib.ubu_sw_load_refs(al._h, C.byref(r)))


def recount(al: SwAligner, reads: SeqSet, ref_idx, pos, cigars, mapq, yq, ym) -> np.ndarray:
    """XM / NM / MAPQ per adjusted read (RECOUNT_DTYPE)."""
    n = reads.count
    pool, off, cn = encode_cigars(cigars)
    placed = np.zeros(n, dtype=PLACED_DTYPE)
    placed["ref_idx"], placed["pos"], placed["cigar_off"], placed["cigar_n"] = ref_idx, pos, off, cn
    placed["mapq"], placed["yq"], placed["ym"] = mapq, yq, ym
    out = np.zeros(n, dtype=RECOUNT_DTYPE)
    rs = reads.as_struct()
    al._check(al._lib.ubu_sw_recount_batch(al._h, C.byref(rs), placed.ctypes.data, pool.ctypes.data, int(off[-1] + cn[-1]) if n else 0, out.ctypes.data))
    return out


def project(al: SwAligner, contig_pos, contig_cigars, contig_idx, position, read_len):
    """Genome POS + CIGAR of each read, pushed through its contig's CIGAR. Returns (PROJ_OUT_DTYPE array, list of CIGARs)."""
    cpool, coff, cn = encode_cigars(contig_cigars)
    contigs = np.zeros(len(contig_cigars), dtype=CONTIG_DTYPE)
    contigs["pos"], contigs["cigar_off"], contigs["cigar_n"] = contig_pos, coff, cn
    n = len(contig_idx)
    pin = np.zeros(n, dtype=PROJ_IN_DTYPE)
    pin["contig_idx"], pin["position"], pin["read_len"] = contig_idx, position, read_len
    out = np.zeros(n, dtype=PROJ_OUT_DTYPE)
    cap = 16 * n + 64
    pool = np.zeros(cap, dtype=np.uint32)
    used = C.c_size_t(0)
    ncp = int(coff[-1] + cn[-1]) if len(contig_cigars) else 0
    al._check(al._lib.ubu_sw_project_batch(al._h, contigs.ctypes.data, len(contig_cigars), cpool.ctypes.data, ncp, pin.ctypes.data, n,
                                           out.ctypes.data, pool.ctypes.data, cap, C.byref(used)))
    cigs = [[(int(o) >> 4, int(o) & 15) for o in pool[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["cigar_n"])]] if r["flags"] == 0 else None
            for r in out]
    return out, cigs
