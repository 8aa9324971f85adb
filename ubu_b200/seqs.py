"""Host-side sequence containers: 2-bit packing in the reference's own convention.

Code A=0 C=1 T=2 G=3, four bases per byte, LSB first
(src/main/java/edu/unc/bioinf/ubu/assembly/Sequence.java:10-13,31-41); N (and any non-ACGT byte,
SPEC.md section 1) goes to a 1-bit side plane. Everything here is vectorised numpy: this is batch
plumbing, not the alignment path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional, Sequence, Union

import numpy as np

from . import _native as N

_CODE_LUT = np.full(256, 4, dtype=np.uint8)
for _ch, _c in (("A", 0), ("C", 1), ("T", 2), ("G", 3)):
    _CODE_LUT[ord(_ch)] = _c
    _CODE_LUT[ord(_ch.lower())] = _c
_CHARS = np.frombuffer(b"ACTGN", dtype=np.uint8)


def codes_from_text(text: Union[str, bytes]) -> np.ndarray:
    b = text.encode("ascii") if isinstance(text, str) else bytes(text)
    return _CODE_LUT[np.frombuffer(b, dtype=np.uint8)]


def text_from_codes(codes: np.ndarray) -> str:
    return _CHARS[np.minimum(codes, 4)].tobytes().decode("ascii")


@dataclass
class SeqSet:
    """A batch of 2-bit packed sequences laid out exactly as include/ubu_sw.h's ubu_sw_seqs wants them."""

    packed: np.ndarray           # uint8
    off: np.ndarray              # uint32 [count]  byte offset of sequence k in packed
    len: np.ndarray              # uint16 [count]
    nmask: Optional[np.ndarray]  # uint8 or None
    nmask_off: Optional[np.ndarray]

    @property
    def count(self) -> int:
        return int(self.len.shape[0])

    def as_struct(self) -> N.Seqs:
        s = N.Seqs()
        s.packed = self.packed.ctypes.data if self.packed.size else None
        s.off = self.off.ctypes.data if self.count else None
        s.len = self.len.ctypes.data if self.count else None
        s.nmask = self.nmask.ctypes.data if self.nmask is not None else None
        s.nmask_off = self.nmask_off.ctypes.data if self.nmask is not None else None
        s.count = self.count
        s.packed_bytes = int(self.packed.size)
        s.nmask_bytes = int(self.nmask.size) if self.nmask is not None else 0
        return s

    def codes(self, k: int) -> np.ndarray:
        """Unpacked codes (0..3, 4 = N) of sequence k (Sequence.java:51-72)."""
        n = int(self.len[k])
        idx = np.arange(n)
        c = (self.packed[int(self.off[k]) + idx // 4] >> ((idx % 4) * 2)) & 3
        if self.nmask is not None:
            nb = (self.nmask[int(self.nmask_off[k]) + idx // 8] >> (idx % 8)) & 1
            c = np.where(nb == 1, 4, c)
        return c.astype(np.uint8)

    def text(self, k: int) -> str:
        return text_from_codes(self.codes(k))

    def slice(self, lo: int, hi: int) -> "SeqSet":
        """Sequences [lo, hi) as a chunk of its own: zero-copy views of the packed bytes those sequences occupy, offsets
        rebased to the view (so a submit() of the chunk uploads only the chunk). Assumes offsets ascend, as every packer
        here produces them."""
        if hi <= lo:
            return SeqSet(self.packed[:0], self.off[:0], self.len[:0], None if self.nmask is None else self.nmask[:0],
                          None if self.nmask is None else self.nmask_off[:0])
        b0 = int(self.off[lo])
        b1 = int(self.off[hi - 1]) + (int(self.len[hi - 1]) + 3) // 4
        nmask = nmask_off = None
        if self.nmask is not None:
            m0 = int(self.nmask_off[lo])
            m1 = int(self.nmask_off[hi - 1]) + (int(self.len[hi - 1]) + 7) // 8
            nmask, nmask_off = self.nmask[m0:m1], (self.nmask_off[lo:hi] - np.uint32(m0)).astype(np.uint32)
        return SeqSet(self.packed[b0:b1], (self.off[lo:hi] - np.uint32(b0)).astype(np.uint32), self.len[lo:hi], nmask, nmask_off)


def pack_codes(code_arrays: Sequence[np.ndarray], force_nmask: bool = False) -> SeqSet:
    """Packs a list of code arrays (values 0..3, >=4 = N)."""
    n = len(code_arrays)
    lens = np.fromiter((a.shape[0] for a in code_arrays), dtype=np.int64, count=n)
    if n and lens.max() > N.MAX_WINDOW:
        raise ValueError("sequence longer than 65535")
    nbytes = (lens + 3) // 4
    off = np.zeros(n, dtype=np.int64)
    if n:
        np.cumsum(nbytes[:-1], out=off[1:])
    total = int(nbytes.sum())
    if total >= 2 ** 32:
        raise ValueError("batch larger than 4 GiB packed; split it")
    flat = np.concatenate(code_arrays).astype(np.uint8) if n else np.zeros(0, np.uint8)
    return _pack_flat(flat, lens, off, total, force_nmask)


def _pack_flat(flat: np.ndarray, lens: np.ndarray, off: np.ndarray, total: int, force_nmask: bool) -> SeqSet:
    n = lens.shape[0]
    # position of every base inside its own sequence
    starts = np.zeros(n, dtype=np.int64)
    if n:
        np.cumsum(lens[:-1], out=starts[1:])
    seq_id = np.repeat(np.arange(n), lens)
    pos = np.arange(flat.shape[0]) - starts[seq_id]
    isn = flat >= 4
    byte_idx = off[seq_id] + pos // 4
    vals = (np.where(isn, 0, flat).astype(np.uint8) << ((pos % 4) * 2).astype(np.uint8)).astype(np.uint8)
    packed = np.zeros(total + 16, dtype=np.uint8)
    np.bitwise_or.at(packed, byte_idx, vals)
    nmask = nmask_off = None
    if force_nmask or bool(isn.any()):
        mbytes = (lens + 7) // 8
        moff = np.zeros(n, dtype=np.int64)
        if n:
            np.cumsum(mbytes[:-1], out=moff[1:])
        nmask = np.zeros(int(mbytes.sum()) + 16, dtype=np.uint8)
        if isn.any():
            sel = np.nonzero(isn)[0]
            np.bitwise_or.at(nmask, moff[seq_id[sel]] + pos[sel] // 8, (1 << (pos[sel] % 8)).astype(np.uint8))
        nmask_off = moff.astype(np.uint32)
    return SeqSet(packed[:total] if total else packed[:0], off.astype(np.uint32), lens.astype(np.uint16), nmask, nmask_off)


def pack_texts(texts: Iterable[Union[str, bytes]], force_nmask: bool = False) -> SeqSet:
    return pack_codes([codes_from_text(t) for t in texts], force_nmask)


def pack_uniform(codes: np.ndarray) -> SeqSet:
    """Fast path for a [count, length] uint8 code matrix (synthetic workloads); codes >= 4 are N and go to the bit plane."""
    n, L = codes.shape
    nb = (L + 3) // 4
    pad = nb * 4 - L
    isn = codes >= 4
    has_n = bool(isn.any())
    c = np.where(isn, 0, codes).astype(np.uint8) if has_n else codes
    c = np.pad(c, ((0, 0), (0, pad))) if pad else c
    c = c.reshape(n, nb, 4).astype(np.uint8)
    packed = (c[:, :, 0] | (c[:, :, 1] << 2) | (c[:, :, 2] << 4) | (c[:, :, 3] << 6)).astype(np.uint8).reshape(-1)
    if n * nb >= 2 ** 32:
        raise ValueError("batch larger than 4 GiB packed; split it")
    off = (np.arange(n, dtype=np.int64) * nb).astype(np.uint32)
    nmask = nmask_off = None
    if has_n:
        mb = (L + 7) // 8
        nmask = np.ascontiguousarray(np.packbits(isn, axis=1, bitorder="little")).reshape(-1)
        assert nmask.size == n * mb
        nmask_off = (np.arange(n, dtype=np.int64) * mb).astype(np.uint32)
    return SeqSet(np.ascontiguousarray(packed), off, np.full(n, L, dtype=np.uint16), nmask, nmask_off)


def pack_text_batch(text: np.ndarray, text_off: np.ndarray, lens: np.ndarray, threads: int = 0, with_nmask: bool = True) -> SeqSet:
    """ubu_sw_pack_batch: packs a whole batch of ASCII sequences on the host's cores (32 or 8 characters per step per thread).
    text: uint8 buffer; sequence k = text[text_off[k] : text_off[k] + lens[k]]. The N plane is dropped when no sequence has one."""
    import ctypes as C

    text = np.ascontiguousarray(text, dtype=np.uint8)
    text_off = np.ascontiguousarray(text_off, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint16)
    n = int(lens.shape[0])
    pb = (lens.astype(np.int64) + 3) // 4
    nb = (lens.astype(np.int64) + 7) // 8
    off = np.zeros(n, dtype=np.int64); noff = np.zeros(n, dtype=np.int64)
    if n:
        off[1:] = np.cumsum(pb)[:-1]; noff[1:] = np.cumsum(nb)[:-1]
    if n and off[-1] + pb[-1] >= 2 ** 32:
        raise ValueError("batch larger than 4 GiB packed; split it")
    packed = np.zeros(int(pb.sum()), dtype=np.uint8)
    nmask = np.zeros(int(nb.sum()), dtype=np.uint8) if with_nmask else None
    off32, noff32 = off.astype(np.uint32), noff.astype(np.uint32)
    cnt = C.c_uint32(0)
    rc = N.lib().ubu_sw_pack_batch(text.ctypes.data, text_off.ctypes.data, lens.ctypes.data, n, packed.ctypes.data, off32.ctypes.data,
                                   nmask.ctypes.data if nmask is not None else None, noff32.ctypes.data if nmask is not None else None,
                                   threads, C.byref(cnt))
    if rc != N.OK:
        raise N.UbuSwError(rc)
    if nmask is not None and cnt.value == 0:
        nmask = None
    return SeqSet(packed, off32, lens, nmask, noff32 if nmask is not None else None)


def reverse_complement(text: str) -> str:
    """ReverseComplementor.java:38-62 through the library helper (A<->T, C<->G, others unchanged)."""
    import ctypes as C

    b = text.encode("ascii")
    out = C.create_string_buffer(len(b))
    N.lib().ubu_sw_revcomp(b, len(b), out)
    return out.raw.decode("ascii")
