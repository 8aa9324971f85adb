"""Second, independent restatement of SPEC.md (pure Python, small cases only).

TEST INFRASTRUCTURE ONLY -- never imported by ubu_b200/. PARITY UNPINNED: the
reference (mozack/ubu) has no Smith-Waterman of its own (its Aligner forks `bwa`,
src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:18-22,46-56); this file and
oracle/sw_oracle.c are two separately written transcriptions of this repo's SPEC.md
and are cross-checked against each other in tests/test_oracle.py.

Written column-major with an explicit arg-max key on purpose, so that it does not
share visiting order (or code) with the C oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence, Tuple

BASE_N = 4
OP_M, OP_I, OP_D, OP_S = 0, 1, 2, 4
OP_CHARS = {0: "M", 1: "I", 2: "D", 4: "S"}
FLAG_UNALIGNED = 1
NEG = -(10 ** 9)

# Sequence.java:10-13
_CODE = {"A": 0, "C": 1, "T": 2, "G": 3}
_CHAR = "ACTGN"


@dataclass
class Params:
    match: int = 1
    mismatch: int = -3
    gap_open: int = 5
    gap_ext: int = 2


@dataclass
class Result:
    score: int = 0
    ref_start: int = 0
    ref_end: int = 0
    q_start: int = 0
    q_end: int = 0
    edit_distance: int = 0
    flags: int = 0
    cigar: List[Tuple[int, int]] = field(default_factory=list)  # (len, op)

    def cigar_string(self) -> str:
        return "".join(f"{n}{OP_CHARS[o]}" for n, o in self.cigar) or "*"

    def cigar_bam(self) -> List[int]:
        return [(n << 4) | o for n, o in self.cigar]


def encode(text: str) -> List[int]:
    """SPEC section 1: case-fold (CompareToReference.java:90-91), non-ACGT -> N."""
    return [_CODE.get(ch.upper(), BASE_N) for ch in text]


def decode(codes: Sequence[int]) -> str:
    return "".join(_CHAR[c if c < 4 else 4] for c in codes)


def pack2(codes: Sequence[int]) -> Tuple[bytes, bytes]:
    """Sequence.java:31-41 (4 bases/byte, LSB first) + N side bit-plane (SPEC section 1)."""
    packed = bytearray((len(codes) + 3) // 4)
    nmask = bytearray((len(codes) + 7) // 8)
    for k, c in enumerate(codes):
        if c < 4:
            packed[(k * 2) // 8] |= c << ((k * 2) % 8)
        else:
            nmask[k // 8] |= 1 << (k % 8)
    return bytes(packed), bytes(nmask)


def unpack2(packed: bytes, nmask: bytes | None, n: int) -> List[int]:
    """Sequence.java:51-72."""
    out = []
    for k in range(n):
        c = (packed[(k * 2) // 8] >> ((k * 2) % 8)) & 3
        if nmask is not None and (nmask[k // 8] >> (k % 8)) & 1:
            c = BASE_N
        out.append(c)
    return out


def revcomp_text(text: str) -> str:
    """ReverseComplementor.java:15-22,38-62: only A,C,G,T (upper case) are complemented."""
    comp = {"A": "T", "T": "A", "C": "G", "G": "C"}
    return "".join(comp.get(ch, ch) for ch in reversed(text))


def _s(a: int, b: int, p: Params) -> int:
    if a >= BASE_N or b >= BASE_N:
        return 0
    return p.match if a == b else p.mismatch


def align(q: Sequence[int], r: Sequence[int], p: Params = Params()) -> Result:
    L, W = len(q), len(r)
    res = Result()
    if L == 0 or W == 0:
        res.flags = FLAG_UNALIGNED
        return res
    # column-major fill: H[j][i]
    H = [[0] * (L + 1) for _ in range(W + 1)]
    E = [[NEG] * (L + 1) for _ in range(W + 1)]
    F = [[NEG] * (L + 1) for _ in range(W + 1)]
    best_key = None  # maximise (score, -j, -i)
    for j in range(1, W + 1):
        Hj, Hl, Ej, El, Fj = H[j], H[j - 1], E[j], E[j - 1], F[j]
        rj = r[j - 1]
        for i in range(1, L + 1):
            e = max(El[i], Hl[i] - p.gap_open) - p.gap_ext
            f = max(Fj[i - 1], Hj[i - 1] - p.gap_open) - p.gap_ext
            h = max(0, Hl[i - 1] + _s(q[i - 1], rj, p), e, f)
            Ej[i], Fj[i], Hj[i] = e, f, h
            if h > 0:
                key = (h, -j, -i)
                if best_key is None or key > best_key:
                    best_key = key
    if best_key is None:
        res.flags = FLAG_UNALIGNED
        return res
    S, ie, je = best_key[0], -best_key[2], -best_key[1]
    i, j, state, ops, edits = ie, je, "H", [], 0
    while True:
        if state == "H":
            if H[j][i] == 0:
                break
            if H[j][i] == H[j - 1][i - 1] + _s(q[i - 1], r[j - 1], p):
                ops.append(OP_M)
                if q[i - 1] < 4 and r[j - 1] < 4 and q[i - 1] != r[j - 1]:
                    edits += 1
                i, j = i - 1, j - 1
            elif H[j][i] == F[j][i]:
                state = "F"
            else:
                assert H[j][i] == E[j][i]
                state = "E"
        elif state == "F":
            ops.append(OP_I)
            edits += 1
            if i > 1 and F[j][i] == F[j][i - 1] - p.gap_ext:
                i -= 1
            else:
                assert F[j][i] == H[j][i - 1] - p.gap_open - p.gap_ext
                i -= 1
                state = "H"
        else:
            ops.append(OP_D)
            edits += 1
            if j > 1 and E[j][i] == E[j - 1][i] - p.gap_ext:
                j -= 1
            else:
                assert E[j][i] == H[j - 1][i] - p.gap_open - p.gap_ext
                j -= 1
                state = "H"
    res.score, res.ref_start, res.ref_end = S, j + 1, je
    res.q_start, res.q_end, res.edit_distance = i + 1, ie, edits
    cig: List[Tuple[int, int]] = []
    if res.q_start > 1:
        cig.append((res.q_start - 1, OP_S))
    for op in reversed(ops):
        if cig and cig[-1][1] == op and not (op == OP_S):
            cig[-1] = (cig[-1][0] + 1, op)
        else:
            cig.append((1, op))
    if L - ie > 0:
        cig.append((L - ie, OP_S))
    # a leading S must not merge with anything (ops never contain S), so the above is safe
    res.cigar = cig
    return res
