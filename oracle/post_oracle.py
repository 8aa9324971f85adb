"""CPU restatement of the reference's post-alignment steps.

TEST INFRASTRUCTURE ONLY. Unlike the Smith-Waterman oracle, these functions restate code that EXISTS in the reference,
line by line, quirks included; they are pinned to that code (file:line below), though the reference has no test or
fixture for any of them, so there are no golden vectors to check against either.

    project_read      ReAligner.updateReadAlignment            src/main/java/edu/unc/bioinf/ubu/assembly/ReAligner.java:1293-1368
                      ReadBlock.getReadBlocks / getSubBlock /   src/main/java/edu/unc/bioinf/ubu/sam/ReadBlock.java:114-156, 74-88,
                      fillToLength / toCigarString / getTotalLength                                         158-175, 90-99, 101-111
    num_differences   CompareToReference.numDifferences         src/main/java/edu/unc/bioinf/ubu/assembly/CompareToReference.java:83-123
    num_indel_bases   ReAligner.getNumIndelBases                ReAligner.java:316-326
    calc_mapq         ReAligner.calcMappingQuality              ReAligner.java:302-314
    recount           the per-read body of updateMismatchAndEditDistance   ReAligner.java:279-287

CIGAR ops are (length, op) with BAM op codes M=0 I=1 D=2 N=3 S=4.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

M, I, D, N, S = 0, 1, 2, 3, 4
OPC = "MIDNS"


@dataclass
class Block:                      # ReadBlock.java:19-30
    read_start: int
    ref_start: int
    length: int
    type: int

    def reference_length(self) -> int:            # ReadBlock.java:61-69 (sic: 0 only for D)
        return 0 if self.type == D else self.length


def _update_base_positions(read_base: int, ref_base: int, op: int, n: int) -> Tuple[int, int]:   # ReadBlock.java:134-156
    if op == S:
        read_base += n
    elif op in (N, D):
        ref_base += n
    elif op == I:
        read_base += n
    elif op == M:
        read_base += n
        ref_base += n
    else:
        raise ValueError("Case statement didn't deal with cigar op")
    return read_base, ref_base


def read_blocks(alignment_start: int, cigar: Sequence[Tuple[int, int]]) -> List[Block]:          # ReadBlock.java:114-132
    out, rb, fb = [], 1, alignment_start
    for n, op in cigar:
        out.append(Block(rb, fb, n, op))
        rb, fb = _update_base_positions(rb, fb, op, n)
    return out


def _sub_block(b: Block, accumulated: int, position_in_read: int, max_length: int) -> Block:     # ReadBlock.java:74-88
    pib = position_in_read + accumulated - b.read_start + 1
    if b.type in (N, D):
        return Block(accumulated + 1, b.ref_start + pib, b.length - pib, b.type)
    if b.type == S:
        return Block(accumulated + 1, b.ref_start, min(max_length, b.length - pib), b.type)
    return Block(accumulated + 1, b.ref_start + pib, min(max_length, b.length - pib), b.type)


def _total_length(blocks: Sequence[Block]) -> int:                                               # ReadBlock.java:101-111
    return sum(b.length for b in blocks if b.type != D)


def _fill_to_length(blocks: List[Block], read_length: int) -> None:                              # ReadBlock.java:158-175
    cl = _total_length(blocks)
    if cl < read_length:
        last = blocks[-1]
        if last.type in (M, S):
            last.length += read_length - cl
        else:
            rs, fs = _update_base_positions(last.read_start, last.ref_start, last.type, last.length)
            blocks.append(Block(rs, fs, read_length - cl, M))


def project_read(contig_pos: int, contig_cigar: Sequence[Tuple[int, int]], position: int, read_length: int
                 ) -> Optional[Tuple[int, List[Tuple[int, int]]]]:
    """ReAligner.java:1293-1368. Returns (new alignment start, new CIGAR) or None (the reference returns a null read).
    Raises ValueError where the reference throws IllegalStateException (ReAligner.java:1334-1339)."""
    blocks: List[Block] = []
    contig_position, acc = position, 0
    for cb in read_blocks(contig_pos, contig_cigar):
        if cb.read_start + cb.reference_length() >= position + 1:                  # :1306-1307 (uses the ORIGINAL position)
            b = _sub_block(cb, acc, contig_position, read_length - acc)            # :1309-1311
            if b.type == I and b.length != 0:                                      # :1314-1316
                contig_position -= cb.length - b.length
                b.ref_start -= cb.length - b.length
            if b.length != 0:                                                      # :1325
                blocks.append(b)
                if b.type != D:
                    acc += b.length
                if acc > read_length:
                    raise ValueError("Accumulated Length is greater than read length")
                if acc == read_length:
                    break
    if not blocks:
        return None                                                                # :1362-1364
    _fill_to_length(blocks, read_length)                                           # :1356
    return blocks[0].ref_start, [(b.length, b.type) for b in blocks]               # :1357-1358 (no merging of equal neighbours)


def cigar_string(cigar: Sequence[Tuple[int, int]]) -> str:                                       # ReadBlock.java:90-99
    return "".join(f"{n}{OPC[o]}" for n, o in cigar)


def num_differences(read: str, ref: str, alignment_start: int, cigar: Sequence[Tuple[int, int]]) -> int:
    """CompareToReference.java:83-123, including its silence on N (skip) elements: they advance nothing (:99-108)."""
    diffs, ri, fi = 0, 0, alignment_start - 1
    for n, op in cigar:
        if op == M:
            for _ in range(n):
                a, b = read[ri].upper(), ref[fi].upper()
                if a != b and a != "N" and b != "N":
                    diffs += 1
                ri += 1
                fi += 1
        elif op == I:
            ri += n
        elif op == D:
            fi += n
        elif op == S:
            ri += n
    return diffs


def num_indel_bases(cigar: Sequence[Tuple[int, int]]) -> int:                                    # ReAligner.java:316-326
    return sum(n for n, op in cigar if op in (D, I))


def calc_mapq(mapq: int, yq: int, ym: int) -> int:                                               # ReAligner.java:302-314
    if mapq > 0:
        return max(min(yq, 50) - ym * 5, 1)
    return 0


def recount(read: str, ref: str, alignment_start: int, cigar, mapq: int, yq: int, ym: int) -> Tuple[int, int, int]:
    """ReAligner.java:279-287: (XM, NM, MAPQ) of an adjusted read."""
    xm = num_differences(read, ref, alignment_start, cigar)
    return xm, xm + num_indel_bases(cigar), calc_mapq(mapq, yq, ym)
