"""CIGAR projection and the XM/NM/MAPQ recount, GPU vs the line-by-line restatement of the
reference's Java (oracle/post_oracle.py). The CPU tests pin the restatement on hand-worked cases."""
import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import post_oracle as PO  # noqa: E402

import ubu_b200 as U
from ubu_b200 import post as POST

M, I, D, N, S = 0, 1, 2, 3, 4


def test_projection_restatement_hand_cases():
    # read inside the first M block
    assert PO.project_read(1000, [(100, M)], 10, 30) == (1010, [(30, M)])
    # read spanning a deletion: the D block is carried whole (ReadBlock.java:79-81)
    assert PO.project_read(1000, [(50, M), (10, D), (50, M)], 40, 30) == (1040, [(10, M), (10, D), (20, M)])
    # read spanning an intron: the reference counts the N block into the accumulated READ length (only D is excluded,
    # ReAligner.java:1328-1330) and then throws IllegalStateException (:1334-1339) -- kept, the GPU path flags it
    with pytest.raises(ValueError):
        PO.project_read(500, [(20, M), (300, N), (40, M)], 5, 30)
    assert PO.project_read(500, [(20, M), (3, N), (40, M)], 5, 30) == (505, [(15, M), (3, N), (12, M)])
    # read starting inside an insertion: start is pulled back by the skipped part (ReAligner.java:1314-1316)
    pos, cig = PO.project_read(1000, [(50, M), (6, I), (50, M)], 52, 20)
    assert cig[0][1] == I and cig[0][0] == 4 and PO.cigar_string(cig) == "4I16M"
    # past the end of the contig: fillToLength extends the last M (ReadBlock.java:163-166)
    assert PO.project_read(1000, [(100, M)], 90, 30) == (1090, [(30, M)])
    # leading soft clip of the contig shifts read coordinates, not the reference start
    assert PO.project_read(1000, [(5, S), (95, M)], 10, 20) == (1005, [(20, M)])
    # placement beyond every block -> null read
    assert PO.project_read(1000, [(50, M)], 80, 20) is None


def test_recount_restatement_hand_cases():
    ref = "AACCGGTTACGTNACGTTTGACCA"
    assert PO.num_differences("CCGGTTAC", ref, 3, [(8, M)]) == 0
    assert PO.num_differences("CCGATTAC", ref, 3, [(8, M)]) == 1
    assert PO.num_differences("ccgnttac", ref, 3, [(8, M)]) == 0                     # case-folded, N never mismatches
    assert PO.num_differences("ACGTAACGT", ref, 9, [(9, M)]) == 0                    # reference N at position 13
    assert PO.num_differences("CCGGAATTAC", ref, 3, [(4, M), (2, I), (4, M)]) == 0
    assert PO.num_differences("CCGGAC", ref, 3, [(4, M), (2, D), (2, M)]) == 0
    # quirk kept: an N (skip) element advances nothing (CompareToReference.java:99-108)
    assert PO.num_differences("CCGGTT", ref, 3, [(2, M), (5, N), (4, M)]) == PO.num_differences("CCGGTT", ref, 3, [(6, M)])
    assert PO.num_indel_bases([(4, M), (2, I), (3, D), (1, S)]) == 5
    assert PO.calc_mapq(0, 60, 0) == 0 and PO.calc_mapq(30, 60, 0) == 50 and PO.calc_mapq(30, 20, 3) == 5 and PO.calc_mapq(30, 10, 3) == 1


def random_contig(rng):
    ops, prev = [], None
    if rng.random() < 0.3:
        ops.append((rng.randint(1, 12), S)); prev = S
    for _ in range(rng.randint(1, 6)):
        ops.append((rng.randint(5, 120), M))
        if rng.random() < 0.7:
            ops.append((rng.randint(1, 30) if rng.random() < 0.8 else rng.randint(50, 400), rng.choice([I, D, N])))
    if ops[-1][1] != M:
        ops.append((rng.randint(5, 80), M))
    if rng.random() < 0.3:
        ops.append((rng.randint(1, 12), S))
    return rng.randint(1, 100000), ops


@pytest.mark.gpu
def test_projection_gpu_equals_restatement():
    rng = random.Random(11)
    contigs = [random_contig(rng) for _ in range(200)]
    cidx, position, rlen = [], [], []
    for _ in range(5000):
        c = rng.randrange(len(contigs))
        clen = sum(n for n, o in contigs[c][1] if o in (M, I, S))
        cidx.append(c); rlen.append(rng.choice([50, 76, 100, 150])); position.append(rng.randint(0, clen + 10))
    with U.SwAligner() as al:
        out, cigs = POST.project(al, [p for p, _ in contigs], [c for _, c in contigs], cidx, position, rlen)
    n_null = n_bad = 0
    for k in range(len(cidx)):
        try:
            want = PO.project_read(contigs[cidx[k]][0], contigs[cidx[k]][1], position[k], rlen[k])
            if want is not None and any(ln <= 0 for ln, _ in want[1]):
                want = "malformed"
        except ValueError:
            want = "malformed"
        if want is None:
            assert out[k]["flags"] == 1; n_null += 1
        elif want == "malformed":
            assert out[k]["flags"] == 2; n_bad += 1
        else:
            assert out[k]["flags"] == 0 and int(out[k]["pos"]) == want[0] and cigs[k] == want[1], (k, contigs[cidx[k]], position[k], rlen[k])
    assert n_null > 0 and n_null + n_bad < len(cidx) // 2


@pytest.mark.gpu
def test_recount_gpu_equals_restatement():
    rng = random.Random(12)
    refs = ["".join(rng.choice("ACGT" if rng.random() > 0.01 else "N") for _ in range(L)) for L in (5000, 12000, 333)]
    reads, ref_idx, pos, cigars, mapq, yq, ym = [], [], [], [], [], [], []
    for _ in range(4000):
        r = rng.randrange(len(refs)); G = len(refs[r])
        ops, rl, fl = [], 0, 0
        if rng.random() < 0.2:
            n = rng.randint(1, 8); ops.append((n, S)); rl += n
        for b in range(rng.randint(1, 3)):
            n = rng.randint(5, 60); ops.append((n, M)); rl += n; fl += n
            if b < 2 and rng.random() < 0.5:
                n = rng.randint(1, 9); op = rng.choice([I, D, N]); ops.append((n, op))
                if op == I: rl += n
                elif op == D: fl += n
        if ops[-1][1] != M:
            ops.append((7, M)); rl += 7; fl += 7
        p = rng.randint(1, max(1, G - fl - 20))
        seq = []
        # build the read from the reference along the CIGAR (N element advances nothing, as in the reference), then mutate
        ri_ref = p - 1
        for n, op in ops:
            if op == M:
                seq.append(refs[r][ri_ref: ri_ref + n]); ri_ref += n
            elif op in (I, S):
                seq.append("".join(rng.choice("ACGT") for _ in range(n)))
            elif op == D:
                ri_ref += n
        seq = list("".join(seq))
        for k in range(len(seq)):
            x = rng.random()
            if x < 0.04: seq[k] = rng.choice("ACGT")
            elif x < 0.05: seq[k] = "N"
            elif x < 0.08: seq[k] = seq[k].lower()
        reads.append("".join(seq)); ref_idx.append(r); pos.append(p); cigars.append(ops)
        mapq.append(rng.choice([0, 1, 20, 37, 60])); yq.append(rng.randint(0, 70)); ym.append(rng.randint(0, 6))
    with U.SwAligner() as al:
        POST.load_refs(al, refs)
        out = POST.recount(al, U.pack_texts(reads, force_nmask=True), ref_idx, pos, cigars, mapq, yq, ym)
    for k in range(len(reads)):
        want = PO.recount(reads[k], refs[ref_idx[k]], pos[k], cigars[k], mapq[k], yq[k], ym[k])
        assert out[k]["flags"] == 0 and (int(out[k]["xm"]), int(out[k]["nm"]), int(out[k]["mapq"])) == want, (k, cigars[k])
