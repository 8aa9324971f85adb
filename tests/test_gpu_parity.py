"""GPU parity: the CUDA path (through the C ABI) vs the CPU oracle, field for field (SPEC.md).
Reference parity is UNPINNED (no Smith-Waterman exists in mozack/ubu); 'exact' here means exact
against this repo's SPEC oracle."""
import numpy as np
import pytest

import oracle_lib as O
import ubu_b200 as U
from ubu_b200 import workloads as WL

pytestmark = pytest.mark.gpu

FIELDS = ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance")


def check_against_oracle(al: U.Alignments, reads: U.SeqSet, windows: U.SeqSet, idx, params=(1, -3, 5, 2), nthreads=8):
    n = reads.count
    q = [reads.codes(k) for k in range(n)]
    r = [windows.codes(k) for k in range(windows.count)]
    ores, ocig = O.align_batch(q, r, idx, params, nthreads=nthreads)
    bad = []
    for k in range(n):
        g = al.results[k]
        ok = all(int(g[f]) == int(ores[k][f]) for f in FIELDS)
        ok = ok and (int(g["flags"]) & 1) == (int(ores[k]["flags"]) & 1)
        ok = ok and al.cigar(k) == ocig[k]
        if not ok:
            bad.append((k, {f: int(g[f]) for f in FIELDS}, al.cigar_string(k), {f: int(ores[k][f]) for f in FIELDS}, ocig[k]))
    assert not bad, f"{len(bad)}/{n} tasks differ; first: {bad[:3]}"


@pytest.fixture(scope="module")
def aligner():
    with U.SwAligner() as a:
        yield a


def test_tiny_known(aligner):
    reads = U.pack_texts(["ACGTACGTTAGC", "TTTTTTTT", "GATTACA"])
    wins = U.pack_texts(["TTTTACGTACGTAGCTTTT", "ACACACACAC", "CCGATTTACAGG"])
    al = aligner.align(reads, wins)
    check_against_oracle(al, reads, wins, None)
    assert al.cigar_string(0) == "8M4S" and int(al.results[0]["ref_start"]) == 5


@pytest.mark.parametrize("name,n,L,W,paired", [("cfg1", 600, 100, 1000, False), ("cfg2", 1000, 76, 500, True),
                                                ("cfg3", 300, 150, 2000, False), ("l250", 200, 250, 700, False),
                                                ("l50", 500, 50, 300, False), ("l33", 300, 33, 200, False)])
def test_uniform_sets(aligner, name, n, L, W, paired):
    w = WL.uniform(name, n, L, W, seed=11, paired=paired, n_windows=None if paired else max(n // 3, 1))
    al = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    check_against_oracle(al, w.reads, w.windows, w.pair_ref_idx)


def test_mixed_lengths_indel_rich(aligner):
    w = WL.cfg4(700, seed=5)
    al = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    check_against_oracle(al, w.reads, w.windows, w.pair_ref_idx)


def test_with_n(aligner):
    base = WL.uniform("n", 400, 100, 600, seed=3, keep_codes=True)
    w = WL.with_n(base, rate=0.02)
    al = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    check_against_oracle(al, w.reads, w.windows, w.pair_ref_idx)


def test_tie_stress(aligner):
    rng = np.random.default_rng(9)
    texts_r, texts_w = [], []
    for k in range(300):
        kind = k % 5
        if kind == 0:      # homopolymers
            b = "ACGT"[k % 4]; texts_r.append(b * int(rng.integers(1, 120))); texts_w.append(b * int(rng.integers(1, 300)))
        elif kind == 1:    # tandem repeats
            u = "".join(rng.choice(list("ACGT"), size=int(rng.integers(2, 7))))
            texts_r.append((u * 60)[: int(rng.integers(20, 200))]); texts_w.append((u * 200)[: int(rng.integers(50, 600))])
        elif kind == 2:    # read occurs twice in the window with equal score
            r = "".join(rng.choice(list("ACGT"), size=int(rng.integers(20, 120))))
            f = "".join(rng.choice(list("ACGT"), size=int(rng.integers(0, 40))))
            texts_r.append(r); texts_w.append(f + r + f + r + f)
        elif kind == 3:    # random vs random (low scores, wide bands)
            texts_r.append("".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 256)))))
            texts_w.append("".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 500)))))
        else:              # read longer than window
            w_ = "".join(rng.choice(list("ACGT"), size=int(rng.integers(5, 60))))
            texts_r.append("".join(rng.choice(list("ACGT"), size=10)) + w_ + "".join(rng.choice(list("ACGT"), size=30)))
            texts_w.append(w_)
    reads, wins = U.pack_texts(texts_r), U.pack_texts(texts_w)
    al = aligner.align(reads, wins)
    check_against_oracle(al, reads, wins, None)


def test_empty_and_degenerate(aligner):
    al = aligner.align(U.pack_texts([]), U.pack_texts([]))
    assert len(al) == 0
    reads = U.pack_texts(["", "A", "ACGT", "NNNN", "ACGT"])
    wins = U.pack_texts(["ACGT", "", "TGCA", "ACGT", "NNNN"])
    al = aligner.align(reads, wins)
    check_against_oracle(al, reads, wins, None)
    assert all(int(al.results[k]["flags"]) & 1 for k in (0, 1, 3, 4))


@pytest.mark.parametrize("params", [(2, -1, 3, 1), (1, -1, 0, 1), (5, -4, 10, 1), (1, -3, 5, 2), (3, 0, 2, 2)])
def test_other_scoring(params):
    w = WL.uniform("p", 300, 90, 400, seed=21, sub_rate=0.05, indel_rate=0.02, geom_p=0.4)
    with U.SwAligner(*params) as a:
        al = a.align(w.reads, w.windows, w.pair_ref_idx)
    check_against_oracle(al, w.reads, w.windows, w.pair_ref_idx, params)
