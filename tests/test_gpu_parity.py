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


def check_every_record_fast(al: U.Alignments, w, params=(1, -3, 5, 2), nthreads=16):
    """Every record and every CIGAR of a large run against the oracle, vectorised (no per-task Python work except unpacking)."""
    import ctypes as C

    n = w.n_tasks
    rl, wl = w.reads.len.astype(np.uint32), w.windows.len.astype(np.uint32)
    def flat(sq):
        ln = sq.len.astype(np.int64)
        seq = np.repeat(np.arange(sq.count), ln)
        start = np.concatenate([[0], np.cumsum(ln)[:-1]])
        pos = np.arange(int(ln.sum())) - start[seq]
        codes = (sq.packed[sq.off.astype(np.int64)[seq] + pos // 4] >> ((pos % 4) * 2).astype(np.uint8)) & 3
        if sq.nmask is not None:
            codes = np.where((sq.nmask[sq.nmask_off.astype(np.int64)[seq] + pos // 8] >> (pos % 8).astype(np.uint8)) & 1, 4, codes)
        return np.ascontiguousarray(codes, dtype=np.uint8), start.astype(np.uint64)
    q, qoff = flat(w.reads); r, roff = flat(w.windows)
    idx = None if w.pair_ref_idx is None else np.ascontiguousarray(w.pair_ref_idx, dtype=np.uint32)
    res = np.zeros(n, dtype=O.SWO_RESULT)
    stride = int(rl.max()) + 8
    cig = np.zeros(n * stride, dtype=np.uint32)
    p = O.SwoParams(*params)
    st = O.lib().swo_align_batch(q.ctypes.data, qoff.ctypes.data, rl.ctypes.data, r.ctypes.data, roff.ctypes.data, wl.ctypes.data,
                                 None if idx is None else idx.ctypes.data, n, C.byref(p), res.ctypes.data, cig.ctypes.data, stride, nthreads)
    assert st == 0
    g = al.results
    for f in FIELDS:
        bad = np.nonzero(g[f] != res[f])[0]
        assert bad.size == 0, f"{bad.size}/{n} records differ in {f}; first task {bad[:5]}: gpu {g[f][bad[:5]]} oracle {res[f][bad[:5]]}"
    assert np.array_equal(g["flags"] & 1, (res["flags"] & 1).astype(g["flags"].dtype))
    assert np.array_equal(g["cigar_n"].astype(np.int64), res["cigar_n"].astype(np.int64))
    cn = res["cigar_n"].astype(np.int64)
    task = np.repeat(np.arange(n), cn)
    k = np.arange(int(cn.sum())) - np.concatenate([[0], np.cumsum(cn)[:-1]])[task]
    assert np.array_equal(al.cigar_pool[g["cigar_off"].astype(np.int64)[task] + k], cig[task * stride + k]), "CIGAR ops differ"
    return res


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


def test_mixed_read_lengths_one_class_equal_windows_exact_buffers(aligner):
    """Regression (uniform fast path): trimmed reads of one length class against equal-length windows, shortest read last."""
    rng = np.random.default_rng(31)
    wins_t = ["".join(rng.choice(list("ACGT"), size=300)) for _ in range(40)]
    reads_t = [wins_t[k][50:50 + L] for k, L in enumerate([104] * 20 + [103, 102, 101] * 6 + [100, 99])]
    reads, wins = U.pack_texts(reads_t), U.pack_texts(wins_t)
    al = aligner.align(reads, wins)
    check_against_oracle(al, reads, wins, None)


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


def test_golden_vectors_through_the_c_abi():
    """Every committed golden case (tests/golden/sw_golden_v1.json), grouped by scoring parameters."""
    import json
    import os

    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sw_golden_v1.json")))
    by_params = {}
    for c in gold["cases"]:
        by_params.setdefault(tuple(c["params"]), []).append(c)
    for params, cases in by_params.items():
        reads = U.pack_texts([c["read"] for c in cases], force_nmask=True)
        wins = U.pack_texts([c["window"] for c in cases], force_nmask=True)
        with U.SwAligner(*params) as a:
            al = a.align(reads, wins)
        for k, c in enumerate(cases):
            e = c["expect"]
            got = {f: int(al.results[k][f]) for f in FIELDS}
            assert got == {f: e[f] for f in FIELDS}, (c["name"], got, e)
            assert al.cigar_string(k) == e["cigar"], c["name"]
            assert (int(al.results[k]["flags"]) & 1) == (e["flags"] & 1)


def test_full_size_cfg2_properties(aligner):
    """BASELINE.json configs[1] at full size (1M reads): properties that need no oracle, plus an oracle-checked sample."""
    w = WL.cfg2(1_000_000, keep_codes=True)
    al = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    r = al.results
    n = len(al)
    assert n == 1_000_000 and int((r["flags"] & 1).sum()) == 0
    assert int(r["score"].max()) <= 76 and int(r["score"].min()) > 20
    assert np.all(r["q_end"] >= r["q_start"]) and np.all(r["ref_end"] >= r["ref_start"]) and np.all(r["ref_end"] <= 500)
    # CIGAR covers the read; reference span matches; every CIGAR slice lies inside the pool and slices do not overlap
    ops = al.cigar_pool
    lens, kinds = (ops >> 4).astype(np.int64), (ops & 15)
    owner = np.repeat(np.arange(n), r["cigar_n"].astype(np.int64)[np.argsort(r["cigar_off"], kind="stable")])
    order = np.argsort(r["cigar_off"], kind="stable")
    assert int(r["cigar_n"].sum()) == ops.shape[0]
    starts = r["cigar_off"][order].astype(np.int64)
    assert np.array_equal(starts, np.concatenate([[0], np.cumsum(r["cigar_n"][order].astype(np.int64))[:-1]]))
    qlen = np.zeros(n, dtype=np.int64); rlen = np.zeros(n, dtype=np.int64); gaps = np.zeros(n, dtype=np.int64)
    task_of_op = order[owner]
    np.add.at(qlen, task_of_op, np.where(np.isin(kinds, (0, 1, 4)), lens, 0))
    np.add.at(rlen, task_of_op, np.where(np.isin(kinds, (0, 2)), lens, 0))
    np.add.at(gaps, task_of_op, np.where(np.isin(kinds, (1, 2)), lens, 0))
    assert np.all(qlen == 76)
    assert np.all(rlen == r["ref_end"].astype(np.int64) - r["ref_start"] + 1)
    assert np.all(r["edit_distance"] >= gaps)
    # idempotence: a second run gives the same records
    al2 = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    for f in FIELDS + ("cigar_n",):
        assert np.array_equal(al2.results[f], r[f])
    # oracle on a random sample of 5000 tasks
    pick = np.sort(np.random.default_rng(1).choice(n, size=5000, replace=False))
    q = [w.read_codes[k] for k in pick]
    wi = [w.window_codes[int(w.pair_ref_idx[k])] for k in pick]
    ores, ocig = O.align_batch(q, wi, None, nthreads=8)
    for j, k in enumerate(pick):
        assert all(int(r[int(k)][f]) == int(ores[j][f]) for f in FIELDS) and al.cigar(int(k)) == ocig[j]


def test_full_cfg1_every_record(aligner):
    """BASELINE.json configs[0] in full: 10,000 x 100 bp reads vs 10,000 distinct 1 kb windows, every record and CIGAR vs the oracle."""
    w = WL.cfg1(10_000)
    al = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    res = check_every_record_fast(al, w)
    assert int((res["flags"] & 1).sum()) == 0 and int(res["score"].min()) > 30


def test_cfg4_120k_every_record_with_tie_stress(aligner):
    """BASELINE.json configs[3]: 120,000 mixed 50-250 bp indel-rich reads (wide traceback bands) including a 20,000-read
    tie-stress subset (homopolymer windows, tandem repeats, reads with two equal-score placements); every record and CIGAR
    vs the oracle. This is the tie-break parity stress config and the slowest path of the library."""
    w = WL.cfg4(120_000, tie_stress=20_000)
    al = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    res = check_every_record_fast(al, w)
    assert int(res["cigar_n"].max()) >= 7                      # gapped CIGARs are in the set


def test_cfg4_300k_resident_batch_runs_its_traceback_in_phases(aligner):
    """A large mixed-length batch through stage / run / fetch: ubu_sw_run cuts its forward launches into groups and runs the traceback
    of each group underneath the next group's forward pass (list ranges from counter snapshots). Every record and CIGAR vs the oracle,
    twice (the second run reuses the slot's counters and scratch)."""
    w = WL.cfg4(300_000, tie_stress=6_000, seed=77)
    aligner.stage(0, w.reads, w.windows, w.pair_ref_idx)
    for _ in range(2):
        aligner.run(0); aligner.sync(0)
        assert aligner.last_run_launches(0) >= 16 + 2 * 7         # forward launches + two traceback phases of 7 kernels
        al = aligner.fetch(0, w.n_tasks)
        check_every_record_fast(al, w)


def test_phased_batch_with_contig_sized_reads(aligner):
    """A phased batch (280k mixed-length reads) that also holds contig-sized reads: sw_long_kernel runs on the phases' side streams
    behind the last group. The short records must equal those of the batch without the long reads, the long ones the oracle's."""
    rng = np.random.default_rng(5)
    w = WL.cfg4(280_000, tie_stress=2_000, seed=78)
    big = "".join(rng.choice(list("ACGT"), size=3000))
    long_reads = [_mutated(rng, big, L) for L in (300, 777, 1500, 2500, 257, 1024)]
    lw = WL.Workload("long", U.pack_texts(long_reads), U.pack_texts([big]), np.zeros(len(long_reads), dtype=np.uint32), 0)
    both = WL.concat("cfg4+long", [w, lw])
    aligner.stage(0, w.reads, w.windows, w.pair_ref_idx); aligner.run(0); aligner.sync(0)
    alone = aligner.fetch(0, w.n_tasks)
    aligner.stage(0, both.reads, both.windows, both.pair_ref_idx); aligner.run(0); aligner.sync(0)
    assert aligner.last_run_launches(0) >= 16 + 2 * 7 + 1
    got = aligner.fetch(0, both.n_tasks)
    n = w.n_tasks
    for f in FIELDS + ("cigar_n", "flags"):
        assert np.array_equal(got.results[f][:n], alone.results[f]), f
    pick = rng.choice(n, size=2000, replace=False)
    assert all(got.cigar(int(k)) == alone.cigar(int(k)) for k in pick)
    from ubu_b200.seqs import codes_from_text
    ores, ocig = O.align_batch([codes_from_text(r) for r in long_reads], [codes_from_text(big)] * len(long_reads), None, nthreads=8)
    for j in range(len(long_reads)):
        assert all(int(got.results[n + j][f]) == int(ores[j][f]) for f in FIELDS) and got.cigar(n + j) == ocig[j]


def test_extreme_scoring_reaches_the_scalar_last_resort():
    """match 31 / ext 1 lets a low score afford thousands of gap bases: the band outgrows every packed class."""
    rng = np.random.default_rng(2)
    reads = ["A" * 256, "".join(rng.choice(list("ACGT"), size=200)), "ACGT" * 50]
    wins = ["C" * 8000 + "AAA" + "C" * 50, "".join(rng.choice(list("ACGT"), size=7000)), "G" * 5000 + "ACGTACGTAC" + "G" * 100]
    r, w = U.pack_texts(reads), U.pack_texts(wins)
    params = (31, 0, 0, 1)
    with U.SwAligner(*params) as a:
        al = a.align(r, w)
    check_against_oracle(al, r, w, None, params)


def test_very_long_windows_use_the_global_selector_scratch(aligner):
    """Short reads against windows too long for the shared-memory selector array (76 bp vs 20-60 kb)."""
    rng = np.random.default_rng(8)
    wins, reads = [], []
    for W in (20000, 33000, 60000, 65535):
        w = "".join(rng.choice(list("ACGT"), size=W))
        wins.append(w)
        s = int(rng.integers(0, W - 100))
        r = list(w[s:s + 76]); r[10] = "A" if r[10] != "A" else "C"
        reads.append("".join(r))
        reads.append(w[W - 76:])                       # alignment ending in the last column
    rs, ws = U.pack_texts(reads), U.pack_texts(wins)
    idx = np.array([0, 0, 1, 1, 2, 2, 3, 3], dtype=np.uint32)
    al = aligner.align(rs, ws, idx)
    check_against_oracle(al, rs, ws, idx)
    assert int(al.results[7]["ref_end"]) == 65535 and al.cigar_string(7) == "76M"


# ---- pairs that share their window: the query-profile forward kernel (csrc/sw_fwd16s.cuh) --------------------------------
@pytest.mark.parametrize("L,W", [(76, 500), (100, 1000), (150, 2000), (250, 700), (33, 200), (50, 300), (128, 64), (208, 900)])
def test_shared_window_pairs(aligner, L, W):
    w = WL.uniform("sh", 601, L, W, seed=31 + L, paired=True, sub_rate=0.03, indel_rate=0.01, geom_p=0.4)      # odd count: last pair is half empty
    al = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    check_against_oracle(al, w.reads, w.windows, w.pair_ref_idx)


def test_shared_window_pairs_with_n(aligner):
    base = WL.uniform("shn", 500, 76, 500, seed=4, paired=True, keep_codes=True)
    w = WL.with_n(base, rate=0.02)
    al = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    check_against_oracle(al, w.reads, w.windows, w.pair_ref_idx)


@pytest.mark.parametrize("params", [(2, -1, 3, 1), (5, -4, 10, 1), (3, 0, 2, 2), (1, -9, 1, 1)])
def test_shared_window_pairs_other_scoring(params):
    w = WL.uniform("shp", 400, 90, 400, seed=22, paired=True, sub_rate=0.05, indel_rate=0.02, geom_p=0.4)
    with U.SwAligner(*params) as a:
        al = a.align(w.reads, w.windows, w.pair_ref_idx)
    check_against_oracle(al, w.reads, w.windows, w.pair_ref_idx, params)


def test_shared_window_ragged_reads_and_many_reads_per_window(aligner):
    """Reads of different lengths (one length class) in a pair, windows of different lengths, and reads of one window
    scattered over the batch (the planner groups them by window)."""
    rng = np.random.default_rng(12)
    wins = ["".join(rng.choice(list("ACGT"), size=int(rng.integers(90, 700)))) for _ in range(40)]
    reads, idx = [], []
    for k in range(500):
        wi = int(rng.integers(0, len(wins)))
        w = wins[wi]
        L = int(rng.integers(65, 77))                   # one length class (<= 76 rows)
        s = int(rng.integers(0, max(len(w) - L, 1)))
        r = list(w[s:s + L]) + list(rng.choice(list("ACGT"), size=max(0, L - len(w[s:s + L]))))
        for p in rng.integers(0, L, size=2):
            r[p] = "ACGT"[int(rng.integers(0, 4))]
        reads.append("".join(r)); idx.append(wi)
    rs, ws = U.pack_texts(reads), U.pack_texts(wins)
    idx = np.array(idx, dtype=np.uint32)
    al = aligner.align(rs, ws, idx)
    check_against_oracle(al, rs, ws, idx)
    # same reads, every window the same length: the uniform-batch plan with its counting sort by window
    wins2 = ["".join(rng.choice(list("ACGT"), size=400)) for _ in range(40)]
    reads2 = []
    for k in range(500):
        w = wins2[int(idx[k])]
        s = int(rng.integers(0, 400 - 76))
        r = list(w[s:s + 76]); r[int(rng.integers(0, 76))] = "A"
        reads2.append("".join(r))
    rs2, ws2 = U.pack_texts(reads2), U.pack_texts(wins2)
    al2 = aligner.align(rs2, ws2, idx)
    check_against_oracle(al2, rs2, ws2, idx)


def test_shared_window_kernel_equals_general_kernel(aligner, monkeypatch):
    """The two forward kernels must give identical records on the same batch (UBU_NO_SHARED_WIN=1 forces the general one)."""
    w = WL.uniform("eq", 20_000, 76, 500, seed=77, paired=True, sub_rate=0.04, indel_rate=0.01)
    a1 = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    monkeypatch.setenv("UBU_NO_SHARED_WIN", "1")
    a2 = aligner.align(w.reads, w.windows, w.pair_ref_idx)
    monkeypatch.delenv("UBU_NO_SHARED_WIN")
    for f in FIELDS + ("cigar_n", "flags"):
        assert np.array_equal(a1.results[f], a2.results[f]), f
    assert all(a1.cigar(k) == a2.cigar(k) for k in range(0, len(a1), 37))


@pytest.mark.parametrize("seed", range(12))
def test_randomized_differential(seed):
    """Random scoring parameters, lengths, repeat structure, N density and window sharing per seed; every record against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    match = int(rng.integers(1, 6)); mismatch = -int(rng.integers(0, 9)); gap_open = int(rng.integers(0, 12)); gap_ext = int(rng.integers(1, 5))
    params = (match, mismatch, gap_open, gap_ext)
    n_win = int(rng.integers(5, 60))
    alphabet = list("ACGT") if seed % 3 else list("ACGTN")
    pw = None if seed % 3 else [0.245, 0.245, 0.245, 0.245, 0.02]
    low_complexity = seed % 4 == 1
    wins = []
    for _ in range(n_win):
        W = int(rng.integers(1, 700))
        if low_complexity:
            u = "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 5))))
            wins.append((u * W)[:W])
        else:
            wins.append("".join(rng.choice(alphabet, size=W, p=pw)))
    reads, idx = [], []
    for _ in range(int(rng.integers(150, 400))):
        wi = int(rng.integers(0, n_win)); w = wins[wi]
        L = int(rng.integers(1, 257))
        s = int(rng.integers(0, len(w)))
        r = list(w[s:s + L])
        r += list(rng.choice(list("ACGT"), size=L - len(r)))
        k = 0
        while k < len(r):                                     # substitutions, insertions, deletions
            u = rng.random()
            if u < 0.03: r[k] = "ACGT"[int(rng.integers(0, 4))]
            elif u < 0.04: r[k:k] = list(rng.choice(list("ACGT"), size=int(rng.integers(1, 6))))
            elif u < 0.05: del r[k:k + int(rng.integers(1, 6))]
            k += 1
        r = r[:256] or ["A"]
        reads.append("".join(r)); idx.append(wi)
    order = np.argsort(np.array(idx), kind="stable") if seed % 2 else np.arange(len(reads))      # grouped by window or scattered
    reads = [reads[i] for i in order]; idx = np.array([idx[i] for i in order], dtype=np.uint32)
    rs, ws = U.pack_texts(reads), U.pack_texts(wins)
    with U.SwAligner(*params) as a:
        al = a.align(rs, ws, idx)
    check_against_oracle(al, rs, ws, idx, params)


def test_full_size_cfg3_properties(aligner):
    """BASELINE.json configs[2] at full size: 5M x 150 bp reads against 2 kb contig windows (1M windows, ~5 reads per window,
    scattered over the batch -> the planner groups them by window), built in five shards to bound host memory. Properties that
    need no oracle over all 5M records, plus an oracle-checked sample."""
    shards, per, L, W, pool = 5, 1_000_000, 150, 2000, 200_000
    parts, codes_r, codes_w = [], [], []
    for k in range(shards):
        w = WL.uniform("cfg3", per, L, W, n_windows=pool, seed=WL.SEEDS["cfg3"] + k, keep_codes=True)
        parts.append(w); codes_r.append(w.read_codes[:400]); codes_w.append((w.window_codes, w.pair_ref_idx[:400]))
    nb_r, nb_w = (L + 3) // 4, (W + 3) // 4
    reads = U.SeqSet(np.concatenate([p.reads.packed[: per * nb_r] for p in parts]), (np.arange(shards * per, dtype=np.int64) * nb_r).astype(np.uint32),
                     np.full(shards * per, L, dtype=np.uint16), None, None)
    wins = U.SeqSet(np.concatenate([p.windows.packed[: pool * nb_w] for p in parts]), (np.arange(shards * pool, dtype=np.int64) * nb_w).astype(np.uint32),
                    np.full(shards * pool, W, dtype=np.uint16), None, None)
    idx = np.concatenate([p.pair_ref_idx.astype(np.uint32) + np.uint32(k * pool) for k, p in enumerate(parts)])
    n = shards * per
    del parts
    al = aligner.align(reads, wins, idx)
    r = al.results
    assert len(al) == n and int((r["flags"] & 1).sum()) == 0
    assert int(r["score"].max()) <= L and int(r["score"].min()) > 40
    assert np.all(r["q_end"] >= r["q_start"]) and np.all(r["ref_end"] >= r["ref_start"]) and np.all(r["ref_end"] <= W) and np.all(r["ref_start"] >= 1)
    ops = al.cigar_pool
    assert int(r["cigar_n"].sum()) == ops.shape[0]
    order = np.argsort(r["cigar_off"], kind="stable")
    starts = r["cigar_off"][order].astype(np.int64)
    assert np.array_equal(starts, np.concatenate([[0], np.cumsum(r["cigar_n"][order].astype(np.int64))[:-1]]))      # slices tile the pool
    task_of_op = np.repeat(order, r["cigar_n"][order].astype(np.int64))
    lens, kinds = (ops >> 4).astype(np.int64), (ops & 15)
    qlen = np.bincount(task_of_op, weights=np.where(np.isin(kinds, (0, 1, 4)), lens, 0), minlength=n)
    rlen = np.bincount(task_of_op, weights=np.where(np.isin(kinds, (0, 2)), lens, 0), minlength=n)
    gaps = np.bincount(task_of_op, weights=np.where(np.isin(kinds, (1, 2)), lens, 0), minlength=n)
    assert np.all(qlen == L)
    assert np.all(rlen == r["ref_end"].astype(np.int64) - r["ref_start"] + 1)
    assert np.all(r["edit_distance"] >= gaps)
    # score is consistent with the record: matches - mismatches - gap costs, given NM = mismatches + gap bases
    aligned = r["q_end"].astype(np.int64) - r["q_start"] + 1
    ins = np.bincount(task_of_op, weights=np.where(kinds == 1, lens, 0), minlength=n)
    n_gap_runs = np.bincount(task_of_op, weights=np.isin(kinds, (1, 2)).astype(np.float64), minlength=n)
    mism = r["edit_distance"] - gaps
    m_cols = aligned - ins
    assert np.array_equal(r["score"].astype(np.int64), (m_cols - mism) * 1 - 3 * mism - 5 * n_gap_runs - 2 * gaps)
    # oracle on the first 400 reads of every shard
    for k in range(shards):
        wc, ix = codes_w[k]
        q = list(codes_r[k]); wi = [wc[int(j)] for j in ix]
        ores, ocig = O.align_batch(q, wi, None, nthreads=8)
        for j in range(400):
            t = k * per + j
            assert all(int(r[t][f]) == int(ores[j][f]) for f in FIELDS) and al.cigar(t) == ocig[j]


def test_multi_gpu_aligner_pipeline_equals_one_call(aligner):
    """ubu_b200.sharding.MultiGpuAligner on the visible GPUs (chunks in flight through pinned staging, in-place ordered merge)
    returns what one align_batch call returns, CIGARs included; windows shared across chunk borders, ragged reads, N."""
    from ubu_b200 import _native as N
    from ubu_b200 import sharding as SH

    rng = np.random.default_rng(4)
    base = WL.uniform("mg", 30_011, 100, 700, n_windows=4_000, seed=91, sub_rate=0.03, indel_rate=0.01, keep_codes=True)
    order = np.argsort(base.pair_ref_idx, kind="stable")
    codes = base.read_codes[order]
    codes[rng.random(codes.shape) < 0.002] = 4
    reads = U.pack_codes([c[: int(rng.integers(60, 101))] for c in codes])
    idx = base.pair_ref_idx[order]
    one = aligner.align(reads, base.windows, idx)
    ndev = N.lib().ubu_sw_device_count()
    with SH.MultiGpuAligner(list(range(ndev)), chunk=2_500) as m:
        for _ in range(2):                                  # second call reuses the staging buffers
            got = m.align(reads, base.windows, idx)
            for f in FIELDS + ("cigar_n", "flags"):
                assert np.array_equal(got.results[f], one.results[f]), f
            assert all(got.cigar(k) == one.cigar(k) for k in range(0, len(one), 13))


def test_align_batch_streams_large_batches_identically(monkeypatch):
    """ubu_sw_align_batch cuts a large batch into chunks that travel through the handle's slots; the records, their order and
    the CIGARs must equal the one-shot form (UBU_NO_STREAM=1): mates sharing windows, one window per read, N, and a CIGAR
    pool that is too small on the first try."""
    cases = []
    w = WL.uniform("st1", 300_001, 76, 500, seed=61, paired=True, sub_rate=0.02, indel_rate=0.004)
    cases.append((w.reads, w.windows, w.pair_ref_idx))
    w = WL.uniform("st2", 150_000, 100, 300, seed=62)                       # idx None: window k belongs to read k
    cases.append((w.reads, w.windows, None))
    b = WL.uniform("st3", 140_000, 60, 200, seed=63, paired=True, keep_codes=True)
    wn = WL.with_n(b, rate=0.003)
    cases.append((wn.reads, wn.windows, wn.pair_ref_idx))
    with U.SwAligner(slots=4) as a:
        for reads, wins, idx in cases:
            got = a.align(reads, wins, idx)
            with pytest.raises(U.UbuSwError) as e:                      # a caller-given pool that is too small: status + required size
                a.align(reads, wins, idx, cigar_cap=reads.count // 2)
            assert e.value.status == -3
            again = a.align(reads, wins, idx)                          # the handle is fine afterwards
            monkeypatch.setenv("UBU_NO_STREAM", "1")
            want = a.align(reads, wins, idx)
            monkeypatch.delenv("UBU_NO_STREAM")
            for other in (got, again):
                for f in FIELDS + ("cigar_n", "flags"):
                    assert np.array_equal(other.results[f], want.results[f]), f
                assert other.cigar_pool.shape[0] == want.cigar_pool.shape[0]
                order = np.argsort(other.results["cigar_off"], kind="stable")              # slices tile the pool
                order = order[other.results["cigar_n"][order] > 0]
                starts = other.results["cigar_off"][order].astype(np.int64)
                assert np.array_equal(starts, np.concatenate([[0], np.cumsum(other.results["cigar_n"][order].astype(np.int64))[:-1]]))
                for k in range(0, len(want), 997):
                    assert other.cigar(k) == want.cigar(k)


def test_mixed_batch_pairs_window_sharing_neighbours(aligner, monkeypatch):
    """The general plan (several read-length classes, windows of different lengths, some with N): neighbours that share a
    window go to the query-profile kernel, the rest to the general one; every record against the oracle and against the
    general-kernel-only run."""
    rng = np.random.default_rng(33)
    wins = ["".join(rng.choice(list("ACGTN"), size=int(rng.integers(120, 900)), p=[.2475, .2475, .2475, .2475, .01])) for _ in range(300)]
    reads, idx = [], []
    for wi, w in enumerate(wins):                          # coordinate-sorted style: the reads of a window follow each other
        for _ in range(int(rng.integers(0, 7))):
            L = int(rng.choice([50, 76, 100, 101, 150, 250]))
            s = int(rng.integers(0, max(len(w) - L, 1)))
            r = list(w[s:s + L]) + list(rng.choice(list("ACGT"), size=max(0, L - len(w[s:s + L]))))
            for p in rng.integers(0, L, size=max(1, L // 40)):
                r[p] = "ACGT"[int(rng.integers(0, 4))]
            reads.append("".join(r)); idx.append(wi)
    rs, ws = U.pack_texts(reads), U.pack_texts(wins)
    idx = np.array(idx, dtype=np.uint32)
    got = aligner.align(rs, ws, idx)
    check_against_oracle(got, rs, ws, idx)
    monkeypatch.setenv("UBU_NO_SHARED_WIN", "1")
    want = aligner.align(rs, ws, idx)
    monkeypatch.delenv("UBU_NO_SHARED_WIN")
    for f in FIELDS + ("cigar_n", "flags"):
        assert np.array_equal(got.results[f], want.results[f]), f


def test_multi_gpu_driver_equals_one_call(aligner):
    """ubu_sw_multi_align (C++ driver inside the library: a host thread + handle per GPU, chunk c -> GPU c mod N, per-chunk slices of
    the caller's CIGAR pool, no host merge pass) on all visible GPUs returns what one ubu_sw_align_batch call returns: records in
    input order, CIGARs equal; windows shared across chunk borders, ragged reads, N; then a scattered (non-compact) layout; then
    SAM text equal to ubu_sw_sam_write's."""
    import os
    import tempfile

    from ubu_b200 import multi as M

    rng = np.random.default_rng(4)
    base = WL.uniform("mg", 30_011, 100, 700, n_windows=4_000, seed=91, sub_rate=0.03, indel_rate=0.01, keep_codes=True)
    order = np.argsort(base.pair_ref_idx, kind="stable")
    codes = base.read_codes[order]
    codes[rng.random(codes.shape) < 0.002] = 4
    reads = U.pack_codes([c[: int(rng.integers(60, 101))] for c in codes])
    idx = base.pair_ref_idx[order]
    one = aligner.align(reads, base.windows, idx)
    with M.MultiAligner(chunk=2_500, slots=3) as m:
        assert m.n_devices >= 1
        for _ in range(2):
            got = m.align(reads, base.windows, idx)
            for f in FIELDS + ("cigar_n", "flags"):
                assert np.array_equal(got.results[f], one.results[f]), f
            assert all(got.cigar(k) == one.cigar(k) for k in range(0, len(one), 7))
        # scattered reads (window indices all over the set): chunks are not compact, every chunk uploads the whole window set
        perm = rng.permutation(reads.count)
        reads_s = U.pack_codes([reads.codes(int(k)) for k in perm])
        got = m.align(reads_s, base.windows, idx[perm])
        for f in FIELDS + ("cigar_n", "flags"):
            assert np.array_equal(got.results[f], one.results[f][perm]), f
        assert all(got.cigar(k) == one.cigar(int(perm[k])) for k in range(0, len(one), 11))
        # SAM text behind the merge == the single-call SAM writer on the single-call results
        with tempfile.TemporaryDirectory() as td:
            p1, p2 = os.path.join(td, "multi.sam"), os.path.join(td, "one.sam")
            fd = os.open(p1, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
            try:
                got = m.align(reads, base.windows, idx, sam_fd=fd, sam_threads=3)
            finally:
                os.close(fd)
            fd = os.open(p2, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
            try:
                U.write_sam(fd, reads, one, idx)
            finally:
                os.close(fd)
            assert open(p1, "rb").read() == open(p2, "rb").read() and m.last_sam_bytes == os.path.getsize(p2)
    # a pool that is too small: status + a capacity that works
    with M.MultiAligner(chunk=2_500) as m:
        rs, ws = reads.as_struct(), base.windows.as_struct()
        import ctypes as C
        from ubu_b200 import _native as N
        res = np.zeros(reads.count, dtype=U.RESULT_DTYPE); pool = np.zeros(reads.count // 2, dtype=np.uint32); used = C.c_size_t(0)
        ix = np.ascontiguousarray(idx, dtype=np.uint32)
        rc = N.lib().ubu_sw_multi_align(m._m, C.byref(rs), C.byref(ws), ix.ctypes.data, res.ctypes.data, pool.ctypes.data, pool.shape[0], C.byref(used))
        assert rc == N.ERR_CIGAR_POOL and used.value > pool.shape[0]
        pool = np.zeros(used.value, dtype=np.uint32)
        rc = N.lib().ubu_sw_multi_align(m._m, C.byref(rs), C.byref(ws), ix.ctypes.data, res.ctypes.data, pool.ctypes.data, pool.shape[0], C.byref(used))
        assert rc == 0
        for f in FIELDS + ("cigar_n", "flags"):
            assert np.array_equal(res[f], one.results[f]), f


# ---- reads longer than 256 bp: the contig-sized queries of Aligner.align (sw_long.cuh, 32-bit, a warp per task) ----------------
def _mutated(rng, w, L):
    s = int(rng.integers(0, max(len(w) - L, 1)))
    r = list(w[s:s + L]) + list(rng.choice(list("ACGT"), size=max(0, L - len(w[s:s + L]))))
    k = 0
    while k < len(r):
        u = rng.random()
        if u < 0.02: r[k] = "ACGT"[int(rng.integers(0, 4))]
        elif u < 0.025: r[k:k] = list(rng.choice(list("ACGT"), size=int(rng.integers(1, 30))))
        elif u < 0.03: del r[k:k + int(rng.integers(1, 30))]
        k += 1
    return "".join(r[:L]) or "A"


@pytest.mark.parametrize("seed", range(4))
def test_long_reads_randomized_differential(seed):
    """Contig-sized queries (257 .. 2500 bp) mixed with short reads in one batch, random scoring, N, shared windows:
    every record and CIGAR against the oracle."""
    rng = np.random.default_rng(3000 + seed)
    params = [(1, -3, 5, 2), (2, -1, 3, 1), (5, -4, 10, 1), (1, -1, 0, 1)][seed]
    alphabet = list("ACGT") if seed % 2 else list("ACGTN")
    pw = None if seed % 2 else [0.2475, 0.2475, 0.2475, 0.2475, 0.01]
    wins = ["".join(rng.choice(alphabet, size=int(rng.integers(300, 4000)), p=pw)) for _ in range(12)]
    reads, idx = [], []
    for k in range(90):
        wi = int(rng.integers(0, len(wins)))
        L = int(rng.choice([int(rng.integers(257, 2501)), int(rng.integers(1, 257)), 257, 2500, 512, 1024]))
        reads.append(_mutated(rng, wins[wi], L)); idx.append(wi)
    rs, ws = U.pack_texts(reads), U.pack_texts(wins)
    idx = np.array(idx, dtype=np.uint32)
    with U.SwAligner(*params) as a:
        al = a.align(rs, ws, idx)
    check_against_oracle(al, rs, ws, idx, params)
    assert max(len(r) for r in reads) > 2000


def test_long_reads_edge_cases(aligner):
    """Longest read the library takes (4096), a long read against a short window, a long homopolymer (every tie), a long read
    that does not align at all, and a batch made only of long reads."""
    rng = np.random.default_rng(77)
    w = "".join(rng.choice(list("ACGT"), size=6000))
    reads = [w[100:4196], "".join(rng.choice(list("ACGT"), size=700)), "A" * 600, "C" * 300, w[2000:2300] + "TTTT" + w[2300:2700]]
    wins = [w, w[3000:3100], "A" * 900, "G" * 500, w]
    rs, ws = U.pack_texts(reads), U.pack_texts(wins)
    al = aligner.align(rs, ws)
    check_against_oracle(al, rs, ws, None)
    assert al.cigar_string(0) == "4096M" and int(al.results[3]["flags"]) & 1 and "4I" in al.cigar_string(4)


def test_extreme_scoring_with_long_reads():
    """match 31: scores far beyond 16 bits' comfort (31 x 2000 = 62000) only exist on the 32-bit long-read path."""
    rng = np.random.default_rng(6)
    w = "".join(rng.choice(list("ACGT"), size=3000))
    reads = [w[500:2500], _mutated(rng, w, 1500), "A" * 300]
    r, ws = U.pack_texts(reads), U.pack_texts([w, w, "C" * 100 + "A" * 40 + "C" * 100])
    params = (31, -64, 63, 31)
    with U.SwAligner(*params) as a:
        al = a.align(r, ws)
    check_against_oracle(al, r, ws, None, params)
    assert int(al.results[0]["score"]) == 31 * 2000


def test_cfg5_shape_through_the_multi_gpu_driver():
    """BASELINE.json configs[4] in shape (150 bp reads vs 2 kb windows, reads grouped by window, generated shard by shard from
    (seed, shard)) at 2 % of its size, through ubu_sw_multi_align on every visible GPU: size-independent properties over all
    2 M records, an oracle-checked sample, SAM line count, and equality with one ubu_sw_align_batch call on one GPU."""
    import os

    from ubu_b200 import multi as M

    n = 2_000_000
    w = WL.cfg3_sharded(n, shard=250_000, seed=WL.SEEDS["cfg5"], threads=8)
    with M.MultiAligner(slots=4) as m:
        fd = os.open(os.devnull, os.O_WRONLY)
        try:
            al = m.align(w.reads, w.windows, w.pair_ref_idx, sam_fd=fd)
        finally:
            os.close(fd)
        sam_bytes = m.last_sam_bytes
    r = al.results
    assert len(al) == n and int((r["flags"] & 1).sum()) == 0 and int(r["score"].min()) > 40 and int(r["score"].max()) <= 150
    assert np.all(r["ref_end"] <= 2000) and np.all(r["ref_start"] >= 1) and np.all(r["q_end"] >= r["q_start"])
    # every CIGAR slice lies inside the pool, slices do not overlap, and every CIGAR covers its read
    order = np.argsort(r["cigar_off"], kind="stable")
    starts, cn = r["cigar_off"][order].astype(np.int64), r["cigar_n"][order].astype(np.int64)
    assert np.all(starts[1:] >= starts[:-1] + cn[:-1]) and starts[-1] + cn[-1] <= al.cigar_pool.shape[0]
    task = np.repeat(np.arange(n), r["cigar_n"].astype(np.int64))
    k = np.arange(task.shape[0]) - np.repeat(np.concatenate([[0], np.cumsum(r["cigar_n"].astype(np.int64))[:-1]]), r["cigar_n"].astype(np.int64))
    ops = al.cigar_pool[r["cigar_off"].astype(np.int64)[task] + k]
    lens, kinds = (ops >> 4).astype(np.int64), ops & 15
    assert np.all(np.bincount(task, weights=np.where(np.isin(kinds, (0, 1, 4)), lens, 0), minlength=n) == 150)
    assert np.all(np.bincount(task, weights=np.where(np.isin(kinds, (0, 2)), lens, 0), minlength=n) == r["ref_end"].astype(np.int64) - r["ref_start"] + 1)
    assert sam_bytes > n * 200                                     # one SAM line per read was formatted
    # oracle on a sample
    pick = np.sort(np.random.default_rng(2).choice(n, size=3000, replace=False))
    q = [w.reads.codes(int(t)) for t in pick]
    wi = [w.windows.codes(int(w.pair_ref_idx[t])) for t in pick]
    ores, ocig = O.align_batch(q, wi, None, nthreads=8)
    for j, t in enumerate(pick):
        assert all(int(r[int(t)][f]) == int(ores[j][f]) for f in FIELDS) and al.cigar(int(t)) == ocig[j]
    # one call on one GPU gives the same records
    with U.SwAligner(slots=4) as a:
        one = a.align(w.reads, w.windows, w.pair_ref_idx)
    for f in FIELDS + ("cigar_n", "flags"):
        assert np.array_equal(one.results[f], r[f]), f


def test_streamed_and_multi_gpu_batches_with_contig_sized_reads(monkeypatch):
    """A large batch (streamed through the slots inside ubu_sw_align_batch, and sharded by ubu_sw_multi_align) that also holds contig-sized
    reads: the chunks plan their long reads onto sw_long_kernel; records must equal the one-shot call's, and the long ones the oracle's."""
    from ubu_b200 import multi as M

    rng = np.random.default_rng(41)
    n_short, L, W = 150_000, 100, 600
    wins = rng.integers(0, 4, size=(n_short // 3, W), dtype=np.uint8)
    idx = np.sort(rng.integers(0, wins.shape[0], size=n_short)).astype(np.uint32)
    start = rng.integers(0, W - L, size=n_short)
    reads = [wins[idx[k], start[k]:start[k] + L].copy() for k in range(n_short)]
    long_at = list(range(500, n_short, 3000))                                 # 50 long reads spread over the batch
    big = rng.integers(0, 4, size=2500, dtype=np.uint8)
    for t in long_at:
        a = int(rng.integers(0, 400)); b = int(rng.integers(300, 1500))
        r = np.concatenate([big[a:a + b // 2], big[a + b // 2 + 7:a + b]])        # a 7 bp deletion
        reads[t] = r
    wl = [w for w in wins]
    win_of = idx.copy()
    for t in long_at:                                                           # the long reads get the big window, appended to the set
        win_of[t] = len(wl)
    wl.append(big)
    # keep window indices ascending enough for streaming: the big window sits at the end, so chunks holding a long read fall back
    rs, ws = U.pack_codes(reads), U.pack_codes(wl)
    with U.SwAligner(slots=4) as a:
        got = a.align(rs, ws, win_of)
        monkeypatch.setenv("UBU_NO_STREAM", "1")
        want = a.align(rs, ws, win_of)
        monkeypatch.delenv("UBU_NO_STREAM")
    for f in FIELDS + ("cigar_n", "flags"):
        assert np.array_equal(got.results[f], want.results[f]), f
    ores, ocig = O.align_batch([reads[t] for t in long_at], [big] * len(long_at), None, nthreads=8)
    for j, t in enumerate(long_at):
        assert all(int(got.results[t][f]) == int(ores[j][f]) for f in FIELDS) and got.cigar(t) == ocig[j] == want.cigar(t)
        assert any(op == 2 for _, op in got.cigar(t))
    with M.MultiAligner(chunk=20_000) as m:
        mg = m.align(rs, ws, win_of)
    for f in FIELDS + ("cigar_n", "flags"):
        assert np.array_equal(mg.results[f], want.results[f]), f
    assert all(mg.cigar(t) == want.cigar(t) for t in long_at)
