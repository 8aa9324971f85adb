"""The host plan of a batch (length classes, pairing, window sharing, launch split) through ubu_sw_debug_plan: runs without a GPU.
What the forward kernels rely on is asserted as invariants over random and hand-made batches."""
import ctypes as C

import numpy as np
import pytest

import ubu_b200 as U
from ubu_b200 import _native as N
from ubu_b200 import workloads as WL

PAD = 0xFFFFFFFF


def plan(reads, wins, idx):
    n = reads.count
    rs, ws = reads.as_struct(), wins.as_struct()
    ix = None if idx is None else np.ascontiguousarray(idx, dtype=np.uint32)
    order = np.zeros(n + 128, dtype=np.uint32)
    launches = np.zeros((512, 6), dtype=np.int32)
    n_slots, ident, nl = C.c_uint32(0), C.c_int(0), C.c_int(0)
    rc = N.lib().ubu_sw_debug_plan(C.byref(rs), C.byref(ws), None if ix is None else ix.ctypes.data, order.ctypes.data, order.shape[0],
                                   C.byref(n_slots), C.byref(ident), launches.ctypes.data, launches.shape[0], C.byref(nl))
    assert rc == 0, rc
    return order[: n_slots.value], bool(ident.value), launches[: nl.value]


def check_invariants(reads, wins, idx, order, launches):
    n = reads.count
    real = order[order != PAD]
    assert np.array_equal(np.sort(real), np.arange(n, dtype=np.uint32))            # every task exactly once
    n_long = sum(int(l[3]) for l in launches if l[0] == -1)
    assert (order.shape[0] - n_long) % 2 == 0 and order.shape[0] - n <= 2 * max(len(launches), 1)
    covered = np.zeros(order.shape[0], dtype=np.int32)
    win_of = (lambda t: int(idx[t])) if idx is not None else (lambda t: int(t))
    for rows, has_n, first, slots, wmax, shared in launches:
        covered[first:first + slots] += 1
        part = order[first:first + slots]
        if rows == -1:                                                             # the long-read kernel's list: reads beyond 256 bp, no padding
            assert first + slots == order.shape[0] and np.all(part != PAD) and all(int(reads.len[t]) > 256 for t in part)
            continue
        assert slots % 2 == 0 and slots > 0
        for t in part[part != PAD]:
            assert int(reads.len[t]) <= rows
            assert int(wins.len[win_of(t)]) <= wmax
        if shared:
            a, b = part[0::2], part[1::2]
            both = (a != PAD) & (b != PAD)
            assert all(win_of(x) == win_of(y) for x, y in zip(a[both], b[both]))
        if wins.nmask is not None and not has_n:                                   # the no-N kernels never see a window with N
            for t in part[part != PAD]:
                assert "N" not in wins.text(win_of(t))
    assert np.all(covered == 1)                                                    # launches tile the slots


def test_paired_uniform_batch_runs_in_input_order_on_the_shared_kernel():
    w = WL.uniform("p", 20_001, 76, 500, seed=3, paired=True)
    order, ident, launches = plan(w.reads, w.windows, w.pair_ref_idx)
    assert ident and len(launches) == 1 and launches[0][5] == 1 and launches[0][0] == 76 and launches[0][4] == 500
    check_invariants(w.reads, w.windows, w.pair_ref_idx, order, launches)


def test_one_window_per_read_is_not_shared():
    w = WL.uniform("s", 5_000, 100, 1000, seed=4)
    order, ident, launches = plan(w.reads, w.windows, None)
    assert ident and len(launches) == 1 and launches[0][5] == 0 and launches[0][0] == 104
    check_invariants(w.reads, w.windows, None, order, launches)


def test_scattered_reads_are_grouped_by_window():
    w = WL.uniform("g", 30_000, 150, 2000, n_windows=6_000, seed=5)
    order, ident, launches = plan(w.reads, w.windows, w.pair_ref_idx)
    assert not ident and len(launches) == 2 and launches[0][5] == 1 and launches[1][5] == 0
    assert launches[0][3] > 0.8 * 30_000                                            # ~5 reads per window: most tasks end up in sharing pairs
    check_invariants(w.reads, w.windows, w.pair_ref_idx, order, launches)


@pytest.mark.parametrize("seed", range(6))
def test_mixed_batches_keep_the_plan_invariants(seed):
    rng = np.random.default_rng(70 + seed)
    alphabet, p = list("ACGTN"), [.2475, .2475, .2475, .2475, .01]
    wins = ["".join(rng.choice(alphabet, size=int(rng.integers(1, 1500)), p=p if seed % 2 else [.25, .25, .25, .25, 0])) for _ in range(200)]
    reads, idx = [], []
    for wi in range(len(wins)):
        for _ in range(int(rng.integers(0, 6))):
            reads.append("".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 257))))); idx.append(wi)
    if seed % 3 == 0:                                                               # scattered instead of window by window
        perm = rng.permutation(len(reads)); reads = [reads[i] for i in perm]; idx = [idx[i] for i in perm]
    rs, ws = U.pack_texts(reads), U.pack_texts(wins)
    idx = np.array(idx, dtype=np.uint32)
    order, ident, launches = plan(rs, ws, idx)
    check_invariants(rs, ws, idx, order, launches)
    if seed % 3:
        assert any(l[5] for l in launches)                                          # neighbours sharing a window were found


def test_plan_rejects_bad_batches_and_handles_empty_ones():
    e = U.pack_texts([])
    order, ident, launches = plan(e, e, None)
    assert order.shape[0] == 0 and len(launches) == 0
    r, w = U.pack_texts(["ACGT", "AC"]), U.pack_texts(["ACGTACGT"])
    rs, ws = r.as_struct(), w.as_struct()
    n_slots, nl = C.c_uint32(0), C.c_int(0)
    bad = np.array([0, 1], dtype=np.uint32)                                         # window index out of range
    assert N.lib().ubu_sw_debug_plan(C.byref(rs), C.byref(ws), bad.ctypes.data, None, 0, C.byref(n_slots), None, None, 0, C.byref(nl)) == N.ERR_ARG
    assert N.lib().ubu_sw_debug_plan(C.byref(rs), C.byref(ws), None, None, 0, C.byref(n_slots), None, None, 0, C.byref(nl)) == N.ERR_ARG   # counts differ
    w1s = U.pack_texts(["ACGTACGT"]); w1 = w1s.as_struct()
    for text, want in (("A" * 300, 0), ("A" * 4096, 0), ("A" * 4097, N.ERR_TOO_LONG)):      # reads beyond 256 bp go to the long-read kernel, up to 4096
        lr = U.pack_texts([text]); ls = lr.as_struct()
        assert N.lib().ubu_sw_debug_plan(C.byref(ls), C.byref(w1), None, None, 0, C.byref(n_slots), None, None, 0, C.byref(nl)) == want
    bws = U.pack_texts(["ACGT" * 16000]); big_w = bws.as_struct()                 # 4096 x 64000 cells > 2^26: too large a matrix
    lr = U.pack_texts(["A" * 4096]); ls = lr.as_struct()
    assert N.lib().ubu_sw_debug_plan(C.byref(ls), C.byref(big_w), None, None, 0, C.byref(n_slots), None, None, 0, C.byref(nl)) == N.ERR_TOO_LONG


def test_long_reads_get_their_own_list_behind_the_slots():
    rng = np.random.default_rng(5)
    lens = [100, 300, 76, 2500, 250, 257, 150, 1024]
    reads = U.pack_texts(["".join(rng.choice(list("ACGT"), size=L)) for L in lens])
    wins = U.pack_texts(["".join(rng.choice(list("ACGT"), size=3000)) for _ in lens])
    order, ident, launches = plan(reads, wins, None)
    assert not ident and launches[-1][0] == -1 and launches[-1][3] == 4
    assert sorted(int(t) for t in order[-4:]) == [1, 3, 5, 7]
    check_invariants(reads, wins, None, order, launches)
    only_long = U.pack_texts(["".join(rng.choice(list("ACGT"), size=400)) for _ in range(3)])
    order, ident, launches = plan(only_long, U.pack_texts(["ACGT" * 200] * 3), None)
    assert len(launches) == 1 and launches[0][0] == -1 and order.shape[0] == 3


def test_mixed_read_lengths_in_one_class_with_exactly_sized_buffers():
    """Regression: the uniform fast path (one length class, one window length) bound-checked every read with the bytes of the
    LONGEST read; with exactly sized packed buffers and the shortest read last it rejected a valid batch."""
    reads = U.pack_texts(["A" * 104, "C" * 104, "G" * 101, "T" * 100])
    wins = U.pack_texts(["ACGT" * 75] * 4)
    assert reads.packed.size == 26 + 26 + 26 + 25                                   # exact: no slack behind the last read
    order, ident, launches = plan(reads, wins, None)
    assert ident and len(launches) == 1 and launches[0][0] == 104
    check_invariants(reads, wins, None, order, launches)
    # the same with an N plane on the reads
    reads_n = U.pack_texts(["A" * 104, "C" * 104, "G" * 101, "TTNT" * 25])
    order, ident, launches = plan(reads_n, wins, None)
    check_invariants(reads_n, wins, None, order, launches)
    # and a genuinely short buffer is still rejected
    rs, ws = reads.as_struct(), wins.as_struct()
    rs.packed_bytes -= 1
    n_slots, nl = C.c_uint32(0), C.c_int(0)
    assert N.lib().ubu_sw_debug_plan(C.byref(rs), C.byref(ws), None, None, 0, C.byref(n_slots), None, None, 0, C.byref(nl)) == N.ERR_ARG


def test_buckets_are_ordered_by_window_length_stably():
    """Inside a read-length class the tasks are ordered by ascending window length (a pair, a CTA sweep similar widths), ties in input
    order: the plan sorts with a counting sort over the bucket's window-length range, narrow or as wide as 16-bit lengths allow."""
    rng = np.random.default_rng(12)
    for wlo, whi in ((200, 1000), (1, 65535)):
        n = 3000
        rl = rng.integers(20, 257, size=n)
        wl = rng.integers(wlo, whi + 1, size=n)
        wl[rng.integers(0, n, size=n // 3)] = wlo + 7                              # many ties
        reads = U.pack_codes([rng.integers(0, 4, size=int(l), dtype=np.uint8) for l in rl])
        wins = U.pack_codes([rng.integers(0, 4, size=int(l), dtype=np.uint8) for l in wl])
        order, ident, launches = plan(reads, wins, None)
        assert not ident
        check_invariants(reads, wins, None, order, launches)
        by_rows = {}
        for rows, has_n, first, slots, wmax, shared in launches:                   # a class may be split into several launches: join them
            by_rows.setdefault(int(rows), []).extend(int(t) for t in order[first:first + slots] if t != PAD)
        assert len(by_rows) > 10
        for rows, tasks in by_rows.items():
            keys = [(int(wl[t]), t) for t in tasks]
            assert keys == sorted(keys), rows                                      # ascending window length, input order among equals
