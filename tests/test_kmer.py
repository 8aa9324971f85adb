"""k-mer graph build: the oracle against the reference's assembler fixture, the boundary
(include/ubu_kmer.h), and -- on the GPU -- the device table against the oracle node for node."""
import ctypes as C
import json
import os
import re
import sys

import numpy as np
import pytest

import ubu_b200 as U
from ubu_b200 import _native as N
from ubu_b200 import kmer as K

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kmer_oracle as KO  # noqa: E402

FIX = json.load(open(os.path.join(ROOT, "tests", "golden", "assembler_input1.json")))


# ------------------------------------------------------------------ CPU: oracle and boundary -----------------------------------
def test_oracle_reproduces_the_reference_fixture_contig():
    """input1.fastq / output1.fasta (AssemblerTest.java:22-44, k=3, minNodeFrequency=2, minContigLength=4) -> GCTCCAG.
    The fixture predates the unique-read rule (Node.java:78-94): its reads are two identical pairs, so with the rule as the
    current code has it only TCC survives and no contig of length >= 4 comes out. Both facts are asserted: the count filter
    (Assembler.java:205) + graph + walk give the expected contig, and the full current filter gives none."""
    a = KO.Assembler(FIX["kmer_size"], FIX["min_node_frequency"])
    assert a.add_reads(FIX["reads"]) == 0
    assert {k: n.count for k, n in a.nodes.items()} == {"GCT": 2, "CTC": 2, "TCC": 4, "CCA": 2, "CAG": 2, "GGG": 3}
    assert [t.sequence for t in a.nodes["GGG"].to_nodes] == ["GGG"] and [t.sequence for t in a.nodes["GGG"].from_nodes] == ["GGG"]
    # count filter only (the code the fixture was made with)
    for n in [n for n in a.nodes.values() if n.count < a.min_node_frequency]:
        n.remove(); a.nodes.pop(n.sequence)
    a.identify_root_nodes()
    assert [n.sequence for n in a.root_nodes] == ["GCT"]          # GGG is its own predecessor: not a root
    assert a.build_contigs(FIX["min_contig_length"]) == FIX["expected_contigs"] == ["GCTCCAG"]
    # the current code
    b = KO.Assembler(FIX["kmer_size"], FIX["min_node_frequency"])
    b.add_reads(FIX["reads"])
    b.filter_low_frequency_nodes(); b.identify_root_nodes()
    assert sorted(b.nodes) == ["TCC"] and b.build_contigs(FIX["min_contig_length"]) == []


def test_oracle_edge_order_and_unique_read_rule():
    a = KO.Assembler(3, 1)
    a.add_reads(["ACGTA", "ACGCA", "ACGTA", "NACGT", "CGT"])
    assert [t.sequence for t in a.nodes["ACG"].to_nodes] == ["CGT", "CGC"]        # creation order
    assert a.nodes["ACG"].has_multiple_unique_reads and a.nodes["CGC"].has_multiple_unique_reads is False
    assert a.nodes["GTA"].count == 2 and not a.nodes["GTA"].has_multiple_unique_reads   # same read string twice
    assert a.nodes["CGT"].count == 3 and a.nodes["CGT"].has_multiple_unique_reads        # ACGTA, ACGTA, CGT
    rows, skipped = KO.build_graphs(["ACGTA", "ACGCA", "ACGTA", "NACGT", "CGT"], None, 3, 1)
    assert skipped == 1 and [r["kmer"] for r in rows][:3] == ["ACG", "CGT", "GTA"]


def test_kmer_header_symbols_exported_and_bound():
    hdr = open(os.path.join(ROOT, "include", "ubu_kmer.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(ubu_kmer_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) == 9
    lib = N.lib()
    bound = {n for n, _, _ in N.SYMBOLS}
    for n in names:
        assert hasattr(lib, n) and n in bound, n
    assert K.NODE_DTYPE.itemsize == 64 and C.sizeof(N.KmerParams) == 8
    assert b"bound" in lib.ubu_sw_strerror(N.ERR_CAPACITY)


@pytest.mark.skipif(N.lib().ubu_sw_device_count() > 0, reason="a GPU is present")
def test_kmer_product_path_fails_loudly_without_a_gpu():
    with pytest.raises(U.UbuSwError) as e:
        K.KmerGraphBuilder(31, 2)
    assert e.value.status == N.ERR_NO_DEVICE


# ------------------------------------------------------------------ GPU: parity against the oracle ------------------------------
def compare_with_oracle(nodes, skipped, reads, regions, k, min_freq):
    rows, oskip = KO.build_graphs(reads, regions, k, min_freq)
    assert skipped == oskip
    assert nodes.shape[0] == len(rows), (nodes.shape[0], len(rows))
    for j, (nd, r) in enumerate(zip(nodes, rows)):
        got = {"region": int(nd["region"]), "kmer": K.kmer_string(nd, k), "count": int(nd["count"]), "first_read": int(nd["first_read"]),
               "first_pos": int(nd["first_pos"]), "multi": bool(nd["flags"] & N.KMER_FLAG_MULTI_READS),
               "filtered": bool(nd["flags"] & N.KMER_FLAG_FILTERED), "root": bool(nd["flags"] & N.KMER_FLAG_ROOT),
               "to": [(int(nodes[int(t)]["region"]), K.kmer_string(nodes[int(t)], k)) for t in nd["to_node"] if t != N.KMER_NONE],
               "from": [(int(nodes[int(t)]["region"]), K.kmer_string(nodes[int(t)], k)) for t in nd["from_node"] if t != N.KMER_NONE]}
        assert got == r, (j, got, r)
        # UBU_KMER_NONE only as padding behind the real entries
        for lst in (nd["to_node"], nd["from_node"]):
            seen_none = False
            for t in lst:
                assert not (seen_none and t != N.KMER_NONE)
                seen_none = seen_none or t == N.KMER_NONE


def sample_reads(rng, n_regions, reads_per_region, L, ref_len, sub=0.01, dup=0.2, with_n=0.02, ragged=False):
    reads, regions = [], []
    for reg in range(n_regions):
        ref = "".join(rng.choice(list("ACGT"), size=ref_len))
        if reg % 3 == 1:                                   # a tandem repeat inside: branching, cycles
            u = "".join(rng.choice(list("ACGT"), size=int(rng.integers(1, 6))))
            ref = ref[: ref_len // 3] + (u * ref_len)[: ref_len // 3] + ref[2 * ref_len // 3:]
        mine = []
        for _ in range(reads_per_region):
            if mine and rng.random() < dup:
                mine.append(mine[int(rng.integers(0, len(mine)))])          # exact duplicate read (unique-read rule)
                continue
            ll = int(rng.integers(max(L // 3, 1), L + 1)) if ragged else L
            s = int(rng.integers(0, max(len(ref) - ll, 0) + 1))
            r = list(ref[s:s + ll])
            for p in range(len(r)):
                if rng.random() < sub:
                    r[p] = "ACGT"[int(rng.integers(0, 4))]
            if rng.random() < with_n and r:
                r[int(rng.integers(0, len(r)))] = "N"
            mine.append("".join(r))
        reads += mine; regions += [reg] * len(mine)
    order = rng.permutation(len(reads))                    # regions interleaved in the batch
    return [reads[i] for i in order], np.array([regions[i] for i in order], dtype=np.uint32)


@pytest.mark.gpu
def test_device_fixture_table_and_contig():
    reads = U.pack_texts(FIX["reads"])
    with K.KmerGraphBuilder(FIX["kmer_size"], FIX["min_node_frequency"]) as b:
        nodes, skipped = b.build(reads)
    compare_with_oracle(nodes, skipped, FIX["reads"], None, 3, 2)
    assert [K.kmer_string(n, 3) for n in nodes] == ["GCT", "CTC", "TCC", "CCA", "CAG", "GGG"]
    by_count = nodes["count"] < FIX["min_node_frequency"]     # the filter the fixture was made with (see the oracle test)
    assert K.contigs_from_table(nodes, 3, FIX["min_contig_length"], filtered=by_count) == ["GCTCCAG"]
    assert K.contigs_from_table(nodes, 3, FIX["min_contig_length"]) == []


@pytest.mark.gpu
@pytest.mark.parametrize("k", [1, 3, 11, 25, 31, 32, 33, 47, 63, 64])
def test_device_graph_equals_oracle(k):
    rng = np.random.default_rng(100 + k)
    reads, regions = sample_reads(rng, n_regions=5, reads_per_region=120, L=100, ref_len=260)
    regions[::17] = N.KMER_NONE                                # reads the caller excludes (X0 > 1)
    with K.KmerGraphBuilder(k, 3) as b:
        nodes, skipped = b.build(U.pack_texts(reads), regions)
    compare_with_oracle(nodes, skipped, reads, regions, k, 3)


@pytest.mark.gpu
def test_device_graph_ragged_reads_short_reads_and_homopolymers():
    rng = np.random.default_rng(5)
    reads, regions = sample_reads(rng, n_regions=3, reads_per_region=150, L=90, ref_len=200, ragged=True)
    reads += ["A" * 80, "A" * 80, "A" * 79, "ACAC" * 20, "CACA" * 20, "", "ACG", "T" * 30]
    regions = np.concatenate([regions, np.array([7, 7, 7, 8, 8, 8, 8, 9], dtype=np.uint32)])
    for k, mf in ((21, 2), (30, 1)):
        with K.KmerGraphBuilder(k, mf) as b:
            nodes, skipped = b.build(U.pack_texts(reads), regions)
        compare_with_oracle(nodes, skipped, reads, regions, k, mf)


@pytest.mark.gpu
def test_device_graph_capacity_error_and_determinism():
    rng = np.random.default_rng(9)
    reads, regions = sample_reads(rng, n_regions=2, reads_per_region=300, L=100, ref_len=400, with_n=0.0)
    packed = U.pack_texts(reads)
    with K.KmerGraphBuilder(33, 2) as b:
        n1, _ = b.build(packed, regions)
        n2, _ = b.build(packed, regions, max_nodes=int(n1.shape[0]))         # exact bound is enough
        assert n1.tobytes() == n2.tobytes()
        with pytest.raises(U.UbuSwError) as e:
            b.build(packed, regions, max_nodes=int(n1.shape[0]) // 2)
        assert e.value.status == N.ERR_CAPACITY
        n3, _ = b.build(packed, regions)                                     # the handle is usable after the error
        assert n1.tobytes() == n3.tobytes()


@pytest.mark.gpu
def test_device_graph_full_size_properties():
    """200k x 100 bp reads over 1000 regions, k = 63 (the value ReAligner.main runs with, ReAligner.java:1723): properties that
    need no oracle, plus two whole regions against the oracle."""
    rng = np.random.default_rng(77)
    n_regions, per, L, k = 1000, 200, 100, 63
    refs = rng.integers(0, 4, size=(n_regions, 400), dtype=np.uint8)
    starts = rng.integers(0, 400 - L + 1, size=(n_regions, per))
    codes = refs[np.arange(n_regions)[:, None, None], starts[:, :, None] + np.arange(L)[None, None, :]].reshape(-1, L)
    sub = rng.random(codes.shape) < 0.005
    codes[sub] = (codes[sub] + rng.integers(1, 4, size=int(sub.sum()), dtype=np.uint8)) & 3
    regions = np.repeat(np.arange(n_regions, dtype=np.uint32), per)
    with K.KmerGraphBuilder(k, 2) as b:
        nodes, skipped = b.build(U.pack_uniform(codes), regions)
        occ, table_bytes, launches = b.last_run_stats()
    assert skipped == 0 and occ == n_regions * per * (L - k + 1) and launches >= 5
    assert int(nodes["count"].sum()) == occ
    o = np.lexsort((nodes["kmer"][:, 1], nodes["kmer"][:, 0], nodes["region"]))          # one node per (region, k-mer)
    same = (nodes["region"][o][1:] == nodes["region"][o][:-1]) & np.all(nodes["kmer"][o][1:] == nodes["kmer"][o][:-1], axis=1)
    assert not same.any()
    first = (nodes["first_read"].astype(np.uint64) << np.uint64(16)) | nodes["first_pos"].astype(np.uint64)
    assert np.all(first[1:] > first[:-1])                                    # insertion order, strictly increasing
    # edges are symmetric and stay inside the region
    n = nodes.shape[0]
    src = np.repeat(np.arange(n, dtype=np.int64), 4)
    to = nodes["to_node"].reshape(-1).astype(np.int64); ok = to != N.KMER_NONE
    fr = nodes["from_node"].reshape(-1).astype(np.int64); ok2 = fr != N.KMER_NONE
    e1 = np.sort(src[ok] * n + to[ok]); e2 = np.sort(fr[ok2] * n + src[ok2])
    assert np.array_equal(e1, e2) and e1.shape[0] > n // 2
    assert np.all(nodes["region"][src[ok]] == nodes["region"][to[ok]])
    roots = (nodes["flags"] & N.KMER_FLAG_ROOT) != 0
    assert not np.any(roots & ((nodes["flags"] & N.KMER_FLAG_FILTERED) != 0)) and roots.sum() >= n_regions // 2
    # two regions against the oracle
    for reg in (0, 777):
        sel = np.nonzero(regions == reg)[0]
        texts = ["".join("ACTG"[c] for c in codes[i]) for i in sel]
        rows, _ = KO.build_graphs(texts, None, k, 2)
        mine = nodes[nodes["region"] == reg]
        assert mine.shape[0] == len(rows)
        for nd, r in zip(mine, rows):
            assert (K.kmer_string(nd, k), int(nd["count"]), int(nd["first_read"]) - int(sel[0]), int(nd["first_pos"]),
                    bool(nd["flags"] & N.KMER_FLAG_FILTERED), bool(nd["flags"] & N.KMER_FLAG_ROOT)) == \
                   (r["kmer"], r["count"], r["first_read"], r["first_pos"], r["filtered"], r["root"])


# ------------------------------------------------------------------ CPU: the host-side contig walk over a node table ----------------
def table_from_oracle_rows(rows, k):
    """The oracle's rows in the device table's format (what ubu_kmer_build returns), so that the host-side walk can be tested
    without a GPU."""
    index = {(r["region"], r["kmer"]): j for j, r in enumerate(rows)}
    code = {"A": 0, "C": 1, "T": 2, "G": 3}
    t = np.zeros(len(rows), dtype=K.NODE_DTYPE)
    t["to_node"] = N.KMER_NONE; t["from_node"] = N.KMER_NONE
    for j, r in enumerate(rows):
        v = sum(code[ch] << (2 * p) for p, ch in enumerate(r["kmer"]))
        t[j]["kmer"] = (v & ((1 << 64) - 1), v >> 64)
        t[j]["region"], t[j]["count"], t[j]["first_read"], t[j]["first_pos"] = r["region"], r["count"], r["first_read"], r["first_pos"]
        t[j]["flags"] = (N.KMER_FLAG_MULTI_READS if r["multi"] else 0) | (N.KMER_FLAG_FILTERED if r["filtered"] else 0) | \
                        (N.KMER_FLAG_ROOT if r["root"] else 0)
        for q, key in enumerate(r["to"]):
            t[j]["to_node"][q] = index[key]
        for q, key in enumerate(r["from"]):
            t[j]["from_node"][q] = index[key]
    return t


@pytest.mark.parametrize("k,min_freq", [(11, 1), (15, 2), (21, 2), (33, 3)])
def test_host_contig_walk_over_a_table_equals_the_oracle_walk(k, min_freq):
    rng = np.random.default_rng(40 + k)
    reads, _ = sample_reads(rng, n_regions=1, reads_per_region=150, L=70, ref_len=160, sub=0.004, dup=0.1, with_n=0.02)
    rows, _ = KO.build_graphs(reads, None, k, min_freq)
    a = KO.Assembler(k, min_freq)
    a.add_reads(reads)
    a.filter_low_frequency_nodes(); a.identify_root_nodes()
    want = a.build_contigs(min_contig_length=k + 5)
    got = K.contigs_from_table(table_from_oracle_rows(rows, k), k, min_contig_length=k + 5)
    assert sorted(got) == sorted(want) and got == want       # same contigs, and in the same order: roots and branches in creation order
    assert len(want) >= 1


@pytest.mark.gpu
def test_device_graph_locked_and_lock_free_slots_agree(monkeypatch):
    """k <= 63 uses lock-free slots (128-bit CAS), k = 64 the locked protocol; UBU_KMER_LOCKED=1 forces the latter for every k:
    both must return the same table byte for byte."""
    rng = np.random.default_rng(21)
    reads, regions = sample_reads(rng, n_regions=40, reads_per_region=400, L=100, ref_len=300, with_n=0.01)
    packed = U.pack_texts(reads)
    for k in (15, 33, 63):
        with K.KmerGraphBuilder(k, 2) as b:
            free, s1 = b.build(packed, regions)
            monkeypatch.setenv("UBU_KMER_LOCKED", "1")
            locked, s2 = b.build(packed, regions)
            monkeypatch.delenv("UBU_KMER_LOCKED")
        assert s1 == s2 and free.tobytes() == locked.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(16))
def test_device_graph_randomized_differential(seed):
    """Random k (1..64), filter threshold, region count, read length, error / duplicate / N rates per seed; the whole node table
    against the oracle."""
    rng = np.random.default_rng(500 + seed)
    k = int(rng.integers(1, 65)); mf = int(rng.integers(1, 5))
    reads, regions = sample_reads(rng, n_regions=int(rng.integers(1, 8)), reads_per_region=int(rng.integers(20, 250)), L=int(rng.integers(30, 140)),
                                  ref_len=int(rng.integers(60, 400)), sub=float(rng.random() * 0.03), dup=float(rng.random() * 0.4),
                                  with_n=float(rng.random() * 0.05), ragged=bool(seed % 2))
    with K.KmerGraphBuilder(k, mf) as b:
        nodes, skipped = b.build(U.pack_texts(reads), regions)
    compare_with_oracle(nodes, skipped, reads, regions, k, mf)
