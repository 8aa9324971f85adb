"""The path-in/path-out facade that mirrors the reference's Aligner (Aligner.java:13-16,18-22,46-56,62-69).
CPU tests run it with the oracle standing in for the device call (they test the text I/O and the record contract the
consumer reads back, ReAligner.java:1088-1164); the gpu test checks the real device path writes the same file."""
import os

import numpy as np
import pytest

import ubu_b200 as U
from ubu_b200.facade import Aligner, read_fastx
from test_sharding import oracle_align_fn

CONTIG_A = "TTGACCATAGGCTAACGTTAGCCATGGATCCGATTACAGGCATCGATTGCAGTCCGATAGCTAGGCTTAACGGATCTAGCATCGGATTTACGGCAT"
CONTIG_B = "GGCATTAGCGGATCCAGTTAGGCTAGCTAAGCGGCTTTAGACGATCGGCTAATCGCGGATATCGCTAGGCTTTAGACCAGT"


def write_inputs(tmp_path, reads, refs):
    ref = tmp_path / "contigs.fasta"
    ref.write_text("".join(f">{n} some description\n{s[:40]}\n{s[40:]}\n" for n, s in refs))
    fq = tmp_path / "reads.fastq"
    fq.write_text("".join(f"@{n}\n{s}\n+\n{'I' * (len(s) - 1)}#\n" for n, s in reads))
    return str(ref), str(fq)


def parse_sam(path):
    recs = []
    for line in open(path):
        if line.startswith("@"):
            continue
        f = line.rstrip("\n").split("\t")
        recs.append({"qname": f[0], "flag": int(f[1]), "rname": f[2], "pos": int(f[3]), "mapq": int(f[4]), "cigar": f[5],
                     "seq": f[9], "qual": f[10], "tags": dict((t[:2], t[5:]) for t in f[11:])})
    return recs


def test_fastx_readers(tmp_path):
    ref, fq = write_inputs(tmp_path, [("r1", "ACGT"), ("r2", "GGCC")], [("cA", CONTIG_A)])
    assert [(r.name, r.seq) for r in read_fastx(ref)] == [("cA", CONTIG_A)]
    assert [(r.name, r.seq, r.qual) for r in read_fastx(fq)] == [("r1", "ACGT", "III#"), ("r2", "GGCC", "III#")]


def run(tmp_path, reads, refs, align_fn=oracle_align_fn):
    ref, fq = write_inputs(tmp_path, reads, refs)
    sam = str(tmp_path / "out.sam")
    a = Aligner(ref, 4, align_fn=align_fn)
    a.index()
    a.shortAlign(fq, sam)
    a.close()
    return parse_sam(sam), open(sam).read()


def test_records_carry_what_adjust_reads_consumes(tmp_path):
    fwd = CONTIG_A[10:60]
    rc = U.reverse_complement(CONTIG_B[5:55])
    mm = CONTIG_A[20:45] + "A" + CONTIG_A[46:70]                  # one mismatch (original base is C or G, never A here?)
    reads = [("fwd", fwd), ("rev", rc), ("mm", mm), ("junk", "ACACACAC" * 0 + "NNNNNNNN"), ("clip", "GGGGGG" + CONTIG_B[30:70])]
    recs, text = run(tmp_path, reads, [("cA", CONTIG_A), ("cB", CONTIG_B)])
    assert text.startswith("@HD") and "@SQ\tSN:cA\tLN:%d" % len(CONTIG_A) in text
    by = {r["qname"]: r for r in recs}
    assert [r["qname"] for r in recs] == [n for n, _ in reads]                      # input order
    assert by["fwd"]["flag"] == 0 and by["fwd"]["rname"] == "cA" and by["fwd"]["pos"] == 11 and by["fwd"]["cigar"] == "50M"
    assert by["fwd"]["tags"]["NM"] == "0" and by["fwd"]["tags"]["XM"] == "0" and by["fwd"]["tags"]["X0"] == "1"
    assert by["fwd"]["mapq"] == 37
    # minus strand: SEQ is written reverse-complemented (= the contig's own bases), QUAL reversed, flag 0x10
    assert by["rev"]["flag"] == 16 and by["rev"]["rname"] == "cB" and by["rev"]["pos"] == 6 and by["rev"]["seq"] == CONTIG_B[5:55]
    assert by["rev"]["qual"] == "#" + "I" * 49
    expect_nm = 0 if CONTIG_A[45] == "A" else 1
    assert by["mm"]["tags"]["NM"] == str(expect_nm) and by["mm"]["cigar"] == "50M"
    assert by["junk"]["flag"] == 4 and by["junk"]["cigar"] == "*" and by["junk"]["rname"] == "*"
    assert by["clip"]["cigar"] == "6S40M" and by["clip"]["pos"] == 31


def test_equal_hits_go_to_xa_and_zero_mapq(tmp_path):
    reads = [("dup", CONTIG_A[15:65])]
    recs, _ = run(tmp_path, reads, [("c1", CONTIG_A), ("c2", "TT" + CONTIG_A), ("c3", CONTIG_B)])
    r = recs[0]
    assert r["rname"] == "c1" and r["tags"]["X0"] == "2" and r["mapq"] == 0
    assert r["tags"]["XA"] == "c2,+18,50M,0;"                                       # name,±pos,CIGAR,NM; (ReAligner.java:1135-1143)
    # the consumer's own parse (split on ';', then ','), applied to our string
    alt = r["tags"]["XA"].split(";")[0].split(",")
    assert alt[0] == "c2" and alt[1][0] == "+" and int(alt[1][1:]) == 18 and alt[2] == "50M" and int(alt[3]) == 0


def test_failure_surfaces_as_runtime_error(tmp_path):
    a = Aligner(str(tmp_path / "missing.fa"), 1, align_fn=oracle_align_fn)
    with pytest.raises(RuntimeError):
        a.index()
    with pytest.raises(RuntimeError):
        a.shortAlign(str(tmp_path / "nope.fq"), str(tmp_path / "o.sam"))


def test_seed_filter_aligns_only_candidate_pairs(tmp_path):
    """seed_k > 0 (BUILD-DEFINED scalability switch): a read meets only the sequences it shares an exact k-mer with. Reads with such a
    seed get the record all-vs-all gives them (X0/X1 may only shrink); the number of alignment tasks drops from reads x contigs x 2."""
    rng = np.random.default_rng(21)
    contigs = [(f"ctg{k}", "".join(rng.choice(list("ACGT"), size=300))) for k in range(12)]
    reads = []
    for k in range(60):
        c = contigs[int(rng.integers(0, 12))][1]
        s0 = int(rng.integers(0, 220))
        r = list(c[s0:s0 + 76]); r[10] = "A" if r[10] != "A" else "C"; r[50] = "G" if r[50] != "G" else "T"
        r = "".join(r)
        reads.append((f"r{k}", U.reverse_complement(r) if k % 2 else r))
    reads.append(("noseed", "ACGTTGCA" * 6))
    ref, fq = write_inputs(tmp_path, reads, contigs)
    counts = []

    def counting(rs, ws, ix):
        counts.append(rs.count)
        return oracle_align_fn(rs, ws, ix)

    out = {}
    for k in (0, 16):
        counts.clear()
        a = Aligner(ref, 1, align_fn=counting, seed_k=k)
        a.shortAlign(fq, str(tmp_path / f"o{k}.sam")); a.close()
        out[k] = (parse_sam(str(tmp_path / f"o{k}.sam")), sum(counts), open(tmp_path / f"o{k}.sam").read())
    assert out[0][1] == len(reads) * len(contigs) * 2 and out[16][1] <= 2 * len(reads)           # ~1 task per read instead of 24
    assert "candidates = sequences sharing an exact 16-mer" in out[16][2] and "candidates" not in out[0][2]
    for full, seeded in zip(out[0][0], out[16][0]):
        if full["qname"] == "noseed":
            assert seeded["flag"] == 4                                                           # no exact 16-mer anywhere: not attempted
            continue
        for f in ("flag", "rname", "pos", "cigar", "seq"):
            assert full[f] == seeded[f], (full["qname"], f)
        assert full["tags"]["NM"] == seeded["tags"]["NM"] and int(seeded["tags"]["X0"]) <= int(full["tags"]["X0"])


@pytest.mark.gpu
def test_device_writes_the_same_sam_as_the_oracle_stand_in(tmp_path):
    rng = np.random.default_rng(4)
    contigs = [(f"ctg{k}", "".join(rng.choice(list("ACGT"), size=int(rng.integers(150, 400))))) for k in range(5)]
    reads = []
    for k in range(120):
        c = contigs[int(rng.integers(0, 5))][1]
        s = int(rng.integers(0, len(c) - 80))
        r = list(c[s:s + 76])
        for p in range(76):
            if rng.random() < 0.03:
                r[p] = "ACGT"[int(rng.integers(0, 4))]
        r = "".join(r)
        if rng.random() < 0.5:
            r = U.reverse_complement(r)
        reads.append((f"read{k}", r))
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    _, dev = run(tmp_path / "a", reads, contigs, align_fn=None)
    _, ora = run(tmp_path / "b", reads, contigs)
    assert dev == ora


CLI = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "host", "ubu_aligner")


def _build_cli():
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(CLI):
        subprocess.run(["g++", "-O2", "-std=c++17", "-o", CLI, os.path.join(root, "host", "ubu_aligner_main.cpp"),
                        os.path.join(root, "host", "aligner.cpp"), "-L" + os.path.join(root, "ubu_b200"), "-lubu_sw",
                        "-Wl,-rpath,$ORIGIN/../ubu_b200"], check=True)


def test_cpp_host_cli_fails_like_the_reference_without_its_inputs(tmp_path):
    """host/aligner.cpp mirrors Aligner.java's names and error behaviour (RuntimeException -> message + exit 1)."""
    import subprocess
    _build_cli()
    r = subprocess.run([CLI, "shortAlign", str(tmp_path / "missing.fa"), "x.fq", str(tmp_path / "o.sam")], capture_output=True, text=True)
    assert r.returncode == 1 and "reference not found" in r.stderr
    assert subprocess.run([CLI], capture_output=True).returncode == 2


@pytest.mark.gpu
def test_cpp_host_writes_the_same_sam_as_the_python_facade(tmp_path):
    import subprocess
    _build_cli()
    rng = np.random.default_rng(14)
    contigs = [(f"ctg{k}", "".join(rng.choice(list("ACGTN"), p=[.245, .245, .245, .245, .02], size=int(rng.integers(200, 600))))) for k in range(4)]
    contigs.append(("dup", contigs[0][1]))
    reads = []
    for k in range(150):
        c = contigs[int(rng.integers(0, 4))][1]
        s = int(rng.integers(0, len(c) - 110))
        r = list(c[s:s + int(rng.integers(40, 100))])
        for p in range(len(r)):
            if rng.random() < 0.04:
                r[p] = "ACGT"[int(rng.integers(0, 4))]
        if rng.random() < 0.2:
            del r[len(r) // 2]
        r = "".join(r)
        reads.append((f"read{k}", U.reverse_complement(r) if rng.random() < 0.5 else r))
    reads.append(("junk", "N" * 30))
    ref, fq = write_inputs(tmp_path, reads, contigs)
    a = Aligner(ref, 1); a.shortAlign(fq, str(tmp_path / "py.sam")); a.close()
    r = subprocess.run([CLI, "shortAlign", ref, fq, str(tmp_path / "cpp.sam")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "py.sam").read() == open(tmp_path / "cpp.sam").read()
    # the same with the seed filter on both sides (ubu::Aligner::setSeedLength / UBU_ALIGNER_SEED_K == facade.Aligner(seed_k=...))
    a = Aligner(ref, 1, seed_k=14); a.shortAlign(fq, str(tmp_path / "py14.sam")); a.close()
    r = subprocess.run([CLI, "shortAlign", ref, fq, str(tmp_path / "cpp14.sam")], capture_output=True, text=True, env=dict(os.environ, UBU_ALIGNER_SEED_K="14"))
    assert r.returncode == 0, r.stderr
    assert open(tmp_path / "py14.sam").read() == open(tmp_path / "cpp14.sam").read()
