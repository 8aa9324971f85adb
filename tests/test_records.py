"""Record-level steps around the aligner: the C++ host functions (host/records.cpp, through the
C ABI of include/ubu_records.h) against the line-by-line Python restatement of the Java (oracle/records_oracle.py), and both against
the one fixture the reference holds for this code (Sam2FastqTest.java:35-44). No GPU needed except for the composites at the end."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import post_oracle as PO  # noqa: E402
import records_oracle as O  # noqa: E402

from ubu_b200 import records as R  # noqa: E402

FIX = os.path.join(ROOT, "tests", "golden", "ref_fixtures")
HDR = "@HD\tVN:1.0\tSO:queryname\n@SQ\tSN:chr1\tLN:100000\n@SQ\tSN:chr2\tLN:100000\n"


def sam_line(q, flag, rname, pos, mapq, cigar, seq="*", qual="*", tags=()):
    return "\t".join([q, str(flag), rname, str(pos), str(mapq), cigar, "*", "0", "0", seq, qual] + list(tags)) + "\n"


def rnd_seq(rng, n):
    return "".join(rng.choice(list("ACGT"), size=n))


# ---- Sam2Fastq: pinned to the reference's own fixture -----------------------------------------------------------------------------
def test_sam2fastq_matches_the_reference_fixture():
    sam = open(os.path.join(FIX, "sam2fastq_test.sam")).read()
    e1 = open(os.path.join(FIX, "sam2fastq_expected_paired1.fastq")).read()
    e2 = open(os.path.join(FIX, "sam2fastq_expected_paired2.fastq")).read()
    assert O.sam2fastq_paired(sam, "/1", "/2") == (e1, e2)                 # the restatement is pinned ...
    assert R.sam2fastq_paired(sam, "/1", "/2") == (e1, e2)                 # ... and so is the C++ mirror
    assert R.sam2fastq(sam) == O.sam2fastq(sam)
    # single-file form: consecutive records of one name are written once, minus-strand reads are reverse-complemented
    one = R.sam2fastq(sam).split("\n")
    assert one[0] == "@XYZ13-SN749:180:D127FACXX:6:1101:1066:2717/1" and one[1].startswith("GTTNGTTTAAATC") and one.count("+") == 8


def test_sam2fastq_flag_based_pairing_and_asymmetry():
    sam = HDR + sam_line("a", 0x40 | 0x10, "chr1", 5, 30, "4M", "ACGN", "IJKL") + sam_line("a", 0x80, "chr1", 50, 30, "4M", "TTGA", "ABCD")
    assert R.sam2fastq_paired(sam) == O.sam2fastq_paired(sam) == ("@a\nNCGT\n+\nLKJI\n", "@a\nTTGA\n+\nABCD\n")
    lonely = HDR + sam_line("a", 0x40, "chr1", 5, 30, "4M", "ACGT", "IIII")
    with pytest.raises(R.RecordsError) as e:
        R.sam2fastq_paired(lonely)
    assert e.value.status == R.ERR_STATE                                     # IllegalStateException in the Java (Sam2Fastq.java:76-78)
    with pytest.raises(ValueError):
        O.sam2fastq_paired(lonely)


# ---- CombineChimera3 -------------------------------------------------------------------------------------------------------------
def test_combine_chimera_hand_made_cases():
    seq = "A" * 100
    # deletion: left 60M40S at 1000 (ref 1000..1059), right 60S40M at 1110 -> 60M50D40M
    dele = sam_line("c1", 0, "chr1", 1000, 30, "60M40S", seq) + sam_line("c1", 0, "chr1", 1110, 20, "60S40M", seq)
    # insertion: the two parts abut on the reference but skip 10 read bases -> 50M10I40M
    ins = sam_line("c2", 0, "chr1", 2000, 40, "50M50S", seq) + sam_line("c2", 0, "chr1", 2050, 50, "60S40M", seq)
    # overlap on the read: trimmed off the left side's last element
    ovl = sam_line("c3", 16, "chr2", 3000, 25, "60M40S", seq) + sam_line("c3", 16, "chr2", 3200, 35, "50S50M", seq)
    # trimmed too far: not an indel, both records stay
    far = sam_line("c4", 0, "chr1", 4000, 30, "20S10M70S", seq) + sam_line("c4", 0, "chr1", 4500, 30, "5S95M", seq)
    # parts that overlap on the reference and the read by the same amount: the Java emits a zero-length insertion
    zero = sam_line("c8", 0, "chr1", 4000, 30, "10M90S", seq) + sam_line("c8", 0, "chr1", 4005, 30, "5S95M", seq)
    # different strands / references: untouched; single-element CIGAR: untouched
    diff = sam_line("c5", 0, "chr1", 5000, 30, "60M40S", seq) + sam_line("c5", 16, "chr1", 5100, 30, "60S40M", seq)
    one = sam_line("c6", 0, "chr1", 6000, 30, "100M", seq) + sam_line("c6", 0, "chr1", 7000, 30, "60S40M", seq)
    # three hits, the middle one a likely insert (FUDGE_FACTOR 5): pruned, then the outer two combine
    three = (sam_line("c7", 0, "chr1", 8000, 30, "40M60S", seq) + sam_line("c7", 0, "chr2", 500, 30, "38S24M38S", seq) +
             sam_line("c7", 0, "chr1", 8100, 30, "60S40M", seq))
    text = HDR + dele + ins + ovl + far + diff + one + three + zero
    got, want = R.combine_chimera(text), O.combine(text)
    assert got == want
    recs = {}
    for ln in got.splitlines():
        if not ln.startswith("@"):
            recs.setdefault(ln.split("\t")[0], []).append(ln.split("\t"))
    assert got.splitlines()[0] == "@HD\tVN:1.0\tSO:unsorted"
    assert [r[5] for r in recs["c1"]] == ["60M50D40M"] and recs["c1"][0][3] == "1000" and recs["c1"][0][4] == "20"      # MAPQ = min
    assert [r[5] for r in recs["c2"]] == ["50M10I40M"]
    assert [r[5] for r in recs["c3"]] == ["50M150D50M"]
    assert len(recs["c4"]) == 2 and len(recs["c5"]) == 2 and len(recs["c6"]) == 2
    assert len(recs["c7"]) == 1 and recs["c7"][0][5] == "40M60D40M"                                    # the chr2 hit in the middle was pruned
    assert [r[5] for r in recs["c8"]] == ["5M0I95M"]


@pytest.mark.parametrize("seed", range(5))
def test_combine_chimera_randomized(seed):
    rng = np.random.default_rng(100 + seed)
    lines = []
    for k in range(300):
        L = int(rng.integers(60, 400))
        seq = rnd_seq(rng, L)
        n_hits = int(rng.choice([1, 2, 2, 2, 3]))
        name = f"contig_{k}"
        cut = sorted(int(x) for x in rng.integers(5, L - 5, size=n_hits - 1))
        bounds = [0] + cut + [L]
        base = int(rng.integers(100, 90000))
        for h in range(n_hits):
            a, b = bounds[h], bounds[h + 1]
            a = max(0, a - int(rng.integers(0, 8)))                      # parts may overlap on the read
            cig = (f"{a}S" if a else "") + f"{b - a}M" + (f"{L - b}S" if L - b else "")
            if rng.random() < 0.2 and b - a > 20:                          # an indel inside a part
                x = int(rng.integers(5, b - a - 5))
                cig = (f"{a}S" if a else "") + f"{x}M3D{b - a - x}M" + (f"{L - b}S" if L - b else "")
            pos = base + a + int(rng.integers(-20, 200)) * (h > 0)
            rn = "chr1" if rng.random() < 0.85 else "chr2"
            fl = 0 if rng.random() < 0.85 else 16
            lines.append(sam_line(name, fl, rn, pos, int(rng.integers(0, 60)), cig, seq))
    text = HDR + "".join(lines)
    assert R.combine_chimera(text) == O.combine(text)


# ---- removeSoftClips / cleanAndOutputContigs --------------------------------------------------------------------------------------
def test_clean_contigs_keeps_the_reference_quirks():
    seq = "".join("ACGT"[k % 4] for k in range(120))
    text = HDR + (sam_line("k1", 0, "chr1", 100, 30, "10S100M10S", seq, "*", ["AS:i:95"]) +      # both clips: drops 10+1 and 10+1 bases
                  sam_line("k2", 0, "chr1", 300, 30, "120M", seq) +
                  sam_line("k3", 0, "chr1", 500, 1, "120M", seq) +                             # MAPQ <= 1: dropped (:887)
                  sam_line("k4", 16, "chr1", 700, 50, "100M20S", seq) +
                  sam_line("k5", 0, "chr1", 900, 50, "30S90M", seq))                          # 90 - 1 bases left: below the minimum of 90
    got = R.clean_contigs(text, 90, True)
    assert got == O.clean_and_output_contigs(text, 90, True)
    assert "k1" not in R.clean_contigs(text, 100, True)[0]                                      # 100 aligned bases, 98 kept: the quirk costs the contig
    fasta, has = got
    assert has
    recs = fasta.split(">")[1:]
    names = [r.split("~")[0] for r in recs]
    assert names == ["k1", "k2", "k4"]
    k1 = recs[0]
    # the name carries the whole SAM record with '~' for TAB, SEQ blanked to '*', and getSAMString's newline ends the header line
    assert k1.startswith("k1~k1~0~chr1~100~30~100M~*~0~0~*~*~AS:i:95\n")
    assert k1.split("\n")[1] == seq[10:119][: 109 - 10 - 1]                                    # off-by-one on both sides (:942, :953)
    assert recs[2].split("\n")[1] == seq[: 120 - 20 - 1]
    # without soft-clip removal the bases stay whole
    f2, _ = R.clean_contigs(text, 90, False)
    assert f2 == O.clean_and_output_contigs(text, 90, False)[0] and f2.split(">")[1].split("\n")[1] == seq


@pytest.mark.parametrize("seed", range(3))
def test_clean_contigs_randomized(seed):
    rng = np.random.default_rng(200 + seed)
    lines = []
    for k in range(400):
        L = int(rng.integers(40, 300))
        a, b = int(rng.integers(0, 15)) * int(rng.random() < 0.4), int(rng.integers(0, 15)) * int(rng.random() < 0.4)
        if a + b + 3 >= L:
            a = b = 0
        cig = (f"{a}S" if a else "") + f"{L - a - b}M" + (f"{b}S" if b else "")
        lines.append(sam_line(f"c{k}", int(rng.choice([0, 16])), "chr1", int(rng.integers(1, 90000)), int(rng.integers(0, 5)) * int(rng.integers(0, 20)),
                              cig, rnd_seq(rng, L), "*", [f"AS:i:{L}", "XS:i:0"]))
    text = HDR + "".join(lines)
    for rm in (True, False):
        assert R.clean_contigs(text, 100, rm) == O.clean_and_output_contigs(text, 100, rm)


# ---- updateReadAlignment: C++ host mirror vs the restatement (the device kernel is checked against the same restatement in test_post.py) ----
def test_update_read_alignment_randomized():
    rng = np.random.default_rng(7)
    checked = raised = nulls = 0
    for _ in range(4000):
        n = int(rng.integers(1, 7))
        cig = []
        for k in range(n):
            op = int(rng.choice([0, 0, 0, 1, 2, 3, 4])) if 0 < k < n - 1 else int(rng.choice([0, 0, 4]))
            cig.append((int(rng.integers(1, 120)), op))
        pos, position, rl = int(rng.integers(1, 5000)), int(rng.integers(0, 300)), int(rng.integers(20, 151))
        try:
            want = PO.project_read(pos, cig, position, rl)
        except ValueError:
            with pytest.raises(R.RecordsError) as e:
                R.update_read_alignment(pos, cig, position, rl)
            assert e.value.status == R.ERR_STATE
            raised += 1
            continue
        got = R.update_read_alignment(pos, cig, position, rl)
        assert got == (None if want is None else (want[0], list(want[1]))), (pos, cig, position, rl)
        checked += 1; nulls += want is None
    assert checked > 3000 and nulls > 0


# ---- adjustReads -----------------------------------------------------------------------------------------------------------------
def contig_name(qname, flag, rname, pos, mapq, cigar, tags=("AS:i:300",)):
    """What cleanAndOutputContigs puts into the FASTA name and bwa then reports as RNAME (up to the first white space)."""
    return qname + "~" + "~".join([qname, str(flag), rname, str(pos), str(mapq), cigar, "*", "0", "0", "*", "*"] + list(tags))


def test_adjust_reads_gate_hits_and_tags():
    L = 100
    s1, q1 = "ACGT" * 25, "I" * 99 + "#"
    cA = contig_name("ctgA", 0, "chr1", 5000, 50, "150M5D150M")
    cB = contig_name("ctgB", 0, "chr1", 9000, 50, "300M")
    cLow = contig_name("ctgL", 0, "chr2", 100, 10, "300M")
    orig = HDR + "".join([
        sam_line("r1", 0, "chr1", 5095, 20, "100M", s1, q1, ["NM:i:6", "X0:i:1", "X1:i:0", "MD:Z:100", "XT:A:U"]),    # adjusted: spans the contig's deletion
        sam_line("r2", 0, "chr1", 5000, 20, "100M", s1, q1, ["NM:i:0"]),                                          # contig NM not smaller: kept
        sam_line("r3", 0, "chr1", 7000, 20, "100M", s1, q1, ["NM:i:5"]),                                          # contig MAPQ (10) <= orig MAPQ (20): kept
        sam_line("r4", 4, "*", 0, 0, "*", s1, q1),                                                                 # unmapped original, two placements -> MAPQ 0
        sam_line("r5", 0, "chr1", 100, 20, "100M", s1, q1, ["NM:i:5"]),                                           # to-contig CIGAR has an indel: kept
        sam_line("r6", 16, "chr1", 100, 20, "100M", s1, q1, ["NM:i:5"]),                                          # hit past the end of the contig: filtered
        sam_line("r7", 0, "chr1", 100, 20, "100M", s1, q1, ["NM:i:5"]),                                           # no record in the to-contig file, and behind its last one
    ])
    to_contig = HDR + "".join([
        sam_line("r1", 0, cA, 96, 37, "100M", s1, q1, ["NM:i:1", "XM:i:1", "X0:i:1", "X1:i:0"]),
        sam_line("r2", 0, cA, 1, 37, "100M", s1, q1, ["NM:i:0", "XM:i:0", "X0:i:1", "X1:i:0"]),
        sam_line("r3", 0, cLow, 10, 37, "100M", s1, q1, ["NM:i:0", "XM:i:0", "X0:i:1", "X1:i:0"]),
        sam_line("r4", 16, cA, 20, 0, "100M", s1, q1, ["NM:i:2", "XM:i:2", "X0:i:2", "X1:i:1",
                                                       f"XA:Z:{cB},+40,100M,2;{cB},+70,100M,2;{cA},-200,100M,2;"]),   # the LAST real entry is skipped (:1137)
        sam_line("r5", 0, cA, 10, 37, "60M2I38M", s1, q1, ["NM:i:2", "XM:i:0", "X0:i:1", "X1:i:0"]),
        sam_line("r6", 0, cB, 250, 37, "100M", s1, q1, ["NM:i:0", "XM:i:0", "X0:i:1", "X1:i:0"]),
    ])
    got, st = R.adjust_reads(orig, to_contig)
    want, wst = O.adjust_reads(orig, to_contig)
    assert got == want and st == wst
    by = {}
    for ln in got.splitlines():
        by.setdefault(ln.split("\t")[0], []).append(ln.split("\t"))
    r1 = by["r1"][0]
    assert (r1[2], r1[3], r1[5]) == ("chr1", "5095", "55M5D45M") and "YO:Z:chr1:5095:+:100M" in r1 and "YM:i:1" in r1 and "YQ:i:50" in r1
    assert "X0:i:1" in r1 and "X1:i:1" in r1 and not any(t.startswith(("MD:", "XT:", "XA:")) for t in r1)          # tags rewritten / cleared (:1254-1262)
    assert by["r2"][0][5] == "100M" and not any(t.startswith("YO") for t in by["r2"][0])
    assert by["r3"][0][3] == "7000"
    r4 = by["r4"]
    assert len(r4) == 3 and all(x[4] == "0" for x in r4) and all("YO:Z:N/A" in x for x in r4)                    # primary + 2 of the 3 XA entries, MAPQ 0 (:1249)
    assert sorted((x[2], int(x[3])) for x in r4) == [("chr1", 5019), ("chr1", 9039), ("chr1", 9069)]
    minus = [x for x in r4 if int(x[1]) & 16][0]
    plus = [x for x in r4 if not int(x[1]) & 16][0]
    assert minus[9] == s1 and plus[9] == O._revcomp(s1) and plus[10] == q1[::-1]                                 # strand flips relative to the primary hit (:1200-1205)
    assert by["r5"][0][3] == "100" and by["r6"][0][3] == "100"
    assert "r7" not in by            # the merge join stops with the to-contig file (:1070): originals behind its last record are not written
    assert st == {"missing_xa": 0, "mismatch_issue": 0, "adjusted": 4}


def test_adjust_reads_literal_100m_gate_and_its_generalisation():
    s = "ACGT" * 19                                                            # 76 bp reads: the literal "100M" gate never fires
    cA = contig_name("ctgA", 0, "chr1", 5000, 50, "300M")
    orig = HDR + sam_line("r1", 0, "chr1", 1, 20, "76M", s, "*", ["NM:i:4"])
    toc = HDR + sam_line("r1", 0, cA, 11, 37, "76M", s, "*", ["NM:i:0", "XM:i:0", "X0:i:1", "X1:i:0"])
    for fm in (None, ""):
        assert R.adjust_reads(orig, toc, fm) == O.adjust_reads(orig, toc, "100M" if fm is None else fm)
    assert R.adjust_reads(orig, toc)[1]["adjusted"] == 0                       # reference behaviour
    out, st = R.adjust_reads(orig, toc, "")                                    # "<read length>M": the flagged generalisation
    assert st["adjusted"] == 1 and out.split("\t")[3] == "5010" and out.split("\t")[5] == "76M"


@pytest.mark.parametrize("seed", range(4))
def test_adjust_reads_randomized(seed):
    rng = np.random.default_rng(300 + seed)
    contigs = []
    for k in range(12):
        cig = rng.choice(["300M", "120M4D180M", "100M7I193M", "20S260M20S", "150M300N150M", "80M2I60M3D158M"])
        contigs.append(contig_name(f"ctg{k}", int(rng.choice([0, 16])), rng.choice(["chr1", "chr2"]), int(rng.integers(100, 80000)), int(rng.integers(0, 60)), cig))
    o_lines, c_lines = [], []
    for k in range(500):
        name = f"read{k:05d}"
        s, q = rnd_seq(rng, 100), "".join(chr(int(x)) for x in rng.integers(35, 74, size=100))
        unm = rng.random() < 0.1
        o_lines.append(sam_line(name, 4 if unm else int(rng.choice([0, 16])), "*" if unm else "chr1", 0 if unm else int(rng.integers(1, 90000)),
                                0 if unm else int(rng.integers(0, 40)), "*" if unm else "100M", s, q,
                                [] if rng.random() < 0.2 else [f"NM:i:{int(rng.integers(0, 8))}", "X0:i:1", "X1:i:0", "MD:Z:100", "XA:Z:foo;"]))
        if rng.random() < 0.1:
            continue                                                          # the to-contig file may miss reads (merge join, :1264-1266)
        c = contigs[int(rng.integers(0, len(contigs)))]
        x0, x1 = int(rng.choice([1, 1, 2, 3])), int(rng.choice([0, 0, 1, 2]))
        tags = [f"NM:i:{int(rng.integers(0, 4))}", f"XM:i:{int(rng.integers(0, 3))}", f"X0:i:{x0}", f"X1:i:{x1}"]
        if x0 + x1 > 1 and rng.random() < 0.9:
            ents = []
            for _ in range(x0 + x1 - 1):
                cc = contigs[int(rng.integers(0, len(contigs)))]
                ents.append(f"{cc},{rng.choice(['+', '-'])}{int(rng.integers(1, 260))},{rng.choice(['100M', '100M', '50M1I49M'])},{int(rng.integers(0, 3))};")
            tags.append("XA:Z:" + "".join(ents))
        cigar = rng.choice(["100M", "100M", "100M", "60M1D40M"])
        fl = 4 if rng.random() < 0.05 else int(rng.choice([0, 16]))
        c_lines.append(sam_line(name, fl, c, int(rng.integers(1, 260)), 37, cigar, s, q, tags))
    orig, toc = HDR + "".join(o_lines), HDR + "".join(c_lines)
    try:
        want = O.adjust_reads(orig, toc)
    except ValueError:                                                         # an N-containing contig made updateReadAlignment throw, as in the Java
        with pytest.raises(R.RecordsError):
            R.adjust_reads(orig, toc)
        return
    got = R.adjust_reads(orig, toc)
    assert got[1] == want[1] and got[0] == want[0]
    assert want[1]["adjusted"] > 20


def test_num_differences_compares_iupac_codes_as_letters():
    """CompareToReference.java:90-94: 'R' vs 'A' is a mismatch, 'N' never is, case is folded. The device recount folds every non-ACGT
    letter into N (documented divergence); the host function is letter-exact and countNonAcgtn says which reads need it."""
    assert R.num_differences("ACGR", "ACGA", 1, "4M") == 1 and R.num_differences("ACGR", "ACGR", 1, "4M") == 0
    assert R.num_differences("acgn", "ACGT", 1, "4M") == 0 and R.num_differences("ACGT", "ACGN", 1, "4M") == 0
    assert R.count_non_acgtn("ACGTNacgtn") == 0 and R.count_non_acgtn("ACRYGT-") == 3
    rng = np.random.default_rng(3)
    for _ in range(300):
        ref = "".join(rng.choice(list("ACGTNRYKMacgt"), size=400))
        cig, rl, fl = [], 0, 0
        for k in range(int(rng.integers(1, 6))):
            op = int(rng.choice([0, 0, 1, 2, 3, 4])) if k else 0
            n = int(rng.integers(1, 40)); cig.append((n, op))
            rl += n if op in (0, 1, 4) else 0; fl += n if op in (0, 2) else 0
        read = "".join(rng.choice(list("ACGTNRWacgt"), size=rl))
        start = int(rng.integers(1, 400 - fl))
        want = PO.num_differences(read, ref, start, cig)
        assert R.num_differences(read, ref, start, O.cigar_text(cig)) == want
    with pytest.raises(R.RecordsError):
        R.num_differences("ACGT", "AC", 1, "4M")                                  # StringIndexOutOfBounds in the Java


def test_records_library_exports_every_symbol_of_the_header():
    l = R.lib()
    hdr = open(os.path.join(ROOT, "include", "ubu_records.h")).read()
    for name in R.SYMBOLS:
        assert name in hdr
        getattr(l, name)
    import re
    assert set(re.findall(r"\b(ubu_(?:rec|aligner)_[a-z0-9_]+)\s*\(", hdr)) == set(R.SYMBOLS)
    assert R.lib().ubu_aligner_index(b"/nonexistent/ref.fa") == R.ERR_ALIGNER        # the reference throws RuntimeException (Aligner.java:41-43)


# ---- composites on the GPU: alignAndCleanContigs -> alignToContigs -> adjustReads --------------------------------------------------
@pytest.mark.gpu
def test_contig_pipeline_through_the_gpu_aligner(tmp_path):
    """ReAligner.alignAndCleanContigs (ReAligner.java:351-367) and alignToContigs (:963-975) with the GPU aligner behind the reference's
    Aligner facade, then adjustReads on what they wrote: a region with a 12 bp deletion, a contig spanning it, 100 bp reads from the
    contig. The reads' original alignments are soft (mismatching) placements; after the pipeline the reads that cross the deletion carry
    it in their CIGAR and the tags the pipeline writes."""
    rng = np.random.default_rng(11)
    region = rnd_seq(rng, 3000)
    contig = region[400:1000] + region[1012:1700]                            # the sample has lost 12 bp
    (tmp_path / "ref.fa").write_text(">chr1\n" + region + "\n")
    (tmp_path / "contigs.fa").write_text(">ctg1\n" + contig + "\n")
    clean = R.align_and_clean_contigs(str(tmp_path / "ref.fa"), str(tmp_path / "contigs.fa"), str(tmp_path), True, min_contig_length=100)
    assert clean is not None and os.path.exists(clean)
    contig_sam = [l for l in open(tmp_path / "all_contigs_chim.sam") if not l.startswith("@")]
    import re
    m = re.fullmatch(r"(\d+)M12D(\d+)M", contig_sam[0].split("\t")[5])         # the gap may sit a base or two left of 600 when the flanks repeat
    assert len(contig_sam) == 1 and m and int(m.group(1)) + int(m.group(2)) == 1288 and contig_sam[0].split("\t")[3] == "401"
    A = int(m.group(1))
    assert 590 <= A <= 600
    # original reads: 100 bp pieces of the contig; their original records claim an ungapped placement with a large NM
    reads, lines = [], []
    for k, s in enumerate(range(0, len(contig) - 100, 37)):
        seq = contig[s:s + 100]
        reads.append((f"rd{k:04d}", s, seq))
        lines.append(sam_line(f"rd{k:04d}", 0, "chr1", 401 + s, 10, "100M", seq, "I" * 100, ["NM:i:9", "X0:i:1", "X1:i:0"]))
    (tmp_path / "orig.sam").write_text("@HD\tVN:1.0\tSO:queryname\n@SQ\tSN:chr1\tLN:3000\n" + "".join(lines))
    R.align_to_contigs(str(tmp_path), str(tmp_path / "orig.sam"), str(tmp_path / "to_contig.sam"), clean)
    toc = open(tmp_path / "to_contig.sam").read()
    orig = open(tmp_path / "orig.sam").read()
    got, st = R.adjust_reads(orig, toc)
    want, wst = O.adjust_reads(orig, toc)
    assert got == want and st == wst and st["adjusted"] == len(reads)
    by = {l.split("\t")[0]: l.split("\t") for l in got.splitlines()}
    for name, s, seq in reads:
        r = by[name]
        # removeSoftClips' off-by-one leaves the contig one base short, nothing else: placements are unchanged
        if s <= A - 100 or s >= A:
            assert r[5] == "100M" and int(r[3]) == (401 + s if s < A else 413 + s)
        else:
            assert r[5] == f"{A - s}M12D{100 - (A - s)}M" and int(r[3]) == 401 + s
        assert "YM:i:0" in r and any(t.startswith("YQ:i:") for t in r)
