"""CPU restatement of the reference's record-level steps around the aligner.

TEST INFRASTRUCTURE ONLY (never imported from ubu_b200/ or host/). Every function restates Java that EXISTS in the
reference, line by line, quirks included (file:line below; paths relative to src/main/java/edu/unc/bioinf/ubu/):

    combine / combine_chimeric_reads / prune_likely_inserts   chimera/CombineChimera3.java:21-73, 75-184, 251-267
    read comparator, mapped read start / end                  chimera/CombineChimera3.java:221-237, 269-296
    remove_soft_clips                                         assembly/ReAligner.java:926-961   (off-by-one at :942 and :953 kept)
    clean_and_output_contigs                                  assembly/ReAligner.java:876-921
    adjust_reads (gate, hit list, XA loop, de-dup, tags)      assembly/ReAligner.java:1048-1291 (XA loop bound :1137 kept)
    update_read_alignment                                     assembly/ReAligner.java:1293-1368 (via post_oracle.project_read)
    sam2fastq / sam2fastq_paired / sam_read_to_fastq          fastq/Sam2Fastq.java:33-79, 85-109, 111-162

Pinning: Sam2Fastq has a fixture in the reference (src/test/java/.../fastq/testdata/sam2fastq_test.sam with
sam2fastq_expected_paired{1,2}.fastq, Sam2FastqTest.java:35-44); tests/test_records.py checks this restatement against it.
The other functions have no test or fixture in the reference: they are restated from reading the
Java and labelled so.

SAM records are modelled after the few net.sf.samtools.SAMRecord behaviours the Java above relies on (samtools 1.54 is not
in this image; the behaviours are the documented ones: getAlignmentEnd = start + reference length - 1, 0 for unmapped
reads; Cigar.getReadLength counts M I S = X; getReadString of an empty sequence is "*"; getSAMString is the tab-joined
text line plus a newline; attributes keep insertion order, setAttribute(tag, null) removes).
"""
from __future__ import annotations

import copy
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import post_oracle as PO

OPS = "MIDNSHP=X"


class SamRecord:
    def __init__(self, line: str):
        f = line.rstrip("\n").split("\t")
        self.qname, self.flag, self.rname, self.pos, self.mapq = f[0], int(f[1]), f[2], int(f[3]), int(f[4])
        self.cigar = parse_cigar(f[5])
        self.rnext, self.pnext, self.tlen = (self.rname if f[6] == "=" else f[6]), int(f[7]), int(f[8])
        self.seq, self.qual = ("" if f[9] == "*" else f[9]), ("" if f[10] == "*" else f[10])
        self.tags: List[List[str]] = [t.split(":", 2) for t in f[11:]]          # [tag, type, value], file order

    # ---- the SAMRecord accessors the reference uses ----
    def unmapped(self) -> bool: return bool(self.flag & 0x4)
    def negative(self) -> bool: return bool(self.flag & 0x10)
    def first_of_pair(self) -> bool: return bool(self.flag & 0x40)
    def second_of_pair(self) -> bool: return bool(self.flag & 0x80)
    def read_string(self) -> str: return self.seq if self.seq else "*"
    def read_length(self) -> int: return len(self.seq)
    def cigar_string(self) -> str: return cigar_text(self.cigar)
    def cigar_length(self) -> int: return len(self.cigar)
    def cigar_read_length(self) -> int: return sum(n for n, o in self.cigar if OPS[o] in "MIS=X")
    def alignment_end(self) -> int:
        if self.unmapped():
            return 0
        return self.pos + sum(n for n, o in self.cigar if OPS[o] in "MDN=X") - 1

    def get_int(self, tag: str) -> Optional[int]:
        for t in self.tags:
            if t[0] == tag:
                return int(t[2])
        return None

    def get_attr(self, tag: str) -> Optional[str]:
        for t in self.tags:
            if t[0] == tag:
                return t[2]
        return None

    def set_attr(self, tag: str, value) -> None:
        for k, t in enumerate(self.tags):
            if t[0] == tag:
                if value is None:
                    del self.tags[k]
                else:
                    self.tags[k] = [tag, "i" if isinstance(value, int) else "Z", str(value)]
                return
        if value is not None:
            self.tags.append([tag, "i" if isinstance(value, int) else "Z", str(value)])

    def set_flag(self, bit: int, on: bool) -> None:
        self.flag = (self.flag | bit) if on else (self.flag & ~bit)

    def sam_string(self) -> str:
        rnext = "=" if (self.rnext == self.rname and self.rnext != "*") else self.rnext
        f = [self.qname, str(self.flag), self.rname, str(self.pos), str(self.mapq), self.cigar_string(), rnext, str(self.pnext),
             str(self.tlen), self.read_string(), self.qual if self.qual else "*"] + [":".join(t) for t in self.tags]
        return "\t".join(f) + "\n"

    def clone(self) -> "SamRecord":
        return copy.deepcopy(self)


def parse_cigar(text: str) -> List[Tuple[int, int]]:
    out, n = [], 0
    if text == "*":
        return out
    for ch in text:
        if ch.isdigit():
            n = n * 10 + ord(ch) - 48
        else:
            out.append((n, OPS.index(ch)))
            n = 0
    return out


def cigar_text(cigar: Sequence[Tuple[int, int]]) -> str:
    return "".join(f"{n}{OPS[o]}" for n, o in cigar) if cigar else "*"


def parse_sam(text: str) -> Tuple[List[str], List[SamRecord]]:
    header, recs = [], []
    for line in text.split("\n"):
        if not line:
            continue
        if line.startswith("@"):
            header.append(line)
        else:
            recs.append(SamRecord(line))
    return header, recs


def _read_blocks(r: SamRecord) -> List[PO.Block]:                 # ReadBlock.getReadBlocks (ReadBlock.java:114-132)
    return PO.read_blocks(r.pos, r.cigar)


# ---------------------------------------------------------------- CombineChimera3 -----------------------------------------
def mapped_read_start(r: SamRecord) -> int:                        # CombineChimera3.java:269-281
    for b in _read_blocks(r):
        if b.type != PO.S:
            return b.read_start
    return -1


def mapped_read_end(r: SamRecord) -> int:                          # CombineChimera3.java:283-296
    for b in reversed(_read_blocks(r)):
        if b.type != PO.S:
            return b.read_start + b.length - 1
    return -1


def sort_reads_by_position(reads: List[SamRecord]) -> None:        # CombineChimera3.java:221-237 (read2 - read1: DESCENDING; stable sort)
    import functools

    def cmp(a, b):
        c = mapped_read_start(b) - mapped_read_start(a)
        if c == 0:
            c = mapped_read_end(b) - mapped_read_end(a)
        return c
    reads.sort(key=functools.cmp_to_key(cmp))


def prune_likely_inserts(reads: List[SamRecord]) -> None:          # CombineChimera3.java:251-267
    if len(reads) == 3:
        first, middle, last = reads
        FUDGE_FACTOR = 5
        if mapped_read_start(middle) <= mapped_read_end(first) + FUDGE_FACTOR and mapped_read_end(middle) >= mapped_read_start(last) - FUDGE_FACTOR:
            reads.remove(middle)


def _total_length(elements) -> int:                                # CombineChimera3.java:203-219
    return sum(n for n, o in elements if o != PO.D)


def combine_chimeric_reads(read1: SamRecord, read2: SamRecord) -> Optional[SamRecord]:      # CombineChimera3.java:75-184
    left, right = (read1, read2) if read1.pos < read2.pos else (read2, read1)
    lc, rc = left.cigar, right.cigar
    if rc[0][1] == PO.S and lc[len(lc) - 1][1] == PO.S:
        left_el = list(lc[:-1])                                     # drop trailing S on the left side
        right_el = list(rc[1:])                                     # drop leading S on the right side
        left_blocks = _read_blocks(left)
        last_left = left_blocks[len(left_blocks) - 2]
        right_blocks = _read_blocks(right)
        first_right = right_blocks[1]
        left_stop = last_left.ref_start + last_left.length - 1      # getReferenceStop (ReadBlock.java:41-44)
        right_start = first_right.ref_start
        left_read_stop = last_left.read_start + last_left.length - 1
        right_read_start = first_right.read_start
        trim = 0
        if left_stop >= right_start or left_read_stop >= right_read_start:
            trim = max(left_stop - right_start + 1, left_read_stop - right_read_start + 1)
            n, o = left_el[-1]                                       # trimRightmostElement (:194-201)
            left_el[-1] = (n - trim, o)
            if n - trim < 1:
                return None
        left_alignment_stop = (last_left.ref_start + last_left.length - 1) - trim
        gap = first_right.ref_start - left_alignment_stop - 1
        if gap > 0:
            indel = (gap, PO.D)
        else:
            indel = (read1.read_length() - (_total_length(left_el) + _total_length(right_el)), PO.I)
        combined = read1.clone()
        combined.pos = left_blocks[0].ref_start
        combined.cigar = left_el + [indel] + right_el
        combined.mapq = min(read1.mapq, read2.mapq)
        return combined
    return None


def combine(sam_text: str) -> str:                                  # CombineChimera3.java:21-73
    header, recs = parse_sam(sam_text)
    out = []
    for h in header:                                                 # header.setSortOrder(unsorted) (:25)
        if h.startswith("@HD"):
            f = [x for x in h.split("\t") if not x.startswith("SO:")] + ["SO:unsorted"]
            h = "\t".join(f)
        out.append(h + "\n")
    k = 0
    while k < len(recs):                                             # SamMultiMappingReader: consecutive records of one name
        j = k
        while j < len(recs) and recs[j].qname == recs[k].qname:
            j += 1
        group = recs[k:j]
        k = j
        combined_flag = False
        sort_reads_by_position(group)
        prune_likely_inserts(group)
        if len(group) == 2:
            r1, r2 = group
            if r1.rname == r2.rname and r1.negative() == r2.negative() and r1.cigar_length() >= 2 and r2.cigar_length() >= 2:
                c = combine_chimeric_reads(r1, r2)
                if c is not None:
                    combined_flag = True
                    out.append(c.sam_string())
        if not combined_flag:
            for r in group:
                out.append(r.sam_string())
    return "".join(out)


# ---------------------------------------------------------------- ReAligner: contigs ------------------------------------
def remove_soft_clips(read: SamRecord) -> None:                     # ReAligner.java:926-961
    cig = read.cigar
    first, last = cig[0], cig[len(cig) - 1]
    if first[1] == PO.S or last[1] == PO.S:
        new = []
        bases = read.read_string()
        if first[1] == PO.S:
            bases = bases[first[0]:len(bases) - 1]                   # :942 substring(len, length-1): drops the last base too
        else:
            new.append(first)
        for i in range(1, len(cig) - 1):
            new.append(cig[i])
        if last[1] == PO.S:
            bases = bases[0:len(bases) - last[0] - 1]                # :953: one base more than the clip
        read.cigar = new
        read.seq = "" if bases == "*" else bases


def clean_and_output_contigs(sam_text: str, min_contig_length: int, should_remove_soft_clips: bool) -> Tuple[str, bool]:   # ReAligner.java:876-921
    _, recs = parse_sam(sam_text)
    out, has = [], False
    for r in recs:
        if r.mapq > 1:                                               # :887
            if should_remove_soft_clips:
                remove_soft_clips(r)
            bases = r.read_string()
            if len(bases) >= min_contig_length:                      # :895
                r.seq = ""                                           # setReadString("") -> "*"
                s = r.sam_string().replace("\t", "~")              # keeps the trailing newline of getSAMString (:904-905)
                out.append(">" + r.qname + "~" + s)                 # :907-909 (no newline written between header and bases)
                out.append(bases)
                out.append("\n")
                has = True
    return "".join(out), has


# ---------------------------------------------------------------- ReAligner.adjustReads ---------------------------------
def _int_attr(r: SamRecord, tag: str) -> int:                       # ReAligner.java:1028-1036
    v = r.get_int(tag)
    return 0 if v is None else v


def _edit_distance(r: SamRecord) -> int:                            # ReAligner.java:1038-1046
    v = r.get_int("NM")
    return r.read_length() if v is None else v


def _revcomp(s: str) -> str:                                        # sam/ReverseComplementor.java:38-62
    comp = {"A": "T", "T": "A", "C": "G", "G": "C"}
    return "".join(comp.get(c, c) for c in reversed(s))


def update_read_alignment(contig: SamRecord, orig: SamRecord, position: int) -> Optional[SamRecord]:     # ReAligner.java:1293-1368
    got = PO.project_read(contig.pos, contig.cigar, position, orig.read_length())
    if got is None:
        return None
    read = orig.clone()
    read.rname = contig.rname
    read.pos, read.cigar = got[0], list(got[1])
    return read


def adjust_reads(orig_sam: str, to_contig_sam: str, full_match_cigar: str = "100M") -> Tuple[str, Dict[str, int]]:
    """ReAligner.java:1048-1291. full_match_cigar is the literal the reference compares against ("100M", :1091,1145,1149);
    "" stands for "<read length>M" per read (the generalisation the 76 / 150 bp configs need; never the default).
    Returns (SAM records text without header, counters)."""
    _, origs = parse_sam(orig_sam)
    _, contigs = parse_sam(to_contig_sam)
    out: List[str] = []
    stats = {"missing_xa": 0, "mismatch_issue": 0, "adjusted": 0}
    ci, oi, cached = 0, 0, None
    while (ci < len(contigs) or cached is not None) and oi < len(origs):          # :1070
        orig = origs[oi]; oi += 1
        if cached is not None:
            read, cached = cached, None
        else:
            read = contigs[ci]; ci += 1
        written = False
        if orig.qname == read.qname:
            fm_in = full_match_cigar
            full_match_cigar = fm_in if fm_in else f"{read.read_length()}M"       # "" = the flagged generalisation "<read length>M"
            if read.cigar_string() == full_match_cigar and not read.unmapped() and _edit_distance(read) < _edit_distance(orig):     # :1091-1093
                best_hits = []
                s = read.rname
                s = s[s.index("~") + 1:] if "~" in s else s[0 + 0:]                # substring(indexOf('~')+1): indexOf -1 -> substring(0)
                contig_read = SamRecord(s.replace("~", "\t"))
                best_mm = _int_attr(read, "XM")
                if read.alignment_end() <= contig_read.cigar_read_length():        # :1112
                    best_hits.append((contig_read, read.pos, "-" if read.negative() else "+", best_mm))
                n_best, n_sub = _int_attr(read, "X0"), _int_attr(read, "X1")
                total_hits = n_best + n_sub
                if 1 < total_hits < 1000:                                          # :1126
                    xa = read.get_attr("XA")
                    if xa is None:
                        stats["missing_xa"] += 1
                    else:
                        alternates = _java_split(xa, ";")
                        for i in range(len(alternates) - 1):                        # :1137 (skips the last entry)
                            info = alternates[i].split(",")
                            alt, strand, position, cigar, mm = info[0], info[1][0], int(info[1][1:]), info[2], int(info[3])
                            if cigar == full_match_cigar and mm < best_mm:
                                stats["mismatch_issue"] += 1
                            if cigar == full_match_cigar and mm == best_mm:
                                a = alt[alt.index("~") + 1:] if "~" in alt else alt
                                cr = SamRecord(a.replace("~", "\t"))
                                if position + read.read_length() <= cr.cigar_read_length():     # :1156
                                    best_hits.append((cr, position, strand, mm))
                info_map: Dict[str, SamRecord] = {}
                for cr, pos1, strand, mm in best_hits:
                    position = pos1 - 1
                    if cr.mapq > orig.mapq:                                        # :1176
                        upd = update_read_alignment(cr, orig, position)
                        if upd is not None:
                            if upd.mapq == 0:
                                upd.mapq = 1
                            if upd.unmapped():
                                upd.set_flag(0x4, False)
                            upd.set_flag(0x10, strand == "-")
                            if (strand == "-") == read.negative():                  # :1200-1205
                                upd.seq, upd.qual = read.seq, read.qual
                            else:
                                upd.seq, upd.qual = _revcomp(read.seq), read.qual[::-1]
                            if (orig.unmapped() or orig.rname != upd.rname or orig.pos != upd.pos or orig.negative() != upd.negative()
                                    or orig.cigar_string() != upd.cigar_string()):   # :1208-1212
                                if orig.unmapped():
                                    oa = "N/A"
                                else:
                                    oa = f"{orig.rname}:{orig.pos}:{'-' if orig.negative() else '+'}:{orig.cigar_string()}"
                                upd.set_attr("YO", oa)
                            upd.set_attr("YM", mm)
                            upd.set_attr("YQ", cr.mapq)
                            key = f"{upd.rname}_{upd.pos}_{'-' if upd.negative() else '+'}_{upd.cigar_string()}"
                            if key not in info_map:
                                info_map[key] = upd
                for r in _hashmap_values(info_map):                                # HashMap.values(): Java's bucket order (:1241)
                    ob, osub = _int_attr(r, "X0"), _int_attr(r, "X1")
                    if len(info_map) > 1 or total_hits > 1000:
                        r.mapq = 0
                    if r.get_attr("YO") is not None:
                        r.set_attr("X0", len(info_map))
                        r.set_attr("X1", ob + osub)
                        for t in ("XO", "XG", "MD", "XA", "XT"):
                            r.set_attr(t, None)
                    out.append(r.sam_string())
                    written = True
                    stats["adjusted"] += 1
            full_match_cigar = fm_in
        else:
            cached = read
        if not written:
            out.append(orig.sam_string())
    return "".join(out), stats


def _java_split(s: str, sep: str) -> List[str]:
    """String.split(sep): trailing empty strings are removed (so "a;b;" -> ["a", "b"])."""
    parts = s.split(sep)
    while parts and parts[-1] == "":
        parts.pop()
    return parts


def _java_string_hash(s: str) -> int:
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h


def _hashmap_values(m: Dict[str, SamRecord]) -> List[SamRecord]:
    """Iteration order of java.util.HashMap<String, ...> (Java 6/7: table of 16 doubling at load factor 0.75, supplemental hash
    h ^= (h >>> 20) ^ (h >>> 12); h ^ (h >>> 7) ^ (h >>> 4), buckets walked from index 0 up, a bucket's entries newest first).
    With one entry -- the overwhelmingly common case -- the order is trivially that entry."""
    keys = list(m.keys())
    if len(keys) <= 1:
        return [m[k] for k in keys]
    cap = 16
    while len(keys) > cap * 0.75:
        cap *= 2

    def spread(h):
        h ^= (h >> 20) ^ (h >> 12)
        return (h ^ (h >> 7) ^ (h >> 4)) & 0xFFFFFFFF
    buckets: Dict[int, List[str]] = {}
    for k in keys:                                                   # insertion order; a bucket's chain has the newest entry first
        buckets.setdefault(spread(_java_string_hash(k)) & (cap - 1), []).insert(0, k)
    return [m[k] for b in sorted(buckets) for k in buckets[b]]


# ---------------------------------------------------------------- Sam2Fastq ---------------------------------------------
def sam_read_to_fastq(r: SamRecord) -> List[str]:                   # Sam2Fastq.java:111-162 (the non-fusion branch)
    bases, quals = r.read_string(), (r.qual if r.qual else "*")
    if r.negative():
        bases, quals = _revcomp(bases), quals[::-1]
    return ["@" + r.qname, bases, "+", quals]                       # FastqRecord(id, sequence, quality): 4 lines, '+' alone


def sam2fastq(sam_text: str) -> str:                                # Sam2Fastq.java:85-109
    _, recs = parse_sam(sam_text)
    out, last = [], ""
    for r in recs:
        if r.qname != last:
            out.extend(sam_read_to_fastq(r))
            last = r.qname
    return "".join(l + "\n" for l in out)


def sam2fastq_paired(sam_text: str, end1: Optional[str] = None, end2: Optional[str] = None) -> Tuple[str, str]:   # Sam2Fastq.java:33-79
    _, recs = parse_sam(sam_text)
    o1, o2, last1, last2, n1, n2 = [], [], "", "", 0, 0
    for r in recs:
        is1 = r.qname.endswith(end1) if end1 is not None else r.first_of_pair()      # :164-188
        is2 = r.qname.endswith(end2) if end2 is not None else r.second_of_pair()
        if is1:
            if r.qname != last1:
                o1.extend(sam_read_to_fastq(r)); last1 = r.qname; n1 += 1
        elif is2:
            if r.qname != last2:
                o2.extend(sam_read_to_fastq(r)); last2 = r.qname; n2 += 1
    if n1 != n2:
        raise ValueError("Non-symmetrical read counts found")                         # :76-78 IllegalStateException
    return "".join(l + "\n" for l in o1), "".join(l + "\n" for l in o2)
