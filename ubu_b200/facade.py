"""Host-side mirror of the reference's aligner facade: same constructor, same method names, path in / path out.

    reference:  new Aligner(String reference, int numThreads)       src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:13-16
                void shortAlign(String input, String outputSam)     Aligner.java:46-56   (bwa aln + bwa samse -n 1000)
                void align(String input, String outputSam)          Aligner.java:18-22   (bwa bwasw)
                void index()                                        Aligner.java:62-69   (bwa index)

The reference aligns every read in `input` against the FASTA it was constructed with and writes SAM. This class does
the same with window Smith-Waterman on the GPU: every read is aligned against every sequence of the FASTA on both
strands (bwa searches both strands too), the best hit becomes the record, equally good hits go to `XA`. The records
carry what the consumer reads back (ReAligner.adjustReads, ReAligner.java:1088-1164): QNAME, FLAG (unmapped 0x4,
strand 0x10), RNAME, POS, CIGAR string, SEQ/QUAL (reverse-complemented on the minus strand, as bwa writes them), and the
tags NM, XM, X0, X1, XA (`name,±pos,CIGAR,NM;` per entry, ReAligner.java:1135-1143).

BUILD-DEFINED parts (the reference delegates them to bwa and pins none of them): "best" = highest SPEC score, ties broken
by reference order then plus strand first; X0 = number of hits with the best score; X1 = number of hits scoring less
than the best but within one mismatch of it (score >= best - (match - mismatch)); MAPQ = 37 for a unique best hit
without close competitors, 23 with them, 0 when the best score is shared (the values bwa aln typically prints).

The alignment itself runs in libubu_sw.so; everything here is text I/O and bookkeeping.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from .aligner import Alignments, SwAligner
from .seqs import SeqSet, codes_from_text, pack_codes, reverse_complement

MAX_XA = 1000          # `bwa samse -n 1000` (Aligner.java:53)


@dataclass
class FastxRecord:
    name: str
    seq: str
    qual: Optional[str] = None


def read_fastx(path: str) -> Iterator[FastxRecord]:
    """FASTA or FASTQ (4-line records, as the reference's FastqInputFile reads them: fastq/FastqInputFile.java)."""
    with open(path) as f:
        first = f.readline()
        if not first:
            return
        if first.startswith(">"):
            name, chunks = first[1:].strip(), []
            for line in f:
                if line.startswith(">"):
                    yield FastxRecord(name.split()[0] if name else "", "".join(chunks))
                    name, chunks = line[1:].strip(), []
                else:
                    chunks.append(line.strip())
            yield FastxRecord(name.split()[0] if name else "", "".join(chunks))
        elif first.startswith("@"):
            head = first
            while head:
                seq = f.readline().rstrip("\n")
                f.readline()
                qual = f.readline().rstrip("\n")
                name = head[1:].strip().split()[0]
                yield FastxRecord(name, seq, qual)
                head = f.readline()
        else:
            raise ValueError(f"{path}: neither FASTA nor FASTQ")


def cigar_text(ops: Sequence[Tuple[int, int]]) -> str:
    return "".join(f"{n}{'MIDNSHP=X'[o]}" for n, o in ops) if ops else "*"


class Aligner:
    """Drop-in for edu.unc.bioinf.ubu.assembly.Aligner. `align_fn` lets tests substitute the device call."""

    def __init__(self, reference: str, numThreads: int = 1, device: int = 0, match: int = 1, mismatch: int = -3,
                 gap_open: int = 5, gap_ext: int = 2, align_fn: Optional[Callable[[SeqSet, SeqSet, np.ndarray], Alignments]] = None,
                 batch_tasks: int = 2_000_000, seed_k: int = 0):
        self.reference, self.numThreads = reference, numThreads        # Aligner.java:9-16 (numThreads unused: one GPU per handle)
        self.scoring = (match, mismatch, gap_open, gap_ext)
        self._device, self._align_fn, self._sw = device, align_fn, None
        self._seeds = None
        self._batch = batch_tasks
        # BUILD-DEFINED scalability switch (mirrors ubu::Aligner::setSeedLength): k > 0 aligns a read (either strand) only against the
        # reference sequences it shares an exact k-mer with; 0 = every read x every sequence x both strands
        self.seed_k = max(0, min(int(seed_k), 31))

    # -- reference API ---------------------------------------------------------------------------
    def index(self) -> None:
        """Aligner.java:62-69 builds a BWT index with bwa; window Smith-Waterman needs none. Kept so callers run unchanged;
        fails like the reference would if the FASTA is unreadable."""
        if not os.path.exists(self.reference):
            raise RuntimeError(f"reference not found: {self.reference}")

    def shortAlign(self, input: str, outputSam: str) -> None:          # noqa: N802 (reference method name)
        self._run(input, outputSam)

    def align(self, input: str, outputSam: str) -> None:
        self._run(input, outputSam)

    # -- implementation --------------------------------------------------------------------------
    def _device_align(self, reads: SeqSet, wins: SeqSet, idx: np.ndarray) -> Alignments:
        if self._align_fn is not None:
            return self._align_fn(reads, wins, idx)
        if self._sw is None:
            m, x, o, e = self.scoring
            self._sw = SwAligner(m, x, o, e, device=self._device)
        return self._sw.align(reads, wins, idx)

    def close(self) -> None:
        if self._sw is not None:
            self._sw.close()
            self._sw = None

    def _run(self, input_path: str, output_sam: str) -> None:
        try:
            refs = list(read_fastx(self.reference))
            reads = list(read_fastx(input_path))
            with open(output_sam, "w") as out:
                out.write("@HD\tVN:1.4\tSO:unsorted\n")
                for r in refs:
                    out.write(f"@SQ\tSN:{r.name}\tLN:{len(r.seq)}\n")
                out.write("@PG\tID:ubu_b200\tPN:ubu_b200\tDS:window Smith-Waterman on B200 (SPEC.md); X0 X1 XA MAPQ rules are BUILD-DEFINED (best = highest score, X1 = within one mismatch of it, MAPQ 37/23/0), not bwa's"
                          + (f"; candidates = sequences sharing an exact {self.seed_k}-mer with the read" if self.seed_k > 0 else "") + "\n")
                if not reads:
                    return
                wins = pack_codes([codes_from_text(r.seq) for r in refs]) if refs else None
                per_batch = max(1, self._batch // max(1, 2 * len(refs))) if self.seed_k == 0 else 500_000
                self._seeds = self._seed_index(refs) if self.seed_k > 0 else None
                for lo in range(0, len(reads), per_batch):
                    self._align_and_write(reads[lo:lo + per_batch], refs, wins, out)
        except (OSError, ValueError) as e:                               # the reference surfaces every failure as RuntimeException
            raise RuntimeError(f"aligner failed: {e}") from e            # (Aligner.java:41-43)

    def _align_and_write(self, reads: List[FastxRecord], refs: List[FastxRecord], wins: Optional[SeqSet], out) -> None:
        nr, nw = len(reads), len(refs)
        if nw == 0:
            for rd in reads:
                out.write(self._unmapped(rd))
            return
        # read entries: [fwd_0, rc_0, fwd_1, rc_1, ...]; tasks: every (entry, window) pair, sharing the packed bytes
        texts = []
        for rd in reads:
            texts.append(rd.seq)
            texts.append(reverse_complement(rd.seq))
        base = pack_codes([codes_from_text(t) for t in texts])
        # tasks: (entry, window) pairs -- all of them, or the seeded candidates (ascending window order per entry)
        cand = [list(range(nw)) if self._seeds is None else self._candidates(t) for t in texts]
        ent = np.repeat(np.arange(2 * nr), [len(c) for c in cand])       # task -> read entry
        widx = np.array([w for c in cand for w in c], dtype=np.uint32)   # task -> window
        if widx.size == 0:
            for rd in reads:
                out.write(self._unmapped(rd))
            return
        first = np.concatenate([[0], np.cumsum([len(c) for c in cand])])
        tasks = SeqSet(base.packed, base.off[ent], base.len[ent], base.nmask, None if base.nmask is None else base.nmask_off[ent])
        al = self._device_align(tasks, wins, widx)
        score = al.results["score"].astype(np.int64)
        m, x = self.scoring[0], self.scoring[1]
        for k, rd in enumerate(reads):
            # this read's tasks as (window, strand) -> task, reference order, plus strand first
            task_of = {}
            for s in (0, 1):
                for j, w in enumerate(cand[2 * k + s]):
                    task_of[(w, s)] = int(first[2 * k + s]) + j
            order = sorted(task_of)
            best = max((int(score[task_of[ws]]) for ws in order), default=0)
            if best <= 0:
                out.write(self._unmapped(rd))
                continue
            best_hits = [ws for ws in order if score[task_of[ws]] == best]
            x1 = sum(1 for ws in order if best - (m - x) <= score[task_of[ws]] < best)
            w0, s0 = best_hits[0]
            t0 = task_of[(w0, s0)]
            r0 = al.results[t0]
            cig = cigar_text(al.cigar(t0))
            nm = int(r0["edit_distance"])
            xm = nm - sum(n for n, o in al.cigar(t0) if o in (1, 2))
            mapq = 0 if len(best_hits) > 1 else (23 if x1 else 37)
            seq = rd.seq if s0 == 0 else reverse_complement(rd.seq)
            qual = (rd.qual if rd.qual else "*")
            if s0 == 1 and rd.qual:
                qual = rd.qual[::-1]
            tags = [f"NM:i:{nm}", f"XM:i:{xm}", f"X0:i:{len(best_hits)}", f"X1:i:{x1}", f"AS:i:{best}"]
            if len(best_hits) > 1:
                xa = []
                for (w, s) in best_hits[1:MAX_XA + 1]:
                    t = task_of[(w, s)]
                    rr = al.results[t]
                    xa.append(f"{refs[w].name},{'-' if s else '+'}{int(rr['ref_start'])},{cigar_text(al.cigar(t))},{int(rr['edit_distance'])};")
                tags.append("XA:Z:" + "".join(xa))
            flag = 16 if s0 else 0
            out.write("\t".join([rd.name, str(flag), refs[w0].name, str(int(r0["ref_start"])), str(mapq), cig, "*", "0", "0", seq, qual] + tags) + "\n")

    # -- exact k-mer seeds (k-mers holding a non-ACGT letter are skipped) ---------------------------
    def _kmers(self, text: str):
        k, t = self.seed_k, text.upper()
        for i in range(len(t) - k + 1):
            km = t[i:i + k]
            if all(c in "ACGT" for c in km):
                yield km

    def _seed_index(self, refs: List[FastxRecord]) -> dict:
        idx: dict = {}
        for w, r in enumerate(refs):
            for km in self._kmers(r.seq):
                lst = idx.setdefault(km, [])
                if not lst or lst[-1] != w:
                    lst.append(w)
        return idx

    def _candidates(self, text: str) -> List[int]:
        out = set()
        for km in self._kmers(text):
            out.update(self._seeds.get(km, ()))
        return sorted(out)

    @staticmethod
    def _unmapped(rd: FastxRecord) -> str:
        return "\t".join([rd.name, "4", "*", "0", "0", "*", "*", "0", "0", rd.seq, rd.qual if rd.qual else "*"]) + "\n"
