#!/usr/bin/env python
"""Writes tests/golden/assembler_input1.json from the reference's own assembler fixture (run in the build container, where
/root/reference is mounted; the GPU box does not have it):
    reads     <- src/test/java/edu/unc/bioinf/ubu/assembly/input1.fastq  (sequence lines)
    expected  <- src/test/java/edu/unc/bioinf/ubu/assembly/output1.fasta (sequence line), the contig the commented-out
                 AssemblerTest.testBasicAssembly (AssemblerTest.java:22-44) asserts with setKmerSize(3), setMinNodeFrequncy(2),
                 setMinContigLength(4).
"""
import json
import os

REF = "/root/reference/src/test/java/edu/unc/bioinf/ubu/assembly"
lines = open(os.path.join(REF, "input1.fastq")).read().split("\n")
reads = [lines[i + 1].strip() for i in range(0, len(lines) - 1, 4) if lines[i].startswith("@")]
fasta = [l.strip() for l in open(os.path.join(REF, "output1.fasta")) if l.strip() and not l.startswith(">")]
out = {"source": "mozack/ubu src/test/java/edu/unc/bioinf/ubu/assembly/{input1.fastq,output1.fasta}; AssemblerTest.java:22-44",
       "kmer_size": 3, "min_node_frequency": 2, "min_contig_length": 4, "reads": reads, "expected_contigs": fasta}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "assembler_input1.json"), "w"), indent=1)
print(out)
