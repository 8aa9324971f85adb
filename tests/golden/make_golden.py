#!/usr/bin/env python
"""Generates tests/golden/sw_golden_v1.json.

There are NO alignment golden vectors in the reference (mozack/ubu has no Smith-Waterman and no test that
pins an alignment), so these vectors pin THIS repo's SPEC.md: each case is produced by
the pure-Python restatement (oracle/sw_oracle.py), cross-checked here against the C oracle
(oracle/sw_oracle.c), and committed. The sequence-level vectors in reference_kats come verbatim from the
reference's own test: src/test/java/edu/unc/bioinf/ubu/sam/ReverseComplementorTest.java:22,30,40,48, plus the
strings Sequence.main round-trips (src/main/java/edu/unc/bioinf/ubu/assembly/Sequence.java:127-131).

    python tests/golden/make_golden.py        # rewrites the JSON next to this script
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

import oracle_lib as O  # noqa: E402
import sw_oracle as P  # noqa: E402

DEFAULT = (1, -3, 5, 2)


def rnd(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def mutate(rng, s, sub=0.03, indel=0.02):
    out = []
    for ch in s:
        x = rng.random()
        if x < sub:
            out.append(rng.choice([c for c in "ACGT" if c != ch]))
        elif x < sub + indel / 2:
            continue
        elif x < sub + indel:
            out.append(ch)
            out.append(rnd(rng, rng.randint(1, 3)))
        else:
            out.append(ch)
    return "".join(out)


def cases():
    rng = random.Random(20261019)
    c = []
    # hand-made: exact, clipped, one mismatch, one insertion, one deletion, N, ties, empties
    w = "TTGACCATAGGCTAACGTTAGCCATGGATCCGATTACAGGCATCGA"
    c += [("exact", w[10:30], w, DEFAULT), ("clip_both", "GGGG" + w[8:28] + "CCCC", w, DEFAULT),
          ("one_mismatch", w[5:17] + "A" + w[18:32], w, DEFAULT),
          ("one_insertion", w[5:20] + "TT" + w[20:40], w, DEFAULT),
          ("one_deletion", w[5:20] + w[23:44], w, DEFAULT),
          ("read_n", w[5:15] + "N" + w[16:30], w, DEFAULT), ("window_n", w[5:30], w[:18] + "N" + w[19:], DEFAULT),
          ("all_n_read", "NNNNNNNN", w, DEFAULT), ("all_n_window", w[3:20], "N" * 30, DEFAULT),
          ("empty_read", "", w, DEFAULT), ("empty_window", w[3:9], "", DEFAULT), ("single_base_hit", "G", "ACGT", DEFAULT),
          ("single_base_miss", "G", "ACAT", DEFAULT), ("lower_case", w[10:30].lower(), w, DEFAULT),
          ("iupac_as_n", w[10:18] + "R" + w[19:30], w, DEFAULT),
          ("homopolymer", "A" * 30, "A" * 50, DEFAULT), ("homopolymer_short_window", "A" * 30, "A" * 7, DEFAULT),
          ("tandem2", "ACACACACACACAC", "GG" + "AC" * 20 + "GG", DEFAULT), ("tandem3", "ACGACGACGACG", "ACG" * 15, DEFAULT),
          ("two_equal_hits", "GATTACAGATTACA", "CC" + "GATTACAGATTACA" + "TTT" + "GATTACAGATTACA" + "CC", DEFAULT),
          ("gap_vs_mismatch_tie", "ACGTACGTTTACGTACGT", "ACGTACGTACGTACGT", (1, -1, 0, 1)),
          ("ins_vs_del_tie", "AAAACCCCGGGG", "AAAACCCGGGG", (2, -2, 1, 1)),
          ("read_longer_than_window", rnd(rng, 10) + w[4:24] + rnd(rng, 30), w[4:24], DEFAULT)]
    # seeded random
    for k in range(40):
        W = rng.randint(30, 400)
        win = rnd(rng, W)
        L = rng.randint(10, min(250, W))
        st = rng.randint(0, W - L)
        read = mutate(rng, win[st:st + L], sub=rng.choice([0.0, 0.01, 0.05]), indel=rng.choice([0.0, 0.01, 0.04]))
        if len(read) > 250:
            read = read[:250]
        params = rng.choice([DEFAULT, DEFAULT, (2, -1, 3, 1), (1, -1, 0, 1), (5, -4, 10, 1), (3, 0, 2, 2)])
        c.append((f"random{k:02d}", read, win, params))
    for k in range(8):
        c.append((f"unrelated{k}", rnd(rng, rng.randint(5, 120)), rnd(rng, rng.randint(5, 200)), DEFAULT))
    return c


def main():
    out = {"spec": "SPEC.md (BUILD-DEFINED; reference parity UNPINNED)", "cases": [], "reference_kats": {
        "revcomp_input": "CTCTNCTGATCCCCACCTCCAAATATCTCATCAACAACCAACTAATCACC",
        "reverse": "CCACTAATCAACCAACAACTACTCTATAAACCTCCACCCCTAGTCNTCTC",
        "complement": "GAGANGACTAGGGGTGGAGGTTTATAGAGTAGTTGTTGGTTGATTAGTGG",
        "reverse_complement": "GGTGATTAGTTGGTTGTTGATGAGATATTTGGAGGTGGGGATCAGNAGAG",
        "sequence_roundtrip": ["ATCGATCG", "ATCGATCC", "ATCGATCGG", "TCGATCGG"],
        "two_bit_code": {"A": 0, "C": 1, "T": 2, "G": 3},
        "packed_ATCGATCG_hex": "d8d8"}}
    for name, read, win, params in cases():
        q, r = P.encode(read), P.encode(win)
        pr = P.align(q, r, P.Params(*params))
        cres, ccig = O.align_one(np.array(q, dtype=np.uint8), np.array(r, dtype=np.uint8), params)
        rec = {"score": pr.score, "ref_start": pr.ref_start, "ref_end": pr.ref_end, "q_start": pr.q_start, "q_end": pr.q_end,
               "edit_distance": pr.edit_distance, "flags": pr.flags, "cigar": pr.cigar_string()}
        for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance"):
            assert int(cres[f]) == rec[f], (name, f, int(cres[f]), rec[f])
        assert ccig == pr.cigar, (name, ccig, pr.cigar)
        out["cases"].append({"name": name, "read": read, "window": win, "params": list(params), "expect": rec})
    with open(os.path.join(HERE, "sw_golden_v1.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {len(out['cases'])} cases")


if __name__ == "__main__":
    main()
