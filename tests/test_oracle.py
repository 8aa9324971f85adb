"""CPU tests of the checker itself: the C oracle vs the independent Python restatement vs the committed
golden vectors, plus SPEC-level properties. No GPU, no /root/reference at run time."""
import json
import os
import random

import numpy as np
import pytest

import oracle_lib as O
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import sw_oracle as P  # noqa: E402

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sw_golden_v1.json")))
FIELDS = ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance")


def c_align(read, win, params):
    q, r = np.array(P.encode(read), dtype=np.uint8), np.array(P.encode(win), dtype=np.uint8)
    res, cig = O.align_one(q, r, tuple(params))
    return res, cig, q, r


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_golden_c_oracle(case):
    res, cig, _, _ = c_align(case["read"], case["window"], case["params"])
    e = case["expect"]
    for f in FIELDS:
        assert int(res[f]) == e[f], f
    assert (int(res["flags"]) & 1) == (e["flags"] & 1)
    assert ("".join(f"{n}{'MIDNS'[o]}" for n, o in cig) or "*") == e["cigar"]


@pytest.mark.parametrize("case", GOLD["cases"][:40], ids=[c["name"] for c in GOLD["cases"][:40]])
def test_golden_python_restatement(case):
    pr = P.align(P.encode(case["read"]), P.encode(case["window"]), P.Params(*case["params"]))
    e = case["expect"]
    assert {f: getattr(pr, f) for f in FIELDS} == {f: e[f] for f in FIELDS}
    assert pr.cigar_string() == e["cigar"] and (pr.flags & 1) == (e["flags"] & 1)


def _rand_pair(rng):
    kind = rng.randint(0, 4)
    if kind == 0:
        b = rng.choice("ACGT")
        return b * rng.randint(1, 40), b * rng.randint(1, 60)
    if kind == 1:
        u = "".join(rng.choice("ACGT") for _ in range(rng.randint(2, 5)))
        return (u * 20)[: rng.randint(4, 50)], (u * 40)[: rng.randint(4, 90)]
    L, W = rng.randint(1, 60), rng.randint(1, 90)
    alpha = "ACGT" if kind < 4 else "ACGTN"
    w = "".join(rng.choice(alpha) for _ in range(W))
    if kind == 2 and W > 4:
        s = rng.randint(0, W - 2)
        r = list(w[s: s + L])
        for k in range(len(r)):
            if rng.random() < 0.08:
                r[k] = rng.choice("ACGT")
        if r and rng.random() < 0.5:
            del r[rng.randrange(len(r))]
        return "".join(r) or "A", w
    return "".join(rng.choice(alpha) for _ in range(L)), w


@pytest.mark.parametrize("params", [(1, -3, 5, 2), (2, -1, 3, 1), (1, -1, 0, 1), (4, -6, 2, 3)])
def test_c_equals_python_on_random_and_adversarial(params):
    rng = random.Random(hash(params) & 0xFFFF)
    for _ in range(150):
        read, win = _rand_pair(rng)
        res, cig, q, r = c_align(read, win, params)
        pr = P.align(list(q), list(r), P.Params(*params))
        assert {f: int(res[f]) for f in FIELDS} == {f: getattr(pr, f) for f in FIELDS}, (read, win)
        assert cig == pr.cigar, (read, win)


def test_properties_rescore_and_shape():
    rng = random.Random(77)
    for _ in range(300):
        read, win = _rand_pair(rng)
        params = rng.choice([(1, -3, 5, 2), (2, -2, 1, 1)])
        res, cig, q, r = c_align(read, win, params)
        if int(res["flags"]) & 1:
            assert int(res["score"]) == 0 and not cig
            continue
        rec = {f: int(res[f]) for f in FIELDS}
        sc, ed = O.rescore(q, r, rec, cig, params)
        assert sc == rec["score"] and ed == rec["edit_distance"]          # CIGAR re-scored under SPEC == reported score
        ops = [o for _, o in cig if o != 4]
        assert ops[0] == 0 and ops[-1] == 0                                 # aligned part begins and ends with M
        assert sum(n for n, o in cig if o in (0, 1, 4)) == len(q)           # CIGAR covers the whole read
        assert sum(n for n, o in cig if o in (0, 2)) == rec["ref_end"] - rec["ref_start"] + 1
        assert all(a[1] != b[1] for a, b in zip(cig, cig[1:]))              # adjacent ops merged


def test_score_symmetric_under_reversal():
    rng = random.Random(5)
    for _ in range(100):
        read, win = _rand_pair(rng)
        a, _, _, _ = c_align(read, win, (1, -3, 5, 2))
        b, _, _, _ = c_align(read[::-1], win[::-1], (1, -3, 5, 2))
        assert int(a["score"]) == int(b["score"])


def test_reference_sequence_kats():
    """Vectors copied from the reference's own test (ReverseComplementorTest.java:22,30,40,48) and from the strings
    Sequence.main round-trips (Sequence.java:127-131)."""
    k = GOLD["reference_kats"]
    import ctypes as C
    l = O.lib()
    src = k["revcomp_input"].encode()
    out = C.create_string_buffer(len(src))
    l.swo_revcomp_text(src, len(src), out)
    assert out.raw.decode() == k["reverse_complement"]
    assert P.revcomp_text(k["revcomp_input"]) == k["reverse_complement"]
    assert k["revcomp_input"][::-1] == k["reverse"]
    comp = {"A": "T", "T": "A", "C": "G", "G": "C"}
    assert "".join(comp.get(c, c) for c in k["revcomp_input"]) == k["complement"]
    for s in k["sequence_roundtrip"]:
        packed, nmask = P.pack2(P.encode(s))
        assert P.decode(P.unpack2(packed, nmask, len(s))) == s
        codes = np.array(P.encode(s), dtype=np.uint8)
        pk = np.zeros((len(s) + 3) // 4, dtype=np.uint8); nm = np.zeros((len(s) + 7) // 8, dtype=np.uint8)
        l.swo_pack2(codes.ctypes.data, len(s), pk.ctypes.data, nm.ctypes.data)
        assert bytes(pk) == packed
    assert P.pack2(P.encode("ATCGATCG"))[0].hex() == k["packed_ATCGATCG_hex"]
    assert {c: P.encode(c)[0] for c in "ACTG"} == k["two_bit_code"]


def test_batch_threads_agree():
    rng = random.Random(3)
    pairs = [_rand_pair(rng) for _ in range(64)]
    q = [np.array(P.encode(a), dtype=np.uint8) for a, _ in pairs]
    r = [np.array(P.encode(b), dtype=np.uint8) for _, b in pairs]
    r1, c1 = O.align_batch(q, r, None, nthreads=1)
    r4, c4 = O.align_batch(q, r, None, nthreads=4)
    assert np.array_equal(r1, r4) and c1 == c4
