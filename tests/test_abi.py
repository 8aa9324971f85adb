"""CPU tests of the boundary: libubu_sw.so loads, exports every symbol include/ubu_sw.h declares, and its
host-only helpers behave; no compute call is made (there is no GPU here and no CPU fallback in the library)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ubu_b200 as U
from ubu_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "ubu_sw.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ubu_sw_[a-z_0-9]+)\s*\(", hdr)))


def test_header_symbols_all_exported_and_bound():
    lib = N.lib()
    names = declared_symbols()
    assert len(names) >= 18
    bound = {n for n, _, _ in N.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ubu_sw.h but not exported"
        assert n in bound, f"{n} declared but missing from the ctypes binding"


def test_struct_layouts_match_header():
    assert C.sizeof(N.Result) == 32 and U.RESULT_DTYPE.itemsize == 32
    assert C.sizeof(N.Params) == 16
    assert N.Seqs.count.offset == 40 and C.sizeof(N.Seqs) == 56


def test_version_and_errors():
    lib = N.lib()
    assert lib.ubu_sw_version() == 0x00010000
    assert lib.ubu_sw_strerror(0) == b"ok"
    assert b"CIGAR" in lib.ubu_sw_strerror(N.ERR_CIGAR_POOL)
    assert lib.ubu_sw_strerror(-99) == b"unknown status"


def test_create_rejects_bad_arguments_without_touching_a_gpu():
    lib = N.lib()
    h = C.c_void_p()
    bad = N.Params(0, -3, 5, 2)
    assert lib.ubu_sw_create(C.byref(bad), 0, 2, C.byref(h)) == N.ERR_ARG
    bad = N.Params(1, 1, 5, 2)
    assert lib.ubu_sw_create(C.byref(bad), 0, 2, C.byref(h)) == N.ERR_ARG
    ok = N.Params(1, -3, 5, 2)
    assert lib.ubu_sw_create(C.byref(ok), 0, 9, C.byref(h)) == N.ERR_ARG
    assert lib.ubu_sw_create(C.byref(ok), 0, 2, None) == N.ERR_ARG


@pytest.mark.skipif(N.lib().ubu_sw_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_product_path_fails_loudly():
    with pytest.raises(U.UbuSwError) as e:
        U.SwAligner()
    assert e.value.status == N.ERR_NO_DEVICE


def test_pack_helper_matches_numpy_packer_and_reference_convention():
    lib = N.lib()
    text = b"ACGTNacgtRYACGTTTGACA"
    packed = np.zeros((len(text) + 3) // 4, dtype=np.uint8)
    nmask = np.zeros((len(text) + 7) // 8, dtype=np.uint8)
    has_n = C.c_int(0)
    lib.ubu_sw_pack(text, len(text), packed.ctypes.data, nmask.ctypes.data, C.byref(has_n))
    s = U.pack_texts([text])
    assert has_n.value == 1
    assert bytes(packed) == bytes(s.packed[: packed.size]) and bytes(nmask) == bytes(s.nmask[: nmask.size])
    assert s.text(0) == "ACGTNACGTNNACGTTTGACA"
    # Sequence.java:10-13,31-41: A=0 C=1 T=2 G=3, LSB first -> "ATCG" = 0b11_01_10_00
    one = U.pack_texts(["ATCG"])
    assert int(one.packed[0]) == 0b11011000 and one.nmask is None


def test_revcomp_and_cigar_string_helpers():
    lib = N.lib()
    # reference vector: ReverseComplementorTest.java:22,48
    assert U.reverse_complement("CTCTNCTGATCCCCACCTCCAAATATCTCATCAACAACCAACTAATCACC") == \
        "GGTGATTAGTTGGTTGTTGATGAGATATTTGGAGGTGGGGATCAGNAGAG"
    assert U.reverse_complement("acgtN") == "Ntgca"          # only upper-case ACGT are complemented (ReverseComplementor.java:15-22)
    ops = np.array([(4 << 4) | 4, (30 << 4) | 0, (2 << 4) | 1, (40 << 4) | 0, (3 << 4) | 2, (1 << 4) | 0], dtype=np.uint32)
    buf = C.create_string_buffer(64)
    n = lib.ubu_sw_cigar_string(ops.ctypes.data, len(ops), buf, 64)
    assert buf.value == b"4S30M2I40M3D1M" and n == 14
    assert lib.ubu_sw_cigar_string(None, 0, buf, 64) == 1 and buf.value == b"*"


def test_seqset_roundtrip_and_uniform_packer():
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 4, size=(17, 77), dtype=np.uint8)
    a, b = U.pack_uniform(codes), U.pack_codes(list(codes))
    assert np.array_equal(a.packed, b.packed) and np.array_equal(a.off, b.off) and np.array_equal(a.len, b.len)
    for k in (0, 5, 16):
        assert np.array_equal(a.codes(k), codes[k])
    ragged = [rng.integers(0, 5, size=int(n), dtype=np.uint8) for n in (0, 1, 3, 4, 5, 250, 9)]
    s = U.pack_codes(ragged)
    for k, c in enumerate(ragged):
        assert np.array_equal(s.codes(k), c)


def test_batch_packer_matches_the_numpy_packer():
    """ubu_sw_pack_batch (8 characters per step, all cores) against the numpy packer: ragged lengths, both cases, N and other
    non-ACGT bytes at every position of the 8-character groups, empty sequences."""
    rng = np.random.default_rng(3)
    alphabet = np.frombuffer(b"ACGTacgtNnRY-*", dtype=np.uint8)
    p = np.array([.2, .2, .2, .2, .04, .04, .04, .04, .01, .005, .005, .005, .005, .01]); p = p / p.sum()
    texts = [alphabet[rng.choice(len(alphabet), size=int(n), p=p)].tobytes() for n in
             list(rng.integers(0, 300, size=4000)) + [0, 1, 7, 8, 9, 15, 16, 17, 255, 256, 65535]]
    texts += [b"ACGT" * 19, b"acgtacgt", b"NNNNNNNNN", b"ACGTACGN", b"NCGTACGT"]
    lens = np.array([len(t) for t in texts], dtype=np.uint16)
    off = np.zeros(len(texts), dtype=np.uint64); off[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
    buf = np.frombuffer(b"".join(texts), dtype=np.uint8)
    ref = U.pack_texts(texts)
    for threads in (1, 0, 3):
        got = U.pack_text_batch(buf, off, lens, threads=threads)
        assert np.array_equal(got.len, ref.len) and np.array_equal(got.off, ref.off)
        assert np.array_equal(got.packed, ref.packed[: got.packed.size])
        assert got.nmask is not None and np.array_equal(got.nmask, ref.nmask[: got.nmask.size]) and np.array_equal(got.nmask_off, ref.nmask_off)
    clean = [t for t in texts if not (set(t) - set(b"ACGTacgt"))]
    lens = np.array([len(t) for t in clean], dtype=np.uint16)
    off = np.zeros(len(clean), dtype=np.uint64); off[1:] = np.cumsum(lens.astype(np.uint64))[:-1]
    got = U.pack_text_batch(np.frombuffer(b"".join(clean), dtype=np.uint8), off, lens)
    assert got.nmask is None and np.array_equal(got.packed, U.pack_texts(clean).packed[: got.packed.size])
