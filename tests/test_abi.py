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
    assert lib.ubu_sw_version() == 0x00020000
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
    from ubu_b200 import multi as M

    with pytest.raises(U.UbuSwError) as e:                      # the multi-GPU driver has no CPU path either
        M.MultiAligner()
    assert e.value.status == N.ERR_NO_DEVICE
    m = C.c_void_p()
    assert N.lib().ubu_sw_multi_create(None, None, 0, 0, 0, None) == N.ERR_ARG
    assert N.lib().ubu_sw_multi_align(None, None, None, None, None, None, 0, None) == N.ERR_ARG
    assert N.lib().ubu_sw_multi_device_count(None) == 0


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


def test_sam_writer_formats_records_in_input_order(tmp_path):
    """ubu_sw_sam_write on hand-made records (no GPU): every line against a Python rendering, across thread counts and the
    block boundary (16384 reads per block)."""
    rng = np.random.default_rng(8)
    n = 40_000
    texts = ["".join(rng.choice(list("ACGTN"), size=int(rng.integers(0, 90)), p=[.24, .24, .24, .24, .04])) for _ in range(n)]
    reads = U.pack_texts(texts)
    res = np.zeros(n, dtype=U.RESULT_DTYPE)
    pool = []
    idx = rng.integers(0, 500, size=n).astype(np.uint32)
    pos0 = rng.integers(0, 10**9, size=500).astype(np.int64)
    for k in range(n):
        if k % 7 == 3 or not texts[k]:
            res[k]["flags"] = 1
            continue
        ops = [(int(rng.integers(1, 60)), int(o)) for o in rng.choice([0, 1, 2, 4], size=int(rng.integers(1, 6)))]
        res[k]["cigar_off"], res[k]["cigar_n"] = len(pool), len(ops)
        pool += [(ln << 4) | op for ln, op in ops]
        res[k]["score"], res[k]["ref_start"], res[k]["edit_distance"] = int(rng.integers(1, 250)), int(rng.integers(1, 2000)), int(rng.integers(0, 30))
    al = U.Alignments(res, np.array(pool, dtype=np.uint32))
    want = []
    for k in range(n):
        seq = texts[k] or "*"
        if res[k]["flags"] & 1:
            want.append(f"q{1000 + k}\t4\t*\t0\t0\t*\t*\t0\t0\t{seq}\t*")
        else:
            want.append(f"q{1000 + k}\t0\tw{idx[k]}\t{int(pos0[idx[k]]) + int(res[k]['ref_start'])}\t37\t{al.cigar_string(k)}\t*\t0\t0\t{seq}\t*"
                        f"\tNM:i:{int(res[k]['edit_distance'])}\tAS:i:{int(res[k]['score'])}")
    for threads in (1, 0, 5):
        path = tmp_path / f"out{threads}.sam"
        fd = os.open(str(path), os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
        nbytes = U.write_sam(fd, reads, al, idx, first_index=1000, qname_prefix="q", window_pos0=pos0, threads=threads)
        os.close(fd)
        text = path.read_text()
        assert nbytes == len(text.encode()) and text.split("\n")[:-1] == want
    assert U.write_sam(-1, reads, al, idx, first_index=1000, qname_prefix="q", window_pos0=pos0) == nbytes      # format only
    # a record whose CIGAR slice leaves the pool is an argument error, not a wild read (ubu_sw_sam_names.cigar_pool_n)
    bad = U.Alignments(al.results.copy(), al.cigar_pool.copy())
    k = int(np.nonzero(bad.results["cigar_n"] > 0)[0][0])
    bad.results["cigar_off"][k] = bad.cigar_pool.shape[0]
    with pytest.raises(U.UbuSwError) as e:
        U.write_sam(-1, reads, bad, idx)
    assert e.value.status == N.ERR_ARG


def test_jni_shim_parses_against_the_c_abi():
    """integration/jni/ubu_sw_jni.c cannot be built here (no JDK), but it must at least be valid C against include/*.h: compiled
    with -fsyntax-only against a minimal stand-in for jni.h, warnings as errors for argument-type mismatches."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    r = subprocess.run([gcc, "-fsyntax-only", "-Wall", "-Werror=incompatible-pointer-types", "-Werror=int-conversion",
                        "-I" + os.path.join(ROOT, "tests", "mock_jni"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "jni", "ubu_sw_jni.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
