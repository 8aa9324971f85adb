"""ctypes face of oracle/libsw_oracle.so for the tests, smoke() and bench.py's CPU baseline leg.
TEST INFRASTRUCTURE: never imported from ubu_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libsw_oracle.so")


class SwoParams(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_ext", C.c_int32)]


SWO_RESULT = np.dtype([("score", "<i4"), ("ref_start", "<i4"), ("ref_end", "<i4"), ("q_start", "<i4"), ("q_end", "<i4"),
                       ("edit_distance", "<i4"), ("cigar_n", "<u4"), ("flags", "<u4")])

_lib = None


def build() -> str:
    src = [os.path.join(ORACLE_DIR, f) for f in ("sw_oracle.c", "sw_oracle.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        l.swo_align.restype = C.c_int
        l.swo_align.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(SwoParams), C.c_void_p, C.c_void_p, C.c_uint32]
        l.swo_align_batch.restype = C.c_int
        l.swo_align_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_uint32, C.POINTER(SwoParams), C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        l.swo_rescore.restype = C.c_int32
        l.swo_rescore.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(SwoParams), C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]
        l.swo_pack2.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        l.swo_unpack2.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        l.swo_revcomp_text.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        l.swo_encode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        _lib = l
    return _lib


def align_one(q: np.ndarray, r: np.ndarray, params=(1, -3, 5, 2)):
    """Returns (result record, [(len, op)...])."""
    q = np.ascontiguousarray(q, dtype=np.uint8); r = np.ascontiguousarray(r, dtype=np.uint8)
    p = SwoParams(*params)
    res = np.zeros(1, dtype=SWO_RESULT)
    cap = q.shape[0] + r.shape[0] + 4
    cig = np.zeros(cap, dtype=np.uint32)
    rc = lib().swo_align(q.ctypes.data, q.shape[0], r.ctypes.data, r.shape[0], C.byref(p), res.ctypes.data, cig.ctypes.data, cap)
    assert rc == 0, rc
    n = int(res[0]["cigar_n"])
    return res[0], [(int(o) >> 4, int(o) & 15) for o in cig[:n]]


def align_batch(q_list, r_list, pair_ref_idx=None, params=(1, -3, 5, 2), nthreads=1, cigar_stride=None):
    """q_list / r_list: lists of uint8 code arrays (or 2-D arrays). Returns (results[n], cigars list of lists)."""
    n = len(q_list)
    qlen = np.array([len(x) for x in q_list], dtype=np.uint32)
    rlen = np.array([len(x) for x in r_list], dtype=np.uint32)
    qoff = np.zeros(n, dtype=np.uint64); roff = np.zeros(len(r_list), dtype=np.uint64)
    if n:
        qoff[1:] = np.cumsum(qlen[:-1])
    if len(r_list):
        roff[1:] = np.cumsum(rlen[:-1])
    qc = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint8) for x in q_list]) if n else np.zeros(1, np.uint8))
    rc_ = np.ascontiguousarray(np.concatenate([np.asarray(x, dtype=np.uint8) for x in r_list]) if len(r_list) else np.zeros(1, np.uint8))
    if qc.size == 0:
        qc = np.zeros(1, np.uint8)
    if rc_.size == 0:
        rc_ = np.zeros(1, np.uint8)
    idx = None if pair_ref_idx is None else np.ascontiguousarray(pair_ref_idx, dtype=np.uint32)
    stride = int(cigar_stride or (int(qlen.max() if n else 0) + 8))
    res = np.zeros(n, dtype=SWO_RESULT)
    cig = np.zeros(max(n * stride, 1), dtype=np.uint32)
    p = SwoParams(*params)
    st = lib().swo_align_batch(qc.ctypes.data, qoff.ctypes.data, qlen.ctypes.data, rc_.ctypes.data, roff.ctypes.data, rlen.ctypes.data,
                               None if idx is None else idx.ctypes.data, n, C.byref(p), res.ctypes.data, cig.ctypes.data, stride, nthreads)
    assert st == 0, st
    cigs = [[(int(o) >> 4, int(o) & 15) for o in cig[k * stride: k * stride + int(res[k]["cigar_n"])]] for k in range(n)]
    return res, cigs


def rescore(q, r, res_fields: dict, cigar, params=(1, -3, 5, 2)):
    """Re-scores a CIGAR under SPEC sections 2/3. Returns (score, edit_distance) or (None, None) if inconsistent."""
    q = np.ascontiguousarray(q, dtype=np.uint8); r = np.ascontiguousarray(r, dtype=np.uint8)
    rec = np.zeros(1, dtype=SWO_RESULT)
    for f in ("score", "ref_start", "ref_end", "q_start", "q_end", "edit_distance"):
        rec[0][f] = res_fields[f]
    rec[0]["cigar_n"] = len(cigar)
    ops = np.array([(n << 4) | o for n, o in cigar] or [0], dtype=np.uint32)
    p = SwoParams(*params); ed = C.c_int32(0)
    sc = lib().swo_rescore(q.ctypes.data, q.shape[0], r.ctypes.data, r.shape[0], C.byref(p), rec.ctypes.data, ops.ctypes.data, C.byref(ed))
    if sc == -2 ** 31:
        return None, None
    return int(sc), int(ed.value)
