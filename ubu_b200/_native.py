"""ctypes binding of libubu_sw.so (include/ubu_sw.h). Fails loudly when the library is absent:
there is no Python / CPU implementation of the alignment path in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("UBU_SW_LIB") or os.path.join(HERE, "libubu_sw.so")      # UBU_SW_LIB: try another build of the library

OK = 0
ERR_ARG, ERR_CUDA, ERR_CIGAR_POOL, ERR_TOO_LONG, ERR_NO_DEVICE, ERR_BUSY, ERR_NOMEM = -1, -2, -3, -4, -5, -6, -7
FLAG_UNALIGNED = 1
MAX_READ = 4096
MAX_LONG_CELLS = 1 << 26
MAX_WINDOW = 65535


class Params(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open", C.c_int32), ("gap_ext", C.c_int32)]


class Result(C.Structure):
    _fields_ = [
        ("score", C.c_int32), ("ref_start", C.c_int32), ("ref_end", C.c_int32),
        ("q_start", C.c_int32), ("q_end", C.c_int32), ("edit_distance", C.c_int32),
        ("cigar_off", C.c_uint32), ("cigar_n", C.c_uint16), ("flags", C.c_uint16),
    ]


class Seqs(C.Structure):
    _fields_ = [
        ("packed", C.c_void_p), ("off", C.c_void_p), ("len", C.c_void_p),
        ("nmask", C.c_void_p), ("nmask_off", C.c_void_p),
        ("count", C.c_uint32), ("packed_bytes", C.c_uint32), ("nmask_bytes", C.c_uint32),
    ]


class Refs(C.Structure):
    _fields_ = [
        ("packed", C.c_void_p), ("off", C.c_void_p), ("len", C.c_void_p), ("nmask", C.c_void_p), ("nmask_off", C.c_void_p),
        ("count", C.c_uint32), ("packed_bytes", C.c_uint64), ("nmask_bytes", C.c_uint64),
    ]


class SamNames(C.Structure):
    _fields_ = [("qname_prefix", C.c_char_p), ("first_index", C.c_uint64), ("pair_ref_idx", C.c_void_p), ("rnames", C.c_void_p),
                ("window_pos0", C.c_void_p), ("cigar_pool_n", C.c_uint64)]


class KmerParams(C.Structure):
    _fields_ = [("kmer_size", C.c_uint32), ("min_node_frequency", C.c_uint32)]


POST_NULL, POST_MALFORMED, POST_POOL = 1, 2, 4
ERR_CAPACITY = -8
KMER_NONE = 0xFFFFFFFF
KMER_FLAG_MULTI_READS, KMER_FLAG_FILTERED, KMER_FLAG_ROOT = 1, 2, 4

# every symbol include/ubu_sw.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("ubu_sw_create", C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("ubu_sw_destroy", None, [C.c_void_p]),
    ("ubu_sw_align_batch", C.c_int, [C.c_void_p, C.POINTER(Seqs), C.POINTER(Seqs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("ubu_sw_submit", C.c_int, [C.c_void_p, C.POINTER(Seqs), C.POINTER(Seqs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]),
    ("ubu_sw_wait", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]),
    ("ubu_sw_stage", C.c_int, [C.c_void_p, C.c_int, C.POINTER(Seqs), C.POINTER(Seqs), C.c_void_p]),
    ("ubu_sw_run", C.c_int, [C.c_void_p, C.c_int]),
    ("ubu_sw_sync", C.c_int, [C.c_void_p, C.c_int]),
    ("ubu_sw_fetch", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("ubu_sw_last_run_ms", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    ("ubu_sw_last_run_launches", C.c_int, [C.c_void_p, C.c_int]),
    ("ubu_sw_debug_counters", C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    ("ubu_sw_multi_create", C.c_int, [C.POINTER(Params), C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]),
    ("ubu_sw_multi_destroy", None, [C.c_void_p]),
    ("ubu_sw_multi_device_count", C.c_int, [C.c_void_p]),
    ("ubu_sw_multi_last_error", C.c_char_p, [C.c_void_p]),
    ("ubu_sw_multi_align", C.c_int, [C.c_void_p, C.POINTER(Seqs), C.POINTER(Seqs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("ubu_sw_multi_align_sam", C.c_int, [C.c_void_p, C.POINTER(Seqs), C.POINTER(Seqs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                         C.c_int, C.POINTER(SamNames), C.c_int, C.POINTER(C.c_uint64)]),
    ("ubu_sw_load_refs", C.c_int, [C.c_void_p, C.POINTER(Refs)]),
    ("ubu_sw_recount_batch", C.c_int, [C.c_void_p, C.POINTER(Seqs), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    ("ubu_sw_project_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint32, C.c_void_p,
                                       C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("ubu_sw_host_alloc", C.c_void_p, [C.c_size_t]),
    ("ubu_sw_host_free", None, [C.c_void_p]),
    ("ubu_sw_pack", None, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    ("ubu_sw_pack_batch", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.POINTER(C.c_uint32)]),
    ("ubu_sw_sam_write", C.c_int, [C.c_int, C.POINTER(Seqs), C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(SamNames), C.c_int, C.POINTER(C.c_uint64)]),
    ("ubu_sw_revcomp", None, [C.c_char_p, C.c_size_t, C.c_char_p]),
    ("ubu_sw_cigar_string", C.c_int, [C.c_void_p, C.c_uint32, C.c_char_p, C.c_size_t]),
    ("ubu_sw_debug_plan", C.c_int, [C.POINTER(Seqs), C.POINTER(Seqs), C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_int),
                                    C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    ("ubu_sw_device_count", C.c_int, []),
    ("ubu_sw_version", C.c_uint, []),
    ("ubu_sw_strerror", C.c_char_p, [C.c_int]),
    ("ubu_sw_last_error", C.c_char_p, [C.c_void_p]),
    # include/ubu_kmer.h
    ("ubu_kmer_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("ubu_kmer_destroy", None, [C.c_void_p]),
    ("ubu_kmer_last_error", C.c_char_p, [C.c_void_p]),
    ("ubu_kmer_build", C.c_int, [C.c_void_p, C.POINTER(KmerParams), C.POINTER(Seqs), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                 C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]),
    ("ubu_kmer_stage", C.c_int, [C.c_void_p, C.POINTER(KmerParams), C.POINTER(Seqs), C.c_void_p, C.c_size_t]),
    ("ubu_kmer_run", C.c_int, [C.c_void_p]),
    ("ubu_kmer_fetch", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32)]),
    ("ubu_kmer_last_run_ms", C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    ("ubu_kmer_last_run_stats", C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
]

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Loads libubu_sw.so (built in-tree by ubu_b200/build.py). No fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                f"{LIB_PATH} not found: build it with `python -m ubu_b200.build` (nvcc, sm_100a). "
                "ubu_b200 has no CPU implementation of the alignment path."
            )
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class UbuSwError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        msg = lib().ubu_sw_strerror(status).decode()
        super().__init__(f"ubu_sw status {status}: {msg}" + (f" ({detail})" if detail else ""))
