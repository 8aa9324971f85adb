"""ctypes face of libubu_host.so (include/ubu_records.h): the reference's record-level steps around the aligner as C++ host code
(host/records.cpp) -- CombineChimera3.combine, ReAligner.cleanAndOutputContigs / removeSoftClips / adjustReads /
updateReadAlignment / alignAndCleanContigs / alignToContigs, Sam2Fastq.convert. Text in, text out; nothing here is Python logic."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "host", "libubu_host.so")
OK, ERR_ARG, ERR_STATE, ERR_IO, ERR_ALIGNER = 0, -1, -2, -3, -4
SYMBOLS = ["ubu_rec_combine_chimera", "ubu_rec_clean_contigs", "ubu_rec_adjust_reads", "ubu_rec_update_read_alignment", "ubu_rec_sam2fastq",
           "ubu_rec_sam2fastq_paired", "ubu_rec_align_and_clean_contigs", "ubu_rec_align_to_contigs", "ubu_rec_free", "ubu_rec_last_error",
           "ubu_aligner_align", "ubu_aligner_short_align", "ubu_aligner_index", "ubu_rec_num_differences", "ubu_rec_count_non_acgtn"]
_lib = None


class RecordsError(RuntimeError):
    def __init__(self, status: int, msg: str):
        self.status = status
        super().__init__(f"ubu_rec status {status}: {msg}")


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python __graft_entry__.py`")
        from . import _native as N

        N.lib()                                   # libubu_sw.so first (libubu_host.so links against it)
        l = C.CDLL(LIB_PATH)
        l.ubu_rec_last_error.restype = C.c_char_p
        l.ubu_rec_free.argtypes = [C.c_void_p]
        l.ubu_rec_count_non_acgtn.restype = C.c_uint64
        _lib = l
    return _lib


def _take(p: C.c_void_p) -> str:
    s = C.string_at(p).decode()
    lib().ubu_rec_free(p)
    return s


def _check(rc: int) -> None:
    if rc != OK:
        raise RecordsError(rc, lib().ubu_rec_last_error().decode())


def combine_chimera(sam_text: str) -> str:
    out = C.c_void_p()
    _check(lib().ubu_rec_combine_chimera(sam_text.encode(), C.byref(out)))
    return _take(out)


def clean_contigs(sam_text: str, min_contig_length: int, should_remove_soft_clips: bool) -> Tuple[str, bool]:
    out, has = C.c_void_p(), C.c_int(0)
    _check(lib().ubu_rec_clean_contigs(sam_text.encode(), min_contig_length, int(should_remove_soft_clips), C.byref(out), C.byref(has)))
    return _take(out), bool(has.value)


def adjust_reads(orig_sam: str, to_contig_sam: str, full_match_cigar: Optional[str] = None) -> Tuple[str, dict]:
    out, st = C.c_void_p(), (C.c_int32 * 3)()
    fm = None if full_match_cigar is None else full_match_cigar.encode()
    _check(lib().ubu_rec_adjust_reads(orig_sam.encode(), to_contig_sam.encode(), fm, C.byref(out), st))
    return _take(out), {"missing_xa": st[0], "mismatch_issue": st[1], "adjusted": st[2]}


def update_read_alignment(contig_pos: int, contig_cigar: Sequence[Tuple[int, int]], position: int, read_length: int):
    cc = np.array([(n << 4) | o for n, o in contig_cigar] or [0], dtype=np.uint32)
    out = np.zeros(len(contig_cigar) + 4, dtype=np.uint32)
    ns, n, isnull = C.c_int32(0), C.c_uint32(0), C.c_int(0)
    _check(lib().ubu_rec_update_read_alignment(contig_pos, cc.ctypes.data_as(C.c_void_p), len(contig_cigar), position, read_length, C.byref(ns),
                                               out.ctypes.data_as(C.c_void_p), out.shape[0], C.byref(n), C.byref(isnull)))
    if isnull.value:
        return None
    return int(ns.value), [(int(o) >> 4, int(o) & 15) for o in out[: n.value].view(np.int32)]      # lengths are signed: the Java keeps negative ones


def sam2fastq(sam_text: str) -> str:
    out = C.c_void_p()
    _check(lib().ubu_rec_sam2fastq(sam_text.encode(), C.byref(out)))
    return _take(out)


def sam2fastq_paired(sam_text: str, end1: Optional[str] = None, end2: Optional[str] = None) -> Tuple[str, str]:
    o1, o2 = C.c_void_p(), C.c_void_p()
    _check(lib().ubu_rec_sam2fastq_paired(sam_text.encode(), None if end1 is None else end1.encode(), None if end2 is None else end2.encode(),
                                          C.byref(o1), C.byref(o2)))
    return _take(o1), _take(o2)


def align_and_clean_contigs(reference_fasta: str, contig_fasta: str, temp_dir: str, should_remove_soft_clips: bool, min_contig_length: int = 100,
                            num_threads: int = 1, device: int = 0) -> Optional[str]:
    out = C.c_void_p()
    _check(lib().ubu_rec_align_and_clean_contigs(reference_fasta.encode(), num_threads, min_contig_length, device, contig_fasta.encode(), temp_dir.encode(),
                                                 int(should_remove_soft_clips), C.byref(out)))
    s = _take(out)
    return s or None


def align_to_contigs(temp_dir: str, input_sam: str, aligned_to_contig_sam: str, contig_fasta: str, num_threads: int = 1, device: int = 0) -> None:
    _check(lib().ubu_rec_align_to_contigs(num_threads, device, temp_dir.encode(), input_sam.encode(), aligned_to_contig_sam.encode(), contig_fasta.encode()))


def aligner_align(reference_fasta: str, input_path: str, output_sam: str, short: bool = False, num_threads: int = 1, device: int = 0) -> None:
    """Aligner.align / Aligner.shortAlign (Aligner.java:18-22, 46-56) through the C ABI a JNI caller binds."""
    fn = lib().ubu_aligner_short_align if short else lib().ubu_aligner_align
    _check(fn(reference_fasta.encode(), num_threads, device, input_path.encode(), output_sam.encode()))


def num_differences(read_bases: str, ref_bases: str, alignment_start: int, cigar: str) -> int:
    """CompareToReference.numDifferences letter for letter (IUPAC codes compared as letters)."""
    out = C.c_int32(0)
    _check(lib().ubu_rec_num_differences(read_bases.encode(), ref_bases.encode(), alignment_start, cigar.encode(), C.byref(out)))
    return int(out.value)


def count_non_acgtn(bases: str) -> int:
    return int(lib().ubu_rec_count_non_acgtn(bases.encode()))
