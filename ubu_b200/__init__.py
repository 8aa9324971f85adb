"""ubu_b200 -- B200-native Smith-Waterman realignment path behind ubu's Aligner boundary.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C ABI of include/ubu_sw.h),
the ctypes binding, 2-bit sequence packing and the host-side mirror of the reference's Aligner.
"""
from ._native import LIB_PATH, NativeLibraryMissing, UbuSwError  # noqa: F401
from .aligner import Alignments, SwAligner, RESULT_DTYPE, pinned_empty, write_sam  # noqa: F401
from .seqs import SeqSet, pack_codes, pack_texts, pack_text_batch, pack_uniform, codes_from_text, text_from_codes, reverse_complement  # noqa: F401

__all__ = [
    "SwAligner", "Alignments", "SeqSet", "pack_codes", "pack_texts", "pack_text_batch", "pack_uniform", "codes_from_text",
    "text_from_codes", "reverse_complement", "RESULT_DTYPE", "pinned_empty", "write_sam", "UbuSwError", "NativeLibraryMissing", "LIB_PATH",
]
