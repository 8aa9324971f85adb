/* sw_oracle.h -- CPU oracle for the SPEC.md Smith-Waterman contract.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under ubu_b200/ may include, link or call
 * this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do (as the checker / reported baseline).
 *
 * PARITY UNPINNED: the reference (mozack/ubu) holds no Smith-Waterman -- its
 * Aligner forks an external `bwa` (src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:18-22,46-56)
 * and no reference test pins an alignment result. This oracle is a direct,
 * deliberately naive transcription of this repo's SPEC.md (BUILD-DEFINED).
 * What IS pinned to the reference are the sequence-level conventions:
 *   2-bit code / bit order      Sequence.java:10-13,31-41
 *   reverse complement          ReverseComplementor.java:15-22,38-62 (+ its test vectors)
 *   mismatch / NM definition    CompareToReference.java:83-123, ReAligner.java:282-285,316-326
 *   CIGAR op alphabet           ReadBlock.java:135-153
 */
#ifndef SW_ORACLE_H
#define SW_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SWO_BASE_N 4u            /* unpacked code for N; A=0 C=1 T=2 G=3 (Sequence.java:10-13) */
#define SWO_OP_M 0u
#define SWO_OP_I 1u
#define SWO_OP_D 2u
#define SWO_OP_S 4u
#define SWO_FLAG_UNALIGNED 1u

typedef struct { int32_t match, mismatch, gap_open, gap_ext; } swo_params;

typedef struct {
  int32_t score, ref_start, ref_end, q_start, q_end, edit_distance;
  uint32_t cigar_n, flags;
} swo_result;

/* One task. q/r are unpacked codes (0..3, 4 = N), one byte per base.
 * cigar receives BAM-encoded ops (len<<4|op). Returns 0, or -1 if cigar_cap is
 * too small (res->cigar_n then holds the needed count), -2 on bad params / OOM. */
int swo_align(const uint8_t* q, int L, const uint8_t* r, int W, const swo_params* p,
              swo_result* res, uint32_t* cigar, uint32_t cigar_cap);

/* n tasks over concatenated code arrays; cigars written at i*cigar_stride.
 * nthreads <= 0 -> 1. Returns 0 or the first non-zero per-task status. */
int swo_align_batch(const uint8_t* q_codes, const uint64_t* q_off, const uint32_t* q_len,
                    const uint8_t* r_codes, const uint64_t* r_off, const uint32_t* r_len,
                    const uint32_t* pair_ref_idx, uint32_t n, const swo_params* p,
                    swo_result* res, uint32_t* cigars, uint32_t cigar_stride, int nthreads);

/* Sequence conventions (each cites the reference in sw_oracle.c). */
uint8_t swo_code_of_char(char c);                                   /* ACGTacgt -> 0,1,3,2 ; else 4 */
char    swo_char_of_code(uint8_t code);
void    swo_encode(const char* s, size_t n, uint8_t* codes);
void    swo_pack2(const uint8_t* codes, size_t n, uint8_t* packed, uint8_t* nmask);
void    swo_unpack2(const uint8_t* packed, const uint8_t* nmask, size_t n, uint8_t* codes);
void    swo_revcomp_text(const char* in, size_t n, char* out);      /* non-ACGT bytes pass through */
void    swo_revcomp_codes(const uint8_t* in, size_t n, uint8_t* out);

/* Re-score a CIGAR under SPEC sections 2/3 (property-test helper): returns the score of the aligned part
 * and writes the edit distance; INT32_MIN if the CIGAR is inconsistent with the coordinates. */
int32_t swo_rescore(const uint8_t* q, int L, const uint8_t* r, int W, const swo_params* p,
                    const swo_result* res, const uint32_t* cigar, int32_t* edit_out);

#ifdef __cplusplus
}
#endif
#endif
