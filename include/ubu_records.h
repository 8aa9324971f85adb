/* ubu_records.h -- C ABI of the record-level steps around the aligner (host/records.{hpp,cpp}; libubu_host.so).
 *
 * Text in, text out: what a JNI / FFM caller hands over is the SAM text its htsjdk stack already holds (BAM, sorting and indexing
 * stay in Java, north_star). Each entry point replaces one reference method (paths relative to src/main/java/edu/unc/bioinf/ubu/):
 *     ubu_rec_combine_chimera         CombineChimera3.combine                     chimera/CombineChimera3.java:21-73 (75-184, 251-267)
 *     ubu_rec_clean_contigs           ReAligner.cleanAndOutputContigs             assembly/ReAligner.java:876-921 (removeSoftClips :926-961)
 *     ubu_rec_adjust_reads            ReAligner.adjustReads                       assembly/ReAligner.java:1048-1291
 *     ubu_rec_update_read_alignment   ReAligner.updateReadAlignment               assembly/ReAligner.java:1293-1368 (batch form on the GPU: ubu_sw_project_batch)
 *     ubu_rec_sam2fastq[_paired]      Sam2Fastq.convert                           fastq/Sam2Fastq.java:85-109, 33-79
 *     ubu_rec_align_and_clean_contigs ReAligner.alignAndCleanContigs              assembly/ReAligner.java:351-367  (paths; aligns on the GPU)
 *     ubu_rec_align_to_contigs        ReAligner.alignToContigs                    assembly/ReAligner.java:963-975  (paths; aligns on the GPU)
 * Returned strings are malloc'd; release them with ubu_rec_free. Status: 0 ok, UBU_REC_ERR_* otherwise (the reference throws:
 * IllegalStateException / StringIndexOutOfBounds -> UBU_REC_ERR_STATE, RuntimeException from the aligner -> UBU_REC_ERR_ALIGNER);
 * ubu_rec_last_error() has the message of the calling thread's last failure.
 */
#ifndef UBU_RECORDS_H
#define UBU_RECORDS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define UBU_REC_OK 0
#define UBU_REC_ERR_ARG (-1)
#define UBU_REC_ERR_STATE (-2)
#define UBU_REC_ERR_IO (-3)
#define UBU_REC_ERR_ALIGNER (-4)

int ubu_rec_combine_chimera(const char* sam_in, char** sam_out);
int ubu_rec_clean_contigs(const char* sam_in, int min_contig_length, int should_remove_soft_clips, char** fasta_out, int* has_clean_contigs);
/* full_match_cigar: NULL = the reference's literal "100M" (ReAligner.java:1091,1145,1149); "" = "<read length>M" per read.
 * stats[3] = {reads whose XA tag was missing, MISMATCH_ISSUE count, records written as adjusted}. out = records only (no header). */
int ubu_rec_adjust_reads(const char* orig_sam, const char* to_contig_sam, const char* full_match_cigar, char** sam_out, int32_t stats[3]);
/* CIGAR elements as BAM-encoded ops (len << 4 | op, MIDNS = 0..4; output lengths are SIGNED -- (int32_t)op >> 4 -- because the Java keeps the
 * negative block lengths its arithmetic can produce). Returns 1 in *is_null where the reference returns a null read. */
int ubu_rec_update_read_alignment(int32_t contig_pos, const uint32_t* contig_cigar, uint32_t contig_cigar_n, int32_t position, int32_t read_length,
                                  int32_t* new_start, uint32_t* cigar_out, uint32_t cigar_cap, uint32_t* cigar_n, int* is_null);
/* CompareToReference.numDifferences (assembly/CompareToReference.java:83-123) letter for letter, IUPAC codes compared as letters ('R' vs 'A' IS a
 * mismatch, N never is). The device form, ubu_sw_recount_batch, folds every non-ACGT letter into N: route reads for which
 * ubu_rec_count_non_acgtn(...) > 0 (or whose reference stretch holds IUPAC codes) through this call. cigar as text ("50M2D26M"). */
int ubu_rec_num_differences(const char* read_bases, const char* ref_bases, int32_t alignment_start, const char* cigar, int32_t* diffs);
uint64_t ubu_rec_count_non_acgtn(const char* bases);
int ubu_rec_sam2fastq(const char* sam_in, char** fastq_out);
/* end1 / end2: read-name suffixes identifying the ends (Sam2Fastq.setEndSuffixes); NULL = use the FLAG bits. */
int ubu_rec_sam2fastq_paired(const char* sam_in, const char* end1, const char* end2, char** fastq1_out, char** fastq2_out);
/* clean_fasta_out: path of clean_contigs.fasta, or an empty string when no contig survived (the reference returns null). */
int ubu_rec_align_and_clean_contigs(const char* reference_fasta, int num_threads, int min_contig_length, int device, const char* contig_fasta,
                                    const char* temp_dir, int should_remove_soft_clips, char** clean_fasta_out);
int ubu_rec_align_to_contigs(int num_threads, int device, const char* temp_dir, const char* input_sam, const char* aligned_to_contig_sam,
                             const char* contig_fasta);
/* The aligner facade itself, path in / path out like the reference (assembly/Aligner.java:13-16, 18-22, 46-56, 62-69): every read of `input`
 * (FASTA / FASTQ) against every sequence of `reference_fasta` on the GPU, SAM with NM XM X0 X1 XA written to `output_sam`.
 * A failure returns UBU_REC_ERR_ALIGNER where the reference throws RuntimeException("BWA exited with non-zero return code"). */
int ubu_aligner_align(const char* reference_fasta, int num_threads, int device, const char* input, const char* output_sam);
int ubu_aligner_short_align(const char* reference_fasta, int num_threads, int device, const char* input, const char* output_sam);
int ubu_aligner_index(const char* reference_fasta);
void ubu_rec_free(char* p);
const char* ubu_rec_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
