// records_abi.cpp -- the C ABI of include/ubu_records.h over host/records.cpp.
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../include/ubu_records.h"
#include "aligner.hpp"
#include "records.hpp"

namespace {
thread_local std::string g_err;
char* dup(const std::string& s) { char* p = (char*)malloc(s.size() + 1); if (p) { memcpy(p, s.data(), s.size()); p[s.size()] = 0; } return p; }
template <class F> int guarded(F f) {
  try { return f(); }
  catch (const std::logic_error& e) { g_err = e.what(); return UBU_REC_ERR_STATE; }            // IllegalStateException / index out of range in the Java
  catch (const std::runtime_error& e) { g_err = e.what(); return strncmp(e.what(), "cannot ", 7) == 0 ? UBU_REC_ERR_IO : UBU_REC_ERR_ALIGNER; }
  catch (const std::exception& e) { g_err = e.what(); return UBU_REC_ERR_STATE; }
}
const int kOps = 5; const char kOpChar[kOps + 1] = "MIDNS";
}  // namespace

extern "C" {

int ubu_rec_combine_chimera(const char* sam_in, char** sam_out) {
  if (!sam_in || !sam_out) return UBU_REC_ERR_ARG;
  return guarded([&] { *sam_out = dup(ubu::CombineChimera3::combineText(sam_in)); return UBU_REC_OK; });
}
int ubu_rec_clean_contigs(const char* sam_in, int min_len, int remove_sc, char** fasta_out, int* has) {
  if (!sam_in || !fasta_out) return UBU_REC_ERR_ARG;
  return guarded([&] { bool h = false; *fasta_out = dup(ubu::ReAligner::cleanAndOutputContigsText(sam_in, min_len, remove_sc != 0, h)); if (has) *has = h ? 1 : 0; return UBU_REC_OK; });
}
int ubu_rec_adjust_reads(const char* orig_sam, const char* to_contig_sam, const char* full_match_cigar, char** sam_out, int32_t stats[3]) {
  if (!orig_sam || !to_contig_sam || !sam_out) return UBU_REC_ERR_ARG;
  return guarded([&] {
    ubu::AdjustStats st;
    *sam_out = dup(ubu::ReAligner::adjustReadsText(orig_sam, to_contig_sam, full_match_cigar ? full_match_cigar : "100M", st));
    if (stats) { stats[0] = st.missingXATag; stats[1] = st.mismatchIssue; stats[2] = st.adjusted; }
    return UBU_REC_OK;
  });
}
int ubu_rec_update_read_alignment(int32_t contig_pos, const uint32_t* cc, uint32_t cn, int32_t position, int32_t read_length, int32_t* new_start,
                                  uint32_t* cigar_out, uint32_t cap, uint32_t* n_out, int* is_null) {
  if ((cn && !cc) || !new_start || !n_out || !is_null) return UBU_REC_ERR_ARG;
  return guarded([&] {
    std::vector<ubu::CigarElement> contig, out;
    for (uint32_t k = 0; k < cn; ++k) { const uint32_t op = cc[k] & 15u; if (op >= (uint32_t)kOps) throw std::logic_error("Case statement didn't deal with cigar op"); contig.push_back({(int)(cc[k] >> 4), kOpChar[op]}); }
    int ns = 0;
    const bool ok = ubu::updateReadAlignment(contig_pos, contig, position, read_length, ns, out);
    *is_null = ok ? 0 : 1; *n_out = 0;
    if (!ok) return UBU_REC_OK;
    *new_start = ns; *n_out = (uint32_t)out.size();
    if (out.size() > cap || (out.size() && !cigar_out)) return UBU_REC_ERR_ARG;
    for (size_t k = 0; k < out.size(); ++k) cigar_out[k] = ((uint32_t)out[k].length << 4) | (uint32_t)(strchr(kOpChar, out[k].op) - kOpChar);
    return UBU_REC_OK;
  });
}
int ubu_rec_num_differences(const char* read_bases, const char* ref_bases, int32_t alignment_start, const char* cigar, int32_t* diffs) {
  if (!read_bases || !ref_bases || !cigar || !diffs) return UBU_REC_ERR_ARG;
  return guarded([&] { *diffs = ubu::numDifferences(read_bases, ref_bases, alignment_start, ubu::parseCigar(cigar)); return UBU_REC_OK; });
}
uint64_t ubu_rec_count_non_acgtn(const char* bases) { return bases ? (uint64_t)ubu::countNonAcgtn(bases) : 0; }
int ubu_rec_sam2fastq(const char* sam_in, char** fastq_out) {
  if (!sam_in || !fastq_out) return UBU_REC_ERR_ARG;
  return guarded([&] { *fastq_out = dup(ubu::Sam2Fastq().convertText(sam_in)); return UBU_REC_OK; });
}
int ubu_rec_sam2fastq_paired(const char* sam_in, const char* end1, const char* end2, char** o1, char** o2) {
  if (!sam_in || !o1 || !o2 || ((end1 == nullptr) != (end2 == nullptr))) return UBU_REC_ERR_ARG;
  return guarded([&] {
    ubu::Sam2Fastq s; if (end1) s.setEndSuffixes(end1, end2);
    std::string a, b; s.convertPairedText(sam_in, a, b); *o1 = dup(a); *o2 = dup(b); return UBU_REC_OK;
  });
}
int ubu_rec_align_and_clean_contigs(const char* reference_fasta, int num_threads, int min_len, int device, const char* contig_fasta, const char* temp_dir,
                                    int remove_sc, char** clean_fasta_out) {
  if (!reference_fasta || !contig_fasta || !temp_dir || !clean_fasta_out) return UBU_REC_ERR_ARG;
  return guarded([&] { ubu::ReAligner r(reference_fasta, num_threads, min_len, device); *clean_fasta_out = dup(r.alignAndCleanContigs(contig_fasta, temp_dir, remove_sc != 0)); return UBU_REC_OK; });
}
int ubu_rec_align_to_contigs(int num_threads, int device, const char* temp_dir, const char* input_sam, const char* aligned, const char* contig_fasta) {
  if (!temp_dir || !input_sam || !aligned || !contig_fasta) return UBU_REC_ERR_ARG;
  return guarded([&] { ubu::ReAligner r("", num_threads, 100, device); r.alignToContigs(temp_dir, input_sam, aligned, contig_fasta); return UBU_REC_OK; });
}
int ubu_aligner_align(const char* reference_fasta, int num_threads, int device, const char* input, const char* output_sam) {
  if (!reference_fasta || !input || !output_sam) return UBU_REC_ERR_ARG;
  return guarded([&] { ubu::Aligner a(reference_fasta, num_threads, device); a.align(input, output_sam); return UBU_REC_OK; });
}
int ubu_aligner_short_align(const char* reference_fasta, int num_threads, int device, const char* input, const char* output_sam) {
  if (!reference_fasta || !input || !output_sam) return UBU_REC_ERR_ARG;
  return guarded([&] { ubu::Aligner a(reference_fasta, num_threads, device); a.shortAlign(input, output_sam); return UBU_REC_OK; });
}
int ubu_aligner_index(const char* reference_fasta) {
  if (!reference_fasta) return UBU_REC_ERR_ARG;
  return guarded([&] { ubu::Aligner a(reference_fasta, 1); a.index(); return UBU_REC_OK; });
}
void ubu_rec_free(char* p) { free(p); }
const char* ubu_rec_last_error(void) { return g_err.c_str(); }

}  // extern "C"
