// records.hpp -- C++ host-side mirror of the reference's record-level steps around the aligner (SURVEY.md section 8a rows a4-a7, a9,
// a13). The reference's host code is Java (no JDK in this image), so these carry the reference's own class and method names,
// argument meaning and error behaviour (paths relative to src/main/java/edu/unc/bioinf/ubu/):
//
//     CombineChimera3.combine / combineChimericReads / pruneLikelyInserts     chimera/CombineChimera3.java:21-73, 75-184, 251-267
//     ReAligner.alignAndCleanContigs                                          assembly/ReAligner.java:351-367
//     ReAligner.cleanAndOutputContigs / removeSoftClips                       assembly/ReAligner.java:876-921, 926-961
//     ReAligner.alignToContigs                                                assembly/ReAligner.java:963-975
//     ReAligner.adjustReads (+ HitInfo, getEditDistance, getIntAttribute)     assembly/ReAligner.java:1048-1291, 976-1046
//     ReAligner.updateReadAlignment + ReadBlock                               assembly/ReAligner.java:1293-1368, sam/ReadBlock.java
//     Sam2Fastq.convert (single and paired) / samReadToFastqRecord            fastq/Sam2Fastq.java:33-79, 85-109, 111-162
//
// Bug-compatible where the Java has quirks a consumer could observe (removeSoftClips' off-by-one, the XA loop that skips the last
// entry, the literal "100M" gate -- the latter is a parameter whose default is the reference's literal). SAM is read and written
// as TEXT: BAM and the Picard sort/index steps stay with the caller's htsjdk stack (north_star), as do file formats other than SAM.
// The alignment inside alignAndCleanContigs / alignToContigs runs on the GPU through ubu::Aligner (aligner.hpp); nothing here
// computes an alignment.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace ubu {

struct CigarElement { int length; char op; };

// The handful of net.sf.samtools.SAMRecord behaviours the Java above relies on.
struct SamRecord {
  std::string qname, rname, rnext, seq, qual;       // seq / qual empty = "*"
  int flag = 0, pos = 0, mapq = 0, pnext = 0, tlen = 0;
  std::vector<CigarElement> cigar;
  struct Tag { std::string tag, type, value; };
  std::vector<Tag> tags;                            // file / insertion order

  static SamRecord parse(const std::string& line);
  std::string getSAMString() const;                 // tab-joined text line + '\n'
  bool getReadUnmappedFlag() const { return flag & 0x4; }
  bool getReadNegativeStrandFlag() const { return flag & 0x10; }
  bool getFirstOfPairFlag() const { return flag & 0x40; }
  bool getSecondOfPairFlag() const { return flag & 0x80; }
  std::string getReadString() const { return seq.empty() ? "*" : seq; }
  int getReadLength() const { return (int)seq.size(); }
  std::string getCigarString() const;
  int getCigarReadLength() const;                   // Cigar.getReadLength(): M I S = X
  int getAlignmentEnd() const;                      // start + reference length - 1; 0 for unmapped reads
  bool getIntegerAttribute(const std::string& tag, int& out) const;
  const std::string* getAttribute(const std::string& tag) const;
  void setAttribute(const std::string& tag, int value);
  void setAttribute(const std::string& tag, const std::string& value);
  void removeAttribute(const std::string& tag);     // setAttribute(tag, null)
};

std::vector<CigarElement> parseCigar(const std::string& text);
void parseSamText(const std::string& text, std::vector<std::string>& header, std::vector<SamRecord>& records);

// sam/ReadBlock.java
struct ReadBlock {
  int readStart, referenceStart, length; char type;
  int getReferenceStop() const { return referenceStart + length - 1; }
  int getReferenceLength() const { return type == 'D' ? 0 : length; }
  static std::vector<ReadBlock> getReadBlocks(int alignmentStart, const std::vector<CigarElement>& cigar);
};

// ReAligner.updateReadAlignment (ReAligner.java:1293-1368): new start + CIGAR of a read placed ungapped at 0-based `position` of a
// contig. Returns false where the reference returns a null read; throws std::logic_error where it throws IllegalStateException.
// (The batch form of the same arithmetic on the GPU is ubu_sw_project_batch, include/ubu_sw.h.)
bool updateReadAlignment(int contigPos, const std::vector<CigarElement>& contigCigar, int position, int readLength, int& newStart,
                         std::vector<CigarElement>& newCigar);

// CompareToReference.numDifferences (assembly/CompareToReference.java:83-123), letter for letter: case-folded, 'N' on either side is never
// a mismatch, every OTHER letter (IUPAC codes included) is compared as a letter. The device form (ubu_sw_recount_batch) works on 2-bit
// codes and folds every non-ACGT letter into N, so it under-counts when a read or the reference holds IUPAC codes: callers route such
// reads here (countNonAcgtn tells which).
int numDifferences(const std::string& readBases, const std::string& refBases, int alignmentStart, const std::vector<CigarElement>& cigar);
size_t countNonAcgtn(const std::string& bases);

class CombineChimera3 {
 public:
  void combine(const std::string& input, const std::string& output);                         // files (SAM text)
  static std::string combineText(const std::string& samText);
  static bool combineChimericReads(const SamRecord& read1, const SamRecord& read2, SamRecord& combined);
  static void sortReadsByPosition(std::vector<SamRecord>& reads);
  static void pruneLikelyInserts(std::vector<SamRecord>& sortedReads);
  static int getMappedReadStart(const SamRecord& read);
  static int getMappedReadEnd(const SamRecord& read);
};

class Sam2Fastq {
 public:
  void convert(const std::string& inputSam, const std::string& outputFastq);
  void convert(const std::string& inputSam, const std::string& outputFastq1, const std::string& outputFastq2);
  void setEndSuffixes(const std::string& end1Suffix, const std::string& end2Suffix);
  std::string convertText(const std::string& samText) const;
  void convertPairedText(const std::string& samText, std::string& out1, std::string& out2) const;   // throws std::logic_error on unequal counts
 private:
  bool byReadId_ = false;
  std::string end1_, end2_;
};

struct AdjustStats { int missingXATag = 0, mismatchIssue = 0, adjusted = 0; };

class ReAligner {
 public:
  ReAligner(const std::string& reference, int numThreads, int minContigLength = 100, int device = 0)
      : reference_(reference), numThreads_(numThreads), minContigLength_(minContigLength), device_(device) {}

  // returns the clean-contig FASTA path, or "" when no contig survived (the reference returns null)
  std::string alignAndCleanContigs(const std::string& contigFasta, const std::string& tempDir, bool shouldRemoveSoftClips);
  bool cleanAndOutputContigs(const std::string& contigsSam, const std::string& cleanContigsFasta, bool shouldRemoveSoftClips);
  void alignToContigs(const std::string& tempDir, const std::string& inputSam, const std::string& alignedToContigSam, const std::string& contigFasta);
  void adjustReads(const std::string& originalReadsSam, const std::string& alignedToContigSam, const std::string& outputSam);

  static void removeSoftClips(SamRecord& read);
  static std::string cleanAndOutputContigsText(const std::string& samText, int minContigLength, bool shouldRemoveSoftClips, bool& hasCleanContigs);
  // fullMatchCigar: the reference compares with the literal "100M" (ReAligner.java:1091,1145,1149); "" = "<read length>M" per read,
  // the generalisation BASELINE's 76 / 150 bp configs need (a flagged behaviour change, never the default)
  static std::string adjustReadsText(const std::string& origSam, const std::string& toContigSam, const std::string& fullMatchCigar, AdjustStats& stats);

  AdjustStats lastAdjustStats;
  std::string fullMatchCigar = "100M";

 private:
  std::string reference_;
  int numThreads_, minContigLength_, device_;
};

}  // namespace ubu
