// aligner.hpp -- C++ host-side mirror of the reference's aligner class, on top of the C ABI (include/ubu_sw.h).
//
// The reference's host code is Java (no JDK in this image), so the host side above the C ABI is written in C++ with the
// reference's own names, argument meaning and error behaviour:
//     new Aligner(String reference, int numThreads)     src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:13-16
//     void align(String input, String outputSam)        Aligner.java:18-22   (bwa bwasw)
//     void shortAlign(String input, String outputSam)   Aligner.java:46-56   (bwa aln + samse -n 1000)
//     void index()                                      Aligner.java:62-69   (bwa index; nothing to build for window SW)
// Failures surface as std::runtime_error, where the reference throws RuntimeException (Aligner.java:41-43).
// Record contract and BUILD-DEFINED choices (best hit, X0/X1, MAPQ): identical to ubu_b200/facade.py, which documents them;
// the two write byte-identical SAM for the same input (tests/test_facade.py).
#pragma once
#include <string>

namespace ubu {

class Aligner {
 public:
  Aligner(const std::string& reference, int numThreads, int device = 0, int match = 1, int mismatch = -3, int gapOpen = 5,
          int gapExt = 2);
  ~Aligner();
  Aligner(const Aligner&) = delete;
  Aligner& operator=(const Aligner&) = delete;

  void align(const std::string& input, const std::string& outputSam);
  void shortAlign(const std::string& input, const std::string& outputSam);
  void index();
  // BUILD-DEFINED scalability switch (the reference hands candidate finding to bwa's index): k > 0 aligns a read (either strand) only
  // against the reference sequences it shares at least one exact k-mer with (k <= 31); a hit without any exact k-mer match is then not
  // reported. 0 (default) = every read x every sequence x both strands. Also read from the environment variable UBU_ALIGNER_SEED_K.
  void setSeedLength(int k) { seedK_ = k; }

 private:
  void run(const std::string& input, const std::string& outputSam);
  std::string reference_;
  int numThreads_, device_, match_, mismatch_, gapOpen_, gapExt_;
  int seedK_ = -1;              // -1: take UBU_ALIGNER_SEED_K (or 0)
  void* handle_ = nullptr;      // ubu_sw_handle*, created on first use
};

}  // namespace ubu
