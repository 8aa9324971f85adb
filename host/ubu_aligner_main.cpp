// ubu_aligner -- command-line face of host/aligner.hpp:   ubu_aligner {align|shortAlign} <reference.fa> <input.fq|fa> <output.sam> [threads]
// Exit code 0 on success; on failure prints the message and exits 1 (the reference's callers see a RuntimeException, Aligner.java:41-43).
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <string>

#include "aligner.hpp"

int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s {align|shortAlign} reference.fa input.{fq,fa} output.sam [threads]\n", argv[0]); return 2; }
  try {
    ubu::Aligner a(argv[2], argc > 5 ? atoi(argv[5]) : 1);
    a.index();
    if (std::string(argv[1]) == "align") a.align(argv[3], argv[4]); else a.shortAlign(argv[3], argv[4]);
  } catch (const std::exception& e) {
    fprintf(stderr, "ubu_aligner: %s\n", e.what());
    return 1;
  }
  return 0;
}
