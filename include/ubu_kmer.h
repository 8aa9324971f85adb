/* ubu_kmer.h -- C ABI of the B200 k-mer graph build for ubu's assembler (part of libubu_sw.so). SURVEY.md section 8 row f-4.
 *
 * Stands in for the graph-building half of the reference's Assembler:
 *     Assembler.addToGraph(String sequence)        src/main/java/edu/unc/bioinf/ubu/assembly/Assembler.java:452-474
 *     Node (count, toNodes, fromNodes, addReadSequence / hasMultipleUniqueReads)         assembly/Node.java:12-26,48-54,78-126
 *     Assembler.filterLowFrequencyNodes()          Assembler.java:201-227
 *     Assembler.identifyRootNodes()                Assembler.java:270-276
 *     Sequence (2-bit code, equality of k-mers)    assembly/Sequence.java:10-13,15-43,79-86
 * What stays on the host, as SURVEY.md section 8f says it should: the recursive contig enumeration
 * (Assembler.buildContigs / buildContig, Assembler.java:278-367), which walks the node table this call returns.
 *
 * One call builds the graphs of many regions at once (the reference assembles every target region with its own
 * Assembler instance, ReAligner.java:328-349): a node is a distinct (region, k-mer).
 *
 * Same conventions as ubu_sw.h: plain pointers and sizes, 0 on success / negative ubu_sw_status on failure, no
 * exceptions, a handle belongs to one host thread at a time, there is no CPU implementation behind these entry points.
 */
#ifndef UBU_KMER_H
#define UBU_KMER_H
#include "ubu_sw.h"
#ifdef __cplusplus
extern "C" {
#endif

#define UBU_KMER_MAX_K 64u
#define UBU_KMER_NONE 0xFFFFFFFFu

/* Assembler.setKmerSize (Assembler.java:165) and Assembler.setMinNodeFrequncy (:181). */
typedef struct { uint32_t kmer_size; uint32_t min_node_frequency; } ubu_kmer_params;

/* Node.hasMultipleUniqueReads() (Node.java:92-94): at least two DIFFERENT read strings contain the k-mer. */
#define UBU_KMER_FLAG_MULTI_READS 1u
/* filterLowFrequencyNodes would remove the node: count < min_node_frequency or !hasMultipleUniqueReads (Assembler.java:204-210). */
#define UBU_KMER_FLAG_FILTERED 2u
/* survives the filter and no predecessor does: a member of Assembler.rootNodes (Assembler.java:270-276, Node.java:128-130). */
#define UBU_KMER_FLAG_ROOT 4u

/* One node of the graph as addToGraph leaves it (before the filter), 64 bytes. Nodes come back in the order the
 * reference inserts them into Assembler.nodes: by (first_read, first_pos) of their first occurrence.
 * to_node / from_node list the successors / predecessors in the order Node.addToNode created the edges
 * (Node.toNodes / Node.fromNodes array order, Node.java:100-126), as indices into the returned table, UBU_KMER_NONE-padded.
 * The lists are the UNFILTERED ones; a consumer that wants the graph after filterLowFrequencyNodes drops the entries whose
 * target carries UBU_KMER_FLAG_FILTERED (Node.remove keeps the order of the others, Node.java:143-171). */
typedef struct {
  uint64_t kmer[2];        /* 2 bits per base, base j at bits 2j..2j+1 (Sequence.java:31-41), A=0 C=1 T=2 G=3 */
  uint32_t region;
  uint32_t count;          /* Node.count: occurrences over all reads and positions */
  uint32_t first_read;     /* index (in the call's read set) and position of the first occurrence */
  uint16_t first_pos;
  uint16_t flags;          /* UBU_KMER_FLAG_* */
  uint32_t to_node[4];
  uint32_t from_node[4];
} ubu_kmer_node;

typedef struct ubu_kmer_handle ubu_kmer_handle;

/* Fails with UBU_SW_ERR_NO_DEVICE off an sm_100 GPU. */
int ubu_kmer_create(int device, ubu_kmer_handle** out);
void ubu_kmer_destroy(ubu_kmer_handle* h);
const char* ubu_kmer_last_error(ubu_kmer_handle* h);

/* Builds the k-mer graphs of all regions of one batch of reads.
 *   reads        2-bit packed reads (ubu_sw_seqs). A read with any N-plane bit set is skipped, like the reference skips reads
 *                whose string contains 'N' (Assembler.java:77-81); the caller applies the X0 > 1 test of the same lines
 *                (it is a SAM tag, not sequence data) by passing UBU_KMER_NONE as that read's region.
 *   read_region  region of every read (NULL = one region, 0); UBU_KMER_NONE = skip the read.
 *   max_nodes    upper bound on distinct (region, k-mer) nodes, 0 = the number of k-mer occurrences (always enough).
 *   out/out_cap  node table; *n_nodes = number of nodes. UBU_SW_ERR_CAPACITY with *n_nodes = required count if out_cap
 *                is too small or max_nodes was exceeded.
 *   n_skipped    (optional) reads skipped for N. */
int ubu_kmer_build(ubu_kmer_handle* h, const ubu_kmer_params* params, const ubu_sw_seqs* reads, const uint32_t* read_region,
                   size_t max_nodes, ubu_kmer_node* out, size_t out_cap, size_t* n_nodes, uint32_t* n_skipped);

/* The same work in steps, so that a benchmark can time the device-resident part: stage = validate + upload,
 * run = all kernels (may be repeated), fetch = download. ms: [0] whole run, [1] table clear + read classes,
 * [2] insert + edges (the hash-table kernel), [3] order + emit. */
int ubu_kmer_stage(ubu_kmer_handle* h, const ubu_kmer_params* params, const ubu_sw_seqs* reads, const uint32_t* read_region, size_t max_nodes);
int ubu_kmer_run(ubu_kmer_handle* h);
int ubu_kmer_fetch(ubu_kmer_handle* h, ubu_kmer_node* out, size_t out_cap, size_t* n_nodes, uint32_t* n_skipped);
int ubu_kmer_last_run_ms(ubu_kmer_handle* h, float ms[4]);
/* k-mer occurrences of the staged batch (the unit the throughput is quoted in) and bytes of the node hash table. */
int ubu_kmer_last_run_stats(ubu_kmer_handle* h, uint64_t* occurrences, uint64_t* table_bytes, uint32_t* launches);

#ifdef __cplusplus
}
#endif
#endif
