/* ubu_sw.h -- C ABI of the B200 Smith-Waterman realignment path (libubu_sw.so).
 *
 * Drop-in boundary for the reference's aligner facade. The reference side of this
 * boundary is path-in/path-out and forks `bwa`:
 *     new Aligner(String reference, int numThreads)   src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:13-16
 *     void align(String input, String outputSam)      Aligner.java:18-22   (bwa bwasw)
 *     void shortAlign(String input, String outputSam) Aligner.java:46-56   (bwa aln + samse)
 *     void index()                                    Aligner.java:62-69   (bwa index; no-op for window SW)
 * called from ReAligner.alignAndCleanContigs (ReAligner.java:351-367) and
 * ReAligner.alignToContigs (ReAligner.java:963-975). Nothing like a batch ABI exists in
 * the reference, so the entry points below are BUILD-DEFINED; each
 * comment names the reference interface it stands in for. INTEGRATION.md shows the JNI
 * stub a maintainer would add under Aligner.
 *
 * Alignment semantics: SPEC.md (BUILD-DEFINED; reference parity UNPINNED).
 * No torch / CUDA types cross this boundary: plain pointers and sizes only.
 * Error convention: 0 on success, negative ubu_sw_status on failure, never throws or aborts
 * (the reference throws RuntimeException on a non-zero bwa exit, Aligner.java:41-43; the JNI
 * shim maps a non-zero status to the same exception).
 * Threading: a handle is owned by one host thread at a time; distinct handles (one per GPU)
 * are fully independent (the reference delegates parallelism with `-t numThreads`, Aligner.java:19,49).
 */
#ifndef UBU_SW_H
#define UBU_SW_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define UBU_SW_VERSION 0x00020000u   /* round 2: multi-GPU driver, long reads, ubu_sw_sam_names.cigar_pool_n */

typedef enum {
  UBU_SW_OK = 0,
  UBU_SW_ERR_ARG = -1,          /* NULL pointer, bad parameter range, bad index          */
  UBU_SW_ERR_CUDA = -2,         /* CUDA runtime failure; ubu_sw_last_error() has the text */
  UBU_SW_ERR_CIGAR_POOL = -3,   /* cigar_pool_cap too small; *cigar_pool_used = required  */
  UBU_SW_ERR_TOO_LONG = -4,     /* read longer than UBU_SW_MAX_READ, or a long read whose matrix exceeds UBU_SW_MAX_LONG_CELLS */
  UBU_SW_ERR_NO_DEVICE = -5,    /* no CUDA device / not an sm_100 device                  */
  UBU_SW_ERR_BUSY = -6,         /* no free pipeline slot / ticket not pending             */
  UBU_SW_ERR_NOMEM = -7,
  UBU_SW_ERR_CAPACITY = -8      /* an output table / node bound was too small; the required count is written back */
} ubu_sw_status;

#define UBU_SW_MAX_READ   4096u     /* reads up to 256 bp run on the packed 16-bit kernels (all BASELINE configs: 50..250 bp); longer   */
                                    /* ones (assembled contigs, Aligner.align's queries: 100 bp .. 2.5 kb) on the 32-bit warp-per-task    */
                                    /* kernel, as long as read x window <= UBU_SW_MAX_LONG_CELLS                                          */
#define UBU_SW_MAX_LONG_CELLS (1u << 26)
#define UBU_SW_MAX_WINDOW 65535u    /* windows are regions / contigs, not chromosomes: lengths are 16-bit in ubu_sw_seqs                  */

/* SPEC.md section 1. Defaults +1 / -3 / 5 / 2. Stands in for bwa's (implicit, default) scoring flags. */
typedef struct { int32_t match, mismatch, gap_open, gap_ext; } ubu_sw_params;

/* SPEC.md section 6. Replaces the SAM fields the callers read back from bwa's output
 * (POS, CIGAR, NM: ReAligner.java:1088-1143). 32 bytes. */
typedef struct {
  int32_t  score;
  int32_t  ref_start, ref_end;     /* 1-based inclusive, within the window */
  int32_t  q_start, q_end;         /* 1-based inclusive, within the read   */
  int32_t  edit_distance;          /* = NM (ReAligner.java:282-285)        */
  uint32_t cigar_off;              /* first op in the CIGAR pool           */
  uint16_t cigar_n;                /* number of ops                        */
  uint16_t flags;                  /* UBU_SW_FLAG_*                        */
} ubu_sw_result;

#define UBU_SW_FLAG_UNALIGNED 1u
#define UBU_SW_CIGAR_M 0u
#define UBU_SW_CIGAR_I 1u
#define UBU_SW_CIGAR_D 2u
#define UBU_SW_CIGAR_S 4u

/* A set of 2-bit packed sequences (Sequence.java:10-13,31-41: A=0 C=1 T=2 G=3, 4 bases/byte, LSB first).
 * Sequence k occupies ceil(len[k]/4) bytes starting at packed + off[k].
 * nmask (optional, NULL = no N anywhere): 1 bit per base, LSB first, sequence k at nmask + nmask_off[k]. */
typedef struct {
  const uint8_t*  packed;
  const uint32_t* off;
  const uint16_t* len;
  const uint8_t*  nmask;
  const uint32_t* nmask_off;
  uint32_t        count;
  uint32_t        packed_bytes;   /* total bytes readable at packed */
  uint32_t        nmask_bytes;    /* total bytes readable at nmask (0 if NULL) */
} ubu_sw_seqs;

typedef struct ubu_sw_handle ubu_sw_handle;

/* Stands in for `new Aligner(reference, numThreads)` (Aligner.java:13-16): one handle per GPU.
 * params == NULL -> SPEC defaults. slots = depth of the H2D/compute/D2H pipeline (1..8; 0 -> 4). */
int ubu_sw_create(const ubu_sw_params* params, int device, int slots, ubu_sw_handle** out);
void ubu_sw_destroy(ubu_sw_handle* h);

/* Stands in for Aligner.shortAlign / Aligner.align (Aligner.java:18-22,46-56): align read k against
 * window pair_ref_idx[k] (NULL = identity, then reads->count must equal windows->count).
 * out[k] and the CIGAR slice it names describe read k: results are in INPUT order.
 * Synchronous: host->device copies, all kernels and device->host copies complete before return.
 * A large batch (>= 131072 reads, sequences laid out in ascending order, every stretch of reads using a compact stretch of
 * windows -- what coordinate-sorted input gives) is cut into chunks that travel through the handle's `slots` internally, so
 * the upload of one chunk, the kernels of the next and the download of a third overlap inside this one call; pinned host
 * buffers (ubu_sw_host_alloc) make those copies truly asynchronous. Other batches go through in one shot. */
int ubu_sw_align_batch(ubu_sw_handle* h, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows,
                       const uint32_t* pair_ref_idx, ubu_sw_result* out,
                       uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used);

/* Pipelined form of the same call: submit() plans the batch, enqueues copies and kernels on the next
 * free slot's stream and returns a ticket; wait() blocks until that batch's results are in `out`.
 * All caller buffers must stay valid until wait() returns; use ubu_sw_host_alloc for overlap. */
int ubu_sw_submit(ubu_sw_handle* h, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows,
                  const uint32_t* pair_ref_idx, ubu_sw_result* out,
                  uint32_t* cigar_pool, size_t cigar_pool_cap, int* ticket);
int ubu_sw_wait(ubu_sw_handle* h, int ticket, size_t* cigar_pool_used);

/* Staged form used by bench.py's device-resident timing (`value`): stage() = plan + H2D into slot
 * `slot`; run() = enqueue every kernel of the path on the slot's stream (inputs and results stay in HBM);
 * sync() = wait for the slot's stream; fetch() = D2H of results + CIGARs of the last run(). */
int ubu_sw_stage(ubu_sw_handle* h, int slot, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows,
                 const uint32_t* pair_ref_idx);
int ubu_sw_run(ubu_sw_handle* h, int slot);
int ubu_sw_sync(ubu_sw_handle* h, int slot);
int ubu_sw_fetch(ubu_sw_handle* h, int slot, ubu_sw_result* out,
                 uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used);
/* Device time in ms of the last run() on the slot (CUDA events on the slot's stream): [0] whole run,
 * [1] forward (score) kernels, [2] traceback + CIGAR kernels. Valid after sync(). A large mixed-length batch
 * runs the traceback of its earlier forward launches underneath the later ones: [1] is then
 * the span of the forward launches with that traceback underneath, [2] the traceback left after them. */
int ubu_sw_last_run_ms(ubu_sw_handle* h, int slot, float ms[3]);
/* Number of kernel launches issued by the last run() on the slot. */
int ubu_sw_last_run_launches(ubu_sw_handle* h, int slot);
/* Diagnostics: device counters of the last run() on the slot (valid after run): out[0..9] = tasks per traceback class (0 gap-free scan,
 * 1-3 register bands of 8/16/32 diagonals, 4-8 strip classes by band width, 9 scalar last resort), out[10] = CIGAR ops produced,
 * out[11] = error flags, out[12] = strip work units drawn. */
int ubu_sw_debug_counters(ubu_sw_handle* h, int slot, uint32_t out[16]);

/* ---- one process, several GPUs (BASELINE.json configs[4]) -------------------------------------------
 * The reference hands parallelism to `bwa -t numThreads` (Aligner.java:19,49) and treats its inputs as independent work
 * (ReAligner.java:180-184,199-200). Reads shard trivially: there is no collective and no peer traffic. The driver owns one
 * handle and one host thread per device; the batch is cut into input-order chunks of `chunk_reads` reads, chunk c runs on
 * device c mod N through that handle's slots. Nothing is merged on the host afterwards: every chunk owns a slice of the
 * caller's CIGAR pool proportional to its reads (slice c starts at cigar_pool_cap * first_read(c) / reads), the kernels
 * write final cigar_off values, and records and ops are copied by the GPUs straight into their input-order places.
 * Consequences for the caller: out[k] / cigar slices are in input order exactly as with ubu_sw_align_batch, but the pool is
 * not dense (*cigar_pool_used = its high-water mark); when a slice is too small the call returns UBU_SW_ERR_CIGAR_POOL and
 * *cigar_pool_used = a capacity with which every slice fits. Pinned buffers (ubu_sw_host_alloc) make the copies asynchronous. */
typedef struct ubu_sw_multi ubu_sw_multi;
/* devices == NULL: the first n_devices visible devices (n_devices <= 0: all). slots as in ubu_sw_create; chunk_reads 0 -> 65536. */
int ubu_sw_multi_create(const ubu_sw_params* params, const int* devices, int n_devices, int slots, uint32_t chunk_reads, ubu_sw_multi** out);
void ubu_sw_multi_destroy(ubu_sw_multi* m);
int ubu_sw_multi_device_count(const ubu_sw_multi* m);
const char* ubu_sw_multi_last_error(ubu_sw_multi* m);
int ubu_sw_multi_align(ubu_sw_multi* m, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx,
                       ubu_sw_result* out, uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used);

/* ---- the steps right behind the aligner; bug-compatible with the Java they replace ---- */

/* Long reference sequences (chromosomes), same 2-bit code + N plane; replaces the per-chromosome StringBuffer that
 * CompareToReference.loadSeqRef rebuilds for every sequence (CompareToReference.java:125-149): uploaded once, kept in HBM. */
typedef struct {
  const uint8_t*  packed;
  const uint64_t* off;          /* byte offset of sequence k in packed */
  const uint32_t* len;          /* bases */
  const uint8_t*  nmask;        /* optional */
  const uint64_t* nmask_off;
  uint32_t        count;
  uint64_t        packed_bytes, nmask_bytes;
} ubu_sw_refs;
int ubu_sw_load_refs(ubu_sw_handle* h, const ubu_sw_refs* refs);

#define UBU_SW_POST_NULL      1u   /* the reference returns a null read (ReAligner.java:1362-1364)                     */
#define UBU_SW_POST_MALFORMED 2u   /* the reference throws (IllegalStateException / index out of bounds) or bad index  */
#define UBU_SW_POST_POOL      4u   /* output CIGAR pool too small                                                      */

/* An aligned read: reference sequence, 1-based POS, CIGAR slice (BAM ops, N = 3 allowed), MAPQ and the YQ / YM tags. */
typedef struct { uint32_t ref_idx; int32_t pos; uint32_t cigar_off; uint16_t cigar_n; int16_t mapq; int16_t yq; int16_t ym; } ubu_sw_placed;
typedef struct { int32_t xm, nm, mapq; uint32_t flags; } ubu_sw_recount;
/* Replaces the per-read body of ReAligner.updateMismatchAndEditDistance (ReAligner.java:279-287): XM = numDifferences
 * (CompareToReference.java:83-123), NM = XM + indel bases (:316-326), MAPQ = calcMappingQuality (:302-314).
 * reads->count records; read k is described by placed[k]. Needs ubu_sw_load_refs first. */
int ubu_sw_recount_batch(ubu_sw_handle* h, const ubu_sw_seqs* reads, const ubu_sw_placed* placed,
                         const uint32_t* cigar_pool, size_t cigar_pool_n, ubu_sw_recount* out);

/* A contig's own alignment to the genome (POS + CIGAR slice) and an ungapped placement of a read inside a contig. */
typedef struct { int32_t pos; uint32_t cigar_off; uint32_t cigar_n; } ubu_sw_contig;
typedef struct { uint32_t contig_idx; int32_t position /* 0-based offset in the contig (ReadPosition) */; uint32_t read_len; } ubu_sw_proj_in;
typedef struct { int32_t pos; uint32_t cigar_off; uint16_t cigar_n; uint16_t flags; } ubu_sw_proj_out;
/* Replaces ReAligner.updateReadAlignment (ReAligner.java:1293-1368) with ReadBlock's CIGAR algebra
 * (ReadBlock.java:74-175): genome POS + CIGAR of each read, projected through its contig's CIGAR. */
int ubu_sw_project_batch(ubu_sw_handle* h, const ubu_sw_contig* contigs, uint32_t n_contigs,
                         const uint32_t* contig_cigar_pool, size_t contig_cigar_pool_n,
                         const ubu_sw_proj_in* in, uint32_t n, ubu_sw_proj_out* out,
                         uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used);

/* Pinned host memory for the caller's buffers (cudaHostAlloc); plain malloc'd buffers also work, slower. */
void* ubu_sw_host_alloc(size_t bytes);
void  ubu_sw_host_free(void* p);

/* Host-side helpers on the reference's sequence conventions (no GPU needed). */
/* text -> 2-bit codes + N bit-plane (Sequence.java:31-41; non-ACGT -> N, case-folded). */
void ubu_sw_pack(const char* text, size_t n, uint8_t* packed, uint8_t* nmask, int* has_n);
/* The same for a whole batch on `threads` host threads (0 = all cores), 32 (AVX2) or 8 characters per step: sequence k is the len[k]
 * characters at text + text_off[k] and is written to packed + off[k] (and nmask + nmask_off[k]; nmask may be NULL, non-ACGT
 * characters are then packed as A and only counted). Output ranges must not overlap. *n_with_n = sequences holding a non-ACGT
 * character. This is the packer of a pipeline that aligns a million reads in ~6 ms. */
int ubu_sw_pack_batch(const char* text, const uint64_t* text_off, const uint16_t* len, uint32_t n, uint8_t* packed, const uint32_t* off,
                      uint8_t* nmask, const uint32_t* nmask_off, int threads, uint32_t* n_with_n);
/* SAM text for a batch of results, the last step of the ordered host merge: one line per read in input order with the fields the
 * reference's consumers read back from the aligner's SAM (QNAME FLAG RNAME POS MAPQ CIGAR SEQ, NM:i, AS:i; ReAligner.java:1088-1143).
 * QNAME = qname_prefix + (first_index + k); RNAME = rnames[window] (or "w<window>"), POS = window_pos0[window] + ref_start. */
typedef struct {
  const char*        qname_prefix;   /* NULL = "r" */
  uint64_t           first_index;    /* index of read 0 of this batch in the whole run */
  const uint32_t*    pair_ref_idx;   /* window of every read (NULL = read k <-> window k) */
  const char* const* rnames;         /* per window, NULL = "w<index>" */
  const int64_t*     window_pos0;    /* 0-based start of every window on its reference sequence, NULL = 0 */
  uint64_t           cigar_pool_n;   /* ops readable at cigar_pool; a record whose slice leaves the pool fails the call with UBU_SW_ERR_ARG
                                      * (0 = unknown: slices are not checked) */
} ubu_sw_sam_names;
/* Formats on `threads` host threads (0 = all cores) and writes the blocks to file descriptor fd in order (fd < 0: format only). */
int ubu_sw_sam_write(int fd, const ubu_sw_seqs* reads, const ubu_sw_result* res, const uint32_t* cigar_pool, uint32_t n,
                     const ubu_sw_sam_names* names, int threads, uint64_t* bytes_written);
/* ubu_sw_multi_align followed by the SAM text of the batch (the ordered host merge of BASELINE.json configs[4]): finished chunks are
 * released in input order, formatted on sam_threads helper threads (0 = all cores) and written to fd (fd < 0: format only) while
 * later chunks are still on the GPUs. */
int ubu_sw_multi_align_sam(ubu_sw_multi* m, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx,
                           ubu_sw_result* out, uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used,
                           int fd, const ubu_sw_sam_names* names, int sam_threads, uint64_t* bytes_written);
/* ReverseComplementor.java:38-62: A<->T, C<->G, every other byte unchanged. */
void ubu_sw_revcomp(const char* in, size_t n, char* out);
/* BAM-encoded ops -> "76M" style text; returns chars written (excluding NUL), "*" for n == 0. */
int  ubu_sw_cigar_string(const uint32_t* ops, uint32_t n, char* buf, size_t cap);

/* Diagnostics, no GPU needed: the host plan of a batch (which forward launches it would get, in which order its tasks would
 * run). launches_out: 6 ints per launch = {rows of the length class, has_n, first slot, slots, widest window, shared_window};
 * order_out[slot] = task (0xFFFFFFFF = padding). UBU_SW_ERR_CAPACITY if a buffer is too small (counts are still written). */
int ubu_sw_debug_plan(const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx, uint32_t* order_out,
                      size_t order_cap, uint32_t* n_slots, int* identity, int32_t* launches_out, int launches_cap, int* n_launches);

int          ubu_sw_device_count(void);
unsigned     ubu_sw_version(void);
const char*  ubu_sw_strerror(int status);
const char*  ubu_sw_last_error(ubu_sw_handle* h);   /* text of the last CUDA failure on this handle */

#ifdef __cplusplus
}
#endif
#endif /* UBU_SW_H */
