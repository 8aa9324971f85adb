/* ubu_sw_jni.c -- JNI shim between ubu's Java host code and libubu_sw.so / libubu_host.so (include/ubu_sw.h, include/ubu_records.h).
 *
 * UNTESTED: this image has no JDK (no jni.h), so this file has never been compiled against a real JNI. tests/test_abi.py
 * syntax-checks it against a stand-in jni.h (tests/mock_jni/jni.h). It contains argument marshalling and validation only.
 *
 * Build on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       -o libubu_sw_jni.so ubu_sw_jni.c -L../../ubu_b200 -lubu_sw -L../../host -lubu_host
 *
 * Java side: integration/jni/GpuAligner.java (drop-in for edu.unc.bioinf.ubu.assembly.Aligner,
 * src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:7-91) and GpuReAlignerSteps.java (the record-level steps).
 * A non-zero status is mapped to RuntimeException, which is what the reference throws when bwa exits non-zero
 * (Aligner.java:41-43); a buffer that is not direct or too small for the counts it is passed with is an
 * IllegalArgumentException BEFORE anything native touches it.
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>
#include "ubu_sw.h"
#include "ubu_records.h"

static void throw_status(JNIEnv* env, ubu_sw_handle* h, int rc) {
  char msg[320];
  snprintf(msg, sizeof msg, "ubu_sw failed: %s (%s)", ubu_sw_strerror(rc), h ? ubu_sw_last_error(h) : "");
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), msg);
}
static void throw_arg(JNIEnv* env, const char* what) {
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/IllegalArgumentException"), what);
}
static void throw_rec(JNIEnv* env, int rc) {
  char msg[400];
  snprintf(msg, sizeof msg, "ubu_rec failed (%d): %s", rc, ubu_rec_last_error());
  (*env)->ThrowNew(env, (*env)->FindClass(env, rc == UBU_REC_ERR_STATE ? "java/lang/IllegalStateException" : "java/lang/RuntimeException"), msg);
}

/* address of a direct buffer that must hold at least `need` bytes; NULL (with an exception pending) otherwise */
static void* direct(JNIEnv* env, jobject buf, size_t need, const char* name, size_t* cap_out) {
  if (!buf) { throw_arg(env, name); return NULL; }
  void* p = (*env)->GetDirectBufferAddress(env, buf);
  const jlong cap = (*env)->GetDirectBufferCapacity(env, buf);
  if (!p || cap < 0 || (size_t)cap < need) { throw_arg(env, name); return NULL; }
  if (cap_out) *cap_out = (size_t)cap;
  return p;
}

JNIEXPORT jlong JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuAligner_create(JNIEnv* env, jclass cls, jint match, jint mismatch,
                                                                          jint gapOpen, jint gapExt, jint device, jint slots) {
  ubu_sw_params p = {match, mismatch, gapOpen, gapExt};
  ubu_sw_handle* h = NULL;
  int rc = ubu_sw_create(&p, device, slots, &h);
  if (rc) { throw_status(env, NULL, rc); return 0; }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuAligner_destroy(JNIEnv* env, jclass cls, jlong handle) {
  ubu_sw_destroy((ubu_sw_handle*)(intptr_t)handle);
}

/* 0 on success; an IllegalArgumentException is pending otherwise */
static int fill(JNIEnv* env, ubu_sw_seqs* s, jobject packed, jobject off, jobject len, jobject nmask, jobject nmaskOff, jint count) {
  size_t cap = 0;
  memset(s, 0, sizeof *s);
  if (count < 0) { throw_arg(env, "negative sequence count"); return -1; }
  s->count = (uint32_t)count;
  s->packed = (const uint8_t*)direct(env, packed, 0, "packed: not a direct buffer", &cap); if (!s->packed) return -1;
  if (cap > 0xFFFFFFFFu) { throw_arg(env, "packed: larger than 4 GiB, split the batch"); return -1; }
  s->packed_bytes = (uint32_t)cap;
  s->off = (const uint32_t*)direct(env, off, (size_t)count * 4, "off: needs 4 bytes per sequence", NULL); if (!s->off) return -1;
  s->len = (const uint16_t*)direct(env, len, (size_t)count * 2, "len: needs 2 bytes per sequence", NULL); if (!s->len) return -1;
  if (nmask) {
    s->nmask = (const uint8_t*)direct(env, nmask, 0, "nmask: not a direct buffer", &cap); if (!s->nmask) return -1;
    s->nmask_bytes = (uint32_t)cap;
    s->nmask_off = (const uint32_t*)direct(env, nmaskOff, (size_t)count * 4, "nmaskOff: needs 4 bytes per sequence", NULL); if (!s->nmask_off) return -1;
  }
  return 0;
}

/* returns the number of CIGAR ops written to cigarPool (ubu_sw_align_batch) */
JNIEXPORT jlong JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuAligner_alignBatch(
    JNIEnv* env, jclass cls, jlong handle,
    jobject rPacked, jobject rOff, jobject rLen, jobject rNmask, jobject rNmaskOff, jint nReads,
    jobject wPacked, jobject wOff, jobject wLen, jobject wNmask, jobject wNmaskOff, jint nWindows,
    jobject pairRefIdx, jobject results, jobject cigarPool) {
  ubu_sw_handle* h = (ubu_sw_handle*)(intptr_t)handle;
  ubu_sw_seqs reads, wins;
  size_t pool_bytes = 0;
  if (!h) { throw_arg(env, "closed aligner"); return -1; }
  if (fill(env, &reads, rPacked, rOff, rLen, rNmask, rNmaskOff, nReads)) return -1;
  if (fill(env, &wins, wPacked, wOff, wLen, wNmask, wNmaskOff, nWindows)) return -1;
  const uint32_t* idx = NULL;
  if (pairRefIdx) { idx = (const uint32_t*)direct(env, pairRefIdx, (size_t)nReads * 4, "pairRefIdx: needs 4 bytes per read", NULL); if (!idx) return -1; }
  ubu_sw_result* out = (ubu_sw_result*)direct(env, results, (size_t)nReads * sizeof(ubu_sw_result), "results: needs 32 bytes per read", NULL);
  if (!out) return -1;
  uint32_t* pool = (uint32_t*)direct(env, cigarPool, 0, "cigarPool: not a direct buffer", &pool_bytes);
  if (!pool) return -1;
  size_t used = 0;
  int rc = ubu_sw_align_batch(h, &reads, &wins, idx, out, pool, pool_bytes / 4, &used);
  if (rc) { throw_status(env, h, rc); return -1; }
  return (jlong)used;
}

/* ---- the reference's own entry points, path in / path out (Aligner.java:18-22, 46-56, 62-69) ---- */
static void aligner_call(JNIEnv* env, jstring reference, jint numThreads, jint device, jstring input, jstring outputSam, int which) {
  const char* ref = (*env)->GetStringUTFChars(env, reference, NULL);
  const char* in = input ? (*env)->GetStringUTFChars(env, input, NULL) : NULL;
  const char* out = outputSam ? (*env)->GetStringUTFChars(env, outputSam, NULL) : NULL;
  int rc = which == 0 ? ubu_aligner_align(ref, numThreads, device, in, out)
         : which == 1 ? ubu_aligner_short_align(ref, numThreads, device, in, out) : ubu_aligner_index(ref);
  (*env)->ReleaseStringUTFChars(env, reference, ref);
  if (in) (*env)->ReleaseStringUTFChars(env, input, in);
  if (out) (*env)->ReleaseStringUTFChars(env, outputSam, out);
  if (rc) throw_rec(env, rc);
}
JNIEXPORT void JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuAligner_alignFiles(JNIEnv* env, jclass cls, jstring reference, jint numThreads, jint device,
                                                                             jstring input, jstring outputSam) { aligner_call(env, reference, numThreads, device, input, outputSam, 0); }
JNIEXPORT void JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuAligner_shortAlignFiles(JNIEnv* env, jclass cls, jstring reference, jint numThreads, jint device,
                                                                                  jstring input, jstring outputSam) { aligner_call(env, reference, numThreads, device, input, outputSam, 1); }
JNIEXPORT void JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuAligner_indexFile(JNIEnv* env, jclass cls, jstring reference) { aligner_call(env, reference, 1, 0, NULL, NULL, 2); }

/* ---- record-level steps (include/ubu_records.h): SAM text in, text out ---- */
static jstring text_call(JNIEnv* env, jstring a, jstring b, jstring c, int which, jint i1, jint i2) {
  const char* sa = (*env)->GetStringUTFChars(env, a, NULL);
  const char* sb = b ? (*env)->GetStringUTFChars(env, b, NULL) : NULL;
  const char* sc = c ? (*env)->GetStringUTFChars(env, c, NULL) : NULL;
  char* out = NULL; int has = 0; int32_t stats[3];
  int rc = which == 0 ? ubu_rec_combine_chimera(sa, &out)
         : which == 1 ? ubu_rec_clean_contigs(sa, i1, i2, &out, &has)
         : which == 2 ? ubu_rec_adjust_reads(sa, sb, sc, &out, stats) : ubu_rec_sam2fastq(sa, &out);
  (*env)->ReleaseStringUTFChars(env, a, sa);
  if (sb) (*env)->ReleaseStringUTFChars(env, b, sb);
  if (sc) (*env)->ReleaseStringUTFChars(env, c, sc);
  if (rc) { throw_rec(env, rc); return NULL; }
  jstring r = (*env)->NewStringUTF(env, out);
  ubu_rec_free(out);
  return r;
}
JNIEXPORT jstring JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuReAlignerSteps_combineChimera(JNIEnv* env, jclass cls, jstring sam) { return text_call(env, sam, NULL, NULL, 0, 0, 0); }
JNIEXPORT jstring JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuReAlignerSteps_cleanContigs(JNIEnv* env, jclass cls, jstring sam, jint minContigLength, jint removeSoftClips) {
  return text_call(env, sam, NULL, NULL, 1, minContigLength, removeSoftClips);
}
JNIEXPORT jstring JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuReAlignerSteps_adjustReads(JNIEnv* env, jclass cls, jstring origSam, jstring toContigSam, jstring fullMatchCigar) {
  return text_call(env, origSam, toContigSam, fullMatchCigar, 2, 0, 0);
}
JNIEXPORT jstring JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuReAlignerSteps_sam2Fastq(JNIEnv* env, jclass cls, jstring sam) { return text_call(env, sam, NULL, NULL, 3, 0, 0); }

/* ---- k-mer graph build (include/ubu_kmer.h), Java side: integration/jni/GpuKmerGraph.java. UNTESTED like the rest of this file. ---- */
#include "ubu_kmer.h"

JNIEXPORT jlong JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuKmerGraph_create(JNIEnv* env, jclass cls, jint device) {
  ubu_kmer_handle* h = NULL;
  int rc = ubu_kmer_create(device, &h);
  if (rc) { throw_status(env, NULL, rc); return 0; }
  return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuKmerGraph_destroy(JNIEnv* env, jclass cls, jlong handle) {
  ubu_kmer_destroy((ubu_kmer_handle*)(intptr_t)handle);
}

JNIEXPORT jlong JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuKmerGraph_build(JNIEnv* env, jclass cls, jlong handle, jint kmerSize,
    jint minNodeFrequency, jobject rPacked, jobject rOff, jobject rLen, jobject rNmask, jobject rNmaskOff, jint nReads,
    jobject readRegion, jlong maxNodes, jobject nodes) {
  ubu_kmer_handle* h = (ubu_kmer_handle*)(intptr_t)handle;
  ubu_sw_seqs reads;
  size_t node_bytes = 0;
  if (!h) { throw_arg(env, "closed graph builder"); return 0; }
  if (fill(env, &reads, rPacked, rOff, rLen, rNmask, rNmaskOff, nReads)) return 0;
  ubu_kmer_params p = {(uint32_t)kmerSize, (uint32_t)minNodeFrequency};
  const uint32_t* region = NULL;
  if (readRegion) { region = (const uint32_t*)direct(env, readRegion, (size_t)nReads * 4, "readRegion: needs 4 bytes per read", NULL); if (!region) return 0; }
  ubu_kmer_node* out = (ubu_kmer_node*)direct(env, nodes, 0, "nodes: not a direct buffer", &node_bytes);
  if (!out) return 0;
  size_t n_nodes = 0;
  int rc = ubu_kmer_build(h, &p, &reads, region, (size_t)maxNodes, out, node_bytes / sizeof(ubu_kmer_node), &n_nodes, NULL);
  if (rc) {
    char msg[320];
    snprintf(msg, sizeof msg, "ubu_kmer failed: %s (%s); nodes needed: %zu", ubu_sw_strerror(rc), ubu_kmer_last_error(h), n_nodes);
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), msg);
    return 0;
  }
  return (jlong)n_nodes;
}
