/* ubu_sw_jni.c -- JNI shim between ubu's Java host code and libubu_sw.so (include/ubu_sw.h).
 *
 * UNTESTED: this image has no JDK (no jni.h), so this file has never been compiled here. It is shipped as
 * source, contains no logic beyond argument marshalling, and is NOT part of build()/tests.
 *
 * Build on a box with a JDK:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include \
 *       -o libubu_sw_jni.so ubu_sw_jni.c -L../../ubu_b200 -lubu_sw
 *
 * Java side: integration/jni/GpuAligner.java (drop-in for edu.unc.bioinf.ubu.assembly.Aligner,
 * src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:7-91).
 * A non-zero status is mapped to RuntimeException, which is what the reference throws when bwa exits
 * non-zero (Aligner.java:41-43). All buffers are direct ByteBuffers (pinned or not is the caller's choice).
 */
#include <jni.h>
#include <stdint.h>
#include "ubu_sw.h"

static void throw_status(JNIEnv* env, ubu_sw_handle* h, int rc) {
  char msg[320];
  snprintf(msg, sizeof msg, "ubu_sw failed: %s (%s)", ubu_sw_strerror(rc), h ? ubu_sw_last_error(h) : "");
  (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), msg);
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

static void fill(JNIEnv* env, ubu_sw_seqs* s, jobject packed, jobject off, jobject len, jobject nmask, jobject nmaskOff, jint count) {
  s->packed = (const uint8_t*)(*env)->GetDirectBufferAddress(env, packed);
  s->packed_bytes = (uint32_t)(*env)->GetDirectBufferCapacity(env, packed);
  s->off = (const uint32_t*)(*env)->GetDirectBufferAddress(env, off);
  s->len = (const uint16_t*)(*env)->GetDirectBufferAddress(env, len);
  s->nmask = nmask ? (const uint8_t*)(*env)->GetDirectBufferAddress(env, nmask) : NULL;
  s->nmask_bytes = nmask ? (uint32_t)(*env)->GetDirectBufferCapacity(env, nmask) : 0;
  s->nmask_off = nmaskOff ? (const uint32_t*)(*env)->GetDirectBufferAddress(env, nmaskOff) : NULL;
  s->count = (uint32_t)count;
}

/* returns the number of CIGAR ops written to cigarPool */
JNIEXPORT jlong JNICALL Java_edu_unc_bioinf_ubu_assembly_GpuAligner_alignBatch(
    JNIEnv* env, jclass cls, jlong handle,
    jobject rPacked, jobject rOff, jobject rLen, jobject rNmask, jobject rNmaskOff, jint nReads,
    jobject wPacked, jobject wOff, jobject wLen, jobject wNmask, jobject wNmaskOff, jint nWindows,
    jobject pairRefIdx, jobject results, jobject cigarPool) {
  ubu_sw_handle* h = (ubu_sw_handle*)(intptr_t)handle;
  ubu_sw_seqs reads, wins;
  fill(env, &reads, rPacked, rOff, rLen, rNmask, rNmaskOff, nReads);
  fill(env, &wins, wPacked, wOff, wLen, wNmask, wNmaskOff, nWindows);
  const uint32_t* idx = pairRefIdx ? (const uint32_t*)(*env)->GetDirectBufferAddress(env, pairRefIdx) : NULL;
  ubu_sw_result* out = (ubu_sw_result*)(*env)->GetDirectBufferAddress(env, results);
  uint32_t* pool = (uint32_t*)(*env)->GetDirectBufferAddress(env, cigarPool);
  size_t cap = (size_t)(*env)->GetDirectBufferCapacity(env, cigarPool) / 4, used = 0;
  int rc = ubu_sw_align_batch(h, &reads, &wins, idx, out, pool, cap, &used);
  if (rc) { throw_status(env, h, rc); return -1; }
  return (jlong)used;
}

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
  fill(env, &reads, rPacked, rOff, rLen, rNmask, rNmaskOff, nReads);
  ubu_kmer_params p = {(uint32_t)kmerSize, (uint32_t)minNodeFrequency};
  const uint32_t* region = readRegion ? (const uint32_t*)(*env)->GetDirectBufferAddress(env, readRegion) : NULL;
  size_t n_nodes = 0;
  int rc = ubu_kmer_build(h, &p, &reads, region, (size_t)maxNodes, (ubu_kmer_node*)(*env)->GetDirectBufferAddress(env, nodes),
                          (size_t)(*env)->GetDirectBufferCapacity(env, nodes) / sizeof(ubu_kmer_node), &n_nodes, NULL);
  if (rc) {
    char msg[320];
    snprintf(msg, sizeof msg, "ubu_kmer failed: %s (%s); nodes needed: %zu", ubu_sw_strerror(rc), ubu_kmer_last_error(h), n_nodes);
    (*env)->ThrowNew(env, (*env)->FindClass(env, "java/lang/RuntimeException"), msg);
    return 0;
  }
  return (jlong)n_nodes;
}
