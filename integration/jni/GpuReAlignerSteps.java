package edu.unc.bioinf.ubu.assembly;

/**
 * The record-level steps around the aligner as native calls (host/records.cpp through include/ubu_records.h): SAM text in, text
 * out, bug-compatible with the Java they stand in for. A maintainer swaps them in one at a time:
 *
 *   combineChimera(sam)                    CombineChimera3.combine                chimera/CombineChimera3.java:21-73
 *   cleanContigs(sam, minLen, removeClips) ReAligner.cleanAndOutputContigs        assembly/ReAligner.java:876-921 (removeSoftClips :926-961)
 *   adjustReads(orig, toContig, "100M")    ReAligner.adjustReads                  assembly/ReAligner.java:1048-1291
 *   sam2Fastq(sam)                         Sam2Fastq.convert                      fastq/Sam2Fastq.java:85-109
 *
 * IllegalStateException where the Java throws one, RuntimeException otherwise. UNTESTED (no JDK in the build image).
 */
public class GpuReAlignerSteps {
    static { System.loadLibrary("ubu_sw_jni"); }
    public static native String combineChimera(String sam);
    public static native String cleanContigs(String sam, int minContigLength, int removeSoftClips);
    public static native String adjustReads(String origSam, String toContigSam, String fullMatchCigar);
    public static native String sam2Fastq(String sam);
}
