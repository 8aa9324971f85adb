package edu.unc.bioinf.ubu.assembly;

import java.io.IOException;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;

/**
 * Drop-in for {@link Aligner} (src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java): same constructor, same three methods, same
 * path-in / path-out contract, same RuntimeException on failure (Aligner.java:41-43). The three `bwa` forks are replaced by calls
 * into libubu_host.so / libubu_sw.so through integration/jni/ubu_sw_jni.c:
 *
 *   align(input, outputSam)        Aligner.java:18-22  (bwa bwasw)          -> ubu_aligner_align
 *   shortAlign(input, outputSam)   Aligner.java:46-56  (bwa aln + samse)    -> ubu_aligner_short_align
 *   index()                        Aligner.java:62-69  (bwa index)          -> ubu_aligner_index (checks the FASTA; window SW needs no index)
 *
 * The two callers (ReAligner.java:353-355, 969-974) compile unchanged against this class after `s/new Aligner(/new GpuAligner(/`.
 * Callers that already hold packed reads (a future in-memory pipeline) can use alignBatch with direct buffers instead.
 *
 * UNTESTED: no JDK exists in the image this repo was built in; see INTEGRATION.md.
 */
public class GpuAligner {
    static { System.loadLibrary("ubu_sw_jni"); }

    private final String reference;      // FASTA of windows / contigs, as today (Aligner.java:9)
    private final int numThreads;        // kept for the signature; one GPU replaces bwa's -t threads
    private final int device;

    public GpuAligner(String reference, int numThreads) { this(reference, numThreads, 0); }          // Aligner.java:13-16
    public GpuAligner(String reference, int numThreads, int device) { this.reference = reference; this.numThreads = numThreads; this.device = device; }

    public void align(String input, String outputSam) throws IOException, InterruptedException {        // Aligner.java:18-22
        alignFiles(reference, numThreads, device, input, outputSam);
    }

    public void shortAlign(String input, String outputSam) throws IOException, InterruptedException {   // Aligner.java:46-56
        shortAlignFiles(reference, numThreads, device, input, outputSam);
    }

    public void index() throws IOException, InterruptedException {                                      // Aligner.java:62-69
        indexFile(reference);
    }

    static native void alignFiles(String reference, int numThreads, int device, String input, String outputSam);
    static native void shortAlignFiles(String reference, int numThreads, int device, String input, String outputSam);
    static native void indexFile(String reference);

    // ---- batch form for callers that hold 2-bit packed sequences (Sequence.java:10-13,31-41: A=0 C=1 T=2 G=3, LSB first) ----
    public static ByteBuffer direct(int bytes) { return ByteBuffer.allocateDirect(bytes).order(ByteOrder.LITTLE_ENDIAN); }
    static native long create(int match, int mismatch, int gapOpen, int gapExt, int device, int slots);
    static native void destroy(long handle);
    /** Every buffer must be direct and hold what the counts imply (4 bytes per off entry, 2 per len entry, 32 per result);
     *  otherwise IllegalArgumentException, before native code touches anything. Returns the CIGAR ops written. */
    static native long alignBatch(long handle,
                                  ByteBuffer rPacked, ByteBuffer rOff, ByteBuffer rLen, ByteBuffer rNmask, ByteBuffer rNmaskOff, int nReads,
                                  ByteBuffer wPacked, ByteBuffer wOff, ByteBuffer wLen, ByteBuffer wNmask, ByteBuffer wNmaskOff, int nWindows,
                                  ByteBuffer pairRefIdx, ByteBuffer results /* 32 B records */, ByteBuffer cigarPool);
}
