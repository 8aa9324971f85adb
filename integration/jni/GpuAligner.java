package edu.unc.bioinf.ubu.assembly;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;

/**
 * Drop-in for {@link Aligner} (src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java) that keeps its
 * constructor and method signatures and replaces the three `bwa` forks by one JNI call into libubu_sw.so.
 *
 * UNTESTED: no JDK exists in the image this repo was built in; see INTEGRATION.md.
 * SAM I/O, read batching and result write-back stay in Java (ubu's existing net.sf.samtools stack);
 * only alignBatch crosses into native code.
 */
public class GpuAligner {
    static { System.loadLibrary("ubu_sw_jni"); }

    private final String reference;      // FASTA of windows / contigs, as today (Aligner.java:9)
    private final long handle;

    public GpuAligner(String reference, int numThreads) {           // Aligner.java:13-16; numThreads is unused (one handle = one GPU)
        this.reference = reference;
        this.handle = create(1, -3, 5, 2, 0, 3);
    }

    /** Aligner.java:62-69 built a BWT index; window SW needs none. Kept so callers compile unchanged. */
    public void index() { }

    // shortAlign(String input, String outputSam) / align(String input, String outputSam) (Aligner.java:18-22,46-56):
    // read `input` (FASTQ/FASTA) with the existing readers, 2-bit pack reads and windows into direct buffers with
    // Sequence's own code (Sequence.java:10-13,31-41: A=0 C=1 T=2 G=3, LSB first), call alignBatch, then write one
    // SAMRecord per result (POS = window offset + ref_start, CIGAR from the pool, NM = edit_distance) to outputSam.

    public static ByteBuffer direct(int bytes) { return ByteBuffer.allocateDirect(bytes).order(ByteOrder.LITTLE_ENDIAN); }

    static native long create(int match, int mismatch, int gapOpen, int gapExt, int device, int slots);
    static native void destroy(long handle);
    static native long alignBatch(long handle,
                                  ByteBuffer rPacked, ByteBuffer rOff, ByteBuffer rLen, ByteBuffer rNmask, ByteBuffer rNmaskOff, int nReads,
                                  ByteBuffer wPacked, ByteBuffer wOff, ByteBuffer wLen, ByteBuffer wNmask, ByteBuffer wNmaskOff, int nWindows,
                                  ByteBuffer pairRefIdx, ByteBuffer results /* 32 B records */, ByteBuffer cigarPool);

    public void close() { destroy(handle); }
}
