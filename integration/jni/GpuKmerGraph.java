package edu.unc.bioinf.ubu.assembly;

import java.nio.ByteBuffer;

/**
 * Graph-building half of {@link Assembler} (src/main/java/edu/unc/bioinf/ubu/assembly/Assembler.java) on the GPU:
 * replaces the addToGraph loop (Assembler.java:74-92, 452-474), filterLowFrequencyNodes (:201-227) and identifyRootNodes
 * (:270-276) by one JNI call into libubu_sw.so (include/ubu_kmer.h). buildContigs (:278-367) stays in Java and walks the
 * returned node table instead of the HashMap: a node is a 64-byte record {kmer[2], region, count, first_read, first_pos,
 * flags, to_node[4], from_node[4]}; successors are to_node[] in Node.toNodes order, minus entries flagged FILTERED.
 *
 * UNTESTED: no JDK exists in the image this repo was built in; see INTEGRATION.md section 6.
 */
public class GpuKmerGraph {
    static { System.loadLibrary("ubu_sw_jni"); }

    public static final int FLAG_MULTI_READS = 1, FLAG_FILTERED = 2, FLAG_ROOT = 4;   // include/ubu_kmer.h
    public static final int NODE_BYTES = 64;

    private final long handle;
    private final int kmerSize, minNodeFrequency;          // Assembler.setKmerSize (:165), setMinNodeFrequncy (:181)

    public GpuKmerGraph(int kmerSize, int minNodeFrequency, int device) {
        this.kmerSize = kmerSize; this.minNodeFrequency = minNodeFrequency;
        this.handle = create(device);
    }

    /**
     * reads: 2-bit packed with Sequence's own code (Sequence.java:10-13,31-41) in direct buffers; a read whose string contains
     * 'N' is skipped by the library (its N-plane bit is set), a read with X0 > 1 by passing region -1 (Assembler.java:77-81).
     * Returns the number of nodes written to `nodes` (capacity nodes.capacity() / NODE_BYTES); throws RuntimeException on
     * failure like the rest of the pipeline.
     */
    public long build(ByteBuffer rPacked, ByteBuffer rOff, ByteBuffer rLen, ByteBuffer rNmask, ByteBuffer rNmaskOff, int nReads,
                      ByteBuffer readRegion, long maxNodes, ByteBuffer nodes) {
        return build(handle, kmerSize, minNodeFrequency, rPacked, rOff, rLen, rNmask, rNmaskOff, nReads, readRegion, maxNodes, nodes);
    }

    public void close() { destroy(handle); }

    static native long create(int device);
    static native void destroy(long handle);
    static native long build(long handle, int kmerSize, int minNodeFrequency, ByteBuffer rPacked, ByteBuffer rOff, ByteBuffer rLen,
                             ByteBuffer rNmask, ByteBuffer rNmaskOff, int nReads, ByteBuffer readRegion, long maxNodes, ByteBuffer nodes);
}
