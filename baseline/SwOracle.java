/**
 * SwOracle.java -- the "Java on host cores" baseline BASELINE.json's metric asks for (SURVEY.md section 8d-ii, BASELINE.md section 3).
 *
 * mozack/ubu holds no Java Smith-Waterman (its Aligner forks bwa, src/main/java/edu/unc/bioinf/ubu/assembly/Aligner.java:18-22,46-56),
 * so this is a line-for-line Java restatement of THIS repo's SPEC.md (sections 2-6), identical in structure to oracle/sw_oracle.c:
 * full matrix, 32-bit ints, row-major fill with the positional arg-max of section 4, the three-state traceback of section 5.
 * It is a baseline, not the reference aligner; every number it produces must be labelled "SPEC oracle (Java), N threads".
 *
 * NEVER COMPILED IN THE BUILD IMAGE (no JDK there). bench.py runs it only when `java` and `javac` are on PATH:
 *     javac -d baseline/_classes baseline/SwOracle.java
 *     java -cp baseline/_classes SwOracle <reads> <readLen> <windowLen> <threads> <seed>
 * and prints one JSON line {"gcups": ..., "reads_per_s": ..., "threads": ..., "checksum": ...}; otherwise bench.py reports
 * "Java baseline not run: no JVM on this box".
 *
 * Synthetic input of the cfg3 shape: windows i.i.d. uniform over the reference's 2-bit code (A=0 C=1 T=2 G=3, Sequence.java:10-13),
 * a read = a stretch of its window with 1 % substitutions (java.util.Random, so the set differs from the numpy-generated GPU set;
 * GCUPS is normalised by cells, which is what the comparison needs).
 */
import java.util.Random;
import java.util.concurrent.atomic.AtomicLong;

public class SwOracle {
    static final int NEG_INF = Integer.MIN_VALUE / 4;
    final int match, mismatch, gapOpen, gapExt;

    SwOracle(int match, int mismatch, int gapOpen, int gapExt) { this.match = match; this.mismatch = mismatch; this.gapOpen = gapOpen; this.gapExt = gapExt; }

    /** SPEC section 2: N (code >= 4) never mismatches (CompareToReference.java:92). */
    int subScore(byte a, byte b) { return (a >= 4 || b >= 4) ? 0 : (a == b ? match : mismatch); }

    /** Returns {score, refStart, refEnd, qStart, qEnd, editDistance, cigarOps...}; ops are BAM-encoded (len << 4 | op, M=0 I=1 D=2 S=4). */
    int[] align(byte[] q, byte[] r) {
        final int L = q.length, W = r.length;
        if (L == 0 || W == 0) return new int[] {0, 0, 0, 0, 0, 0};
        final int stride = W + 1;
        int[] H = new int[(L + 1) * stride], E = new int[(L + 1) * stride], F = new int[(L + 1) * stride];
        for (int j = 0; j <= W; ++j) { H[j] = 0; E[j] = NEG_INF; F[j] = NEG_INF; }                      // SPEC section 3 boundaries
        for (int i = 1; i <= L; ++i) { H[i * stride] = 0; E[i * stride] = NEG_INF; F[i * stride] = NEG_INF; }
        int S = 0, ie = 0, je = 0;
        for (int i = 1; i <= L; ++i) {
            for (int j = 1; j <= W; ++j) {
                final int c = i * stride + j;
                final int e = Math.max(E[c - 1], H[c - 1] - gapOpen) - gapExt;
                final int f = Math.max(F[c - stride], H[c - stride] - gapOpen) - gapExt;
                int h = H[c - stride - 1] + subScore(q[i - 1], r[j - 1]);
                h = Math.max(Math.max(h, e), Math.max(f, 0));
                E[c] = e; F[c] = f; H[c] = h;
                if (h > S || (h == S && h > 0 && (j < je || (j == je && i < ie)))) { S = h; ie = i; je = j; }   // section 4: smallest j, then smallest i
            }
        }
        if (S == 0) return new int[] {0, 0, 0, 0, 0, 0};
        byte[] ops = new byte[L + W + 2];
        int nops = 0, i = ie, j = je, edits = 0, st = 0;                                                  // section 5; st: 0 = H, 1 = F, 2 = E
        for (;;) {
            final int c = i * stride + j;
            if (st == 0) {
                if (H[c] == 0) break;
                if (H[c] == H[c - stride - 1] + subScore(q[i - 1], r[j - 1])) {
                    ops[nops++] = 0;
                    if (q[i - 1] < 4 && r[j - 1] < 4 && q[i - 1] != r[j - 1]) ++edits;
                    --i; --j;
                } else if (H[c] == F[c]) st = 1;
                else st = 2;
            } else if (st == 1) {
                ops[nops++] = 1; ++edits;
                if (i > 1 && F[c] == F[c - stride] - gapExt) { --i; } else { --i; st = 0; }
            } else {
                ops[nops++] = 2; ++edits;
                if (j > 1 && E[c] == E[c - 1] - gapExt) { --j; } else { --j; st = 0; }
            }
        }
        int[] out = new int[6 + L + 2];
        int n = 6;
        if (i > 0) out[n++] = (i << 4) | 4;                                                               // section 6: leading S
        for (int k = nops; k > 0;) {
            final byte op = ops[k - 1]; int len = 0;
            while (k > 0 && ops[k - 1] == op) { ++len; --k; }
            out[n++] = (len << 4) | op;
        }
        if (L - ie > 0) out[n++] = ((L - ie) << 4) | 4;
        out[0] = S; out[1] = j + 1; out[2] = je; out[3] = i + 1; out[4] = ie; out[5] = edits;
        return java.util.Arrays.copyOf(out, n);
    }

    public static void main(String[] args) throws Exception {
        final int reads = Integer.parseInt(args[0]), L = Integer.parseInt(args[1]), W = Integer.parseInt(args[2]);
        final int threads = Integer.parseInt(args[3]); final long seed = Long.parseLong(args[4]);
        final int nWin = Math.max(reads / 5, 1);
        final byte[][] wins = new byte[nWin][W], qs = new byte[reads][L];
        final int[] widx = new int[reads];
        Random rng = new Random(seed);
        for (byte[] w : wins) for (int k = 0; k < W; ++k) w[k] = (byte) rng.nextInt(4);
        for (int t = 0; t < reads; ++t) {
            widx[t] = t * nWin / reads;                                                                    // reads grouped by window
            final int s = rng.nextInt(W - L + 1);
            for (int k = 0; k < L; ++k) { byte b = wins[widx[t]][s + k]; if (rng.nextInt(100) == 0) b = (byte) ((b + 1 + rng.nextInt(3)) & 3); qs[t][k] = b; }
        }
        final AtomicLong next = new AtomicLong(0), checksum = new AtomicLong(0);
        final SwOracle sw = new SwOracle(1, -3, 5, 2);
        Thread[] pool = new Thread[threads];
        final long t0 = System.nanoTime();
        for (int k = 0; k < threads; ++k) {
            pool[k] = new Thread(() -> {
                long local = 0;
                for (long t; (t = next.getAndIncrement()) < reads;) { int[] res = sw.align(qs[(int) t], wins[widx[(int) t]]); local += res[0] + 31L * res[1] + res.length; }
                checksum.addAndGet(local);
            });
            pool[k].start();
        }
        for (Thread th : pool) th.join();
        final double dt = (System.nanoTime() - t0) / 1e9;
        System.out.printf("{\"gcups\": %.4f, \"reads_per_s\": %.1f, \"threads\": %d, \"seconds\": %.3f, \"reads\": %d, \"checksum\": %d}%n",
                          (double) reads * L * W / dt / 1e9, reads / dt, threads, dt, reads, checksum.get());
    }
}
