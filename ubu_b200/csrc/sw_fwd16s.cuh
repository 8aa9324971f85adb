// sw_fwd16s.cuh -- forward pass for task pairs that SHARE their window (many reads are realigned against one
// contig / window: the two mates of cfg2, the reads of a region in ReAligner.alignToContigs, ReAligner.java:963-975).
//
// Same systolic layout as sw_fwd16.cuh (a group of G lanes, K rows per lane, task A in the low and task B in
// the high 16-bit half of every register). What changes is where the substitution scores come from. With one
// window for both halves, the score pair (sB:sA) of a row depends on ONE window base, so the group keeps a
// query profile in shared memory,
//     prof[b][row] = ( s(readB[row], b) + oe : s(readA[row], b) + oe ),   b in {A, C, T, G, (N), outside-the-window},
// and a lane fetches the K words of its rows for the base of its current column with ceil(K/4) LDS.128. The
// per-cell PRMT of the general kernel (ALU pipe, the pipe that limits the sweep) disappears: the ALU work per
// packed cell drops from 4.5 to 3.5 instructions (E, H, F, half a running max), the profile registers (2K) are
// freed, and a window N needs no extra LOP3 (it is just one more profile row).
//
// Shared-memory layout per group, in 16-byte units: lane stride LU = ceil(K/4) made odd, base-row stride BU =
// G*LU rounded up to a multiple of 8, group stride = 4 mod 8 when G == 4. Then the eight lanes of every LDS.128
// phase fall into eight distinct 16-byte bank groups whatever bases their columns hold: no bank conflicts.
#pragma once
#include "sw_fwd16.cuh"

namespace ubu {

#ifndef UBU_FWDS_MIN_BLOCKS
#define UBU_FWDS_MIN_BLOCKS 12
#endif
#ifndef UBU_FWDS_CHAIN3
#define UBU_FWDS_CHAIN3 0
#endif
#ifndef UBU_FWDS_UNROLL
#define UBU_FWDS_UNROLL 2
#endif

template <int G, int K, bool HAS_N> struct FwdsLayout {
  static constexpr int KQ = (K + 3) / 4;                     // LDS.128 per lane per step
  static constexpr int LU = KQ | 1;                          // lane stride (16-byte units), odd
  static constexpr int BU = ((G * LU + 7) / 8) * 8;          // base-row stride (units), multiple of 8
  static constexpr int NB = HAS_N ? 6 : 5;                   // rows: A C T G [N] outside
  static constexpr int PAD = NB - 1;
  // table stride between groups in units: 4 mod 8 when G == 4 (two groups share an LDS.128 phase)
  static constexpr int TU = NB * BU + (G == 4 ? 4 : 0);
  // selector arrays (one byte per column, sel_stride % 16 == 0) follow the tables; their stride in words is odd, so the
  // one-byte loads of the groups of a warp fall into different banks
  __host__ __device__ static constexpr int sel_words(int sel_stride) { return (sel_stride / 4) | 1; }
  __host__ __device__ static constexpr size_t smem_bytes(int groups, int sel_stride) {
    return (size_t)groups * TU * 16 + (size_t)groups * sel_words(sel_stride) * 4;
  }
};

template <int G, int K, bool HAS_N, bool DEFAULT_SC, int THREADS>
__global__ void __launch_bounds__(THREADS, UBU_FWDS_MIN_BLOCKS)
sw_fwd16s_kernel(DevSeqs reads, DevSeqs wins, const uint32_t* __restrict__ ref_idx,
                 const uint32_t* __restrict__ order /* nullptr = identity */, uint32_t n_slots, uint32_t n_tasks, Scoring sc,
                 FwdOut* __restrict__ out, int sel_stride, TbPlan plan) {
  static_assert(G == 4 || G == 8 || G == 16 || G == 32, "G must divide the warp and cover an LDS.128 phase half");
  using LY = FwdsLayout<G, K, HAS_N>;
  constexpr int GP = THREADS / G, KQ = LY::KQ, LU = LY::LU, BU = LY::BU, NB = LY::NB, PAD = LY::PAD;
  extern __shared__ uint4 smem4[];
  __shared__ int s_wmax;

  const int tid = threadIdx.x;
  const int grp = tid / G, g = tid % G;
  if (tid == 0) s_wmax = 0;
  uint32_t* tbl = reinterpret_cast<uint32_t*>(smem4 + grp * LY::TU);                                   // [NB][BU * 4] words
  uint32_t* selw = reinterpret_cast<uint32_t*>(smem4 + GP * LY::TU) + grp * LY::sel_words(sel_stride);   // [sel_stride] bytes
  uint8_t* sel = reinterpret_cast<uint8_t*>(selw);                                                     // index = column + G

  // ---- the pair of tasks of this group; both halves sweep the same window ---------------------
  const uint32_t slotA = 2u * (blockIdx.x * GP + grp), slotB = slotA + 1u;
  uint32_t tA = 0xFFFFFFFFu, tB = 0xFFFFFFFFu;
  if (order) { if (slotA < n_slots) tA = order[slotA]; if (slotB < n_slots) tB = order[slotB]; }
  else { if (slotA < n_tasks) tA = slotA; if (slotB < n_tasks) tB = slotB; }
  int LA = 0, LB = 0, W = 0;
  const uint8_t *qA = nullptr, *qB = nullptr, *qnA = nullptr, *qnB = nullptr, *rW = nullptr, *rnW = nullptr;
  if (tA != 0xFFFFFFFFu) {
    LA = reads.len[tA]; qA = reads.packed + reads.off[tA];
    if (reads.nmask) qnA = reads.nmask + reads.nmask_off[tA];
  }
  if (tB != 0xFFFFFFFFu) {
    LB = reads.len[tB]; qB = reads.packed + reads.off[tB];
    if (reads.nmask) qnB = reads.nmask + reads.nmask_off[tB];
  }
  {
    const uint32_t tw = tA != 0xFFFFFFFFu ? tA : tB;
    if (tw != 0xFFFFFFFFu) {
      const uint32_t ri = ref_idx ? ref_idx[tw] : tw;
      W = wins.len[ri]; rW = wins.packed + wins.off[ri];
      if (HAS_N && wins.nmask) rnW = wins.nmask + wins.nmask_off[ri];
    }
  }

  if (DEFAULT_SC) { sc.match = 1; sc.mismatch = -3; sc.gap_open = 5; sc.gap_ext = 2; }
  const int oe = sc.gap_open + sc.gap_ext;

  // ---- query profile of the pair (shared memory) ----------------------------------------------
  {
    constexpr int NQ = (K + 15) / 16;
    uint32_t qa[NQ], qb[NQ], na[NQ], nb[NQ];
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
      const int r0 = g * K + 16 * c;
      qa[c] = (r0 < LA) ? load16(qA, r0) : 0u; qb[c] = (r0 < LB) ? load16(qB, r0) : 0u;
      na[c] = (qnA && r0 < LA) ? load16n(qnA, r0) : 0u; nb[c] = (qnB && r0 < LB) ? load16n(qnB, r0) : 0u;
    }
    const uint32_t ma = (uint32_t)(sc.match + oe) & 0xFFFFu, mm = (uint32_t)(sc.mismatch + oe) & 0xFFFFu, nn = (uint32_t)oe & 0xFFFFu;
    uint32_t* lt = tbl + g * (LU * 4);
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int row = g * K + r;
      const bool va = row < LA, vb = row < LB;
      const uint32_t ca = (qa[r / 16] >> (2 * (r % 16))) & 3u, cb = (qb[r / 16] >> (2 * (r % 16))) & 3u;
      const bool an = (na[r / 16] >> (r % 16)) & 1u, bn = (nb[r / 16] >> (r % 16)) & 1u;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const uint32_t sa = !va ? 0xFF80u : an ? nn : (ca == (uint32_t)b ? ma : mm);
        const uint32_t sb = !vb ? 0xFF80u : bn ? nn : (cb == (uint32_t)b ? ma : mm);
        lt[b * (BU * 4) + r] = sa | (sb << 16);
      }
      if (HAS_N) lt[4 * (BU * 4) + r] = (va ? nn : 0xFF80u) | ((vb ? nn : 0xFF80u) << 16);      // window N: score 0
      lt[PAD * (BU * 4) + r] = (va ? 0u : 0xFF80u) | ((vb ? 0u : 0xFF80u) << 16);               // outside the window: -oe
    }
#pragma unroll
    for (int r = K; r < KQ * 4; ++r) {                      // words an LDS.128 reads past row K-1
#pragma unroll
      for (int b = 0; b < NB; ++b) lt[b * (BU * 4) + r] = 0xFF80FF80u;
    }
  }

  // ---- per-column profile-row index (one byte per column), index = column + G -------------------
  {
    constexpr int CH = 4;                                   // chunks of 16 columns fetched together
    const int nch = sel_stride / 16;
    for (int cb = g; cb < nch; cb += G * CH) {
      uint32_t wv[CH], fv[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int c0 = 16 * (cb + q * G) - G;               // first column of the chunk
        const bool on = cb + q * G < nch && rW && c0 < W && c0 + 16 > 0;
        wv[q] = on ? load16(rW, c0) : 0u;
        fv[q] = (HAS_N && rnW && on) ? load16n(rnW, c0) : 0u;
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int ch = cb + q * G;
        if (ch < nch) {
          const int c0 = 16 * ch - G;
          uint32_t w4[4];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            uint32_t v = 0u;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int c = c0 + 4 * k4 + u;
              uint32_t b = (uint32_t)PAD;
              if (c >= 0 && c < W) { b = (wv[q] >> (2 * (4 * k4 + u))) & 3u; if (HAS_N && ((fv[q] >> (4 * k4 + u)) & 1u)) b = 4u; }
              v |= b << (8 * u);
            }
            w4[k4] = v;
          }
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) selw[4 * ch + k4] = w4[k4];
        }
      }
    }
  }
  __syncthreads();
  if (g == 0) atomicMax(&s_wmax, W);
  __syncthreads();
  const int nsteps = s_wmax > 0 ? s_wmax + G - 1 : 0;

  // ---- systolic sweep ---------------------------------------------------------------------
  const uint32_t NEG_OE = DEFAULT_SC ? 0xFFF9FFF9u : sc.neg_oe2, NEG_EXT = DEFAULT_SC ? 0xFFFEFFFEu : sc.neg_ext2;
  uint32_t Hoe[K], E[K];
#pragma unroll
  for (int r = 0; r < K; ++r) { Hoe[r] = NEG_OE; E[r] = NEG_OE; }
  uint32_t hoe_bot = NEG_OE, f_bot = NEG_OE, hoe_up_prev = NEG_OE;
  uint32_t best = 0u; int stepA = 0, stepB = 0;
  const uint8_t* selp = sel + (G - g);                    // selp[s] = profile row of column s - g
  const uint4* ltab = reinterpret_cast<const uint4*>(tbl) + g * LU;

  // boundary fix-up of lane 0 done with one IMAD per shuffle (FMA pipe) instead of a select (ALU pipe)
  uint32_t bmul = (g == 0) ? 0u : 1u;
  const uint32_t badd = (g == 0) ? NEG_OE : 0u;
  asm volatile("" : "+r"(bmul));                          // opaque, so the multiply is not folded back into a select

  constexpr int STEP_UNROLL = UBU_FWDS_UNROLL;
#pragma unroll STEP_UNROLL
  for (int s = 0; s < nsteps; ++s) {
    const uint4* prow = ltab + (uint32_t)selp[s] * (uint32_t)BU;
    uint32_t sp[KQ * 4];
#pragma unroll
    for (int q = 0; q < KQ; ++q) { const uint4 v = prow[q]; sp[4 * q] = v.x; sp[4 * q + 1] = v.y; sp[4 * q + 2] = v.z; sp[4 * q + 3] = v.w; }
    const uint32_t up = __shfl_up_sync(0xffffffffu, hoe_bot, 1, G) * bmul + badd;   // row 0: H[0][j] = 0
    uint32_t F = __shfl_up_sync(0xffffffffu, f_bot, 1, G) * bmul + badd;            // row 0: F[1][j] = -oe
    uint32_t hd = hoe_up_prev;                            // H[top-1][j-1] - oe
    hoe_up_prev = up;
    // running max over t = Hdiag + s (see sw_fwd16.cuh): the arg-max cell is always entered diagonally
    uint32_t colmax = 0u, tprev = 0u;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const uint32_t t = __vadd2(hd, sp[r]);                                 // H[i-1][j-1] + s      [pipe 2]
      hd = Hoe[r];
      E[r] = __viaddmax_s16x2(E[r], NEG_EXT, Hoe[r]);                        // E[i][j]              [ALU]
#if UBU_FWDS_CHAIN3
      const uint32_t h = __vimax3_s16x2_relu(t, E[r], F);                    // H[i][j]              [ALU]
      Hoe[r] = __vadd2(h, NEG_OE);                                           //                      [pipe 2]
      F = __viaddmax_s16x2(F, NEG_EXT, Hoe[r]);                              // F[i+1][j]            [ALU]
#else
      const uint32_t fx = __vadd2(F, NEG_EXT);                               // F[i][j] - ext        [pipe 2]
      const uint32_t h = __vimax3_s16x2_relu(t, E[r], F);                    // H[i][j]              [ALU]
      Hoe[r] = __vadd2(h, NEG_OE);                                           //                      [pipe 2]
      F = __viaddmax_s16x2(h, NEG_OE, fx);                                   // F[i+1][j]            [ALU]
#endif
      if (r & 1) colmax = __vimax3_s16x2(colmax, tprev, t); else tprev = t;  //                      [ALU, 1 per 2 rows]
    }
    if (K & 1) colmax = __vmaxs2(colmax, tprev);
    hoe_bot = Hoe[K - 1]; f_bot = F;
    bool keep_hi, keep_lo;                                                   // true = old best >= column max
    best = vmax2_keep(best, colmax, keep_hi, keep_lo);
    if (!keep_lo) stepA = s;
    if (!keep_hi) stepB = s;
  }

  fwd_group_argmax_store<G, K>(best, stepA, stepB, g, tA, tB, LA, LB, sc, out, plan);
}

}  // namespace ubu
