// sw_fwd16.cuh -- forward (score + end cell) pass, packed 16-bit lanes, sm_100a.
//
// Layout of the work: a GROUP of G consecutive lanes of one warp aligns a PAIR of tasks at once,
// task A in the low and task B in the high 16-bit half of every register (inter-task SIMD).
// Lane g owns K consecutive read rows [g*K, g*K+K) of both tasks in registers; the group sweeps
// the window as a systolic wavefront: at step s lane g updates its K cells of column s-g, then
// hands its bottom H (as H-oe) and the F entering the row below to lane g+1 with two warp shuffles.
//
// Per packed cell (two DP cells) the ALU work is:
//     PRMT              substitution score of (A,B) from the per-row byte profiles      [ALU]
//     VIADDMNMX.S16x2   E = max(E - ext, Hleft - oe)                                    [ALU]
//     VIADD.16x2        t = Hdiag - oe + (s + oe)                                       [second int pipe]
//     VIMNMX3.S16x2.RELU H = max(0, t, E, F)                                            [ALU]
//     VIADD.16x2        Hoe = H - oe     (feeds E of the next column, F of the next row, diag)
//     VIADDMNMX.S16x2   F = max(F - ext, Hoe)                                           [ALU]
//     VIMNMX3.S16x2     running column max, one op per two rows                         [ALU]
// which is the 7-op cell of SURVEY.md section 8d with the score lookup added and the running
// max halved. No shared-memory traffic in the cell loop except one selector load per lane per step.
//
// Substitution scores without a table lookup: row r of lane g keeps ONE register per task holding
// four int8 scores (vs A,C,T,G, each pre-biased by +oe); the window base of the current column is
// turned (once per pair, in the prologue) into a 16-bit PRMT selector kept in shared memory, so a
// single PRMT yields the sign-extended (scoreB:scoreA) pair. Padding rows use -128 bytes, padding
// columns use the sign-replicate selector (score 0 or -1 after bias => H stays below every real
// cell), so ragged pairs need no masks. N in the read = a row of (0+oe) bytes; N in the window needs
// one extra LOP3 per cell and is compiled only into the HAS_N variant.
//
// Arg-max: each lane keeps the running packed best and the step at which it first strictly
// improved; the group reduction orders candidates by (score, smaller column, smaller lane), which
// is exactly SPEC.md section 4 at lane granularity. The row inside the winning lane's K rows is
// resolved by the traceback pass.
#pragma once
#include "sw_common.cuh"
#include <type_traits>

namespace ubu {

// DEFAULT_SC = true compiles SPEC.md's default scoring (+1/-3/5/2) in as immediates: the packed
// instructions then read two registers instead of three, which matters because the register file,
// not just the ALU pipe, limits the dual-pipe issue rate (bench/rf_probe.cu).
#ifndef UBU_FWD_MIN_BLOCKS
#define UBU_FWD_MIN_BLOCKS 1
#endif
#ifndef UBU_FWD_UNROLL
#define UBU_FWD_UNROLL 2
#endif
// Group arg-max of the forward sweep and the FwdOut records of the pair: candidates are ordered by
// (score, smaller column, smaller lane), which is SPEC.md section 4 at lane granularity. A path of score S ending in row i
// holds at most i matches, so the end row is at least ceil(S / match): rows of the winning lane below that are dropped
// from the candidate range (a narrower traceback band).
template <int G, int K>
__device__ __forceinline__ void fwd_group_argmax_store(uint32_t best, int stepA, int stepB, int g, uint32_t tA, uint32_t tB, int LA, int LB,
                                                       const Scoring& sc, FwdOut* __restrict__ out, const TbPlan& plan) {
  const int match = sc.match;
  const int scA = (int)(int16_t)(best & 0xFFFFu), scB = (int)(int16_t)(best >> 16);
  unsigned long long keyA = 0ull, keyB = 0ull;
  if (scA > 0) keyA = ((unsigned long long)scA << 32) | ((unsigned long long)(0xFFFFu - (uint32_t)(stepA - g + 1)) << 8) | (unsigned long long)(31 - g);
  if (scB > 0) keyB = ((unsigned long long)scB << 32) | ((unsigned long long)(0xFFFFu - (uint32_t)(stepB - g + 1)) << 8) | (unsigned long long)(31 - g);
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const unsigned long long oa = __shfl_xor_sync(0xffffffffu, keyA, o, G);
    const unsigned long long ob = __shfl_xor_sync(0xffffffffu, keyB, o, G);
    keyA = oa > keyA ? oa : keyA; keyB = ob > keyB ? ob : keyB;
  }
  // The group's first lane writes the two records and decides how each task will be traced back (its class list), so no
  // separate planning pass has to re-read them; an unaligned task (score 0) gets its final record here.
  int clsA = -1, clsB = -1;
  if (g == 0) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const uint32_t t = half ? tB : tA;
      if (t == 0xFFFFFFFFu) continue;
      const unsigned long long key = half ? keyB : keyA;
      FwdOut o = {0, 0, 0, 0};
      int cls = -1;
      if (key) {
        const int lane = 31 - (int)(key & 0xFFu);
        o.score = (int)(key >> 32); o.je = (int)(0xFFFFu - (uint32_t)((key >> 8) & 0xFFFFu));
        o.i_lo = max(lane * K + 1, (o.score + match - 1) / match); o.i_hi = min(half ? LB : LA, lane * K + K);
        cls = tb_class_of(sc, o, plan.caps);
      } else {
        ubu_sw_result r; r.score = 0; r.ref_start = r.ref_end = r.q_start = r.q_end = 0; r.edit_distance = 0;
        r.cigar_off = 0; r.cigar_n = 0; r.flags = UBU_SW_FLAG_UNALIGNED;
        plan.res[t] = r;
      }
      out[t] = o;
      if (half) clsB = cls; else clsA = cls;
    }
  }
  const unsigned lane = threadIdx.x & 31u, lt = (1u << lane) - 1u;
#pragma unroll
  for (int c = 0; c < TB_CLASSES; ++c) {                                   // one counter update per warp and class
    const unsigned mA = __ballot_sync(0xffffffffu, clsA == c), mB = __ballot_sync(0xffffffffu, clsB == c);
    if (mA | mB) {
      const int leader = __ffs(mA | mB) - 1;
      unsigned base = 0;
      if ((int)lane == leader) base = atomicAdd(&plan.counts[c], (unsigned)(__popc(mA) + __popc(mB)));
      base = __shfl_sync(0xffffffffu, base, leader);
      uint32_t* list = plan.lists + (size_t)c * plan.n_slots;
      if (clsA == c) list[base + __popc(mA & lt)] = tA;
      if (clsB == c) list[base + __popc(mA) + __popc(mB & lt)] = tB;
    }
  }
}

// SEL_GLOBAL = true keeps the per-column selectors in a global scratch instead of shared memory: used only when the
// windows of a launch are too long for shared memory (the loads are sequential 2-byte reads, served by L1).
template <int G, int K, bool HAS_N, bool DEFAULT_SC, bool SEL_GLOBAL, int THREADS>
__global__ void __launch_bounds__(THREADS, UBU_FWD_MIN_BLOCKS)
sw_fwd16_kernel(DevSeqs reads, DevSeqs wins, const uint32_t* __restrict__ ref_idx,
                const uint32_t* __restrict__ order /* nullptr = identity */, uint32_t n_slots, uint32_t n_tasks, Scoring sc,
                FwdOut* __restrict__ out, int sel_stride, void* __restrict__ gsel, TbPlan plan) {
  static_assert(G == 1 || G == 2 || G == 4 || G == 8 || G == 16 || G == 32, "G must divide the warp");
  constexpr int GP = THREADS / G;
  using SelT = typename SelType<HAS_N>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelT* sel_all = SEL_GLOBAL ? reinterpret_cast<SelT*>(gsel) + (size_t)blockIdx.x * (THREADS / G) * sel_stride
                             : reinterpret_cast<SelT*>(smem_raw);
  __shared__ int s_wmax;

  const int tid = threadIdx.x;
  const int grp = tid / G, g = tid % G;
  if (tid == 0) s_wmax = 0;

  // ---- the pair of tasks of this group --------------------------------------------------
  const uint32_t slotA = 2u * (blockIdx.x * GP + grp), slotB = slotA + 1u;
  uint32_t tA = 0xFFFFFFFFu, tB = 0xFFFFFFFFu;
  if (order) { if (slotA < n_slots) tA = order[slotA]; if (slotB < n_slots) tB = order[slotB]; }
  else { if (slotA < n_tasks) tA = slotA; if (slotB < n_tasks) tB = slotB; }
  int LA = 0, WA = 0, LB = 0, WB = 0;
  const uint8_t *qA = nullptr, *qB = nullptr, *rA = nullptr, *rB = nullptr;
  const uint8_t *qnA = nullptr, *qnB = nullptr, *rnA = nullptr, *rnB = nullptr;
  if (tA != 0xFFFFFFFFu) {
    const uint32_t ri = ref_idx ? ref_idx[tA] : tA;
    LA = reads.len[tA]; qA = reads.packed + reads.off[tA];
    WA = wins.len[ri];  rA = wins.packed + wins.off[ri];
    if (reads.nmask) qnA = reads.nmask + reads.nmask_off[tA];
    if (HAS_N && wins.nmask) rnA = wins.nmask + wins.nmask_off[ri];
  }
  if (tB != 0xFFFFFFFFu) {
    const uint32_t ri = ref_idx ? ref_idx[tB] : tB;
    LB = reads.len[tB]; qB = reads.packed + reads.off[tB];
    WB = wins.len[ri];  rB = wins.packed + wins.off[ri];
    if (reads.nmask) qnB = reads.nmask + reads.nmask_off[tB];
    if (HAS_N && wins.nmask) rnB = wins.nmask + wins.nmask_off[ri];
  }

  // ---- per-row byte profiles (registers) ------------------------------------------------
  if (DEFAULT_SC) { sc.match = 1; sc.mismatch = -3; sc.gap_open = 5; sc.gap_ext = 2; }
  const int oe = sc.gap_open + sc.gap_ext;
  const uint32_t mmb = (uint32_t)(sc.mismatch + oe) & 0xFFu, mab = (uint32_t)(sc.match + oe) & 0xFFu;
  const uint32_t MM4 = mmb * 0x01010101u, MX = mab ^ mmb, N4 = ((uint32_t)oe & 0xFFu) * 0x01010101u;
  // The prologue is written for latency: a warp that is not yet in the sweep cannot feed the ALU pipe, so everything is
  // fetched 16 bases per pair of 32-bit loads (load16) and the loads of several chunks are issued before any is used.
  uint32_t profA[K], profB[K];
  {
    constexpr int NQ = (K + 15) / 16;
    uint32_t qa[NQ], qb[NQ], na[NQ], nb[NQ];
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
      const int r0 = g * K + 16 * c;
      qa[c] = (r0 < LA) ? load16(qA, r0) : 0u; qb[c] = (r0 < LB) ? load16(qB, r0) : 0u;
      na[c] = (qnA && r0 < LA) ? load16n(qnA, r0) : 0u; nb[c] = (qnB && r0 < LB) ? load16n(qnB, r0) : 0u;
    }
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int row = g * K + r;
      uint32_t pa = 0x80808080u, pb = 0x80808080u;          // padding row: -128 against every base
      if (row < LA) { pa = MM4 ^ (MX << (8u * ((qa[r / 16] >> (2 * (r % 16))) & 3u))); if ((na[r / 16] >> (r % 16)) & 1u) pa = N4; }
      if (row < LB) { pb = MM4 ^ (MX << (8u * ((qb[r / 16] >> (2 * (r % 16))) & 3u))); if ((nb[r / 16] >> (r % 16)) & 1u) pb = N4; }
      profA[r] = pa; profB[r] = pb;
    }
  }

  // ---- per-column PRMT selectors (shared memory or global scratch), index = column + G ----------------------
  SelT* sel = sel_all + (size_t)grp * sel_stride;
  {
    constexpr int CH = 4;                                   // chunks of 16 columns fetched together
    const int nch = (sel_stride + 15) / 16;
    for (int cb = g; cb < nch; cb += G * CH) {
      uint32_t wa[CH], wb[CH], fa[CH], fb[CH];
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int c0 = 16 * (cb + q * G) - G;               // first column of the chunk
        const bool on = cb + q * G < nch;
        wa[q] = (on && rA && c0 < WA && c0 + 16 > 0) ? load16(rA, c0) : 0u;   // rA == nullptr: no task in this half
        wb[q] = (on && rB && c0 < WB && c0 + 16 > 0) ? load16(rB, c0) : 0u;
        fa[q] = (HAS_N && rnA && on && c0 < WA && c0 + 16 > 0) ? load16n(rnA, c0) : 0u;
        fb[q] = (HAS_N && rnB && on && c0 < WB && c0 + 16 > 0) ? load16n(rnB, c0) : 0u;
      }
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int ch = cb + q * G;
        if (ch < nch) {
          const int c0 = 16 * ch - G;
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            const int c = c0 + u;
            uint32_t lo = 0x88u, hi = 0xCCu, nf = 0u;       // padding column: sign-replicate selectors
            if (c >= 0 && c < WA) { lo = ((wa[q] >> (2 * u)) & 3u) * 0x11u + 0x80u; if (HAS_N && ((fa[q] >> u) & 1u)) nf |= 0x00800000u; }
            if (c >= 0 && c < WB) { hi = ((wb[q] >> (2 * u)) & 3u) * 0x11u + 0xC4u; if (HAS_N && ((fb[q] >> u) & 1u)) nf |= 0x80000000u; }
            if (16 * ch + u < sel_stride) sel[16 * ch + u] = (SelT)(lo | (hi << 8) | nf);
          }
        }
      }
    }
  }
  __syncthreads();
  if (g == 0) atomicMax(&s_wmax, max(WA, WB));
  __syncthreads();
  const int nsteps = s_wmax > 0 ? s_wmax + G - 1 : 0;

  // ---- systolic sweep ---------------------------------------------------------------------
  const uint32_t NEG_OE = DEFAULT_SC ? 0xFFF9FFF9u : sc.neg_oe2, NEG_EXT = DEFAULT_SC ? 0xFFFEFFFEu : sc.neg_ext2,
                 POS_OE = DEFAULT_SC ? 0x00070007u : sc.pos_oe2;
  uint32_t Hoe[K], E[K];
#pragma unroll
  for (int r = 0; r < K; ++r) { Hoe[r] = NEG_OE; E[r] = NEG_OE; }
  uint32_t hoe_bot = NEG_OE, f_bot = NEG_OE, hoe_up_prev = NEG_OE;
  uint32_t best = 0u; int stepA = 0, stepB = 0;
  const SelT* selp = sel + (G - g);                       // selp[s] = selector of column s - g

  // boundary fix-up of lane 0 done with one IMAD per shuffle (FMA pipe) instead of a select (ALU pipe)
  uint32_t bmul = (g == 0) ? 0u : 1u;
  const uint32_t badd = (g == 0) ? NEG_OE : 0u;
  asm volatile("" : "+r"(bmul));                          // opaque, so the multiply is not folded back into a select

  constexpr int STEP_UNROLL = UBU_FWD_UNROLL;
#pragma unroll STEP_UNROLL
  for (int s = 0; s < nsteps; ++s) {
    const uint32_t selv = selp[s];
    const uint32_t up = __shfl_up_sync(0xffffffffu, hoe_bot, 1, G) * bmul + badd;   // row 0: H[0][j] = 0
    uint32_t F = __shfl_up_sync(0xffffffffu, f_bot, 1, G) * bmul + badd;            // row 0: F[1][j] = -oe
    uint32_t hd = hoe_up_prev;                            // H[top-1][j-1] - oe
    hoe_up_prev = up;
    uint32_t isn = 0u;
    if (HAS_N) isn = prmt(selv, 0u, 0xBBAAu);             // 0xFFFF per half whose window base is N
    // The running max is taken over t = Hdiag + s instead of H: the arg-max cell of a local alignment is
    // always entered diagonally (SPEC.md section 5), so max t == max H with identical positions, and
    // giving t a second consumer keeps the add on the second integer pipe (ptxas would otherwise fuse
    // add+max into a second ALU-pipe VIADDMNMX).
    uint32_t colmax = 0u, tprev = 0u;
#pragma unroll
    for (int r = 0; r < K; ++r) {
      uint32_t sp = prmt(profA[r], profB[r], selv & 0xFFFFu);               // (s+oe) of B : A      [ALU]
      if (HAS_N) sp = (sp & ~isn) | (POS_OE & isn);                          // window N scores 0
      const uint32_t t = __vadd2(hd, sp);                                    // H[i-1][j-1] + s      [pipe 2]
      hd = Hoe[r];
      E[r] = __viaddmax_s16x2(E[r], NEG_EXT, Hoe[r]);                        // E[i][j]              [ALU]
      const uint32_t fx = __vadd2(F, NEG_EXT);                               // F[i][j] - ext        [pipe 2]
      const uint32_t h = __vimax3_s16x2_relu(t, E[r], F);                    // H[i][j]              [ALU]
      Hoe[r] = __vadd2(h, NEG_OE);                                           //                      [pipe 2]
      F = __viaddmax_s16x2(h, NEG_OE, fx);                                   // F[i+1][j]            [ALU]
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
