// sw_tbrect.cuh -- traceback fill for WIDE bands, in the forward kernel's own systolic structure.
//
// The one-thread-per-pair band kernels of sw_tb.cuh are latency-bound once the band no longer fits registers.
// A wide band (K + 2Gb of 60..130 diagonals for indel-rich reads) covers a large part of the rectangle
// rows 1..i_hi x columns [clo, je], clo = max(1, je - i_hi - Gb + 1), so this kernel simply re-runs the forward
// recurrence over that rectangle with the forward pass's work decomposition (a group of G lanes per task PAIR,
// K rows per lane in registers, two warp shuffles per step, no running max) and stores H of every cell whose
// lane-column touches the band: K consecutive 16-bit pairs per lane per step, column-major, 16-byte stores.
// Values in the rectangle with a zero boundary at column clo-1 are sandwiched between the banded and the
// full-matrix values, so the exactness argument of DESIGN.md section 4.2 applies unchanged.
// Afterwards lane 0 (task A) and lane 1 (task B) of the group replay the SPEC walk on the stored H
// (walk_h_and_emit<.., USE_FLAG = false>: the diagonal test reads the sequences, 16 bases per load).
#pragma once
#include "sw_common.cuh"
#include "sw_tb.cuh"

namespace ubu {

struct RectGeom { int S, je, i_lo, i_hi, gb, d_lo, d_hi, clo, w; };
__device__ __forceinline__ RectGeom make_rect(const Scoring& sc, const FwdOut& f) {
  RectGeom g;
  g.S = f.score; g.je = f.je; g.i_lo = f.i_lo; g.i_hi = f.i_hi;
  g.gb = gap_bound(sc, f.i_hi, f.score);
  g.d_lo = f.je - f.i_hi - g.gb; g.d_hi = f.je - f.i_lo + g.gb;
  g.clo = max(1, 1 + g.d_lo);                                 // leftmost column any band cell of rows >= 1 can have
  g.w = f.je - g.clo + 1;
  return g;
}

// scratch: one block of wcap * (G*K) words per GROUP, cell (row0, k) at block[k * (G*K) + row0]
#ifndef UBU_TBR_MIN_BLOCKS
#define UBU_TBR_MIN_BLOCKS 8
#endif
template <int G, int K, bool HAS_N, bool DEFAULT_SC, int THREADS>
__global__ void __launch_bounds__(THREADS, UBU_TBR_MIN_BLOCKS)
tb_rect_kernel(TbArgs a, const uint32_t* __restrict__ list, const unsigned int* __restrict__ count,
               uint32_t* __restrict__ scratch, uint32_t* __restrict__ opscratch, int wcap, int sel_stride) {
  static_assert(K % 4 == 0, "K must be a multiple of 4 (16-byte stores)");
  constexpr int GP = THREADS / G, RS = G * K;
  using SelT = typename SelType<HAS_N>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelT* sel_all = reinterpret_cast<SelT*>(smem_raw);

  const int tid = threadIdx.x, grp = tid / G, g = tid % G;
  const uint32_t n_tasks = *count, n_pairs = (n_tasks + 1u) / 2u;
  const uint32_t n_groups = gridDim.x * GP, gid = blockIdx.x * GP + grp;
  Scoring sc = a.sc;
  if (DEFAULT_SC) { sc.match = 1; sc.mismatch = -3; sc.gap_open = 5; sc.gap_ext = 2; }
  const int oe = sc.gap_open + sc.gap_ext;
  const uint32_t mmb = (uint32_t)(sc.mismatch + oe) & 0xFFu, mab = (uint32_t)(sc.match + oe) & 0xFFu;
  const uint32_t MM4 = mmb * 0x01010101u, MX = mab ^ mmb, N4 = ((uint32_t)oe & 0xFFu) * 0x01010101u;
  const uint32_t NEG_OE = DEFAULT_SC ? 0xFFF9FFF9u : sc.neg_oe2, NEG_EXT = DEFAULT_SC ? 0xFFFEFFFEu : sc.neg_ext2,
                 POS_OE = DEFAULT_SC ? 0x00070007u : sc.pos_oe2;
  SelT* sel = sel_all + (size_t)grp * sel_stride;
  uint32_t* block = scratch + (size_t)gid * ((size_t)wcap * RS);
  uint32_t* opbuf = opscratch + (size_t)(blockIdx.x * THREADS + tid) * (size_t)(2 * wcap + 16);
  uint32_t bmul = (g == 0) ? 0u : 1u;
  const uint32_t badd = (g == 0) ? NEG_OE : 0u;
  asm volatile("" : "+r"(bmul));

  // all groups of a warp iterate together (the shuffles use the full mask); a group without a pair runs idle
  const uint32_t rounds = (n_pairs + n_groups - 1) / n_groups;
  for (uint32_t round = 0; round < rounds; ++round) {
    const uint32_t pair = round * n_groups + gid;
    const bool have = pair < n_pairs;
    const bool haveB = have && 2u * pair + 1u < n_tasks;
    const uint32_t tA = have ? list[2u * pair] : 0u, tB = haveB ? list[2u * pair + 1u] : tA;
    TaskView va, vb; RectGeom ra, rb;
    if (have) { va = load_task(a, tA); vb = load_task(a, tB); ra = make_rect(sc, a.fwd[tA]); rb = make_rect(sc, a.fwd[tB]); }
    else { va.L = vb.L = 0; va.W = vb.W = 0; va.q = vb.q = va.r = vb.r = va.qn = vb.qn = va.rn = vb.rn = nullptr;
           ra.S = rb.S = 0; ra.je = rb.je = 0; ra.i_lo = rb.i_lo = 1; ra.i_hi = rb.i_hi = 0; ra.gb = rb.gb = 0; ra.d_lo = rb.d_lo = 0;
           ra.d_hi = rb.d_hi = -1; ra.clo = rb.clo = 1; ra.w = rb.w = 0; }
    bool fits = !have || (ra.w <= wcap && rb.w <= wcap && ra.i_hi <= RS && rb.i_hi <= RS);
    if (!fits) { if (g == 0) atomicOr(a.err_flag, 2u); ra.w = rb.w = 0; }
    const int wpair = fits ? max(ra.w, rb.w) : 0;
    int wmax = wpair;                                           // warp-wide step count
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));

    // ---- per-row profiles (rows beyond i_hi are padding: their cells are never read) ----
    uint32_t profA[K], profB[K];
#pragma unroll
    for (int r = 0; r < K; ++r) {
      const int row = g * K + r;
      uint32_t pa = 0x80808080u, pb = 0x80808080u;
      if (row < ra.i_hi) { pa = MM4 ^ (MX << (8u * base2(va.q, row))); if (va.qn && nbit(va.qn, row)) pa = N4; }
      if (row < rb.i_hi) { pb = MM4 ^ (MX << (8u * base2(vb.q, row))); if (vb.qn && nbit(vb.qn, row)) pb = N4; }
      profA[r] = pa; profB[r] = pb;
    }
    // ---- selectors: index = k + G, k = rectangle column (window column clo + k, 1-based) ----
    __syncwarp();
    for (int k0 = 16 * g; k0 < wmax + 2 * G; k0 += 16 * G) {
      const int ca = ra.clo - 1 - G + k0, cb = rb.clo - 1 - G + k0;           // 0-based window columns of index k0
      const bool anyA = have && ca < ra.je && ca + 16 > 0, anyB = have && cb < rb.je && cb + 16 > 0;
      const uint32_t wa = anyA ? load16(va.r, ca) : 0u, wb = anyB ? load16(vb.r, cb) : 0u;
      uint32_t na = 0u, nb = 0u;
      if (HAS_N) { if (va.rn && anyA) na = load16n(va.rn, ca); if (vb.rn && anyB) nb = load16n(vb.rn, cb); }
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        if (k0 + u < sel_stride) {
          uint32_t lo = 0x88u, hi = 0xCCu, nf = 0u;
          if (anyA && ca + u >= ra.clo - 1 && ca + u < ra.je) { const uint32_t c = (wa >> (2 * u)) & 3u; lo = c | ((c | 8u) << 4); if (HAS_N && ((na >> u) & 1u)) nf |= 0x00800000u; }
          if (anyB && cb + u >= rb.clo - 1 && cb + u < rb.je) { const uint32_t c = ((wb >> (2 * u)) & 3u) + 4u; hi = c | ((c | 8u) << 4); if (HAS_N && ((nb >> u) & 1u)) nf |= 0x80000000u; }
          sel[k0 + u] = (SelT)(lo | (hi << 8) | nf);
        }
      }
    }
    __syncwarp();

    // ---- systolic sweep of the rectangle ----
    uint32_t Hoe[K], E[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { Hoe[r] = NEG_OE; E[r] = NEG_OE; }
    uint32_t hoe_bot = NEG_OE, f_bot = NEG_OE, hoe_up_prev = NEG_OE;
    const SelT* selp = sel + (G - g);
    const int nsteps = wmax > 0 ? wmax + G - 1 : 0;
    // lane-level band test: rows (1-based) g*K+1 .. g*K+K at column clo + k touch [d_lo, d_hi] of A or of B
    const int rowlo = g * K + 1, rowhi = g * K + K;
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
      const uint32_t selv = selp[s];
      const uint32_t up = __shfl_up_sync(0xffffffffu, hoe_bot, 1, G) * bmul + badd;
      uint32_t F = __shfl_up_sync(0xffffffffu, f_bot, 1, G) * bmul + badd;
      uint32_t hd = hoe_up_prev;
      hoe_up_prev = up;
      uint32_t isn = 0u;
      if (HAS_N) isn = prmt(selv, 0u, 0xBBAAu);
      uint32_t hv[K];
#pragma unroll
      for (int r = 0; r < K; ++r) {
        uint32_t sp = prmt(profA[r], profB[r], selv & 0xFFFFu);
        if (HAS_N) sp = (sp & ~isn) | (POS_OE & isn);
        const uint32_t t = __vadd2(hd, sp);
        hd = Hoe[r];
        E[r] = __viaddmax_s16x2(E[r], NEG_EXT, Hoe[r]);
        const uint32_t fx = __vadd2(F, NEG_EXT);
        const uint32_t h = __vimax3_s16x2_relu(t, E[r], F);
        Hoe[r] = __vadd2(h, NEG_OE);
        F = __viaddmax_s16x2(h, NEG_OE, fx);
        hv[r] = h;
      }
      hoe_bot = Hoe[K - 1]; f_bot = F;
      const int k = s - g;
      if (k >= 0 && k < wpair) {
        const int ja = ra.clo + k, jb = rb.clo + k;
        const bool inA = (ja - rowlo >= ra.d_lo) && (ja - rowhi <= ra.d_hi) && rowlo <= ra.i_hi && k < ra.w;
        const bool inB = (jb - rowlo >= rb.d_lo) && (jb - rowhi <= rb.d_hi) && rowlo <= rb.i_hi && k < rb.w;
        if (inA || inB) {
          uint4* dst = reinterpret_cast<uint4*>(block + (size_t)k * RS + (size_t)g * K);
#pragma unroll
          for (int r4 = 0; r4 < K / 4; ++r4) dst[r4] = make_uint4(hv[4 * r4], hv[4 * r4 + 1], hv[4 * r4 + 2], hv[4 * r4 + 3]);
        }
      }
    }
    __syncwarp();

    // ---- walks: lane 0 -> task A, lane 1 -> task B ----
    if (have && fits && g < 2 && (g == 0 || haveB)) {
      const RectGeom& rg = (g == 0) ? ra : rb;
      const TaskView& tv = (g == 0) ? va : vb;
      Band bd;
      bd.S = rg.S; bd.je = rg.je; bd.i_lo = rg.i_lo; bd.i_hi = rg.i_hi; bd.d_lo = rg.d_lo; bd.bw = rg.d_hi - rg.d_lo + 1;
      bd.i_first = 1; bd.rows = rg.i_hi; bd.c0 = 0; bd.gb = rg.gb;
      const int shift = (g == 0) ? 0 : 16;
      const int clo = rg.clo, dlo = rg.d_lo;
      walk_h_and_emit<HAS_N, false, 16>(a, (g == 0) ? tA : tB, tv, bd, bd.bw, opbuf, [&](int n, int b) -> uint32_t {
        const int row0 = n, kk = (n + 1) + dlo + b - clo;         // row i = n + 1 (i_first = 1), column j = i + d_lo + b
        const uint32_t v = (__ldcg(block + (size_t)max(kk, 0) * RS + row0) >> shift) & 0xFFFFu;   // unconditional, clamped
        return kk < 0 ? 0u : v;                                    // left of the rectangle: zero boundary
      });
    }
    __syncwarp();
  }
}

}  // namespace ubu
