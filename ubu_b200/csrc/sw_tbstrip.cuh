// sw_tbstrip.cuh -- traceback fill + walk for WIDE bands (more than 32 diagonals): one THREAD per task pair, the band
// swept in horizontal strips of K rows that live in registers.
//
// Why not the forward kernel's systolic layout (a group of lanes per pair)? A band of bw diagonals over `rows` rows holds
// rows x bw cells, but a wavefront over it takes rows + bw steps whatever the lane count, so the lanes idle most of the
// time, and afterwards one lane walks while the others wait (round 1's tb_rect_kernel: 0.24 of the roofline on cfg4).
// Here every lane owns a different pair: no shuffles, no idle lanes in the fill, and all 32 lanes walk at once.
//
// FILL. Band row n (read row i_first + n) holds the cells b = 0 .. bw-1 at window column c0 + n + b (Band, sw_tb.cuh).
// Strip r = band rows [rK, rK + K); it is swept column by column over the K + bw - 1 columns its rows touch, left to
// right, with the forward pass's packed 16-bit cell (task A in the low, task B in the high half; Hoe[K], E[K] in
// registers). What a strip needs from the strip above -- H - oe and the F entering its first row, per column -- goes
// through a small per-thread boundary array in global memory (read one step ahead, written in place). Columns right of
// the end column je are never visited by the walk and are skipped. The region actually computed (the union of the strip
// rectangles, zero boundary around it) lies between the band and the full matrix, so its values are sandwiched between
// the banded and the full-matrix values and the exactness argument of DESIGN.md section 4.2 applies unchanged.
// Every computed H is stored: one byte per task and cell when no score of the batch can exceed 255 (H <= S <= match * L:
// all BASELINE configs), else 16 bits; a warp's stores are 16-byte chunks, lane-minor (512 contiguous bytes per store).
// Selectors are derived on the fly from the window words (16 columns per pair of loads), so no shared memory is used.
//
// WALK. walk_h_and_emit (sw_tb.cuh) on the stored H, diagonal test from the sequences, 17 cells of the current diagonal per
// round trip; each lane walks its task A, then its task B.
//
// SCHEDULING. Tasks are filed by band width in the forward epilogue; a launch serves a range of those classes, widest first, so
// that a warp's 32 pairs have similar trip counts; warps draw units of 32 pairs from a device counter. The host issues two
// launches side by side: one for bands up to 120 diagonals (scratch sized for that width, full grid) and a small one for the
// rare wider bands (scratch sized for the widest band the batch's scoring allows).
#pragma once
#include "sw_common.cuh"
#include "sw_tb.cuh"

namespace ubu {

constexpr int TBS_K = 16;                 // rows per strip
constexpr int TBS_THREADS = 128;
constexpr int TBS_BW_MAX = 288;          // widest band taken (the walk keeps <= bw + 8 CIGAR runs in local memory)
#ifndef UBU_TBS_WB
#define UBU_TBS_WB 16                    // cells of the current diagonal fetched per round trip of the walk
#endif
#ifndef UBU_TBS_MIN_BLOCKS
#define UBU_TBS_MIN_BLOCKS 3
#endif

// per-warp scratch in bytes: H store of 32 pairs + their boundary arrays
__host__ __device__ inline size_t tbs_hstore_bytes_per_warp(int rows_cap, int bw_cap, int bytes_per_cell) {
  const size_t strips = (size_t)(rows_cap + TBS_K - 1) / TBS_K, steps = (size_t)(TBS_K + bw_cap - 1);
  return strips * steps * TBS_K * 2u * (size_t)bytes_per_cell * 32u;
}
__host__ __device__ inline size_t tbs_boundary_bytes_per_warp(int rows_cap, int bw_cap) {
  return ((size_t)((rows_cap + TBS_K - 1) / TBS_K) * TBS_K + (size_t)(TBS_K + bw_cap)) * 8u * 32u;
}

struct StripArgs {
  const uint32_t* lists;          // [TB_CLASSES][n_slots]
  const unsigned int* firsts;     // [TB_CLASSES] this launch serves lists[c][firsts[c], counts[c])
  const unsigned int* counts;     // [TB_CLASSES]
  unsigned int* work;             // unit counter of this launch (zeroed before it)
  int cls_lo, cls_hi;             // traceback classes served (a launch for the common widths, one for the rare very wide bands)
  uint32_t n_slots;
  unsigned char* hstore;          // [warps of the grid] x tbs_hstore_bytes_per_warp
  uint2* boundary;                // [warps of the grid] x tbs_boundary_bytes_per_warp
  int rows_cap, bw_cap;
  size_t h_stride, b_stride;      // bytes per warp of hstore / boundary (>= tbs_*_bytes_per_warp)
};

template <bool HAS_N, bool DEFAULT_SC, int BYTES>
__global__ void __launch_bounds__(TBS_THREADS, UBU_TBS_MIN_BLOCKS)
tb_strip_kernel(TbArgs a, StripArgs sa) {
  constexpr int K = TBS_K;
  static_assert(BYTES == 1 || BYTES == 2, "one or two bytes per stored H");
  constexpr int CHUNKS = K * 2 * BYTES / 16;                     // 16-byte chunks per lane per step (2 or 4)
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * TBS_THREADS + threadIdx.x) >> 5;
  Scoring sc = a.sc;
  if (DEFAULT_SC) { sc.match = 1; sc.mismatch = -3; sc.gap_open = 5; sc.gap_ext = 2; }
  const int oe = sc.gap_open + sc.gap_ext;
  const uint32_t mmb = (uint32_t)(sc.mismatch + oe) & 0xFFu, mab = (uint32_t)(sc.match + oe) & 0xFFu;
  const uint32_t MM4 = mmb * 0x01010101u, MX = mab ^ mmb, N4 = ((uint32_t)oe & 0xFFu) * 0x01010101u;
  const uint32_t NEG_OE = DEFAULT_SC ? 0xFFF9FFF9u : sc.neg_oe2, NEG_EXT = DEFAULT_SC ? 0xFFFEFFFEu : sc.neg_ext2,
                 POS_OE = DEFAULT_SC ? 0x00070007u : sc.pos_oe2;
  unsigned char* hwarp = sa.hstore + (size_t)warp_global * sa.h_stride;
  uint2* bnd = sa.boundary + (size_t)warp_global * (sa.b_stride / 8u) + lane;
  uint4* hlane = reinterpret_cast<uint4*>(hwarp) + lane;         // chunk (step, c) of this lane at hlane[(step * CHUNKS + c) * 32]
  uint32_t opbuf[TBS_BW_MAX + 16];                               // CIGAR runs of one walk (local memory; <= 2 * gaps + 3 <= bw + 2 runs)
  // boundary staging: per warp a double buffer of 16 steps x 32 lanes x 8 bytes
  __shared__ __align__(16) uint2 sm_all[(TBS_THREADS / 32) * 2 * 16 * 32];
  const uint2* sm_mine = sm_all + (threadIdx.x >> 5) * (2 * 16 * 32) + lane;
  const uint32_t sm_bnd = (uint32_t)__cvta_generic_to_shared(sm_mine);

  // units of 32 pairs over the classes this launch serves, widest class first
  unsigned int total_units = 0;
  for (int c = sa.cls_hi; c >= sa.cls_lo; --c) total_units += ((sa.counts[c] - sa.firsts[c] + 1u) / 2u + 31u) / 32u;

  for (;;) {
    unsigned int unit = 0;
    if (lane == 0) unit = atomicAdd(sa.work, 1u);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= total_units) break;
    int cls = sa.cls_hi; unsigned int ubase = 0, n_tasks = sa.counts[cls] - sa.firsts[cls];
    for (;;) {                                                   // the class this unit falls into
      const unsigned int nu = ((n_tasks + 1u) / 2u + 31u) / 32u;
      if (unit < ubase + nu || cls == sa.cls_lo) break;
      ubase += nu; --cls; n_tasks = sa.counts[cls] - sa.firsts[cls];
    }
    const uint32_t* list = sa.lists + (size_t)cls * sa.n_slots + sa.firsts[cls];
    const unsigned int n_pairs = (n_tasks + 1u) / 2u;
    const unsigned int pair = (unit - ubase) * 32u + (unsigned)lane;
    const bool have = pair < n_pairs, haveB = have && 2u * pair + 1u < n_tasks;
    const uint32_t tA = have ? list[2u * pair] : 0u, tB = haveB ? list[2u * pair + 1u] : tA;

    TaskView va, vb; Band ga, gb;
    int bw = 0, nrows = 0;
    bool ok = have;
    if (have) {
      va = load_task(a, tA); vb = load_task(a, tB);
      const FwdOut fa = a.fwd[tA], fb = a.fwd[tB];
      bw = max((fa.i_hi - fa.i_lo) + 2 * gap_bound(sc, fa.i_hi, fa.score) + 1, (fb.i_hi - fb.i_lo) + 2 * gap_bound(sc, fb.i_hi, fb.score) + 1);
      ga = make_band(sc, fa, bw); gb = make_band(sc, fb, bw);
      nrows = max(ga.rows, gb.rows);
      if (bw > sa.bw_cap || nrows > sa.rows_cap || bw > TBS_BW_MAX) { atomicOr(a.err_flag, 2u); ok = false; }
    }
    if (!ok) { bw = 0; nrows = 0; }
    int pw = ok ? K + bw - 1 : 0;                                // steps per strip, warp-wide maximum = row pitch of the store
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pw = max(pw, __shfl_xor_sync(0xffffffffu, pw, o));
    const int nstrips = (nrows + K - 1) / K;
    // columns right of je are never visited: per strip, steps p < clip - rK only (clip = steps until the pair's last end column)
    const int clip = ok ? max(ga.je - ga.c0, gb.je - gb.c0) : 0;

    int p_prev = 0;
    uint32_t corner = NEG_OE;                                    // H[last row of the strip above][column left of step 0] - oe
    uint32_t qa_n = 0u, qb_n = 0u, na_n = 0u, nb_n = 0u;         // read words of the next strip, fetched a strip ahead
    auto load_rows = [&](int r) {
      const int ia = ga.i_first - 1 + r * K, ib = gb.i_first - 1 + r * K;      // 0-based read position of row 0 of the strip
      qa_n = (ok && ia < va.L) ? load16(va.q, ia) : 0u; qb_n = (ok && ib < vb.L) ? load16(vb.q, ib) : 0u;
      if (HAS_N) { na_n = (ok && va.qn && ia < va.L) ? load16n(va.qn, ia) : 0u; nb_n = (ok && vb.qn && ib < vb.L) ? load16n(vb.qn, ib) : 0u; }
    };
    load_rows(0);
    for (int r = 0; r < nstrips; ++r) {
      const int psteps = min(K + bw - 1, clip - r * K);
      const uint32_t qa = qa_n, qb = qb_n, na = na_n, nb = nb_n;
      if (r + 1 < nstrips) load_rows(r + 1);
      if (psteps <= 0) { p_prev = 0; continue; }
      // ---- row profiles of the strip (16 rows = one load16 per task) ----
      uint32_t profA[K], profB[K];
      {
        const int ia = ga.i_first - 1 + r * K, ib = gb.i_first - 1 + r * K;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          profA[k] = profile_of((qa >> (2 * k)) & 3u, (na >> k) & 1u, ia + k < va.L, MM4, MX, N4);
          profB[k] = profile_of((qb >> (2 * k)) & 3u, (nb >> k) & 1u, ib + k < vb.L, MM4, MX, N4);
        }
      }
      uint32_t Hoe[K], E[K];
#pragma unroll
      for (int k = 0; k < K; ++k) { Hoe[k] = NEG_OE; E[k] = NEG_OE; }
      const int ca0 = ga.c0 + r * K, cb0 = gb.c0 + r * K;       // 0-based window column of step 0
      uint2* brow = bnd + (size_t)(r * K) * 32;                  // boundary slot of step p: brow[p * 32]
      const int top_valid = (r > 0) ? p_prev - K : 0;            // steps p < top_valid have a real strip above them
      uint32_t hoe_up_prev = (r > 0 && p_prev >= K) ? corner : NEG_OE;
      uint4* hstrip = hlane + (size_t)(r * pw) * CHUNKS * 32;
      // Everything a block of 16 steps reads from memory is requested one block ahead: the boundary values by cp.async into
      // this warp's double buffer in shared memory (no registers held while they fly), the window words into registers.
      Load16 la, lb;                                             // window words of the next block, in flight
      uint32_t fa_n = 0u, fb_n = 0u;
      auto prefetch = [&](int blk) {
        const int p0 = blk * 16;
        if (p0 < top_valid) {
          const uint32_t dst = sm_bnd + (uint32_t)((blk & 1) * 16 * 32 * 8);
          const uint2* src = brow + (size_t)p0 * 32;
#pragma unroll
          for (int u = 0; u < 16; ++u)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst + (uint32_t)(u * 32 * 8)), "l"(src + (size_t)u * 32) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        const int ca = ca0 + p0, cb = cb0 + p0;
        la.issue(va.r, ca, ca < va.W && ca + 16 > 0); lb.issue(vb.r, cb, cb < vb.W && cb + 16 > 0);
        if (HAS_N) { fa_n = (va.rn && ca < va.W && ca + 16 > 0) ? load16n(va.rn, ca) : 0u; fb_n = (vb.rn && cb < vb.W && cb + 16 > 0) ? load16n(vb.rn, cb) : 0u; }
      };
      prefetch(0);
      const int nblk = (psteps + 15) / 16;
#pragma unroll 1
      for (int blk = 0; blk < nblk; ++blk) {
        const uint32_t wa = la.finish(), wb = lb.finish(), fa_ = fa_n, fb_ = fb_n;
        if (blk + 1 < nblk) prefetch(blk + 1); else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // this block's boundary values have landed
        const uint2* sblk = sm_mine + (blk & 1) * 16 * 32;
        const int pend = min(16, psteps - blk * 16);
#pragma unroll 1
        for (int u = 0; u < pend; ++u) {
          const int p = blk * 16 + u;
          uint2 top = sblk[u * 32];
          if (p >= top_valid) top = make_uint2(NEG_OE, NEG_OE);
          const int ca = ca0 + p, cb = cb0 + p;
          uint32_t lo = 0x88u, hi = 0xCCu, nf = 0u;               // outside the window: sign-replicate selectors (score <= 0 after bias)
          if (ca >= 0 && ca < va.W) { lo = ((wa >> (2 * u)) & 3u) * 0x11u + 0x80u; if (HAS_N && ((fa_ >> u) & 1u)) nf |= 0x00800000u; }
          if (cb >= 0 && cb < vb.W) { hi = ((wb >> (2 * u)) & 3u) * 0x11u + 0xC4u; if (HAS_N && ((fb_ >> u) & 1u)) nf |= 0x80000000u; }
          const uint32_t selv = lo | (hi << 8);
          uint32_t isn = 0u;
          if (HAS_N) isn = prmt(nf, 0u, 0xBBAAu);
          uint32_t hd = hoe_up_prev;
          hoe_up_prev = top.x;
          uint32_t F = top.y;
          uint32_t hv[K];
#pragma unroll
          for (int k = 0; k < K; ++k) {
            uint32_t sp = prmt(profA[k], profB[k], selv);
            if (HAS_N) sp = (sp & ~isn) | (POS_OE & isn);
            const uint32_t t = __vadd2(hd, sp);
            hd = Hoe[k];
            E[k] = __viaddmax_s16x2(E[k], NEG_EXT, Hoe[k]);
            const uint32_t fx = __vadd2(F, NEG_EXT);
            const uint32_t h = __vimax3_s16x2_relu(t, E[k], F);
            Hoe[k] = __vadd2(h, NEG_OE);
            F = __viaddmax_s16x2(h, NEG_OE, fx);
            hv[k] = h;
          }
          brow[p * 32] = make_uint2(Hoe[K - 1], F);
          if (p == K - 1) corner = Hoe[K - 1];
          uint4* dst = hstrip + (size_t)p * CHUNKS * 32;
          if (BYTES == 1) {
            // rows 2j, 2j+1 share a word: bytes (A 2j, A 2j+1, B 2j, B 2j+1); one multiply-add per word on the FMA pipe
            uint32_t wv[K / 2];
#pragma unroll
            for (int j = 0; j < K / 2; ++j) wv[j] = hv[2 * j + 1] * 256u + hv[2 * j];
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) dst[c * 32] = make_uint4(wv[4 * c], wv[4 * c + 1], wv[4 * c + 2], wv[4 * c + 3]);
          } else {
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) dst[c * 32] = make_uint4(hv[4 * c], hv[4 * c + 1], hv[4 * c + 2], hv[4 * c + 3]);
          }
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      p_prev = psteps;
    }

    // ---- walks: every lane walks its task A, then its task B, on its own stored cells ----
    if (ok) {
      const unsigned char* hbytes = reinterpret_cast<const unsigned char*>(hlane);
      auto cell = [&](int n, int b, int half) -> uint32_t {
        const int r = n / K, k = n % K;
        const size_t step = (size_t)r * pw + (size_t)(k + b);
        if (BYTES == 1) {
          const size_t chunk = (step * CHUNKS + (size_t)(k >> 3)) * 32;                     // in 16-byte units, relative to hlane
          return (uint32_t)hbytes[chunk * 16 + (size_t)(((k & 7) >> 1) * 4 + half * 2 + (k & 1))];
        } else {
          const size_t chunk = (step * CHUNKS + (size_t)(k >> 2)) * 32;
          return (uint32_t)reinterpret_cast<const uint16_t*>(hbytes + chunk * 16)[(k & 3) * 2 + half];
        }
      };
#pragma unroll 1
      for (int half = 0; half < (haveB ? 2 : 1); ++half)         // one copy of the walk's code for both tasks
        walk_h_and_emit<HAS_N, false, UBU_TBS_WB>(a, half ? tB : tA, half ? vb : va, half ? gb : ga, bw, opbuf,
                                          [&](int n, int b) -> uint32_t { return cell(n, b, half); });
    }
    __syncwarp();
  }
}

}  // namespace ubu
