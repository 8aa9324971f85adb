// sw_long.cuh -- reads longer than the packed 16-bit classes take (L > 256 rows): the contig-sized queries of the reference's
// Aligner.align (bwa bwasw on assembled contigs of 100 bp .. 2.5 kb, Aligner.java:18-22, ReAligner.java:44-46,351-367).
//
// These are few (one per assembled contig, not one per read) and long, so the kernel is built for generality, not for the
// roofline: 32-bit arithmetic (no score can overflow: match <= 31, L <= 4096), one WARP per task, the whole matrix.
// Lane l owns LONG_K consecutive rows of the current stripe of 32 x LONG_K rows and sweeps the window as a skewed wavefront
// (lane l is at column step - l + 1); what a lane needs from the row above its rows (H, F) and the window base of its column
// arrive by warp shuffle from lane l-1, which was at that column one step earlier. Lane 0 takes them from the stripe above (a per-warp boundary row in global
// memory, written in place by lane 31; read 32 columns per coalesced load and handed to lane 0 by shuffle) and from the packed
// window (same scheme). Every cell stores a direction nibble (8 cells per 32-bit word per lane) into a per-warp L x W
// matrix: bits 0-1 = where H came from (0 stop, 1 diagonal, 2 F, 3 E), bit 2 = F was extended, bit 3 = E was extended --
// the encoding of tb_wide_kernel, so the walk replays SPEC.md section 5 without any value comparison. The end cell is the
// warp maximum of (score, -j, -i) = SPEC section 4. Lane 0 walks.
#pragma once
#include "sw_common.cuh"
#include "sw_tb.cuh"

namespace ubu {

constexpr int LONG_THREADS = 128;
#ifndef UBU_LONG_K
#define UBU_LONG_K 8
#endif
constexpr int LONG_K = UBU_LONG_K;                                // rows per lane: a stripe is 32 x LONG_K rows
constexpr uint32_t LONG_MAX_READ = 4096;                 // = UBU_SW_MAX_READ
constexpr uint64_t LONG_MAX_CELLS = (uint64_t)1 << 26;   // L x W of one task: 32 MB of direction nibbles per warp at most

__host__ __device__ inline size_t long_scratch_bytes(uint32_t L, uint32_t W) {
  const size_t wpr = ((size_t)W + 7) / 8;
  return ((size_t)W + 33) * 8 + ((size_t)L + W + 8) * 4 + (size_t)L * wpr * 4;          // boundary H,F | CIGAR runs | direction words
}

struct LongArgs {
  const uint32_t* list; uint32_t n;          // the long tasks of the batch
  unsigned int* work;                        // task counter (zeroed before the launch)
  unsigned char* scratch; size_t per_warp;   // per-warp scratch, sized for the largest long task of the batch
};

template <bool HAS_N>
__global__ void __launch_bounds__(LONG_THREADS)
sw_long_kernel(TbArgs a, LongArgs la) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp_global = (blockIdx.x * LONG_THREADS + threadIdx.x) >> 5;
  const int ma = a.sc.match, mm = a.sc.mismatch, ext = a.sc.gap_ext, oe = a.sc.gap_open + a.sc.gap_ext;
  unsigned char* mine = la.scratch + (size_t)warp_global * la.per_warp;

  for (;;) {
    unsigned int u = 0;
    if (lane == 0) u = atomicAdd(la.work, 1u);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= la.n) break;
    const uint32_t task = la.list[u];
    const TaskView t = load_task(a, task);
    const int L = t.L, W = t.W;
    if (L == 0 || W == 0) {
      if (lane == 0) {
        ubu_sw_result r; r.score = 0; r.ref_start = r.ref_end = r.q_start = r.q_end = 0; r.edit_distance = 0;
        r.cigar_off = 0; r.cigar_n = 0; r.flags = UBU_SW_FLAG_UNALIGNED; a.res[task] = r;
      }
      continue;
    }
    const int wpr = (W + 7) / 8;
    int* Hb = reinterpret_cast<int*>(mine);                    // boundary row: H[stripe above][j], j = 0 .. W (+32 slack)
    int* Fb = Hb + (W + 33);
    uint32_t* opbuf = reinterpret_cast<uint32_t*>(Fb + (W + 33));
    uint32_t* dir = opbuf + (L + W + 8);
    for (int j = lane; j < W + 33; j += 32) { Hb[j] = 0; Fb[j] = TB_NEG; }
    __syncwarp();

    int best_h = 0, best_j = 0, best_i = 0;                    // this lane's best cell: highest score, then smallest column, then smallest row
    for (int s0 = 0; s0 < L; s0 += 32 * LONG_K) {
      const int i0 = s0 + lane * LONG_K + 1;                   // 1-based first row of this lane
      const int nact = max(0, min(LONG_K, L - (i0 - 1)));      // rows of this lane inside the read
      uint32_t qcodes = 0u, qnbits = 0u;                       // 2-bit codes / N flags of the lane's rows
#pragma unroll
      for (int k = 0; k < LONG_K; ++k) {
        if (k < nact) {
          qcodes |= base2(t.q, (uint32_t)(i0 - 1 + k)) << (2 * k);
          if (HAS_N && t.qn && nbit(t.qn, (uint32_t)(i0 - 1 + k))) qnbits |= 1u << k;
        }
      }
      int hl[LONG_K], ee[LONG_K]; uint32_t wd[LONG_K];         // H[i][j-1], E[i][j-1], direction nibbles being collected, per row
#pragma unroll
      for (int k = 0; k < LONG_K; ++k) { hl[k] = 0; ee[k] = TB_NEG; wd[k] = 0u; }
      int dgtop = 0;                                           // H[i0-1][j-1]
      int out_h = 0, out_f = TB_NEG; uint32_t out_c = 0u;      // what this lane hands to the lane below: H, F of its bottom row, window base (+4 = N)
      uint32_t* drow = dir + (size_t)(i0 - 1) * wpr;
      // blocks of 32 columns for lane 0: boundary values and window bases, one coalesced load per block, one block ahead
      int hb_n = Hb[1 + lane], fb_n = Fb[1 + lane];
      uint32_t rc_n = (lane < W) ? (base2(t.r, (uint32_t)lane) | ((HAS_N && t.rn && nbit(t.rn, (uint32_t)lane)) ? 4u : 0u)) : 0u;
      int hb = 0, fb = TB_NEG; uint32_t rcb = 0u;
      const int nsteps = W + 31;
      for (int step = 0; step < nsteps; ++step) {
        if ((step & 31) == 0) {
          hb = hb_n; fb = fb_n; rcb = rc_n;
          const int jn = step + 32 + lane;                     // 0-based column of the next block's entry of this lane
          hb_n = Hb[min(jn + 1, W + 32)]; fb_n = Fb[min(jn + 1, W + 32)];
          rc_n = (jn < W) ? (base2(t.r, (uint32_t)jn) | ((HAS_N && t.rn && nbit(t.rn, (uint32_t)jn)) ? 4u : 0u)) : 0u;
        }
        // from the rows above: lane l-1's bottom-row output of the previous step is column j of row i0-1
        int uh = __shfl_up_sync(0xffffffffu, out_h, 1), uf = __shfl_up_sync(0xffffffffu, out_f, 1);
        uint32_t rc = __shfl_up_sync(0xffffffffu, out_c, 1);
        const int b_h = __shfl_sync(0xffffffffu, hb, step & 31), b_f = __shfl_sync(0xffffffffu, fb, step & 31);
        const uint32_t b_c = __shfl_sync(0xffffffffu, rcb, step & 31);
        if (lane == 0) { uh = b_h; uf = b_f; rc = b_c; }
        const int j = step - lane + 1;                         // 1-based column of this lane
        const bool colvalid = j >= 1 && j <= W;
        const uint32_t sh = 4u * (uint32_t)((j - 1) & 7);
        int dg = dgtop;                                        // H[i-1][j-1] of the row being updated
        dgtop = uh;
#pragma unroll
        for (int k = 0; k < LONG_K; ++k) {
          const int fe = uf - ext, fo = uh - oe, f = max(fe, fo);
          const int e2 = ee[k] - ext, eo = hl[k] - oe, ev = max(e2, eo);
          int s = ((rc & 3u) == ((qcodes >> (2 * k)) & 3u)) ? ma : mm;
          if (HAS_N && (((qnbits >> k) & 1u) || (rc & 4u))) s = 0;
          const int d = dg + s;
          const int h = max(max(d, ev), max(f, 0));
          uint32_t nib = (h == 0) ? 0u : (h == d) ? 1u : (h == f) ? 2u : 3u;
          nib |= (fe >= fo ? 4u : 0u) | (e2 >= eo ? 8u : 0u);
          dg = hl[k];                                          // the old left value is the diagonal of the row below
          const bool v = colvalid && k < nact;
          if (v) {
            hl[k] = h; ee[k] = ev;
            if (h > best_h || (h == best_h && h > 0 && j < best_j)) { best_h = h; best_j = j; best_i = i0 + k; }   // rows ascend: the smaller row keeps a tie
            wd[k] |= nib << sh;
          } else if (!colvalid) { hl[k] = 0; ee[k] = TB_NEG; }   // column 0 / outside: the matrix boundary
          uh = v ? h : 0; uf = v ? f : TB_NEG;
        }
        out_h = uh; out_f = uf; out_c = rc;
        if (colvalid && (((j - 1) & 7) == 7 || j == W)) {
#pragma unroll
          for (int k = 0; k < LONG_K; ++k) if (k < nact) { drow[(size_t)k * wpr + ((j - 1) >> 3)] = wd[k]; wd[k] = 0u; }
        }
        if (lane == 31 && colvalid) { Hb[j] = out_h; Fb[j] = out_f; }   // boundary for the next stripe (lane 0 is 31 columns ahead: in place)
      }
      __syncwarp();
    }
    // SPEC section 4 across the lanes: (score, -j, -i) as one key
    unsigned long long best = best_h > 0 ? (((unsigned long long)(uint32_t)best_h << 32) | ((unsigned long long)(0xFFFFu - (uint32_t)best_j) << 16) |
                                            (unsigned long long)(0xFFFFu - (uint32_t)best_i)) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o); best = other > best ? other : best; }
    __syncwarp();
    if (lane != 0) continue;
    // ---- lane 0: end cell and walk (SPEC.md sections 4 and 5) ----
    if (best == 0ull) {
      ubu_sw_result r; r.score = 0; r.ref_start = r.ref_end = r.q_start = r.q_end = 0; r.edit_distance = 0;
      r.cigar_off = 0; r.cigar_n = 0; r.flags = UBU_SW_FLAG_UNALIGNED; a.res[task] = r;
      continue;
    }
    const int S = (int)(best >> 32), je = (int)(0xFFFFu - (uint32_t)((best >> 16) & 0xFFFFu)), ie = (int)(0xFFFFu - (uint32_t)(best & 0xFFFFu));
    OpRuns o; o.init(opbuf);
    o.push_run(UBU_SW_CIGAR_S, (uint32_t)(L - ie));
    int i = ie, j = je, state = 0, edits = 0;
    bool bad = false;
    while (true) {
      if (i < 1 || j < 1) { if (state != 0) bad = true; break; }
      const uint32_t nib = (__ldcg(dir + (size_t)(i - 1) * wpr + ((j - 1) >> 3)) >> (4 * ((j - 1) & 7))) & 15u;
      if (state == 0) {
        const uint32_t src = nib & 3u;
        if (src == 0u) break;
        if (src == 1u) {
          o.push(UBU_SW_CIGAR_M, 1u);
          const bool isn = HAS_N && ((t.qn && nbit(t.qn, (uint32_t)(i - 1))) || (t.rn && nbit(t.rn, (uint32_t)(j - 1))));
          if (!isn && base2(t.q, (uint32_t)(i - 1)) != base2(t.r, (uint32_t)(j - 1))) ++edits;
          --i; --j;
        } else state = (src == 2u) ? 1 : 2;
      } else if (state == 1) {
        o.push(UBU_SW_CIGAR_I, 1u); ++edits;
        --i; if (!(nib & 4u)) state = 0;
      } else {
        o.push(UBU_SW_CIGAR_D, 1u); ++edits;
        --j; if (!(nib & 8u)) state = 0;
      }
    }
    o.flush();
    o.push_run(UBU_SW_CIGAR_S, (uint32_t)i);
    if (bad || o.n > 65535) { atomicOr(a.err_flag, 2u); continue; }
    emit_result(a, task, o, S, j + 1, je, i + 1, ie, edits);
  }
}

}  // namespace ubu
