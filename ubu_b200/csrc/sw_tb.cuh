// sw_tb.cuh -- banded traceback pass: CIGAR, start offsets, edit distance (SPEC.md sections 5-7).
//
// Input per task is the forward pass's (S, je, [i_lo,i_hi]). By the gap bound of SPEC.md section 7
// the SPEC traceback path from any candidate end cell (i, je), i in [i_lo,i_hi], stays inside the
// diagonal band  d = j - i  in  [je - i_hi - Gb, je - i_lo + Gb],  Gb = gap_bound(i_hi, S).
// The band is re-filled with exact 32-bit SPEC arithmetic (cells outside the band / matrix act as
// H = 0, E = F = -inf), 4 direction bits per cell are kept, and the SPEC state machine is replayed
// on those bits. Banded values are <= the full-matrix values everywhere and equal on the SPEC path
// (DESIGN.md section 4), so every decision -- including tie-breaks -- is the full-matrix decision.
//
// Kernels: tb_plan_kernel sorts tasks into band-width classes; tb_band_kernel<BW> keeps the band
// rows in registers (one thread per task, BW = 16 or 32 diagonals); tb_wide_kernel handles any
// wider band with the band rows in a per-thread global scratch (rare: low-scoring tasks).
// Direction bits live in a per-thread global scratch, interleaved so that a warp's row write is
// one contiguous 256-byte store.
#pragma once
#include "sw_common.cuh"

namespace ubu {

constexpr int TB_NEG = -0x10000000;
constexpr int TB_CLASSES = 3;     // 0: band <= 16 (registers), 1: band <= 32 (registers), 2: wide (global scratch)

struct TbArgs {
  DevSeqs reads, wins;
  const uint32_t* ref_idx;
  Scoring sc;
  const FwdOut* fwd;
  ubu_sw_result* res;
  uint32_t* cig_pool;            // device CIGAR pool (bump-allocated per task)
  unsigned int cig_cap;
  unsigned int* cig_cursor;
  unsigned int* err_flag;        // bit 0: pool overflow, bit 1: internal inconsistency
};

// ---- classify ------------------------------------------------------------------------------
__global__ void tb_plan_kernel(const uint32_t* __restrict__ order, uint32_t n_slots, const FwdOut* __restrict__ fwd,
                               Scoring sc, ubu_sw_result* __restrict__ res,
                               uint32_t* __restrict__ lists /* [TB_CLASSES][n_slots] */, unsigned int* __restrict__ counts) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  int cls = -1; uint32_t task = 0xFFFFFFFFu;
  if (idx < n_slots) task = order[idx];
  if (task != 0xFFFFFFFFu) {
    const FwdOut f = fwd[task];
    if (f.score <= 0) {
      ubu_sw_result r; r.score = 0; r.ref_start = r.ref_end = r.q_start = r.q_end = 0; r.edit_distance = 0;
      r.cigar_off = 0; r.cigar_n = 0; r.flags = UBU_SW_FLAG_UNALIGNED;
      res[task] = r;
    } else {
      const int bw = (f.i_hi - f.i_lo) + 2 * gap_bound(sc, f.i_hi, f.score) + 1;
      cls = bw <= 16 ? 0 : (bw <= 32 ? 1 : 2);
    }
  }
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int c = 0; c < TB_CLASSES; ++c) {
    const unsigned m = __ballot_sync(0xffffffffu, cls == c);
    if (m) {
      unsigned base = 0;
      const int leader = __ffs(m) - 1;
      if ((int)lane == leader) base = atomicAdd(&counts[c], __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (cls == c) lists[(size_t)c * n_slots + base + __popc(m & ((1u << lane) - 1u))] = task;
    }
  }
}

// ---- helpers -------------------------------------------------------------------------------
struct TaskView {
  int L, W; const uint8_t *q, *r, *qn, *rn;
};
__device__ __forceinline__ TaskView load_task(const TbArgs& a, uint32_t task) {
  TaskView t;
  const uint32_t ri = a.ref_idx ? a.ref_idx[task] : task;
  t.L = a.reads.len[task]; t.q = a.reads.packed + a.reads.off[task];
  t.W = a.wins.len[ri];    t.r = a.wins.packed + a.wins.off[ri];
  t.qn = a.reads.nmask ? a.reads.nmask + a.reads.nmask_off[task] : nullptr;
  t.rn = a.wins.nmask ? a.wins.nmask + a.wins.nmask_off[ri] : nullptr;
  return t;
}

// Appends (len, op) runs in walk order (back to front); emit() reverses into the pool.
struct OpRuns {
  uint32_t* ops; int n; int cur_op; uint32_t cur_len;
  __device__ __forceinline__ void init(uint32_t* buf) { ops = buf; n = 0; cur_op = -1; cur_len = 0; }
  __device__ __forceinline__ void flush() { if (cur_len) ops[n++] = (cur_len << 4) | (uint32_t)cur_op; cur_len = 0; }
  __device__ __forceinline__ void push(int op) { if (op != cur_op) { flush(); cur_op = op; } ++cur_len; }
  __device__ __forceinline__ void push_run(int op, uint32_t len) { if (len) { flush(); cur_op = -1; ops[n++] = (len << 4) | (uint32_t)op; } }
};

__device__ __forceinline__ void emit_result(const TbArgs& a, uint32_t task, const OpRuns& o, int S, int ref_start, int je,
                                            int q_start, int ie, int edits) {
  ubu_sw_result r;
  r.score = S; r.ref_start = ref_start; r.ref_end = je; r.q_start = q_start; r.q_end = ie; r.edit_distance = edits;
  r.flags = 0; r.cigar_n = (uint16_t)o.n;
  const unsigned off = atomicAdd(a.cig_cursor, (unsigned)o.n);
  r.cigar_off = off;
  if (off + (unsigned)o.n <= a.cig_cap) {
    for (int k = 0; k < o.n; ++k) a.cig_pool[off + k] = o.ops[o.n - 1 - k];
  } else {
    atomicOr(a.err_flag, 1u);
  }
  a.res[task] = r;
}

// Replays the SPEC section 5 state machine on the stored direction nibbles.
//   nibble: bits 0-1 = H source (0 stop, 1 diagonal, 2 F, 3 E), bit 2 = F came by extension, bit 3 = E came by extension
// DirAt(row_index, b) returns the nibble of band cell b in row i (row_index = i - i_first).
template <class DirAt>
__device__ __forceinline__ void walk_and_emit(const TbArgs& a, uint32_t task, const TaskView& t, int S, int ie, int je,
                                              int d_lo, int i_first, int bw, uint32_t* opbuf, DirAt dir_at) {
  OpRuns o; o.init(opbuf);
  o.push_run(UBU_SW_CIGAR_S, (uint32_t)(t.L - ie));
  int i = ie, j = je, state = 0, edits = 0;
  bool bad = false;
  while (true) {
    if (i < 1 || j < 1) { if (state != 0) bad = true; break; }
    const int b = j - i - d_lo;
    if (b < 0 || b >= bw || i < i_first) { bad = true; break; }
    const uint32_t nib = dir_at(i - i_first, b);
    if (state == 0) {
      const uint32_t src = nib & 3u;
      if (src == 0u) break;
      if (src == 1u) {
        o.push(UBU_SW_CIGAR_M);
        const uint32_t qc = base2(t.q, (uint32_t)(i - 1)), rc = base2(t.r, (uint32_t)(j - 1));
        const bool isn = (t.qn && nbit(t.qn, (uint32_t)(i - 1))) || (t.rn && nbit(t.rn, (uint32_t)(j - 1)));
        if (qc != rc && !isn) ++edits;
        --i; --j;
      } else state = (src == 2u) ? 1 : 2;
    } else if (state == 1) {
      o.push(UBU_SW_CIGAR_I); ++edits;
      --i; if (!(nib & 4u)) state = 0;
    } else {
      o.push(UBU_SW_CIGAR_D); ++edits;
      --j; if (!(nib & 8u)) state = 0;
    }
  }
  o.flush();
  o.push_run(UBU_SW_CIGAR_S, (uint32_t)i);     // q_start - 1 = i
  if (bad) atomicOr(a.err_flag, 2u);
  emit_result(a, task, o, S, j + 1, je, i + 1, ie, edits);
}

// ---- register band kernel ------------------------------------------------------------------
// dir scratch: word (row*NW + w) of thread gtid at dir[((row*NW + w) * nthreads) + gtid]
template <int BW, bool HAS_N, int THREADS>
__global__ void __launch_bounds__(THREADS)
tb_band_kernel(TbArgs a, const uint32_t* __restrict__ list, const unsigned int* __restrict__ count,
               unsigned long long* __restrict__ dir, int max_rows) {
  constexpr int NW = BW / 16;
  const uint32_t nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = *count;
  const int ma = a.sc.match, mm = a.sc.mismatch, ext = a.sc.gap_ext, oe = a.sc.gap_open + a.sc.gap_ext;
  uint32_t opbuf[BW + 8];

  for (uint32_t idx = gtid; idx < n; idx += nthreads) {
    const uint32_t task = list[idx];
    const FwdOut f = a.fwd[task];
    const TaskView t = load_task(a, task);
    const int S = f.score, je = f.je, i_lo = f.i_lo, i_hi = f.i_hi;
    const int Gb = gap_bound(a.sc, i_hi, S);
    const int d_lo = je - i_hi - Gb;
    const int i_first = max(1, 2 - d_lo - BW);                 // first row whose band touches column 1
    if (i_hi - i_first + 1 > max_rows) { atomicOr(a.err_flag, 2u); continue; }

    // sliding window of the window's bases for the band cells of the current row (16 bases / word)
    uint32_t rwin[NW]; uint32_t nwin = 0u;
#pragma unroll
    for (int w = 0; w < NW; ++w) rwin[w] = 0u;
#pragma unroll
    for (int b = 0; b < BW; ++b) {
      const int c = i_first + d_lo + b - 1;
      if (c >= 0 && c < t.W) {
        rwin[b / 16] |= base2(t.r, (uint32_t)c) << (2 * (b % 16));
        if (HAS_N && t.rn) nwin |= nbit(t.rn, (uint32_t)c) << b;
      }
    }
    int H[BW], F[BW];
#pragma unroll
    for (int b = 0; b < BW; ++b) { H[b] = 0; F[b] = TB_NEG; }
    int ie = 0;

    for (int i = i_first; i <= i_hi; ++i) {
      const uint32_t qc = base2(t.q, (uint32_t)(i - 1));
      const bool qn = HAS_N && t.qn && nbit(t.qn, (uint32_t)(i - 1));
      const int bmin = max(0, 1 - i - d_lo), bmax = min(BW - 1, je - i - d_lo);
      int hleft = 0, eleft = TB_NEG, hend = -1;
      unsigned long long word = 0ull;
      const size_t rowbase = (size_t)(i - i_first) * NW;
#pragma unroll
      for (int b = 0; b < BW; ++b) {
        const int upH = (b + 1 < BW) ? H[b + 1] : 0, upF = (b + 1 < BW) ? F[b + 1] : TB_NEG;
        const int fe = upF - ext, fo = upH - oe;
        int fv = max(fe, fo);
        const int ee = eleft - ext, eo = hleft - oe;
        int ev = max(ee, eo);
        const bool ismatch = (((rwin[b / 16] >> (2 * (b % 16))) & 3u) == qc);
        int s = ismatch ? ma : mm;
        if (HAS_N) { if (qn || ((nwin >> b) & 1u)) s = 0; }
        const int d = H[b] + s;
        int h = max(max(d, ev), max(fv, 0));
        uint32_t nib = (h == 0) ? 0u : (h == d) ? 1u : (h == fv) ? 2u : 3u;
        nib |= (fe >= fo ? 4u : 0u) | (ee >= eo ? 8u : 0u);
        if (b < bmin || b > bmax) { h = 0; ev = TB_NEG; fv = TB_NEG; nib = 0u; }
        if (b == bmax) hend = h;
        word |= (unsigned long long)nib << (4 * (b % 16));
        H[b] = h; F[b] = fv; hleft = h; eleft = ev;
        if ((b % 16) == 15) { dir[(rowbase + b / 16) * nthreads + gtid] = word; word = 0ull; }
      }
      if (ie == 0 && i >= i_lo && hend == S) ie = i;           // smallest i with H[i][je] == S
      // slide the base window one column to the right
      const int cn = i + d_lo + BW - 1;                        // 0-based column entering at b = BW-1 for row i+1
      uint32_t nb = 0u, nn = 0u;
      if (cn >= 0 && cn < t.W) { nb = base2(t.r, (uint32_t)cn); if (HAS_N && t.rn) nn = nbit(t.rn, (uint32_t)cn); }
#pragma unroll
      for (int w = 0; w < NW; ++w) rwin[w] = (rwin[w] >> 2) | ((w + 1 < NW ? rwin[w + 1] : nb) << 30);
      if (HAS_N) nwin = (nwin >> 1) | (nn << (BW - 1));
    }
    if (ie == 0) { atomicOr(a.err_flag, 2u); continue; }
    walk_and_emit(a, task, t, S, ie, je, d_lo, i_first, BW, opbuf,
                  [&](int row, int b) -> uint32_t {
                    return (uint32_t)(dir[((size_t)row * NW + b / 16) * nthreads + gtid] >> (4 * (b % 16))) & 15u;
                  });
  }
}

// ---- wide band kernel (band rows in global scratch; any width) --------------------------------
// Per-thread scratch, contiguous: [Hs: bw_cap ints][Fs: bw_cap ints][ops: max_rows + 2*bw_cap + 8][dir: max_rows * nw_cap u64]
__host__ __device__ inline size_t tb_wide_scratch_bytes(int max_rows, int bw_cap) {
  const size_t nw = (size_t)(bw_cap + 15) / 16;
  size_t b = (size_t)bw_cap * 8 + ((size_t)max_rows + 2 * (size_t)bw_cap + 8) * 4;
  b = (b + 7) & ~(size_t)7;
  return b + (size_t)max_rows * nw * 8;
}

__global__ void __launch_bounds__(64)
tb_wide_kernel(TbArgs a, const uint32_t* __restrict__ list, const unsigned int* __restrict__ count,
               unsigned char* __restrict__ scratch, int max_rows, int bw_cap) {
  const uint32_t nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = *count;
  const int ma = a.sc.match, mm = a.sc.mismatch, ext = a.sc.gap_ext, oe = a.sc.gap_open + a.sc.gap_ext;
  const size_t per = tb_wide_scratch_bytes(max_rows, bw_cap);
  unsigned char* mine = scratch + (size_t)gtid * per;
  int* Hs = reinterpret_cast<int*>(mine);
  int* Fs = Hs + bw_cap;
  uint32_t* opbuf = reinterpret_cast<uint32_t*>(Fs + bw_cap);
  size_t head = (size_t)bw_cap * 8 + ((size_t)max_rows + 2 * (size_t)bw_cap + 8) * 4; head = (head + 7) & ~(size_t)7;
  unsigned long long* dir = reinterpret_cast<unsigned long long*>(mine + head);

  for (uint32_t idx = gtid; idx < n; idx += nthreads) {
    const uint32_t task = list[idx];
    const FwdOut f = a.fwd[task];
    const TaskView t = load_task(a, task);
    const int S = f.score, je = f.je, i_lo = f.i_lo, i_hi = f.i_hi;
    const int Gb = gap_bound(a.sc, i_hi, S);
    // clip the band to the diagonals that exist left of / above the end cells: d in [1 - i_hi, je - 1]
    const int d_lo = max(je - i_hi - Gb, 1 - i_hi);
    const int d_hi = min(je - i_lo + Gb, je - 1);
    const int bw = d_hi - d_lo + 1;
    const int nw = (bw + 15) / 16;
    const int i_first = max(1, 2 - d_lo - bw);
    if (bw > bw_cap || bw < 1 || i_hi - i_first + 1 > max_rows) { atomicOr(a.err_flag, 2u); continue; }
    for (int b = 0; b < bw; ++b) { Hs[b] = 0; Fs[b] = TB_NEG; }
    int ie = 0;
    for (int i = i_first; i <= i_hi; ++i) {
      const uint32_t qc = base2(t.q, (uint32_t)(i - 1));
      const bool qn = t.qn && nbit(t.qn, (uint32_t)(i - 1));
      const int bmin = max(0, 1 - i - d_lo), bmax = min(bw - 1, je - i - d_lo);
      int hleft = 0, eleft = TB_NEG, hend = -1;
      unsigned long long word = 0ull;
      unsigned long long* drow = dir + (size_t)(i - i_first) * nw;
      // cells left of bmin are outside the matrix: they keep H = 0, F = -inf from the reset below
      for (int b = 0; b < bmin && b < bw; ++b) { Hs[b] = 0; Fs[b] = TB_NEG; }
      for (int w = 0; w < nw; ++w) drow[w] = 0ull;
      int diagH = (bmin < bw) ? Hs[bmin] : 0;                // H[i-1][j-1] for b = bmin (old value at b)
      for (int b = bmin; b <= bmax; ++b) {
        const int upH = (b + 1 < bw) ? Hs[b + 1] : 0, upF = (b + 1 < bw) ? Fs[b + 1] : TB_NEG;
        const int fe = upF - ext, fo = upH - oe, fv = max(fe, fo);
        const int ee = eleft - ext, eo = hleft - oe, ev = max(ee, eo);
        const int c = i + d_lo + b - 1;                       // 0-based window column
        int s = (base2(t.r, (uint32_t)c) == qc) ? ma : mm;
        if (qn || (t.rn && nbit(t.rn, (uint32_t)c))) s = 0;
        const int d = diagH + s;
        const int h = max(max(d, ev), max(fv, 0));
        uint32_t nib = (h == 0) ? 0u : (h == d) ? 1u : (h == fv) ? 2u : 3u;
        nib |= (fe >= fo ? 4u : 0u) | (ee >= eo ? 8u : 0u);
        if (b == bmax) hend = h;
        word |= (unsigned long long)nib << (4 * (b % 16));
        if ((b % 16) == 15 || b == bmax) { drow[b / 16] |= word; word = 0ull; }
        diagH = upH;
        Hs[b] = h; Fs[b] = fv; hleft = h; eleft = ev;
      }
      for (int b = bmax + 1; b < bw; ++b) { Hs[b] = 0; Fs[b] = TB_NEG; }   // right of column je: outside
      if (ie == 0 && i >= i_lo && hend == S) ie = i;
    }
    if (ie == 0) { atomicOr(a.err_flag, 2u); continue; }
    walk_and_emit(a, task, t, S, ie, je, d_lo, i_first, bw, opbuf,
                  [&](int row, int b) -> uint32_t {
                    return (uint32_t)(dir[(size_t)row * nw + b / 16] >> (4 * (b % 16))) & 15u;
                  });
  }
}

}  // namespace ubu
