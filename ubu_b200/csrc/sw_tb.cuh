// sw_tb.cuh -- banded traceback pass: CIGAR, start offsets, edit distance (SPEC.md sections 5-7).
//
// Input per task is the forward pass's (S, je, [i_lo,i_hi]). By the gap bound of SPEC.md section 7
// the SPEC traceback path from any candidate end cell (i, je), i in [i_lo,i_hi], stays inside the
// diagonal band  d = j - i  in  [je - i_hi - Gb, je - i_lo + Gb],  Gb = gap_bound(i_hi, S).
//
// FILL. The band is re-filled with the same packed 16-bit arithmetic as the forward pass, two tasks
// per thread (task A in the low, task B in the high half), one band row per loop trip; per cell one 16-bit
// value per task is stored: H in bits 0-14 and, in bit 15, "the diagonal does NOT reproduce H" (h != Hdiag + s). Cells left of column 1 see padding selectors (score <= 0
// after bias), so they evaluate to H = 0, E = F = -oe, which is what the matrix boundary needs;
// cells right of column je only feed cells right of column je. Banded values are <= the full-matrix
// values everywhere and equal on the SPEC path.
//
// WALK. The SPEC state machine is replayed from the stored values alone (no sequence access on the way):
//   * stop when the running score reaches 0 (H along the path is the running score);
//   * diagonal first:  H[i-1][j-1] + s(q_i, r_j) == H[i][j]  (the stored flag);
//   * else read-gap (I) if some k has H[i-k][j] - open - k*ext == H[i][j]; SPEC's "extension preferred
//     over gap-close" picks the LARGEST such k (an extension step at row x exists iff some k' >= 2
//     reproduces the value from row x, so the chain runs to the farthest opener);
//   * else window-gap (D), same rule along the row.
// A comparison that holds in the full matrix involves a cell of the SPEC path (exact in the band);
// one that fails in the full matrix fails a fortiori with banded values (they are lower bounds).
//
// Kernels: the forward kernels' epilogue sorts tasks into band-width classes (tb_class_of, sw_common.cuh); tb_diag_kernel
// handles the gap-free class;
// tb_fill_reg_kernel<BW> keeps the band row in registers (BW = 8, 16 or 32 diagonals); wider bands go to
// tb_strip_kernel (sw_tbstrip.cuh); tb_wide_kernel (scalar, global scratch, direction bits) is the last
// resort (extreme scoring parameters / very long windows only).
#pragma once
#include "sw_common.cuh"

#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>

namespace ubu {

constexpr int TB_NEG = -0x10000000;
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
  unsigned int cig_base;         // added to every record's cigar_off: where this batch's ops will sit in the CALLER's pool
                                 // (the multi-GPU driver gives every chunk its own slice, so nothing is rebased on the host)
};

// Band geometry of one task (SPEC.md section 7).
struct Band {
  int S, je, i_lo, i_hi, d_lo, bw, i_first, rows, c0, gb;
};
__host__ __device__ __forceinline__ Band make_band(const Scoring& sc, const FwdOut& f, int bw_fixed) {
  Band g;
  g.S = f.score; g.je = f.je; g.i_lo = f.i_lo; g.i_hi = f.i_hi;
  const int Gb = gap_bound(sc, f.i_hi, f.score);
  g.gb = Gb;
  g.d_lo = f.je - f.i_hi - Gb;
  g.bw = bw_fixed > 0 ? bw_fixed : (f.i_hi - f.i_lo) + 2 * Gb + 1;
  g.i_first = 2 - g.d_lo - g.bw; if (g.i_first < 1) g.i_first = 1;     // first row whose band touches column 1
  g.rows = f.i_hi - g.i_first + 1;
  g.c0 = g.i_first + g.d_lo - 1;                                       // 0-based window column of cell (row 0, b = 0)
  return g;
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

// Appends (len, op) runs in walk order (back to front); emit_result() reverses into the pool.
struct OpRuns {
  uint32_t* ops; int n; int cur_op; uint32_t cur_len;
  __device__ __forceinline__ void init(uint32_t* buf) { ops = buf; n = 0; cur_op = -1; cur_len = 0; }
  __device__ __forceinline__ void flush() { if (cur_len) ops[n++] = (cur_len << 4) | (uint32_t)cur_op; cur_len = 0; }
  __device__ __forceinline__ void push(int op, uint32_t len) { if (op != cur_op) { flush(); cur_op = op; } cur_len += len; }
  __device__ __forceinline__ void push_run(int op, uint32_t len) { if (len) { flush(); cur_op = -1; ops[n++] = (len << 4) | (uint32_t)op; } }
};

__device__ __forceinline__ void emit_result(const TbArgs& a, uint32_t task, const OpRuns& o, int S, int ref_start, int je,
                                            int q_start, int ie, int edits) {
  ubu_sw_result r;
  r.score = S; r.ref_start = ref_start; r.ref_end = je; r.q_start = q_start; r.q_end = ie; r.edit_distance = edits;
  r.flags = 0; r.cigar_n = (uint16_t)o.n;
  const unsigned off = atomicAdd(a.cig_cursor, (unsigned)o.n);
  r.cigar_off = off + a.cig_base;
  if (off + (unsigned)o.n <= a.cig_cap) {
    for (int k = 0; k < o.n; ++k) a.cig_pool[off + k] = o.ops[o.n - 1 - k];
  } else {
    atomicOr(a.err_flag, 1u);
  }
  a.res[task] = r;
}

// Walk on the stored band values (see the header comment). WordAt(n, b) returns the 16-bit value of band cell b in
// band row n (n = i - i_first) for THIS task's half: bits 0-14 = H, bit 15 = diagonal does not reproduce H.
// Rows above i_first are the zero boundary; cells left of column 1 were filled as H = 0.
// USE_FLAG = false: the store holds H only and the diagonal test is made from the sequences (16 bases per load).
// WB = cells fetched per round trip along the current diagonal (a multiple of 16 when !USE_FLAG).
template <bool HAS_N, bool USE_FLAG, int WB, class WordAt>
__device__ __forceinline__ void walk_h_and_emit(const TbArgs& a, uint32_t task, const TaskView& t, const Band& g, int bw,
                                                uint32_t* opbuf, WordAt word_at) {
  const int ma = a.sc.match, mm = a.sc.mismatch, open = a.sc.gap_open, ext = a.sc.gap_ext;
  static_assert(USE_FLAG || WB % 16 == 0 || WB == 8, "sequence words cover 16 bases");
  // SPEC section 4: smallest row among the candidates that reaches S in column je
  constexpr int SB = (WB > 8) ? 16 : 8;                   // batch of the candidate / gap scans
  int ie = 0;
  for (int i0 = g.i_lo; i0 <= g.i_hi && !ie; i0 += SB) {
    int v[SB];
#pragma unroll
    for (int u = 0; u < SB; ++u) {                          // loads are unconditional (clamped) so that they all go out together
      const int ii = min(i0 + u, g.i_hi);
      v[u] = (int)(word_at(ii - g.i_first, g.je - ii - g.d_lo) & 0x7FFFu);
      if (i0 + u > g.i_hi) v[u] = -1;
    }
#pragma unroll
    for (int u = SB - 1; u >= 0; --u) if (v[u] == g.S) ie = i0 + u;
  }
  if (ie == 0) { atomicOr(a.err_flag, 2u); return; }
  OpRuns o; o.init(opbuf);
  o.push_run(UBU_SW_CIGAR_S, (uint32_t)(t.L - ie));
  int i = ie, j = g.je, h = g.S, n_m = 0, gap_bases = 0, gap_cost = 0, grem = gap_bound(a.sc, ie, g.S);
  bool bad = false;
  while (h > 0) {
    const int n = i - g.i_first, b = j - i - g.d_lo;
    if (i < 1 || j < 1 || n < 0 || b < 0 || b >= bw) { bad = true; break; }
    // this cell and the next WB up its diagonal, fetched together: the walk is latency-bound and most paths are
    // long diagonal runs, so one round trip to the H store serves up to WB steps
    uint32_t w[WB + 1];
#pragma unroll
    for (int k = 0; k <= WB; ++k) { w[k] = word_at(max(n - k, 0), b); if (n - k < 0) w[k] = 0u; }   // unconditional, clamped loads
    // !USE_FLAG: word c covers steps k = 16c .. 16c+15: base (i-1-k) vs (j-1-k) differ / involve N, at bit 2*(15-(k%16)) / (15-(k%16))
    constexpr int NXW = (WB + 15) / 16;
    uint32_t xdiff[NXW], xn[NXW];
#pragma unroll
    for (int c = 0; c < NXW; ++c) {
      xdiff[c] = 0u; xn[c] = 0u;
      if (!USE_FLAG) {
        xdiff[c] = load16(t.q, i - 16 - 16 * c) ^ load16(t.r, j - 16 - 16 * c);
        xdiff[c] = (xdiff[c] | (xdiff[c] >> 1)) & 0x55555555u;
        if (HAS_N) { if (t.qn) xn[c] |= load16n(t.qn, i - 16 - 16 * c); if (t.rn) xn[c] |= load16n(t.rn, j - 16 - 16 * c); }
      }
    }
    bool stop = false;
#pragma unroll
    for (int k = 0; k < WB; ++k) {
      if (!stop) {
        bool diag;
        if (USE_FLAG) diag = !(w[k] & 0x8000u);
        else {
          const bool isn = HAS_N && ((xn[k / 16] >> (15 - (k % 16))) & 1u);
          const int sk = isn ? 0 : (((xdiff[k / 16] >> (2 * (15 - (k % 16)))) & 1u) ? mm : ma);
          diag = (i >= 1 && j >= 1) && ((int)(w[k + 1] & 0x7FFFu) + sk == h);
        }
        if (h > 0 && diag) { o.push(UBU_SW_CIGAR_M, 1u); ++n_m; --i; --j; h = (int)(w[k + 1] & 0x7FFFu); }
        else stop = true;
      }
    }
    if (!stop || h <= 0) continue;
    // (i, j) is not entered diagonally: a gap run. Runs are bounded by the band, by what is left of the SPEC
    // section 7 gap budget, and by need <= S; candidates are fetched in batches of independent loads.
    const int n2 = i - g.i_first, b2 = j - i - g.d_lo;
    if (i < 1 || j < 1 || n2 < 0 || b2 < 0 || b2 >= bw) { bad = true; break; }
    const int kcap = min(grem, (g.S - h - open) / ext);
    // read gap (I): largest k with H[i-k][j] - open - k*ext == h;  window gap (D): same along the row; I has priority.
    // The first SB candidates of BOTH directions are fetched in one round trip (the scans are latency-bound).
    const int kmaxI = min(kcap, min(n2, bw - 1 - b2)), kmaxD = min(kcap, min(b2, j - 1));
    int kI = 0, kD = 0;
    {
      int hi_[SB], hd_[SB];
#pragma unroll
      for (int u = 0; u < SB; ++u) {
        const int ui = min(1 + u, max(kmaxI, 0)), ud = min(1 + u, max(kmaxD, 0));     // clamped: every load is issued
        hi_[u] = (int)(word_at(n2 - ui, b2 + ui) & 0x7FFFu); if (1 + u > kmaxI) hi_[u] = -1;
        hd_[u] = (int)(word_at(n2, b2 - ud) & 0x7FFFu); if (1 + u > kmaxD) hd_[u] = -1;
      }
#pragma unroll
      for (int u = 0; u < SB; ++u) {
        if (hi_[u] == h + open + (1 + u) * ext) kI = 1 + u;
        if (hd_[u] == h + open + (1 + u) * ext) kD = 1 + u;
      }
    }
    for (int k0 = 1 + SB; k0 <= kmaxI; k0 += SB) {
      int hv[SB];
#pragma unroll
      for (int u = 0; u < SB; ++u) { const int kk = min(k0 + u, kmaxI); hv[u] = (int)(word_at(n2 - kk, b2 + kk) & 0x7FFFu); if (k0 + u > kmaxI) hv[u] = -1; }
#pragma unroll
      for (int u = 0; u < SB; ++u) if (hv[u] == h + open + (k0 + u) * ext) kI = k0 + u;
    }
    if (kI) {
      o.push(UBU_SW_CIGAR_I, (uint32_t)kI); gap_bases += kI; gap_cost += open + kI * ext; grem -= kI;
      i -= kI; h += open + kI * ext;
      continue;
    }
    for (int k0 = 1 + SB; k0 <= kmaxD; k0 += SB) {
      int hv[SB];
#pragma unroll
      for (int u = 0; u < SB; ++u) { const int kk = min(k0 + u, kmaxD); hv[u] = (int)(word_at(n2, b2 - kk) & 0x7FFFu); if (k0 + u > kmaxD) hv[u] = -1; }
#pragma unroll
      for (int u = 0; u < SB; ++u) if (hv[u] == h + open + (k0 + u) * ext) kD = k0 + u;
    }
    if (kD) {
      o.push(UBU_SW_CIGAR_D, (uint32_t)kD); gap_bases += kD; gap_cost += open + kD * ext; grem -= kD;
      j -= kD; h += open + kD * ext;
      continue;
    }
    bad = true; break;
  }
  o.flush();
  if (bad) { atomicOr(a.err_flag, 2u); return; }
  // edit distance = mismatching M columns + gap bases (SPEC section 6)
  int mism;
  if (!HAS_N || (!t.qn && !t.rn)) {
    // without N every M column scores match or mismatch: S = ma*(n_m - x) + mm*x - gap_cost
    const int num = ma * n_m - gap_cost - g.S;
    mism = num / (ma - mm);
    if (num < 0 || mism * (ma - mm) != num) { atomicOr(a.err_flag, 2u); return; }
  } else {
    mism = 0;                                               // replay the runs front to back against the sequences
    int qi = i, rj = j;                                     // 0-based positions of the first aligned bases
    for (int k = o.n - 1; k >= 0; --k) {
      const int len = (int)(o.ops[k] >> 4); const uint32_t op = o.ops[k] & 15u;
      if (op == UBU_SW_CIGAR_M) {                           // 16 columns per XOR + popcount
        for (int u = 0; u < len; u += 16) mism += mismatches16(t.q, t.r, t.qn, t.rn, qi + u, rj + u, min(16, len - u));
        qi += len; rj += len;
      } else if (op == UBU_SW_CIGAR_I) qi += len;
      else if (op == UBU_SW_CIGAR_D) rj += len;
    }
  }
  o.push_run(UBU_SW_CIGAR_S, (uint32_t)i);     // q_start - 1 = i
  emit_result(a, task, o, g.S, j + 1, g.je, i + 1, ie, mism + gap_bases);
}

// ---- class 0: no gap fits under the score (Gb == 0) ---------------------------------------------
// Every path of score S ending in a candidate cell is gap-free, so H[i][je] == S iff the best purely
// diagonal local score ending there is S, and the SPEC walk is the diagonal itself: accumulate scores
// backwards from (i, je); the first time the sum reaches S the running H hits 0 (SPEC stop).
template <bool HAS_N>
__global__ void __launch_bounds__(128)
tb_diag_kernel(TbArgs a, const uint32_t* __restrict__ list, const unsigned int* __restrict__ first, const unsigned int* __restrict__ count) {
  const uint32_t nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = *count - *first;                  // this launch serves list[*first, *count) (see do_run: traceback phases)
  list += *first;
  const int ma = a.sc.match, mm = a.sc.mismatch;
  for (uint32_t idx = gtid; idx < n; idx += nthreads) {
    const uint32_t task = list[idx];
    const FwdOut f = a.fwd[task];
    const TaskView t = load_task(a, task);
    const int S = f.score, je = f.je;
    int ie = 0, len = 0, edits = 0;
    // Phase 1 (uniform across the warp): drop the candidate diagonals that cannot reach S, judged on their last 16
    // columns only (popcount of differing bases; N counted as a match, so the bound stays an upper bound). Without it
    // every lane of a warp pays for every other lane's full-length scan of ITS true diagonal.
    uint32_t alive = 0u;
    for (int i = f.i_lo; i <= f.i_hi; ++i) {
      const int maxk = min(i, je);
      bool keep = true;
      if (maxk > 16 && S > 16 * ma) {
        uint32_t x = load16(t.q, i - 16) ^ load16(t.r, je - 16);
        x = (x | (x >> 1)) & 0x55555555u;
        if (HAS_N) {
          uint32_t nm = 0u;
          if (t.qn) nm |= load16n(t.qn, i - 16);
          if (t.rn) nm |= load16n(t.rn, je - 16);
          x &= ~spread16(nm);
        }
        const int d = __popc(x);
        keep = (16 - d) * ma + d * mm + (maxk - 16) * ma >= S;
      }
      alive |= (keep ? 1u : 0u) << (i - f.i_lo);
    }
    // Phase 2: exact scan of the survivors in row order (usually exactly one).
    for (int i = f.i_lo; i <= f.i_hi && !ie; ++i) {
      if (!((alive >> (i - f.i_lo)) & 1u)) continue;
      const int maxk = min(i, je);
      int run = 0, mism = 0;
      bool done = false;
      for (int k0 = 0; k0 < maxk && !done; k0 += 16) {
        // 16 bases of the read ending at row i-k0 and of the window ending at column je-k0; base k0+15-u sits at bits 2u
        const uint32_t x = load16(t.q, i - 16 - k0) ^ load16(t.r, je - 16 - k0);
        uint32_t nm = 0u;
        if (HAS_N) {
          if (t.qn) nm |= load16n(t.qn, i - 16 - k0);
          if (t.rn) nm |= load16n(t.rn, je - 16 - k0);
        }
        const int rem = min(16, maxk - k0);
        if (run + rem * ma < S) {                               // S is out of reach inside this word: take it in one step
          const uint32_t fields = rem == 16 ? 0x55555555u : (0x55555555u << (2 * (16 - rem)));
          uint32_t dbits = (x | (x >> 1)) & fields, nbits = 0u;
          if (HAS_N) { nbits = spread16(nm) & fields; dbits &= ~nbits; }
          const int d = __popc(dbits), nn = __popc(nbits);
          run += (rem - d - nn) * ma + d * mm; mism += d;
          if (run + (maxk - k0 - rem) * ma < S) done = true;
          continue;
        }
#pragma unroll
        for (int u = 15; u >= 0; --u) {
          const int k = k0 + 15 - u;
          if (!done && k < maxk) {
            const bool diff = ((x >> (2 * u)) & 3u) != 0u;
            const bool isn = HAS_N && ((nm >> u) & 1u);
            run += isn ? 0 : (diff ? mm : ma);
            mism += (!isn && diff) ? 1 : 0;
            if (run == S) { ie = i; len = k + 1; edits = mism; done = true; }
            else if (run + (maxk - 1 - k) * ma < S) done = true;           // S is out of reach on this diagonal
          }
        }
      }
    }
    if (!ie) { atomicOr(a.err_flag, 2u); continue; }
    uint32_t ops[3]; int no = 0;                              // front to back
    if (ie - len > 0) ops[no++] = ((uint32_t)(ie - len) << 4) | UBU_SW_CIGAR_S;
    ops[no++] = ((uint32_t)len << 4) | UBU_SW_CIGAR_M;
    if (t.L - ie > 0) ops[no++] = ((uint32_t)(t.L - ie) << 4) | UBU_SW_CIGAR_S;
    ubu_sw_result r;
    r.score = S; r.ref_start = je - len + 1; r.ref_end = je; r.q_start = ie - len + 1; r.q_end = ie; r.edit_distance = edits;
    r.flags = 0; r.cigar_n = (uint16_t)no;
    // CIGAR pool space for the lanes that got here together: one cursor update per warp
    const cooperative_groups::coalesced_group cg_ = cooperative_groups::coalesced_threads();
    const unsigned pre = cooperative_groups::exclusive_scan(cg_, (unsigned)no);
    unsigned base = 0u;
    if (cg_.thread_rank() == cg_.size() - 1) base = atomicAdd(a.cig_cursor, pre + (unsigned)no);
    const unsigned off = cg_.shfl(base, cg_.size() - 1) + pre;
    r.cigar_off = off + a.cig_base;
    if (off + (unsigned)no <= a.cig_cap) { for (int k = 0; k < no; ++k) a.cig_pool[off + k] = ops[k]; }
    else atomicOr(a.err_flag, 1u);
    a.res[task] = r;
  }
}

// ---- pieces shared by the packed fill kernels ---------------------------------------------------
// Byte profile of read base `code` (same encoding as the forward pass).
__device__ __forceinline__ uint32_t profile_of(uint32_t code, bool isn, bool valid, uint32_t MM4, uint32_t MX, uint32_t N4) {
  return !valid ? 0x80808080u : (isn ? N4 : (MM4 ^ (MX << (8u * code))));
}

// Streams the read bases of band rows n = 0, 1, ... (row n is read row i_first + n), 16 rows per 32-bit load.
struct RowCodes {
  const TaskView* t; int i_first; uint32_t w, nw;
  __device__ __forceinline__ void init(const TaskView* tv, int ifirst) { t = tv; i_first = ifirst; w = 0u; nw = 0u; }
  template <bool HAS_N>
  __device__ __forceinline__ uint32_t profile(int n, uint32_t MM4, uint32_t MX, uint32_t N4) {
    const int i0 = i_first - 1 + n;                        // 0-based read position of row n
    if ((n & 15) == 0) {
      const bool any = i0 < t->L;
      w = any ? load16(t->q, i0) : 0u;
      nw = (HAS_N && t->qn && any) ? load16n(t->qn, i0) : 0u;
    }
    const int u = n & 15;
    return profile_of((w >> (2 * u)) & 3u, (nw >> u) & 1u, i0 < t->L, MM4, MX, N4);
  }
};

// Selectors of band columns k = 0 .. ncols-1 (0-based window columns c0A + k / c0B + k) for the pair, 16 per pair of
// 32-bit loads; columns outside [0, W) get the padding selector (score <= 0 after bias).
template <bool HAS_N, class Store>
__device__ __forceinline__ void build_selectors(const TaskView& ta, const TaskView& tb, int c0A, int c0B, int ncols, Store store) {
  for (int k0 = 0; k0 < ncols; k0 += 16) {
    const int ca = c0A + k0, cb = c0B + k0;
    const bool anyA = ca < ta.W && ca + 16 > 0, anyB = cb < tb.W && cb + 16 > 0;
    const uint32_t wa = anyA ? load16(ta.r, ca) : 0u, wb = anyB ? load16(tb.r, cb) : 0u;
    uint32_t na = 0u, nb = 0u;
    if (HAS_N) { if (ta.rn && anyA) na = load16n(ta.rn, ca); if (tb.rn && anyB) nb = load16n(tb.rn, cb); }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (k0 + u < ncols) {
        uint32_t lo = 0x88u, hi = 0xCCu, nf = 0u;
        if (ca + u >= 0 && ca + u < ta.W) { const uint32_t c = (wa >> (2 * u)) & 3u; lo = c | ((c | 8u) << 4); if (HAS_N && ((na >> u) & 1u)) nf |= 0x00800000u; }
        if (cb + u >= 0 && cb + u < tb.W) { const uint32_t c = ((wb >> (2 * u)) & 3u) + 4u; hi = c | ((c | 8u) << 4); if (HAS_N && ((nb >> u) & 1u)) nf |= 0x80000000u; }
        store(k0 + u, lo | (hi << 8) | nf);
      }
    }
  }
}

// One packed band cell. Returns the value to store: H | (diagonal does not reproduce H) << 15, per half.
template <bool HAS_N, bool WITH_FLAG = true>
__device__ __forceinline__ uint32_t band_cell(uint32_t pa, uint32_t pb, uint32_t selv, uint32_t diag_hoe, uint32_t up_hoe, uint32_t up_f,
                                              uint32_t& e, uint32_t& hoe_left, uint32_t& f_out, uint32_t NEG_OE, uint32_t NEG_EXT,
                                              uint32_t POS_OE) {
  uint32_t sp = prmt(pa, pb, selv & 0xFFFFu);
  if (HAS_N) { const uint32_t isn = prmt(selv, 0u, 0xBBAAu); sp = (sp & ~isn) | (POS_OE & isn); }
  const uint32_t t = __vadd2(diag_hoe, sp);                                 // H[i-1][j-1] + s
  const uint32_t f = __viaddmax_s16x2(up_f, NEG_EXT, up_hoe);               // F[i][j]
  e = __viaddmax_s16x2(e, NEG_EXT, hoe_left);                               // E[i][j]
  const uint32_t h = __vimax3_s16x2_relu(t, e, f);
  hoe_left = __vadd2(h, NEG_OE);
  f_out = f;
  if (!WITH_FLAG) return h;                                                 // byte store: H only, the walk tests the diagonal from the sequences
  const uint32_t notdiag = __vmins2(__vsub2(h, t), 0x00010001u);           // h >= t always: 0 iff the diagonal reproduces H
  return h + notdiag * 0x8000u;
}

// ---- register band fill + walk: one thread = one PAIR of tasks ---------------------------------
// hstore: one contiguous block per warp, packed value of band cell (n, b) of lane l at block[(n*BW + b) * 32 + l]: a warp's
//         store is one 128-byte line, consecutive cells are adjacent (streaming writes, TLB-friendly walk reads)
// selectors: shared memory, sel[k * THREADS + tid], k = n + b
// BYTES = 2: 16 bits per task and cell, H | not-diagonal flag << 15, the walk needs no sequence access.
// BYTES = 1: when no score of the batch can exceed 255 (H <= S <= match * L), H alone in one byte, two cells per word (bytes A_b, A_b+1,
//            B_b, B_b+1): half the store traffic -- these kernels are bound by HBM writes -- and three instructions less per cell; the walk
//            then tests the diagonal from the sequences, 16 bases per load, as the strip kernel's does.
template <int BW, bool HAS_N, int THREADS, int BYTES>
__global__ void __launch_bounds__(THREADS)
tb_fill_reg_kernel(TbArgs a, const uint32_t* __restrict__ list, const unsigned int* __restrict__ first, const unsigned int* __restrict__ count,
                   uint32_t* __restrict__ hstore, int max_rows) {
  using SelT = typename SelType<HAS_N>::type;
  extern __shared__ __align__(16) unsigned char tb_smem[];
  SelT* sel = reinterpret_cast<SelT*>(tb_smem);
  const uint32_t nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int tid = threadIdx.x;
  const uint32_t n_tasks = *count - *first, n_pairs = (n_tasks + 1u) / 2u;
  list += *first;
  const int oe = a.sc.gap_open + a.sc.gap_ext;
  const uint32_t mmb = (uint32_t)(a.sc.mismatch + oe) & 0xFFu, mab = (uint32_t)(a.sc.match + oe) & 0xFFu;
  const uint32_t MM4 = mmb * 0x01010101u, MX = mab ^ mmb, N4 = ((uint32_t)oe & 0xFFu) * 0x01010101u;
  const uint32_t NEG_OE = a.sc.neg_oe2, NEG_EXT = a.sc.neg_ext2, POS_OE = a.sc.pos_oe2;
  uint32_t opbuf[BW + 8];
  constexpr int WPR = BYTES == 1 ? BW / 2 : BW;                 // words per band row and lane
  uint32_t* hmine = hstore + (size_t)(gtid >> 5) * ((size_t)max_rows * WPR * 32) + (gtid & 31u);

  for (uint32_t pair = gtid; pair < n_pairs; pair += nthreads) {
    const uint32_t tA = list[2u * pair];
    const bool haveB = 2u * pair + 1u < n_tasks;
    const uint32_t tB = haveB ? list[2u * pair + 1u] : tA;
    const TaskView va = load_task(a, tA), vb = load_task(a, tB);
    const Band ga = make_band(a.sc, a.fwd[tA], BW), gb = make_band(a.sc, a.fwd[tB], BW);
    const int nrows = max(ga.rows, gb.rows);
    if (nrows > max_rows) { atomicOr(a.err_flag, 2u); continue; }

    build_selectors<HAS_N>(va, vb, ga.c0, gb.c0, nrows + BW - 1, [&](int k, uint32_t v) { sel[k * THREADS + tid] = (SelT)v; });

    uint32_t Hd[BW], Fv[BW];
#pragma unroll
    for (int b = 0; b < BW; ++b) { Hd[b] = NEG_OE; Fv[b] = NEG_OE; }
    RowCodes ra, rb; ra.init(&va, ga.i_first); rb.init(&vb, gb.i_first);
    for (int n = 0; n < nrows; ++n) {
      const uint32_t pa = ra.template profile<HAS_N>(n, MM4, MX, N4);
      const uint32_t pb = rb.template profile<HAS_N>(n, MM4, MX, N4);
      uint32_t e = NEG_OE, hoe_left = NEG_OE;
      const SelT* srow = sel + (size_t)n * THREADS + tid;
      uint32_t* hrow = hmine + (size_t)n * WPR * 32;
      uint32_t even = 0u;
#pragma unroll
      for (int b = 0; b < BW; ++b) {
        const uint32_t up_hoe = (b + 1 < BW) ? Hd[b + 1] : NEG_OE, up_f = (b + 1 < BW) ? Fv[b + 1] : NEG_OE;
        uint32_t f;
        const uint32_t v = band_cell<HAS_N, BYTES == 2>(pa, pb, srow[b * THREADS], Hd[b], up_hoe, up_f, e, hoe_left, f, NEG_OE, NEG_EXT, POS_OE);
        if (BYTES == 2) hrow[b * 32] = v;
        else if (b & 1) hrow[(b >> 1) * 32] = v * 256u + even;   // bytes (A_b-1, A_b, B_b-1, B_b): one multiply-add on the FMA pipe
        else even = v;
        Hd[b] = hoe_left; Fv[b] = f;
      }
    }
    if (BYTES == 2) {
      walk_h_and_emit<HAS_N, true, 8>(a, tA, va, ga, BW, opbuf, [&](int n, int b) -> uint32_t { return hmine[(n * BW + b) * 32] & 0xFFFFu; });
      if (haveB)
        walk_h_and_emit<HAS_N, true, 8>(a, tB, vb, gb, BW, opbuf, [&](int n, int b) -> uint32_t { return hmine[(n * BW + b) * 32] >> 16; });
    } else {
      walk_h_and_emit<HAS_N, false, 8>(a, tA, va, ga, BW, opbuf, [&](int n, int b) -> uint32_t { return (hmine[(n * WPR + (b >> 1)) * 32] >> (8 * (b & 1))) & 0xFFu; });
      if (haveB)
        walk_h_and_emit<HAS_N, false, 8>(a, tB, vb, gb, BW, opbuf, [&](int n, int b) -> uint32_t { return (hmine[(n * WPR + (b >> 1)) * 32] >> (16 + 8 * (b & 1))) & 0xFFu; });
    }
  }
}

// ---- last resort: scalar, band rows + direction bits in global scratch (any width) ----------------
// Replays the SPEC state machine on stored direction nibbles:
//   bits 0-1 = H source (0 stop, 1 diagonal, 2 F, 3 E), bit 2 = F came by extension, bit 3 = E came by extension
// Per-thread scratch, contiguous: [Hs: bw_cap ints][Fs: bw_cap ints][ops: max_rows + 2*bw_cap + 8][dir: max_rows * nw_cap u64]
__host__ __device__ inline size_t tb_wide_scratch_bytes(int max_rows, int bw_cap) {
  const size_t nw = (size_t)(bw_cap + 15) / 16;
  size_t b = (size_t)bw_cap * 8 + ((size_t)max_rows + 2 * (size_t)bw_cap + 8) * 4;
  b = (b + 7) & ~(size_t)7;
  return b + (size_t)max_rows * nw * 8;
}

__global__ void __launch_bounds__(64)
tb_wide_kernel(TbArgs a, const uint32_t* __restrict__ list, const unsigned int* __restrict__ first, const unsigned int* __restrict__ count,
               unsigned char* __restrict__ scratch, int max_rows, int bw_cap) {
  const uint32_t nthreads = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = *count - *first;
  list += *first;
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
    OpRuns o; o.init(opbuf);
    o.push_run(UBU_SW_CIGAR_S, (uint32_t)(t.L - ie));
    int i = ie, j = je, state = 0, edits = 0;
    bool bad = false;
    while (true) {
      if (i < 1 || j < 1) { if (state != 0) bad = true; break; }
      const int b = j - i - d_lo;
      if (b < 0 || b >= bw || i < i_first) { bad = true; break; }
      const uint32_t nib = (uint32_t)(dir[(size_t)(i - i_first) * nw + b / 16] >> (4 * (b % 16))) & 15u;
      if (state == 0) {
        const uint32_t src = nib & 3u;
        if (src == 0u) break;
        if (src == 1u) {
          o.push(UBU_SW_CIGAR_M, 1u);
          const uint32_t qc = base2(t.q, (uint32_t)(i - 1)), rc = base2(t.r, (uint32_t)(j - 1));
          const bool isn = (t.qn && nbit(t.qn, (uint32_t)(i - 1))) || (t.rn && nbit(t.rn, (uint32_t)(j - 1)));
          if (qc != rc && !isn) ++edits;
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
    if (bad) { atomicOr(a.err_flag, 2u); continue; }
    emit_result(a, task, o, S, j + 1, je, i + 1, ie, edits);
  }
}

}  // namespace ubu
