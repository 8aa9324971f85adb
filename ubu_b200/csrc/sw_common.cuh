// sw_common.cuh -- shared device-side types for the SPEC.md Smith-Waterman path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ubu_sw.h"

namespace ubu {

// Device view of a ubu_sw_seqs (include/ubu_sw.h): 2-bit codes A=0 C=1 T=2 G=3, LSB first
// (reference convention, Sequence.java:10-13,31-41) + optional N bit-plane.
struct DevSeqs {
  const uint8_t* packed;
  const uint32_t* off;
  const uint16_t* len;
  const uint8_t* nmask;        // nullptr = no N anywhere in this set
  const uint32_t* nmask_off;
};

struct Scoring {
  int match, mismatch, gap_open, gap_ext;
  // packed (v:v) 16-bit constants precomputed on the host: as kernel parameters they reach the packed
  // instructions as constant-bank operands instead of costing a register-file read per use
  uint32_t neg_oe2, neg_ext2, pos_oe2;
};
inline Scoring make_scoring(int match, int mismatch, int gap_open, int gap_ext) {
  Scoring s; s.match = match; s.mismatch = mismatch; s.gap_open = gap_open; s.gap_ext = gap_ext;
  const int oe = gap_open + gap_ext;
  s.neg_oe2 = ((uint32_t)(-oe) & 0xFFFFu) * 0x00010001u;
  s.neg_ext2 = ((uint32_t)(-gap_ext) & 0xFFFFu) * 0x00010001u;
  s.pos_oe2 = ((uint32_t)oe & 0xFFFFu) * 0x00010001u;
  return s;
}

// Per-column PRMT selector as kept in shared memory: 16 bits, or 32 when the window N flags ride along.
template <bool HAS_N> struct SelType { using type = uint16_t; };
template <> struct SelType<true> { using type = uint32_t; };

// Output of the forward (score) pass, input of the traceback pass. 16 bytes.
// The end column is exact; the end row is known to lie in [i_lo, i_hi] (the K rows one
// lane owns); the traceback pass resolves it (SPEC.md section 4: smallest j, then smallest i).
struct FwdOut { int32_t score, je, i_lo, i_hi; };

__device__ __forceinline__ uint32_t base2(const uint8_t* __restrict__ p, uint32_t k) {
  return (uint32_t)(p[k >> 2] >> ((k & 3u) * 2u)) & 3u;
}
__device__ __forceinline__ uint32_t nbit(const uint8_t* __restrict__ m, uint32_t k) {
  return (uint32_t)(m[k >> 3] >> (k & 7u)) & 1u;
}
__device__ __forceinline__ uint32_t pack16x2(int v) {
  return ((uint32_t)v & 0xFFFFu) | ((uint32_t)v << 16);
}

// 16 consecutive bases (2 bits each, base `start` in bits 0-1) of a byte-aligned 2-bit sequence. `start` may be
// negative: bases before the sequence come back as zeros in the low positions (callers mask them by range).
// Reads at most 7 bytes past the last needed base: every packed buffer is uploaded with 16 bytes of slack.
__device__ __forceinline__ uint32_t load16(const uint8_t* __restrict__ seq, int start) {
  const int adj = start < 0 ? -start : 0;                    // bases to skip at the front
  const int s0 = start + adj;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(seq) + (uintptr_t)(s0 >> 2);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(addr & 3u) * 8u + (uint32_t)(s0 & 3) * 2u;
  const uint32_t v = __funnelshift_r(w[0], w[1], sh);
  return adj >= 16 ? 0u : (v << (2 * adj));
}
// load16 in two halves, for software prefetch: issue() puts the two 32-bit loads in flight, finish() (called a block of work
// later) shifts the 16 bases out. Same result as load16(seq, start).
struct Load16 {
  uint32_t w0, w1, sh; int adj;
  __device__ __forceinline__ void issue(const uint8_t* __restrict__ seq, int start, bool on) {
    adj = start < 0 ? -start : 0;
    const int s0 = start + adj;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(seq) + (uintptr_t)(s0 >> 2);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
    sh = (uint32_t)(addr & 3u) * 8u + (uint32_t)(s0 & 3) * 2u;
    w0 = 0u; w1 = 0u;
    if (on) { w0 = w[0]; w1 = w[1]; }
  }
  __device__ __forceinline__ uint32_t finish() const {
    const uint32_t v = __funnelshift_r(w0, w1, sh);
    return adj >= 16 ? 0u : (v << (2 * adj));
  }
};
// 16 consecutive N-plane bits (bit `start` in bit 0); same conventions.
__device__ __forceinline__ uint32_t load16n(const uint8_t* __restrict__ nm, int start) {
  const int adj = start < 0 ? -start : 0;
  const int s0 = start + adj;
  const uintptr_t addr = reinterpret_cast<uintptr_t>(nm) + (uintptr_t)(s0 >> 3);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(addr & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(addr & 3u) * 8u + (uint32_t)(s0 & 7);
  const uint32_t v = __funnelshift_r(w[0], w[1], sh) & 0xFFFFu;
  return adj >= 16 ? 0u : ((v << adj) & 0xFFFFu);
}

// bit u of a 16-bit mask -> bit 2u (the position of base u's low code bit in a load16 word)
__device__ __forceinline__ uint32_t spread16(uint32_t m) {
  m = (m | (m << 8)) & 0x00FF00FFu;
  m = (m | (m << 4)) & 0x0F0F0F0Fu;
  m = (m | (m << 2)) & 0x33333333u;
  return (m | (m << 1)) & 0x55555555u;
}
// mismatching, N-free columns among the `len` (<= 16) bases starting at read position qi / window position rj (0-based)
__device__ __forceinline__ int mismatches16(const uint8_t* __restrict__ q, const uint8_t* __restrict__ r, const uint8_t* __restrict__ qn,
                                            const uint8_t* __restrict__ rn, int qi, int rj, int len) {
  uint32_t x = load16(q, qi) ^ load16(r, rj);
  x = (x | (x >> 1)) & 0x55555555u;
  uint32_t nm = 0u;
  if (qn) nm |= load16n(qn, qi);
  if (rn) nm |= load16n(rn, rj);
  x &= ~spread16(nm);
  if (len < 16) x &= (1u << (2 * len)) - 1u;
  return __popc(x);
}

// prmt.b32 in its default mode: selector nibble bit 3 set = replicate the selected byte's sign bit.
// (Inline PTX rather than __byte_perm, whose documented contract only covers selector bits 2:0.)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
  uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s)); return d;
}

// Packed max with "old value kept" predicates: returns max(a,b) per half, keep_* = (a >= b).
// Same PTX idiom as the toolkit's __vibmax_s16x2 (which ptxas turns into one VIMNMX.S16x2 with two
// predicate outputs), but with an early-clobber output: the toolkit version lets the result alias
// input a, and `best = __vibmax_s16x2(best, ...)` then compares the result with itself.
__device__ __forceinline__ uint32_t vmax2_keep(uint32_t a, uint32_t b, bool& keep_hi, bool& keep_lo) {
  uint32_t val, ph, pl;
  asm("{.reg .pred pu, pv;\n\t"
      ".reg .s16 rs0, rs1, rs2, rs3;\n\t"
      "max.s16x2 %0, %3, %4;\n\t"
      "mov.b32 {rs0, rs1}, %0;\n\t"
      "mov.b32 {rs2, rs3}, %3;\n\t"
      "setp.eq.s16 pv, rs0, rs2;\n\t"
      "setp.eq.s16 pu, rs1, rs3;\n\t"
      "selp.b32 %1, 1, 0, pu;\n\t"
      "selp.b32 %2, 1, 0, pv;}"
      : "=&r"(val), "=r"(ph), "=r"(pl) : "r"(a), "r"(b));
  keep_hi = ph != 0u; keep_lo = pl != 0u;
  return val;
}

// Gap bound of SPEC.md section 7: max gap bases on a path of score S ending in row <= i.
__host__ __device__ __forceinline__ int gap_bound(const Scoring& sc, int i, int S) {
  const int num = sc.match * i - sc.gap_open - S;
  return num > 0 ? num / sc.gap_ext : 0;
}

// ---- traceback classes (sw_tb.cuh, sw_tbstrip.cuh). The forward kernels' epilogue sorts every task into one of them ----------
// 0: no gap possible (Gb == 0, pure diagonal scan), 1/2/3: band <= 8/16/32 diagonals (band row in registers),
// 4..8: wide bands by width (<= 48 / 64 / 88 / 120 / bw_cap diagonals): row strips in registers, one thread per pair
// (sw_tbstrip.cuh; the width classes only group similar tasks into the same warps), 9: scalar last resort
constexpr int TB_CLASSES = 10;
constexpr int TB_STRIP_FIRST = 4, TB_STRIP_CLASSES = 5, TB_CLASS_SCALAR = 9;
struct StripCaps { int rows_cap, bw_cap; };         // what the strip kernel's scratch of this batch was sized for (0 = strips off)
struct TbPlan {                        // where the epilogue puts its verdicts
  StripCaps caps;
  ubu_sw_result* res;                  // unaligned tasks (score <= 0) get their final record right away
  uint32_t* lists;                     // [TB_CLASSES][n_slots] task lists
  unsigned int* counts;                // [TB_CLASSES]
  uint32_t n_slots;
};
__device__ __forceinline__ int tb_class_of(const Scoring& sc, const FwdOut& f, const StripCaps& caps) {
  const int gb = gap_bound(sc, f.i_hi, f.score);
  const int bw = (f.i_hi - f.i_lo) + 2 * gb + 1;
  if (gb == 0) return 0;
  if (bw <= 8) return 1;
  if (bw <= 16) return 2;
  if (bw <= 32) return 3;
  if (bw > caps.bw_cap || f.i_hi > caps.rows_cap) return TB_CLASS_SCALAR;
  return bw <= 48 ? 4 : bw <= 64 ? 5 : bw <= 88 ? 6 : bw <= 120 ? 7 : 8;
}

}  // namespace ubu
