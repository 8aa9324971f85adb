// sw_post.cuh -- the two steps right behind the aligner in the reference's realign pipeline,
// one thread per read, bug-compatible with the Java they replace (oracle/post_oracle.py restates it line by line).
//
//   post_project_kernel : ReAligner.updateReadAlignment (ReAligner.java:1293-1368) + ReadBlock.getReadBlocks / getSubBlock /
//                         fillToLength / toCigarString (ReadBlock.java:74-175): an ungapped placement of a read at offset
//                         `position` of a contig is pushed through the contig's own genome CIGAR -> genome POS + CIGAR.
//   post_recount_kernel : CompareToReference.numDifferences (CompareToReference.java:83-123) + ReAligner.getNumIndelBases
//                         (ReAligner.java:316-326) + ReAligner.calcMappingQuality (:302-314): XM, NM, MAPQ of an adjusted read.
//                         Mismatches are counted 16 bases per XOR + popcount on the 2-bit codes (N plane masked out: "N never
//                         mismatches", CompareToReference.java:92). The reference's silence on CIGAR N (skip) elements -- they
//                         advance neither index (:99-108) -- is kept.
// Both are integer/byte work bounded by HBM (a few hundred bytes per read), not by issue rate.
#pragma once
#include "sw_common.cuh"

namespace ubu {

struct DevRefs {                 // long reference sequences (chromosomes): 64-bit offsets, 32-bit lengths
  const uint8_t* packed; const uint64_t* off; const uint32_t* len;
  const uint8_t* nmask; const uint64_t* nmask_off;
};

constexpr int POST_MAX_OPS = 64;

__device__ __forceinline__ bool post_update_positions(int& rb, int& fb, uint32_t op, int n) {   // ReadBlock.java:134-156
  switch (op) {
    case UBU_SW_CIGAR_S: rb += n; return true;
    case 3u /*N*/: case UBU_SW_CIGAR_D: fb += n; return true;
    case UBU_SW_CIGAR_I: rb += n; return true;
    case UBU_SW_CIGAR_M: rb += n; fb += n; return true;
    default: return false;                                                                       // IllegalStateException
  }
}

__global__ void post_project_kernel(const ubu_sw_contig* __restrict__ contigs, uint32_t n_contigs, const uint32_t* __restrict__ ccig,
                                    const ubu_sw_proj_in* __restrict__ in, uint32_t n, ubu_sw_proj_out* __restrict__ out,
                                    uint32_t* __restrict__ pool, unsigned int pool_cap, unsigned int* __restrict__ cursor) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const ubu_sw_proj_in pin = in[t];
  ubu_sw_proj_out po; po.pos = 0; po.cigar_off = 0; po.cigar_n = 0; po.flags = 0;
  if (pin.contig_idx >= n_contigs) { po.flags = UBU_SW_POST_MALFORMED; out[t] = po; return; }
  const ubu_sw_contig c = contigs[pin.contig_idx];
  const int L = (int)pin.read_len, position = pin.position;
  int olen[POST_MAX_OPS]; uint8_t otype[POST_MAX_OPS];
  int nb = 0, first_ref = 0, last_rs = 0, last_fs = 0;
  int contig_position = position, acc = 0;
  int rb = 1, fb = c.pos;                                        // ReadBlock.java:118-119
  bool bad = false;
  for (uint32_t k = 0; k < c.cigar_n && !bad; ++k) {
    const uint32_t e = ccig[c.cigar_off + k];
    const int blen = (int)(e >> 4); const uint32_t op = e & 15u;
    const int b_rs = rb, b_fs = fb;
    if (!post_update_positions(rb, fb, op, blen)) { bad = true; break; }
    const int ref_len = (op == UBU_SW_CIGAR_D) ? 0 : blen;       // ReadBlock.java:61-69
    if (b_rs + ref_len >= position + 1) {                        // ReAligner.java:1306-1307
      const int pib = contig_position + acc - b_rs + 1;          // ReadBlock.java:77
      int s_fs, s_len;
      if (op == 3u || op == UBU_SW_CIGAR_D) { s_fs = b_fs + pib; s_len = blen - pib; }
      else if (op == UBU_SW_CIGAR_S) { s_fs = b_fs; s_len = min(L - acc, blen - pib); }
      else { s_fs = b_fs + pib; s_len = min(L - acc, blen - pib); }
      const int s_rs = acc + 1;
      if (op == UBU_SW_CIGAR_I && s_len != 0) {                  // ReAligner.java:1314-1316
        contig_position -= blen - s_len;
        s_fs -= blen - s_len;
      }
      if (s_len != 0) {                                          // :1325
        if (nb >= POST_MAX_OPS - 1) { bad = true; break; }       // (a negative length is kept, as in the Java; see below)
        if (nb == 0) first_ref = s_fs;
        olen[nb] = s_len; otype[nb] = (uint8_t)op; ++nb;
        last_rs = s_rs; last_fs = s_fs;
        if (op != UBU_SW_CIGAR_D) acc += s_len;
        if (acc > L) { bad = true; break; }                      // :1334-1339
        if (acc == L) break;
      }
    }
  }
  if (bad) { po.flags = UBU_SW_POST_MALFORMED; out[t] = po; return; }
  if (nb == 0) { po.flags = UBU_SW_POST_NULL; out[t] = po; return; }     // :1362-1364
  // ReadBlock.fillToLength (ReadBlock.java:158-175); getTotalLength counts every non-D block (:101-111)
  int cl = 0;
  for (int k = 0; k < nb; ++k) if (otype[k] != UBU_SW_CIGAR_D) cl += olen[k];
  if (cl < L) {
    if (otype[nb - 1] == UBU_SW_CIGAR_M || otype[nb - 1] == UBU_SW_CIGAR_S) olen[nb - 1] += L - cl;
    else {
      int r2 = last_rs, f2 = last_fs;
      post_update_positions(r2, f2, otype[nb - 1], olen[nb - 1]);
      olen[nb] = L - cl; otype[nb] = UBU_SW_CIGAR_M; ++nb;
    }
  }
  // a block length that is still <= 0 here would make the reference's setCigarString throw
  for (int k = 0; k < nb; ++k) if (olen[k] <= 0) bad = true;
  if (bad) { po.flags = UBU_SW_POST_MALFORMED; out[t] = po; return; }
  const unsigned off = atomicAdd(cursor, (unsigned)nb);
  po.pos = first_ref; po.cigar_off = off; po.cigar_n = (uint16_t)nb;
  if (off + (unsigned)nb <= pool_cap) { for (int k = 0; k < nb; ++k) pool[off + k] = ((uint32_t)olen[k] << 4) | otype[k]; }
  else po.flags = UBU_SW_POST_POOL;
  out[t] = po;
}

__global__ void post_recount_kernel(DevSeqs reads, DevRefs refs, uint32_t n_refs, const ubu_sw_placed* __restrict__ placed,
                                    const uint32_t* __restrict__ cig, uint32_t n, ubu_sw_recount* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const ubu_sw_placed p = placed[t];
  ubu_sw_recount r; r.xm = 0; r.nm = 0; r.mapq = 0; r.flags = 0;
  if (p.ref_idx >= n_refs) { r.flags = UBU_SW_POST_MALFORMED; out[t] = r; return; }
  const uint8_t* q = reads.packed + reads.off[t];
  const uint8_t* qn = reads.nmask ? reads.nmask + reads.nmask_off[t] : nullptr;
  const int L = reads.len[t];
  const uint8_t* g = refs.packed + refs.off[p.ref_idx];
  const uint8_t* gn = refs.nmask ? refs.nmask + refs.nmask_off[p.ref_idx] : nullptr;
  const long long G = refs.len[p.ref_idx];
  int ri = 0; long long fi = (long long)p.pos - 1;               // CompareToReference.java:85-86
  int diffs = 0, indel = 0; bool bad = false;
  for (uint32_t k = 0; k < p.cigar_n && !bad; ++k) {
    const uint32_t e = cig[p.cigar_off + k];
    const int len = (int)(e >> 4); const uint32_t op = e & 15u;
    if (op == UBU_SW_CIGAR_M) {
      if (ri + len > L || fi < 0 || fi + len > G) { bad = true; break; }       // Java: StringIndexOutOfBoundsException
      for (int c0 = 0; c0 < len; c0 += 16) {
        const int m = min(16, len - c0);
        uint32_t x = load16(q, ri + c0) ^ load16(g, (int)(fi + c0));
        x = (x | (x >> 1)) & 0x55555555u;
        uint32_t nm = 0u;
        if (qn) nm |= load16n(qn, ri + c0);
        if (gn) nm |= load16n(gn, (int)(fi + c0));
        if (nm) { uint32_t spread = 0u; for (int u = 0; u < 16; ++u) spread |= ((nm >> u) & 1u) << (2 * u); x &= ~spread; }
        if (m < 16) x &= (1u << (2 * m)) - 1u;
        diffs += __popc(x);
      }
      ri += len; fi += len;
    } else if (op == UBU_SW_CIGAR_I) { ri += len; indel += len; }
    else if (op == UBU_SW_CIGAR_D) { fi += len; indel += len; }
    else if (op == UBU_SW_CIGAR_S) { ri += len; }
    // every other element (N, H, P, =, X) advances nothing: CompareToReference.java:99-108
  }
  if (bad) { r.flags = UBU_SW_POST_MALFORMED; out[t] = r; return; }
  r.xm = diffs; r.nm = diffs + indel;                                          // ReAligner.java:280-285
  r.mapq = p.mapq > 0 ? max(min((int)p.yq, 50) - (int)p.ym * 5, 1) : 0;          // ReAligner.java:302-314
  out[t] = r;
}

}  // namespace ubu
