// ubu_pack.cu -- host-only part of libubu_sw.so: batch 2-bit packing of read / window text on the host's cores.
// (No device code here; it is a .cu only so that the one nvcc build line covers the whole library.)
//
// The reference packs one k-mer at a time, char by char (Sequence.java:15-43). A realignment batch is millions of reads, and
// at ~6 ms of GPU time per million reads the packer sits on the critical path of the pipeline, so the
// batch form works on 32 (AVX2, picked at run time) or 8 (64-bit words) characters per step and on all cores: with A=0x41 C=0x43 G=0x47 T=0x54 the reference's code
// A=0 C=1 T=2 G=3 (Sequence.java:10-13) is simply (c >> 1) & 3, for both cases.
#include "../../include/ubu_sw.h"

#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline uint64_t zero_bytes(uint64_t z) {                 // 0x80 in every byte of z that is 0
  return ~(((z & 0x7F7F7F7F7F7F7F7Full) + 0x7F7F7F7F7F7F7F7Full) | z) & 0x8080808080808080ull;
}

#if defined(__x86_64__)
// 32 characters per step with AVX2 (chosen at run time, see use_avx2): returns the characters consumed (a multiple of 32).
__attribute__((target("avx2"))) uint32_t pack_avx2(const char* t, uint32_t n, uint8_t* packed, uint8_t* nmask, bool& any) {
  const __m256i m3 = _mm256_set1_epi8(3), fold = _mm256_set1_epi8((char)0xDF);
  const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
  const __m256i w14 = _mm256_set1_epi16(0x0401), w116 = _mm256_set1_epi32(0x00100001);
  uint32_t k = 0;
  for (; k + 32u <= n; k += 32u) {
    const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(t + k));
    const __m256i y = _mm256_and_si256(x, fold);
    const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(y, cA), _mm256_cmpeq_epi8(y, cC)),
                                       _mm256_or_si256(_mm256_cmpeq_epi8(y, cG), _mm256_cmpeq_epi8(y, cT)));
    const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(ok);
    __m256i c = _mm256_and_si256(_mm256_and_si256(_mm256_srli_epi16(x, 1), m3), ok);      // codes, N -> 0
    c = _mm256_maddubs_epi16(c, w14);                       // c0 + 4 c1 per 16 bits
    c = _mm256_madd_epi16(c, w116);                         // + 16 (c2 + 4 c3) per 32 bits: one packed byte each
    c = _mm256_packus_epi32(c, c);                          // per 128-bit lane: 4 x 16 bits (twice)
    c = _mm256_packus_epi16(c, c);                          // per lane: 4 bytes in the low dword
    const uint32_t lo = (uint32_t)_mm256_extract_epi32(c, 0), hi = (uint32_t)_mm256_extract_epi32(c, 4);
    memcpy(packed + k / 4, &lo, 4); memcpy(packed + k / 4 + 4, &hi, 4);
    if (bad) any = true;
    if (nmask) memcpy(nmask + k / 8, &bad, 4);
  }
  return k;
}
bool use_avx2() { static const bool v = __builtin_cpu_supports("avx2"); return v; }
#endif

// Packs one sequence; returns true when it holds a non-ACGT character. nmask may be nullptr.
bool pack_one(const char* t, uint32_t n, uint8_t* packed, uint8_t* nmask) {
  bool any = false;
  uint32_t k = 0;
  const uint32_t pbytes = (n + 3u) / 4u, nbytes = (n + 7u) / 8u;
#if defined(__x86_64__)
  if (use_avx2()) k = pack_avx2(t, n, packed, nmask, any);
#endif
  for (; k + 8u <= n; k += 8u) {
    uint64_t x; memcpy(&x, t + k, 8);
    uint64_t c = (x >> 1) & 0x0303030303030303ull;                        // 2-bit codes, one per byte
    c = (c | (c >> 6)) & 0x000F000F000F000Full;                           // two codes per 16 bits
    c = (c | (c >> 12)) & 0x000000FF000000FFull;                          // four codes per 32 bits
    const uint32_t w = (uint32_t)((c | (c >> 24)) & 0xFFFFull);           // eight codes
    const uint64_t y = x & 0xDFDFDFDFDFDFDFDFull;                          // case-folded
    const uint64_t ok = zero_bytes(y ^ 0x4141414141414141ull) | zero_bytes(y ^ 0x4343434343434343ull) |
                        zero_bytes(y ^ 0x4747474747474747ull) | zero_bytes(y ^ 0x5454545454545454ull);
    uint32_t bad = 0;
    if (ok != 0x8080808080808080ull) {
      const uint64_t inv = ~ok & 0x8080808080808080ull;
      bad = (uint32_t)(((inv >> 7) * 0x0102040810204080ull) >> 56);       // one bit per character, LSB first
      any = true;
    }
    uint32_t ww = w;
    if (bad) for (int j = 0; j < 8; ++j) if ((bad >> j) & 1u) ww &= ~(3u << (2 * j));      // N is packed as code 0
    packed[k / 4] = (uint8_t)ww; packed[k / 4 + 1] = (uint8_t)(ww >> 8);
    if (nmask) nmask[k / 8] = (uint8_t)bad;
  }
  if (k < n) {                                                            // tail of fewer than 8 characters
    uint32_t ww = 0, bad = 0;
    for (uint32_t j = 0; k + j < n; ++j) {
      const unsigned char ch = (unsigned char)t[k + j], f = ch & 0xDFu;
      if (f == 'A' || f == 'C' || f == 'G' || f == 'T') ww |= ((uint32_t)(ch >> 1) & 3u) << (2 * j);
      else { bad |= 1u << j; any = true; }
    }
    packed[k / 4] = (uint8_t)ww;
    if (k / 4 + 1 < pbytes) packed[k / 4 + 1] = (uint8_t)(ww >> 8);
    if (nmask && k / 8 < nbytes) nmask[k / 8] = (uint8_t)bad;
  }
  return any;
}

}  // namespace

extern "C" int ubu_sw_pack_batch(const char* text, const uint64_t* text_off, const uint16_t* len, uint32_t n, uint8_t* packed,
                                 const uint32_t* off, uint8_t* nmask, const uint32_t* nmask_off, int threads, uint32_t* n_with_n) {
  if (n_with_n) *n_with_n = 0;
  if (n == 0) return UBU_SW_OK;
  if (!text || !text_off || !len || !packed || !off || (nmask && !nmask_off)) return UBU_SW_ERR_ARG;
  unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
  nt = std::min<unsigned>(nt, std::max<uint32_t>(1u, n / 4096u));
  std::atomic<uint32_t> with_n{0};
  auto work = [&](uint32_t lo, uint32_t hi) {
    uint32_t cnt = 0;
    for (uint32_t k = lo; k < hi; ++k)
      if (pack_one(text + text_off[k], len[k], packed + off[k], nmask ? nmask + nmask_off[k] : nullptr)) ++cnt;
    with_n += cnt;
  };
  if (nt <= 1) work(0, n);
  else {
    std::vector<std::thread> pool;
    try {
      for (unsigned t = 0; t < nt; ++t) pool.emplace_back(work, (uint32_t)((uint64_t)n * t / nt), (uint32_t)((uint64_t)n * (t + 1) / nt));
    } catch (...) { for (auto& th : pool) th.join(); return UBU_SW_ERR_NOMEM; }
    for (auto& th : pool) th.join();
  }
  if (n_with_n) *n_with_n = with_n.load();
  return UBU_SW_OK;
}

// ---- SAM text for a batch of results (the ordered host merge's last step) --------------------------------
// One line per read, input order: QNAME FLAG RNAME POS MAPQ CIGAR * 0 0 SEQ * NM:i AS:i -- the fields the reference's
// consumers read back from the aligner's SAM (ReAligner.java:1088-1143). Formatting runs on all cores in blocks of reads;
// blocks are written to the descriptor in order.
namespace {

inline char* put_u(char* p, uint64_t v) {
  char tmp[24]; int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}
inline char* put_s(char* p, const char* s) { while (*s) *p++ = *s++; return p; }

}  // namespace

size_t ubu_sam_block(const ubu_sw_seqs* reads, const ubu_sw_result* res, const uint32_t* pool, uint32_t lo, uint32_t hi, const ubu_sw_sam_names* nm,
                     std::vector<char>& out) {
  static const char bases[4] = {'A', 'C', 'T', 'G'};      // Sequence.java:10-13
  static const char opc[] = "MIDNSHP=X";
  size_t need = 0;
  const size_t prefix = nm->qname_prefix ? strlen(nm->qname_prefix) : 1;
  for (uint32_t k = lo; k < hi; ++k) {                    // worst case per line: fixed fields and five 20-digit numbers, the names, SEQ, 11 characters per op
    const uint32_t w = nm->pair_ref_idx ? nm->pair_ref_idx[k] : k;
    need += 192 + prefix + (nm->rnames ? strlen(nm->rnames[w]) : 11) + (size_t)reads->len[k] + 11 * (size_t)res[k].cigar_n;
  }
  out.resize(need);
  char* p = out.data();
  for (uint32_t k = lo; k < hi; ++k) {
    const ubu_sw_result& r = res[k];
    const bool unal = (r.flags & UBU_SW_FLAG_UNALIGNED) != 0;
    const uint32_t w = nm->pair_ref_idx ? nm->pair_ref_idx[k] : k;
    p = put_s(p, nm->qname_prefix ? nm->qname_prefix : "r"); p = put_u(p, nm->first_index + k); *p++ = '\t';
    p = put_u(p, unal ? 4u : 0u); *p++ = '\t';
    if (unal) { *p++ = '*'; *p++ = '\t'; *p++ = '0'; *p++ = '\t'; *p++ = '0'; *p++ = '\t'; *p++ = '*'; }
    else {
      if (nm->rnames) p = put_s(p, nm->rnames[w]); else { *p++ = 'w'; p = put_u(p, w); }
      *p++ = '\t';
      p = put_u(p, (uint64_t)((nm->window_pos0 ? nm->window_pos0[w] : 0) + (int64_t)r.ref_start)); *p++ = '\t';
      p = put_u(p, 37u); *p++ = '\t';
      if (r.cigar_n && (!pool || (nm->cigar_pool_n && (uint64_t)r.cigar_off + r.cigar_n > nm->cigar_pool_n))) { out.clear(); return (size_t)-1; }   // slice outside the pool
      for (uint32_t c = 0; c < r.cigar_n; ++c) { const uint32_t op = pool[r.cigar_off + c]; p = put_u(p, op >> 4); *p++ = opc[(op & 15u) < 9 ? (op & 15u) : 0]; }
    }
    *p++ = '\t'; *p++ = '*'; *p++ = '\t'; *p++ = '0'; *p++ = '\t'; *p++ = '0'; *p++ = '\t';
    const uint8_t* q = reads->packed + reads->off[k];
    const uint8_t* qn = reads->nmask ? reads->nmask + reads->nmask_off[k] : nullptr;
    const uint32_t L = reads->len[k];
    for (uint32_t i = 0; i < L; ++i) *p++ = (qn && ((qn[i >> 3] >> (i & 7)) & 1)) ? 'N' : bases[(q[i >> 2] >> ((i & 3) * 2)) & 3];
    if (L == 0) *p++ = '*';
    *p++ = '\t'; *p++ = '*';
    if (!unal) { p = put_s(p, "\tNM:i:"); p = put_u(p, (uint64_t)r.edit_distance); p = put_s(p, "\tAS:i:"); p = put_u(p, (uint64_t)r.score); }
    *p++ = '\n';
  }
  return (size_t)(p - out.data());
}

#include <unistd.h>

extern "C" int ubu_sw_sam_write(int fd, const ubu_sw_seqs* reads, const ubu_sw_result* res, const uint32_t* cigar_pool, uint32_t n,
                                const ubu_sw_sam_names* names, int threads, uint64_t* bytes_written) {
  if (bytes_written) *bytes_written = 0;
  if (n == 0) return UBU_SW_OK;
  if (!reads || !res || !names || reads->count < n || (n && (!reads->off || !reads->len))) return UBU_SW_ERR_ARG;
  unsigned nt = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
  const uint32_t block = 16384;
  const uint32_t n_blocks = (n + block - 1) / block;
  nt = std::min<unsigned>(nt, n_blocks);
  uint64_t total = 0;
  // rounds of nt blocks: formatted in parallel, written in order
  std::vector<std::vector<char>> bufs(nt);
  std::vector<size_t> used(nt, 0);
  for (uint32_t b0 = 0; b0 < n_blocks; b0 += nt) {
    const unsigned m = std::min<unsigned>(nt, n_blocks - b0);
    std::vector<std::thread> pool;
    try {
      for (unsigned t = 0; t < m; ++t)
        pool.emplace_back([&, t] { const uint32_t lo = (b0 + t) * block, hi = std::min<uint32_t>(n, lo + block); used[t] = ubu_sam_block(reads, res, cigar_pool, lo, hi, names, bufs[t]); });
    } catch (...) { for (auto& th : pool) th.join(); return UBU_SW_ERR_NOMEM; }
    for (auto& th : pool) th.join();
    for (unsigned t = 0; t < m; ++t) if (used[t] == (size_t)-1) return UBU_SW_ERR_ARG;          // a CIGAR slice outside the pool
    for (unsigned t = 0; t < m; ++t) {
      size_t off = 0;
      while (fd >= 0 && off < used[t]) { const ssize_t wr = write(fd, bufs[t].data() + off, used[t] - off); if (wr <= 0) return UBU_SW_ERR_ARG; off += (size_t)wr; }
      total += used[t];
    }
  }
  if (bytes_written) *bytes_written = total;
  return UBU_SW_OK;
}
