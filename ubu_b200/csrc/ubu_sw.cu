// ubu_sw.cu -- host side of libubu_sw.so: the C ABI of include/ubu_sw.h, the per-batch plan
// (length classes, pairing, window-length ordering), the stream pipeline and the kernel launches.
// Product path: there is NO CPU alignment code in this library and no fallback; every entry point
// that needs the GPU fails with UBU_SW_ERR_NO_DEVICE / UBU_SW_ERR_CUDA when it is not there.
#include "sw_fwd16.cuh"
#include "sw_fwd16s.cuh"
#include "sw_tb.cuh"
#include "sw_tbstrip.cuh"
#include "sw_long.cuh"
#include "sw_post.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace ubu;

namespace {

// ---- length classes of the packed 16-bit forward kernel: rows capacity = G*K -------------------
struct FwdCfg { int lmax, G, K; };
#ifndef UBU_FWD_THREADS
#define UBU_FWD_THREADS 64
#endif
constexpr int FWD_THREADS = UBU_FWD_THREADS;
// (finer steps above 80 rows since round 2: a mixed-length batch pays for the rows a class pads, ~10 % with the eleven classes of round 1)
#define UBU_FWD_CFGS(X) X(0, 32, 4, 8) X(1, 48, 4, 12) X(2, 64, 4, 16) X(3, 76, 4, 19) X(4, 80, 4, 20) X(5, 96, 8, 12) X(6, 104, 8, 13) \
  X(7, 112, 8, 14) X(8, 128, 8, 16) X(9, 136, 8, 17) X(10, 152, 8, 19) X(11, 160, 8, 20) X(12, 176, 16, 11) X(13, 192, 16, 12) \
  X(14, 208, 16, 13) X(15, 224, 16, 14) X(16, 240, 16, 15) X(17, 256, 16, 16)
#define X(i, l, g, k) {l, g, k},
const FwdCfg kFwdCfgs[] = {UBU_FWD_CFGS(X)};
#undef X
constexpr int kNumCfgs = (int)(sizeof(kFwdCfgs) / sizeof(kFwdCfgs[0]));
constexpr size_t kMaxDynSmem = 200 * 1024;

constexpr int TB_THREADS = 128, TB_GRID_PER_SM = 3;      // register-band traceback: one thread per task pair
constexpr int TBS_GRID_PER_SM = UBU_TBS_MIN_BLOCKS;      // wide bands: row strips in registers, one thread per task pair (sw_tbstrip.cuh)
constexpr int kStripBwCap = TBS_BW_MAX;                  // widest band the strip kernel takes: every band default scoring allows for reads <= 256 bp
constexpr size_t kStripScratchBudget = (size_t)6 << 30;  // per pipeline slot; fewer CTAs are launched when a batch's worst case would need more
constexpr int CTR_CIG = TB_CLASSES, CTR_ERR = TB_CLASSES + 1, CTR_WORK = TB_CLASSES + 2;   // d_ctr: class counts, CIGAR cursor, error flags, work counters (+2: long reads)
constexpr int kTbPhases = 8;                       // traceback phases of a large batch (do_run)
constexpr int CTR_SNAP = 16;                       // [16 * g, +TB_CLASSES): class counts after forward group g (g < kTbPhases - 1)
constexpr int CTR_ZERO = 16 * kTbPhases;           // TB_CLASSES zeros: the first phase's list starts
constexpr int CTR_STRIP_WORK = CTR_ZERO + 16;      // [2 * phase + launch] unit counters of the strip kernel's launches
constexpr int CTR_WORDS = CTR_STRIP_WORK + 2 * kTbPhases + 8;
constexpr int kStripBwCommon = 120;                      // bands up to this width go to the full-grid strip launch (TB classes 4..7)
constexpr int TBW_THREADS = 64, TBW_GRID_PER_SM = 2;     // scalar last resort
constexpr int kMaxK = 20;                                // largest K in UBU_FWD_CFGS: i_hi - i_lo <= kMaxK - 1

template <class T>
struct DevBuf {
  T* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct PinBuf {
  T* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 256;
    cudaError_t e = cudaHostAlloc((void**)&p, want * sizeof(T), cudaHostAllocPortable);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr int kAux = 4;
constexpr int kMaxSlots = 8;                          // depth of the H2D / compute / D2H pipeline (ubu_sw_create)
struct FwdLaunch { int cfg; bool has_n; uint32_t begin, n_slots; int wmax; uint32_t n_tasks; bool shared_win = false; };   // shared_win: both tasks of every pair sweep the same window   // selectors go to global scratch when they do not fit shared memory

struct Slot {
  cudaStream_t stream = nullptr;
  cudaStream_t hi[kAux] = {};                            // side streams above the main stream's priority, for traceback phases (made on first use)
  cudaStream_t aux[kAux] = {};                           // the traceback classes are independent: they run side by side
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {};
  DevBuf<uint8_t> d_rp, d_rn, d_wp, d_wn;
  DevBuf<uint32_t> d_roff, d_rnoff, d_woff, d_wnoff, d_idx, d_order, d_lists, d_cig;
  DevBuf<uint16_t> d_rlen, d_wlen;
  DevBuf<FwdOut> d_fwd;
  DevBuf<ubu_sw_result> d_res;
  DevBuf<unsigned int> d_ctr;            // [0..TB_CLASSES) class counts, then CTR_CIG, CTR_ERR, CTR_WORK
  DevBuf<uint32_t> d_hstore;             // packed H of every band cell of the register-band kernels: [BW=8 | BW=16 | BW=32] regions
  size_t h_off[3] = {0, 0, 0};
  int tb_grid = 1;                       // CTAs of the register-band kernels (sized by the batch)
  StripCaps caps{0, 0};
  DevBuf<unsigned char> d_strip;         // H store + boundary arrays of the strip kernel's warps
  struct StripLaunch { int grid = 0, bw_cap = 0, cls_lo = 0, cls_hi = 0; size_t h_off = 0, b_off = 0, h_stride = 0, b_stride = 0; } strip[2];   // common widths | very wide
  int strip_bytes = 1;
  int tbw_grid = 1;
  bool pre_queued = false; size_t pre_words = 0;   // early device->host copies queued by submit()
  DevBuf<unsigned char> d_wide;          // scratch of the scalar last-resort traceback kernel
  DevBuf<unsigned char> d_gsel;          // per-column selectors of forward launches with very long windows
  PinBuf<uint32_t> h_order;
  PinBuf<uint32_t> h_pool;               // pinned staging of a streamed chunk's CIGAR ops
  uint32_t st_lo = 0, st_hi = 0;         // chunk of the streamed batch this slot holds
  // A streamed chunk keeps the caller's ABSOLUTE offsets and window indices: the device copies hold only the chunk's stretch,
  // and the pointers handed to the kernels are moved back by these amounts instead (no host pass rewrites anything).
  size_t bias_rp = 0, bias_rn = 0, bias_wp = 0, bias_wn = 0; uint32_t bias_w = 0;
  PinBuf<unsigned int> h_ctr;
  std::vector<FwdLaunch> launches;
  std::vector<uint32_t> long_tasks;      // reads beyond the packed classes (L > 256): sw_long_kernel, one warp per task
  DevBuf<unsigned char> d_long; size_t long_per_warp = 0; int long_grid = 0;
  std::vector<uint32_t> tmp_keys, tmp_end, tmp_sort, tmp_hist, plain_order;     // plain_order: the order array when planning without a device
  // staged batch
  bool staged = false, ran = false, pending = false;
  uint32_t n_tasks = 0, n_slots = 0;
  bool any_n = false, have_idx = false, reads_n = false, wins_n = false;
  bool identity_order = false;           // uniform batch: tasks run in input order, no order array at all
  int lmax = 0, bw_cap = 0;
  unsigned int cig_cap = 0;
  unsigned int cig_base = 0;             // offset of this batch's ops in the caller's pool (multi-GPU driver), 0 otherwise
  int n_launches = 0;
  // submit()/wait() bookkeeping
  ubu_sw_result* out = nullptr; uint32_t* pool = nullptr; size_t pool_cap = 0;
};

}  // namespace

struct ubu_sw_handle {
  int device = 0, n_slots = 2, sm_count = 148, prio_hi = 0;
  Scoring sc = make_scoring(1, -3, 5, 2);
  Slot slots[kMaxSlots];
  std::string last_error;
  int next_slot = 0;
  // resident reference sequences + scratch of the post-alignment calls (ubu_sw_load_refs / recount / project)
  DevBuf<uint8_t> g_packed, g_nmask, p_bytes;
  DevBuf<uint64_t> g_off, g_noff;
  DevBuf<uint32_t> g_len, p_words;
  uint32_t g_count = 0; bool g_has_n = false;
};

namespace {

int fail_cuda(ubu_sw_handle* h, cudaError_t e, const char* what) {
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  if (h) h->last_error = buf;
  return UBU_SW_ERR_CUDA;
}
#define CU(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return fail_cuda((h), e__, #call); } while (0)

bool env_off(const char* name) { const char* v = getenv(name); return v && *v && *v != '0'; }

// ---- host passes over a batch's metadata: they sit in front of every upload, so they get AVX2 bodies when the CPU has them ----
#if defined(__x86_64__)
#define UBU_AVX2 __attribute__((target("avx2")))
bool cpu_has_avx2() { static const bool v = __builtin_cpu_supports("avx2"); return v; }
#else
#define UBU_AVX2
bool cpu_has_avx2() { return false; }
#endif
template <class T> inline void minmax_body(const T* p, size_t n, T& mn, T& mx) {
  T a = mn, b = mx;
  for (size_t k = 0; k < n; ++k) { a = p[k] < a ? p[k] : a; b = p[k] > b ? p[k] : b; }
  mn = a; mx = b;
}
UBU_AVX2 void minmax_u32_avx2(const uint32_t* p, size_t n, uint32_t& mn, uint32_t& mx) { minmax_body(p, n, mn, mx); }
UBU_AVX2 void minmax_u16_avx2(const uint16_t* p, size_t n, uint16_t& mn, uint16_t& mx) { minmax_body(p, n, mn, mx); }
inline void minmax_u32(const uint32_t* p, size_t n, uint32_t& mn, uint32_t& mx) { if (cpu_has_avx2()) minmax_u32_avx2(p, n, mn, mx); else minmax_body(p, n, mn, mx); }
inline void minmax_u16(const uint16_t* p, size_t n, uint16_t& mn, uint16_t& mx) { if (cpu_has_avx2()) minmax_u16_avx2(p, n, mn, mx); else minmax_body(p, n, mn, mx); }
// sequences k = 0..m-1 at off[k], len[k] bases at `per` bases per byte: laid out in ascending, non-overlapping order?
inline uint32_t ascending_body(const uint32_t* off, const uint16_t* len, uint32_t m, uint32_t per) {
  uint32_t bad = 0;
  const uint32_t sh = per == 4u ? 2u : 3u, add = per - 1u;            // per is 4 (2-bit codes) or 8 (N plane)
  for (uint32_t k = 0; k + 1 < m; ++k) {
    const uint32_t d = off[k + 1] - off[k], need = ((uint32_t)len[k] + add) >> sh;
    bad |= (uint32_t)(off[k + 1] < off[k]) | (uint32_t)(d < need);
  }
  return bad;
}
UBU_AVX2 uint32_t ascending_avx2(const uint32_t* off, const uint16_t* len, uint32_t m, uint32_t per) { return ascending_body(off, len, m, per); }
inline bool ascending(const uint32_t* off, const uint16_t* len, uint32_t m, uint32_t per) {
  return (cpu_has_avx2() ? ascending_avx2(off, len, m, per) : ascending_body(off, len, m, per)) == 0u;
}

int cfg_of_len(int L) {                 // read-length class of a read of L bases (-1: longer than the packed kernels take)
  static const std::vector<int8_t> lut = [] {
    std::vector<int8_t> t((size_t)kFwdCfgs[kNumCfgs - 1].lmax + 1, (int8_t)-1);
    for (int len = 0; len < (int)t.size(); ++len)
      for (int c = 0; c < kNumCfgs; ++c) if (len <= kFwdCfgs[c].lmax) { t[len] = (int8_t)c; break; }
    return t;
  }();
  return L >= 0 && L < (int)lut.size() ? lut[L] : -1;
}

int sel_stride_for(int wmax, int G) { return ((wmax + 2 * G + 63) / 64) * 64 + G; }


template <int G, int K, bool HAS_N, bool DEF>
cudaError_t launch_fwd_t(const FwdLaunch& L, const DevSeqs& reads, const DevSeqs& wins, const uint32_t* idx,
                         const uint32_t* order, const Scoring& sc, FwdOut* out, void* gsel, const TbPlan& plan, cudaStream_t st) {
  constexpr int GP = FWD_THREADS / G;
  const int stride = sel_stride_for(L.wmax, G);
  const size_t smem = (size_t)GP * stride * (HAS_N ? 4u : 2u);
  const uint32_t pairs = L.n_slots / 2;
  const uint32_t grid = (pairs + GP - 1) / GP;
  const uint32_t* ord = order ? order + L.begin : nullptr;
  if (smem > kMaxDynSmem) {            // long windows: selectors in global scratch (sized by the caller)
    sw_fwd16_kernel<G, K, HAS_N, DEF, true, FWD_THREADS><<<grid, FWD_THREADS, 0, st>>>(reads, wins, idx, ord, L.n_slots, L.n_tasks, sc, out, stride, gsel, plan);
    return cudaGetLastError();
  }
  auto kern = sw_fwd16_kernel<G, K, HAS_N, DEF, false, FWD_THREADS>;
  if (smem > 48 * 1024) {      // per device, so set on every large launch (cheap)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return e;
  }
  kern<<<grid, FWD_THREADS, smem, st>>>(reads, wins, idx, ord, L.n_slots, L.n_tasks, sc, out, stride, nullptr, plan);
  return cudaGetLastError();
}

#ifndef UBU_FWDS_THREADS
#define UBU_FWDS_THREADS 32
#endif
constexpr int FWDS_THREADS = UBU_FWDS_THREADS;
int fwds_sel_stride(int wmax, int G) { return ((wmax + 2 * G + 15) / 16) * 16; }

// Shared-window pairs: query profile in shared memory (sw_fwd16s.cuh). Returns cudaErrorInvalidConfiguration when the profile +
// selectors of a CTA do not fit shared memory; the caller then uses the general kernel.
template <int G, int K, bool HAS_N, bool DEF>
cudaError_t launch_fwds_t(const FwdLaunch& L, const DevSeqs& reads, const DevSeqs& wins, const uint32_t* idx,
                          const uint32_t* order, const Scoring& sc, FwdOut* out, const TbPlan& plan, cudaStream_t st) {
  constexpr int GP = FWDS_THREADS / G;
  const int stride = fwds_sel_stride(L.wmax, G);
  const size_t smem = FwdsLayout<G, K, HAS_N>::smem_bytes(GP, stride);
  if (smem > kMaxDynSmem) return cudaErrorInvalidConfiguration;
  const uint32_t pairs = L.n_slots / 2;
  const uint32_t grid = (pairs + GP - 1) / GP;
  const uint32_t* ord = order ? order + L.begin : nullptr;
  auto kern = sw_fwd16s_kernel<G, K, HAS_N, DEF, FWDS_THREADS>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem);
    if (e != cudaSuccess) return e;
  }
  kern<<<grid, FWDS_THREADS, smem, st>>>(reads, wins, idx, ord, L.n_slots, L.n_tasks, sc, out, stride, plan);
  return cudaGetLastError();
}

cudaError_t launch_fwd(const FwdLaunch& L, const DevSeqs& reads, const DevSeqs& wins, const uint32_t* idx,
                       const uint32_t* order, const Scoring& sc, FwdOut* out, void* gsel, const TbPlan& plan, cudaStream_t st) {
  const bool def = sc.match == 1 && sc.mismatch == -3 && sc.gap_open == 5 && sc.gap_ext == 2;
  if (L.shared_win) {
    cudaError_t e = cudaErrorInvalidValue;
    switch (L.cfg * 4 + (L.has_n ? 1 : 0) + (def ? 2 : 0)) {
#define X(i, l, g, k) \
      case i * 4: e = launch_fwds_t<g, k, false, false>(L, reads, wins, idx, order, sc, out, plan, st); break; \
      case i * 4 + 1: e = launch_fwds_t<g, k, true, false>(L, reads, wins, idx, order, sc, out, plan, st); break; \
      case i * 4 + 2: e = launch_fwds_t<g, k, false, true>(L, reads, wins, idx, order, sc, out, plan, st); break; \
      case i * 4 + 3: e = launch_fwds_t<g, k, true, true>(L, reads, wins, idx, order, sc, out, plan, st); break;
      UBU_FWD_CFGS(X)
#undef X
      default: break;
    }
    if (e != cudaErrorInvalidConfiguration) return e;          // does not fit shared memory: general kernel below
  }
  switch (L.cfg * 4 + (L.has_n ? 1 : 0) + (def ? 2 : 0)) {
#define X(i, l, g, k) \
    case i * 4: return launch_fwd_t<g, k, false, false>(L, reads, wins, idx, order, sc, out, gsel, plan, st); \
    case i * 4 + 1: return launch_fwd_t<g, k, true, false>(L, reads, wins, idx, order, sc, out, gsel, plan, st); \
    case i * 4 + 2: return launch_fwd_t<g, k, false, true>(L, reads, wins, idx, order, sc, out, gsel, plan, st); \
    case i * 4 + 3: return launch_fwd_t<g, k, true, true>(L, reads, wins, idx, order, sc, out, gsel, plan, st);
    UBU_FWD_CFGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_strip(bool has_n, bool def, int bytes, int grid, const TbArgs& a, const StripArgs& sa, cudaStream_t st) {
  switch ((has_n ? 1 : 0) + (def ? 2 : 0) + (bytes == 2 ? 4 : 0)) {
    case 0: tb_strip_kernel<false, false, 1><<<grid, TBS_THREADS, 0, st>>>(a, sa); break;
    case 1: tb_strip_kernel<true, false, 1><<<grid, TBS_THREADS, 0, st>>>(a, sa); break;
    case 2: tb_strip_kernel<false, true, 1><<<grid, TBS_THREADS, 0, st>>>(a, sa); break;
    case 3: tb_strip_kernel<true, true, 1><<<grid, TBS_THREADS, 0, st>>>(a, sa); break;
    case 4: tb_strip_kernel<false, false, 2><<<grid, TBS_THREADS, 0, st>>>(a, sa); break;
    case 5: tb_strip_kernel<true, false, 2><<<grid, TBS_THREADS, 0, st>>>(a, sa); break;
    case 6: tb_strip_kernel<false, true, 2><<<grid, TBS_THREADS, 0, st>>>(a, sa); break;
    default: tb_strip_kernel<true, true, 2><<<grid, TBS_THREADS, 0, st>>>(a, sa); break;
  }
  return cudaGetLastError();
}

bool params_ok(const ubu_sw_params& p) {
  return p.match >= 1 && p.match <= 31 && p.mismatch <= 0 && p.mismatch >= -64 && p.gap_open >= 0 && p.gap_open <= 63 &&
         p.gap_ext >= 1 && p.gap_ext <= 31;
}

int validate_seqs(const ubu_sw_seqs* s) {
  if (!s) return UBU_SW_ERR_ARG;
  if (s->count && (!s->off || !s->len)) return UBU_SW_ERR_ARG;
  if (s->count && s->packed_bytes && !s->packed) return UBU_SW_ERR_ARG;
  if (s->nmask && !s->nmask_off) return UBU_SW_ERR_ARG;
  return UBU_SW_OK;
}

// ---- plan: class per task, order by (class, window length), pad classes to pairs ------------------
// A chunk of a streamed batch (align_stream): the seqs structs keep the caller's full buffers and ABSOLUTE offsets / window
// indices, only their off / len pointers are advanced to the chunk; the stretch of bytes the chunk really uses (and that is
// copied to the device) is described here, and the device pointers are moved back by the same amounts in do_run.
struct StageBias {
  size_t rp, r_bytes, rn, rn_bytes;      // reads: first byte and byte count in packed / nmask
  size_t wp, w_bytes, wn, wn_bytes;      // windows
  uint32_t w;                            // index of the chunk's first window
};

// trusted = the caller (the streamed align_batch) has already checked every offset and index of this chunk
int plan_batch(ubu_sw_handle* h, Slot& s, const ubu_sw_seqs* reads, const ubu_sw_seqs* wins, const uint32_t* idx, bool trusted = false,
               const StageBias* bias = nullptr) {
  const uint32_t idx_bias = bias ? bias->w : 0u;
  const uint32_t n = reads->count;
  s.identity_order = false;
  // ---- uniform batch (one read-length class, one window length, no N in any window): the common shape of a
  //      realignment batch. Only bounds are checked (vectorisable passes); tasks keep their input order.
  {
    uint16_t lmin = 0xFFFF, lmax = 0, wmin = 0xFFFF, wmax = 0;
    minmax_u16(reads->len, n, lmin, lmax);
    minmax_u16(wins->len, wins->count, wmin, wmax);
    bool win_n = false;
    if (wins->nmask) {
      if (wins->count && (uint64_t)wins->nmask_off[wins->count - 1] + ((uint32_t)wins->len[wins->count - 1] + 7u) / 8u > wins->nmask_bytes) return UBU_SW_ERR_ARG;
      // any N at all in the window planes? (bytes between sequences are zero in every packer of this repo; a stray bit only costs speed)
      const uint8_t* m = wins->nmask + (bias ? bias->wn : 0); uint8_t any = 0;
      const size_t mbytes = bias ? bias->wn_bytes : wins->nmask_bytes;
      for (size_t k = 0; k < mbytes; ++k) any |= m[k];
      win_n = any != 0;
    }
    const int c0 = cfg_of_len(lmin), c1 = cfg_of_len(lmax);
    if (n && wins->count && lmax > LONG_MAX_READ) return UBU_SW_ERR_TOO_LONG;
    s.long_tasks.clear();
    if (n && wins->count && c1 >= 0 && c0 == c1 && wmin == wmax && !win_n) {
      uint32_t bad = 0, imax = 0;
      const uint32_t wbytes = ((uint32_t)wmax + 3u) / 4u;
      if (!trusted) {
        // every read is checked with ITS OWN length: reads of one class may differ in length, and packed_bytes is exact
        for (uint32_t t = 0; t < n; ++t) bad |= (uint32_t)((uint64_t)reads->off[t] + ((uint32_t)reads->len[t] + 3u) / 4u > reads->packed_bytes);
        for (uint32_t w = 0; w < wins->count; ++w) bad |= (uint32_t)((uint64_t)wins->off[w] + wbytes > wins->packed_bytes);
        if (idx) { for (uint32_t t = 0; t < n; ++t) imax = std::max(imax, idx[t]); if (imax >= wins->count) bad = 1; }
        if (reads->nmask) for (uint32_t t = 0; t < n; ++t) bad |= (uint32_t)((uint64_t)reads->nmask_off[t] + ((uint32_t)reads->len[t] + 7u) / 8u > reads->nmask_bytes);
        if (wins->nmask) for (uint32_t w = 0; w < wins->count; ++w) bad |= (uint32_t)((uint64_t)wins->nmask_off[w] + ((uint32_t)wmax + 7u) / 8u > wins->nmask_bytes);
      }
      if (bad) return UBU_SW_ERR_ARG;
      s.lmax = lmax;
      s.n_slots = (n + 1u) & ~1u;
      s.identity_order = true;
      s.launches.clear();
      FwdLaunch L{c1, false, 0u, s.n_slots, (int)wmax, n};
      if (idx && !env_off("UBU_NO_SHARED_WIN")) {             // do the two tasks of every pair share their window?
        uint32_t diff = 0;
        for (uint32_t t = 0; t + 1 < n; t += 2) diff |= idx[t] ^ idx[t + 1];
        L.shared_win = diff == 0;
        if (!L.shared_win && (uint64_t)wins->count * 3u <= (uint64_t)n * 2u) {
          // Many reads per window, but not as neighbours: group the reads by window (counting sort, two passes) so that
          // most pairs share their window and run on the query-profile kernel; the odd read of a window joins the
          // general launch behind them.
          const uint32_t nw = wins->count;
          s.tmp_keys.assign((size_t)2 * nw + 2, 0u);
          uint32_t* cnt = s.tmp_keys.data(); uint32_t* pos = cnt + nw + 1;
          for (uint32_t t = 0; t < n; ++t) ++cnt[idx[t] - idx_bias];
          uint32_t n_pair_slots = 0, n_left = 0;
          for (uint32_t w = 0; w < nw; ++w) { n_pair_slots += cnt[w] & ~1u; n_left += cnt[w] & 1u; }
          const uint32_t left_slots = (n_left + 1u) & ~1u;
          s.n_slots = n_pair_slots + left_slots;
          if (h) { cudaError_t e = s.h_order.reserve(std::max<size_t>(s.n_slots, 2)); if (e != cudaSuccess) return fail_cuda(h, e, "cudaHostAlloc(order)"); }
          else s.plain_order.assign(std::max<size_t>(s.n_slots, 2), 0u);
          uint32_t* order = h ? s.h_order.p : s.plain_order.data();
          std::vector<uint32_t>& endv = s.tmp_end;           // endv[w]: end of window w's pair slots
          endv.resize(nw);
          uint32_t pp = 0, lp = n_pair_slots;
          for (uint32_t w = 0; w < nw; ++w) {                // pos[w]: next pair slot of window w
            const uint32_t c = cnt[w];
            pos[w] = pp; pp += c & ~1u; endv[w] = pp;
            cnt[w] = (c & 1u) ? lp++ : 0xFFFFFFFFu;          // cnt[w] now = the slot of window w's odd read (general launch)
          }
          for (uint32_t t = 0; t < n; ++t) {
            const uint32_t w = idx[t] - idx_bias;
            if (pos[w] < endv[w]) order[pos[w]++] = t; else order[cnt[w]] = t;
          }
          if (n_left & 1u) order[s.n_slots - 1] = 0xFFFFFFFFu;
          s.launches.clear();
          if (n_pair_slots) { FwdLaunch A{c1, false, 0u, n_pair_slots, (int)wmax, 0u}; A.shared_win = true; s.launches.push_back(A); }
          if (left_slots) s.launches.push_back(FwdLaunch{c1, false, n_pair_slots, left_slots, (int)wmax, 0u});
          s.identity_order = false;
          return UBU_SW_OK;
        }
      }
      s.launches.push_back(L);
      return UBU_SW_OK;
    }
  }
  std::vector<uint8_t> win_has_n;
  if (wins->nmask) {
    win_has_n.assign(wins->count, 0);
    for (uint32_t w = 0; w < wins->count; ++w) {
      const uint32_t nb = ((uint32_t)wins->len[w] + 7u) / 8u;
      if ((uint64_t)wins->nmask_off[w] + nb > wins->nmask_bytes) return UBU_SW_ERR_ARG;
      const uint8_t* m = wins->nmask + wins->nmask_off[w];
      uint8_t any = 0;
      for (uint32_t k = 0; k < nb; ++k) any |= m[k];
      win_has_n[w] = any ? 1 : 0;
    }
  }
  for (uint32_t w = 0; w < wins->count; ++w)
    if ((uint64_t)wins->off[w] + ((uint32_t)wins->len[w] + 3u) / 4u > wins->packed_bytes) return UBU_SW_ERR_ARG;

  constexpr int NB = kNumCfgs * 2;
  uint32_t cnt[NB] = {0};
  int wmin[NB], wmaxv[NB];
  for (int b = 0; b < NB; ++b) { wmin[b] = 1 << 30; wmaxv[b] = 0; }
  s.tmp_keys.resize(n);
  int lmax = 0;
  for (uint32_t t = 0; t < n; ++t) {
    const int L = reads->len[t];
    const uint32_t ri = idx ? idx[t] - idx_bias : t;
    if (ri >= wins->count) return UBU_SW_ERR_ARG;
    if (!trusted) {
      if ((uint64_t)reads->off[t] + ((uint32_t)L + 3u) / 4u > reads->packed_bytes) return UBU_SW_ERR_ARG;
      if (reads->nmask && (uint64_t)reads->nmask_off[t] + ((uint32_t)L + 7u) / 8u > reads->nmask_bytes) return UBU_SW_ERR_ARG;
    }
    const int c = cfg_of_len(L);
    const int W = wins->len[ri];
    if (c < 0) {                          // a contig-sized query: its own kernel, outside the length classes
      if ((uint32_t)L > LONG_MAX_READ || (uint64_t)L * (uint64_t)W > LONG_MAX_CELLS) return UBU_SW_ERR_TOO_LONG;
      s.long_tasks.push_back(t);
      s.tmp_keys[t] = 0xFFFFFFFFu;
      continue;
    }
    const int b = c * 2 + (wins->nmask && win_has_n[ri] ? 1 : 0);
    ++cnt[b];
    wmin[b] = std::min(wmin[b], W); wmaxv[b] = std::max(wmaxv[b], W);
    lmax = std::max(lmax, L);
    s.tmp_keys[t] = ((uint32_t)b << 16) | (uint32_t)W;
  }
  s.lmax = lmax;

  // bucket starts (each bucket padded to an even number of slots)
  uint32_t start[NB + 1]; start[0] = 0;
  for (int b = 0; b < NB; ++b) start[b + 1] = start[b] + ((cnt[b] + 1u) & ~1u);
  s.n_slots = start[NB];
  if (h) { cudaError_t e = s.h_order.reserve(std::max<size_t>(s.n_slots + s.long_tasks.size(), 2)); if (e != cudaSuccess) return fail_cuda(h, e, "cudaHostAlloc(order)"); }
  else s.plain_order.assign(std::max<size_t>(s.n_slots + s.long_tasks.size(), 2), 0u);
  uint32_t* order = h ? s.h_order.p : s.plain_order.data();
  std::copy(s.long_tasks.begin(), s.long_tasks.end(), order + s.n_slots);          // the long tasks follow the slots
  uint32_t fill[NB];
  for (int b = 0; b < NB; ++b) fill[b] = start[b];
  for (uint32_t t = 0; t < n; ++t) if (s.tmp_keys[t] != 0xFFFFFFFFu) order[fill[s.tmp_keys[t] >> 16]++] = t;
  for (int b = 0; b < NB; ++b) if (cnt[b] & 1u) order[fill[b]] = 0xFFFFFFFFu;

  // inside a bucket, order by window length when it varies (so a pair / a warp sweeps similar widths)
  s.launches.clear();
  for (int b = 0; b < NB; ++b) {
    if (!cnt[b]) continue;
    uint32_t* lo = order + start[b];
    if (wmin[b] != wmaxv[b]) {
      // stable counting sort by window length (this runs per streamed chunk, in front of its upload: a comparison sort through the
      // key array cost more host time than the chunk takes on the GPU)
      const uint32_t range = (uint32_t)(wmaxv[b] - wmin[b]) + 1u;
      s.tmp_hist.assign((size_t)range + 1, 0u);
      s.tmp_sort.resize(cnt[b]);
      uint32_t* hist = s.tmp_hist.data();
      for (uint32_t k = 0; k < cnt[b]; ++k) ++hist[(s.tmp_keys[lo[k]] & 0xFFFFu) - (uint32_t)wmin[b] + 1u];
      for (uint32_t r = 0; r < range; ++r) hist[r + 1] += hist[r];
      for (uint32_t k = 0; k < cnt[b]; ++k) s.tmp_sort[hist[(s.tmp_keys[lo[k]] & 0xFFFFu) - (uint32_t)wmin[b]]++] = lo[k];
      std::copy(s.tmp_sort.begin(), s.tmp_sort.end(), lo);
    }
    const int cfg = b / 2; const bool has_n = b & 1;
    const uint32_t total = (cnt[b] + 1u) & ~1u;
    // Neighbours that share a window (mates, coordinate-sorted reads of one region) are paired up for the query-profile kernel;
    // the others follow in a general-kernel stretch. Both stretches keep the bucket's order (ascending window length).
    uint32_t n_shared = 0;
    if (idx && !env_off("UBU_NO_SHARED_WIN")) {
      std::vector<uint32_t>& rest = s.tmp_end;
      rest.clear();
      uint32_t k = 0;
      while (k < cnt[b]) {
        if (k + 1 < cnt[b] && idx[lo[k]] == idx[lo[k + 1]]) { lo[n_shared] = lo[k]; lo[n_shared + 1] = lo[k + 1]; n_shared += 2; k += 2; }
        else rest.push_back(lo[k++]);                       // (writes to lo[] trail the read position: n_shared <= k)
      }
      std::copy(rest.begin(), rest.end(), lo + n_shared);
      if (rest.size() & 1u) lo[n_shared + rest.size()] = 0xFFFFFFFFu;
    }
    // split a stretch whose widths vary a lot so that short windows do not pay the long ones' shared memory
    auto add_launches = [&](uint32_t first, uint32_t count_real, bool shared) {
      const uint32_t stretch_total = (count_real + 1u) & ~1u;
      uint32_t beg = 0;
      while (beg < stretch_total) {
        uint32_t end = stretch_total;
        if (wmin[b] != wmaxv[b]) {
          const int w0 = (int)(s.tmp_keys[lo[first + beg]] & 0xFFFFu);
          const int limit = std::max(w0 * 2, 512);
          uint32_t e = beg;
          while (e < count_real && (int)(s.tmp_keys[lo[first + e]] & 0xFFFFu) <= limit) ++e;
          e = (e + 1u) & ~1u;                 // keep pairs intact
          if (e > stretch_total) e = stretch_total;
          if (e < stretch_total && e > beg) end = e;
        }
        int wm = 0;
        for (uint32_t k = beg; k < end && k < count_real; ++k) wm = std::max(wm, (int)(s.tmp_keys[lo[first + k]] & 0xFFFFu));
        FwdLaunch L{cfg, has_n, start[b] + first + beg, end - beg, wm, 0u};
        L.shared_win = shared;
        s.launches.push_back(L);
        beg = end;
      }
    };
    if (n_shared) add_launches(0u, n_shared, true);
    if (cnt[b] > n_shared) add_launches(n_shared, cnt[b] - n_shared, false);
    (void)total;
  }
  return UBU_SW_OK;
}

int do_stage(ubu_sw_handle* h, Slot& s, const ubu_sw_seqs* reads, const ubu_sw_seqs* wins, const uint32_t* idx, bool trusted = false,
             const StageBias* bias = nullptr) {
  int rc = validate_seqs(reads); if (rc) return rc;
  rc = validate_seqs(wins); if (rc) return rc;
  if (!idx && reads->count != wins->count) return UBU_SW_ERR_ARG;
  s.staged = false; s.ran = false; s.pre_queued = false; s.cig_base = 0;
  s.n_tasks = reads->count;
  s.have_idx = idx != nullptr;
  s.reads_n = reads->nmask != nullptr; s.wins_n = wins->nmask != nullptr;
  s.any_n = s.reads_n || s.wins_n;
  if (s.n_tasks == 0) { s.n_slots = 0; s.launches.clear(); s.staged = true; return UBU_SW_OK; }
  rc = plan_batch(h, s, reads, wins, idx, trusted, bias); if (rc) return rc;
  // what goes to the device: the whole buffers, or only the stretch a streamed chunk uses
  const size_t r_lo = bias ? bias->rp : 0, r_bytes = bias ? bias->r_bytes : reads->packed_bytes;
  const size_t rn_lo = bias ? bias->rn : 0, rn_bytes = bias ? bias->rn_bytes : reads->nmask_bytes;
  const size_t w_lo = bias ? bias->wp : 0, w_bytes = bias ? bias->w_bytes : wins->packed_bytes;
  const size_t wn_lo = bias ? bias->wn : 0, wn_bytes = bias ? bias->wn_bytes : wins->nmask_bytes;
  s.bias_rp = r_lo; s.bias_rn = rn_lo; s.bias_wp = w_lo; s.bias_wn = wn_lo; s.bias_w = bias ? bias->w : 0u;

  const uint32_t n = s.n_tasks, nw = wins->count;
  {   // launches whose selectors do not fit shared memory keep them in a global scratch (one launch at a time on the stream)
    size_t need = 0;
    for (const FwdLaunch& L : s.launches) {
      const FwdCfg& c = kFwdCfgs[L.cfg];
      const size_t gp = FWD_THREADS / c.G, stride = (size_t)sel_stride_for(L.wmax, c.G), selsz = L.has_n ? 4 : 2;
      if (gp * stride * selsz > kMaxDynSmem) need = std::max(need, (((size_t)L.n_slots / 2 + gp - 1) / gp) * gp * stride * selsz);
    }
    if (need) CU(h, s.d_gsel.reserve(need + 64));
  }
  CU(h, s.d_rp.reserve(r_bytes + 16)); CU(h, s.d_roff.reserve(n)); CU(h, s.d_rlen.reserve(n));
  CU(h, s.d_wp.reserve(w_bytes + 16)); CU(h, s.d_woff.reserve(nw)); CU(h, s.d_wlen.reserve(nw));
  if (s.reads_n) { CU(h, s.d_rn.reserve(rn_bytes + 16)); CU(h, s.d_rnoff.reserve(n)); }
  if (s.wins_n) { CU(h, s.d_wn.reserve(wn_bytes + 16)); CU(h, s.d_wnoff.reserve(nw)); }
  if (idx) CU(h, s.d_idx.reserve(n));
  CU(h, s.d_order.reserve(s.n_slots + s.long_tasks.size())); CU(h, s.d_lists.reserve((size_t)TB_CLASSES * s.n_slots));
  CU(h, s.d_fwd.reserve(n)); CU(h, s.d_res.reserve(n));
  CU(h, s.d_ctr.reserve(CTR_WORDS)); CU(h, s.h_ctr.reserve(16));
  if (s.cig_cap < 6u * n + 4096u) { s.cig_cap = 6u * n + 4096u; }
  CU(h, s.d_cig.reserve(s.cig_cap));
  // traceback scratch
  const int max_rows = std::max(s.lmax, 1);
  {
    const int gbmax = gap_bound(h->sc, max_rows, 1);
    s.bw_cap = std::min(kMaxK + 2 * gbmax + 1, max_rows + 65535);
    // grids are sized by the batch (a small batch must not pin the scratch of a full one)
    const size_t pair_ctas = ((size_t)(n + 1) / 2 + TB_THREADS - 1) / TB_THREADS;
    s.tb_grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)h->sm_count * TB_GRID_PER_SM, pair_ctas));
    const size_t reg_threads = (size_t)s.tb_grid * TB_THREADS;
    s.h_off[0] = 0;
    s.h_off[1] = s.h_off[0] + reg_threads * (size_t)max_rows * 8;
    s.h_off[2] = s.h_off[1] + reg_threads * (size_t)max_rows * 16;
    const size_t hwords = s.h_off[2] + reg_threads * (size_t)max_rows * 32;
    // wide bands: strip kernel. One byte per stored H when no score of the batch can exceed 255, else two.
    s.caps.rows_cap = 0; s.caps.bw_cap = 0; s.strip[0].grid = s.strip[1].grid = 0;
    if (s.bw_cap > 32) {
      s.caps.rows_cap = max_rows; s.caps.bw_cap = std::min(s.bw_cap, kStripBwCap);
      s.strip_bytes = (h->sc.match * max_rows <= 255) ? 1 : 2;
      const size_t strip_ctas = ((size_t)(n + 1) / 2 + TBS_THREADS - 1) / TBS_THREADS;
      size_t bytes = 0;
      for (int k = 0; k < 2; ++k) {
        Slot::StripLaunch& L = s.strip[k];
        L.cls_lo = k == 0 ? TB_STRIP_FIRST : TB_STRIP_FIRST + TB_STRIP_CLASSES - 1;
        L.cls_hi = k == 0 ? TB_STRIP_FIRST + TB_STRIP_CLASSES - 2 : TB_STRIP_FIRST + TB_STRIP_CLASSES - 1;
        L.bw_cap = k == 0 ? std::min(s.caps.bw_cap, kStripBwCommon) : s.caps.bw_cap;
        if (k == 1 && s.caps.bw_cap <= kStripBwCommon) continue;            // no band of this batch can be that wide
        const size_t per_warp_h = tbs_hstore_bytes_per_warp(s.caps.rows_cap, L.bw_cap, s.strip_bytes),
                     per_warp_b = tbs_boundary_bytes_per_warp(s.caps.rows_cap, L.bw_cap);
        L.h_stride = per_warp_h; L.b_stride = per_warp_b;
        const size_t per_cta = (per_warp_h + per_warp_b) * (TBS_THREADS / 32);
        // the wide launch sees ~2 % of an indel-rich batch: a quarter of the SMs' worth of CTAs is plenty
        size_t grid = std::min<size_t>(k == 0 ? (size_t)h->sm_count * TBS_GRID_PER_SM : (size_t)(h->sm_count + 1) / 2, strip_ctas);
        grid = std::max<size_t>(1, std::min<size_t>(grid, kStripScratchBudget / per_cta));
        L.grid = (int)grid;
        L.h_off = bytes; bytes += grid * (TBS_THREADS / 32) * per_warp_h;
        L.b_off = bytes; bytes += grid * (TBS_THREADS / 32) * per_warp_b;
      }
      CU(h, s.d_strip.reserve(bytes + 256));
    }
    CU(h, s.d_hstore.reserve(hwords));
    {
      // the scalar last resort only sees bands no other class takes (extreme scoring / very long windows): its band can be
      // as wide as the whole window, so its scratch is capped at 2 GB by running fewer CTAs
      int wcap_all = 0;
      for (uint32_t w = 0; w < wins->count; ++w) wcap_all = std::max(wcap_all, (int)wins->len[w]);
      s.bw_cap = std::min(s.bw_cap, max_rows + wcap_all);
      const size_t per = tb_wide_scratch_bytes(max_rows, s.bw_cap) * TBW_THREADS;
      s.tbw_grid = (int)std::max<size_t>(1, std::min<size_t>((size_t)h->sm_count * TBW_GRID_PER_SM, ((size_t)2 << 30) / std::max<size_t>(per, 1)));
      CU(h, s.d_wide.reserve((size_t)s.tbw_grid * per));
    }
  }

  s.long_grid = 0;
  if (!s.long_tasks.empty()) {            // contig-sized queries: per-warp scratch sized for the largest of them, a warp per task
    size_t per = 0;
    for (uint32_t t : s.long_tasks) {
      const uint32_t ri = idx ? idx[t] - (bias ? bias->w : 0u) : t;
      per = std::max(per, long_scratch_bytes(reads->len[t], wins->len[ri]));
    }
    per = (per + 255) & ~(size_t)255;
    const size_t warps_per_cta = LONG_THREADS / 32;
    // the kernel is latency-bound (a dependent chain per 32-cell step): it wants ~16 warps per SM, which only the scratch limits
    size_t ctas = std::min<size_t>((s.long_tasks.size() + warps_per_cta - 1) / warps_per_cta, (size_t)h->sm_count * 4);
    ctas = std::max<size_t>(1, std::min<size_t>(ctas, ((size_t)12 << 30) / (per * warps_per_cta)));
    s.long_per_warp = per; s.long_grid = (int)ctas;
    CU(h, s.d_long.reserve(ctas * warps_per_cta * per + 256));
  }
  cudaStream_t st = s.stream;
  CU(h, cudaMemcpyAsync(s.d_rp.p, reads->packed + r_lo, r_bytes, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(s.d_roff.p, reads->off, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(s.d_rlen.p, reads->len, sizeof(uint16_t) * n, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(s.d_wp.p, wins->packed + w_lo, w_bytes, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(s.d_woff.p, wins->off, sizeof(uint32_t) * nw, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(s.d_wlen.p, wins->len, sizeof(uint16_t) * nw, cudaMemcpyHostToDevice, st));
  if (s.reads_n) {
    CU(h, cudaMemcpyAsync(s.d_rn.p, reads->nmask + rn_lo, rn_bytes, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(s.d_rnoff.p, reads->nmask_off, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  }
  if (s.wins_n) {
    CU(h, cudaMemcpyAsync(s.d_wn.p, wins->nmask + wn_lo, wn_bytes, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(s.d_wnoff.p, wins->nmask_off, sizeof(uint32_t) * nw, cudaMemcpyHostToDevice, st));
  }
  if (idx) CU(h, cudaMemcpyAsync(s.d_idx.p, idx, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  if (!s.identity_order) CU(h, cudaMemcpyAsync(s.d_order.p, s.h_order.p, sizeof(uint32_t) * (s.n_slots + s.long_tasks.size()), cudaMemcpyHostToDevice, st));
  s.staged = true;
  return UBU_SW_OK;
}

int do_run(ubu_sw_handle* h, Slot& s) {
  if (!s.staged) return UBU_SW_ERR_ARG;
  cudaStream_t st = s.stream;
  s.n_launches = 0;
  CU(h, cudaEventRecord(s.ev[0], st));
  if (s.n_tasks == 0) { CU(h, cudaEventRecord(s.ev[1], st)); CU(h, cudaEventRecord(s.ev[2], st)); s.ran = true; return UBU_SW_OK; }
  CU(h, cudaMemsetAsync(s.d_ctr.p, 0, CTR_WORDS * sizeof(unsigned int), st));
  // (a streamed chunk keeps absolute offsets and window indices: the pointers are moved back instead, see Slot::bias_*)
  DevSeqs reads{s.d_rp.p - s.bias_rp, s.d_roff.p, s.d_rlen.p, s.reads_n ? s.d_rn.p - s.bias_rn : nullptr, s.reads_n ? s.d_rnoff.p : nullptr};
  DevSeqs wins{s.d_wp.p - s.bias_wp, s.d_woff.p - s.bias_w, s.d_wlen.p - s.bias_w, s.wins_n ? s.d_wn.p - s.bias_wn : nullptr,
               s.wins_n ? s.d_wnoff.p - s.bias_w : nullptr};
  const uint32_t* idx = s.have_idx ? s.d_idx.p : nullptr;
  const uint32_t* d_order = s.identity_order ? nullptr : s.d_order.p;
  // the forward kernels' epilogue also sorts the tasks into the traceback class lists (and finishes unaligned ones)
  const TbPlan plan{s.caps, s.d_res.p, s.d_lists.p, s.d_ctr.p, s.n_slots};
  TbArgs a;
  a.reads = reads; a.wins = wins; a.ref_idx = idx; a.sc = h->sc; a.fwd = s.d_fwd.p; a.res = s.d_res.p;
  a.cig_pool = s.d_cig.p; a.cig_cap = s.cig_cap; a.cig_cursor = s.d_ctr.p + CTR_CIG; a.err_flag = s.d_ctr.p + CTR_ERR; a.cig_base = s.cig_base;
  const int max_rows = std::max(s.lmax, 1);
  const size_t NS = s.n_slots;

  // Traceback phases. The class lists grow in launch order, and a traceback kernel at a third of its grid loses little (it is bound by
  // HBM writes) while it leaves most of each SM to the forward kernels (bound by integer issue). On large mixed-length batches (many
  // forward launches) the launches are cut into groups of about equal cell count, and the traceback of group g (list entries [counts
  // after group g-1, counts after group g), from counter snapshots taken in stream order) runs on the side streams, at a third of its
  // grid, underneath the forward launches of group g+1; only the last group's traceback stays exposed (cfg4, 1M reads: 25.6 - 28.9 ms
  // from handle to handle -> 23.6 ms with the stream schedule and the launch order below). Batches with one or a few forward launches run unphased: with 2 launches cfg3 gains 0.7 % and cfg2+N loses
  // 3 % (a reduced-grid phase outlives the short forward launch behind it), profiles/tb_phases_r02.txt.
  // (development switches, read per run so that a bench can time the phased and the unphased form of one staged batch)
  const int phases_env = [] { const char* v = getenv("UBU_TB_PHASES"); return v && *v ? std::max(1, std::min(kTbPhases, atoi(v))) : 0; }();
  const int gdiv_env = [] { const char* v = getenv("UBU_TB_PHASE_GRID_DIV"); return v && *v ? std::max(1, atoi(v)) : 3; }();
  int n_phases = 1;
  if (phases_env) { if (s.launches.size() > 1) n_phases = std::min<int>(phases_env, (int)s.launches.size()); }   // development switch
  else if (s.n_tasks >= (256u << 10) && s.launches.size() >= 8)
    n_phases = (int)std::min<size_t>(std::min<size_t>(kTbPhases, s.launches.size() / 2), s.n_tasks >> 17);
  // Phased batches run their forward launches longest reads first: the last group, whose traceback nothing hides, is then the one
  // of the shortest reads (narrow bands, cheap traceback): cfg4 24.35 -> 23.6 ms. Groups of equal cell count (shrinking groups were tried:
  // no gain, profiles/tb_phases_r02.txt).
  std::vector<size_t> phase_end(n_phases, s.launches.size());         // phase g = launches fwd_order[phase_end[g-1] .. phase_end[g])
  std::vector<size_t> fwd_order(s.launches.size());
  for (size_t i = 0; i < fwd_order.size(); ++i) fwd_order[i] = i;
  if (n_phases > 1) {
    if (!env_off("UBU_TB_PHASE_NOREV")) std::reverse(fwd_order.begin(), fwd_order.end());
    double total = 0, acc = 0;
    auto cost = [&](const FwdLaunch& L) { return (double)L.n_slots * kFwdCfgs[L.cfg].lmax * std::max(L.wmax, 1); };
    for (const FwdLaunch& L : s.launches) total += cost(L);
    int g = 0;
    for (size_t i = 0; i < fwd_order.size() && g + 1 < n_phases; ++i) {
      acc += cost(s.launches[fwd_order[i]]);
      if (acc >= total * (g + 1) / n_phases) phase_end[g++] = i + 1;
    }
    if (g > 0 && phase_end[g - 1] == s.launches.size()) --g;        // no forward launch left for a last group
    n_phases = g + 1; phase_end.resize(n_phases); phase_end[n_phases - 1] = s.launches.size();
  }
  // traceback kernels side by side on the side streams hide their tails and pay on batches up to ~1M tasks (cfg2: 0.47 vs 0.65 ms,
  // cfg4: 7.3 vs 7.8 ms). On very large batches every kernel fills the GPU by itself and the store-heavy ones only compete for HBM
  // (cfg3, 5M tasks: 9.6 ms side by side, 8.2 ms one after the other): those run in turn, on one stream.
  static const bool tb_serial_env = env_off("UBU_TB_SERIAL");  // development switch: always one after the other
  const bool tb_serial = tb_serial_env || s.n_tasks >= (3u << 20);
  const bool phased = n_phases > 1;
  // Which stream each traceback kernel goes to (issue order below = execution order within a stream: wide strips, common strips,
  // band 32, 16, 8, diagonal, scalar). Uniform-length batches (few forward launches, narrow bands): the register-band kernels side by
  // side, the strip launches on streams of their own (they see a handful of tasks and must not hold anything back). Mixed-length
  // batches (>= 8 forward launches: cfg4), where the strip kernel is most of the traceback: the two store-heavy kernels one after
  // the other on one stream (common strips, then band 32), everything light on a second one. With five streams per slot the first form
  // also depends on which hardware queues the streams share (cfg4: 6.7 - 10.4 ms between handles of ONE process, a fixed sequence);
  // the second gives 6.6 ms on every handle (profiles/tb_phases_r02.txt). Small mixed batches keep the first form: their tails matter
  // more (cfg4 at 100k reads: 3.8 - 4.5 ms, mean 4.0, vs 4.1 - 4.2).
  enum { K_DIAG, K_8, K_16, K_32, K_S0, K_S1, K_W, K_LONG, K_N };
  const int sched_env = [&] { const char* v = getenv("UBU_TB_SCHED"); return v && *v ? atoi(v) : -1; }();     // development switch: 0 / 7
  const int sched = sched_env >= 0 ? sched_env : (s.n_tasks >= (256u << 10) && s.launches.size() >= 8 ? 7 : 0);
  // A traceback phase underneath later forward launches must get its CTAs placed: phased batches use side streams above the main
  // stream's priority (cfg4: 23.6 vs 24.4 ms). Everything else stays on streams of equal priority: the streamed chunks of the
  // end-to-end path lose 1.3 % when each chunk's traceback pushes in front of the next chunk's forward pass.
  if (phased)
    for (int k = 0; k < kAux; ++k) if (!s.hi[k]) CU(h, cudaStreamCreateWithPriority(&s.hi[k], cudaStreamNonBlocking, h->prio_hi));
  cudaStream_t* const sx = phased ? s.hi : s.aux;
  cudaStream_t ks[K_N];
  if (tb_serial && sched_env < 0) for (int k = 0; k < K_N; ++k) ks[k] = phased ? sx[0] : st;
  else if (sched == 7) {
    for (int k = 0; k < K_N; ++k) ks[k] = sx[1];
    ks[K_S0] = sx[0]; ks[K_32] = sx[0]; ks[K_DIAG] = phased ? sx[2] : st;
  } else {
    ks[K_DIAG] = phased ? sx[0] : st; ks[K_8] = ks[K_16] = ks[K_LONG] = sx[0]; ks[K_32] = ks[K_W] = sx[1];
    ks[K_S0] = sx[2]; ks[K_S1] = sx[3];
  }
  const bool def = h->sc.match == 1 && h->sc.mismatch == -3 && h->sc.gap_open == 5 && h->sc.gap_ext == 2;
  const size_t selsz = s.any_n ? 4 : 2;
  const size_t sm8 = (size_t)(max_rows + 8) * TB_THREADS * selsz, sm16 = (size_t)(max_rows + 16) * TB_THREADS * selsz,
               sm32 = (size_t)(max_rows + 32) * TB_THREADS * selsz;
  uint32_t *h8 = s.d_hstore.p + s.h_off[0], *h16 = s.d_hstore.p + s.h_off[1], *h32 = s.d_hstore.p + s.h_off[2];
  static const bool reg16 = env_off("UBU_TB_REG16");            // development switch: the 16-bit store + flag form whatever the scores
  const bool hbyte = h->sc.match * max_rows <= 255 && !reg16;  // no score of the batch exceeds 255: one byte per stored H

  size_t li = 0;
  for (int g = 0; g < n_phases; ++g) {
    for (; li < phase_end[g]; ++li) {
      CU(h, launch_fwd(s.launches[fwd_order[li]], reads, wins, idx, d_order, h->sc, s.d_fwd.p, s.d_gsel.p, plan, st));
      ++s.n_launches;
    }
    const bool last = g + 1 == n_phases;
    const int gdiv = last ? 1 : gdiv_env;
    const int grid = std::max(1, s.tb_grid / gdiv);
    // list ranges of this phase: [first, count) per class
    const unsigned int* first = g == 0 ? s.d_ctr.p + CTR_ZERO : s.d_ctr.p + CTR_SNAP + 16 * (g - 1);
    const unsigned int* count = last ? s.d_ctr.p : s.d_ctr.p + CTR_SNAP + 16 * g;
    if (!last) CU(h, cudaMemcpyAsync(s.d_ctr.p + CTR_SNAP + 16 * g, s.d_ctr.p, TB_CLASSES * sizeof(unsigned int), cudaMemcpyDeviceToDevice, st));
    if (last) CU(h, cudaEventRecord(s.ev[1], st));
    CU(h, cudaEventRecord(s.ev_fork, st));
    for (int k = 0; k < kAux; ++k) CU(h, cudaStreamWaitEvent(sx[k], s.ev_fork, 0));
    auto issue_strip = [&](int k) -> int {             // wide bands: the rare very wide ones (k = 1) are issued first
      const Slot::StripLaunch& L = s.strip[k];
      if (L.grid <= 0) return UBU_SW_OK;
      StripArgs sa;
      sa.lists = s.d_lists.p; sa.firsts = first; sa.counts = count; sa.work = s.d_ctr.p + CTR_STRIP_WORK + 2 * g + k; sa.n_slots = s.n_slots;
      sa.cls_lo = L.cls_lo; sa.cls_hi = L.cls_hi;
      sa.hstore = s.d_strip.p + L.h_off; sa.boundary = reinterpret_cast<uint2*>(s.d_strip.p + L.b_off);
      sa.rows_cap = s.caps.rows_cap; sa.bw_cap = L.bw_cap; sa.h_stride = L.h_stride; sa.b_stride = L.b_stride;
      CU(h, launch_strip(s.any_n, def, s.strip_bytes, std::max(1, L.grid / gdiv), a, sa, k == 0 ? ks[K_S0] : ks[K_S1]));
      ++s.n_launches;
      return UBU_SW_OK;
    };
    auto issue_reg = [&]() -> int {
      auto launch3 = [&](auto k8, auto k16, auto k32) -> int {
        if (sm32 > 48 * 1024) {
          CU(h, cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
          CU(h, cudaFuncSetAttribute(k16, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
          CU(h, cudaFuncSetAttribute(k32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
        }
        if (sched != 7) { k8<<<grid, TB_THREADS, sm8, ks[K_8]>>>(a, s.d_lists.p + 1 * NS, first + 1, count + 1, h8, max_rows); CU(h, cudaGetLastError()); }
        if (sched != 7) { k16<<<grid, TB_THREADS, sm16, ks[K_16]>>>(a, s.d_lists.p + 2 * NS, first + 2, count + 2, h16, max_rows); CU(h, cudaGetLastError()); }
        k32<<<grid, TB_THREADS, sm32, ks[K_32]>>>(a, s.d_lists.p + 3 * NS, first + 3, count + 3, h32, max_rows); CU(h, cudaGetLastError());
        if (sched == 7) { k16<<<grid, TB_THREADS, sm16, ks[K_16]>>>(a, s.d_lists.p + 2 * NS, first + 2, count + 2, h16, max_rows); CU(h, cudaGetLastError()); }
        if (sched == 7) { k8<<<grid, TB_THREADS, sm8, ks[K_8]>>>(a, s.d_lists.p + 1 * NS, first + 1, count + 1, h8, max_rows); CU(h, cudaGetLastError()); }
        return UBU_SW_OK;
      };
      int rc3;
      if (s.any_n && hbyte) rc3 = launch3(tb_fill_reg_kernel<8, true, TB_THREADS, 1>, tb_fill_reg_kernel<16, true, TB_THREADS, 1>, tb_fill_reg_kernel<32, true, TB_THREADS, 1>);
      else if (s.any_n) rc3 = launch3(tb_fill_reg_kernel<8, true, TB_THREADS, 2>, tb_fill_reg_kernel<16, true, TB_THREADS, 2>, tb_fill_reg_kernel<32, true, TB_THREADS, 2>);
      else if (hbyte) rc3 = launch3(tb_fill_reg_kernel<8, false, TB_THREADS, 1>, tb_fill_reg_kernel<16, false, TB_THREADS, 1>, tb_fill_reg_kernel<32, false, TB_THREADS, 1>);
      else rc3 = launch3(tb_fill_reg_kernel<8, false, TB_THREADS, 2>, tb_fill_reg_kernel<16, false, TB_THREADS, 2>, tb_fill_reg_kernel<32, false, TB_THREADS, 2>);
      if (rc3) return rc3;
      s.n_launches += 3;
      return UBU_SW_OK;
    };
    auto issue_diag = [&]() -> int {
      if (s.any_n) tb_diag_kernel<true><<<h->sm_count * 8, 128, 0, ks[K_DIAG]>>>(a, s.d_lists.p, first, count);
      else tb_diag_kernel<false><<<h->sm_count * 8, 128, 0, ks[K_DIAG]>>>(a, s.d_lists.p, first, count);
      CU(h, cudaGetLastError()); ++s.n_launches;
      return UBU_SW_OK;
    };
    // issue order (it is the execution order within a stream, and earlier launches get their CTAs placed first). Side-by-side form:
    // diagonal, band 8, 16, 32, strips -- the streamed chunks of the end-to-end path lose 2.6 % with the heavy kernels issued first.
    // Two-stream form: wide strips, common strips, band 32, 16, 8, diagonal.
    {
      int rc = UBU_SW_OK;
      if (sched == 7) { rc = issue_strip(1); if (!rc) rc = issue_strip(0); if (!rc) rc = issue_reg(); if (!rc) rc = issue_diag(); }
      else { rc = issue_diag(); if (!rc) rc = issue_reg(); if (!rc) rc = issue_strip(1); if (!rc) rc = issue_strip(0); }
      if (rc) return rc;
    }
    tb_wide_kernel<<<s.tbw_grid, TBW_THREADS, 0, ks[K_W]>>>(a, s.d_lists.p + (size_t)TB_CLASS_SCALAR * NS, first + TB_CLASS_SCALAR, count + TB_CLASS_SCALAR,
                                                            s.d_wide.p, max_rows, s.bw_cap);
    CU(h, cudaGetLastError()); ++s.n_launches;
  }
  if (s.long_grid > 0) {                              // contig-sized queries (L > 256): forward, direction matrix and walk in one kernel
    LongArgs la;
    la.list = s.d_order.p + s.n_slots; la.n = (uint32_t)s.long_tasks.size(); la.work = s.d_ctr.p + CTR_WORK + 2;
    la.scratch = s.d_long.p; la.per_warp = s.long_per_warp;
    if (s.any_n) sw_long_kernel<true><<<s.long_grid, LONG_THREADS, 0, ks[K_LONG]>>>(a, la);
    else sw_long_kernel<false><<<s.long_grid, LONG_THREADS, 0, ks[K_LONG]>>>(a, la);
    CU(h, cudaGetLastError()); ++s.n_launches;
  }
  for (int k = 0; k < kAux; ++k) {                    // join
    CU(h, cudaEventRecord(s.ev_join[k], sx[k]));
    CU(h, cudaStreamWaitEvent(st, s.ev_join[k], 0));
  }
  CU(h, cudaEventRecord(s.ev[2], st));
  s.ran = true;
  return UBU_SW_OK;
}

// submit(): the device->host copies that do not depend on the CIGAR count are queued right behind the kernels
// (records, counters, and a first slice of the pool sized for ~2 ops per task), so wait() usually needs one sync.
int enqueue_early_d2h(ubu_sw_handle* h, Slot& s, ubu_sw_result* out, uint32_t* pool, size_t pool_cap) {
  s.pre_words = 0;
  if (s.n_tasks == 0 || !out) return UBU_SW_OK;
  cudaStream_t st = s.stream;
  CU(h, cudaMemcpyAsync(s.h_ctr.p, s.d_ctr.p, 16 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CU(h, cudaMemcpyAsync(out, s.d_res.p, sizeof(ubu_sw_result) * s.n_tasks, cudaMemcpyDeviceToHost, st));
  const size_t guess = std::min<size_t>(std::min<size_t>(pool_cap, s.cig_cap), (size_t)2 * s.n_tasks + 64);
  if (pool && guess) { CU(h, cudaMemcpyAsync(pool, s.d_cig.p, sizeof(uint32_t) * guess, cudaMemcpyDeviceToHost, st)); s.pre_words = guess; }
  s.pre_queued = true;
  return UBU_SW_OK;
}

int do_fetch(ubu_sw_handle* h, Slot& s, ubu_sw_result* out, uint32_t* pool, size_t pool_cap, size_t* used) {
  if (!s.ran) return UBU_SW_ERR_ARG;
  if (used) *used = 0;
  if (s.n_tasks == 0) return UBU_SW_OK;
  if (!out) return UBU_SW_ERR_ARG;
  cudaStream_t st = s.stream;
  if (s.pre_queued) {                  // fast path of wait(): everything may already be on its way
    s.pre_queued = false;
    CU(h, cudaStreamSynchronize(st));
    const unsigned int need = s.h_ctr.p[CTR_CIG], err = s.h_ctr.p[CTR_ERR];
    if (!err && need <= pool_cap && (need == 0 || pool)) {
      if (need > s.pre_words) {
        CU(h, cudaMemcpyAsync(pool + s.pre_words, s.d_cig.p + s.pre_words, sizeof(uint32_t) * (need - s.pre_words), cudaMemcpyDeviceToHost, st));
        CU(h, cudaStreamSynchronize(st));
      }
      if (used) *used = need;
      return UBU_SW_OK;
    }
  }
  for (int attempt = 0;; ++attempt) {
    CU(h, cudaMemcpyAsync(s.h_ctr.p, s.d_ctr.p, 16 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    const unsigned int need = s.h_ctr.p[CTR_CIG], err = s.h_ctr.p[CTR_ERR];
    if (err & 2u) { h->last_error = "internal: traceback left its band (please report)"; return UBU_SW_ERR_CUDA; }
    if (err & 1u) {                    // device CIGAR pool was too small: grow it and redo the device work once
      if (attempt) { h->last_error = "internal: CIGAR pool overflow after regrow"; return UBU_SW_ERR_CUDA; }
      s.cig_cap = need + 4096u;
      CU(h, s.d_cig.reserve(s.cig_cap));
      int rc = do_run(h, s); if (rc) return rc;
      continue;
    }
    if (used) *used = need;
    if (need > pool_cap || (need && !pool)) return UBU_SW_ERR_CIGAR_POOL;
    CU(h, cudaMemcpyAsync(out, s.d_res.p, sizeof(ubu_sw_result) * s.n_tasks, cudaMemcpyDeviceToHost, st));
    if (need) CU(h, cudaMemcpyAsync(pool, s.d_cig.p, sizeof(uint32_t) * need, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return UBU_SW_OK;
  }
}

}  // namespace

// ================================= C ABI ========================================================
extern "C" {

int ubu_sw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

unsigned ubu_sw_version(void) { return UBU_SW_VERSION; }

const char* ubu_sw_strerror(int status) {
  switch (status) {
    case UBU_SW_OK: return "ok";
    case UBU_SW_ERR_ARG: return "invalid argument";
    case UBU_SW_ERR_CUDA: return "CUDA failure";
    case UBU_SW_ERR_CIGAR_POOL: return "CIGAR pool too small";
    case UBU_SW_ERR_TOO_LONG: return "read or window too long for this build";
    case UBU_SW_ERR_NO_DEVICE: return "no usable CUDA device (sm_100 required)";
    case UBU_SW_ERR_BUSY: return "no free pipeline slot / ticket not pending";
    case UBU_SW_ERR_NOMEM: return "out of host memory";
    case UBU_SW_ERR_CAPACITY: return "output table / node bound too small";
    default: return "unknown status";
  }
}

const char* ubu_sw_last_error(ubu_sw_handle* h) { return h ? h->last_error.c_str() : ""; }

int ubu_sw_create(const ubu_sw_params* params, int device, int slots, ubu_sw_handle** out) {
  if (!out) return UBU_SW_ERR_ARG;
  *out = nullptr;
  ubu_sw_params p = params ? *params : ubu_sw_params{1, -3, 5, 2};
  if (!params_ok(p)) return UBU_SW_ERR_ARG;
  if (slots == 0) slots = 4;
  if (slots < 1 || slots > kMaxSlots) return UBU_SW_ERR_ARG;
  int ndev = ubu_sw_device_count();
  if (device < 0 || device >= ndev) return UBU_SW_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return UBU_SW_ERR_NO_DEVICE; }
  if (prop.major != 10) return UBU_SW_ERR_NO_DEVICE;            // sm_100a code only: no other code path exists
  ubu_sw_handle* h = new (std::nothrow) ubu_sw_handle();
  if (!h) return UBU_SW_ERR_NOMEM;
  h->device = device; h->n_slots = slots; h->sm_count = prop.multiProcessorCount;
  h->sc = make_scoring(p.match, p.mismatch, p.gap_open, p.gap_ext);
  if (cudaSetDevice(device) != cudaSuccess) { delete h; return UBU_SW_ERR_CUDA; }
  {
    int prio_lo = 0, prio_hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess) { cudaGetLastError(); prio_lo = prio_hi = 0; }
    h->prio_hi = env_off("UBU_TB_NOPRIO") ? prio_lo : prio_hi;  // development switch
  }
  for (int i = 0; i < slots; ++i) {
    if (cudaStreamCreateWithFlags(&h->slots[i].stream, cudaStreamNonBlocking) != cudaSuccess) { ubu_sw_destroy(h); return UBU_SW_ERR_CUDA; }
    for (int e = 0; e < 3; ++e)
      if (cudaEventCreate(&h->slots[i].ev[e]) != cudaSuccess) { ubu_sw_destroy(h); return UBU_SW_ERR_CUDA; }
    if (cudaEventCreateWithFlags(&h->slots[i].ev_fork, cudaEventDisableTiming) != cudaSuccess) { ubu_sw_destroy(h); return UBU_SW_ERR_CUDA; }
    for (int e = 0; e < kAux; ++e) {
      if (cudaStreamCreateWithFlags(&h->slots[i].aux[e], cudaStreamNonBlocking) != cudaSuccess) { ubu_sw_destroy(h); return UBU_SW_ERR_CUDA; }
      if (cudaEventCreateWithFlags(&h->slots[i].ev_join[e], cudaEventDisableTiming) != cudaSuccess) { ubu_sw_destroy(h); return UBU_SW_ERR_CUDA; }
    }
  }
  *out = h;
  return UBU_SW_OK;
}

void ubu_sw_destroy(ubu_sw_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (int i = 0; i < kMaxSlots; ++i) {
    Slot& s = h->slots[i];
    if (s.stream) cudaStreamSynchronize(s.stream);
    s.d_rp.release(); s.d_rn.release(); s.d_wp.release(); s.d_wn.release();
    s.d_roff.release(); s.d_rnoff.release(); s.d_woff.release(); s.d_wnoff.release(); s.d_idx.release();
    s.d_order.release(); s.d_lists.release(); s.d_cig.release(); s.d_rlen.release(); s.d_wlen.release();
    s.d_fwd.release(); s.d_res.release(); s.d_ctr.release(); s.d_hstore.release(); s.d_strip.release(); s.d_long.release(); s.d_wide.release(); s.d_gsel.release();
    s.h_order.release(); s.h_ctr.release();
    s.h_pool.release();
    if (i == 0) { h->g_packed.release(); h->g_nmask.release(); h->p_bytes.release(); h->g_off.release(); h->g_noff.release(); h->g_len.release(); h->p_words.release(); }
    for (int e = 0; e < 3; ++e) if (s.ev[e]) cudaEventDestroy(s.ev[e]);
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    for (int e = 0; e < kAux; ++e) { if (s.ev_join[e]) cudaEventDestroy(s.ev_join[e]); if (s.aux[e]) cudaStreamDestroy(s.aux[e]); if (s.hi[e]) cudaStreamDestroy(s.hi[e]); }
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  delete h;
}

int ubu_sw_stage(ubu_sw_handle* h, int slot, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx) {
  if (!h || slot < 0 || slot >= h->n_slots) return UBU_SW_ERR_ARG;
  if (h->slots[slot].pending) return UBU_SW_ERR_BUSY;
  CU(h, cudaSetDevice(h->device));
  return do_stage(h, h->slots[slot], reads, windows, pair_ref_idx);
}

int ubu_sw_run(ubu_sw_handle* h, int slot) {
  if (!h || slot < 0 || slot >= h->n_slots) return UBU_SW_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  return do_run(h, h->slots[slot]);
}

int ubu_sw_sync(ubu_sw_handle* h, int slot) {
  if (!h || slot < 0 || slot >= h->n_slots) return UBU_SW_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->slots[slot].stream));
  return UBU_SW_OK;
}

int ubu_sw_fetch(ubu_sw_handle* h, int slot, ubu_sw_result* out, uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used) {
  if (!h || slot < 0 || slot >= h->n_slots) return UBU_SW_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  return do_fetch(h, h->slots[slot], out, cigar_pool, cigar_pool_cap, cigar_pool_used);
}

int ubu_sw_last_run_ms(ubu_sw_handle* h, int slot, float ms[3]) {
  if (!h || slot < 0 || slot >= h->n_slots || !ms) return UBU_SW_ERR_ARG;
  Slot& s = h->slots[slot];
  if (!s.ran) return UBU_SW_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaEventSynchronize(s.ev[2]));
  CU(h, cudaEventElapsedTime(&ms[0], s.ev[0], s.ev[2]));
  CU(h, cudaEventElapsedTime(&ms[1], s.ev[0], s.ev[1]));
  CU(h, cudaEventElapsedTime(&ms[2], s.ev[1], s.ev[2]));
  return UBU_SW_OK;
}

// Diagnostics: the device counters of the last run() on the slot, copied out after a sync: [0..TB_CLASSES) tasks per traceback
// class (0 gap-free, 1-3 register bands, 4-8 strip classes, 9 scalar), then CIGAR ops produced, error flags, strip work units.
int ubu_sw_debug_counters(ubu_sw_handle* h, int slot, uint32_t out[16]) {
  if (!h || slot < 0 || slot >= h->n_slots || !out) return UBU_SW_ERR_ARG;
  Slot& s = h->slots[slot];
  if (!s.ran || !s.d_ctr.p) return UBU_SW_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(s.stream));
  CU(h, cudaMemcpy(out, s.d_ctr.p, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return UBU_SW_OK;
}

int ubu_sw_last_run_launches(ubu_sw_handle* h, int slot) {
  if (!h || slot < 0 || slot >= h->n_slots) return UBU_SW_ERR_ARG;
  return h->slots[slot].n_launches;
}

int ubu_sw_submit(ubu_sw_handle* h, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx,
                  ubu_sw_result* out, uint32_t* cigar_pool, size_t cigar_pool_cap, int* ticket) {
  if (!h || !ticket) return UBU_SW_ERR_ARG;
  int slot = -1;
  for (int k = 0; k < h->n_slots; ++k) {
    const int c = (h->next_slot + k) % h->n_slots;
    if (!h->slots[c].pending) { slot = c; break; }
  }
  if (slot < 0) return UBU_SW_ERR_BUSY;
  CU(h, cudaSetDevice(h->device));
  Slot& s = h->slots[slot];
  int rc = do_stage(h, s, reads, windows, pair_ref_idx); if (rc) return rc;
  rc = do_run(h, s); if (rc) return rc;
  rc = enqueue_early_d2h(h, s, out, cigar_pool, cigar_pool_cap); if (rc) return rc;
  s.out = out; s.pool = cigar_pool; s.pool_cap = cigar_pool_cap; s.pending = true;
  h->next_slot = (slot + 1) % h->n_slots;
  *ticket = slot;
  return UBU_SW_OK;
}

int ubu_sw_wait(ubu_sw_handle* h, int ticket, size_t* cigar_pool_used) {
  if (!h || ticket < 0 || ticket >= h->n_slots) return UBU_SW_ERR_ARG;
  Slot& s = h->slots[ticket];
  if (!s.pending) return UBU_SW_ERR_BUSY;
  CU(h, cudaSetDevice(h->device));
  s.pending = false;
  return do_fetch(h, s, s.out, s.pool, s.pool_cap, cigar_pool_used);
}

// A large batch is cut into chunks that travel through the handle's slots (upload of chunk c+1, kernels of chunk c and download
// of chunk c-1 overlap); records land in out[] directly, every chunk's CIGAR ops come back through the slot's pinned staging
// and are appended to the caller's pool in input order with the records' offsets rebased. Needs what every packer produces:
// sequences laid out in ascending, non-overlapping order. Returns UBU_SW_ERR_BUSY_FALLBACK (internal) when the batch does not
// qualify and the one-shot path should run instead.
namespace {
constexpr int kNoStream = 1;                               // internal: not a streamable batch
constexpr uint32_t kStreamMinReads = 1u << 17, kStreamChunkMin = 1u << 15, kStreamChunkMax = 1u << 17;

int align_stream(ubu_sw_handle* h, const ubu_sw_seqs* reads, const ubu_sw_seqs* wins, const uint32_t* idx, ubu_sw_result* out,
                 uint32_t* pool, size_t pool_cap, size_t* used_out) {
  const uint32_t n = reads->count;
  if (h->n_slots < 2 || n < kStreamMinReads || !out) return kNoStream;
  for (int k = 0; k < h->n_slots; ++k) if (h->slots[k].pending) return kNoStream;
  if (validate_seqs(reads) || validate_seqs(wins) || (!idx && reads->count != wins->count)) return kNoStream;   // the one-shot path reports it
  uint32_t chunk = std::min(kStreamChunkMax, std::max(kStreamChunkMin, n / (uint32_t)(2 * h->n_slots)));
  {
    // Mixed read lengths (judged on the first reads): a chunk is a dozen or more forward launches, and from 256k reads on its traceback
    // runs in phases underneath them (do_run): larger chunks (cfg4, 4M reads from pinned host buffers: 110 -> 102 ms).
    uint16_t lmin = 0xFFFF, lmax = 0;
    minmax_u16(reads->len, std::min<uint32_t>(n, 1u << 16), lmin, lmax);
    if (cfg_of_len(lmin) != cfg_of_len(lmax)) chunk = std::min(1u << 18, std::max(1u << 16, n / (uint32_t)h->n_slots));
  }
  if (const char* v = getenv("UBU_STREAM_CHUNK")) if (*v) chunk = std::max<uint32_t>(1024u, (uint32_t)atol(v));   // development switch
  // What streaming needs is checked chunk by chunk, right before a chunk is staged (a pass over the whole batch up front
  // would sit in front of the first upload): its reads and the windows they use are laid out in ascending order, so that
  // one stretch of bytes holds them all, and the windows form a compact stretch (reads grouped by window, as
  // coordinate-sorted input is). A chunk that does not qualify sends the whole batch to the one-shot path, which has no
  // such needs. Nothing is rewritten: a chunk keeps the caller's absolute offsets and window indices (StageBias).
  auto stretch = [](const uint32_t* off, const uint16_t* len, uint32_t m, uint32_t per, uint64_t limit, size_t& b0, size_t& bytes) -> bool {
    const uint64_t end = (uint64_t)off[m - 1] + ((uint32_t)len[m - 1] + per - 1u) / per;
    b0 = off[0]; bytes = (size_t)(end - off[0]);
    return end <= limit && ascending(off, len, m, per);
  };
  CU(h, cudaSetDevice(h->device));
  auto fail = [&](int rc) {                                // nothing of this call may still be writing into the caller's buffers
    for (int k = 0; k < h->n_slots; ++k) if (h->slots[k].stream) cudaStreamSynchronize(h->slots[k].stream);
    return rc;
  };
  size_t base = 0, total = 0;
  bool overflow = false;
  int head = 0, inflight = 0;
  const bool trace = env_off("UBU_STREAM_TRACE");
  double t_prep = 0, t_stage = 0, t_run = 0, t_d2h = 0, t_fin = 0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };                              // slots are used round-robin: head = oldest in flight

  auto finish = [&](Slot& s) -> int {
    size_t used = 0;
    const size_t cap = s.h_pool.cap;
    int rc = do_fetch(h, s, out + s.st_lo, s.h_pool.p, cap, &used);
    if (rc == UBU_SW_ERR_CIGAR_POOL) {                     // the 6-ops-per-read guess was too small for this chunk
      CU(h, s.h_pool.reserve(used + 64));
      rc = do_fetch(h, s, out + s.st_lo, s.h_pool.p, s.h_pool.cap, &used);
    }
    if (rc) return rc;
    total += used;
    if (base + used > pool_cap || (used && !pool)) overflow = true;
    else {
      if (used) memcpy(pool + base, s.h_pool.p, used * sizeof(uint32_t));
      if (base) for (uint32_t t = s.st_lo; t < s.st_hi; ++t) out[t].cigar_off += (uint32_t)base;
      base += used;
    }
    return UBU_SW_OK;
  };

  // chunk borders: quarter- and half-size chunks at both ends shorten the pipeline's fill (nothing overlaps the first upload)
  // and drain (nothing overlaps the last traceback and download)
  std::vector<uint32_t> cuts{0u};
  {
    const uint32_t ramp[3] = {chunk / 4u, chunk / 4u, chunk / 2u};
    uint32_t head_end = 0, tail = 0;
    for (uint32_t r : ramp) { head_end += r; tail += r; }
    if ((uint64_t)head_end + tail + chunk <= n && !env_off("UBU_STREAM_NORAMP")) {
      for (uint32_t r : ramp) cuts.push_back(cuts.back() + r);
      while ((uint64_t)cuts.back() + chunk + tail <= n) cuts.push_back(cuts.back() + chunk);
      const uint32_t rest = n - tail - cuts.back();
      if (rest) cuts.push_back(cuts.back() + rest);
      for (int r = 2; r >= 0; --r) cuts.push_back(cuts.back() + ramp[r]);
    } else {
      while (cuts.back() < n) cuts.push_back(std::min<uint64_t>(n, (uint64_t)cuts.back() + chunk));
    }
  }
  for (size_t ci = 0; ci + 1 < cuts.size(); ++ci) {
    const uint32_t lo = cuts[ci], hi = cuts[ci + 1], m = hi - lo;
    double t0 = now();
    if (inflight == h->n_slots) { int rc = finish(h->slots[head]); if (rc) return fail(rc); head = (head + 1) % h->n_slots; --inflight; }
    t_fin += now() - t0; t0 = now();
    Slot& s = h->slots[(head + inflight) % h->n_slots];
    // ---- this chunk's views: reads [lo, hi) and the stretch of windows they use ----
    ubu_sw_seqs rs = *reads, ws = *wins;
    StageBias bias{};
    rs.off = reads->off + lo; rs.len = reads->len + lo; rs.count = m;
    if (!stretch(rs.off, rs.len, m, 4, reads->packed_bytes, bias.rp, bias.r_bytes)) return fail(kNoStream);
    if (reads->nmask) {
      rs.nmask_off = reads->nmask_off + lo;
      if (!stretch(rs.nmask_off, rs.len, m, 8, reads->nmask_bytes, bias.rn, bias.rn_bytes)) return fail(kNoStream);
    }
    uint32_t w0 = lo, w1 = hi;                             // windows [w0, w1)
    if (idx) {
      uint32_t mn = 0xFFFFFFFFu, mx = 0;
      minmax_u32(idx + lo, m, mn, mx);
      if (mx >= wins->count || (uint64_t)(mx - mn) + 1u > 2ull * m + 16ull) return fail(kNoStream);
      w0 = mn; w1 = mx + 1u;
      bias.w = w0;
    }
    const uint32_t mw = w1 - w0;
    ws.off = wins->off + w0; ws.len = wins->len + w0; ws.count = mw;
    if (!stretch(ws.off, ws.len, mw, 4, wins->packed_bytes, bias.wp, bias.w_bytes)) return fail(kNoStream);
    if (wins->nmask) {
      ws.nmask_off = wins->nmask_off + w0;
      if (!stretch(ws.nmask_off, ws.len, mw, 8, wins->nmask_bytes, bias.wn, bias.wn_bytes)) return fail(kNoStream);
    }
    const uint32_t* ix = idx ? idx + lo : nullptr;
    CU(h, s.h_pool.reserve(6u * (size_t)m + 4096u));
    t_prep += now() - t0; t0 = now();
    int rc = do_stage(h, s, &rs, &ws, ix, true, &bias); if (rc) return fail(rc);
    t_stage += now() - t0; t0 = now();
    rc = do_run(h, s); if (rc) return fail(rc);
    t_run += now() - t0; t0 = now();
    rc = enqueue_early_d2h(h, s, out + lo, s.h_pool.p, s.h_pool.cap); if (rc) return fail(rc);
    t_d2h += now() - t0;
    s.st_lo = lo; s.st_hi = hi;
    ++inflight;
  }
  { const double t0 = now();
    while (inflight) { int rc = finish(h->slots[head]); if (rc) return fail(rc); head = (head + 1) % h->n_slots; --inflight; }
    t_fin += now() - t0; }
  if (trace) fprintf(stderr, "[ubu_sw stream] %zu chunks: prep %.2f ms, stage %.2f, run %.2f, d2h-enqueue %.2f, finish (incl. waiting) %.2f\n", cuts.size() - 1,
                     t_prep * 1e3, t_stage * 1e3, t_run * 1e3, t_d2h * 1e3, t_fin * 1e3);
  if (used_out) *used_out = overflow ? total : base;
  return overflow ? UBU_SW_ERR_CIGAR_POOL : UBU_SW_OK;
}
}  // namespace

int ubu_sw_align_batch(ubu_sw_handle* h, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx,
                       ubu_sw_result* out, uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used) {
  if (h && reads && windows && !env_off("UBU_NO_STREAM")) {
    const int rc = align_stream(h, reads, windows, pair_ref_idx, out, cigar_pool, cigar_pool_cap, cigar_pool_used);
    if (rc != kNoStream) return rc;
  }
  if (h) {                                // sequential synchronous calls reuse slot 0 (and its scratch) instead of rotating through the slots
    bool any_pending = false;
    for (int k = 0; k < h->n_slots; ++k) any_pending = any_pending || h->slots[k].pending;
    if (!any_pending) h->next_slot = 0;
  }
  int ticket = -1;
  int rc = ubu_sw_submit(h, reads, windows, pair_ref_idx, out, cigar_pool, cigar_pool_cap, &ticket);
  if (rc) return rc;
  return ubu_sw_wait(h, ticket, cigar_pool_used);
}

// ================================= multi-GPU driver ==============================================
// One process, N GPUs: a host thread and a handle per device; the batch is cut into input-order chunks,
// chunk c belongs to device c mod N and travels through that handle's slots. There is no merge pass on the host: every chunk
// owns a slice of the caller's CIGAR pool proportional to its reads, the kernels write final cigar_off values (TbArgs::cig_base)
// and both the records and the ops are copied by the GPU straight into their input-order places in the caller's arrays.
// "Ordered merge" is then only the order in which finished chunks are released to the SAM writer.
}  // extern "C"

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unistd.h>

size_t ubu_sam_block(const ubu_sw_seqs* reads, const ubu_sw_result* res, const uint32_t* pool, uint32_t lo, uint32_t hi,
                     const ubu_sw_sam_names* nm, std::vector<char>& out);       // ubu_pack.cu

struct ubu_sw_multi {
  std::vector<ubu_sw_handle*> hs;
  uint32_t chunk = 1u << 16;
  std::string last_error;
};

namespace {

// Views of reads [lo, hi) and of the windows they use. True: the chunk's bytes are ascending stretches (what every packer
// produces from coordinate-sorted input) and only those stretches go to the device (bias). False: the views name the
// caller's whole buffers (correct for any layout, but every chunk then uploads all of them).
bool chunk_views(const ubu_sw_seqs* reads, const ubu_sw_seqs* wins, const uint32_t* idx, uint32_t lo, uint32_t hi,
                 ubu_sw_seqs& rs, ubu_sw_seqs& ws, StageBias& bias, const uint32_t*& ix) {
  const uint32_t m = hi - lo;
  auto stretch = [](const uint32_t* off, const uint16_t* len, uint32_t k, uint32_t per, uint64_t limit, size_t& b0, size_t& bytes) -> bool {
    const uint64_t end = (uint64_t)off[k - 1] + ((uint32_t)len[k - 1] + per - 1u) / per;
    b0 = off[0]; bytes = (size_t)(end - off[0]);
    return end <= limit && ascending(off, len, k, per);
  };
  rs = *reads; ws = *wins; bias = StageBias{};
  rs.off = reads->off + lo; rs.len = reads->len + lo; rs.count = m;
  if (reads->nmask) rs.nmask_off = reads->nmask_off + lo;
  ix = idx ? idx + lo : nullptr;
  bool ok = stretch(rs.off, rs.len, m, 4, reads->packed_bytes, bias.rp, bias.r_bytes);
  if (ok && reads->nmask) ok = stretch(rs.nmask_off, rs.len, m, 8, reads->nmask_bytes, bias.rn, bias.rn_bytes);
  uint32_t w0 = lo, w1 = hi;
  if (ok && idx) {
    uint32_t mn = 0xFFFFFFFFu, mx = 0;
    minmax_u32(idx + lo, m, mn, mx);
    ok = mx < wins->count && (uint64_t)(mx - mn) + 1u <= 2ull * m + 16ull;
    w0 = mn; w1 = mx + 1u;
  }
  if (ok) {
    const uint32_t mw = w1 - w0;
    ws.off = wins->off + w0; ws.len = wins->len + w0; ws.count = mw;
    bias.w = idx ? w0 : 0u;
    ok = stretch(ws.off, ws.len, mw, 4, wins->packed_bytes, bias.wp, bias.w_bytes);
    if (ok && wins->nmask) { ws.nmask_off = wins->nmask_off + w0; ok = stretch(ws.nmask_off, ws.len, mw, 8, wins->nmask_bytes, bias.wn, bias.wn_bytes); }
  }
  if (ok) return true;
  rs.packed = reads->packed; ws = *wins; bias = StageBias{};            // whole buffers
  if (!idx) { ws.off = wins->off + lo; ws.len = wins->len + lo; ws.count = m; if (wins->nmask) ws.nmask_off = wins->nmask_off + lo; }
  return false;
}

struct MultiRun {
  ubu_sw_multi* m;
  const ubu_sw_seqs *reads, *wins; const uint32_t* idx;
  ubu_sw_result* out; uint32_t* pool; size_t pool_cap;
  uint32_t n, chunk, n_chunks;
  std::vector<std::atomic<int>> done;          // per chunk: 1 = records and ops are in the caller's arrays
  std::vector<uint32_t> need;                  // per chunk: CIGAR ops it produced
  std::atomic<int> failed{0};                  // first failing status
  std::mutex mu; std::condition_variable cv;
  MultiRun(uint32_t nc) : done(nc), need(nc, 0u) { for (auto& d : done) d.store(0); }
  uint64_t base_of(uint32_t c) const { return (uint64_t)pool_cap * std::min<uint64_t>((uint64_t)c * chunk, n) / n; }
  void fail(int rc) { int z = 0; failed.compare_exchange_strong(z, rc); cv.notify_all(); }
};

// the worker of one device: its chunks, in order, through the handle's slots
void multi_worker(MultiRun* R, int d) {
  ubu_sw_handle* h = R->m->hs[d];
  const int nd = (int)R->m->hs.size();
  if (cudaSetDevice(h->device) != cudaSuccess) { R->fail(UBU_SW_ERR_CUDA); return; }
  int head = 0, inflight = 0;
  uint32_t chunk_of_slot[kMaxSlots] = {};
  auto finish = [&](int slot) -> int {
    Slot& s = h->slots[slot];
    const uint32_t c = chunk_of_slot[slot];
    const uint64_t base = R->base_of(c), cap = R->base_of(c + 1) - base;
    size_t used = 0;
    int rc = do_fetch(h, s, R->out + s.st_lo, R->pool ? R->pool + base : nullptr, (size_t)cap, &used);
    R->need[c] = (uint32_t)used;
    if (rc && rc != UBU_SW_ERR_CIGAR_POOL) return rc;
    if (rc == UBU_SW_ERR_CIGAR_POOL) R->fail(rc);                       // the slice was too small: the call reports the capacity it needs
    R->done[c].store(1, std::memory_order_release);
    { std::lock_guard<std::mutex> lk(R->mu); }
    R->cv.notify_all();
    return UBU_SW_OK;
  };
  int rc = UBU_SW_OK;
  for (uint32_t c = (uint32_t)d; c < R->n_chunks && !rc; c += (uint32_t)nd) {
    if (R->failed.load() && R->failed.load() != UBU_SW_ERR_CIGAR_POOL) break;
    if (inflight == h->n_slots) { rc = finish(head); head = (head + 1) % h->n_slots; --inflight; if (rc) break; }
    const int slot = (head + inflight) % h->n_slots;
    Slot& s = h->slots[slot];
    const uint32_t lo = c * R->chunk, hi = (uint32_t)std::min<uint64_t>(R->n, (uint64_t)lo + R->chunk);
    ubu_sw_seqs rs, ws; StageBias bias; const uint32_t* ix;
    const bool compact = chunk_views(R->reads, R->wins, R->idx, lo, hi, rs, ws, bias, ix);
    rc = do_stage(h, s, &rs, &ws, ix, compact, compact ? &bias : nullptr); if (rc) break;
    const uint64_t base = R->base_of(c), cap = R->base_of(c + 1) - base;
    s.cig_base = (unsigned int)base;
    rc = do_run(h, s); if (rc) break;
    rc = enqueue_early_d2h(h, s, R->out + lo, R->pool ? R->pool + base : nullptr, (size_t)cap); if (rc) break;
    s.st_lo = lo; s.st_hi = hi; chunk_of_slot[slot] = c;
    ++inflight;
  }
  while (inflight) { const int r2 = finish(head); if (r2 && !rc) rc = r2; head = (head + 1) % h->n_slots; --inflight; }
  if (rc) {
    for (int k = 0; k < h->n_slots; ++k) if (h->slots[k].stream) cudaStreamSynchronize(h->slots[k].stream);
    R->m->last_error = h->last_error;
    R->fail(rc);
  }
}

int multi_align(ubu_sw_multi* m, const ubu_sw_seqs* reads, const ubu_sw_seqs* wins, const uint32_t* idx, ubu_sw_result* out,
                uint32_t* pool, size_t pool_cap, size_t* used_out, int fd, const ubu_sw_sam_names* names, int sam_threads,
                uint64_t* bytes_out, bool want_sam) {
  if (used_out) *used_out = 0;
  if (bytes_out) *bytes_out = 0;
  if (!m || m->hs.empty() || !reads || !wins) return UBU_SW_ERR_ARG;
  int rc = validate_seqs(reads); if (rc) return rc;
  rc = validate_seqs(wins); if (rc) return rc;
  if (!idx && reads->count != wins->count) return UBU_SW_ERR_ARG;
  const uint32_t n = reads->count;
  if (n == 0) return UBU_SW_OK;
  if (!out || (want_sam && !names)) return UBU_SW_ERR_ARG;
  for (ubu_sw_handle* h : m->hs) for (int k = 0; k < h->n_slots; ++k) if (h->slots[k].pending) return UBU_SW_ERR_BUSY;
  pool_cap = std::min<size_t>(pool_cap, 0xFFFFFFFFull);                      // cigar_off is 32 bits
  const uint32_t n_chunks = (uint32_t)(((uint64_t)n + m->chunk - 1) / m->chunk);
  MultiRun R(n_chunks);
  R.m = m; R.reads = reads; R.wins = wins; R.idx = idx; R.out = out; R.pool = pool; R.pool_cap = pool ? pool_cap : 0;
  R.n = n; R.chunk = m->chunk; R.n_chunks = n_chunks;
  std::vector<std::thread> workers;
  try {
    for (size_t d = 0; d < m->hs.size(); ++d) workers.emplace_back(multi_worker, &R, (int)d);
  } catch (...) { R.fail(UBU_SW_ERR_NOMEM); }

  // ---- ordered release of finished chunks: SAM text formatted on helper threads, written in input order ----
  uint64_t total_bytes = 0;
  if (want_sam) {
    unsigned nt = sam_threads > 0 ? (unsigned)sam_threads : std::max(2u, std::thread::hardware_concurrency());
    nt = std::max(1u, std::min<unsigned>(nt, n_chunks));
    const unsigned ring = nt + 2;
    std::vector<std::vector<char>> bufs(ring);
    std::vector<size_t> blen(ring, 0);
    std::vector<std::atomic<int>> formatted(n_chunks);
    for (auto& f : formatted) f.store(0);
    std::atomic<uint32_t> next_fmt{0}, written{0};
    std::vector<std::thread> fmt;
    auto formatter = [&] {
      for (;;) {
        const uint32_t c = next_fmt.fetch_add(1);
        if (c >= n_chunks) return;
        {   // wait for the chunk's results and for its ring buffer to be free
          std::unique_lock<std::mutex> lk(R.mu);
          R.cv.wait(lk, [&] { return (R.failed.load() != 0) || (R.done[c].load(std::memory_order_acquire) && c < written.load() + ring); });
          if (R.failed.load()) return;
        }
        const uint32_t lo = c * R.chunk, hi = (uint32_t)std::min<uint64_t>(n, (uint64_t)lo + R.chunk);
        blen[c % ring] = ubu_sam_block(reads, out, pool, lo, hi, names, bufs[c % ring]);
        if (blen[c % ring] == (size_t)-1) { blen[c % ring] = 0; R.fail(UBU_SW_ERR_ARG); return; }
        formatted[c].store(1, std::memory_order_release);
        { std::lock_guard<std::mutex> lk(R.mu); }
        R.cv.notify_all();
      }
    };
    try { for (unsigned t = 0; t < nt; ++t) fmt.emplace_back(formatter); } catch (...) { R.fail(UBU_SW_ERR_NOMEM); }
    for (uint32_t c = 0; c < n_chunks; ++c) {
      {
        std::unique_lock<std::mutex> lk(R.mu);
        R.cv.wait(lk, [&] { return R.failed.load() != 0 || formatted[c].load(std::memory_order_acquire) != 0; });
        if (R.failed.load()) break;
      }
      const std::vector<char>& b = bufs[c % ring];
      size_t off = 0;
      while (fd >= 0 && off < blen[c % ring]) {
        const ssize_t wr = write(fd, b.data() + off, blen[c % ring] - off);
        if (wr <= 0) { R.fail(UBU_SW_ERR_ARG); break; }
        off += (size_t)wr;
      }
      total_bytes += blen[c % ring];
      written.store(c + 1);
      { std::lock_guard<std::mutex> lk(R.mu); }
      R.cv.notify_all();
    }
    for (auto& t : fmt) t.join();
  }
  for (auto& t : workers) t.join();
  uint64_t high = 0; double per_read = 0.0;
  for (uint32_t c = 0; c < n_chunks; ++c) {
    const uint32_t lo = c * R.chunk, hi = (uint32_t)std::min<uint64_t>(n, (uint64_t)lo + R.chunk);
    high = std::max<uint64_t>(high, R.base_of(c) + R.need[c]);
    per_read = std::max(per_read, (double)R.need[c] / (double)(hi - lo));
  }
  const int failed = R.failed.load();
  if (failed == UBU_SW_ERR_CIGAR_POOL) {                                     // capacity with which every slice would have been large enough
    if (used_out) *used_out = (size_t)(per_read * (double)n) + n_chunks + 64;
    return failed;
  }
  if (failed) return failed;
  if (used_out) *used_out = (size_t)high;
  if (bytes_out) *bytes_out = total_bytes;
  return UBU_SW_OK;
}

}  // namespace

extern "C" {

int ubu_sw_multi_create(const ubu_sw_params* params, const int* devices, int n_devices, int slots, uint32_t chunk_reads, ubu_sw_multi** out) {
  if (!out) return UBU_SW_ERR_ARG;
  *out = nullptr;
  const int visible = ubu_sw_device_count();
  if (visible <= 0) return UBU_SW_ERR_NO_DEVICE;
  if (!devices) n_devices = n_devices > 0 ? std::min(n_devices, visible) : visible;
  if (n_devices <= 0) return UBU_SW_ERR_ARG;
  ubu_sw_multi* m = new (std::nothrow) ubu_sw_multi();
  if (!m) return UBU_SW_ERR_NOMEM;
  if (chunk_reads) m->chunk = std::max<uint32_t>(chunk_reads, 256u);
  for (int k = 0; k < n_devices; ++k) {
    ubu_sw_handle* h = nullptr;
    const int rc = ubu_sw_create(params, devices ? devices[k] : k, slots, &h);
    if (rc) { ubu_sw_multi_destroy(m); return rc; }
    m->hs.push_back(h);
  }
  *out = m;
  return UBU_SW_OK;
}

void ubu_sw_multi_destroy(ubu_sw_multi* m) {
  if (!m) return;
  for (ubu_sw_handle* h : m->hs) ubu_sw_destroy(h);
  delete m;
}

int ubu_sw_multi_device_count(const ubu_sw_multi* m) { return m ? (int)m->hs.size() : 0; }
const char* ubu_sw_multi_last_error(ubu_sw_multi* m) { return m ? m->last_error.c_str() : ""; }

int ubu_sw_multi_align(ubu_sw_multi* m, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx,
                       ubu_sw_result* out, uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used) {
  return multi_align(m, reads, windows, pair_ref_idx, out, cigar_pool, cigar_pool_cap, cigar_pool_used, -1, nullptr, 0, nullptr, false);
}

int ubu_sw_multi_align_sam(ubu_sw_multi* m, const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx,
                           ubu_sw_result* out, uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used,
                           int fd, const ubu_sw_sam_names* names, int sam_threads, uint64_t* bytes_written) {
  return multi_align(m, reads, windows, pair_ref_idx, out, cigar_pool, cigar_pool_cap, cigar_pool_used, fd, names, sam_threads, bytes_written, true);
}

int ubu_sw_load_refs(ubu_sw_handle* h, const ubu_sw_refs* r) {
  if (!h || !r || (r->count && (!r->packed || !r->off || !r->len)) || (r->nmask && !r->nmask_off)) return UBU_SW_ERR_ARG;
  for (uint32_t k = 0; k < r->count; ++k) {
    if (r->off[k] + ((uint64_t)r->len[k] + 3) / 4 > r->packed_bytes) return UBU_SW_ERR_ARG;
    if (r->nmask && r->nmask_off[k] + ((uint64_t)r->len[k] + 7) / 8 > r->nmask_bytes) return UBU_SW_ERR_ARG;
  }
  CU(h, cudaSetDevice(h->device));
  cudaStream_t st = h->slots[0].stream;
  CU(h, h->g_packed.reserve(r->packed_bytes + 16)); CU(h, h->g_off.reserve(r->count)); CU(h, h->g_len.reserve(r->count));
  CU(h, cudaMemcpyAsync(h->g_packed.p, r->packed, r->packed_bytes, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->g_off.p, r->off, sizeof(uint64_t) * r->count, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(h->g_len.p, r->len, sizeof(uint32_t) * r->count, cudaMemcpyHostToDevice, st));
  h->g_has_n = r->nmask != nullptr;
  if (h->g_has_n) {
    CU(h, h->g_nmask.reserve(r->nmask_bytes + 16)); CU(h, h->g_noff.reserve(r->count));
    CU(h, cudaMemcpyAsync(h->g_nmask.p, r->nmask, r->nmask_bytes, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(h->g_noff.p, r->nmask_off, sizeof(uint64_t) * r->count, cudaMemcpyHostToDevice, st));
  }
  CU(h, cudaStreamSynchronize(st));
  h->g_count = r->count;
  return UBU_SW_OK;
}

int ubu_sw_recount_batch(ubu_sw_handle* h, const ubu_sw_seqs* reads, const ubu_sw_placed* placed, const uint32_t* cigar_pool,
                         size_t cigar_pool_n, ubu_sw_recount* out) {
  if (!h || !reads) return UBU_SW_ERR_ARG;
  int rc = validate_seqs(reads); if (rc) return rc;
  const uint32_t n = reads->count;
  if (n == 0) return UBU_SW_OK;
  if (!placed || !out || (cigar_pool_n && !cigar_pool)) return UBU_SW_ERR_ARG;
  for (uint32_t t = 0; t < n; ++t) {
    if ((uint64_t)reads->off[t] + ((uint32_t)reads->len[t] + 3u) / 4u > reads->packed_bytes) return UBU_SW_ERR_ARG;
    if (reads->nmask && (uint64_t)reads->nmask_off[t] + ((uint32_t)reads->len[t] + 7u) / 8u > reads->nmask_bytes) return UBU_SW_ERR_ARG;
    if ((uint64_t)placed[t].cigar_off + placed[t].cigar_n > cigar_pool_n) return UBU_SW_ERR_ARG;
  }
  CU(h, cudaSetDevice(h->device));
  Slot& s = h->slots[0];
  if (s.pending) return UBU_SW_ERR_BUSY;
  cudaStream_t st = s.stream;
  const bool rn = reads->nmask != nullptr;
  CU(h, s.d_rp.reserve(reads->packed_bytes + 16)); CU(h, s.d_roff.reserve(n)); CU(h, s.d_rlen.reserve(n));
  if (rn) { CU(h, s.d_rn.reserve(reads->nmask_bytes + 16)); CU(h, s.d_rnoff.reserve(n)); }
  const size_t b_placed = sizeof(ubu_sw_placed) * (size_t)n, b_out = sizeof(ubu_sw_recount) * (size_t)n;
  CU(h, h->p_bytes.reserve(b_placed + b_out + 64)); CU(h, h->p_words.reserve(cigar_pool_n + 1));
  ubu_sw_placed* d_placed = reinterpret_cast<ubu_sw_placed*>(h->p_bytes.p);
  ubu_sw_recount* d_out = reinterpret_cast<ubu_sw_recount*>(h->p_bytes.p + ((b_placed + 15) & ~(size_t)15));
  CU(h, cudaMemcpyAsync(s.d_rp.p, reads->packed, reads->packed_bytes, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(s.d_roff.p, reads->off, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(s.d_rlen.p, reads->len, sizeof(uint16_t) * n, cudaMemcpyHostToDevice, st));
  if (rn) {
    CU(h, cudaMemcpyAsync(s.d_rn.p, reads->nmask, reads->nmask_bytes, cudaMemcpyHostToDevice, st));
    CU(h, cudaMemcpyAsync(s.d_rnoff.p, reads->nmask_off, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  }
  CU(h, cudaMemcpyAsync(d_placed, placed, b_placed, cudaMemcpyHostToDevice, st));
  if (cigar_pool_n) CU(h, cudaMemcpyAsync(h->p_words.p, cigar_pool, sizeof(uint32_t) * cigar_pool_n, cudaMemcpyHostToDevice, st));
  DevSeqs dr{s.d_rp.p, s.d_roff.p, s.d_rlen.p, rn ? s.d_rn.p : nullptr, rn ? s.d_rnoff.p : nullptr};
  DevRefs dg{h->g_packed.p, h->g_off.p, h->g_len.p, h->g_has_n ? h->g_nmask.p : nullptr, h->g_has_n ? h->g_noff.p : nullptr};
  post_recount_kernel<<<(n + 127) / 128, 128, 0, st>>>(dr, dg, h->g_count, d_placed, h->p_words.p, n, d_out);
  CU(h, cudaGetLastError());
  CU(h, cudaMemcpyAsync(out, d_out, b_out, cudaMemcpyDeviceToHost, st));
  CU(h, cudaStreamSynchronize(st));
  return UBU_SW_OK;
}

int ubu_sw_project_batch(ubu_sw_handle* h, const ubu_sw_contig* contigs, uint32_t n_contigs, const uint32_t* contig_cigar_pool,
                         size_t contig_cigar_pool_n, const ubu_sw_proj_in* in, uint32_t n, ubu_sw_proj_out* out,
                         uint32_t* cigar_pool, size_t cigar_pool_cap, size_t* cigar_pool_used) {
  if (!h) return UBU_SW_ERR_ARG;
  if (cigar_pool_used) *cigar_pool_used = 0;
  if (n == 0) return UBU_SW_OK;
  if (!in || !out || (n_contigs && !contigs) || (contig_cigar_pool_n && !contig_cigar_pool) || (cigar_pool_cap && !cigar_pool)) return UBU_SW_ERR_ARG;
  for (uint32_t c = 0; c < n_contigs; ++c)
    if ((uint64_t)contigs[c].cigar_off + contigs[c].cigar_n > contig_cigar_pool_n) return UBU_SW_ERR_ARG;
  CU(h, cudaSetDevice(h->device));
  Slot& s = h->slots[0];
  if (s.pending) return UBU_SW_ERR_BUSY;
  cudaStream_t st = s.stream;
  const size_t b_c = (sizeof(ubu_sw_contig) * (size_t)n_contigs + 15) & ~(size_t)15, b_in = (sizeof(ubu_sw_proj_in) * (size_t)n + 15) & ~(size_t)15,
               b_out = (sizeof(ubu_sw_proj_out) * (size_t)n + 15) & ~(size_t)15;
  CU(h, h->p_bytes.reserve(b_c + b_in + b_out + 64));
  CU(h, h->p_words.reserve(contig_cigar_pool_n + cigar_pool_cap + 8));
  ubu_sw_contig* d_c = reinterpret_cast<ubu_sw_contig*>(h->p_bytes.p);
  ubu_sw_proj_in* d_in = reinterpret_cast<ubu_sw_proj_in*>(h->p_bytes.p + b_c);
  ubu_sw_proj_out* d_out = reinterpret_cast<ubu_sw_proj_out*>(h->p_bytes.p + b_c + b_in);
  uint32_t* d_ccig = h->p_words.p; uint32_t* d_pool = h->p_words.p + contig_cigar_pool_n;
  CU(h, s.d_ctr.reserve(64));
  CU(h, cudaMemsetAsync(s.d_ctr.p, 0, sizeof(unsigned int), st));
  if (n_contigs) CU(h, cudaMemcpyAsync(d_c, contigs, sizeof(ubu_sw_contig) * (size_t)n_contigs, cudaMemcpyHostToDevice, st));
  if (contig_cigar_pool_n) CU(h, cudaMemcpyAsync(d_ccig, contig_cigar_pool, sizeof(uint32_t) * contig_cigar_pool_n, cudaMemcpyHostToDevice, st));
  CU(h, cudaMemcpyAsync(d_in, in, sizeof(ubu_sw_proj_in) * (size_t)n, cudaMemcpyHostToDevice, st));
  post_project_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_c, n_contigs, d_ccig, d_in, n, d_out, d_pool, (unsigned)cigar_pool_cap, s.d_ctr.p);
  CU(h, cudaGetLastError());
  unsigned int used = 0;
  CU(h, cudaMemcpyAsync(&used, s.d_ctr.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CU(h, cudaMemcpyAsync(out, d_out, sizeof(ubu_sw_proj_out) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(h, cudaStreamSynchronize(st));
  if (cigar_pool_used) *cigar_pool_used = used;
  if (used > cigar_pool_cap) return UBU_SW_ERR_CIGAR_POOL;
  if (used) CU(h, cudaMemcpy(cigar_pool, d_pool, sizeof(uint32_t) * used, cudaMemcpyDeviceToHost));
  return UBU_SW_OK;
}

// Device-free view of the host plan (diagnostics and CPU tests): which forward launches a batch would get and in which order
// its tasks would run. launches_out: 6 ints per launch = {rows of the length class (G*K), has_n, first slot, slots, widest window,
// shared_window}; order_out: task of every slot (0xFFFFFFFF = padding), identity when *identity is set.
int ubu_sw_debug_plan(const ubu_sw_seqs* reads, const ubu_sw_seqs* windows, const uint32_t* pair_ref_idx, uint32_t* order_out,
                      size_t order_cap, uint32_t* n_slots, int* identity, int32_t* launches_out, int launches_cap, int* n_launches) {
  if (!reads || !windows || !n_slots || !n_launches) return UBU_SW_ERR_ARG;
  int rc = validate_seqs(reads); if (rc) return rc;
  rc = validate_seqs(windows); if (rc) return rc;
  if (!pair_ref_idx && reads->count != windows->count) return UBU_SW_ERR_ARG;
  Slot s;
  *n_slots = 0; *n_launches = 0;
  if (identity) *identity = 0;
  if (reads->count == 0) return UBU_SW_OK;
  rc = plan_batch(nullptr, s, reads, windows, pair_ref_idx); if (rc) return rc;
  const uint32_t n_long = (uint32_t)s.long_tasks.size();
  *n_slots = s.n_slots + n_long; *n_launches = (int)s.launches.size() + (n_long ? 1 : 0);
  if (identity) *identity = s.identity_order ? 1 : 0;
  if (order_out) {
    if (order_cap < s.n_slots + n_long) return UBU_SW_ERR_CAPACITY;
    for (uint32_t k = 0; k < s.n_slots + n_long; ++k) order_out[k] = s.identity_order ? (k < reads->count ? k : 0xFFFFFFFFu) : s.plain_order[k];
  }
  if (launches_out) {
    if (launches_cap < *n_launches) return UBU_SW_ERR_CAPACITY;
    if (n_long) {                      // the long tasks (sw_long_kernel) follow the slots: rows = -1 marks the entry
      int32_t* o = launches_out + 6 * s.launches.size();
      o[0] = -1; o[1] = 0; o[2] = (int32_t)s.n_slots; o[3] = (int32_t)n_long; o[4] = 0; o[5] = 0;
    }
    for (size_t k = 0; k < s.launches.size(); ++k) {
      const FwdLaunch& L = s.launches[k];
      int32_t* o = launches_out + 6 * k;
      o[0] = kFwdCfgs[L.cfg].G * kFwdCfgs[L.cfg].K; o[1] = L.has_n ? 1 : 0; o[2] = (int32_t)L.begin; o[3] = (int32_t)L.n_slots; o[4] = L.wmax; o[5] = L.shared_win ? 1 : 0;
    }
  }
  return UBU_SW_OK;
}

void* ubu_sw_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void ubu_sw_host_free(void* p) { if (p) cudaFreeHost(p); }

// Sequence.java:10-13 (A=0 C=1 T=2 G=3), :31-41 (LSB first); case-folded, non-ACGT -> N plane (SPEC.md section 1).
void ubu_sw_pack(const char* text, size_t n, uint8_t* packed, uint8_t* nmask, int* has_n) {
  struct Table { int8_t v[256]; };
  static const Table tab = [] {                            // initialised once, thread-safe (C++11 static)
    Table t;
    for (int k = 0; k < 256; ++k) t.v[k] = -1;
    t.v['A'] = t.v['a'] = 0; t.v['C'] = t.v['c'] = 1; t.v['T'] = t.v['t'] = 2; t.v['G'] = t.v['g'] = 3;
    return t;
  }();
  const int8_t* code = tab.v;
  memset(packed, 0, (n + 3) / 4);
  if (nmask) memset(nmask, 0, (n + 7) / 8);
  int any = 0;
  for (size_t k = 0; k < n; ++k) {
    const int c = code[(unsigned char)text[k]];
    if (c >= 0) packed[k >> 2] |= (uint8_t)(c << ((k & 3) * 2));
    else { any = 1; if (nmask) nmask[k >> 3] |= (uint8_t)(1u << (k & 7)); }
  }
  if (has_n) *has_n = any;
}

// ReverseComplementor.java:15-22,38-62
void ubu_sw_revcomp(const char* in, size_t n, char* out) {
  for (size_t k = 0; k < n; ++k) {
    char c = in[n - 1 - k];
    switch (c) { case 'A': c = 'T'; break; case 'T': c = 'A'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; default: break; }
    out[k] = c;
  }
}

int ubu_sw_cigar_string(const uint32_t* ops, uint32_t n, char* buf, size_t cap) {
  if (!buf || cap == 0) return 0;
  if (n == 0) { if (cap >= 2) { buf[0] = '*'; buf[1] = 0; return 1; } buf[0] = 0; return 0; }
  size_t w = 0;
  for (uint32_t k = 0; k < n; ++k) {
    const char opc = "MIDNSHP=X"[(ops[k] & 15u) < 9 ? (ops[k] & 15u) : 0];
    int m = snprintf(buf + w, cap - w, "%u%c", ops[k] >> 4, opc);
    if (m < 0 || (size_t)m >= cap - w) { buf[w] = 0; return (int)w; }
    w += (size_t)m;
  }
  return (int)w;
}

}  // extern "C"
