// ubu_kmer.cu -- k-mer graph build for ubu's assembler on B200 (include/ubu_kmer.h).
//
// What the reference does (one Assembler per region, single-threaded, HashMap<Sequence, Node>):
//   addToGraph            Assembler.java:452-474   every k-mer of every read: find-or-create the node, count it, note the
//                                                  contributing read, link it to the node of the previous position
//   Node                  Node.java:78-126         count, "at least two different read strings" flag, ordered toNodes / fromNodes
//   filterLowFrequencyNodes / identifyRootNodes    Assembler.java:201-227, 270-276
//
// B200 design: HBM-bound hashing, not arithmetic. One 128-byte slot per node (= one L2 line / one DRAM burst pair) in an
// open-addressing table; one thread per k-mer occurrence. Everything sequential in the Java becomes an order-free atomic:
//   count                         atomicAdd
//   insertion order of nodes      atomicMin of the occurrence index (read * max_occ + pos) -> output order = the order the
//                                                                                            Java inserts into Assembler.nodes
//   hasMultipleUniqueReads        reads are first put into exact equality classes (a hash table over whole reads with a
//                                 full compare on every hit), named by their smallest member. The first read that
//                                 contributes to a node is then its own class, so the node only tracks the LARGEST class
//                                 (one atomicMin on the complement): it differs from the first read iff two different read
//                                 strings contributed
//   toNodes / fromNodes order     the successor (predecessor) of a k-mer is determined by one base, so a node has four
//                                 to-slots and four from-slots; each keeps the 64-bit atomicMin of (occurrence that walks
//                                 the edge << 32 | neighbour's slot); sorted by that key they are the Java's array order
// All results are therefore independent of the order in which threads run, and bit-identical to the sequential Java.
// Passes: clear table -> read classes -> insert (+count/first/class/edges) -> sort nodes by first occurrence (cub radix
// sort, 32-bit keys) -> rank -> emit (filter flag, root flag, neighbour indices).
#include "sw_common.cuh"
#include "../../include/ubu_kmer.h"

#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>

using namespace ubu;

namespace {

constexpr uint32_t ST_EMPTY = 0xFFFFFFFFu, ST_LOCKED = 1u, ST_FULL = 2u;
constexpr uint32_t NONE = 0xFFFFFFFFu;

// The table is cleared with one memset to 0xFF: state = EMPTY, count_m1 = -1, first / to / from / class minima = +inf.
// Occurrence keys are the 32-bit occurrence index read * max_occ + pos (max_occ = most k-mers any read of the batch has):
// ordered like (read, pos), and half the width of the pair for the atomics and the sort.
struct __align__(128) NodeSlot {
  unsigned long long k0, k1;
  uint32_t region, state;
  uint32_t count_m1;                 // occurrences - 1
  uint32_t first;                    // min occurrence key; first / max_occ = the first contributing read = its own class
  uint32_t ncls_min;                 // min over ~class = ~(max class): two different read strings <=> max class != first read
  uint32_t rank;                     // index in the output table
  unsigned long long to[4];          // by appended base: (key of the occurrence that first walked the edge) << 32 | successor's slot
  unsigned long long from[4];        // by the predecessor's first base: key << 32 | predecessor's slot
  uint32_t pad[6];
};
static_assert(sizeof(NodeSlot) == 128, "one node = one 128-byte line");

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
__device__ __forceinline__ unsigned long long node_hash(unsigned long long k0, unsigned long long k1, uint32_t region) {
  return mix64(k0 ^ mix64(k1 + 0x9E3779B97F4A7C15ULL * (unsigned long long)(region + 1u)));
}

// k bases of a read starting at base i, 2 bits per base, base j of the k-mer at bits 2j (Sequence.java:31-41).
__device__ __forceinline__ void load_kmer(const uint8_t* __restrict__ seq, int i, int k, unsigned long long& k0, unsigned long long& k1) {
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int c = 0; c < 4; ++c)
    if (16 * c < k) w[c] = load16(seq, i + 16 * c);
  k0 = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
  k1 = (unsigned long long)w[2] | ((unsigned long long)w[3] << 32);
  if (k < 32) { k0 &= (1ull << (2 * k)) - 1ull; k1 = 0ull; }
  else if (k == 32) k1 = 0ull;
  else if (k < 64) k1 &= (1ull << (2 * (k - 32))) - 1ull;
}

__device__ __forceinline__ uint32_t ld_state(const NodeSlot* s) { return *reinterpret_cast<const volatile uint32_t*>(&s->state); }

// Find-or-create. Returns the slot, NONE when the table is full. `created` tells the caller it made the node.
// The first 32 bytes of a slot (key, region, state) are fetched in one round trip. A full slot's key never changes, so a
// match read together with state == FULL is final; a MISMATCH may be a key read before its owner published it, and is
// confirmed by a second read that is ordered after the FULL state.
// `overflow` (set by whoever finds the node bound exceeded) is polled every 64 probes: once the batch is known to need a
// bigger table nobody keeps walking a nearly full one.
__device__ __forceinline__ uint32_t find_or_insert(NodeSlot* __restrict__ tab, unsigned long long mask, unsigned long long k0,
                                                   unsigned long long k1, uint32_t region, bool& created,
                                                   const unsigned int* overflow) {
  created = false;
  unsigned long long s = node_hash(k0, k1, region) & mask;
  for (unsigned long long probes = 0; probes <= mask; ++probes, s = (s + 1ull) & mask) {
    if ((probes & 63ull) == 63ull && *reinterpret_cast<const volatile unsigned int*>(overflow)) return NONE;
    NodeSlot* p = tab + s;
    const uint4 a = __ldcv(reinterpret_cast<const uint4*>(p));          // k0, k1
    const uint2 b = __ldcv(reinterpret_cast<const uint2*>(p) + 2);      // region, state
    uint32_t st = b.y;
    if (st == ST_FULL && a.x == (uint32_t)k0 && a.y == (uint32_t)(k0 >> 32) && a.z == (uint32_t)k1 && a.w == (uint32_t)(k1 >> 32) &&
        b.x == region)
      return (uint32_t)s;
    if (st == ST_EMPTY) {
      st = atomicCAS(&p->state, ST_EMPTY, ST_LOCKED);
      if (st == ST_EMPTY) {
        p->k0 = k0; p->k1 = k1; p->region = region;
        __threadfence();
        atomicExch(&p->state, ST_FULL);
        created = true;
        return (uint32_t)s;
      }
    }
    while (st == ST_LOCKED) st = ld_state(p);            // the owner publishes within a few instructions
    __threadfence();
    const volatile NodeSlot* q = p;
    if (q->k0 == k0 && q->k1 == k1 && q->region == region) return (uint32_t)s;
  }
  return NONE;
}

// 128-bit compare-and-swap on a 16-byte aligned key (atom.cas.b128, sm_90+): returns the previous value.
__device__ __forceinline__ void cas128(void* addr, unsigned long long cmp_lo, unsigned long long cmp_hi, unsigned long long new_lo,
                                       unsigned long long new_hi, unsigned long long& old_lo, unsigned long long& old_hi) {
  asm volatile("{\n\t"
               ".reg .b128 c, n, o;\n\t"
               "mov.b128 c, {%2, %3};\n\t"
               "mov.b128 n, {%4, %5};\n\t"
               "atom.global.cas.b128 o, [%6], c, n;\n\t"
               "mov.b128 {%0, %1}, o;\n\t"
               "}"
               : "=l"(old_lo), "=l"(old_hi) : "l"(cmp_lo), "l"(cmp_hi), "l"(new_lo), "l"(new_hi), "l"(addr) : "memory");
}

// Lock-free find-or-create for k <= 63 (a valid k-mer then never equals the all-ones EMPTY key, and its high word is never
// all ones). A 128-bit CAS installs the k-mer in an empty slot; a slot whose key matches belongs to the first region that
// claims it with a 32-bit CAS on the region word (the same k-mer of another region moves on to the next slot). No lock, no
// fence, no spinning: key and region are each written by exactly one atomic and never change afterwards.
// The common case (the node exists) costs one round trip of plain loads and no atomic: a key read whose two 64-bit words
// are both different from all-ones is a complete, final key, whatever the hardware does about 128-bit tearing; any read with
// an all-ones word (empty, possibly mid-install, or a k-mer starting with 32 G) is settled by the CAS instead.
__device__ __forceinline__ uint32_t find_or_insert_lockfree(NodeSlot* __restrict__ tab, unsigned long long mask, unsigned long long k0,
                                                            unsigned long long k1, uint32_t region, bool& created,
                                                            const unsigned int* overflow) {
  created = false;
  unsigned long long s = node_hash(k0, k1, region) & mask;
  for (unsigned long long probes = 0; probes <= mask; ++probes, s = (s + 1ull) & mask) {
    if ((probes & 63ull) == 63ull && *reinterpret_cast<const volatile unsigned int*>(overflow)) return NONE;
    NodeSlot* p = tab + s;
    const ulonglong2 a = __ldcv(reinterpret_cast<const ulonglong2*>(p));
    uint32_t r = __ldcv(&p->region);
    unsigned long long o0 = a.x, o1 = a.y;
    if (o0 == ~0ull || o1 == ~0ull) {                                     // not a settled key: let the CAS decide
      cas128(p, ~0ull, ~0ull, k0, k1, o0, o1);
      if ((o0 & o1) == ~0ull) { o0 = k0; o1 = k1; r = NONE; }             // installed just now: nobody has claimed it yet ...
      else r = __ldcv(&p->region);                                        // ... unless the CAS below says otherwise
    }
    if (o0 != k0 || o1 != k1) continue;                                   // another k-mer lives here
    if (r == NONE) {
      r = atomicCAS(&p->region, NONE, region);
      if (r == NONE) { created = true; return (uint32_t)s; }
    }
    if (r == region) return (uint32_t)s;
  }
  return NONE;
}

// ---- pass 1: exact equality classes of the reads (Node.addReadSequence compares read STRINGS, Node.java:78-90) -----------
__device__ __forceinline__ uint32_t read_chunk(const uint8_t* __restrict__ seq, int L, int c) {   // 16 bases, tail masked
  const int rem = L - 16 * c;
  uint32_t w = load16(seq, 16 * c);
  if (rem < 16) w &= (1u << (2 * rem)) - 1u;
  return w;
}

__global__ void kmer_read_class_kernel(DevSeqs reads, const uint32_t* __restrict__ region, uint32_t n, uint32_t* __restrict__ rtab,
                                       uint32_t rmask, uint32_t* __restrict__ cls, unsigned int* __restrict__ ctr) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  if (region && region[r] == NONE) { cls[r] = NONE; return; }
  const int L = reads.len[r];
  const uint8_t* q = reads.packed + reads.off[r];
  if (reads.nmask) {                                       // a read with an N is not added to the graph (Assembler.java:77-81)
    const uint8_t* m = reads.nmask + reads.nmask_off[r];
    uint32_t any = 0u;
    for (int c = 0; 16 * c < L; ++c) { uint32_t w = load16n(m, 16 * c); const int rem = L - 16 * c; if (rem < 16) w &= (1u << rem) - 1u; any |= w; }
    if (any) { cls[r] = NONE; atomicAdd(&ctr[2], 1u); return; }
  }
  // A class = (region, read string): its members contribute to the same nodes. The table slot ends up holding the
  // smallest member; cls[r] is the slot for now and is resolved to that member by kmer_read_class_resolve_kernel.
  const uint32_t reg = region ? region[r] : 0u;
  unsigned long long h = 0x51ED270B1ull + (unsigned long long)L + ((unsigned long long)reg << 20);
  for (int c = 0; 16 * c < L; ++c) h = mix64(h ^ (unsigned long long)read_chunk(q, L, c));
  uint32_t s = (uint32_t)h & rmask;
  for (;;) {
    const uint32_t old = atomicCAS(&rtab[s], NONE, r);
    if (old == NONE) { cls[r] = s; return; }
    bool same = reads.len[old] == L && (region ? region[old] : 0u) == reg;
    if (same) {
      const uint8_t* o = reads.packed + reads.off[old];
      for (int c = 0; same && 16 * c < L; ++c) same = read_chunk(q, L, c) == read_chunk(o, L, c);
    }
    if (same) { atomicMin(&rtab[s], r); cls[r] = s; return; }      // any member serves as the one to compare with
    s = (s + 1u) & rmask;
  }
}

__global__ void kmer_read_class_resolve_kernel(const uint32_t* __restrict__ rtab, uint32_t n, uint32_t* __restrict__ cls) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n && cls[r] != NONE) cls[r] = rtab[cls[r]];
}

// ---- pass 2: the hash-table kernel. addToGraph's loop body (Assembler.java:456-472): find-or-create the node, count it, note
// the read, link it to the node of the previous position.
// A warp takes 31 consecutive occurrences (occurrence index = read * max_occ + position) on lanes 1..31; lane 0 only LOOKS UP
// the occurrence before them, so that every lane gets its predecessor's slot from the lane below by shuffle and all 32 lookups
// run side by side (no lane does a second, serialised lookup). After the lookup everything is a fire-and-forget reduction.
constexpr int KMER_OCC_PER_WARP = 31;
constexpr unsigned long long KMER_PERM_MUL = 1000003ull;       // prime, larger than any permutation block

template <bool LOCKFREE>
__global__ void __launch_bounds__(256)
kmer_insert_kernel(DevSeqs reads, const uint32_t* __restrict__ region, const uint32_t* __restrict__ cls, uint32_t n_reads,
                   uint32_t max_occ, int k, NodeSlot* __restrict__ tab, unsigned long long mask, uint32_t* __restrict__ node_list,
                   unsigned int* __restrict__ ctr, uint32_t max_nodes, unsigned long long n_warps, unsigned long long perm_warps) {
  // Warps that run at the same time take occurrences that are far apart (a multiplicative permutation of the warp index
  // inside blocks of KMER_PERM_WARPS warps): neighbouring reads of a region cover the same k-mers, and spreading them in
  // time keeps their reductions from queueing on the same L2 lines; keeping the permutation inside a block of ~4M
  // occurrences keeps that block's nodes L2-resident while it is being worked on.
  const unsigned long long hw_warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (hw_warp >= n_warps) return;
  const unsigned long long blk = hw_warp / perm_warps, base = blk * perm_warps;
  const unsigned long long span = n_warps - base < perm_warps ? n_warps - base : perm_warps;
  const unsigned long long gwarp = base + ((hw_warp - base) * KMER_PERM_MUL) % span;      // KMER_PERM_MUL is prime and > span
  const uint32_t lane = threadIdx.x & 31u;
  const long long occ = (long long)gwarp * KMER_OCC_PER_WARP + (long long)lane - 1;     // lane 0: the occurrence before the warp's
  uint32_t r = 0u, i = 0u, c = NONE;
  bool valid = false;
  if (occ >= 0) {
    r = (uint32_t)((unsigned long long)occ / max_occ); i = (uint32_t)((unsigned long long)occ % max_occ);
    if (r < n_reads) { c = cls[r]; valid = c != NONE && (int)i + k <= (int)reads.len[r]; }
  }
  // lane 0 is needed only as the predecessor of lane 1's occurrence (same read, so lane 1 is at a position > 0)
  const bool owner = lane != 0u;
  if (!owner) valid = valid && i + 1u < max_occ;
  if (*reinterpret_cast<const volatile unsigned int*>(ctr + 1)) valid = false;          // the run already failed (node bound exceeded)
  const uint8_t* q = valid ? reads.packed + reads.off[r] : nullptr;
  const uint32_t reg = (valid && region) ? region[r] : 0u;
  uint32_t s = NONE;
  bool created = false;
  if (valid) {
    unsigned long long k0, k1;
    load_kmer(q, (int)i, k, k0, k1);
    s = LOCKFREE ? find_or_insert_lockfree(tab, mask, k0, k1, reg, created, ctr + 1) : find_or_insert(tab, mask, k0, k1, reg, created, ctr + 1);
  }
  const uint32_t prev = __shfl_up_sync(0xffffffffu, s, 1);                 // slot of occurrence occ - 1 (same read when i > 0)
  // new nodes are appended to the node list with one counter update per warp
  const uint32_t m = __ballot_sync(0xffffffffu, created);
  if (m) {
    uint32_t base = 0u;
    if (lane == 0u) base = atomicAdd(&ctr[0], (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (created) {
      const uint32_t idx = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
      if (idx < max_nodes) node_list[idx] = s; else atomicOr(&ctr[1], 1u);
    }
  }
  if (!valid) return;
  if (s == NONE) { atomicOr(&ctr[1], 1u); return; }                        // table full
  if (!owner) return;
  NodeSlot* p = tab + s;
  const uint32_t key = (uint32_t)occ;                                       // ordered like (read, position)
  atomicAdd(&p->count_m1, 1u);
  atomicMin(&p->first, key);
  atomicMin(&p->ncls_min, ~c);                                              // Node.addReadSequence, Node.java:78-90
  if (i > 0u) {                                                             // prev.addToNode(node), Node.java:118-130
    if (prev == NONE) { atomicOr(&ctr[1], 1u); return; }
    const uint32_t b_next = base2(q, i + (uint32_t)k - 1u), b_prev = base2(q, i - 1u);
    atomicMin(&tab[prev].to[b_next], ((unsigned long long)key << 32) | (unsigned long long)s);      // the first walk of an edge
    atomicMin(&p->from[b_prev], ((unsigned long long)key << 32) | (unsigned long long)prev);        // fixes its place in the list
  }
}

__global__ void kmer_keys_kernel(const NodeSlot* __restrict__ tab, const uint32_t* __restrict__ node_list, uint32_t n,
                                 uint32_t* __restrict__ keys) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) keys[j] = tab[node_list[j]].first;
}

__global__ void kmer_rank_kernel(NodeSlot* __restrict__ tab, const uint32_t* __restrict__ sorted_slot, uint32_t n) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) tab[sorted_slot[j]].rank = j;
}

__device__ __forceinline__ bool slot_multi(const NodeSlot* p, uint32_t max_occ) {        // Node.hasMultipleUniqueReads
  return ~p->ncls_min != p->first / max_occ;
}
__device__ __forceinline__ bool slot_filtered(const NodeSlot* p, uint32_t min_freq, uint32_t max_occ) {   // Assembler.java:204-210
  return p->count_m1 + 1u < min_freq || !slot_multi(p, max_occ);
}

// ---- pass 5: one thread per node, in output order --------------------------------------------------------------------------
__global__ void kmer_emit_kernel(const NodeSlot* __restrict__ tab, const uint32_t* __restrict__ sorted_slot, uint32_t n,
                                 uint32_t max_occ, uint32_t min_freq, ubu_kmer_node* __restrict__ out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const NodeSlot* p = tab + sorted_slot[j];
  ubu_kmer_node o;
  o.kmer[0] = p->k0; o.kmer[1] = p->k1; o.region = p->region; o.count = p->count_m1 + 1u;
  o.first_read = p->first / max_occ; o.first_pos = (uint16_t)(p->first % max_occ);
  const bool filtered = slot_filtered(p, min_freq, max_occ);
  uint32_t flags = (slot_multi(p, max_occ) ? UBU_KMER_FLAG_MULTI_READS : 0u) | (filtered ? UBU_KMER_FLAG_FILTERED : 0u);
  uint32_t tkey[4], fkey[4], tnode[4], fnode[4];
  bool live_pred = false;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const unsigned long long t = p->to[b], f = p->from[b];
    tkey[b] = (uint32_t)(t >> 32); tnode[b] = NONE;
    if (tkey[b] != NONE) tnode[b] = tab[(uint32_t)t].rank;
    fkey[b] = (uint32_t)(f >> 32); fnode[b] = NONE;
    if (fkey[b] != NONE) {
      const NodeSlot* fp = tab + (uint32_t)f;
      fnode[b] = fp->rank;
      if (!slot_filtered(fp, min_freq, max_occ)) live_pred = true;
    }
  }
  if (!filtered && !live_pred) flags |= UBU_KMER_FLAG_ROOT;                 // Assembler.java:270-276 after the filter
  // order of creation = array order of Node.toNodes / Node.fromNodes (4-element insertion sort; absent entries sort last)
#pragma unroll
  for (int a = 1; a < 4; ++a)
#pragma unroll
    for (int c = a; c > 0; --c) {
      if (tkey[c] < tkey[c - 1]) { uint32_t t = tkey[c]; tkey[c] = tkey[c - 1]; tkey[c - 1] = t; t = tnode[c]; tnode[c] = tnode[c - 1]; tnode[c - 1] = t; }
      if (fkey[c] < fkey[c - 1]) { uint32_t t = fkey[c]; fkey[c] = fkey[c - 1]; fkey[c - 1] = t; t = fnode[c]; fnode[c] = fnode[c - 1]; fnode[c - 1] = t; }
    }
#pragma unroll
  for (int b = 0; b < 4; ++b) { o.to_node[b] = tnode[b]; o.from_node[b] = fnode[b]; }
  o.flags = (uint16_t)flags;
  out[j] = o;
}

template <class T>
struct Buf {
  T* p = nullptr; size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, (n + 64) * sizeof(T));
    if (e == cudaSuccess) cap = n + 64;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct ubu_kmer_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string last_error;
  Buf<uint8_t> d_packed, d_nmask, d_sort_tmp;
  Buf<uint32_t> d_off, d_noff, d_region, d_cls, d_rtab, d_list, d_sorted;
  Buf<uint16_t> d_len;
  Buf<uint32_t> d_keys, d_keys2;
  Buf<NodeSlot> d_tab;
  Buf<ubu_kmer_node> d_out;
  Buf<unsigned int> d_ctr;
  // staged batch
  bool staged = false, ran = false, has_region = false, has_n = false;
  uint32_t n_reads = 0, max_occ = 0, rmask = 0, max_nodes = 0, launches = 0;
  unsigned long long mask = 0, occurrences = 0;
  size_t sort_tmp_bytes = 0;
  ubu_kmer_params params{0, 0};
  unsigned int h_ctr[4] = {0, 0, 0, 0};
};

namespace {
int kfail(ubu_kmer_handle* h, cudaError_t e, const char* what) {
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  if (h) h->last_error = buf;
  return UBU_SW_ERR_CUDA;
}
#define KCU(h, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return kfail((h), e__, #call); } while (0)

unsigned long long pow2_at_least(unsigned long long x) { unsigned long long p = 1; while (p < x) p <<= 1; return p; }
}  // namespace

extern "C" {

int ubu_kmer_create(int device, ubu_kmer_handle** out) {
  if (!out) return UBU_SW_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return UBU_SW_ERR_NO_DEVICE; }
  if (device < 0 || device >= ndev) return UBU_SW_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return UBU_SW_ERR_NO_DEVICE; }
  if (prop.major != 10) return UBU_SW_ERR_NO_DEVICE;            // sm_100a code only
  ubu_kmer_handle* h = new (std::nothrow) ubu_kmer_handle();
  if (!h) return UBU_SW_ERR_NOMEM;
  h->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return UBU_SW_ERR_CUDA; }
  for (int e = 0; e < 4; ++e)
    if (cudaEventCreate(&h->ev[e]) != cudaSuccess) { ubu_kmer_destroy(h); return UBU_SW_ERR_CUDA; }
  *out = h;
  return UBU_SW_OK;
}

void ubu_kmer_destroy(ubu_kmer_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  h->d_packed.release(); h->d_nmask.release(); h->d_sort_tmp.release(); h->d_off.release(); h->d_noff.release(); h->d_region.release();
  h->d_cls.release(); h->d_rtab.release(); h->d_list.release(); h->d_sorted.release(); h->d_len.release();
  h->d_keys.release(); h->d_keys2.release(); h->d_tab.release(); h->d_out.release(); h->d_ctr.release();
  for (int e = 0; e < 4; ++e) if (h->ev[e]) cudaEventDestroy(h->ev[e]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

const char* ubu_kmer_last_error(ubu_kmer_handle* h) { return h ? h->last_error.c_str() : ""; }

int ubu_kmer_stage(ubu_kmer_handle* h, const ubu_kmer_params* params, const ubu_sw_seqs* reads, const uint32_t* read_region, size_t max_nodes) {
  if (!h || !params || !reads) return UBU_SW_ERR_ARG;
  if (params->kmer_size < 1 || params->kmer_size > UBU_KMER_MAX_K) return UBU_SW_ERR_ARG;
  if (reads->count && (!reads->off || !reads->len || (reads->packed_bytes && !reads->packed))) return UBU_SW_ERR_ARG;
  if (reads->nmask && !reads->nmask_off) return UBU_SW_ERR_ARG;
  h->staged = false; h->ran = false;
  const uint32_t n = reads->count;
  const int k = (int)params->kmer_size;
  unsigned long long occ = 0; uint32_t max_occ = 0;
  for (uint32_t r = 0; r < n; ++r) {
    const uint32_t L = reads->len[r];
    if ((uint64_t)reads->off[r] + (L + 3u) / 4u > reads->packed_bytes) return UBU_SW_ERR_ARG;
    if (reads->nmask && (uint64_t)reads->nmask_off[r] + (L + 7u) / 8u > reads->nmask_bytes) return UBU_SW_ERR_ARG;
    if (read_region && read_region[r] == NONE) continue;
    if ((int)L >= k) { const uint32_t o = L - (uint32_t)k + 1u; occ += o; max_occ = std::max(max_occ, o); }
  }
  if (occ > 0xFFFFFFF0ull || (unsigned long long)n * std::max<uint32_t>(max_occ, 1u) > 0xFFFFFFF0ull) return UBU_SW_ERR_TOO_LONG;   // 32-bit slot ids / occurrence grid
  h->params = *params; h->n_reads = n; h->max_occ = max_occ; h->occurrences = occ;
  h->has_region = read_region != nullptr; h->has_n = reads->nmask != nullptr;
  const unsigned long long want_nodes = std::max<unsigned long long>(std::min<unsigned long long>(max_nodes ? max_nodes : occ, occ), 1ull);
  h->max_nodes = (uint32_t)want_nodes;
  const unsigned long long cap = pow2_at_least(std::max<unsigned long long>(1024ull, want_nodes + want_nodes / 2ull));
  h->mask = cap - 1ull;
  h->rmask = (uint32_t)pow2_at_least(std::max<unsigned long long>(1024ull, 2ull * n)) - 1u;
  KCU(h, cudaSetDevice(h->device));
  KCU(h, h->d_packed.reserve((size_t)reads->packed_bytes + 16)); KCU(h, h->d_off.reserve(n)); KCU(h, h->d_len.reserve(n));
  if (h->has_n) { KCU(h, h->d_nmask.reserve((size_t)reads->nmask_bytes + 16)); KCU(h, h->d_noff.reserve(n)); }
  if (h->has_region) KCU(h, h->d_region.reserve(n));
  KCU(h, h->d_cls.reserve(n)); KCU(h, h->d_rtab.reserve((size_t)h->rmask + 1));
  KCU(h, h->d_list.reserve(h->max_nodes)); KCU(h, h->d_sorted.reserve(h->max_nodes));
  KCU(h, h->d_keys.reserve(h->max_nodes)); KCU(h, h->d_keys2.reserve(h->max_nodes));
  KCU(h, h->d_tab.reserve((size_t)cap)); KCU(h, h->d_out.reserve(h->max_nodes)); KCU(h, h->d_ctr.reserve(8));
  size_t tmp = 0;
  KCU(h, cub::DeviceRadixSort::SortPairs(nullptr, tmp, h->d_keys.p, h->d_keys2.p, h->d_list.p, h->d_sorted.p, (int)h->max_nodes, 0, 32, h->stream));
  h->sort_tmp_bytes = tmp;
  KCU(h, h->d_sort_tmp.reserve(tmp + 16));
  cudaStream_t st = h->stream;
  if (n) {
    KCU(h, cudaMemcpyAsync(h->d_packed.p, reads->packed, reads->packed_bytes, cudaMemcpyHostToDevice, st));
    KCU(h, cudaMemcpyAsync(h->d_off.p, reads->off, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    KCU(h, cudaMemcpyAsync(h->d_len.p, reads->len, sizeof(uint16_t) * n, cudaMemcpyHostToDevice, st));
    if (h->has_n) {
      KCU(h, cudaMemcpyAsync(h->d_nmask.p, reads->nmask, reads->nmask_bytes, cudaMemcpyHostToDevice, st));
      KCU(h, cudaMemcpyAsync(h->d_noff.p, reads->nmask_off, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    }
    if (h->has_region) KCU(h, cudaMemcpyAsync(h->d_region.p, read_region, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  }
  KCU(h, cudaStreamSynchronize(st));
  h->staged = true;
  return UBU_SW_OK;
}

int ubu_kmer_run(ubu_kmer_handle* h) {
  if (!h || !h->staged) return UBU_SW_ERR_ARG;
  KCU(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const uint32_t n = h->n_reads;
  const int k = (int)h->params.kmer_size;
  h->launches = 0;
  KCU(h, cudaEventRecord(h->ev[0], st));
  KCU(h, cudaMemsetAsync(h->d_ctr.p, 0, 8 * sizeof(unsigned int), st));
  KCU(h, cudaMemsetAsync(h->d_tab.p, 0xFF, (size_t)(h->mask + 1ull) * sizeof(NodeSlot), st));
  KCU(h, cudaMemsetAsync(h->d_rtab.p, 0xFF, ((size_t)h->rmask + 1) * sizeof(uint32_t), st));
  DevSeqs reads{h->d_packed.p, h->d_off.p, h->d_len.p, h->has_n ? h->d_nmask.p : nullptr, h->has_n ? h->d_noff.p : nullptr};
  const uint32_t* region = h->has_region ? h->d_region.p : nullptr;
  if (n) {
    kmer_read_class_kernel<<<(n + 127) / 128, 128, 0, st>>>(reads, region, n, h->d_rtab.p, h->rmask, h->d_cls.p, h->d_ctr.p);
    KCU(h, cudaGetLastError()); ++h->launches;
    kmer_read_class_resolve_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->d_rtab.p, n, h->d_cls.p);
    KCU(h, cudaGetLastError()); ++h->launches;
  }
  KCU(h, cudaEventRecord(h->ev[1], st));
  const unsigned long long total_occ = (unsigned long long)n * h->max_occ;     // occurrence index space (ragged reads leave holes)
  if (total_occ) {
    const unsigned long long warps = (total_occ + KMER_OCC_PER_WARP - 1) / KMER_OCC_PER_WARP;
    const unsigned int grid = (unsigned int)((warps * 32ull + 255ull) / 256ull);
    unsigned long long perm = 131072ull;                                      // warps per permutation block (~4M occurrences)
    if (const char* e = getenv("UBU_KMER_PERM_WARPS")) { const unsigned long long v = strtoull(e, nullptr, 10); if (v >= 1ull && v < KMER_PERM_MUL) perm = v; }
    // k <= 63: lock-free slots (a k-mer never equals the empty key); k = 64 uses every bit pattern, so it keeps the state word
    if (k <= 63 && !getenv("UBU_KMER_LOCKED"))
      kmer_insert_kernel<true><<<grid, 256, 0, st>>>(reads, region, h->d_cls.p, n, h->max_occ, k, h->d_tab.p, h->mask, h->d_list.p,
                                                    h->d_ctr.p, h->max_nodes, warps, perm);
    else
      kmer_insert_kernel<false><<<grid, 256, 0, st>>>(reads, region, h->d_cls.p, n, h->max_occ, k, h->d_tab.p, h->mask, h->d_list.p,
                                                     h->d_ctr.p, h->max_nodes, warps, perm);
    KCU(h, cudaGetLastError()); ++h->launches;
  }
  KCU(h, cudaEventRecord(h->ev[2], st));
  KCU(h, cudaMemcpyAsync(h->h_ctr, h->d_ctr.p, 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  KCU(h, cudaStreamSynchronize(st));                       // the node count sizes the remaining launches
  const unsigned int nn = h->h_ctr[0];
  if ((h->h_ctr[1] & 1u) || nn > h->max_nodes) {
    KCU(h, cudaEventRecord(h->ev[3], st));
    h->ran = true;
    return UBU_SW_ERR_CAPACITY;
  }
  if (nn) {
    kmer_keys_kernel<<<(nn + 255) / 256, 256, 0, st>>>(h->d_tab.p, h->d_list.p, nn, h->d_keys.p);
    KCU(h, cudaGetLastError()); ++h->launches;
    size_t tmp = h->sort_tmp_bytes;
    KCU(h, cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp, h->d_keys.p, h->d_keys2.p, h->d_list.p, h->d_sorted.p, (int)nn, 0, 32, st));
    kmer_rank_kernel<<<(nn + 255) / 256, 256, 0, st>>>(h->d_tab.p, h->d_sorted.p, nn);
    KCU(h, cudaGetLastError()); ++h->launches;
    kmer_emit_kernel<<<(nn + 127) / 128, 128, 0, st>>>(h->d_tab.p, h->d_sorted.p, nn, std::max<uint32_t>(h->max_occ, 1u), h->params.min_node_frequency, h->d_out.p);
    KCU(h, cudaGetLastError()); ++h->launches;
  }
  KCU(h, cudaEventRecord(h->ev[3], st));
  h->ran = true;
  return UBU_SW_OK;
}

int ubu_kmer_fetch(ubu_kmer_handle* h, ubu_kmer_node* out, size_t out_cap, size_t* n_nodes, uint32_t* n_skipped) {
  if (!h || !h->ran) return UBU_SW_ERR_ARG;
  KCU(h, cudaSetDevice(h->device));
  KCU(h, cudaStreamSynchronize(h->stream));
  const unsigned int nn = h->h_ctr[0];
  if (n_nodes) *n_nodes = nn;
  if (n_skipped) *n_skipped = h->h_ctr[2];
  if ((h->h_ctr[1] & 1u) || nn > h->max_nodes || nn > out_cap) return UBU_SW_ERR_CAPACITY;
  if (nn && !out) return UBU_SW_ERR_ARG;
  if (nn) KCU(h, cudaMemcpy(out, h->d_out.p, sizeof(ubu_kmer_node) * (size_t)nn, cudaMemcpyDeviceToHost));
  return UBU_SW_OK;
}

int ubu_kmer_build(ubu_kmer_handle* h, const ubu_kmer_params* params, const ubu_sw_seqs* reads, const uint32_t* read_region,
                   size_t max_nodes, ubu_kmer_node* out, size_t out_cap, size_t* n_nodes, uint32_t* n_skipped) {
  if (n_nodes) *n_nodes = 0;
  // No bound given: start from an eighth of the occurrences (typical coverage is far above 8) and fall back to the
  // occurrence count, which always suffices.
  for (int attempt = 0; attempt < 2; ++attempt) {
    size_t bound = max_nodes;
    if (!bound && attempt == 0 && reads) {
      unsigned long long occ = 0;
      if (params && params->kmer_size >= 1 && reads->len)
        for (uint32_t r = 0; r < reads->count; ++r) if (reads->len[r] >= params->kmer_size) occ += reads->len[r] - params->kmer_size + 1u;
      bound = (size_t)(occ / 8ull + 4096ull);
    }
    int rc = ubu_kmer_stage(h, params, reads, read_region, bound); if (rc) return rc;
    rc = ubu_kmer_run(h);
    if (rc == UBU_SW_ERR_CAPACITY && !max_nodes && attempt == 0) continue;
    if (rc && rc != UBU_SW_ERR_CAPACITY) return rc;
    return ubu_kmer_fetch(h, out, out_cap, n_nodes, n_skipped);
  }
  return UBU_SW_ERR_CAPACITY;
}

int ubu_kmer_last_run_ms(ubu_kmer_handle* h, float ms[4]) {
  if (!h || !h->ran || !ms) return UBU_SW_ERR_ARG;
  KCU(h, cudaSetDevice(h->device));
  KCU(h, cudaEventSynchronize(h->ev[3]));
  KCU(h, cudaEventElapsedTime(&ms[0], h->ev[0], h->ev[3]));
  KCU(h, cudaEventElapsedTime(&ms[1], h->ev[0], h->ev[1]));
  KCU(h, cudaEventElapsedTime(&ms[2], h->ev[1], h->ev[2]));
  KCU(h, cudaEventElapsedTime(&ms[3], h->ev[2], h->ev[3]));
  return UBU_SW_OK;
}

int ubu_kmer_last_run_stats(ubu_kmer_handle* h, uint64_t* occurrences, uint64_t* table_bytes, uint32_t* launches) {
  if (!h || !h->staged) return UBU_SW_ERR_ARG;
  if (occurrences) *occurrences = h->occurrences;
  if (table_bytes) *table_bytes = (h->mask + 1ull) * sizeof(NodeSlot);
  if (launches) *launches = h->launches;
  return UBU_SW_OK;
}

}  // extern "C"
