// peak_int16x2.cu -- measures the B200 (sm_100a) issue rate of the packed 16-bit x2
// integer instructions the Smith-Waterman cell update is built from, plus the
// neighbours it has to share issue slots with (PRMT, IADD3, IMAD, SHFL, LDS).
//
// This is the denominator of the roofline in bench.py / DESIGN.md:
//     P_int16x2      = sustained packed-16x2 warp-instructions/s x 32 lanes (whole chip)
//     roofline_GCUPS = P_int16x2 * 2 cells / 7 ops / 1e9            
//
// Build : nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o peak_int16x2 peak_int16x2.cu
// Run   : ./peak_int16x2            (prints one JSON object)
//
// Method: every thread owns CH independent register chains; the loop body applies
// one op to each chain, so there are no RAW stalls at >=4 warps per scheduler.
// Cycles come from clock64() around the loop (per CTA, averaged), wall time from
// CUDA events, SM MHz from clock64/globaltimer inside the kernel.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

enum Op { OP_VIADDMAX = 0, OP_VIMAX3RELU, OP_VADD2, OP_VIMAX, OP_PRMT, OP_IADD3, OP_IMAD, OP_LOP3,
          OP_SHFL, OP_LDS, OP_FFMA, OP_VIBMAX, OP_NONE };

static const char* op_name(int o) {
  switch (o) {
    case OP_VIADDMAX: return "VIADDMNMX.S16x2";
    case OP_VIMAX3RELU: return "VIMNMX3.S16x2.RELU";
    case OP_VADD2: return "VIADD.16x2";
    case OP_VIMAX: return "VIMNMX.S16x2";
    case OP_PRMT: return "PRMT";
    case OP_IADD3: return "IADD3";
    case OP_IMAD: return "IMAD";
    case OP_LOP3: return "LOP3";
    case OP_SHFL: return "SHFL";
    case OP_LDS: return "LDS.32";
    case OP_FFMA: return "FFMA";
    case OP_VIBMAX: return "VIMNMX.S16x2+pred";
    default: return "-";
  }
}

template <int OP>
__device__ __forceinline__ unsigned apply(unsigned a, unsigned b, unsigned c, const unsigned* sm) {
  if constexpr (OP == OP_VIADDMAX) return __viaddmax_s16x2(a, b, c);
  else if constexpr (OP == OP_VIMAX3RELU) return __vimax3_s16x2_relu(a, b, c);
  else if constexpr (OP == OP_VADD2) return __vadd2(a, b);
  else if constexpr (OP == OP_VIMAX) return __vmaxs2(a, b);
  else if constexpr (OP == OP_PRMT) return __byte_perm(a, b, c);
  else if constexpr (OP == OP_IADD3) { unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
  else if constexpr (OP == OP_IMAD) { unsigned r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  else if constexpr (OP == OP_LOP3) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  else if constexpr (OP == OP_SHFL) return __shfl_up_sync(0xffffffffu, a, 1);
  else if constexpr (OP == OP_LDS) return sm[(a + threadIdx.x) & 1023];
  else if constexpr (OP == OP_FFMA) return __float_as_uint(fmaf(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c)));
  else if constexpr (OP == OP_VIBMAX) { bool ph, pl; unsigned r = __vibmax_s16x2(a, b, &ph, &pl); return ph ? r : (pl ? c : r); }
  else return a;
}

// NA ops of type A and NB ops of type B per loop trip, each on its own chain.
template <int OPA, int NA, int OPB, int NB>
__global__ void __launch_bounds__(512, 2)
mix_kernel(unsigned* out, long long* cyc, unsigned long long* ns, unsigned b, unsigned c, int iters) {
  __shared__ unsigned sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7u) & 1023u;
  __syncthreads();
  unsigned a[NA > 0 ? NA : 1], d[NB > 0 ? NB : 1];
#pragma unroll
  for (int u = 0; u < NA; ++u) a[u] = threadIdx.x * 0x00010001u + u * b;
#pragma unroll
  for (int u = 0; u < NB; ++u) d[u] = threadIdx.x + u * c;
  unsigned long long t0g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0g));
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      // interleave A and B as evenly as the counts allow
      constexpr int N = (NA > NB ? NA : NB);
#pragma unroll
      for (int u = 0; u < N; ++u) {
        if (u < NA) a[u] = apply<OPA>(a[u], b, c, sm);
        if (u < NB) d[u] = apply<OPB>(d[u], b, c, sm);
      }
    }
  }
  long long t1 = clock64();
  unsigned long long t1g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1g));
  unsigned acc = 0;
#pragma unroll
  for (int u = 0; u < NA; ++u) acc ^= a[u];
#pragma unroll
  for (int u = 0; u < NB; ++u) acc ^= d[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; ns[blockIdx.x] = t1g - t0g; }
}

struct Row { std::string name; int na, nb; double a_per_clk_sm, b_per_clk_sm, tot_per_clk_sm, mhz, ms; };

template <int OPA, int NA, int OPB, int NB>
static Row run(int sms, int iters) {
  const int threads = 512, blocks = sms * 2;   // 32 warps per SM, 8 per scheduler, one wave
  unsigned* out; long long* cyc; unsigned long long* ns;
  CK(cudaMalloc(&out, sizeof(unsigned) * threads * blocks));
  CK(cudaMalloc(&cyc, sizeof(long long) * blocks));
  CK(cudaMalloc(&ns, sizeof(unsigned long long) * blocks));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) mix_kernel<OPA, NA, OPB, NB><<<blocks, threads>>>(out, cyc, ns, 0x00030005u, 0x5410u, iters / 4);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  mix_kernel<OPA, NA, OPB, NB><<<blocks, threads>>>(out, cyc, ns, 0x00030005u, 0x5410u, iters);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> hc(blocks); std::vector<unsigned long long> hn(blocks);
  CK(cudaMemcpy(hc.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hn.data(), ns, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost));
  double sc = 0, sn = 0; for (int i = 0; i < blocks; ++i) { sc += (double)hc[i]; sn += (double)hn[i]; }
  double cycles = sc / blocks, nsec = sn / blocks;
  const double warps_per_sm = 2.0 * threads / 32.0;
  Row r; r.na = NA; r.nb = NB;
  r.name = std::string(op_name(OPA)) + (NB ? std::string(" + ") + op_name(OPB) : std::string(""));
  r.a_per_clk_sm = (double)iters * 4 * NA * warps_per_sm / cycles;
  r.b_per_clk_sm = (double)iters * 4 * NB * warps_per_sm / cycles;
  r.tot_per_clk_sm = r.a_per_clk_sm + r.b_per_clk_sm;
  r.mhz = cycles / nsec * 1e3; r.ms = ms;
  CK(cudaFree(out)); CK(cudaFree(cyc)); CK(cudaFree(ns));
  return r;
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  std::vector<Row> rows;
  // single-op issue rates (8 independent chains per thread)
  rows.push_back(run<OP_VIADDMAX, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_VIMAX3RELU, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_VADD2, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_VIMAX, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_VIBMAX, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_PRMT, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_IADD3, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_IMAD, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_LOP3, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_FFMA, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_SHFL, 8, OP_NONE, 0>(sms, iters));
  rows.push_back(run<OP_LDS, 8, OP_NONE, 0>(sms, iters));
  // pairings: does the second op steal issue slots from the packed-integer pipe?
  rows.push_back(run<OP_VIADDMAX, 4, OP_VIMAX3RELU, 4>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 4, OP_VADD2, 4>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 4, OP_PRMT, 4>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 4, OP_IADD3, 4>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 4, OP_IMAD, 4>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 4, OP_LOP3, 4>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 4, OP_FFMA, 4>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 8, OP_SHFL, 1>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 8, OP_LDS, 1>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 8, OP_LDS, 2>(sms, iters));
  rows.push_back(run<OP_VADD2, 4, OP_IMAD, 4>(sms, iters));
  rows.push_back(run<OP_PRMT, 4, OP_IMAD, 4>(sms, iters));
  rows.push_back(run<OP_VIMAX3RELU, 4, OP_IMAD, 4>(sms, iters));
  rows.push_back(run<OP_VIADDMAX, 6, OP_IMAD, 2>(sms, iters));

  // P_int16x2: the best sustained rate over the three native packed ops the cell update uses.
  double best = 0, mhz = 0;
  for (int i = 0; i < 4; ++i) if (rows[i].tot_per_clk_sm > best) { best = rows[i].tot_per_clk_sm; mhz = rows[i].mhz; }
  double mixed = rows[12].tot_per_clk_sm;
  double p_lane_ops = best * 32.0 * sms * mhz * 1e6;          // packed lane-instr/s at the observed clock
  double p_at_max = best * 32.0 * sms * (double)p.clockRate * 1e3;
  printf("{\n \"device\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.0f, \"iters\": %d,\n", p.name, sms, p.clockRate / 1e3, iters);
  printf(" \"rows\": [\n");
  for (size_t i = 0; i < rows.size(); ++i)
    printf("  {\"mix\": \"%s\", \"nA\": %d, \"nB\": %d, \"A_warp_instr_per_clk_per_sm\": %.3f, \"B_warp_instr_per_clk_per_sm\": %.3f, \"total\": %.3f, \"sm_mhz\": %.0f, \"ms\": %.3f}%s\n",
           rows[i].name.c_str(), rows[i].na, rows[i].nb, rows[i].a_per_clk_sm, rows[i].b_per_clk_sm, rows[i].tot_per_clk_sm, rows[i].mhz, rows[i].ms,
           i + 1 < rows.size() ? "," : "");
  printf(" ],\n");
  printf(" \"packed16x2_warp_instr_per_clk_per_sm\": %.3f, \"packed16x2_mixed_warp_instr_per_clk_per_sm\": %.3f, \"observed_sm_mhz\": %.0f,\n", best, mixed, mhz);
  printf(" \"P_int16x2_lane_instr_per_s_observed_clock\": %.4e, \"P_int16x2_lane_instr_per_s_at_max_clock\": %.4e,\n", p_lane_ops, p_at_max);
  printf(" \"roofline_gcups_observed_clock\": %.1f, \"roofline_gcups_at_max_clock\": %.1f\n}\n",
         p_lane_ops * 2.0 / 7.0 / 1e9, p_at_max * 2.0 / 7.0 / 1e9);
  return 0;
}
