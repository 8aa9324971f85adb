// peak_int16x2.cu -- measures the B200 (sm_100a) issue rate of the packed 16-bit x2 integer
// instructions the Smith-Waterman cell update is built from. This is the denominator of the
// roofline in bench.py / DESIGN.md (MEASURED_PEAKS.json has no integer figure):
//
//     P_int16x2      = packed-16x2 lane-instructions/s, whole chip, on the cell's own op mix
//     roofline_GCUPS = P_int16x2 * 2 cells / 7 ops / 1e9
//
// The 7-op cell of SURVEY.md section 8d is 4 min/max-class ops (2x VIADDMNMX, VIMNMX3.RELU, VIMNMX:
// the "ALU" pipe, 16 lanes/clk/SMSP) + 3 packed adds (VIADD.16x2). On sm_100a VIADD.16x2 issues
// down a second integer pipe, so the 4:3 mix sustains MORE than the 2 warp-instr/clk/SM of the ALU
// pipe alone. The roofline uses the measured rate of that 4:3 mix (the harder denominator); the
// ALU-pipe-only figure is printed next to it.
//
// Build : nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o peak_int16x2 peak_int16x2.cu
// Run   : ./peak_int16x2 [iters]        (prints one JSON object)
//
// Method: every thread owns one independent register chain per op instance, so there are no RAW
// stalls at 8 warps per scheduler; rate = warp-instructions per SM / (CUDA-event time x SM clock),
// SM clock measured inside the kernel as clock64 / globaltimer.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2);} } while (0)

enum { K_VIADDMAX = 0, K_VIMAX3RELU, K_VIMAX, K_VADD2, K_PRMT, K_IMAD, K_LOP3, K_FFMA, K_SHFL, NKINDS };
static const char* kind_name[NKINDS] = {"VIADDMNMX.S16x2", "VIMNMX3.S16x2.RELU", "VIMNMX.S16x2", "VIADD.16x2", "PRMT",
                                        "IMAD", "LOP3", "FFMA", "SHFL"};

template <int KIND>
__device__ __forceinline__ unsigned apply(unsigned a, unsigned b, unsigned c, int rep) {
  if constexpr (KIND == K_VIADDMAX) return __viaddmax_s16x2(a, b, c);
  else if constexpr (KIND == K_VIMAX3RELU) return __vimax3_s16x2_relu(a, b, c);
  else if constexpr (KIND == K_VIMAX) return (rep & 1) ? __vmins2(a, c) : __vmaxs2(a, b);     // alternate so it cannot fold
  else if constexpr (KIND == K_VADD2) return __vadd2(a, (rep & 1) ? b : c);
  else if constexpr (KIND == K_PRMT) return __byte_perm(a, b, c);
  else if constexpr (KIND == K_IMAD) { unsigned r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  else if constexpr (KIND == K_LOP3) { unsigned r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
  else if constexpr (KIND == K_FFMA) return __float_as_uint(fmaf(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c)));
  else return __shfl_up_sync(0xffffffffu, a, 1);
}

template <int KIND, int N>
struct Chains {
  unsigned v[N > 0 ? N : 1];
  __device__ __forceinline__ void init(unsigned seed) {
#pragma unroll
    for (int u = 0; u < N; ++u) v[u] = seed * 0x00010001u + u * 0x00070003u + KIND;
  }
  __device__ __forceinline__ void step(unsigned b, unsigned c, int rep) {
#pragma unroll
    for (int u = 0; u < N; ++u) v[u] = apply<KIND>(v[u], b, c, rep);
  }
  __device__ __forceinline__ unsigned fold() const {
    unsigned a = 0;
#pragma unroll
    for (int u = 0; u < N; ++u) a ^= v[u];
    return a;
  }
};

// C0..C8 = number of independent chains (= instructions per rep) of each kind; 4 reps per loop trip.
template <int C0, int C1, int C2, int C3, int C4, int C5, int C6, int C7, int C8>
__global__ void __launch_bounds__(512, 2)
mix_kernel(unsigned* out, long long* cyc, unsigned long long* ns, unsigned b, unsigned c, int iters) {
  Chains<0, C0> k0; Chains<1, C1> k1; Chains<2, C2> k2; Chains<3, C3> k3; Chains<4, C4> k4;
  Chains<5, C5> k5; Chains<6, C6> k6; Chains<7, C7> k7; Chains<8, C8> k8;
  const unsigned seed = threadIdx.x + 1;
  k0.init(seed); k1.init(seed); k2.init(seed); k3.init(seed); k4.init(seed); k5.init(seed); k6.init(seed); k7.init(seed); k8.init(seed);
  unsigned long long t0g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0g));
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      k0.step(b, c, rep); k3.step(b, c, rep); k1.step(b, c, rep); k2.step(b, c, rep); k4.step(b, c, rep);
      k5.step(b, c, rep); k6.step(b, c, rep); k7.step(b, c, rep); k8.step(b, c, rep);
    }
  }
  const long long t1 = clock64();
  unsigned long long t1g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1g));
  out[blockIdx.x * blockDim.x + threadIdx.x] = k0.fold() ^ k1.fold() ^ k2.fold() ^ k3.fold() ^ k4.fold() ^ k5.fold() ^ k6.fold() ^ k7.fold() ^ k8.fold();
  if (threadIdx.x == 0) { cyc[blockIdx.x] = t1 - t0; ns[blockIdx.x] = t1g - t0g; }
}

struct Row { std::string name; int n_packed, n_total; double per_clk_sm, packed_per_clk_sm, mhz, ms; };

template <int C0, int C1, int C2, int C3, int C4, int C5, int C6, int C7, int C8>
static Row run(int sms, int iters) {
  const int threads = 512, blocks = sms * 2;   // 32 warps per SM, 8 per scheduler, one wave
  unsigned* out; long long* cyc; unsigned long long* ns;
  CK(cudaMalloc(&out, sizeof(unsigned) * threads * blocks));
  CK(cudaMalloc(&cyc, sizeof(long long) * blocks));
  CK(cudaMalloc(&ns, sizeof(unsigned long long) * blocks));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto kern = mix_kernel<C0, C1, C2, C3, C4, C5, C6, C7, C8>;
  for (int w = 0; w < 2; ++w) kern<<<blocks, threads>>>(out, cyc, ns, 0x00030005u, 0x5410u, iters / 4);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  kern<<<blocks, threads>>>(out, cyc, ns, 0x00030005u, 0x5410u, iters);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> hc(blocks); std::vector<unsigned long long> hn(blocks);
  CK(cudaMemcpy(hc.data(), cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hn.data(), ns, sizeof(unsigned long long) * blocks, cudaMemcpyDeviceToHost));
  double sc = 0, sn = 0; for (int i = 0; i < blocks; ++i) { sc += (double)hc[i]; sn += (double)hn[i]; }
  const double mhz = sc / sn * 1e3;
  const int counts[NKINDS] = {C0, C1, C2, C3, C4, C5, C6, C7, C8};
  Row r; r.n_total = 0; r.n_packed = C0 + C1 + C2 + C3;
  for (int k = 0; k < NKINDS; ++k) if (counts[k]) {
    r.n_total += counts[k];
    r.name += (r.name.empty() ? "" : " + ") + std::to_string(counts[k]) + "x " + kind_name[k];
  }
  const double warps_per_sm = 2.0 * threads / 32.0;
  const double clks = (double)ms * 1e-3 * mhz * 1e6;              // wall time x measured SM clock
  r.per_clk_sm = (double)iters * 4 * r.n_total * warps_per_sm / clks;
  r.packed_per_clk_sm = (double)iters * 4 * r.n_packed * warps_per_sm / clks;
  r.mhz = mhz; r.ms = ms;
  CK(cudaFree(out)); CK(cudaFree(cyc)); CK(cudaFree(ns));
  return r;
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  std::vector<Row> rows;
  //                 VIADDMAX VIMAX3 VIMAX VADD2 PRMT IMAD LOP3 FFMA SHFL
  Row sw_mix = run<2, 1, 1, 3, 0, 0, 0, 0, 0>(sms, iters);      // the 7-op cell of SURVEY section 8d: 4 ALU + 3 VIADD
  rows.push_back(sw_mix);
  Row alu4 = run<4, 2, 2, 0, 0, 0, 0, 0, 0>(sms, iters);        // min/max-class ops only (one pipe)
  rows.push_back(alu4);
  rows.push_back(run<8, 0, 0, 0, 0, 0, 0, 0, 0>(sms, iters));
  rows.push_back(run<0, 8, 0, 0, 0, 0, 0, 0, 0>(sms, iters));
  rows.push_back(run<0, 0, 8, 0, 0, 0, 0, 0, 0>(sms, iters));
  rows.push_back(run<0, 0, 0, 8, 0, 0, 0, 0, 0>(sms, iters));
  rows.push_back(run<0, 0, 0, 0, 8, 0, 0, 0, 0>(sms, iters));
  rows.push_back(run<0, 0, 0, 0, 0, 8, 0, 0, 0>(sms, iters));
  rows.push_back(run<0, 0, 0, 0, 0, 0, 8, 0, 0>(sms, iters));
  rows.push_back(run<0, 0, 0, 0, 0, 0, 0, 8, 0>(sms, iters));
  rows.push_back(run<0, 0, 0, 0, 0, 0, 0, 0, 8>(sms, iters));
  rows.push_back(run<4, 0, 0, 4, 0, 0, 0, 0, 0>(sms, iters));   // VIADDMNMX : VIADD.16x2 1:1
  rows.push_back(run<4, 0, 0, 0, 4, 0, 0, 0, 0>(sms, iters));   // + PRMT
  rows.push_back(run<4, 0, 0, 0, 0, 4, 0, 0, 0>(sms, iters));   // + IMAD
  rows.push_back(run<0, 0, 0, 4, 0, 4, 0, 0, 0>(sms, iters));   // VIADD.16x2 + IMAD
  rows.push_back(run<0, 0, 0, 0, 4, 4, 0, 0, 0>(sms, iters));   // PRMT + IMAD
  rows.push_back(run<2, 2, 1, 2, 2, 0, 0, 0, 0>(sms, iters));   // this repo's kernel mix per packed cell (x2): see DESIGN.md
  rows.push_back(run<6, 0, 0, 0, 0, 0, 0, 0, 1>(sms, iters));   // + SHFL

  const double mix_rate = sw_mix.packed_per_clk_sm, mhz = sw_mix.mhz;
  const double p_obs = mix_rate * 32.0 * sms * mhz * 1e6;
  const double p_alu = alu4.packed_per_clk_sm * 32.0 * sms * alu4.mhz * 1e6;
  printf("{\n \"device\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.0f, \"iters\": %d,\n", p.name, sms, p.clockRate / 1e3, iters);
  printf(" \"rows\": [\n");
  for (size_t i = 0; i < rows.size(); ++i)
    printf("  {\"mix\": \"%s\", \"warp_instr_per_clk_per_sm\": %.3f, \"packed16x2_per_clk_per_sm\": %.3f, \"sm_mhz\": %.0f, \"ms\": %.3f}%s\n",
           rows[i].name.c_str(), rows[i].per_clk_sm, rows[i].packed_per_clk_sm, rows[i].mhz, rows[i].ms, i + 1 < rows.size() ? "," : "");
  printf(" ],\n");
  printf(" \"sw_mix_warp_instr_per_clk_per_sm\": %.3f, \"alu_only_warp_instr_per_clk_per_sm\": %.3f, \"observed_sm_mhz\": %.0f,\n",
         mix_rate, alu4.packed_per_clk_sm, mhz);
  printf(" \"P_int16x2_lane_instr_per_s\": %.4e, \"P_int16x2_alu_only_lane_instr_per_s\": %.4e,\n", p_obs, p_alu);
  printf(" \"roofline_gcups\": %.1f, \"alu_only_roofline_gcups\": %.1f\n}\n", p_obs * 2.0 / 7.0 / 1e9, p_alu * 2.0 / 7.0 / 1e9);
  return 0;
}
