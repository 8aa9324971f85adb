#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
// RF-bandwidth probe: packed ops with DISTINCT register operands (no reuse), alone and mixed with VIADD.16x2
template <int NA, int NB, int MODE>
__global__ void __launch_bounds__(256, 2) k(unsigned* out, unsigned seed, int iters, long long* cyc) {
  unsigned a[NA > 0 ? NA : 1], x[NA > 0 ? NA : 1], y[NA > 0 ? NA : 1], d[NB > 0 ? NB : 1], z[NB > 0 ? NB : 1];
#pragma unroll
  for (int u = 0; u < NA; ++u) { a[u] = threadIdx.x * 3 + u * seed; x[u] = seed * (u + 7); y[u] = seed ^ (u * 0x10001u); }
#pragma unroll
  for (int u = 0; u < NB; ++u) { d[u] = threadIdx.x + u; z[u] = seed + u * 5; }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      constexpr int N = NA > NB ? NA : NB;
#pragma unroll
      for (int u = 0; u < N; ++u) {
        if (u < NA) {
          if (MODE == 0) a[u] = __viaddmax_s16x2(a[u], x[u], y[u]);           // 3 distinct regs
          else if (MODE == 1) a[u] = __viaddmax_s16x2(a[u], x[0], y[u]);      // shared b
          else a[u] = __vimax3_s16x2_relu(a[u], x[u], y[(u + 1) % NA]);
        }
        if (u < NB) d[u] = __vadd2(d[u], (rep & 1) ? z[u] : z[(u + 1) % (NB > 0 ? NB : 1)]);
      }
    }
  }
  long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int u = 0; u < NA; ++u) acc ^= a[u] ^ x[u] ^ y[u];
#pragma unroll
  for (int u = 0; u < NB; ++u) acc ^= d[u] ^ z[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int NA, int NB, int MODE> void run(const char* name, int warps_per_sm) {
  int sms = 148, iters = 20000; unsigned* out; long long* cyc;
  int threads = 256, blocks_per_sm = warps_per_sm * 32 / threads; int blocks = sms * blocks_per_sm;
  cudaMalloc(&out, 4 * threads * blocks); cudaMalloc(&cyc, 8 * blocks);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NA, NB, MODE><<<blocks, threads>>>(out, 12345u, iters / 4, cyc); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<NA, NB, MODE><<<blocks, threads>>>(out, 12345u, iters, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double clk = ms * 1e-3 * 1.965e9; double instr = (double)iters * 4 * (NA + NB) * warps_per_sm;
  printf("%-50s warps/SM=%2d  %.3f instr/clk/SM (A %.3f, B %.3f)  %.3f ms\n", name, warps_per_sm, instr / clk, instr / clk * NA / (NA + NB), instr / clk * NB / (NA + NB), ms);
}
int main() {
  run<8, 0, 0>("VIADDMNMX 3 distinct regs", 32);
  run<8, 0, 1>("VIADDMNMX shared b", 32);
  run<8, 0, 2>("VIMNMX3.RELU 3 distinct", 32);
  run<8, 6, 0>("8 VIADDMNMX(3 distinct) + 6 VIADD.16x2 (2 distinct)", 32);
  run<8, 6, 1>("8 VIADDMNMX(shared b) + 6 VIADD.16x2", 32);
  run<8, 6, 0>("8 VIADDMNMX(3 distinct) + 6 VIADD.16x2 (2 distinct)", 16);
  run<8, 6, 0>("8 VIADDMNMX(3 distinct) + 6 VIADD.16x2 (2 distinct)", 8);
  run<8, 8, 0>("8 VIADDMNMX(3 distinct) + 8 VIADD.16x2", 16);
  run<8, 4, 0>("8 VIADDMNMX(3 distinct) + 4 VIADD.16x2", 16);
  run<8, 0, 0>("VIADDMNMX 3 distinct regs", 16);
  run<8, 0, 0>("VIADDMNMX 3 distinct regs", 8);
  return 0;
}
