// fwd_probe.cu -- times one instantiation of the forward kernels on a cfg2-shaped synthetic batch (development tool:
// lets several compile-time variants be built as separate binaries and compared in one GPU call).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DPG=4 -DPK=19 -DPT=64 [-DUBU_FWDS_...] -o fwd_probe tools/fwd_probe.cu
//   ./fwd_probe [pairs=500000] [L=76] [W=500] [reps=5]
#include "../ubu_b200/csrc/sw_fwd16s.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace ubu;
#ifndef PG
#define PG 4
#endif
#ifndef PK
#define PK 19
#endif
#ifndef PT
#define PT 64
#endif
#ifndef PGT
#define PGT 64           // threads per CTA of the general kernel
#endif
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

int main(int argc, char** argv) {
  const int pairs = argc > 1 ? atoi(argv[1]) : 500000, L = argc > 2 ? atoi(argv[2]) : 76, W = argc > 3 ? atoi(argv[3]) : 500, reps = argc > 4 ? atoi(argv[4]) : 5;
  const uint32_t n = 2u * pairs;
  const uint32_t rb = (L + 3) / 4, wb = (W + 3) / 4;
  std::vector<uint8_t> hw((size_t)pairs * wb + 16), hr((size_t)n * rb + 16);
  std::vector<uint32_t> roff(n), woff(pairs), idx(n);
  std::vector<uint16_t> rlen(n, (uint16_t)L), wlen(pairs, (uint16_t)W);
  uint64_t x = 88172645463325252ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (auto& b : hw) b = (uint8_t)rnd();
  for (int p = 0; p < pairs; ++p) woff[p] = (uint32_t)p * wb;
  for (uint32_t t = 0; t < n; ++t) {                       // read = a slice of its window with ~2 % substitutions
    roff[t] = t * rb; idx[t] = t / 2;
    const int st = (int)(rnd() % (uint64_t)(W - L + 1));
    for (int k = 0; k < L; ++k) {
      const uint8_t* w = &hw[(size_t)(t / 2) * wb];
      uint32_t c = (w[(st + k) >> 2] >> (((st + k) & 3) * 2)) & 3u;
      if (rnd() % 50 == 0) c = (c + 1 + (uint32_t)(rnd() % 3)) & 3u;
      hr[(size_t)t * rb + (k >> 2)] = (uint8_t)((hr[(size_t)t * rb + (k >> 2)] & ~(3u << ((k & 3) * 2))) | (c << ((k & 3) * 2)));
    }
  }
  uint8_t *d_w, *d_r; uint32_t *d_roff, *d_woff, *d_idx; uint16_t *d_rlen, *d_wlen; FwdOut *d_o1, *d_o2;
  CK(cudaMalloc(&d_w, hw.size())); CK(cudaMalloc(&d_r, hr.size()));
  CK(cudaMalloc(&d_roff, 4 * n)); CK(cudaMalloc(&d_woff, 4 * pairs)); CK(cudaMalloc(&d_idx, 4 * n));
  CK(cudaMalloc(&d_rlen, 2 * n)); CK(cudaMalloc(&d_wlen, 2 * pairs)); CK(cudaMalloc(&d_o1, sizeof(FwdOut) * n)); CK(cudaMalloc(&d_o2, sizeof(FwdOut) * n));
  CK(cudaMemcpy(d_w, hw.data(), hw.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_r, hr.data(), hr.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_roff, roff.data(), 4 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_woff, woff.data(), 4 * pairs, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, idx.data(), 4 * n, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_rlen, rlen.data(), 2 * n, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_wlen, wlen.data(), 2 * pairs, cudaMemcpyHostToDevice));
  DevSeqs reads{d_r, d_roff, d_rlen, nullptr, nullptr}, wins{d_w, d_woff, d_wlen, nullptr, nullptr};
  Scoring sc = make_scoring(1, -3, 5, 2);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // the kernels' epilogue also files every task into a traceback class list (TbPlan, sw_common.cuh)
  uint32_t* d_lists; unsigned int* d_counts; ubu_sw_result* d_res;
  CK(cudaMalloc(&d_lists, sizeof(uint32_t) * (size_t)TB_CLASSES * n)); CK(cudaMalloc(&d_counts, sizeof(unsigned int) * TB_CLASSES)); CK(cudaMalloc(&d_res, sizeof(ubu_sw_result) * n));
  const TbPlan plan{StripCaps{256, 200}, d_res, d_lists, d_counts, n};
  const double cells = (double)n * L * W;

  // general kernel (reference for the checksum)
  {
    constexpr int GP = PGT / PG;
    const int stride = ((W + 2 * PG + 63) / 64) * 64 + PG;
    const size_t smem = (size_t)GP * stride * 2;
    auto k = sw_fwd16_kernel<PG, PK, false, true, false, PGT>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    float best = 1e9f;
    for (int r = 0; r < reps; ++r) {
      CK(cudaMemset(d_counts, 0, sizeof(unsigned int) * TB_CLASSES));
      CK(cudaEventRecord(e0));
      k<<<(pairs + GP - 1) / GP, PGT, smem>>>(reads, wins, d_idx, nullptr, n, n, sc, d_o1, stride, nullptr, plan);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    int nb = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, PGT, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    printf("general  G=%d K=%d T=%d regs=%d blocks/SM=%d warps/SM=%d : %.3f ms  %.0f GCUPS\n", PG, PK, PGT, fa.numRegs, nb, nb * PGT / 32, best, cells / best / 1e6);
  }
  {
    constexpr int GP = PT / PG;
    const int stride = ((W + 2 * PG + 15) / 16) * 16;
    const size_t smem = FwdsLayout<PG, PK, false>::smem_bytes(GP, stride);
    auto k = sw_fwd16s_kernel<PG, PK, false, true, PT>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int nb = 0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, PT, smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k));
    float best = 1e9f;
    for (int r = 0; r < reps; ++r) {
      CK(cudaMemset(d_counts, 0, sizeof(unsigned int) * TB_CLASSES));
      CK(cudaEventRecord(e0));
      k<<<(pairs + GP - 1) / GP, PT, smem>>>(reads, wins, d_idx, nullptr, n, n, sc, d_o2, stride, plan);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("shared   G=%d K=%d T=%d regs=%d smem=%zu blocks/SM=%d warps/SM=%d : %.3f ms  %.0f GCUPS\n", PG, PK, PT, fa.numRegs, smem, nb,
           nb * PT / 32, best, cells / best / 1e6);
  }
  std::vector<FwdOut> o1(n), o2(n);
  CK(cudaMemcpy(o1.data(), d_o1, sizeof(FwdOut) * n, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o2.data(), d_o2, sizeof(FwdOut) * n, cudaMemcpyDeviceToHost));
  size_t bad = 0; double sum = 0;
  for (uint32_t t = 0; t < n; ++t) { sum += o1[t].score; if (o1[t].score != o2[t].score || o1[t].je != o2[t].je || o1[t].i_lo != o2[t].i_lo || o1[t].i_hi != o2[t].i_hi) ++bad; }
  printf("mismatching records: %zu of %u, mean score %.2f\n", bad, n, sum / n);
  return bad ? 2 : 0;
}
