// Host side of the GSO ingest: the dense graph-shift operator usually sits in HOST memory (the reference's dataloader
// and simulator build it on the CPU, dataloader/Dataloader_dcplocal_notTF_onlineExpert.py:172, utils/new_simulator.py:317)
// and the layer only ever uses it as an edge mask (|s| > 1e-9, graphML.py:1274-1276).  Shipping 4 N^2 bytes per instance
// over PCIe to test them on the device makes the end-to-end step PCIe bound (2.3 GB per 512 x 1000-agent batch); this
// packs the mask on the host cores instead -- one streaming pass, multi-threaded, AVX2 where the CPU has it -- so that
// N^2 / 8 bytes cross the link.  Plain C++ (no CUDA); the device side continues with magat_gso_from_rowbits.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/magat_gat.h"

namespace {

template <typename T> inline bool edge(T v) { return fabs((double)v) > 1e-9; }
template <> inline bool edge<float>(float v) { return fabsf(v) > 1e-9f; }

template <typename T>
void pack_rows_scalar(const T* S, long r0, long r1, int N, int W, uint32_t* bits) {
  for (long r = r0; r < r1; ++r) {
    const T* row = S + r * (long)N;
    uint32_t* out = bits + r * (long)W;
    for (int w = 0; w < W; ++w) {
      uint32_t word = 0;
      const int j1 = std::min(N, (w + 1) * 32);
      for (int j = w * 32; j < j1; ++j) word |= (uint32_t)edge<T>(row[j]) << (j & 31);
      out[w] = word;
    }
  }
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) void pack_rows_avx2_f32(const float* S, long r0, long r1, int N, int W, uint32_t* bits) {
  const __m256 thr = _mm256_set1_ps(1e-9f);
  const __m256 absmask = _mm256_castsi256_ps(_mm256_set1_epi32(0x7fffffff));
  const int full = N / 32;
  for (long r = r0; r < r1; ++r) {
    const float* row = S + r * (long)N;
    uint32_t* out = bits + r * (long)W;
    for (int w = 0; w < full; ++w) {
      const float* q = row + w * 32;
      uint32_t word = 0;
      for (int k = 0; k < 4; ++k) {
        const __m256 v = _mm256_and_ps(_mm256_loadu_ps(q + 8 * k), absmask);
        word |= (uint32_t)_mm256_movemask_ps(_mm256_cmp_ps(v, thr, _CMP_GT_OQ)) << (8 * k);   // ordered: NaN is "no edge"
      }
      out[w] = word;
    }
    if (full < W) {
      uint32_t word = 0;
      for (int j = full * 32; j < N; ++j) word |= (uint32_t)(fabsf(row[j]) > 1e-9f) << (j & 31);
      out[full] = word;
    }
  }
}
__attribute__((target("avx2"))) void pack_rows_avx2_f64(const double* S, long r0, long r1, int N, int W, uint32_t* bits) {
  const __m256d thr = _mm256_set1_pd(1e-9);
  const __m256d absmask = _mm256_castsi256_pd(_mm256_set1_epi64x(0x7fffffffffffffffll));
  const int full = N / 32;
  for (long r = r0; r < r1; ++r) {
    const double* row = S + r * (long)N;
    uint32_t* out = bits + r * (long)W;
    for (int w = 0; w < full; ++w) {
      const double* q = row + w * 32;
      uint32_t word = 0;
      for (int k = 0; k < 8; ++k) {
        const __m256d v = _mm256_and_pd(_mm256_loadu_pd(q + 4 * k), absmask);
        word |= (uint32_t)_mm256_movemask_pd(_mm256_cmp_pd(v, thr, _CMP_GT_OQ)) << (4 * k);
      }
      out[w] = word;
    }
    if (full < W) {
      uint32_t word = 0;
      for (int j = full * 32; j < N; ++j) word |= (uint32_t)(fabs(row[j]) > 1e-9) << (j & 31);
      out[full] = word;
    }
  }
}
#endif

template <typename T>
void pack_rows(const T* S, long r0, long r1, int N, int W, uint32_t* bits, bool avx2) {
#if defined(__x86_64__)
  if (avx2) {
    if (sizeof(T) == 4) pack_rows_avx2_f32(reinterpret_cast<const float*>(S), r0, r1, N, W, bits);
    else pack_rows_avx2_f64(reinterpret_cast<const double*>(S), r0, r1, N, W, bits);
    return;
  }
#endif
  pack_rows_scalar<T>(S, r0, r1, N, W, bits);
}

}  // namespace

// rowbits_host[row][w]: bit j % 32 of word j / 32 set iff |S[row][j]| > 1e-9 (same layout as magat_gso_scan's rowbits);
// rows = B * N.  threads <= 0: one per hardware thread (capped at 64).  Returns 0, or MAGAT_E_BAD_ARG.
extern "C" int magat_gso_pack_host(const void* S_host, int s_dtype, long rows, int N, uint32_t* rowbits_host, int threads) {
  if (S_host == nullptr || rowbits_host == nullptr || rows < 1 || N < 1 ||
      (s_dtype != MAGAT_DT_F32 && s_dtype != MAGAT_DT_F64))
    return MAGAT_E_BAD_ARG;
  const int W = (N + 31) / 32;
  bool avx2 = false;
#if defined(__x86_64__)
  avx2 = __builtin_cpu_supports("avx2");
#endif
  int nt = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if (nt > 64) nt = 64;
  if ((long)nt > rows) nt = (int)rows;
  auto work = [&](int t) {
    const long r0 = rows * t / nt, r1 = rows * (t + 1) / nt;
    if (s_dtype == MAGAT_DT_F32) pack_rows<float>((const float*)S_host, r0, r1, N, W, rowbits_host, avx2);
    else pack_rows<double>((const double*)S_host, r0, r1, N, W, rowbits_host, avx2);
  };
  if (nt == 1) {
    work(0);
    return MAGAT_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(nt - 1);
  for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  return MAGAT_OK;
}
