// Generic fp32 SIMT tile GEMMs used by the baseline (non tensor-core) path and by every
// shape the tcgen05 path does not cover.  Operands and results are described by functors so
// the same two kernels serve all node-wise projections and all weight-gradient reductions.
#pragma once
#include "common.cuh"

namespace magat {

// C(m, n, z) = sum_{k < Kred} A(m, k, z) * B(k, n, z), m < M, n < Ncols, z = blockIdx.z.
// 64x64 tile per CTA, 256 threads, 4x4 outputs per thread, k-tile of 16.
template <class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__(256) k_node_gemm(long M, int Ncols, int Kred, ALoad A, BLoad Bm,
                                                   Epi epi) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const long m0 = (long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int z = blockIdx.z;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Kred; k0 += 16) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int idx = tid + t * 256;
      const int kk = idx & 15, r = idx >> 4;
      const int k = k0 + kk;
      const long m = m0 + r;
      const int n = n0 + r;
      As[kk][r] = (k < Kred && m < M) ? A(m, k, z) : 0.f;
      Bs[kk][r] = (k < Kred && n < Ncols) ? Bm(k, n, z) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < Ncols) epi(m, n, z, acc[i][j]);
    }
  }
}

// Reduction over rows ("weight gradient"):
//   C(i, j, z) = sum_{r < R} A(r, i, z) * B(r, j, z),  i < Mi, j < Nj.
// grid = (ceil(Mi/64), ceil(Nj/64), Z * splits); the row range is cut into `splits` chunks and
// each chunk writes its own partial tile (deterministic two-pass reduction, no atomics):
//   partial[((z * splits + s) * Mi + i) * Nj + j]
template <class ALoad, class BLoad>
__global__ void __launch_bounds__(256) k_rowred_gemm(long R, int Mi, int Nj, int splits, ALoad A,
                                                     BLoad Bm, float* __restrict__ partial) {
  __shared__ float As[16][68];
  __shared__ float Bs[16][68];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  const int z = blockIdx.z / splits, s = blockIdx.z - z * splits;
  const long chunk = (R + splits - 1) / splits;
  const long r_begin = (long)s * chunk;
  const long r_end = min(R, r_begin + chunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long r0 = r_begin; r0 < r_end; r0 += 16) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int idx = tid + t * 256;
      const int c = idx & 63, rr = idx >> 6;
      const long r = r0 + rr;
      As[rr][c] = (r < r_end && i0 + c < Mi) ? A(r, i0 + c, z) : 0.f;
      Bs[rr][c] = (r < r_end && j0 + c < Nj) ? Bm(r, j0 + c, z) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + ((size_t)(z * splits + s) * Mi) * Nj;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ii = i0 + ty * 4 + i;
    if (ii >= Mi) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = j0 + tx * 4 + j;
      if (jj < Nj) out[(size_t)ii * Nj + jj] = acc[i][j];
    }
  }
}

// out[e] = scale * sum_s partial[(zmap(e) * splits + s) * per + local(e)] for a [Z][per] result.
static __global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ partial, int splits,
                                                         long per, long total, float* __restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long z = e / per, l = e - z * per;
  const float* p = partial + (size_t)z * splits * per + l;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += p[(size_t)k * per];
  out[e] = s;
}

}  // namespace magat
