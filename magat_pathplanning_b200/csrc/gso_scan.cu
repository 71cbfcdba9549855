// GSO -> adjacency.  The dense graph-shift operator is only ever used as an edge mask
// (|S| > 1e-9, graphML.py:1274-1276 and :808-809); this file reads it exactly once and
// turns it into bit masks and padded neighbour lists.
#include "common.cuh"

namespace magat {

template <typename T> __device__ __forceinline__ bool is_edge(T v);
template <> __device__ __forceinline__ bool is_edge<float>(float v) { return fabsf(v) > 1e-9f; }
template <> __device__ __forceinline__ bool is_edge<double>(double v) { return fabs(v) > 1e-9; }

// One warp owns a band of 32 sender rows and sweeps it in 32x32 blocks: every load is a full
// 128 B (fp32) row segment, ballots give the row words, each lane accumulates the transposed
// (column) word of its own column.  A CTA is 8 consecutive bands.
template <typename T>
__global__ void __launch_bounds__(256) k_gso_scan(const T* __restrict__ S, int N, int W,
                                                  uint32_t* __restrict__ rowbits,
                                                  uint32_t* __restrict__ colbits) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rb = blockIdx.x * 8 + warp;
  if (rb >= W) return;
  const T* Sb = S + (size_t)b * N * N;
  const int i_own = rb * 32 + lane;
  for (int cb = 0; cb < W; ++cb) {
    const int j = cb * 32 + lane;
    uint32_t myrow = 0, mycol = 0;
    const bool jin = j < N;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const int i = rb * 32 + r;
      T v = T(0);
      if (jin && i < N) v = __ldg(Sb + (size_t)i * N + j);
      const bool e = is_edge<T>(v);
      const uint32_t word = __ballot_sync(0xffffffffu, e);
      if (lane == r) myrow = word;
      mycol |= (uint32_t)e << r;
    }
    if (i_own < N) rowbits[((size_t)b * N + i_own) * W + cb] = myrow;
    if (jin) colbits[((size_t)b * N + j) * W + rb] = mycol;
  }
}

// Vectorised scan (N % 4 == 0, 16 B aligned S): one warp per (32-row band, 128-column segment); every
// load is a full 512 B (fp32) row piece, 8 rows in flight per lane.  Row words are assembled with three
// xor-shuffles per row, column words accumulate in registers (bit r = row r of the band).
template <typename T> struct Ld4;
template <> struct Ld4<float> {
  static __device__ __forceinline__ uint32_t edges(const float* p) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
    return (uint32_t)(fabsf(v.x) > 1e-9f) | ((uint32_t)(fabsf(v.y) > 1e-9f) << 1) |
           ((uint32_t)(fabsf(v.z) > 1e-9f) << 2) | ((uint32_t)(fabsf(v.w) > 1e-9f) << 3);
  }
};
template <> struct Ld4<double> {
  static __device__ __forceinline__ uint32_t edges(const double* p) {
    const double2 a = __ldcs(reinterpret_cast<const double2*>(p));
    const double2 b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
    return (uint32_t)(fabs(a.x) > 1e-9) | ((uint32_t)(fabs(a.y) > 1e-9) << 1) | ((uint32_t)(fabs(b.x) > 1e-9) << 2) |
           ((uint32_t)(fabs(b.y) > 1e-9) << 3);
  }
};

template <typename T>
__global__ void __launch_bounds__(256) k_gso_scan_v4(const T* __restrict__ S, int N, int W, int segs, long units,
                                                     uint32_t* __restrict__ rowbits,
                                                     uint32_t* __restrict__ colbits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long unit = (long)blockIdx.x * 8 + warp;
  if (unit >= units) return;
  const int seg = (int)(unit % segs);
  const int band = (int)((unit / segs) % W);
  const long b = unit / ((long)segs * W);
  const int j0 = seg * 128 + lane * 4;
  const bool jin = j0 < N;                       // N % 4 == 0: a lane is entirely inside or outside
  const T* Sb = S + (size_t)b * N * N + j0;
  uint32_t col0 = 0, col1 = 0, col2 = 0, col3 = 0;
  const int i0 = band * 32;
#pragma unroll
  for (int rr = 0; rr < 32; rr += 8) {
    uint32_t nib[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + rr + u;
      nib[u] = (jin && i < N) ? Ld4<T>::edges(Sb + (size_t)i * N) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = rr + u;
      col0 |= (nib[u] & 1u) << r;
      col1 |= ((nib[u] >> 1) & 1u) << r;
      col2 |= ((nib[u] >> 2) & 1u) << r;
      col3 |= ((nib[u] >> 3) & 1u) << r;
      uint32_t v = nib[u] << (4 * (lane & 7));
      v |= __shfl_xor_sync(0xffffffffu, v, 1);
      v |= __shfl_xor_sync(0xffffffffu, v, 2);
      v |= __shfl_xor_sync(0xffffffffu, v, 4);
      const int i = i0 + r;
      const int w = seg * 4 + (lane >> 3);
      if ((lane & 7) == 0 && i < N && w < W) rowbits[((size_t)b * N + i) * W + w] = v;
    }
  }
  if (jin) {
    uint32_t* cb = colbits + ((size_t)b * N + j0) * W + band;
    cb[0] = col0;
    cb[W] = col1;
    cb[2 * (size_t)W] = col2;
    cb[3 * (size_t)W] = col3;
  }
}

// stats[0] max out-degree, [1] max in-degree, [2] number of edges, [3] symmetric flag.
__global__ void __launch_bounds__(256) k_gso_stats(const uint32_t* __restrict__ rowbits,
                                                   const uint32_t* __restrict__ colbits, long rows,
                                                   int W, int32_t* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  int mo = 0, mi = 0, tot = 0;
  bool sym = true;
  for (long row = warp0; row < rows; row += nwarps) {
    int od = 0, id = 0;
    for (int w = lane; w < W; w += 32) {
      const uint32_t r = rowbits[row * W + w], c = colbits[row * W + w];
      od += __popc(r);
      id += __popc(c);
      sym = sym && (r == c);
    }
    od = warp_sum_i(od);
    id = warp_sum_i(id);
    mo = max(mo, od);
    mi = max(mi, id);
    tot += od;
  }
  sym = __all_sync(0xffffffffu, sym);
  if (lane == 0) {
    atomicMax(&stats[0], mo);
    atomicMax(&stats[1], mi);
    if (tot) atomicAdd(&stats[2], tot);
    if (!sym) atomicExch(&stats[3], 0);
  }
}

// Enumerate the set bits of a W-word bit row into out[0..D) (ascending), -1 padded.
// rank_of != nullptr additionally stores, for every listed index i, the position of `self`
// inside row i of `rank_bits` (number of set bits below `self`).
__device__ __forceinline__ void list_bits(const uint32_t* __restrict__ bits, int W, int D, int lane,
                                          int32_t* __restrict__ out,
                                          const uint32_t* __restrict__ rank_bits_b, int self,
                                          int32_t* __restrict__ rank_out) {
  int base = 0;
  for (int w0 = 0; w0 < W; w0 += 32) {
    const int w = w0 + lane;
    uint32_t word = (w < W) ? bits[w] : 0u;
    const int cnt = __popc(word);
    int pos = base + warp_excl_scan_i(cnt, lane);
    base += warp_sum_i(cnt);
    while (word) {
      const int bit = __ffs(word) - 1;
      word &= word - 1;
      const int idx = w * 32 + bit;
      if (pos < D) {
        out[pos] = idx;
        if (rank_out) {
          const uint32_t* r = rank_bits_b + (size_t)idx * W;
          int rank = 0;
          const int sw = self >> 5;
          for (int q = 0; q < sw; ++q) rank += __popc(r[q]);
          rank += __popc(r[sw] & ((1u << (self & 31)) - 1u));
          rank_out[pos] = rank;
        }
      }
      ++pos;
    }
  }
  for (int s = base + lane; s < D; s += 32) {
    out[s] = -1;
    if (rank_out) rank_out[s] = 0;
  }
}

__global__ void __launch_bounds__(256) k_build_ell(const uint32_t* __restrict__ rowbits,
                                                   const uint32_t* __restrict__ colbits, long rows,
                                                   int N, int W, int D, int32_t* __restrict__ nbr_out,
                                                   int32_t* __restrict__ nbr_in,
                                                   int32_t* __restrict__ slot_in) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = row / N;
  const int n = (int)(row - b * N);
  list_bits(rowbits + row * W, W, D, lane, nbr_out + row * D, nullptr, 0, nullptr);
  list_bits(colbits + row * W, W, D, lane, nbr_in + row * D, rowbits + (size_t)b * N * W, n,
            slot_in + row * D);
}

// att[B][N][D][P] -> dense aij[B][P][N][N] (mean_heads == 0) or head-mean [B][N][N].
// `out` must be zero filled by the caller.
__global__ void __launch_bounds__(256) k_att_dense(const float* __restrict__ att,
                                                   const int32_t* __restrict__ nbr_out, long rows,
                                                   int N, int D, int P, int mean_heads,
                                                   float* __restrict__ out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * D) return;
  const long row = t / D;
  const int j = nbr_out[t];
  if (j < 0) return;
  const long b = row / N;
  const int i = (int)(row - b * N);
  const float* a = att + t * P;
  if (mean_heads) {
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += a[p];
    out[((size_t)b * N + i) * N + j] = s / (float)P;
  } else {
    for (int p = 0; p < P; ++p) out[(((size_t)b * P + p) * N + i) * N + j] = a[p];
  }
}

}  // namespace magat

using namespace magat;

extern "C" int magat_gso_scan(const void* S, int s_dtype, int B, int N, uint32_t* rowbits,
                              uint32_t* colbits, int32_t* stats, void* stream) {
  MAGAT_REQUIRE(S && rowbits && colbits && stats, MAGAT_E_BAD_ARG, "magat_gso_scan: null pointer");
  MAGAT_REQUIRE(B >= 1 && N >= 1, MAGAT_E_BAD_ARG, "magat_gso_scan: B=%d N=%d", B, N);
  MAGAT_REQUIRE(B <= 65535, MAGAT_E_BAD_ARG, "magat_gso_scan: B=%d exceeds 65535", B);
  MAGAT_REQUIRE(s_dtype == MAGAT_DT_F32 || s_dtype == MAGAT_DT_F64, MAGAT_E_BAD_ARG,
                "magat_gso_scan: GSO dtype must be fp32 or fp64");
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const int W = (N + 31) / 32;
  if (N % 4 == 0 && ((uintptr_t)S % 16) == 0) {
    const int segs = cdiv(N, 128);
    const long units = (long)B * W * segs;
    const int blocks = cdiv(units, 8);
    if (s_dtype == MAGAT_DT_F32)
      k_gso_scan_v4<float><<<blocks, 256, 0, st>>>((const float*)S, N, W, segs, units, rowbits, colbits);
    else
      k_gso_scan_v4<double><<<blocks, 256, 0, st>>>((const double*)S, N, W, segs, units, rowbits, colbits);
  } else {
    dim3 grid(cdiv(W, 8), B);
    if (s_dtype == MAGAT_DT_F32)
      k_gso_scan<float><<<grid, 256, 0, st>>>((const float*)S, N, W, rowbits, colbits);
    else
      k_gso_scan<double><<<grid, 256, 0, st>>>((const double*)S, N, W, rowbits, colbits);
  }
  int rc = check_launch("k_gso_scan", (cudaStream_t)stream);
  if (rc) return rc;
  const long rows = (long)B * N;
  int blocks = (int)min((long)148 * 8, (rows + 7) / 8);
  k_gso_stats<<<blocks, 256, 0, st>>>(rowbits, colbits, rows, W, stats);
  return check_launch("k_gso_stats", (cudaStream_t)stream);
}

extern "C" int magat_gso_build_ell(const uint32_t* rowbits, const uint32_t* colbits, int B, int N,
                                   int D, int32_t* nbr_out, int32_t* nbr_in, int32_t* slot_in,
                                   void* stream) {
  MAGAT_REQUIRE(rowbits && colbits && nbr_out && nbr_in && slot_in, MAGAT_E_BAD_ARG,
                "magat_gso_build_ell: null pointer");
  MAGAT_REQUIRE(B >= 1 && N >= 1 && D >= 1, MAGAT_E_BAD_ARG, "magat_gso_build_ell: B=%d N=%d D=%d", B, N, D);
  const long rows = (long)B * N;
  const int W = (N + 31) / 32;
  prof_begin((cudaStream_t)stream);
  k_build_ell<<<cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(rowbits, colbits, rows, N, W, D,
                                                               nbr_out, nbr_in, slot_in);
  return check_launch("k_build_ell", (cudaStream_t)stream);
}

extern "C" int magat_gat_attention_dense(const float* att, const int32_t* nbr_out, int B, int N,
                                         int D, int P, int mean_heads, float* out, void* stream) {
  MAGAT_REQUIRE(att && nbr_out && out, MAGAT_E_BAD_ARG, "magat_gat_attention_dense: null pointer");
  const long total = (long)B * N * D;
  prof_begin((cudaStream_t)stream);
  k_att_dense<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(att, nbr_out, (long)B * N, N, D, P,
                                                                 mean_heads, out);
  return check_launch("k_att_dense", (cudaStream_t)stream);
}
