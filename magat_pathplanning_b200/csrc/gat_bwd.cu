// Backward of the batched graph-attention layer (fp32 SIMT; sparse parts are edge-wise).
//
// Given dY (same logical shape as y), with dP = dY * act'(y) (mean mode: / P, shared by heads):
//   dH_{p,k} = dP_p^T u_k^p ;  dbias = sum dP ;  gU_k = dP_p H_{p,k}  (per node, [K][G])
//   for k = K-1 .. 1:  dA_p[i,j] += u_{k-1}[i] . gU_k[j] ;  gU_{k-1}[i] += sum_j A_p[i,j] gU_k[j]
//   dx += sum_p gU_0 ;  de = A * (dA - rowsum(A * dA))
//   KeyQuery:     dR_i = sum_j de[i,j] x_j ; dx_j += sum_i de[i,j] R_i ; dx_i += W dR_i ; dW = sum_i x_i dR_i^T
//   GAT_modified: ds = de * lrelu'(s) ; r_i = rowsum, c_j = colsum ; dx_n += c_n cvec_0 + r_n cvec_1 ;
//                 dcvec_t = sum_n {c,r}_n x_n ; ddvec_t = sum_n {c,r}_n ; then the chain through
//                 cvec_t = W^T a_t, dvec_t = a_t . wb  (gat_fwd.cu k_gm_prep)
// mixer / weight_bias receive no gradient in KeyQuery mode (the reference leaves grad = None).
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"
#include "simt_gemm.cuh"

namespace magat {

struct DPre {   // dP(m, p, f)
  const float* y; long y_sb, y_sn, y_sc; const float* dy; long dy_sb, dy_sn, dy_sc;
  int N, F, concat, relu; float scale;
  __device__ __forceinline__ float operator()(long m, int p, int f) const {
    const long b = m / N;
    const long n = m - b * N;
    const long c = concat ? (long)p * F + f : f;
    const float g = __ldg(dy + b * dy_sb + n * dy_sn + c * dy_sc) * scale;
    if (!relu) return g;
    return __ldg(y + b * y_sb + n * y_sn + c * y_sc) > 0.f ? g : 0.f;
  }
};

struct ZNode {   // u_k^p of node m, feature g
  const float* x; long x_sb, x_sn; const float* taps; int N, G, K, P;
  __device__ __forceinline__ float operator()(long m, int p, int k, int g) const {
    if (k == 0) {
      const long b = m / N;
      return __ldg(x + b * x_sb + (m - b * N) * x_sn + g);
    }
    return __ldg(taps + (((size_t)m * P + p) * (K - 1) + (k - 1)) * G + g);
  }
};

// gz[m][p][k][g] = sum_f dP(m,p,f) H[p][f][k][g]
struct DPreA { DPre d; __device__ __forceinline__ float operator()(long m, int f, int z) const { return d(m, z, f); } };
struct HtLoad {
  const float* H; int KG;
  __device__ __forceinline__ float operator()(int f, int n, int z) const {
    return __ldg(H + ((size_t)z * gridF + f) * KG + n);
  }
  int gridF;
};
struct GzEpi {
  float* gz; int P, KG;
  __device__ __forceinline__ void operator()(long m, int n, int z, float v) const { gz[((size_t)m * P + z) * KG + n] = v; }
};
// dH[p][f][kg] = sum_m dP(m,p,f) Z(m,p,kg)
struct DPreR { DPre d; __device__ __forceinline__ float operator()(long r, int f, int z) const { return d(r, z, f); } };
struct ZRed {
  ZNode zn; int G;
  __device__ __forceinline__ float operator()(long r, int kg, int z) const {
    const int k = kg / G;
    return zn(r, z, k, kg - k * G);
  }
};
// dbias[f] = sum_{m,p} dP(m,p,f)
struct OneLoad { __device__ __forceinline__ float operator()(long, int, int) const { return 1.f; } };
struct DPreSumP {
  DPre d; int P;
  __device__ __forceinline__ float operator()(long r, int f, int) const {
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += d(r, p, f);
    return s;
  }
};

// ---- tap recursion backward, level k (k = K-1 .. 1) ---------------------------------------
// warp per sender row i: dA[i,j] (+)= u_{k-1}[i] . gU_k[j];  gU_{k-1}[i] += sum_j A[i,j] gU_k[j].
__global__ void __launch_bounds__(256) k_tap_bwd(ZNode zn, const float* __restrict__ att,
                                                 const int32_t* __restrict__ nbr_out, long rows, int N, int G,
                                                 int P, int K, int D, int k, int first,
                                                 float* __restrict__ gz, float* __restrict__ datt) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = row / N;
  const int32_t* nb = nbr_out + row * D;
  for (int p = 0; p < P; ++p) {
    float* gdst = gz + (((size_t)row * P + p) * K + (k - 1)) * G;
    for (int s = 0; s < D; ++s) {
      const int j = nb[s];
      if (j < 0) break;
      const float a = att[((size_t)row * D + s) * P + p];
      const float* gsrc = gz + ((((size_t)(b * N + j)) * P + p) * K + k) * G;
      float d = 0.f;
      for (int g = lane; g < G; g += 32) {
        const float gv = gsrc[g];
        d = fmaf(zn(row, p, k - 1, g), gv, d);
        gdst[g] = fmaf(a, gv, gdst[g]);
      }
      d = warp_sum(d);
      if (lane == 0) {
        float* da = datt + ((size_t)row * D + s) * P + p;
        *da = first ? d : *da + d;
      }
    }
  }
}

// ---- softmax backward per row + the row-side score gradients ---------------------------------
// datt <- de (KeyQuery) or ds (GAT_modified), in place.
template <int MODE>
__global__ void __launch_bounds__(256) k_softmax_bwd(const float* __restrict__ x, long x_sb, long x_sn,
                                                     const float* __restrict__ sproj,
                                                     const float* __restrict__ att,
                                                     const int32_t* __restrict__ nbr_out, long rows, int N,
                                                     int G, int P, int D, int has_datt,
                                                     float* __restrict__ datt, float* __restrict__ rc) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = row / N;
  const int32_t* nb = nbr_out + row * D;
  const float* arow = att + row * (size_t)D * P;
  float* drow = datt + row * (size_t)D * P;
  for (int p = 0; p < P; ++p) {
    float dot = 0.f;
    int deg = 0;
    for (int s0 = 0; s0 < D; s0 += 32) {
      const int s = s0 + lane;
      const bool v = s < D && nb[s] >= 0;
      if (v) {
        if (!has_datt) drow[(size_t)s * P + p] = 0.f;
        dot = fmaf(arow[(size_t)s * P + p], drow[(size_t)s * P + p], dot);
      }
      deg += __popc(__ballot_sync(0xffffffffu, v));
    }
    dot = warp_sum(dot);
    float rsum = 0.f;
    const float si = (MODE == MAGAT_MODE_GAT_MODIFIED) ? sproj[((size_t)row * P + p) * 2 + 1] : 0.f;
    for (int s = lane; s < deg; s += 32) {
      float de = arow[(size_t)s * P + p] * (drow[(size_t)s * P + p] - dot);
      if (MODE == MAGAT_MODE_GAT_MODIFIED) {
        const float sr = si + sproj[((size_t)(b * N + nb[s]) * P + p) * 2 + 0];
        de *= sr > 0.f ? 1.f : kLeaky;
        rsum += de;
      }
      drow[(size_t)s * P + p] = de;
    }
    if (MODE == MAGAT_MODE_GAT_MODIFIED) {
      rsum = warp_sum(rsum);
      if (lane == 0) rc[((size_t)row * P + p) * 2 + 1] = rsum;
    } else {
      __syncwarp();
      // dR_i^p = sum_j de[i,j] x_j
      for (int g = lane; g < G; g += 32) {
        float acc = 0.f;
        for (int s = 0; s < deg; ++s)
          acc = fmaf(drow[(size_t)s * P + p], x[b * x_sb + (long)nb[s] * x_sn + g], acc);
        rc[((size_t)row * P + p) * G + g] = acc;
      }
    }
  }
}

// ---- column side: dx_j = sum_p gU_0[j] + score terms gathered over the in-neighbours -----------
template <int MODE>
__global__ void __launch_bounds__(256) k_col_bwd(const float* __restrict__ gz, const float* __restrict__ datt,
                                                 const float* __restrict__ sproj, const float* __restrict__ cvec,
                                                 const int32_t* __restrict__ nbr_in,
                                                 const int32_t* __restrict__ slot_in, long rows, int N, int G,
                                                 int P, int K, int D, float* __restrict__ rc,
                                                 float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = row / N;
  const int32_t* nb = nbr_in + row * D;
  const int32_t* sl = slot_in + row * D;
  if (MODE == MAGAT_MODE_GAT_MODIFIED) {
    for (int p = 0; p < P; ++p) {
      float c = 0.f;
      for (int s = lane; s < D; s += 32) {
        const int i = nb[s];
        if (i >= 0) c += datt[((size_t)(b * N + i) * D + sl[s]) * P + p];
      }
      c = warp_sum(c);
      if (lane == 0) rc[((size_t)row * P + p) * 2 + 0] = c;
    }
    __syncwarp();
  }
  if (dx == nullptr) return;
  for (int g = lane; g < G; g += 32) {
    float acc = 0.f;
    for (int p = 0; p < P; ++p) {
      acc += gz[(((size_t)row * P + p) * K + 0) * G + g];
      if (MODE == MAGAT_MODE_KEYQUERY) {
        for (int s = 0; s < D; ++s) {
          const int i = nb[s];
          if (i < 0) break;
          const long ri = b * N + i;
          acc = fmaf(datt[((size_t)ri * D + sl[s]) * P + p], sproj[((size_t)ri * P + p) * G + g], acc);
        }
      } else {
        acc = fmaf(rc[((size_t)row * P + p) * 2 + 0], cvec[((size_t)p * 2 + 0) * G + g], acc);
        acc = fmaf(rc[((size_t)row * P + p) * 2 + 1], cvec[((size_t)p * 2 + 1) * G + g], acc);
      }
    }
    dx[(size_t)row * G + g] = acc;
  }
}

// ---- vectorised variants (G == 128, D <= 32, P in {1,2,4}, 16 B aligned rows) --------------------------
__device__ __forceinline__ void bfma4(float4& acc, float a, const float4& v) {
  acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y); acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
}
__device__ __forceinline__ float bdot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
template <int PT>
__device__ __forceinline__ void ldp(const float* p, float* a) {
  if (PT == 4) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    a[0] = v.x; a[1 % PT] = v.y; a[2 % PT] = v.z; a[3 % PT] = v.w;
  } else {
#pragma unroll
    for (int q = 0; q < PT; ++q) a[q] = p[q];
  }
}
template <int PT>
__device__ __forceinline__ void stp(float* p, const float* a) {
  if (PT == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1 % PT], a[2 % PT], a[3 % PT]);
  } else {
#pragma unroll
    for (int q = 0; q < PT; ++q) p[q] = a[q];
  }
}

// Tap recursion backward, level k: one warp per sender row, lane l owns features 4l..4l+3 and out-slot l.  HP heads
// at a time (PT / HP passes over the row's edges), EF edges per step.  (A first version kept all heads of two edges in
// registers: 80 registers plus an 80 B spill frame -- ncu: a third of its L2 traffic was local memory -- at three CTAs
// per SM, and reduced every (edge, head) dot product on its own: 2.33 ms for both levels at B = 512, N = 1000.)
// Half the heads need half the accumulators -- 64 registers, four CTAs per SM -- and the EF x HP dot products of a
// step share one joint reduction.
// Everything a pass needs that does not depend on the neighbour list is requested up front: the accumulators START
// as the row's own plane k-1 (the read of its read-modify-write) and dsum as the datt it adds to.
template <int PT, int HP, int EF, int MINB>
__global__ void __launch_bounds__(256, MINB) k_tap_bwd_p(const float* __restrict__ x, long x_sb, long x_sn,
                                                      const float* __restrict__ taps, const float* __restrict__ att,
                                                      const int32_t* __restrict__ nbr_out, long rows, int N, int K, int D,
                                                      int k, int first, float* __restrict__ gz, float* __restrict__ datt,
                                                      float* __restrict__ g0sum, float* __restrict__ rc_out,
                                                      const float* __restrict__ gm_sproj, float* __restrict__ gm_rc) {
  constexpr int G = 128, NV = EF * HP;
  constexpr int SH = NV == 16 ? 1 : NV == 8 ? 2 : NV == 4 ? 3 : 4;      // lane l holds the total of value l >> SH
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = batch_of32(row, N);
  const int g0 = lane * 4;
  const int my_j = lane < D ? nbr_out[row * D + lane] : -1;
  const int deg = __popc(__ballot_sync(0xffffffffu, my_j >= 0));
  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);            // g0sum: head sum of g_0
#pragma unroll 1
  for (int h0 = 0; h0 < PT; h0 += HP) {
    float* da = datt + ((size_t)row * D + lane) * PT + h0;
    float am[HP], dsum[HP];
    float4 u[HP], acc[HP];
#pragma unroll
    for (int q = 0; q < HP; ++q) {
      am[q] = my_j >= 0 ? att[((size_t)row * D + lane) * PT + h0 + q] : 0.f;
      dsum[q] = (!first && lane < D) ? da[q] : 0.f;
      acc[q] = *reinterpret_cast<const float4*>(gz + (((size_t)row * PT + h0 + q) * K + (k - 1)) * G + g0);
      u[q] = (k == 1) ? __ldg(reinterpret_cast<const float4*>(x + b * x_sb + (row - b * N) * x_sn + g0))
                      : __ldg(reinterpret_cast<const float4*>(taps + (((size_t)row * PT + h0 + q) * (K - 1) + (k - 2)) * G + g0));
    }
    for (int s = 0; s < deg; s += EF) {
      float4 gv[EF][HP];
#pragma unroll
      for (int w = 0; w < EF; ++w) {
        const int j = __shfl_sync(0xffffffffu, my_j, (s + w) & 31);
#pragma unroll
        for (int q = 0; q < HP; ++q)
          gv[w][q] = s + w < deg ? *reinterpret_cast<const float4*>(gz + ((((size_t)(b * N + j)) * PT + h0 + q) * K + k) * G + g0)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float d[NV];
#pragma unroll
      for (int w = 0; w < EF; ++w)
#pragma unroll
        for (int q = 0; q < HP; ++q) {
          d[w * HP + q] = bdot4(u[q], gv[w][q]);
          bfma4(acc[q], __shfl_sync(0xffffffffu, am[q], (s + w) & 31), gv[w][q]);
        }
      const float tot = warp_multi_sum<NV>(d, lane);
      // lane s + w takes the HP totals of edge w
      const int wme = (lane - s) & (EF - 1);
      const bool mine = lane >= s && lane < s + EF;
#pragma unroll
      for (int q = 0; q < HP; ++q) {
        const float dq = __shfl_sync(0xffffffffu, tot, (wme * HP + q) << SH);
        if (mine) dsum[q] += dq;
      }
    }
    if (g0sum != nullptr) {
#pragma unroll
      for (int q = 0; q < HP; ++q) { t.x += acc[q].x; t.y += acc[q].y; t.z += acc[q].z; t.w += acc[q].w; }
    } else {
#pragma unroll
      for (int q = 0; q < HP; ++q)
        *reinterpret_cast<float4*>(gz + (((size_t)row * PT + h0 + q) * K + (k - 1)) * G + g0) = acc[q];
    }
    if (gm_rc != nullptr) {
      // Last level, GAT_modified: dA of this row is complete, so its softmax + LeakyReLU backward follows here (same
      // arithmetic as k_softmax_bwd_gm_v, which then need not run): datt <- ds, rc[row][p][1] = sum_j ds[i,j].
#pragma unroll
      for (int q = 0; q < HP; ++q) {
        const float o = lane < D ? dsum[q] : 0.f;
        const float dot = warp_sum(am[q] * o);
        float de = am[q] * (o - dot);
        if (my_j >= 0) {
          const float sr = gm_sproj[((size_t)row * PT + h0 + q) * 2 + 1] +
                           gm_sproj[((size_t)(b * N + my_j) * PT + h0 + q) * 2 + 0];
          de *= sr > 0.f ? 1.f : kLeaky;
        }
        const float rsum = warp_sum(de);
        if (lane == 0) gm_rc[((size_t)row * PT + h0 + q) * 2 + 1] = rsum;
        if (lane < D) da[q] = de;
      }
      continue;
    }
    if (rc_out == nullptr) {
      if (lane < D && (first || deg > 0)) {
#pragma unroll
        for (int q = 0; q < HP; ++q) da[q] = dsum[q];
      }
      continue;
    }
    // Last level, KeyQuery: dA of this row is complete, so its softmax backward and dR_i = sum_j de[i,j] x_j follow
    // here (same arithmetic as k_softmax_bwd_kq_v, which then need not run).  datt <- de.
    float de[HP];
#pragma unroll
    for (int q = 0; q < HP; ++q) {
      const float o = lane < D ? dsum[q] : 0.f;
      const float dot = warp_sum(am[q] * o);
      de[q] = am[q] * (o - dot);
      if (lane < D) da[q] = de[q];
    }
    float4 racc[HP];
#pragma unroll
    for (int q = 0; q < HP; ++q) racc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < deg; s += 4) {
      float4 xv[4];
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int j = __shfl_sync(0xffffffffu, my_j, (s + w) & 31);
        xv[w] = (s + w < deg) ? __ldg(reinterpret_cast<const float4*>(x + b * x_sb + (long)j * x_sn + g0))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int q = 0; q < HP; ++q) {
          const float dd = __shfl_sync(0xffffffffu, de[q], (s + w) & 31);
          bfma4(racc[q], s + w < deg ? dd : 0.f, xv[w]);
        }
    }
#pragma unroll
    for (int q = 0; q < HP; ++q) *reinterpret_cast<float4*>(rc_out + ((size_t)row * PT + h0 + q) * G + g0) = racc[q];
  }
  if (g0sum != nullptr) *reinterpret_cast<float4*>(g0sum + (size_t)row * G + g0) = t;
}

// KeyQuery softmax backward + dR_i = sum_j de[i,j] x_j.  datt <- de in place.
template <int PT>
__global__ void __launch_bounds__(256) k_softmax_bwd_kq_v(const float* __restrict__ x, long x_sb, long x_sn,
                                                          const float* __restrict__ att,
                                                          const int32_t* __restrict__ nbr_out, long rows, int N,
                                                          int D, int has_datt, float* __restrict__ datt,
                                                          float* __restrict__ rc) {
  constexpr int G = 128;
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = batch_of32(row, N);
  const int g0 = lane * 4;
  const int my_j = lane < D ? nbr_out[row * D + lane] : -1;
  const int deg = __popc(__ballot_sync(0xffffffffu, my_j >= 0));
  float a[PT], da[PT], de[PT];
#pragma unroll
  for (int q = 0; q < PT; ++q) { a[q] = 0.f; da[q] = 0.f; }
  if (my_j >= 0) {
    ldp<PT>(att + ((size_t)row * D + lane) * PT, a);
    if (has_datt) ldp<PT>(datt + ((size_t)row * D + lane) * PT, da);
  }
#pragma unroll
  for (int q = 0; q < PT; ++q) {
    const float dot = warp_sum(a[q] * da[q]);
    de[q] = a[q] * (da[q] - dot);
  }
  if (lane < D) stp<PT>(datt + ((size_t)row * D + lane) * PT, de);
  float4 acc[PT];
#pragma unroll
  for (int q = 0; q < PT; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < deg; s += 4) {
    float4 xv[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int j = __shfl_sync(0xffffffffu, my_j, (s + w) & 31);
      xv[w] = (s + w < deg) ? __ldg(reinterpret_cast<const float4*>(x + b * x_sb + (long)j * x_sn + g0))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int q = 0; q < PT; ++q) {
        const float d = __shfl_sync(0xffffffffu, de[q], (s + w) & 31);
        bfma4(acc[q], s + w < deg ? d : 0.f, xv[w]);
      }
  }
#pragma unroll
  for (int q = 0; q < PT; ++q) *reinterpret_cast<float4*>(rc + ((size_t)row * PT + q) * G + g0) = acc[q];
}

// KeyQuery column side: dx_j = sum_p gU_0[j] + sum_{i in in(j)} sum_p de_p[i,j] R_i^p
template <int PT, int MINB>
__global__ void __launch_bounds__(256, MINB) k_col_bwd_kq_v(const float* __restrict__ gz, const float* __restrict__ datt,
                                                      const float* __restrict__ sproj,
                                                      const int32_t* __restrict__ nbr_in,
                                                      const int32_t* __restrict__ slot_in, long rows, int N, int K,
                                                      int D, int g0_in_dx, const float* __restrict__ dxp, int ndxp,
                                                      float* __restrict__ dx) {
  constexpr int G = 128;
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = batch_of32(row, N);
  const int g0 = lane * 4;
  const int my_i = lane < D ? nbr_in[row * D + lane] : -1;
  const int deg = __popc(__ballot_sync(0xffffffffu, my_i >= 0));
  float de[PT];
#pragma unroll
  for (int q = 0; q < PT; ++q) de[q] = 0.f;
  if (my_i >= 0) ldp<PT>(datt + ((size_t)(b * N + my_i) * D + slot_in[row * D + lane]) * PT, de);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g0_in_dx) {                          // k_tap_bwd_p already left sum_p g_0^p in dx
    acc = *reinterpret_cast<const float4*>(dx + (size_t)row * G + g0);
  } else {
#pragma unroll
    for (int q = 0; q < PT; ++q) {
      const float4 v = *reinterpret_cast<const float4*>(gz + (((size_t)row * PT + q) * K + 0) * G + g0);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  for (int h = 0; h < ndxp; ++h) {         // dense part sum_p dR_p W_p^T, one partial row per head pair
    const float4 v = __ldcs(reinterpret_cast<const float4*>(dxp + ((size_t)row * ndxp + h) * G + g0));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  for (int s = 0; s < deg; s += 2) {
    float4 rv[2][PT];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int i = __shfl_sync(0xffffffffu, my_i, (s + w) & 31);
#pragma unroll
      for (int q = 0; q < PT; ++q)
        rv[w][q] = (s + w < deg) ? __ldg(reinterpret_cast<const float4*>(sproj + ((size_t)(b * N + i) * PT + q) * G + g0))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int w = 0; w < 2; ++w)
#pragma unroll
      for (int q = 0; q < PT; ++q) {
        const float d = __shfl_sync(0xffffffffu, de[q], (s + w) & 31);
        bfma4(acc, s + w < deg ? d : 0.f, rv[w][q]);
      }
  }
  *reinterpret_cast<float4*>(dx + (size_t)row * G + g0) = acc;
}

// ---- GAT_modified vector kernels (G == 128, D <= 32, P in {1,2,4}) -------------------------------------------
// softmax + LeakyReLU backward of a sender row: datt <- ds, rc[row][p][1] = sum_j ds[i,j] (the a2 / row term)
template <int PT>
__global__ void __launch_bounds__(256) k_softmax_bwd_gm_v(const float* __restrict__ sproj, const float* __restrict__ att,
                                                          const int32_t* __restrict__ nbr_out, long rows, int N, int D,
                                                          int has_datt, float* __restrict__ datt,
                                                          float* __restrict__ rc) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = batch_of32(row, N);
  const int my_j = lane < D ? nbr_out[row * D + lane] : -1;
  float a[PT], da[PT], ds[PT];
#pragma unroll
  for (int q = 0; q < PT; ++q) { a[q] = 0.f; da[q] = 0.f; }
  if (my_j >= 0) {
    ldp<PT>(att + ((size_t)row * D + lane) * PT, a);
    if (has_datt) ldp<PT>(datt + ((size_t)row * D + lane) * PT, da);
  }
#pragma unroll
  for (int q = 0; q < PT; ++q) {
    const float dot = warp_sum(a[q] * da[q]);
    float de = a[q] * (da[q] - dot);
    if (my_j >= 0) {
      const float sr = sproj[((size_t)row * PT + q) * 2 + 1] + sproj[((size_t)(b * N + my_j) * PT + q) * 2 + 0];
      de *= sr > 0.f ? 1.f : kLeaky;
    }
    ds[q] = de;
    const float rsum = warp_sum(de);
    if (lane == 0) rc[((size_t)row * PT + q) * 2 + 1] = rsum;
  }
  if (lane < D) stp<PT>(datt + ((size_t)row * D + lane) * PT, ds);
}

// column side: rc[row][p][0] = sum_{i in in(row)} ds[i,row]; dx_row = sum_p gU_0 + sum_p (c_p cvec[p][0] + r_p cvec[p][1])
// Grid-stride over the rows with the next row's in-list entries loaded one iteration ahead; the P column sums share one
// joint warp reduction.
template <int PT>
__global__ void __launch_bounds__(256) k_col_bwd_gm_v(const float* __restrict__ gz, const float* __restrict__ datt,
                                                      const float* __restrict__ cvec,
                                                      const int32_t* __restrict__ nbr_in,
                                                      const int32_t* __restrict__ slot_in, long rows, int N, int K,
                                                      int D, int g0_in_dx, float* __restrict__ rc,
                                                      float* __restrict__ dx) {
  constexpr int G = 128;
  constexpr int SH = PT == 4 ? 3 : PT == 2 ? 4 : 5;           // warp_multi_sum<PT>: lane l holds value l >> SH
  const int lane = threadIdx.x & 31;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const int g0 = lane * 4;
  float4 cv[PT][2];                                            // the 2 P projection vectors of this lane's features
#pragma unroll
  for (int q = 0; q < PT; ++q) {
    cv[q][0] = __ldg(reinterpret_cast<const float4*>(cvec + ((size_t)q * 2 + 0) * G + g0));
    cv[q][1] = __ldg(reinterpret_cast<const float4*>(cvec + ((size_t)q * 2 + 1) * G + g0));
  }
  int n_i = lane < D ? __ldg(nbr_in + row * D + lane) : -1;
  int n_s = lane < D ? __ldg(slot_in + row * D + lane) : 0;
  for (; row < rows; row += nwarps) {
    const int my_i = n_i, my_s = n_s;
    const long nxt = row + nwarps;
    n_i = (nxt < rows && lane < D) ? __ldg(nbr_in + nxt * D + lane) : -1;
    n_s = (nxt < rows && lane < D) ? __ldg(slot_in + nxt * D + lane) : 0;
    const long b = batch_of32(row, N);
    float ds[PT];
#pragma unroll
    for (int q = 0; q < PT; ++q) ds[q] = 0.f;
    if (my_i >= 0) ldp<PT>(datt + ((size_t)(b * N + my_i) * D + my_s) * PT, ds);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (dx != nullptr) {                                       // requested before the reduction below needs it
      if (g0_in_dx) {
        acc = *reinterpret_cast<const float4*>(dx + (size_t)row * G + g0);
      } else {
#pragma unroll
        for (int q = 0; q < PT; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(gz + (((size_t)row * PT + q) * K + 0) * G + g0);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    float rq = lane < PT ? rc[((size_t)row * PT + lane) * 2 + 1] : 0.f;
    const float tot = warp_multi_sum<PT>(ds, lane);
    if ((lane & ((1 << SH) - 1)) == 0) rc[((size_t)row * PT + (lane >> SH)) * 2 + 0] = tot;
    if (dx == nullptr) continue;
#pragma unroll
    for (int q = 0; q < PT; ++q) {
      bfma4(acc, __shfl_sync(0xffffffffu, tot, q << SH), cv[q][0]);
      bfma4(acc, __shfl_sync(0xffffffffu, rq, q), cv[q][1]);
    }
    *reinterpret_cast<float4*>(dx + (size_t)row * G + g0) = acc;
  }
}

// dc[n][g] = sum_m rc[m][n] x[m][g] (g < G) and dc[n][G] = sum_m rc[m][n], n = 2p + t.  Every warp walks a contiguous
// run of nodes with 2P float4 accumulators per lane; blocks write one partial [2P][G+1] each (summed in block order).
template <int PT>
__global__ void __launch_bounds__(256) k_gm_dcvec_v(const float* __restrict__ x, long x_sb, long x_sn,
                                                    const float* __restrict__ rc, long rows, int N, long per_block,
                                                    float* __restrict__ partial) {
  constexpr int G = 128, NV = 2 * PT;
  __shared__ float red[8][NV][G + 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long m0 = (long)blockIdx.x * per_block;
  const long m1 = min(rows, m0 + per_block);
  float4 acc[NV];
  float cs[NV];
#pragma unroll
  for (int n = 0; n < NV; ++n) { acc[n] = make_float4(0.f, 0.f, 0.f, 0.f); cs[n] = 0.f; }
  // four rows per step and warp: their loads are issued together (with one row in flight per warp the kernel streamed
  // x at 1.7 TB/s)
  for (long mq = m0 + warp; mq < m1; mq += 32) {
    float4 xv[4];
    float w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long m = mq + 8 * u;
      xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      w[u] = 0.f;
      if (m < m1) {
        const long b = batch_of32(m, N);
        xv[u] = __ldg(reinterpret_cast<const float4*>(x + b * x_sb + (m - b * N) * x_sn + lane * 4));
        if (lane < NV) w[u] = __ldg(rc + (size_t)m * NV + lane);          // lane n keeps rc[m][n]
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int n = 0; n < NV; ++n) {
        const float wn = __shfl_sync(0xffffffffu, w[u], n);
        bfma4(acc[n], wn, xv[u]);
        cs[n] += wn;                       // identical in every lane; lane 0 reports it
      }
  }
#pragma unroll
  for (int n = 0; n < NV; ++n) {
    *reinterpret_cast<float4*>(&red[warp][n][lane * 4]) = acc[n];
    if (lane == 0) red[warp][n][G] = cs[n];
  }
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * NV * (G + 1);
  for (int e = threadIdx.x; e < NV * (G + 1); e += blockDim.x) {
    const int n = e / (G + 1), g = e - n * (G + 1);
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][n][g];
    out[e] = s;
  }
}

// dbias partials: each block sums `chunk` rows of dP for every channel (coalesced, 8 rows in flight per
// thread), one partial row per block
__global__ void __launch_bounds__(256) k_dbias_partial(DPre dp, long rows, int C, int P, int F, int chunk,
                                                       float* __restrict__ partial) {
  const long r_begin = (long)blockIdx.x * chunk;
  const long r_end = min(rows, r_begin + chunk);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int p = dp.concat ? c / F : 0, f = dp.concat ? c - p * F : c;
    float s[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) s[u] = 0.f;
    long r = r_begin;
    for (; r + 8 <= r_end; r += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] += dp(r + u, p, f);
    }
    for (; r < r_end; ++r) s[0] += dp(r, p, f);
    partial[(size_t)blockIdx.x * C + c] = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
  }
}
// dbias[f] = (concat ? sum_p : P *) sum_blocks partial[block][p*F + f]; one block per f, fixed-order tree
__global__ void __launch_bounds__(256) k_dbias_final(const float* __restrict__ partial, int nblocks, int C, int P,
                                                     int F, int concat, float* __restrict__ dbias) {
  __shared__ float red[256];
  const int f = blockIdx.x;
  const int np = concat ? P : 1;
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks * np; i += blockDim.x) {
    const int k = i / np, p = i - k * np;
    s += partial[(size_t)k * C + p * F + f];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) dbias[f] = concat ? red[0] : red[0] * (float)P;   // mean: every head sees dP = dY / P
}

// KeyQuery: dx[m][g] += sum_{p,g'} W[p][g][g'] dR[m][p][g']
struct RcLoad { const float* rc; int PG; __device__ __forceinline__ float operator()(long m, int k, int) const { return __ldg(rc + (size_t)m * PG + k); } };
struct WtLoad {   // B(k = p*G+g', n = g) = W[p][g][g']
  const float* W; int G;
  __device__ __forceinline__ float operator()(int k, int n, int) const {
    const int p = k / G;
    return __ldg(W + ((size_t)p * G + n) * G + (k - p * G));
  }
};
struct AccEpi { float* out; int ld; __device__ __forceinline__ void operator()(long m, int n, int, float v) const { out[m * ld + n] += v; } };
// dW[p][g][g'] = sum_m x[m][g] dR[m][p][g']
struct XRed {
  const float* x; long x_sb, x_sn; int N;
  __device__ __forceinline__ float operator()(long r, int g, int) const {
    const long b = r / N;
    return __ldg(x + b * x_sb + (r - b * N) * x_sn + g);
  }
};
struct RcRed { const float* rc; int P, G; __device__ __forceinline__ float operator()(long r, int g, int z) const { return __ldg(rc + ((size_t)r * P + z) * G + g); } };
// GAT_modified: dc[n = 2p+t][g] = sum_m rc[m][n] x[m][g], and column G carries sum_m rc[m][n]
struct RcRed2 { const float* rc; int P2; __device__ __forceinline__ float operator()(long r, int n, int) const { return __ldg(rc + (size_t)r * P2 + n); } };
struct XRed1 {
  XRed xr; int G;
  __device__ __forceinline__ float operator()(long r, int g, int z) const { return g < G ? xr(r, g, z) : 1.f; }
};

// chain through cvec = W^T a, dvec = a . wb  (dc: [P][2][G+1], last column = ddvec)
__global__ void __launch_bounds__(128) k_gm_param_bwd(const float* __restrict__ W, const float* __restrict__ mixer,
                                                      const float* __restrict__ wb, const float* __restrict__ dc,
                                                      int G, int F, int P, float* __restrict__ dW,
                                                      float* __restrict__ dmixer, float* __restrict__ dwb) {
  const int p = blockIdx.x;
  const float* dc0 = dc + ((size_t)p * 2 + 0) * (G + 1);
  const float* dc1 = dc + ((size_t)p * 2 + 1) * (G + 1);
  const float dd0 = dc0[G], dd1 = dc1[G];
  const float* a1 = mixer + (size_t)p * 2 * F;
  const float* a2 = a1 + F;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float m0 = dd0 * wb[(size_t)p * F + f], m1 = dd1 * wb[(size_t)p * F + f];
    const float* w = W + ((size_t)p * F + f) * G;
    for (int g = 0; g < G; ++g) {
      m0 = fmaf(w[g], dc0[g], m0);
      m1 = fmaf(w[g], dc1[g], m1);
    }
    if (dmixer) {
      dmixer[(size_t)p * 2 * F + f] = m0;
      dmixer[(size_t)p * 2 * F + F + f] = m1;
    }
    if (dwb) dwb[(size_t)p * F + f] = dd0 * a1[f] + dd1 * a2[f];
    if (dW)
      for (int g = 0; g < G; ++g) dW[((size_t)p * F + f) * G + g] = a1[f] * dc0[g] + a2[f] * dc1[g];
  }
}

// dx[m][g] = sum_p gz[m][p][0][g]
__global__ void __launch_bounds__(256) k_gz0_to_dx(const float* __restrict__ gz, long rows, int G, int P, int K,
                                                   float* __restrict__ dx) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * G) return;
  const long m = t / G;
  const int g = (int)(t - m * G);
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += gz[((m * P + p) * K + 0) * G + g];
  dx[t] = s;
}

int run_tap_gather(const float* x, long x_sb, long x_sn, const float* att, const int32_t* nbr_in,
                   const int32_t* slot_in, int B, int N, int G, int K, int P, int D, int k, float* taps,
                   float* ain, int ain_ready, cudaStream_t st);   // gat_fwd.cu

// gat_wgrad_tc.cu
size_t wgrad_tc_partial_floats(int G, int F, int K, int P);
bool wgrad_tc_supported(const magat_gat_bwd_args* a);
int wgrad_tc_dfilter(const magat_gat_bwd_args* a, bool with_dbias, cudaStream_t st);
int wgrad_tc_dweight(const magat_gat_bwd_args* a, cudaStream_t st);

// gat_tap_tc.cu / gat_tc.cu
bool gz_tc_supported(const magat_gat_bwd_args* a);
int gz_tc_backward(const magat_gat_bwd_args* a, float* ht, cudaStream_t st);
bool dx_tc_supported(const magat_gat_bwd_args* a);
bool dx_tap_supported(const magat_gat_bwd_args* a);                 // gat_tap_tc.cu
int dx_tap_partials(const magat_gat_bwd_args* a, float* wcat, float* dxp, cudaStream_t st);
int tc_dx_accumulate(const magat_gat_bwd_args* a, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo,
                     cudaStream_t st);
int tc_split_weights(const float* src, long n, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st);

// SMs of the current device; the B200 count where there is none (size helpers called on a build box)
static long sm_count_or_default() {
  const int n = device_sm_count();
  return n > 0 ? n : 148;
}

static int pick_splits(long R, int tiles) {
  long s = (sm_count_or_default() * 4 + tiles - 1) / tiles;
  const long cap = (R + 255) / 256;
  if (s > cap) s = cap;
  if (s > 256) s = 256;
  if (s < 1) s = 1;
  return (int)s;
}

template <class ALoad, class BLoad>
static int rowred(long R, int Mi, int Nj, int Z, ALoad A, BLoad Bm, float* partial, float* out, cudaStream_t st,
                  const char* what) {
  const int ti = cdiv(Mi, 64), tj = cdiv(Nj, 64);
  const int splits = pick_splits(R, ti * tj * Z);
  dim3 grid(ti, tj, Z * splits);
  k_rowred_gemm<<<grid, 256, 0, st>>>(R, Mi, Nj, splits, A, Bm, partial);
  int rc = check_launch(what, st);
  if (rc) return rc;
  const long per = (long)Mi * Nj, total = per * Z;
  k_reduce_partials<<<cdiv(total, 256), 256, 0, st>>>(partial, splits, per, total, out);
  return check_launch("k_reduce_partials", st);
}

}  // namespace magat

using namespace magat;

extern "C" size_t magat_gat_bwd_partial_floats(int B, int N, int G, int F, int K, int P, int mode) {
  (void)B; (void)N;
  // every row-reduction writes at most Z * splits * Mi * Nj floats with Z*splits*tiles <= ~148*4 + Z*tiles
  auto need = [](long Mi, long Nj, long Z) {
    const long tiles = ((Mi + 63) / 64) * ((Nj + 63) / 64) * Z;
    long s = (sm_count_or_default() * 4 + tiles - 1) / tiles;
    if (s > 256) s = 256;
    if (s < 1) s = 1;
    return (size_t)(Z * s * Mi * Nj);
  };
  size_t m = need(F, (long)K * G, P);
  m = max(m, need(1, F, 1));
  if (mode == MAGAT_MODE_KEYQUERY) m = max(m, need(G, G, P));
  else m = max(m, need(2l * P, G + 1, 1) + (size_t)2 * P * (G + 1));
  m = max(m, wgrad_tc_partial_floats(G, F, K, P));
  return m;
}

// ---- one-workspace form of the backward call: gz, datt, rc and partial are carved out of a single caller buffer ----
static size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

extern "C" size_t magat_gat_backward_workspace_bytes(int B, int N, int G, int F, int K, int P, int D, int mode) {
  if (B < 1 || N < 1 || G < 1 || F < 1 || K < 1 || P < 1 || D < 1) return 0;
  const size_t rows = (size_t)B * N;
  const size_t rcw = mode == MAGAT_MODE_KEYQUERY ? (size_t)G : 2;
  return up256(rows * P * K * G * 4) + up256(rows * D * P * 4) + up256(rows * P * rcw * 4) +
         up256(magat_gat_bwd_partial_floats(B, N, G, F, K, P, mode) * 4);
}

extern "C" int magat_gat_backward_ws(const magat_gat_bwd_args* a, void* workspace, size_t ws_bytes, void* stream) {
  MAGAT_REQUIRE(a != nullptr && workspace != nullptr, MAGAT_E_BAD_ARG, "magat_gat_backward_ws: null argument");
  MAGAT_REQUIRE(((uintptr_t)workspace % 256) == 0, MAGAT_E_ALIGN, "magat_gat_backward_ws: workspace must be 256 B aligned");
  const size_t need = magat_gat_backward_workspace_bytes(a->B, a->N, a->G, a->F, a->K, a->P, a->D, a->mode);
  MAGAT_REQUIRE(need != 0 && ws_bytes >= need, MAGAT_E_BAD_ARG, "magat_gat_backward_ws: workspace %zu B < %zu B", ws_bytes, need);
  const size_t rows = (size_t)a->B * a->N;
  const size_t rcw = a->mode == MAGAT_MODE_KEYQUERY ? (size_t)a->G : 2;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  magat_gat_bwd_args b = *a;
  b.gz = reinterpret_cast<float*>(w); w += up256(rows * a->P * a->K * a->G * 4);
  b.datt = reinterpret_cast<float*>(w); w += up256(rows * a->D * a->P * 4);
  b.rc = reinterpret_cast<float*>(w); w += up256(rows * a->P * rcw * 4);
  b.partial = reinterpret_cast<float*>(w);
  return magat_gat_backward(&b, stream);
}

extern "C" int magat_gat_backward(const magat_gat_bwd_args* a, void* stream) {
  MAGAT_REQUIRE(a != nullptr, MAGAT_E_BAD_ARG, "magat_gat_backward: null args");
  const int B = a->B, N = a->N, G = a->G, F = a->F, K = a->K, P = a->P, D = a->D;
  MAGAT_REQUIRE(B >= 1 && N >= 1 && G >= 1 && F >= 1 && K >= 1 && P >= 1 && D >= 1, MAGAT_E_BAD_ARG,
                "magat_gat_backward: bad shape");
  MAGAT_REQUIRE(a->mode == MAGAT_MODE_KEYQUERY || a->mode == MAGAT_MODE_GAT_MODIFIED || a->mode == MAGAT_MODE_GSO_VALUES,
                MAGAT_E_BAD_ARG, "magat_gat_backward: unknown mode %d", a->mode);
  const bool plain = a->mode == MAGAT_MODE_GSO_VALUES;       // non-attentional filter: no score / softmax part
  MAGAT_REQUIRE(!plain || P == 1, MAGAT_E_BAD_ARG, "magat_gat_backward: MAGAT_MODE_GSO_VALUES needs P == 1");
  MAGAT_REQUIRE(a->mode != MAGAT_MODE_KEYQUERY || F == G, MAGAT_E_UNSUPPORTED, "KeyQuery needs F == G");
  MAGAT_REQUIRE(a->x && a->nbr_out && a->nbr_in && a->slot_in && a->weight && a->filterWeight && a->y && a->att &&
                    a->sproj && a->dy && a->gz && a->datt && a->rc && a->partial,
                MAGAT_E_BAD_ARG, "magat_gat_backward: null pointer");
  MAGAT_REQUIRE(K == 1 || a->taps, MAGAT_E_BAD_ARG, "magat_gat_backward: taps missing");
  MAGAT_REQUIRE(!a->need_dx || a->dx, MAGAT_E_BAD_ARG, "magat_gat_backward: dx missing");
  MAGAT_REQUIRE(!a->need_dfilter || a->dfilterWeight, MAGAT_E_BAD_ARG, "magat_gat_backward: dfilterWeight missing");
  MAGAT_REQUIRE(!a->need_dbias || a->dbias, MAGAT_E_BAD_ARG, "magat_gat_backward: dbias missing");
  MAGAT_REQUIRE(!a->need_dweight || a->dweight, MAGAT_E_BAD_ARG, "magat_gat_backward: dweight missing");
  const bool gm = a->mode == MAGAT_MODE_GAT_MODIFIED;
  MAGAT_REQUIRE(!(gm && a->need_dmixer) || (a->dmixer && a->dweight_bias), MAGAT_E_BAD_ARG,
                "magat_gat_backward: dmixer / dweight_bias missing");
  MAGAT_REQUIRE(!gm || (a->mixer && a->weight_bias && a->wprep), MAGAT_E_BAD_ARG,
                "magat_gat_backward: GAT_modified needs mixer, weight_bias, wprep");
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const long rows = (long)B * N;
  const int row_blocks = cdiv(rows, 8);
  const int KG = K * G;
  int rc;
  const DPre dp{a->y, a->y_sb, a->y_sn, a->y_sc, a->dy, a->dy_sb, a->dy_sn, a->dy_sc,
                N, F, a->concat, a->relu, a->concat ? 1.f : 1.f / (float)P};
  const ZNode zn{a->x, a->x_sb, a->x_sn, a->taps, N, G, K, P};
  const bool need_scores = plain ? (a->need_dx != 0) : (a->need_dx || a->need_dweight || (gm && a->need_dmixer));
  // the fused forward keeps only u_1 in memory: rebuild the later taps before anything reads them
  for (int k = (a->taps_valid < 1 ? 1 : a->taps_valid + 1); k < K; ++k)
    if ((rc = run_tap_gather(a->x, a->x_sb, a->x_sn, a->att, a->nbr_in, a->slot_in, B, N, G, K, P, D, k,
                             a->taps, nullptr, 0, st)))
      return rc;

  const bool tc_wgrad = a->path != MAGAT_PATH_SIMT && wgrad_tc_supported(a);
  const bool dbias_in_wgrad = a->need_dbias && a->need_dfilter && tc_wgrad;
  if (a->need_dbias && !dbias_in_wgrad) {
    const int C = a->concat ? P * F : F;
    const int nblocks = (int)sm_count_or_default() * 8;
    const int chunk = cdiv(rows, nblocks);
    k_dbias_partial<<<nblocks, 256, 0, st>>>(dp, rows, C, P, F, chunk, a->partial);
    if ((rc = check_launch("k_dbias_partial", st))) return rc;
    k_dbias_final<<<F, 256, 0, st>>>(a->partial, nblocks, C, P, F, a->concat, a->dbias);
    if ((rc = check_launch("k_dbias_final", st))) return rc;
  }
  if (a->need_dfilter) {
    if (tc_wgrad) {
      if ((rc = wgrad_tc_dfilter(a, dbias_in_wgrad, st))) return rc;
    } else if ((rc = rowred(rows, F, KG, P, DPreR{dp}, ZRed{zn, G}, a->partial, a->dfilterWeight, st,
                            "k_rowred_gemm(dfilterWeight)")))
      return rc;
  }
  if (!need_scores) return MAGAT_OK;

  if (a->path != MAGAT_PATH_SIMT && gz_tc_supported(a)) {
    // scratch: transposed taps in `partial` (free between the weight-gradient launches)
    if ((rc = gz_tc_backward(a, a->partial, st))) return rc;
  } else {  // gz = dP H
    dim3 grid(cdiv(rows, 64), cdiv(KG, 64), P);
    HtLoad hl{a->filterWeight, KG, F};
    k_node_gemm<<<grid, 256, 0, st>>>(rows, KG, F, DPreA{dp}, hl, GzEpi{a->gz, P, KG});
    if ((rc = check_launch("k_node_gemm(gz)", st))) return rc;
  }
  const bool vec = G == 128 && D <= 32 && (a->x_sn % 4) == 0 && (a->x_sb % 4) == 0 && ((uintptr_t)a->x % 16) == 0 &&
                   ((uintptr_t)a->gz % 16) == 0 && ((uintptr_t)a->att % 16) == 0 && ((uintptr_t)a->datt % 16) == 0 &&
                   ((uintptr_t)a->rc % 16) == 0 && ((uintptr_t)a->sproj % 16) == 0 &&
                   (K == 1 || ((uintptr_t)a->taps % 16) == 0) && (P == 1 || P == 2 || P == 4) && rows < (1l << 31);
  // KeyQuery, vector kernels: g_0 is only ever read as its head sum (for dx), so the last level of the recursion
  // writes that sum straight into dx and the column kernel picks it up there
  const bool gm_vec = gm && vec;      // GAT_modified vector kernels
  const bool g0_in_dx = vec && (!gm || gm_vec) && K > 1 && a->need_dx;
  // ... and the row softmax backward (+ dR) rides on the same last level
  const bool fuse_softmax = vec && !gm && !plain && K > 1;
  const bool fuse_softmax_gm = gm_vec && K > 1;
  for (int k = K - 1; k >= 1; --k) {
    const int first = k == K - 1 ? 1 : 0;
    float* g0sum = (k == 1 && g0_in_dx) ? a->dx : nullptr;
    float* rc_fused = (k == 1 && fuse_softmax) ? a->rc : nullptr;
    float* gm_rc = (k == 1 && fuse_softmax_gm) ? a->rc : nullptr;
#define MAGAT_TBX(PT, HPV, EFV, MB) \
  k_tap_bwd_p<PT, HPV, EFV, MB><<<row_blocks, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->taps, a->att, a->nbr_out, rows, N, K, D, k, \
                                                   first, a->gz, a->datt, g0sum, rc_fused, a->sproj, gm_rc)
    // measured at B = 512, N = 1000, P = 4 (both levels): 2 heads x 2 edges at 4 CTAs/SM 1.76 ms; 2 x 4 at 4 CTAs/SM
    // 1.80; 2 x 4 at 3 CTAs/SM 2.08; 4 x 2 at 3 CTAs/SM 1.93; one head per pass >= 2.07
    if (vec && P == 4) MAGAT_TBX(4, 2, 2, 4);
    else if (vec && P == 2) MAGAT_TBX(2, 2, 2, 4);
    else if (vec && P == 1) MAGAT_TBX(1, 1, 4, 4);
    else
      k_tap_bwd<<<row_blocks, 256, 0, st>>>(zn, a->att, a->nbr_out, rows, N, G, P, K, D, k, first, a->gz, a->datt);
#undef MAGAT_TBX
    if ((rc = check_launch("k_tap_bwd", st))) return rc;
  }
  const int has_datt = K > 1 ? 1 : 0;
  if (plain) {
    // dx = gU_0 (one head): already in dx when the last recursion level wrote the head sum there
    if (!g0_in_dx) {
      k_gz0_to_dx<<<cdiv(rows * G, 256), 256, 0, st>>>(a->gz, rows, G, P, K, a->dx);
      if ((rc = check_launch("k_gz0_to_dx", st))) return rc;
    }
    return MAGAT_OK;
  }
  if (!gm) {
#define MAGAT_SB(PT) \
  k_softmax_bwd_kq_v<PT><<<row_blocks, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->att, a->nbr_out, rows, N, D, has_datt, \
                                                     a->datt, a->rc)
    if (fuse_softmax) {
    } else if (vec && P == 4) MAGAT_SB(4);
    else if (vec && P == 2) MAGAT_SB(2);
    else if (vec && P == 1) MAGAT_SB(1);
    else
      k_softmax_bwd<MAGAT_MODE_KEYQUERY><<<row_blocks, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->sproj, a->att,
                                                                      a->nbr_out, rows, N, G, P, D, has_datt,
                                                                      a->datt, a->rc);
#undef MAGAT_SB
    if (!fuse_softmax && (rc = check_launch("k_softmax_bwd", st))) return rc;
    // dense part of dx first (needs only dR): head-pair partials into the dead planes of gz, added by the column kernel
    int dx_done = 0;
    const float* dxp = nullptr;
    if (vec && g0_in_dx && a->need_dx && a->path != MAGAT_PATH_SIMT && dx_tap_supported(a) &&
        (size_t)(P / 2) * G <= (size_t)P * K * G) {
      rc = dx_tap_partials(a, a->partial, a->gz, st);
      if (rc > 0) return rc;
      if (rc == 0) { dx_done = 1; dxp = a->gz; }
    }
#define MAGAT_CBX(PT, MB) \
  k_col_bwd_kq_v<PT, MB><<<row_blocks, 256, 0, st>>>(a->gz, a->datt, a->sproj, a->nbr_in, a->slot_in, rows, N, K, D, \
                                                     g0_in_dx ? 1 : 0, dxp, dxp ? P / 2 : 0, a->dx)
    // eight CTAs per SM (32 registers, the gathers consumed one by one) beat fewer, fuller warps here: 0.44 ms against
    // 0.46 at five and 0.49 at four CTAs per SM
#define MAGAT_CB(PT) MAGAT_CBX(PT, 8)
    if (vec && a->need_dx && P == 4) MAGAT_CB(4);
    else if (vec && a->need_dx && P == 2) MAGAT_CB(2);
    else if (vec && a->need_dx && P == 1) MAGAT_CB(1);
    else
      k_col_bwd<MAGAT_MODE_KEYQUERY><<<row_blocks, 256, 0, st>>>(a->gz, a->datt, a->sproj, nullptr, a->nbr_in,
                                                                  a->slot_in, rows, N, G, P, K, D, a->rc,
                                                                  a->need_dx ? a->dx : nullptr);
#undef MAGAT_CB
#undef MAGAT_CBX
    if ((rc = check_launch("k_col_bwd", st))) return rc;
    if (dx_done) {
    } else if (a->need_dx && a->path != MAGAT_PATH_SIMT && dx_tc_supported(a)) {
      __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(a->partial);
      __nv_bfloat16* w_lo = w_hi + (size_t)P * G * G;
      if ((rc = tc_split_weights(a->weight, (long)P * G * G, w_hi, w_lo, st))) return rc;
      if ((rc = tc_dx_accumulate(a, w_hi, w_lo, st))) return rc;
    } else if (a->need_dx) {
      dim3 grid(cdiv(rows, 64), cdiv(G, 64), 1);
      k_node_gemm<<<grid, 256, 0, st>>>(rows, G, P * G, RcLoad{a->rc, P * G}, WtLoad{a->weight, G},
                                        AccEpi{a->dx, G});
      if ((rc = check_launch("k_node_gemm(dx += W dR)", st))) return rc;
    }
    if (a->need_dweight) {
      if (tc_wgrad) {
        if ((rc = wgrad_tc_dweight(a, st))) return rc;
      } else if ((rc = rowred(rows, G, G, P, XRed{a->x, a->x_sb, a->x_sn, N}, RcRed{a->rc, P, G}, a->partial,
                              a->dweight, st, "k_rowred_gemm(dweight)")))
        return rc;
    }
  } else {
    const float* cvec = a->wprep;
    if (gm_vec) {
#define MAGAT_GSB(PT) \
  k_softmax_bwd_gm_v<PT><<<row_blocks, 256, 0, st>>>(a->sproj, a->att, a->nbr_out, rows, N, D, has_datt, a->datt, a->rc)
      if (fuse_softmax_gm) {
        // done by the last level of the recursion
      } else if (P == 4) MAGAT_GSB(4);
      else if (P == 2) MAGAT_GSB(2);
      else MAGAT_GSB(1);
#undef MAGAT_GSB
      if (!fuse_softmax_gm && (rc = check_launch("k_softmax_bwd", st))) return rc;
      const int col_grid = row_blocks < sm_count_or_default() * 48 ? row_blocks : sm_count_or_default() * 48;
#define MAGAT_GCB(PT) \
  k_col_bwd_gm_v<PT><<<col_grid, 256, 0, st>>>(a->gz, a->datt, cvec, a->nbr_in, a->slot_in, rows, N, K, D, \
                                                 g0_in_dx ? 1 : 0, a->rc, a->need_dx ? a->dx : nullptr)
      if (P == 4) MAGAT_GCB(4);
      else if (P == 2) MAGAT_GCB(2);
      else MAGAT_GCB(1);
#undef MAGAT_GCB
      if ((rc = check_launch("k_col_bwd", st))) return rc;
    } else {
      k_softmax_bwd<MAGAT_MODE_GAT_MODIFIED><<<row_blocks, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->sproj, a->att,
                                                                          a->nbr_out, rows, N, G, P, D, has_datt,
                                                                          a->datt, a->rc);
      if ((rc = check_launch("k_softmax_bwd", st))) return rc;
      k_col_bwd<MAGAT_MODE_GAT_MODIFIED><<<row_blocks, 256, 0, st>>>(a->gz, a->datt, a->sproj, cvec, a->nbr_in,
                                                                      a->slot_in, rows, N, G, P, K, D, a->rc,
                                                                      a->need_dx ? a->dx : nullptr);
      if ((rc = check_launch("k_col_bwd", st))) return rc;
    }
    if (a->need_dweight || a->need_dmixer) {
      // dc [2P][G+1] lives at the tail of `partial`
      const size_t head = (size_t)magat_gat_bwd_partial_floats(B, N, G, F, K, P, a->mode) - (size_t)2 * P * (G + 1);
      float* dc = a->partial + head;
      const int nblk = 2 * (int)sm_count_or_default();
      if (gm_vec && (size_t)nblk * 2 * P * (G + 1) <= head) {
        const long per_block = (rows + nblk - 1) / nblk;
#define MAGAT_GDC(PT) k_gm_dcvec_v<PT><<<nblk, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->rc, rows, N, per_block, a->partial)
        if (P == 4) MAGAT_GDC(4);
        else if (P == 2) MAGAT_GDC(2);
        else MAGAT_GDC(1);
#undef MAGAT_GDC
        if ((rc = check_launch("k_gm_dcvec", st))) return rc;
        const long per = 2l * P * (G + 1);
        k_reduce_partials<<<cdiv(per, 256), 256, 0, st>>>(a->partial, nblk, per, per, dc);
        if ((rc = check_launch("k_reduce_partials", st))) return rc;
      } else if ((rc = rowred(rows, 2 * P, G + 1, 1, RcRed2{a->rc, 2 * P}, XRed1{XRed{a->x, a->x_sb, a->x_sn, N}, G},
                              a->partial, dc, st, "k_rowred_gemm(dcvec)")))
        return rc;
      k_gm_param_bwd<<<P, 128, 0, st>>>(a->weight, a->mixer, a->weight_bias, dc, G, F, P,
                                        a->need_dweight ? a->dweight : nullptr,
                                        a->need_dmixer ? a->dmixer : nullptr,
                                        a->need_dmixer ? a->dweight_bias : nullptr);
      if ((rc = check_launch("k_gm_param_bwd", st))) return rc;
    }
  }
  return MAGAT_OK;
}
