// Forward of the batched graph-attention layer -- generic fp32 SIMT path.
//
// Math (validated against the reference autograd):
//   scores   KeyQuery      e_p[i,j] = x_i^T W_p x_j = R_i^p . x_j,  R_i^p = W_p^T x_i      (graphML.py:1246-1262)
//            GAT_modified  e_p[i,j] = LeakyReLU_0.2(a2_p.z_i + a1_p.z_j), z = W_p x + wb_p  (graphML.py:777-796)
//   A_p[i,:] = softmax of e_p[i,:] over the out-neighbours of i; empty row -> zeros      (graphML.py:1278-1286)
//   u_0 = x, u_k[j] = sum_i A_p[i,j] u_{k-1}[i]                                          (graphML.py:1756-1759)
//   Y_p[n] = sum_k H_{p,k} u_k^p[n] + bias; concat: relu(Y_p) at channel p*F+f; mean: relu(mean_p Y_p)
//
// Nothing N x N is ever materialised: the GSO arrives here already as padded neighbour lists
// (gso_scan.cu).  This file is the path for every shape; the tcgen05 path (gat_tc.cu) replaces
// the two dense projections for the shapes it covers.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "gat_sparse.cuh"
#include "simt_gemm.cuh"

namespace magat {

// ---- GAT_modified parameter folding -------------------------------------------------------
// a_t.z_n = a_t.(W x_n + wb) = (W^T a_t).x_n + a_t.wb, so per head two G-vectors and two scalars
// replace the F x N projection:  cvec[p][t][g] = sum_f mixer[p][t*F+f] W[p][f][g],
// dvec[p][t] = sum_f mixer[p][t*F+f] wb[p][f];  t = 0 is a1 (column/receiver term), t = 1 is a2.
__global__ void __launch_bounds__(128) k_gm_prep(const float* __restrict__ W, const float* __restrict__ mixer,
                                                 const float* __restrict__ wb, int G, int F, int P,
                                                 float* __restrict__ cvec, float* __restrict__ dvec) {
  const int pt = blockIdx.x;            // p*2 + t
  const int p = pt >> 1, t = pt & 1;
  const float* a = mixer + (size_t)p * 2 * F + (size_t)t * F;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    float s = 0.f;
    for (int f = 0; f < F; ++f) s = fmaf(a[f], W[((size_t)p * F + f) * G + g], s);
    cvec[(size_t)pt * G + g] = s;
  }
  if (threadIdx.x < 32) {
    float s = 0.f;
    for (int f = threadIdx.x; f < F; f += 32) s = fmaf(a[f], wb[(size_t)p * F + f], s);
    s = warp_sum(s);
    if (threadIdx.x == 0) dvec[pt] = s;
  }
}

// ---- row softmax over the neighbour list ---------------------------------------------------
// One warp per sender row i.  Scores go through `att` (owned by this warp) so arbitrary degree
// works; masked entries never enter the softmax, which is what softmax(e*M - 1e12(1-M))*M gives
// in fp32 whenever the row has at least one edge, and an empty row stays all zero.
template <int MODE>
__global__ void __launch_bounds__(256) k_attention(const float* __restrict__ x, long x_sb, long x_sn,
                                                   const float* __restrict__ sproj,
                                                   const int32_t* __restrict__ nbr_out, long rows, int N,
                                                   int G, int P, int D, float* __restrict__ att,
                                                   const int32_t* __restrict__ slot_out, float* __restrict__ ain) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = row / N;
  const int32_t* nb = nbr_out + row * D;
  float* arow = att + row * (size_t)D * P;
  int deg = 0;
  for (int s0 = 0; s0 < D; s0 += 32) {
    const int s = s0 + lane;
    deg += __popc(__ballot_sync(0xffffffffu, s < D && nb[s] >= 0));
  }
  for (int s = deg + lane; s < D; s += 32)
    for (int p = 0; p < P; ++p) arow[(size_t)s * P + p] = 0.f;
  if (deg == 0) return;
  const float* xb = x + b * x_sb;
  for (int p = 0; p < P; ++p) {
    float mx = -INFINITY;
    if (MODE == MAGAT_MODE_KEYQUERY) {
      const float* r = sproj + ((size_t)row * P + p) * G;
      for (int s = 0; s < deg; ++s) {
        const float* xj = xb + (long)nb[s] * x_sn;
        float d = 0.f;
        for (int g = lane; g < G; g += 32) d = fmaf(r[g], xj[g], d);
        d = warp_sum(d);
        if (lane == 0) arow[(size_t)s * P + p] = d;
        mx = fmaxf(mx, d);
      }
      __syncwarp();
    } else {
      const float si = sproj[((size_t)row * P + p) * 2 + 1];
      for (int s = lane; s < deg; s += 32) {
        const long j = b * N + nb[s];
        float e = si + sproj[((size_t)j * P + p) * 2 + 0];
        e = e > 0.f ? e : kLeaky * e;
        arow[(size_t)s * P + p] = e;
        mx = fmaxf(mx, e);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    float sum = 0.f;
    for (int s = lane; s < deg; s += 32) {
      const float ex = expf(arow[(size_t)s * P + p] - mx);
      arow[(size_t)s * P + p] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int s = lane; s < deg; s += 32) {
      const float v = arow[(size_t)s * P + p] * inv;
      arow[(size_t)s * P + p] = v;
      if (ain != nullptr) ain[((size_t)(b * N + nb[s]) * P + p) * D + slot_out[row * D + s]] = v;
    }
  }
}

// ---- one tap of the recursion: u_k[j] = sum_{i in in(j)} A_p[i,j] u_{k-1}[i] ----------------
// One warp per receiver j; lanes sweep the feature axis.  k >= 1; u_0 = x.
__global__ void __launch_bounds__(256) k_tap_gather(const float* __restrict__ x, long x_sb, long x_sn,
                                                    const float* __restrict__ att,
                                                    const int32_t* __restrict__ nbr_in,
                                                    const int32_t* __restrict__ slot_in, long rows, int N,
                                                    int G, int P, int K, int D, int k,
                                                    float* __restrict__ taps) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = row / N;
  const int32_t* nb = nbr_in + row * D;
  const int32_t* sl = slot_in + row * D;
  const int Km1 = K - 1;
  for (int p = 0; p < P; ++p) {
    for (int g0 = 0; g0 < G; g0 += 32) {
      const int g = g0 + lane;
      float acc = 0.f;
      for (int s = 0; s < D; ++s) {
        const int i = nb[s];
        if (i < 0) break;
        const long ri = b * N + i;
        const float a = att[((size_t)ri * D + sl[s]) * P + p];
        const float* src = (k == 1) ? (x + b * x_sb + (long)i * x_sn)
                                    : (taps + (((size_t)ri * P + p) * Km1 + (k - 2)) * G);
        if (g < G) acc = fmaf(a, src[g], acc);
      }
      if (g < G) taps[(((size_t)row * P + p) * Km1 + (k - 1)) * G + g] = acc;
    }
  }
}

// ---- vectorised variants for the common shapes (P in {1,2,4}, 16 B aligned rows) -----------------
__device__ __forceinline__ void fma4(float4& acc, float a, const float4& v) {
  acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y); acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// u_k[j] for all heads at once: one warp per receiver, each lane owns 4 consecutive features of
// every 128-feature slab; lane s fetches the attention values of in-edge s (all heads) and they are
// broadcast by shuffle.  When `ain` is given the in-edge values are also stored receiver-major,
// ain[j][p][s] = A_p[nbr_in[j][s], j] (0 beyond the degree), which is what the fused tcgen05 kernel reads.
template <int PT, int MINB>
__global__ void __launch_bounds__(256, MINB) k_tap_gather_v(const float* __restrict__ x, long x_sb, long x_sn,
                                                      const float* __restrict__ att,
                                                      const int32_t* __restrict__ nbr_in,
                                                      const int32_t* __restrict__ slot_in, long rows, int N,
                                                      int G, int K, int D, int k, float* __restrict__ taps,
                                                      float* __restrict__ ain, int ain_ready) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = batch_of32(row, N);
  const int32_t* nb = nbr_in + row * D;
  const int32_t* sl = slot_in + row * D;
  const int Km1 = K - 1;
  for (int gb = 0; gb < G; gb += 128) {       // warp-uniform trip count: the shuffles below need every lane
    const int g0 = gb + lane * 4;
    const bool act = g0 < G;
    float4 acc[PT];
#pragma unroll
    for (int p = 0; p < PT; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s0 = 0; s0 < D; s0 += 32) {
      const int my_i = (s0 + lane < D) ? nb[s0 + lane] : -1;
      float am[PT];
#pragma unroll
      for (int p = 0; p < PT; ++p) am[p] = 0.f;
      if (my_i >= 0 && ain_ready) {
#pragma unroll
        for (int p = 0; p < PT; ++p) am[p] = __ldg(ain + ((size_t)row * PT + p) * D + s0 + lane);
      } else if (my_i >= 0) {
        const float* ap = att + ((size_t)(b * N + my_i) * D + sl[s0 + lane]) * PT;
        if (PT == 4) {
          const float4 a4 = __ldg(reinterpret_cast<const float4*>(ap));
          am[0] = a4.x; am[1 % PT] = a4.y; am[2 % PT] = a4.z; am[3 % PT] = a4.w;
        } else {
#pragma unroll
          for (int p = 0; p < PT; ++p) am[p] = __ldg(ap + p);
        }
      }
      if (ain != nullptr && !ain_ready && gb == 0 && s0 + lane < D) {
#pragma unroll
        for (int p = 0; p < PT; ++p) ain[((size_t)row * PT + p) * D + s0 + lane] = am[p];
      }
      const int cnt = __popc(__ballot_sync(0xffffffffu, my_i >= 0));
      for (int s = 0; s < cnt; s += 4) {     // four neighbours per round: their row loads overlap
        int iu[4];
        float au[4][PT];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          iu[u] = __shfl_sync(0xffffffffu, my_i, (s + u) & 31);
#pragma unroll
          for (int p = 0; p < PT; ++p) au[u][p] = __shfl_sync(0xffffffffu, am[p], (s + u) & 31);
          if (s + u >= cnt) iu[u] = -1;
        }
        if (!act) continue;
        if (k == 1) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            v[u] = iu[u] >= 0 ? __ldg(reinterpret_cast<const float4*>(x + b * x_sb + (long)iu[u] * x_sn + g0))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int p = 0; p < PT; ++p) fma4(acc[p], au[u][p], v[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 4; u += 2) {
            float4 v[2][PT];
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
              for (int p = 0; p < PT; ++p)
                v[w][p] = iu[u + w] >= 0
                              ? *reinterpret_cast<const float4*>(
                                    taps + (((size_t)(b * N + iu[u + w]) * PT + p) * Km1 + (k - 2)) * G + g0)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int w = 0; w < 2; ++w)
#pragma unroll
              for (int p = 0; p < PT; ++p) fma4(acc[p], au[u + w][p], v[w][p]);
          }
        }
      }
      if (cnt < 32) {
        // zero the rest of ain beyond this block
        if (ain != nullptr && !ain_ready && gb == 0)
          for (int s = s0 + 32 + lane; s < D; s += 32)
#pragma unroll
            for (int p = 0; p < PT; ++p) ain[((size_t)row * PT + p) * D + s] = 0.f;
        break;
      }
    }
    if (act) {
#pragma unroll
      for (int p = 0; p < PT; ++p)
        *reinterpret_cast<float4*>(taps + (((size_t)row * PT + p) * Km1 + (k - 1)) * G + g0) = acc[p];
    }
  }
}

// KeyQuery scores + row softmax, D <= 32: lane s owns slot s; G = 128 * GV.
template <int PT, int GV, int MINB>
__global__ void __launch_bounds__(256, MINB) k_attention_kq_v(const float* __restrict__ x, long x_sb, long x_sn,
                                                        const float* __restrict__ sproj,
                                                        const int32_t* __restrict__ nbr_out, long rows, int N,
                                                        int D, float* __restrict__ att,
                                                        const int32_t* __restrict__ slot_out,
                                                        float* __restrict__ ain) {
  constexpr int G = 128 * GV;
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = batch_of32(row, N);
  const int my_j = lane < D ? nbr_out[row * D + lane] : -1;
  const int deg = __popc(__ballot_sync(0xffffffffu, my_j >= 0));
  float e[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) e[p] = -INFINITY;
  if (deg > 0) {
    float4 r[PT][GV];
#pragma unroll
    for (int p = 0; p < PT; ++p)
#pragma unroll
      for (int v = 0; v < GV; ++v)
        r[p][v] = __ldg(reinterpret_cast<const float4*>(sproj + ((size_t)row * PT + p) * G + v * 128 + lane * 4));
    for (int s = 0; s < deg; ++s) {
      const int j = __shfl_sync(0xffffffffu, my_j, s);
      const float* xj = x + b * x_sb + (long)j * x_sn + lane * 4;
      float4 xv[GV];
#pragma unroll
      for (int v = 0; v < GV; ++v) xv[v] = __ldg(reinterpret_cast<const float4*>(xj + v * 128));
#pragma unroll
      for (int p = 0; p < PT; ++p) {
        float d = 0.f;
#pragma unroll
        for (int v = 0; v < GV; ++v) d += dot4(r[p][v], xv[v]);
        d = warp_sum(d);
        if (lane == s) e[p] = d;
      }
    }
  }
  float a[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) {
    const float mx = warp_max(e[p]);
    const float ex = lane < deg ? expf(e[p] - mx) : 0.f;
    const float sum = warp_sum(ex);
    a[p] = lane < deg ? ex / sum : 0.f;
  }
  if (lane < D) {
    float* dst = att + ((size_t)row * D + lane) * PT;
    if (PT == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1 % PT], a[2 % PT], a[3 % PT]);
    } else {
#pragma unroll
      for (int p = 0; p < PT; ++p) dst[p] = a[p];
    }
    // receiver-major copy: A_p[i, j] lands at slot (position of i in j's in-list) of ain[j][p][:]
    if (ain != nullptr && lane < deg) {
      float* r = ain + ((size_t)(b * N + my_j) * PT) * D + slot_out[row * D + lane];
#pragma unroll
      for (int p = 0; p < PT; ++p) r[(size_t)p * D] = a[p];
    }
  }
}

// ---- GAT_modified fast path (G = 128, P in {1,2,4}, D <= 32) ----------------------------------------
// sproj[m][2p + t] = cvec[p][t] . x_m + dvec[p][t]: the whole "projection" of this mode is 2P dot products per node.
// A warp keeps its slice of the 2P vectors in registers and walks over `per_warp` consecutive nodes, one 512 B row
// load each -- the generic 64 x 64 tile GEMM spent 0.40 ms on an 8-column output.
template <int PT>
__global__ void __launch_bounds__(256) k_gm_mixer_v(const float* __restrict__ x, long x_sb, long x_sn,
                                                    const float* __restrict__ cvec, const float* __restrict__ dvec,
                                                    long rows, int N, int per_warp, float* __restrict__ sproj) {
  constexpr int G = 128, NV = 2 * PT;
  constexpr int SH = NV == 8 ? 2 : NV == 4 ? 3 : 4;            // 5 - log2 NV
  const int lane = threadIdx.x & 31;
  const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  float4 c[NV];
#pragma unroll
  for (int n = 0; n < NV; ++n) c[n] = __ldg(reinterpret_cast<const float4*>(cvec + (size_t)n * G + lane * 4));
  const int mine = lane >> SH;                                   // output this lane ends up holding
  const float dv = __ldg(dvec + mine);
  const long m_end = min(rows, (w + 1) * per_warp);
  for (long m0 = w * per_warp; m0 < m_end; m0 += 4) {          // four row loads in flight per warp
    float4 xv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long m = m0 + u < m_end ? m0 + u : m_end - 1;
      const long b = batch_of32(m, N);
      xv[u] = __ldg(reinterpret_cast<const float4*>(x + b * x_sb + (m - b * N) * x_sn + lane * 4));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float d[NV];
#pragma unroll
      for (int n = 0; n < NV; ++n) d[n] = dot4(c[n], xv[u]);
      const float tot = warp_multi_sum<NV>(d, lane);
      if (m0 + u < m_end && (lane & ((1 << SH) - 1)) == 0) sproj[(size_t)(m0 + u) * NV + mine] = tot + dv;
    }
  }
}

// scores e = LeakyReLU(a2.z_i + a1.z_j) from the two scalars per (node, head), row softmax, att + receiver-major ain
template <int PT>
__global__ void __launch_bounds__(256) k_attention_gm_v(const float* __restrict__ sproj,
                                                        const int32_t* __restrict__ nbr_out, long rows, int N, int D,
                                                        float* __restrict__ att, const int32_t* __restrict__ slot_out,
                                                        float* __restrict__ ain) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = batch_of32(row, N);
  const int my_j = lane < D ? __ldg(nbr_out + row * D + lane) : -1;
  const bool v = my_j >= 0;
  float a[PT];
#pragma unroll
  for (int p = 0; p < PT; ++p) {
    const float si = __ldg(sproj + ((size_t)row * PT + p) * 2 + 1);
    float e = -INFINITY;
    if (v) {
      e = si + __ldg(sproj + ((size_t)(b * N + my_j) * PT + p) * 2 + 0);
      e = e > 0.f ? e : kLeaky * e;
    }
    const float mx = warp_max(e);
    const float ex = v ? expf(e - mx) : 0.f;
    const float sum = warp_sum(ex);
    a[p] = v ? ex / sum : 0.f;
  }
  if (lane < D) {
    float* dst = att + ((size_t)row * D + lane) * PT;
    if (PT == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(a[0], a[1 % PT], a[2 % PT], a[3 % PT]);
    } else {
#pragma unroll
      for (int p = 0; p < PT; ++p) dst[p] = a[p];
    }
    if (ain != nullptr && v) {
      float* r = ain + ((size_t)(b * N + my_j) * PT) * D + slot_out[row * D + lane];
#pragma unroll
      for (int p = 0; p < PT; ++p) r[(size_t)p * D] = a[p];
    }
  }
}

// The same attention with lane = (slot, head): 32 / PT out-slots x PT heads per chunk, so a row of degree <= 8 (P = 4) is
// one pass: one gathered score per lane, the max and the sum over the slots are xor-shuffles over the upper lane bits
// for all heads at once (6 shuffles instead of 40), one coalesced store.  Grid-stride, the next row's list entry of the
// lane loaded one iteration ahead.  No receiver-major copy (the lean gathers do not read one).
template <int PT>
__global__ void __launch_bounds__(256) k_attention_gm_h(const float* __restrict__ sproj,
                                                        const int32_t* __restrict__ nbr_out, long rows, int N, int D,
                                                        float* __restrict__ att) {
  constexpr int LOGP = PT == 4 ? 2 : PT == 2 ? 1 : 0;
  constexpr int SPC = 32 / PT, NC = 32 / SPC;
  const int lane = threadIdx.x & 31;
  const int sl = lane >> LOGP, hd = lane & (PT - 1);
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  int nj[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) nj[c] = (c * SPC + sl < D) ? __ldg(nbr_out + row * D + c * SPC + sl) : -1;
  for (; row < rows; row += nwarps) {
    int j[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) j[c] = nj[c];
    const long nxt = row + nwarps;
#pragma unroll
    for (int c = 0; c < NC; ++c) nj[c] = (nxt < rows && c * SPC + sl < D) ? __ldg(nbr_out + nxt * D + c * SPC + sl) : -1;
    const long b = batch_of32(row, N);
    const float si = __ldg(sproj + ((size_t)row * PT + hd) * 2 + 1);
    float e[NC];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      e[c] = -INFINITY;
      if (c * SPC < D && j[c] >= 0) {
        const float v = si + __ldg(sproj + ((size_t)(b * N + j[c]) * PT + hd) * 2 + 0);
        e[c] = v > 0.f ? v : kLeaky * v;
      }
      m = fmaxf(m, e[c]);
    }
#pragma unroll
    for (int o = PT; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      e[c] = (c * SPC < D && j[c] >= 0) ? expf(e[c] - m) : 0.f;
      sum += e[c];
    }
#pragma unroll
    for (int o = PT; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float* dst = att + (size_t)row * D * PT;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (c * SPC + sl < D) dst[c * SPC * PT + lane] = j[c] >= 0 ? e[c] / sum : 0.f;
  }
}

// ---- lean sparse kernels (G = 128, D <= 32, P in {1,2,4}): the warp-level routines of gat_sparse.cuh, one row per warp ----
// Attention: half a warp per edge, joint head reduction, softmax through shared memory; no receiver-major copy.
template <int PT, int MINB>
__global__ void __launch_bounds__(256, MINB) k_attention_kq_h(const float* __restrict__ x, long x_sb, long x_sn,
                                                        const float* __restrict__ sproj,
                                                        const int32_t* __restrict__ nbr_out, long rows, int N, int D,
                                                        float* __restrict__ att) {
  __shared__ float esc[8 * 32 * PT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long nwarps = (long)gridDim.x * 8;
  long row = (long)blockIdx.x * 8 + warp;
  if (row >= rows) return;
  // grid-stride over the rows, the next row's list entry loaded one iteration ahead
  int my_j = sparse::list_entry<false>(nbr_out + row * D, D, lane, -1);
  for (; row < rows; row += nwarps) {
    const long nxt = row + nwarps;
    const int nxt_j = nxt < rows ? sparse::list_entry<false>(nbr_out + nxt * D, D, lane, -1) : -1;
    const long b = batch_of32(row, N);
    sparse::attention_kq_row<PT, false>(x + b * x_sb, (unsigned)x_sn, sproj + (size_t)row * PT * 128, my_j,
                                        att + (size_t)row * D * PT, D, lane, esc + warp * 32 * PT);
    my_j = nxt_j;
  }
}

// Tap level k for all heads; the in-edge attention values come straight from the senders' softmax rows (slot_in).
template <int PT, bool K1>
__global__ void __launch_bounds__(256) k_tap_gather_s(const float* __restrict__ x, long x_sb, long x_sn,
                                                      const float* __restrict__ att, const int32_t* __restrict__ nbr_in,
                                                      const int32_t* __restrict__ slot_in, long rows, int N, int K, int D,
                                                      int k, float* __restrict__ taps) {
  const int lane = threadIdx.x & 31;
  const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const unsigned trow = (unsigned)((K - 1) * 128);
  // grid-stride over the rows; the next row's list entries are loaded one iteration ahead (a second stage -- the edges'
  // attention values one row ahead as well -- measured slower: 0.72 against 0.68 ms for both levels)
  int my_i = sparse::list_entry<false>(nbr_in + row * D, D, lane, -1);
  int my_sl = sparse::list_entry<false>(slot_in + row * D, D, lane, 0);
  for (; row < rows; row += nwarps) {
    const long nxt = row + nwarps;
    const int nxt_i = nxt < rows ? sparse::list_entry<false>(nbr_in + nxt * D, D, lane, -1) : -1;
    const int nxt_sl = nxt < rows ? sparse::list_entry<false>(slot_in + nxt * D, D, lane, 0) : 0;
    const long b = batch_of32(row, N);
    float am[PT];
    sparse::edge_weights<PT, false>(att + (size_t)b * N * D * PT, my_i, my_sl, D, am);
    const float* src = taps + (size_t)b * N * PT * trow + (k >= 2 ? (k - 2) * 128 : 0);
    float4 acc[PT];
    sparse::gather_row<PT, K1, false>(x + b * x_sb, (unsigned)x_sn, src, trow, my_i, am, lane, acc);
    float* out = taps + (size_t)row * PT * trow + (k - 1) * 128 + lane * 4;
#pragma unroll
    for (int h = 0; h < PT; ++h) *reinterpret_cast<float4*>(out + h * trow) = acc[h];
    my_i = nxt_i;
    my_sl = nxt_sl;
  }
}

// grid of the grid-stride sparse kernels: enough blocks to fill every SM a few times over (the rows differ in degree),
// few enough that a warp sees many rows and its one-row-ahead list loads pay
static int lean_grid(int row_blocks) {
  const int sms = device_sm_count() > 0 ? device_sm_count() : 148;
  const int cap = sms * 8 * 6;
  return row_blocks < cap ? row_blocks : cap;
}

// shapes the lean sparse kernels take (32-bit offsets inside an instance)
static bool lean_sparse_ok(long x_sb, long x_sn, int N, int G, int K, int P, int D) {
  return G == 128 && D <= 32 && D % 4 == 0 && (P == 1 || P == 2 || P == 4) && x_sn < (1l << 24) &&
         (long)N * x_sn < (1l << 31) && (long)N * P * (K > 1 ? K - 1 : 1) * 128 < (1l << 31) && (x_sn % 4) == 0 &&
         (x_sb % 4) == 0;
}

// ---- functors for the tile GEMMs ----------------------------------------------------------
struct XLoad {   // A(m, g): node features
  const float* x; long x_sb, x_sn; int N;
  __device__ __forceinline__ float operator()(long m, int g, int) const {
    const long b = m / N;
    return __ldg(x + b * x_sb + (m - b * N) * x_sn + g);
  }
};
struct KqWLoad {   // B(g, n = p*G + g'): W[p][g][g']
  const float* W; int G;
  __device__ __forceinline__ float operator()(int g, int n, int) const {
    const int p = n / G;
    return __ldg(W + ((size_t)p * G + g) * G + (n - p * G));
  }
};
struct StoreEpi {   // C[m][n]
  float* out; int ld;
  __device__ __forceinline__ void operator()(long m, int n, int, float v) const { out[m * ld + n] = v; }
};
struct GmCLoad {   // B(g, n = 2p+t): cvec[p][t][g]
  const float* cvec; int G;
  __device__ __forceinline__ float operator()(int g, int n, int) const { return __ldg(cvec + (size_t)n * G + g); }
};
struct GmEpi {   // sproj[m][n] = acc + dvec[n]
  float* out; const float* dvec; int ld;
  __device__ __forceinline__ void operator()(long m, int n, int, float v) const { out[m * ld + n] = v + dvec[n]; }
};

// Z(m, kk, z): the stacked taps [x | u_1^p | ... | u_{K-1}^p] of node m.
// concat: z = p, kk in [0, K*G); mean: z = 0, kk in [0, P*K*G) with p = kk / (K*G).
struct ZLoad {
  const float* x; long x_sb, x_sn; const float* taps; int N, G, K, P, per_head;
  __device__ __forceinline__ float operator()(long m, int kk, int z) const {
    int p = z;
    if (!per_head) { p = kk / (K * G); kk -= p * K * G; }
    const int k = kk / G, g = kk - k * G;
    if (k == 0) {
      const long b = m / N;
      return __ldg(x + b * x_sb + (m - b * N) * x_sn + g);
    }
    return __ldg(taps + (((size_t)m * P + p) * (K - 1) + (k - 1)) * G + g);
  }
};
struct HLoad {   // B(kk, f, z): filterWeight[p][f][k][g]
  const float* H; int G, K, F, per_head;
  __device__ __forceinline__ float operator()(int kk, int f, int z) const {
    int p = z;
    if (!per_head) { p = kk / (K * G); kk -= p * K * G; }
    return __ldg(H + ((size_t)p * F + f) * K * G + kk);
  }
};
struct YEpi {
  float* y; long y_sb, y_sn, y_sc; const float* bias; int N, F, relu; float scale;
  __device__ __forceinline__ void operator()(long m, int f, int z, float v) const {
    v = v * scale + (bias ? bias[f] : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    const long b = m / N;
    y[b * y_sb + (m - b * N) * y_sn + ((long)z * F + f) * y_sc] = v;
  }
};

static int validate_common(int B, int N, int G, int F, int K, int P, int D, int mode) {
  MAGAT_REQUIRE(B >= 1 && N >= 1 && G >= 1 && F >= 1 && K >= 1 && P >= 1 && D >= 1, MAGAT_E_BAD_ARG,
                "bad shape B=%d N=%d G=%d F=%d K=%d P=%d D=%d", B, N, G, F, K, P, D);
  MAGAT_REQUIRE(mode == MAGAT_MODE_KEYQUERY || mode == MAGAT_MODE_GAT_MODIFIED || mode == MAGAT_MODE_GSO_VALUES,
                MAGAT_E_BAD_ARG, "unknown attention mode %d", mode);
  MAGAT_REQUIRE(mode != MAGAT_MODE_GSO_VALUES || P == 1, MAGAT_E_BAD_ARG, "MAGAT_MODE_GSO_VALUES needs P == 1 (got %d)", P);
  MAGAT_REQUIRE(mode != MAGAT_MODE_KEYQUERY || F == G, MAGAT_E_UNSUPPORTED,
                "KeyQuery needs F == G (got F=%d G=%d; graphML.py:1728,1765)", F, G);
  MAGAT_REQUIRE(P <= 65535 && (long)P * K * G < (1l << 31), MAGAT_E_UNSUPPORTED, "P/K/G too large");
  return MAGAT_OK;
}

// gat_tc.cu
bool tc_supported(const magat_gat_fwd_args* a);
size_t tc_wprep_floats(int G, int F, int K, int P, int mode);
int tc_score_projection(const magat_gat_fwd_args* a, const __nv_bfloat16* wt_hi, const __nv_bfloat16* wt_lo,
                        cudaStream_t st);
int tc_tap_projection(const magat_gat_fwd_args* a, const __nv_bfloat16* h_hi, const __nv_bfloat16* h_lo,
                      cudaStream_t st);
int tc_split_weights(const float* src, long n, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st);
int tc_split_weights_t(const float* W, int G, int P, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st);

bool tap_tc_supported(const magat_gat_fwd_args* a);              // gat_tap_tc.cu
bool tap_tc_supported_no_y(const magat_gat_fwd_args* a);
bool score_tc_supported(const magat_gat_fwd_args* a);
int score_tc_forward(const magat_gat_fwd_args* a, float* wt, cudaStream_t st);
int tap_tc_forward(const magat_gat_fwd_args* a, cudaStream_t st, const HeadArgs* head);

// one level of the tap recursion, u_k from u_{k-1} (k >= 1), for every head
int run_tap_gather(const float* x, long x_sb, long x_sn, const float* att, const int32_t* nbr_in,
                   const int32_t* slot_in, int B, int N, int G, int K, int P, int D, int k, float* taps,
                   float* ain, int ain_ready, cudaStream_t st) {
  const long rows = (long)B * N;
  const int row_blocks = cdiv(rows, 8);
  const bool vec_ok = (G % 4 == 0) && (x_sn % 4 == 0) && (x_sb % 4 == 0) && (((uintptr_t)x) % 16 == 0) &&
                      (((uintptr_t)att) % 16 == 0) && (((uintptr_t)taps) % 16 == 0) && rows < (1l << 31);
  if (vec_ok && slot_in != nullptr && lean_sparse_ok(x_sb, x_sn, N, G, K, P, D) && ((uintptr_t)nbr_in % 16) == 0) {
#define MAGAT_GS(PT)                                                                                                   \
  do {                                                                                                                 \
    if (k == 1) k_tap_gather_s<PT, true><<<lean_grid(row_blocks), 256, 0, st>>>(x, x_sb, x_sn, att, nbr_in, slot_in, rows, N, K, D, k, taps); \
    else k_tap_gather_s<PT, false><<<lean_grid(row_blocks), 256, 0, st>>>(x, x_sb, x_sn, att, nbr_in, slot_in, rows, N, K, D, k, taps);       \
  } while (0)
    if (P == 4) MAGAT_GS(4);
    else if (P == 2) MAGAT_GS(2);
    else MAGAT_GS(1);
#undef MAGAT_GS
    return check_launch("k_tap_gather", st);
  }
#define MAGAT_GATHER(PT) \
  k_tap_gather_v<PT, 4><<<row_blocks, 256, 0, st>>>(x, x_sb, x_sn, att, nbr_in, slot_in, rows, N, G, K, D, k, taps, ain, \
                                                    ain_ready)
  if (vec_ok && P == 4) MAGAT_GATHER(4);
  else if (vec_ok && P == 2) MAGAT_GATHER(2);
  else if (vec_ok && P == 1) MAGAT_GATHER(1);
  else
    k_tap_gather<<<row_blocks, 256, 0, st>>>(x, x_sb, x_sn, att, nbr_in, slot_in, rows, N, G, P, K, D, k, taps);
#undef MAGAT_GATHER
  return check_launch("k_tap_gather", st);
}

static size_t simt_wprep_floats(int G, int P, int mode) {
  const size_t n = mode == MAGAT_MODE_GAT_MODIFIED ? (size_t)P * 2 * G + (size_t)P * 2 : 0;
  return (n + 3) & ~(size_t)3;     // keeps the bf16 region behind it 16 B aligned
}

static int forward_impl(const magat_gat_fwd_args* a, cudaStream_t st, bool use_tc, const HeadArgs* head = nullptr) {
  const int B = a->B, N = a->N, G = a->G, F = a->F, K = a->K, P = a->P, D = a->D;
  const long rows = (long)B * N;
  const int row_blocks = cdiv(rows, 8);
  const XLoad xl{a->x, a->x_sb, a->x_sn, N};
  const bool vec_ok = (G % 4 == 0) && (a->x_sn % 4 == 0) && (a->x_sb % 4 == 0) && (((uintptr_t)a->x) % 16 == 0) &&
                      (((uintptr_t)a->att) % 16 == 0) && (((uintptr_t)a->taps) % 16 == 0) &&
                      (((uintptr_t)a->sproj) % 16 == 0) && rows < (1l << 31);
  int rc;
  // scratch for the tcgen05 projections: transposed fp32 W (KeyQuery), then bf16 hi/lo copies of the weights
  const size_t nW = (size_t)P * G * G, nH = (size_t)P * F * K * G;
  float* wt_f32 = a->wprep + simt_wprep_floats(G, P, a->mode);
  __nv_bfloat16* tcw = reinterpret_cast<__nv_bfloat16*>(wt_f32 + (a->mode == MAGAT_MODE_KEYQUERY ? nW : 0));
  const bool fused = use_tc && tap_tc_supported(a);
  const bool score_fused = use_tc && score_tc_supported(a);
  __nv_bfloat16 *wt_hi = nullptr, *wt_lo = nullptr, *h_hi = tcw, *h_lo = tcw + nH;
  if (a->mode == MAGAT_MODE_KEYQUERY) { wt_hi = tcw; wt_lo = tcw + nW; h_hi = tcw + 2 * nW; h_lo = h_hi + nH; }
  if (use_tc) {
    if (a->mode == MAGAT_MODE_KEYQUERY && !score_fused &&
        (rc = tc_split_weights_t(a->weight, G, P, wt_hi, wt_lo, st)))
      return rc;
    if (!fused && (rc = tc_split_weights(a->filterWeight, (long)nH, h_hi, h_lo, st))) return rc;
  }
  // the attention kernel also scatters the receiver-major copy `ain` when the caller provides it
  // (the lean gather kernels read the senders' softmax rows through slot_in and need no receiver-major copy)
  const bool lean = vec_ok && lean_sparse_ok(a->x_sb, a->x_sn, N, G, K, P, D) &&
                    ((uintptr_t)a->nbr_in % 16) == 0 && ((uintptr_t)a->nbr_out % 16) == 0;
  const int32_t* so = (a->ain && a->slot_out && a->mode != MAGAT_MODE_GSO_VALUES && !lean) ? a->slot_out : nullptr;
  float* ain_w = so ? a->ain : nullptr;
  // 1. score projection
  if (a->mode == MAGAT_MODE_GSO_VALUES) {
    // non-attentional filter: att already holds the GSO values of the edges (magat_gso_edge_values)
  } else if (a->mode == MAGAT_MODE_KEYQUERY) {
    if (score_fused) {
      if ((rc = score_tc_forward(a, wt_f32, st))) return rc;
    } else if (use_tc) {
      if ((rc = tc_score_projection(a, wt_hi, wt_lo, st))) return rc;
    } else {
      dim3 grid(cdiv(rows, 64), cdiv((long)P * G, 64), 1);
      k_node_gemm<<<grid, 256, 0, st>>>(rows, P * G, G, xl, KqWLoad{a->weight, G}, StoreEpi{a->sproj, P * G});
      if ((rc = check_launch("k_node_gemm(score projection)", st))) return rc;
    }
    bool fast = vec_ok && D <= 32 && (G == 128 || G == 256);
#define MAGAT_ATT(PT, GV) \
  k_attention_kq_v<PT, GV, 8><<<row_blocks, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->sproj, a->nbr_out, rows, N, D, \
                                                          a->att, so, ain_w)
    // (four CTAs per SM = 64 registers without spills: 0.34 ms against 0.37 with ptxas left to itself, 0.38 at 3 or 5)
    if (lean && P == 4) k_attention_kq_h<4, 4><<<lean_grid(row_blocks), 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->sproj, a->nbr_out, rows, N, D, a->att);
    else if (lean && P == 2) k_attention_kq_h<2, 4><<<lean_grid(row_blocks), 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->sproj, a->nbr_out, rows, N, D, a->att);
    else if (lean && P == 1) k_attention_kq_h<1, 4><<<lean_grid(row_blocks), 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->sproj, a->nbr_out, rows, N, D, a->att);
    else if (fast && P == 4 && G == 128) MAGAT_ATT(4, 1);
    else if (fast && P == 4 && G == 256) MAGAT_ATT(4, 2);
    else if (fast && P == 2 && G == 128) MAGAT_ATT(2, 1);
    else if (fast && P == 2 && G == 256) MAGAT_ATT(2, 2);
    else if (fast && P == 1 && G == 128) MAGAT_ATT(1, 1);
    else if (fast && P == 1 && G == 256) MAGAT_ATT(1, 2);
    else
      k_attention<MAGAT_MODE_KEYQUERY><<<row_blocks, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->sproj, a->nbr_out,
                                                                    rows, N, G, P, D, a->att, so, ain_w);
#undef MAGAT_ATT
  } else {
    float* cvec = a->wprep;
    float* dvec = a->wprep + (size_t)P * 2 * G;
    k_gm_prep<<<P * 2, 128, 0, st>>>(a->weight, a->mixer, a->weight_bias, G, F, P, cvec, dvec);
    if ((rc = check_launch("k_gm_prep", st))) return rc;
    const bool gm_fast = vec_ok && G == 128 && D <= 32 && (P == 1 || P == 2 || P == 4) &&
                         (((uintptr_t)cvec) % 16 == 0);
    if (gm_fast) {
      const int per_warp = 16;
      const int blocks = cdiv(cdiv(rows, per_warp), 8);
#define MAGAT_GMX(PT) k_gm_mixer_v<PT><<<blocks, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, cvec, dvec, rows, N, per_warp, a->sproj)
      if (P == 4) MAGAT_GMX(4);
      else if (P == 2) MAGAT_GMX(2);
      else MAGAT_GMX(1);
#undef MAGAT_GMX
      if ((rc = check_launch("k_gm_mixer", st))) return rc;
#define MAGAT_GMA(PT) k_attention_gm_v<PT><<<row_blocks, 256, 0, st>>>(a->sproj, a->nbr_out, rows, N, D, a->att, so, ain_w)
#define MAGAT_GMH(PT) k_attention_gm_h<PT><<<lean_grid(row_blocks), 256, 0, st>>>(a->sproj, a->nbr_out, rows, N, D, a->att)
      if (so == nullptr && P == 4) MAGAT_GMH(4);
      else if (so == nullptr && P == 2) MAGAT_GMH(2);
      else if (so == nullptr && P == 1) MAGAT_GMH(1);
      else if (P == 4) MAGAT_GMA(4);
      else if (P == 2) MAGAT_GMA(2);
      else MAGAT_GMA(1);
#undef MAGAT_GMA
#undef MAGAT_GMH
    } else {
      dim3 grid(cdiv(rows, 64), cdiv(2l * P, 64), 1);
      k_node_gemm<<<grid, 256, 0, st>>>(rows, 2 * P, G, xl, GmCLoad{cvec, G}, GmEpi{a->sproj, dvec, 2 * P});
      if ((rc = check_launch("k_node_gemm(mixer projection)", st))) return rc;
      k_attention<MAGAT_MODE_GAT_MODIFIED><<<row_blocks, 256, 0, st>>>(a->x, a->x_sb, a->x_sn, a->sproj,
                                                                        a->nbr_out, rows, N, G, P, D, a->att, so, ain_w);
    }
  }
  if (a->mode != MAGAT_MODE_GSO_VALUES && (rc = check_launch("k_attention", st))) return rc;
  // 2. taps
  for (int k = 1; k < K; ++k)
    if ((rc = run_tap_gather(a->x, a->x_sb, a->x_sn, a->att, a->nbr_in, a->slot_in, B, N, G, K, P, D, k, a->taps,
                             ain_w ? a->ain : nullptr, ain_w ? 1 : 0, st)))
      return rc;
  // 3. per-(head, tap) projection + bias + activation + concat / head mean
  if (fused) return tap_tc_forward(a, st, head);
  if (head != nullptr) {
    set_error("magat_gat_forward_actions: shape not covered by the tcgen05 K-tap projection");
    return MAGAT_E_UNSUPPORTED;
  }
  if (use_tc) return tc_tap_projection(a, h_hi, h_lo, st);
  const int per_head = a->concat ? 1 : 0;
  const ZLoad zl{a->x, a->x_sb, a->x_sn, a->taps, N, G, K, P, per_head};
  const HLoad hl{a->filterWeight, G, K, F, per_head};
  const YEpi ye{a->y, a->y_sb, a->y_sn, a->y_sc, a->bias, N, F, a->relu, per_head ? 1.f : 1.f / (float)P};
  dim3 grid(cdiv(rows, 64), cdiv(F, 64), per_head ? P : 1);
  k_node_gemm<<<grid, 256, 0, st>>>(rows, F, per_head ? K * G : P * K * G, zl, hl, ye);
  return check_launch("k_node_gemm(tap projection)", st);
}

}  // namespace magat

using namespace magat;

extern "C" size_t magat_gat_wprep_floats(int G, int F, int K, int P, int mode) {
  return simt_wprep_floats(G, P, mode) + tc_wprep_floats(G, F, K, P, mode);
}

// number of tap planes (k = 1 .. K-1) magat_gat_forward leaves valid in a->taps for these arguments
extern "C" int magat_gat_forward_taps_valid(const magat_gat_fwd_args* a) {
  if (a == nullptr || a->K <= 1) return 0;
  return a->K - 1;
}

// the tcgen05 K-tap projection writes the ReLU bit mask from its epilogue; no other projection route does
extern "C" int magat_gat_forward_relu_bits_valid(const magat_gat_fwd_args* a) {
  if (a == nullptr || a->relu_bits == nullptr || !a->relu || a->path == MAGAT_PATH_SIMT) return 0;
  if (((uintptr_t)a->relu_bits % 16) != 0 || ((a->P * a->F) % 4) != 0) return 0;
  return tc_supported(a) && tap_tc_supported(a) ? 1 : 0;
}

// words of the bit mask: two 32-row groups per 64-row tile of the projection kernel
extern "C" size_t magat_gat_relu_bits_words(int B, int N, int P, int F) {
  const size_t rows = (size_t)B * N;
  return ((rows + 63) / 64) * 2 * (size_t)P * F;
}

// ---- heads averaged on top of the concat path (graphML.py:4665-4667) ------------------------------------------------
// y[b][f][n] = act((1/P) sum_p ycat[b*N + n][p*F + f]) as a contiguous [B][F][N] tensor: 32 x 32 tiles, read along f,
// transposed through shared memory, written along n.
__global__ void __launch_bounds__(256) k_head_mean_fwd(const float* __restrict__ ycat, int N, int P, int F, int relu,
                                                       float* __restrict__ y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // 32 x 8
  const float inv = 1.f / (float)P;
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r;
    float s = 0.f;
    if (n < N) {
      const float* src = ycat + ((size_t)b * N + n) * ((size_t)P * F) + f0 + tx;
      for (int p = 0; p < P; ++p) s += __ldcs(src + (size_t)p * F);
      s *= inv;
      if (relu) s = fmaxf(s, 0.f);
    }
    tile[r][tx] = s;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + tx;
    if (n < N) y[((size_t)b * F + f0 + r) * N + n] = tile[tx][r];
  }
}

// dycat[b*N + n][p*F + f] = dy[b][f][n] * (relu ? y[b][f][n] > 0 : 1) / P for every head p
__global__ void __launch_bounds__(256) k_head_mean_bwd(const float* __restrict__ dy, const float* __restrict__ y, int N,
                                                       int P, int F, int relu, float* __restrict__ dycat) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float inv = 1.f / (float)P;
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + tx;
    float g = 0.f;
    if (n < N) {
      const size_t i = ((size_t)b * F + f0 + r) * N + n;
      g = dy[i] * inv;
      if (relu && !(y[i] > 0.f)) g = 0.f;
    }
    tile[r][tx] = g;                                                 // [f][n]
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r;
    if (n < N) {
      const float g = tile[tx][r];
      float* dst = dycat + ((size_t)b * N + n) * ((size_t)P * F) + f0 + tx;
      for (int p = 0; p < P; ++p) __stcs(dst + (size_t)p * F, g);
    }
  }
}

// the layer and the planner's linear action head in one pass (inference).  y never reaches memory.
extern "C" int magat_gat_actions_supported(const magat_gat_fwd_args* a, int A) {
  if (a == nullptr || A < 1 || A > 8 || a->path == MAGAT_PATH_SIMT) return 0;
  return tc_supported(a) && tap_tc_supported_no_y(a) ? 1 : 0;
}

extern "C" int magat_gat_forward_actions(const magat_gat_fwd_args* a, const float* head_weight, const float* head_bias,
                                         int A, float* partial, float* logits, int32_t* actions_or_null, void* stream) {
  MAGAT_REQUIRE(a != nullptr && head_weight && partial && logits, MAGAT_E_BAD_ARG,
                "magat_gat_forward_actions: null pointer");
  int rc = validate_common(a->B, a->N, a->G, a->F, a->K, a->P, a->D, a->mode);
  if (rc) return rc;
  MAGAT_REQUIRE(a->x && a->nbr_out && a->nbr_in && a->slot_in && a->weight && a->filterWeight && a->att && a->sproj &&
                    a->wprep,
                MAGAT_E_BAD_ARG, "magat_gat_forward_actions: null pointer");
  MAGAT_REQUIRE(a->K == 1 || a->taps, MAGAT_E_BAD_ARG, "magat_gat_forward_actions: taps buffer missing for K=%d", a->K);
  MAGAT_REQUIRE(a->mode != MAGAT_MODE_GAT_MODIFIED || (a->mixer && a->weight_bias), MAGAT_E_BAD_ARG,
                "magat_gat_forward_actions: GAT_modified needs mixer and weight_bias");
  MAGAT_REQUIRE(magat_gat_actions_supported(a, A), MAGAT_E_UNSUPPORTED,
                "magat_gat_forward_actions: needs concatenated heads, F = 128, G multiple of 128, K <= 3, A <= 8");
  MAGAT_REQUIRE(((uintptr_t)partial % 16) == 0, MAGAT_E_BAD_ARG, "magat_gat_forward_actions: partial not 16 B aligned");
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const HeadArgs head{head_weight, head_bias, A, partial, logits, actions_or_null};
  return forward_impl(a, st, true, &head);
}

extern "C" int magat_head_mean_forward(const float* ycat, int B, int N, int P, int F, int relu, float* y, void* stream) {
  MAGAT_REQUIRE(ycat && y, MAGAT_E_BAD_ARG, "magat_head_mean_forward: null pointer");
  MAGAT_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && P >= 1 && F >= 32 && F % 32 == 0, MAGAT_E_BAD_ARG,
                "magat_head_mean_forward: B=%d N=%d P=%d F=%d (F must be a multiple of 32)", B, N, P, F);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  k_head_mean_fwd<<<dim3(cdiv(N, 32), F / 32, B), 256, 0, st>>>(ycat, N, P, F, relu, y);
  return check_launch("k_head_mean_fwd", st);
}

extern "C" int magat_head_mean_backward(const float* dy, const float* y, int B, int N, int P, int F, int relu,
                                        float* dycat, void* stream) {
  MAGAT_REQUIRE(dy && y && dycat, MAGAT_E_BAD_ARG, "magat_head_mean_backward: null pointer");
  MAGAT_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && P >= 1 && F >= 32 && F % 32 == 0, MAGAT_E_BAD_ARG,
                "magat_head_mean_backward: B=%d N=%d P=%d F=%d (F must be a multiple of 32)", B, N, P, F);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  k_head_mean_bwd<<<dim3(cdiv(N, 32), F / 32, B), 256, 0, st>>>(dy, y, N, P, F, relu, dycat);
  return check_launch("k_head_mean_bwd", st);
}

extern "C" int magat_gat_forward(const magat_gat_fwd_args* a, void* stream) {
  MAGAT_REQUIRE(a != nullptr, MAGAT_E_BAD_ARG, "magat_gat_forward: null args");
  int rc = validate_common(a->B, a->N, a->G, a->F, a->K, a->P, a->D, a->mode);
  if (rc) return rc;
  MAGAT_REQUIRE(a->x && a->nbr_out && a->nbr_in && a->slot_in && a->weight && a->filterWeight && a->y &&
                    a->att && a->sproj && a->wprep,
                MAGAT_E_BAD_ARG, "magat_gat_forward: null pointer");
  MAGAT_REQUIRE(a->K == 1 || a->taps, MAGAT_E_BAD_ARG, "magat_gat_forward: taps buffer missing for K=%d", a->K);
  MAGAT_REQUIRE(a->mode != MAGAT_MODE_GAT_MODIFIED || (a->mixer && a->weight_bias), MAGAT_E_BAD_ARG,
                "magat_gat_forward: GAT_modified needs mixer and weight_bias");
  MAGAT_REQUIRE(a->x_sn >= a->G, MAGAT_E_BAD_ARG, "magat_gat_forward: x row stride %ld < G", (long)a->x_sn);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  if (a->path == MAGAT_PATH_TCGEN05) {
    MAGAT_REQUIRE(tc_supported(a), MAGAT_E_UNSUPPORTED, "magat_gat_forward: shape not covered by the tcgen05 path");
    return forward_impl(a, st, true);
  }
  return forward_impl(a, st, a->path == MAGAT_PATH_AUTO && tc_supported(a));
}
