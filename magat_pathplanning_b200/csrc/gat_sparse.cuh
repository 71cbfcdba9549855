// Warp-level building blocks of the sparse phases, shared by the multi-launch forward (gat_fwd.cu) and the fused
// forward (gat_fused.cu).
#pragma once
#include "common.cuh"

namespace magat {
namespace sparse {

__device__ __forceinline__ void fma4(float4& acc, float a, const float4& v) {
  acc.x = fmaf(a, v.x, acc.x); acc.y = fmaf(a, v.y, acc.y); acc.z = fmaf(a, v.z, acc.z); acc.w = fmaf(a, v.w, acc.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// ---- row softmax over the out-neighbours + store -----------------------------------------------------------------
// The raw scores of the row sit in shared memory, e_s[slot * PT + head].  Lane = (slot within a chunk of 32 / PT
// slots, head): the max and the sum over the slots are xor-shuffles over the upper lane bits, all heads at once, and
// att[row][slot][head] leaves with one coalesced store per chunk (zeros beyond the degree).
template <int PT>
__device__ __forceinline__ void softmax_store(float* att_row, int D, int lane, int deg, const float* e_s) {
  constexpr int LOGP = PT == 4 ? 2 : PT == 2 ? 1 : 0;
  constexpr int SPC = 32 / PT;
  const int sl = lane >> LOGP;
  float m = -INFINITY;
  for (int c0 = 0; c0 < deg; c0 += SPC)
    if (c0 + sl < deg) m = fmaxf(m, e_s[c0 * PT + lane]);
#pragma unroll
  for (int o = PT; o < 32; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int c0 = 0; c0 < deg; c0 += SPC)
    if (c0 + sl < deg) sum += expf(e_s[c0 * PT + lane] - m);
#pragma unroll
  for (int o = PT; o < 32; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for (int c0 = 0; c0 < D; c0 += SPC) {
    const int sidx = c0 + sl;
    if (sidx < D) att_row[c0 * PT + lane] = sidx < deg ? expf(e_s[c0 * PT + lane] - m) / sum : 0.f;
  }
}

// Sum NV per-lane values over the 16 lanes of a half warp at once: every step halves the number of live values and
// doubles the lanes each has absorbed.  Lane t of the half ends up with the total of value t >> (4 - log2 NV).
template <int NV>
__device__ __forceinline__ float half_multi_sum(float (&v)[NV], int t) {
  int off = 8;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1, off >>= 1) {
    const bool hi = (t & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float keep = hi ? v[i + n / 2] : v[i];
      const float send = hi ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

template <bool CG>
__device__ __forceinline__ int list_entry(const int32_t* __restrict__ list_row, int D, int lane, int fill) {
  return lane < D ? (CG ? __ldcg(list_row + lane) : __ldg(list_row + lane)) : fill;
}

// KeyQuery scores + softmax of ONE sender row by one warp (graphML.py:1246-1286).  HALF a warp per edge: lane t of a
// half owns features 8t .. 8t+7 of R_i (all heads, in registers for the whole row) and of x_j, so the two halves
// score two edges per step and the cross-lane sum of the PT head dots is one joint reduction (16 instructions for 4
// heads instead of 40).  CG = true: R was written inside this launch by another CTA (bypass L1).
// The caller loads the row's out-list entry of this lane (my_j, -1 beyond D) -- one row ahead, so that the list load
// of the next row is in flight while this one is being scored.
template <int PT, bool CG>
__device__ __forceinline__ void attention_kq_row(const float* __restrict__ xb, unsigned x_sn, const float* __restrict__ r_row,
                                                 int my_j, float* att_row, int D, int lane, float* e_s) {
  constexpr int LOGP = PT == 4 ? 2 : PT == 2 ? 1 : 0;
  constexpr int G = 128;
  const int half = lane >> 4, t = lane & 15;
  const int deg = __popc(__ballot_sync(0xffffffffu, my_j >= 0));
  if (deg > 0) {
    float4 rv[PT][2];
#pragma unroll
    for (int h = 0; h < PT; ++h) {
      const float4* q = reinterpret_cast<const float4*>(r_row + h * G + t * 8);
      rv[h][0] = CG ? __ldcg(q) : __ldg(q);
      rv[h][1] = CG ? __ldcg(q + 1) : __ldg(q + 1);
    }
    for (int s0 = 0; s0 < deg; s0 += 2) {
      const int sidx = s0 + half;
      const int j = __shfl_sync(0xffffffffu, my_j, sidx & 31);
      float d[PT];
      if (sidx < deg) {
        const float* xr = xb + (unsigned)j * x_sn + t * 8;
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(xr));
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(xr + 4));
#pragma unroll
        for (int h = 0; h < PT; ++h) d[h] = dot4(rv[h][0], x0) + dot4(rv[h][1], x1);
      } else {
#pragma unroll
        for (int h = 0; h < PT; ++h) d[h] = 0.f;
      }
      const float tot = half_multi_sum<PT>(d, t);
      if (sidx < deg && (t & (16 / PT - 1)) == 0) e_s[sidx * PT + (t >> (4 - LOGP))] = tot;
    }
  }
  __syncwarp();
  softmax_store<PT>(att_row, D, lane, deg, e_s);
  __syncwarp();
}

// One level of the tap recursion for ONE receiver by one warp: u_k[j] = sum_{i in in(j)} A_p[i,j] u_{k-1}[i]
// (graphML.py:1756-1759), all heads at once; lane l owns features 4l .. 4l+3.  Lane s keeps in-edge s: the sender id
// and -- read through slot_in, the position of j in the sender's out-list -- its PT attention values (one 16 B
// load); the edge loop broadcasts them by shuffle.  K1 = true gathers rows of x (one row feeds every head, two edges
// in flight), else the heads' rows src[(i * PT + h) * trow] of the previous level.
// in-edge `lane` of a receiver: A_p[i, j] for all heads sits at slot my_sl of sender my_i's softmax row
template <int PT, bool CG>
__device__ __forceinline__ void edge_weights(const float* __restrict__ att_b, int my_i, int my_sl, int D, float (&am)[PT]) {
  if (PT == 4) {
    const float4* q = reinterpret_cast<const float4*>(att_b + (unsigned)((my_i * D + my_sl) * PT));
    const float4 wv = my_i >= 0 ? (CG ? __ldcg(q) : __ldg(q)) : make_float4(0.f, 0.f, 0.f, 0.f);
    am[0] = wv.x; am[1 % PT] = wv.y; am[2 % PT] = wv.z; am[3 % PT] = wv.w;
  } else {
#pragma unroll
    for (int h = 0; h < PT; ++h) {
      const float* q = att_b + (unsigned)((my_i * D + my_sl) * PT + h);
      am[h] = my_i >= 0 ? (CG ? __ldcg(q) : __ldg(q)) : 0.f;
    }
  }
}

// my_i: this lane's in-list entry (loaded by the caller one row ahead), am: its attention values (edge_weights).
template <int PT, bool K1, bool CG>
__device__ __forceinline__ void gather_row(const float* __restrict__ xb, unsigned x_sn, const float* __restrict__ src,
                                           unsigned trow, int my_i, const float (&am)[PT], int lane, float4 (&acc)[PT]) {
  const int cnt = __popc(__ballot_sync(0xffffffffu, my_i >= 0));
#pragma unroll
  for (int h = 0; h < PT; ++h) acc[h] = make_float4(0.f, 0.f, 0.f, 0.f);
  int s = 0;
  if (K1) {
    for (; s + 1 < cnt; s += 2) {
      const int i0 = __shfl_sync(0xffffffffu, my_i, s), i1 = __shfl_sync(0xffffffffu, my_i, s + 1);
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(xb + (unsigned)i0 * x_sn + lane * 4));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(xb + (unsigned)i1 * x_sn + lane * 4));
#pragma unroll
      for (int h = 0; h < PT; ++h) {
        fma4(acc[h], __shfl_sync(0xffffffffu, am[h], s), v0);
        fma4(acc[h], __shfl_sync(0xffffffffu, am[h], s + 1), v1);
      }
    }
  }
  for (; s < cnt; ++s) {
    const int i0 = __shfl_sync(0xffffffffu, my_i, s);
    if (K1) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(xb + (unsigned)i0 * x_sn + lane * 4));
#pragma unroll
      for (int h = 0; h < PT; ++h) fma4(acc[h], __shfl_sync(0xffffffffu, am[h], s), v0);
    } else {
      float4 v0[PT];
#pragma unroll
      for (int h = 0; h < PT; ++h) {
        const float4* q = reinterpret_cast<const float4*>(src + ((unsigned)i0 * PT + h) * trow + lane * 4);
        v0[h] = CG ? __ldcg(q) : __ldg(q);
      }
#pragma unroll
      for (int h = 0; h < PT; ++h) fma4(acc[h], __shfl_sync(0xffffffffu, am[h], s), v0[h]);
    }
  }
}

}  // namespace sparse
}  // namespace magat
