// Fused K-tap projection kernel (concat heads, K <= 3, F = 128, K*G <= 384), tcgen05 "weights in TMEM":
//
//   Y_p^T [F x nodes] = H_p [F x K*G] . Z_p^T,   Z_p[n] = [ x_n | u_1^p[n] | u_2^p[n] ]
//
// Each persistent CTA owns ONE weight-block group (a head, or `nout` blocks that share one operand tile).  Its weights,
// split into bf16 hi/lo, are written once into TMEM (two K elements per 32-bit column) and used as the A operand of
// every tcgen05.mma (TS form); shared memory is therefore free for the node-side operand, which arrives by TMA tensor
// copies and is converted to bf16 hi/lo by shared-memory-only converter warps.  Three MMAs per 16-wide K step
// (hi.hi + lo.hi + hi.lo) give fp32-level accuracy.  The accumulator tile [F=128 lanes x 64 nodes] is double buffered
// in the remaining 128 TMEM columns; the epilogue warps add bias, apply ReLU and store 128 B lines straight from
// registers.  Also used with K = 1 for the KeyQuery score projection, the backward gz = dP H and the dense part of dx.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma_host.cuh"

namespace magat {

namespace {

constexpr int TN = 64;                       // nodes per tile = UMMA N
constexpr int SK = 128;                      // K elements per stage = two 64-wide swizzle atoms
constexpr int ATOM_BYTES = TN * 64 * 2;      // 8 KB: bf16 [64 rows x 64 k], SWIZZLE_128B
constexpr int STAGE_BYTES = 4 * ATOM_BYTES;  // hi atom 0/1, lo atom 0/1
constexpr int FT = 128;                      // out-features = UMMA M = TMEM lanes
constexpr int ACC_COL0 = 384;                // accumulators: columns 384 + 64 a

struct TapParams {
  // v2 kernel: 2-D views [rows][row stride] of x, the optional mask and the taps buffer; one box = 64 rows x 128 floats
  alignas(64) CUtensorMap tm_x;
  alignas(64) CUtensorMap tm_m;
  alignas(64) CUtensorMap tm_u;
  long rows;             // B * N
  int N, G, K, P, D;
  const float* x; long x_sb, x_sn;
  int x_hdiv; long x_hmul;               // head h reads x + (h / x_hdiv) * x_hmul (0 in the forward)
  const float* mask; long m_sb, m_sn;    // optional: x is zeroed where mask <= 0 (same head offset rule)
  // The same mask as one bit per element (what the forward epilogue below leaves in relu_bits_out): word
  // [m >> 5][c] holds rows 32 (m >> 5) .. + 31 of column c.  Takes precedence over `mask`: 1/32 of its bytes, no second
  // tensor copy per stage, and the raw ring keeps its four slots.
  const uint32_t* mask_bits; int bits_C;
  const float* u1;       // taps buffer, tap k = 1 of head p of node m at u1 + (m*P + p)*(K-1)*G
  const float* H;        // filterWeight [P][F][K*G]
  const float* bias; int relu;
  float* y; long y_sb, y_sn;     // channel stride 1
  int nout;                      // P counts weight blocks; nout consecutive blocks share one operand tile
  uint32_t* relu_bits_out; int bits_out_C;   // optional: bit (y > 0) per output element, layout as mask_bits
  // optional fused linear head (the planner's actionsMLP, decentralplanner_GAT.py:329-334): y is NOT
  // stored; head_partial[(p * rows + m) * 8 + a] = sum_f y[m][p*F + f] * head_w[a][p*F + f]  (a < head_A <= 8)
  const float* head_w; int head_A; float* head_partial;
  int accum;                     // y += result (the caller guarantees one weight-block group, i.e. no two CTAs ever touch
                                 // the same output row)
};

// ================================================================================================================
// Roles.  (A first version built the operand with SIMT producer warps straight from global memory; its source-level
// profile -- profiles/r01b_tap_tc_lsu.md -- showed the LSU data pipe 71 % busy and every role queueing behind it.)
// Here nothing but the unavoidable fp32 -> bf16 hi/lo conversion goes through the LSU:
//   warp 21      copy issuer: raw fp32 node rows -> shared-memory ring, one cp.async.bulk.tensor.2d per 64 x 128 box
//                (per-row 512 B bulk copies capped the K-tap use at ~70 cycles per row), mbarrier completion;
//                128 KB in flight
//   warps 0-15   converters, two groups of eight: LDS.128 of a raw row piece (conflict free), optional ReLU mask,
//                hi/lo split, two 8 B shared stores into the SWIZZLE_128B operand stage of the group
//   warp 20      MMA issuer; with nout > 1 the SAME operand stage is multiplied by nout weight blocks held in TMEM
//                (backward: the K taps of a head share dP; score projection: two heads share x), so the input is
//                read and converted once instead of nout times
//   warps 16-19  epilogue: tcgen05.ld, bias / ReLU, one coalesced 128 B store per node and warp
constexpr int V2_NOP = 2;                               // operand stages (one per converter group)
constexpr int V2_RAW_BYTES = 128 * 1024;                // raw ring: 4 slots of 32 KB, or 2 of 64 KB with a mask
constexpr int V2_CONV_WARPS = 16, V2_GROUP = 256;
constexpr int V2_EPI_WARP0 = 16, V2_EPI_WARPS = 4, V2_MMA_WARP = 20, V2_TMA_WARP = 21, V2_THREADS = 22 * 32;
// linear-head variant: warps 22-25 take the dot products over from the four epilogue warps (which then only drain TMEM
// into shared memory), 16 nodes at a time through a double-buffered tile
constexpr int V2_HEAD_WARP0 = 22, V2_HEAD_WARPS = 4, V2_THREADS_HEAD = 26 * 32;
// linear head: two [16 nodes][128 features] fp32 tiles of y (row stride 144 floats) + the head's [8][128] weight slice
constexpr int V2_YS_STRIDE = 144;
constexpr int V2_YS_NODES = 16;                               // nodes per hand-over
constexpr int V2_HEAD_BYTES = 2 * V2_YS_NODES * V2_YS_STRIDE * 4 + 8 * 128 * 4;
constexpr size_t V2_SMEM_BYTES = (size_t)V2_NOP * STAGE_BYTES + V2_RAW_BYTES + 1024 + 256 + V2_HEAD_BYTES;
static_assert(V2_SMEM_BYTES <= 227 * 1024, "k_tap_tc2 shared memory");


template <int STRIDE>
__device__ __forceinline__ void v2_store(float* dst, long stride_rt, const float (&v)[32], long left, float bias,
                                         bool relu) {
  if (relu) {
    if (left >= 32) {
#pragma unroll
      for (int n = 0; n < 32; ++n) __stcs(STRIDE ? dst + n * STRIDE : dst + n * stride_rt, fmaxf(v[n] + bias, 0.f));
    } else {
#pragma unroll
      for (int n = 0; n < 32; ++n)
        if (n < left) __stcs(STRIDE ? dst + n * STRIDE : dst + n * stride_rt, fmaxf(v[n] + bias, 0.f));
    }
  } else {
    if (left >= 32) {
#pragma unroll
      for (int n = 0; n < 32; ++n) __stcs(STRIDE ? dst + n * STRIDE : dst + n * stride_rt, v[n] + bias);
    } else {
#pragma unroll
      for (int n = 0; n < 32; ++n)
        if (n < left) __stcs(STRIDE ? dst + n * STRIDE : dst + n * stride_rt, v[n] + bias);
    }
  }
}

// MASK: 0 none, 1 fp32 mask tile next to every x tile, 2 bit mask read straight from global memory
// HEAD: the epilogue feeds the linear action head instead of storing y (its own instantiation: the register
// allocation of the other variants stays what it was)
template <int MASK, bool HEAD = false>
__global__ void __launch_bounds__(HEAD ? V2_THREADS_HEAD : V2_THREADS, 1) k_tap_tc2(const __grid_constant__ TapParams p) {
  constexpr bool MASKED = MASK == 1;
  constexpr int NRAW = MASKED ? 2 : 4;
  constexpr int SLOT_BYTES = V2_RAW_BYTES / NRAW;
  constexpr int TILE_BYTES = TN * SK * 4;               // 32 KB of fp32
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* raw = smem + (size_t)V2_NOP * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + V2_RAW_BYTES);
  uint64_t* op_full = bars;                  // [2]  256 converter arrivals
  uint64_t* op_empty = bars + 2;             // [2]  tcgen05.commit
  uint64_t* raw_full = bars + 4;             // [4]  expect_tx + bulk copies
  uint64_t* raw_empty = bars + 8;            // [4]  8 warp arrivals
  uint64_t* acc_full = bars + 12;            // [2]
  uint64_t* acc_empty = bars + 14;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  uint64_t* ys_full = bars + 20;             // [2]  128 drain threads have written the tile   (HEAD only)
  uint64_t* ys_empty = bars + 22;            // [2]  128 head threads have taken it
  float* ys = reinterpret_cast<float*>(raw + V2_RAW_BYTES + 1024 + 256);      // [2][16][V2_YS_STRIDE] (HEAD only)
  float* was = ys + 2 * V2_YS_NODES * V2_YS_STRIDE;                           // [8][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nout = p.nout;
  const int ngroups = p.P / nout;                        // head groups: a CTA serves heads hg*nout .. hg*nout+nout-1
  const int hg = blockIdx.x % ngroups;
  const int slot = blockIdx.x / ngroups, nslots = gridDim.x / ngroups;
  const int nst = (p.K * p.G) / SK;                      // 128-wide K slices per output tile
  const int sps = p.G / SK;
  const int KS = nst * SK;                               // reduction length of one output
  const int KGtot = nout * KS;                           // <= 384: weight columns hi [0, KGtot/2), lo [KGtot/2, KGtot)
  const long tiles = (p.rows + TN - 1) / TN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < V2_NOP; ++s) {
      tc::mbar_init(&op_full[s], V2_GROUP);
      tc::mbar_init(&op_empty[s], 1);
    }
    for (int s = 0; s < NRAW; ++s) {
      tc::mbar_init(&raw_full[s], 1);
      tc::mbar_init(&raw_empty[s], V2_GROUP / 32);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&acc_full[a], 1);
      tc::mbar_init(&acc_empty[a], V2_EPI_WARPS * 32);
      tc::mbar_init(&ys_full[a], V2_EPI_WARPS * 32);
      tc::mbar_init(&ys_empty[a], V2_HEAD_WARPS * 32);
    }
    tc::fence_barrier_init();
  }
  if (warp == V2_MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- weight blocks of this head group -> TMEM --------------------------------------------------------
  if (warp < 4) {
    const int f = warp * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int o = 0; o < nout; ++o) {
      const float* hrow = p.H + ((size_t)(hg * nout + o) * FT + f) * KS;
      for (int k0 = 0; k0 < KS; k0 += 32) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(hrow + k0 + 2 * j));
          tc::split2(v.x, v.y, hi[j], lo[j]);
          tc::split2(v.z, v.w, hi[j + 1], lo[j + 1]);
        }
        const uint32_t col = (uint32_t)((o * KS + k0) / 2);
        tc::tmem_st16(lane_addr + col, hi);
        tc::tmem_st16(lane_addr + (uint32_t)(KGtot / 2) + col, lo);
      }
    }
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  if (warp == V2_TMA_WARP) {
    // ===== copy issuer ====================================================================================
    // One tensor copy (cp.async.bulk.tensor.2d) per 64-row x 128-float box; rows past the end are zero filled by
    // the TMA unit, so every stage carries the same byte count.
    if (tc::elect_one()) {
      const int x_c0 = p.x_hmul ? (int)(((hg * nout) / p.x_hdiv) * p.x_hmul) : 0;
      const int u_c0 = (hg * nout) * (p.K - 1) * p.G;
      const uint64_t tmx = reinterpret_cast<uint64_t>(&p.tm_x);
      const uint64_t tmm = reinterpret_cast<uint64_t>(&p.tm_m);
      const uint64_t tmu = reinterpret_cast<uint64_t>(&p.tm_u);
      unsigned q = 0;
      for (long tile = slot; tile < tiles; tile += nslots) {
        const int m0 = (int)(tile * TN);
        for (int s = 0; s < nst; ++s, ++q) {
          const int r = (int)(q % NRAW);
          tc::mbar_wait(&raw_empty[r], ((q / NRAW) & 1u) ^ 1u);
          const uint32_t dst = tc::smem_u32(raw + (size_t)r * SLOT_BYTES);
          const int seg = s / sps;
          const int koff = (s - seg * sps) * SK;
          tc::mbar_arrive_expect_tx(&raw_full[r], (uint32_t)TILE_BYTES * (MASKED ? 2u : 1u));
          if (seg == 0) {
            tc::tensor_g2s_2d(dst, tmx, x_c0 + koff, m0, &raw_full[r]);
            if (MASKED) tc::tensor_g2s_2d(dst + TILE_BYTES, tmm, x_c0 + koff, m0, &raw_full[r]);
          } else {
            tc::tensor_g2s_2d(dst, tmu, u_c0 + (seg - 1) * p.G + koff, m0, &raw_full[r]);
          }
        }
      }
    }
  } else if (warp < V2_CONV_WARPS) {
    // ===== converters =====================================================================================
    // Group g (8 warps) owns operand stage g and builds every stage-chunk q with q % 2 == g.  Warp w of the group
    // converts rows w, w + 8, ...; lane l owns k = 4l .. 4l+3 of the 128-wide slice (one conflict-free LDS.128 per
    // row), i.e. half h = l & 1 of 16 B chunk c = (l >> 1) & 7 of swizzle atom l >> 4.
    const unsigned grp = warp >> 3;
    const int wg = warp & 7;
    const uint32_t op_s = tc::smem_u32(smem + (size_t)grp * STAGE_BYTES) + (uint32_t)((lane >> 4) * ATOM_BYTES) +
                          (uint32_t)((lane & 1) * 8);
    const int c = (lane >> 1) & 7;
    const uint32_t raw_s = tc::smem_u32(raw) + (uint32_t)(lane * 16);
    const int x_c0 = p.x_hmul ? (int)(((hg * nout) / p.x_hdiv) * p.x_hmul) : 0;
    unsigned q = 0;
    for (long tile = slot; tile < tiles; tile += nslots) {
      for (int s = 0; s < nst; ++s, ++q) {
        if ((q & 1u) != grp) continue;
        const int r = (int)(q % NRAW);
        // bit mask of this lane's four columns: rows 0-31 and 32-63 of the tile (issued before the wait on the copy)
        uint4 bw0 = make_uint4(~0u, ~0u, ~0u, ~0u), bw1 = bw0;
        if (MASK == 2 && s < sps) {
          const uint32_t* bp = p.mask_bits + (size_t)(tile * 2) * p.bits_C + x_c0 + s * SK + lane * 4;
          bw0 = __ldg(reinterpret_cast<const uint4*>(bp));
          bw1 = __ldg(reinterpret_cast<const uint4*>(bp + p.bits_C));
        }
        tc::mbar_wait(&raw_full[r], (q / NRAW) & 1u);
        tc::mbar_wait(&op_empty[grp], ((q >> 1) & 1u) ^ 1u);
        const uint32_t src = raw_s + (uint32_t)(r * SLOT_BYTES);
        constexpr int RB = MASKED ? 4 : 8;              // rows in flight per thread
#pragma unroll
        for (int r0 = 0; r0 < 8; r0 += RB) {
          float4 v[RB];
#pragma unroll
          for (int i = 0; i < RB; ++i) v[i] = tc::ld_shared_v4(src + (uint32_t)((wg + 8 * (r0 + i)) * (SK * 4)));
          if (MASKED) {
#pragma unroll
            for (int i = 0; i < RB; ++i) {
              const float4 mk = tc::ld_shared_v4(src + (uint32_t)TILE_BYTES + (uint32_t)((wg + 8 * (r0 + i)) * (SK * 4)));
              v[i].x = mk.x > 0.f ? v[i].x : 0.f; v[i].y = mk.y > 0.f ? v[i].y : 0.f;
              v[i].z = mk.z > 0.f ? v[i].z : 0.f; v[i].w = mk.w > 0.f ? v[i].w : 0.f;
            }
          }
          if (MASK == 2) {
#pragma unroll
            for (int i = 0; i < RB; ++i) {
              const uint4& bw = (r0 + i) < 4 ? bw0 : bw1;      // row wg + 8 (r0 + i): bit wg + 8 ((r0 + i) & 3) of its word
              const uint32_t bit = 1u << (wg + 8 * ((r0 + i) & 3));
              v[i].x = (bw.x & bit) ? v[i].x : 0.f; v[i].y = (bw.y & bit) ? v[i].y : 0.f;
              v[i].z = (bw.z & bit) ? v[i].z : 0.f; v[i].w = (bw.w & bit) ? v[i].w : 0.f;
            }
          }
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int row = wg + 8 * (r0 + i);
            uint32_t h0, l0, h1, l1;
            tc::split2(v[i].x, v[i].y, h0, l0);
            tc::split2(v[i].z, v[i].w, h1, l1);
            const uint32_t off = tc::sw128_offset(row, c);
            tc::st_shared_v2(op_s + off, h0, h1);
            tc::st_shared_v2(op_s + 2 * ATOM_BYTES + off, l0, l1);
          }
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&op_full[grp]);
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&raw_empty[r]);
      }
    }
  } else if (HEAD && warp >= V2_HEAD_WARP0) {
    // ===== linear head (nout == 1: one head per CTA) =======================================================
    // Warp hw takes nodes 4 hw .. + 3 of every 16-node tile the epilogue warps hand over, its lane l features
    // 4l .. 4l+3 of each.  The head's slice of every weight row sits in shared memory (zero rows past head_A).
    const int hw = warp - V2_HEAD_WARP0;
    const int t = hw * 32 + lane;
    {
      const int C = p.P * FT;
#pragma unroll
      for (int a = 0; a < 8; ++a) was[a * 128 + t] = a < p.head_A ? __ldg(p.head_w + (size_t)a * C + hg * FT + t) : 0.f;
      tc::named_bar_sync(1, V2_HEAD_WARPS * 32);
    }
    const uint32_t was_s = tc::smem_u32(was + lane * 4);
    unsigned part = 0;
    for (long tile = slot; tile < tiles; tile += nslots) {
      for (int s = 0; s < TN / V2_YS_NODES; ++s, ++part) {
        const int buf = (int)(part & 1u);
        tc::mbar_wait(&ys_full[buf], (part >> 1) & 1u);
        float4 yv[4];
        const uint32_t ys_s = tc::smem_u32(ys + (buf * V2_YS_NODES + hw * 4) * V2_YS_STRIDE + lane * 4);
#pragma unroll
        for (int n = 0; n < 4; ++n) yv[n] = tc::ld_shared_v4(ys_s + n * (V2_YS_STRIDE * 4));
        // all 8 (zero padded) actions x 4 nodes = 32 dot products, straight-line, then ONE joint reduction: lane
        // 4a + n ends up with action a of node n
        float d[32];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const float4 w = tc::ld_shared_v4(was_s + a * 512);
#pragma unroll
          for (int n = 0; n < 4; ++n)
            d[a * 4 + n] = fmaf(yv[n].x, w.x, fmaf(yv[n].y, w.y, fmaf(yv[n].z, w.z, yv[n].w * w.w)));
        }
        tc::mbar_arrive(&ys_empty[buf]);                  // every yv has been used: the tile may be overwritten
        const float tot = warp_multi_sum<32>(d, lane);
        const long m = tile * TN + s * V2_YS_NODES + hw * 4 + (lane & 3);
        if (m < p.rows) p.head_partial[((size_t)hg * p.rows + m) * 8 + (lane >> 2)] = tot;
      }
    }
  } else if (warp < V2_MMA_WARP) {
    // ===== epilogue =======================================================================================
    const int qd = warp - V2_EPI_WARP0;                  // TMEM lane quarter (= warp % 4)
    const int f = qd * 32 + lane;
    const float bias = p.bias ? __ldg(p.bias + f) : 0.f;
    unsigned oc = 0, part = 0;
    for (long tile = slot; tile < tiles; tile += nslots) {
      const long m0 = tile * TN;
      for (int o = 0; o < nout; ++o, ++oc) {
        const int acc = (int)(oc & 1u);
        tc::mbar_wait(&acc_full[acc], (oc >> 1) & 1u);
        tc::tc_fence_after();
        float* ybase = p.y + (long)(hg * nout + o) * FT + f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ACC_COL0 + acc * TN + 32 * hh);
          float v[32];
          tc::tmem_ld32(taddr, v);
          tc::tmem_ld_wait();
          if (hh == 1) {
            tc::tc_fence_before();
            tc::mbar_arrive(&acc_empty[acc]);
          }
          const long mrow = m0 + 32 * hh;
          float* dst = ybase + mrow * p.y_sn;
          const long left = p.rows - mrow;
          if (HEAD) {
            // y goes to the head warps through shared memory, feature-major in (thread = feature), 16 nodes at a time
#pragma unroll
            for (int s = 0; s < 32 / V2_YS_NODES; ++s, ++part) {
              const int buf = (int)(part & 1u);
              tc::mbar_wait(&ys_empty[buf], ((part >> 1) & 1u) ^ 1u);
              float* yd = ys + buf * V2_YS_NODES * V2_YS_STRIDE + f;
#pragma unroll
              for (int n = 0; n < V2_YS_NODES; ++n) {
                const float yv = v[s * V2_YS_NODES + n] + bias;
                yd[n * V2_YS_STRIDE] = p.relu ? fmaxf(yv, 0.f) : yv;
              }
              tc::mbar_arrive(&ys_full[buf]);
            }
            continue;
          }
          if (p.accum) {
#pragma unroll
            for (int n = 0; n < 32; ++n)
              if (n < left) v[n] += dst[(long)n * p.y_sn];
          }
          if (p.relu_bits_out) {
            // ReLU mask for the backward (one word per feature and 32 nodes; bits past the last row stay 0)
            uint32_t w = 0;
#pragma unroll
            for (int n = 0; n < 32; ++n) w |= (v[n] + bias > 0.f ? 1u : 0u) << n;
            if (left < 32) w = left > 0 ? w & ((1u << left) - 1u) : 0u;
            p.relu_bits_out[(size_t)(tile * 2 + hh) * p.bits_out_C + (hg * nout + o) * FT + f] = w;
          }
          // the usual row strides as compile-time constants: every store then carries its offset as an immediate
          // (FADD + FMNMX + STG per node instead of a 64-bit multiply-add chain -- the four epilogue warps pace the
          // one-slice uses of this kernel)
          if (p.y_sn == 512) v2_store<512>(dst, 0, v, left, bias, p.relu != 0);
          else if (p.y_sn == 1536) v2_store<1536>(dst, 0, v, left, bias, p.relu != 0);
          else if (p.y_sn == 256) v2_store<256>(dst, 0, v, left, bias, p.relu != 0);
          else v2_store<0>(dst, p.y_sn, v, left, bias, p.relu != 0);
        }
      }
    }
  } else if (warp == V2_MMA_WARP) {
    // ===== MMA issuer =====================================================================================
    constexpr uint32_t idesc = tc::make_idesc_bf16(FT, TN);
    unsigned q = 0, oc = 0;
    for (long tile = slot; tile < tiles; tile += nslots) {
      for (int o = 0; o < nout; ++o, ++oc) {
        const int acc = (int)(oc & 1u);
        tc::mbar_wait(&acc_empty[acc], ((oc >> 1) & 1u) ^ 1u);
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(ACC_COL0 + acc * TN);
        for (int s = 0; s < nst; ++s) {
          const unsigned qq = q + (unsigned)s;
          const int stage = (int)(qq & 1u);
          if (o == 0) {
            tc::mbar_wait(&op_full[stage], (qq >> 1) & 1u);
            tc::tc_fence_after();
          }
          if (tc::elect_one()) {
            const uint32_t sb = tc::smem_u32(smem + (size_t)stage * STAGE_BYTES);
            const uint32_t h_hi = tmem_base + (uint32_t)((o * nst + s) * (SK / 2));
            const uint32_t h_lo = h_hi + (uint32_t)(KGtot / 2);
#pragma unroll
            for (int at = 0; at < 2; ++at) {
              const uint64_t z_hi = tc::make_sw128_desc(sb + at * ATOM_BYTES);
              const uint64_t z_lo = tc::make_sw128_desc(sb + (2 + at) * ATOM_BYTES);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                const uint32_t col = (uint32_t)(at * 32 + kk * 8);
                tc::umma_bf16_ts(tmem_d, h_hi + col, z_hi + adv, idesc, (s | at | kk) != 0);
                tc::umma_bf16_ts(tmem_d, h_lo + col, z_hi + adv, idesc, 1);
                tc::umma_bf16_ts(tmem_d, h_hi + col, z_lo + adv, idesc, 1);
              }
            }
            if (o == nout - 1) tc::umma_commit(&op_empty[stage]);
            if (s == nst - 1) tc::umma_commit(&acc_full[acc]);
          }
          __syncwarp();
        }
      }
      q += (unsigned)nst;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == V2_MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// Wt[p][g'][g] = W[p][g][g']: the KeyQuery score projection R = X W_p as the same "weights in TMEM" GEMM
__global__ void __launch_bounds__(256) k_transpose_w(const float* __restrict__ W, int G, long n,
                                                     float* __restrict__ Wt) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // i = (p*G + g') * G + g
  if (i >= n) return;
  const int g = (int)(i % G);
  const long pg = i / G;
  const int gp = (int)(pg % G);
  const long p = pg / G;
  Wt[i] = W[(p * G + g) * G + gp];
}

// The kernel needs flat inputs and outputs (one row stride over the whole batch); nout > 1 only with one K slice.
// Fills the tensor maps of tq.
bool tap_tc2_prepare(TapParams& tq, int x_width) {
  if (tq.nout < 1 || tq.P % tq.nout != 0 || tq.rows >= (1l << 31)) return false;
  const int nst = tq.K * tq.G / SK;
  if (tq.G % SK != 0 || tq.nout * nst * SK > ACC_COL0 || (tq.nout > 1 && nst != 1)) return false;
  if (tq.y_sb != (long)tq.N * tq.y_sn) return false;
  if (tq.x_sb != (long)tq.N * tq.x_sn) return false;
  if (!tma::make_row_map(&tq.tm_x, tq.x, tq.rows, x_width, tq.x_sn, TN, SK)) return false;
  if (tq.mask_bits) {
    if (tq.bits_C % 4 != 0 || ((uintptr_t)tq.mask_bits % 16) != 0) return false;
    tq.mask = nullptr;
  }
  if (tq.mask) {
    if (tq.m_sb != (long)tq.N * tq.m_sn || !tma::make_row_map(&tq.tm_m, tq.mask, tq.rows, x_width, tq.m_sn, TN, SK)) return false;
  }
  if (tq.K > 1) {
    const long uw = (long)tq.P * (tq.K - 1) * tq.G;
    if (!tma::make_row_map(&tq.tm_u, tq.u1, tq.rows, uw, uw, TN, SK)) return false;
  }
  return true;
}

// Returns -1 when the layout cannot be described to the kernel (the *_supported predicates below rule that out for
// the callers that have no other route).
int launch_tap_tc(const TapParams& tp, cudaStream_t st, const char* what) {
  const int sm_count = device_sm_count();
  TapParams tq = tp;
  const long tiles = (tp.rows + TN - 1) / TN;
  // logical width of an x row: every weight-block group reads its own 128-wide (G-wide) column window
  const int x_width = tp.x_hmul ? (int)(((tp.P - 1) / tp.x_hdiv) * tp.x_hmul) + tp.G : tp.G;
  if (!tap_tc2_prepare(tq, x_width)) return -1;
  int rc0 = ensure_dyn_smem(KID_TAP_TC2, (const void*)k_tap_tc2<0>, V2_SMEM_BYTES, "k_tap_tc2<0>");
  if (!rc0) rc0 = ensure_dyn_smem(KID_TAP_TC2M, (const void*)k_tap_tc2<1>, V2_SMEM_BYTES, "k_tap_tc2<1>");
  if (!rc0) rc0 = ensure_dyn_smem(KID_TAP_TC2B, (const void*)k_tap_tc2<2>, V2_SMEM_BYTES, "k_tap_tc2<2>");
  if (!rc0) rc0 = ensure_dyn_smem(KID_TAP_TC2H, (const void*)k_tap_tc2<0, true>, V2_SMEM_BYTES, "k_tap_tc2<0, head>");
  if (rc0) return rc0;
  const int ngroups = tp.P / tp.nout;
  long slots = sm_count / ngroups;
  if (slots < 1) slots = 1;
  if (slots > tiles) slots = tiles;
  if (tq.head_partial) k_tap_tc2<0, true><<<(int)(slots * ngroups), V2_THREADS_HEAD, V2_SMEM_BYTES, st>>>(tq);
  else if (tq.mask_bits) k_tap_tc2<2><<<(int)(slots * ngroups), V2_THREADS, V2_SMEM_BYTES, st>>>(tq);
  else if (tq.mask) k_tap_tc2<1><<<(int)(slots * ngroups), V2_THREADS, V2_SMEM_BYTES, st>>>(tq);
  else k_tap_tc2<0><<<(int)(slots * ngroups), V2_THREADS, V2_SMEM_BYTES, st>>>(tq);
  return check_launch(what, st);
}

// -1 (layout not describable after the *_supported check passed: a tensor map could not be encoded) becomes an error
int must_launch(int rc, const char* what) {
  if (rc >= 0) return rc;
  set_error("%s: cuTensorMapEncodeTiled rejected the operand layout", what);
  return MAGAT_E_CUDA;
}

}  // namespace

// KeyQuery score projection sproj[m][p*G + g'] = sum_g x[m][g] W[p][g][g'] on the fused kernel (K = 1, no
// bias / activation).  wt is P*G*G floats of scratch.
bool score_tc_supported(const magat_gat_fwd_args* a) {
  if (a->mode != MAGAT_MODE_KEYQUERY || a->G != FT || a->P > 64) return false;
  if ((a->x_sn % 4) != 0 || (a->x_sb % 4) != 0 || ((uintptr_t)a->x % 16) != 0 || ((uintptr_t)a->sproj % 16) != 0)
    return false;
  if ((long)a->B * a->N * a->D * a->P >= (1l << 31)) return false;
  if (a->x_sb != (long)a->N * a->x_sn) return false;                   // flat rows (one stride over the whole batch)
  return true;
}

int score_tc_forward(const magat_gat_fwd_args* a, float* wt, cudaStream_t st) {
  const long n = (long)a->P * a->G * a->G;
  k_transpose_w<<<cdiv(n, 256), 256, 0, st>>>(a->weight, a->G, n, wt);
  int rc = check_launch("k_transpose_w", st);
  if (rc) return rc;
  TapParams tp{};
  tp.rows = (long)a->B * a->N;
  tp.N = a->N; tp.G = a->G; tp.K = 1; tp.P = a->P; tp.D = a->D;
  tp.x = a->x; tp.x_sb = a->x_sb; tp.x_sn = a->x_sn;
  tp.x_hdiv = 1; tp.x_hmul = 0; tp.mask = nullptr;
  tp.u1 = nullptr;
  tp.H = wt;
  tp.bias = nullptr; tp.relu = 0;
  tp.y = a->sproj; tp.y_sb = (long)a->N * a->P * a->G; tp.y_sn = (long)a->P * a->G;
  // heads that share one converted x tile (their W_p^T blocks sit side by side in TMEM)
  tp.nout = (a->P % 3 == 0) ? 3 : (a->P % 2 == 0) ? 2 : 1;
  return must_launch(launch_tap_tc(tp, st, "k_tap_tc(score projection)"), "score projection");
}

// Ht[(p*K + k)][g][f] = H[p][f][k][g]: the backward projection gz = dP H as the same kernel
__global__ void __launch_bounds__(256) k_transpose_h(const float* __restrict__ H, int F, int K, int G, long n,
                                                     float* __restrict__ Ht) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // i = ((p*K + k)*G + g)*F + f
  if (i >= n) return;
  const int f = (int)(i % F);
  long r = i / F;
  const int g = (int)(r % G);
  r /= G;
  const int k = (int)(r % K);
  const long p = r / K;
  Ht[i] = H[((p * F + f) * K + k) * G + g];
}

bool gz_tc_supported(const magat_gat_bwd_args* a) {
  if (!a->concat || a->G != FT || a->F % SK != 0 || a->F > ACC_COL0 || a->P * a->K > 64) return false;
  if ((long)a->K * a->F > ACC_COL0 || (a->K > 1 && a->F != SK)) return false;     // the K blocks of a head sit side by side in TMEM
  if (a->dy_sc != 1 || (a->dy_sn % 4) || (a->dy_sb % 4) || ((uintptr_t)a->dy % 16)) return false;
  if (a->relu && (a->y_sc != 1 || (a->y_sn % 4) || (a->y_sb % 4) || ((uintptr_t)a->y % 16))) return false;
  if (((uintptr_t)a->gz % 16) != 0) return false;
  if ((long)a->B * a->N * a->D * a->P >= (1l << 31)) return false;
  if (a->dy_sb != (long)a->N * a->dy_sn || (a->relu && a->y_sb != (long)a->N * a->y_sn)) return false;   // flat rows
  return true;
}

// gz[m][p][k][g] = sum_f dP[m][p*F + f] H[p][f][k][g], dP = dY * relu'(y).  ht: P*K*G*F floats of scratch.
int gz_tc_backward(const magat_gat_bwd_args* a, float* ht, cudaStream_t st) {
  const long n = (long)a->P * a->K * a->G * a->F;
  k_transpose_h<<<cdiv(n, 256), 256, 0, st>>>(a->filterWeight, a->F, a->K, a->G, n, ht);
  int rc = check_launch("k_transpose_h", st);
  if (rc) return rc;
  TapParams tp{};
  tp.rows = (long)a->B * a->N;
  // P*K weight blocks Ht[p*K + k]; the K blocks of a head share one converted dP_p tile (nout = K)
  tp.N = a->N; tp.G = a->F; tp.K = 1; tp.P = a->P * a->K; tp.D = a->D;
  tp.nout = a->K;
  tp.x = a->dy; tp.x_sb = a->dy_sb; tp.x_sn = a->dy_sn;
  tp.x_hdiv = a->K; tp.x_hmul = a->F;
  tp.mask = a->relu ? a->y : nullptr; tp.m_sb = a->y_sb; tp.m_sn = a->y_sn;
  if (a->relu && a->relu_bits) { tp.mask_bits = a->relu_bits; tp.bits_C = a->P * a->F; }
  tp.H = ht;
  tp.bias = nullptr; tp.relu = 0;
  tp.y = a->gz; tp.y_sn = (long)a->P * a->K * a->G; tp.y_sb = (long)a->N * tp.y_sn;
  return must_launch(launch_tap_tc(tp, st, "k_tap_tc(gz = dP H)"), "gz = dP H");
}

// Wcat[h][g][pl*G + g'] = W[2h + pl][g][g'] (block h starts at h * G * 2G; its rows are G * nsl long, nsl = heads in it)
__global__ void __launch_bounds__(256) k_pack_wcat(const float* __restrict__ W, int G, int P, float* __restrict__ wcat) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;       // i = (p*G + g)*G + g'
  if (i >= (long)P * G * G) return;
  const int gp = (int)(i % G);
  const long r = i / G;
  const int g = (int)(r % G);
  const int p = (int)(r / G);
  const int h = p >> 1, pl = p & 1;
  const int nsl = (P - 2 * h) >= 2 ? 2 : 1;
  wcat[(size_t)h * G * 2 * G + (size_t)g * (G * nsl) + pl * G + gp] = W[i];
}

// KeyQuery: the dense part of dx, sum_p dR_p W_p^T, as head-PAIR partial products on the TMA-fed kernel:
//   dxp[m][h][g] = sum_{pl < 2, g'} dR[m][2h + pl][g'] W[2h + pl][g][g']
// (four heads of weights do not fit TMEM next to the accumulators; two do).  One launch, plain stores -- the
// accumulating variant stalled its four epilogue warps on the read-modify-write latency (1.00 ms for two launches).
// The column kernel, which runs next and owns dx anyway, adds the P/2 partial rows.
bool dx_tap_supported(const magat_gat_bwd_args* a) {
  if (a->mode != MAGAT_MODE_KEYQUERY || a->G != FT || a->P % 2 != 0) return false;
  if (((uintptr_t)a->rc % 16) != 0 || ((uintptr_t)a->gz % 16) != 0 || ((uintptr_t)a->partial % 16) != 0) return false;
  if ((long)a->B * a->N >= (1l << 31)) return false;
  return true;
}

// wcat: P*G*G floats of scratch; dxp: rows * (P/2) * G floats.  Returns -1 when the v2 kernel cannot take the shapes.
int dx_tap_partials(const magat_gat_bwd_args* a, float* wcat, float* dxp, cudaStream_t st) {
  const int G = a->G, P = a->P;
  TapParams tp{};
  tp.rows = (long)a->B * a->N;
  tp.N = a->N; tp.G = 2 * G; tp.K = 1; tp.P = P / 2; tp.D = a->D; tp.nout = 1;
  tp.x = a->rc; tp.x_sn = (long)P * G; tp.x_sb = (long)a->N * tp.x_sn;
  tp.x_hdiv = 1; tp.x_hmul = 2 * G;
  tp.H = wcat;
  tp.y = dxp; tp.y_sn = (long)(P / 2) * G; tp.y_sb = (long)a->N * tp.y_sn;
  {
    TapParams t0 = tp;
    if (!tap_tc2_prepare(t0, P * G)) return -1;
  }
  k_pack_wcat<<<cdiv((long)P * G * G, 256), 256, 0, st>>>(a->weight, G, P, wcat);
  int rc = check_launch("k_pack_wcat", st);
  if (rc) return rc;
  return launch_tap_tc(tp, st, "k_tap_tc(dx partials = dR W^T)");
}

bool tap_tc_supported(const magat_gat_fwd_args* a) {
  if (!a->concat || a->K > 3 || a->F != FT) return false;
  if (a->G % SK != 0 || a->K * a->G > ACC_COL0) return false;
  if (a->y_sc != 1 || (a->y_sn % 4) != 0 || (a->y_sb % 4) != 0 || ((uintptr_t)a->y % 16) != 0) return false;
  if ((a->x_sn % 4) != 0 || (a->x_sb % 4) != 0 || ((uintptr_t)a->x % 16) != 0) return false;
  if (a->K > 1 && ((uintptr_t)a->taps % 16) != 0) return false;
  if (a->P > 64) return false;
  if ((long)a->B * a->N * a->D * a->P >= (1l << 31)) return false;     // 32-bit index math in the kernel
  if (a->x_sb != (long)a->N * a->x_sn || a->y_sb != (long)a->N * a->y_sn) return false;   // flat rows
  return true;
}

// logits[m][a] = b[a] + sum_p partial[p][m][a]; optional decode = argmax_a (softmax is monotone:
// utils/new_simulator.py:863-869), first maximum on ties like torch.max
__global__ void __launch_bounds__(256) k_actions_finalize(const float* __restrict__ partial, const float* __restrict__ b,
                                                          long rows, int P, int A, float* __restrict__ logits,
                                                          int32_t* __restrict__ actions) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  float acc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) acc[a] = (b != nullptr && a < A) ? __ldg(b + a) : 0.f;
  for (int h = 0; h < P; ++h) {
    const float4* q = reinterpret_cast<const float4*>(partial + ((size_t)h * rows + m) * 8);
    const float4 lo = __ldcs(q), hi = __ldcs(q + 1);
    acc[0] += lo.x; acc[1] += lo.y; acc[2] += lo.z; acc[3] += lo.w;
    acc[4] += hi.x; acc[5] += hi.y; acc[6] += hi.z; acc[7] += hi.w;
  }
  int best = 0;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    if (a < A) {
      logits[m * A + a] = acc[a];
      if (acc[a] > acc[best]) best = a;
    }
  }
  if (actions != nullptr) actions[m] = best;
}

// the same shapes when y is not written (linear head): its layout does not matter
bool tap_tc_supported_no_y(const magat_gat_fwd_args* a) {
  magat_gat_fwd_args t = *a;
  t.y = reinterpret_cast<float*>(uintptr_t(256));
  t.y_sc = 1; t.y_sn = (long)a->P * a->F; t.y_sb = (long)a->N * t.y_sn;
  return tap_tc_supported(&t);
}

// needs the taps k = 1 .. K-1 in a->taps.  head != null: y is not written, the logits of the linear head are.
int tap_tc_forward(const magat_gat_fwd_args* a, cudaStream_t st, const HeadArgs* head) {
  TapParams tp{};
  tp.rows = (long)a->B * a->N;
  tp.N = a->N; tp.G = a->G; tp.K = a->K; tp.P = a->P; tp.D = a->D;
  tp.x = a->x; tp.x_sb = a->x_sb; tp.x_sn = a->x_sn;
  tp.x_hdiv = 1; tp.x_hmul = 0; tp.mask = nullptr;
  tp.u1 = a->taps;
  tp.H = a->filterWeight;
  tp.bias = a->bias; tp.relu = a->relu;
  tp.y = a->y; tp.y_sb = a->y_sb; tp.y_sn = a->y_sn;
  if (a->relu && a->relu_bits) { tp.relu_bits_out = a->relu_bits; tp.bits_out_C = a->P * a->F; }
  tp.nout = 1;
  if (head != nullptr) {
    tp.head_w = head->w; tp.head_A = head->A; tp.head_partial = head->partial;
    tp.relu_bits_out = nullptr;
    tp.y = nullptr; tp.y_sn = (long)a->P * a->F; tp.y_sb = (long)a->N * tp.y_sn;      // never dereferenced
    int rc = must_launch(launch_tap_tc(tp, st, "k_tap_tc(fused taps + projection + linear head)"), "K-tap projection");
    if (rc) return rc;
    k_actions_finalize<<<cdiv(tp.rows, 256), 256, 0, st>>>(head->partial, head->b, tp.rows, a->P, head->A, head->logits,
                                                          head->actions);
    return check_launch("k_actions_finalize", st);
  }
  return must_launch(launch_tap_tc(tp, st, "k_tap_tc(fused taps + projection)"), "K-tap projection");
}

}  // namespace magat
