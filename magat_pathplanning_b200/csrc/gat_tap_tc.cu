// Fused K-tap projection kernel (concat heads, K <= 3, F = 128, K*G <= 384), tcgen05 "weights in TMEM":
//
//   Y_p^T [F x nodes] = H_p [F x K*G] . Z_p^T,   Z_p[n] = [ x_n | u_1^p[n] | u_2^p[n] ]
//
// Each persistent CTA owns ONE head p.  Its filter taps H_p, split into bf16 hi/lo, are written once
// into TMEM (K*G 32-bit columns: two K elements per column) and used as the A operand of every
// tcgen05.mma (TS form); shared memory is therefore free for the node-side operand, which the
// producer warps BUILD instead of load: x and u_1 rows straight from global, u_2 rows gathered on the
// fly, u_2[j] = sum_{i in in(j)} A_p[i,j] u_1^p[i], so the second tap never exists in HBM.  Three MMAs
// per 16-wide K step (hi.hi + lo.hi + hi.lo) give fp32-level accuracy.  The accumulator tile
// [F=128 lanes x 64 nodes] is double buffered in the remaining 128 TMEM columns; the epilogue warps
// add bias, apply ReLU, transpose through shared memory and store full 512 B rows of y.
//
//   warps 0-7   producers, two groups of four that alternate K chunks ([64 nodes x 64 k] -> hi/lo tiles,
//               SWIZZLE_128B K-major, 8-stage ring)
//   warps 8-11  epilogue (TMEM lanes q*32.. for warp q)
//   warp  12    MMA issuer (one thread)
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace magat {

namespace {

constexpr int TN = 64;                       // nodes per tile = UMMA N
constexpr int BK = 64;                       // K elements per chunk
constexpr int STAGES = 8;
constexpr int CH_BYTES = TN * BK * 2;        // 8 KB: one bf16 [64 x 64] tile
constexpr int STAGE_BYTES = 2 * CH_BYTES;    // hi + lo
constexpr int FT = 128;                      // out-features = UMMA M = TMEM lanes
constexpr int EPI_BYTES = TN * FT * 4;       // 32 KB transposition buffer
constexpr int PROD_WARPS = 8, PROD_GROUP = 128;
constexpr int EPI_WARP0 = 8, MMA_WARP = 12;
constexpr int THREADS = 13 * 32;
constexpr int ACC_COL0 = 384;                // accumulators: columns 384 + 64 a
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 256;

struct TapParams {
  long rows;             // B * N
  int N, G, K, P, D;
  const float* x; long x_sb, x_sn;
  const float* u1;       // taps buffer, tap k = 1 of head p of node m at u1 + (m*P + p)*(K-1)*G
  const float* ain; const int32_t* nbr_in;   // ain[m][p][s] = A_p[nbr_in[m][s], m]
  const float* H;        // filterWeight [P][F][K*G]
  const float* bias; int relu;
  float* y; long y_sb, y_sn;     // channel stride 1
};

__device__ __forceinline__ void ld8(const float* p, float* v) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__global__ void __launch_bounds__(THREADS, 1) k_tap_tc(const __grid_constant__ TapParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.P;
  const int slot = blockIdx.x / p.P, nslots = gridDim.x / p.P;
  const int KG = p.K * p.G;
  const int nchunks = KG / BK;
  const int cps = p.G / BK;                  // chunks per K segment
  const long tiles = (p.rows + TN - 1) / TN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full[s], PROD_GROUP);
      tc::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&acc_full[a], 1);
      tc::mbar_init(&acc_empty[a], 128);
    }
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- filter taps of this head -> TMEM (hi at columns [0, KG/2), lo at [KG/2, KG)) --------------
  if (warp < 4) {
    const int f = warp * 32 + lane;          // TMEM lane = output feature
    const float* hrow = p.H + ((size_t)head * FT + f) * KG;
    for (int k0 = 0; k0 < KG; k0 += 64) {
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(hrow + k0 + 2 * j));
        tc::split2(v.x, v.y, hi[j], lo[j]);
        tc::split2(v.z, v.w, hi[j + 1], lo[j + 1]);
      }
      const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
      tc::tmem_st32(lane_addr + (uint32_t)(k0 / 2), hi);
      tc::tmem_st32(lane_addr + (uint32_t)(KG / 2 + k0 / 2), lo);
    }
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  if (warp < PROD_WARPS) {
    // ===== producers ======================================================================
    const int grp = warp >> 2;
    const int tg = threadIdx.x & (PROD_GROUP - 1);
    const int c16 = tg & 7;
    const int r0 = tg >> 3;                   // rows r0 + 16 i, i < 4
    const long u1_row = (long)p.P * (p.K - 1) * p.G;
    long q = 0;                               // running chunk counter of the CTA
    for (long tile = slot; tile < tiles; tile += nslots) {
      const long m0 = tile * TN;
      for (int c = 0; c < nchunks; ++c, ++q) {
        if ((q & 1) != grp) continue;
        const int stage = (int)(q % STAGES);
        const uint32_t phase = (uint32_t)((q / STAGES) & 1);
        const int seg = c / cps;
        const int k0 = (c - seg * cps) * BK + c16 * 8;
        float v[4][8];
        if (seg < 2) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const long m = m0 + r0 + 16 * i;
            if (m < p.rows) {
              const float* src;
              if (seg == 0) {
                const long b = m / p.N;
                src = p.x + b * p.x_sb + (m - b * p.N) * p.x_sn + k0;
              } else {
                src = p.u1 + (m * p.P + head) * (long)(p.K - 1) * p.G + k0;
              }
              ld8(src, v[i]);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
            }
          }
        } else {
          // second tap gathered on the fly: u_2[j] = sum_i A_p[i,j] u_1^p[i].  Index and weight lists of
          // a row are contiguous (nbr_in, ain), read four entries at a time; the row loads of two rows x
          // four entries are issued together so a thread keeps 256 B in flight.
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
          }
          const float* u1h = p.u1 + (long)head * (p.K - 1) * p.G + k0;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            long mm[2], bN[2];
            bool valid[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              mm[j] = m0 + r0 + 16 * (2 * h + j);
              valid[j] = mm[j] < p.rows;
              bN[j] = valid[j] ? (mm[j] / p.N) * p.N : 0;
            }
            for (int s0 = 0; s0 < p.D; s0 += 4) {
              int id[2][4];
              float aw[2][4];
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                if (valid[j]) {
                  const int4 i4 = __ldg(reinterpret_cast<const int4*>(p.nbr_in + mm[j] * p.D + s0));
                  const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.ain + (mm[j] * p.P + head) * p.D + s0));
                  id[j][0] = i4.x; id[j][1] = i4.y; id[j][2] = i4.z; id[j][3] = i4.w;
                  aw[j][0] = a4.x; aw[j][1] = a4.y; aw[j][2] = a4.z; aw[j][3] = a4.w;
                } else {
#pragma unroll
                  for (int e = 0; e < 4; ++e) { id[j][e] = -1; aw[j][e] = 0.f; }
                }
              }
              if (id[0][0] < 0 && id[1][0] < 0) break;      // lists are packed: nothing further in either row
              float t[2][4][8];
#pragma unroll
              for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (id[j][e] >= 0) {
                    ld8(u1h + (bN[j] + id[j][e]) * u1_row, t[j][e]);
                  } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) t[j][e][c] = 0.f;
                  }
                }
#pragma unroll
              for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                  for (int c = 0; c < 8; ++c) v[2 * h + j][c] = fmaf(aw[j][e], t[j][e][c], v[2 * h + j][c]);
            }
          }
        }
        tc::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 hi, lo;
          tc::split2(v[i][0], v[i][1], hi.x, lo.x);
          tc::split2(v[i][2], v[i][3], hi.y, lo.y);
          tc::split2(v[i][4], v[i][5], hi.z, lo.z);
          tc::split2(v[i][6], v[i][7], hi.w, lo.w);
          const uint32_t off = tc::sw128_offset(r0 + 16 * i, c16);
          *reinterpret_cast<uint4*>(st + off) = hi;
          *reinterpret_cast<uint4*>(st + CH_BYTES + off) = lo;
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&full[stage]);
      }
    }
  } else if (warp < MMA_WARP) {
    // ===== epilogue =======================================================================
    const int qd = warp - EPI_WARP0;          // == warp % 4: TMEM lane quarter
    const int f = qd * 32 + lane;
    const float bias = p.bias ? __ldg(p.bias + f) : 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long tile = slot; tile < tiles; tile += nslots) {
      const long m0 = tile * TN;
      tc::mbar_wait(&acc_full[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(ACC_COL0 + acc * TN);
      float v[64];
      tc::tmem_ld32(taddr, v);
      tc::tmem_ld32(taddr + 32, v + 32);
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[acc]);        // accumulator drained into registers
#pragma unroll
      for (int n = 0; n < TN; ++n) {
        float o = v[n] + bias;
        if (p.relu) o = fmaxf(o, 0.f);
        epi[n * FT + f] = o;
      }
      tc::named_bar_sync(1, 128);
      // 64 rows x 512 B, each warp 16 rows, one float4 per lane
#pragma unroll 4
      for (int i = 0; i < 16; ++i) {
        const int n = qd * 16 + i;
        const long m = m0 + n;
        if (m < p.rows) {
          const long b = m / p.N;
          const float4 o = *reinterpret_cast<const float4*>(epi + n * FT + lane * 4);
          __stcs(reinterpret_cast<float4*>(p.y + b * p.y_sb + (m - b * p.N) * p.y_sn + (long)head * FT + lane * 4), o);
        }
      }
      tc::named_bar_sync(1, 128);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===== MMA issuer =====================================================================
    constexpr uint32_t idesc = tc::make_idesc_bf16(FT, TN);
    int acc = 0;
    uint32_t acc_phase = 0;
    long q = 0;
    for (long tile = slot; tile < tiles; tile += nslots) {
      tc::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(ACC_COL0 + acc * TN);
      for (int c = 0; c < nchunks; ++c, ++q) {
        const int stage = (int)(q % STAGES);
        const uint32_t phase = (uint32_t)((q / STAGES) & 1);
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        if (lane == 0) {
          const uint32_t sb = tc::smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint64_t z_hi = tc::make_sw128_desc(sb), z_lo = tc::make_sw128_desc(sb + CH_BYTES);
          const uint32_t h_hi = tmem_base + (uint32_t)(c * (BK / 2));
          const uint32_t h_lo = h_hi + (uint32_t)(KG / 2);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adv = (uint64_t)((kk * 32) >> 4);
            tc::umma_bf16_ts(tmem_d, h_hi + kk * 8, z_hi + adv, idesc, (c | kk) != 0);
            tc::umma_bf16_ts(tmem_d, h_lo + kk * 8, z_hi + adv, idesc, 1);
            tc::umma_bf16_ts(tmem_d, h_hi + kk * 8, z_lo + adv, idesc, 1);
          }
          tc::umma_commit(&empty[stage]);
          if (c == nchunks - 1) tc::umma_commit(&acc_full[acc]);
        }
        __syncwarp();
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

bool tap_tc_supported(const magat_gat_fwd_args* a) {
  if (!a->concat || a->K > 3 || a->F != FT) return false;
  if (a->G % BK != 0 || a->K * a->G > ACC_COL0) return false;
  if (a->y_sc != 1 || (a->y_sn % 4) != 0 || (a->y_sb % 4) != 0 || ((uintptr_t)a->y % 16) != 0) return false;
  if ((a->x_sn % 4) != 0 || (a->x_sb % 4) != 0 || ((uintptr_t)a->x % 16) != 0) return false;
  if (a->K > 1 && ((uintptr_t)a->taps % 16) != 0) return false;
  if (a->P > 64) return false;
  if (a->K > 2 && (a->ain == nullptr || a->D % 4 != 0 || ((uintptr_t)a->ain % 16) != 0 ||
                   ((uintptr_t)a->nbr_in % 16) != 0))
    return false;
  return true;
}

// needs tap k = 1 (u_1) in a->taps when K >= 2; never reads or writes tap k = 2
int tap_tc_forward(const magat_gat_fwd_args* a, cudaStream_t st) {
  static int sm_count = 0;
  static bool attr_set = false;
  if (!attr_set) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(k_tap_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(k_tap_tc): %s", cudaGetErrorString(e));
      return MAGAT_E_CUDA;
    }
    attr_set = true;
  }
  TapParams tp{};
  tp.rows = (long)a->B * a->N;
  tp.N = a->N; tp.G = a->G; tp.K = a->K; tp.P = a->P; tp.D = a->D;
  tp.x = a->x; tp.x_sb = a->x_sb; tp.x_sn = a->x_sn;
  tp.u1 = a->taps;
  tp.ain = a->ain; tp.nbr_in = a->nbr_in;
  tp.H = a->filterWeight;
  tp.bias = a->bias; tp.relu = a->relu;
  tp.y = a->y; tp.y_sb = a->y_sb; tp.y_sn = a->y_sn;
  const long tiles = (tp.rows + TN - 1) / TN;
  long slots = sm_count / a->P;
  if (slots < 1) slots = 1;
  if (slots > tiles) slots = tiles;
  k_tap_tc<<<(int)(slots * a->P), THREADS, SMEM_BYTES, st>>>(tp);
  return check_launch("k_tap_tc(fused taps + projection)", st);
}

}  // namespace magat
