// tcgen05 path of the two dense projections (score projection X W_p and the per-(head,tap)
// projection [x | u_1 | ... ] H_p^T), fp32-accurate through a bf16 hi/lo split:
//     a b ~= a_hi b_hi + a_lo b_hi + a_hi b_lo      (three UMMA passes, fp32 accumulation in TMEM)
// which measured 4-8e-6 max-norm relative error against the fp64 reference where a
// single TF32 or BF16 pass fails the 1e-4 bar.
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-3  producers: fp32 rows of A from global (coalesced 16 B loads) -> hi/lo bf16 -> shared memory
//              in the canonical SWIZZLE_128B K-major layout; pre-split B (weights) rows copied likewise
//   warps 4-7  epilogue: TMEM -> registers (tcgen05.ld) -> bias / ReLU -> global
//   warp  8    one elected thread issues tcgen05.mma and commits to the mbarriers
// Shared-memory ring of STAGES x {A_hi, A_lo, B_hi, B_lo} (64 KB each), two TMEM accumulators so the
// epilogue of tile t overlaps the MMAs of tile t+1.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace magat {

namespace {

constexpr int BM = 128;          // rows (nodes) per tile = UMMA M
constexpr int BN = 128;          // output columns per tile = UMMA N
constexpr int BK = 64;           // fp32 K elements per stage = one 128 B swizzle row of bf16
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 2;                 // 16 KB: one bf16 operand tile
constexpr int STAGE_BYTES = 4 * TILE_BYTES;             // A_hi, A_lo, B_hi, B_lo
constexpr int NUM_PRODUCER_WARPS = 4;
constexpr int NUM_EPI_WARPS = 4;
constexpr int MMA_WARP = NUM_PRODUCER_WARPS + NUM_EPI_WARPS;
constexpr int NUM_THREADS = (MMA_WARP + 1) * 32;
constexpr int TMEM_COLS = 2 * BN;
constexpr int MAX_SEGS = 16;
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

// One K-segment (G consecutive reduction indices) of a tile: where the fp32 A rows and the
// pre-split bf16 B rows come from.
struct Seg {
  const float* a;          // row m at a + (m / n_per_b) * a_sb + (m % n_per_b) * a_sn
  long a_sb, a_sn;
  const __nv_bfloat16* b_hi;   // B row n (output column) at b + n * b_ld, K contiguous
  const __nv_bfloat16* b_lo;
  long b_ld;
};

struct GemmParams {
  long M;                  // rows
  int n_per_b;             // rows per batch element (for the strided x addressing)
  int chunks_per_seg;      // G / BK
  int segs_per_tile;       // K-segments reduced into one output tile
  int Z;                   // independent z slices (heads); z picks segs [z*segs_per_tile, ...)
  int n_tiles;             // BN-wide column tiles per z
  int epi;                 // 0: store to c[m * ldc + z * z_cols + n]; 2: accumulate into it; 1: y epilogue
  float* c; long ldc; int z_cols;
  // y epilogue: v * scale + bias[n]; relu; y[b*y_sb + node*y_sn + (z*z_cols + n)*y_sc]
  float* y; long y_sb, y_sn, y_sc; const float* bias; int relu; float scale;
  Seg seg[MAX_SEGS];
};

__global__ void __launch_bounds__(NUM_THREADS, 1) k_tc_gemm(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]   producers -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES]   MMA (commit) -> producers
  uint64_t* acc_full = bars + 2 * STAGES;   // [2]        MMA (commit) -> epilogue
  uint64_t* acc_empty = acc_full + 2;       // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full[s], NUM_PRODUCER_WARPS * 32);
      tc::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&acc_full[a], 1);
      tc::mbar_init(&acc_empty[a], NUM_EPI_WARPS * 32);
    }
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc<TMEM_COLS>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long m_tiles = (p.M + BM - 1) / BM;
  const long tiles = m_tiles * p.Z * p.n_tiles;
  const int chunks = p.segs_per_tile * p.chunks_per_seg;

  if (warp < NUM_PRODUCER_WARPS) {
    // ===== producers ======================================================================
    const int t = threadIdx.x;               // 0..127
    const int c16 = t & 7;                   // 16 B chunk (8 bf16 = 8 fp32 source elements) of the 128 B row
    const int r0 = t >> 3;                   // rows r0 + 16 i
    int stage = 0;
    uint32_t phase = 0;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int nt = (int)(tile % p.n_tiles);
      const int z = (int)((tile / p.n_tiles) % p.Z);
      const long m0 = (tile / ((long)p.n_tiles * p.Z)) * BM;
      const int n0 = nt * BN;
      for (int ch = 0; ch < chunks; ++ch) {
        const int si = ch / p.chunks_per_seg;
        const int k0 = (ch - si * p.chunks_per_seg) * BK;
        const Seg& sg = p.seg[z * p.segs_per_tile + si];
        tc::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
        // A: 128 rows x 64 fp32 -> hi / lo tiles
        float4 va[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long m = m0 + r0 + 16 * i;
          if (m < p.M) {
            const unsigned b = (unsigned)m / (unsigned)p.n_per_b;          // 32-bit: M < 2^31 (checked on the host)
            const float4* src = reinterpret_cast<const float4*>(sg.a + (long)b * sg.a_sb +
                                                                (long)((unsigned)m - b * (unsigned)p.n_per_b) * sg.a_sn +
                                                                k0 + c16 * 8);
            va[i][0] = __ldg(src);
            va[i][1] = __ldg(src + 1);
          } else {
            va[i][0] = va[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        // B: 128 rows (output columns) x 64 bf16, already split
        uint4 vbh[8], vbl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const long off = (long)(n0 + r0 + 16 * i) * sg.b_ld + k0 + c16 * 8;
          vbh[i] = __ldg(reinterpret_cast<const uint4*>(sg.b_hi + off));
          vbl[i] = __ldg(reinterpret_cast<const uint4*>(sg.b_lo + off));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          uint4 hi, lo;
          tc::split2(va[i][0].x, va[i][0].y, hi.x, lo.x);
          tc::split2(va[i][0].z, va[i][0].w, hi.y, lo.y);
          tc::split2(va[i][1].x, va[i][1].y, hi.z, lo.z);
          tc::split2(va[i][1].z, va[i][1].w, hi.w, lo.w);
          const uint32_t off = tc::sw128_offset(r, c16);
          *reinterpret_cast<uint4*>(st + off) = hi;
          *reinterpret_cast<uint4*>(st + TILE_BYTES + off) = lo;
          *reinterpret_cast<uint4*>(st + 2 * TILE_BYTES + off) = vbh[i];
          *reinterpret_cast<uint4*>(st + 3 * TILE_BYTES + off) = vbl[i];
        }
        tc::fence_proxy_async();
        tc::mbar_arrive(&full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < MMA_WARP) {
    // ===== epilogue =======================================================================
    const int q = warp & 3;                   // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const int nt = (int)(tile % p.n_tiles);
      const int z = (int)((tile / p.n_tiles) % p.Z);
      const long m0 = (tile / ((long)p.n_tiles * p.Z)) * BM;
      const int n0 = nt * BN;
      const long m = m0 + row;
      tc::mbar_wait(&acc_full[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
      float* crow = nullptr;
      long ybase = 0;
      if (m < p.M) {
        if (p.epi != 1) {
          crow = p.c + m * p.ldc + (long)z * p.z_cols + n0;
        } else {
          const unsigned b = (unsigned)m / (unsigned)p.n_per_b;
          ybase = (long)b * p.y_sb + (long)((unsigned)m - b * (unsigned)p.n_per_b) * p.y_sn +
                  ((long)z * p.z_cols + n0) * p.y_sc;
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tc::tmem_ld32(taddr + c0, v);
        tc::tmem_ld_wait();
        if (m < p.M) {
          if (p.epi == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(crow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else if (p.epi == 2) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = *reinterpret_cast<float4*>(crow + c0 + j);
              o.x += v[j]; o.y += v[j + 1]; o.z += v[j + 2]; o.w += v[j + 3];
              *reinterpret_cast<float4*>(crow + c0 + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float o = v[j] * p.scale + (p.bias ? __ldg(p.bias + n0 + c0 + j) : 0.f);
              v[j] = p.relu ? fmaxf(o, 0.f) : o;
            }
            if (p.y_sc == 1) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(p.y + ybase + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) p.y[ybase + (long)(c0 + j) * p.y_sc] = v[j];
            }
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===== MMA issuer =====================================================================
    constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      tc::mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc::tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int ch = 0; ch < chunks; ++ch) {
        tc::mbar_wait(&full[stage], phase);
        tc::tc_fence_after();
        if (tc::elect_one()) {
          const uint32_t sa = tc::smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const uint64_t a_hi = tc::make_sw128_desc(sa), a_lo = tc::make_sw128_desc(sa + TILE_BYTES);
          const uint64_t b_hi = tc::make_sw128_desc(sa + 2 * TILE_BYTES), b_lo = tc::make_sw128_desc(sa + 3 * TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adv = (uint64_t)((kk * 32) >> 4);       // 16 bf16 = 32 B along K inside the swizzle row
            tc::umma_bf16(tmem_d, a_hi + adv, b_hi + adv, idesc, (ch | kk) != 0);
            tc::umma_bf16(tmem_d, a_lo + adv, b_hi + adv, idesc, 1);
            tc::umma_bf16(tmem_d, a_hi + adv, b_lo + adv, idesc, 1);
          }
          tc::umma_commit(&empty[stage]);                          // frees the smem stage when these MMAs retire
          if (ch == chunks - 1) tc::umma_commit(&acc_full[acc]);   // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// fp32 weights -> bf16 hi / lo (same element order)
__global__ void __launch_bounds__(256) k_split_weights(const float* __restrict__ src, long n,
                                                       __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}
// KeyQuery W[p][g][g'] -> B rows n = p*G + g' with K = g contiguous (R = X W_p), split
__global__ void __launch_bounds__(256) k_split_weights_t(const float* __restrict__ W, int G, long n,
                                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // i = (p*G + g') * G + g
  if (i >= n) return;
  const int g = (int)(i % G);
  const long pg = i / G;
  const int gp = (int)(pg % G);
  const long p = pg / G;
  const float v = W[(p * G + g) * G + gp];
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

int launch_gemm(const GemmParams& gp, cudaStream_t st, const char* what) {
  const int sm_count = device_sm_count();
  int rc0 = ensure_dyn_smem(KID_TC_GEMM, (const void*)k_tc_gemm, SMEM_BYTES, "k_tc_gemm");
  if (rc0) return rc0;
  const long tiles = ((gp.M + BM - 1) / BM) * gp.Z * gp.n_tiles;
  const int grid = (int)(tiles < sm_count ? tiles : sm_count);
  k_tc_gemm<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(gp);
  return check_launch(what, st);
}

}  // namespace

// wprep layout for the tcgen05 path (bf16 elements after the SIMT region, 16 B aligned):
//   KeyQuery: Wt_hi [P*G][G], Wt_lo, then H_hi [P][F][K*G], H_lo.   GAT_modified: cvec/dvec floats, then H_hi, H_lo.
size_t tc_wprep_floats(int G, int F, int K, int P, int mode) {
  size_t bf = 2 * (size_t)P * F * K * G;
  if (mode == MAGAT_MODE_KEYQUERY) bf += 2 * (size_t)P * G * G;
  return bf / 2 + 8 + (mode == MAGAT_MODE_KEYQUERY ? (size_t)P * G * G : 0);
}

bool tc_shape_ok(int G, int F, int K, int P, int concat) {
  if (G % BK != 0 || G > 1024) return false;
  if (F % BN != 0) return false;
  if (concat ? (K > MAX_SEGS / P ? (P * K > MAX_SEGS) : false) : (P * K > MAX_SEGS)) return false;
  if (P * K > MAX_SEGS) return false;
  return true;
}

bool tc_supported(const magat_gat_fwd_args* a) {
  if ((long)a->B * a->N >= (1l << 31)) return false;
  if (!tc_shape_ok(a->G, a->F, a->K, a->P, a->concat)) return false;
  if ((a->x_sn % 4) != 0 || (a->x_sb % 4) != 0 || ((uintptr_t)a->x % 16) != 0) return false;
  if (a->mode == MAGAT_MODE_KEYQUERY && ((long)a->P * a->G) % BN != 0) return false;
  return true;
}

// score projection (KeyQuery): sproj[m][p*G + g'] = sum_g x[m][g] W[p][g][g']
int tc_score_projection(const magat_gat_fwd_args* a, const __nv_bfloat16* wt_hi, const __nv_bfloat16* wt_lo,
                        cudaStream_t st) {
  GemmParams gp{};
  gp.M = (long)a->B * a->N;
  gp.n_per_b = a->N;
  gp.chunks_per_seg = a->G / BK;
  gp.segs_per_tile = 1;
  gp.Z = 1;
  gp.n_tiles = a->P * a->G / BN;
  gp.epi = 0;
  gp.c = a->sproj;
  gp.ldc = (long)a->P * a->G;
  gp.z_cols = 0;
  gp.seg[0] = Seg{a->x, a->x_sb, a->x_sn, wt_hi, wt_lo, (long)a->G};
  return launch_gemm(gp, st, "k_tc_gemm(score projection)");
}

// tap projection + bias + activation (+ concat / head mean)
int tc_tap_projection(const magat_gat_fwd_args* a, const __nv_bfloat16* h_hi, const __nv_bfloat16* h_lo,
                      cudaStream_t st) {
  const int G = a->G, F = a->F, K = a->K, P = a->P;
  GemmParams gp{};
  gp.M = (long)a->B * a->N;
  gp.n_per_b = a->N;
  gp.chunks_per_seg = G / BK;
  gp.segs_per_tile = a->concat ? K : P * K;
  gp.Z = a->concat ? P : 1;
  gp.n_tiles = F / BN;
  gp.epi = 1;
  gp.z_cols = F;
  gp.y = a->y; gp.y_sb = a->y_sb; gp.y_sn = a->y_sn; gp.y_sc = a->y_sc;
  gp.bias = a->bias;
  gp.relu = a->relu;
  gp.scale = a->concat ? 1.f : 1.f / (float)P;
  const long tap_row = (long)P * (K - 1) * G;           // floats per node in taps
  for (int p = 0; p < P; ++p)
    for (int k = 0; k < K; ++k) {
      Seg s;
      if (k == 0) {
        s.a = a->x; s.a_sb = a->x_sb; s.a_sn = a->x_sn;
      } else {
        s.a = a->taps + ((long)p * (K - 1) + (k - 1)) * G;
        s.a_sb = (long)a->N * tap_row; s.a_sn = tap_row;
      }
      const long boff = (long)p * F * K * G + (long)k * G;   // filterWeight[p][f][k][g], row stride K*G
      s.b_hi = h_hi + boff; s.b_lo = h_lo + boff; s.b_ld = (long)K * G;
      gp.seg[p * K + k] = s;
    }
  return launch_gemm(gp, st, "k_tc_gemm(tap projection)");
}

bool dx_tc_supported(const magat_gat_bwd_args* a) {
  if ((long)a->B * a->N >= (1l << 31)) return false;
  if (a->mode != MAGAT_MODE_KEYQUERY || a->G % BK != 0 || a->G % BN != 0 || a->P > MAX_SEGS) return false;
  if (((uintptr_t)a->rc % 16) != 0 || ((uintptr_t)a->dx % 16) != 0) return false;
  return true;
}

// KeyQuery: dx[m][g] += sum_{p,g'} dR[m][p][g'] W[p][g][g'];  w_hi / w_lo: bf16 split of `weight` as stored
int tc_dx_accumulate(const magat_gat_bwd_args* a, const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo,
                     cudaStream_t st) {
  const int G = a->G, P = a->P;
  GemmParams gp{};
  gp.M = (long)a->B * a->N;
  gp.n_per_b = a->N;
  gp.chunks_per_seg = G / BK;
  gp.segs_per_tile = P;
  gp.Z = 1;
  gp.n_tiles = G / BN;
  gp.epi = 2;
  gp.c = a->dx;
  gp.ldc = G;
  gp.z_cols = 0;
  for (int p = 0; p < P; ++p)
    gp.seg[p] = Seg{a->rc + (long)p * G, (long)a->N * P * G, (long)P * G, w_hi + (long)p * G * G,
                    w_lo + (long)p * G * G, (long)G};
  return launch_gemm(gp, st, "k_tc_gemm(dx += dR W^T)");
}

int tc_split_weights(const float* src, long n, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st) {
  k_split_weights<<<cdiv(n, 256), 256, 0, st>>>(src, n, hi, lo);
  return check_launch("k_split_weights", st);
}
int tc_split_weights_t(const float* W, int G, int P, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t st) {
  const long n = (long)P * G * G;
  k_split_weights_t<<<cdiv(n, 256), 256, 0, st>>>(W, G, n, hi, lo);
  return check_launch("k_split_weights_t", st);
}

}  // namespace magat
