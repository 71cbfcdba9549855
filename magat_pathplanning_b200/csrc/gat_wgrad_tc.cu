// Weight gradients on tcgen05: reductions over the NODE axis,
//     dC[z][i][j] = sum_m  A[m][z*az + i] * B_z[m][j],        i < 128, j < NB <= 384
// which is dfilterWeight (A = dP = dY * relu'(y), B = the stacked taps [x | u_1 | u_2] of head z) and, for
// KeyQuery, dweight (A = x, B = dR_z).  Both operands live in memory node-major (features contiguous),
// so the reduction index is the SLOW index: the tiles are stored exactly as they are read -- 128 B rows of
// 64 bf16 features per node -- and handed to the tensor core as MN-major operands (SWIZZLE_128B, LBO =
// stride between 64-feature atoms, SBO = stride between 8-node groups); no transposition anywhere.
// fp32 accuracy comes from the same three-pass bf16 hi/lo split as the forward kernels.
//
// Persistent CTA per (z, slot): accumulates its share of the nodes into TMEM (128 lanes x NB columns, fp32)
// over the whole kernel, then writes ONE partial tile; a tiny kernel sums the partials (deterministic).
//   warps 0-15 producers, two groups of eight that build alternate 32-node stages {A hi/lo [32 x 128], B hi/lo
//              [32 x NB]} of a 3-stage ring.  A thread's loads sit in registers until its stage is free, so ONE group
//              keeps only one stage of loads (80 KB per SM) in flight; the ncu profile showed the kernel waiting on
//              exactly those loads at 3.7 TB/s, hence the second group
//   warp  16   MMA issuer: per 16-node step 3 x (N=256 [+ N=128]) tcgen05.mma, both operands from smem
//   warps 0-3  double as the epilogue at the end (TMEM -> partial tile)
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "tma_host.cuh"

namespace magat {

namespace {

constexpr int SN = 32;                         // nodes per stage (MMA K = 16 -> 2 steps)
constexpr int MI = 128;                        // rows of dC = UMMA M
constexpr int MAX_NB = 384;
constexpr int A_BYTES = SN * MI * 2;           // 8 KB per hi or lo
constexpr int B_BYTES = SN * MAX_NB * 2;       // 24 KB per hi or lo
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 64 KB
constexpr int STAGES = 3;
constexpr int ATOM_STRIDE = SN * 128;          // bytes between 64-feature atoms (LBO)
constexpr int PROD_THREADS = 256;              // per group
constexpr int PROD_GROUPS = 2;
constexpr int MMA_WARP = 16;
constexpr int THREADS = 17 * 32;
constexpr int RED_BYTES = PROD_GROUPS * 16 * MI * 4;          // column-sum staging (dbias)
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + RED_BYTES + 1024 + 256;

struct Src {                 // node-major fp32 rows: row m at p + (m / N) * sb + (m % N) * sn  (+ z * zoff)
  const float* p; long sb, sn, zoff;
};

struct WgradParams {
  long rows; int N;
  int Z, NB;                 // slices (heads), columns of dC per slice (multiple of 128, <= 384)
  Src a;                     // 128 features per node
  Src b[3];                  // NB / 128 segments of 128 features
  const float* mask_y; long my_sb, my_sn, my_zoff;   // optional: A = a * (mask_y > 0)  (dP from dY and y)
  const uint32_t* mask_bits; int bits_C;             // the same mask, one bit per element (v2 kernel only; preferred)
  float a_scale;
  float* partial;            // [Z][nslots][128][NB]
  float* colsum_partial;     // optional [Z][nslots][128]: per-CTA column sums of the (masked, scaled) A rows
};

// MN-major SWIZZLE_128B operand: LBO = bytes between 64-element atoms along M/N, SBO = bytes between
// 8-row groups along K (1024), version 1, layout type 2.
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(ATOM_STRIDE >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, D = F32, A = B = BF16, both MN-major (bits 15, 16 = 1)
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(THREADS, 1) k_wgrad_tc(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  float* red = reinterpret_cast<float*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES + RED_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* done = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.x % p.Z;
  const int slot = blockIdx.x / p.Z, nslots = gridDim.x / p.Z;
  const long nstages = (p.rows + SN - 1) / SN;
  const int nseg = p.NB / 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full[s], PROD_THREADS);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < MMA_WARP) {
    // ===== producers ======================================================================
    const int t = threadIdx.x & (PROD_THREADS - 1);     // 0..255 inside the group
    const int grp = threadIdx.x / PROD_THREADS;
    const int c = t & 15;                      // 16 B chunk (8 features) of the 128-feature row
    const int r0 = t >> 4;                     // rows r0, r0 + 16
    const unsigned N = (unsigned)p.N;
    const int rows = (int)p.rows;
    // byte offset inside an operand region: atom (c / 8), row r, chunk (c % 8) swizzled by the row
    uint32_t off[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = r0 + 16 * i;
      off[i] = (uint32_t)((c >> 3) * ATOM_STRIDE + r * 128 + (((c & 7) ^ (r & 7)) << 4));
    }
    int stage = 0;
    uint32_t phase = 0;
    float4 cs0 = make_float4(0.f, 0.f, 0.f, 0.f), cs1 = cs0;      // column sums of A (features c*8 .. c*8+7)
    int it = 0;
    for (long sidx = slot; sidx < nstages; sidx += nslots, ++it) {
      if ((it & (PROD_GROUPS - 1)) != grp) {   // the other group's stage
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        continue;
      }
      const int m0 = (int)(sidx * SN);
      float4 va[4][2][2];                      // [source][row][half]
      float4 vm[2][2];
      bool has_mask = p.mask_y != nullptr;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int m = m0 + r0 + 16 * i;
        const bool live = m < rows;
        const unsigned mu = live ? (unsigned)m : 0u;
        const unsigned b = mu / N, n = mu - b * N;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const Src& sc = s == 0 ? p.a : p.b[s - 1];
          if (live && (s == 0 || s - 1 < nseg)) {
            const float4* src = reinterpret_cast<const float4*>(sc.p + (long)b * sc.sb + (long)n * sc.sn +
                                                                (long)z * sc.zoff + c * 8);
            va[s][i][0] = __ldg(src);
            va[s][i][1] = __ldg(src + 1);
          } else {
            va[s][i][0] = va[s][i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if (has_mask && live) {
          const float4* src = reinterpret_cast<const float4*>(p.mask_y + (long)b * p.my_sb + (long)n * p.my_sn +
                                                              (long)z * p.my_zoff + c * 8);
          vm[i][0] = __ldg(src);
          vm[i][1] = __ldg(src + 1);
        } else {
          vm[i][0] = vm[i][1] = make_float4(1.f, 1.f, 1.f, 1.f);
        }
      }
      // A = a * scale * (y > 0)
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4& v = va[0][i][h];
          const float4 mk = vm[i][h];
          v.x = mk.x > 0.f ? v.x * p.a_scale : 0.f;
          v.y = mk.y > 0.f ? v.y * p.a_scale : 0.f;
          v.z = mk.z > 0.f ? v.z * p.a_scale : 0.f;
          v.w = mk.w > 0.f ? v.w * p.a_scale : 0.f;
          float4& cs = h == 0 ? cs0 : cs1;
          cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
        }
      tc::mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* st = smem + (size_t)stage * STAGE_BYTES;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s > 0 && s - 1 >= nseg) continue;
        uint8_t* hi_base = s == 0 ? st : st + 2 * A_BYTES + (s - 1) * 2 * ATOM_STRIDE;
        uint8_t* lo_base = s == 0 ? st + A_BYTES : hi_base + B_BYTES;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint4 hi, lo;
          tc::split2(va[s][i][0].x, va[s][i][0].y, hi.x, lo.x);
          tc::split2(va[s][i][0].z, va[s][i][0].w, hi.y, lo.y);
          tc::split2(va[s][i][1].x, va[s][i][1].y, hi.z, lo.z);
          tc::split2(va[s][i][1].z, va[s][i][1].w, hi.w, lo.w);
          *reinterpret_cast<uint4*>(hi_base + off[i]) = hi;
          *reinterpret_cast<uint4*>(lo_base + off[i]) = lo;
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&full[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    // column sums (dbias): 16 row-threads per feature chunk -> smem -> fixed-order sum
    if (p.colsum_partial != nullptr) {
      *reinterpret_cast<float4*>(red + (grp * 16 + r0) * MI + c * 8) = cs0;
      *reinterpret_cast<float4*>(red + (grp * 16 + r0) * MI + c * 8 + 4) = cs1;
      tc::named_bar_sync(2, PROD_GROUPS * PROD_THREADS);
      if (threadIdx.x < MI) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < PROD_GROUPS * 16; ++r) s += red[r * MI + threadIdx.x];
        p.colsum_partial[((size_t)z * nslots + slot) * MI + threadIdx.x] = s;
      }
    }
    // ===== epilogue (warps 0-3): the accumulated tile -> this CTA's partial =====================
    if (warp < 4) {
      tc::mbar_wait(done, 0);
      tc::tc_fence_after();
      const int i_row = warp * 32 + lane;
      float* dst = p.partial + (((size_t)z * nslots + slot) * MI + i_row) * p.NB;
      const bool any = slot < nstages;           // a CTA without work never touched TMEM: write zeros
      for (int c0 = 0; c0 < p.NB; c0 += 32) {
        float v[32];
        if (any) {
          tc::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  } else {
    // ===== MMA issuer =====================================================================
    const uint32_t idesc256 = make_idesc_mn(MI, 256), idesc128 = make_idesc_mn(MI, 128);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (long sidx = slot; sidx < nstages; sidx += nslots) {
      tc::mbar_wait(&full[stage], phase);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t sb = tc::smem_u32(smem + (size_t)stage * STAGE_BYTES);
        const uint32_t a_hi = sb, a_lo = sb + A_BYTES, b_hi = sb + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
        for (int kk = 0; kk < SN / 16; ++kk) {
          const uint32_t kofs = (uint32_t)(kk * 2 * 1024);          // 16 nodes = two 8-row groups
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {
            const uint32_t aa = (pass == 1 ? a_lo : a_hi) + kofs;
            const uint32_t bb = (pass == 2 ? b_lo : b_hi) + kofs;
            const uint32_t accum = (first && kk == 0 && pass == 0) ? 0u : 1u;
            if (p.NB >= 256) {
              tc::umma_bf16(tmem_base, make_mn_desc(aa), make_mn_desc(bb), idesc256, accum);
              if (p.NB > 256)
                tc::umma_bf16(tmem_base + 256, make_mn_desc(aa), make_mn_desc(bb + 4 * ATOM_STRIDE), idesc128, accum);
            } else {
              tc::umma_bf16(tmem_base, make_mn_desc(aa), make_mn_desc(bb), idesc128, accum);
            }
          }
        }
        tc::umma_commit(&empty[stage]);
        first = false;
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
    if (tc::elect_one()) tc::umma_commit(done);
    __syncwarp();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================================================================
// v2: the same reduction with a TMA-fed operand pipeline (flat, 16 B aligned sources).  The register-staged
// producers above stream at 3.7 TB/s however many of them are in flight (a second producer group changed
// nothing); tensor copies into shared memory plus converter warps that only touch shared memory reached
// 5-6 TB/s in the forward projection kernel (gat_tap_tc.cu, k_tap_tc2), so the same structure is used here:
//   warp 17      copy issuer: per 16-node stage one box [16 x 128] fp32 per source (A, mask, <= 3 B segments)
//   warps 0-15   converters, two groups of eight building alternate operand stages; warp w converts row pieces
//                w and w + 8 of every source (lane = 4 features), A rows also masked, scaled and column-summed
//   warp 16      MMA issuer (one 16-node K step per stage), warps 0-3 drain TMEM at the end
constexpr int V2_SN = 16;
constexpr int V2_A_BYTES = V2_SN * MI * 2;              // 4 KB per hi or lo
constexpr int V2_B_BYTES = V2_SN * MAX_NB * 2;          // 12 KB per hi or lo
constexpr int V2_OP_BYTES = 2 * V2_A_BYTES + 2 * V2_B_BYTES;     // 32 KB
constexpr int V2_ATOM_STRIDE = V2_SN * 128;             // 2 KB between 64-feature atoms
constexpr int V2_BOX_BYTES = V2_SN * 128 * 4;           // 8 KB: one raw source box
constexpr int V2_RAW_SLOT = 5 * V2_BOX_BYTES;           // A, mask, B0, B1, B2
// An even ring: slot q % 4 is then always consumed by the same converter group (q % 2).  With three slots the groups
// alternated on a slot and a group waiting for fill m could be satisfied by the parity of fill m - 2 while fill m - 1
// (the other group's) was still in flight -- it read the wrong stage, released the slot early and the copy issuer
// re-armed a barrier whose phase was still open (illegal-instruction trap at B*N >= ~10^5 rows).
constexpr int V2_NRAW = 4;
constexpr int V2_NOP = 2;
constexpr int V2_GROUP = 256;
constexpr int V2_MMA_WARP = 16, V2_TMA_WARP = 17, V2_THREADS = 18 * 32;
constexpr int V2_RED_BYTES = 16 * MI * 4;
// the column-sum staging (V2_RED_BYTES) reuses raw slot 0 once every stage has been converted
constexpr size_t V2_SMEM_BYTES = (size_t)V2_NOP * V2_OP_BYTES + (size_t)V2_NRAW * V2_RAW_SLOT + 1024 + 256;
static_assert(V2_RED_BYTES <= V2_RAW_SLOT, "column-sum staging must fit a raw slot");
static_assert(V2_SMEM_BYTES <= 227 * 1024, "k_wgrad_tc2 shared memory");

struct Wgrad2Params {
  alignas(64) CUtensorMap tm_a;       // A source rows (dY or x)
  alignas(64) CUtensorMap tm_m;       // mask rows (y), same column window as A
  alignas(64) CUtensorMap tm_b0;      // B segment 0 (x or dR)
  alignas(64) CUtensorMap tm_u;       // B segments 1.. (taps buffer)
  long rows;
  int Z, NB;
  int a_zoff, b0_zoff, u_zoff, u_seg;      // column offsets: z * zoff (+ (seg - 1) * u_seg for the taps)
  int has_mask;                       // 0 none, 1 fp32 mask rows (tm_m), 2 bit mask (mask_bits)
  const uint32_t* mask_bits; int bits_C;      // word [m >> 5][c]: bit m & 31 = (y[m][c] > 0), written by the forward projection
  float a_scale;
  float* partial;
  float* colsum_partial;
};

__device__ __forceinline__ uint64_t make_mn_desc2(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(V2_ATOM_STRIDE >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(V2_THREADS, 1) k_wgrad_tc2(const __grid_constant__ Wgrad2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* raw = smem + (size_t)V2_NOP * V2_OP_BYTES;
  float* red = reinterpret_cast<float*>(raw);          // valid only after the last stage (see the column sums below)
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw + (size_t)V2_NRAW * V2_RAW_SLOT);
  uint64_t* op_full = bars;            // [2]
  uint64_t* op_empty = bars + 2;       // [2]
  uint64_t* raw_full = bars + 4;       // [4]
  uint64_t* raw_empty = bars + 8;      // [4]
  uint64_t* done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int z = blockIdx.x % p.Z;
  const int slot = blockIdx.x / p.Z, nslots = gridDim.x / p.Z;
  const long nstages = (p.rows + V2_SN - 1) / V2_SN;
  const int nseg = p.NB / 128;
  const int nsrc = 1 + nseg;                         // converted sources per stage (the mask is not one)

  if (threadIdx.x == 0) {
    for (int s = 0; s < V2_NOP; ++s) {
      tc::mbar_init(&op_full[s], V2_GROUP);
      tc::mbar_init(&op_empty[s], 1);
    }
    for (int s = 0; s < V2_NRAW; ++s) {
      tc::mbar_init(&raw_full[s], 1);
      tc::mbar_init(&raw_empty[s], V2_GROUP / 32);
    }
    tc::mbar_init(done, 1);
    tc::fence_barrier_init();
  }
  if (warp == V2_MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == V2_TMA_WARP) {
    // ===== copy issuer ====================================================================================
    if (tc::elect_one()) {
      const uint64_t tma_ = reinterpret_cast<uint64_t>(&p.tm_a), tmm = reinterpret_cast<uint64_t>(&p.tm_m);
      const uint64_t tmb = reinterpret_cast<uint64_t>(&p.tm_b0), tmu = reinterpret_cast<uint64_t>(&p.tm_u);
      const uint32_t bytes = (uint32_t)V2_BOX_BYTES * (uint32_t)(nsrc + (p.has_mask == 1 ? 1 : 0));
      unsigned q = 0;
      for (long sidx = slot; sidx < nstages; sidx += nslots, ++q) {
        const int r = (int)(q % V2_NRAW);
        tc::mbar_wait(&raw_empty[r], ((q / V2_NRAW) & 1u) ^ 1u);
        const uint32_t dst = tc::smem_u32(raw + (size_t)r * V2_RAW_SLOT);
        const int m0 = (int)(sidx * V2_SN);
        tc::mbar_arrive_expect_tx(&raw_full[r], bytes);
        tc::tensor_g2s_2d(dst, tma_, z * p.a_zoff, m0, &raw_full[r]);
        if (p.has_mask == 1) tc::tensor_g2s_2d(dst + V2_BOX_BYTES, tmm, z * p.a_zoff, m0, &raw_full[r]);
        tc::tensor_g2s_2d(dst + 2 * V2_BOX_BYTES, tmb, z * p.b0_zoff, m0, &raw_full[r]);
        for (int s = 1; s < nseg; ++s)
          tc::tensor_g2s_2d(dst + (2 + s) * V2_BOX_BYTES, tmu, z * p.u_zoff + (s - 1) * p.u_seg, m0, &raw_full[r]);
      }
    }
  } else if (warp < V2_MMA_WARP) {
    // ===== converters =====================================================================================
    const unsigned grp = warp >> 3;
    const int wg = warp & 7;
    // lane l owns features 4l .. 4l+3 of a 128-feature row: half (l & 1) of chunk (l >> 1) & 7 of atom l >> 4
    const uint32_t lane_off = (uint32_t)((lane >> 4) * V2_ATOM_STRIDE) + (uint32_t)((lane & 1) * 8);
    const int c8 = (lane >> 1) & 7;
    const uint32_t op_s = tc::smem_u32(smem + (size_t)grp * V2_OP_BYTES);
    const uint32_t raw_s = tc::smem_u32(raw) + (uint32_t)(lane * 16);
    float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
    unsigned q = 0;
    for (long sidx = slot; sidx < nstages; sidx += nslots, ++q) {
      if ((q & 1u) != grp) continue;
      const int r = (int)(q % V2_NRAW);
      // bit mask of this lane's four A columns for the 32-row group the stage lies in (issued before the wait)
      uint4 bw = make_uint4(0u, 0u, 0u, 0u);
      if (p.has_mask == 2)
        bw = __ldg(reinterpret_cast<const uint4*>(p.mask_bits + (size_t)(sidx >> 1) * p.bits_C + z * p.a_zoff + lane * 4));
      tc::mbar_wait(&raw_full[r], (q / V2_NRAW) & 1u);
      tc::mbar_wait(&op_empty[grp], ((q >> 1) & 1u) ^ 1u);
      const uint32_t src = raw_s + (uint32_t)(r * V2_RAW_SLOT);
      // rows wg and wg + 8 of every source; all loads first
      float4 v[4][2], mk[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const uint32_t ro = (uint32_t)((wg + 8 * i) * 512);
        v[0][i] = tc::ld_shared_v4(src + ro);
        if (p.has_mask == 2) {
          const uint32_t bit = 1u << ((int)(sidx & 1) * 16 + wg + 8 * i);
          mk[i] = make_float4((bw.x & bit) ? 1.f : 0.f, (bw.y & bit) ? 1.f : 0.f, (bw.z & bit) ? 1.f : 0.f,
                              (bw.w & bit) ? 1.f : 0.f);
        } else {
          mk[i] = p.has_mask ? tc::ld_shared_v4(src + V2_BOX_BYTES + ro) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
#pragma unroll
        for (int s = 0; s < 3; ++s)
          v[1 + s][i] = s < nseg ? tc::ld_shared_v4(src + (uint32_t)((2 + s) * V2_BOX_BYTES) + ro)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4& a = v[0][i];
        a.x = mk[i].x > 0.f ? a.x * p.a_scale : 0.f;
        a.y = mk[i].y > 0.f ? a.y * p.a_scale : 0.f;
        a.z = mk[i].z > 0.f ? a.z * p.a_scale : 0.f;
        a.w = mk[i].w > 0.f ? a.w * p.a_scale : 0.f;
        cs.x += a.x; cs.y += a.y; cs.z += a.z; cs.w += a.w;
      }
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (s > 0 && s - 1 >= nseg) continue;
        const uint32_t hi_base = op_s + (s == 0 ? 0u : (uint32_t)(2 * V2_A_BYTES + (s - 1) * 2 * V2_ATOM_STRIDE));
        const uint32_t lo_base = s == 0 ? op_s + (uint32_t)V2_A_BYTES : hi_base + (uint32_t)V2_B_BYTES;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int row = wg + 8 * i;
          uint32_t h0, l0, h1, l1;
          tc::split2(v[s][i].x, v[s][i].y, h0, l0);
          tc::split2(v[s][i].z, v[s][i].w, h1, l1);
          const uint32_t off = lane_off + (uint32_t)(row * 128 + ((c8 ^ (row & 7)) << 4));
          tc::st_shared_v2(hi_base + off, h0, h1);
          tc::st_shared_v2(lo_base + off, l0, l1);
        }
      }
      tc::fence_proxy_async();
      tc::mbar_arrive(&op_full[grp]);
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&raw_empty[r]);
    }
    // column sums (dbias): 16 warps x 128 features -> smem -> fixed-order sum
    if (p.colsum_partial != nullptr) {
      tc::named_bar_sync(2, 2 * V2_GROUP);       // every stage converted: the raw ring is free to hold the sums
      *reinterpret_cast<float4*>(red + warp * MI + lane * 4) = cs;
      tc::named_bar_sync(2, 2 * V2_GROUP);
      if (threadIdx.x < MI) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) s += red[r * MI + threadIdx.x];
        p.colsum_partial[((size_t)z * nslots + slot) * MI + threadIdx.x] = s;
      }
    }
    // ===== epilogue (warps 0-3) ===========================================================================
    if (warp < 4) {
      tc::mbar_wait(done, 0);
      tc::tc_fence_after();
      const int i_row = warp * 32 + lane;
      float* dst = p.partial + (((size_t)z * nslots + slot) * MI + i_row) * p.NB;
      const bool any = slot < nstages;
      for (int c0 = 0; c0 < p.NB; c0 += 32) {
        float v[32];
        if (any) {
          tc::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
          tc::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  } else {
    // ===== MMA issuer =====================================================================================
    const uint32_t idesc256 = make_idesc_mn(MI, 256), idesc128 = make_idesc_mn(MI, 128);
    unsigned q = 0;
    for (long sidx = slot; sidx < nstages; sidx += nslots, ++q) {
      const int stage = (int)(q & 1u);
      tc::mbar_wait(&op_full[stage], (q >> 1) & 1u);
      tc::tc_fence_after();
      if (tc::elect_one()) {
        const uint32_t sb = tc::smem_u32(smem + (size_t)stage * V2_OP_BYTES);
        const uint32_t a_hi = sb, a_lo = sb + V2_A_BYTES, b_hi = sb + 2 * V2_A_BYTES, b_lo = b_hi + V2_B_BYTES;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t aa = pass == 1 ? a_lo : a_hi;
          const uint32_t bb = pass == 2 ? b_lo : b_hi;
          const uint32_t accum = (q == 0 && pass == 0) ? 0u : 1u;
          if (p.NB >= 256) {
            tc::umma_bf16(tmem_base, make_mn_desc2(aa), make_mn_desc2(bb), idesc256, accum);
            if (p.NB > 256)
              tc::umma_bf16(tmem_base + 256, make_mn_desc2(aa), make_mn_desc2(bb + 4 * V2_ATOM_STRIDE), idesc128, accum);
          } else {
            tc::umma_bf16(tmem_base, make_mn_desc2(aa), make_mn_desc2(bb), idesc128, accum);
          }
        }
        tc::umma_commit(&op_empty[stage]);
      }
      __syncwarp();
    }
    if (tc::elect_one()) tc::umma_commit(done);
    __syncwarp();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == V2_MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// out[z][e] = sum_s partial[(z * nslots + s) * per + e]
__global__ void __launch_bounds__(256) k_wgrad_reduce(const float* __restrict__ partial, int nslots, long per,
                                                      long total, float* __restrict__ out) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const long zz = e / per, l = e - zz * per;
  const float* q = partial + (size_t)zz * nslots * per + l;
  float s = 0.f;
  for (int k = 0; k < nslots; ++k) s += q[(size_t)k * per];
  out[e] = s;
}

// dbias[f] = sum over heads z and CTA slots of the column-sum partials
__global__ void __launch_bounds__(128) k_dbias_from_colsums(const float* __restrict__ cp, int n, float* __restrict__ dbias) {
  const int f = threadIdx.x;
  float s = 0.f;
  for (int i = 0; i < n; ++i) s += cp[(size_t)i * MI + f];
  dbias[f] = s;
}

int wgrad_nslots(int Z) {
  const int s = device_sm_count() / Z;
  return s < 1 ? 1 : s;
}

// v2 when every source is a flat [rows][stride] matrix; fills the tensor maps
bool wgrad2_prepare(const WgradParams& wp, Wgrad2Params& q) {
  if (wp.rows >= (1l << 31) || wp.NB % 128 != 0 || wp.NB > MAX_NB) return false;
  auto flat = [&](const Src& s) { return s.sb == (long)wp.N * s.sn; };
  const int nseg = wp.NB / 128;
  if (!flat(wp.a) || !flat(wp.b[0])) return false;
  // widths: the last slice's 128-column window must lie inside the row
  const long a_w = (long)(wp.Z - 1) * wp.a.zoff + MI;
  const long b0_w = (long)(wp.Z - 1) * wp.b[0].zoff + 128;
  if (a_w > wp.a.sn || b0_w > wp.b[0].sn) return false;
  if (!tma::make_row_map(&q.tm_a, wp.a.p, wp.rows, a_w, wp.a.sn, V2_SN, 128)) return false;
  if (!tma::make_row_map(&q.tm_b0, wp.b[0].p, wp.rows, b0_w, wp.b[0].sn, V2_SN, 128)) return false;
  q.has_mask = wp.mask_bits != nullptr ? 2 : wp.mask_y != nullptr ? 1 : 0;
  if (q.has_mask == 2) {
    if (wp.bits_C % 4 != 0 || wp.a.zoff % 4 != 0 || ((uintptr_t)wp.mask_bits % 16) != 0) return false;
    q.mask_bits = wp.mask_bits; q.bits_C = wp.bits_C;
  }
  if (q.has_mask == 1) {
    if (wp.my_sb != (long)wp.N * wp.my_sn || wp.my_zoff != wp.a.zoff || a_w > wp.my_sn) return false;
    if (!tma::make_row_map(&q.tm_m, wp.mask_y, wp.rows, a_w, wp.my_sn, V2_SN, 128)) return false;
  }
  q.u_zoff = 0; q.u_seg = 0;
  if (nseg > 1) {
    // segments 1.. are consecutive 128-wide windows of one buffer (the taps rows)
    const Src& u = wp.b[1];
    if (!flat(u)) return false;
    for (int s = 2; s < nseg; ++s)
      if (wp.b[s].sn != u.sn || wp.b[s].sb != u.sb || wp.b[s].zoff != u.zoff || wp.b[s].p != u.p + (long)(s - 1) * 128)
        return false;
    const long u_w = (long)(wp.Z - 1) * u.zoff + (long)(nseg - 1) * 128;
    if (u_w > u.sn || !tma::make_row_map(&q.tm_u, u.p, wp.rows, u_w, u.sn, V2_SN, 128)) return false;
    q.u_zoff = (int)u.zoff; q.u_seg = 128;
  }
  q.rows = wp.rows; q.Z = wp.Z; q.NB = wp.NB;
  q.a_zoff = (int)wp.a.zoff; q.b0_zoff = (int)wp.b[0].zoff;
  q.a_scale = wp.a_scale; q.partial = wp.partial; q.colsum_partial = wp.colsum_partial;
  return true;
}

int launch_wgrad(WgradParams& wp, float* out, cudaStream_t st, const char* what) {
  const int nslots = wgrad_nslots(wp.Z);
  Wgrad2Params q{};
  if (wgrad2_prepare(wp, q)) {
    int rc0 = ensure_dyn_smem(KID_WGRAD2, (const void*)k_wgrad_tc2, V2_SMEM_BYTES, "k_wgrad_tc2");
    if (rc0) return rc0;
    k_wgrad_tc2<<<nslots * wp.Z, V2_THREADS, V2_SMEM_BYTES, st>>>(q);
  } else {
    int rc0 = ensure_dyn_smem(KID_WGRAD, (const void*)k_wgrad_tc, SMEM_BYTES, "k_wgrad_tc");
    if (rc0) return rc0;
    k_wgrad_tc<<<nslots * wp.Z, THREADS, SMEM_BYTES, st>>>(wp);
  }
  int rc = check_launch(what, st);
  if (rc) return rc;
  const long per = (long)MI * wp.NB, total = per * wp.Z;
  k_wgrad_reduce<<<cdiv(total, 256), 256, 0, st>>>(wp.partial, nslots, per, total, out);
  return check_launch("k_wgrad_reduce", st);
}

bool aligned16(const void* p) { return ((uintptr_t)p % 16) == 0; }

}  // namespace

// floats of `partial` scratch the tcgen05 weight-gradient kernels need
size_t wgrad_tc_partial_floats(int G, int F, int K, int P) {
  (void)G;
  return (size_t)wgrad_nslots(P) * P * MI * (size_t)(K * G > 128 ? K * G : 128) + (size_t)wgrad_nslots(P) * P * MI +
         (size_t)F * 0;
}

bool wgrad_tc_supported(const magat_gat_bwd_args* a) {
  if (a->F != MI || a->G != 128 || a->K > 3 || !a->concat) return false;
  if ((a->x_sn % 4) || (a->x_sb % 4) || !aligned16(a->x) || !aligned16(a->dy) || !aligned16(a->y)) return false;
  if (a->dy_sc != 1 || a->y_sc != 1 || (a->dy_sn % 4) || (a->dy_sb % 4) || (a->y_sn % 4) || (a->y_sb % 4)) return false;
  if (a->K > 1 && !aligned16(a->taps)) return false;
  if ((long)a->B * a->N >= (1l << 31)) return false;
  return true;
}

// dfilterWeight[p][f][k*G + g] = sum_m dP[m][p*F + f] * u_k^p[m][g]
int wgrad_tc_dfilter(const magat_gat_bwd_args* a, bool with_dbias, cudaStream_t st) {
  WgradParams wp{};
  wp.rows = (long)a->B * a->N;
  wp.N = a->N;
  wp.Z = a->P;
  wp.NB = a->K * a->G;
  wp.a = Src{a->dy, a->dy_sb, a->dy_sn, (long)a->F};
  wp.mask_y = a->relu ? a->y : nullptr;
  wp.my_sb = a->y_sb; wp.my_sn = a->y_sn; wp.my_zoff = a->F;
  if (a->relu && a->relu_bits) { wp.mask_bits = a->relu_bits; wp.bits_C = a->P * a->F; }
  wp.a_scale = 1.f;
  wp.b[0] = Src{a->x, a->x_sb, a->x_sn, 0};
  const long tap_row = (long)a->P * (a->K - 1) * a->G;
  for (int k = 1; k < a->K; ++k)
    wp.b[k] = Src{a->taps + (long)(k - 1) * a->G, (long)a->N * tap_row, tap_row, (long)(a->K - 1) * a->G};
  wp.partial = a->partial;
  const int nslots = wgrad_nslots(a->P);
  // concat mode: dbias[f] = sum_{m,p} dP[m][p*F+f] falls out of the A tiles this kernel already streams
  wp.colsum_partial = with_dbias ? a->partial + (size_t)nslots * a->P * MI * wp.NB : nullptr;
  int rc = launch_wgrad(wp, a->dfilterWeight, st, "k_wgrad_tc(dfilterWeight)");
  if (rc || !with_dbias) return rc;
  k_dbias_from_colsums<<<1, 128, 0, st>>>(wp.colsum_partial, nslots * a->P, a->dbias);
  return check_launch("k_dbias_from_colsums", st);
}

// KeyQuery: dweight[p][g][g'] = sum_m x[m][g] * dR[m][p][g']
int wgrad_tc_dweight(const magat_gat_bwd_args* a, cudaStream_t st) {
  WgradParams wp{};
  wp.rows = (long)a->B * a->N;
  wp.N = a->N;
  wp.Z = a->P;
  wp.NB = 128;
  wp.a = Src{a->x, a->x_sb, a->x_sn, 0};
  wp.mask_y = nullptr;
  wp.a_scale = 1.f;
  const long rc_row = (long)a->P * a->G;
  wp.b[0] = Src{a->rc, (long)a->N * rc_row, rc_row, (long)a->G};
  wp.partial = a->partial;
  wp.colsum_partial = nullptr;
  return launch_wgrad(wp, a->dweight, st, "k_wgrad_tc(dweight)");
}

}  // namespace magat
