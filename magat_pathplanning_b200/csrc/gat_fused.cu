// Single-launch fused forward of the batched graph-attention layer (graphML.py:4636-4667 with the functionals
// :1724-1827, :1180-1286, :713-823 behind it) for the shapes the planners train at scale: G = F = 128, K <= 3,
// P in {1,2,4}, heads concatenated.
//
// One cooperative, persistent launch.  A TEAM of 8 (or 16) CTAs owns one planning instance at a time; everything the
// instance needs between reading its dense GSO and writing y lives in a few MB of per-team scratch that is reused
// for the team's next instance, i.e. stays in L2 and never travels to HBM:
//
//   scan      S[b] (4N^2 bytes, the only large read) -> row / column bit masks; x -> bf16 hi/lo operand image
//   lists     bit masks -> padded neighbour lists + the slot of every edge in the list of its other end point
//   score     KeyQuery: R_p = X W_p on tcgen05 (W_p^T bf16 hi/lo image resident in shared memory, SS form)
//             GAT_modified: the two folded mixer dots per (node, head)
//   attention per-edge scores, row softmax over the out-neighbours (warp per sender, lane per slot)
//   taps      u_1 = A^T x, u_2 = A^T u_1 (warp per receiver), written as bf16 hi/lo operand images
//   project   Y_p = [x | u_1 | u_2] H_p^T + b, ReLU, concat: tcgen05 with H_p resident in TMEM (TS form),
//             operands arrive by TMA tensor copies straight in the SWIZZLE_128B layout -- no conversion pass
//
// CTA r of a team is (head p = r % P, node split h = r / P): in the two tensor-core phases it owns head p for the
// 64-node tiles t = h, h + NSPLIT, ...; in the sparse phases it owns the node range [r * CH, (r + 1) * CH) for all
// heads.  Phases are separated by a team barrier (one global counter per team, release / acquire); teams never
// synchronise with each other, so while one team streams its GSO another one is on the tensor cores and a third
// gathers from L2.  bf16 hi/lo split operands, three MMAs per k step (hi.hi + lo.hi + hi.lo): ~5e-6 max-norm
// relative against the fp32 reference (bar 1e-4).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "gat_sparse.cuh"
#include "tc_common.cuh"
#include "tma_host.cuh"

namespace magat {

namespace {

constexpr int FT = 128;                      // features per head: G == F == UMMA M == TMEM lanes
constexpr int TN = 64;                       // nodes per MMA tile == UMMA N
constexpr int SK = 128;                      // K elements per pipeline stage
constexpr int ATOM_B = TN * 128;             // 8 KB: [64 nodes x 64 k] bf16, SWIZZLE_128B
constexpr int STAGE_BYTES = 4 * ATOM_B;      // hi atom 0/1, lo atom 0/1
constexpr int NST = 4;                       // operand stages
constexpr int WATOM = FT * 128;              // 16 KB: [128 features x 64 k] bf16, SWIZZLE_128B
constexpr int WIMG_BYTES = 4 * WATOM;        // W_p^T image: hi atom 0/1, lo atom 0/1
constexpr int ACC_COL0 = 384;                // TMEM: filter taps in columns [0, K*G), accumulators at 384 + 64 a
constexpr int NTHREADS = 768, NWARPS = NTHREADS / 32;
constexpr int TMA_WARP = 0, MMA_WARP = 1, EPI_WARP0 = 4;       // warps 4..7: TMEM lane quarter = warp % 4
constexpr int TEAM_DEFAULT = 8;              // CTAs per planning instance (a multiple of P; 16 halves the instances in flight)
constexpr size_t SMEM_BYTES = (size_t)NST * STAGE_BYTES + WIMG_BYTES + 1024 + 256;
constexpr long long WATCHDOG_CYCLES = 6000000000ll;            // ~3 s: a lost arrival traps instead of hanging the GPU

struct FusedParams {
  alignas(64) CUtensorMap tm_x;              // x image   [teams * 2][N][256] bf16, box 64 x 64
  alignas(64) CUtensorMap tm_u;              // tap image [teams][N][P * (K-1) * 256] bf16, box 64 x 64
  int B, N, K, P, D, W, WS;                  // W = ceil(N / 32) mask words per row, WS = W rounded up to 4
  int mode, relu, save, s_f64;
  int nteams, team_size, nsplit, chunk, tiles;
  const void* S;
  const float* x; long x_sb, x_sn;
  const float* weight; const float* mixer; const float* wb; const float* H; const float* bias;
  float* y; long y_sb, y_sn;
  int32_t* nbr_out; int32_t* nbr_in; int32_t* slot_in;                         // [B][N][D]
  float* att;                                // [B][N][D][P]
  float* taps; long taps_inst, taps_team;    // fp32 taps for backward (save) -- [N][P][K-1][G] per instance
  float* sproj; long sproj_inst, sproj_team; // KeyQuery R [N][P][G] / GAT_modified [N][P][2]
  float* wprep_out;                          // GAT_modified, save: folded mixer vectors [2P][G] + [2P] for backward
  uint32_t* rowbits; uint32_t* colbits;      // [team][2][N][WS] (double buffered by instance parity)
  uint16_t* ximg; uint16_t* uimg;            // images as above
  unsigned* bar;                             // [team][32] barrier counters (128 B apart)
  int32_t* status;                           // [0] max out-degree [1] max in-degree [2] rows over the cap D [3] watchdog
  long long* prof;                           // [CTA][16] cycles per phase (work / barrier wait), see PROF_* below
};

// phase clock (thread 0 of every CTA): cycles since the previous mark are added to slot i
#define PROF_MARK(i)                         \
  do {                                       \
    if (threadIdx.x == 0) {                  \
      const long long t_ = clock64();        \
      prof[i] += t_ - prof_t;                \
      prof_t = t_;                           \
    }                                        \
  } while (0)

// ---- small device helpers --------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic-proxy writes (global or shared) <-> async-proxy accesses (TMA, UMMA)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct ScanJob {                              // one instance's GSO on its way into one parity of the mask buffers
  const void* S; uint32_t* rowbits; uint32_t* colbits;
};

__device__ __forceinline__ void tensor_g2s_3d(uint32_t dst_smem, uint64_t tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          dst_smem),
      "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(tc::smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_wait_wd(uint64_t* bar, uint32_t parity, int32_t* status, int code) {
  if (tc::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  unsigned n = 0;
  while (!tc::mbar_try_wait(bar, parity)) {
    if ((++n & 255u) == 0 && clock64() - t0 > WATCHDOG_CYCLES) {
      status[3] = code;
      __threadfence_system();
      __trap();
    }
  }
}

__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, int32_t* status, int code, long long& acc) {
  const long long t0 = clock64();
  mbar_wait_wd(bar, parity, status, code);
  acc += clock64() - t0;
}

// Data produced inside this launch by another CTA: L2 is the point of coherence, so bypass L1 (ld.global.cg).

// Scratch that has been consumed is dead, but L2 does not know: left alone, its dirty lines are written back to HBM when
// the next instances push them out (4.5 GB per forward at the default workload -- more than the layer's output).
// discard.global.L2 drops the lines without a write-back.  All threads of the CTA; base and size multiples of 128 B.
__device__ __forceinline__ void discard_lines(const void* base, size_t bytes) {
  const char* q = reinterpret_cast<const char*>(base);
  for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)NTHREADS * 128)
    asm volatile("discard.global.L2 [%0], 128;" ::"l"(q + off) : "memory");
}

// one image row (256 bf16: hi[128] | lo[128]); lane l owns features 4l .. 4l+3
__device__ __forceinline__ void image_store(uint16_t* row, int lane, const float4& v) {
  uint2 hi, lo;
  tc::split2(v.x, v.y, hi.x, lo.x);
  tc::split2(v.z, v.w, hi.y, lo.y);
  *reinterpret_cast<uint2*>(row + lane * 4) = hi;
  *reinterpret_cast<uint2*>(row + 128 + lane * 4) = lo;
}
// Team barrier: every thread publishes its writes (also towards the TMA engine of the other CTAs), one thread per CTA
// arrives on the team's counter and waits until all TEAM CTAs have.
__device__ __forceinline__ void team_barrier(unsigned* cnt, unsigned& target, int team_size, int32_t* status,
                                             long long* prof, long long& prof_t, int slot) {
  fence_proxy_async_all();
  __syncthreads();
  if (threadIdx.x == 0) {
    PROF_MARK(slot);                          // work of the phase that ends here
    target += (unsigned)team_size;
    __threadfence();
    red_release_add(cnt, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(cnt) < target) {
      __nanosleep(40);
      if (clock64() - t0 > WATCHDOG_CYCLES) {
        status[3] = 100;
        __threadfence_system();
        __trap();
      }
    }
    __threadfence();
    fence_proxy_async_all();
    PROF_MARK(slot + 1);                      // waiting for the slowest CTA of the team
  }
  __syncthreads();
}

// ---- phase: GSO scan (graphML.py:1274-1276: only |S| > 1e-9 matters; NaN is "no edge") --------------------
// One warp per (32-row band, 128-column segment): 512 B row pieces, 8 rows in flight per lane.  Most row pieces of a
// sparse GSO hold no edge at all: one vote per row piece decides that, and only pieces with an edge pay for the three
// xor-shuffles that assemble the row words and for the column-word updates (bit r of a column word = row r of the
// band).
template <typename T> struct Nib4;
template <> struct Nib4<float> {
  static __device__ __forceinline__ uint32_t edges(const float* p) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
    return (uint32_t)(fabsf(v.x) > 1e-9f) | ((uint32_t)(fabsf(v.y) > 1e-9f) << 1) |
           ((uint32_t)(fabsf(v.z) > 1e-9f) << 2) | ((uint32_t)(fabsf(v.w) > 1e-9f) << 3);
  }
};
template <> struct Nib4<double> {
  static __device__ __forceinline__ uint32_t edges(const double* p) {
    const double2 a = __ldcs(reinterpret_cast<const double2*>(p));
    const double2 b = __ldcs(reinterpret_cast<const double2*>(p) + 1);
    return (uint32_t)(fabs(a.x) > 1e-9) | ((uint32_t)(fabs(a.y) > 1e-9) << 1) | ((uint32_t)(fabs(b.x) > 1e-9) << 2) |
           ((uint32_t)(fabs(b.y) > 1e-9) << 3);
  }
};

// The units of an instance are dealt out to the team's CTAs round robin (unit u belongs to CTA u % TEAM); inside the
// CTA the warps draw from a shared counter, so whoever has nothing else to do -- every warp in the scan phase, the
// warps without a tensor-core role during the two projection phases of the PREVIOUS instance -- takes the next one.
template <typename T>
__device__ __forceinline__ void scan_units(const FusedParams& p, const T* Sb, int r, int lane, uint32_t* rowbits,
                                           uint32_t* colbits, int* counter, const volatile int* stop, int stop_at) {
  const int N = p.N, W = p.W, WS = p.WS;
  const int segs = (N + 127) >> 7;
  const int units = W * segs;
  for (;;) {
    int kq = 0;
    if (lane == 0) kq = (stop != nullptr && *stop >= stop_at) ? -1 : atomicAdd(counter, 1);   // background pass: until the tensor-core roles are through
    kq = __shfl_sync(0xffffffffu, kq, 0);
    const int u = r + p.team_size * kq;
    if (kq < 0 || kq >= units || u >= units) break;
    const int seg = u % segs, band = u / segs;
    const int j0 = seg * 128 + lane * 4;
    const bool jin = j0 < N;                     // N % 4 == 0: a lane is entirely inside or outside
    const T* Sc = Sb + j0;
    uint32_t col0 = 0, col1 = 0, col2 = 0, col3 = 0;
    const int i0 = band * 32;
    const int w = seg * 4 + (lane >> 3);
    uint32_t* rb = rowbits + (size_t)i0 * WS + w;
    const bool rw = (lane & 7) == 0 && w < WS;
#pragma unroll
    for (int rr = 0; rr < 32; rr += 8) {
      uint32_t nib[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        nib[q] = (jin && i0 + rr + q < N) ? Nib4<T>::edges(Sc + (size_t)(i0 + rr + q) * N) : 0u;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int rw_i = rr + q;
        uint32_t v = 0u;
        if (__any_sync(0xffffffffu, nib[q] != 0u)) {
          const uint32_t nb = nib[q];
          col0 |= (nb & 1u) << rw_i;
          col1 |= ((nb >> 1) & 1u) << rw_i;
          col2 |= ((nb >> 2) & 1u) << rw_i;
          col3 |= ((nb >> 3) & 1u) << rw_i;
          v = nb << (4 * (lane & 7));
          v |= __shfl_xor_sync(0xffffffffu, v, 1);
          v |= __shfl_xor_sync(0xffffffffu, v, 2);
          v |= __shfl_xor_sync(0xffffffffu, v, 4);
        }
        if (rw && i0 + rw_i < N) rb[(size_t)rw_i * WS] = v;
      }
    }
    if (jin) {
      uint32_t* cb = colbits + (size_t)j0 * WS + band;
      cb[0] = col0;
      cb[WS] = col1;
      cb[2 * (size_t)WS] = col2;
      cb[3 * (size_t)WS] = col3;
    }
  }
}

__device__ __forceinline__ void scan_job(const FusedParams& p, const ScanJob& jb, int r, int lane, int* counter,
                                         const volatile int* stop = nullptr, int stop_at = 0) {
  if (p.s_f64)
    scan_units<double>(p, reinterpret_cast<const double*>(jb.S), r, lane, jb.rowbits, jb.colbits, counter, stop, stop_at);
  else
    scan_units<float>(p, reinterpret_cast<const float*>(jb.S), r, lane, jb.rowbits, jb.colbits, counter, stop, stop_at);
}

// ---- phase: neighbour lists of this CTA's nodes ---------------------------------------------------------------
// One task per (node, direction), a quarter warp (8 lanes) each -- lane l of the group owns mask words 4l .. 4l+3 of a
// 32-word pass:
//   direction 0: out-list of n from its row bits
//   direction 1: in-list of n from its column bits; slot_in = position of n in the out-list of each sender (where the
//                sender's softmax leaves A[i, n] in att[i][:])
// The position of n in the list of m is the number of set bits below n in the other mask's row m (one 16 B load per
// lane and three shuffles per edge).  scratch: 2 x 32 ints per group in shared memory.
__device__ __forceinline__ int rank_partial(const uint32_t* __restrict__ row, int wi, int sw, uint32_t below) {
  int rk = 0;
  if (wi <= sw) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(row + wi));
    const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int idx = wi + c;
      if (idx < sw) rk += __popc(ws[c]);
      else if (idx == sw) rk += __popc(ws[c] & below);
    }
  }
  return rk;
}

__device__ __forceinline__ void phase_lists(const FusedParams& p, long rowbase, int n0, int n1,
                                            const uint32_t* rowbits, const uint32_t* colbits, int32_t* scratch,
                                            int warp, int lane) {
  const int cn = n1 - n0, D = p.D, W = p.W, WS = p.WS;
  const int grp = lane >> 3, gl = lane & 7;
  const unsigned gmask = 0xffu << (grp * 8);
  int32_t* my = scratch + (size_t)(warp * 4 + grp) * 64;       // [0,32) list, [32,64) slots
  int mo = 0, mi = 0, over = 0;
  const int ntask = 2 * cn;
  for (int idx0 = 0; idx0 < ntask; idx0 += NWARPS * 4) {       // warp-uniform trip count
    const int idx = idx0 + warp * 4 + grp;
    const bool live = idx < ntask;
    const int dir = (live && idx >= cn) ? 1 : 0;
    const int n = live ? n0 + (dir ? idx - cn : idx) : n0;
    const uint32_t* mine = (dir ? colbits : rowbits) + (size_t)n * WS;
    const uint32_t* other = dir ? rowbits : colbits;
    *reinterpret_cast<int4*>(my + gl * 4) = make_int4(-1, -1, -1, -1);
    *reinterpret_cast<int4*>(my + 32 + gl * 4) = make_int4(0, 0, 0, 0);
    __syncwarp();
    int deg = 0;
    for (int w0 = 0; w0 < W; w0 += 32) {
      const int wi = w0 + 4 * gl;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (live && wi < W) v = __ldcg(reinterpret_cast<const uint4*>(mine + wi));
      uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (wi + c >= W) ws[c] = 0u;
      const int cnt = __popc(ws[0]) + __popc(ws[1]) + __popc(ws[2]) + __popc(ws[3]);
      int incl = cnt;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o, 8);
        if (gl >= o) incl += t;
      }
      int pos = deg + incl - cnt;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t word = ws[c];
        while (word) {
          const int bit = __ffs(word) - 1;
          word &= word - 1;
          if (pos < 32) my[pos] = (wi + c) * 32 + bit;
          ++pos;
        }
      }
      deg += __shfl_sync(0xffffffffu, incl, 7, 8);
    }
    __syncwarp();
    const int dcap = dir ? min(deg, D) : 0;      // only in-edges carry a slot (position in the sender's out-list)
    const int sw = n >> 5;
    const uint32_t below = (1u << (n & 31)) - 1u;
    if (W <= 32) {
      for (int s = 0; s < dcap; s += 4) {                       // four edges in flight per group
        int rk[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int m = (s + u < dcap) ? my[s + u] : -1;
          rk[u] = m >= 0 ? rank_partial(other + (size_t)m * WS, 4 * gl, sw, below) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          rk[u] += __shfl_xor_sync(gmask, rk[u], 1);
          rk[u] += __shfl_xor_sync(gmask, rk[u], 2);
          rk[u] += __shfl_xor_sync(gmask, rk[u], 4);
          if (gl == u && s + u < dcap) my[32 + s + u] = rk[u];
        }
      }
    } else {
      for (int s = 0; s < dcap; ++s) {
        const int m = my[s];
        int rk = 0;
        for (int w0 = 0; w0 <= sw; w0 += 32) rk += rank_partial(other + (size_t)m * WS, w0 + 4 * gl, sw, below);
        rk += __shfl_xor_sync(gmask, rk, 1);
        rk += __shfl_xor_sync(gmask, rk, 2);
        rk += __shfl_xor_sync(gmask, rk, 4);
        if (gl == 0) my[32 + s] = rk;
      }
    }
    __syncwarp(gmask);
    if (live) {
      int32_t* lst = (dir ? p.nbr_in : p.nbr_out) + (rowbase + n) * D;
      if (gl * 4 < D) {
        *reinterpret_cast<int4*>(lst + gl * 4) = *reinterpret_cast<const int4*>(my + gl * 4);
        if (dir)
          *reinterpret_cast<int4*>(p.slot_in + (rowbase + n) * D + gl * 4) = *reinterpret_cast<const int4*>(my + 32 + gl * 4);
      }
      if (dir) mi = max(mi, deg); else mo = max(mo, deg);
      if (deg > D) over = 1;
    }
    __syncwarp();
  }
  mo = warp_max_i(mo);
  mi = warp_max_i(mi);
  over = __any_sync(0xffffffffu, over);
  if (lane == 0) {
    if (mo > 0) atomicMax(&p.status[0], mo);
    if (mi > 0) atomicMax(&p.status[1], mi);
    if (over) atomicAdd(&p.status[2], 1);
  }
}

// ---- phase: KeyQuery scores + row softmax (graphML.py:1246-1286): warp per sender row (gat_sparse.cuh) ----------
template <int PT>
__device__ __forceinline__ void phase_attention_kq(const FusedParams& p, long rowbase, const float* xb, int n0, int n1,
                                                   int warp, int lane, const float* sproj, float* esc) {
  const int D = p.D;
  float* e_s = esc + warp * 32 * PT;
  const int32_t* nbo = p.nbr_out + rowbase * D;
  float* att_b = p.att + (size_t)rowbase * D * PT;
  int i = n0 + warp;
  if (i >= n1) return;
  int my_j = sparse::list_entry<true>(nbo + (unsigned)(i * D), D, lane, -1);
  for (; i < n1; i += NWARPS) {                 // the next row's list entry is loaded one row ahead
    const int nxt = i + NWARPS;
    const int nxt_j = nxt < n1 ? sparse::list_entry<true>(nbo + (unsigned)(nxt * D), D, lane, -1) : -1;
    sparse::attention_kq_row<PT, true>(xb, (unsigned)p.x_sn, sproj + (unsigned)(i * PT * FT), my_j,
                                       att_b + (unsigned)(i * D * PT), D, lane, e_s);
    my_j = nxt_j;
  }
}

// ---- GAT_modified (graphML.py:713-823) ---------------------------------------------------------------------
// a_t.z_n = (W^T a_t).x_n + a_t.wb: two G-vectors and two scalars per head replace the F x N projection.
// cvec[p][t][g] = sum_f mixer[p][t*F+f] W[p][f][g], dvec[p][t] = sum_f mixer[p][t*F+f] wb[p][f]; t = 0 is a1 (receiver
// term), t = 1 is a2 (sender term).  Every CTA builds them in shared memory once.
__device__ __forceinline__ void gm_prep(const FusedParams& p, float* cd /* smem [2P][G] + [2P] */) {
  const int P = p.P;
  for (int o = threadIdx.x; o < 2 * P * FT; o += NTHREADS) {
    const int pt = o / FT, g = o - pt * FT;
    const int h = pt >> 1, t = pt & 1;
    const float* a = p.mixer + (size_t)h * 2 * FT + (size_t)t * FT;
    float s = 0.f;
    for (int f = 0; f < FT; ++f) s = fmaf(__ldg(a + f), __ldg(p.weight + ((size_t)h * FT + f) * FT + g), s);
    cd[o] = s;
  }
  for (int pt = threadIdx.x; pt < 2 * P; pt += NTHREADS) {
    const int h = pt >> 1, t = pt & 1;
    const float* a = p.mixer + (size_t)h * 2 * FT + (size_t)t * FT;
    float s = 0.f;
    for (int f = 0; f < FT; ++f) s = fmaf(__ldg(a + f), __ldg(p.wb + (size_t)h * FT + f), s);
    cd[2 * P * FT + pt] = s;
  }
}

// sproj[n][2h + t] = cvec[h][t] . x_n + dvec[h][t] for this CTA's nodes (warp per node)
template <int PT>
__device__ __forceinline__ void phase_mixer_gm(const FusedParams& p, const float* xb, int n0, int n1, int warp, int lane,
                                               const float* cd, float* sproj) {
  float4 c[2 * PT];
#pragma unroll
  for (int q = 0; q < 2 * PT; ++q) c[q] = *reinterpret_cast<const float4*>(cd + q * FT + lane * 4);
  for (int n = n0 + warp; n < n1; n += NWARPS) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(xb + (long)n * p.x_sn + lane * 4));
#pragma unroll
    for (int q = 0; q < 2 * PT; ++q) {
      const float d = warp_sum(sparse::dot4(c[q], xv));
      if (lane == q) sproj[(size_t)n * 2 * PT + q] = d + cd[2 * PT * FT + q];
    }
  }
}

template <int PT>
__device__ __forceinline__ void phase_attention_gm(const FusedParams& p, long rowbase, int n0, int n1, int warp, int lane,
                                                   const float* sproj, float* esc) {
  const int D = p.D;
  float* e_s = esc + warp * 32 * PT;
  float* att_b = p.att + (size_t)rowbase * D * PT;
  for (int i = n0 + warp; i < n1; i += NWARPS) {
    const long row = rowbase + i;
    const int my_j = lane < D ? __ldcg(p.nbr_out + row * D + lane) : -1;
    const int deg = __popc(__ballot_sync(0xffffffffu, my_j >= 0));
    if (my_j >= 0) {                         // lane = slot: e = LeakyReLU(a2.z_i + a1.z_j), graphML.py:785-796
#pragma unroll
      for (int h = 0; h < PT; ++h) {
        float e = __ldcg(sproj + ((size_t)i * PT + h) * 2 + 1) + __ldcg(sproj + ((size_t)my_j * PT + h) * 2 + 0);
        e = e > 0.f ? e : kLeaky * e;
        e_s[lane * PT + h] = e;
      }
    }
    __syncwarp();
    sparse::softmax_store<PT>(att_b + (unsigned)(i * D * PT), D, lane, deg, e_s);
    __syncwarp();
  }
}

// ---- phase: one level of the tap recursion, u_k[j] = sum_{i in in(j)} A_p[i,j] u_{k-1}[i] (graphML.py:1756-1759) ----
// Warp per receiver, all heads at once; lane l owns features 4l .. 4l+3.  Lane s keeps in-edge s (sender id and the
// PT attention values, one 16 B load); the edge loop broadcasts them by shuffle, two edges in flight.  k = 1 gathers
// rows of x (one row feeds every head), k = 2 the heads' fp32 u_1 rows.  Output: bf16 hi/lo image row for the
// projection (+ the fp32 copy the next level and backward read).
template <int PT>
__device__ __forceinline__ void phase_gather(const FusedParams& p, long rowbase, const float* xb, int n0, int n1,
                                             int warp, int lane, int k, uint16_t* uimg, float* taps) {
  const int D = p.D, Km1 = p.K - 1;
  const int32_t* nbi = p.nbr_in + rowbase * D;
  const int32_t* sli = p.slot_in + rowbase * D;
  const float* att_b = p.att + (size_t)rowbase * D * PT;
  const unsigned trow = (unsigned)(Km1 * FT);                  // floats between the heads of a node in the taps buffer
  const float* tsrc = taps + (k >= 2 ? (k - 2) * FT : 0);       // plane k-1 of every (node, head)
  const bool keep32 = p.save || k < Km1;                       // fp32 copy: for backward, and as the source of the next level
  int j = n0 + warp;
  if (j >= n1) return;
  // the next row's list entries are loaded one row ahead
  int my_i = sparse::list_entry<true>(nbi + (unsigned)(j * D), D, lane, -1);
  int my_sl = sparse::list_entry<true>(sli + (unsigned)(j * D), D, lane, 0);
  for (; j < n1; j += NWARPS) {
    const int nxt = j + NWARPS;
    const int nxt_i = nxt < n1 ? sparse::list_entry<true>(nbi + (unsigned)(nxt * D), D, lane, -1) : -1;
    const int nxt_sl = nxt < n1 ? sparse::list_entry<true>(sli + (unsigned)(nxt * D), D, lane, 0) : 0;
    float am[PT];
    sparse::edge_weights<PT, true>(att_b, my_i, my_sl, D, am);
    float4 acc[PT];
    if (k == 1) sparse::gather_row<PT, true, true>(xb, (unsigned)p.x_sn, tsrc, trow, my_i, am, lane, acc);
    else sparse::gather_row<PT, false, true>(xb, (unsigned)p.x_sn, tsrc, trow, my_i, am, lane, acc);
    uint16_t* irow = uimg + (unsigned)((j * PT * Km1 + (k - 1)) * 256);
    float* trow_out = taps + (unsigned)(j * PT) * trow + (k - 1) * FT + lane * 4;
#pragma unroll
    for (int h = 0; h < PT; ++h) {
      image_store(irow + h * Km1 * 256, lane, acc[h]);
      if (keep32) *reinterpret_cast<float4*>(trow_out + h * trow) = acc[h];
    }
    my_i = nxt_i;
    my_sl = nxt_sl;
  }
}

// ---- epilogue stores: lane f of the warp holds feature f of 32 consecutive nodes, one 128 B line per node -----------
// STRIDE = floats between consecutive node rows, a compile-time constant so every store carries its offset as an
// immediate (one FADD, one FMNMX and one STG per node instead of a 64-bit multiply-add chain); 0 = runtime stride.
template <int STRIDE, bool ACT, bool STREAM>
__device__ __forceinline__ void epi_store(float* dst, long stride_rt, const float (&v)[32], int left, float bias,
                                          bool relu) {
  if (ACT && !relu) {                          // bias only (a caller-supplied nonlinearity follows in torch)
#pragma unroll
    for (int n = 0; n < 32; ++n) {
      float* q = STRIDE ? dst + n * STRIDE : dst + n * stride_rt;
      if (n < left) __stcs(q, v[n] + bias);
    }
  } else if (left >= 32) {
#pragma unroll
    for (int n = 0; n < 32; ++n) {
      const float o = ACT ? fmaxf(v[n] + bias, 0.f) : v[n];
      float* q = STRIDE ? dst + n * STRIDE : dst + n * stride_rt;
      if (STREAM) __stcs(q, o); else *q = o;
    }
  } else {
#pragma unroll
    for (int n = 0; n < 32; ++n) {
      const float o = ACT ? fmaxf(v[n] + bias, 0.f) : v[n];
      float* q = STRIDE ? dst + n * STRIDE : dst + n * stride_rt;
      if (n < left) { if (STREAM) __stcs(q, o); else *q = o; }
    }
  }
}

// ---- the kernel -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1) k_gat_fused(const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* wimg = smem + (size_t)NST * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wimg + WIMG_BYTES);
  uint64_t* full = bars;                     // [NST] expect_tx + tensor copies
  uint64_t* empty = bars + NST;              // [NST] tcgen05.commit
  uint64_t* acc_full = bars + 2 * NST;       // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int* scan_ctr = reinterpret_cast<int*>(tmem_slot + 1);     // next scan unit of this CTA (of the instance being scanned)
  volatile int* gemm_done = reinterpret_cast<volatile int*>(tmem_slot + 2);   // tensor-core phases this CTA has finished
  int gphase = 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int TEAM = p.team_size;
  const int team = blockIdx.x / TEAM, r = blockIdx.x % TEAM;
  const int P = p.P, K = p.K, N = p.N;
  const int head = r % P, split = r / P;
  const int KG = K * FT;
  const bool kq = p.mode == MAGAT_MODE_KEYQUERY;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      tc::mbar_init(&acc_full[a], 1);
      tc::mbar_init(&acc_empty[a], 4 * 32);
    }
    *scan_ctr = 0;
    *gemm_done = 0;
    tc::fence_barrier_init();
  }
  if (warp == MMA_WARP) tc::tmem_alloc<512>(tmem_slot);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- resident weights: filter taps of this head -> TMEM (hi at columns [0, KG/2), lo at [KG/2, KG)) ----
  if (warp < 4) {
    const int f = warp * 32 + lane;
    const float* hrow = p.H + ((size_t)head * FT + f) * KG;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < KG; k0 += 32) {
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(hrow + k0 + 2 * j));
        tc::split2(v.x, v.y, hi[j], lo[j]);
        tc::split2(v.z, v.w, hi[j + 1], lo[j + 1]);
      }
      tc::tmem_st16(lane_addr + (uint32_t)(k0 / 2), hi);
      tc::tmem_st16(lane_addr + (uint32_t)(KG / 2 + k0 / 2), lo);
    }
    tc::tmem_st_wait();
  }
  // ---- KeyQuery: W_p^T as a K-major bf16 hi/lo image in shared memory (A operand of the score MMAs):
  //      row m = g' (output feature of R), k = g; chunk task = (g', 8 consecutive g)
  float* gm_cd = reinterpret_cast<float*>(wimg);            // GAT_modified reuses the region for cvec / dvec
  if (kq) {
    const float* Wp = p.weight + (size_t)head * FT * FT;
    for (int c = threadIdx.x; c < FT * 16; c += NTHREADS) {
      const int gp = c & (FT - 1), kc = c >> 7;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(Wp + (size_t)(kc * 8 + i) * FT + gp);
      uint4 hi, lo;
      tc::split2(v[0], v[1], hi.x, lo.x);
      tc::split2(v[2], v[3], hi.y, lo.y);
      tc::split2(v[4], v[5], hi.z, lo.z);
      tc::split2(v[6], v[7], hi.w, lo.w);
      const uint32_t off = (uint32_t)((kc >> 3) * WATOM) + tc::sw128_offset(gp, kc & 7);
      *reinterpret_cast<uint4*>(wimg + off) = hi;
      *reinterpret_cast<uint4*>(wimg + 2 * WATOM + off) = lo;
    }
  } else {
    gm_prep(p, gm_cd);
    __syncthreads();
    if (blockIdx.x == 0 && p.wprep_out != nullptr)
      for (int o = threadIdx.x; o < 2 * P * FT + 2 * P; o += NTHREADS) p.wprep_out[o] = gm_cd[o];
  }
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();

  float* esc = reinterpret_cast<float*>(smem);       // sparse phases: per-warp score scratch in the (idle) stage ring
  unsigned* bar = p.bar + (size_t)team * 32;
  unsigned bar_target = 0;
  long long* prof = p.prof + (size_t)blockIdx.x * 16;
  long long prof_t = clock64();
  long long wait_a = 0, wait_b = 0;          // role-private mbarrier wait clocks (slots 12..15 at exit)
  uint32_t q = 0, oc = 0;                    // running stage / accumulator counters of the tensor-core roles
  const int n0 = min(N, r * p.chunk), n1 = min(N, n0 + p.chunk);
  const size_t esz = p.s_f64 ? 8 : 4;
  const bool bg_warp = warp != TMA_WARP && warp != MMA_WARP && !(warp >= EPI_WARP0 && warp < EPI_WARP0 + 4);
  uint16_t* uimg = p.uimg + (size_t)team * N * P * (K > 1 ? K - 1 : 1) * 256;
  const uint64_t tmx = reinterpret_cast<uint64_t>(&p.tm_x);
  const uint64_t tmu = reinterpret_cast<uint64_t>(&p.tm_u);
  constexpr uint32_t idesc = tc::make_idesc_bf16(FT, TN);

  int it = 0;
  for (int b = team; b < p.B; b += p.nteams, ++it) {
    const int par = it & 1;
    const long rowbase = (long)b * N;
    const float* xb = p.x + (long)b * p.x_sb;
    uint16_t* ximg = p.ximg + ((size_t)team * 2 + par) * N * 256;
    float* sproj = p.sproj + (long)b * p.sproj_inst + (long)team * p.sproj_team;
    float* taps = p.taps ? p.taps + (long)b * p.taps_inst + (long)team * p.taps_team : nullptr;

    // ================= scan: x image of my nodes, then my share of the GSO ==============================
    for (int n = n0 + warp; n < n1; n += NWARPS)
      image_store(ximg + (size_t)n * 256, lane,
                  __ldg(reinterpret_cast<const float4*>(xb + (long)n * p.x_sn + lane * 4)));
    // mask buffers are double buffered by instance parity: the next instance is scanned in the background while this
    // one is still being read by the list builders of slower CTAs
    uint32_t* rowbits = p.rowbits + ((size_t)team * 2 + par) * N * p.WS;
    uint32_t* colbits = p.colbits + ((size_t)team * 2 + par) * N * p.WS;
    const ScanJob cur{reinterpret_cast<const uint8_t*>(p.S) + (size_t)b * N * N * esz, rowbits, colbits};
    const int bn = b + p.nteams;
    const ScanJob nxt{reinterpret_cast<const uint8_t*>(p.S) + (size_t)(bn < p.B ? bn : b) * N * N * esz,
                      p.rowbits + ((size_t)team * 2 + (par ^ 1)) * N * p.WS,
                      p.colbits + ((size_t)team * 2 + (par ^ 1)) * N * p.WS};
    const bool bg = bn < p.B && bg_warp;
    scan_job(p, cur, r, lane, scan_ctr);                       // whatever the background passes left over (all of it for it = 0)
    team_barrier(bar, bar_target, TEAM, p.status, prof, prof_t, 0);
    if (threadIdx.x == 0) *scan_ctr = 0;                        // (published by the __syncthreads below)
    if (it > 0 && n1 > n0) {
      // every CTA of the team is past the previous instance's projection: its operand images are dead
      discard_lines(p.ximg + (((size_t)team * 2 + (par ^ 1)) * N + n0) * 256, (size_t)(n1 - n0) * 512);
      if (K > 1) discard_lines(uimg + (size_t)n0 * P * (K - 1) * 256, (size_t)(n1 - n0) * P * (K - 1) * 512);
    }

    // ================= neighbour lists of my nodes =========================================================
    phase_lists(p, rowbase, n0, n1, rowbits, colbits, reinterpret_cast<int32_t*>(smem), warp, lane);
    __syncthreads();
    PROF_MARK(2);

    // ================= scores ==================================================================================
    if (kq) {
      // R_p[tile] = W_p^T x^T on tcgen05 (SS form): D[128 features x 64 nodes]
      ++gphase;
      if (warp == TMA_WARP) {
        if (tc::elect_one()) {
          for (int t = split; t < p.tiles; t += p.nsplit) {
            const int st = (int)(q % NST);
            mbar_wait_timed(&empty[st], ((q / NST) & 1u) ^ 1u, p.status, 1, wait_a);
            tc::mbar_arrive_expect_tx(&full[st], (uint32_t)STAGE_BYTES);
            const uint32_t dst = tc::smem_u32(smem + (size_t)st * STAGE_BYTES);
#pragma unroll
            for (int a = 0; a < 4; ++a) tensor_g2s_3d(dst + a * ATOM_B, tmx, a * 64, t * TN, team * 2 + par, &full[st]);
            ++q;
          }
        }
        __syncwarp();
      } else if (warp == MMA_WARP) {
        for (int t = split; t < p.tiles; t += p.nsplit) {
          const int acc = (int)(oc & 1u);
          if (lane == 0) mbar_wait_timed(&acc_empty[acc], ((oc >> 1) & 1u) ^ 1u, p.status, 2, wait_b);
          __syncwarp();
          tc::tc_fence_after();
          const int st = (int)(q % NST);
          if (lane == 0) mbar_wait_timed(&full[st], (q / NST) & 1u, p.status, 3, wait_a);
          __syncwarp();
          tc::tc_fence_after();
          if (tc::elect_one()) {
            const uint32_t tmem_d = tmem_base + (uint32_t)(ACC_COL0 + acc * TN);
            const uint32_t sb = tc::smem_u32(smem + (size_t)st * STAGE_BYTES);
            const uint32_t wb = tc::smem_u32(wimg);
#pragma unroll
            for (int at = 0; at < 2; ++at) {
              const uint64_t z_hi = tc::make_sw128_desc(sb + at * ATOM_B);
              const uint64_t z_lo = tc::make_sw128_desc(sb + (2 + at) * ATOM_B);
              const uint64_t w_hi = tc::make_sw128_desc(wb + at * WATOM);
              const uint64_t w_lo = tc::make_sw128_desc(wb + (2 + at) * WATOM);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                tc::umma_bf16(tmem_d, w_hi + adv, z_hi + adv, idesc, (at | kk) != 0);
                tc::umma_bf16(tmem_d, w_lo + adv, z_hi + adv, idesc, 1);
                tc::umma_bf16(tmem_d, w_hi + adv, z_lo + adv, idesc, 1);
              }
            }
            tc::umma_commit(&empty[st]);
            tc::umma_commit(&acc_full[acc]);
          }
          __syncwarp();
          ++q;
          ++oc;
        }
      } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
        const int qd = warp & 3;
        const int f = qd * 32 + lane;
        for (int t = split; t < p.tiles; t += p.nsplit) {
          const int acc = (int)(oc & 1u);
          if (lane == 0) mbar_wait_timed(&acc_full[acc], (oc >> 1) & 1u, p.status, 4, wait_a);
          __syncwarp();
          tc::tc_fence_after();
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
            const int m0 = t * TN + 32 * hh;
            float* dst = sproj + (unsigned)((m0 * P + head) * FT + f);
            const int left = N - m0;
            if (P == 4) epi_store<4 * FT, false, false>(dst, 0, v, left, 0.f, false);
            else if (P == 2) epi_store<2 * FT, false, false>(dst, 0, v, left, 0.f, false);
            else epi_store<FT, false, false>(dst, 0, v, left, 0.f, false);
          }
          ++oc;
        }
        if (warp == EPI_WARP0 && lane == 0) *gemm_done = gphase;
      } else if (bg) {
        scan_job(p, nxt, r, lane, scan_ctr, gemm_done, gphase);
      }
    } else {
      if (P == 4) phase_mixer_gm<4>(p, xb, n0, n1, warp, lane, gm_cd, sproj);
      else if (P == 2) phase_mixer_gm<2>(p, xb, n0, n1, warp, lane, gm_cd, sproj);
      else phase_mixer_gm<1>(p, xb, n0, n1, warp, lane, gm_cd, sproj);
    }
    team_barrier(bar, bar_target, TEAM, p.status, prof, prof_t, 3);

    // ================= attention ===================================================================================
    if (kq) {
      if (P == 4) phase_attention_kq<4>(p, rowbase, xb, n0, n1, warp, lane, sproj, esc);
      else if (P == 2) phase_attention_kq<2>(p, rowbase, xb, n0, n1, warp, lane, sproj, esc);
      else phase_attention_kq<1>(p, rowbase, xb, n0, n1, warp, lane, sproj, esc);
    } else {
      if (P == 4) phase_attention_gm<4>(p, rowbase, n0, n1, warp, lane, sproj, esc);
      else if (P == 2) phase_attention_gm<2>(p, rowbase, n0, n1, warp, lane, sproj, esc);
      else phase_attention_gm<1>(p, rowbase, n0, n1, warp, lane, sproj, esc);
    }

    // ================= taps ========================================================================================
    for (int k = 1; k < K; ++k) {
      team_barrier(bar, bar_target, TEAM, p.status, prof, prof_t, 3 + 2 * k);
      if (k == 1 && kq && !p.save && n1 > n0)        // R of my rows: read by this CTA's attention only, and that is over
        discard_lines(sproj + (size_t)n0 * P * FT, (size_t)(n1 - n0) * P * FT * 4);
      if (P == 4) phase_gather<4>(p, rowbase, xb, n0, n1, warp, lane, k, uimg, taps);
      else if (P == 2) phase_gather<2>(p, rowbase, xb, n0, n1, warp, lane, k, uimg, taps);
      else phase_gather<1>(p, rowbase, xb, n0, n1, warp, lane, k, uimg, taps);
    }
    if (K > 1) team_barrier(bar, bar_target, TEAM, p.status, prof, prof_t, 3 + 2 * K);
    if (!p.save && K > 2 && n1 > n0)                 // fp32 u_1 of my rows: the last gather level has read it
      discard_lines(taps + (size_t)n0 * P * (K - 1) * FT, (size_t)(n1 - n0) * P * (K - 1) * FT * 4);

    // ================= projection: Y_p[tile] = H_p [x | u_1 | u_2]^T + b, ReLU (TS form, H_p in TMEM) ==========
    ++gphase;
    if (warp == TMA_WARP) {
      if (tc::elect_one()) {
        for (int t = split; t < p.tiles; t += p.nsplit) {
          for (int s = 0; s < K; ++s) {
            const int st = (int)(q % NST);
            mbar_wait_timed(&empty[st], ((q / NST) & 1u) ^ 1u, p.status, 5, wait_a);
            tc::mbar_arrive_expect_tx(&full[st], (uint32_t)STAGE_BYTES);
            const uint32_t dst = tc::smem_u32(smem + (size_t)st * STAGE_BYTES);
            if (s == 0) {
#pragma unroll
              for (int a = 0; a < 4; ++a) tensor_g2s_3d(dst + a * ATOM_B, tmx, a * 64, t * TN, team * 2 + par, &full[st]);
            } else {
              const int c0 = (head * (K - 1) + (s - 1)) * 256;
#pragma unroll
              for (int a = 0; a < 4; ++a) tensor_g2s_3d(dst + a * ATOM_B, tmu, c0 + a * 64, t * TN, team, &full[st]);
            }
            ++q;
          }
        }
      }
      __syncwarp();
    } else if (warp == MMA_WARP) {
      for (int t = split; t < p.tiles; t += p.nsplit) {
        const int acc = (int)(oc & 1u);
        if (lane == 0) mbar_wait_timed(&acc_empty[acc], ((oc >> 1) & 1u) ^ 1u, p.status, 6, wait_b);
        __syncwarp();
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(ACC_COL0 + acc * TN);
        for (int s = 0; s < K; ++s) {
          const int st = (int)(q % NST);
          if (lane == 0) mbar_wait_timed(&full[st], (q / NST) & 1u, p.status, 7, wait_a);
          __syncwarp();
          tc::tc_fence_after();
          if (tc::elect_one()) {
            const uint32_t sb = tc::smem_u32(smem + (size_t)st * STAGE_BYTES);
            const uint32_t h_hi = tmem_base + (uint32_t)(s * (SK / 2));
            const uint32_t h_lo = h_hi + (uint32_t)(KG / 2);
#pragma unroll
            for (int at = 0; at < 2; ++at) {
              const uint64_t z_hi = tc::make_sw128_desc(sb + at * ATOM_B);
              const uint64_t z_lo = tc::make_sw128_desc(sb + (2 + at) * ATOM_B);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t adv = (uint64_t)((kk * 32) >> 4);
                const uint32_t col = (uint32_t)(at * 32 + kk * 8);
                tc::umma_bf16_ts(tmem_d, h_hi + col, z_hi + adv, idesc, (s | at | kk) != 0);
                tc::umma_bf16_ts(tmem_d, h_lo + col, z_hi + adv, idesc, 1);
                tc::umma_bf16_ts(tmem_d, h_hi + col, z_lo + adv, idesc, 1);
              }
            }
            tc::umma_commit(&empty[st]);
            if (s == K - 1) tc::umma_commit(&acc_full[acc]);
          }
          __syncwarp();
          ++q;
        }
        ++oc;
      }
    } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
      const int qd = warp & 3;
      const int f = qd * 32 + lane;
      const float bias = p.bias ? __ldg(p.bias + f) : 0.f;
      const bool y_dense = p.y_sn == (long)P * FT;
      float* yb = p.y + (long)b * p.y_sb + (long)head * FT + f;
      for (int t = split; t < p.tiles; t += p.nsplit) {
        const int acc = (int)(oc & 1u);
        if (lane == 0) mbar_wait_timed(&acc_full[acc], (oc >> 1) & 1u, p.status, 8, wait_a);
        __syncwarp();
        tc::tc_fence_after();
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
          const int m0 = t * TN + 32 * hh;
          float* dst = yb + (long)m0 * p.y_sn;
          const int left = N - m0;
          if (!y_dense) epi_store<0, true, true>(dst, p.y_sn, v, left, bias, p.relu != 0);
          else if (P == 4) epi_store<4 * FT, true, true>(dst, 0, v, left, bias, p.relu != 0);
          else if (P == 2) epi_store<2 * FT, true, true>(dst, 0, v, left, bias, p.relu != 0);
          else epi_store<FT, true, true>(dst, 0, v, left, bias, p.relu != 0);
        }
        ++oc;
      }
      if (warp == EPI_WARP0 && lane == 0) *gemm_done = gphase;
    } else if (bg) {
      scan_job(p, nxt, r, lane, scan_ctr, gemm_done, gphase);
    }
    // no barrier here: the next instance's scan only writes buffers nobody reads any more (the x image is double
    // buffered), and every later phase of it sits behind a team barrier all CTAs reach after this projection
    __syncthreads();
    PROF_MARK(11);
    if (K == 1 && kq && !p.save && n1 > n0) discard_lines(sproj + (size_t)n0 * P * FT, (size_t)(n1 - n0) * P * FT * 4);
  }

  if (lane == 0) {
    if (warp == TMA_WARP) prof[15] = wait_a;                       // producer: waiting for a free stage
    if (warp == MMA_WARP) { prof[12] = wait_a; prof[13] = wait_b; } // issuer: waiting for operands / for a free accumulator
    if (warp == EPI_WARP0) prof[14] = wait_a;                      // epilogue: waiting for a finished accumulator
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
}

// bf16 image tensor map: dims {cols, rows, slabs}, box 64 x 64 x 1, SWIZZLE_128B, zero fill out of bounds
bool make_image_map(CUtensorMap* tm, const void* base, long cols, long rows, long slabs) {
  tma::EncodeTiledFn enc = tma::encode_tiled_fn();
  if (enc == nullptr || base == nullptr || ((uintptr_t)base % 128) != 0) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slabs};
  const cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)cols * 2 * (cuuint64_t)rows};
  const cuuint32_t box[3] = {64, 64, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct WsLayout {
  int nteams;
  size_t off_bar, off_prof, off_rowbits, off_colbits, off_ximg, off_uimg, off_sproj, off_taps, total;
};

// status (64 B) | team counters | per-team scratch
WsLayout ws_layout(int B, int N, int K, int P, int D, int mode, int save, int TEAM) {
  WsLayout L{};
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  L.nteams = sms / TEAM;
  if (L.nteams > B) L.nteams = B;
  if (L.nteams < 1) L.nteams = 1;
  const size_t T = (size_t)L.nteams;
  const int WS = ((N + 31) / 32 + 3) / 4 * 4;
  size_t o = 64;
  L.off_bar = o; o = align_up(o + T * 128, 1024);
  L.off_prof = o; o = align_up(o + T * TEAM * 16 * 8, 1024);
  L.off_rowbits = o; o = align_up(o + T * 2 * N * WS * 4, 1024);
  L.off_colbits = o; o = align_up(o + T * 2 * N * WS * 4, 1024);
  L.off_ximg = o; o = align_up(o + T * 2 * N * 512, 1024);
  L.off_uimg = o; o = align_up(o + T * N * P * (size_t)(K > 1 ? K - 1 : 1) * 512, 1024);
  L.off_sproj = o;
  if (!save) o = align_up(o + T * N * P * (mode == MAGAT_MODE_KEYQUERY ? FT : 2) * 4, 1024);
  L.off_taps = o;
  if (!save && K > 2) o = align_up(o + T * N * P * (size_t)(K - 1) * FT * 4, 1024);
  L.total = o;
  return L;
}

}  // namespace

}  // namespace magat

using namespace magat;

extern "C" int magat_gat_fused_supported(int N, int G, int F, int K, int P, int D, int mode, int concat) {
  if (G != FT || F != FT || K < 1 || K > 3 || !concat) return 0;
  if (P != 1 && P != 2 && P != 4) return 0;
  if (mode != MAGAT_MODE_KEYQUERY && mode != MAGAT_MODE_GAT_MODIFIED) return 0;
  if (N < TN || N % 4 != 0 || N > 16384) return 0;
  if (D < 4 || D > 32 || D % 4 != 0) return 0;
  return 1;
}

static int team_of(int requested, int P) {
  const int t = requested > 0 ? requested : TEAM_DEFAULT;
  return (t == 8 || t == 16) && t % P == 0 ? t : 0;
}

extern "C" size_t magat_gat_fused_workspace_bytes(int B, int N, int K, int P, int D, int mode, int save, int team) {
  if (B < 1 || N < 1 || K < 1 || P < 1 || D < 1 || team_of(team, P) == 0) return 0;
  return ws_layout(B, N, K, P, D, mode, save, team_of(team, P)).total;
}

extern "C" int magat_gat_forward_fused(const magat_gat_fused_args* a, void* stream) {
  MAGAT_REQUIRE(a != nullptr, MAGAT_E_BAD_ARG, "magat_gat_forward_fused: null args");
  MAGAT_REQUIRE(a->B >= 1 && a->N >= 1, MAGAT_E_BAD_ARG, "magat_gat_forward_fused: B=%d N=%d", a->B, a->N);
  MAGAT_REQUIRE(magat_gat_fused_supported(a->N, a->G, a->F, a->K, a->P, a->D, a->mode, a->concat), MAGAT_E_UNSUPPORTED,
                "magat_gat_forward_fused: shape not covered (needs G=F=128, K<=3, P in {1,2,4}, concat, N%%4==0, "
                "N>=64, D%%4==0, D<=32; got N=%d G=%d F=%d K=%d P=%d D=%d)", a->N, a->G, a->F, a->K, a->P, a->D);
  MAGAT_REQUIRE(a->S && a->x && a->weight && a->filterWeight && a->y && a->nbr_out && a->nbr_in && a->slot_in &&
                    a->att && a->workspace,
                MAGAT_E_BAD_ARG, "magat_gat_forward_fused: null pointer");
  MAGAT_REQUIRE(a->s_dtype == MAGAT_DT_F32 || a->s_dtype == MAGAT_DT_F64, MAGAT_E_BAD_ARG,
                "magat_gat_forward_fused: GSO dtype must be fp32 or fp64");
  MAGAT_REQUIRE(a->mode != MAGAT_MODE_GAT_MODIFIED || (a->mixer && a->weight_bias), MAGAT_E_BAD_ARG,
                "magat_gat_forward_fused: GAT_modified needs mixer and weight_bias");
  MAGAT_REQUIRE(!a->save || ((a->K == 1 || a->taps) && a->sproj), MAGAT_E_BAD_ARG,
                "magat_gat_forward_fused: save = 1 needs the taps and sproj buffers");
  auto al16 = [](const void* q) { return ((uintptr_t)q % 16) == 0; };
  MAGAT_REQUIRE(al16(a->S) && al16(a->x) && al16(a->y) && al16(a->att) && al16(a->nbr_out) && al16(a->nbr_in) &&
                    al16(a->slot_in) && (a->x_sn % 4) == 0 && (a->x_sb % 4) == 0 &&
                    (a->y_sn % 4) == 0 && (a->y_sb % 4) == 0 && a->y_sc == 1 && a->x_sn >= FT &&
                    (!a->taps || al16(a->taps)) && (!a->sproj || al16(a->sproj)) && ((uintptr_t)a->workspace % 1024) == 0,
                MAGAT_E_ALIGN, "magat_gat_forward_fused: pointers must be 16 B aligned (workspace 1024 B), strides "
                "multiples of 4 floats, unit channel stride");
  MAGAT_REQUIRE((long)a->B * a->N * a->D * a->P < (1l << 31) && (long)a->N * a->x_sn < (1l << 31), MAGAT_E_UNSUPPORTED,
                "magat_gat_forward_fused: batch or row stride too large for the 32-bit index math");
  const int TEAM = team_of(a->team, a->P);
  MAGAT_REQUIRE(TEAM != 0, MAGAT_E_BAD_ARG, "magat_gat_forward_fused: team must be 0 (default), 8 or 16 (got %d)", a->team);
  const WsLayout L = ws_layout(a->B, a->N, a->K, a->P, a->D, a->mode, a->save, TEAM);
  MAGAT_REQUIRE(a->ws_bytes >= L.total, MAGAT_E_BAD_ARG, "magat_gat_forward_fused: workspace %zu B < %zu B", a->ws_bytes,
                L.total);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  int rc = ensure_dyn_smem(KID_FUSED, (const void*)k_gat_fused, SMEM_BYTES, "k_gat_fused");
  if (rc) return rc;

  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  FusedParams fp{};
  fp.B = a->B; fp.N = a->N; fp.K = a->K; fp.P = a->P; fp.D = a->D;
  fp.W = (a->N + 31) / 32; fp.WS = (fp.W + 3) / 4 * 4;
  fp.mode = a->mode; fp.relu = a->relu; fp.save = a->save; fp.s_f64 = a->s_dtype == MAGAT_DT_F64;
  fp.nteams = L.nteams; fp.team_size = TEAM; fp.nsplit = TEAM / a->P;
  fp.chunk = (a->N + TEAM - 1) / TEAM;
  fp.tiles = (a->N + TN - 1) / TN;
  fp.S = a->S; fp.x = a->x; fp.x_sb = a->x_sb; fp.x_sn = a->x_sn;
  fp.weight = a->weight; fp.mixer = a->mixer; fp.wb = a->weight_bias; fp.H = a->filterWeight; fp.bias = a->bias;
  fp.y = a->y; fp.y_sb = a->y_sb; fp.y_sn = a->y_sn;
  fp.nbr_out = a->nbr_out; fp.nbr_in = a->nbr_in; fp.slot_in = a->slot_in;
  fp.att = a->att;
  const int sw = a->mode == MAGAT_MODE_KEYQUERY ? FT : 2;
  if (a->save) {
    fp.taps = a->K > 1 ? a->taps : nullptr;
    fp.taps_inst = (long)a->N * a->P * (a->K - 1) * FT; fp.taps_team = 0;
    fp.sproj = a->sproj; fp.sproj_inst = (long)a->N * a->P * sw; fp.sproj_team = 0;
  } else {
    fp.taps = a->K > 2 ? reinterpret_cast<float*>(ws + L.off_taps) : nullptr;
    fp.taps_inst = 0; fp.taps_team = (long)a->N * a->P * (a->K - 1) * FT;
    fp.sproj = reinterpret_cast<float*>(ws + L.off_sproj); fp.sproj_inst = 0; fp.sproj_team = (long)a->N * a->P * sw;
  }
  fp.wprep_out = (a->save && a->mode == MAGAT_MODE_GAT_MODIFIED) ? a->wprep : nullptr;
  fp.rowbits = reinterpret_cast<uint32_t*>(ws + L.off_rowbits);
  fp.colbits = reinterpret_cast<uint32_t*>(ws + L.off_colbits);
  fp.ximg = reinterpret_cast<uint16_t*>(ws + L.off_ximg);
  fp.uimg = reinterpret_cast<uint16_t*>(ws + L.off_uimg);
  fp.bar = reinterpret_cast<unsigned*>(ws + L.off_bar);
  fp.status = reinterpret_cast<int32_t*>(ws);
  fp.prof = reinterpret_cast<long long*>(ws + L.off_prof);
  MAGAT_REQUIRE(make_image_map(&fp.tm_x, fp.ximg, 256, a->N, (long)L.nteams * 2), MAGAT_E_CUDA,
                "magat_gat_forward_fused: cuTensorMapEncodeTiled (x image) failed");
  MAGAT_REQUIRE(make_image_map(&fp.tm_u, fp.uimg, (long)a->P * (a->K > 1 ? a->K - 1 : 1) * 256, a->N, L.nteams),
                MAGAT_E_CUDA, "magat_gat_forward_fused: cuTensorMapEncodeTiled (tap image) failed");
  cudaError_t e = cudaMemsetAsync(ws, 0, L.off_prof + (size_t)L.nteams * TEAM * 16 * 8, st);
  if (e != cudaSuccess) {
    set_error("magat_gat_forward_fused: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return MAGAT_E_CUDA;
  }
  void* args[] = {(void*)&fp};
  e = cudaLaunchCooperativeKernel((const void*)k_gat_fused, dim3(L.nteams * TEAM), dim3(NTHREADS), args, SMEM_BYTES, st);
  if (e != cudaSuccess) {
    set_error("magat_gat_forward_fused: cooperative launch: %s", cudaGetErrorString(e));
    return MAGAT_E_CUDA;
  }
  return check_launch("k_gat_fused", st);
}
