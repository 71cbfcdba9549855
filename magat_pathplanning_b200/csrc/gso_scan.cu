// GSO -> adjacency.  The dense graph-shift operator is only ever used as an edge mask
// (|S| > 1e-9, graphML.py:1274-1276 and :808-809); this file reads it exactly once and
// turns it into bit masks and padded neighbour lists.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace magat {

template <typename T> __device__ __forceinline__ bool is_edge(T v);
template <> __device__ __forceinline__ bool is_edge<float>(float v) { return fabsf(v) > 1e-9f; }
template <> __device__ __forceinline__ bool is_edge<double>(double v) { return fabs(v) > 1e-9; }

// One warp owns a band of 32 sender rows and sweeps it in 32x32 blocks: every load is a full
// 128 B (fp32) row segment, ballots give the row words, each lane accumulates the transposed
// (column) word of its own column.  A CTA is 8 consecutive bands.
// NZ = true: the predicate of the non-attentional graph filter (BatchLSIGF multiplies by S itself, graphML.py:5571:
// every entry that is not exactly zero takes part, NaN included).
template <typename T, bool NZ = false>
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
      const bool e = NZ ? (v != T(0)) : is_edge<T>(v);
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

// ---- TMA-staged scan ----------------------------------------------------------------------------
// Persistent CTAs; one elected thread streams whole GSO rows into a shared-memory ring with bulk async copies
// (cp.async.bulk, completion counted on an mbarrier), so ~190 KB per SM are in flight no matter how few
// registers or warps are resident; eight consumer warps turn each staged chunk into row words (three
// xor-shuffles per row) and accumulate the transposed column words in registers over the 32 rows of a band.
template <typename T> struct Edge4;
template <> struct Edge4<float> {
  static __device__ __forceinline__ uint32_t nib(const float* p) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    return (uint32_t)(fabsf(v.x) > 1e-9f) | ((uint32_t)(fabsf(v.y) > 1e-9f) << 1) |
           ((uint32_t)(fabsf(v.z) > 1e-9f) << 2) | ((uint32_t)(fabsf(v.w) > 1e-9f) << 3);
  }
};
template <> struct Edge4<double> {
  static __device__ __forceinline__ uint32_t nib(const double* p) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
    return (uint32_t)(fabs(a.x) > 1e-9) | ((uint32_t)(fabs(a.y) > 1e-9) << 1) | ((uint32_t)(fabs(b.x) > 1e-9) << 2) |
           ((uint32_t)(fabs(b.y) > 1e-9) << 3);
  }
};

// one staged element (shared-space load; the generic-address loads the compiler emitted for a
// plain pointer into the ring were measurably slower)
template <typename T> struct EdgeS;
template <> struct EdgeS<float> {
  static __device__ __forceinline__ float ld(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
  }
};
template <> struct EdgeS<double> {
  static __device__ __forceinline__ double ld(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
  }
};

// R staged rows x one segment of 32 KK columns.  FULL: all R rows are there (no per-row guards around the ballots).
// Columns past N (last segment only) read whatever follows in the ring: their bits are cut out of the row words by
// the masks below and their column words are never stored.
// a[k] = (lane == k) ? mask of the in-range lanes of column group k : 0 -- lane k < KK stores word k.
template <typename T, int R, int KK, bool FULL, bool NZ>
__device__ __forceinline__ void scan_rows(uint32_t addr, uint32_t row_bytes, int nrows, uint32_t bit0,
                                          const uint32_t (&a)[KK], bool wr, uint32_t* rw, int W, uint32_t (&col)[KK]) {
  constexpr int RB = R < 8 ? R : 8;                   // rows whose loads are issued back to back
  unsigned woff = 0;
#pragma unroll
  for (int r0 = 0; r0 < R; r0 += RB) {
    if (!FULL && r0 >= nrows) break;
    T v[RB][KK];
    uint32_t ad = addr;
#pragma unroll
    for (int u = 0; u < RB; ++u) {
#pragma unroll
      for (int k = 0; k < KK; ++k) v[u][k] = EdgeS<T>::ld(ad + (uint32_t)(32 * k * sizeof(T)));
      ad += row_bytes;
    }
    addr = ad;
#pragma unroll
    for (int u = 0; u < RB; ++u) {
      const int r = r0 + u;
      if (FULL || r < nrows) {
        const uint32_t bit = bit0 << r;
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < KK; ++k) {
          const bool pk = NZ ? (v[u][k] != T(0)) : is_edge<T>(v[u][k]);
          w |= __ballot_sync(0xffffffffu, pk) & a[k];
          col[k] |= pk ? bit : 0u;
        }
        if (wr) rw[woff] = w;
      }
      woff += (unsigned)W;
    }
  }
}

constexpr int kScanWarps = 16;                // consumer warps; warp 16 is the copy issuer
constexpr int kScanRing = 192 * 1024;

// KK: 32-column groups per consumer warp (2 for N <= 1024 -- sixteen warps on 64-column segments --, 4 up to N = 2048)
// NZ: the non-attentional filter's predicate (every entry that is not exactly zero, NaN included) instead of |s| > 1e-9
template <typename T, int R, int KK, bool NZ>
__global__ void __launch_bounds__((kScanWarps + 1) * 32, 1) k_gso_scan_tma(const T* __restrict__ S, int N, int W,
                                                                           long bands, int nstages,
                                                                           uint32_t* __restrict__ rowbits,
                                                                           uint32_t* __restrict__ colbits) {
  extern __shared__ uint8_t scan_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)scan_smem_raw + 127) & ~(uintptr_t)127);
  const size_t row_bytes = (size_t)N * sizeof(T);
  const size_t chunk_bytes = (size_t)R * row_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)nstages * chunk_bytes);
  uint64_t* empty = full + nstages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], kScanWarps);
    }
    tc::fence_barrier_init();
  }
  __syncthreads();
  const int chunks_per_band = 32 / R;
  const int segs = (N + 32 * KK - 1) / (32 * KK);       // <= kScanWarps

  if (warp == kScanWarps) {
    // ===== copy issuer ======================================================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long band = blockIdx.x; band < bands; band += gridDim.x) {
        const long b = band / W;
        const int i0 = (int)(band - b * W) * 32;
        for (int c = 0; c < chunks_per_band; ++c) {
          const int r_first = i0 + c * R;
          int nrows = N - r_first;
          if (nrows > R) nrows = R;
          if (nrows <= 0) break;
          tc::mbar_wait(&empty[stage], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full[stage], (uint32_t)(nrows * row_bytes));
          const T* src = S + ((size_t)b * N + r_first) * N;
          // rows of one instance are contiguous: one bulk copy for the whole chunk
          tc::bulk_g2s(smem + (size_t)stage * chunk_bytes, src, (uint32_t)(nrows * row_bytes), &full[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===== consumers ========================================================================
    // Warp w owns the 32 KK-column segment w of every staged row.  Lane l reads columns 32k + l of the segment
    // (conflict-free scalar shared loads): one ballot per 32 columns IS the row word, and the lane's own predicate is
    // its bit of the transposed column word -- no nibble packing, no shuffles (the first version of this loop spent 106
    // instructions per 128 columns of a row and paced the kernel at 56 % of the DRAM peak; this one ~28).
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t smem_s = tc::smem_u32(smem);
    const int seg = warp;
    const bool active = seg < segs;
    const int j = seg * (32 * KK) + lane;
    uint32_t am[KK];
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      const uint32_t vm = __ballot_sync(0xffffffffu, j + 32 * k < N);
      am[k] = lane == k ? vm : 0u;
      asm volatile("" : "+r"(am[k]));                   // keep it a register value (not re-derived per row)
    }
    const bool wr = active && lane < KK && seg * KK + lane < W;
    for (long band = blockIdx.x; band < bands; band += gridDim.x) {
      const long b = band / W;
      const int rb = (int)(band - b * W);
      const int i0 = rb * 32;
      uint32_t col[KK];
#pragma unroll
      for (int k = 0; k < KK; ++k) col[k] = 0u;
      for (int c = 0; c < chunks_per_band; ++c) {
        const int r_first = i0 + c * R;
        int nrows = N - r_first;
        if (nrows > R) nrows = R;
        if (nrows <= 0) break;
        tc::mbar_wait(&full[stage], phase);
        if (active) {
          const uint32_t addr = smem_s + (uint32_t)((size_t)stage * chunk_bytes) + (uint32_t)(j * sizeof(T));
          const uint32_t bit0 = 1u << (c * R);                // bit of the chunk's first row inside the band
          uint32_t* rw = rowbits + ((size_t)b * N + r_first) * W + seg * KK + lane;
          if (nrows == R) scan_rows<T, R, KK, true, NZ>(addr, (uint32_t)row_bytes, nrows, bit0, am, wr, rw, W, col);
          else scan_rows<T, R, KK, false, NZ>(addr, (uint32_t)row_bytes, nrows, bit0, am, wr, rw, W, col);
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&empty[stage]);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
      if (active) {
#pragma unroll
        for (int k = 0; k < KK; ++k) {
          const int jj = j + 32 * k;
          if (jj < N) colbits[((size_t)b * N + jj) * W + rb] = col[k];
        }
      }
    }
  }
}

// ---- f1: edge mask straight from agent positions ------------------------------------------------
// utils/new_simulator.py:823-827 builds the GSO as squareform(pdist(pos)) < commR with a zeroed diagonal (the
// normalisation that follows only scales it, and the layer only tests |S| > 1e-9).  Same predicate here, in fp64 as
// scipy computes it: sqrt(dx^2 + dy^2) < R, i != j -- evaluated as d2 < T with T the smallest double whose (correctly
// rounded, monotone) square root reaches R, found on the host, so no fp64 sqrt runs per pair (it made this kernel
// slower than scanning the dense GSO); squares and sum are rounded separately like the x86 build of scipy does.
// One block = 32 rows of one instance, positions staged in shared memory; a warp emits the W words of its rows with
// one ballot each.  The mask is symmetric: colbits = rowbits.
template <typename T>
__global__ void __launch_bounds__(256) k_gso_from_positions(const T* __restrict__ pos, int N, int W, double thr2,
                                                            uint32_t* __restrict__ rowbits,
                                                            uint32_t* __restrict__ colbits) {
  extern __shared__ double pos_s[];                    // [N][2]
  const int b = blockIdx.y;
  const T* pb = pos + (size_t)b * N * 2;
  for (int e = threadIdx.x; e < 2 * N; e += blockDim.x) pos_s[e] = (double)pb[e];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rr = warp; rr < 32; rr += 8) {
    const int i = blockIdx.x * 32 + rr;
    if (i >= N) break;
    const double xi = pos_s[2 * i], yi = pos_s[2 * i + 1];
    const size_t base = ((size_t)b * N + i) * W;
#pragma unroll 4
    for (int w = 0; w < W; ++w) {
      const int j = w * 32 + lane;
      bool e = false;
      if (j < N && j != i) {
        const double dx = xi - pos_s[2 * j], dy = yi - pos_s[2 * j + 1];
        e = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < thr2;
      }
      const uint32_t word = __ballot_sync(0xffffffffu, e);
      if (lane == 0) rowbits[base + w] = word;
      if (lane == 1) colbits[base + w] = word;
    }
  }
}

// The same mask through a cell list (SURVEY 8f row f1 asks for it): one CTA per instance bins its agents into square
// cells a hair wider than the radius, and every agent tests only the agents of its 3 x 3 cells with the SAME fp64
// predicate as above -- ~11 candidates instead of N - 1 at the reference's density.  Agent i owns row i of the
// (zero-initialised) masks and sets its bits with plain read-modify-writes; the mask is symmetric, so colbits takes the
// same words.  An instance whose positions are not all finite, or whose bounding box needs more than kCellMax cells,
// takes the all-pairs loop inside the same CTA.
constexpr int kCellMax = 4096;

template <typename T>
__global__ void __launch_bounds__(256) k_gso_cells_from_positions(const T* __restrict__ pos, int N, int W, double thr2,
                                                                  double cell, uint32_t* __restrict__ rowbits,
                                                                  uint32_t* __restrict__ colbits) {
  extern __shared__ double cell_smem[];
  double* pos_s = cell_smem;                                        // [N][2]
  int* cell_of = reinterpret_cast<int*>(pos_s + 2 * (size_t)N);     // [N]
  int* ids = cell_of + N;                                           // [N] agents sorted by cell
  int* start = ids + N;                                             // [kCellMax + 1]
  int* fill = start + kCellMax + 1;                                 // [kCellMax]
  __shared__ double red[4][8];
  __shared__ int bad_s, gx_s, gy_s;
  __shared__ double x0_s, y0_s;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* pb = pos + (size_t)b * N * 2;
  double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  int bad = 0;
  for (int i = tid; i < N; i += blockDim.x) {
    const double x = (double)pb[2 * i], y = (double)pb[2 * i + 1];
    pos_s[2 * i] = x;
    pos_s[2 * i + 1] = y;
    if (!(fabs(x) < INFINITY) || !(fabs(y) < INFINITY)) bad = 1;
    xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
    ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if (tid == 0) bad_s = 0;
  __syncthreads();
  if (lane == 0) {
    red[0][warp] = xmin; red[1][warp] = xmax; red[2][warp] = ymin; red[3][warp] = ymax;
    if (bad) bad_s = 1;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) {
      red[0][0] = fmin(red[0][0], red[0][w]); red[1][0] = fmax(red[1][0], red[1][w]);
      red[2][0] = fmin(red[2][0], red[2][w]); red[3][0] = fmax(red[3][0], red[3][w]);
    }
    const double nx = floor((red[1][0] - red[0][0]) / cell) + 1.0, ny = floor((red[3][0] - red[2][0]) / cell) + 1.0;
    if (bad_s || !(nx * ny <= (double)kCellMax)) {
      bad_s = 1;
    } else {
      gx_s = (int)nx; gy_s = (int)ny; x0_s = red[0][0]; y0_s = red[2][0];
    }
  }
  __syncthreads();
  uint32_t* rb = rowbits + (size_t)b * N * W;
  uint32_t* cb = colbits + (size_t)b * N * W;
  if (bad_s) {
    // all pairs (NaN compares false: no edge), one warp per row
    for (int i = warp; i < N; i += 8) {
      const double xi = pos_s[2 * i], yi = pos_s[2 * i + 1];
      for (int w = 0; w < W; ++w) {
        const int j = w * 32 + lane;
        bool e = false;
        if (j < N && j != i) {
          const double dx = xi - pos_s[2 * j], dy = yi - pos_s[2 * j + 1];
          e = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < thr2;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, e);
        if (lane == 0) rb[(size_t)i * W + w] = word;
        if (lane == 1) cb[(size_t)i * W + w] = word;
      }
    }
    return;
  }
  const int gx = gx_s, gy = gy_s, ncell = gx * gy;
  const double x0 = x0_s, y0 = y0_s;
  for (int c = tid; c <= ncell; c += blockDim.x) start[c] = 0;
  for (int c = tid; c < ncell; c += blockDim.x) fill[c] = 0;
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) {
    int cx = (int)floor((pos_s[2 * i] - x0) / cell), cy = (int)floor((pos_s[2 * i + 1] - y0) / cell);
    cx = min(max(cx, 0), gx - 1);
    cy = min(max(cy, 0), gy - 1);
    const int c = cy * gx + cx;
    cell_of[i] = c;
    atomicAdd(&start[c + 1], 1);
  }
  __syncthreads();
  if (warp == 0) {                                   // inclusive scan of the counts: start[c] = first slot of cell c
    int carry = 0;
    for (int c0 = 1; c0 <= ncell; c0 += 32) {
      const int c = c0 + lane;
      int v = c <= ncell ? start[c] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
      }
      if (c <= ncell) start[c] = v + carry;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) {
    const int c = cell_of[i];
    ids[start[c] + atomicAdd(&fill[c], 1)] = i;
  }
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) {
    const double xi = pos_s[2 * i], yi = pos_s[2 * i + 1];
    const int c = cell_of[i], cx = c % gx, cy = c / gx;
    uint32_t* ri = rb + (size_t)i * W;
    uint32_t* ci = cb + (size_t)i * W;
    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, gy - 1); ++yy) {
      const int c_lo = yy * gx + max(cx - 1, 0), c_hi = yy * gx + min(cx + 1, gx - 1);
      for (int k = start[c_lo]; k < start[c_hi + 1]; ++k) {             // the three cells of a row are contiguous
        const int j = ids[k];
        if (j == i) continue;
        const double dx = xi - pos_s[2 * j], dy = yi - pos_s[2 * j + 1];
        if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < thr2) {
          const uint32_t word = ri[j >> 5] | (1u << (j & 31));
          ri[j >> 5] = word;
          ci[j >> 5] = word;
        }
      }
    }
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

// Same statistics, W == 4 * LPR: LPR lanes share a row (one uint4 of each mask per lane), so a warp digests
// 32 / LPR rows per load pair and the per-row sums cost log2(LPR) shuffles instead of five.
template <int LPR>
__global__ void __launch_bounds__(256) k_gso_stats_v(const uint4* __restrict__ rowbits4,
                                                     const uint4* __restrict__ colbits4, long n4,
                                                     int32_t* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long stride = (long)gridDim.x * blockDim.x;
  int mo = 0, mi = 0, tot = 0;
  bool sym = true;
  for (long base = (long)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < n4; base += stride) {
    const long t = base + lane;
    uint4 r = make_uint4(0u, 0u, 0u, 0u), c = r;
    if (t < n4) {
      r = __ldcs(rowbits4 + t);
      c = __ldcs(colbits4 + t);
    }
    int od = __popc(r.x) + __popc(r.y) + __popc(r.z) + __popc(r.w);
    int id = __popc(c.x) + __popc(c.y) + __popc(c.z) + __popc(c.w);
    sym = sym && r.x == c.x && r.y == c.y && r.z == c.z && r.w == c.w;
    tot += od;
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) {
      od += __shfl_xor_sync(0xffffffffu, od, o);
      id += __shfl_xor_sync(0xffffffffu, id, o);
    }
    mo = max(mo, od);
    mi = max(mi, id);
  }
  mo = warp_max_i(mo);
  mi = warp_max_i(mi);
  tot = warp_sum_i(tot);
  sym = __all_sync(0xffffffffu, sym);
  if (lane == 0) {
    atomicMax(&stats[0], mo);
    atomicMax(&stats[1], mi);
    if (tot) atomicAdd(&stats[2], tot);
    if (!sym) atomicExch(&stats[3], 0);
  }
}

// Enumerate the set bits of a W-word bit row into out[0..D) (ascending), -1 padded; returns the count.
__device__ __forceinline__ int list_bits(const uint32_t* __restrict__ bits, int W, int D, int lane,
                                         int32_t* __restrict__ out) {
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
      if (pos < D) out[pos] = w * 32 + bit;
      ++pos;
    }
  }
  for (int s = base + lane; s < D; s += 32) out[s] = -1;
  return base;
}

// One warp per node: out-list from its row bits, in-list from its column bits, and for every in-edge
// (i -> n) the slot of n inside row i's out-list = number of set bits of row i below n (one coalesced
// read of row i's words + a warp reduction per edge).
__global__ void __launch_bounds__(256) k_build_ell(const uint32_t* __restrict__ rowbits,
                                                   const uint32_t* __restrict__ colbits, long rows,
                                                   int N, int W, int D, int32_t* __restrict__ nbr_out,
                                                   int32_t* __restrict__ nbr_in,
                                                   int32_t* __restrict__ slot_in,
                                                   int32_t* __restrict__ slot_out) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const long b = row / N;
  const int n = (int)(row - b * N);
  int deg_out = list_bits(rowbits + row * W, W, D, lane, nbr_out + row * D);
  if (deg_out > D) deg_out = D;
  int32_t* nin = nbr_in + row * D;
  int deg_in = list_bits(colbits + row * W, W, D, lane, nin);
  if (deg_in > D) deg_in = D;
  __syncwarp();
  const int sw = n >> 5;
  const uint32_t below = (1u << (n & 31)) - 1u;
  const int grp = lane >> 3, gl = lane & 7;           // four in-edges at a time, eight lanes each
  for (int s0 = 0; s0 < deg_in; s0 += 4) {
    const int s = s0 + grp;
    int rank = 0;
    if (s < deg_in) {
      const int i = nin[s];
      const uint32_t* r = rowbits + ((size_t)b * N + i) * W;
      for (int w = gl; w <= sw; w += 8) rank += __popc(w < sw ? r[w] : (r[w] & below));
    }
    rank += __shfl_xor_sync(0xffffffffu, rank, 1);
    rank += __shfl_xor_sync(0xffffffffu, rank, 2);
    rank += __shfl_xor_sync(0xffffffffu, rank, 4);
    if (gl == 0 && s < deg_in) slot_in[row * D + s] = rank;
  }
  for (int s = deg_in + lane; s < D; s += 32) slot_in[row * D + s] = 0;
  if (slot_out == nullptr) return;
  // out-edge (n -> j): position of n inside column j's in-list = set bits of column j below n
  const int32_t* nout = nbr_out + row * D;
  for (int s0 = 0; s0 < deg_out; s0 += 4) {
    const int s = s0 + grp;
    int rank = 0;
    if (s < deg_out) {
      const int j = nout[s];
      const uint32_t* c = colbits + ((size_t)b * N + j) * W;
      for (int w = gl; w <= sw; w += 8) rank += __popc(w < sw ? c[w] : (c[w] & below));
    }
    rank += __shfl_xor_sync(0xffffffffu, rank, 1);
    rank += __shfl_xor_sync(0xffffffffu, rank, 2);
    rank += __shfl_xor_sync(0xffffffffu, rank, 4);
    if (gl == 0 && s < deg_out) slot_out[row * D + s] = rank;
  }
  for (int s = deg_out + lane; s < D; s += 32) slot_out[row * D + s] = 0;
}

// ---- two-step list builder --------------------------------------------------------------------------
// Step 1, one warp per node: out-list from the row bits, in-list from the column bits.
__global__ void __launch_bounds__(256) k_build_lists(const uint32_t* __restrict__ rowbits,
                                                     const uint32_t* __restrict__ colbits, long rows, int W, int D,
                                                     int32_t* __restrict__ nbr_out, int32_t* __restrict__ nbr_in) {
  const int lane = threadIdx.x & 31;
  const long row = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  list_bits(rowbits + row * W, W, D, lane, nbr_out + row * D);
  list_bits(colbits + row * W, W, D, lane, nbr_in + row * D);
}

// position of `key` in the ascending, -1 padded list l[0..D) (the key is known to be present)
__device__ __forceinline__ int find_slot(const int32_t* __restrict__ l, int D, int key) {
  if ((D & 3) == 0 && D <= 32 && ((uintptr_t)l & 15) == 0) {
    int pos = 0;
    for (int s0 = 0; s0 < D; s0 += 4) {
      const int4 v = __ldg(reinterpret_cast<const int4*>(l + s0));
      // entries below the key (padding is -1: mask it out with the unsigned compare)
      pos += ((unsigned)v.x < (unsigned)key) + ((unsigned)v.y < (unsigned)key) + ((unsigned)v.z < (unsigned)key) +
             ((unsigned)v.w < (unsigned)key);
    }
    return pos;
  }
  int lo = 0, hi = D;                       // lower bound on the unsigned order (-1 sorts last)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((unsigned)__ldg(l + mid) < (unsigned)key) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Step 2, one thread per list entry: where the same edge sits in the list of its other end point.
__global__ void __launch_bounds__(256) k_build_slots(const int32_t* __restrict__ nbr_out,
                                                     const int32_t* __restrict__ nbr_in, long rows, int N, int D,
                                                     int32_t* __restrict__ slot_in, int32_t* __restrict__ slot_out) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * D) return;
  long row, b;
  if (rows * D < (1l << 32)) {              // 32-bit divisions (the 64-bit ones are emulated)
    row = (long)((unsigned)t / (unsigned)D);
    b = (long)((unsigned)row / (unsigned)N);
  } else {
    row = t / D;
    b = row / N;
  }
  const int n = (int)(row - b * N);
  const int i = nbr_in[t];
  slot_in[t] = i >= 0 ? find_slot(nbr_out + (b * N + i) * D, D, n) : 0;
  if (slot_out != nullptr) {
    const int j = nbr_out[t];
    slot_out[t] = j >= 0 ? find_slot(nbr_in + (b * N + j) * D, D, n) : 0;
  }
}

// ---- thread-per-row variants (W % 4 == 0, D % 4 == 0, D <= 32, 16 B aligned buffers) ---------------------
// The warp-per-row kernels above spend ~100 instructions per row on shuffles for ~3.5 edges; here a lane owns
// a whole row: WV independent 16 B loads, then the few set bits are written behind a -1 fill of the row.
// blockIdx.y selects the list: 0 = out (row bits), 1 = in (column bits).
__global__ void __launch_bounds__(128) k_build_lists_t(const uint4* __restrict__ rowbits4,
                                                       const uint4* __restrict__ colbits4, long rows, int WV, int D,
                                                       int32_t* __restrict__ nbr_out, int32_t* __restrict__ nbr_in) {
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const uint4* src = (blockIdx.y ? colbits4 : rowbits4) + row * WV;
  int32_t* out = (blockIdx.y ? nbr_in : nbr_out) + row * D;
  for (int s = 0; s < D; s += 4) *reinterpret_cast<int4*>(out + s) = make_int4(-1, -1, -1, -1);
  int pos = 0;
  for (int v0 = 0; v0 < WV; v0 += 8) {
    uint4 w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = (v0 + i < WV) ? __ldcs(src + v0 + i) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t ws[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t word = ws[c];
        const int base = ((v0 + i) * 4 + c) * 32;
        while (word) {
          const int bit = __ffs(word) - 1;
          word &= word - 1;
          if (pos < D) out[pos] = base + bit;      // same thread, program order: lands on top of the fill
          ++pos;
        }
      }
    }
  }
}

// slots of one list row per thread; blockIdx.y: 0 = slot_in (search the senders' out-lists), 1 = slot_out
__global__ void __launch_bounds__(128) k_build_slots_t(const int32_t* __restrict__ nbr_out,
                                                       const int32_t* __restrict__ nbr_in, long rows, int N, int D,
                                                       int32_t* __restrict__ slot_in, int32_t* __restrict__ slot_out) {
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  const bool second = blockIdx.y != 0;
  const int32_t* mine = (second ? nbr_out : nbr_in) + row * D;
  const int32_t* other = second ? nbr_in : nbr_out;
  int32_t* slot = (second ? slot_out : slot_in) + row * D;
  const long b = batch_of32(row, N);
  const int n = (int)(row - b * N);
  const int32_t* ob = other + b * N * D;
  for (int s0 = 0; s0 < D; s0 += 4) {
    const int4 id = __ldg(reinterpret_cast<const int4*>(mine + s0));
    const int ids[4] = {id.x, id.y, id.z, id.w};
    int r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = ids[u] >= 0 ? find_slot(ob + (long)ids[u] * D, D, n) : 0;
    *reinterpret_cast<int4*>(slot + s0) = make_int4(r[0], r[1], r[2], r[3]);
  }
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

static int launch_gso_stats(const uint32_t* rowbits, const uint32_t* colbits, int B, int N, int W, int32_t* stats,
                            cudaStream_t st) {
  const long rows = (long)B * N;
  const int lpr = W / 4;
  if (W % 4 == 0 && lpr <= 32 && (lpr & (lpr - 1)) == 0 && ((uintptr_t)rowbits % 16) == 0 &&
      ((uintptr_t)colbits % 16) == 0) {
    const long n4 = rows * lpr;
    const int blocks = (int)min((long)(device_sm_count() > 0 ? device_sm_count() : 148) * 8, (n4 + 255) / 256);
    const uint4* r4 = reinterpret_cast<const uint4*>(rowbits);
    const uint4* c4 = reinterpret_cast<const uint4*>(colbits);
    switch (lpr) {
      case 1: k_gso_stats_v<1><<<blocks, 256, 0, st>>>(r4, c4, n4, stats); break;
      case 2: k_gso_stats_v<2><<<blocks, 256, 0, st>>>(r4, c4, n4, stats); break;
      case 4: k_gso_stats_v<4><<<blocks, 256, 0, st>>>(r4, c4, n4, stats); break;
      case 8: k_gso_stats_v<8><<<blocks, 256, 0, st>>>(r4, c4, n4, stats); break;
      case 16: k_gso_stats_v<16><<<blocks, 256, 0, st>>>(r4, c4, n4, stats); break;
      default: k_gso_stats_v<32><<<blocks, 256, 0, st>>>(r4, c4, n4, stats); break;
    }
  } else {
    int blocks = (int)min((long)(device_sm_count() > 0 ? device_sm_count() : 148) * 8, (rows + 7) / 8);
    k_gso_stats<<<blocks, 256, 0, st>>>(rowbits, colbits, rows, W, stats);
  }
  return check_launch("k_gso_stats", st);
}

// nz: the non-attentional filter's predicate (entry != 0, NaN included) instead of |entry| > 1e-9
static int gso_scan_impl(const void* S, int s_dtype, int B, int N, uint32_t* rowbits, uint32_t* colbits, int32_t* stats,
                         void* stream, bool nz) {
  MAGAT_REQUIRE(S && rowbits && colbits && stats, MAGAT_E_BAD_ARG, "magat_gso_scan: null pointer");
  MAGAT_REQUIRE(B >= 1 && N >= 1, MAGAT_E_BAD_ARG, "magat_gso_scan: B=%d N=%d", B, N);
  MAGAT_REQUIRE(B <= 65535, MAGAT_E_BAD_ARG, "magat_gso_scan: B=%d exceeds 65535", B);
  MAGAT_REQUIRE(s_dtype == MAGAT_DT_F32 || s_dtype == MAGAT_DT_F64, MAGAT_E_BAD_ARG,
                "magat_gso_scan: GSO dtype must be fp32 or fp64");
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const int W = (N + 31) / 32;
  const size_t esz = s_dtype == MAGAT_DT_F32 ? 4 : 8;
  int R = 32;
  while (R > 1 && (size_t)R * N * esz > 32 * 1024) R >>= 1;
  const bool tma_ok = N % 4 == 0 && ((uintptr_t)S % 16) == 0 && N <= 2048 && (size_t)R * N * esz <= 48 * 1024;
  if (tma_ok) {
    const int sm_count = device_sm_count();
    const size_t chunk_bytes = (size_t)R * N * esz;
    int nstages = (int)(kScanRing / chunk_bytes);
    if (nstages > 16) nstages = 16;
    const long bands = (long)B * W;
    const int grid = (int)(bands < sm_count ? bands : sm_count);
    // + 1280: the consumers of the last column segment read up to 127 elements past a row's end (masked afterwards);
    // for the last row of the last stage that lands behind the ring
    const size_t smem = (size_t)nstages * chunk_bytes + 2 * nstages * 8 + 256 + 1280;
#define MAGAT_SCAN_K(TT, RR, KKV)                                                                               \
  do {                                                                                                         \
    const int kid = KID_SCAN_BASE + (nz ? 24 : 0) + (KKV == 4 ? 12 : 0) + (sizeof(TT) == 8 ? 6 : 0) +           \
                    (RR >= 32 ? 5 : RR >= 16 ? 4 : RR >= 8 ? 3 : RR >= 4 ? 2 : RR >= 2 ? 1 : 0);              \
    if (nz) {                                                                                                  \
      if (ensure_dyn_smem(kid, (const void*)k_gso_scan_tma<TT, RR, KKV, true>, kScanRing + 2048, "k_gso_scan_tma")) \
        return MAGAT_E_CUDA;                                                                                   \
      k_gso_scan_tma<TT, RR, KKV, true><<<grid, (kScanWarps + 1) * 32, smem, st>>>((const TT*)S, N, W, bands,  \
                                                                                   nstages, rowbits, colbits); \
    } else {                                                                                                   \
      if (ensure_dyn_smem(kid, (const void*)k_gso_scan_tma<TT, RR, KKV, false>, kScanRing + 2048, "k_gso_scan_tma")) \
        return MAGAT_E_CUDA;                                                                                   \
      k_gso_scan_tma<TT, RR, KKV, false><<<grid, (kScanWarps + 1) * 32, smem, st>>>((const TT*)S, N, W, bands, \
                                                                                    nstages, rowbits, colbits); \
    }                                                                                                          \
  } while (0)
#define MAGAT_SCAN(TT, RR)             \
  do {                                 \
    if (N <= 1024) MAGAT_SCAN_K(TT, RR, 2); \
    else MAGAT_SCAN_K(TT, RR, 4);      \
  } while (0)
#define MAGAT_SCAN_R(TT)                          \
  switch (R) {                                    \
    case 32: MAGAT_SCAN(TT, 32); break;           \
    case 16: MAGAT_SCAN(TT, 16); break;           \
    case 8: MAGAT_SCAN(TT, 8); break;             \
    case 4: MAGAT_SCAN(TT, 4); break;             \
    case 2: MAGAT_SCAN(TT, 2); break;             \
    default: MAGAT_SCAN(TT, 1); break;            \
  }
    if (s_dtype == MAGAT_DT_F32) MAGAT_SCAN_R(float) else MAGAT_SCAN_R(double)
#undef MAGAT_SCAN_R
#undef MAGAT_SCAN
#undef MAGAT_SCAN_K
  } else if (!nz && N % 4 == 0 && ((uintptr_t)S % 16) == 0) {
    const int segs = cdiv(N, 128);
    const long units = (long)B * W * segs;
    const int blocks = cdiv(units, 8);
    if (s_dtype == MAGAT_DT_F32)
      k_gso_scan_v4<float><<<blocks, 256, 0, st>>>((const float*)S, N, W, segs, units, rowbits, colbits);
    else
      k_gso_scan_v4<double><<<blocks, 256, 0, st>>>((const double*)S, N, W, segs, units, rowbits, colbits);
  } else {
    dim3 grid(cdiv(W, 8), B);
    if (s_dtype == MAGAT_DT_F32) {
      if (nz) k_gso_scan<float, true><<<grid, 256, 0, st>>>((const float*)S, N, W, rowbits, colbits);
      else k_gso_scan<float><<<grid, 256, 0, st>>>((const float*)S, N, W, rowbits, colbits);
    } else {
      if (nz) k_gso_scan<double, true><<<grid, 256, 0, st>>>((const double*)S, N, W, rowbits, colbits);
      else k_gso_scan<double><<<grid, 256, 0, st>>>((const double*)S, N, W, rowbits, colbits);
    }
  }
  int rc = check_launch(nz ? "k_gso_scan(nonzero)" : "k_gso_scan", (cudaStream_t)stream);
  if (rc) return rc;
  return launch_gso_stats(rowbits, colbits, B, N, W, stats, st);
}

extern "C" int magat_gso_scan(const void* S, int s_dtype, int B, int N, uint32_t* rowbits,
                              uint32_t* colbits, int32_t* stats, void* stream) {
  return gso_scan_impl(S, s_dtype, B, N, rowbits, colbits, stats, stream, false);
}

// att[b][i][s][0] = (float) S[b][i][nbr_out[b][i][s]], 0 beyond the degree: the "attention" of the non-attentional
// filter is the GSO itself (graphML.py:5569-5572)
template <typename T>
__global__ void __launch_bounds__(256) k_gso_edge_values(const T* __restrict__ S, const int32_t* __restrict__ nbr_out,
                                                         long total, int N, int D, float* __restrict__ att) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long row = t / D;
  const int j = nbr_out[t];
  att[t] = j >= 0 ? (float)S[row * N + j] : 0.f;
}

// colbits = transpose of rowbits, one warp per 32 x 32 bit block: lane r holds word w of row i0 + r; ballot c collects
// bit c of all 32 rows, i.e. the word (band i0 / 32) of column w * 32 + c.
__global__ void __launch_bounds__(256) k_bits_transpose(const uint32_t* __restrict__ rowbits, int N, int W, long blocks,
                                                        uint32_t* __restrict__ colbits) {
  const int lane = threadIdx.x & 31;
  const long blk = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (blk >= blocks) return;
  const int w = (int)(blk % W);
  const int band = (int)((blk / W) % W);
  const long b = blk / ((long)W * W);
  const int i = band * 32 + lane;
  const uint32_t x = i < N ? rowbits[((size_t)b * N + i) * W + w] : 0u;
  uint32_t mine = 0u;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const uint32_t v = __ballot_sync(0xffffffffu, (x >> c) & 1u);
    if (lane == c) mine = v;
  }
  const int j = w * 32 + lane;
  if (j < N) colbits[((size_t)b * N + j) * W + band] = mine;
}

extern "C" int magat_gso_from_rowbits(const uint32_t* rowbits, int B, int N, uint32_t* colbits, int32_t* stats,
                                      void* stream) {
  MAGAT_REQUIRE(rowbits && colbits && stats, MAGAT_E_BAD_ARG, "magat_gso_from_rowbits: null pointer");
  MAGAT_REQUIRE(B >= 1 && N >= 1, MAGAT_E_BAD_ARG, "magat_gso_from_rowbits: B=%d N=%d", B, N);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const int W = (N + 31) / 32;
  const long blocks = (long)B * W * W;
  k_bits_transpose<<<cdiv(blocks, 8), 256, 0, st>>>(rowbits, N, W, blocks, colbits);
  int rc = check_launch("k_bits_transpose", st);
  if (rc) return rc;
  return launch_gso_stats(rowbits, colbits, B, N, W, stats, st);
}

extern "C" int magat_gso_scan_nonzero(const void* S, int s_dtype, int B, int N, uint32_t* rowbits, uint32_t* colbits,
                                      int32_t* stats, void* stream) {
  return gso_scan_impl(S, s_dtype, B, N, rowbits, colbits, stats, stream, true);
}

extern "C" int magat_gso_edge_values(const void* S, int s_dtype, const int32_t* nbr_out, int B, int N, int D, float* att,
                                     void* stream) {
  MAGAT_REQUIRE(S && nbr_out && att, MAGAT_E_BAD_ARG, "magat_gso_edge_values: null pointer");
  MAGAT_REQUIRE(B >= 1 && N >= 1 && D >= 1, MAGAT_E_BAD_ARG, "magat_gso_edge_values: B=%d N=%d D=%d", B, N, D);
  MAGAT_REQUIRE(s_dtype == MAGAT_DT_F32 || s_dtype == MAGAT_DT_F64, MAGAT_E_BAD_ARG,
                "magat_gso_edge_values: GSO dtype must be fp32 or fp64");
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const long total = (long)B * N * D;
  if (s_dtype == MAGAT_DT_F32)
    k_gso_edge_values<float><<<cdiv(total, 256), 256, 0, st>>>((const float*)S, nbr_out, total, N, D, att);
  else
    k_gso_edge_values<double><<<cdiv(total, 256), 256, 0, st>>>((const double*)S, nbr_out, total, N, D, att);
  return check_launch("k_gso_edge_values", st);
}

extern "C" int magat_gso_build_ell(const uint32_t* rowbits, const uint32_t* colbits, int B, int N,
                                   int D, int32_t* nbr_out, int32_t* nbr_in, int32_t* slot_in,
                                   int32_t* slot_out, void* stream) {
  MAGAT_REQUIRE(rowbits && colbits && nbr_out && nbr_in && slot_in, MAGAT_E_BAD_ARG,
                "magat_gso_build_ell: null pointer");
  MAGAT_REQUIRE(B >= 1 && N >= 1 && D >= 1, MAGAT_E_BAD_ARG, "magat_gso_build_ell: B=%d N=%d D=%d", B, N, D);
  const long rows = (long)B * N;
  const int W = (N + 31) / 32;
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  MAGAT_REQUIRE(rows * D < (1l << 40), MAGAT_E_UNSUPPORTED, "magat_gso_build_ell: B*N*D too large");
  auto al16 = [](const void* q) { return ((uintptr_t)q % 16) == 0; };
  const bool per_thread = W % 4 == 0 && D % 4 == 0 && D <= 32 && rows < (1l << 31) && al16(rowbits) && al16(colbits) &&
                          al16(nbr_out) && al16(nbr_in) && al16(slot_in) && (slot_out == nullptr || al16(slot_out));
  int rc;
  if (per_thread) {
    k_build_lists_t<<<dim3(cdiv(rows, 128), 2), 128, 0, st>>>(reinterpret_cast<const uint4*>(rowbits),
                                                              reinterpret_cast<const uint4*>(colbits), rows, W / 4, D,
                                                              nbr_out, nbr_in);
    if ((rc = check_launch("k_build_lists", st))) return rc;
    k_build_slots_t<<<dim3(cdiv(rows, 128), slot_out ? 2 : 1), 128, 0, st>>>(nbr_out, nbr_in, rows, N, D, slot_in,
                                                                            slot_out);
    return check_launch("k_build_slots", st);
  }
  k_build_lists<<<cdiv(rows, 8), 256, 0, st>>>(rowbits, colbits, rows, W, D, nbr_out, nbr_in);
  if ((rc = check_launch("k_build_lists", st))) return rc;
  k_build_slots<<<cdiv(rows * D, 256), 256, 0, st>>>(nbr_out, nbr_in, rows, N, D, slot_in, slot_out);
  return check_launch("k_build_slots", st);
}

extern "C" int magat_gso_from_positions(const void* pos, int pos_dtype, int B, int N, double comm_radius,
                                        uint32_t* rowbits, uint32_t* colbits, int32_t* stats, void* stream) {
  MAGAT_REQUIRE(pos && rowbits && colbits && stats, MAGAT_E_BAD_ARG, "magat_gso_from_positions: null pointer");
  MAGAT_REQUIRE(B >= 1 && N >= 1 && B <= 65535, MAGAT_E_BAD_ARG, "magat_gso_from_positions: B=%d N=%d", B, N);
  MAGAT_REQUIRE(pos_dtype == MAGAT_DT_F32 || pos_dtype == MAGAT_DT_F64, MAGAT_E_BAD_ARG,
                "magat_gso_from_positions: positions must be fp32 or fp64");
  MAGAT_REQUIRE(comm_radius == comm_radius, MAGAT_E_BAD_ARG, "magat_gso_from_positions: radius is NaN");
  const size_t smem = (size_t)N * 2 * sizeof(double);
  MAGAT_REQUIRE(smem <= 48 * 1024, MAGAT_E_UNSUPPORTED, "magat_gso_from_positions: N=%d exceeds 3072 agents", N);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const int W = (N + 31) / 32;
  // smallest double T with sqrt(T) >= R: then sqrt(d2) < R  <=>  d2 < T for every d2 >= 0 (sqrt is monotone)
  double thr2 = 0.0;
  if (comm_radius > 0.0) {
    thr2 = comm_radius * comm_radius;
    if (thr2 < INFINITY) {
      while (thr2 > 0.0 && sqrt(nextafter(thr2, 0.0)) >= comm_radius) thr2 = nextafter(thr2, 0.0);
      while (sqrt(thr2) < comm_radius) thr2 = nextafter(thr2, INFINITY);
    }
  }
  int rc;
  // cell list when the radius makes sense as a cell size (cells a hair wider than the radius: two agents closer than
  // the radius can then never sit two cells apart, whatever the rounding of the divisions)
  const bool cells = comm_radius > 0.0 && comm_radius < INFINITY;
  if (cells) {
    const size_t words = (size_t)B * N * W;
    cudaMemsetAsync(rowbits, 0, words * 4, st);
    cudaMemsetAsync(colbits, 0, words * 4, st);
    const size_t csm = (size_t)N * 2 * sizeof(double) + (size_t)N * 2 * sizeof(int) + (size_t)(2 * kCellMax + 1) * sizeof(int);
    const double cell = comm_radius * (1.0 + 1e-9);
    if (pos_dtype == MAGAT_DT_F32) {
      if ((rc = ensure_dyn_smem(KID_CELLS_F32, (const void*)k_gso_cells_from_positions<float>, 128 * 1024, "k_gso_cells_from_positions"))) return rc;
      k_gso_cells_from_positions<float><<<B, 256, csm, st>>>((const float*)pos, N, W, thr2, cell, rowbits, colbits);
    } else {
      if ((rc = ensure_dyn_smem(KID_CELLS_F64, (const void*)k_gso_cells_from_positions<double>, 128 * 1024, "k_gso_cells_from_positions"))) return rc;
      k_gso_cells_from_positions<double><<<B, 256, csm, st>>>((const double*)pos, N, W, thr2, cell, rowbits, colbits);
    }
    if ((rc = check_launch("k_gso_from_positions(cells)", st))) return rc;
    return launch_gso_stats(rowbits, colbits, B, N, W, stats, st);
  }
  dim3 grid(cdiv(N, 32), B);
  if (pos_dtype == MAGAT_DT_F32)
    k_gso_from_positions<float><<<grid, 256, smem, st>>>((const float*)pos, N, W, thr2, rowbits, colbits);
  else
    k_gso_from_positions<double><<<grid, 256, smem, st>>>((const double*)pos, N, W, thr2, rowbits, colbits);
  if ((rc = check_launch("k_gso_from_positions", st))) return rc;
  return launch_gso_stats(rowbits, colbits, B, N, W, stats, st);
}

// GAT_origin's edge test runs on S + I (graphML.py:1019): off the diagonal nothing changes, on it the bit becomes
// |float(S_ii) + 1| > 1e-9 (a -1 cancels the loop).  Fixes bit i of row / column word i in place.
template <typename T>
__global__ void __launch_bounds__(256) k_gso_self_loops(const T* __restrict__ S, long rows, int N, int W,
                                                        uint32_t* __restrict__ rowbits, uint32_t* __restrict__ colbits) {
  const long m = (long)blockIdx.x * blockDim.x + threadIdx.x;            // m = b * N + i
  if (m >= rows) return;
  const int i = (int)(m % N);
  const float d = (float)S[(size_t)m * N + i] + 1.0f;
  const bool e = fabsf(d) > 1e-9f;
  const uint32_t bit = 1u << (i & 31);
  const size_t w = (size_t)m * W + (i >> 5);
  rowbits[w] = e ? (rowbits[w] | bit) : (rowbits[w] & ~bit);
  colbits[w] = e ? (colbits[w] | bit) : (colbits[w] & ~bit);
}

extern "C" int magat_gso_self_loops(const void* S, int s_dtype, int B, int N, uint32_t* rowbits, uint32_t* colbits,
                                    int32_t* stats, void* stream) {
  MAGAT_REQUIRE(S && rowbits && colbits && stats, MAGAT_E_BAD_ARG, "magat_gso_self_loops: null pointer");
  MAGAT_REQUIRE(B >= 1 && N >= 1, MAGAT_E_BAD_ARG, "magat_gso_self_loops: B=%d N=%d", B, N);
  MAGAT_REQUIRE(s_dtype == MAGAT_DT_F32 || s_dtype == MAGAT_DT_F64, MAGAT_E_BAD_ARG,
                "magat_gso_self_loops: GSO dtype must be fp32 or fp64");
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const int W = (N + 31) / 32;
  const long rows = (long)B * N;
  if (s_dtype == MAGAT_DT_F32)
    k_gso_self_loops<float><<<cdiv(rows, 256), 256, 0, st>>>((const float*)S, rows, N, W, rowbits, colbits);
  else
    k_gso_self_loops<double><<<cdiv(rows, 256), 256, 0, st>>>((const double*)S, rows, N, W, rowbits, colbits);
  int rc = check_launch("k_gso_self_loops", st);
  if (rc) return rc;
  // the degree statistics of the scan are stale now: {0, 0, 0, 1} again, then recount
  static const int32_t init[4] = {0, 0, 0, 1};
  cudaMemcpyAsync(stats, init, sizeof(init), cudaMemcpyHostToDevice, st);
  return launch_gso_stats(rowbits, colbits, B, N, W, stats, st);
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
