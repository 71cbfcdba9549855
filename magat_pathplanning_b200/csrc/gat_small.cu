// Small-graph forward: the whole layer for ONE planning instance in ONE CTA, one launch per call.
//
// The reference's simulator loop calls the layer with B = 1 and N = 10..64 agents at every time step under
// torch.no_grad() (agents/decentralplannerlocal_OnlineExpert_GAT.py:1039-1044); there the multi-kernel path is
// pure launch latency (9 launches, ~0.19 ms).  This kernel reads S[b] and x[b] once into shared memory and does
// mask -> scores -> row softmax -> K taps -> per-(head,tap) projection -> bias / ReLU / concat or head mean with
// fp32 FMAs (same math as gat_fwd.cu), writing y and, optionally, the dense attention aij[b][p][i][j] that
// returnAttentionGSO() needs.  Inference only (nothing is saved for backward).
#include "common.cuh"

namespace magat {

namespace {

constexpr int SMALL_MAX_N = 64;
constexpr int SMALL_THREADS = 256;

struct SmallParams {
  int N, G, F, K, P, mode, concat, relu, s_dtype;
  const void* S;
  const float* x; long x_sb, x_sn;
  const float* weight; const float* mixer; const float* weight_bias; const float* filterWeight; const float* bias;
  float* y; long y_sb, y_sn, y_sc;
  float* aij;          // [B][P][N][N] or null
};

__global__ void __launch_bounds__(SMALL_THREADS) k_gat_small_fwd(const SmallParams p) {
  extern __shared__ float sm[];
  const int N = p.N, G = p.G, F = p.F, K = p.K, P = p.P;
  const int b = blockIdx.x, tid = threadIdx.x;
  float* xs = sm;                          // [N][G]
  float* R = xs + N * G;                   // [N][G] (KeyQuery) / [N][2] (GAT_modified) scores projection
  const int r_floats = (max(N * G, 2 * N + 2 * G + 2) + 3) & ~3;      // regions stay 16 B aligned when G % 4 == 0
  float* A = R + r_floats;                 // [N][N] attention of the current head
  float* U = A + ((N * N + 3) & ~3);       // [K-1][N][G] taps of the current head
  float* Ys = U + (K > 1 ? (K - 1) : 0) * N * G;   // [N][F] head-mean accumulator (mean mode only)
  unsigned char* M = reinterpret_cast<unsigned char*>(Ys + (p.concat ? 0 : N * F));   // [N][N] edge mask

  // ---- x[b] and the edge mask of S[b] --------------------------------------------------------------
  for (int e = tid; e < N * G; e += SMALL_THREADS) {
    const int n = e / G, g = e - n * G;
    xs[e] = p.x[(long)b * p.x_sb + (long)n * p.x_sn + g];
  }
  for (int e = tid; e < N * N; e += SMALL_THREADS) {
    bool edge;
    if (p.s_dtype == MAGAT_DT_F32) edge = fabsf(static_cast<const float*>(p.S)[(size_t)b * N * N + e]) > 1e-9f;
    else edge = fabs(static_cast<const double*>(p.S)[(size_t)b * N * N + e]) > 1e-9;
    M[e] = edge ? 1 : 0;
  }
  if (!p.concat)
    for (int e = tid; e < N * F; e += SMALL_THREADS) Ys[e] = 0.f;
  __syncthreads();

  for (int h = 0; h < P; ++h) {
    // ---- scores projection ---------------------------------------------------------------------
    if (p.mode == MAGAT_MODE_KEYQUERY) {
      // R[i][g'] = sum_g x[i][g] W[h][g][g']
      const float* W = p.weight + (size_t)h * G * G;
      if (SMALL_THREADS % G == 0 && N * G <= 32 * SMALL_THREADS) {
        // a thread owns one output column g' for every (256 / G)-th node: each W element it loads feeds all of them
        const int gp = tid % G, i0 = tid / G, istep = SMALL_THREADS / G;
        float acc[32];
#pragma unroll
        for (int o = 0; o < 32; ++o) acc[o] = 0.f;
#pragma unroll 8
        for (int g = 0; g < G; ++g) {
          const float w = __ldg(W + (size_t)g * G + gp);
#pragma unroll
          for (int o = 0; o < 32; ++o) {
            const int i = i0 + o * istep;
            if (i < N) acc[o] = fmaf(xs[i * G + g], w, acc[o]);
          }
        }
#pragma unroll
        for (int o = 0; o < 32; ++o) {
          const int i = i0 + o * istep;
          if (i < N) R[i * G + gp] = acc[o];
        }
      } else
      for (int e = tid; e < N * G; e += SMALL_THREADS) {
        const int i = e / G, gp = e - i * G;
        float acc = 0.f;
#pragma unroll 16
        for (int g = 0; g < G; ++g) acc = fmaf(xs[i * G + g], __ldg(W + (size_t)g * G + gp), acc);   // 16 loads in flight
        R[e] = acc;
      }
    } else {
      // R[i][t] = a_t . (W x_i + wb) = (W^T a_t) . x_i + a_t . wb;  t = 0: a1 (receiver term), t = 1: a2 (sender)
      const float* W = p.weight + (size_t)h * F * G;
      const float* a = p.mixer + (size_t)h * 2 * F;
      const float* wb = p.weight_bias + (size_t)h * F;
      float* cvec = R + 2 * N;               // [2][G]
      float* dvec = cvec + 2 * G;            // [2]
      for (int e = tid; e < 2 * G + 2; e += SMALL_THREADS) {
        float acc = 0.f;
        if (e < 2 * G) {
          const int t = e / G, g = e - t * G;
#pragma unroll 16
          for (int f = 0; f < F; ++f) acc = fmaf(__ldg(a + t * F + f), __ldg(W + (size_t)f * G + g), acc);
        } else {
          const int t = e - 2 * G;
          for (int f = 0; f < F; ++f) acc = fmaf(__ldg(a + t * F + f), __ldg(wb + f), acc);
        }
        cvec[e] = acc;
      }
      __syncthreads();
      for (int e = tid; e < N * 2; e += SMALL_THREADS) {
        const int i = e >> 1, t = e & 1;
        float acc = dvec[t];
        for (int g = 0; g < G; ++g) acc = fmaf(cvec[t * G + g], xs[i * G + g], acc);
        R[e] = acc;
      }
    }
    __syncthreads();
    // ---- e[i][j] on the mask, row softmax ----------------------------------------------------------
    for (int e = tid; e < N * N; e += SMALL_THREADS) {
      const int i = e / N, j = e - i * N;
      float s = -INFINITY;
      if (M[e]) {
        if (p.mode == MAGAT_MODE_KEYQUERY) {
          s = 0.f;
          for (int g = 0; g < G; ++g) s = fmaf(R[i * G + g], xs[j * G + g], s);
        } else {
          s = R[i * 2 + 1] + R[j * 2 + 0];
          s = s > 0.f ? s : kLeaky * s;
        }
      }
      A[e] = s;
    }
    __syncthreads();
    for (int i = tid; i < N; i += SMALL_THREADS) {
      float mx = -INFINITY;
      for (int j = 0; j < N; ++j) mx = fmaxf(mx, A[i * N + j]);
      float sum = 0.f;
      for (int j = 0; j < N; ++j) {
        const float ex = M[i * N + j] ? expf(A[i * N + j] - mx) : 0.f;
        A[i * N + j] = ex;
        sum += ex;
      }
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      for (int j = 0; j < N; ++j) A[i * N + j] *= inv;
    }
    __syncthreads();
    if (p.aij != nullptr)
      for (int e = tid; e < N * N; e += SMALL_THREADS) p.aij[(((size_t)b * P + h) * N) * N + e] = A[e];
    // ---- taps u_k[j] = sum_i A[i][j] u_{k-1}[i] -----------------------------------------------------
    for (int k = 1; k < K; ++k) {
      const float* src = k == 1 ? xs : U + (k - 2) * N * G;
      float* dst = U + (k - 1) * N * G;
      for (int e = tid; e < N * G; e += SMALL_THREADS) {
        const int j = e / G, g = e - j * G;
        float acc = 0.f;
        for (int i = 0; i < N; ++i) acc = fmaf(A[i * N + j], src[i * G + g], acc);
        dst[e] = acc;
      }
      __syncthreads();
    }
    // ---- Y_h[n][f] = sum_{k,g} H[h][f][k][g] u_k[n][g] ---------------------------------------------
    const float* H = p.filterWeight + (size_t)h * F * K * G;
    const bool vec_ok = (((uintptr_t)p.filterWeight) & 15) == 0;
    if (SMALL_THREADS % F == 0 && N * F <= 32 * SMALL_THREADS && (G & 3) == 0 && vec_ok) {
      // a thread owns output feature f for every (256 / F)-th node: one float4 of H feeds all of its nodes
      const int f = tid % F, n0 = tid / F, nstep = SMALL_THREADS / F;
      const float4* h4 = reinterpret_cast<const float4*>(H + (size_t)f * K * G);
      float acc[32];
#pragma unroll
      for (int o = 0; o < 32; ++o) acc[o] = 0.f;
      for (int k = 0; k < K; ++k) {
        const float* ub = k == 0 ? xs : U + (k - 1) * N * G;
#pragma unroll 4
        for (int g = 0; g < G / 4; ++g) {
          const float4 hv = __ldg(h4 + k * (G / 4) + g);
#pragma unroll
          for (int o = 0; o < 32; ++o) {
            const int n = n0 + o * nstep;
            if (n < N) {
              const float4 uv = *reinterpret_cast<const float4*>(ub + n * G + 4 * g);
              acc[o] = fmaf(hv.x, uv.x, fmaf(hv.y, uv.y, fmaf(hv.z, uv.z, fmaf(hv.w, uv.w, acc[o]))));
            }
          }
        }
      }
#pragma unroll
      for (int o = 0; o < 32; ++o) {
        const int n = n0 + o * nstep;
        if (n >= N) continue;
        if (p.concat) {
          float v = acc[o] + (p.bias ? __ldg(p.bias + f) : 0.f);
          if (p.relu) v = fmaxf(v, 0.f);
          p.y[(long)b * p.y_sb + (long)n * p.y_sn + ((long)h * F + f) * p.y_sc] = v;
        } else {
          Ys[n * F + f] += acc[o];
        }
      }
    } else
    for (int e = tid; e < N * F; e += SMALL_THREADS) {
      const int f = e % F, n = e / F;          // consecutive threads: consecutive f, same node
      const float* hrow = H + (size_t)f * K * G;
      float acc = 0.f;
      for (int k = 0; k < K; ++k) {
        const float* u = k == 0 ? xs + n * G : U + ((k - 1) * N + n) * G;
        if ((G & 3) == 0 && vec_ok) {
          const float4* h4 = reinterpret_cast<const float4*>(hrow + k * G);
          const float4* u4 = reinterpret_cast<const float4*>(u);
#pragma unroll 8
          for (int g = 0; g < G / 4; ++g) {
            const float4 hv = __ldg(h4 + g), uv = u4[g];
            acc = fmaf(hv.x, uv.x, fmaf(hv.y, uv.y, fmaf(hv.z, uv.z, fmaf(hv.w, uv.w, acc))));
          }
        } else {
          for (int g = 0; g < G; ++g) acc = fmaf(__ldg(hrow + k * G + g), u[g], acc);
        }
      }
      if (p.concat) {
        float o = acc + (p.bias ? __ldg(p.bias + f) : 0.f);
        if (p.relu) o = fmaxf(o, 0.f);
        p.y[(long)b * p.y_sb + (long)n * p.y_sn + ((long)h * F + f) * p.y_sc] = o;
      } else {
        Ys[n * F + f] += acc;                  // this thread owns (n, f) for every head
      }
    }
    __syncthreads();
  }
  if (!p.concat) {
    for (int e = tid; e < N * F; e += SMALL_THREADS) {
      const int f = e % F, n = e / F;
      float o = Ys[e] / (float)P + (p.bias ? __ldg(p.bias + f) : 0.f);
      if (p.relu) o = fmaxf(o, 0.f);
      p.y[(long)b * p.y_sb + (long)n * p.y_sn + (long)f * p.y_sc] = o;
    }
  }
}

size_t small_smem_bytes(int N, int G, int F, int K, int concat) {
  size_t r_floats = (size_t)N * G > (size_t)(2 * N + 2 * G + 2) ? (size_t)N * G : (size_t)(2 * N + 2 * G + 2);
  r_floats = (r_floats + 3) & ~(size_t)3;
  size_t fl = (size_t)N * G + r_floats + (((size_t)N * N + 3) & ~(size_t)3) + (size_t)(K > 1 ? K - 1 : 0) * N * G +
              (concat ? 0 : (size_t)N * F);
  return fl * 4 + (size_t)N * N + 16;
}

}  // namespace
}  // namespace magat

using namespace magat;

extern "C" int magat_gat_small_supported(int N, int G, int F, int K, int P, int concat) {
  if (N < 1 || N > SMALL_MAX_N || G < 1 || F < 1 || K < 1 || P < 1) return 0;
  return small_smem_bytes(N, G, F, K, concat) <= 200 * 1024 ? 1 : 0;
}

extern "C" int magat_gat_forward_small(const void* S, int s_dtype, const float* x, int64_t x_sb, int64_t x_sn,
                                       const float* weight, const float* mixer, const float* weight_bias,
                                       const float* filterWeight, const float* bias, float* y, int64_t y_sb,
                                       int64_t y_sn, int64_t y_sc, float* aij_or_null, int B, int N, int G, int F,
                                       int K, int P, int mode, int concat, int relu, void* stream) {
  MAGAT_REQUIRE(S && x && weight && filterWeight && y, MAGAT_E_BAD_ARG, "magat_gat_forward_small: null pointer");
  MAGAT_REQUIRE(B >= 1 && B <= 2147483647, MAGAT_E_BAD_ARG, "magat_gat_forward_small: B=%d", B);
  MAGAT_REQUIRE(mode == MAGAT_MODE_KEYQUERY || mode == MAGAT_MODE_GAT_MODIFIED, MAGAT_E_BAD_ARG, "unknown mode %d", mode);
  MAGAT_REQUIRE(mode != MAGAT_MODE_KEYQUERY || F == G, MAGAT_E_UNSUPPORTED,
                "KeyQuery needs F == G (got F=%d G=%d; graphML.py:1728,1765)", F, G);
  MAGAT_REQUIRE(mode != MAGAT_MODE_GAT_MODIFIED || (mixer && weight_bias), MAGAT_E_BAD_ARG,
                "magat_gat_forward_small: GAT_modified needs mixer and weight_bias");
  MAGAT_REQUIRE(s_dtype == MAGAT_DT_F32 || s_dtype == MAGAT_DT_F64, MAGAT_E_BAD_ARG, "GSO dtype must be fp32 or fp64");
  MAGAT_REQUIRE(magat_gat_small_supported(N, G, F, K, P, concat), MAGAT_E_UNSUPPORTED,
                "magat_gat_forward_small: N=%d G=%d F=%d K=%d outside the single-CTA limits", N, G, F, K);
  cudaStream_t st = (cudaStream_t)stream;
  prof_begin(st);
  const size_t smem = small_smem_bytes(N, G, F, K, concat);
  if (smem > 48 * 1024) {
    int rc0 = ensure_dyn_smem(KID_SMALL, (const void*)k_gat_small_fwd, 200 * 1024, "k_gat_small_fwd");
    if (rc0) return rc0;
  }
  SmallParams sp{N, G, F, K, P, mode, concat, relu, s_dtype, S, x, x_sb, x_sn, weight, mixer, weight_bias,
                 filterWeight, bias, y, y_sb, y_sn, y_sc, aij_or_null};
  k_gat_small_fwd<<<B, SMALL_THREADS, smem, st>>>(sp);
  return check_launch("k_gat_small_fwd", st);
}
