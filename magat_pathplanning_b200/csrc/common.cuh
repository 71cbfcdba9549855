// Shared helpers for the magat_gat library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/magat_gat.h"

namespace magat {

void set_error(const char* fmt, ...);
int check_launch(const char* what, cudaStream_t st);
void prof_begin(cudaStream_t st);
int device_sm_count();                       // SMs of the current device (cached per device), 0 without a device
// once per (kernel id, device): raise the dynamic shared memory limit of `func`
enum { KID_TC_GEMM = 0, KID_TAP_TC4, KID_TAP_TC8, KID_WGRAD, KID_SMALL, KID_TAP_TC2, KID_TAP_TC2M, KID_WGRAD2, KID_FUSED, KID_TAP_TC2B, KID_TAP_TC2H, KID_CELLS_F32, KID_CELLS_F64,
       KID_SCAN_BASE /* + 48 dtype / R / KK / predicate variants */ = 16, KID_MAX = 72 };
int ensure_dyn_smem(int kernel_id, const void* func, size_t bytes, const char* name);

// fused linear head of the K-tap projection (magat_gat_forward_actions)
struct HeadArgs {
  const float* w;      // [A][P*F]
  const float* b;      // [A] or null
  int A;
  float* partial;      // [P][rows][8] scratch
  float* logits;       // [rows][A]
  int32_t* actions;    // [rows] argmax, or null
};

#define MAGAT_REQUIRE(cond, code, ...)        \
  do {                                        \
    if (!(cond)) {                            \
      ::magat::set_error(__VA_ARGS__);        \
      return (code);                          \
    }                                         \
  } while (0)

constexpr float kZeroTol = 1e-9f;   // graphML.py:45 zeroTolerance
constexpr float kLeaky = 0.2f;      // graphML.py:713 negative_slope

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Sum NV per-lane values over the warp at once: every step halves the number of live values and doubles the lanes
// each has absorbed (NV - 1 + 5 - log2 NV shuffles instead of 5 NV).  The lane's result is the sum with index
// lane >> (5 - log2 NV).
template <int NV>
__device__ __forceinline__ float warp_multi_sum(float (&v)[NV], int lane) {
  int off = 16;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1, off >>= 1) {
    const bool hi = (lane & off) != 0;
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
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_excl_scan_i(int v, int lane) {
  int s = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, s, o);
    if (lane >= o) s += t;
  }
  return s - v;
}

// batch index of a flat node row when the host has checked B * N < 2^31: a 32-bit division (about a fifth of
// the instructions of the emulated 64-bit one, which used to be ~15 % of the warp-per-node kernels)
__device__ __forceinline__ long batch_of32(long row, int N) { return (long)((unsigned)row / (unsigned)N); }

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace magat
