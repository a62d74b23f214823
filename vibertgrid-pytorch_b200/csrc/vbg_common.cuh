// Shared helpers for libvbg_sm100a (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/vbg.h"

namespace vbg {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return VBG_ECUDA;
  }
  return VBG_OK;
}

#define VBG_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      vbg::set_error(__VA_ARGS__);        \
      return VBG_EINVAL;                  \
    }                                     \
  } while (0)

inline cudaStream_t as_stream(vbg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// sample index of global segment k: largest b with seg_off[b] <= k   (B is small)
__device__ __forceinline__ int sample_of(const int32_t* __restrict__ seg_off, int B, int k) {
  int lo = 0, hi = B - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(seg_off + mid) <= k) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == VBG_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == VBG_ACT_GELU) return gelu_erf(v);
  return v;
}

// Python slice index normalisation for one bound: i<0 -> i+dim, then clamp to [0,dim]
__device__ __forceinline__ int py_slice_bound(int i, int dim) {
  if (i < 0) { i += dim; if (i < 0) i = 0; }
  return i > dim ? dim : i;
}

}  // namespace vbg
