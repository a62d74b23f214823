// Shared helpers for libvbg_sm100a (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
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

// erf-GELU for the tensor-core epilogues: Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, i.e. <= 5e-7 absolute on the
// output -- far inside the tensor-core mode's own error), branch-free: rcp.approx + ex2.approx + ~10 FMA-pipe
// instructions per element instead of erff's two divergent branches (~26).  The exact-fp32 CUDA-core path keeps erff.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float q = p * t * e;                                   // 1 - erf(|x| / sqrt 2)
  const float h = 0.5f * x;
  return x >= 0.f ? fmaf(-h, q, x) : h * q;                    // x >= 0: 0.5x(2 - q);  x < 0: 0.5x q
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == VBG_ACT_RELU) return fmaxf(v, 0.0f);
  if (act == VBG_ACT_GELU) return gelu_erf(v);
  return v;
}

// ------------------------------------------------------------------ activation storage formats
// An activation is either fp32 (plane == 0: `base` is float*) or a pair of bf16 planes (plane > 0: `base` is the hi
// plane = bf16_rn(x), the lo plane = bf16_rn(x - hi) starts `plane` ELEMENTS later): hi + lo carries 16 mantissa bits
// and is the operand format the bf16x3 tensor-core kernels consume without conversion.  i4 indexes groups of 4 elements.
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// (a, b) -> packed bf16 pairs hi = bf16_rn(.), lo = bf16_rn(. - hi): two packed F2FP conversions, no scalar F2F
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);                 // .x = a in the low half
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  split2(v.x, v.y, hi.x, lo.x);
  split2(v.z, v.w, hi.y, lo.y);
}
__device__ __forceinline__ float bf16lo_f(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi_f(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float4 merge4(const uint2 hi, const uint2 lo) {
  return make_float4(bf16lo_f(hi.x) + bf16lo_f(lo.x), bf16hi_f(hi.x) + bf16hi_f(lo.x),
                     bf16lo_f(hi.y) + bf16lo_f(lo.y), bf16hi_f(hi.y) + bf16hi_f(lo.y));
}
__device__ __forceinline__ float4 ld4_fmt(const void* __restrict__ base, long long plane, size_t i4) {
  if (plane == 0) return __ldg(reinterpret_cast<const float4*>(base) + i4);
  const uint2* h = reinterpret_cast<const uint2*>(base);
  const uint2* l = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + plane);
  return merge4(__ldg(h + i4), __ldg(l + i4));
}
__device__ __forceinline__ void st4_fmt(void* __restrict__ base, long long plane, size_t i4, const float4 v) {
  if (plane == 0) { reinterpret_cast<float4*>(base)[i4] = v; return; }
  uint2 hi, lo;
  split4(v, hi, lo);
  reinterpret_cast<uint2*>(base)[i4] = hi;
  reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + plane)[i4] = lo;
}
inline bool fmt_ok(const void* p, long long plane) { return p && plane >= 0 && (plane % 8) == 0 && aligned16(p); }

// Python slice index normalisation for one bound: i<0 -> i+dim, then clamp to [0,dim]
__device__ __forceinline__ int py_slice_bound(int i, int dim) {
  if (i < 0) { i += dim; if (i < 0) i = 0; }
  return i > dim ? dim : i;
}

}  // namespace vbg
