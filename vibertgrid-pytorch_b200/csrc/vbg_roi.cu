// GridROIAlign forward over channels-last features.
//
// Replaces torchvision.ops.roi_align as called by reference model/grid_roi_align.py:37-41,81
// (output 7x7, spatial_scale 1/4, sampling_ratio=-1 -> adaptive ceil(roi/7) grid, aligned=False).
// The published algorithm (Mask R-CNN ROIAlign, torchvision legacy mode) is restated in
// oracle/oracle_ops.py::roi_align; bin geometry here follows it operation for operation so the
// integer sample-grid table is bit-exact and values agree to fp32 rounding.
//
// Mapping: one warp per (roi, bin).  Lanes stride the channel dimension with 128-bit loads, so
// each bilinear tap is a fully coalesced C*4-byte read (C=256 -> 1 KiB) instead of torchvision's
// one-thread-per-output NCHW gather.  Geometry is computed once per warp (uniform registers).
#include "vbg_common.cuh"
#include <stdlib.h>

namespace vbg {

template <int kVec>  // float4 vectors per lane (C = 128 * kVec)
__global__ void __launch_bounds__(256)
roi_align_kernel(const float* __restrict__ feat, int B, int Hf, int Wf, int C, const int32_t* __restrict__ boxes,
                 const int32_t* __restrict__ seg_off, int K, float scale, int P, float* __restrict__ out,
                 int32_t* __restrict__ sample_grid) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)K * P * P;
  if (warp >= total) return;
  const int k = (int)(warp / (P * P));
  const int bin = (int)(warp - (long long)k * P * P);
  const int ph = bin / P, pw = bin - ph * P;
  const int b = sample_of(seg_off, B, k);

  const int4 bx = __ldg(reinterpret_cast<const int4*>(boxes) + k);   // (x1, y1, x2, y2) int32 -> .float()
  const float sw = __fmul_rn((float)bx.x, scale), sh = __fmul_rn((float)bx.y, scale);
  const float ew = __fmul_rn((float)bx.z, scale), eh = __fmul_rn((float)bx.w, scale);
  const float rw = fmaxf(__fsub_rn(ew, sw), 1.0f), rh = fmaxf(__fsub_rn(eh, sh), 1.0f);
  const float bw = __fdiv_rn(rw, (float)P), bh = __fdiv_rn(rh, (float)P);
  const int gh = (int)ceilf(__fdiv_rn(rh, (float)P));
  const int gw = (int)ceilf(__fdiv_rn(rw, (float)P));
  if (sample_grid && bin == 0 && lane == 0) { sample_grid[2 * k] = gh; sample_grid[2 * k + 1] = gw; }
  const float count = (float)max(gh * gw, 1);

  const int C4 = C >> 2;
  const float4* f4 = reinterpret_cast<const float4*>(feat) + (size_t)b * Hf * Wf * C4;
  float4 acc[kVec];
#pragma unroll
  for (int v = 0; v < kVec; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int iy = 0; iy < gh; ++iy) {
    // y = roi_start_h + ph*bin_h + (iy + .5) * bin_h / grid_h      (same operation order as torchvision)
    float y = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)ph, bh)),
                        __fdiv_rn(__fmul_rn((float)iy + 0.5f, bh), (float)gh));
    for (int ix = 0; ix < gw; ++ix) {
      float x = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)pw, bw)),
                          __fdiv_rn(__fmul_rn((float)ix + 0.5f, bw), (float)gw));
      if (y < -1.0f || y > (float)Hf || x < -1.0f || x > (float)Wf) continue;
      float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
      int yl = (int)yy, xl = (int)xx, yh, xh;
      if (yl >= Hf - 1) { yh = yl = Hf - 1; yy = (float)yl; } else { yh = yl + 1; }
      if (xl >= Wf - 1) { xh = xl = Wf - 1; xx = (float)xl; } else { xh = xl + 1; }
      const float ly = yy - (float)yl, lx = xx - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
      const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
      const float4* p1 = f4 + ((size_t)yl * Wf + xl) * C4;
      const float4* p2 = f4 + ((size_t)yl * Wf + xh) * C4;
      const float4* p3 = f4 + ((size_t)yh * Wf + xl) * C4;
      const float4* p4 = f4 + ((size_t)yh * Wf + xh) * C4;
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        const int c = lane + 32 * v;
        if (c < C4) {
          float4 a = __ldg(p1 + c), bb = __ldg(p2 + c), cc = __ldg(p3 + c), d = __ldg(p4 + c);
          acc[v].x += w1 * a.x + w2 * bb.x + w3 * cc.x + w4 * d.x;
          acc[v].y += w1 * a.y + w2 * bb.y + w3 * cc.y + w4 * d.y;
          acc[v].z += w1 * a.z + w2 * bb.z + w3 * cc.z + w4 * d.z;
          acc[v].w += w1 * a.w + w2 * bb.w + w3 * cc.w + w4 * d.w;
        }
      }
    }
  }
  float4* o4 = reinterpret_cast<float4*>(out) + (size_t)warp * C4;   // [K,P,P,C]
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int c = lane + 32 * v;
    if (c < C4) {
      float4 r = acc[v];
      r.x = __fdiv_rn(r.x, count); r.y = __fdiv_rn(r.y, count);
      r.z = __fdiv_rn(r.z, count); r.w = __fdiv_rn(r.w, count);
      o4[c] = r;
    }
  }
}

// ------------------------------------------------------------------ separable form (the default for P == 7)
// Bilinear weights factor (w = wy * wx) and so do the skip / clamp rules (each depends on one coordinate only), hence
//   out[ph,pw,c] = 1/count * sum_yy Wy[ph][yy] * ( sum_xx Wx[pw][xx] * f[yy][xx][c] ),
// with Wy[ph][.] / Wx[pw][.] the per-bin sums of the samples' row / column weights.  One CTA per ROI builds the two small
// weight tables in shared memory (same coordinate arithmetic, operation for operation, as the direct kernel, so the
// sample-grid table stays bit-exact), then every thread owns 4 channels and streams the rows of its bin-row: each
// feature value is loaded once per (bin-row, bin-column) it contributes to instead of once per sample -- about half the
// L1 traffic of the direct form at typical line boxes, which is what bounds that kernel (ncu: 33% L1, 42% of HBM peak).
// Values differ from torchvision's summation order by fp32 re-association only (<= 1e-6 rel).
template <int P>
__global__ void __launch_bounds__(256)
roi_align_sep_kernel(const float* __restrict__ feat, int B, int Hf, int Wf, int C, const int32_t* __restrict__ boxes,
                     const int32_t* __restrict__ seg_off, float scale, float* __restrict__ out,
                     int32_t* __restrict__ sample_grid, int tab_stride) {
  extern __shared__ float tab[];                  // [2][P][tab_stride]: x weights, then y weights
  __shared__ int t_start[2][P], t_cnt[2][P];
  const int k = blockIdx.x, tid = threadIdx.x;
  const int b = sample_of(seg_off, B, k);
  const int4 bx = __ldg(reinterpret_cast<const int4*>(boxes) + k);
  const float sw = __fmul_rn((float)bx.x, scale), sh = __fmul_rn((float)bx.y, scale);
  const float ew = __fmul_rn((float)bx.z, scale), eh = __fmul_rn((float)bx.w, scale);
  const float rw = fmaxf(__fsub_rn(ew, sw), 1.0f), rh = fmaxf(__fsub_rn(eh, sh), 1.0f);
  const float bw = __fdiv_rn(rw, (float)P), bh = __fdiv_rn(rh, (float)P);
  const int gh = (int)ceilf(__fdiv_rn(rh, (float)P));
  const int gw = (int)ceilf(__fdiv_rn(rw, (float)P));
  if (sample_grid && tid == 0) { sample_grid[2 * k] = gh; sample_grid[2 * k + 1] = gw; }
  const float count = (float)max(gh * gw, 1);

  for (int i = tid; i < 2 * P * tab_stride; i += blockDim.x) tab[i] = 0.f;
  __syncthreads();
  if (tid < 2 * P) {
    const int axis = tid / P, pb = tid - axis * P;            // axis 0: x (columns), 1: y (rows)
    const int g = axis ? gh : gw, dim = axis ? Hf : Wf;
    const float start = axis ? sh : sw, bin = axis ? bh : bw;
    float* w = tab + (size_t)(axis * P + pb) * tab_stride;
    int base = 0, cnt = 0;
    for (int i = 0; i < g; ++i) {
      float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)pb, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)g));
      if (c < -1.0f || c > (float)dim) continue;
      c = fmaxf(c, 0.f);
      int lo = (int)c, hi;
      if (lo >= dim - 1) { hi = lo = dim - 1; c = (float)lo; } else { hi = lo + 1; }
      const float l = c - (float)lo, h = 1.f - l;
      if (cnt == 0) base = lo;
      w[lo - base] += h;
      w[hi - base] += l;
      cnt = hi - base + 1;
    }
    t_start[axis][pb] = base;
    t_cnt[axis][pb] = cnt;
  }
  __syncthreads();

  const int C4 = C >> 2, c4 = tid % C4, grp = tid / C4, G = blockDim.x / C4;
  const float4* f4 = reinterpret_cast<const float4*>(feat) + (size_t)b * Hf * Wf * C4 + c4;
  float4* o4 = reinterpret_cast<float4*>(out) + (size_t)k * P * P * C4 + c4;
  for (int ph = grp; ph < P; ph += G) {
    float4 acc[P];
#pragma unroll
    for (int pw = 0; pw < P; ++pw) acc[pw] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* wyp = tab + (size_t)(P + ph) * tab_stride;
    const int y0 = t_start[1][ph], ny = t_cnt[1][ph];
    for (int j = 0; j < ny; ++j) {
      const float wy = wyp[j];
      const float4* rowp = f4 + (size_t)(y0 + j) * Wf * C4;
#pragma unroll
      for (int pw = 0; pw < P; ++pw) {
        const float* wxp = tab + (size_t)pw * tab_stride;
        const int x0 = t_start[0][pw], nx = t_cnt[0][pw];
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < nx; ++i) {
          const float w = wxp[i];
          const float4 v = __ldg(rowp + (size_t)(x0 + i) * C4);
          t.x = fmaf(w, v.x, t.x); t.y = fmaf(w, v.y, t.y); t.z = fmaf(w, v.z, t.z); t.w = fmaf(w, v.w, t.w);
        }
        acc[pw].x = fmaf(wy, t.x, acc[pw].x); acc[pw].y = fmaf(wy, t.y, acc[pw].y);
        acc[pw].z = fmaf(wy, t.z, acc[pw].z); acc[pw].w = fmaf(wy, t.w, acc[pw].w);
      }
    }
#pragma unroll
    for (int pw = 0; pw < P; ++pw) {
      float4 r = acc[pw];
      r.x = __fdiv_rn(r.x, count); r.y = __fdiv_rn(r.y, count); r.z = __fdiv_rn(r.z, count); r.w = __fdiv_rn(r.w, count);
      o4[(size_t)(ph * P + pw) * C4] = r;
    }
  }
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_roi_align_fwd(const float* feat, int B, int Hf, int Wf, int C, const int32_t* boxes,
                                 const int32_t* seg_off, int K, float spatial_scale, int P, float* out,
                                 int32_t* sample_grid, vbg_stream_t stream) {
  VBG_REQUIRE(feat && boxes && seg_off && out && B > 0 && Hf > 0 && Wf > 0 && P > 0 && K >= 0,
              "vbg_roi_align_fwd: bad arguments");
  VBG_REQUIRE(C > 0 && C % 4 == 0 && C <= 1024 && aligned16(feat) && aligned16(out) && aligned16(boxes),
              "vbg_roi_align_fwd: C %% 4 == 0, C <= 1024 and 16B alignment required (C=%d)", C);
  if (K == 0) return VBG_OK;
  {
    const int C4 = C / 4;
    static const bool direct = [] { const char* e = getenv("VBG_ROI_DIRECT"); return e && e[0] == '1'; }();
    const int tab_stride = (Hf > Wf ? Hf : Wf) + 1;
    const size_t smem = (size_t)2 * 7 * tab_stride * sizeof(float);
    if (!direct && P == 7 && C4 <= 256 && 256 % C4 == 0 && smem <= 48 * 1024) {
      roi_align_sep_kernel<7><<<K, 256, smem, as_stream(stream)>>>(feat, B, Hf, Wf, C, boxes, seg_off, spatial_scale, out,
                                                                  sample_grid, tab_stride);
      return check_launch("vbg_roi_align_fwd");
    }
  }
  long long warps = (long long)K * P * P;
  int blocks = (int)((warps + 7) / 8);
  cudaStream_t s = as_stream(stream);
  int vec = (C / 4 + 31) / 32;
#define LAUNCH(V) roi_align_kernel<V><<<blocks, 256, 0, s>>>(feat, B, Hf, Wf, C, boxes, seg_off, K, spatial_scale, P, out, sample_grid)
  if (vec <= 1) LAUNCH(1);
  else if (vec == 2) LAUNCH(2);
  else if (vec <= 4) LAUNCH(4);
  else LAUNCH(8);
#undef LAUNCH
  return check_launch("vbg_roi_align_fwd");
}
