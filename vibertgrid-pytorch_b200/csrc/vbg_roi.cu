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

long long* tc_debug_timeline();
static long long* roi_debug_timeline() { return tc_debug_timeline(); }

template <int kVec>  // float4 vectors per lane (C = 128 * kVec)
__global__ void __launch_bounds__(256)
roi_align_kernel(const void* __restrict__ feat, long long feat_plane, int B, int Hf, int Wf, int C,
                 const int32_t* __restrict__ boxes, const int32_t* __restrict__ seg_off, int K, float scale, int P,
                 void* __restrict__ out, long long out_plane, int32_t* __restrict__ sample_grid) {
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
  const size_t f0 = (size_t)b * Hf * Wf * C4;
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
      const size_t p1 = f0 + ((size_t)yl * Wf + xl) * C4;
      const size_t p2 = f0 + ((size_t)yl * Wf + xh) * C4;
      const size_t p3 = f0 + ((size_t)yh * Wf + xl) * C4;
      const size_t p4 = f0 + ((size_t)yh * Wf + xh) * C4;
#pragma unroll
      for (int v = 0; v < kVec; ++v) {
        const int c = lane + 32 * v;
        if (c < C4) {
          float4 a = ld4_fmt(feat, feat_plane, p1 + c), bb = ld4_fmt(feat, feat_plane, p2 + c),
                 cc = ld4_fmt(feat, feat_plane, p3 + c), d = ld4_fmt(feat, feat_plane, p4 + c);
          acc[v].x += w1 * a.x + w2 * bb.x + w3 * cc.x + w4 * d.x;
          acc[v].y += w1 * a.y + w2 * bb.y + w3 * cc.y + w4 * d.y;
          acc[v].z += w1 * a.z + w2 * bb.z + w3 * cc.z + w4 * d.z;
          acc[v].w += w1 * a.w + w2 * bb.w + w3 * cc.w + w4 * d.w;
        }
      }
    }
  }
  const size_t o4 = (size_t)warp * C4;                                 // [K,P,P,C]
#pragma unroll
  for (int v = 0; v < kVec; ++v) {
    const int c = lane + 32 * v;
    if (c < C4) {
      float4 r = acc[v];
      r.x = __fdiv_rn(r.x, count); r.y = __fdiv_rn(r.y, count);
      r.z = __fdiv_rn(r.z, count); r.w = __fdiv_rn(r.w, count);
      st4_fmt(out, out_plane, o4 + c, r);
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

// ------------------------------------------------------------------ windowed form (the default)
// One CTA per (ROI, 64-channel chunk).  The feature window the ROI's samples can touch -- rows [y_lo, y_hi] x columns
// [x_lo, x_hi] -- is staged ONCE into shared memory with coalesced 256-byte reads (either storage format, merged to fp32
// on the way in), then every (bin, channel-quad) item accumulates from shared memory with per-bin SEPARABLE weight tables
// (see below: each window pixel is read once per bin it contributes to, not once per sample tap).  HBM / L2 -> SM traffic
// per ROI drops from (samples x 4 taps x C) per bin -- ~24 KB per 1 KB of output at line-sized boxes -- to the window
// itself.  ROIs whose window does not fit `win_floats` fall back to per-sample global taps inside the same kernel.
constexpr int kRoiCh = 64, kRoiCq = kRoiCh / 4, kRoiP = 8, kRoiSpan = 32;

__global__ void __launch_bounds__(256)
roi_align_win_kernel(const void* __restrict__ feat, long long feat_plane, int B, int Hf, int Wf, int C,
                     const int32_t* __restrict__ boxes, const int32_t* __restrict__ seg_off, float scale, int P,
                     void* __restrict__ out, long long out_plane, int32_t* __restrict__ sample_grid, int win_floats,
                     long long* __restrict__ dbg) {
#define ROI_STAMP(slot) do { if (dbg && blockIdx.x == 2000 && threadIdx.x == 0) dbg[slot] = clock64(); } while (0)
  extern __shared__ __align__(16) float win[];
  ROI_STAMP(0);
  // the channel chunks of one ROI are neighbouring CTAs: they run together, so each 512-byte pixel row is fetched once
  const int nchunk = C / kRoiCh, k = blockIdx.x / nchunk, chunk = blockIdx.x - k * nchunk, tid = threadIdx.x;
  int b;
  if (B <= 31) {       // one load latency instead of a dependent binary search: lane i holds seg_off[i]; b = #(seg_off[1..B-1] <= k)
    const int lane = tid & 31;
    const int so = (lane >= 1 && lane < B) ? __ldg(seg_off + lane) : 0x7fffffff;
    b = __popc(__ballot_sync(0xffffffffu, so <= k));
  } else {
    b = sample_of(seg_off, B, k);
  }
  const int4 bx = __ldg(reinterpret_cast<const int4*>(boxes) + k);
  const float sw = __fmul_rn((float)bx.x, scale), sh = __fmul_rn((float)bx.y, scale);
  const float ew = __fmul_rn((float)bx.z, scale), eh = __fmul_rn((float)bx.w, scale);
  const float rw = fmaxf(__fsub_rn(ew, sw), 1.0f), rh = fmaxf(__fsub_rn(eh, sh), 1.0f);
  const float bw = __fdiv_rn(rw, (float)P), bh = __fdiv_rn(rh, (float)P);
  const int gh = (int)ceilf(__fdiv_rn(rh, (float)P));
  const int gw = (int)ceilf(__fdiv_rn(rw, (float)P));
  if (sample_grid && chunk == 0 && tid == 0) { sample_grid[2 * k] = gh; sample_grid[2 * k + 1] = gw; }
  const float count = (float)max(gh * gw, 1);

  // window bounds from the first / last sample coordinate of each axis (coordinates are monotone in (bin, sample))
  const float y_first = __fadd_rn(sh, __fdiv_rn(__fmul_rn(0.5f, bh), (float)gh));
  const float y_last = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)(P - 1), bh)), __fdiv_rn(__fmul_rn((float)gh - 0.5f, bh), (float)gh));
  const float x_first = __fadd_rn(sw, __fdiv_rn(__fmul_rn(0.5f, bw), (float)gw));
  const float x_last = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)(P - 1), bw)), __fdiv_rn(__fmul_rn((float)gw - 0.5f, bw), (float)gw));
  const int y_lo = min(max((int)fmaxf(y_first, 0.f), 0), Hf - 1), y_hi = min(max((int)fmaxf(y_last, 0.f) + 1, y_lo), Hf - 1);
  const int x_lo = min(max((int)fmaxf(x_first, 0.f), 0), Wf - 1), x_hi = min(max((int)fmaxf(x_last, 0.f) + 1, x_lo), Wf - 1);
  const int rows = y_hi - y_lo + 1, cols = x_hi - x_lo + 1;
  const bool staged = (long long)rows * cols * kRoiCh <= (long long)win_floats;

  const int C4 = C >> 2;
  const size_t f0 = (size_t)b * Hf * Wf * C4 + (size_t)chunk * kRoiCq;
  // Separable form: bilinear weights factor (w = wy * wx) and so do the skip / clamp rules, hence
  //   out[ph,pw,c] = 1/count * sum_j Wy[ph][j] * ( sum_i Wx[pw][i] * f[y0+j][x0+i][c] )
  // with Wy[ph][.] / Wx[pw][.] the per-bin sums of the samples' row / column weights.  The tables are built once per ROI
  // (same coordinate arithmetic, operation for operation, as the direct kernel: the sample-grid table stays bit-exact);
  // every feature value is then read once per bin it contributes to, not once per sample tap.  Values differ from
  // torchvision's summation order by fp32 re-association only (<= 1e-6 rel).
  __shared__ float wtab[2][kRoiP][kRoiSpan];
  __shared__ int t_start[2][kRoiP], t_cnt[2][kRoiP];
  __shared__ int t_ok;
  ROI_STAMP(1);
  const bool try_tab = staged && P <= kRoiP;
  if (try_tab) {
    if (tid == 0) t_ok = 1;
    for (int i = tid; i < 2 * kRoiP * kRoiSpan; i += 256) (&wtab[0][0][0])[i] = 0.f;
    __syncthreads();
    if (tid < 2 * P) {
      const int axis = tid / P, pb = tid - axis * P;            // axis 0: x (columns), 1: y (rows)
      const int g = axis ? gh : gw, dim = axis ? Hf : Wf, lo_w = axis ? y_lo : x_lo, hi_w = axis ? y_hi : x_hi;
      const float start = axis ? sh : sw, bin = axis ? bh : bw;
      float* w = wtab[axis][pb];
      int base = 0, cnt = 0, ok = 1;
      for (int i = 0; i < g; ++i) {
        float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)pb, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)g));
        if (c < -1.0f || c > (float)dim) continue;
        c = fmaxf(c, 0.f);
        int lo = (int)c, hi;
        if (lo >= dim - 1) { hi = lo = dim - 1; c = (float)lo; } else { hi = lo + 1; }
        const float l = c - (float)lo, h = 1.f - l;
        if (cnt == 0) base = lo;
        if (hi - base >= kRoiSpan || lo < base || lo < lo_w || hi > hi_w) { ok = 0; break; }   // span / window assumptions
        w[lo - base] += h;
        w[hi - base] += l;
        cnt = hi - base + 1;
      }
      t_start[axis][pb] = base - lo_w;
      t_cnt[axis][pb] = cnt;
      if (!ok) t_ok = 0;
    }
    __syncthreads();
  }
  const bool tabled = try_tab && t_ok;
  ROI_STAMP(2);
  if (staged) {
    // Window load with cp.async: every 8 / 16-byte piece is issued before anything is waited for (the register-staged
    // version exposed one DRAM round trip per batch of 4 loads: 12.5 us per CTA, ncu: 41 % of stalls on the merge right
    // after the loads).  16 lanes cover one pixel's 64 channels; pixel p of the window lives at win + p * 64 floats.
    // fp32 source: 16-byte copies straight into place.  bf16-plane source: the hi / lo halves of the pixel (128 B each) land
    // in the two halves of the pixel's 256 bytes and are merged to fp32 in place afterwards.
    const int cq = tid & (kRoiCq - 1), slot = tid >> 4, npix = rows * cols;
    const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win);
    {
      int y = slot / cols, x = slot - y * cols;
      for (int pix = slot; pix < npix; pix += 16) {
        const size_t g4 = f0 + ((size_t)(y_lo + y) * Wf + (x_lo + x)) * C4 + cq;       // index in units of 4 elements
        const uint32_t dst = win_s + (uint32_t)pix * 256u;
        if (feat_plane == 0) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)cq * 16u),
                       "l"(reinterpret_cast<const float4*>(feat) + g4) : "memory");
        } else {
          const uint2* hp = reinterpret_cast<const uint2*>(feat) + g4;
          const uint2* lp = reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(feat) + feat_plane) + g4;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + (uint32_t)cq * 8u), "l"(hp) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 128u + (uint32_t)cq * 8u), "l"(lp) : "memory");
        }
        x += 16;
        while (x >= cols) { x -= cols; ++y; }
      }
    }
    ROI_STAMP(3);
    asm volatile("cp.async.wait_all;" ::: "memory");
    ROI_STAMP(4);
    if (feat_plane != 0) {
      // in-place merge: a half-warp owns one pixel; all 16 lanes read their hi / lo pieces before any of them writes
      // (the loop trip count is warp-uniform up to the last pass, so the full-mask __syncwarp is reached by every lane)
      const int passes = (npix + 15) >> 4;
      for (int it = 0; it < passes; ++it) {
        const int pix = slot + 16 * it;
        uint2 h = make_uint2(0u, 0u), l = h;
        uint8_t* base = reinterpret_cast<uint8_t*>(win) + (size_t)pix * 256;
        if (pix < npix) {
          h = *reinterpret_cast<const uint2*>(base + cq * 8);
          l = *reinterpret_cast<const uint2*>(base + 128 + cq * 8);
        }
        __syncwarp();
        if (pix < npix) *reinterpret_cast<float4*>(base + cq * 16) = merge4(h, l);
        __syncwarp();
      }
    }
    __syncthreads();
  }

  ROI_STAMP(5);
  const int items = P * P * kRoiCq;
  for (int item = tid; item < items; item += 256) {
    const int bin = item / kRoiCq, cq = item - bin * kRoiCq;
    const int ph = bin / P, pw = bin - ph * P;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tabled) {
      const int y0 = t_start[1][ph], ny = t_cnt[1][ph], x0 = t_start[0][pw], nx = t_cnt[0][pw];
      const float* wy = wtab[1][ph];
      const float* wx = wtab[0][pw];
      const float4* base4 = reinterpret_cast<const float4*>(win) + ((size_t)y0 * cols + x0) * kRoiCq + cq;
      if (nx > 0 && nx <= 4) {
        // the common case (bins narrower than 3 px): four independent taps per row, weights hoisted out of the row loop;
        // taps beyond nx re-read the last valid column with weight 0
        const float w0 = wx[0], w1 = nx > 1 ? wx[1] : 0.f, w2 = nx > 2 ? wx[2] : 0.f, w3 = nx > 3 ? wx[3] : 0.f;
        const int o1 = min(1, nx - 1) * kRoiCq, o2 = min(2, nx - 1) * kRoiCq, o3 = min(3, nx - 1) * kRoiCq;
        for (int j = 0; j < ny; ++j) {
          const float4* rowp = base4 + (size_t)j * cols * kRoiCq;
          const float4 v0 = rowp[0], v1 = rowp[o1], v2 = rowp[o2], v3 = rowp[o3];
          const float wj = wy[j];
          float4 t;
          t.x = fmaf(w3, v3.x, fmaf(w2, v2.x, fmaf(w1, v1.x, w0 * v0.x)));
          t.y = fmaf(w3, v3.y, fmaf(w2, v2.y, fmaf(w1, v1.y, w0 * v0.y)));
          t.z = fmaf(w3, v3.z, fmaf(w2, v2.z, fmaf(w1, v1.z, w0 * v0.z)));
          t.w = fmaf(w3, v3.w, fmaf(w2, v2.w, fmaf(w1, v1.w, w0 * v0.w)));
          acc.x = fmaf(wj, t.x, acc.x); acc.y = fmaf(wj, t.y, acc.y); acc.z = fmaf(wj, t.z, acc.z); acc.w = fmaf(wj, t.w, acc.w);
        }
      } else {
        for (int j = 0; j < ny; ++j) {
          const float4* rowp = base4 + (size_t)j * cols * kRoiCq;
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int i = 0; i < nx; ++i) {
            const float w = wx[i];
            const float4 v = rowp[i * kRoiCq];
            t.x = fmaf(w, v.x, t.x); t.y = fmaf(w, v.y, t.y); t.z = fmaf(w, v.z, t.z); t.w = fmaf(w, v.w, t.w);
          }
          const float wj = wy[j];
          acc.x = fmaf(wj, t.x, acc.x); acc.y = fmaf(wj, t.y, acc.y); acc.z = fmaf(wj, t.z, acc.z); acc.w = fmaf(wj, t.w, acc.w);
        }
      }
    } else {
      for (int iy = 0; iy < gh; ++iy) {
        float y = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)ph, bh)), __fdiv_rn(__fmul_rn((float)iy + 0.5f, bh), (float)gh));
        for (int ix = 0; ix < gw; ++ix) {
          float x = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)pw, bw)), __fdiv_rn(__fmul_rn((float)ix + 0.5f, bw), (float)gw));
          if (y < -1.0f || y > (float)Hf || x < -1.0f || x > (float)Wf) continue;
          float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
          int yl = (int)yy, xl = (int)xx, yh, xh;
          if (yl >= Hf - 1) { yh = yl = Hf - 1; yy = (float)yl; } else { yh = yl + 1; }
          if (xl >= Wf - 1) { xh = xl = Wf - 1; xx = (float)xl; } else { xh = xl + 1; }
          const float ly = yy - (float)yl, lx = xx - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
          const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
          float4 a, bb, cc, d;
          if (staged) {
            // clamped into the window: never out of bounds even if a rounding corner case widened the sample range
            const int r0 = min(max(yl, y_lo), y_hi) - y_lo, r1 = min(max(yh, y_lo), y_hi) - y_lo;
            const int c0 = min(max(xl, x_lo), x_hi) - x_lo, c1 = min(max(xh, x_lo), x_hi) - x_lo;
            const float4* w4p = reinterpret_cast<const float4*>(win) + cq;
            a = w4p[(r0 * cols + c0) * kRoiCq]; bb = w4p[(r0 * cols + c1) * kRoiCq];
            cc = w4p[(r1 * cols + c0) * kRoiCq]; d = w4p[(r1 * cols + c1) * kRoiCq];
          } else {
            a = ld4_fmt(feat, feat_plane, f0 + ((size_t)yl * Wf + xl) * C4 + cq);
            bb = ld4_fmt(feat, feat_plane, f0 + ((size_t)yl * Wf + xh) * C4 + cq);
            cc = ld4_fmt(feat, feat_plane, f0 + ((size_t)yh * Wf + xl) * C4 + cq);
            d = ld4_fmt(feat, feat_plane, f0 + ((size_t)yh * Wf + xh) * C4 + cq);
          }
          acc.x += w1 * a.x + w2 * bb.x + w3 * cc.x + w4 * d.x;
          acc.y += w1 * a.y + w2 * bb.y + w3 * cc.y + w4 * d.y;
          acc.z += w1 * a.z + w2 * bb.z + w3 * cc.z + w4 * d.z;
          acc.w += w1 * a.w + w2 * bb.w + w3 * cc.w + w4 * d.w;
        }
      }
    }
    acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count);
    acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count);
    st4_fmt(out, out_plane, ((size_t)k * P * P + bin) * C4 + (size_t)chunk * kRoiCq + cq, acc);
  }
  ROI_STAMP(6);
#undef ROI_STAMP
}


// ------------------------------------------------------------------ row-per-warp windowed kernel (P = 7, C % 128 == 0)
// The same separable arithmetic as roi_align_win_kernel (bit-identical results), re-mapped so that nothing is recomputed per
// thread and no index is divided: a CTA owns (ROI, 128-channel chunk); warp w owns output row ph = w % 7 (two warps per row,
// splitting the 7 columns 4 + 3) and lane l owns channel quad l, so every tap is one conflict-free 512-byte LDS.128 per warp
// and the y-table of the row is warp-uniform.  Warp 0 alone does the ROI geometry and (while the window is in flight) the
// weight tables; the window is fetched pixel-per-warp with 16-byte cp.async pieces (32 lanes = the pixel's 512 bytes: fp32,
// or 256 B of the hi plane + 256 B of the lo plane, merged to fp32 in place by the warp that fetched it).
// Instruction budget per 128 channels: ~8 k warp-instructions against ~20 k for two CTAs of the 64-channel kernel, which was
// issue-bound (profiles/r1_timeline_roi.log).
constexpr int kR2P = 7, kR2Span = 32;

// CH = 128: 14 warps, lane = channel quad, warps w and w + 7 share output row w (columns 0-3 / 4-6); 100 KB window, 2 CTAs per SM.
// CH = 64:   7 warps, half-warps share the row (lanes 0-15: columns 0-3, lanes 16-31: columns 4-6), lane % 16 = channel quad;
//            48 KB window, 4 CTAs per SM -- twice the CTAs in flight to hide the box -> geometry -> window -> bins chain.
template <int CH>
__global__ void __launch_bounds__(CH == 128 ? 448 : 224, CH == 128 ? 2 : 4)
roi_align_row_kernel(const void* __restrict__ feat, long long feat_plane, int B, int Hf, int Wf, int C,
                     const int32_t* __restrict__ boxes, const int32_t* __restrict__ seg_off, float scale,
                     void* __restrict__ out, long long out_plane, int32_t* __restrict__ sample_grid, int win_bytes) {
  constexpr int P = kR2P;
  constexpr int kCq = CH / 4;                 // channel quads per chunk = 16-byte pieces per window pixel
  constexpr int kPix = 32 / kCq;              // window pixels one warp moves per iteration (1 or 2)
  constexpr int kWarps = 7 * (2 / kPix);      // 14 or 7
  constexpr int kPixBytes = CH * 4;
  extern __shared__ __align__(16) unsigned char win_raw[];
  __shared__ float wtab[2][kR2P][kR2Span];
  __shared__ int t_start[2][kR2P], t_cnt[2][kR2P];
  __shared__ int g_i[8];        // y_lo, x_lo, rows, cols, staged, b, gh, gw
  __shared__ float g_f[4];      // sh, sw, bh, bw
  __shared__ int t_ok;
  const int nchunk = C / CH, k = blockIdx.x / nchunk, chunk = blockIdx.x - k * nchunk;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int piece = lane % kCq, sub = lane / kCq;

  // ---- warp 0: geometry (operation for operation as in the other kernels: the sample grid stays bit-exact)
  float sh = 0.f, sw = 0.f, bh = 0.f, bw = 0.f;
  int gh = 0, gw = 0, y_lo = 0, y_hi = 0, x_lo = 0, x_hi = 0;
  if (warp == 0) {
    int b;
    if (B <= 31) {
      const int so = (lane >= 1 && lane < B) ? __ldg(seg_off + lane) : 0x7fffffff;
      b = __popc(__ballot_sync(0xffffffffu, so <= k));
    } else {
      b = sample_of(seg_off, B, k);
    }
    const int4 bx = __ldg(reinterpret_cast<const int4*>(boxes) + k);
    sw = __fmul_rn((float)bx.x, scale); sh = __fmul_rn((float)bx.y, scale);
    const float ew = __fmul_rn((float)bx.z, scale), eh = __fmul_rn((float)bx.w, scale);
    const float rw = fmaxf(__fsub_rn(ew, sw), 1.0f), rh = fmaxf(__fsub_rn(eh, sh), 1.0f);
    bw = __fdiv_rn(rw, (float)P); bh = __fdiv_rn(rh, (float)P);
    gh = (int)ceilf(__fdiv_rn(rh, (float)P));
    gw = (int)ceilf(__fdiv_rn(rw, (float)P));
    const float y_first = __fadd_rn(sh, __fdiv_rn(__fmul_rn(0.5f, bh), (float)gh));
    const float y_last = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)(P - 1), bh)), __fdiv_rn(__fmul_rn((float)gh - 0.5f, bh), (float)gh));
    const float x_first = __fadd_rn(sw, __fdiv_rn(__fmul_rn(0.5f, bw), (float)gw));
    const float x_last = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)(P - 1), bw)), __fdiv_rn(__fmul_rn((float)gw - 0.5f, bw), (float)gw));
    y_lo = min(max((int)fmaxf(y_first, 0.f), 0), Hf - 1); y_hi = min(max((int)fmaxf(y_last, 0.f) + 1, y_lo), Hf - 1);
    x_lo = min(max((int)fmaxf(x_first, 0.f), 0), Wf - 1); x_hi = min(max((int)fmaxf(x_last, 0.f) + 1, x_lo), Wf - 1);
    const int rows = y_hi - y_lo + 1, cols = x_hi - x_lo + 1;
    if (lane == 0) {
      g_i[0] = y_lo; g_i[1] = x_lo; g_i[2] = rows; g_i[3] = cols;
      g_i[4] = ((long long)rows * cols * kPixBytes <= (long long)win_bytes) ? 1 : 0;
      g_i[5] = b; g_i[6] = gh; g_i[7] = gw;
      g_f[0] = sh; g_f[1] = sw; g_f[2] = bh; g_f[3] = bw;
      if (sample_grid && chunk == 0) { sample_grid[2 * k] = gh; sample_grid[2 * k + 1] = gw; }
    }
  }
  __syncthreads();
  const int rows = g_i[2], cols = g_i[3], b = g_i[5];
  const bool staged = g_i[4] != 0;
  y_lo = g_i[0]; x_lo = g_i[1];
  const int C4 = C >> 2, npix = rows * cols;
  const size_t f0 = (size_t)b * Hf * Wf * C4 + (size_t)chunk * kCq;       // in units of 4 elements

  // ---- window fetch: a warp moves kPix pixels per iteration; lane -> (pixel sub, 16-byte piece of its CH*4 bytes)
  if (staged) {
    const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win_raw);
    const int first = warp * kPix + sub;
    int y = first / cols, x = first - y * cols;
    for (int pix = first; pix < npix; pix += kWarps * kPix) {
      const size_t g4 = f0 + ((size_t)(y_lo + y) * Wf + (x_lo + x)) * C4;
      const uint32_t dst = win_s + (uint32_t)pix * (uint32_t)kPixBytes + (uint32_t)piece * 16u;
      if (feat_plane == 0) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const float4*>(feat) + g4 + piece) : "memory");
      } else {
        // the chunk's CH channels are CH*2 bytes = kCq/2 pieces in each plane: hi pieces first, then lo pieces
        const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(feat) + (piece < kCq / 2 ? 0 : feat_plane) + g4 * 4 +
                                   (size_t)(piece % (kCq / 2)) * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
      x += kWarps * kPix;
      while (x >= cols) { x -= cols; ++y; }
    }
  }

  // ---- warp 0: per-bin separable weight tables (same code path as roi_align_win_kernel), while the window is in flight
  if (warp == 0) {
    for (int i = lane; i < 2 * kR2P * kR2Span; i += 32) (&wtab[0][0][0])[i] = 0.f;
    __syncwarp();
    int ok = 1;
    if (lane < 2 * P) {
      const int axis = lane / P, pb = lane - axis * P;            // axis 0: x (columns), 1: y (rows)
      const int g = axis ? gh : gw, dim = axis ? Hf : Wf, lo_w = axis ? y_lo : x_lo, hi_w = axis ? y_hi : x_hi;
      const float start = axis ? sh : sw, bin = axis ? bh : bw;
      float* w = wtab[axis][pb];
      int base = 0, cnt = 0;
      for (int i = 0; i < g; ++i) {
        float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)pb, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)g));
        if (c < -1.0f || c > (float)dim) continue;
        c = fmaxf(c, 0.f);
        int lo = (int)c, hi;
        if (lo >= dim - 1) { hi = lo = dim - 1; c = (float)lo; } else { hi = lo + 1; }
        const float l = c - (float)lo, h = 1.f - l;
        if (cnt == 0) base = lo;
        if (hi - base >= kR2Span || lo < base || lo < lo_w || hi > hi_w) { ok = 0; break; }
        w[lo - base] += h;
        w[hi - base] += l;
        cnt = hi - base + 1;
      }
      t_start[axis][pb] = base - lo_w;
      t_cnt[axis][pb] = cnt;
    }
    const unsigned okm = __ballot_sync(0xffffffffu, ok != 0);
    if (lane == 0) t_ok = (okm == 0xffffffffu) ? 1 : 0;
  }

  // ---- land + merge: each warp waits for ITS pixels only
  if (staged) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    if (feat_plane != 0) {
      for (int p0 = warp * kPix; p0 < npix; p0 += kWarps * kPix) {      // warp-uniform trip count
        const int pix = p0 + sub;
        unsigned char* base = win_raw + (size_t)pix * kPixBytes;
        uint2 h = make_uint2(0u, 0u), l = h;
        if (pix < npix) {
          h = *reinterpret_cast<const uint2*>(base + piece * 8);
          l = *reinterpret_cast<const uint2*>(base + CH * 2 + piece * 8);
        }
        __syncwarp();                       // every lane holds its pieces before the fp32 values overwrite them
        if (pix < npix) *reinterpret_cast<float4*>(base + piece * 16) = merge4(h, l);
      }
    }
  }
  __syncthreads();

  // ---- bins: (warp, half-warp) -> (row ph, column range), lane -> channel quad
  const int ph = warp % P, part = (kPix == 2) ? sub : warp / P;
  const int pw0 = part ? 4 : 0, pw1 = part ? P : 4;
  const float count = (float)max(g_i[6] * g_i[7], 1);
  const size_t o_base = ((size_t)k * P * P + (size_t)ph * P) * C4 + (size_t)chunk * kCq + piece;
  if (staged && t_ok) {
    const int y0 = t_start[1][ph], ny = t_cnt[1][ph];
    const float* wy = wtab[1][ph];
    const float4* win4 = reinterpret_cast<const float4*>(win_raw) + piece;
    for (int pw = pw0; pw < pw1; ++pw) {
      const int x0 = t_start[0][pw], nx = t_cnt[0][pw];
      const float* wx = wtab[0][pw];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < ny; ++j) {
        const float4* rowp = win4 + (size_t)((y0 + j) * cols + x0) * kCq;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int i = 0; i < nx; ++i) {
          const float w = wx[i];
          const float4 v = rowp[i * kCq];
          t.x = fmaf(w, v.x, t.x); t.y = fmaf(w, v.y, t.y); t.z = fmaf(w, v.z, t.z); t.w = fmaf(w, v.w, t.w);
        }
        const float wj = wy[j];
        acc.x = fmaf(wj, t.x, acc.x); acc.y = fmaf(wj, t.y, acc.y); acc.z = fmaf(wj, t.z, acc.z); acc.w = fmaf(wj, t.w, acc.w);
      }
      acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count);
      acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count);
      st4_fmt(out, out_plane, o_base + (size_t)pw * C4, acc);
    }
  } else {
    // window too large for shared memory, or a table assumption failed: per-sample taps straight from global memory
    const float s_h = g_f[0], s_w = g_f[1], b_h = g_f[2], b_w = g_f[3];
    const int g_h = g_i[6], g_w = g_i[7];
    const size_t fq = f0 + piece;
    for (int pw = pw0; pw < pw1; ++pw) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int iy = 0; iy < g_h; ++iy) {
        float y = __fadd_rn(__fadd_rn(s_h, __fmul_rn((float)ph, b_h)), __fdiv_rn(__fmul_rn((float)iy + 0.5f, b_h), (float)g_h));
        for (int ix = 0; ix < g_w; ++ix) {
          float x = __fadd_rn(__fadd_rn(s_w, __fmul_rn((float)pw, b_w)), __fdiv_rn(__fmul_rn((float)ix + 0.5f, b_w), (float)g_w));
          if (y < -1.0f || y > (float)Hf || x < -1.0f || x > (float)Wf) continue;
          float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
          int yl = (int)yy, xl = (int)xx, yh, xh;
          if (yl >= Hf - 1) { yh = yl = Hf - 1; yy = (float)yl; } else { yh = yl + 1; }
          if (xl >= Wf - 1) { xh = xl = Wf - 1; xx = (float)xl; } else { xh = xl + 1; }
          const float ly = yy - (float)yl, lx = xx - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
          const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
          const float4 a = ld4_fmt(feat, feat_plane, fq + ((size_t)yl * Wf + xl) * C4);
          const float4 bb = ld4_fmt(feat, feat_plane, fq + ((size_t)yl * Wf + xh) * C4);
          const float4 cc = ld4_fmt(feat, feat_plane, fq + ((size_t)yh * Wf + xl) * C4);
          const float4 d = ld4_fmt(feat, feat_plane, fq + ((size_t)yh * Wf + xh) * C4);
          acc.x += w1 * a.x + w2 * bb.x + w3 * cc.x + w4 * d.x;
          acc.y += w1 * a.y + w2 * bb.y + w3 * cc.y + w4 * d.y;
          acc.z += w1 * a.z + w2 * bb.z + w3 * cc.z + w4 * d.z;
          acc.w += w1 * a.w + w2 * bb.w + w3 * cc.w + w4 * d.w;
        }
      }
      acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count);
      acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count);
      st4_fmt(out, out_plane, o_base + (size_t)pw * C4, acc);
    }
  }
}


// ------------------------------------------------------------------ persistent, double-buffered form of the row-per-warp kernel
// NOT YET RUN ON A GPU (written after this round's GPU budget was spent; opt-in: VBG_ROI_ROW=3 / 4; bit-equality with the
// kernels above and the timing are one `python scripts/roi_compare.py` away).  Why: roi_align_row_kernel keeps its window loads
// in flight for only ~2.5 of a CTA's ~9 us, and bytes in flight per SM x SMs / DRAM latency is exactly the 2.9 TB/s it reaches
// (DESIGN.md section 8).  Here a CTA is persistent over (ROI, chunk) items and owns TWO window buffers: the geometry and the
// cp.async fetch of item i + 1 are issued before the merge + bins of item i (cp.async commit groups, wait_group 1), so a window
// is in flight all the time.  The arithmetic is the row kernel's, statement for statement.
// Slots: per-item geometry / tables live in 3 rotating slots (item i + 1 is written while items i - 1 and i may still be read),
// windows in 2 buffers (buffer of item i + 1 is rewritten only after barrier A, i.e. after every warp left bins(i - 1)).
struct RoiItemGeo {
  int y_lo, x_lo, rows, cols, staged, b, gh, gw, ok;
  float sh, sw, bh, bw;
};

template <int CH>
__global__ void __launch_bounds__(CH == 128 ? 448 : 224, CH == 128 ? 1 : 2)
roi_align_pipe_kernel(const void* __restrict__ feat, long long feat_plane, int B, int Hf, int Wf, int C,
                      const int32_t* __restrict__ boxes, const int32_t* __restrict__ seg_off, int K, float scale,
                      void* __restrict__ out, long long out_plane, int32_t* __restrict__ sample_grid, int win_bytes) {
  constexpr int P = kR2P;
  constexpr int kCq = CH / 4, kPix = 32 / kCq, kWarps = 7 * (2 / kPix), kPixBytes = CH * 4;
  extern __shared__ __align__(16) unsigned char win_all[];          // two windows of win_bytes each
  __shared__ float wtab[3][2][kR2P][kR2Span];
  __shared__ int t_start[3][2][kR2P], t_cnt[3][2][kR2P];
  __shared__ RoiItemGeo geo[3];
  const int nchunk = C / CH, total = K * nchunk;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int piece = lane % kCq, sub = lane / kCq;
  const int C4 = C >> 2;

  // warp 0: geometry of `item` into slot `sl` (same operations as every other ROI kernel here: bit-exact sample grid)
  auto geometry = [&](int item, int sl) {
    const int k = item / nchunk, chunk = item - k * nchunk;
    int b;
    if (B <= 31) {
      const int so = (lane >= 1 && lane < B) ? __ldg(seg_off + lane) : 0x7fffffff;
      b = __popc(__ballot_sync(0xffffffffu, so <= k));
    } else {
      b = sample_of(seg_off, B, k);
    }
    const int4 bx = __ldg(reinterpret_cast<const int4*>(boxes) + k);
    const float sw = __fmul_rn((float)bx.x, scale), sh = __fmul_rn((float)bx.y, scale);
    const float ew = __fmul_rn((float)bx.z, scale), eh = __fmul_rn((float)bx.w, scale);
    const float rw = fmaxf(__fsub_rn(ew, sw), 1.0f), rh = fmaxf(__fsub_rn(eh, sh), 1.0f);
    const float bw = __fdiv_rn(rw, (float)P), bh = __fdiv_rn(rh, (float)P);
    const int gh = (int)ceilf(__fdiv_rn(rh, (float)P));
    const int gw = (int)ceilf(__fdiv_rn(rw, (float)P));
    const float y_first = __fadd_rn(sh, __fdiv_rn(__fmul_rn(0.5f, bh), (float)gh));
    const float y_last = __fadd_rn(__fadd_rn(sh, __fmul_rn((float)(P - 1), bh)), __fdiv_rn(__fmul_rn((float)gh - 0.5f, bh), (float)gh));
    const float x_first = __fadd_rn(sw, __fdiv_rn(__fmul_rn(0.5f, bw), (float)gw));
    const float x_last = __fadd_rn(__fadd_rn(sw, __fmul_rn((float)(P - 1), bw)), __fdiv_rn(__fmul_rn((float)gw - 0.5f, bw), (float)gw));
    const int y_lo = min(max((int)fmaxf(y_first, 0.f), 0), Hf - 1), y_hi = min(max((int)fmaxf(y_last, 0.f) + 1, y_lo), Hf - 1);
    const int x_lo = min(max((int)fmaxf(x_first, 0.f), 0), Wf - 1), x_hi = min(max((int)fmaxf(x_last, 0.f) + 1, x_lo), Wf - 1);
    const int rows = y_hi - y_lo + 1, cols = x_hi - x_lo + 1;
    if (lane == 0) {
      RoiItemGeo g;
      g.y_lo = y_lo; g.x_lo = x_lo; g.rows = rows; g.cols = cols;
      g.staged = ((long long)rows * cols * kPixBytes <= (long long)win_bytes) ? 1 : 0;
      g.b = b; g.gh = gh; g.gw = gw; g.ok = 0;
      g.sh = sh; g.sw = sw; g.bh = bh; g.bw = bw;
      geo[sl] = g;
      if (sample_grid && chunk == 0) { sample_grid[2 * k] = gh; sample_grid[2 * k + 1] = gw; }
    }
    __syncwarp();
  };

  // warp 0: separable per-bin weight tables of the item in slot `sl` (geometry already published there)
  auto tables = [&](int sl) {
    const RoiItemGeo g = geo[sl];
    const int y_hi = g.y_lo + g.rows - 1, x_hi = g.x_lo + g.cols - 1;
    for (int i = lane; i < 2 * kR2P * kR2Span; i += 32) (&wtab[sl][0][0][0])[i] = 0.f;
    __syncwarp();
    int ok = 1;
    if (lane < 2 * P) {
      const int axis = lane / P, pb = lane - axis * P;            // axis 0: x (columns), 1: y (rows)
      const int gn = axis ? g.gh : g.gw, dim = axis ? Hf : Wf, lo_w = axis ? g.y_lo : g.x_lo, hi_w = axis ? y_hi : x_hi;
      const float start = axis ? g.sh : g.sw, bin = axis ? g.bh : g.bw;
      float* w = wtab[sl][axis][pb];
      int base = 0, cnt = 0;
      for (int i = 0; i < gn; ++i) {
        float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)pb, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)gn));
        if (c < -1.0f || c > (float)dim) continue;
        c = fmaxf(c, 0.f);
        int lo = (int)c, hi;
        if (lo >= dim - 1) { hi = lo = dim - 1; c = (float)lo; } else { hi = lo + 1; }
        const float l = c - (float)lo, h = 1.f - l;
        if (cnt == 0) base = lo;
        if (hi - base >= kR2Span || lo < base || lo < lo_w || hi > hi_w) { ok = 0; break; }
        w[lo - base] += h;
        w[hi - base] += l;
        cnt = hi - base + 1;
      }
      t_start[sl][axis][pb] = base - lo_w;
      t_cnt[sl][axis][pb] = cnt;
    }
    const unsigned okm = __ballot_sync(0xffffffffu, ok != 0);
    if (lane == 0) geo[sl].ok = (okm == 0xffffffffu) ? 1 : 0;
    __syncwarp();
  };

  // all warps: cp.async fetch of the item's window into `win`; always commits one group per thread
  auto fetch = [&](int item, int sl, unsigned char* win) {
    const RoiItemGeo g = geo[sl];
    if (g.staged) {
      const int chunk = item % nchunk, npix = g.rows * g.cols, cols = g.cols;
      const size_t f0 = (size_t)g.b * Hf * Wf * C4 + (size_t)chunk * kCq;
      const uint32_t win_s = (uint32_t)__cvta_generic_to_shared(win);
      const int first = warp * kPix + sub;
      int y = first / cols, x = first - y * cols;
      for (int pix = first; pix < npix; pix += kWarps * kPix) {
        const size_t g4 = f0 + ((size_t)(g.y_lo + y) * Wf + (g.x_lo + x)) * C4;
        const uint32_t dst = win_s + (uint32_t)pix * (uint32_t)kPixBytes + (uint32_t)piece * 16u;
        if (feat_plane == 0) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const float4*>(feat) + g4 + piece) : "memory");
        } else {
          const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(feat) + (piece < kCq / 2 ? 0 : feat_plane) + g4 * 4 +
                                     (size_t)(piece % (kCq / 2)) * 8;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        x += kWarps * kPix;
        while (x >= cols) { x -= cols; ++y; }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ---- prologue: first item of this CTA
  int item = blockIdx.x;
  if (item >= total) return;
  int sl = 0, buf = 0;
  if (warp == 0) geometry(item, sl);
  __syncthreads();
  fetch(item, sl, win_all);
  if (warp == 0) tables(sl);

  for (; item < total; item += gridDim.x) {
    const int next = item + gridDim.x;
    const bool has_next = next < total;                               // CTA-uniform
    const int sl_n = (sl + 1) % 3;
    if (has_next) {
      if (warp == 0) geometry(next, sl_n);
      __syncthreads();                                                // (A) geometry(next) visible; every warp has left bins(item - stride)
      fetch(next, sl_n, win_all + (size_t)(buf ^ 1) * win_bytes);
      if (warp == 0) tables(sl_n);
      asm volatile("cp.async.wait_group 1;" ::: "memory");            // this item's group has landed (the newest one may be in flight)
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();

    const RoiItemGeo g = geo[sl];
    unsigned char* win = win_all + (size_t)buf * win_bytes;
    const int npix = g.rows * g.cols, cols = g.cols;
    if (g.staged && feat_plane != 0) {                                // merge the warp's own pixels: hi + lo -> fp32 in place
      for (int p0 = warp * kPix; p0 < npix; p0 += kWarps * kPix) {
        const int pix = p0 + sub;
        unsigned char* base = win + (size_t)pix * kPixBytes;
        uint2 h = make_uint2(0u, 0u), l = h;
        if (pix < npix) {
          h = *reinterpret_cast<const uint2*>(base + piece * 8);
          l = *reinterpret_cast<const uint2*>(base + CH * 2 + piece * 8);
        }
        __syncwarp();
        if (pix < npix) *reinterpret_cast<float4*>(base + piece * 16) = merge4(h, l);
      }
    }
    __syncthreads();                                                  // (B) window complete, tables of this item visible

    const int k = item / nchunk, chunk = item - k * nchunk;
    const int ph = warp % P, part = (kPix == 2) ? sub : warp / P;
    const int pw0 = part ? 4 : 0, pw1 = part ? P : 4;
    const float count = (float)max(g.gh * g.gw, 1);
    const size_t o_base = ((size_t)k * P * P + (size_t)ph * P) * C4 + (size_t)chunk * kCq + piece;
    if (g.staged && geo[sl].ok) {
      const int y0 = t_start[sl][1][ph], ny = t_cnt[sl][1][ph];
      const float* wy = wtab[sl][1][ph];
      const float4* win4 = reinterpret_cast<const float4*>(win) + piece;
      for (int pw = pw0; pw < pw1; ++pw) {
        const int x0 = t_start[sl][0][pw], nx = t_cnt[sl][0][pw];
        const float* wx = wtab[sl][0][pw];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < ny; ++j) {
          const float4* rowp = win4 + (size_t)((y0 + j) * cols + x0) * kCq;
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
          for (int i = 0; i < nx; ++i) {
            const float w = wx[i];
            const float4 v = rowp[i * kCq];
            t.x = fmaf(w, v.x, t.x); t.y = fmaf(w, v.y, t.y); t.z = fmaf(w, v.z, t.z); t.w = fmaf(w, v.w, t.w);
          }
          const float wj = wy[j];
          acc.x = fmaf(wj, t.x, acc.x); acc.y = fmaf(wj, t.y, acc.y); acc.z = fmaf(wj, t.z, acc.z); acc.w = fmaf(wj, t.w, acc.w);
        }
        acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count);
        acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count);
        st4_fmt(out, out_plane, o_base + (size_t)pw * C4, acc);
      }
    } else {
      const size_t fq = (size_t)g.b * Hf * Wf * C4 + (size_t)chunk * kCq + piece;
      for (int pw = pw0; pw < pw1; ++pw) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int iy = 0; iy < g.gh; ++iy) {
          float y = __fadd_rn(__fadd_rn(g.sh, __fmul_rn((float)ph, g.bh)), __fdiv_rn(__fmul_rn((float)iy + 0.5f, g.bh), (float)g.gh));
          for (int ix = 0; ix < g.gw; ++ix) {
            float x = __fadd_rn(__fadd_rn(g.sw, __fmul_rn((float)pw, g.bw)), __fdiv_rn(__fmul_rn((float)ix + 0.5f, g.bw), (float)g.gw));
            if (y < -1.0f || y > (float)Hf || x < -1.0f || x > (float)Wf) continue;
            float yy = fmaxf(y, 0.f), xx = fmaxf(x, 0.f);
            int yl = (int)yy, xl = (int)xx, yh, xh;
            if (yl >= Hf - 1) { yh = yl = Hf - 1; yy = (float)yl; } else { yh = yl + 1; }
            if (xl >= Wf - 1) { xh = xl = Wf - 1; xx = (float)xl; } else { xh = xl + 1; }
            const float ly = yy - (float)yl, lx = xx - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
            const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
            const float4 a = ld4_fmt(feat, feat_plane, fq + ((size_t)yl * Wf + xl) * C4);
            const float4 bb = ld4_fmt(feat, feat_plane, fq + ((size_t)yl * Wf + xh) * C4);
            const float4 cc = ld4_fmt(feat, feat_plane, fq + ((size_t)yh * Wf + xl) * C4);
            const float4 d = ld4_fmt(feat, feat_plane, fq + ((size_t)yh * Wf + xh) * C4);
            acc.x += w1 * a.x + w2 * bb.x + w3 * cc.x + w4 * d.x;
            acc.y += w1 * a.y + w2 * bb.y + w3 * cc.y + w4 * d.y;
            acc.z += w1 * a.z + w2 * bb.z + w3 * cc.z + w4 * d.z;
            acc.w += w1 * a.w + w2 * bb.w + w3 * cc.w + w4 * d.w;
          }
        }
        acc.x = __fdiv_rn(acc.x, count); acc.y = __fdiv_rn(acc.y, count);
        acc.z = __fdiv_rn(acc.z, count); acc.w = __fdiv_rn(acc.w, count);
        st4_fmt(out, out_plane, o_base + (size_t)pw * C4, acc);
      }
    }
    sl = sl_n;
    buf ^= 1;
  }
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_roi_align_fwd(const float* feat, int B, int Hf, int Wf, int C, const int32_t* boxes,
                                 const int32_t* seg_off, int K, float spatial_scale, int P, float* out,
                                 int32_t* sample_grid, vbg_stream_t stream) {
  return vbg_roi_align_x(feat, 0, B, Hf, Wf, C, boxes, seg_off, K, spatial_scale, P, out, 0, sample_grid, stream);
}

extern "C" int vbg_roi_align_x(const void* feat, long long feat_plane, int B, int Hf, int Wf, int C, const int32_t* boxes,
                               const int32_t* seg_off, int K, float spatial_scale, int P, void* out, long long out_plane,
                               int32_t* sample_grid, vbg_stream_t stream) {
  VBG_REQUIRE(feat && boxes && seg_off && out && B > 0 && Hf > 0 && Wf > 0 && P > 0 && K >= 0,
              "vbg_roi_align_fwd: bad arguments");
  VBG_REQUIRE(C > 0 && C % 4 == 0 && C <= 1024 && fmt_ok(feat, feat_plane) && fmt_ok(out, out_plane) && aligned16(boxes),
              "vbg_roi_align_fwd: C %% 4 == 0, C <= 1024 and 16B alignment required (C=%d)", C);
  if (K == 0) return VBG_OK;
  if (feat_plane == 0 && out_plane == 0) {
    // The separable form measured SLOWER than the direct one on B200 (cfg2: 101 us vs 67 us): opt-in for experiments only.
    const int C4 = C / 4;
    static const bool sep = [] { const char* e = getenv("VBG_ROI_SEPARABLE"); return e && e[0] == '1'; }();
    const int tab_stride = (Hf > Wf ? Hf : Wf) + 1;
    const size_t smem = (size_t)2 * 7 * tab_stride * sizeof(float);
    if (sep && P == 7 && C4 <= 256 && 256 % C4 == 0 && smem <= 48 * 1024) {
      roi_align_sep_kernel<7><<<K, 256, smem, as_stream(stream)>>>(reinterpret_cast<const float*>(feat), B, Hf, Wf, C, boxes, seg_off,
                                                                  spatial_scale, reinterpret_cast<float*>(out), sample_grid, tab_stride);
      return check_launch("vbg_roi_align_fwd");
    }
  }
  cudaStream_t s = as_stream(stream);
  static const bool direct = [] { const char* e = getenv("VBG_ROI_DIRECT"); return e && e[0] == '1'; }();
  // VBG_ROI_ROW: 1 (default) = row-per-warp kernel, 128-channel chunks; 2 = its 64-channel form; 0 = the 64-channel windowed
  // kernel below.  Read per call: tests and scripts/roi_compare.py switch kernels inside one process.
  const char* row_env = getenv("VBG_ROI_ROW");
  const int rowk = row_env ? (row_env[0] - '0') : 1;
  if (!direct && rowk == 1 && P == kR2P && C % 128 == 0) {
    // 100 KB window (two resident CTAs of 14 warps per SM): a line-sized ROI at stride 4 needs 60-92 KB per 128-channel chunk
    constexpr int win_bytes = 100 * 1024;
    static bool attr2 = false;
    if (!attr2) {
      cudaError_t e = cudaFuncSetAttribute(roi_align_row_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, win_bytes);
      if (e != cudaSuccess) { set_error("vbg_roi_align_fwd: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
      attr2 = true;
    }
    roi_align_row_kernel<128><<<(unsigned)((long long)K * (C / 128)), 448, win_bytes, s>>>(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off,
                                                                                        spatial_scale, out, out_plane, sample_grid, win_bytes);
    return check_launch("vbg_roi_align_fwd");
  }
  if (!direct && rowk == 2 && P == kR2P && C % 64 == 0) {
    constexpr int win_bytes = 48 * 1024;      // four resident CTAs of 7 warps per SM
    static bool attr3 = false;
    if (!attr3) {
      cudaError_t e = cudaFuncSetAttribute(roi_align_row_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, win_bytes);
      if (e != cudaSuccess) { set_error("vbg_roi_align_fwd: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
      attr3 = true;
    }
    roi_align_row_kernel<64><<<(unsigned)((long long)K * (C / 64)), 224, win_bytes, s>>>(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off,
                                                                                      spatial_scale, out, out_plane, sample_grid, win_bytes);
    return check_launch("vbg_roi_align_fwd");
  }
  if (!direct && (rowk == 3 || rowk == 4) && P == kR2P && C % (rowk == 3 ? 64 : 128) == 0) {
    // persistent double-buffered form (opt-in, not yet run on a GPU): 3 = 64-channel chunks, two CTAs per SM with 2 x 48 KB
    // windows each; 4 = 128-channel chunks, one CTA per SM with 2 x 100 KB windows
    const int win_bytes = rowk == 3 ? 48 * 1024 : 100 * 1024;
    const int ctas = rowk == 3 ? 2 * kNumSMs : kNumSMs;
    const long long items = (long long)K * (C / (rowk == 3 ? 64 : 128));
    const unsigned grid = (unsigned)(items < ctas ? items : ctas);
    static bool attr4 = false;
    if (!attr4) {
      cudaError_t e1 = cudaFuncSetAttribute(roi_align_pipe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 48 * 1024);
      cudaError_t e2 = cudaFuncSetAttribute(roi_align_pipe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 100 * 1024);
      if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("vbg_roi_align_fwd: smem opt-in failed"); return VBG_ECUDA; }
      attr4 = true;
    }
    if (rowk == 3)
      roi_align_pipe_kernel<64><<<grid, 224, 2 * win_bytes, s>>>(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off, K, spatial_scale, out, out_plane,
                                                               sample_grid, win_bytes);
    else
      roi_align_pipe_kernel<128><<<grid, 448, 2 * win_bytes, s>>>(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off, K, spatial_scale, out, out_plane,
                                                                sample_grid, win_bytes);
    return check_launch("vbg_roi_align_fwd");
  }
  if (!direct && C % kRoiCh == 0) {
    // 60 KB window (three resident CTAs per SM): a line-sized ROI at stride 4 needs 30-46 KB per 64-channel chunk
    constexpr int win_floats = 60 * 1024 / 4;
    static bool attr = false;
    if (!attr) {
      cudaError_t e = cudaFuncSetAttribute(roi_align_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, win_floats * 4);
      if (e != cudaSuccess) { set_error("vbg_roi_align_fwd: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
      attr = true;
    }
    roi_align_win_kernel<<<(unsigned)((long long)K * (C / kRoiCh)), 256, win_floats * sizeof(float), s>>>(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off,
                                                                                       spatial_scale, P, out, out_plane, sample_grid,
                                                                                       win_floats, roi_debug_timeline());
    return check_launch("vbg_roi_align_fwd");
  }
  long long warps = (long long)K * P * P;
  int blocks = (int)((warps + 7) / 8);
  int vec = (C / 4 + 31) / 32;
#define LAUNCH(V) roi_align_kernel<V><<<blocks, 256, 0, s>>>(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off, K, spatial_scale, P, out, out_plane, sample_grid)
  if (vec <= 1) LAUNCH(1);
  else if (vec == 2) LAUNCH(2);
  else if (vec <= 4) LAUNCH(4);
  else LAUNCH(8);
#undef LAUNCH
  return check_launch("vbg_roi_align_fwd");
}
