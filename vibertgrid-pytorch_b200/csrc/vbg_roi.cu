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

namespace vbg {


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




}  // namespace vbg
using namespace vbg;

namespace vbg {
int launch_roi_align_stream(const void* feat, long long feat_plane, int B, int Hf, int Wf, int C, const int32_t* boxes,
                            const int32_t* seg_off, int K, float scale, void* out, long long out_plane, int32_t* sample_grid,
                            cudaStream_t s);
}

extern "C" int vbg_roi_align_fwd(const float* feat, int B, int Hf, int Wf, int C, const int32_t* boxes,
                                 const int32_t* seg_off, int K, float spatial_scale, int P, float* out,
                                 int32_t* sample_grid, vbg_stream_t stream) {
  return vbg_roi_align_sel(feat, 0, B, Hf, Wf, C, boxes, seg_off, K, spatial_scale, P, out, 0, sample_grid, VBG_ROI_AUTO, stream);
}

extern "C" int vbg_roi_align_x(const void* feat, long long feat_plane, int B, int Hf, int Wf, int C, const int32_t* boxes,
                               const int32_t* seg_off, int K, float spatial_scale, int P, void* out, long long out_plane,
                               int32_t* sample_grid, vbg_stream_t stream) {
  return vbg_roi_align_sel(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off, K, spatial_scale, P, out, out_plane, sample_grid,
                           VBG_ROI_AUTO, stream);
}

// Three kernels, chosen by shape (VBG_ROI_AUTO) or by the caller (tests / scripts compare them on identical inputs):
//   STREAM  persistent TMA row-streaming kernel (vbg_roi_stream.cu): P == 7, C in {128, 256}          -- the product path
//   ROW     row-per-warp kernel over a window staged in shared memory: P == 7, C % 128 == 0             -- round-1 default
//   DIRECT  one warp per (ROI, bin), per-sample taps from global memory: any P, any C % 4 == 0
extern "C" int vbg_roi_align_sel(const void* feat, long long feat_plane, int B, int Hf, int Wf, int C, const int32_t* boxes,
                                 const int32_t* seg_off, int K, float spatial_scale, int P, void* out, long long out_plane,
                                 int32_t* sample_grid, int variant, vbg_stream_t stream) {
  VBG_REQUIRE(feat && boxes && seg_off && out && B > 0 && Hf > 0 && Wf > 0 && P > 0 && K >= 0,
              "vbg_roi_align_fwd: bad arguments");
  VBG_REQUIRE(C > 0 && C % 4 == 0 && C <= 1024 && fmt_ok(feat, feat_plane) && fmt_ok(out, out_plane) && aligned16(boxes),
              "vbg_roi_align_fwd: C %% 4 == 0, C <= 1024 and 16B alignment required (C=%d)", C);
  VBG_REQUIRE(variant >= VBG_ROI_AUTO && variant <= VBG_ROI_DIRECT, "vbg_roi_align_fwd: unknown kernel variant %d", variant);
  if (K == 0) return VBG_OK;
  cudaStream_t s = as_stream(stream);
  const bool stream_ok = P == kR2P && (C == 128 || C == 256);
  const bool row_ok = P == kR2P && C % 128 == 0;
  VBG_REQUIRE(variant != VBG_ROI_STREAM || stream_ok, "vbg_roi_align_fwd: the streaming kernel takes P == 7, C in {128, 256}");
  VBG_REQUIRE(variant != VBG_ROI_ROW || row_ok, "vbg_roi_align_fwd: the row kernel takes P == 7, C %% 128 == 0");
  if (variant == VBG_ROI_AUTO) variant = stream_ok ? VBG_ROI_STREAM : (row_ok ? VBG_ROI_ROW : VBG_ROI_DIRECT);
  if (variant == VBG_ROI_STREAM)
    return launch_roi_align_stream(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off, K, spatial_scale, out, out_plane, sample_grid, s);
  if (variant == VBG_ROI_ROW) {
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
