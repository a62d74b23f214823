// Image-side memory-bound kernels: fused normalise+resize+pad (a1), pooling, BatchNorm folding,
// weight repacking, seg-head broadcast store, layout conversion.
#include "vbg_common.cuh"

namespace vbg {

// ------------------------------------------------------------------ a1
// (x - mean)/std per tap, then bilinear (align_corners=False) exactly as ATen's
// upsample_bilinear2d with scale = in/out; writes the zero-bordered NHWC4 batch [B, H+6, W+6, 4] (3-pixel
// border = the stem conv's padding, 4th channel = 0) that the tensor-core stem reads through one TMA map
// (vbg_gemm_tc3.cu::stem_tc3).  Replaces the per-image
// normalize/interpolate/copy_ kernels of reference pipeline/transform.py:122,149-155,261-269.
// Pixel sources: fp32 CHW planes in [0, 1] (what torchvision's ToTensor hands the reference model, data/SROIE_dataset.py:84-86)
// or the decoded uint8 HWC bytes themselves (the shard format of the input pipeline, shards.py): ToTensor is
// `byte.to(float32).div(255)`, one IEEE division, so doing it here is bit-identical and the host never touches pixels.
struct PixF32 {
  const float* img; int h, w;
  __device__ __forceinline__ float at(int c, int y, int x) const { return __ldg(img + ((size_t)c * h + y) * w + x); }
};
struct PixU8 {
  const unsigned char* img; int h, w;
  __device__ __forceinline__ float at(int c, int y, int x) const {
    return __fdiv_rn((float)__ldg(img + ((size_t)y * w + x) * 3 + c), 255.f);
  }
};

template <class Pix>
__device__ __forceinline__ void normalize_resize_pixel(const Pix& px, int h, int w, float* __restrict__ out, int W, int oh, int ow,
                                                       int x, int y, const float3& mean, const float3& stdv) {
  const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;
  float fy = sy * ((float)y + 0.5f) - 0.5f; if (fy < 0.f) fy = 0.f;
  float fx = sx * ((float)x + 0.5f) - 0.5f; if (fx < 0.f) fx = 0.f;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
  const float m[3] = {mean.x, mean.y, mean.z}, s[3] = {stdv.x, stdv.y, stdv.z};
  float r[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float p00 = __fdiv_rn(px.at(c, y0, x0) - m[c], s[c]);
    float p01 = __fdiv_rn(px.at(c, y0, x1) - m[c], s[c]);
    float p10 = __fdiv_rn(px.at(c, y1, x0) - m[c], s[c]);
    float p11 = __fdiv_rn(px.at(c, y1, x1) - m[c], s[c]);
    r[c] = hy * (hx * p00 + lx * p01) + ly * (hx * p10 + lx * p11);
  }
  *reinterpret_cast<float4*>(out + ((size_t)(y + 3) * (W + 6) + (x + 3)) * 4) = make_float4(r[0], r[1], r[2], 0.f);
}

template <bool kU8>
__global__ void normalize_resize_kernel(const void* __restrict__ img, int h, int w, float* __restrict__ out, int H,
                                        int W, int oh, int ow, float3 mean, float3 stdv) {
  // blockIdx.z = image of a same-shape batch ([n,3,h,w] fp32 or [n,h,w,3] uint8 contiguous in, consecutive samples of the padded batch out)
  out += (size_t)blockIdx.z * (H + 6) * (W + 6) * 4;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  if (kU8) {
    PixU8 px{reinterpret_cast<const unsigned char*>(img) + (size_t)blockIdx.z * 3 * h * w, h, w};
    normalize_resize_pixel(px, h, w, out, W, oh, ow, x, y, mean, stdv);
  } else {
    PixF32 px{reinterpret_cast<const float*>(img) + (size_t)blockIdx.z * 3 * h * w, h, w};
    normalize_resize_pixel(px, h, w, out, W, oh, ow, x, y, mean, stdv);
  }
}

// The whole batch of a collated shard batch in ONE launch, documents of different sizes: `tab` holds per document
// {byte offset into `arena` (lo, hi 32 bits), h, w, oh, ow}; blockIdx.z = document, the x / y grid covers the largest output.
__global__ void decode_batch_u8_kernel(const unsigned char* __restrict__ arena, const int* __restrict__ tab, float* __restrict__ out,
                                       int H, int W, float3 mean, float3 stdv) {
  const int* t = tab + 6 * blockIdx.z;
  const long long off = (long long)(unsigned)t[0] | ((long long)t[1] << 32);
  const int h = t[2], w = t[3], oh = t[4], ow = t[5];
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= ow || y >= oh) return;
  PixU8 px{arena + off, h, w};
  normalize_resize_pixel(px, h, w, out + (size_t)blockIdx.z * (H + 6) * (W + 6) * 4, W, oh, ow, x, y, mean, stdv);
}

// PyTorch stem weight [O,3,7,7] -> (a) [O,7,7,4] (4th channel 0) for the CUDA-core path, (b) [O,8,8,4] with zero
// 8th filter row / 8th pixel / 4th channel: the K = 256 operand of the tensor-core stem GEMM.
__global__ void stem_pack_kernel(const float* __restrict__ w, int O, float* __restrict__ w774, float* __restrict__ w884) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= O * 256) return;
  const int c = i & 3, px = (i >> 2) & 7, r = (i >> 5) & 7, o = i >> 8;
  const float v = (c < 3 && px < 7 && r < 7) ? w[(((size_t)o * 3 + c) * 7 + r) * 7 + px] : 0.f;
  w884[i] = v;
  if (px < 7 && r < 7) w774[(((size_t)o * 7 + r) * 7 + px) * 4 + c] = v;
}

// ------------------------------------------------------------------ pooling (NHWC, float4 over channels)
__global__ void maxpool3x3s2_kernel(const void* __restrict__ x, long long x_plane, int B, int H, int W, int C4, int Ho, int Wo,
                                    void* __restrict__ y, long long y_plane) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * Ho * Wo * C4;
  if (i >= total) return;
  int c = (int)(i % C4); long long t = i / C4;
  int wo = (int)(t % Wo); t /= Wo;
  int ho = (int)(t % Ho); int b = (int)(t / Ho);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int hi = ho * 2 - 1 + dy;
    if (hi < 0 || hi >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      int wi = wo * 2 - 1 + dx;
      if (wi < 0 || wi >= W) continue;
      float4 v = ld4_fmt(x, x_plane, (((size_t)b * H + hi) * W + wi) * C4 + c);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  }
  st4_fmt(y, y_plane, (size_t)i, m);
}

__global__ void avgpool2x2_kernel(const void* __restrict__ x, long long x_plane, int B, int H, int W, int C4,
                                  void* __restrict__ y, long long y_plane) {
  const int Ho = H / 2, Wo = W / 2;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)B * Ho * Wo * C4;
  if (i >= total) return;
  int c = (int)(i % C4); long long t = i / C4;
  int wo = (int)(t % Wo); t /= Wo;
  int ho = (int)(t % Ho); int b = (int)(t / Ho);
  const size_t p = (((size_t)b * H + 2 * ho) * W + 2 * wo) * C4 + c;
  float4 a = ld4_fmt(x, x_plane, p), bb = ld4_fmt(x, x_plane, p + C4), cc = ld4_fmt(x, x_plane, p + (size_t)W * C4),
         d = ld4_fmt(x, x_plane, p + (size_t)W * C4 + C4);
  float4 r;
  r.x = (a.x + bb.x + cc.x + d.x) * 0.25f; r.y = (a.y + bb.y + cc.y + d.y) * 0.25f;
  r.z = (a.z + bb.z + cc.z + d.z) * 0.25f; r.w = (a.w + bb.w + cc.w + d.w) * 0.25f;
  st4_fmt(y, y_plane, (size_t)i, r);
}

// ------------------------------------------------------------------ parameter preparation
__global__ void bn_fold_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ mean,
                               const float* __restrict__ var, float eps, int C, float* __restrict__ scale,
                               float* __restrict__ shift) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float inv = __fdiv_rn(1.0f, sqrtf(var[c] + eps));
  float a = w[c] * inv;
  scale[c] = a;
  shift[c] = b[c] - mean[c] * a;
}

__global__ void repack_oihw_kernel(const float* __restrict__ w, int O, int I, int H, int W, float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)O * I * H * W;
  if (i >= total) return;
  int ci = (int)(i % I); long long t = i / I;     // out index = ((o*H + h)*W + w)*I + ci
  int ww = (int)(t % W); t /= W;
  int hh = (int)(t % H); int o = (int)(t / H);
  out[i] = __ldg(w + (((size_t)o * I + ci) * H + hh) * W + ww);
}

// ------------------------------------------------------------------ outputs
// One CTA per (sample, low-res row): stage the row's [w, Ct] logits in shared memory, then emit `up`
// full-resolution rows per channel with coalesced stores into the two NCHW outputs.
__global__ void upsample_split_kernel(const float* __restrict__ x, int h, int w, int Ct, int up, int c_split,
                                      float* __restrict__ out1, float* __restrict__ out2) {
  extern __shared__ float row[];          // [Ct][w] (channel-major: the emit loop reads consecutive x)
  const int b = blockIdx.y, yi = blockIdx.x;
  const float* src = x + ((size_t)b * h + yi) * w * Ct;
  for (int i = threadIdx.x; i < w * Ct; i += blockDim.x) {
    const int xs = i / Ct, c = i - xs * Ct;
    row[c * w + xs] = __ldg(src + i);
  }
  __syncthreads();
  const int Wf = w * up, Hf = h * up;
  if ((up & 3) == 0) {
    // 128-bit stores: the `up` output columns of one source pixel hold the same value, up/4 float4 per pixel
    const int q = up >> 2, n4 = Ct * up * w * q;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const int xq = i % (w * q); int t = i / (w * q);
      const int dy = t % up, c = t / up;
      const float v = row[c * w + xq / q];
      const int Y = yi * up + dy;
      float* dst = c < c_split ? out1 + (((size_t)b * c_split + c) * Hf + Y) * Wf
                               : out2 + (((size_t)b * (Ct - c_split) + (c - c_split)) * Hf + Y) * Wf;
      reinterpret_cast<float4*>(dst)[xq] = make_float4(v, v, v, v);
    }
    return;
  }
  const int n_out = Ct * up * Wf;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
    int X = i % Wf; int t = i / Wf;
    int dy = t % up; int c = t / up;
    float v = row[c * w + X / up];
    int Y = yi * up + dy;
    if (c < c_split) out1[(((size_t)b * c_split + c) * Hf + Y) * Wf + X] = v;
    else out2[(((size_t)b * (Ct - c_split) + (c - c_split)) * Hf + Y) * Wf + X] = v;
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int HW, int C, float* __restrict__ y) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* src = x + (size_t)b * HW * C;
  float* dst = y + (size_t)b * HW * C;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < HW && c < C) ? __ldg(src + (size_t)p * C + c) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    if (p < HW && c < C) dst[(size_t)c * HW + p] = tile[threadIdx.x][j];
  }
}

// ------------------------------------------------------------------ backward building blocks (first bricks of the training step)
// out[c][r] = x[r][c] as bf16 hi/lo planes with leading dimension ld_out >= rows (columns rows .. ld_out-1 zero): turns a
// [rows, cols] activation / gradient / weight into the K-major operand of a GEMM that reduces over `rows` (wgrad:
// dW = dY^T X) or of the dgrad GEMM (W^T).  32x32 tiles through shared memory, coalesced on both sides.
__global__ void __launch_bounds__(256)
transpose_split_kernel(const void* __restrict__ x, long long x_plane, int rows, int cols, __nv_bfloat16* __restrict__ out,
                       long long out_plane, int ld_out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      const size_t idx = (size_t)r * cols + c;
      v = x_plane == 0 ? __ldg(reinterpret_cast<const float*>(x) + idx)
                       : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[idx]) +
                             __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[idx + x_plane]);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;                              // output row = input column
    if (c < cols && r < ld_out) {
      const float v = tile[tx][i];
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      out[(size_t)c * ld_out + r] = h;
      out[(size_t)c * ld_out + r + out_plane] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}


// w [Cout, kh, kw, Cin] fp32 -> planes of w'[Cin, kh, kw, Cout] with w'[ci, r, s, co] = w[co, kh-1-r, kw-1-s, ci]: the weight of
// the data-gradient convolution dX = conv(dY, w', stride 1, pad k-1-p) of a stride-1 convolution
__global__ void conv_dgrad_weight_kernel(const float* __restrict__ w, int Cout, int kh, int kw, int Cin, __nv_bfloat16* __restrict__ out,
                                         long long out_plane) {
  const long long total = (long long)Cin * kh * kw * Cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout); long long t = i / Cout;
    const int s_ = (int)(t % kw); t /= kw;
    const int r = (int)(t % kh); const int ci = (int)(t / kh);
    const float v = __ldg(w + (((size_t)co * kh + (kh - 1 - r)) * kw + (kw - 1 - s_)) * Cin + ci);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = h;
    out[i + out_plane] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// LayerNorm backward, data gradient: one warp per row, row in registers (hidden <= 1024); statistics recomputed from x
__global__ void __launch_bounds__(256)
ln_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma, float eps, int R,
                 int H4, float* __restrict__ dx, float2* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  constexpr int kMax = 8;
  float4 xv[kMax], gv[kMax];
  const float4* x4 = reinterpret_cast<const float4*>(x) + (size_t)r * H4;
  const float4* d4 = reinterpret_cast<const float4*>(dy) + (size_t)r * H4;
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < H4) { xv[i] = __ldg(x4 + c); sum += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w); }
  }
  const float n = (float)(H4 * 4);
  const float mean = warp_sum(sum) / n;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < H4) {
      xv[i].x -= mean; xv[i].y -= mean; xv[i].z -= mean; xv[i].w -= mean;
      sq += (xv[i].x * xv[i].x + xv[i].y * xv[i].y) + (xv[i].z * xv[i].z + xv[i].w * xv[i].w);
    }
  }
  const float rstd = __fdiv_rn(1.0f, sqrtf(warp_sum(sq) / n + eps));
  if (stats && lane == 0) stats[r] = make_float2(mean, rstd);        // reused by the parameter-gradient reduction
  float s1 = 0.f, s2 = 0.f;                 // sum(dxhat), sum(dxhat * xhat)
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < H4) {
      const float4 d = __ldg(d4 + c), g = __ldg(g4 + c);
      xv[i].x *= rstd; xv[i].y *= rstd; xv[i].z *= rstd; xv[i].w *= rstd;          // xhat
      gv[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);              // dxhat
      s1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
      s2 += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
    }
  }
  const float m1 = warp_sum(s1) / n, m2 = warp_sum(s2) / n;
  float4* o4 = reinterpret_cast<float4*>(dx) + (size_t)r * H4;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < H4)
      o4[c] = make_float4(rstd * (gv[i].x - m1 - xv[i].x * m2), rstd * (gv[i].y - m1 - xv[i].y * m2),
                          rstd * (gv[i].z - m1 - xv[i].z * m2), rstd * (gv[i].w - m1 - xv[i].w * m2));
  }
}

// LayerNorm backward, parameter gradients: partial[blk][2][H] over blocks of rows (each CTA: 32 columns x its row block, the row
// statistics come from ln_bwd_dx_kernel), summed in block order by ln_bwd_param_finish_kernel => deterministic
__global__ void __launch_bounds__(256)
ln_bwd_param_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float2* __restrict__ stats, int R, int H, int rows_per_blk,
                    float* __restrict__ partial) {
  __shared__ float pg[8][33], pb[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
  const int r0 = blockIdx.y * rows_per_blk, r1 = min(r0 + rows_per_blk, R);
  float ag = 0.f, ab = 0.f;
  if (c < H)
    for (int r = r0 + ty; r < r1; r += 8) {
      const float2 st = __ldg(stats + r);
      const float d = __ldg(dy + (size_t)r * H + c);
      ag = fmaf(d, (__ldg(x + (size_t)r * H + c) - st.x) * st.y, ag);
      ab += d;
    }
  pg[ty][tx] = ag; pb[ty][tx] = ab;
  __syncthreads();
  if (ty == 0 && c < H) {
    float g = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { g += pg[i][tx]; b += pb[i][tx]; }
    partial[((size_t)blockIdx.y * 2 + 0) * H + c] = g;
    partial[((size_t)blockIdx.y * 2 + 1) * H + c] = b;
  }
}

__global__ void ln_bwd_param_finish_kernel(const float* __restrict__ partial, int nblk, int H, float* __restrict__ dgamma,
                                           float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float g = 0.f, b = 0.f;
  for (int k = 0; k < nblk; ++k) { g += partial[((size_t)k * 2 + 0) * H + c]; b += partial[((size_t)k * 2 + 1) * H + c]; }
  dgamma[c] = g; dbeta[c] = b;
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_normalize_resize_pad(const float* img_chw, int h, int w, float* batch_nhwc, int b, int H, int W,
                                        int oh, int ow, const float* h_mean3, const float* h_std3, vbg_stream_t stream) {
  VBG_REQUIRE(img_chw && batch_nhwc && h_mean3 && h_std3 && aligned16(batch_nhwc), "vbg_normalize_resize_pad: null / unaligned pointer");
  VBG_REQUIRE(h > 0 && w > 0 && oh > 0 && ow > 0 && oh <= H && ow <= W && b >= 0,
              "vbg_normalize_resize_pad: bad geometry h=%d w=%d oh=%d ow=%d H=%d W=%d", h, w, oh, ow, H, W);
  dim3 blk(32, 8), grd(cdiv(ow, 32), cdiv(oh, 8));
  normalize_resize_kernel<false><<<grd, blk, 0, as_stream(stream)>>>(
      img_chw, h, w, batch_nhwc + (size_t)b * (H + 6) * (W + 6) * 4, H, W, oh, ow,
      make_float3(h_mean3[0], h_mean3[1], h_mean3[2]), make_float3(h_std3[0], h_std3[1], h_std3[2]));
  return check_launch("vbg_normalize_resize_pad");
}

extern "C" int vbg_normalize_resize_pad_batch(const float* imgs, int n, int h, int w, float* batch_nhwc, int b0, int H, int W,
                                              int oh, int ow, const float* h_mean3, const float* h_std3, vbg_stream_t stream) {
  VBG_REQUIRE(imgs && batch_nhwc && h_mean3 && h_std3 && aligned16(batch_nhwc), "vbg_normalize_resize_pad_batch: null / unaligned pointer");
  VBG_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0 && oh > 0 && ow > 0 && oh <= H && ow <= W && b0 >= 0,
              "vbg_normalize_resize_pad_batch: bad geometry n=%d h=%d w=%d oh=%d ow=%d H=%d W=%d", n, h, w, oh, ow, H, W);
  dim3 blk(32, 8), grd(cdiv(ow, 32), cdiv(oh, 8), n);
  normalize_resize_kernel<false><<<grd, blk, 0, as_stream(stream)>>>(
      imgs, h, w, batch_nhwc + (size_t)b0 * (H + 6) * (W + 6) * 4, H, W, oh, ow,
      make_float3(h_mean3[0], h_mean3[1], h_mean3[2]), make_float3(h_std3[0], h_std3[1], h_std3[2]));
  return check_launch("vbg_normalize_resize_pad_batch");
}

extern "C" int vbg_normalize_resize_pad_u8(const uint8_t* imgs_hwc, int n, int h, int w, float* batch_nhwc, int b0, int H, int W,
                                           int oh, int ow, const float* h_mean3, const float* h_std3, vbg_stream_t stream) {
  VBG_REQUIRE(imgs_hwc && batch_nhwc && h_mean3 && h_std3 && aligned16(batch_nhwc), "vbg_normalize_resize_pad_u8: null / unaligned pointer");
  VBG_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0 && oh > 0 && ow > 0 && oh <= H && ow <= W && b0 >= 0,
              "vbg_normalize_resize_pad_u8: bad geometry n=%d h=%d w=%d oh=%d ow=%d H=%d W=%d", n, h, w, oh, ow, H, W);
  dim3 blk(32, 8), grd(cdiv(ow, 32), cdiv(oh, 8), n);
  normalize_resize_kernel<true><<<grd, blk, 0, as_stream(stream)>>>(
      imgs_hwc, h, w, batch_nhwc + (size_t)b0 * (H + 6) * (W + 6) * 4, H, W, oh, ow,
      make_float3(h_mean3[0], h_mean3[1], h_mean3[2]), make_float3(h_std3[0], h_std3[1], h_std3[2]));
  return check_launch("vbg_normalize_resize_pad_u8");
}

extern "C" int vbg_decode_batch_u8(const uint8_t* arena, const int32_t* tab, int B, int max_oh, int max_ow, float* batch_nhwc, int H,
                                   int W, const float* h_mean3, const float* h_std3, vbg_stream_t stream) {
  VBG_REQUIRE(arena && tab && batch_nhwc && h_mean3 && h_std3 && aligned16(batch_nhwc), "vbg_decode_batch_u8: null / unaligned pointer");
  VBG_REQUIRE(B > 0 && B <= 65535 && max_oh > 0 && max_ow > 0 && max_oh <= H && max_ow <= W,
              "vbg_decode_batch_u8: bad geometry B=%d max_oh=%d max_ow=%d H=%d W=%d", B, max_oh, max_ow, H, W);
  dim3 blk(32, 8), grd(cdiv(max_ow, 32), cdiv(max_oh, 8), B);
  decode_batch_u8_kernel<<<grd, blk, 0, as_stream(stream)>>>(
      arena, tab, batch_nhwc, H, W, make_float3(h_mean3[0], h_mean3[1], h_mean3[2]), make_float3(h_std3[0], h_std3[1], h_std3[2]));
  return check_launch("vbg_decode_batch_u8");
}

extern "C" int vbg_stem_pack_weights(const float* w_oihw, int O, float* w_ohwi4, float* w_k256, vbg_stream_t stream) {
  VBG_REQUIRE(w_oihw && w_ohwi4 && w_k256 && O > 0, "vbg_stem_pack_weights: bad arguments");
  stem_pack_kernel<<<cdiv((long long)O * 256, 256), 256, 0, as_stream(stream)>>>(w_oihw, O, w_ohwi4, w_k256);
  return check_launch("vbg_stem_pack_weights");
}

extern "C" int vbg_maxpool3x3s2_x(const void* x, long long x_plane, int B, int H, int W, int C, void* y, long long y_plane,
                                  vbg_stream_t stream) {
  VBG_REQUIRE(B > 0 && H > 0 && W > 0 && C % 4 == 0 && fmt_ok(x, x_plane) && fmt_ok(y, y_plane), "vbg_maxpool3x3s2: bad arguments");
  int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  long long total = (long long)B * Ho * Wo * (C / 4);
  maxpool3x3s2_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(x, x_plane, B, H, W, C / 4, Ho, Wo, y, y_plane);
  return check_launch("vbg_maxpool3x3s2");
}

extern "C" int vbg_maxpool3x3s2(const float* x, int B, int H, int W, int C, float* y, vbg_stream_t stream) {
  return vbg_maxpool3x3s2_x(x, 0, B, H, W, C, y, 0, stream);
}

extern "C" int vbg_avgpool2x2_x(const void* x, long long x_plane, int B, int H, int W, int C, void* y, long long y_plane,
                                vbg_stream_t stream) {
  VBG_REQUIRE(B > 0 && H > 1 && W > 1 && C % 4 == 0 && fmt_ok(x, x_plane) && fmt_ok(y, y_plane), "vbg_avgpool2x2: bad arguments");
  long long total = (long long)B * (H / 2) * (W / 2) * (C / 4);
  avgpool2x2_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(x, x_plane, B, H, W, C / 4, y, y_plane);
  return check_launch("vbg_avgpool2x2");
}

extern "C" int vbg_avgpool2x2(const float* x, int B, int H, int W, int C, float* y, vbg_stream_t stream) {
  return vbg_avgpool2x2_x(x, 0, B, H, W, C, y, 0, stream);
}

extern "C" int vbg_bn_fold(const float* weight, const float* bias, const float* mean, const float* var, float eps, int C,
                           float* scale, float* shift, vbg_stream_t stream) {
  VBG_REQUIRE(weight && bias && mean && var && scale && shift && C > 0, "vbg_bn_fold: bad arguments");
  bn_fold_kernel<<<cdiv(C, 128), 128, 0, as_stream(stream)>>>(weight, bias, mean, var, eps, C, scale, shift);
  return check_launch("vbg_bn_fold");
}

extern "C" int vbg_repack_oihw_to_ohwi(const float* w, int O, int I, int H, int W, float* out, vbg_stream_t stream) {
  VBG_REQUIRE(w && out && O > 0 && I > 0 && H > 0 && W > 0, "vbg_repack_oihw_to_ohwi: bad arguments");
  long long total = (long long)O * I * H * W;
  repack_oihw_kernel<<<cdiv(total, 256), 256, 0, as_stream(stream)>>>(w, O, I, H, W, out);
  return check_launch("vbg_repack_oihw_to_ohwi");
}

extern "C" int vbg_upsample_split_nchw(const float* x, int B, int h, int w, int Ct, int up, int c_split, float* out1,
                                       float* out2, vbg_stream_t stream) {
  VBG_REQUIRE(x && out1 && out2 && B > 0 && h > 0 && w > 0 && Ct > 0 && up > 0 && c_split > 0 && c_split < Ct,
              "vbg_upsample_split_nchw: bad arguments");
  size_t smem = (size_t)w * Ct * sizeof(float);
  VBG_REQUIRE(smem <= 48 * 1024, "vbg_upsample_split_nchw: row of %zu bytes exceeds 48 KiB", smem);
  upsample_split_kernel<<<dim3(h, B), 256, smem, as_stream(stream)>>>(x, h, w, Ct, up, c_split, out1, out2);
  return check_launch("vbg_upsample_split_nchw");
}

extern "C" int vbg_nhwc_to_nchw(const float* x, int B, int H, int W, int C, float* y, vbg_stream_t stream) {
  VBG_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, "vbg_nhwc_to_nchw: bad arguments");
  dim3 grd(cdiv((long long)H * W, 32), cdiv(C, 32), B), blk(32, 8);
  nhwc_to_nchw_kernel<<<grd, blk, 0, as_stream(stream)>>>(x, H * W, C, y);
  return check_launch("vbg_nhwc_to_nchw");
}

extern "C" int vbg_transpose_split(const void* x, long long x_plane, int rows, int cols, void* out_hi, long long out_plane, int ld_out,
                                   vbg_stream_t stream) {
  VBG_REQUIRE(x && out_hi && rows > 0 && cols > 0 && ld_out >= rows && x_plane >= 0 && out_plane > 0, "vbg_transpose_split: bad arguments");
  dim3 grd(cdiv(cols, 32), cdiv(ld_out, 32));
  transpose_split_kernel<<<grd, 256, 0, as_stream(stream)>>>(x, x_plane, rows, cols, reinterpret_cast<__nv_bfloat16*>(out_hi), out_plane, ld_out);
  return check_launch("vbg_transpose_split");
}

extern "C" int vbg_conv_dgrad_weight(const float* w_ohwi, int Cout, int kh, int kw, int Cin, void* out_hi, long long out_plane,
                                     vbg_stream_t stream) {
  VBG_REQUIRE(w_ohwi && out_hi && Cout > 0 && kh > 0 && kw > 0 && Cin > 0 && out_plane > 0, "vbg_conv_dgrad_weight: bad arguments");
  const long long total = (long long)Cin * kh * kw * Cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  conv_dgrad_weight_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w_ohwi, Cout, kh, kw, Cin, reinterpret_cast<__nv_bfloat16*>(out_hi), out_plane);
  return check_launch("vbg_conv_dgrad_weight");
}

extern "C" int vbg_layernorm_bwd(const float* x, const float* dy, const float* gamma, float eps, int R, int hidden, float* dx,
                                 float* dgamma, float* dbeta, float* workspace, size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(x && dy && gamma && dx && R > 0 && hidden % 4 == 0 && hidden <= 1024 && aligned16(x) && aligned16(dy) && aligned16(dx),
              "vbg_layernorm_bwd: hidden %% 4 == 0, <= 1024, 16B alignment");
  cudaStream_t s = as_stream(stream);
  // workspace = [nblk][2][hidden] partial sums, then [R] (mean, rstd) pairs
  const int rows_per_blk = 64, nblk = cdiv(R, rows_per_blk);
  const size_t need = ((size_t)nblk * 2 * hidden + 2 * (size_t)R) * sizeof(float);
  if (dgamma) {
    VBG_REQUIRE(dbeta && workspace && aligned16(workspace), "vbg_layernorm_bwd: dbeta and a 16B-aligned workspace required with dgamma");
    if (need > ws_bytes) { set_error("vbg_layernorm_bwd: workspace of %zu bytes needed", need); return VBG_EWORKSPACE; }
  }
  float2* stats = dgamma ? reinterpret_cast<float2*>(workspace + (size_t)nblk * 2 * hidden) : nullptr;
  ln_bwd_dx_kernel<<<cdiv(R, 8), 256, 0, s>>>(x, dy, gamma, eps, R, hidden / 4, dx, stats);
  int rc = check_launch("vbg_layernorm_bwd(dx)");
  if (rc || !dgamma) return rc;
  ln_bwd_param_kernel<<<dim3(cdiv(hidden, 32), nblk), 256, 0, s>>>(x, dy, stats, R, hidden, rows_per_blk, workspace);
  rc = check_launch("vbg_layernorm_bwd(params)");
  if (rc) return rc;
  ln_bwd_param_finish_kernel<<<cdiv(hidden, 128), 128, 0, s>>>(workspace, nblk, hidden, dgamma, dbeta);
  return check_launch("vbg_layernorm_bwd(finish)");
}
