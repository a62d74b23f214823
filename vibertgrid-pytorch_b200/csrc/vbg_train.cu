// Memory-bound kernels of the training step (reference: loss.backward() in pipeline/train_val_utils.py:277 through the modules
// of SURVEY.md section 8a).  fp32, channels-last.  Every reduction is two-stage with a fixed summation order (deterministic);
// the only atomics are the scatter-adds of vbg_roi_align_bwd and vbg_embed_bwd (the same places torch uses them).
//
//   batch-norm (training)  model/ResNetFPN_ViBERTgrid.py:106-186 (nn.BatchNorm2d in train mode): batch statistics, normalise
//                          (+ residual, + ReLU), and the three backward pieces (two column reductions + one elementwise pass)
//   max-pool 3x3/2 bwd     first-maximum rule of torch's max_pool2d (strict '>' scan in (h, w) order)
//   2x2 sum-pool / nearest-x2 broadcast: the backward of the FPN's nearest-x2 upsample-add and of the D-variant's avg-pool
//   zero insertion         dY of a stride-2 convolution spread onto the stride-1 lattice (its data gradient is then a stride-1 conv)
//   GELU fwd / bwd, dropout (counter-based mask, same call forward and backward)
//   BERTgrid scatter bwd, segment-mean bwd, embedding-table scatter-add
//   ROI-align bwd          same sampling arithmetic as vbg_roi.cu::roi_align_kernel
//   auxiliary CE bwd       gradient of the two mean cross entropies w.r.t. the LOW-resolution logits (nearest upsampling folded in)
//   small-N weight gradient  dW[N<=16, K] = dY^T X for the heads the tensor-core wgrad kernel does not take
//   stem weight gradient   7x7/2 convolution over the zero-bordered NHWC4 image batch
#include "vbg_common.cuh"

namespace vbg {

static inline int grid_for(long long n, int per_block, int max_blocks = kNumSMs * 8) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  return (int)(b > max_blocks ? max_blocks : b);
}

// ------------------------------------------------------------------ column reductions over [rows, C]
// A CTA of 256 threads is (256 / C4) row lanes x C4 float4 channel lanes (C4 = C / 4 divides 256).  Each CTA reduces a
// contiguous row chunk into partial[cta][2][C]; a finish kernel sums the partials per channel in double, fixed order.
struct BnStatsF {   // sum x, sum x^2
  const float4* x;
  __device__ __forceinline__ void operator()(size_t i, int, float4& a, float4& b) const {
    const float4 v = __ldg(x + i);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y); b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
  }
};
struct BnBwdF {     // sum dy', sum dy' * xhat   with dy' = dy * (y > 0) when y is given
  const float4 *x, *dy, *y, *mean, *rstd;
  __device__ __forceinline__ void operator()(size_t i, int c, float4& a, float4& b) const {
    float4 g = __ldg(dy + i);
    if (y) { const float4 o = __ldg(y + i); g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f; }
    const float4 v = __ldg(x + i), m = __ldg(mean + c), r = __ldg(rstd + c);
    a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
    b.x = fmaf(g.x, (v.x - m.x) * r.x, b.x); b.y = fmaf(g.y, (v.y - m.y) * r.y, b.y);
    b.z = fmaf(g.z, (v.z - m.z) * r.z, b.z); b.w = fmaf(g.w, (v.w - m.w) * r.w, b.w);
  }
};

template <class F>
__global__ void __launch_bounds__(256) colreduce2_kernel(F f, long long rows, int C4, long long rows_per_cta, float* __restrict__ partial) {
  __shared__ float4 sh[2][256];
  const int c = threadIdx.x % C4, rl = threadIdx.x / C4, RL = 256 / C4;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta; if (r1 > rows) r1 = rows;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll 4
  for (long long r = r0 + rl; r < r1; r += RL) f((size_t)r * C4 + c, c, a, b);
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = b;
  __syncthreads();
  if (rl == 0) {
    for (int i = 1; i < RL; ++i) {
      const float4 u = sh[0][i * C4 + c], v = sh[1][i * C4 + c];
      a.x += u.x; a.y += u.y; a.z += u.z; a.w += u.w;
      b.x += v.x; b.y += v.y; b.z += v.z; b.w += v.w;
    }
    float4* p = reinterpret_cast<float4*>(partial) + (size_t)blockIdx.x * 2 * C4;
    p[c] = a; p[C4 + c] = b;
  }
}

// mode 0: (mean, biased var, rstd) from (sum, sumsq); mode 1: plain sums (dbeta, dgamma).  One warp per channel: lanes stride the
// partials, double accumulators, fixed-order shuffle tree.
__global__ void colreduce2_finish_kernel(const float* __restrict__ partial, int nblk, int C, double inv_rows, float eps, int mode,
                                         float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  double a = 0.0, b = 0.0;
  for (int i = lane; i < nblk; i += 32) { a += (double)partial[(size_t)i * 2 * C + c]; b += (double)partial[(size_t)i * 2 * C + C + c]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  if (lane) return;
  if (mode == 0) {
    const double m = a * inv_rows;
    double v = b * inv_rows - m * m; if (v < 0.0) v = 0.0;
    o0[c] = (float)m; o1[c] = (float)v; o2[c] = (float)(1.0 / sqrt(v + (double)eps));
  } else {
    o0[c] = (float)a; o1[c] = (float)b;
  }
}

static int colreduce_geometry(long long rows, int C, long long& rows_per_cta) {
  const int RL = 256 / (C / 4);
  long long want = (rows + kNumSMs * 4 - 1) / (kNumSMs * 4);
  long long rpc = ((want + RL - 1) / RL) * RL; if (rpc < 4 * RL) rpc = 4 * RL;
  rows_per_cta = rpc;
  return (int)((rows + rpc - 1) / rpc);
}

__global__ void bn_apply_kernel(const float4* __restrict__ x, long long n4, int C4, const float4* __restrict__ mean, const float4* __restrict__ rstd,
                                const float4* __restrict__ gamma, const float4* __restrict__ beta, const float4* __restrict__ res, int relu,
                                float4* __restrict__ y) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    const float4 v = __ldg(x + i), m = __ldg(mean + c), r = __ldg(rstd + c), g = __ldg(gamma + c), b = __ldg(beta + c);
    float4 o;
    o.x = fmaf((v.x - m.x) * r.x, g.x, b.x); o.y = fmaf((v.y - m.y) * r.y, g.y, b.y);
    o.z = fmaf((v.z - m.z) * r.z, g.z, b.z); o.w = fmaf((v.w - m.w) * r.w, g.w, b.w);
    if (res) { const float4 q = __ldg(res + i); o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w; }
    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    y[i] = o;
  }
}

// dx = gamma * rstd * (dy' - dbeta / n - xhat * dgamma / n);  dres = dy' (the residual branch's gradient) when asked
__global__ void bn_bwd_dx_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const float4* __restrict__ y, long long n4, int C4,
                                 float inv_rows, const float4* __restrict__ mean, const float4* __restrict__ rstd, const float4* __restrict__ gamma,
                                 const float4* __restrict__ dgamma, const float4* __restrict__ dbeta, float4* __restrict__ dx,
                                 float4* __restrict__ dres) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    float4 g = __ldg(dy + i);
    if (y) { const float4 o = __ldg(y + i); g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f; }
    if (dres) dres[i] = g;
    const float4 v = __ldg(x + i), m = __ldg(mean + c), r = __ldg(rstd + c), ga = __ldg(gamma + c), dg = __ldg(dgamma + c), db = __ldg(dbeta + c);
    float4 o;
    o.x = ga.x * r.x * (g.x - db.x * inv_rows - (v.x - m.x) * r.x * dg.x * inv_rows);
    o.y = ga.y * r.y * (g.y - db.y * inv_rows - (v.y - m.y) * r.y * dg.y * inv_rows);
    o.z = ga.z * r.z * (g.z - db.z * inv_rows - (v.z - m.z) * r.z * dg.z * inv_rows);
    o.w = ga.w * r.w * (g.w - db.w * inv_rows - (v.w - m.w) * r.w * dg.w * inv_rows);
    dx[i] = o;
  }
}

// ------------------------------------------------------------------ pooling / resampling backward
// dX of max_pool2d(3, 2, 1): every input pixel looks at the (up to 2 x 2) windows that contain it and takes dY of those whose
// FIRST maximum (scan order r, s; strict '>') it is -- a gather, so no atomics and ties resolve exactly like torch.
__global__ void maxpool3x3s2_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, int B, int H, int W, int C4, int Ho, int Wo,
                                        float4* __restrict__ dx) {
  const long long total = (long long)B * H * W * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4); long long t = i / C4;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H); const int b = (int)(t / H);
    const float4 self = __ldg(x + i);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ho0 = h >> 1, ho1 = (h + 1) >> 1, wo0 = w >> 1, wo1 = (w + 1) >> 1;    // windows 2*ho-1 .. 2*ho+1 containing h
    for (int ho = ho0; ho <= ho1; ++ho) {
      if (ho >= Ho) continue;
      for (int wo = wo0; wo <= wo1; ++wo) {
        if (wo >= Wo) continue;
        // does (h, w) hold the first maximum of window (ho, wo)?  per channel: no earlier element >= self, no later element > self
        bool kx = true, ky = true, kz = true, kw = true;
        for (int r = 0; r < 3; ++r) {
          const int hh = 2 * ho - 1 + r;
          if (hh < 0 || hh >= H) continue;
          for (int s = 0; s < 3; ++s) {
            const int ww = 2 * wo - 1 + s;
            if (ww < 0 || ww >= W || (hh == h && ww == w)) continue;
            const float4 v = __ldg(x + (((size_t)b * H + hh) * W + ww) * C4 + c);
            const bool earlier = hh < h || (hh == h && ww < w);
            if (earlier) { kx &= !(v.x >= self.x); ky &= !(v.y >= self.y); kz &= !(v.z >= self.z); kw &= !(v.w >= self.w); }
            else         { kx &= !(v.x > self.x);  ky &= !(v.y > self.y);  kz &= !(v.z > self.z);  kw &= !(v.w > self.w); }
          }
        }
        const float4 g = __ldg(dy + (((size_t)b * Ho + ho) * Wo + wo) * C4 + c);
        if (kx) acc.x += g.x; if (ky) acc.y += g.y; if (kz) acc.z += g.z; if (kw) acc.w += g.w;
      }
    }
    dx[i] = acc;
  }
}

// The same gather with the pooled OUTPUT at hand: (h, w) can hold a window's maximum only where x == y, so the window is
// scanned (earlier positions only: is there an earlier element that also equals the maximum?) just for those channels --
// ~14 loads per input pixel instead of ~40.  Same first-maximum rule, same results.
__global__ void maxpool3x3s2_bwd_y_kernel(const float4* __restrict__ x, const float4* __restrict__ y, const float4* __restrict__ dy, int B,
                                          int H, int W, int C4, int Ho, int Wo, float4* __restrict__ dx) {
  const long long total = (long long)B * H * W * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4); long long t = i / C4;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H); const int b = (int)(t / H);
    const float4 self = __ldg(x + i);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ho0 = h >> 1, ho1 = (h + 1) >> 1, wo0 = w >> 1, wo1 = (w + 1) >> 1;
    for (int ho = ho0; ho <= ho1; ++ho) {
      if (ho >= Ho) continue;
      for (int wo = wo0; wo <= wo1; ++wo) {
        if (wo >= Wo) continue;
        const size_t o = (((size_t)b * Ho + ho) * Wo + wo) * C4 + c;
        const float4 m = __ldg(y + o);
        bool kx = self.x == m.x, ky = self.y == m.y, kz = self.z == m.z, kw = self.w == m.w;
        if (!(kx | ky | kz | kw)) continue;
        // an EARLIER element of the window (scan order r, s) equal to the maximum takes the gradient instead
        for (int r = 0; r < 3; ++r) {
          const int hh = 2 * ho - 1 + r;
          if (hh < 0 || hh > h) continue;
          for (int s2 = 0; s2 < 3; ++s2) {
            const int ww = 2 * wo - 1 + s2;
            if (ww < 0 || ww >= W || !(hh < h || ww < w)) continue;
            const float4 v = __ldg(x + (((size_t)b * H + hh) * W + ww) * C4 + c);
            kx &= !(v.x == m.x); ky &= !(v.y == m.y); kz &= !(v.z == m.z); kw &= !(v.w == m.w);
          }
        }
        const float4 g = __ldg(dy + o);
        if (kx) acc.x += g.x; if (ky) acc.y += g.y; if (kz) acc.z += g.z; if (kw) acc.w += g.w;
      }
    }
    dx[i] = acc;
  }
}

// y[b, h, w] = scale * sum of the 2x2 block of x (H, W even)
__global__ void sumpool2x2_kernel(const float4* __restrict__ x, int B, int H, int W, int C4, float scale, float4* __restrict__ y) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)B * Ho * Wo * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4); long long t = i / C4;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho); const int b = (int)(t / Ho);
    const float4* p = x + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C4 + c;
    const float4 a = __ldg(p), bb = __ldg(p + C4), cc = __ldg(p + (size_t)W * C4), d = __ldg(p + (size_t)W * C4 + C4);
    y[i] = make_float4(scale * ((a.x + bb.x) + (cc.x + d.x)), scale * ((a.y + bb.y) + (cc.y + d.y)), scale * ((a.z + bb.z) + (cc.z + d.z)),
                       scale * ((a.w + bb.w) + (cc.w + d.w)));
  }
}

// y[b, h, w] = scale * x[b, h/2, w/2] when `every` == 0 (nearest x2), or x[b, h/2, w/2] on even (h, w) and 0 elsewhere (`every` == 1:
// zero insertion); output H x W, input ceil(H/2) x ceil(W/2)
__global__ void expand2x_kernel(const float4* __restrict__ x, int B, int H, int W, int C4, int Hi, int Wi, float scale, int zero_insert,
                                float4* __restrict__ y) {
  const long long total = (long long)B * H * W * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4); long long t = i / C4;
    const int w = (int)(t % W); t /= W;
    const int h = (int)(t % H); const int b = (int)(t / H);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    const int hi = h >> 1, wi = w >> 1;
    if (hi < Hi && wi < Wi && !(zero_insert && ((h | w) & 1))) {
      const float4 v = __ldg(x + (((size_t)b * Hi + hi) * Wi + wi) * C4 + c);
      o = make_float4(scale * v.x, scale * v.y, scale * v.z, scale * v.w);
    }
    y[i] = o;
  }
}

// ------------------------------------------------------------------ elementwise
// x, dy and out in either storage format (plane == 0: fp32; > 0: bf16 hi/lo planes `plane` elements apart): the FFN of the
// training step keeps its [rows, 3072] activations and gradients in the plane format end to end (autograd.py "planes protocol")
__global__ void gelu_kernel(const void* __restrict__ x, long long x_plane, const void* __restrict__ dy, long long dy_plane, long long n4,
                            void* __restrict__ out, long long out_plane) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = ld4_fmt(x, x_plane, (size_t)i);
    float4 o;
    if (!dy) {
      o = make_float4(gelu_erf(v.x), gelu_erf(v.y), gelu_erf(v.z), gelu_erf(v.w));
    } else {
      const float4 g = ld4_fmt(dy, dy_plane, (size_t)i);
      auto d = [](float u) { return 0.5f * (1.0f + erff(u * 0.70710678118654752440f)) + u * 0.3989422804014327f * expf(-0.5f * u * u); };
      o = make_float4(g.x * d(v.x), g.y * d(v.y), g.z * d(v.z), g.w * d(v.w));
    }
    st4_fmt(out, out_plane, (size_t)i, o);
  }
}

__device__ __forceinline__ float keep_of(unsigned long long seed, unsigned long long i, float p) {
  unsigned long long z = seed + (i + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return ((float)(z >> 40) * (1.0f / 16777216.0f)) >= p ? 1.0f : 0.0f;
}
// y = x * keep(seed, index) / (1 - p): the same call regenerates the mask for the backward
// `step_seed` (optional, DEVICE memory) is folded into the by-value seed: a captured CUDA graph bakes the by-value seed of
// each call site, the host refreshes the one device word before every replay, so every step draws a new mask
__global__ void dropout_kernel(const float* __restrict__ x, long long n, float p, unsigned long long seed,
                               const unsigned long long* __restrict__ step_seed, float* __restrict__ y) {
  if (step_seed) seed ^= __ldg(step_seed) * 0xD6E8FEB86659FD93ull;
  const float s = 1.0f / (1.0f - p);
  if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {     // same mask, 16-byte accesses
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += (long long)gridDim.x * blockDim.x) {
      float4 v = __ldg(x4 + i);
      const unsigned long long e = 4ull * (unsigned long long)i;
      v.x = v.x * keep_of(seed, e, p) * s; v.y = v.y * keep_of(seed, e + 1, p) * s;
      v.z = v.z * keep_of(seed, e + 2, p) * s; v.w = v.w * keep_of(seed, e + 3, p) * s;
      y4[i] = v;
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = __ldg(x + i) * keep_of(seed, (unsigned long long)i, p) * s;
}

// ------------------------------------------------------------------ BERTgrid backward
// One CTA per segment: sum dgrid over the cells of its box that it WON (index map), cell order, then channels across threads.
__global__ void __launch_bounds__(256)
grid_scatter_bwd_kernel(const float* __restrict__ dgrid, long long ld, const int32_t* __restrict__ idx, const int32_t* __restrict__ boxes,
                        const int32_t* __restrict__ seg_off, int B, int stride, int Hg, int Wg, int C4, float* __restrict__ demb) {
  const int k = blockIdx.x;
  const int b = sample_of(seg_off, B, k);
  const int local = k - __ldg(seg_off + b);
  const int4 c = __ldg(reinterpret_cast<const int4*>(boxes) + k);
  const int x1 = py_slice_bound(c.x / stride, Wg), y1 = py_slice_bound(c.y / stride, Hg);
  const int x2 = py_slice_bound(c.z / stride, Wg), y2 = py_slice_bound(c.w / stride, Hg);
  for (int ch = threadIdx.x; ch < C4; ch += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = y1; y < y2; ++y)
      for (int x = x1; x < x2; ++x) {
        const size_t cell = ((size_t)b * Hg + y) * Wg + x;
        if (__ldg(idx + cell) != local) continue;
        const float4 v = __ldg(reinterpret_cast<const float4*>(dgrid + cell * ld) + ch);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    reinterpret_cast<float4*>(demb)[(size_t)k * C4 + ch] = acc;
  }
}

// dhidden rows of a segment's tokens = dseg / n (mean) or dseg on the first token (first); rows of no segment stay as the caller
// initialised them (zero)
__global__ void segment_reduce_bwd_kernel(const float* __restrict__ dseg, const int32_t* __restrict__ tok_row, const int32_t* __restrict__ seg_start,
                                          int C4, int mode, float* __restrict__ dhidden) {
  const int k = blockIdx.x;
  const int a = seg_start[k], e = seg_start[k + 1];
  if (e <= a) return;
  const float inv = mode == VBG_AGG_MEAN ? 1.0f / (float)(e - a) : 1.0f;
  const int last = mode == VBG_AGG_MEAN ? e : a + 1;
  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    float4 g = __ldg(reinterpret_cast<const float4*>(dseg) + (size_t)k * C4 + c);
    g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
    for (int t = a; t < last; ++t) reinterpret_cast<float4*>(dhidden)[(size_t)tok_row[t] * C4 + c] = g;
  }
}

// d(word table)[ids[r]] += dx[r], d(position table)[pos[r]] += dx[r]   (tables zeroed by the caller).  DETERMINISTIC, no atomics:
// the CTA of row r owns table row key = idx[r] iff r is the FIRST packed row with that key (a scan of idx[0, r)); the owner then
// sums dx[j] over every j >= r with idx[j] == key in ascending j (matches compacted in order by warp ballots), and stores the row.
// blockIdx.y = 0: word table through ids, 1: position table through pos.  O(R^2) integer compares per table -- R is a few
// thousand packed rows -- against R * H floats of real traffic.
__global__ void __launch_bounds__(256)
embed_bwd_kernel(const float* __restrict__ dx, const int32_t* __restrict__ ids, const int32_t* __restrict__ pos, int R, int H,
                 float* __restrict__ dword, float* __restrict__ dpos) {
  const int32_t* __restrict__ idx = blockIdx.y ? pos : ids;
  float* __restrict__ out = blockIdx.y ? dpos : dword;
  const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int key = __ldg(idx + r);
  int seen = 0;
  for (int j = tid; j < r; j += blockDim.x) seen |= (__ldg(idx + j) == key);
  if (__syncthreads_or(seen)) return;                      // an earlier row owns this table row
  __shared__ int list[256];
  __shared__ int wcnt[8];
  const int H4 = H >> 2;
  float4 acc[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};       // H <= 2048: two float4 per thread
  for (int base = r; base < R; base += blockDim.x) {
    const int j = base + tid;
    const bool hit = j < R && __ldg(idx + j) == key;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nwarp; ++w) { const int c = wcnt[w]; if (w < warp) before += c; total += c; }
    if (hit) list[before + __popc(bal & ((1u << lane) - 1u))] = j;
    __syncthreads();
    for (int e = 0; e < total; ++e) {
      const float4* src = reinterpret_cast<const float4*>(dx + (size_t)list[e] * H);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c = tid + u * blockDim.x;
        if (c < H4) { const float4 v = __ldg(src + c); acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w; }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c = tid + u * blockDim.x;
    if (c < H4) reinterpret_cast<float4*>(out + (size_t)key * H)[c] = acc[u];
  }
}

// ------------------------------------------------------------------ ROI-align backward (deterministic gather form)
// dfeat[b, y, x, :] = sum over the ROIs k of document b (ascending k), bins (ph, pw):  Wy_k[ph](y) * Wx_k[pw](x) / count_k * dout[k, ph, pw, :]
// where Wy_k[ph](y) = the summed bilinear row weights of bin ph's samples on feature row y (torchvision's sampling rule, the
// forward's arithmetic operation for operation).  A CTA owns an 8 x 8 tile of feature pixels of one document: it lists the ROIs
// whose sample window meets the tile (ascending k), builds for each the 7 x 8 row weights and 7 x 8 column weights of ITS rows /
// columns straight from the sample coordinates (only the few samples near a row can touch it, so no per-ROI table of bounded
// span is needed: any ROI size takes this path), and accumulates -- every pixel is written exactly once: no atomics, no
// zero-fill, bitwise reproducible.  (torchvision's CUDA backward, like round 1's kernel, scatters with atomics.)
struct RoiGeo { float sh, sw, bh, bw, inv_count; int gh, gw, y_lo, y_hi, x_lo, x_hi, pad; };

__global__ void roi_geo_kernel(const int32_t* __restrict__ boxes, int K, float scale, int P, int Hf, int Wf, RoiGeo* __restrict__ geo) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const int4 bx = __ldg(reinterpret_cast<const int4*>(boxes) + k);
  RoiGeo g;
  g.sw = __fmul_rn((float)bx.x, scale); g.sh = __fmul_rn((float)bx.y, scale);
  const float ew = __fmul_rn((float)bx.z, scale), eh = __fmul_rn((float)bx.w, scale);
  const float rw = fmaxf(__fsub_rn(ew, g.sw), 1.0f), rh = fmaxf(__fsub_rn(eh, g.sh), 1.0f);
  g.bw = __fdiv_rn(rw, (float)P); g.bh = __fdiv_rn(rh, (float)P);
  g.gh = (int)ceilf(__fdiv_rn(rh, (float)P));
  g.gw = (int)ceilf(__fdiv_rn(rw, (float)P));
  g.inv_count = 1.0f / (float)max(g.gh * g.gw, 1);
  // rows / columns any sample can touch (first and last sample coordinate of each axis; the +1 row of the bilinear pair)
  const float y_first = __fadd_rn(g.sh, __fdiv_rn(__fmul_rn(0.5f, g.bh), (float)g.gh));
  const float y_last = __fadd_rn(__fadd_rn(g.sh, __fmul_rn((float)(P - 1), g.bh)), __fdiv_rn(__fmul_rn((float)g.gh - 0.5f, g.bh), (float)g.gh));
  const float x_first = __fadd_rn(g.sw, __fdiv_rn(__fmul_rn(0.5f, g.bw), (float)g.gw));
  const float x_last = __fadd_rn(__fadd_rn(g.sw, __fmul_rn((float)(P - 1), g.bw)), __fdiv_rn(__fmul_rn((float)g.gw - 0.5f, g.bw), (float)g.gw));
  g.y_lo = min(max((int)fmaxf(y_first, 0.f), 0), Hf - 1); g.y_hi = min(max((int)fmaxf(y_last, 0.f) + 1, g.y_lo), Hf - 1);
  g.x_lo = min(max((int)fmaxf(x_first, 0.f), 0), Wf - 1); g.x_hi = min(max((int)fmaxf(x_last, 0.f) + 1, g.x_lo), Wf - 1);
  g.pad = 0;
  geo[k] = g;
}

// summed weight of bin `pb`'s samples on row / column `t` of an axis of size `dim` (start, bin size, samples per bin as in the forward)
__device__ __forceinline__ float roi_axis_weight(float start, float bin, int g, int pb, int t, int dim) {
  const float s0 = __fadd_rn(start, __fmul_rn((float)pb, bin));
  // candidate samples: coordinates within (t - 1, t + 1), widened for the clamped edges and by a safety margin of two samples
  const float inv_step = (float)g / bin;
  const float lo_c = (t == 0 ? -1.5f : (float)t - 1.0f), hi_c = (t == dim - 1 ? (float)dim + 0.5f : (float)t + 1.0f);
  int i0 = (int)floorf((lo_c - s0) * inv_step - 0.5f) - 2, i1 = (int)ceilf((hi_c - s0) * inv_step - 0.5f) + 2;
  i0 = max(i0, 0); i1 = min(i1, g - 1);
  float w = 0.f;
  for (int i = i0; i <= i1; ++i) {
    float c = __fadd_rn(s0, __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)g));
    if (c < -1.0f || c > (float)dim) continue;
    c = fmaxf(c, 0.f);
    int lo = (int)c, hi;
    if (lo >= dim - 1) { hi = lo = dim - 1; c = (float)lo; } else { hi = lo + 1; }
    const float l = c - (float)lo, h = 1.f - l;
    if (lo == t) w += h;
    if (hi == t) w += l;
  }
  return w;
}

constexpr int kRbT = 8;          // tile edge (feature pixels)
constexpr int kRbMax = 32;       // ROIs processed per round of a tile
__global__ void __launch_bounds__(256)
roi_align_bwd_kernel(const float* __restrict__ dout, int B, int Hf, int Wf, int C, const int32_t* __restrict__ seg_off, const RoiGeo* __restrict__ geo,
                     int P, float* __restrict__ dfeat) {
  __shared__ int list[kRbMax];
  __shared__ int n_list, next_k;
  __shared__ float wy[kRbMax][8][kRbT], wx[kRbMax][8][kRbT];       // [roi][bin][tile row / column]; bin index 7 unused
  const int b = blockIdx.z, ty0 = blockIdx.y * kRbT, tx0 = blockIdx.x * kRbT, tid = threadIdx.x;
  const int k_begin = __ldg(seg_off + b), k_end = __ldg(seg_off + b + 1);
  const int C4 = C >> 2;
  const int quads_per_round = 256 / 4;                    // 64 channel quads x 4 pixel groups
  const int pg = tid >> 6, q0 = tid & 63;
  // each thread owns pixels pg, pg + 4, ... of the tile (16 pixels) for channel quads q0, q0 + 64, ... : accumulators in registers
  // for C <= 256 (one quad per thread); larger C loops the whole tile again per 256-channel slab
  for (int cbase = 0; cbase < C4; cbase += quads_per_round) {
    const int cq = cbase + q0;
    float4 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) next_k = k_begin;
    __syncthreads();
    while (true) {
      // ---- list up to kRbMax ROIs (ascending k) whose window meets the tile: one warp scans with ballots
      if (tid < 32) {
        int cnt = 0, k = next_k;
        while (k < k_end && cnt < kRbMax) {
          const int kk = k + tid;
          bool hit = false;
          if (kk < k_end) {
            const RoiGeo g = geo[kk];
            hit = g.y_lo < ty0 + kRbT && g.y_hi >= ty0 && g.x_lo < tx0 + kRbT && g.x_hi >= tx0;
          }
          const unsigned bal = __ballot_sync(0xffffffffu, hit);
          const int room = kRbMax - cnt;
          const int npop = __popc(bal);
          if (npop <= room) {
            if (hit) list[cnt + __popc(bal & ((1u << tid) - 1u))] = kk;
            cnt += npop; k += 32;
          } else {                                   // take the first `room` hits only; resume after the last one taken
            const int rank = __popc(bal & ((1u << tid) - 1u));
            if (hit && rank < room) list[cnt + rank] = kk;
            int last = 0;
            for (int l = 0, seen = 0; l < 32; ++l) if ((bal >> l) & 1u) { if (++seen == room) { last = l; break; } }
            cnt += room; k += last + 1;
          }
        }
        __syncwarp();                                  // every lane has read next_k (a zero-trip scan has no ballot to converge on)
        if (tid == 0) { n_list = cnt; next_k = k; }
      }
      __syncthreads();
      const int n = n_list;
      if (n == 0) break;
      // ---- row / column weights of the listed ROIs for this tile's rows and columns
      for (int item = tid; item < n * 2 * 7 * kRbT; item += 256) {
        const int t = item % kRbT, pb = (item / kRbT) % 7, axis = (item / (kRbT * 7)) & 1, li = item / (kRbT * 7 * 2);
        const RoiGeo g = geo[list[li]];
        if (axis) wy[li][pb][t] = ty0 + t < Hf ? roi_axis_weight(g.sh, g.bh, g.gh, pb, ty0 + t, Hf) * g.inv_count : 0.f;
        else wx[li][pb][t] = tx0 + t < Wf ? roi_axis_weight(g.sw, g.bw, g.gw, pb, tx0 + t, Wf) : 0.f;
      }
      __syncthreads();
      // ---- accumulate
      if (cq < C4) {
        for (int li = 0; li < n; ++li) {
          const float4* g4 = reinterpret_cast<const float4*>(dout) + (size_t)list[li] * P * P * C4 + cq;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int pix = pg + 4 * i, ty = pix >> 3, tx = pix & 7;
#pragma unroll 1
            for (int ph = 0; ph < 7; ++ph) {
              const float a = wy[li][ph][ty];
              if (a == 0.f) continue;
#pragma unroll 1
              for (int pw = 0; pw < 7; ++pw) {
                const float w = a * wx[li][pw][tx];
                if (w == 0.f) continue;
                const float4 v = __ldg(g4 + (size_t)(ph * 7 + pw) * C4);
                acc[i].x = fmaf(w, v.x, acc[i].x); acc[i].y = fmaf(w, v.y, acc[i].y);
                acc[i].z = fmaf(w, v.z, acc[i].z); acc[i].w = fmaf(w, v.w, acc[i].w);
              }
            }
          }
        }
      }
      __syncthreads();
      if (n < kRbMax) break;
    }
    if (cq < C4) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int pix = pg + 4 * i, y = ty0 + (pix >> 3), x = tx0 + (pix & 7);
        if (y < Hf && x < Wf) reinterpret_cast<float4*>(dfeat)[(((size_t)b * Hf + y) * Wf + x) * C4 + cq] = acc[i];
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ auxiliary cross entropy backward
// loss = g[0] * mean_px CE(mask logits, pos_neg) + g[1] * mean_px CE(class logits, cls) over the FULL-resolution pixels, whose
// logits are the nearest-upsampled low-resolution ones: d/dz(cell) = sum over the cell's up x up pixels of (softmax - onehot) / n_px.
// One thread per low-resolution cell.
__global__ void seg_ce_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ pos_neg, const long long* __restrict__ cls, int B, int H,
                                  int W, int up, int Ct, int c_split, const float* __restrict__ gscale, float inv_n, float* __restrict__ dlogits) {
  const int h = H / up, w = W / up;
  const long long total = (long long)B * h * w;
  const float g1 = __ldg(gscale) * inv_n, g2 = __ldg(gscale + 1) * inv_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xw = (int)(i % w); long long t = i / w;
    const int yh = (int)(t % h); const int b = (int)(t / h);
    const float* z = logits + (size_t)i * Ct;
    float m1 = -INFINITY, m2 = -INFINITY;
    for (int c = 0; c < c_split; ++c) m1 = fmaxf(m1, __ldg(z + c));
    for (int c = c_split; c < Ct; ++c) m2 = fmaxf(m2, __ldg(z + c));
    float e1 = 0.f, e2 = 0.f;
    for (int c = 0; c < c_split; ++c) e1 += expf(__ldg(z + c) - m1);
    for (int c = c_split; c < Ct; ++c) e2 += expf(__ldg(z + c) - m2);
    const float n_px = (float)(up * up);
    float* d = dlogits + (size_t)i * Ct;
    for (int c = 0; c < Ct; ++c) {
      const bool first = c < c_split;
      const float p = first ? expf(__ldg(z + c) - m1) / e1 : expf(__ldg(z + c) - m2) / e2;
      int hits = 0;
      for (int dy = 0; dy < up; ++dy)
        for (int dx = 0; dx < up; ++dx) {
          const size_t px = ((size_t)b * H + (size_t)yh * up + dy) * W + (size_t)xw * up + dx;
          hits += first ? (__ldg(pos_neg + px) == c) : (__ldg(cls + px) == c - c_split);
        }
      d[c] = (first ? g1 : g2) * (n_px * p - (float)hits);
    }
  }
}

// d(logits)[b, h, w, c] = sum over the up x up block of the NCHW full-resolution gradients (backward of vbg_upsample_split_nchw)
__global__ void upsample_split_bwd_kernel(const float* __restrict__ d1, const float* __restrict__ d2, int B, int h, int w, int Ct, int up, int c_split,
                                          float* __restrict__ dl) {
  const long long total = (long long)B * h * w * Ct;
  const int H = h * up, W = w * up;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Ct); long long t = i / Ct;
    const int x = (int)(t % w); t /= w;
    const int y = (int)(t % h); const int b = (int)(t / h);
    const bool first = c < c_split;
    const float* src = first ? d1 + ((size_t)b * c_split + c) * H * W : d2 + ((size_t)b * (Ct - c_split) + (c - c_split)) * H * W;
    float acc = 0.f;
    for (int dy = 0; dy < up; ++dy)
      for (int dx = 0; dx < up; ++dx) acc += __ldg(src + ((size_t)y * up + dy) * W + (size_t)x * up + dx);
    dl[i] = acc;
  }
}

// ------------------------------------------------------------------ small-N weight gradient
// dW[n, k] = sum_r dY[r, n] X[r, k], N <= 16: thread = one column k, 16 accumulators; a CTA walks a row chunk with the dY rows
// staged through shared memory; per-CTA partials are summed by sum_slabs_kernel (fixed order).
constexpr int kSmallN = 16;
__global__ void __launch_bounds__(256)
small_wgrad_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ x, int ldx, long long M, int N, int K, long long rows_per_cta,
                   float* __restrict__ partial) {
  __shared__ float sdy[64][kSmallN];
  const int k = blockIdx.y * 256 + threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  long long r1 = r0 + rows_per_cta; if (r1 > M) r1 = M;
  float acc[kSmallN];
#pragma unroll
  for (int n = 0; n < kSmallN; ++n) acc[n] = 0.f;
  for (long long rb = r0; rb < r1; rb += 64) {
    const int nr = (int)((r1 - rb) < 64 ? (r1 - rb) : 64);
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * kSmallN; i += 256) {
      const int rr = i / kSmallN, n = i % kSmallN;
      sdy[rr][n] = (rr < nr && n < N) ? __ldg(dy + (size_t)(rb + rr) * ldy + n) : 0.f;
    }
    __syncthreads();
    if (k < K)
      for (int rr = 0; rr < nr; ++rr) {
        const float xv = __ldg(x + (size_t)(rb + rr) * ldx + k);
#pragma unroll
        for (int n = 0; n < kSmallN; ++n) acc[n] = fmaf(sdy[rr][n], xv, acc[n]);
      }
  }
  if (k < K) {
#pragma unroll
    for (int n = 0; n < kSmallN; ++n)
      if (n < N) partial[((size_t)blockIdx.x * N + n) * K + k] = acc[n];
  }
}

// column sums of [rows, cols] (fp32, or bf16 hi/lo planes): (cols / 32) x G CTAs, each 8 row lanes x 32 columns over a row chunk;
// the G partial rows are summed in order by sum_slabs_kernel
__global__ void __launch_bounds__(256)
colsum_kernel(const void* __restrict__ x, long long x_plane, long long rows, int cols, long long rows_per_cta, float* __restrict__ partial) {
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta; if (r1 > rows) r1 = rows;
  float acc = 0.f;
  if (c < cols)
    for (long long r = r0 + ty; r < r1; r += 8) {
      const size_t idx = (size_t)r * cols + c;
      acc += x_plane == 0 ? __ldg(reinterpret_cast<const float*>(x) + idx)
                          : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[idx]) +
                                __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[idx + x_plane]);
    }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][tx];
    partial[(size_t)blockIdx.y * cols + c] = t;
  }
}

// cols % 4 == 0, either storage format: a warp reads 512 contiguous bytes of a row (lane = 4 columns: one float4, or one
// 8-byte load from each bf16 plane), 8 row lanes per CTA, four rows in flight per thread; per column the rows are still
// added in ascending order within a lane, lanes in fixed order.
__global__ void __launch_bounds__(256)
colsum4_kernel(const void* __restrict__ x, long long x_plane, long long rows, int cols4, long long rows_per_cta, float4* __restrict__ partial) {
  __shared__ float4 part[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5, c = blockIdx.x * 32 + tx;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta; if (r1 > rows) r1 = rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols4) {
    long long r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      const float4 v0 = ld4_fmt(x, x_plane, (size_t)r * cols4 + c), v1 = ld4_fmt(x, x_plane, (size_t)(r + 8) * cols4 + c);
      const float4 v2 = ld4_fmt(x, x_plane, (size_t)(r + 16) * cols4 + c), v3 = ld4_fmt(x, x_plane, (size_t)(r + 24) * cols4 + c);
      acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
      acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
      acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
      acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
    }
    for (; r < r1; r += 8) {
      const float4 v = ld4_fmt(x, x_plane, (size_t)r * cols4 + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols4) {
    float4 t = part[0][tx];
#pragma unroll
    for (int i = 1; i < 8; ++i) { const float4 u = part[i][tx]; t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w; }
    partial[(size_t)blockIdx.y * cols4 + c] = t;
  }
}

__global__ void sum_slabs_kernel(const float* __restrict__ partial, int slabs, long long n, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < slabs; ++s) a += __ldg(partial + (size_t)s * n + i);
    out[i] = a;
  }
}

// ------------------------------------------------------------------ stem weight gradient
// y[b,i,j,co] = sum_{r,s,c} x4[b, 2i + r, 2j + s, c] w[co,r,s,c] over the zero-bordered NHWC4 batch (border 3 = the padding), so
// dW[co, (r,s,c)] = sum_px dy[px, co] * patch[px, (r,s,c)]: a 64 x 196 accumulator per CTA, 4 x 13 per thread, over 8x8-pixel tiles
// staged in shared memory; persistent CTAs, partials summed in fixed order.
constexpr int kStemCols = 196, kStemColsPerThread = 13, kStemPatch = 21;
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ x4, const float* __restrict__ dy, int B, int Hp, int Wp, int Ho, int Wo, float* __restrict__ partial) {
  __shared__ float4 patch[kStemPatch * kStemPatch];
  __shared__ float4 sdy[64][16];
  const int cg = threadIdx.x >> 4, jg = threadIdx.x & 15;
  int off[kStemColsPerThread];
  bool live[kStemColsPerThread];
#pragma unroll
  for (int t = 0; t < kStemColsPerThread; ++t) {
    const int col = jg * kStemColsPerThread + t;
    live[t] = col < kStemCols;
    const int cc = live[t] ? col : 0;
    const int r = cc / 28, s = (cc % 28) / 4, c = cc % 4;
    off[t] = (r * kStemPatch + s) * 4 + c;
  }
  float acc[4][kStemColsPerThread];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int t = 0; t < kStemColsPerThread; ++t) acc[i][t] = 0.f;
  const int tiles_w = (Wo + 7) >> 3, tiles_h = (Ho + 7) >> 3;
  const int total = B * tiles_h * tiles_w;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, b = tile / (tiles_w * tiles_h);
    const int i0 = th * 8, j0 = tw * 8;
    __syncthreads();
    for (int i = threadIdx.x; i < kStemPatch * kStemPatch; i += 256) {
      const int pr = i / kStemPatch, pc = i % kStemPatch;
      const int hh = 2 * i0 + pr, ww = 2 * j0 + pc;
      patch[i] = (hh < Hp && ww < Wp) ? __ldg(reinterpret_cast<const float4*>(x4) + ((size_t)b * Hp + hh) * Wp + ww) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int px = i >> 4, c4 = i & 15;
      const int ii = i0 + (px >> 3), jj = j0 + (px & 7);
      sdy[px][c4] = (ii < Ho && jj < Wo) ? __ldg(reinterpret_cast<const float4*>(dy) + (((size_t)b * Ho + ii) * Wo + jj) * 16 + c4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const float* pf = reinterpret_cast<const float*>(patch);
    for (int px = 0; px < 64; ++px) {
      const float4 g = sdy[px][cg];
      const int base = ((2 * (px >> 3)) * kStemPatch + 2 * (px & 7)) * 4;
#pragma unroll
      for (int t = 0; t < kStemColsPerThread; ++t) {
        const float xv = pf[base + off[t]];
        acc[0][t] = fmaf(g.x, xv, acc[0][t]); acc[1][t] = fmaf(g.y, xv, acc[1][t]);
        acc[2][t] = fmaf(g.z, xv, acc[2][t]); acc[3][t] = fmaf(g.w, xv, acc[3][t]);
      }
    }
  }
  float* p = partial + (size_t)blockIdx.x * 64 * kStemCols;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int t = 0; t < kStemColsPerThread; ++t)
      if (live[t]) p[(cg * 4 + i) * kStemCols + jg * kStemColsPerThread + t] = acc[i][t];
}

}  // namespace vbg

using namespace vbg;

// =================================================================== C ABI
extern "C" int vbg_bn_stats(const float* x, long long rows, int C, float eps, float* mean, float* var, float* rstd, float* workspace,
                            size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(x && mean && var && rstd && workspace && rows > 0 && C >= 4 && C % 4 == 0 && 256 % (C / 4) == 0 && aligned16(x),
              "vbg_bn_stats: C must be 4 * a divisor of 256, 16B-aligned input");
  long long rpc; const int nblk = colreduce_geometry(rows, C, rpc);
  if ((size_t)nblk * 2 * C * 4 > ws_bytes) { set_error("vbg_bn_stats: workspace of %zu bytes needed", (size_t)nblk * 2 * C * 4); return VBG_EWORKSPACE; }
  cudaStream_t s = as_stream(stream);
  colreduce2_kernel<<<nblk, 256, 0, s>>>(BnStatsF{reinterpret_cast<const float4*>(x)}, rows, C / 4, rpc, workspace);
  colreduce2_finish_kernel<<<cdiv(C, 8), 256, 0, s>>>(workspace, nblk, C, 1.0 / (double)rows, eps, 0, mean, var, rstd);
  return check_launch("vbg_bn_stats");
}

extern "C" long long vbg_bn_workspace(long long rows, int C) {
  if (C < 4 || C % 4 || 256 % (C / 4)) return 0;
  long long rpc; const int nblk = colreduce_geometry(rows, C, rpc);
  return (long long)nblk * 2 * C * 4;
}

extern "C" int vbg_bn_apply(const float* x, long long rows, int C, const float* mean, const float* rstd, const float* gamma, const float* beta,
                            const float* residual, int relu, float* y, vbg_stream_t stream) {
  VBG_REQUIRE(x && mean && rstd && gamma && beta && y && rows > 0 && C % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(gamma) &&
                  aligned16(beta) && aligned16(mean) && aligned16(rstd) && (!residual || aligned16(residual)),
              "vbg_bn_apply: C %% 4 == 0 and 16B-aligned pointers");
  const long long n4 = rows * (C / 4);
  bn_apply_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), n4, C / 4, reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(rstd),
      reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(beta), reinterpret_cast<const float4*>(residual), relu,
      reinterpret_cast<float4*>(y));
  return check_launch("vbg_bn_apply");
}

extern "C" int vbg_bn_bwd(const float* x, const float* dy, const float* y_relu, long long rows, int C, const float* mean, const float* rstd,
                          const float* gamma, float* dx, float* dres, float* dgamma, float* dbeta, float* workspace, size_t ws_bytes,
                          vbg_stream_t stream) {
  VBG_REQUIRE(x && dy && mean && rstd && gamma && dx && dgamma && dbeta && workspace && rows > 0 && C >= 4 && C % 4 == 0 &&
                  256 % (C / 4) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx) && (!y_relu || aligned16(y_relu)) &&
                  (!dres || aligned16(dres)) && aligned16(gamma) && aligned16(mean) && aligned16(rstd) && aligned16(dgamma) && aligned16(dbeta),
              "vbg_bn_bwd: C must be 4 * a divisor of 256, 16B-aligned pointers");
  long long rpc; const int nblk = colreduce_geometry(rows, C, rpc);
  if ((size_t)nblk * 2 * C * 4 > ws_bytes) { set_error("vbg_bn_bwd: workspace of %zu bytes needed", (size_t)nblk * 2 * C * 4); return VBG_EWORKSPACE; }
  cudaStream_t s = as_stream(stream);
  BnBwdF f{reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y_relu),
           reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(rstd)};
  colreduce2_kernel<<<nblk, 256, 0, s>>>(f, rows, C / 4, rpc, workspace);
  colreduce2_finish_kernel<<<cdiv(C, 8), 256, 0, s>>>(workspace, nblk, C, 0.0, 0.f, 1, dbeta, dgamma, nullptr);
  const long long n4 = rows * (C / 4);
  bn_bwd_dx_kernel<<<grid_for(n4, 256), 256, 0, s>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y_relu), n4, C / 4,
      (float)(1.0 / (double)rows), reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(rstd),
      reinterpret_cast<const float4*>(gamma), reinterpret_cast<const float4*>(dgamma), reinterpret_cast<const float4*>(dbeta),
      reinterpret_cast<float4*>(dx), reinterpret_cast<float4*>(dres));
  return check_launch("vbg_bn_bwd");
}

// The two halves of vbg_bn_bwd as separate entry points, for SyncBatchNorm: the per-channel sums are all-reduced between them.
extern "C" int vbg_bn_bwd_reduce(const float* x, const float* dy, const float* y_relu, long long rows, int C, const float* mean,
                                 const float* rstd, float* sum_dy_xhat, float* sum_dy, float* workspace, size_t ws_bytes,
                                 vbg_stream_t stream) {
  VBG_REQUIRE(x && dy && mean && rstd && sum_dy_xhat && sum_dy && workspace && rows > 0 && C >= 4 && C % 4 == 0 && 256 % (C / 4) == 0 &&
                  aligned16(x) && aligned16(dy) && (!y_relu || aligned16(y_relu)) && aligned16(mean) && aligned16(rstd),
              "vbg_bn_bwd_reduce: C must be 4 * a divisor of 256, 16B-aligned pointers");
  long long rpc; const int nblk = colreduce_geometry(rows, C, rpc);
  if ((size_t)nblk * 2 * C * 4 > ws_bytes) { set_error("vbg_bn_bwd_reduce: workspace of %zu bytes needed", (size_t)nblk * 2 * C * 4); return VBG_EWORKSPACE; }
  cudaStream_t s = as_stream(stream);
  BnBwdF f{reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y_relu),
           reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(rstd)};
  colreduce2_kernel<<<nblk, 256, 0, s>>>(f, rows, C / 4, rpc, workspace);
  colreduce2_finish_kernel<<<cdiv(C, 8), 256, 0, s>>>(workspace, nblk, C, 0.0, 0.f, 1, sum_dy, sum_dy_xhat, nullptr);
  return check_launch("vbg_bn_bwd_reduce");
}

extern "C" int vbg_bn_bwd_dx(const float* x, const float* dy, const float* y_relu, long long rows, int C, float inv_count, const float* mean,
                             const float* rstd, const float* gamma, const float* sum_dy_xhat, const float* sum_dy, float* dx, float* dres,
                             vbg_stream_t stream) {
  VBG_REQUIRE(x && dy && mean && rstd && gamma && sum_dy_xhat && sum_dy && dx && rows > 0 && C >= 4 && C % 4 == 0 && aligned16(x) &&
                  aligned16(dy) && aligned16(dx) && (!y_relu || aligned16(y_relu)) && (!dres || aligned16(dres)) && aligned16(gamma) &&
                  aligned16(mean) && aligned16(rstd) && aligned16(sum_dy_xhat) && aligned16(sum_dy) && inv_count > 0.f,
              "vbg_bn_bwd_dx: C %% 4 == 0, 16B-aligned pointers, inv_count = 1 / (rows summed into the two sums)");
  const long long n4 = rows * (C / 4);
  bn_bwd_dx_kernel<<<grid_for(n4, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), reinterpret_cast<const float4*>(y_relu), n4, C / 4, inv_count,
      reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(rstd), reinterpret_cast<const float4*>(gamma),
      reinterpret_cast<const float4*>(sum_dy_xhat), reinterpret_cast<const float4*>(sum_dy), reinterpret_cast<float4*>(dx),
      reinterpret_cast<float4*>(dres));
  return check_launch("vbg_bn_bwd_dx");
}

extern "C" int vbg_maxpool3x3s2_bwd(const float* x, const float* dy, int B, int H, int W, int C, float* dx, vbg_stream_t stream) {
  VBG_REQUIRE(x && dy && dx && B > 0 && H > 0 && W > 0 && C % 4 == 0 && aligned16(x) && aligned16(dy) && aligned16(dx),
              "vbg_maxpool3x3s2_bwd: C %% 4 == 0 and 16B-aligned pointers");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  maxpool3x3s2_bwd_kernel<<<grid_for((long long)B * H * W * (C / 4), 256, kNumSMs * 16), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), B, H, W, C / 4, Ho, Wo, reinterpret_cast<float4*>(dx));
  return check_launch("vbg_maxpool3x3s2_bwd");
}

extern "C" int vbg_maxpool3x3s2_bwd_y(const float* x, const float* y, const float* dy, int B, int H, int W, int C, float* dx,
                                      vbg_stream_t stream) {
  VBG_REQUIRE(x && y && dy && dx && B > 0 && H > 0 && W > 0 && C % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(dy) && aligned16(dx),
              "vbg_maxpool3x3s2_bwd_y: C %% 4 == 0 and 16B-aligned pointers");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  maxpool3x3s2_bwd_y_kernel<<<grid_for((long long)B * H * W * (C / 4), 256, kNumSMs * 16), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(y), reinterpret_cast<const float4*>(dy), B, H, W, C / 4, Ho, Wo,
      reinterpret_cast<float4*>(dx));
  return check_launch("vbg_maxpool3x3s2_bwd_y");
}

extern "C" int vbg_sumpool2x2(const float* x, int B, int H, int W, int C, float scale, float* y, vbg_stream_t stream) {
  VBG_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0 && aligned16(x) && aligned16(y),
              "vbg_sumpool2x2: even H, W; C %% 4 == 0; 16B-aligned pointers");
  sumpool2x2_kernel<<<grid_for((long long)B * (H / 2) * (W / 2) * (C / 4), 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), B, H, W, C / 4, scale, reinterpret_cast<float4*>(y));
  return check_launch("vbg_sumpool2x2");
}

extern "C" int vbg_expand2x(const float* x, int B, int Hi, int Wi, int C, int H, int W, float scale, int zero_insert, float* y,
                            vbg_stream_t stream) {
  VBG_REQUIRE(x && y && B > 0 && Hi > 0 && Wi > 0 && H > 0 && W > 0 && C % 4 == 0 && aligned16(x) && aligned16(y),
              "vbg_expand2x: C %% 4 == 0 and 16B-aligned pointers");
  expand2x_kernel<<<grid_for((long long)B * H * W * (C / 4), 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(x), B, H, W, C / 4, Hi, Wi, scale, zero_insert, reinterpret_cast<float4*>(y));
  return check_launch("vbg_expand2x");
}

extern "C" int vbg_gelu(const float* x, const float* dy, long long n, float* out, vbg_stream_t stream) {
  return vbg_gelu_x(x, 0, dy, 0, n, out, 0, stream);
}

extern "C" int vbg_gelu_x(const void* x, long long x_plane, const void* dy, long long dy_plane, long long n, void* out, long long out_plane,
                          vbg_stream_t stream) {
  VBG_REQUIRE(n > 0 && n % 4 == 0 && fmt_ok(x, x_plane) && fmt_ok(out, out_plane) && (!dy || fmt_ok(dy, dy_plane)),
              "vbg_gelu: n %% 4 == 0, 16B-aligned tensors, planes a multiple of 8 elements apart");
  gelu_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>(x, x_plane, dy, dy_plane, n / 4, out, out_plane);
  return check_launch("vbg_gelu");
}

extern "C" int vbg_dropout(const float* x, long long n, float p, unsigned long long seed, float* y, vbg_stream_t stream) {
  return vbg_dropout_ds(x, n, p, seed, nullptr, y, stream);
}

extern "C" int vbg_dropout_ds(const float* x, long long n, float p, unsigned long long seed, const unsigned long long* step_seed,
                              float* y, vbg_stream_t stream) {
  VBG_REQUIRE(x && y && n > 0 && p >= 0.f && p < 1.f, "vbg_dropout: 0 <= p < 1");
  dropout_kernel<<<grid_for(n, 1024), 256, 0, as_stream(stream)>>>(x, n, p, seed, step_seed, y);
  return check_launch("vbg_dropout");
}

extern "C" int vbg_grid_scatter_bwd(const float* dgrid, long long ld, const int32_t* idx, const int32_t* boxes, const int32_t* seg_off, int B, int K,
                                    int stride, int Hg, int Wg, int C, float* demb, vbg_stream_t stream) {
  VBG_REQUIRE(dgrid && idx && boxes && seg_off && demb && B > 0 && K >= 0 && stride > 0 && C % 4 == 0 && ld >= C && ld % 4 == 0 &&
                  aligned16(dgrid) && aligned16(demb) && aligned16(boxes),
              "vbg_grid_scatter_bwd: C, ld multiples of 4, 16B-aligned pointers");
  if (K == 0) return VBG_OK;
  grid_scatter_bwd_kernel<<<K, 256, 0, as_stream(stream)>>>(dgrid, ld, idx, boxes, seg_off, B, stride, Hg, Wg, C / 4, demb);
  return check_launch("vbg_grid_scatter_bwd");
}

extern "C" int vbg_segment_reduce_bwd(const float* dseg, const int32_t* tok_row, const int32_t* seg_start, int K, int C, int mode, float* dhidden,
                                      vbg_stream_t stream) {
  VBG_REQUIRE(dseg && tok_row && seg_start && dhidden && K >= 0 && C % 4 == 0 && aligned16(dseg) && aligned16(dhidden),
              "vbg_segment_reduce_bwd: C %% 4 == 0, 16B-aligned pointers");
  if (K == 0) return VBG_OK;
  segment_reduce_bwd_kernel<<<K, 192, 0, as_stream(stream)>>>(dseg, tok_row, seg_start, C / 4, mode, dhidden);
  return check_launch("vbg_segment_reduce_bwd");
}

extern "C" int vbg_embed_bwd(const float* dx, const int32_t* ids, const int32_t* pos, int R, int hidden, float* dword, float* dpos,
                             vbg_stream_t stream) {
  VBG_REQUIRE(dx && ids && pos && dword && dpos && R > 0 && hidden > 0 && hidden % 4 == 0 && hidden <= 2048 && aligned16(dx) &&
                  aligned16(dword) && aligned16(dpos),
              "vbg_embed_bwd: hidden %% 4 == 0, hidden <= 2048, 16B-aligned pointers");
  embed_bwd_kernel<<<dim3((unsigned)R, 2), 256, 0, as_stream(stream)>>>(dx, ids, pos, R, hidden, dword, dpos);
  return check_launch("vbg_embed_bwd");
}

extern "C" long long vbg_roi_align_bwd_workspace(int K) { return (long long)(K > 0 ? K : 1) * (long long)sizeof(RoiGeo); }

extern "C" int vbg_roi_align_bwd(const float* dout, int B, int Hf, int Wf, int C, const int32_t* boxes, const int32_t* seg_off, int K, float scale,
                                 int P, float* dfeat, void* workspace, size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(dout && boxes && seg_off && dfeat && workspace && B > 0 && Hf > 0 && Wf > 0 && C % 4 == 0 && K >= 0 && aligned16(dout) &&
                  aligned16(dfeat) && aligned16(boxes) && aligned16(workspace),
              "vbg_roi_align_bwd: C %% 4 == 0, 16B-aligned pointers");
  VBG_REQUIRE(P == 7, "vbg_roi_align_bwd: 7 x 7 bins only (got %d)", P);
  if ((size_t)vbg_roi_align_bwd_workspace(K) > ws_bytes) { set_error("vbg_roi_align_bwd: workspace of %lld bytes needed", vbg_roi_align_bwd_workspace(K)); return VBG_EWORKSPACE; }
  cudaStream_t s = as_stream(stream);
  RoiGeo* geo = reinterpret_cast<RoiGeo*>(workspace);
  if (K > 0) roi_geo_kernel<<<cdiv(K, 128), 128, 0, s>>>(boxes, K, scale, P, Hf, Wf, geo);
  roi_align_bwd_kernel<<<dim3(cdiv(Wf, kRbT), cdiv(Hf, kRbT), B), 256, 0, s>>>(dout, B, Hf, Wf, C, seg_off, geo, P, dfeat);
  return check_launch("vbg_roi_align_bwd");
}

extern "C" int vbg_seg_ce_bwd(const float* logits, const long long* pos_neg, const long long* cls, int B, int H, int W, int up, int Ct,
                              int c_split, const float* gscale, float* dlogits, vbg_stream_t stream) {
  VBG_REQUIRE(logits && pos_neg && cls && gscale && dlogits && B > 0 && up > 0 && H % up == 0 && W % up == 0 && Ct > c_split && c_split > 0,
              "vbg_seg_ce_bwd: bad arguments");
  seg_ce_bwd_kernel<<<grid_for((long long)B * (H / up) * (W / up), 128), 128, 0, as_stream(stream)>>>(
      logits, pos_neg, cls, B, H, W, up, Ct, c_split, gscale, (float)(1.0 / ((double)B * H * W)), dlogits);
  return check_launch("vbg_seg_ce_bwd");
}

extern "C" int vbg_upsample_split_bwd(const float* d1, const float* d2, int B, int h, int w, int Ct, int up, int c_split, float* dlogits,
                                      vbg_stream_t stream) {
  VBG_REQUIRE(d1 && d2 && dlogits && B > 0 && h > 0 && w > 0 && up > 0 && Ct > c_split && c_split > 0, "vbg_upsample_split_bwd: bad arguments");
  upsample_split_bwd_kernel<<<grid_for((long long)B * h * w * Ct, 256), 256, 0, as_stream(stream)>>>(d1, d2, B, h, w, Ct, up, c_split, dlogits);
  return check_launch("vbg_upsample_split_bwd");
}

static bool colsum_vec(const void* x, long long x_plane, int cols) { return (cols & 3) == 0 && fmt_ok(x, x_plane); }
static int colsum_geometry(long long rows, int cols, long long& rows_per_cta, bool vec = false) {
  const int col_tiles = vec ? cdiv(cols, 128) : cdiv(cols, 32);
  long long g = (kNumSMs * 8 + col_tiles - 1) / col_tiles;
  const long long max_g = (rows + 63) / 64;
  if (g > max_g) g = max_g;
  if (g < 1) g = 1;
  rows_per_cta = ((rows + g - 1) / g + 7) / 8 * 8;
  return (int)((rows + rows_per_cta - 1) / rows_per_cta);
}

extern "C" long long vbg_colsum_workspace(long long rows, int cols) {
  long long rpc; int g = colsum_geometry(rows, cols, rpc);
  if ((cols & 3) == 0) { const int gv = colsum_geometry(rows, cols, rpc, true); if (gv > g) g = gv; }    // the float4 form splits rows finer
  return g > 1 ? (long long)g * cols * 4 : 0;
}

extern "C" int vbg_colsum(const void* x, long long x_plane, long long rows, int cols, float* out, float* workspace, size_t ws_bytes,
                          vbg_stream_t stream) {
  VBG_REQUIRE(x && out && rows > 0 && cols > 0 && x_plane >= 0, "vbg_colsum: bad arguments");
  const bool vec = colsum_vec(x, x_plane, cols) && aligned16(out) && (!workspace || aligned16(workspace));
  long long rpc; int g = colsum_geometry(rows, cols, rpc, vec);
  if (g > 1 && (!workspace || (size_t)g * cols * 4 > ws_bytes)) { g = 1; rpc = rows; }       // no workspace: one CTA row per column tile
  cudaStream_t s = as_stream(stream);
  if (vec)
    colsum4_kernel<<<dim3(cdiv(cols, 128), g), 256, 0, s>>>(x, x_plane, rows, cols / 4, rpc, reinterpret_cast<float4*>(g > 1 ? workspace : out));
  else
    colsum_kernel<<<dim3(cdiv(cols, 32), g), 256, 0, s>>>(x, x_plane, rows, cols, rpc, g > 1 ? workspace : out);
  if (g > 1) sum_slabs_kernel<<<grid_for(cols, 256), 256, 0, s>>>(workspace, g, cols, out);
  return check_launch("vbg_colsum");
}

static int small_wgrad_geometry(long long M, long long& rows_per_cta) {
  long long want = (M + kNumSMs * 2 - 1) / (kNumSMs * 2);
  rows_per_cta = ((want + 63) / 64) * 64;
  return (int)((M + rows_per_cta - 1) / rows_per_cta);
}

extern "C" long long vbg_small_wgrad_workspace(long long M, int N, int K) {
  long long rpc; const int nblk = small_wgrad_geometry(M, rpc);
  return (long long)nblk * N * K * 4;
}

extern "C" int vbg_small_wgrad(const float* dy, int ldy, const float* x, int ldx, long long M, int N, int K, float* dw, float* workspace,
                               size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(dy && x && dw && workspace && M > 0 && N > 0 && N <= kSmallN && K > 0 && ldy >= N && ldx >= K, "vbg_small_wgrad: 1 <= N <= 16");
  long long rpc; const int nblk = small_wgrad_geometry(M, rpc);
  if ((size_t)nblk * N * K * 4 > ws_bytes) { set_error("vbg_small_wgrad: workspace of %zu bytes needed", (size_t)nblk * N * K * 4); return VBG_EWORKSPACE; }
  cudaStream_t s = as_stream(stream);
  small_wgrad_kernel<<<dim3(nblk, cdiv(K, 256)), 256, 0, s>>>(dy, ldy, x, ldx, M, N, K, rpc, workspace);
  sum_slabs_kernel<<<grid_for((long long)N * K, 256), 256, 0, s>>>(workspace, nblk, (long long)N * K, dw);
  return check_launch("vbg_small_wgrad");
}

extern "C" long long vbg_stem_wgrad_workspace(void) { return (long long)kNumSMs * 2 * 64 * kStemCols * 4; }

extern "C" int vbg_stem_wgrad(const float* x4, const float* dy, int B, int Hp, int Wp, int Ho, int Wo, float* dw774, float* workspace,
                              size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(x4 && dy && dw774 && workspace && B > 0 && Hp >= 2 * Ho + 5 && Wp >= 2 * Wo + 5 && aligned16(x4) && aligned16(dy),
              "vbg_stem_wgrad: x4 is the zero-bordered NHWC4 batch [B, Hp, Wp, 4] with Hp >= 2 Ho + 5");
  const int nblk = kNumSMs * 2;
  if ((size_t)nblk * 64 * kStemCols * 4 > ws_bytes) { set_error("vbg_stem_wgrad: workspace too small"); return VBG_EWORKSPACE; }
  cudaStream_t s = as_stream(stream);
  stem_wgrad_kernel<<<nblk, 256, 0, s>>>(x4, dy, B, Hp, Wp, Ho, Wo, workspace);
  sum_slabs_kernel<<<grid_for(64 * kStemCols, 256), 256, 0, s>>>(workspace, nblk, 64 * kStemCols, dw774);
  return check_launch("vbg_stem_wgrad");
}

// ------------------------------------------------------------------ sampling keys of the device-side sampled / OHEM losses
// key[i] = 31-bit hash of (seed, i): "element i is kept iff its key is among the k smallest of its group" (losses_device.py) --
// uniform k-subsets without host randomness, so the loss tail needs no device->host sync and can live in a CUDA graph
// (step_seed: the device word refreshed per replay, see vbg_dropout_ds).
namespace vbg {
__global__ void uniform_keys_kernel(long long n, unsigned long long seed, const unsigned long long* __restrict__ step_seed, int32_t* __restrict__ out) {
  if (step_seed) seed ^= __ldg(step_seed) * 0xD6E8FEB86659FD93ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    unsigned long long z = seed + (unsigned long long)(i + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    out[i] = (int32_t)(z >> 33);
  }
}
}  // namespace vbg

extern "C" int vbg_uniform_keys(long long n, unsigned long long seed, const unsigned long long* step_seed, int32_t* out, vbg_stream_t stream) {
  VBG_REQUIRE(out && n > 0, "vbg_uniform_keys: bad arguments");
  vbg::uniform_keys_kernel<<<vbg::grid_for(n, 1024), 256, 0, vbg::as_stream(stream)>>>(n, seed, step_seed, out);
  return vbg::check_launch("vbg_uniform_keys");
}
