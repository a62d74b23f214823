// fp32 CUDA-core implicit-GEMM (VBG_PREC_FP32): the exact-fp32 arithmetic path for every dense
// contraction of the forward -- linears, 1x1 / 3x3 / 7x7 NHWC convolutions -- and the fallback for
// shapes the tcgen05 path (vbg_gemm_tc.cu) does not take (tiny N, Cin=3 stem, strided convs).
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T ),  M = B*Ho*Wo pixels (or rows), N = Cout, K = kh*kw*Cin
//
// A is gathered on the fly (im2col never materialised; channel-concatenation of two sources for
// the cat-free early-fusion / late-fusion forms).  Epilogue fuses folded-BN scale/shift or bias,
// residual add (same-shape or nearest-x2-upsampled, which fuses the FPN top-down add), ReLU/GELU.
//
// Tiling: BMxBNx16 CTA tile, 256 threads, (BM/16)x(BN/16) register micro-tile, register-staged
// double buffering (global loads for tile k+1 are in flight while tile k is multiplied).
#include "vbg_common.cuh"

namespace vbg {

struct GemmArgs {
  const float* A; const float* A2; const float* W; float* C;
  int lda, lda2, K1, ldw, ldc;
  int M, N, K;
  // conv geometry (conv == 1): A is x[B,H,W,Cin]
  int conv, H, Wd, Cin, Ho, Wo, kh, kw, stride, pad;
  vbg_epilogue_t ep;
};

constexpr int BK = 16;

template <int BM, int BN, bool VEC>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const GemmArgs g) {
  constexpr int TM = BM / 16, TN = BN / 16;         // per-thread micro tile
  constexpr int A_PER = BM * BK / 4 / 256;          // float4 (or 4 scalars) per thread per tile
  constexpr int B_PER = BN * BK / 4 / 256;
  static_assert(A_PER >= 1 && B_PER >= 1, "tile too small for 256 threads");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  // ---- per-thread load descriptors: element group i covers tile row (tid/4 + i*64), k-quad (tid%4)
  const int kq = (tid & 3) * 4;
  int a_row[A_PER]; long long a_base[A_PER]; int a_h0[A_PER], a_w0[A_PER]; bool a_ok[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    const int r = (tid >> 2) + i * 64;
    a_row[i] = r;
    const int m = m0 + r;
    a_ok[i] = m < g.M;
    if (g.conv) {
      const int mm = a_ok[i] ? m : 0;
      const int wo = mm % g.Wo; const int t = mm / g.Wo;
      const int ho = t % g.Ho; const int b = t / g.Ho;
      a_h0[i] = ho * g.stride - g.pad; a_w0[i] = wo * g.stride - g.pad;
      a_base[i] = (long long)b * g.H * g.Wd * g.Cin;
    } else {
      a_h0[i] = a_w0[i] = 0;
      a_base[i] = (long long)(a_ok[i] ? m : 0);
    }
  }
  int b_row[B_PER]; bool b_ok[B_PER];
#pragma unroll
  for (int i = 0; i < B_PER; ++i) {
    b_row[i] = (tid >> 2) + i * 64;
    b_ok[i] = (n0 + b_row[i]) < g.N;
  }

  float4 ra[A_PER], rb[B_PER];
  auto load_tile = [&](int k0) {
    const int k = k0 + kq;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a_ok[i]) {
        if (g.conv) {
          if (VEC) {
            if (k < g.K) {
              const int c = k % g.Cin; const int rs = k / g.Cin;
              const int s = rs % g.kw; const int r = rs / g.kw;
              const int hi = a_h0[i] + r, wi = a_w0[i] + s;
              if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.Wd)
                v = __ldg(reinterpret_cast<const float4*>(g.A + a_base[i] + ((long long)hi * g.Wd + wi) * g.Cin + c));
            }
          } else {
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              t[j] = 0.f;
              const int kk = k + j;
              if (kk < g.K) {
                const int c = kk % g.Cin; const int rs = kk / g.Cin;
                const int s = rs % g.kw; const int r = rs / g.kw;
                const int hi = a_h0[i] + r, wi = a_w0[i] + s;
                if (hi >= 0 && hi < g.H && wi >= 0 && wi < g.Wd)
                  t[j] = __ldg(g.A + a_base[i] + ((long long)hi * g.Wd + wi) * g.Cin + c);
              }
            }
            v = make_float4(t[0], t[1], t[2], t[3]);
          }
        } else {
          if (VEC) {
            if (k < g.K) {
              v = (k < g.K1) ? __ldg(reinterpret_cast<const float4*>(g.A + a_base[i] * g.lda + k))
                             : __ldg(reinterpret_cast<const float4*>(g.A2 + a_base[i] * g.lda2 + (k - g.K1)));
            }
          } else {
            float t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int kk = k + j;
              t[j] = 0.f;
              if (kk < g.K) t[j] = (kk < g.K1) ? __ldg(g.A + a_base[i] * g.lda + kk) : __ldg(g.A2 + a_base[i] * g.lda2 + (kk - g.K1));
            }
            v = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b_ok[i]) {
        const float* p = g.W + (long long)(n0 + b_row[i]) * g.ldw + k;
        if (VEC) {
          if (k < g.K) v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          float t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] = (k + j < g.K) ? __ldg(p + j) : 0.f;
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      rb[i] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      As[buf][kq + 0][a_row[i]] = ra[i].x; As[buf][kq + 1][a_row[i]] = ra[i].y;
      As[buf][kq + 2][a_row[i]] = ra[i].z; As[buf][kq + 3][a_row[i]] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      Bs[buf][kq + 0][b_row[i]] = rb[i].x; Bs[buf][kq + 1][b_row[i]] = rb[i].y;
      Bs[buf][kq + 2][b_row[i]] = rb[i].z; Bs[buf][kq + 3][b_row[i]] = rb[i].w;
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (g.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
      // rows ty*4..+3 and (BM/2)+ty*4..+3 (for TM=8); cols likewise: conflict-free float4 reads
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][(i / 4) * (BM / (TM / 4)) + ty * 4]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][(j / 4) * (BN / (TN / 4)) + tx * 4]);
        b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue
  const vbg_epilogue_t& ep = g.ep;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + (i / 4) * (BM / (TM / 4)) + ty * 4 + (i & 3);
    if (m >= g.M) continue;
    long long res_row = 0;
    if (ep.residual) {
      if (ep.res_mode == VBG_RES_UP2) {
        const int wo = m % ep.out_w; const int t = m / ep.out_w;
        const int ho = t % ep.out_h; const int b = t / ep.out_h;
        res_row = (((long long)b * (ep.out_h >> 1) + (ho >> 1)) * (ep.out_w >> 1) + (wo >> 1)) * (long long)g.N;
      } else {
        res_row = (long long)m * ep.ldr;
      }
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + (j / 4) * (BN / (TN / 4)) + tx * 4 + (j & 3);
      if (n >= g.N) continue;
      float v = acc[i][j];
      if (ep.scale) v *= __ldg(ep.scale + n);
      if (ep.shift) v += __ldg(ep.shift + n);
      if (ep.residual) v += __ldg(ep.residual + res_row + n);
      g.C[(long long)m * g.ldc + n] = apply_act(v, ep.act);
    }
  }
}

static int launch_simt(const GemmArgs& g, cudaStream_t s) {
  if (g.M == 0 || g.N == 0) return VBG_OK;
  bool vec;
  if (g.conv) vec = (g.Cin % 4 == 0) && aligned16(g.A) && aligned16(g.W) && (g.ldw % 4 == 0);
  else vec = (g.K % 4 == 0) && (g.K1 % 4 == 0) && (g.lda % 4 == 0) && (g.A2 == nullptr || (g.lda2 % 4 == 0 && aligned16(g.A2))) &&
             (g.ldw % 4 == 0) && aligned16(g.A) && aligned16(g.W);
  // big tile when it still fills the machine, else 64x64 for more CTAs
  long long big_ctas = (long long)cdiv(g.M, 128) * cdiv(g.N, 128);
  bool big = g.N > 64 && big_ctas >= kNumSMs;
  if (big) {
    dim3 grid(cdiv(g.M, 128), cdiv(g.N, 128));
    if (vec) gemm_simt_kernel<128, 128, true><<<grid, 256, 0, s>>>(g);
    else gemm_simt_kernel<128, 128, false><<<grid, 256, 0, s>>>(g);
  } else {
    dim3 grid(cdiv(g.M, 64), cdiv(g.N, 64));
    if (vec) gemm_simt_kernel<64, 64, true><<<grid, 256, 0, s>>>(g);
    else gemm_simt_kernel<64, 64, false><<<grid, 256, 0, s>>>(g);
  }
  return check_launch("vbg_gemm(simt)");
}

int gemm_simt(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int ldw, float* C, int ldc,
              int M, int N, int K, const vbg_epilogue_t* ep, cudaStream_t s) {
  GemmArgs g{};
  g.A = A; g.A2 = A2; g.W = W; g.C = C; g.lda = lda; g.lda2 = lda2; g.K1 = K1; g.ldw = ldw; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K; g.conv = 0;
  if (ep) g.ep = *ep;
  return launch_simt(g, s);
}

int conv_simt(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int kh, int kw, int stride, int pad,
              float* y, const vbg_epilogue_t* ep, cudaStream_t s) {
  GemmArgs g{};
  g.conv = 1; g.A = x; g.W = w; g.C = y;
  g.H = H; g.Wd = W; g.Cin = Cin; g.kh = kh; g.kw = kw; g.stride = stride; g.pad = pad;
  g.Ho = (H + 2 * pad - kh) / stride + 1; g.Wo = (W + 2 * pad - kw) / stride + 1;
  g.M = B * g.Ho * g.Wo; g.N = Cout; g.K = kh * kw * Cin; g.K1 = g.K; g.ldw = g.K; g.ldc = Cout;
  if (ep) g.ep = *ep;
  if (g.ep.res_mode == VBG_RES_UP2) { g.ep.out_h = g.Ho; g.ep.out_w = g.Wo; }
  if (g.ep.res_mode == VBG_RES_SAME && g.ep.ldr == 0) g.ep.ldr = Cout;
  return launch_simt(g, s);
}

}  // namespace vbg
