// C-ABI glue: error state, version, and the precision dispatch of the dense contractions.
#include "vbg_common.cuh"
#include <string.h>

namespace vbg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int gemm_simt(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int ldw, float* C, int ldc,
              int M, int N, int K, const vbg_epilogue_t* ep, cudaStream_t s);
int conv_simt(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int kh, int kw, int stride, int pad,
              float* y, const vbg_epilogue_t* ep, cudaStream_t s);
// tcgen05 path: returns VBG_EUNSUPPORTED (and enqueues nothing) for shapes it does not take
int gemm_tc(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int ldw, float* C, int ldc,
            int M, int N, int K, const vbg_epilogue_t* ep, cudaStream_t s);
int conv_tc(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int kh, int kw, int stride, int pad,
            float* y, const vbg_epilogue_t* ep, cudaStream_t s);
int gemm_tc3(const float* A, int lda, const float* A2, int lda2, int K1, const void* w_split, long long plane, int ldw,
             float* C, int ldc, int M, int N, int K, const vbg_epilogue_t* ep, cudaStream_t s);
int conv_tc3(const float* x, int B, int H, int W, int Cin, const void* w_split, long long plane, int Cout, int kh, int kw,
             int stride, int pad, float* y, const vbg_epilogue_t* ep, cudaStream_t s);
int stem_tc3(const float* x4, int B, int H, int W, const void* w_split, long long plane, int Cout, float* y,
             const vbg_epilogue_t* ep, cudaStream_t s);
int split_bf16(const float* w, long long n, void* hi, void* lo, cudaStream_t s);
int merge_bf16(const void* hi, const void* lo, long long n, float* out, cudaStream_t s);
int gemm_ps(const void* A, long long a_plane, int lda, const void* A2, long long a2_plane, int lda2, int K1, const void* w_hi,
            long long w_plane, int ldw, void* C, int ldc, int M, int N, int K, const vbg_epilogue_t* ep, void* workspace,
            size_t ws_bytes, cudaStream_t s);
int conv_ps(const void* x, long long x_plane, int B, int H, int W, int Cin, const void* w_hi, long long w_plane, int Cout, int kh,
            int kw, int stride, int pad, void* y, const vbg_epilogue_t* ep, void* workspace, size_t ws_bytes, cudaStream_t s);
size_t ps_workspace_bytes(int m_tiles, long long M, int N, int K, int tune);
size_t linear_wgrad_workspace(int M, int N, int K);
size_t conv_wgrad_workspace(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad);
int conv_wgrad(const void* dY, long long y_plane, const void* X, long long x_plane, int B, int H, int W, int Cin, int Cout, int kh, int kw,
               int stride, int pad, float* dW, void* workspace, size_t ws_bytes, cudaStream_t s);
int linear_wgrad(const void* dY, long long y_plane, const void* X, long long x_plane, int M, int N, int K, float* dW, void* workspace,
                 size_t ws_bytes, cudaStream_t s);
size_t conv_ps_workspace_bytes(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int tune);
bool tc_available();

// ------------------------------------------------------------------ CRF Viterbi (model/crf.py:96-146)
// One warp per sample; lane t owns tag t (T <= 32).  Back-pointers live in global scratch-free form:
// the path is re-derived by storing per-step argmax in shared memory in chunks is unnecessary for the
// sequence lengths here (S <= 4096): back-pointers are packed 8-bit in dynamic shared memory.
__global__ void crf_viterbi_kernel(const float* __restrict__ feats, const float* __restrict__ trans,
                                   const int32_t* __restrict__ seg_off, int T, float* __restrict__ tags,
                                   float* __restrict__ scores, unsigned char* __restrict__ bp_scratch) {
  const int b = blockIdx.x, lane = threadIdx.x;
  const int s0 = seg_off[b], n = seg_off[b + 1] - s0;
  const int start = T - 2, stop = T - 1;
  unsigned char* bp = bp_scratch + (size_t)s0 * T;
  float tr[32];
#pragma unroll
  for (int p = 0; p < 32; ++p) tr[p] = (lane < T && p < T) ? trans[lane * T + p] : 0.f;   // to lane from p
  float fv = (lane == start) ? 0.f : -10000.f;
  for (int t = 0; t < n; ++t) {
    float best = -INFINITY; int arg = 0;
#pragma unroll
    for (int p = 0; p < 32; ++p) {
      float prev = __shfl_sync(0xffffffffu, fv, p);
      if (p < T) {
        float v = prev + tr[p];
        if (v > best) { best = v; arg = p; }        // first max wins, like torch.max
      }
    }
    if (lane < T) {
      bp[(size_t)t * T + lane] = (unsigned char)arg;
      fv = best + feats[(size_t)(s0 + t) * T + lane];
    }
  }
  float term = (lane < T) ? fv + trans[stop * T + lane] : -INFINITY;
  float best = term; int arg = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
  }
  __syncwarp();
  if (lane == 0) {
    scores[b] = best;
    int cur = arg;
    for (int t = n - 1; t >= 0; --t) {
      tags[s0 + t] = (float)cur;
      cur = bp[(size_t)t * T + cur];
    }
  }
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_version(void) { return VBG_VERSION; }

extern "C" int vbg_last_error(char* buf, size_t n) {
  size_t len = strlen(g_err);
  if (buf && n > 0) {
    size_t c = len < n - 1 ? len : n - 1;
    memcpy(buf, g_err, c);
    buf[c] = 0;
  }
  return (int)len;
}

extern "C" int vbg_tc_available(void) { return tc_available() ? 1 : 0; }

static int check_epilogue(const vbg_epilogue_t* ep, const char* who) {
  if (!ep) return VBG_OK;
  VBG_REQUIRE(ep->act >= VBG_ACT_NONE && ep->act <= VBG_ACT_GELU, "%s: bad activation %d", who, ep->act);
  VBG_REQUIRE(ep->res_mode >= VBG_RES_NONE && ep->res_mode <= VBG_RES_UP2, "%s: bad residual mode %d", who, ep->res_mode);
  VBG_REQUIRE((ep->residual != nullptr) == (ep->res_mode != VBG_RES_NONE), "%s: residual pointer / mode mismatch", who);
  VBG_REQUIRE(ep->out_mode == VBG_OUT_F32 || ep->out_mode == VBG_OUT_SPLIT_BF16, "%s: bad out_mode %d", who, ep->out_mode);
  VBG_REQUIRE(ep->out_mode == VBG_OUT_F32 || (ep->out_plane > 0 && ep->out_plane % 8 == 0), "%s: out_plane must be a positive multiple of 8", who);
  VBG_REQUIRE(ep->res_plane >= 0 && ep->res_plane % 8 == 0 && (ep->res_plane == 0 || ep->residual), "%s: bad res_plane", who);
  return VBG_OK;
}

static bool valid_precision(int p) { return p == VBG_PREC_FP32 || p == VBG_PREC_TF32 || p == VBG_PREC_BF16X3; }

extern "C" int vbg_gemm(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int ldw, const void* W_split,
                        long long split_plane, float* C, int ldc, int M, int N, int K, const vbg_epilogue_t* ep,
                        int precision, vbg_stream_t stream) {
  VBG_REQUIRE(A && W && C, "vbg_gemm: null pointer");
  VBG_REQUIRE(M >= 0 && N > 0 && K > 0 && K1 > 0 && K1 <= K, "vbg_gemm: bad shape M=%d N=%d K=%d K1=%d", M, N, K, K1);
  VBG_REQUIRE((K1 == K) || A2, "vbg_gemm: A2 required when K1 < K");
  VBG_REQUIRE(lda >= K1 && ldw >= K && ldc >= N && (K1 == K || lda2 >= K - K1), "vbg_gemm: leading dimension too small");
  VBG_REQUIRE(valid_precision(precision), "vbg_gemm: bad precision %d", precision);
  int rc = check_epilogue(ep, "vbg_gemm");
  if (rc) return rc;
  if (ep && ep->res_mode == VBG_RES_UP2)
    VBG_REQUIRE(ep->out_h > 0 && ep->out_w > 0 && M % (ep->out_h * ep->out_w) == 0 && ep->out_h % 2 == 0 && ep->out_w % 2 == 0,
                "vbg_gemm: VBG_RES_UP2 needs even out_h/out_w dividing M");
  if (ep && ep->res_mode == VBG_RES_SAME) VBG_REQUIRE(ep->ldr >= N, "vbg_gemm: ldr too small");
  if (M == 0) return VBG_OK;
  cudaStream_t s = as_stream(stream);
  if (precision == VBG_PREC_BF16X3) {
    rc = gemm_tc3(A, lda, A2, lda2, K1, W_split, split_plane, ldw, C, ldc, M, N, K, ep, s);
    if (rc != VBG_EUNSUPPORTED) return rc;
  } else if (precision == VBG_PREC_TF32) {
    rc = gemm_tc(A, lda, A2, lda2, K1, W, ldw, C, ldc, M, N, K, ep, s);
    if (rc != VBG_EUNSUPPORTED) return rc;
  }
  VBG_REQUIRE(!ep || (ep->out_mode == VBG_OUT_F32 && ep->res_plane == 0),
              "vbg_gemm: bf16-plane output / residual needs a tensor-core path (shape / precision not eligible)");
  return gemm_simt(A, lda, A2, lda2, K1, W, ldw, C, ldc, M, N, K, ep, s);
}

extern "C" int vbg_conv2d(const float* x, int B, int H, int W, int Cin, const float* w, const void* w_split, long long split_plane,
                          int Cout, int kh, int kw, int stride, int pad, float* y, const vbg_epilogue_t* ep, int precision,
                          vbg_stream_t stream) {
  VBG_REQUIRE(x && w && y, "vbg_conv2d: null pointer");
  VBG_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0,
              "vbg_conv2d: bad geometry");
  VBG_REQUIRE(H + 2 * pad >= kh && W + 2 * pad >= kw, "vbg_conv2d: kernel larger than padded input");
  VBG_REQUIRE(valid_precision(precision), "vbg_conv2d: bad precision %d", precision);
  int rc = check_epilogue(ep, "vbg_conv2d");
  if (rc) return rc;
  if (ep && ep->res_mode == VBG_RES_UP2) {
    int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    VBG_REQUIRE(Ho % 2 == 0 && Wo % 2 == 0, "vbg_conv2d: VBG_RES_UP2 needs even output dims");
  }
  cudaStream_t s = as_stream(stream);
  if (precision == VBG_PREC_BF16X3) {
    rc = conv_tc3(x, B, H, W, Cin, w_split, split_plane, Cout, kh, kw, stride, pad, y, ep, s);
    if (rc != VBG_EUNSUPPORTED) return rc;
  } else if (precision == VBG_PREC_TF32) {
    rc = conv_tc(x, B, H, W, Cin, w, Cout, kh, kw, stride, pad, y, ep, s);
    if (rc != VBG_EUNSUPPORTED) return rc;
  }
  VBG_REQUIRE(!ep || (ep->out_mode == VBG_OUT_F32 && ep->res_plane == 0), "vbg_conv2d: bf16-plane output / residual needs a tensor-core path");
  return conv_simt(x, B, H, W, Cin, w, Cout, kh, kw, stride, pad, y, ep, s);
}

extern "C" int vbg_gemm_ps(const void* A_hi, long long a_plane, int lda, const void* A2_hi, long long a2_plane, int lda2, int K1,
                           const void* W_hi, long long w_plane, int ldw, void* C, int ldc, int M, int N, int K,
                           const vbg_epilogue_t* ep, void* workspace, size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(A_hi && W_hi && C, "vbg_gemm_ps: null pointer");
  VBG_REQUIRE(M >= 0 && N > 0 && K > 0 && K1 > 0 && K1 <= K, "vbg_gemm_ps: bad shape M=%d N=%d K=%d K1=%d", M, N, K, K1);
  VBG_REQUIRE((K1 == K) || A2_hi, "vbg_gemm_ps: A2 required when K1 < K");
  VBG_REQUIRE(lda >= K1 && ldw >= K && ldc >= N && (K1 == K || lda2 >= K - K1), "vbg_gemm_ps: leading dimension too small");
  int rc = check_epilogue(ep, "vbg_gemm_ps");
  if (rc) return rc;
  if (ep && ep->res_mode == VBG_RES_UP2)
    VBG_REQUIRE(ep->out_h > 0 && ep->out_w > 0 && M % (ep->out_h * ep->out_w) == 0 && ep->out_h % 2 == 0 && ep->out_w % 2 == 0,
                "vbg_gemm_ps: VBG_RES_UP2 needs even out_h/out_w dividing M");
  if (ep && ep->res_mode == VBG_RES_SAME) VBG_REQUIRE(ep->ldr >= N, "vbg_gemm_ps: ldr too small");
  if (M == 0) return VBG_OK;
  rc = gemm_ps(A_hi, a_plane, lda, A2_hi, a2_plane, lda2, K1, W_hi, w_plane, ldw, C, ldc, M, N, K, ep, workspace, ws_bytes,
               as_stream(stream));
  if (rc == VBG_EUNSUPPORTED)
    set_error("vbg_gemm_ps: needs sm_100a, N >= 64, K %% 64 == 0, K1 %% 64 == 0, lda/ldw %% 8 == 0, 16B-aligned planes with plane %% 8 == 0 "
              "(M=%d N=%d K=%d K1=%d)", M, N, K, K1);
  return rc;
}

extern "C" int vbg_conv2d_ps(const void* x_hi, long long x_plane, int B, int H, int W, int Cin, const void* w_hi, long long w_plane,
                             int Cout, int kh, int kw, int stride, int pad, void* y, const vbg_epilogue_t* ep, void* workspace,
                             size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(x_hi && w_hi && y, "vbg_conv2d_ps: null pointer");
  VBG_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "vbg_conv2d_ps: bad geometry");
  VBG_REQUIRE(H + 2 * pad >= kh && W + 2 * pad >= kw, "vbg_conv2d_ps: kernel larger than padded input");
  int rc = check_epilogue(ep, "vbg_conv2d_ps");
  if (rc) return rc;
  if (ep && ep->res_mode == VBG_RES_UP2) {
    int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    VBG_REQUIRE(Ho % 2 == 0 && Wo % 2 == 0, "vbg_conv2d_ps: VBG_RES_UP2 needs even output dims");
  }
  rc = conv_ps(x_hi, x_plane, B, H, W, Cin, w_hi, w_plane, Cout, kh, kw, stride, pad, y, ep, workspace, ws_bytes, as_stream(stream));
  if (rc == VBG_EUNSUPPORTED)
    set_error("vbg_conv2d_ps: needs sm_100a, Cin %% 64 == 0, Cout >= 64, stride 1 or 2, 16B-aligned planes (Cin=%d Cout=%d stride=%d)",
              Cin, Cout, stride);
  return rc;
}

namespace vbg { void tc_debug_set_timeline(long long* buf); }
extern "C" int vbg_debug_set_timeline(long long* dev_buf) {
  tc_debug_set_timeline(dev_buf);
  return VBG_OK;
}

extern "C" long long vbg_gemm_ps_workspace(int M, int N, int K, int tune) {
  return (M > 0 && N > 0 && K > 0) ? (long long)ps_workspace_bytes(cdiv(M, 128), M, N, K, tune) : 0;
}

extern "C" long long vbg_conv2d_ps_workspace(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int tune) {
  if (B <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0) return 0;
  return (long long)conv_ps_workspace_bytes(B, H, W, Cin, Cout, kh, kw, stride, pad, tune);
}

extern "C" long long vbg_linear_wgrad_workspace(int M, int N, int K) { return (long long)linear_wgrad_workspace(M, N, K); }

extern "C" int vbg_linear_wgrad(const void* dY_hi, long long y_plane, const void* X_hi, long long x_plane, int M, int N, int K, float* dW,
                                void* workspace, size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(dY_hi && X_hi && dW && M > 0 && N > 0 && K > 0, "vbg_linear_wgrad: bad arguments");
  int rc = linear_wgrad(dY_hi, y_plane, X_hi, x_plane, M, N, K, dW, workspace, ws_bytes, as_stream(stream));
  if (rc == VBG_EUNSUPPORTED)
    set_error("vbg_linear_wgrad: needs sm_100a, N %% 128 == 0, K %% 64 == 0, 16B-aligned bf16 planes (M=%d N=%d K=%d)", M, N, K);
  return rc;
}

extern "C" long long vbg_conv2d_wgrad_workspace(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad) {
  return (long long)conv_wgrad_workspace(B, H, W, Cin, Cout, kh, kw, stride, pad);
}

extern "C" int vbg_conv2d_wgrad(const void* dY_hi, long long y_plane, const void* X_hi, long long x_plane, int B, int H, int W, int Cin,
                                int Cout, int kh, int kw, int stride, int pad, float* dW, void* workspace, size_t ws_bytes,
                                vbg_stream_t stream) {
  VBG_REQUIRE(dY_hi && X_hi && dW && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0,
              "vbg_conv2d_wgrad: bad arguments");
  int rc = conv_wgrad(dY_hi, y_plane, X_hi, x_plane, B, H, W, Cin, Cout, kh, kw, stride, pad, dW, workspace, ws_bytes, as_stream(stream));
  if (rc == VBG_EUNSUPPORTED)
    set_error("vbg_conv2d_wgrad: needs sm_100a, Cout %% 128 == 0, Cin %% 64 == 0, stride 1 or 2, output tiles of exactly 64 pixels "
              "(Cin=%d Cout=%d H=%d W=%d)", Cin, Cout, H, W);
  return rc;
}

extern "C" int vbg_merge_bf16(const void* hi, const void* lo, long long n, float* out, vbg_stream_t stream) {
  VBG_REQUIRE(hi && lo && out && n >= 0, "vbg_merge_bf16: bad arguments");
  return merge_bf16(hi, lo, n, out, as_stream(stream));
}

extern "C" int vbg_stem_conv(const float* x4, int B, int H, int W, const float* w_ohwi4, const void* w_split, long long split_plane,
                             int Cout, float* y, const vbg_epilogue_t* ep, int precision, vbg_stream_t stream) {
  VBG_REQUIRE(x4 && w_ohwi4 && y && B > 0 && H > 0 && W > 0 && Cout > 0, "vbg_stem_conv: bad arguments");
  VBG_REQUIRE(valid_precision(precision), "vbg_stem_conv: bad precision %d", precision);
  int rc = check_epilogue(ep, "vbg_stem_conv");
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  if (precision == VBG_PREC_BF16X3) {
    rc = stem_tc3(x4, B, H, W, w_split, split_plane, Cout, y, ep, s);
    if (rc != VBG_EUNSUPPORTED) return rc;
  }
  // the border is already in the buffer: a plain 7x7 / stride-2 / pad-0 conv over [B, H+6, W+6, 4]
  return conv_simt(x4, B, H + 6, W + 6, 4, w_ohwi4, Cout, 7, 7, 2, 0, y, ep, s);
}

extern "C" int vbg_split_bf16(const float* w, long long n, void* hi, void* lo, vbg_stream_t stream) {
  VBG_REQUIRE(w && hi && lo && n >= 0, "vbg_split_bf16: bad arguments");
  return split_bf16(w, n, hi, lo, as_stream(stream));
}

extern "C" int vbg_crf_viterbi(const float* feats, const float* trans, const int32_t* seg_off, int B, int K, int T,
                                  float* tags, float* scores, void* workspace, size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(feats && trans && seg_off && tags && scores && B > 0 && T >= 3 && T <= 32, "vbg_crf_viterbi: bad arguments (T<=32)");
  if ((size_t)K * T > ws_bytes || !workspace) {
    set_error("vbg_crf_viterbi: workspace of %zu bytes needed", (size_t)K * T);
    return VBG_EWORKSPACE;
  }
  crf_viterbi_kernel<<<B, 32, 0, as_stream(stream)>>>(feats, trans, seg_off, T, tags, scores,
                                                     reinterpret_cast<unsigned char*>(workspace));
  return check_launch("vbg_crf_viterbi");
}
