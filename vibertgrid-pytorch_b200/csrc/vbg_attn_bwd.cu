// Self-attention backward over the packed varlen batch (HF BertSelfAttention inside model/BERTgrid_generator.py:81-146, reached by
// loss.backward() at pipeline/train_val_utils.py:277).  Head dimension 64, fp32 CUDA cores: the first correct version of this
// stage; the probabilities are recomputed from Q, K and the row log-sum-exp (nothing of size L x L is ever stored).
//
//   S = Q K^T / 8, P = softmax(S), O = P V                       (forward, vbg_attention_split_fwd)
//   delta_q = sum_d dO[q,d] O[q,d]
//   dP = dO V^T, dS = P o (dP - delta) / 8
//   dQ = dS K            kernel A: one CTA per (64 queries, head, sequence): pass 1 row LSE, pass 2 dQ
//   dK = dS^T Q, dV = P^T dO   kernel B: one CTA per (64 keys, head, sequence)
//
// Every product is a 64x64x64 tile product with a 4x4 register block per thread (256 threads); operands sit in shared memory
// with the reduction index as the row ("transposed" tiles for Q K^T and dO V^T), so each step is two 128-bit shared loads for
// 16 FMAs.  Deterministic (no atomics).
#include "vbg_common.cuh"

namespace vbg {

constexpr int kT = 64;            // tile edge = head dimension
constexpr int kLd = 68;           // shared row stride in floats (16-byte aligned rows)
constexpr int kTile = kT * kLd;   // floats per tile

// global [rows r0.., 64 columns at col0] -> shared tile, natural ([row][d]) and / or transposed ([d][row]); rows >= n_valid are zero
__device__ __forceinline__ void load_tile(const float* __restrict__ g, long long ld, long long r0, int n_valid, int col0, float* nat, float* tr) {
  for (int idx = threadIdx.x; idx < kT * 16; idx += 256) {
    const int r = idx >> 4, d4 = idx & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < n_valid) v = __ldg(reinterpret_cast<const float4*>(g + (r0 + r) * ld + col0) + d4);
    if (nat) *reinterpret_cast<float4*>(nat + r * kLd + 4 * d4) = v;
    if (tr) {
      tr[(4 * d4 + 0) * kLd + r] = v.x; tr[(4 * d4 + 1) * kLd + r] = v.y;
      tr[(4 * d4 + 2) * kLd + r] = v.z; tr[(4 * d4 + 3) * kLd + r] = v.w;
    }
  }
}

// acc[i][j] += sum_k A[k][4*ty + i] * B[k][4*tx + j]     (both tiles reduction-major)
__device__ __forceinline__ void mm64(const float* __restrict__ A, const float* __restrict__ B, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll 8
  for (int k = 0; k < kT; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(A + k * kLd + 4 * ty);
    const float4 b = *reinterpret_cast<const float4*>(B + k * kLd + 4 * tx);
    acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]); acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
    acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]); acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
    acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]); acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
    acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]); acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
  }
}

__device__ __forceinline__ void zero16(float (&a)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}

// reductions across the 16 threads that share a row block (lanes with equal lane / 16)
__device__ __forceinline__ float row_max16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float row_sum16(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ kernel A: LSE, delta, dQ
__global__ void __launch_bounds__(256)
attn_bwd_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ o, const float* __restrict__ d_o, const int32_t* __restrict__ cu, int heads,
                   float scale, float* __restrict__ dqkv, float* __restrict__ lse, float* __restrict__ delta) {
  extern __shared__ float sm[];
  float *Qt = sm, *dOt = sm + kTile, *Kt = sm + 2 * kTile, *Vt = sm + 3 * kTile, *Kn = sm + 4 * kTile, *dSt = sm + 5 * kTile;
  float* delta_s = sm + 6 * kTile;      // [64]
  float* lse_s = delta_s + kT;          // [64]
  const int seq = blockIdx.z, h = blockIdx.y;
  const long long rb = cu[seq];
  const int len = cu[seq + 1] - (int)rb;
  const int q0 = blockIdx.x * kT;
  if (q0 >= len) return;
  const int nq = min(kT, len - q0);
  const int hid = heads * kT;
  const long long ld3 = 3LL * hid;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  load_tile(qkv, ld3, rb + q0, nq, h * kT, nullptr, Qt);
  load_tile(d_o, hid, rb + q0, nq, h * kT, nullptr, dOt);
  // delta: 16 lanes per row
  for (int idx = threadIdx.x; idx < kT * 16; idx += 256) {
    const int r = idx >> 4, d4 = idx & 15;
    float part = 0.f;
    if (r < nq) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(d_o + (rb + q0 + r) * hid + h * kT) + d4);
      const float4 b = __ldg(reinterpret_cast<const float4*>(o + (rb + q0 + r) * hid + h * kT) + d4);
      part = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
    part = row_sum16(part);
    if (d4 == 0) delta_s[r] = part;
  }

  // pass 1: row log-sum-exp
  float m[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[i] = -INFINITY; l[i] = 0.f; }
  for (int k0 = 0; k0 < len; k0 += kT) {
    __syncthreads();
    load_tile(qkv, ld3, rb + k0, min(kT, len - k0), hid + h * kT, nullptr, Kt);
    __syncthreads();
    float s[4][4]; zero16(s);
    mm64(Qt, Kt, ty, tx, s);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + 4 * tx + j < len) ? s[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
      mx = row_max16(mx);
      const float mn = fmaxf(m[i], mx);
      float e = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) e += expf(s[i][j] - mn);
      e = row_sum16(e);
      l[i] = l[i] * expf(m[i] - mn) + e;
      m[i] = mn;
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) lse_s[4 * ty + i] = m[i] + logf(l[i]);
  }
  __syncthreads();
  if (threadIdx.x < nq) {
    lse[(rb + q0 + threadIdx.x) * heads + h] = lse_s[threadIdx.x];
    delta[(rb + q0 + threadIdx.x) * heads + h] = delta_s[threadIdx.x];
  }

  // pass 2: dQ
  float dq[4][4]; zero16(dq);
  float lq[4], dl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { lq[i] = lse_s[4 * ty + i]; dl[i] = delta_s[4 * ty + i]; }
  for (int k0 = 0; k0 < len; k0 += kT) {
    const int nk = min(kT, len - k0);
    __syncthreads();
    load_tile(qkv, ld3, rb + k0, nk, hid + h * kT, Kn, Kt);
    load_tile(qkv, ld3, rb + k0, nk, 2 * hid + h * kT, nullptr, Vt);
    __syncthreads();
    float s[4][4], dp[4][4]; zero16(s); zero16(dp);
    mm64(Qt, Kt, ty, tx, s);
    mm64(dOt, Vt, ty, tx, dp);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (k0 + 4 * tx + j < len) && (4 * ty + i < nq);
        const float p = ok ? expf(s[i][j] * scale - lq[i]) : 0.f;
        dSt[(4 * tx + j) * kLd + 4 * ty + i] = p * (dp[i][j] - dl[i]) * scale;
      }
    __syncthreads();
    mm64(dSt, Kn, ty, tx, dq);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (4 * ty + i < nq)
      *reinterpret_cast<float4*>(dqkv + (rb + q0 + 4 * ty + i) * ld3 + h * kT + 4 * tx) = make_float4(dq[i][0], dq[i][1], dq[i][2], dq[i][3]);
}

// ------------------------------------------------------------------ kernel B: dK, dV
__global__ void __launch_bounds__(256)
attn_bwd_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, const int32_t* __restrict__ cu, int heads, float scale,
                    const float* __restrict__ lse, const float* __restrict__ delta, float* __restrict__ dqkv) {
  extern __shared__ float sm[];
  // six tiles (105 KB: two CTAs per SM): P and dS overwrite the transposed Q / dO tiles once S and dP have been formed
  float *Kt = sm, *Vt = sm + kTile, *Qt = sm + 2 * kTile, *dOt = sm + 3 * kTile, *Qn = sm + 4 * kTile, *dOn = sm + 5 * kTile, *Pn = Qt, *dSn = dOt;
  float* lse_s = sm + 6 * kTile;
  float* delta_s = lse_s + kT;
  const int seq = blockIdx.z, h = blockIdx.y;
  const long long rb = cu[seq];
  const int len = cu[seq + 1] - (int)rb;
  const int k0 = blockIdx.x * kT;
  if (k0 >= len) return;
  const int nk = min(kT, len - k0);
  const int hid = heads * kT;
  const long long ld3 = 3LL * hid;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  load_tile(qkv, ld3, rb + k0, nk, hid + h * kT, nullptr, Kt);
  load_tile(qkv, ld3, rb + k0, nk, 2 * hid + h * kT, nullptr, Vt);
  float dk[4][4], dv[4][4]; zero16(dk); zero16(dv);
  for (int q0 = 0; q0 < len; q0 += kT) {
    const int nq = min(kT, len - q0);
    __syncthreads();
    load_tile(qkv, ld3, rb + q0, nq, h * kT, Qn, Qt);
    load_tile(d_o, hid, rb + q0, nq, h * kT, dOn, dOt);
    if (threadIdx.x < kT) {
      const bool ok = threadIdx.x < nq;
      lse_s[threadIdx.x] = ok ? __ldg(lse + (rb + q0 + threadIdx.x) * heads + h) : 0.f;
      delta_s[threadIdx.x] = ok ? __ldg(delta + (rb + q0 + threadIdx.x) * heads + h) : 0.f;
    }
    __syncthreads();
    float s[4][4], dp[4][4]; zero16(s); zero16(dp);
    mm64(Qt, Kt, ty, tx, s);           // rows q = 4 ty + i, columns k = 4 tx + j
    mm64(dOt, Vt, ty, tx, dp);
    __syncthreads();                   // every thread is done reading Qt / dOt: their storage becomes P / dS
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float pr[4], ds[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (4 * tx + j < nk) && (4 * ty + i < nq);
        pr[j] = ok ? expf(s[i][j] * scale - lse_s[4 * ty + i]) : 0.f;
        ds[j] = pr[j] * (dp[i][j] - delta_s[4 * ty + i]) * scale;
      }
      *reinterpret_cast<float4*>(Pn + (4 * ty + i) * kLd + 4 * tx) = make_float4(pr[0], pr[1], pr[2], pr[3]);
      *reinterpret_cast<float4*>(dSn + (4 * ty + i) * kLd + 4 * tx) = make_float4(ds[0], ds[1], ds[2], ds[3]);
    }
    __syncthreads();
    mm64(Pn, dOn, ty, tx, dv);         // rows k = 4 ty + i, columns d = 4 tx + j, reduction over q
    mm64(dSn, Qn, ty, tx, dk);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (4 * ty + i < nk) {
      float* row = dqkv + (rb + k0 + 4 * ty + i) * ld3 + h * kT + 4 * tx;
      *reinterpret_cast<float4*>(row + hid) = make_float4(dk[i][0], dk[i][1], dk[i][2], dk[i][3]);
      *reinterpret_cast<float4*>(row + 2 * hid) = make_float4(dv[i][0], dv[i][1], dv[i][2], dv[i][3]);
    }
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_attention_bwd(const float* qkv, const float* out, const float* d_out, const int32_t* cu, int nseq, int max_len, int heads,
                                 int head_dim, long long rows, float* dqkv, float* workspace, size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(qkv && out && d_out && cu && dqkv && workspace && nseq > 0 && max_len > 0 && heads > 0 && rows > 0 && aligned16(qkv) &&
                  aligned16(out) && aligned16(d_out) && aligned16(dqkv),
              "vbg_attention_bwd: bad arguments");
  VBG_REQUIRE(head_dim == kT, "vbg_attention_bwd: head dimension 64 only");
  if ((size_t)rows * heads * 2 * sizeof(float) > ws_bytes) {
    set_error("vbg_attention_bwd: workspace of %zu bytes needed", (size_t)rows * heads * 2 * sizeof(float));
    return VBG_EWORKSPACE;
  }
  float* lse = workspace;
  float* delta = workspace + (size_t)rows * heads;
  const float scale = 1.0f / sqrtf((float)head_dim);
  const size_t smem_a = (size_t)(6 * kTile + 2 * kT) * sizeof(float), smem_b = (size_t)(6 * kTile + 2 * kT) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a);
    cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
    attr_done = true;
  }
  cudaStream_t s = as_stream(stream);
  dim3 grid((max_len + kT - 1) / kT, heads, nseq);
  attn_bwd_dq_kernel<<<grid, 256, smem_a, s>>>(qkv, out, d_out, cu, heads, scale, dqkv, lse, delta);
  int rc = check_launch("vbg_attention_bwd(dQ)");
  if (rc) return rc;
  attn_bwd_dkv_kernel<<<grid, 256, smem_b, s>>>(qkv, d_out, cu, heads, scale, lse, delta, dqkv);
  return check_launch("vbg_attention_bwd(dK, dV)");
}
