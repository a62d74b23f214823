// BERT encoder support kernels: window packing, embedding+LayerNorm, LayerNorm, variable-length
// attention, row softmax.  The GEMMs (QKV / out / FFN) go through vbg_gemm.
//
// Replaces reference model/BERTgrid_generator.py:81-146 (Python loop over 510-token windows, each
// padded to 512 and run through HF BertModel sequentially).  Here all windows of all samples form
// ONE packed variable-length batch holding only real rows ([CLS] tokens [SEP]); padded keys have
// softmax weight exactly 0 in the reference (additive finfo.min mask), so dropping them changes
// results only by fp32 re-association.
#include "vbg_common.cuh"

namespace vbg {

// row r of sequence q: 0 -> [CLS]=101 (pos 0); 1..n -> corpus[b, col0 + r-1] (pos r); n+1 -> [SEP]=102 (pos sep_pos)
__global__ void bert_assemble_kernel(const int64_t* __restrict__ corpus, int L, const int32_t* __restrict__ seq_tab,
                                     const int32_t* __restrict__ cu, int nseq, int32_t* __restrict__ ids,
                                     int32_t* __restrict__ pos) {
  const int q = blockIdx.x;
  if (q >= nseq) return;
  const int b = seq_tab[4 * q], col0 = seq_tab[4 * q + 1], n = seq_tab[4 * q + 2], sep = seq_tab[4 * q + 3];
  const int r0 = cu[q];
  for (int r = threadIdx.x; r < n + 2; r += blockDim.x) {
    int id, p;
    if (r == 0) { id = 101; p = 0; }
    else if (r == n + 1) { id = 102; p = sep; }
    else { id = (int)corpus[(size_t)b * L + col0 + r - 1]; p = r; }
    ids[r0 + r] = id;
    pos[r0 + r] = p;
  }
}

// One warp per row, row kept in registers (hidden <= 1024), two-pass mean / variance in fp32.
template <bool kEmbed>
__global__ void __launch_bounds__(256)
ln_kernel(const float* __restrict__ x, const int32_t* __restrict__ ids, const int32_t* __restrict__ pos,
          const float* __restrict__ word, const float* __restrict__ position, const float* __restrict__ type0,
          const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int R, int H4, int vocab,
          int max_pos, void* __restrict__ out, long long out_plane) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  constexpr int kMax = 8;              // 8 float4 per lane -> hidden <= 1024
  float4 v[kMax];
  float sum = 0.f;
  const float4 *w4 = nullptr, *p4 = nullptr, *x4 = nullptr;
  if (kEmbed) {
    int id = ids[r], ps = pos[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);      // clamp instead of faulting on a bad token id
    ps = ps < 0 ? 0 : (ps >= max_pos ? max_pos - 1 : ps);
    w4 = reinterpret_cast<const float4*>(word) + (size_t)id * H4;
    p4 = reinterpret_cast<const float4*>(position) + (size_t)ps * H4;
  } else {
    x4 = reinterpret_cast<const float4*>(x) + (size_t)r * H4;
  }
  const float4* t4 = reinterpret_cast<const float4*>(type0);
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < H4) {
      float4 a;
      if (kEmbed) {
        float4 w = __ldg(w4 + c), t = __ldg(t4 + c), p = __ldg(p4 + c);
        // HF order: inputs_embeds + token_type_embeddings, then + position_embeddings
        a.x = (w.x + t.x) + p.x; a.y = (w.y + t.y) + p.y; a.z = (w.z + t.z) + p.z; a.w = (w.w + t.w) + p.w;
      } else {
        a = __ldg(x4 + c);
      }
      v[i] = a;
      sum += (a.x + a.y) + (a.z + a.w);
    }
  }
  const float n = (float)(H4 * 4);
  const float mean = warp_sum(sum) / n;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < H4) {
      float dx = v[i].x - mean, dy = v[i].y - mean, dz = v[i].z - mean, dw = v[i].w - mean;
      sq += (dx * dx + dy * dy) + (dz * dz + dw * dw);
    }
  }
  const float rstd = __fdiv_rn(1.0f, sqrtf(warp_sum(sq) / n + eps));
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < kMax; ++i) {
    const int c = lane + 32 * i;
    if (c < H4) {
      float4 g = __ldg(g4 + c), b = __ldg(b4 + c), o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x; o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z; o.w = (v[i].w - mean) * rstd * g.w + b.w;
      st4_fmt(out, out_plane, (size_t)r * H4 + c, o);
    }
  }
}

// ------------------------------------------------------------------ attention (fp32 CUDA-core path)
// CTA = 64 queries of one (sequence, head); streams 64-key tiles with an online softmax.
// 256 threads as 16x16; thread (ty,tx) owns S[4ty..][4tx..] and O[4ty..][4tx..] (head_dim 64).
constexpr int kAttQ = 64, kAttK = 64, kAttD = 64;

__global__ void __launch_bounds__(256)
attention_simt_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu, int heads, float inv_sqrt_d,
                      float* __restrict__ out) {
  extern __shared__ __align__(16) float att_smem[];
  float (*Vs)[kAttD] = reinterpret_cast<float (*)[kAttD]>(att_smem);                       // 16B aligned rows
  float (*Qs)[kAttD + 1] = reinterpret_cast<float (*)[kAttD + 1]>(att_smem + kAttK * kAttD);
  float (*Ks)[kAttD + 1] = reinterpret_cast<float (*)[kAttD + 1]>(att_smem + kAttK * kAttD + kAttQ * (kAttD + 1));
  float (*Ps)[kAttK + 1] = reinterpret_cast<float (*)[kAttK + 1]>(att_smem + kAttK * kAttD + (kAttQ + kAttK) * (kAttD + 1));
  const int seq = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int r0 = cu[seq], len = cu[seq + 1] - r0;
  const int q0 = qt * kAttQ;
  if (q0 >= len) return;
  const int hidden = heads * kAttD, ld = 3 * hidden;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  for (int i = tid; i < kAttQ * (kAttD / 4); i += 256) {
    int r = i / (kAttD / 4), c4 = i % (kAttD / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < len) v = __ldg(reinterpret_cast<const float4*>(qkv + (size_t)(r0 + q0 + r) * ld + head * kAttD) + c4);
    Qs[r][4 * c4] = v.x; Qs[r][4 * c4 + 1] = v.y; Qs[r][4 * c4 + 2] = v.z; Qs[r][4 * c4 + 3] = v.w;
  }
  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  for (int k0 = 0; k0 < len; k0 += kAttK) {
    __syncthreads();
    for (int i = tid; i < kAttK * (kAttD / 4); i += 256) {
      int r = i / (kAttD / 4), c4 = i % (kAttD / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < len) {
        const float* base = qkv + (size_t)(r0 + k0 + r) * ld + head * kAttD;
        kv = __ldg(reinterpret_cast<const float4*>(base + hidden) + c4);
        vv = __ldg(reinterpret_cast<const float4*>(base + 2 * hidden) + c4);
      }
      Ks[r][4 * c4] = kv.x; Ks[r][4 * c4 + 1] = kv.y; Ks[r][4 * c4 + 2] = kv.z; Ks[r][4 * c4 + 3] = kv.w;
      *reinterpret_cast<float4*>(&Vs[r][4 * c4]) = vv;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < kAttD; ++d) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = Qs[4 * ty + i][d];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Ks[4 * tx + j][d];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], b[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + 4 * tx + j < len) ? s[i][j] * inv_sqrt_d : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);          // finite: key 0 ([CLS]) is always valid
      const float alpha = expf(m_run[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float p = expf(s[i][j] - m_new);
        Ps[4 * ty + i][4 * tx + j] = p;
        rs += p;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[i] = l_run[i] * alpha + rs;
      m_run[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= alpha;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < kAttK; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(&Vs[j][4 * tx]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = Ps[4 * ty + i][j];
        o[i][0] = fmaf(p, v.x, o[i][0]); o[i][1] = fmaf(p, v.y, o[i][1]);
        o[i][2] = fmaf(p, v.z, o[i][2]); o[i][3] = fmaf(p, v.w, o[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + 4 * ty + i;
    if (q < len) {
      const float inv = __fdiv_rn(1.0f, l_run[i]);
      float4 r = make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
      *reinterpret_cast<float4*>(out + (size_t)(r0 + q) * hidden + head * kAttD + 4 * tx) = r;
    }
  }
}

// ------------------------------------------------------------------ small row ops
__global__ void softmax_rows_kernel(const float* __restrict__ x, int R, int C, float* __restrict__ y) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float* p = x + (size_t)r * C;
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, p[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(p[c] - m);
  for (int c = 0; c < C; ++c) y[(size_t)r * C + c] = __fdiv_rn(expf(p[c] - m), s);
}

__global__ void full_head_scores_kernel(const float* __restrict__ pn, const float* __restrict__ cls, int R, int C,
                                        float* __restrict__ out) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float p = __fdiv_rn(1.0f, 1.0f + expf(-pn[r]));
  out[(size_t)r * C] = p;
  bool gate = p >= 0.5f;
  for (int c = 1; c < C; ++c)
    out[(size_t)r * C + c] = gate ? __fdiv_rn(1.0f, 1.0f + expf(-cls[(size_t)r * (C - 1) + c - 1])) : 0.f;
}

int attention_tc(const float* qkv, const int32_t* cu, int nseq, int max_len, int heads, int head_dim, float* out,
                 cudaStream_t s);
int attention_split(const void* qkv_hi, long long plane, const int32_t* cu, int nseq, int R, int max_len, int heads,
                    int head_dim, void* out, long long out_plane, float* lse2, float p_drop, unsigned long long seed,
                    const unsigned long long* step_seed, cudaStream_t s);

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_bert_assemble(const int64_t* corpus, int L, const int32_t* seq_tab, const int32_t* cu, int nseq, int R,
                                 int32_t* ids, int32_t* pos, vbg_stream_t stream) {
  VBG_REQUIRE(corpus && seq_tab && cu && ids && pos && L > 0 && nseq >= 0 && R >= 0, "vbg_bert_assemble: bad arguments");
  if (nseq == 0) return VBG_OK;
  bert_assemble_kernel<<<nseq, 256, 0, as_stream(stream)>>>(corpus, L, seq_tab, cu, nseq, ids, pos);
  return check_launch("vbg_bert_assemble");
}

extern "C" int vbg_embed_ln_x(const int32_t* ids, const int32_t* pos, const float* word, const float* position,
                              const float* type0, const float* gamma, const float* beta, float eps, int R, int hidden,
                              int vocab, int max_pos, void* out, long long out_plane, vbg_stream_t stream) {
  VBG_REQUIRE(ids && pos && word && position && type0 && gamma && beta && out, "vbg_embed_ln: null pointer");
  VBG_REQUIRE(hidden % 4 == 0 && hidden <= 1024 && vocab > 0 && max_pos > 0, "vbg_embed_ln: hidden %% 4 == 0 and <= 1024 required");
  VBG_REQUIRE(fmt_ok(out, out_plane), "vbg_embed_ln: output must be 16B aligned, plane %% 8 == 0");
  if (R == 0) return VBG_OK;
  ln_kernel<true><<<cdiv(R, 8), 256, 0, as_stream(stream)>>>(nullptr, ids, pos, word, position, type0, gamma, beta, eps,
                                                            R, hidden / 4, vocab, max_pos, out, out_plane);
  return check_launch("vbg_embed_ln");
}

extern "C" int vbg_embed_ln(const int32_t* ids, const int32_t* pos, const float* word, const float* position,
                            const float* type0, const float* gamma, const float* beta, float eps, int R, int hidden,
                            int vocab, int max_pos, float* out, vbg_stream_t stream) {
  return vbg_embed_ln_x(ids, pos, word, position, type0, gamma, beta, eps, R, hidden, vocab, max_pos, out, 0, stream);
}

extern "C" int vbg_layernorm_x(const float* x, const float* gamma, const float* beta, float eps, int R, int hidden,
                               void* out, long long out_plane, vbg_stream_t stream) {
  VBG_REQUIRE(x && gamma && beta && out, "vbg_layernorm: null pointer");
  VBG_REQUIRE(hidden % 4 == 0 && hidden <= 1024 && aligned16(x) && fmt_ok(out, out_plane),
              "vbg_layernorm: hidden %% 4 == 0, <= 1024, 16B alignment, plane %% 8 == 0");
  if (R == 0) return VBG_OK;
  ln_kernel<false><<<cdiv(R, 8), 256, 0, as_stream(stream)>>>(x, nullptr, nullptr, nullptr, nullptr, nullptr, gamma, beta,
                                                             eps, R, hidden / 4, 0, 0, out, out_plane);
  return check_launch("vbg_layernorm");
}

extern "C" int vbg_layernorm(const float* x, const float* gamma, const float* beta, float eps, int R, int hidden,
                             float* out, vbg_stream_t stream) {
  return vbg_layernorm_x(x, gamma, beta, eps, R, hidden, out, 0, stream);
}

extern "C" int vbg_attention_fwd(const float* qkv, const int32_t* cu, int nseq, int max_len, int heads, int head_dim,
                                 float* out, int precision, vbg_stream_t stream) {
  VBG_REQUIRE(qkv && cu && out && nseq >= 0 && heads > 0, "vbg_attention_fwd: bad arguments");
  VBG_REQUIRE(head_dim == kAttD, "vbg_attention_fwd: head_dim must be 64 (got %d)", head_dim);
  VBG_REQUIRE(aligned16(qkv) && aligned16(out), "vbg_attention_fwd: 16B alignment required");
  if (nseq == 0 || max_len == 0) return VBG_OK;
  if (precision != VBG_PREC_FP32) {              // tensor-core kernel (fp32-class 3-term bf16 split) when it takes the shape
    int rc = attention_tc(qkv, cu, nseq, max_len, heads, head_dim, out, as_stream(stream));
    if (rc != VBG_EUNSUPPORTED) return rc;
  }
  dim3 g(cdiv(max_len, kAttQ), heads, nseq);
  constexpr size_t smem = sizeof(float) * (kAttK * kAttD + (kAttQ + kAttK) * (kAttD + 1) + kAttQ * (kAttK + 1));
  static bool attr_set = false;   // idempotent; a race only repeats the same call
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("vbg_attention_fwd: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
    attr_set = true;
  }
  attention_simt_kernel<<<g, 256, smem, as_stream(stream)>>>(qkv, cu, heads, 1.0f / sqrtf((float)head_dim), out);
  return check_launch("vbg_attention_fwd");
}

extern "C" int vbg_attention_split_fwd(const void* qkv_hi, long long plane, const int32_t* cu, int nseq, int R, int max_len,
                                       int heads, int head_dim, void* out, long long out_plane, vbg_stream_t stream) {
  VBG_REQUIRE(qkv_hi && cu && out && nseq >= 0 && R >= 0 && heads > 0 && plane > 0, "vbg_attention_split_fwd: bad arguments");
  VBG_REQUIRE(fmt_ok(out, out_plane), "vbg_attention_split_fwd: output must be 16B aligned, plane %% 8 == 0");
  if (nseq == 0 || max_len == 0) return VBG_OK;
  int rc = attention_split(qkv_hi, plane, cu, nseq, R, max_len, heads, head_dim, out, out_plane, nullptr, 0.f, 0ull, nullptr, as_stream(stream));
  if (rc == VBG_EUNSUPPORTED)
    set_error("vbg_attention_split_fwd: needs sm_100a, head_dim 64, max_len <= 512 (got head_dim %d, max_len %d)", head_dim, max_len);
  return rc;
}

extern "C" int vbg_attention_split_train_fwd(const void* qkv_hi, long long plane, const int32_t* cu, int nseq, int R, int max_len,
                                             int heads, int head_dim, void* out, long long out_plane, float* lse2, float p_drop,
                                             unsigned long long seed, const unsigned long long* step_seed, vbg_stream_t stream) {
  VBG_REQUIRE(qkv_hi && cu && out && lse2 && nseq >= 0 && R >= 0 && heads > 0 && plane > 0, "vbg_attention_split_train_fwd: bad arguments");
  VBG_REQUIRE(fmt_ok(out, out_plane), "vbg_attention_split_train_fwd: output must be 16B aligned, plane %% 8 == 0");
  VBG_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "vbg_attention_split_train_fwd: 0 <= p_drop < 1");
  if (nseq == 0 || max_len == 0) return VBG_OK;
  int rc = attention_split(qkv_hi, plane, cu, nseq, R, max_len, heads, head_dim, out, out_plane, lse2, p_drop, seed, step_seed, as_stream(stream));
  if (rc == VBG_EUNSUPPORTED)
    set_error("vbg_attention_split_train_fwd: needs sm_100a, head_dim 64, max_len <= 512 (got head_dim %d, max_len %d)", head_dim, max_len);
  return rc;
}

extern "C" int vbg_softmax_rows(const float* x, int R, int C, float* y, vbg_stream_t stream) {
  VBG_REQUIRE(x && y && R >= 0 && C > 0, "vbg_softmax_rows: bad arguments");
  if (R == 0) return VBG_OK;
  softmax_rows_kernel<<<cdiv(R, 128), 128, 0, as_stream(stream)>>>(x, R, C, y);
  return check_launch("vbg_softmax_rows");
}

extern "C" int vbg_full_head_scores(const float* pos_neg, const float* cls, int R, int C, float* out, vbg_stream_t stream) {
  VBG_REQUIRE(pos_neg && cls && out && R >= 0 && C > 1, "vbg_full_head_scores: bad arguments");
  if (R == 0) return VBG_OK;
  full_head_scores_kernel<<<cdiv(R, 128), 128, 0, as_stream(stream)>>>(pos_neg, cls, R, C, out);
  return check_launch("vbg_full_head_scores");
}
