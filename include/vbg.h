/*
 * vbg.h -- C ABI of libvbg_sm100a.so: the ViBERTgrid joint-forward hot path as
 * hand-written CUDA for NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  The reference is pure Python
 * and reaches its arithmetic through torch / torchvision / transformers; each
 * entry point below names the reference call site it replaces (paths relative
 * to the reference repository root).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every function returns 0 on success or a negative VBG_E* code; it never
 *     throws, exits or synchronises the device.  vbg_last_error() returns the
 *     calling thread's last message.
 *   - all pointers are DEVICE pointers unless the name starts with h_.
 *     The caller owns every buffer (including workspaces); the library
 *     allocates nothing persistent.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); the legacy
 *     default stream is never used implicitly.  Functions are re-entrant.
 *   - activations are channels-last (NHWC), stored as fp32 or -- between the
 *     tensor-core kernels of VBG_PREC_BF16X3 -- as a PAIR OF BF16 PLANES
 *     (hi = bf16_rn(x), lo = bf16_rn(x - hi); the lo plane starts `plane`
 *     ELEMENTS after the hi plane, plane % 8 == 0).  Entry points ending in _x
 *     and _ps take (pointer, plane) pairs: plane == 0 means fp32.  Weights of
 *     convolutions are [Cout, kh, kw, Cin] (see vbg_repack_oihw_to_ohwi);
 *     linear weights keep PyTorch's [out, in] layout.
 *   - "seg_off" is an int32 [B+1] exclusive prefix sum of segments per sample,
 *     built by the host from tensor SHAPES (no device sync).
 */
#ifndef VBG_H_
#define VBG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBG_VERSION 100 /* 0.1.0 */

enum { VBG_OK = 0, VBG_EINVAL = -1, VBG_ECUDA = -2, VBG_EUNSUPPORTED = -3, VBG_EWORKSPACE = -4 };

/* arithmetic path of the dense contractions */
enum { VBG_PREC_FP32 = 0,   /* CUDA-core FFMA, fp32 operands and accumulate                                   */
       VBG_PREC_TF32 = 1,   /* tcgen05.mma kind::tf32 (operands truncated to 10-bit mantissa): fast, ~5e-3 fwd  */
       VBG_PREC_BF16X3 = 2  /* 3 x tcgen05.mma kind::f16 on bf16 hi/lo splits, fp32 accumulate: fp32-class, the
                               parity-grade tensor-core mode (needs the split weights of vbg_split_bf16)       */ };

enum { VBG_ACT_NONE = 0, VBG_ACT_RELU = 1, VBG_ACT_GELU = 2 /* erf form, as HF "gelu" */ };
enum { VBG_RES_NONE = 0, VBG_RES_SAME = 1, /* residual[m*ldr + n]                                  */
       VBG_RES_UP2 = 2   /* residual is NHWC [B, out_h/2, out_w/2, N]: nearest x2 upsample-add   */ };
enum { VBG_AGG_MEAN = 0, VBG_AGG_FIRST = 1 };
enum { VBG_OUT_F32 = 0,        /* C is float [M, ldc]                                                                  */
       VBG_OUT_SPLIT_BF16 = 1  /* C is bf16: hi plane [M, ldc] at C, lo plane out_plane ELEMENTS after it (tensor-core
                                  paths only): the operand format of vbg_attention_split_fwd                          */ };

typedef void* vbg_stream_t;

#if defined(__GNUC__)
#define VBG_API __attribute__((visibility("default")))
#else
#define VBG_API
#endif

/* Fused epilogue of vbg_gemm / vbg_conv2d:  y = act( acc * scale[n] + shift[n] + residual ) */
typedef struct vbg_epilogue {
  const float* scale;    /* [N] or NULL (= 1)  -- folded BatchNorm gamma/sqrt(var+eps)          */
  const float* shift;    /* [N] or NULL (= 0)  -- bias, or folded BatchNorm beta - mean*scale   */
  const float* residual; /* or NULL */
  int res_mode;          /* VBG_RES_* */
  int ldr;               /* row stride of residual for VBG_RES_SAME */
  int out_h, out_w;      /* output spatial dims, needed by VBG_RES_UP2 for vbg_gemm            */
  int act;               /* VBG_ACT_* */
  int out_mode;          /* VBG_OUT_* */
  long long out_plane;   /* VBG_OUT_SPLIT_BF16: elements between the hi and the lo plane */
  long long res_plane;   /* > 0: `residual` points at the bf16 hi plane of a split activation (same indexing), the lo
                            plane res_plane elements later (tensor-core paths only); 0: residual is float */
  int tune;              /* VBG_TUNE_* bits, 0 = the library's own tile / pipeline heuristics (what ships) */
} vbg_epilogue_t;

/* Per-call overrides of the pre-split tensor-core kernels' variant choice (vbg_gemm_ps / vbg_conv2d_ps and their
 * *_workspace queries).  The library reads no environment variable and keeps no mutable tuning state: a sweep or a
 * parity test that wants a specific variant says so in the call. */
enum { VBG_TUNE_KB32 = 1,      /* 32-element K ring stages (SWIZZLE_64B) instead of 64 (SWIZZLE_128B)        */
       VBG_TUNE_PAIRS_OFF = 2, /* never use CTA-pair (cta_group::2) tiles                                    */
       VBG_TUNE_PAIRS_ON = 4,  /* CTA-pair tiles whenever the problem has two row tiles                      */
       VBG_TUNE_SPLITK = 8,    /* allow split-K work units for under-filled grids (needs the workspace)      */
       VBG_TUNE_NO_PDL = 16    /* launch without programmatic stream serialization                           */ };

VBG_API int vbg_version(void);
/* copies the calling thread's last error text into buf (NUL terminated); returns its length */
VBG_API int vbg_last_error(char* buf, size_t n);
/* 1 if the tcgen05/TMA path can run on the current device + driver, else 0 */
VBG_API int vbg_tc_available(void);

/* ---- a1: GeneralizedViBERTgridTransform (pipeline/transform.py:104-171, 225-312) ---------- */
/* One source image [3,h,w] (CHW, [0,1]) -> normalised, bilinearly resized to (oh,ow)
 * (align_corners=False, scale = in/out as F.interpolate(recompute_scale_factor=True) uses),
 * written into sample b of the zero-initialised batch [B, H+6, W+6, 4]: NHWC with a 4th zero channel and a
 * 3-pixel zero border (= the stem conv's padding), pixel (y,x) at [b, y+3, x+3, :].            */
VBG_API int vbg_normalize_resize_pad(const float* img_chw, int h, int w, float* batch_nhwc, int b, int H, int W,
                             int oh, int ow, const float* h_mean3, const float* h_std3, vbg_stream_t stream);
/* the same for n same-shape images [n,3,h,w] (contiguous) resized to the same (oh,ow): samples b0 .. b0+n-1, one launch */
VBG_API int vbg_normalize_resize_pad_batch(const float* imgs, int n, int h, int w, float* batch_nhwc, int b0, int H, int W,
                                   int oh, int ow, const float* h_mean3, const float* h_std3, vbg_stream_t stream);
/* The same transform reading DECODED uint8 pixels [n, h, w, 3] (HWC, RGB) instead of ToTensor's fp32 planes: the kernel applies
 * ToTensor's `byte / 255` (data/SROIE_dataset.py:84-86, one IEEE division) itself, so results are bit-identical to
 * vbg_normalize_resize_pad_batch over ToTensor(img) while the host never touches a pixel and the upload is a quarter of the bytes. */
VBG_API int vbg_normalize_resize_pad_u8(const uint8_t* imgs_hwc, int n, int h, int w, float* batch_nhwc, int b0, int H, int W,
                                int oh, int ow, const float* h_mean3, const float* h_std3, vbg_stream_t stream);
/* A whole batch of differently-sized uint8 documents in ONE launch: tab int32 [B, 6] = {byte offset of the document's pixels
 * relative to `arena` (low word, high word of a signed 64-bit offset), h, w, oh, ow}; max_oh / max_ow size the grid. */
VBG_API int vbg_decode_batch_u8(const uint8_t* arena, const int32_t* tab, int B, int max_oh, int max_ow, float* batch_nhwc, int H,
                        int W, const float* h_mean3, const float* h_std3, vbg_stream_t stream);
/* coords int64 [K,4] (l,t,r,b) -> int32 [K,4]: cols 0,2 *= ratio[b][0] (height ratio), cols 1,3 *= ratio[b][1]
 * (width ratio) in fp32, then truncation (pipeline/transform.py:163-169, axis swap included).    */
VBG_API int vbg_resize_coords(const int64_t* coors, const int32_t* seg_off, const float* ratios /*[B,2]*/, int B, int K,
                      int32_t* out, vbg_stream_t stream);

/* ---- a2: windowed BERT encoder (model/BERTgrid_generator.py:81-146 -> HF BertModel) -------- */
/* Packs the real rows of every 510-token window: [CLS] tokens [SEP].  seq_tab int32 [nseq,4] =
 * (sample, first corpus column, n real tokens, position id of [SEP]); cu int32 [nseq+1] row offsets. */
VBG_API int vbg_bert_assemble(const int64_t* corpus, int L, const int32_t* seq_tab, const int32_t* cu, int nseq, int R,
                      int32_t* ids, int32_t* pos, vbg_stream_t stream);
/* x[r] = LayerNorm(word[ids[r]] + position[pos[r]] + token_type[0]) */
VBG_API int vbg_embed_ln(const int32_t* ids, const int32_t* pos, const float* word, const float* position,
                 const float* type0, const float* gamma, const float* beta, float eps, int R, int hidden,
                 int vocab, int max_pos, float* out, vbg_stream_t stream);
VBG_API int vbg_layernorm(const float* x, const float* gamma, const float* beta, float eps, int R, int hidden, float* out,
                  vbg_stream_t stream);
/* same two kernels writing either storage format (out_plane == 0: fp32) */
VBG_API int vbg_embed_ln_x(const int32_t* ids, const int32_t* pos, const float* word, const float* position,
                   const float* type0, const float* gamma, const float* beta, float eps, int R, int hidden,
                   int vocab, int max_pos, void* out, long long out_plane, vbg_stream_t stream);
VBG_API int vbg_layernorm_x(const float* x, const float* gamma, const float* beta, float eps, int R, int hidden, void* out,
                    long long out_plane, vbg_stream_t stream);
/* softmax(Q K^T / sqrt(d)) V per sequence and head over packed qkv [R, 3*heads*d] (q | k | v). */
VBG_API int vbg_attention_fwd(const float* qkv, const int32_t* cu, int nseq, int max_len, int heads, int head_dim,
                      float* out, int precision, vbg_stream_t stream);

/* Same attention over the bf16 hi/lo planes written by vbg_gemm(..., VBG_OUT_SPLIT_BF16): qkv_hi is bf16 [R, 3*heads*64],
 * the lo plane starts `plane` elements later.  TMA-fed tcgen05 kernel, fp32-class 3-term products; out [R, heads*64] in
 * either storage format (out_plane == 0: fp32). */
VBG_API int vbg_attention_split_fwd(const void* qkv_hi, long long plane, const int32_t* cu, int nseq, int R, int max_len,
                            int heads, int head_dim, void* out, long long out_plane, vbg_stream_t stream);

/* Input-contract check of forward()'s `mask` argument.  The reference selects the real token rows with it
 * (model/BERTgrid_generator.py:152-158) and asserts their count against seg_indices (:233); the packed layout assumes the
 * prefix mask its collate produces (data/SROIE_dataset.py:141-148).  mask i32[B,L]; tok_off i32[B+1] (token offsets from the
 * SHAPES of seg_indices).  ORs bit 2 into *status when mask[b] is not exactly tok_off[b+1]-tok_off[b] leading ones. */
VBG_API int vbg_mask_check(const int32_t* mask, int B, int L, const int32_t* tok_off, int32_t* status, vbg_stream_t stream);

/* ---- a3: token -> segment aggregation (model/BERTgrid_generator.py:148-189) ----------------- */
/* Run starts of consecutive-equal ids inside each sample.  status[0] |= 1 if #runs != K.        */
VBG_API int vbg_segment_starts(const int32_t* seg_ids, const int32_t* tok_off, int B, int n_tok, int K,
                       int32_t* seg_start /*[K+1]*/, int32_t* status, vbg_stream_t stream);
/* out[k] = mean (sequential fp32 sum, one divide) or first of hidden[tok_row[t]], t in run k.   */
VBG_API int vbg_segment_reduce(const float* hidden, const int32_t* tok_row, const int32_t* seg_start, int K, int C,
                       int mode, float* out, vbg_stream_t stream);

/* ---- a4 / a6: box -> index map, BERTgrid scatter, label painting ---------------------------- */
/* idx[b,y,x] = max{ s : cell in [int(t/stride):int(b/stride)) x [int(l/stride):int(r/stride)) } or -1,
 * Python slice semantics (model/BERTgrid_generator.py:230-243; "last writer wins").            */
VBG_API int vbg_box_index_map(const int32_t* boxes, const int32_t* seg_off, int B, int stride, int Hg, int Wg,
                      int32_t* idx, vbg_stream_t stream);
/* grid[b,y,x,:] = idx<0 ? 0 : seg_emb[seg_off[b]+idx]   (NHWC BERTgrid, C % 4 == 0)             */
VBG_API int vbg_grid_scatter(const float* seg_emb, const int32_t* idx, const int32_t* seg_off, int B, int cells, int C,
                     float* grid, vbg_stream_t stream);
/* either storage format for the source rows and the grid (a bf16-plane source needs a bf16-plane grid: plane-wise copy) */
VBG_API int vbg_grid_scatter_x(const void* seg_emb, long long emb_plane, const int32_t* idx, const int32_t* seg_off, int B,
                       int cells, int C, void* grid, long long grid_plane, vbg_stream_t stream);
/* full-resolution labels (model/semantic_segmentation_head.py:199-214): pos_neg = 1 if cls>0 else 2. */
VBG_API int vbg_label_paint(const int32_t* boxes, const int32_t* seg_off, const int32_t* seg_cls, int B, int H, int W,
                    int64_t* pos_neg, int64_t* cls, vbg_stream_t stream);

/* Fused auxiliary-segmentation loss (model/semantic_segmentation_head.py:199-214 labels + :343-347 / :216-233 CE, default
 * mean reduction, no class weights / sampling): out2[0] = mean CE(pred_mask, pos_neg), out2[1] = mean CE(pred_ss, class)
 * over all B*H*W pixels, computed from the LOW-resolution logits [B, H/up, W/up, Ct] (channels [0,c_split) = mask head)
 * without materialising the label maps.  workspace: >= 2 * ceil(W/32) * ceil(H/8) * B floats.          */
VBG_API int vbg_seg_ce_loss(const int32_t* boxes, const int32_t* seg_off, const int32_t* seg_cls, const float* logits, int B,
                    int H, int W, int up, int Ct, int c_split, float* workspace, size_t ws_bytes, float* out2,
                    vbg_stream_t stream);

/* ---- a5 / a6 / a8 / a9: dense contractions ------------------------------------------------- */
/* C[M,N] = epilogue( [A | A2][M,K] * W[N,K]^T ).  A supplies columns [0,K1), A2 (may be NULL when
 * K1 == K) columns [K1,K): the torch.cat-free form of ResNetFPN_ViBERTgrid.py:317-318 and
 * field_type_classification_head.py:185-188.                                                     */
/* W_split (may be NULL unless precision == VBG_PREC_BF16X3): bf16 hi plane of vbg_split_bf16(W), same [N, ldw]
 * geometry as W; the lo plane starts split_plane ELEMENTS after it.                                  */
VBG_API int vbg_gemm(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int ldw, const void* W_split,
             long long split_plane, float* C, int ldc, int M, int N, int K, const vbg_epilogue_t* ep, int precision,
             vbg_stream_t stream);
/* The same contraction with A (and A2) already stored as bf16 hi/lo planes (VBG_PREC_BF16X3 arithmetic, tensor cores only:
 * N >= 64, K % 64 == 0, K1 % 64 == 0): TMA drops the planes straight into the tcgen05 operand tiles, nothing is converted in
 * the kernel.  W_hi / w_plane: planes of vbg_split_bf16(W).  C is float [M, ldc] or, with ep->out_mode ==
 * VBG_OUT_SPLIT_BF16, bf16 planes; ep->res_plane > 0 reads the residual from planes too.                          */
VBG_API int vbg_gemm_ps(const void* A_hi, long long a_plane, int lda, const void* A2_hi, long long a2_plane, int lda2, int K1,
                const void* W_hi, long long w_plane, int ldw, void* C, int ldc, int M, int N, int K,
                const vbg_epilogue_t* ep, void* workspace, size_t ws_bytes, vbg_stream_t stream);
/* implicit-GEMM convolution over a split NHWC activation (Cin % 64 == 0, Cout >= 64, stride 1 or 2) */
VBG_API int vbg_conv2d_ps(const void* x_hi, long long x_plane, int B, int H, int W, int Cin, const void* w_hi, long long w_plane,
                  int Cout, int kh, int kw, int stride, int pad, void* y, const vbg_epilogue_t* ep, void* workspace,
                  size_t ws_bytes, vbg_stream_t stream);
/* Split-K: a shape with few output tiles and a long K (the 16x16 / 32x32 ResNet stages, the ROI FC) is computed as
 * (tile, K-range) work units whose fp32 partial tiles go to `workspace` and are summed in a fixed order by a finishing
 * kernel that applies the epilogue (deterministic).  These return the workspace bytes such a call can use (0 = the shape
 * never splits); passing NULL / fewer bytes just disables the split.  Host-only arithmetic, no device work.  Opt-in via
 * VBG_TUNE_SPLITK in `tune` (measured slower than the CTA-pair tiles at the BASELINE shapes: without it these return 0). */
VBG_API long long vbg_gemm_ps_workspace(int M, int N, int K, int tune);
VBG_API long long vbg_conv2d_ps_workspace(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int tune);
/* Tuning aid: when dev_buf (>= 16 int64, device memory) is non-NULL, CTA 0 of every following CTA-pair GEMM launch writes
 * clock64 stamps of its pipeline milestones into it (entry, setup done, first TMA issued, first operands landed, last MMA
 * committed, epilogue start / end, exit).  NULL (the default) disables it.  Not thread safe; never used on the hot path. */
VBG_API int vbg_debug_set_timeline(long long* dev_buf);
/* out[i] = float(hi[i]) + float(lo[i])  (inspection / tests: split activation -> fp32) */
VBG_API int vbg_merge_bf16(const void* hi, const void* lo, long long n, float* out, vbg_stream_t stream);
/* NHWC convolution as implicit GEMM: y[B,Ho,Wo,Cout] = epilogue(conv(x[B,H,W,Cin], w[Cout,kh,kw,Cin])) */
VBG_API int vbg_conv2d(const float* x, int B, int H, int W, int Cin, const float* w, const void* w_split, long long split_plane,
               int Cout, int kh, int kw, int stride, int pad, float* y, const vbg_epilogue_t* ep, int precision,
               vbg_stream_t stream);
/* bf16 hi/lo split of n fp32 values: hi = bf16_rn(w), lo = bf16_rn(w - hi)  (one-time weight preparation) */
VBG_API int vbg_split_bf16(const float* w, long long n, void* hi, void* lo, vbg_stream_t stream);
/* ResNet stem (7x7, stride 2, pad 3, 3->Cout; model/ResNetFPN_ViBERTgrid.py:351-361 / torchvision conv1) over the
 * padded NHWC4 batch of vbg_normalize_resize_pad.  w_ohwi4 [Cout,7,7,4] feeds the CUDA-core path; w_split
 * (bf16 planes of the [Cout,256] operand from vbg_stem_pack_weights) the tensor-core path.            */
VBG_API int vbg_stem_conv(const float* x4, int B, int H, int W, const float* w_ohwi4, const void* w_split, long long split_plane,
                  int Cout, float* y, const vbg_epilogue_t* ep, int precision, vbg_stream_t stream);
VBG_API int vbg_stem_pack_weights(const float* w_oihw, int Cout, float* w_ohwi4, float* w_k256, vbg_stream_t stream);
VBG_API int vbg_maxpool3x3s2(const float* x, int B, int H, int W, int C, float* y, vbg_stream_t stream);
VBG_API int vbg_avgpool2x2(const float* x, int B, int H, int W, int C, float* y, vbg_stream_t stream);
VBG_API int vbg_maxpool3x3s2_x(const void* x, long long x_plane, int B, int H, int W, int C, void* y, long long y_plane,
                       vbg_stream_t stream);
VBG_API int vbg_avgpool2x2_x(const void* x, long long x_plane, int B, int H, int W, int C, void* y, long long y_plane,
                     vbg_stream_t stream);
/* eval-mode BatchNorm folded to y = x*scale + shift */
VBG_API int vbg_bn_fold(const float* weight, const float* bias, const float* mean, const float* var, float eps, int C,
                float* scale, float* shift, vbg_stream_t stream);
/* PyTorch conv weight [O,I,H,W] -> [O,H,W,I]; also permutes the ROI FC weight [1024,(C,7,7)] -> [1024,(7,7,C)] */
VBG_API int vbg_repack_oihw_to_ohwi(const float* w, int O, int I, int H, int W, float* out, vbg_stream_t stream);

/* ---- first bricks of the training step (linear-layer backward) --------------------------------
 * out[c][r] = x[r][c] written as bf16 hi/lo planes with leading dimension ld_out >= rows (the tail is zero filled): the
 * K-major operand of a GEMM that reduces over `rows`.  x is fp32 (x_plane == 0) or bf16 planes.  With these,
 *   dgrad  dX[M,K] = vbg_gemm_ps(A = dY planes [M,N],          W = transpose_split(W [N,K])       -> [K,N] planes)
 *   wgrad  dW[N,K] = vbg_gemm_ps(A = transpose_split(dY) [N,Mp], W = transpose_split(X) [K,Mp]), Mp = M rounded up to 64
 *   bgrad  db[N]   = vbg_colsum(dY)
 * replace torch.autograd's addmm backward of nn.Linear (HF BertSelfOutput / BertIntermediate / ..., head MLPs).      */
VBG_API int vbg_transpose_split(const void* x, long long x_plane, int rows, int cols, void* out_hi, long long out_plane, int ld_out,
                        vbg_stream_t stream);
VBG_API long long vbg_colsum_workspace(long long rows, int cols);
VBG_API int vbg_colsum(const void* x, long long x_plane, long long rows, int cols, float* out, float* workspace, size_t ws_bytes,
               vbg_stream_t stream);
/* dW[N,K] = dY[M,N]^T X[M,K] without transposes: both plane operands are read as MN-major tcgen05 operands (reduction index =
 * the slow, row index), split over the row range with a deterministic finish.  N % 128 == 0, K % 64 == 0; workspace bytes from
 * vbg_linear_wgrad_workspace (0: none needed).                                                                      */
/* the same for an NHWC convolution: dW[Cout,kh,kw,Cin] from dY planes [B,Ho,Wo,Cout] and X planes [B,H,W,Cin] (the X tile of a
 * filter tap is the tap-shifted window through a rank-5 TMA map: padding = out-of-bounds zero fill).  Cout % 128 == 0,
 * Cin % 64 == 0, stride 1 or 2, output rows tiled in 64-pixel blocks (Wo >= 64, or Wo * k == 64).                       */
VBG_API long long vbg_conv2d_wgrad_workspace(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad);
VBG_API int vbg_conv2d_wgrad(const void* dY_hi, long long y_plane, const void* X_hi, long long x_plane, int B, int H, int W, int Cin,
                     int Cout, int kh, int kw, int stride, int pad, float* dW, void* workspace, size_t ws_bytes,
                     vbg_stream_t stream);
VBG_API long long vbg_linear_wgrad_workspace(int M, int N, int K);
VBG_API int vbg_linear_wgrad(const void* dY_hi, long long y_plane, const void* X_hi, long long x_plane, int M, int N, int K, float* dW,
                     void* workspace, size_t ws_bytes, vbg_stream_t stream);
/* planes of w'[Cin,kh,kw,Cout] = w[Cout,kh,kw,Cin] flipped in (kh,kw): with it the data gradient of a stride-1 convolution is
 * vbg_conv2d_ps(dY planes, w' planes, stride 1, pad k-1-p)  (replaces torch's conv2d backward-input of the ResNet / FPN convs) */
VBG_API int vbg_conv_dgrad_weight(const float* w_ohwi, int Cout, int kh, int kw, int Cin, void* out_hi, long long out_plane,
                          vbg_stream_t stream);
/* LayerNorm backward (HF BertSelfOutput / BertOutput / embeddings LayerNorm): dx [R,hidden]; dgamma / dbeta [hidden] (may be
 * NULL together) through `workspace` >= ceil(R/64) * 2 * hidden + 2 * R floats, summed in a fixed order.               */
VBG_API int vbg_layernorm_bwd(const float* x, const float* dy, const float* gamma, float eps, int R, int hidden, float* dx,
                      float* dgamma, float* dbeta, float* workspace, size_t ws_bytes, vbg_stream_t stream);

/* ---- training step: the backward of loss.backward() (pipeline/train_val_utils.py:277) through the modules of the path, and the
 *      training-mode forward pieces (batch-statistics BatchNorm, dropout).  fp32 channels-last; reductions have a fixed order;
 *      every reduction has a fixed order: no atomics anywhere in the training step (bitwise reproducible gradients).          */
/* nn.BatchNorm2d in train mode (model/ResNetFPN_ViBERTgrid.py:106-186, semantic_segmentation_head.py:66, field_type_..._head.py:64)
 * over [rows, C]: batch mean / biased variance / 1/sqrt(var+eps); C = 4 * a divisor of 256; workspace from vbg_bn_workspace     */
VBG_API long long vbg_bn_workspace(long long rows, int C);
VBG_API int vbg_bn_stats(const float* x, long long rows, int C, float eps, float* mean, float* var, float* rstd, float* workspace,
                 size_t ws_bytes, vbg_stream_t stream);
/* y = (x - mean) * rstd * gamma + beta (+ residual) (ReLU when relu != 0) */
VBG_API int vbg_bn_apply(const float* x, long long rows, int C, const float* mean, const float* rstd, const float* gamma, const float* beta,
                 const float* residual, int relu, float* y, vbg_stream_t stream);
/* backward of vbg_bn_apply(+stats): y_relu = the forward output when ReLU was applied (mask), else NULL; dres (may be NULL) receives
 * the masked dy = the gradient of the residual input                                                                       */
VBG_API int vbg_bn_bwd(const float* x, const float* dy, const float* y_relu, long long rows, int C, const float* mean, const float* rstd,
               const float* gamma, float* dx, float* dres, float* dgamma, float* dbeta, float* workspace, size_t ws_bytes,
               vbg_stream_t stream);
/* vbg_bn_bwd in two halves, for nn.SyncBatchNorm (the reference converts the model when syncBN is set: train_SROIE.py:203-205):
 * _reduce leaves this rank's sum_dy_xhat = sum dy' * xhat (= dgamma) and sum_dy = sum dy' (= dbeta); the caller all-reduces both
 * over the ranks; _dx then forms dx with inv_count = 1 / (total rows over all ranks) and the global mean / rstd               */
VBG_API int vbg_bn_bwd_reduce(const float* x, const float* dy, const float* y_relu, long long rows, int C, const float* mean,
                      const float* rstd, float* sum_dy_xhat, float* sum_dy, float* workspace, size_t ws_bytes, vbg_stream_t stream);
VBG_API int vbg_bn_bwd_dx(const float* x, const float* dy, const float* y_relu, long long rows, int C, float inv_count, const float* mean,
                  const float* rstd, const float* gamma, const float* sum_dy_xhat, const float* sum_dy, float* dx, float* dres,
                  vbg_stream_t stream);
/* dX of max_pool2d(3, 2, 1) (first-maximum rule), x [B,H,W,C], dy [B,Ho,Wo,C] */
VBG_API int vbg_maxpool3x3s2_bwd(const float* x, const float* dy, int B, int H, int W, int C, float* dx, vbg_stream_t stream);
/* the same with the pooled output y [B, Ho, Wo, C] at hand (a position can take a window's gradient only where x == y): ~3x fewer loads */
VBG_API int vbg_maxpool3x3s2_bwd_y(const float* x, const float* y, const float* dy, int B, int H, int W, int C, float* dx,
                           vbg_stream_t stream);
/* y [B,H/2,W/2,C] = scale * 2x2 block sums: backward of the nearest-x2 upsample (scale 1) */
VBG_API int vbg_sumpool2x2(const float* x, int B, int H, int W, int C, float scale, float* y, vbg_stream_t stream);
/* y [B,H,W,C] from x [B,Hi,Wi,C]: nearest x2 times scale (backward of avg_pool2d(2): scale 0.25), or zero insertion (zero_insert != 0:
 * dY of a stride-2 convolution spread onto the stride-1 lattice)                                                           */
VBG_API int vbg_expand2x(const float* x, int B, int Hi, int Wi, int C, int H, int W, float scale, int zero_insert, float* y,
                 vbg_stream_t stream);
/* erf-GELU: out = gelu(x) when dy == NULL, else out = dy * gelu'(x) */
VBG_API int vbg_gelu(const float* x, const float* dy, long long n, float* out, vbg_stream_t stream);
/* the same with x, dy and out each in either storage format (plane == 0: fp32; > 0: bf16 hi/lo planes `plane` elements apart) */
VBG_API int vbg_gelu_x(const void* x, long long x_plane, const void* dy, long long dy_plane, long long n, void* out, long long out_plane,
               vbg_stream_t stream);
/* inverted dropout with a counter-based mask: y[i] = x[i] * keep(seed, i) / (1 - p); the same call is its own backward */
VBG_API int vbg_dropout(const float* x, long long n, float p, unsigned long long seed, float* y, vbg_stream_t stream);
/* 31-bit sampling keys key[i] = hash(seed, step_seed, i) for the device-side sampled / OHEM losses (pipeline/custom_loss.py:9-382
 * restated without host randomness: an element is kept iff its key is among the k smallest of its group) */
VBG_API int vbg_uniform_keys(long long n, unsigned long long seed, const unsigned long long* step_seed /* device, or NULL */,
                     int32_t* out, vbg_stream_t stream);
/* the same with an optional DEVICE word `step_seed` folded into the seed at run time: a captured CUDA graph bakes the by-value
 * seed of each call site; the host refreshes the one device word before every replay, so each step draws new masks */
VBG_API int vbg_dropout_ds(const float* x, long long n, float p, unsigned long long seed, const unsigned long long* step_seed,
                   float* y, vbg_stream_t stream);
/* backward of vbg_grid_scatter: demb[k] = sum of dgrid over the cells segment k won (row stride ld floats between cells) */
VBG_API int vbg_grid_scatter_bwd(const float* dgrid, long long ld, const int32_t* idx, const int32_t* boxes, const int32_t* seg_off, int B,
                         int K, int stride, int Hg, int Wg, int C, float* demb, vbg_stream_t stream);
/* backward of vbg_segment_reduce into the rows of dhidden that belong to a segment (caller zero-fills dhidden) */
VBG_API int vbg_segment_reduce_bwd(const float* dseg, const int32_t* tok_row, const int32_t* seg_start, int K, int C, int mode,
                           float* dhidden, vbg_stream_t stream);
/* embedding tables: dword[ids[r]] += dx[r], dpos[pos[r]] += dx[r] (caller zero-fills the tables); deterministic: the first packed
 * row of each table row sums its matches in ascending row order */
VBG_API int vbg_embed_bwd(const float* dx, const int32_t* ids, const int32_t* pos, int R, int hidden, float* dword, float* dpos,
                  vbg_stream_t stream);
/* backward of vbg_roi_align_fwd into dfeat [B,Hf,Wf,C]: deterministic gather form (every pixel written exactly once: no atomics,
 * no zero-fill); workspace >= vbg_roi_align_bwd_workspace(K) bytes (per-ROI geometry) */
VBG_API long long vbg_roi_align_bwd_workspace(int K);
VBG_API int vbg_roi_align_bwd(const float* dout, int B, int Hf, int Wf, int C, const int32_t* boxes, const int32_t* seg_off, int K,
                      float spatial_scale, int P, float* dfeat, void* workspace, size_t ws_bytes, vbg_stream_t stream);
/* gradient of gscale[0] * mean CE(mask head) + gscale[1] * mean CE(class head) (semantic_segmentation_head.py:343-347) w.r.t. the
 * LOW-resolution logits [B,H/up,W/up,Ct], labels = the int64 maps of vbg_label_paint                                       */
VBG_API int vbg_seg_ce_bwd(const float* logits, const long long* pos_neg, const long long* cls, int B, int H, int W, int up, int Ct,
                   int c_split, const float* gscale, float* dlogits, vbg_stream_t stream);
/* backward of vbg_upsample_split_nchw */
VBG_API int vbg_upsample_split_bwd(const float* d1, const float* d2, int B, int h, int w, int Ct, int up, int c_split, float* dlogits,
                           vbg_stream_t stream);
/* dW[N<=16, K] = dY^T X for the narrow heads (row strides ldy / ldx) */
VBG_API long long vbg_small_wgrad_workspace(long long M, int N, int K);
VBG_API int vbg_small_wgrad(const float* dy, int ldy, const float* x, int ldx, long long M, int N, int K, float* dw, float* workspace,
                    size_t ws_bytes, vbg_stream_t stream);
/* weight gradient of the 7x7/2 stem over the zero-bordered NHWC4 batch [B,Hp,Wp,4]: dw774 [64,7,7,4] */
VBG_API long long vbg_stem_wgrad_workspace(void);
VBG_API int vbg_stem_wgrad(const float* x4, const float* dy, int B, int Hp, int Wp, int Ho, int Wo, float* dw774, float* workspace,
                   size_t ws_bytes, vbg_stream_t stream);
/* self-attention backward over the packed varlen batch (head dimension 64): dqkv [rows, 3*heads*64] from qkv, the forward output
 * and its gradient; workspace >= rows * heads * 2 floats (row log-sum-exp and delta)                                      */
VBG_API int vbg_attention_bwd(const float* qkv, const float* out, const float* d_out, const int32_t* cu, int nseq, int max_len, int heads,
                      int head_dim, long long rows, float* dqkv, float* workspace, size_t ws_bytes, vbg_stream_t stream);

/* ---- training-mode attention on the tensor cores (HF BertSelfAttention under model.train(), model/BERTgrid_generator.py:134)
 * forward: vbg_attention_split_fwd that ALSO (1) applies dropout to the attention probabilities (attention_probs_dropout_prob)
 *   with a counter-based mask -- a pure function of (seed, packed query row, key index, head) -- and (2) stores the base-2 row
 *   log-sum-exp lse2 [R, heads] the backward rebuilds the probabilities from.
 * backward: dQKV fp32 [R, 3*hidden] from the bf16 hi/lo planes of QKV and of dO (plus fp32 O, dO for delta = rowsum(dO o O)),
 *   tcgen05 bf16x3 products, the same dropout mask regenerated; workspace >= R * heads floats (delta).  Deterministic.
 * vbg_attention_dropout_mask: the keep mask [len, len] of one (sequence starting at packed row row0, head) and the scale
 *   1 / (1 - p_effective) written to the HOST float *inv_keep -- test infrastructure for an exact reference. */
VBG_API int vbg_attention_split_train_fwd(const void* qkv_hi, long long plane, const int32_t* cu, int nseq, int R, int max_len,
                                  int heads, int head_dim, void* out, long long out_plane, float* lse2, float p_drop,
                                  unsigned long long seed, const unsigned long long* step_seed /* device, or NULL */,
                                  vbg_stream_t stream);
VBG_API int vbg_attention_bwd_tc(const void* qkv_hi, long long qkv_plane, const void* do_hi, long long do_plane, const float* out,
                         const float* d_out, const float* lse2, const int32_t* cu, int nseq, int R, int max_len, int heads,
                         int head_dim, float p_drop, unsigned long long seed, const unsigned long long* step_seed /* device, or NULL */,
                         float* dqkv, float* workspace, size_t ws_bytes, vbg_stream_t stream);
VBG_API int vbg_attention_dropout_mask(unsigned long long seed, unsigned long long step_seed /* host value */, int has_step_seed,
                               float p_drop, int row0, int len, int head, float* mask, float* inv_keep, vbg_stream_t stream);

/* ---- multi-tensor optimizer steps (the reference's torch.optim.SGD / torch.optim.AdamW, train_SROIE.py:217-235, stepped at
 * pipeline/train_val_utils.py:272-284): one launch updates every tensor of a param group with torch's single-tensor update
 * rules.  `table`: DEVICE array of 6 int64 per tensor {param ptr, grad ptr, state1 ptr, state2 ptr, numel, first_chunk};
 * a CTA owns one chunk of vbg_optim_chunk() elements; first_chunk = running sum of ceil(numel / chunk).  SGD: state1 =
 * momentum_buffer (unused when momentum == 0), first_step != 0 initialises it to the gradient; AdamW: state1 / state2 =
 * exp_avg / exp_avg_sq, bias_correction1 = 1 - beta1^t, sqrt_bias_correction2 = sqrt(1 - beta2^t).  grad_scale multiplies
 * the gradients on the fly (1 / world size of an un-averaged all-reduce, or 1). */
VBG_API int vbg_optim_chunk(void);
VBG_API int vbg_sgd_step_mt(const void* table, int n_tensors, long long total_chunks, float lr, float momentum, float weight_decay,
                    int first_step, float grad_scale, vbg_stream_t stream);
VBG_API int vbg_adamw_step_mt(const void* table, int n_tensors, long long total_chunks, float lr, float beta1, float beta2, float eps,
                      float weight_decay, float bias_correction1, float sqrt_bias_correction2, float grad_scale, vbg_stream_t stream);

/* ---- a7: GridROIAlign (model/grid_roi_align.py:37-41,81 -> torchvision roi_align, aligned=False,
 *          sampling_ratio=-1) over NHWC features; boxes are the int32 transformed coords.         */
VBG_API int vbg_roi_align_fwd(const float* feat, int B, int Hf, int Wf, int C, const int32_t* boxes, const int32_t* seg_off,
                      int K, float spatial_scale, int P, float* out /*[K,P,P,C]*/,
                      int32_t* sample_grid /*[K,2] (gh,gw) or NULL*/, vbg_stream_t stream);

VBG_API int vbg_roi_align_x(const void* feat, long long feat_plane, int B, int Hf, int Wf, int C, const int32_t* boxes,
                    const int32_t* seg_off, int K, float spatial_scale, int P, void* out, long long out_plane,
                    int32_t* sample_grid, vbg_stream_t stream);

/* The same operation with the kernel chosen by the CALLER (an argument, not process state): AUTO picks by shape -- the
 * persistent TMA row-streaming kernel for P == 7, C in {128, 256}; tests and scripts pass ROW / DIRECT to compare the kernels
 * on identical inputs.  All variants produce the same bit-exact sample grid; values agree to fp32 re-association. */
enum { VBG_ROI_AUTO = 0, VBG_ROI_STREAM = 1, VBG_ROI_ROW = 2, VBG_ROI_DIRECT = 3 };
VBG_API int vbg_roi_align_sel(const void* feat, long long feat_plane, int B, int Hf, int Wf, int C, const int32_t* boxes,
                      const int32_t* seg_off, int K, float spatial_scale, int P, void* out, long long out_plane,
                      int32_t* sample_grid, int variant, vbg_stream_t stream);

/* ---- heads / outputs ---------------------------------------------------------------------- */
VBG_API int vbg_softmax_rows(const float* x, int R, int C, float* y, vbg_stream_t stream);
/* sigmoid cascade of the "full" head (field_type_classification_head.py:312-332) */
VBG_API int vbg_full_head_scores(const float* pos_neg, const float* cls /*[R,C-1]*/, int R, int C, float* out /*[R,C]*/,
                         vbg_stream_t stream);
/* NHWC [B,h,w,Ct] --nearest x up--> NCHW out1 [B,c_split,h*up,w*up], out2 [B,Ct-c_split,h*up,w*up]
 * (semantic_segmentation_head.py:73-78 with the 1x1 convs commuted before the upsample)          */
VBG_API int vbg_upsample_split_nchw(const float* x, int B, int h, int w, int Ct, int up, int c_split, float* out1,
                            float* out2, vbg_stream_t stream);
VBG_API int vbg_nhwc_to_nchw(const float* x, int B, int H, int W, int C, float* y, vbg_stream_t stream);
/* Viterbi decode per sample (model/crf.py:96-146); tags written as float like the reference returns them */
VBG_API int vbg_crf_viterbi(const float* feats /*[K,T]*/, const float* trans /*[T,T] to<-from*/, const int32_t* seg_off, int B,
                    int K, int T, float* tags /*[K]*/, float* scores /*[B]*/, void* workspace /*>= K*T bytes*/,
                    size_t ws_bytes, vbg_stream_t stream);
/* CRF negative log-likelihood per sample, nll[b] = (logZ_b - gold_b) / S_b: the training branch of the `crf` head
 * (model/crf.py:47-93 `_forward_alg` / `_score_sentence`, :148-152 `forward`; looped over samples by
 * model/field_type_classification_head.py:686-699).  START = T-2, STOP = T-1.  alpha [K,T] receives the forward variables
 * of every step with their maximum subtracted (all recursions run normalised, the normalisers are summed in double into
 * logz, so the result is fp32-exact at any sequence length); the caller keeps it for the backward.
 * tags = gold labels int32 [K] in [0, T-2).                                                                        */
VBG_API int vbg_crf_nll_fwd(const float* feats /*[K,T]*/, const float* trans /*[T,T] to<-from*/, const int32_t* tags /*[K]*/,
                    const int32_t* seg_off, int B, int K, int T, float* alpha /*[K,T]*/, float* logz /*[B]*/,
                    float* nll /*[B]*/, vbg_stream_t stream);
/* gradient of sum_b dnll[b] * nll[b]: posterior marginals minus gold counts (what autograd derives in the reference).
 * dtrans_part [B,T,T] holds each sample's share; the caller sums over B (fixed order => deterministic).           */
VBG_API int vbg_crf_nll_bwd(const float* feats, const float* trans, const int32_t* tags, const int32_t* seg_off, int B, int K,
                    int T, const float* alpha, const float* dnll /*[B]*/, float* dfeats /*[K,T]*/,
                    float* dtrans_part /*[B,T,T]*/, vbg_stream_t stream);

/* ---- f3: input pipeline -- shards of pre-tokenised, pre-decoded documents (HOST functions, no device work) ------------
 * Replaces the per-item work of data/SROIE_dataset.py:94-162 (__getitem__: PIL decode, pandas CSV rows, Python tokenisation)
 * and :165-208 (_ViBERTgrid_coll_func: pad_sequence + mask) on the way INTO the hot path.  A shard (csrc/vbg_shard.cpp documents
 * the file layout; shards.py writes it once, offline) is memory-mapped; a batch is collated into ONE caller-owned (pinned)
 * staging buffer that crosses to the device as one copy.  Handles are not thread safe against close; collate may be called
 * concurrently on one handle (read-only mapping). */
VBG_API int vbg_shard_open(const char* path, void** handle);
VBG_API int vbg_shard_close(void* handle);
VBG_API int vbg_shard_num_docs(void* handle);                                     /* < 0: bad handle */
VBG_API int vbg_shard_doc_shape(void* handle, int doc, int32_t* out4 /* h, w, n_tok, n_seg */);
/* per-document side data the eval loop reads (ocr text list + key dict, data/SROIE_dataset.py:150-162) as a JSON blob */
VBG_API int vbg_shard_doc_meta(void* handle, int doc, const char** ptr, int64_t* bytes);
/* layout13: [0] total staging bytes, [1] L = padded corpus width, [2] sum n_tok, [3] sum n_seg, then byte offsets of
 * [4] corpus int64 [B,L], [5] mask int32 [B,L], [6] seg_ids int32 [sum n_tok], [7] classes int32 [sum n_seg],
 * [8] coors int64 [sum n_seg,4], [9] shapes int32 [B,4] = (h,w,n_tok,n_seg), [10] image offsets int64 [B] (relative to the
 * arena), [11] the uint8 image arena, and [12] the arena's bytes. */
VBG_API int vbg_shard_batch_layout(void* handle, const int32_t* docs, int B, int64_t* layout13);
VBG_API int vbg_shard_collate(void* handle, const int32_t* docs, int B, void* staging, size_t staging_bytes, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* VBG_H_ */
