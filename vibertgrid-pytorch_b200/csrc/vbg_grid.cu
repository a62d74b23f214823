// BERTgrid construction: coordinate rescale, token->segment aggregation, box index map,
// BERTgrid scatter, label painting.  All HBM/latency-bound integer + copy work:
// coalesced 128-bit accesses, no atomics, deterministic ("last writer wins" == max covering id).
//
// Replaces the Python loops of reference model/BERTgrid_generator.py:148-189 (aggregation, one
// .item() sync per token), :220-243 (scatter, 4 int() syncs per segment) and
// model/semantic_segmentation_head.py:199-214 (label painting), plus pipeline/transform.py:163-169.
#include "vbg_common.cuh"

namespace vbg {

// ------------------------------------------------------------------ coords
__global__ void resize_coords_kernel(const int64_t* __restrict__ coors, const int32_t* __restrict__ seg_off,
                                     const float* __restrict__ ratios, int B, int K, int32_t* __restrict__ out) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int b = sample_of(seg_off, B, k);
  float rh = ratios[2 * b], rw = ratios[2 * b + 1];
  // x columns (0,2) use the HEIGHT ratio, y columns (1,3) the WIDTH ratio: reference quirk kept.
  float v0 = __fmul_rn((float)coors[4 * k + 0], rh);
  float v1 = __fmul_rn((float)coors[4 * k + 1], rw);
  float v2 = __fmul_rn((float)coors[4 * k + 2], rh);
  float v3 = __fmul_rn((float)coors[4 * k + 3], rw);
  int4 o = make_int4((int)v0, (int)v1, (int)v2, (int)v3);  // float->int conversion truncates toward zero
  reinterpret_cast<int4*>(out)[k] = o;
}

// ------------------------------------------------------------------ segment runs
// One CTA scans all tokens in chunks of blockDim.x; flag = first token of a sample or id change.
__global__ void segment_starts_kernel(const int32_t* __restrict__ seg_ids, const int32_t* __restrict__ tok_off, int B,
                                      int n_tok, int K, int32_t* __restrict__ seg_start, int32_t* __restrict__ status) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n_tok; t0 += blockDim.x) {
    int t = t0 + tid;
    int flag = 0;
    if (t < n_tok) {
      int b = sample_of(tok_off, B, t);
      flag = (t == tok_off[b]) || (seg_ids[t] != seg_ids[t - 1]);
    }
    unsigned bal = __ballot_sync(0xffffffffu, flag);
    int excl = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_tot[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < nw; ++w) {
      int c = warp_tot[w];
      if (w < wid) before += c;
      total += c;
    }
    int base = base_s;
    if (flag) {
      int s = base + before + excl;
      if (s <= K) seg_start[s] = t;
    }
    __syncthreads();
    if (tid == 0) base_s = base + total;
    __syncthreads();
  }
  if (tid == 0) {
    if (base_s == K) seg_start[K] = n_tok;
    else if (status) atomicOr(status, 1);   // reference asserts #runs == #boxes (BERTgrid_generator.py:233)
  }
}

// One CTA per segment; each thread owns 4 channels and adds the run's tokens in order
// (bit-compatible with the reference's sequential `mean += e; mean /= n`).
__global__ void segment_reduce_kernel(const float* __restrict__ hidden, const int32_t* __restrict__ tok_row,
                                      const int32_t* __restrict__ seg_start, int C4, int mode,
                                      float* __restrict__ out) {
  const int k = blockIdx.x;
  const int a = seg_start[k], e = seg_start[k + 1];
  const float4* h4 = reinterpret_cast<const float4*>(hidden);
  float4* o4 = reinterpret_cast<float4*>(out);
  for (int c = threadIdx.x; c < C4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e > a) {
      acc = __ldg(h4 + (size_t)tok_row[a] * C4 + c);
      if (mode == VBG_AGG_MEAN) {
        for (int t = a + 1; t < e; ++t) {
          float4 v = __ldg(h4 + (size_t)tok_row[t] * C4 + c);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float n = (float)(e - a);
        acc.x = __fdiv_rn(acc.x, n); acc.y = __fdiv_rn(acc.y, n);
        acc.z = __fdiv_rn(acc.z, n); acc.w = __fdiv_rn(acc.w, n);
      }
    }
    o4[(size_t)k * C4 + c] = acc;
  }
}

// ------------------------------------------------------------------ index map / label paint
// CTA = 32x8 tile of cells of one sample.  Boxes are consumed in chunks of 256 from the LAST
// segment backwards; each chunk is culled against the tile and compacted (order preserved) into
// shared memory; a cell takes the first hit scanning backwards.  Early exit when the tile is full.
constexpr int kTileW = 32, kTileH = 8;

// kMode 0: index map; 1: paint the two label maps; 2: fused auxiliary-segmentation cross entropy -- the labels are
// consumed in registers (never written) against the low-resolution logits, per-CTA partial sums go to `partial`.
struct SegCeArgs { const float* logits; int h, w, Ct, up, c_split; float* partial; };

template <int kMode>
__global__ void __launch_bounds__(256)
box_map_kernel(const int32_t* __restrict__ boxes, const int32_t* __restrict__ seg_off,
               const int32_t* __restrict__ seg_cls, int stride, int Hg, int Wg, int32_t* __restrict__ idx,
               int64_t* __restrict__ pos_neg, int64_t* __restrict__ cls, const SegCeArgs ce) {
  constexpr bool kPaint = kMode == 1;
  __shared__ int4 hit_box[256];
  __shared__ int hit_id[256];
  __shared__ int warp_cnt[8];
  const int b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const int x = x0 + lane, y = y0 + wid;
  const int s0 = seg_off[b], S = seg_off[b + 1] - s0;
  int best = -1;
  const bool live = (x < Wg) && (y < Hg);
  for (int hi = S; hi > 0; hi -= 256) {
    const int lo = hi > 256 ? hi - 256 : 0;
    const int s = lo + tid;
    int4 bx = make_int4(0, 0, 0, 0);
    bool ov = false;
    if (s < hi) {
      int4 c = __ldg(reinterpret_cast<const int4*>(boxes) + s0 + s);   // (l, t, r, b)
      bx.x = py_slice_bound(c.x / stride, Wg);
      bx.y = py_slice_bound(c.y / stride, Hg);
      bx.z = py_slice_bound(c.z / stride, Wg);
      bx.w = py_slice_bound(c.w / stride, Hg);
      ov = bx.x < bx.z && bx.y < bx.w && bx.x < x0 + kTileW && bx.z > x0 && bx.y < y0 + kTileH && bx.w > y0;
    }
    unsigned bal = __ballot_sync(0xffffffffu, ov);
    if (lane == 0) warp_cnt[wid] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      int c = warp_cnt[w];
      if (w < wid) before += c;
      total += c;
    }
    if (ov) {
      int p = before + __popc(bal & ((1u << lane) - 1));
      hit_box[p] = bx;
      hit_id[p] = s;
    }
    __syncthreads();
    if (live && best < 0) {
      for (int p = total - 1; p >= 0; --p) {
        int4 q = hit_box[p];
        if (x >= q.x && x < q.z && y >= q.y && y < q.w) { best = hit_id[p]; break; }
      }
    }
    if (__syncthreads_and(best >= 0 || !live)) break;
  }
  if (kMode == 2) {
    // CE of this pixel for the 3-way mask head and the C-way class head (semantic_segmentation_head.py:343-347 with the
    // default mean reduction): logits of low-res pixel (y/up, x/up); nearest upsampling commutes with the 1x1 heads.
    float l1 = 0.f, l2 = 0.f;
    if (live) {
      int c = 0, pn = 0;
      if (best >= 0) { c = seg_cls[s0 + best]; pn = c > 0 ? 1 : 2; }
      const float* z = ce.logits + (((size_t)b * ce.h + y / ce.up) * ce.w + x / ce.up) * ce.Ct;
      float m1 = -INFINITY, m2 = -INFINITY;
      for (int i = 0; i < ce.c_split; ++i) m1 = fmaxf(m1, __ldg(z + i));
      for (int i = ce.c_split; i < ce.Ct; ++i) m2 = fmaxf(m2, __ldg(z + i));
      float e1 = 0.f, e2 = 0.f;
      for (int i = 0; i < ce.c_split; ++i) e1 += expf(__ldg(z + i) - m1);
      for (int i = ce.c_split; i < ce.Ct; ++i) e2 += expf(__ldg(z + i) - m2);
      l1 = (m1 + logf(e1)) - __ldg(z + pn);
      l2 = (m2 + logf(e2)) - __ldg(z + ce.c_split + c);
    }
    l1 = warp_sum(l1); l2 = warp_sum(l2);
    __shared__ float red[2][8];
    if (lane == 0) { red[0][wid] = l1; red[1][wid] = l2; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, c2 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) { a += red[0][w]; c2 += red[1][w]; }
      const size_t blk = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      ce.partial[2 * blk] = a; ce.partial[2 * blk + 1] = c2;
    }
    return;
  }
  if (!live) return;
  const size_t o = ((size_t)b * Hg + y) * Wg + x;
  if (kPaint) {
    long long c = 0, pn = 0;
    if (best >= 0) {
      c = seg_cls[s0 + best];
      pn = c > 0 ? 1 : 2;
    }
    pos_neg[o] = pn;
    cls[o] = c;
  } else {
    idx[o] = best;
  }
}

// deterministic final reduction of the per-CTA partial sums (fixed order, double accumulator)
__global__ void seg_ce_finish_kernel(const float* __restrict__ partial, int nblk, double inv_n, float* __restrict__ out2) {
  __shared__ double sh[2][256];
  double a = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 256) { a += (double)partial[2 * i]; c += (double)partial[2 * i + 1]; }
  sh[0][threadIdx.x] = a; sh[1][threadIdx.x] = c;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out2[0] = (float)(sh[0][0] * inv_n); out2[1] = (float)(sh[1][0] * inv_n); }
}

// ------------------------------------------------------------------ scatter
// One warp per grid cell: 128-bit coalesced copy of the winning segment's embedding (L2 resident,
// K*C*4 bytes) or zeros.  Pure write stream: B*Hg*Wg*C*4 bytes.
__global__ void __launch_bounds__(256)
grid_scatter_kernel(const float* __restrict__ seg_emb, const int32_t* __restrict__ idx,
                    const int32_t* __restrict__ seg_off, int cells_per_img, long long total_cells, int C4,
                    void* __restrict__ grid, long long grid_plane) {
  const int lane = threadIdx.x & 31;
  long long cell = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long step = (long long)gridDim.x * (blockDim.x >> 5);
  for (; cell < total_cells; cell += step) {
    const int b = (int)(cell / cells_per_img);
    const int s = __ldg(idx + cell);
    const size_t dst = (size_t)cell * C4;
    if (s >= 0) {
      const float4* src = reinterpret_cast<const float4*>(seg_emb) + (size_t)(__ldg(seg_off + b) + s) * C4;
      for (int c = lane; c < C4; c += 32) st4_fmt(grid, grid_plane, dst + c, __ldg(src + c));
    } else {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = lane; c < C4; c += 32) st4_fmt(grid, grid_plane, dst + c, z);
    }
  }
}

// Split source -> Split grid: a pure copy of the two bf16 planes, 16 bytes (8 elements) per lane.
__global__ void __launch_bounds__(256)
grid_scatter_planes_kernel(const __nv_bfloat16* __restrict__ seg_emb, long long emb_plane, const int32_t* __restrict__ idx,
                           const int32_t* __restrict__ seg_off, int cells_per_img, long long total_cells, int C8,
                           __nv_bfloat16* __restrict__ grid, long long grid_plane) {
  const int lane = threadIdx.x & 31;
  long long cell = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long step = (long long)gridDim.x * (blockDim.x >> 5);
  for (; cell < total_cells; cell += step) {
    const int b = (int)(cell / cells_per_img);
    const int s = __ldg(idx + cell);
    uint4* d0 = reinterpret_cast<uint4*>(grid) + cell * C8;
    uint4* d1 = reinterpret_cast<uint4*>(grid + grid_plane) + cell * C8;
    if (s >= 0) {
      const size_t row = (size_t)(__ldg(seg_off + b) + s) * C8;
      const uint4* s0 = reinterpret_cast<const uint4*>(seg_emb) + row;
      const uint4* s1 = reinterpret_cast<const uint4*>(seg_emb + emb_plane) + row;
      for (int c = lane; c < C8; c += 32) { d0[c] = __ldg(s0 + c); d1[c] = __ldg(s1 + c); }
    } else {
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      for (int c = lane; c < C8; c += 32) { d0[c] = z; d1[c] = z; }
    }
  }
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_resize_coords(const int64_t* coors, const int32_t* seg_off, const float* ratios, int B, int K,
                                 int32_t* out, vbg_stream_t stream) {
  VBG_REQUIRE(coors && seg_off && ratios && out && B > 0 && K >= 0, "vbg_resize_coords: bad arguments");
  if (K == 0) return VBG_OK;
  resize_coords_kernel<<<cdiv(K, 128), 128, 0, as_stream(stream)>>>(coors, seg_off, ratios, B, K, out);
  return check_launch("vbg_resize_coords");
}

extern "C" int vbg_segment_starts(const int32_t* seg_ids, const int32_t* tok_off, int B, int n_tok, int K,
                                  int32_t* seg_start, int32_t* status, vbg_stream_t stream) {
  VBG_REQUIRE(seg_ids && tok_off && seg_start && B > 0 && n_tok >= 0 && K >= 0, "vbg_segment_starts: bad arguments");
  segment_starts_kernel<<<1, 1024, 0, as_stream(stream)>>>(seg_ids, tok_off, B, n_tok, K, seg_start, status);
  return check_launch("vbg_segment_starts");
}

// The attention mask the reference gathers real rows with (model/BERTgrid_generator.py:152-158: emb[b][mask[b] == 1], then
// the assert that the rows match seg_indices[b]).  The packed layout assumes the mask is the PREFIX of n_tok[b] ones the
// reference's collate produces (data/SROIE_dataset.py:141-148); anything else raises status bit 2 instead of a syncing assert.
__global__ void mask_check_kernel(const int32_t* __restrict__ mask, int B, int L, const int32_t* __restrict__ tok_off,
                                  int32_t* __restrict__ status) {
  const int b = blockIdx.x;
  const int n = tok_off[b + 1] - tok_off[b];
  int bad = 0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) bad |= ((mask[(size_t)b * L + i] == 1) != (i < n));
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(status, 2);
}

extern "C" int vbg_mask_check(const int32_t* mask, int B, int L, const int32_t* tok_off, int32_t* status, vbg_stream_t stream) {
  VBG_REQUIRE(mask && tok_off && status && B > 0 && L > 0, "vbg_mask_check: bad arguments");
  mask_check_kernel<<<B, 256, 0, as_stream(stream)>>>(mask, B, L, tok_off, status);
  return check_launch("vbg_mask_check");
}

extern "C" int vbg_segment_reduce(const float* hidden, const int32_t* tok_row, const int32_t* seg_start, int K, int C,
                                  int mode, float* out, vbg_stream_t stream) {
  VBG_REQUIRE(hidden && tok_row && seg_start && out, "vbg_segment_reduce: null pointer");
  VBG_REQUIRE(C > 0 && C % 4 == 0 && aligned16(hidden) && aligned16(out), "vbg_segment_reduce: C %% 4 and 16B alignment required");
  VBG_REQUIRE(mode == VBG_AGG_MEAN || mode == VBG_AGG_FIRST, "vbg_segment_reduce: bad mode %d", mode);
  if (K == 0) return VBG_OK;
  int threads = C / 4 >= 256 ? 256 : ((C / 4 + 31) / 32) * 32;
  segment_reduce_kernel<<<K, threads, 0, as_stream(stream)>>>(hidden, tok_row, seg_start, C / 4, mode, out);
  return check_launch("vbg_segment_reduce");
}

extern "C" int vbg_box_index_map(const int32_t* boxes, const int32_t* seg_off, int B, int stride, int Hg, int Wg,
                                 int32_t* idx, vbg_stream_t stream) {
  VBG_REQUIRE(boxes && seg_off && idx && B > 0 && stride > 0 && Hg > 0 && Wg > 0, "vbg_box_index_map: bad arguments");
  VBG_REQUIRE(aligned16(boxes), "vbg_box_index_map: boxes must be 16B aligned");
  dim3 g(cdiv(Wg, kTileW), cdiv(Hg, kTileH), B);
  box_map_kernel<0><<<g, 256, 0, as_stream(stream)>>>(boxes, seg_off, nullptr, stride, Hg, Wg, idx, nullptr, nullptr, SegCeArgs{});
  return check_launch("vbg_box_index_map");
}

extern "C" int vbg_label_paint(const int32_t* boxes, const int32_t* seg_off, const int32_t* seg_cls, int B, int H, int W,
                               int64_t* pos_neg, int64_t* cls, vbg_stream_t stream) {
  VBG_REQUIRE(boxes && seg_off && seg_cls && pos_neg && cls && B > 0 && H > 0 && W > 0, "vbg_label_paint: bad arguments");
  VBG_REQUIRE(aligned16(boxes), "vbg_label_paint: boxes must be 16B aligned");
  dim3 g(cdiv(W, kTileW), cdiv(H, kTileH), B);
  box_map_kernel<1><<<g, 256, 0, as_stream(stream)>>>(boxes, seg_off, seg_cls, 1, H, W, nullptr, pos_neg, cls, SegCeArgs{});
  return check_launch("vbg_label_paint");
}

extern "C" int vbg_seg_ce_loss(const int32_t* boxes, const int32_t* seg_off, const int32_t* seg_cls, const float* logits, int B,
                               int H, int W, int up, int Ct, int c_split, float* workspace, size_t ws_bytes, float* out2,
                               vbg_stream_t stream) {
  VBG_REQUIRE(boxes && seg_off && seg_cls && logits && out2 && B > 0 && H > 0 && W > 0 && up > 0 && H % up == 0 && W % up == 0,
              "vbg_seg_ce_loss: bad arguments");
  VBG_REQUIRE(c_split > 0 && c_split < Ct && aligned16(boxes), "vbg_seg_ce_loss: bad channel split / alignment");
  dim3 g(cdiv(W, kTileW), cdiv(H, kTileH), B);
  const size_t nblk = (size_t)g.x * g.y * g.z;
  if (!workspace || ws_bytes < nblk * 2 * sizeof(float)) {
    set_error("vbg_seg_ce_loss: workspace of %zu bytes needed", nblk * 2 * sizeof(float));
    return VBG_EWORKSPACE;
  }
  SegCeArgs ce{logits, H / up, W / up, Ct, up, c_split, workspace};
  box_map_kernel<2><<<g, 256, 0, as_stream(stream)>>>(boxes, seg_off, seg_cls, 1, H, W, nullptr, nullptr, nullptr, ce);
  int rc = check_launch("vbg_seg_ce_loss");
  if (rc) return rc;
  seg_ce_finish_kernel<<<1, 256, 0, as_stream(stream)>>>(workspace, (int)nblk, 1.0 / ((double)B * H * W), out2);
  return check_launch("vbg_seg_ce_loss(finish)");
}

extern "C" int vbg_grid_scatter(const float* seg_emb, const int32_t* idx, const int32_t* seg_off, int B, int cells, int C,
                                float* grid, vbg_stream_t stream) {
  return vbg_grid_scatter_x(seg_emb, 0, idx, seg_off, B, cells, C, grid, 0, stream);
}

extern "C" int vbg_grid_scatter_x(const void* seg_emb_v, long long emb_plane, const int32_t* idx, const int32_t* seg_off, int B,
                                  int cells, int C, void* grid, long long grid_plane, vbg_stream_t stream) {
  VBG_REQUIRE(seg_emb_v && idx && seg_off && grid && B > 0 && cells > 0, "vbg_grid_scatter: bad arguments");
  VBG_REQUIRE(C > 0 && C % 4 == 0 && fmt_ok(seg_emb_v, emb_plane) && fmt_ok(grid, grid_plane), "vbg_grid_scatter: C %% 4 and 16B alignment required");
  VBG_REQUIRE(emb_plane == 0 || (grid_plane > 0 && C % 8 == 0), "vbg_grid_scatter: a bf16-plane source needs a bf16-plane grid and C %% 8 == 0");
  if (emb_plane > 0) {
    long long total = (long long)B * cells;
    int blocks = (int)((total + 7) / 8);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    grid_scatter_planes_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(seg_emb_v), emb_plane, idx, seg_off,
                                                                     cells, total, C / 8, reinterpret_cast<__nv_bfloat16*>(grid), grid_plane);
    return check_launch("vbg_grid_scatter");
  }
  const float* seg_emb = reinterpret_cast<const float*>(seg_emb_v);
  long long total = (long long)B * cells;
  int blocks = (int)((total + 7) / 8);
  int cap = kNumSMs * 16;                      // 8 resident CTAs/SM x 2 waves, then grid-stride
  if (blocks > cap) blocks = cap;
  grid_scatter_kernel<<<blocks, 256, 0, as_stream(stream)>>>(seg_emb, idx, seg_off, cells, total, C / 4, grid, grid_plane);
  return check_launch("vbg_grid_scatter");
}
