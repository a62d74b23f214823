// GridROIAlign forward, streamed: the default kernel for P = 7, C in {128, 256}.
//
// Replaces torchvision.ops.roi_align as called by reference model/grid_roi_align.py:37-41,81 (see vbg_roi.cu for the
// algorithm restatement; the geometry below is the same arithmetic, operation for operation, so the sample-grid table stays
// bit-exact, and the separable accumulation order is the one of roi_align_row_kernel, so values are identical to it up to
// the final multiplication by 1 / count).
//
// Why a new structure: ncu of roi_align_row_kernel (profiles/r2_ncu_roi_kernels.txt) shows it ISSUE-bound (41 M warp
// instructions, 66 % issue-active, DRAM at 28 %): every window pixel costs address arithmetic + a cp.async on the way in, a
// merge pass (planes -> fp32) in shared memory and then generic per-bin tap loops, all serialised by two CTA barriers.
// Here the feature map streams through shared memory ONE WINDOW ROW AT A TIME and nothing else touches it:
//
//   * persistent CTAs (two per SM), each walks ROIs k = blockIdx.x, + gridDim.x, ...
//   * one PREPARE warp: computes the ROI geometry and the separable per-bin weight tables up to three ROIs ahead into a
//     ring of records (so no consumer ever executes a division and the table latency never sits between two ROIs), and one
//     ISSUER thread that issues ONE `cp.async.bulk` (TMA, 1-D) per window row and plane -- in NHWC a window row (cols x C channels) is a
//     contiguous run of cols * C * 2 bytes in each bf16 plane (cols * C * 4 in fp32) -- into a byte ring guarded by
//     full / empty mbarriers.  The fetch costs ~6 instructions per ROW instead of ~15 per PIXEL and is always in flight.
//   * 7 * (C / 128) CONSUMER warps, warp = (bin column pw, 128-channel group), lane = channel quad.  Per landed row a warp
//     forms t = sum_i wx[pw][i] * f[y][x0 + i] straight from the bf16 planes (hi + lo is exact in fp32: packed
//     add.f32x2 + fma.f32x2, no merge pass, no CTA barrier), then adds wy[ph][y - y0[ph]] * t into the 7 bin-row
//     accumulators it keeps in registers.  Each window row is read from HBM once and from shared memory ~2.6 times
//     (the direct separable form re-reads it ~4.6 times); weights come from lane registers via shuffles.
//   * ROIs wider than the ring allows are processed as several column SEGMENTS (each a group of neighbouring bin columns);
//     ROIs whose tables do not fit (a bin spanning more than 32 pixels) take per-sample global taps inside the same kernel.
#include "vbg_tc.cuh"

namespace vbg {

constexpr int kSP = 7;              // output bins per side
constexpr int kSSpan = 32;          // table width: pixels one bin may touch per axis
constexpr int kSNE = 16;            // ring entries (rows in flight)
constexpr int kSRing = 96 * 1024;   // ring bytes per CTA (two CTAs per SM)

struct RoiRec {
  int b, y_lo, x_lo, rows, cols, mode, gh, gw;          // mode 0: streamed rows, 1: per-sample global taps
  int nseg, pad0, pad1, pad2;
  int seg_pa[kSP], seg_pb[kSP], seg_xa[kSP], seg_cols[kSP];   // segment s covers bin columns [pa, pb), pixels [xa, xa + cols)
  int x0[kSP], nx[kSP], y0[kSP], ny[kSP];                      // table origin (relative to x_lo / y_lo) and length per bin
  float inv_count, sh, sw, bh, bw, padf[3];
  float wx[kSP][kSSpan], wy[kSP][kSSpan];
};

// Bounded wait without the printf of mbar_wait (its call frame costs the streaming loop registers): a protocol bug traps.
__device__ __forceinline__ void mbar_wait_lean(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}
// mbarrier operations on precomputed 32-bit shared addresses (no generic -> shared conversion inside the row loop)
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// d += w * (a + b) on two packed fp32 lanes (FADD2 + FFMA2)
__device__ __forceinline__ void acc2(float& d0, float& d1, float w, float a0, float a1, float b0, float b1) {
  uint64_t a, b, c, ww;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(b));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(ww), "l"(a));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}
__device__ __forceinline__ void fma2(float& d0, float& d1, float w, float a0, float a1) {
  uint64_t a, c, ww;
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(ww) : "f"(w));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c) : "l"(ww), "l"(a));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(c));
}

// ---- producer: geometry + tables of ROI k into `rec` (whole warp), operation for operation as roi_align_row_kernel
__device__ __noinline__ void roi_prepare(RoiRec* rec, int k, int lane, int B, int Hf, int Wf, int C, const int4 bx, int b,
                                            float scale, int32_t* __restrict__ sample_grid, int max_cols) {
  constexpr int P = kSP;
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
  for (int i = lane; i < 2 * P * kSSpan; i += 32) (&rec->wx[0][0])[i] = 0.f;     // wx and wy are contiguous
  __syncwarp();
  int ok = 1;
  if (lane < 2 * P) {
    const int axis = lane / P, pb = lane - axis * P;            // axis 0: x (columns), 1: y (rows)
    const int g = axis ? gh : gw, dim = axis ? Hf : Wf, lo_w = axis ? y_lo : x_lo, hi_w = axis ? y_hi : x_hi;
    const float start = axis ? sh : sw, bin = axis ? bh : bw;
    float* w = axis ? rec->wy[pb] : rec->wx[pb];
    int base = 0, cnt = 0;
    for (int i = 0; i < g; ++i) {
      float c = __fadd_rn(__fadd_rn(start, __fmul_rn((float)pb, bin)), __fdiv_rn(__fmul_rn((float)i + 0.5f, bin), (float)g));
      if (c < -1.0f || c > (float)dim) continue;
      c = fmaxf(c, 0.f);
      int lo = (int)c, hi;
      if (lo >= dim - 1) { hi = lo = dim - 1; c = (float)lo; } else { hi = lo + 1; }
      const float l = c - (float)lo, h = 1.f - l;
      if (cnt == 0) base = lo;
      if (hi - base >= kSSpan || lo < base || lo < lo_w || hi > hi_w) { ok = 0; break; }
      w[lo - base] += h;
      w[hi - base] += l;
      cnt = hi - base + 1;
    }
    if (axis) { rec->y0[pb] = base - lo_w; rec->ny[pb] = cnt; } else { rec->x0[pb] = cnt ? base - lo_w : 0; rec->nx[pb] = cnt; }
  }
  const bool all_ok = __ballot_sync(0xffffffffu, ok != 0) == 0xffffffffu;
  __syncwarp();
  if (lane == 0) {
    rec->b = b; rec->y_lo = y_lo; rec->x_lo = x_lo; rec->rows = y_hi - y_lo + 1; rec->cols = x_hi - x_lo + 1;
    rec->gh = gh; rec->gw = gw;
    rec->inv_count = __fdiv_rn(1.0f, (float)max(gh * gw, 1));
    rec->sh = sh; rec->sw = sw; rec->bh = bh; rec->bw = bw;
    if (sample_grid) { sample_grid[2 * k] = gh; sample_grid[2 * k + 1] = gw; }
    // column segments: greedy groups of neighbouring non-empty bin columns whose pixel extent fits `max_cols`
    int mode = all_ok ? 0 : 1, nseg = 0;
    bool any_row = false;
    for (int p = 0; p < P; ++p) any_row |= rec->ny[p] > 0;
    if (mode == 0 && any_row) {
      int p = 0;
      while (p < P) {
        if (rec->nx[p] == 0) { ++p; continue; }
        const int xa = rec->x0[p];
        int xb = xa + rec->nx[p], q = p + 1;
        if (xb - xa > max_cols) { mode = 1; break; }
        while (q < P && (rec->nx[q] == 0 || max(xb, rec->x0[q] + rec->nx[q]) - xa <= max_cols)) {
          if (rec->nx[q]) xb = max(xb, rec->x0[q] + rec->nx[q]);
          ++q;
        }
        rec->seg_pa[nseg] = p; rec->seg_pb[nseg] = q; rec->seg_xa[nseg] = xa; rec->seg_cols[nseg] = xb - xa;
        ++nseg;
        p = q;
      }
    }
    rec->mode = mode;
    rec->nseg = mode == 0 ? nseg : 0;
  }
  __syncwarp();
}

// A bin spans more than kSSpan pixels (or a table assumption failed): per-sample taps straight from global memory, the
// arithmetic of roi_align_kernel.  Out of line: rare, and its registers must not weigh on the streaming loop.
__device__ __noinline__ void roi_direct_taps(const RoiRec* rec, const void* __restrict__ feat, long long feat_plane, int Hf, int Wf,
                                             int C4, int k, int pw, int grp, int lane, void* __restrict__ out, long long out_plane) {
  constexpr int P = kSP;
  const float s_h = rec->sh, s_w = rec->sw, b_h = rec->bh, b_w = rec->bw;
  const int g_h = rec->gh, g_w = rec->gw;
  const size_t fq = (size_t)rec->b * Hf * Wf * C4 + grp * 32 + lane;
#pragma unroll 1
  for (int ph = 0; ph < P; ++ph) {
    float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
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
        a4.x += w1 * a.x + w2 * bb.x + w3 * cc.x + w4 * d.x;
        a4.y += w1 * a.y + w2 * bb.y + w3 * cc.y + w4 * d.y;
        a4.z += w1 * a.z + w2 * bb.z + w3 * cc.z + w4 * d.z;
        a4.w += w1 * a.w + w2 * bb.w + w3 * cc.w + w4 * d.w;
      }
    }
    const float cnt = (float)max(g_h * g_w, 1);
    a4.x = __fdiv_rn(a4.x, cnt); a4.y = __fdiv_rn(a4.y, cnt); a4.z = __fdiv_rn(a4.z, cnt); a4.w = __fdiv_rn(a4.w, cnt);
    st4_fmt(out, out_plane, ((size_t)k * P * P + (size_t)ph * P + pw) * C4 + grp * 32 + lane, a4);
  }
}

constexpr int kSRecs = 4;           // ROI records in flight: the prepare warp runs up to three ROIs ahead of the consumers

template <bool kPlanes>
__global__ void __launch_bounds__(512, 2)
roi_align_stream_kernel(const void* __restrict__ feat, long long feat_plane, int B, int Hf, int Wf, int C,
                        const int32_t* __restrict__ boxes, const int32_t* __restrict__ seg_off, int K, float scale,
                        void* __restrict__ out, long long out_plane, int32_t* __restrict__ sample_grid) {
  constexpr int P = kSP;
  extern __shared__ __align__(128) unsigned char ring[];
  __shared__ __align__(16) RoiRec recs[kSRecs];
  __shared__ __align__(8) uint64_t full[kSNE], empty[kSNE], rec_full[kSRecs], rec_empty[kSRecs];
  __shared__ uint32_t p_vstart[kSNE];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncons = P * (C >> 7);                      // consumer warps; warp `ncons` prepares records, warp `ncons + 1` issues rows
  const uint32_t px_bytes = (uint32_t)C * 4u;          // one pixel, both planes (or fp32)
  const int max_cols = (int)((uint32_t)(kSRing / 2) / px_bytes);

  if (tid == 0) {
    for (int i = 0; i < kSNE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], (uint32_t)ncons); }
    for (int i = 0; i < kSRecs; ++i) { mbar_init(&rec_full[i], 1); mbar_init(&rec_empty[i], (uint32_t)ncons + 1u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == ncons) {
    // ================================================================ prepare warp: geometry + weight tables, ROIs ahead
    int it = 0;
    int4 bx_next = make_int4(0, 0, 0, 0);
    if ((int)blockIdx.x < K) bx_next = __ldg(reinterpret_cast<const int4*>(boxes) + blockIdx.x);
    for (int k = blockIdx.x; k < K; k += gridDim.x, ++it) {
      const int rs = it % kSRecs;
      const int4 bx = bx_next;
      if (k + (int)gridDim.x < K) bx_next = __ldg(reinterpret_cast<const int4*>(boxes) + k + gridDim.x);
      int b;
      if (B <= 31) {
        const int so = (lane >= 1 && lane < B) ? __ldg(seg_off + lane) : 0x7fffffff;
        b = __popc(__ballot_sync(0xffffffffu, so <= k));
      } else {
        b = sample_of(seg_off, B, k);
      }
      mbar_wait_lean(&rec_empty[rs], ((it / kSRecs) & 1) ^ 1);      // consumers and issuer are done with this slot's last ROI
      roi_prepare(&recs[rs], k, lane, B, Hf, Wf, C, bx, b, scale, sample_grid, max_cols);
      if (lane == 0) mbar_arrive(&rec_full[rs]);
      __syncwarp();
    }
    return;
  }
  if (warp == ncons + 1) {
    // ================================================================ issuer: one bulk copy per window row and plane
    if (lane != 0) return;
    uint32_t n = 0, tail = 0, poff = 0, voff = 0;
    int it = 0;
    for (int k = blockIdx.x; k < K; k += gridDim.x, ++it) {
      const int rs = it % kSRecs;
      mbar_wait_lean(&rec_full[rs], (it / kSRecs) & 1);
      const RoiRec* rec = &recs[rs];
      const int nseg = rec->nseg, rows = rec->rows;
      const size_t row0 = ((size_t)rec->b * Hf + rec->y_lo) * Wf + rec->x_lo;        // pixel index of the window origin
      for (int s = 0; s < nseg; ++s) {
        const uint32_t sz = (uint32_t)rec->seg_cols[s] * px_bytes;
        const size_t pix0 = row0 + rec->seg_xa[s];
        for (int y = 0; y < rows; ++y) {
          if (poff + sz > (uint32_t)kSRing) { voff += (uint32_t)kSRing - poff; poff = 0; }
          while (tail != n && (n - tail >= (uint32_t)kSNE || voff + sz - p_vstart[tail % kSNE] > (uint32_t)kSRing)) {
            mbar_wait_lean(&empty[tail % kSNE], (tail / kSNE) & 1);
            ++tail;
          }
          const uint32_t e = n % kSNE;
          p_vstart[e] = voff;
          mbar_expect_tx(&full[e], sz);
          const size_t pix = pix0 + (size_t)y * Wf;
          if (kPlanes) {
            const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(feat) + pix * C;
            bulk_g2s(ring + poff, src, sz / 2, &full[e]);
            bulk_g2s(ring + poff + sz / 2, src + feat_plane, sz / 2, &full[e]);
          } else {
            bulk_g2s(ring + poff, reinterpret_cast<const float*>(feat) + pix * C, sz, &full[e]);
          }
          poff += sz; voff += sz; ++n;
        }
      }
      mbar_arrive(&rec_empty[rs]);
    }
    return;
  }
  if (warp > ncons + 1) return;

  // ================================================================== consumer warps
  const int pw = warp % P, grp = warp / P;
  const int C4 = C >> 2;
  const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(&full[0]), empty_s = smem_u32(&empty[0]);
  uint32_t n = 0, poff = 0;
  int it = 0;
  for (int k = blockIdx.x; k < K; k += gridDim.x, ++it) {
    const int rs = it % kSRecs;
    mbar_wait_lean(&rec_full[rs], (it / kSRecs) & 1);
    const RoiRec* rec = &recs[rs];
    const int rows = rec->rows, nseg = rec->nseg, mode = rec->mode;
    if (mode == 0) {
      float4 acc[P];
#pragma unroll
      for (int p = 0; p < P; ++p) acc[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int nx = rec->nx[pw], x0 = rec->x0[pw];
      const float wxr = rec->wx[pw][lane];
      // lane p (mod 8) looks up the bin-row weight Wy[p][y] of the current window row; the row update then broadcasts the
      // seven weights with shuffles (0 where the row does not contribute): branch-free, 7 shuffles + 14 packed FMAs per row
      const int myp = lane & 7;
      const int y0p = myp < P ? rec->y0[myp] : 0, nyp = myp < P ? rec->ny[myp] : 0;
      const float* wyp = rec->wy[myp < P ? myp : 0];
      for (int s = 0; s < nseg; ++s) {
        const uint32_t sz = (uint32_t)rec->seg_cols[s] * px_bytes;
        const bool mine = pw >= rec->seg_pa[s] && pw < rec->seg_pb[s] && nx > 0;
        const uint32_t tap0 = (uint32_t)((x0 - rec->seg_xa[s]) * C4 + grp * 32 + lane);    // first table pixel, this lane's quad
        for (int y = 0; y < rows; ++y) {
          if (poff + sz > (uint32_t)kSRing) poff = 0;
          const uint32_t e = n % kSNE;
          const int j = y - y0p;
          const float wrow = ((unsigned)j < (unsigned)nyp) ? wyp[j] : 0.f;
          mbar_wait_s(full_s + e * 8u, (n / kSNE) & 1);
          if (mine) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kPlanes) {
              // 32-bit shared-memory addresses: hi plane row, lo plane row = + sz / 2; one pixel = C * 2 bytes per plane
              uint32_t a_hi = ring_s + poff + tap0 * 8u;
              const uint32_t half = sz >> 1;
#pragma unroll 2
              for (int i = 0; i < nx; ++i, a_hi += (uint32_t)C * 2u) {
                const float w = __shfl_sync(0xffffffffu, wxr, i);
                const uint2 h = lds64(a_hi), l = lds64(a_hi + half);
                acc2(t.x, t.y, w, bf16lo_f(h.x), bf16hi_f(h.x), bf16lo_f(l.x), bf16hi_f(l.x));
                acc2(t.z, t.w, w, bf16lo_f(h.y), bf16hi_f(h.y), bf16lo_f(l.y), bf16hi_f(l.y));
              }
            } else {
              uint32_t a_row = ring_s + poff + tap0 * 16u;
#pragma unroll 2
              for (int i = 0; i < nx; ++i, a_row += (uint32_t)C * 4u) {
                const float w = __shfl_sync(0xffffffffu, wxr, i);
                const float4 v = lds128(a_row);
                fma2(t.x, t.y, w, v.x, v.y);
                fma2(t.z, t.w, w, v.z, v.w);
              }
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
              const float w = __shfl_sync(0xffffffffu, wrow, p);
              fma2(acc[p].x, acc[p].y, w, t.x, t.y);
              fma2(acc[p].z, acc[p].w, w, t.z, t.w);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive_s(empty_s + e * 8u);
          poff += sz; ++n;
        }
      }
      const float inv_count = rec->inv_count;
      __syncwarp();
      if (lane == 0) mbar_arrive(&rec_empty[rs]);                   // the record may be rewritten; acc is in registers
      const size_t o_base = ((size_t)k * P * P + pw) * C4 + grp * 32 + lane;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        acc[p].x *= inv_count; acc[p].y *= inv_count; acc[p].z *= inv_count; acc[p].w *= inv_count;
        st4_fmt(out, out_plane, o_base + (size_t)p * P * C4, acc[p]);
      }
    } else {
      roi_direct_taps(rec, feat, feat_plane, Hf, Wf, C4, k, pw, grp, lane, out, out_plane);
      __syncwarp();
      if (lane == 0) mbar_arrive(&rec_empty[rs]);
    }
  }
}

int launch_roi_align_stream(const void* feat, long long feat_plane, int B, int Hf, int Wf, int C, const int32_t* boxes,
                            const int32_t* seg_off, int K, float scale, void* out, long long out_plane, int32_t* sample_grid,
                            cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e1 = cudaFuncSetAttribute(roi_align_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSRing);
    cudaError_t e2 = cudaFuncSetAttribute(roi_align_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSRing);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("vbg_roi_align_fwd: smem opt-in failed"); return VBG_ECUDA; }
    attr = true;
  }
  const int ctas = 2 * kNumSMs;
  const unsigned grid = (unsigned)(K < ctas ? K : ctas);
  const unsigned threads = 32u * (unsigned)(kSP * (C / 128) + 2);
  if (feat_plane != 0)
    roi_align_stream_kernel<true><<<grid, threads, kSRing, s>>>(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off, K, scale, out, out_plane, sample_grid);
  else
    roi_align_stream_kernel<false><<<grid, threads, kSRing, s>>>(feat, feat_plane, B, Hf, Wf, C, boxes, seg_off, K, scale, out, out_plane, sample_grid);
  return check_launch("vbg_roi_align_fwd");
}

}  // namespace vbg
