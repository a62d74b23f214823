// Variable-length multi-head attention on the 5th-generation tensor cores (VBG_PREC_BF16X3 / TF32 requests).
//
//   out[r, h*64:(h+1)*64] = softmax(Q K^T / 8) V      per (sequence, head), packed real rows only
//
// Replaces the eager attention inside HF BertModel as called by reference model/BERTgrid_generator.py:134.
// One CTA = 128 queries of one (sequence, head); sequences hold at most 512 rows (BERT's position limit), so the
// whole score row block S[128, len] lives in TMEM (<= 512 fp32 columns) and the softmax is exact two-pass
// (row max, then exp / sum), not an online approximation schedule.
//
//   warp 0      one thread issues tcgen05.mma.kind::f16: S = Q K^T (N = 128 keys per step), later O += P V (N = 64)
//   warps 1-4   128 threads, thread == query row == TMEM lane:
//               phase A  load Q / K rows from HBM (256 B per thread), split fp32 -> bf16 hi/lo, write the K-major
//                        SWIZZLE_128B operand tiles;
//               phase B  row max over S (tcgen05.ld);
//               phase C  per 64-key unit: p = exp2((s - max) * log2e/8) -> bf16 hi/lo -> P operand tiles; the matching
//                        V rows are transposed into the V^T operand tiles by the same threads; row sum in fp32;
//               phase D  O (TMEM) * 1/sum -> HBM.
// Every product uses the 3-term bf16 split (a1*b1 + a2*b1 + a1*b2, fp32 accumulate) of vbg_gemm_tc3.cu, so the
// result is fp32-class.  O re-uses TMEM columns [0,64) once unit 0 of S has been consumed.
#include "vbg_tc.cuh"
#include <cuda_bf16.h>

namespace vbg {

constexpr int kAtThreads = 160;
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
constexpr uint32_t kTileQ = 128 * 128;       // one bf16 operand tile of 128 rows x 64 elements
constexpr uint32_t kVtTile = 64 * 128;       // V^T tile: 64 d-rows x 64 keys (bf16)

// 64 fp32 of one row (src, 256 B, or zeros) -> row r of the hi / lo operand tiles (128 B each, swizzled 16-byte chunks)
__device__ __forceinline__ void load_split_row(const float* __restrict__ src, bool valid, uint8_t* t1, uint8_t* t2, int r) {
  float4 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = valid ? __ldg(reinterpret_cast<const float4*>(src) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t xr = (uint32_t)(r & 7);
#pragma unroll
  for (int c = 0; c < 8; ++c) {                 // bf16 chunk c = elements 8c .. 8c+7 = float4 2c, 2c+1
    uint4 h, l;
    split2(v[2 * c].x, v[2 * c].y, h.x, l.x); split2(v[2 * c].z, v[2 * c].w, h.y, l.y);
    split2(v[2 * c + 1].x, v[2 * c + 1].y, h.z, l.z); split2(v[2 * c + 1].z, v[2 * c + 1].w, h.w, l.w);
    const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)c) ^ xr) << 4);
    *reinterpret_cast<uint4*>(t1 + off) = h;
    *reinterpret_cast<uint4*>(t2 + off) = l;
  }
}

__global__ void __launch_bounds__(kAtThreads)
attention_tc_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ cu, int heads, float scale_log2e,
                    float* __restrict__ out) {
  const int seq = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * 128;
  const int row0 = cu[seq], len = cu[seq + 1] - row0;
  if (q0 >= len) return;                                        // whole CTA, before any barrier / allocation
  const int n_chunks = (len + 127) >> 7, n_units = (len + 63) >> 6;
  const int hidden = heads * 64, ld = 3 * hidden;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_op = smem;                                         // Q1 | Q2
  uint8_t* kp = smem + 2 * kTileQ;                              // 2 x (K1 | K2), later 2 x (P1 | P2)
  uint8_t* vt = kp + 4 * kTileQ;                                // 2 x (VT1 | VT2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(vt + 4 * kVtTile);
  uint64_t *q_ready = bars, *k_ready = bars + 1, *k_empty = bars + 3, *s_full = bars + 5, *p_ready = bars + 6,
           *p_empty = bars + 8, *o_full = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(q_ready, 128);
      for (int i = 0; i < 2; ++i) { mbar_init(&k_ready[i], 128); mbar_init(&k_empty[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&p_empty[i], 1); }
      mbar_init(s_full, 1); mbar_init(o_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== MMA issuer
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, 128, 128), idesc_o = make_idesc(kFmtBF16, 128, 64);
      const uint64_t q1 = make_sw128_desc(smem_u32(q_op)), q2 = make_sw128_desc(smem_u32(q_op + kTileQ));
      mbar_wait(q_ready, 0);
      for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1;
        mbar_wait(&k_ready[b], (c >> 1) & 1);
        tc_fence_after();
        const uint32_t kb = smem_u32(kp + b * 2 * kTileQ);
        const uint64_t k1 = make_sw128_desc(kb), k2 = make_sw128_desc(kb + kTileQ);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t o = (uint64_t)(2 * k);
          umma_bf16(tmem_base + (uint32_t)(c * 128), q1 + o, k1 + o, idesc_s, k != 0);
          umma_bf16(tmem_base + (uint32_t)(c * 128), q2 + o, k1 + o, idesc_s, 1);
          umma_bf16(tmem_base + (uint32_t)(c * 128), q1 + o, k2 + o, idesc_s, 1);
        }
        umma_commit(&k_empty[b]);
      }
      umma_commit(s_full);
      for (int u = 0; u < n_units; ++u) {
        const int b = u & 1;
        mbar_wait(&p_ready[b], (u >> 1) & 1);
        tc_fence_after();
        const uint32_t pb = smem_u32(kp + b * 2 * kTileQ), vb = smem_u32(vt + b * 2 * kVtTile);
        const uint64_t p1 = make_sw128_desc(pb), p2 = make_sw128_desc(pb + kTileQ);
        const uint64_t v1 = make_sw128_desc(vb), v2 = make_sw128_desc(vb + kVtTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t o = (uint64_t)(2 * k);
          umma_bf16(tmem_base, p1 + o, v1 + o, idesc_o, (u | k) != 0);
          umma_bf16(tmem_base, p2 + o, v1 + o, idesc_o, 1);
          umma_bf16(tmem_base, p1 + o, v2 + o, idesc_o, 1);
        }
        umma_commit(&p_empty[b]);
      }
      umma_commit(o_full);
    }
  } else {
    // ===== loaders / softmax: thread == query row == TMEM lane
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int t = threadIdx.x - 32;                             // 0..127, used for the V^T transposer
    const float* base = qkv + (size_t)row0 * ld + head * 64;
    // --- phase A: Q, then K chunks
    load_split_row(base + (size_t)(q0 + r) * ld, q0 + r < len, q_op, q_op + kTileQ, r);
    fence_proxy_async_smem();
    mbar_arrive(q_ready);
    for (int c = 0; c < n_chunks; ++c) {
      const int b = c & 1;
      if (c >= 2) mbar_wait(&k_empty[b], ((c >> 1) - 1) & 1);
      const int key = c * 128 + r;
      uint8_t* k1 = kp + b * 2 * kTileQ;
      load_split_row(base + hidden + (size_t)key * ld, key < len, k1, k1 + kTileQ, r);
      fence_proxy_async_smem();
      mbar_arrive(&k_ready[b]);
    }
    // --- phase B: exact row max
    mbar_wait(s_full, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float m = -INFINITY;
    for (int c0 = 0; c0 < n_chunks * 128; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(lane_addr + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) if (c0 + j < len) m = fmaxf(m, __uint_as_float(v[j]));
    }
    // --- phase C: probabilities and V^T, 64 keys per unit
    float sum = 0.f;
    const float m_scaled = m * scale_log2e;
    const uint32_t xr = (uint32_t)(r & 7);
    const int vkey = t & 63, vd0 = (t >> 6) * 32;
    for (int u = 0; u < n_units; ++u) {
      const int b = u & 1;
      if (u >= 2) mbar_wait(&p_empty[b], ((u >> 1) - 1) & 1);
      {   // V^T tiles of this unit: element (d, key) <- V[key][d]
        const int key = u * 64 + vkey;
        const float* vsrc = base + 2 * hidden + (size_t)key * ld + vd0;
        float4 vv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[i] = key < len ? __ldg(reinterpret_cast<const float4*>(vsrc) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint8_t* v1 = vt + b * 2 * kVtTile;
        const uint32_t kchunk = (uint32_t)(vkey >> 3), kin = (uint32_t)(vkey & 7) * 2u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float f[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int d = vd0 + 4 * i + e;
            const __nv_bfloat16 h = __float2bfloat16_rn(f[e]);
            const __nv_bfloat16 l = __float2bfloat16_rn(f[e] - __bfloat162float(h));
            const uint32_t off = (uint32_t)d * 128u + ((kchunk ^ (uint32_t)(d & 7)) << 4) + kin;
            *reinterpret_cast<__nv_bfloat16*>(v1 + off) = h;
            *reinterpret_cast<__nv_bfloat16*>(v1 + kVtTile + off) = l;
          }
        }
      }
      uint8_t* p1 = kp + b * 2 * kTileQ + (uint32_t)r * 128u;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        const int c0 = u * 64 + half * 32;
        tmem_ld32(lane_addr + (uint32_t)c0, v);
        uint32_t hi[16], lo[16];
        if (c0 + 32 <= len) {                     // CTA-uniform fast path: arguments are <= 0, ex2.approx underflows to 0
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float pa = ex2_approx(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -m_scaled));
            const float pb = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -m_scaled));
            sum += pa + pb;
            split2(pa, pb, hi[j], lo[j]);
          }
        } else {                                  // chunk holding the sequence end: masked keys get exponent -inf -> p = 0
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float xa = (c0 + 2 * j < len) ? fmaf(__uint_as_float(v[2 * j]), scale_log2e, -m_scaled) : -INFINITY;
            const float xb = (c0 + 2 * j + 1 < len) ? fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -m_scaled) : -INFINITY;
            const float pa = ex2_approx(xa), pb = ex2_approx(xb);
            sum += pa + pb;
            split2(pa, pb, hi[j], lo[j]);
          }
        }
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const uint32_t off = (((uint32_t)(4 * half + qd)) ^ xr) << 4;
          *reinterpret_cast<uint4*>(p1 + off) = make_uint4(hi[4 * qd], hi[4 * qd + 1], hi[4 * qd + 2], hi[4 * qd + 3]);
          *reinterpret_cast<uint4*>(p1 + kTileQ + off) = make_uint4(lo[4 * qd], lo[4 * qd + 1], lo[4 * qd + 2], lo[4 * qd + 3]);
        }
      }
      tc_fence_before();                         // S columns of this unit are consumed before O may overwrite them
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[b]);
    }
    // --- phase D: normalise and store
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = __fdiv_rn(1.0f, sum);
    float* orow = out + (size_t)(row0 + q0 + r) * hidden + head * 64;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld32(lane_addr + (uint32_t)(half * 32), v);
      if (q0 + r < len) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(orow + half * 32 + 4 * j) =
              make_float4(__uint_as_float(v[4 * j]) * inv, __uint_as_float(v[4 * j + 1]) * inv,
                          __uint_as_float(v[4 * j + 2]) * inv, __uint_as_float(v[4 * j + 3]) * inv);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------ v2: operands arrive pre-split from the QKV GEMM
// The QKV projection writes bf16 hi / lo planes (VBG_OUT_SPLIT_BF16), so nothing is converted here: TMA drops Q, K
// (K-major: row = token, 64 head dims = 128 B) and V (the same rows, used as an MN-major B operand -- no transpose)
// straight into SWIZZLE_128B operand tiles.  The 128 softmax threads only do row max, exp, and the hi/lo split of P.
//   warp 0 TMA producer | warp 1 MMA issuer | warps 2-9 softmax (two groups; thread == query row == TMEM lane)
constexpr int kAt2Threads = 320;
constexpr uint32_t kVTile = 64 * 128;        // 64 keys x 64 dims (bf16)

// Training mode (HF BertSelfAttention under model.train(): dropout on the attention probabilities,
// attention_probs_dropout_prob, reached through model/BERTgrid_generator.py:134): the kernel optionally
//   * drops probabilities with a counter-based mask -- a pure function of (seed, packed query row, key index, head), so the
//     backward kernels (vbg_attn_bwd_tc.cu) regenerate it; kept entries are scaled by 1 / (1 - p).  The row sum (softmax
//     normalisation) is taken over the UN-dropped probabilities, as softmax -> dropout does;
//   * stores the base-2 row log-sum-exp  lse2[row, head] = max * scale * log2(e) + log2(sum)  the backward rebuilds P from.
__global__ void __launch_bounds__(kAt2Threads, 1)
attention_split_kernel(const __grid_constant__ CUtensorMap tmQKh, const __grid_constant__ CUtensorMap tmQKl,
                       const __grid_constant__ CUtensorMap tmVh, const __grid_constant__ CUtensorMap tmVl,
                       const int32_t* __restrict__ cu, int heads, float scale_log2e, void* __restrict__ out,
                       long long out_plane, long long* __restrict__ dbg, float* __restrict__ lse2, uint32_t drop_thr,
                       float drop_inv_keep, uint32_t seed, const unsigned long long* __restrict__ step_seed) {
  if (drop_thr && step_seed) seed = attn_fold_step(seed, __ldg(step_seed));
#define AT_STAMP(slot) do { if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) dbg[slot] = clock64(); } while (0)
  const int seq = blockIdx.z, head = blockIdx.y, q0 = blockIdx.x * 128;
  const int row0 = cu[seq], len = cu[seq + 1] - row0;
  if (q0 >= len) return;
  const int n_chunks = (len + 127) >> 7, n_units = (len + 63) >> 6;
  const int hidden = heads * 64;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* q_op = smem;                                         // Q1 | Q2                32 KB
  uint8_t* k_op = q_op + 2 * kTileQ;                            // 2 x (K1 | K2)          64 KB
  uint8_t* v_op = k_op + 4 * kTileQ;                            // 2 x (V1 | V2)          32 KB
  uint8_t* p_op = v_op + 4 * kVTile;                            // 2 x (P1 | P2)          64 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_op + 4 * kTileQ);
  uint64_t *q_full = bars, *k_full = bars + 1, *k_empty = bars + 3, *s_full = bars + 5, *v_full = bars + 6,
           *p_ready = bars + 8, *pv_done = bars + 10, *o_full = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* xch = reinterpret_cast<float*>(bars + 14);             // [2][128] row max / row sum exchange between the groups

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) AT_STAMP(0);
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmQKh); prefetch_tmap(&tmQKl); prefetch_tmap(&tmVh); prefetch_tmap(&tmVl); }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(o_full, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); mbar_init(&v_full[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&pv_done[i], 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      AT_STAMP(1);
      mbar_expect_tx(q_full, 2 * kTileQ);
      tma_load_2d(&tmQKh, q_full, q_op, head * 64, row0 + q0);
      tma_load_2d(&tmQKl, q_full, q_op + kTileQ, head * 64, row0 + q0);
      for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1;
        mbar_wait(&k_empty[b], ((c >> 1) & 1) ^ 1);
        uint8_t* dst = k_op + b * 2 * kTileQ;
        mbar_expect_tx(&k_full[b], 2 * kTileQ);
        tma_load_2d(&tmQKh, &k_full[b], dst, hidden + head * 64, row0 + c * 128);
        tma_load_2d(&tmQKl, &k_full[b], dst + kTileQ, hidden + head * 64, row0 + c * 128);
      }
      for (int u = 0; u < n_units; ++u) {
        const int b = u & 1;
        mbar_wait(&pv_done[b], ((u >> 1) & 1) ^ 1);
        uint8_t* dst = v_op + b * 2 * kVTile;
        mbar_expect_tx(&v_full[b], 2 * kVTile);
        tma_load_2d(&tmVh, &v_full[b], dst, 2 * hidden + head * 64, row0 + u * 64);
        tma_load_2d(&tmVl, &v_full[b], dst + kVTile, 2 * hidden + head * 64, row0 + u * 64);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer
      constexpr uint32_t idesc_s = make_idesc(kFmtBF16, 128, 128);
      constexpr uint32_t idesc_o = make_idesc(kFmtBF16, 128, 64, /*b_mn_major=*/1);
      const uint64_t q1 = make_sw128_desc(smem_u32(q_op)), q2 = make_sw128_desc(smem_u32(q_op + kTileQ));
      mbar_wait(q_full, 0);
      AT_STAMP(2);
      for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1;
        mbar_wait(&k_full[b], (c >> 1) & 1);
        tc_fence_after();
        const uint32_t kb = smem_u32(k_op + b * 2 * kTileQ);
        const uint64_t k1 = make_sw128_desc(kb), k2 = make_sw128_desc(kb + kTileQ);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t o = (uint64_t)(2 * k);
          umma_bf16(tmem_base + (uint32_t)(c * 128), q1 + o, k1 + o, idesc_s, k != 0);
          umma_bf16(tmem_base + (uint32_t)(c * 128), q2 + o, k1 + o, idesc_s, 1);
          umma_bf16(tmem_base + (uint32_t)(c * 128), q1 + o, k2 + o, idesc_s, 1);
        }
        umma_commit(&k_empty[b]);
      }
      umma_commit(s_full);
      AT_STAMP(3);
      for (int u = 0; u < n_units; ++u) {
        const int b = u & 1;
        const uint32_t ph = (u >> 1) & 1;
        mbar_wait(&v_full[b], ph);
        mbar_wait(&p_ready[b], ph);
        tc_fence_after();
        const uint32_t pb = smem_u32(p_op + b * 2 * kTileQ), vb = smem_u32(v_op + b * 2 * kVTile);
        const uint64_t p1 = make_sw128_desc(pb), p2 = make_sw128_desc(pb + kTileQ);
        const uint64_t v1 = make_sw128_desc(vb), v2 = make_sw128_desc(vb + kVTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t oa = (uint64_t)(2 * k);        // P: K-major, 16 keys = 32 B along the row
          const uint64_t ob = (uint64_t)(128 * k);      // V: MN-major, 16 keys = 16 rows x 128 B = 2048 B
          umma_bf16(tmem_base, p1 + oa, v1 + ob, idesc_o, (u | k) != 0);
          umma_bf16(tmem_base, p2 + oa, v1 + ob, idesc_o, 1);
          umma_bf16(tmem_base, p1 + oa, v2 + ob, idesc_o, 1);
        }
        umma_commit(&pv_done[b]);
      }
      umma_commit(o_full);
      AT_STAMP(4);
    }
  } else {
    // ===== softmax: two groups of 4 warps (one warp of each group per scheduler, so TMEM-load and exp latencies of one
    // group hide behind the other).  Group g owns key chunks / 64-key units with index == g (mod 2) and P buffer g.
    const int q = warp & 3, grp = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    mbar_wait(s_full, 0);
    tc_fence_after();
    if (threadIdx.x == 64) AT_STAMP(5);
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    float m = -INFINITY;
    for (int c = grp; c < n_chunks; c += 2) {
#pragma unroll 1
      for (int c0 = c * 128; c0 < c * 128 + 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + (uint32_t)c0, v);
        if (c0 + 32 <= len) {                     // CTA-uniform: interior chunks carry no key mask (no predicate chains)
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, (c0 + j < len) ? __uint_as_float(v[j]) : -INFINITY);
        }
      }
    }
    xch[grp * 128 + r] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(xch[r], xch[128 + r]);
    asm volatile("bar.sync 1, 256;" ::: "memory");               // xch is reused for the row sums below
    if (threadIdx.x == 64) AT_STAMP(6);
    float sum = 0.f;
    const float m_scaled = m * scale_log2e;       // p = 2^(s * scale - m * scale): one FFMA + one MUFU.EX2 per element
    const uint32_t xr = (uint32_t)(r & 7);
    uint8_t* p1 = p_op + grp * 2 * kTileQ + (uint32_t)r * 128u;
    for (int u = grp, it = 0; u < n_units; u += 2, ++it) {
      uint32_t vv[2][32];                         // both 32-column halves of the unit in flight, one wait
      tmem_ld32_nowait(lane_addr + (uint32_t)(u * 64), vv[0]);
      tmem_ld32_nowait(lane_addr + (uint32_t)(u * 64 + 32), vv[1]);
      tmem_wait_ld();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t (&v)[32] = vv[half];
        const int c0 = u * 64 + half * 32;
        uint32_t hi[16], lo[16];
        if (c0 + 32 <= len) {                     // CTA-uniform fast path: arguments are <= 0, ex2.approx underflows to 0
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float pa = ex2_approx(fmaf(__uint_as_float(v[2 * j]), scale_log2e, -m_scaled));
            float pb = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -m_scaled));
            sum += pa + pb;
            if (drop_thr) attn_drop_pair(seed, (uint32_t)(row0 + q0 + r), (uint32_t)(c0 + 2 * j) >> 1, (uint32_t)head, drop_thr, drop_inv_keep, pa, pb);
            split2(pa, pb, hi[j], lo[j]);
          }
        } else {                                  // chunk holding the sequence end: masked keys get exponent -inf -> p = 0
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float xa = (c0 + 2 * j < len) ? fmaf(__uint_as_float(v[2 * j]), scale_log2e, -m_scaled) : -INFINITY;
            const float xb = (c0 + 2 * j + 1 < len) ? fmaf(__uint_as_float(v[2 * j + 1]), scale_log2e, -m_scaled) : -INFINITY;
            float pa = ex2_approx(xa), pb = ex2_approx(xb);
            sum += pa + pb;
            if (drop_thr) attn_drop_pair(seed, (uint32_t)(row0 + q0 + r), (uint32_t)(c0 + 2 * j) >> 1, (uint32_t)head, drop_thr, drop_inv_keep, pa, pb);
            split2(pa, pb, hi[j], lo[j]);
          }
        }
        // the group's P buffer is free once the PV MMAs of its previous unit have read it: wait only now, after the first
        // half's exponentials are already in registers, so the wait overlaps that arithmetic
        if (half == 0 && it >= 1) mbar_wait(&pv_done[grp], (it - 1) & 1);
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          const uint32_t off = (((uint32_t)(4 * half + qd)) ^ xr) << 4;
          *reinterpret_cast<uint4*>(p1 + off) = make_uint4(hi[4 * qd], hi[4 * qd + 1], hi[4 * qd + 2], hi[4 * qd + 3]);
          *reinterpret_cast<uint4*>(p1 + kTileQ + off) = make_uint4(lo[4 * qd], lo[4 * qd + 1], lo[4 * qd + 2], lo[4 * qd + 3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&p_ready[grp]);
    }
    xch[grp * 128 + r] = sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    sum = xch[r] + xch[128 + r];
    if (lse2 && grp == 0 && q0 + r < len) lse2[(size_t)(row0 + q0 + r) * heads + head] = m_scaled + log2f(sum);
    if (threadIdx.x == 64) AT_STAMP(7);
    mbar_wait(o_full, 0);
    tc_fence_after();
    if (threadIdx.x == 64) AT_STAMP(8);
    const float inv = __fdiv_rn(1.0f, sum);
    const size_t o4 = ((size_t)(row0 + q0 + r) * hidden + head * 64 + grp * 32) >> 2;   // group g stores dims [32g, 32g+32)
    {
      uint32_t v[32];
      tmem_ld32(lane_addr + (uint32_t)(grp * 32), v);
      if (q0 + r < len) {
        if (out_plane > 0) {                      // bf16 planes: 16-byte stores (8 elements), full 32-byte sectors per thread
          uint4* hp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + o4 * 4);
          uint4* lp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + out_plane + o4 * 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 h, l;
            split2(__uint_as_float(v[8 * j]) * inv, __uint_as_float(v[8 * j + 1]) * inv, h.x, l.x);
            split2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv, h.y, l.y);
            split2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv, h.z, l.z);
            split2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv, h.w, l.w);
            hp[j] = h; lp[j] = l;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st4_fmt(out, out_plane, o4 + j,
                    make_float4(__uint_as_float(v[4 * j]) * inv, __uint_as_float(v[4 * j + 1]) * inv,
                                __uint_as_float(v[4 * j + 2]) * inv, __uint_as_float(v[4 * j + 3]) * inv));
        }
      }
    }
  }

  tc_fence_before();
  if (threadIdx.x == 64) AT_STAMP(9);
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    if (lane == 0) AT_STAMP(10);
  }
#undef AT_STAMP
}

int attention_split(const void* qkv_hi, long long plane, const int32_t* cu, int nseq, int R, int max_len, int heads,
                    int head_dim, void* out, long long out_plane, float* lse2, float p_drop, unsigned long long seed,
                    const unsigned long long* step_seed, cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  if (head_dim != 64 || max_len > 512 || !aligned16(qkv_hi) || ((plane * 2) & 15)) return VBG_EUNSUPPORTED;
  const long long ld = 3LL * heads * 64;
  const __nv_bfloat16* hi = reinterpret_cast<const __nv_bfloat16*>(qkv_hi);
  CUtensorMap mq[2], mv[2];
  for (int i = 0; i < 2; ++i) {
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box_qk[2] = {64u, 128u}, box_v[2] = {64u, 64u};
    if (!tc_encode(&mq[i], hi + i * plane, 2, dims, strides, box_qk, nullptr, true)) return VBG_EUNSUPPORTED;
    if (!tc_encode(&mv[i], hi + i * plane, 2, dims, strides, box_v, nullptr, true)) return VBG_EUNSUPPORTED;
  }
  constexpr size_t smem = 2 * kTileQ + 4 * kTileQ + 4 * kVTile + 4 * kTileQ + 1024 + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("attention_split: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
    attr = true;
  }
  dim3 grid(cdiv(max_len, 128), heads, nseq);
  uint32_t thr; float inv_keep;
  attn_drop_params(p_drop, thr, inv_keep);
  attention_split_kernel<<<grid, kAt2Threads, smem, s>>>(mq[0], mq[1], mv[0], mv[1], cu, heads, 0.125f * 1.4426950408889634f, out, out_plane,
                                                         tc_debug_timeline(), lse2, thr, inv_keep, attn_seed32(seed), step_seed);
  return check_launch("vbg_attention_split_fwd(tcgen05)");
}

int attention_tc(const float* qkv, const int32_t* cu, int nseq, int max_len, int heads, int head_dim, float* out,
                 cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  if (head_dim != 64 || max_len > 512) return VBG_EUNSUPPORTED;
  constexpr size_t smem = 2 * kTileQ + 4 * kTileQ + 4 * kVtTile + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("attention_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
    attr = true;
  }
  dim3 grid(cdiv(max_len, 128), heads, nseq);
  attention_tc_kernel<<<grid, kAtThreads, smem, s>>>(qkv, cu, heads, 0.125f * 1.4426950408889634f, out);
  return check_launch("vbg_attention_fwd(tcgen05)");
}

}  // namespace vbg
