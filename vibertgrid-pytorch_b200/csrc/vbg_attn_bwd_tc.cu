// Self-attention backward on the 5th-generation tensor cores (tcgen05.mma kind::f16, bf16x3 products, fp32 accumulate in TMEM).
//
// The backward of HF BertSelfAttention inside the reference's BERT (model/BERTgrid_generator.py:134, reached by loss.backward()
// at pipeline/train_val_utils.py:277), over the packed varlen batch, head dimension 64, sequences of at most 512 rows:
//
//   S = Q K^T / 8,  P = softmax(S),  Pd = dropout(P),  O = Pd V                          (forward, vbg_attn_tc.cu)
//   delta_q = sum_d dO[q,d] O[q,d]
//   dPd = dO V^T,  dP = dPd o keep / (1 - p),  dS = P o (dP - delta) / 8
//   dV = Pd^T dO,  dK = dS^T Q,  dQ = dS K
//
// P is rebuilt from Q, K and the forward's base-2 row log-sum-exp (nothing of size L x L is stored); the dropout mask is the
// forward's counter hash (vbg_tc.cuh::attn_drop_bits).  ONE kernel template does both halves of the backward:
//
//   kKV = true   CTA = (128 KEYS j, head, sequence); streams 64-query tiles i.       thread == key row == TMEM lane
//        X  = K_j Q_i^T   (= S^T tile),   Y = V_j dO_i^T (= dPd^T tile)              A, B K-major (row = token, 128 B of head dims)
//        dV_j += Pd^T . dO_i,   dK_j += dS^T . Q_i                                    A = the tile written below (K-major: row =
//                                                                                    key, K = query), B = dO_i / Q_i MN-major
//   kKV = false  CTA = (128 QUERIES i, head, sequence); streams 64-key tiles j.      thread == query row == TMEM lane
//        X  = Q_i K_j^T   (= S tile),     Y = dO_i V_j^T (= dPd tile)
//        dQ_i += dS . K_j                                                            B = K_j MN-major
//
// so every operand flavour is one the forward kernel already uses: TMA drops [64 rows x 64 dims] boxes of the bf16 hi / lo
// planes straight into SWIZZLE_128B tiles; a streamed tile serves first as the K-major B operand of X / Y and then, untouched,
// as the MN-major B operand of the accumulating products (like V in the forward).  Roles: warp 0 TMA producer, warp 1 MMA
// issuer, warps 2-9 = 256 elementwise threads (two groups of four warps; group g owns columns [32g, 32g + 32) of every X / Y
// tile).  X / Y are double-buffered in TMEM (4 x 64 columns) so the products of tile t + 1 run while the threads turn tile t
// into Pd / dS; the accumulators (dV, dK or dQ: 64 columns each) stay in TMEM for the whole CTA.  Deterministic: no atomics.
#include "vbg_tc.cuh"
#include <cuda_bf16.h>

namespace vbg {

constexpr int kAbThreads = 320;
constexpr uint32_t kAbRT = 128 * 128;      // resident operand tile: 128 rows x 64 bf16
constexpr uint32_t kAbTT = 64 * 128;       // streamed operand tile:  64 rows x 64 bf16
__device__ __forceinline__ float ab_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// delta[row, head] = sum_d dO[row, head*64 + d] * O[row, head*64 + d]      (one warp per row; a half-warp per head pair step)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const float* __restrict__ o, const float* __restrict__ d_o, long long rows, int heads, float* __restrict__ delta) {
  const int lane = threadIdx.x & 31;
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float4* o4 = reinterpret_cast<const float4*>(o + r * heads * 64);
  const float4* d4 = reinterpret_cast<const float4*>(d_o + r * heads * 64);
  for (int c0 = 0; c0 < heads * 16; c0 += 32) {          // float4 chunk c belongs to head c / 16: lanes 0-15 and 16-31 hold two heads
    const int c = c0 + lane;
    float s = 0.f;
    if (c < heads * 16) {
      const float4 a = __ldg(o4 + c), b = __ldg(d4 + c);
      s = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((lane & 15) == 0 && c < heads * 16) delta[r * heads + (c >> 4)] = s;
  }
}

template <bool kKV>
__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl,
                   const __grid_constant__ CUtensorMap tmDh, const __grid_constant__ CUtensorMap tmDl,
                   const int32_t* __restrict__ cu, int heads, float scale, float scale_log2e, const float* __restrict__ lse2,
                   const float* __restrict__ delta, float* __restrict__ dqkv, uint32_t drop_thr, float drop_inv_keep, uint32_t seed,
                   const unsigned long long* __restrict__ step_seed) {
  if (drop_thr && step_seed) seed = attn_fold_step(seed, __ldg(step_seed));
  const int seq = blockIdx.z, head = blockIdx.y, r0 = blockIdx.x * 128;
  const int row0 = cu[seq], len = cu[seq + 1] - row0;
  if (r0 >= len) return;                                        // whole CTA, before any barrier / allocation
  const int n_t = (len + 63) >> 6;
  const int hidden = heads * 64;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* r_op = smem;                                         // R1h | R1l | R2h | R2l            64 KB
  uint8_t* t_op = r_op + 4 * kAbRT;                             // 2 x (T1h | T1l | T2h | T2l)      64 KB
  uint8_t* e_op = t_op + 8 * kAbTT;                             // Pdh | Pdl | dSh | dSl            64 KB
  float* stats = reinterpret_cast<float*>(e_op + 4 * kAbRT);    // [2][2][64]: lse2 / delta of the streamed queries (kKV)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stats + 256);
  uint64_t *r_full = bars, *t_full = bars + 1, *t_empty = bars + 3, *xy_full = bars + 5, *e_ready = bars + 7, *e_free = bars + 8,
           *acc_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmQh); prefetch_tmap(&tmQl); prefetch_tmap(&tmDh); prefetch_tmap(&tmDl); }
  if (warp == 1) {
    if (lane == 0) {
      mbar_init(r_full, 1); mbar_init(e_ready, 256); mbar_init(e_free, 1); mbar_init(acc_full, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 1); mbar_init(&xy_full[i], 1); }
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
  // TMEM columns: X[b] = 64 b, Y[b] = 128 + 64 b (b = 0, 1), accumulators at 256 (dV) and 320 (dK | dQ)
  constexpr uint32_t kColY = 128, kColA1 = 256, kColA2 = 320;
  // column of each operand inside the packed [R, 3 * hidden] QKV planes / the [R, hidden] dO planes
  const int colQ = head * 64, colK = hidden + head * 64, colV = 2 * hidden + head * 64, colD = head * 64;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: the resident 128-row tiles as two 64-row boxes each, then the streamed tiles
      mbar_expect_tx(r_full, 4 * kAbRT);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = row0 + r0 + 64 * h;
        uint8_t* d = r_op + h * kAbTT;
        if (kKV) {
          tma_load_2d(&tmQh, r_full, d, colK, row);              tma_load_2d(&tmQl, r_full, d + kAbRT, colK, row);
          tma_load_2d(&tmQh, r_full, d + 2 * kAbRT, colV, row);  tma_load_2d(&tmQl, r_full, d + 3 * kAbRT, colV, row);
        } else {
          tma_load_2d(&tmQh, r_full, d, colQ, row);              tma_load_2d(&tmQl, r_full, d + kAbRT, colQ, row);
          tma_load_2d(&tmDh, r_full, d + 2 * kAbRT, colD, row);  tma_load_2d(&tmDl, r_full, d + 3 * kAbRT, colD, row);
        }
      }
      for (int t = 0; t < n_t; ++t) {
        const int b = t & 1;
        mbar_wait(&t_empty[b], ((t >> 1) & 1) ^ 1);
        uint8_t* d = t_op + b * 4 * kAbTT;
        const int row = row0 + t * 64;
        mbar_expect_tx(&t_full[b], 4 * kAbTT);
        if (kKV) {
          tma_load_2d(&tmQh, &t_full[b], d, colQ, row);              tma_load_2d(&tmQl, &t_full[b], d + kAbTT, colQ, row);
          tma_load_2d(&tmDh, &t_full[b], d + 2 * kAbTT, colD, row);  tma_load_2d(&tmDl, &t_full[b], d + 3 * kAbTT, colD, row);
        } else {
          tma_load_2d(&tmQh, &t_full[b], d, colK, row);              tma_load_2d(&tmQl, &t_full[b], d + kAbTT, colK, row);
          tma_load_2d(&tmQh, &t_full[b], d + 2 * kAbTT, colV, row);  tma_load_2d(&tmQl, &t_full[b], d + 3 * kAbTT, colV, row);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer
      constexpr uint32_t idesc_xy = make_idesc(kFmtBF16, 128, 64);
      constexpr uint32_t idesc_acc = make_idesc(kFmtBF16, 128, 64, /*b_mn_major=*/1);
      const uint64_t r1h = make_sw128_desc(smem_u32(r_op)), r1l = make_sw128_desc(smem_u32(r_op + kAbRT));
      const uint64_t r2h = make_sw128_desc(smem_u32(r_op + 2 * kAbRT)), r2l = make_sw128_desc(smem_u32(r_op + 3 * kAbRT));
      const uint64_t pdh = make_sw128_desc(smem_u32(e_op)), pdl = make_sw128_desc(smem_u32(e_op + kAbRT));
      const uint64_t dsh = make_sw128_desc(smem_u32(e_op + 2 * kAbRT)), dsl = make_sw128_desc(smem_u32(e_op + 3 * kAbRT));
      mbar_wait(r_full, 0);
      for (int t = 0; t <= n_t; ++t) {
        if (t < n_t) {
          const int b = t & 1;
          mbar_wait(&t_full[b], (t >> 1) & 1);
          tc_fence_after();
          const uint32_t tb = smem_u32(t_op + b * 4 * kAbTT);
          const uint64_t t1h = make_sw128_desc(tb), t1l = make_sw128_desc(tb + kAbTT);
          const uint64_t t2h = make_sw128_desc(tb + 2 * kAbTT), t2l = make_sw128_desc(tb + 3 * kAbTT);
          const uint32_t dx = tmem_base + (uint32_t)(64 * b), dy = tmem_base + kColY + (uint32_t)(64 * b);
#pragma unroll
          for (int k = 0; k < 4; ++k) {                // X = R1 . T1^T over the 64 head dims, 16 per step
            const uint64_t o = (uint64_t)(2 * k);
            umma_bf16(dx, r1h + o, t1h + o, idesc_xy, k != 0);
            umma_bf16(dx, r1l + o, t1h + o, idesc_xy, 1);
            umma_bf16(dx, r1h + o, t1l + o, idesc_xy, 1);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {                // Y = R2 . T2^T
            const uint64_t o = (uint64_t)(2 * k);
            umma_bf16(dy, r2h + o, t2h + o, idesc_xy, k != 0);
            umma_bf16(dy, r2l + o, t2h + o, idesc_xy, 1);
            umma_bf16(dy, r2h + o, t2l + o, idesc_xy, 1);
          }
          umma_commit(&xy_full[b]);
        }
        if (t >= 1) {
          const int u = t - 1, bu = u & 1;
          mbar_wait(e_ready, u & 1);
          tc_fence_after();
          const uint32_t tb = smem_u32(t_op + bu * 4 * kAbTT);
          const uint64_t t1h = make_sw128_desc(tb), t1l = make_sw128_desc(tb + kAbTT);
          const uint64_t t2h = make_sw128_desc(tb + 2 * kAbTT), t2l = make_sw128_desc(tb + 3 * kAbTT);
#pragma unroll
          for (int k = 0; k < 4; ++k) {                // reduction over the 64 streamed rows, 16 per step
            const uint64_t oa = (uint64_t)(2 * k);     // A (Pd / dS tile, K-major): 16 columns = 32 B along the row
            const uint64_t ob = (uint64_t)(128 * k);   // B (streamed tile, MN-major): 16 rows x 128 B = 2048 B
            if (kKV) {
              umma_bf16(tmem_base + kColA1, pdh + oa, t2h + ob, idesc_acc, (u | k) != 0);
              umma_bf16(tmem_base + kColA1, pdl + oa, t2h + ob, idesc_acc, 1);
              umma_bf16(tmem_base + kColA1, pdh + oa, t2l + ob, idesc_acc, 1);
            }
            umma_bf16(tmem_base + kColA2, dsh + oa, t1h + ob, idesc_acc, (u | k) != 0);
            umma_bf16(tmem_base + kColA2, dsl + oa, t1h + ob, idesc_acc, 1);
            umma_bf16(tmem_base + kColA2, dsh + oa, t1l + ob, idesc_acc, 1);
          }
          umma_commit(&t_empty[bu]);                   // the streamed stage may be refilled ...
          umma_commit(e_free);                         // ... and the Pd / dS tiles rewritten
        }
      }
      umma_commit(acc_full);
    }
  } else {
    // ===== elementwise: thread == resident row == TMEM lane; group g owns columns [32 g, 32 g + 32) of every X / Y tile
    const int q = warp & 3, g = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const int et = threadIdx.x - 64;                               // 0..255
    const int my = r0 + r;                                         // resident row inside the sequence
    const bool my_ok = my < len;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t xr = (uint32_t)(r & 7);
    float lse_r = 0.f, del_r = 0.f;
    if (!kKV && my_ok) {
      lse_r = __ldg(lse2 + (size_t)(row0 + my) * heads + head);
      del_r = __ldg(delta + (size_t)(row0 + my) * heads + head);
    }
    uint8_t* pd_row = e_op + (uint32_t)r * 128u;
    uint8_t* ds_row = e_op + 2 * kAbRT + (uint32_t)r * 128u;
    for (int t = 0; t < n_t; ++t) {
      const int b = t & 1;
      const int c_base = t * 64 + 32 * g;                          // first streamed row (column) this thread handles
      if (kKV) {
        // per-column statistics of the streamed queries -> shared memory (one bar.sync of the 256 threads per tile)
        if (et < 128) {
          const int qq = t * 64 + (et & 63);
          const float* src = et < 64 ? lse2 : delta;
          stats[b * 128 + et] = qq < len ? __ldg(src + (size_t)(row0 + qq) * heads + head) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(&xy_full[b], (t >> 1) & 1);
      tc_fence_after();
      uint32_t xv[32], yv[32];
      tmem_ld32_nowait(lane_addr + (uint32_t)(64 * b + 32 * g), xv);
      tmem_ld32_nowait(lane_addr + kColY + (uint32_t)(64 * b + 32 * g), yv);
      tmem_wait_ld();
      uint32_t ph[16], pl[16], dh[16], dl[16];
      const float* st_l = stats + b * 128 + 32 * g;
      const float* st_d = st_l + 64;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int ca = c_base + 2 * j, cb = ca + 1;                // streamed indices of this pair
        float la, lb, da, db;
        if (kKV) { la = st_l[2 * j]; lb = st_l[2 * j + 1]; da = st_d[2 * j]; db = st_d[2 * j + 1]; }
        else { la = lb = lse_r; da = db = del_r; }
        const bool va = my_ok && ca < len, vb = my_ok && cb < len;
        float pa = va ? ab_ex2(fmaf(__uint_as_float(xv[2 * j]), scale_log2e, -la)) : 0.f;
        float pb = vb ? ab_ex2(fmaf(__uint_as_float(xv[2 * j + 1]), scale_log2e, -lb)) : 0.f;
        float ya = __uint_as_float(yv[2 * j]), yb = __uint_as_float(yv[2 * j + 1]);
        float pda = pa, pdb = pb;
        if (drop_thr) {
          if (kKV) {     // thread == key `my`; the streamed index is the query: one hash per element, the key's half of the pair word
            const uint32_t wa = attn_drop_bits(seed, (uint32_t)(row0 + ca), (uint32_t)my >> 1, (uint32_t)head);
            const uint32_t wb = attn_drop_bits(seed, (uint32_t)(row0 + cb), (uint32_t)my >> 1, (uint32_t)head);
            const uint32_t ka = (my & 1) ? (wa >> 16) : (wa & 0xffffu), kb = (my & 1) ? (wb >> 16) : (wb & 0xffffu);
            const float fa = ka >= drop_thr ? drop_inv_keep : 0.f, fb = kb >= drop_thr ? drop_inv_keep : 0.f;
            pda = pa * fa; pdb = pb * fb; ya *= fa; yb *= fb;
          } else {       // thread == query `my`; the streamed pair (ca, cb) is one key pair: one hash
            const uint32_t w = attn_drop_bits(seed, (uint32_t)(row0 + my), (uint32_t)ca >> 1, (uint32_t)head);
            const float fa = (w & 0xffffu) >= drop_thr ? drop_inv_keep : 0.f, fb = (w >> 16) >= drop_thr ? drop_inv_keep : 0.f;
            pda = pa * fa; pdb = pb * fb; ya *= fa; yb *= fb;
          }
        }
        const float dsa = pa * (ya - da) * scale, dsb = pb * (yb - db) * scale;
        if (kKV) split2(pda, pdb, ph[j], pl[j]);
        split2(dsa, dsb, dh[j], dl[j]);
      }
      if (t >= 1) mbar_wait(e_free, (t - 1) & 1);                  // the products of the previous tile have read Pd / dS
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        const uint32_t off = (((uint32_t)(4 * g + qd)) ^ xr) << 4;
        if (kKV) {
          *reinterpret_cast<uint4*>(pd_row + off) = make_uint4(ph[4 * qd], ph[4 * qd + 1], ph[4 * qd + 2], ph[4 * qd + 3]);
          *reinterpret_cast<uint4*>(pd_row + kAbRT + off) = make_uint4(pl[4 * qd], pl[4 * qd + 1], pl[4 * qd + 2], pl[4 * qd + 3]);
        }
        *reinterpret_cast<uint4*>(ds_row + off) = make_uint4(dh[4 * qd], dh[4 * qd + 1], dh[4 * qd + 2], dh[4 * qd + 3]);
        *reinterpret_cast<uint4*>(ds_row + kAbRT + off) = make_uint4(dl[4 * qd], dl[4 * qd + 1], dl[4 * qd + 2], dl[4 * qd + 3]);
      }
      tc_fence_before();                                           // X / Y of this tile are consumed before they are overwritten
      fence_proxy_async_smem();
      mbar_arrive(e_ready);
    }
    // ---- accumulators -> dqkv (fp32 rows; group g stores head dims [32 g, 32 g + 32))
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float* orow = dqkv + (size_t)(row0 + my) * (3 * hidden) + 32 * g;
    if (kKV) {
      uint32_t v[32];
      tmem_ld32(lane_addr + kColA1 + (uint32_t)(32 * g), v);
      if (my_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(orow + colV + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
    {
      uint32_t v[32];
      tmem_ld32(lane_addr + kColA2 + (uint32_t)(32 * g), v);
      if (my_ok) {
        float* dst = orow + (kKV ? colK : colQ);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                                 __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int attention_bwd_tc(const void* qkv_hi, long long qkv_plane, const void* do_hi, long long do_plane, const float* o, const float* d_o,
                     const float* lse2, const int32_t* cu, int nseq, int R, int max_len, int heads, float p_drop,
                     unsigned long long seed, const unsigned long long* step_seed, float* dqkv, float* delta, cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  if (max_len > 512 || !aligned16(qkv_hi) || !aligned16(do_hi) || ((qkv_plane * 2) & 15) || ((do_plane * 2) & 15)) return VBG_EUNSUPPORTED;
  const long long ld3 = 3LL * heads * 64, ld1 = 1LL * heads * 64;
  CUtensorMap mq[2], md[2];
  for (int i = 0; i < 2; ++i) {
    cuuint32_t box[2] = {64u, 64u};
    cuuint64_t dq[2] = {(cuuint64_t)ld3, (cuuint64_t)R}, sq[1] = {(cuuint64_t)ld3 * 2};
    cuuint64_t dd[2] = {(cuuint64_t)ld1, (cuuint64_t)R}, sd[1] = {(cuuint64_t)ld1 * 2};
    if (!tc_encode(&mq[i], reinterpret_cast<const __nv_bfloat16*>(qkv_hi) + i * qkv_plane, 2, dq, sq, box, nullptr, true)) return VBG_EUNSUPPORTED;
    if (!tc_encode(&md[i], reinterpret_cast<const __nv_bfloat16*>(do_hi) + i * do_plane, 2, dd, sd, box, nullptr, true)) return VBG_EUNSUPPORTED;
  }
  attn_delta_kernel<<<cdiv(R, 8), 256, 0, s>>>(o, d_o, R, heads, delta);
  int rc = check_launch("vbg_attention_bwd_tc(delta)");
  if (rc) return rc;
  constexpr size_t smem = 4 * kAbRT + 8 * kAbTT + 4 * kAbRT + 1024 + 256 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attention_bwd_tc: smem opt-in failed"); return VBG_ECUDA; }
    attr = true;
  }
  uint32_t thr; float inv_keep;
  attn_drop_params(p_drop, thr, inv_keep);
  const uint32_t seed32 = attn_seed32(seed);
  const float scale = 0.125f;
  dim3 grid(cdiv(max_len, 128), heads, nseq);
  attn_bwd_tc_kernel<true><<<grid, kAbThreads, smem, s>>>(mq[0], mq[1], md[0], md[1], cu, heads, scale, scale * 1.4426950408889634f, lse2, delta,
                                                          dqkv, thr, inv_keep, seed32, step_seed);
  rc = check_launch("vbg_attention_bwd_tc(dK, dV)");
  if (rc) return rc;
  attn_bwd_tc_kernel<false><<<grid, kAbThreads, smem, s>>>(mq[0], mq[1], md[0], md[1], cu, heads, scale, scale * 1.4426950408889634f, lse2, delta,
                                                           dqkv, thr, inv_keep, seed32, step_seed);
  return check_launch("vbg_attention_bwd_tc(dQ)");
}

// keep mask of the attention dropout for one (sequence, head): mask[q, k] = 1 / 0 -- tests build the exact reference with it
__global__ void attn_drop_mask_kernel(uint32_t seed, int row0, int len, int head, uint32_t thr, float* __restrict__ mask) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= len * len) return;
  const int qi = idx / len, k = idx - qi * len;
  const uint32_t w = attn_drop_bits(seed, (uint32_t)(row0 + qi), (uint32_t)k >> 1, (uint32_t)head);
  mask[idx] = ((k & 1) ? (w >> 16) : (w & 0xffffu)) >= thr ? 1.f : 0.f;
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_attention_bwd_tc(const void* qkv_hi, long long qkv_plane, const void* do_hi, long long do_plane, const float* out,
                                    const float* d_out, const float* lse2, const int32_t* cu, int nseq, int R, int max_len, int heads,
                                    int head_dim, float p_drop, unsigned long long seed, const unsigned long long* step_seed, float* dqkv,
                                    float* workspace, size_t ws_bytes, vbg_stream_t stream) {
  VBG_REQUIRE(qkv_hi && do_hi && out && d_out && lse2 && cu && dqkv && workspace && nseq > 0 && R > 0 && max_len > 0 && heads > 0 &&
                  qkv_plane > 0 && do_plane > 0 && aligned16(out) && aligned16(d_out) && aligned16(dqkv),
              "vbg_attention_bwd_tc: bad arguments");
  VBG_REQUIRE(head_dim == 64, "vbg_attention_bwd_tc: head dimension 64 only");
  VBG_REQUIRE(p_drop >= 0.f && p_drop < 1.f, "vbg_attention_bwd_tc: 0 <= p_drop < 1");
  if ((size_t)R * heads * sizeof(float) > ws_bytes) {
    set_error("vbg_attention_bwd_tc: workspace of %zu bytes needed", (size_t)R * heads * sizeof(float));
    return VBG_EWORKSPACE;
  }
  int rc = attention_bwd_tc(qkv_hi, qkv_plane, do_hi, do_plane, out, d_out, lse2, cu, nseq, R, max_len, heads, p_drop, seed, step_seed, dqkv,
                            workspace, as_stream(stream));
  if (rc == VBG_EUNSUPPORTED) set_error("vbg_attention_bwd_tc: needs sm_100a, max_len <= 512, 16-byte aligned planes (max_len %d)", max_len);
  return rc;
}

extern "C" int vbg_attention_dropout_mask(unsigned long long seed, unsigned long long step_seed, int has_step_seed, float p_drop, int row0,
                                          int len, int head, float* mask, float* inv_keep, vbg_stream_t stream) {
  VBG_REQUIRE(mask && inv_keep && len > 0 && p_drop >= 0.f && p_drop < 1.f, "vbg_attention_dropout_mask: bad arguments");
  uint32_t thr; float ik;
  attn_drop_params(p_drop, thr, ik);
  *inv_keep = ik;
  uint32_t s32 = attn_seed32(seed);
  if (has_step_seed) s32 = attn_fold_step(s32, step_seed);
  attn_drop_mask_kernel<<<cdiv((long long)len * len, 256), 256, 0, as_stream(stream)>>>(s32, row0, len, head, thr, mask);
  return check_launch("vbg_attention_dropout_mask");
}
