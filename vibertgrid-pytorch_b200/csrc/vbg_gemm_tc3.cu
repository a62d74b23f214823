// VBG_PREC_BF16X3: the parity-grade tensor-core mode of the dense contractions.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T )     fp32 in HBM, three bf16 tcgen05 products, fp32 accumulate in TMEM
//
// Why: kind::tf32 ignores the low 13 mantissa bits of each fp32 operand (measured on B200: single-GEMM
// error 8e-4, whole-forward logits error 5e-3 -- outside the 1e-3 parity bar), and bf16 is worse.  Writing
// a = a1 + a2 (+ a3), a1 = bf16(a), a2 = bf16(a - a1), |a3| <= 2^-18 |a|, and likewise for w,
//
//   a*w = a1*w1 + a2*w1 + a1*w2  + O(2^-17 |a||w|)
//
// gives fp32-class products from three bf16 MMAs -- 1.5x the tensor time of one TF32 MMA, 2x cheaper than the
// 3xTF32 split -- with the fp32 accumulator of tcgen05 untouched.
//
// Data flow per CTA (one 128 x BN output tile):
//   warp 0      TMA producer: fp32 A blocks (32 floats = 128 B per row, SWIZZLE_128B) into a small landing ring;
//               pre-split bf16 weight planes W1 / W2 (64 bf16 = 128 B per row) straight into the operand ring.
//   warps 2-5   converters: one thread per tile row reads its 128-byte fp32 row from the landing slot, splits it
//               into bf16 hi / lo and writes both into the K-major SWIZZLE_128B operand tiles A1 / A2
//               (two landing blocks fill one 64-wide bf16 K block); later the same warps run the epilogue.
//   warp 1      one thread issues 3 x tcgen05.mma.kind::f16 (M=128, N=BN, K=16) per 16-wide K step.
// A may be a row-major matrix (optionally two K-concatenated sources), or an NHWC activation addressed by a
// 4-D TMA map: one filter tap per K block, conv padding by TMA out-of-bounds zero fill, stride-2 convs by the
// map's traversal stride (implicit GEMM; im2col is never materialised).
#include "vbg_tc.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace vbg {

constexpr uint32_t kAopBytes = BM * 128;       // 128 rows x 128 B: one 32-wide fp32 landing block == one 64-wide bf16 operand tile
constexpr int kTc3Threads = 320;               // warp 0 TMA, warp 1 MMA, warps 2-5 converters, warps 6-9 epilogue
constexpr uint32_t kEpiBytes = 4 * kEpiStageFloats * 4;

__device__ __forceinline__ TcTile tc3_tile(const TcParams& p, int tile, int bn) {
  const int xt = tile % p.m_tiles;             // m fastest: CTAs running together share the weight tile
  TcTile t{0, (tile / p.m_tiles) * bn, 0, 0, 0};
  if (p.conv) {
    int i = xt;
    t.w0 = (i % p.tiles_w) * p.tw; i /= p.tiles_w;
    t.h0 = (i % p.tiles_h) * p.th; i /= p.tiles_h;
    t.b0 = i * p.tb;
  } else {
    t.m0 = xt * BM;
  }
  return t;
}

// Persistent: grid = min(#tiles, #SMs); each CTA walks tiles blockIdx.x, += gridDim.x with the operand ring running
// straight through tile boundaries and two TMEM accumulators, so the epilogue of tile i overlaps the mainloop of tile i+1.
template <int BN, int STAGES>
__global__ void __launch_bounds__(kTc3Threads, 1)
gemm_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const TcParams p) {
  constexpr uint32_t B_BYTES = BN * 128;                              // one bf16 plane tile: BN rows x 64 bf16
  constexpr uint32_t STAGE_BYTES = 2 * kAopBytes + 2 * B_BYTES;       // A (fp32 landing, then A1 | A2 in place) | W1 | W2
  constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(ring + STAGES * STAGE_BYTES);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES + kEpiBytes);
  uint64_t* b_full = a_full + STAGES;
  uint64_t* a_ready = b_full + STAGES;
  uint64_t* st_empty = a_ready + STAGES;
  uint64_t* tmem_full = st_empty + STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2);
    if (p.kb_split < p.num_kb) prefetch_tmap(&tmA2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&a_full[s], 1); mbar_init(&b_full[s], 1); mbar_init(&a_ready[s], 128); mbar_init(&st_empty[s], 1);
      }
      for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer: per 64-wide K block, two fp32 A boxes (32 floats each) land in the A1 / A2 regions of the
      // stage (the converters split them in place) and the two bf16 weight planes land in W1 / W2.
      const uint32_t a_bytes = 2u * (p.conv ? (uint32_t)(p.tw * p.th * p.tb) * 128u : kAopBytes);
      int g = 0;                                                       // K blocks issued so far (runs through tiles)
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TcTile t = tc3_tile(p, tile, BN);
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&st_empty[s], ((g / STAGES) & 1) ^ 1);
          uint8_t* sa = ring + s * STAGE_BYTES;
          uint8_t* sb = sa + 2 * kAopBytes;
          mbar_expect_tx(&a_full[s], a_bytes);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int j = 2 * kb + h;
            uint8_t* dst = sa + h * kAopBytes;
            if (p.conv) {
              const int tap = j / p.cin_blocks, cb = j - tap * p.cin_blocks;
              const int fr = tap / p.kw, fs = tap - fr * p.kw;
              tma_load_4d(&tmA, &a_full[s], dst, cb * BKE, t.w0 * p.sw + fs - p.pad_w, t.h0 * p.sh + fr - p.pad_h, t.b0);
            } else if (kb < p.kb_split) {
              tma_load_2d(&tmA, &a_full[s], dst, j * BKE, t.m0);
            } else {
              tma_load_2d(&tmA2, &a_full[s], dst, (j - 2 * p.kb_split) * BKE, t.m0);
            }
          }
          mbar_expect_tx(&b_full[s], 2 * B_BYTES);
          tma_load_2d(&tmW1, &b_full[s], sb, kb * 64, t.n0);
          tma_load_2d(&tmW2, &b_full[s], sb + B_BYTES, kb * 64, t.n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: per 16-wide K step  D += A1*W1 ; D += A2*W1 ; D += A1*W2
      constexpr uint32_t idesc = make_idesc(kFmtBF16, BM, BN);
      int g = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);             // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          const uint32_t ph = (g / STAGES) & 1;
          mbar_wait(&b_full[s], ph);
          mbar_wait(&a_ready[s], ph);
          tc_fence_after();
          const uint32_t base = smem_u32(ring + s * STAGE_BYTES);
          const uint64_t a1 = make_sw128_desc(base), a2 = make_sw128_desc(base + kAopBytes);
          const uint64_t w1 = make_sw128_desc(base + 2 * kAopBytes), w2 = make_sw128_desc(base + 2 * kAopBytes + B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {            // 16 bf16 = 32 B per MMA: +2 in the (addr>>4) field
            const uint64_t o = (uint64_t)(2 * k);
            umma_bf16(d, a1 + o, w1 + o, idesc, (kb | k) != 0);
            umma_bf16(d, a2 + o, w1 + o, idesc, 1);
            umma_bf16(d, a1 + o, w2 + o, idesc, 1);
          }
          umma_commit(&st_empty[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp < 6) {
    // ===== converters (thread == tile row; each thread rewrites only its own two 128-byte rows)
    const int r = threadIdx.x - 64;
    const uint32_t xr = (uint32_t)(r & 7);
    int g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
        const int s = g % STAGES;
        mbar_wait(&a_full[s], (g / STAGES) & 1);
        uint8_t* row1 = ring + s * STAGE_BYTES + r * 128;     // fp32 K-block 0 of this row  -> bf16 hi of all 64
        uint8_t* row2 = row1 + kAopBytes;                      // fp32 K-block 1 of this row  -> bf16 lo of all 64
        float4 v[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          v[c] = *reinterpret_cast<const float4*>(row1 + (((uint32_t)c ^ xr) << 4));
          v[8 + c] = *reinterpret_cast<const float4*>(row2 + (((uint32_t)c ^ xr) << 4));
        }
#pragma unroll
        for (int qd = 0; qd < 8; ++qd) {                       // bf16 chunk qd = elements 8qd .. 8qd+7 = float4 2qd, 2qd+1
          const float f[8] = {v[2 * qd].x, v[2 * qd].y, v[2 * qd].z, v[2 * qd].w, v[2 * qd + 1].x, v[2 * qd + 1].y, v[2 * qd + 1].z, v[2 * qd + 1].w};
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 hh = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);   // .x = even element (low half)
            const __nv_bfloat162 ll = __floats2bfloat162_rn(f[2 * e] - __bfloat162float(hh.x), f[2 * e + 1] - __bfloat162float(hh.y));
            hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
            lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
          }
          const uint32_t off = (((uint32_t)qd) ^ xr) << 4;
          *reinterpret_cast<uint4*>(row1 + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(row2 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();                  // operand tiles are read by the tensor core through the async proxy
        mbar_arrive(&a_ready[s]);
      }
    }
  } else {
    // ===== epilogue warps: TMEM lanes [32q, 32q+32), q = warp & 3
    const int q = warp & 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const TcTile t = tc3_tile(p, tile, BN);
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      tc_epilogue<BN>(p, t, tmem_base + (uint32_t)(acc * BN), q, lane, epi + q * kEpiStageFloats);
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ weight splitting (one-time parameter preparation)
__global__ void split_bf16_kernel(const float* __restrict__ w, long long n, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    const float a = w[i];
    const __nv_bfloat16 h = __float2bfloat16_rn(a);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(a - __bfloat162float(h));
  }
}

// The same split, four elements per thread: one 16-byte load, two 8-byte stores (the training step re-splits every fp32
// activation and gradient that feeds a tensor-core kernel: ~340 launches per step, so this is a bandwidth kernel).
__global__ void __launch_bounds__(256) split_bf16_vec_kernel(const float4* __restrict__ w, long long n4, uint2* __restrict__ hi,
                                                             uint2* __restrict__ lo) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i + step < n4; i += 2 * step) {                    // two independent loads in flight per thread
    const float4 a = __ldg(w + i), b = __ldg(w + i + step);
    uint2 h, l;
    split4(a, h, l); hi[i] = h; lo[i] = l;
    split4(b, h, l); hi[i + step] = h; lo[i + step] = l;
  }
  if (i < n4) {
    uint2 h, l;
    split4(__ldg(w + i), h, l); hi[i] = h; lo[i] = l;
  }
}

// ------------------------------------------------------------------ host side
static bool map_w_bf16(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  return tc_encode(tm, base, 2, dims, strides, box, nullptr, true);
}
static bool map_a_f32(CUtensorMap* tm, const float* base, long long rows, long long cols, long long ld) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BKE, (cuuint32_t)BM};
  return tc_encode(tm, base, 2, dims, strides, box, nullptr, false);
}

template <int BN, int STAGES>
static int launch3(const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& w1, const CUtensorMap& w2, const TcParams& p,
                   cudaStream_t s) {
  constexpr size_t smem = (size_t)STAGES * (2 * kAopBytes + 2 * BN * 128) + kEpiBytes + 1024 + 256;
  static_assert(smem <= 232448, "bf16x3 tile does not fit the 227 KB shared-memory limit");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc3_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gemm_tc3: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
    attr = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  gemm_tc3_kernel<BN, STAGES><<<tiles < kNumSMs ? tiles : kNumSMs, kTc3Threads, smem, s>>>(a, a2, w1, w2, p);
  return check_launch("vbg_gemm(tcgen05 bf16x3)");
}

// N-tile width: the candidate with the best (wave efficiency x tile efficiency).  Wider tiles re-read A less often;
// the persistent grid makes the cost of a partial last wave explicit.
int pick_bn3(int m_tiles, int N) {
  const int cand[4] = {256, 192, 128, 64};
  int best = 64; double best_score = -1.0;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (bn > 64 && N < bn - 32) continue;                           // mostly-empty tile
    const long long tiles = (long long)m_tiles * cdiv(N, bn);
    const double waves = (double)tiles / kNumSMs;
    const double wave_eff = waves / (double)((tiles + kNumSMs - 1) / kNumSMs);
    const double fill = (double)N / ((double)cdiv(N, bn) * bn);     // useful columns in the N tiling
    const double tile_eff = (double)bn / (bn + 96.0);               // per-tile A re-read / fixed-cost model
    const double score = wave_eff * fill * tile_eff;
    if (score > best_score) { best_score = score; best = bn; }
  }
  return best;
}

static int dispatch3(const CUtensorMap& a, const CUtensorMap& a2, const void* w_hi, long long plane, int ldw, int K, TcParams& p,
                     int m_tiles, cudaStream_t s) {
  const int bn = pick_bn3(m_tiles, p.N);
  CUtensorMap w1, w2;
  const __nv_bfloat16* hi = reinterpret_cast<const __nv_bfloat16*>(w_hi);
  if (!map_w_bf16(&w1, hi, p.N, K, ldw, bn) || !map_w_bf16(&w2, hi + plane, p.N, K, ldw, bn)) return VBG_EUNSUPPORTED;
  p.m_tiles = m_tiles; p.n_tiles = cdiv(p.N, bn);
  if (bn == 256) return launch3<256, 2>(a, a2, w1, w2, p, s);
  if (bn == 192) return launch3<192, 2>(a, a2, w1, w2, p, s);
  if (bn == 128) return launch3<128, 3>(a, a2, w1, w2, p, s);
  return launch3<64, 4>(a, a2, w1, w2, p, s);
}

bool tc_conv_geometry(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, TcParams& p,
                      cuuint64_t dims[4], cuuint64_t strides_b[3], cuuint32_t box[4], cuuint32_t estr[4]);

int gemm_tc3(const float* A, int lda, const float* A2, int lda2, int K1, const void* w_split, long long plane, int ldw,
             float* C, int ldc, int M, int N, int K, const vbg_epilogue_t* ep, cudaStream_t s) {
  if (!w_split || !tc_available()) return VBG_EUNSUPPORTED;
  if (N < 64 || K % 64 || K1 % 64 || (lda & 3) || (ldw & 7) || !aligned16(A) || !aligned16(w_split) || ((plane * 2) & 15))
    return VBG_EUNSUPPORTED;
  if (K1 < K && ((lda2 & 3) || !aligned16(A2))) return VBG_EUNSUPPORTED;
  CUtensorMap ta, ta2;
  if (!map_a_f32(&ta, A, M, K1, lda)) return VBG_EUNSUPPORTED;
  if (K1 < K) { if (!map_a_f32(&ta2, A2, M, K - K1, lda2)) return VBG_EUNSUPPORTED; } else ta2 = ta;
  TcParams p{};
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.num_kb = K / 64; p.kb_split = K1 / 64; p.conv = 0;
  if (ep) p.ep = *ep;
  return dispatch3(ta, ta2, w_split, plane, ldw, K, p, cdiv(M, BM), s);
}

int conv_tc3(const float* x, int B, int H, int W, int Cin, const void* w_split, long long plane, int Cout, int kh, int kw,
             int stride, int pad, float* y, const vbg_epilogue_t* ep, cudaStream_t s) {
  if (!w_split || !tc_available()) return VBG_EUNSUPPORTED;
  if (Cin % 64 || Cout < 64 || !aligned16(x) || !aligned16(w_split) || ((plane * 2) & 15)) return VBG_EUNSUPPORTED;
  TcParams p{};
  cuuint64_t dims[4], strides[3]; cuuint32_t box[4], estr[4];
  if (!tc_conv_geometry(B, H, W, Cin, Cout, kh, kw, stride, pad, p, dims, strides, box, estr)) return VBG_EUNSUPPORTED;
  const int tiles_b = cdiv(B, p.tb);
  p.C = y;
  const int K = kh * kw * Cin;
  p.num_kb = K / 64; p.kb_split = p.num_kb;
  if (ep) p.ep = *ep;
  if (p.ep.res_mode == VBG_RES_UP2) { p.ep.out_h = p.Ho; p.ep.out_w = p.Wo; }
  if (p.ep.res_mode == VBG_RES_SAME && p.ep.ldr == 0) p.ep.ldr = Cout;
  CUtensorMap ta;
  if (!tc_encode(&ta, x, 4, dims, strides, box, estr, false)) return VBG_EUNSUPPORTED;
  return dispatch3(ta, ta, w_split, plane, K, K, p, p.tiles_w * p.tiles_h * tiles_b, s);
}

// Stem: 7x7 / stride 2 / pad 3 over a 3-channel image, as a tensor-core GEMM with K = 8 filter rows x (8 px x 4 ch).
// Input is the zero-bordered NHWC4 batch written by vbg_normalize_resize_pad: [B, H+6, W+6, 4].  For output pixel
// (ho, wo) and filter row r, the 7 taps x 4 channels are 28 CONTIGUOUS floats starting at padded pixel (2ho + r, 2wo);
// a 32-float window (the 8th pixel meets zero weights) is one 128-byte TMA row.  Consecutive wo windows overlap by
// 24 floats, which a tensor map expresses directly: dim1 = wo with a 32-byte stride.  Weights: [Cout][8][8][4] with
// zeros at r = 7, px = 7, ch = 3 (vbg_stem_pack_weights).
int stem_tc3(const float* x4, int B, int H, int W, const void* w_split, long long plane, int Cout, float* y,
             const vbg_epilogue_t* ep, cudaStream_t s) {
  if (!w_split || !tc_available()) return VBG_EUNSUPPORTED;
  if ((H & 1) || (W & 1) || Cout < 64 || !aligned16(x4) || !aligned16(w_split) || ((plane * 2) & 15)) return VBG_EUNSUPPORTED;
  const int Ho = H / 2, Wo = W / 2, Hp = H + 6, Wp = W + 6;
  TcParams p{};
  p.conv = 1; p.Ho = Ho; p.Wo = Wo; p.Bn = B; p.kw = 1; p.cin_blocks = 1; p.pad_w = p.pad_h = 0; p.sw = 1; p.sh = 2;
  p.tw = Wo < BM ? Wo : BM;
  p.th = (BM / p.tw) < Ho ? (BM / p.tw) : Ho;
  p.tb = (p.th == Ho && p.tw == Wo) ? ((BM / (p.tw * p.th)) < B ? (BM / (p.tw * p.th)) : B) : 1;
  if (p.th * 2 > 256) return VBG_EUNSUPPORTED;
  p.tiles_w = cdiv(Wo, p.tw); p.tiles_h = cdiv(Ho, p.th);
  p.M = B * Ho * Wo; p.N = Cout; p.ldc = Cout; p.C = y;
  p.num_kb = 4; p.kb_split = 4;                  // K = 8 rows x 32 floats = 256 = 4 blocks of 64
  if (ep) p.ep = *ep;
  if (p.ep.res_mode == VBG_RES_SAME && p.ep.ldr == 0) p.ep.ldr = Cout;
  cuuint64_t dims[4] = {32, (cuuint64_t)Wo, (cuuint64_t)Hp, (cuuint64_t)B};
  cuuint64_t strides[3] = {32, (cuuint64_t)Wp * 16, (cuuint64_t)Hp * Wp * 16};
  cuuint32_t box[4] = {32, (cuuint32_t)p.tw, (cuuint32_t)(p.th * 2), (cuuint32_t)p.tb};
  cuuint32_t estr[4] = {1, 1, 2, 1};
  CUtensorMap ta;
  if (!tc_encode(&ta, x4, 4, dims, strides, box, estr, false)) return VBG_EUNSUPPORTED;
  return dispatch3(ta, ta, w_split, plane, 256, 256, p, p.tiles_w * p.tiles_h * cdiv(B, p.tb), s);
}

int split_bf16(const float* w, long long n, void* hi, void* lo, cudaStream_t s) {
  if (n == 0) return VBG_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if ((n & 3) == 0 && aligned16(w) && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo) & 7) == 0) {
    const long long n4 = n / 4;
    int b4 = (int)((n4 + 511) / 512);                        // two float4 per thread
    if (b4 > kNumSMs * 8) b4 = kNumSMs * 8;
    if (b4 < 1) b4 = 1;
    split_bf16_vec_kernel<<<b4, 256, 0, s>>>(reinterpret_cast<const float4*>(w), n4, reinterpret_cast<uint2*>(hi), reinterpret_cast<uint2*>(lo));
    return check_launch("vbg_split_bf16");
  }
  split_bf16_kernel<<<blocks, 256, 0, s>>>(w, n, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo));
  return check_launch("vbg_split_bf16");
}

}  // namespace vbg
