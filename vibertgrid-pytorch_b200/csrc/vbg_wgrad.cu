// Weight gradient of a linear layer on the tensor cores, no transposes:
//
//   dW[N, K] = dY[M, N]^T . X[M, K]          (reduction over the M rows; bf16x3: dY1 X1 + dY2 X1 + dY1 X2, fp32 accumulate)
//
// Both operands are stored with the REDUCTION index as the slow one (dY rows / X rows are the GEMM's K), i.e. they are
// "MN-major" in tcgen05 terms: a TMA box of 64 rows x 64 columns lands as 64 rows of 128 bytes (SWIZZLE_128B) and is used as is
// through an MN-major shared-memory descriptor (LBO = the 8 KB between 64-column blocks, SBO = the 1 KB between 8-row groups) with
// the a_major / b_major bits of the instruction descriptor set -- the layout the attention kernel already uses for V.  The
// forward's transposed-operand route (vbg_transpose_split twice + the forward GEMM) costs two extra passes over dY and X.
//
// Few output tiles (N/128 x K/BN) and a long reduction: the row range is split over `splits` CTAs per tile, partial tiles go to a
// workspace and are summed in a fixed order by splitk_finish_kernel (deterministic).
//   warp 0 TMA producer | warp 1 MMA issuer | warps 2-5 epilogue (vbg_tc.cuh::tc_epilogue)
#include "vbg_tc.cuh"

namespace vbg {

constexpr int kWgThreads = 192;
constexpr uint32_t kWgBox = 64 * 128;                 // one 64-row x 64-column bf16 box
constexpr uint32_t kWgEpiBytes = 4 * kEpiStageFloats * 4;

// MN-major SWIZZLE_128B descriptor: 64-element (128-byte) column blocks `lbo_bytes` apart, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)64 << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

struct WgParams {
  TcParams tc;          // epilogue view: C = out or workspace, ldc, M = N (rows of dW), N = K (columns of dW)
  int rows;             // reduction length (rows of dY / X); conv: number of 64-pixel blocks * 64
  int blocks_per_split; // 64-row blocks per split
  int n_tiles, k_tiles; // output tiles
  // conv (conv == 1): a 64-row block is tb images x th rows x tw columns of OUTPUT pixels (tw*th*tb == 64); the X operand of
  // output-column tile kt is the input window of filter tap kt / cin_tiles shifted by (r - pad, s - pad), stride `st`
  int conv, tw, th, tb, tiles_w, tiles_h, kw, cin_tiles, st, pad;
  int box_rows;         // rows one TMA box delivers (64, or tw*th*tb < 64: the remaining rows of a stage are zeroed once and never written)
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, const WgParams wp) {
  constexpr uint32_t A_PLANE = 2 * kWgBox;                        // 128 dW rows = two 64-column boxes of dY
  constexpr uint32_t B_PLANE = (BN / 64) * kWgBox;
  constexpr uint32_t STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;     // Y1 | Y2 | X1 | X2
  constexpr uint32_t TMEM_COLS = BN <= 128 ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(ring + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES + kWgEpiBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = wp.n_tiles * wp.k_tiles;
  const int tile = blockIdx.x % tiles, split = blockIdx.x / tiles;
  const int n0 = (tile % wp.n_tiles) * 128, k0 = (tile / wp.n_tiles) * BN;
  const int n_blocks = (wp.rows + 63) >> 6;
  const int b0 = split * wp.blocks_per_split, b1 = min(b0 + wp.blocks_per_split, n_blocks);

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmY); prefetch_tmap(&tmX); }
  if (wp.box_rows < 64) {                          // short boxes: rows past the box stay zero for the whole launch
    uint4* z = reinterpret_cast<uint4*>(ring);
    for (uint32_t i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += kWgThreads) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(acc_full, 1);
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
      for (int b = b0, g = 0; b < b1; ++b, ++g) {
        const int s = g % STAGES;
        mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
        uint8_t* sy = ring + s * STAGE_BYTES;
        uint8_t* sx = sy + 2 * A_PLANE;
        mbar_expect_tx(&full[s], (STAGE_BYTES / 64) * (uint32_t)wp.box_rows);
        if (wp.conv) {
          int i = b;
          const int w0 = (i % wp.tiles_w) * wp.tw; i /= wp.tiles_w;
          const int h0 = (i % wp.tiles_h) * wp.th; const int bb = (i / wp.tiles_h) * wp.tb;
          const int kt = k0 / BN, tap = kt / wp.cin_tiles, c0 = (kt - tap * wp.cin_tiles) * BN;
          const int fr = tap / wp.kw, fs = tap - fr * wp.kw;
          const int xw = w0 * wp.st + fs - wp.pad, xh = h0 * wp.st + fr - wp.pad;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_5d(&tmY, &full[s], sy + pl * A_PLANE + j * kWgBox, n0 + 64 * j, w0, h0, bb, pl);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_5d(&tmX, &full[s], sx + pl * B_PLANE + j * kWgBox, c0 + 64 * j, xw, xh, bb, pl);
          }
        } else {
#pragma unroll
          for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_3d(&tmY, &full[s], sy + pl * A_PLANE + j * kWgBox, n0 + 64 * j, b * 64, pl);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_3d(&tmX, &full[s], sx + pl * B_PLANE + j * kWgBox, k0 + 64 * j, b * 64, pl);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kFmtBF16, 128, BN, 1) | (1u << 15);        // A and B MN-major
      for (int b = b0, g = 0; b < b1; ++b, ++g) {
        const int s = g % STAGES;
        mbar_wait(&full[s], (g / STAGES) & 1);
        tc_fence_after();
        const uint32_t base = smem_u32(ring + s * STAGE_BYTES);
        const uint64_t y1 = make_sw128_mn_desc(base, kWgBox), y2 = make_sw128_mn_desc(base + A_PLANE, kWgBox);
        const uint64_t x1 = make_sw128_mn_desc(base + 2 * A_PLANE, kWgBox), x2 = make_sw128_mn_desc(base + 2 * A_PLANE + B_PLANE, kWgBox);
#pragma unroll
        for (int k = 0; k < 4; ++k) {              // 16 reduction rows = 16 x 128 B = 2048 B: +128 in the (addr >> 4) field
          const uint64_t o = (uint64_t)(128 * k);
          umma_bf16(tmem_base, y1 + o, x1 + o, idesc, (b != b0) || (k != 0));
          umma_bf16(tmem_base, y2 + o, x1 + o, idesc, 1);
          umma_bf16(tmem_base, y1 + o, x2 + o, idesc, 1);
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    TcTile t{n0, k0, 0, 0, 0, 0, 0, (long long)split * wp.tc.M};
    if (b1 > b0) tc_epilogue<BN>(wp.tc, t, tmem_base, q, lane, epi + q * kEpiStageFloats);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

int splitk_finish(const float* ws, int splits, long long M, int N, const vbg_epilogue_t& ep, void* C, int ldc, cudaStream_t s);   // vbg_gemm_ps.cu

static bool map_rows(CUtensorMap* tm, const void* hi, long long plane, long long rows, long long cols, long long ld) {
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
  cuuint32_t box[3] = {64u, 64u, 1u};
  return tc_encode(tm, hi, 3, dims, strides, box, nullptr, true);
}

template <int BN, int STAGES>
static int launch_wg(const CUtensorMap& ty, const CUtensorMap& tx, const WgParams& wp, int ctas, cudaStream_t s) {
  constexpr size_t smem = (size_t)STAGES * (2 * 2 * kWgBox + 2 * (BN / 64) * kWgBox) + kWgEpiBytes + 1024 + 256;
  static_assert(smem <= 232448, "wgrad tile does not fit shared memory");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("wgrad: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
    attr = true;
  }
  wgrad_kernel<BN, STAGES><<<ctas, kWgThreads, smem, s>>>(ty, tx, wp);
  return check_launch("vbg_linear_wgrad(tcgen05 bf16x3, MN-major operands)");
}

// splits chosen so that tiles * splits fills the chip, each split keeping >= 4 row blocks
static int wg_splits(int tiles, int n_blocks) {
  int s = kNumSMs / (tiles > 0 ? tiles : 1);
  if (s < 1) s = 1;
  if (s > n_blocks / 4) s = n_blocks / 4 > 0 ? n_blocks / 4 : 1;
  if (s > 64) s = 64;
  return s;
}

size_t linear_wgrad_workspace(int M, int N, int K) {
  if (N % 64 || K % 64 || M <= 0) return 0;
  const int bn = K % 128 == 0 ? 128 : 64;
  const int tiles = cdiv(N, 128) * cdiv(K, bn), n_blocks = cdiv(M, 64);
  const int s = wg_splits(tiles, n_blocks);
  return s > 1 ? (size_t)s * N * K * 4 : 0;
}

int linear_wgrad(const void* dY, long long y_plane, const void* X, long long x_plane, int M, int N, int K, float* dW, void* workspace,
                 size_t ws_bytes, cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  if (N % 64 || K % 64 || !aligned16(dY) || !aligned16(X) || !aligned16(dW) || (y_plane & 7) || (x_plane & 7) || y_plane <= 0 || x_plane <= 0)
    return VBG_EUNSUPPORTED;
  const int bn = K % 128 == 0 ? 128 : 64;
  CUtensorMap ty, tx;
  if (!map_rows(&ty, dY, y_plane, M, N, N) || !map_rows(&tx, X, x_plane, M, K, K)) return VBG_EUNSUPPORTED;
  WgParams wp{};
  wp.rows = M; wp.box_rows = 64; wp.n_tiles = cdiv(N, 128); wp.k_tiles = cdiv(K, bn);
  const int tiles = wp.n_tiles * wp.k_tiles, n_blocks = cdiv(M, 64);
  int splits = wg_splits(tiles, n_blocks);
  if (splits > 1 && (!workspace || !aligned16(workspace) || (size_t)splits * N * K * 4 > ws_bytes)) splits = 1;
  wp.blocks_per_split = cdiv(n_blocks, splits);
  splits = cdiv(n_blocks, wp.blocks_per_split);
  wp.tc.M = N; wp.tc.N = K; wp.tc.ldc = K; wp.tc.conv = 0; wp.tc.num_kb = 0;
  wp.tc.C = splits > 1 ? reinterpret_cast<float*>(workspace) : dW;
  int rc = bn == 128 ? launch_wg<128, 3>(ty, tx, wp, tiles * splits, s) : launch_wg<64, 4>(ty, tx, wp, tiles * splits, s);
  if (rc != VBG_OK || splits == 1) return rc;
  return splitk_finish(reinterpret_cast<const float*>(workspace), splits, N, K, vbg_epilogue_t{}, dW, K, s);
}

// dW[Cout, kh, kw, Cin] = sum over output pixels of dY[b,ho,wo,Cout] x X[b, ho*st + r - pad, wo*st + s - pad, Cin]: the same kernel, the
// X boxes come from a rank-5 NHWC map at the tap's shift (padding = TMA out-of-bounds zero fill, stride = traversal stride).
static bool wg_conv_geometry(int B, int Ho, int Wo, WgParams& wp) {
  wp.tw = Wo < 64 ? Wo : 64;
  wp.th = (64 / wp.tw) < Ho ? (64 / wp.tw) : Ho;
  wp.tb = (wp.th == Ho && wp.tw == Wo) ? ((64 / (wp.tw * wp.th)) < B ? (64 / (wp.tw * wp.th)) : B) : 1;
  wp.box_rows = wp.tw * wp.th * wp.tb;                  // <= 64 by construction
  wp.tiles_w = cdiv(Wo, wp.tw); wp.tiles_h = cdiv(Ho, wp.th);
  wp.rows = wp.tiles_w * wp.tiles_h * cdiv(B, wp.tb) * 64;
  return true;
}

size_t conv_wgrad_workspace(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad) {
  if (Cout % 64 || Cin % 64 || stride < 1 || stride > 2) return 0;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  WgParams wp{};
  if (Ho <= 0 || Wo <= 0 || !wg_conv_geometry(B, Ho, Wo, wp)) return 0;
  const int bn = Cin % 128 == 0 ? 128 : 64;
  const int tiles = cdiv(Cout, 128) * kh * kw * (Cin / bn);
  const int s = wg_splits(tiles, wp.rows / 64);
  return s > 1 ? (size_t)s * Cout * kh * kw * Cin * 4 : 0;
}

int conv_wgrad(const void* dY, long long y_plane, const void* X, long long x_plane, int B, int H, int W, int Cin, int Cout, int kh, int kw,
               int stride, int pad, float* dW, void* workspace, size_t ws_bytes, cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  if (Cout % 64 || Cin % 64 || stride < 1 || stride > 2 || !aligned16(dY) || !aligned16(X) || !aligned16(dW) || (y_plane & 7) ||
      (x_plane & 7) || y_plane <= 0 || x_plane <= 0)
    return VBG_EUNSUPPORTED;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  WgParams wp{};
  if (Ho <= 0 || Wo <= 0 || !wg_conv_geometry(B, Ho, Wo, wp)) return VBG_EUNSUPPORTED;
  if (wp.tw * stride > 256 || wp.th * stride > 256) return VBG_EUNSUPPORTED;
  const int bn = Cin % 128 == 0 ? 128 : 64, Kt = kh * kw * Cin;
  wp.conv = 1; wp.kw = kw; wp.cin_tiles = Cin / bn; wp.st = stride; wp.pad = pad;
  wp.n_tiles = cdiv(Cout, 128); wp.k_tiles = kh * kw * wp.cin_tiles;   // Cout % 128 == 64: the upper half tile is TMA zero fill
  CUtensorMap ty, tx;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cout, (cuuint64_t)Wo, (cuuint64_t)Ho, (cuuint64_t)B, 2};
    cuuint64_t str[4] = {(cuuint64_t)Cout * 2, (cuuint64_t)Wo * Cout * 2, (cuuint64_t)Ho * Wo * Cout * 2, (cuuint64_t)y_plane * 2};
    cuuint32_t box[5] = {64u, (cuuint32_t)wp.tw, (cuuint32_t)wp.th, (cuuint32_t)wp.tb, 1u};
    if (!tc_encode(&ty, dY, 5, dims, str, box, nullptr, true)) return VBG_EUNSUPPORTED;
  }
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
    cuuint64_t str[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2, (cuuint64_t)x_plane * 2};
    cuuint32_t box[5] = {64u, (cuuint32_t)(wp.tw * stride), (cuuint32_t)(wp.th * stride), (cuuint32_t)wp.tb, 1u};
    cuuint32_t estr[5] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u, 1u};
    if (!tc_encode(&tx, X, 5, dims, str, box, estr, true)) return VBG_EUNSUPPORTED;
  }
  const int tiles = wp.n_tiles * wp.k_tiles, n_blocks = wp.rows / 64;
  int splits = wg_splits(tiles, n_blocks);
  if (splits > 1 && (!workspace || !aligned16(workspace) || (size_t)splits * Cout * Kt * 4 > ws_bytes)) splits = 1;
  wp.blocks_per_split = cdiv(n_blocks, splits);
  splits = cdiv(n_blocks, wp.blocks_per_split);
  wp.tc.M = Cout; wp.tc.N = Kt; wp.tc.ldc = Kt; wp.tc.conv = 0; wp.tc.num_kb = 0;
  wp.tc.C = splits > 1 ? reinterpret_cast<float*>(workspace) : dW;
  int rc = bn == 128 ? launch_wg<128, 3>(ty, tx, wp, tiles * splits, s) : launch_wg<64, 4>(ty, tx, wp, tiles * splits, s);
  if (rc != VBG_OK || splits == 1) return rc;
  return splitk_finish(reinterpret_cast<const float*>(workspace), splits, Cout, Kt, vbg_epilogue_t{}, dW, Kt, s);
}

}  // namespace vbg
