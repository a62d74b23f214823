// VBG_PREC_BF16X3 with PRE-SPLIT activations: the steady-state GEMM / implicit-GEMM conv of the joint forward.
//
//   C[M,N] = epilogue( (A1 + A2)[M,K] * (W1 + W2)[N,K]^T )  ~=  A1*W1 + A2*W1 + A1*W2      (fp32 accumulate in TMEM)
//
// Between tensor-core kernels every activation lives in HBM as two bf16 planes (hi = bf16_rn(x), lo = bf16_rn(x - hi);
// the same bytes as fp32), written by the producing kernel's epilogue.  So, unlike vbg_gemm_tc3.cu (fp32 A, converted
// inside the kernel by four extra warps through shared memory), nothing is converted here: TMA drops the A planes and the
// weight planes straight into the swizzled K-major operand tiles and one thread issues three tcgen05.mma.kind::f16 per
// 16-wide K step.  That removes the converter's shared-memory traffic (read 32 KB + write 32 KB per 64-wide K block on
// top of the tensor core's own operand reads) and its latency from the TMA -> MMA chain.
//
//   warp 0      TMA producer   A planes: rank-3 map (k, row, plane) or rank-5 NHWC map (c, w, h, b, plane): one filter tap
//                              per K block, conv padding = TMA out-of-bounds zero fill, stride 2 = traversal stride
//                              W planes: rank-3 map (k, n, plane)
//   warp 1      MMA issuer     3 x tcgen05.mma (M=128, N=BN, K=16) per K step into one of two TMEM accumulators
//   warps 2-9   epilogue       tcgen05.ld -> scale/shift/residual/activation -> fp32 or bf16 hi/lo planes (vbg_tc.cuh);
//                              two warps per TMEM lane quarter, alternating 32-column chunks
//
// Persistent over output tiles (grid = min(tiles, 148)); the operand ring runs straight through tile boundaries and the
// epilogue of tile i overlaps the mainloop of tile i+1.  KB = K elements per ring stage: 64 (SWIZZLE_128B rows) or
// 32 (SWIZZLE_64B rows, twice as many, finer stages for the same shared memory).
#include "vbg_tc.cuh"
#include <stdlib.h>

namespace vbg {

constexpr int kPsThreads = 320;                  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr uint32_t kPsEpiBytes = 8 * kEpiStageFloats * 4;
constexpr int kPsDefaultKB = 64;

__device__ __forceinline__ TcTile ps_tile(const TcParams& p, int tile, int bn) {
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

template <int KB>
__device__ __forceinline__ uint64_t ps_desc(uint32_t addr) { return KB == 64 ? make_sw128_desc(addr) : make_sw64_desc(addr); }

template <int BN, int KB, int STAGES>
__global__ void __launch_bounds__(kPsThreads, 1)
gemm_ps_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  constexpr uint32_t A_BYTES = BM * KB * 2;                           // one plane tile: 128 rows x KB bf16
  constexpr uint32_t B_BYTES = BN * KB * 2;
  constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;         // A1 | A2 | W1 | W2
  constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(ring + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES + kPsEpiBytes);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW);
    if (p.kb_split < p.num_kb) prefetch_tmap(&tmA2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 256); }
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
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) may run while
  // the previous kernel of the stream is still draining; nothing below touches global memory before that kernel has
  // completed and flushed.  Our own dependents may be scheduled as soon as SMs free up (they wait the same way).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      const uint32_t a_rows = p.conv ? (uint32_t)(p.tw * p.th * p.tb) : (uint32_t)BM;
      const uint32_t tx_bytes = 2u * a_rows * (uint32_t)(KB * 2) + 2u * B_BYTES;
      int g = 0;                                                       // K blocks issued so far (runs through tiles)
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const TcTile t = ps_tile(p, tile, BN);
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
          uint8_t* sa = ring + s * STAGE_BYTES;
          uint8_t* sb = sa + 2 * A_BYTES;
          mbar_expect_tx(&full[s], tx_bytes);
          if (p.conv) {
            const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
            const int fr = tap / p.kw, fs = tap - fr * p.kw;
            const int cw = t.w0 * p.sw + fs - p.pad_w, ch = t.h0 * p.sh + fr - p.pad_h;
            tma_load_5d(&tmA, &full[s], sa, cb * KB, cw, ch, t.b0, 0);
            tma_load_5d(&tmA, &full[s], sa + A_BYTES, cb * KB, cw, ch, t.b0, 1);
          } else if (kb < p.kb_split) {
            tma_load_3d(&tmA, &full[s], sa, kb * KB, t.m0, 0);
            tma_load_3d(&tmA, &full[s], sa + A_BYTES, kb * KB, t.m0, 1);
          } else {
            tma_load_3d(&tmA2, &full[s], sa, (kb - p.kb_split) * KB, t.m0, 0);
            tma_load_3d(&tmA2, &full[s], sa + A_BYTES, (kb - p.kb_split) * KB, t.m0, 1);
          }
          tma_load_3d(&tmW, &full[s], sb, kb * KB, t.n0, 0);
          tma_load_3d(&tmW, &full[s], sb + B_BYTES, kb * KB, t.n0, 1);
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
          mbar_wait(&full[s], (g / STAGES) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(ring + s * STAGE_BYTES);
          const uint64_t a1 = ps_desc<KB>(base), a2 = ps_desc<KB>(base + A_BYTES);
          const uint64_t w1 = ps_desc<KB>(base + 2 * A_BYTES), w2 = ps_desc<KB>(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {          // 16 bf16 = 32 B per MMA: +2 in the (addr>>4) field
            const uint64_t o = (uint64_t)(2 * k);
            umma_bf16(d, a1 + o, w1 + o, idesc, (kb | k) != 0);
            umma_bf16(d, a2 + o, w1 + o, idesc, 1);
            umma_bf16(d, a1 + o, w2 + o, idesc, 1);
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else {
    // ===== epilogue warps: TMEM lanes [32q, 32q+32), q = warp & 3; warps 2-5 take the even 32-column chunks, 6-9 the odd ones
    const int q = warp & 3, half = (warp - 2) >> 2;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const TcTile t = ps_tile(p, tile, BN);
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      tc_epilogue<BN>(p, t, tmem_base + (uint32_t)(acc * BN), q, lane, epi + (warp - 2) * kEpiStageFloats, half * 32, 64);
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

// ------------------------------------------------------------------ split -> fp32 (inspection / tests)
__global__ void merge_bf16_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, long long n,
                                  float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) out[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}

// ------------------------------------------------------------------ host side
// K elements per ring stage.  Read per call (not cached) so the parity tests can exercise both variants in one process.
static int ps_kb() {
  const char* e = getenv("VBG_PS_KB");
  const int k = e ? atoi(e) : 0;
  return (k == 32 || k == 64) ? k : kPsDefaultKB;
}

// rank-3 (k, row, plane) map over bf16 planes [rows, ld] + [rows, ld] `plane` elements later
static bool map_ps_2d(CUtensorMap* tm, const void* hi, long long plane, long long rows, long long cols, long long ld, int box_rows,
                      int kb) {
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, 2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane * 2};
  cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)box_rows, 1};
  return tc_encode(tm, hi, 3, dims, strides, box, nullptr, true, kb == 32);
}

template <int BN, int KB, int STAGES>
static int launch_ps(const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& w, const TcParams& p, cudaStream_t s) {
  constexpr size_t smem = (size_t)STAGES * (2 * BM * KB * 2 + 2 * BN * KB * 2) + kPsEpiBytes + 1024 + 256;
  static_assert(smem <= 232448, "pre-split tile does not fit the 227 KB shared-memory limit");
  static_assert(8 * (2 * STAGES + 4) + 4 <= 256, "barrier block overflows its 256 bytes");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ps_kernel<BN, KB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gemm_ps: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
    attr = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  static const bool pdl = [] { const char* e = getenv("VBG_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles < kNumSMs ? tiles : kNumSMs); cfg.blockDim = dim3(kPsThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_ps_kernel<BN, KB, STAGES>, a, a2, w, p);
  if (e != cudaSuccess) { set_error("vbg_gemm_ps launch failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return VBG_ECUDA; }
  return check_launch("vbg_gemm_ps(tcgen05 bf16x3, pre-split)");
}

int pick_bn3(int m_tiles, int N);

static int dispatch_ps(const CUtensorMap& a, const CUtensorMap& a2, const void* w_hi, long long w_plane, int ldw, int K, TcParams& p,
                       int m_tiles, int kb, cudaStream_t s) {
  const int bn = pick_bn3(m_tiles, p.N);
  CUtensorMap w;
  if (!map_ps_2d(&w, w_hi, w_plane, p.N, K, ldw, bn, kb)) return VBG_EUNSUPPORTED;
  p.m_tiles = m_tiles; p.n_tiles = cdiv(p.N, bn);
  if (kb == 64) {
    if (bn == 256) return launch_ps<256, 64, 2>(a, a2, w, p, s);
    if (bn == 192) return launch_ps<192, 64, 2>(a, a2, w, p, s);
    if (bn == 128) return launch_ps<128, 64, 3>(a, a2, w, p, s);
    return launch_ps<64, 64, 4>(a, a2, w, p, s);
  }
  if (bn == 256) return launch_ps<256, 32, 4>(a, a2, w, p, s);
  if (bn == 192) return launch_ps<192, 32, 4>(a, a2, w, p, s);
  if (bn == 128) return launch_ps<128, 32, 6>(a, a2, w, p, s);
  return launch_ps<64, 32, 8>(a, a2, w, p, s);
}

static bool planes_ok(const void* p, long long plane) { return p && aligned16(p) && plane > 0 && (plane % 8) == 0; }

int gemm_ps(const void* A, long long a_plane, int lda, const void* A2, long long a2_plane, int lda2, int K1, const void* w_hi,
            long long w_plane, int ldw, void* C, int ldc, int M, int N, int K, const vbg_epilogue_t* ep, cudaStream_t s) {
  if (tc_disabled_by_env() || !tc_available()) return VBG_EUNSUPPORTED;
  if (N < 64 || K % 64 || K1 % 64 || (lda & 7) || (ldw & 7) || !planes_ok(A, a_plane) || !planes_ok(w_hi, w_plane)) return VBG_EUNSUPPORTED;
  if (K1 < K && ((lda2 & 7) || !planes_ok(A2, a2_plane))) return VBG_EUNSUPPORTED;
  const int kb = ps_kb();
  CUtensorMap ta, ta2;
  if (!map_ps_2d(&ta, A, a_plane, M, K1, lda, BM, kb)) return VBG_EUNSUPPORTED;
  if (K1 < K) { if (!map_ps_2d(&ta2, A2, a2_plane, M, K - K1, lda2, BM, kb)) return VBG_EUNSUPPORTED; } else ta2 = ta;
  TcParams p{};
  p.C = reinterpret_cast<float*>(C); p.ldc = ldc; p.M = M; p.N = N; p.num_kb = K / kb; p.kb_split = K1 / kb; p.conv = 0;
  if (ep) p.ep = *ep;
  return dispatch_ps(ta, ta2, w_hi, w_plane, ldw, K, p, cdiv(M, BM), kb, s);
}

bool tc_conv_geometry(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, TcParams& p,
                      cuuint64_t dims[4], cuuint64_t strides_b[3], cuuint32_t box[4], cuuint32_t estr[4]);

int conv_ps(const void* x, long long x_plane, int B, int H, int W, int Cin, const void* w_hi, long long w_plane, int Cout, int kh,
            int kw, int stride, int pad, void* y, const vbg_epilogue_t* ep, cudaStream_t s) {
  if (tc_disabled_by_env() || !tc_available()) return VBG_EUNSUPPORTED;
  if (Cin % 64 || Cout < 64 || !planes_ok(x, x_plane) || !planes_ok(w_hi, w_plane)) return VBG_EUNSUPPORTED;
  const int kb = ps_kb();
  TcParams p{};
  cuuint64_t d4[4], s4[3]; cuuint32_t b4[4], e4[4];
  if (!tc_conv_geometry(B, H, W, Cin, Cout, kh, kw, stride, pad, p, d4, s4, b4, e4)) return VBG_EUNSUPPORTED;
  // the fp32 geometry above, re-expressed for bf16 planes: (c, w, h, b, plane)
  cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2, (cuuint64_t)x_plane * 2};
  cuuint32_t box[5] = {(cuuint32_t)kb, b4[1], b4[2], b4[3], 1};
  cuuint32_t estr[5] = {1, e4[1], e4[2], 1, 1};
  p.cin_blocks = Cin / kb;
  p.C = reinterpret_cast<float*>(y);
  const int K = kh * kw * Cin;
  p.num_kb = K / kb; p.kb_split = p.num_kb;
  if (ep) p.ep = *ep;
  if (p.ep.res_mode == VBG_RES_UP2) { p.ep.out_h = p.Ho; p.ep.out_w = p.Wo; }
  if (p.ep.res_mode == VBG_RES_SAME && p.ep.ldr == 0) p.ep.ldr = Cout;
  CUtensorMap ta;
  if (!tc_encode(&ta, x, 5, dims, strides, box, estr, true, kb == 32)) return VBG_EUNSUPPORTED;
  return dispatch_ps(ta, ta, w_hi, w_plane, K, K, p, p.tiles_w * p.tiles_h * cdiv(B, p.tb), kb, s);
}

int merge_bf16(const void* hi, const void* lo, long long n, float* out, cudaStream_t s) {
  if (n == 0) return VBG_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  merge_bf16_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(hi), reinterpret_cast<const __nv_bfloat16*>(lo), n, out);
  return check_launch("vbg_merge_bf16");
}

}  // namespace vbg
