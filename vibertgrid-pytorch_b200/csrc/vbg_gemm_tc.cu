// VBG_PREC_TF32: the dense contractions on 5th-generation tensor cores.
//
//   C[M,N] = epilogue( A[M,K] * W[N,K]^T )          fp32 in HBM, TF32 multiply, fp32 accumulate
//
// One CTA computes a 128 x BN output tile.  Warp-specialised, mbarrier-pipelined:
//   warp 0    TMA producer : cp.async.bulk.tensor (SWIZZLE_128B boxes of 32 fp32 = 128 B along K)
//   warp 1    MMA issuer   : one thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8), accumulator in TMEM
//   warps 2-5 epilogue     : tcgen05.ld 32x32b (lane == output row) -> scale/shift/residual/activation -> global
//
// Operand A comes from one of
//   * a 2-D map over a row-major [M,K] matrix, optionally two maps split along K
//     (the torch.cat-free early-fusion / late-fusion GEMMs), or
//   * a 4-D map over the NHWC activation [B,H,W,C] for stride-1 convolutions: for filter tap (r,s) the
//     A tile of an output patch is the SAME patch shifted by (r-pad, s-pad); TMA's out-of-bounds
//     zero fill implements the padding, so im2col is never materialised (implicit GEMM).
// Operand B is the weight [N,K] (K-major), a 2-D map.  Both operands are K-major in shared memory in
// the canonical 128-byte-swizzled layout (8 rows x 128 B atoms, SBO = 1024 B) that UMMA descriptors
// address; stepping K by 8 elements inside the swizzle row advances the descriptor start by 32 B.
//
// Shapes this path does not take (K % 32, Cin % 32, N < 64, strided / 7x7 convs) return
// VBG_EUNSUPPORTED and the dispatcher uses the fp32 CUDA-core kernel (vbg_gemm_simt.cu).
#include "vbg_tc.cuh"
#include <dlfcn.h>
#include <mutex>

namespace vbg {

template <int BN, int STAGES>
__global__ void __launch_bounds__(kTcThreads)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(4 * kEpiStageFloats * 4 <= STAGES * STAGE_BYTES, "epilogue staging must fit in the operand ring");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const TcTile t = tc_tile_origin(p, BN);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW);
    if (p.kb_split < p.num_kb) prefetch_tmap(&tmA2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
      mbar_init(tmem_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer
      const uint32_t a_bytes = p.conv ? (uint32_t)(p.tw * p.th * p.tb) * 128u : A_BYTES;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        mbar_expect_tx(&full_bar[s], a_bytes + B_BYTES);
        if (p.conv) {
          const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
          const int fr = tap / p.kw, fs = tap - fr * p.kw;
          tma_load_4d(&tmA, &full_bar[s], sa, cb * BKE, t.w0 * p.sw + fs - p.pad_w, t.h0 * p.sh + fr - p.pad_h, t.b0);
        } else if (kb < p.kb_split) {
          tma_load_2d(&tmA, &full_bar[s], sa, kb * BKE, t.m0);
        } else {
          tma_load_2d(&tmA2, &full_bar[s], sa, (kb - p.kb_split) * BKE, t.m0);
        }
        tma_load_2d(&tmW, &full_bar[s], sb, kb * BKE, t.n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer
      constexpr uint32_t idesc = make_idesc(kFmtTF32, BM, BN);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full_bar[s], (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t da = make_sw128_desc(sa), db = make_sw128_desc(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BKE / 8; ++k)        // 8 tf32 = 32 B per MMA: +2 in the (addr>>4) field
          umma_tf32(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
        umma_commit(&empty_bar[s]);              // frees the smem slot once these MMAs have read it
      }
      umma_commit(tmem_full);                    // accumulator complete
    }
  } else {
    // ===== epilogue (the operand ring is idle once tmem_full fires: reuse it as the transpose buffers)
    const int q = warp & 3;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    tc_epilogue<BN>(p, t, tmem_base, q, lane, reinterpret_cast<float*>(smem) + q * kEpiStageFloats);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static int g_tc_state = -1;           // -1 unknown, 0 unavailable, 1 available
static std::mutex g_tc_mu;

bool tc_available() {
  std::lock_guard<std::mutex> lk(g_tc_mu);
  if (g_tc_state >= 0) return g_tc_state == 1;
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cudaGetLastError(); set_error("tcgen05 path: cudaGetDevice: %s", cudaGetErrorString(e)); return false; }  // not cached
  g_tc_state = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess || major != 10) {
    cudaGetLastError();
    set_error("tcgen05 path: device %d has compute capability major %d (sm_100a required)", dev, major);
    return false;
  }
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres = cudaDriverEntryPointSymbolNotFound;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess) {
    cudaGetLastError();
    fn = nullptr;
    // fall back to the driver library itself
    void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (h) fn = dlsym(h, "cuTensorMapEncodeTiled");
    if (!fn) {
      set_error("tcgen05 path: cuTensorMapEncodeTiled not found (cudaGetDriverEntryPoint: %s, query result %d)",
                cudaGetErrorString(e), (int)qres);
      return false;
    }
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  g_tc_state = 1;
  return true;
}

bool tc_encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides, bool dtype_bf16, bool swizzle64) {
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(tm, dtype_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                        const_cast<void*>(base), dims, strides_bytes, box, elem_strides ? elem_strides : ones,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return false; }
  return true;
}

static bool encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box) {
  return tc_encode(tm, base, rank, dims, strides_bytes, box, nullptr, false);
}

static bool map_2d(CUtensorMap* tm, const float* base, long long rows, long long cols, long long ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BKE, (cuuint32_t)box_rows};
  return encode(tm, base, 2, dims, strides, box);
}

template <int BN, int STAGES>
static int launch(const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& w, const TcParams& p, dim3 grid, cudaStream_t s) {
  constexpr size_t smem = (size_t)STAGES * (BM * 128 + BN * 128) + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gemm_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
    attr = true;
  }
  gemm_tc_kernel<BN, STAGES><<<grid, kTcThreads, smem, s>>>(a, a2, w, p);
  return check_launch("vbg_gemm(tcgen05)");
}

static int pick_bn(int m_tiles, int N) {
  // largest N tile that still gives every SM a CTA; 256-wide tiles halve A re-reads from L2
  if (N >= 256 && (long long)m_tiles * cdiv(N, 256) >= kNumSMs) return 256;
  if (N >= 128 && (long long)m_tiles * cdiv(N, 128) >= kNumSMs / 2) return 128;
  return N >= 128 && N % 128 == 0 && (long long)m_tiles * (N / 64) > 2 * kNumSMs ? 128 : 64;
}

static int dispatch(const CUtensorMap& a, const CUtensorMap& a2, CUtensorMap& w, const float* W, int ldw, int K, TcParams& p,
                    int m_tiles, cudaStream_t s) {
  const int bn = pick_bn(m_tiles, p.N);
  if (!map_2d(&w, W, p.N, K, ldw, bn)) return VBG_ECUDA;
  dim3 grid(m_tiles, cdiv(p.N, bn));
  if (bn == 256) return launch<256, 4>(a, a2, w, p, grid, s);
  if (bn == 128) return launch<128, 3>(a, a2, w, p, grid, s);
  return launch<64, 4>(a, a2, w, p, grid, s);
}

int gemm_tc(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int ldw, float* C, int ldc,
            int M, int N, int K, const vbg_epilogue_t* ep, cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  if (N < 64 || K % BKE || K1 % BKE || (lda & 3) || (ldw & 3) || !aligned16(A) || !aligned16(W)) return VBG_EUNSUPPORTED;
  if (K1 < K && ((lda2 & 3) || !aligned16(A2))) return VBG_EUNSUPPORTED;
  CUtensorMap ta, ta2, tw;
  if (!map_2d(&ta, A, M, K1, lda, BM)) return VBG_ECUDA;
  if (K1 < K) { if (!map_2d(&ta2, A2, M, K - K1, lda2, BM)) return VBG_ECUDA; } else ta2 = ta;
  TcParams p{};
  p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.num_kb = K / BKE; p.kb_split = K1 / BKE; p.conv = 0;
  if (ep) p.ep = *ep;
  return dispatch(ta, ta2, tw, W, ldw, K, p, cdiv(M, BM), s);
}

// Implicit-GEMM geometry shared by the kind::tf32 and bf16x3 conv paths.  Returns false for shapes TMA cannot tile.
bool tc_conv_geometry(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, TcParams& p,
                      cuuint64_t dims[4], cuuint64_t strides_b[3], cuuint32_t box[4], cuuint32_t estr[4]) {
  if (stride < 1 || stride > 2) return false;
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  if (Ho <= 0 || Wo <= 0) return false;
  p.conv = 1; p.Ho = Ho; p.Wo = Wo; p.Bn = B; p.kw = kw; p.pad_w = p.pad_h = pad; p.sw = p.sh = stride;
  p.cin_blocks = Cin / BKE;
  p.tw = Wo < BM ? Wo : BM;
  p.th = (BM / p.tw) < Ho ? (BM / p.tw) : Ho;
  p.tb = (p.th == Ho && p.tw == Wo) ? ((BM / (p.tw * p.th)) < B ? (BM / (p.tw * p.th)) : B) : 1;
  if (p.tw * stride > 256 || p.th * stride > 256 || p.tb > 256) return false;
  p.tiles_w = cdiv(Wo, p.tw); p.tiles_h = cdiv(Ho, p.th);
  p.M = B * Ho * Wo; p.N = Cout; p.ldc = Cout;
  dims[0] = (cuuint64_t)Cin; dims[1] = (cuuint64_t)W; dims[2] = (cuuint64_t)H; dims[3] = (cuuint64_t)B;
  strides_b[0] = (cuuint64_t)Cin * 4; strides_b[1] = (cuuint64_t)W * Cin * 4; strides_b[2] = (cuuint64_t)H * W * Cin * 4;
  // with a traversal stride s, TMA loads ceil(box / s) elements: box = n * s loads n
  box[0] = (cuuint32_t)BKE; box[1] = (cuuint32_t)(p.tw * stride); box[2] = (cuuint32_t)(p.th * stride); box[3] = (cuuint32_t)p.tb;
  estr[0] = 1; estr[1] = (cuuint32_t)stride; estr[2] = (cuuint32_t)stride; estr[3] = 1;
  return true;
}

int conv_tc(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int kh, int kw, int stride, int pad,
            float* y, const vbg_epilogue_t* ep, cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  if (Cin % BKE || Cout < 64 || !aligned16(x) || !aligned16(w)) return VBG_EUNSUPPORTED;
  TcParams p{};
  cuuint64_t dims[4], strides[3]; cuuint32_t box[4], estr[4];
  if (!tc_conv_geometry(B, H, W, Cin, Cout, kh, kw, stride, pad, p, dims, strides, box, estr)) return VBG_EUNSUPPORTED;
  const int tiles_b = cdiv(B, p.tb);
  p.C = y;
  const int K = kh * kw * Cin;
  p.num_kb = K / BKE; p.kb_split = p.num_kb;
  if (ep) p.ep = *ep;
  if (p.ep.res_mode == VBG_RES_UP2) { p.ep.out_h = p.Ho; p.ep.out_w = p.Wo; }
  if (p.ep.res_mode == VBG_RES_SAME && p.ep.ldr == 0) p.ep.ldr = Cout;
  CUtensorMap ta, tw;
  if (!tc_encode(&ta, x, 4, dims, strides, box, estr, false)) return VBG_EUNSUPPORTED;
  return dispatch(ta, ta, tw, w, K, K, p, p.tiles_w * p.tiles_h * tiles_b, s);
}

}  // namespace vbg
