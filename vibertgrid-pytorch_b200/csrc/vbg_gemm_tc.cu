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
#include "vbg_common.cuh"
#include <cuda.h>
#include <dlfcn.h>
#include <mutex>

namespace vbg {

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("vbg tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (tile base 1024-B aligned):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64: 8 rows x 128 B)
//   [46,48) version = 1 (Blackwell) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// ------------------------------------------------------------------ kernel
constexpr int BM = 128, BKE = 32;          // 32 fp32 = one 128-byte swizzle row
constexpr int kTcThreads = 192;

struct TcParams {
  float* C; int ldc;
  int M, N, num_kb, kb_split;
  // conv tiling (conv == 1)
  int conv, tw, th, tb, tiles_w, tiles_h, Ho, Wo, Bn, cin_blocks, kw, pad;
  vbg_epilogue_t ep;
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kTcThreads)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  constexpr uint32_t A_BYTES = BM * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;

  // ---- tile origin
  int m0 = 0, w0 = 0, h0 = 0, b0 = 0;
  if (p.conv) {
    int t = blockIdx.x;
    w0 = (t % p.tiles_w) * p.tw; t /= p.tiles_w;
    h0 = (t % p.tiles_h) * p.th; t /= p.tiles_h;
    b0 = t * p.tb;
  } else {
    m0 = blockIdx.x * BM;
  }

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
          tma_load_4d(&tmA, &full_bar[s], sa, cb * BKE, w0 + fs - p.pad, h0 + fr - p.pad, b0);
        } else if (kb < p.kb_split) {
          tma_load_2d(&tmA, &full_bar[s], sa, kb * BKE, m0);
        } else {
          tma_load_2d(&tmA2, &full_bar[s], sa, (kb - p.kb_split) * BKE, m0);
        }
        tma_load_2d(&tmW, &full_bar[s], sb, kb * BKE, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer.  Instruction descriptor: D=f32 [4,6)=1, A=tf32 [7,10)=2, B=tf32 [10,13)=2,
      // A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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
    // ===== epilogue: warp (w % 4) owns TMEM lanes [32*(w%4), +32); lane == accumulator row
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    long long m_out; bool row_ok;
    if (p.conv) {
      const int wi = w0 + r % p.tw, t = r / p.tw;
      const int hi = h0 + t % p.th, bi = b0 + t / p.th;
      row_ok = (r < p.tw * p.th * p.tb) && wi < p.Wo && hi < p.Ho && bi < p.Bn;
      m_out = ((long long)bi * p.Ho + hi) * p.Wo + wi;
    } else {
      m_out = m0 + r;
      row_ok = m_out < p.M;
    }
    const vbg_epilogue_t& ep = p.ep;
    long long res_row = 0;
    if (ep.residual && row_ok) {
      if (ep.res_mode == VBG_RES_UP2) {
        const int wo = (int)(m_out % ep.out_w); const long long t = m_out / ep.out_w;
        const int ho = (int)(t % ep.out_h); const long long b = t / ep.out_h;
        res_row = ((b * (ep.out_h >> 1) + (ho >> 1)) * (ep.out_w >> 1) + (wo >> 1)) * (long long)p.N;
      } else {
        res_row = m_out * ep.ldr;
      }
    }
    float* crow = p.C + m_out * p.ldc;
    const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.N & 3) == 0);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);     // warp-collective: all lanes execute
      if (!row_ok || n0 + c0 >= p.N) continue;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int n = n0 + c0 + j;
        if (n >= p.N) break;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float x = __uint_as_float(v[j + e]);
          if (n + e < p.N) {
            if (ep.scale) x *= __ldg(ep.scale + n + e);
            if (ep.shift) x += __ldg(ep.shift + n + e);
            if (ep.residual) x += __ldg(ep.residual + res_row + n + e);
            x = apply_act(x, ep.act);
          }
          o[e] = x;
        }
        if (vec_ok && n + 3 < p.N) {
          *reinterpret_cast<float4*>(crow + n) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < p.N) crow[n + e] = o[e];
        }
      }
    }
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

static bool encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box) {
  cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, ones,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return false; }
  return true;
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

static bool tc_disabled_by_env() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VBG_DISABLE_TC"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

int gemm_tc(const float* A, int lda, const float* A2, int lda2, int K1, const float* W, int ldw, float* C, int ldc,
            int M, int N, int K, const vbg_epilogue_t* ep, cudaStream_t s) {
  if (tc_disabled_by_env() || !tc_available()) return VBG_EUNSUPPORTED;
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

int conv_tc(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int kh, int kw, int stride, int pad,
            float* y, const vbg_epilogue_t* ep, cudaStream_t s) {
  if (tc_disabled_by_env() || !tc_available()) return VBG_EUNSUPPORTED;
  if (stride != 1 || Cin % BKE || Cout < 64 || !aligned16(x) || !aligned16(w)) return VBG_EUNSUPPORTED;
  const int Ho = H + 2 * pad - kh + 1, Wo = W + 2 * pad - kw + 1;
  if (Ho <= 0 || Wo <= 0) return VBG_EUNSUPPORTED;
  TcParams p{};
  p.conv = 1; p.Ho = Ho; p.Wo = Wo; p.Bn = B; p.kw = kw; p.pad = pad; p.cin_blocks = Cin / BKE;
  p.tw = Wo < BM ? Wo : BM;
  p.th = (BM / p.tw) < Ho ? (BM / p.tw) : Ho;
  p.tb = (p.th == Ho && p.tw == Wo) ? ((BM / (p.tw * p.th)) < B ? (BM / (p.tw * p.th)) : B) : 1;
  if (p.tw > 256 || p.th > 256 || p.tb > 256) return VBG_EUNSUPPORTED;
  p.tiles_w = cdiv(Wo, p.tw); p.tiles_h = cdiv(Ho, p.th);
  const int tiles_b = cdiv(B, p.tb);
  p.C = y; p.ldc = Cout; p.M = B * Ho * Wo; p.N = Cout;
  const int K = kh * kw * Cin;
  p.num_kb = K / BKE; p.kb_split = p.num_kb;
  if (ep) p.ep = *ep;
  if (p.ep.res_mode == VBG_RES_UP2) { p.ep.out_h = Ho; p.ep.out_w = Wo; }
  if (p.ep.res_mode == VBG_RES_SAME && p.ep.ldr == 0) p.ep.ldr = Cout;
  CUtensorMap ta, tw;
  cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)W * Cin * 4, (cuuint64_t)H * W * Cin * 4};
  cuuint32_t box[4] = {(cuuint32_t)BKE, (cuuint32_t)p.tw, (cuuint32_t)p.th, (cuuint32_t)p.tb};
  if (!encode(&ta, x, 4, dims, strides, box)) return VBG_ECUDA;
  return dispatch(ta, ta, tw, w, K, K, p, p.tiles_w * p.tiles_h * tiles_b, s);
}

}  // namespace vbg
