// Shared device/host helpers of the tcgen05 / TMA kernels (sm_100a).
#pragma once
#include "vbg_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

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

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Instruction descriptor (32-bit): D=f32 [4,6)=1; A fmt [7,10), B fmt [10,13) (kind::tf32: TF32=2; kind::f16: F16=0, BF16=1);
// A,B K-major (bits 15,16 = 0); N>>3 at [17,23); M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(uint32_t ab_fmt, int m, int n, uint32_t b_mn_major = 0) {
  return (1u << 4) | (ab_fmt << 7) | (ab_fmt << 10) | (b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr uint32_t kFmtTF32 = 2, kFmtBF16 = 1;

// ------------------------------------------------------------------ tile geometry shared by the tcgen05 GEMM kernels
constexpr int BM = 128, BKE = 32;          // 32 fp32 = one 128-byte swizzle row
constexpr int kTcThreads = 192;            // warp 0 TMA, warp 1 MMA, warps 2-5 convert / epilogue
constexpr int kEpiStageFloats = 32 * 36;   // per-warp 32x32 transpose buffer, row stride 36 floats (conflict-free float4)

struct TcParams {
  float* C; int ldc;
  int M, N;
  int m_tiles, n_tiles;   // persistent kernels: tile grid (m_tiles = row / spatial tiles)
  int num_kb;      // K blocks: of 32 fp32 (kind::tf32 kernel) or of 64 (bf16x3 kernel)
  int kb_split;    // first K block served by the second A map (cat-free two-source GEMM); == num_kb when unused
  // implicit-GEMM conv (conv == 1): output tile = tb images x th rows x tw cols (tw*th*tb <= 128)
  int conv, tw, th, tb, tiles_w, tiles_h, Ho, Wo, Bn;
  int cin_blocks;  // 32-float blocks per filter tap
  int kw;          // filter taps per filter row
  int sw, sh;      // TMA start-coordinate stride of the output tile origin along W / H
  int pad_w, pad_h;
  vbg_epilogue_t ep;
};

struct TcTile { int m0, n0, w0, h0, b0; };

__device__ __forceinline__ TcTile tc_tile_origin(const TcParams& p, int bn) {
  TcTile t{0, (int)blockIdx.y * bn, 0, 0, 0};
  if (p.conv) {
    int i = blockIdx.x;
    t.w0 = (i % p.tiles_w) * p.tw; i /= p.tiles_w;
    t.h0 = (i % p.tiles_h) * p.th; i /= p.tiles_h;
    t.b0 = i * p.tb;
  } else {
    t.m0 = blockIdx.x * BM;
  }
  return t;
}

// Epilogue of one 128 x BN accumulator tile (called by warps 2-5; q = warp & 3 owns TMEM lanes [32q, 32q+32)).
// Phase 1: tcgen05.ld 32x32b (lane == accumulator row) -> per-warp smem transpose buffer.
// Phase 2: lane owns 4 consecutive columns of 4 rows per pass -> scale/shift/residual/activation with 128-bit
//          loads and fully coalesced 128-bit stores (each 8-lane group writes one 128-byte row segment).
template <int BN>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, const TcTile& t, uint32_t tmem_base, int q, int lane,
                                            float* __restrict__ stage) {
  const vbg_epilogue_t& ep = p.ep;
  const int sub = lane >> 3, c4 = (lane & 7) * 4;
  long long out_off[8], res_off[8];
  uint32_t ok_mask = 0;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = q * 32 + it * 4 + sub;
    long long m_out; bool ok;
    if (p.conv) {
      const int wi = t.w0 + r % p.tw, tt = r / p.tw;
      const int hi = t.h0 + tt % p.th, bi = t.b0 + tt / p.th;
      ok = (r < p.tw * p.th * p.tb) && wi < p.Wo && hi < p.Ho && bi < p.Bn;
      m_out = ((long long)bi * p.Ho + hi) * p.Wo + wi;
    } else {
      m_out = t.m0 + r;
      ok = m_out < p.M;
    }
    long long rr = 0;
    if (ep.residual && ok) {
      if (ep.res_mode == VBG_RES_UP2) {
        const int wo = (int)(m_out % ep.out_w); const long long u = m_out / ep.out_w;
        const int ho = (int)(u % ep.out_h); const long long b = u / ep.out_h;
        rr = ((b * (ep.out_h >> 1) + (ho >> 1)) * (ep.out_w >> 1) + (wo >> 1)) * (long long)p.N;
      } else {
        rr = m_out * ep.ldr;
      }
    }
    out_off[it] = m_out * p.ldc; res_off[it] = rr;
    if (ok) ok_mask |= 1u << it;
  }
  const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.N & 3) == 0) &&
                      (!ep.residual || (((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0) &&
                                        (ep.res_mode == VBG_RES_UP2 || (ep.ldr & 3) == 0)));
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    if (t.n0 + c0 >= p.N) break;                       // warp-uniform
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    __syncwarp();                                      // previous pass finished reading the buffer
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(stage + lane * 36 + j * 4) =
          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    __syncwarp();
    const int n = t.n0 + c0 + c4;
    if (n >= p.N) continue;
    float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (n + e < p.N) {
        if (ep.scale) sc[e] = __ldg(ep.scale + n + e);
        if (ep.shift) sh[e] = __ldg(ep.shift + n + e);
      }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if (!((ok_mask >> it) & 1u)) continue;
      const float4 a = *reinterpret_cast<const float4*>(stage + (it * 4 + sub) * 36 + c4);
      float o[4] = {a.x * sc[0] + sh[0], a.y * sc[1] + sh[1], a.z * sc[2] + sh[2], a.w * sc[3] + sh[3]};
      if (ep.residual) {
        const float* rp = ep.residual + res_off[it] + n;
        if (vec_ok && n + 3 < p.N) {
          const float4 rv = __ldg(reinterpret_cast<const float4*>(rp));
          o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < p.N) o[e] += __ldg(rp + e);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = apply_act(o[e], ep.act);
      if (ep.out_mode == VBG_OUT_SPLIT_BF16) {          // bf16 hi / lo planes (operands of the split attention kernel)
        __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(p.C) + out_off[it] + n;
        __nv_bfloat16* lp = hp + ep.out_plane;
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { h[e] = __float2bfloat16_rn(o[e]); l[e] = __float2bfloat16_rn(o[e] - __bfloat162float(h[e])); }
        if (vec_ok && n + 3 < p.N && (ep.out_plane & 3) == 0) {
          *reinterpret_cast<uint2*>(hp) = make_uint2((uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16),
                                                     (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16));
          *reinterpret_cast<uint2*>(lp) = make_uint2((uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16),
                                                     (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) if (n + e < p.N) { hp[e] = h[e]; lp[e] = l[e]; }
        }
        continue;
      }
      float* cp = p.C + out_off[it] + n;
      if (vec_ok && n + 3 < p.N) {
        *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (n + e < p.N) cp[e] = o[e];
      }
    }
  }
}

// ---- host helpers (vbg_gemm_tc.cu)
bool tc_available();
bool tc_disabled_by_env();
// rank-N tiled map with SWIZZLE_128B; dtype_bf16 selects 2-byte elements
bool tc_encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides, bool dtype_bf16);

}  // namespace vbg
