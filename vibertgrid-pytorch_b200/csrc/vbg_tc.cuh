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
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* tm, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
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

// the same load without the wait: issue several, then tmem_wait_ld() once
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (tile base 1024-B aligned):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64: 8 rows x 128 B)
//   [46,48) version = 1 (Blackwell) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}

// K-major SWIZZLE_64B variant (rows of 64 B = 32 bf16; 8-row atom = 512 B, tile base 512-B aligned): layout = 4, SBO = 512 B
__device__ __forceinline__ uint64_t make_sw64_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)32 << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)4 << 61);
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

// ------------------------------------------------------------------ attention-probability dropout (training)
// keep(query row, key, head) is a pure function of a 32-bit seed: one lowbias32 hash per PAIR of keys (2j, 2j + 1), the low /
// high 16 bits decide the even / odd key; an entry is kept when its 16 bits are >= thr = round(p * 65536), so the effective
// drop probability is thr / 65536 (p = 0.1 -> 0.100006) and kept entries are scaled by 65536 / (65536 - thr).  Shared by the
// forward kernel (vbg_attn_tc.cu) and the two backward kernels (vbg_attn_bwd_tc.cu), which must agree bit for bit.
__host__ __device__ __forceinline__ uint32_t attn_drop_bits(uint32_t seed, uint32_t qrow, uint32_t kpair, uint32_t head) {
  uint32_t x = seed ^ (qrow * 0x9E3779B1u) ^ (kpair * 0x85EBCA77u) ^ (head * 0xC2B2AE3Du);
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ void attn_drop_pair(uint32_t seed, uint32_t qrow, uint32_t kpair, uint32_t head, uint32_t thr, float inv_keep,
                                               float& pa, float& pb) {
  const uint32_t bits = attn_drop_bits(seed, qrow, kpair, head);
  pa = (bits & 0xffffu) >= thr ? pa * inv_keep : 0.f;
  pb = (bits >> 16) >= thr ? pb * inv_keep : 0.f;
}
inline void attn_drop_params(float p, uint32_t& thr, float& inv_keep) {
  thr = p > 0.f ? (uint32_t)(p * 65536.0f + 0.5f) : 0u;
  if (thr > 65535u) thr = 65535u;
  inv_keep = 65536.0f / (65536.0f - (float)thr);
}
// fold the optional DEVICE step seed (refreshed by the host before each CUDA-graph replay) into the launch's 32-bit seed
__host__ __device__ __forceinline__ uint32_t attn_fold_step(uint32_t seed, unsigned long long step) {
  return seed ^ ((uint32_t)step * 0x9E3779B1u) ^ ((uint32_t)(step >> 32) * 0x85EBCA77u);
}
inline uint32_t attn_seed32(unsigned long long seed) { return (uint32_t)(seed ^ (seed >> 32)) * 0x9E3779B1u + 0x7F4A7C15u; }

// ------------------------------------------------------------------ tile geometry shared by the tcgen05 GEMM kernels
constexpr int BM = 128, BKE = 32;          // 32 fp32 = one 128-byte swizzle row
constexpr int kTcThreads = 192;            // warp 0 TMA, warp 1 MMA, warps 2-5 convert / epilogue
constexpr int kEpiStageFloats = 32 * 32;   // per-warp 32x32 transpose buffer; 16-byte chunks XOR-swizzled by row (conflict-free float4)

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
  int splits, kb_per_split;   // split-K (single-CTA pre-split kernel): work unit = (tile, split)
  long long* dbg;  // optional device buffer for clock64 stamps of CTA 0 (vbg_debug_set_timeline); nullptr in production
};

__device__ __forceinline__ void tc_stamp(const TcParams& p, int slot) {
  if (p.dbg && blockIdx.x == 0) p.dbg[slot] = clock64();
}
long long* tc_debug_timeline();

struct TcTile {
  int m0, n0, w0, h0, b0;
  int kb0, kb1;          // K-block range of this work unit (split-K; [0, num_kb) otherwise)
  long long row_shift;   // split-K: partial results of split s go to rows [s * M, (s+1) * M) of the workspace
};

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

// Epilogue of one 128 x BN accumulator tile (called by 4 warps; q = warp & 3 owns TMEM lanes [32q, 32q+32)).
// Per 32-column chunk: tcgen05.ld 32x32b (lane == accumulator row) -> per-warp smem transpose buffer -> each lane owns
// 4 consecutive columns of 8 rows, so scale / shift are two 128-bit loads per chunk, residual reads and output stores are
// 128-bit (fp32) or 64-bit x 2 planes (bf16 hi/lo) and every 8-lane group writes one contiguous row segment.
// The generic variant (any N / alignment) is kept out of line; the fast variant needs N % 4 == 0 and 16-byte aligned
// pointers, which every shape of the joint forward has.  The tile time of short-K GEMMs is THIS function (the mainloop of
// a K = 64 tile is 1.5k cycles), so it is written for instruction count: flags hoisted, no per-element branches.
// Output / residual element offsets of the 8 rows a lane owns after the transpose (row = 32q + 4it + lane/8); bit `it` of the
// returned mask is set when that row exists.
__device__ __forceinline__ uint32_t tc_row_offsets(const TcParams& p, const TcTile& t, int q, int sub, long long (&out_off)[8],
                                                   long long (&res_off)[8]) {
  const vbg_epilogue_t& ep = p.ep;
  const bool up2 = ep.residual && ep.res_mode == VBG_RES_UP2, same = ep.residual && ep.res_mode == VBG_RES_SAME;
  uint32_t ok_mask = 0;
  const int r0 = q * 32 + sub;
  // one division per lane, then rows advance by 4 with carries (integer division is ~25 instructions each; this runs
  // once per tile per lane and used to cost more than the mainloop of a short-K tile)
  if (p.conv) {
    int w = r0 % p.tw, tt = r0 / p.tw;
    int h = tt % p.th, b = tt / p.th;
    const int Ho2 = p.Ho >> 1, Wo2 = p.Wo >> 1;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int wi = t.w0 + w, hi = t.h0 + h, bi = t.b0 + b;
      const bool ok = b < p.tb && wi < p.Wo && hi < p.Ho && bi < p.Bn;
      const long long m_out = ((long long)bi * p.Ho + hi) * p.Wo + wi;
      out_off[it] = (m_out + t.row_shift) * p.ldc;
      res_off[it] = up2 ? (((long long)bi * Ho2 + (hi >> 1)) * Wo2 + (wi >> 1)) * (long long)p.N : (same ? m_out * ep.ldr : 0);
      if (ok) ok_mask |= 1u << it;
      w += 4;
      while (w >= p.tw) { w -= p.tw; if (++h == p.th) { h = 0; ++b; } }
    }
  } else {
    const int m0 = t.m0 + r0;
    int wo = 0, ho = 0, b = 0;
    if (up2) { wo = m0 % ep.out_w; const int u = m0 / ep.out_w; ho = u % ep.out_h; b = u / ep.out_h; }
    const int oh2 = ep.out_h >> 1, ow2 = ep.out_w >> 1;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const long long m_out = m0 + 4 * it;
      out_off[it] = (m_out + t.row_shift) * p.ldc;
      res_off[it] = up2 ? (((long long)b * oh2 + (ho >> 1)) * ow2 + (wo >> 1)) * (long long)p.N : (same ? m_out * ep.ldr : 0);
      if (m_out < p.M) ok_mask |= 1u << it;
      if (up2) {
        wo += 4;
        while (wo >= ep.out_w) { wo -= ep.out_w; if (++ho == ep.out_h) { ho = 0; ++b; } }
      }
    }
  }
  return ok_mask;
}

// staging buffer addressing: element (row, 4-column group g) lives at row*32 + ((g ^ (row & 7)) << 2)
__device__ __forceinline__ int epi_sw(int row, int g) { return row * 32 + ((g ^ (row & 7)) << 2); }

template <int BN>
__device__ __noinline__ void tc_epilogue_generic(const TcParams& p, const TcTile& t, uint32_t tmem_base, int q, int lane,
                                                 float* __restrict__ stage, int c_begin, int c_step) {
  const vbg_epilogue_t& ep = p.ep;
  const int sub = lane >> 3, g = lane & 7, c4 = g * 4;
  long long out_off[8], res_off[8];
  const uint32_t ok_mask = tc_row_offsets(p, t, q, sub, out_off, res_off);
#pragma unroll 1
  for (int c0 = c_begin; c0 < BN; c0 += c_step) {
    if (t.n0 + c0 >= p.N) break;                       // warp-uniform
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(stage + epi_sw(lane, j)) =
          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    __syncwarp();
    const int n = t.n0 + c0 + c4;
    if (n >= p.N) continue;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
      if (!((ok_mask >> it) & 1u)) continue;
#pragma unroll 1
      for (int e = 0; e < 4; ++e) {
        if (n + e >= p.N) break;
        float o = stage[epi_sw(it * 4 + sub, g) + e];
        o = o * (ep.scale ? __ldg(ep.scale + n + e) : 1.f) + (ep.shift ? __ldg(ep.shift + n + e) : 0.f);
        if (ep.residual) {
          if (ep.res_plane > 0) {
            const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + res_off[it] + n + e;
            o += __bfloat162float(hp[0]) + __bfloat162float(hp[ep.res_plane]);
          } else {
            o += __ldg(ep.residual + res_off[it] + n + e);
          }
        }
        o = apply_act(o, ep.act);
        if (ep.out_mode == VBG_OUT_SPLIT_BF16) {
          __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(p.C) + out_off[it] + n + e;
          const __nv_bfloat16 h = __float2bfloat16_rn(o);
          hp[0] = h;
          hp[ep.out_plane] = __float2bfloat16_rn(o - __bfloat162float(h));
        } else {
          p.C[out_off[it] + n + e] = o;
        }
      }
    }
  }
}

// kAllOk: all 8 rows of the lane exist (every full tile) -> no per-row predication.
template <int BN, bool kAllOk>
__device__ __forceinline__ void tc_epilogue_fast(const TcParams& p, const TcTile& t, uint32_t tmem_base, int q, int lane,
                                                 float* __restrict__ stage, int c_begin, int c_step, const long long (&out_off)[8],
                                                 const long long (&res_off)[8], uint32_t ok_mask) {
  const vbg_epilogue_t& ep = p.ep;
  const int sub = lane >> 3, g = lane & 7, c4 = g * 4;
  const bool split_out = ep.out_mode == VBG_OUT_SPLIT_BF16;
  const int res_kind = !ep.residual ? 0 : (ep.res_plane > 0 ? 2 : 1);
  const float* __restrict__ scale = ep.scale;
  const float* __restrict__ shift = ep.shift;
  const int act = ep.act;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  // scale / shift of the first chunk; the next chunk's are fetched while this one is processed
  float4 sc = one, sh = zero;
  {
    const int n = t.n0 + c_begin + c4;
    if (c_begin < BN && n < p.N) {
      if (scale) sc = __ldg(reinterpret_cast<const float4*>(scale + n));
      if (shift) sh = __ldg(reinterpret_cast<const float4*>(shift + n));
    }
  }
#pragma unroll 1
  for (int c0 = c_begin; c0 < BN; c0 += c_step) {
    if (t.n0 + c0 >= p.N) break;                       // warp-uniform
    float4 sc_n = one, sh_n = zero;
    {
      const int nn = t.n0 + c0 + c_step + c4;
      if (c0 + c_step < BN && nn < p.N) {
        if (scale) sc_n = __ldg(reinterpret_cast<const float4*>(scale + nn));
        if (shift) sh_n = __ldg(reinterpret_cast<const float4*>(shift + nn));
      }
    }
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    __syncwarp();                                      // previous chunk finished reading the buffer
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(stage + epi_sw(lane, j)) =
          make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    __syncwarp();
    const int n = t.n0 + c0 + c4;
    if (n < p.N) {                                     // N % 4 == 0: a 4-column group is entirely in or out
      float4 o[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const float4 a = *reinterpret_cast<const float4*>(stage + epi_sw(it * 4 + sub, g));
        o[it] = make_float4(fmaf(a.x, sc.x, sh.x), fmaf(a.y, sc.y, sh.y), fmaf(a.z, sc.z, sh.z), fmaf(a.w, sc.w, sh.w));
      }
      if (res_kind == 1) {
#pragma unroll
        for (int it = 0; it < 8; ++it)
          if (kAllOk || ((ok_mask >> it) & 1u)) {
            const float4 rv = __ldg(reinterpret_cast<const float4*>(ep.residual + res_off[it] + n));
            o[it].x += rv.x; o[it].y += rv.y; o[it].z += rv.z; o[it].w += rv.w;
          }
      } else if (res_kind == 2) {
        const __nv_bfloat16* rh = reinterpret_cast<const __nv_bfloat16*>(ep.residual) + n;
#pragma unroll
        for (int it = 0; it < 8; ++it)
          if (kAllOk || ((ok_mask >> it) & 1u)) {
            const float4 rv = merge4(__ldg(reinterpret_cast<const uint2*>(rh + res_off[it])),
                                     __ldg(reinterpret_cast<const uint2*>(rh + res_off[it] + ep.res_plane)));
            o[it].x += rv.x; o[it].y += rv.y; o[it].z += rv.z; o[it].w += rv.w;
          }
      }
      if (act == VBG_ACT_RELU) {
#pragma unroll
        for (int it = 0; it < 8; ++it)
          o[it] = make_float4(fmaxf(o[it].x, 0.f), fmaxf(o[it].y, 0.f), fmaxf(o[it].z, 0.f), fmaxf(o[it].w, 0.f));
      } else if (act == VBG_ACT_GELU) {
#pragma unroll
        for (int it = 0; it < 8; ++it)
          o[it] = make_float4(gelu_fast(o[it].x), gelu_fast(o[it].y), gelu_fast(o[it].z), gelu_fast(o[it].w));
      }
      if (split_out) {
        __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(p.C) + n;
#pragma unroll
        for (int it = 0; it < 8; ++it)
          if (kAllOk || ((ok_mask >> it) & 1u)) {
            uint2 hi, lo;
            split4(o[it], hi, lo);
            *reinterpret_cast<uint2*>(hp + out_off[it]) = hi;
            *reinterpret_cast<uint2*>(hp + out_off[it] + ep.out_plane) = lo;
          }
      } else {
        float* cp = p.C + n;
#pragma unroll
        for (int it = 0; it < 8; ++it)
          if (kAllOk || ((ok_mask >> it) & 1u)) *reinterpret_cast<float4*>(cp + out_off[it]) = o[it];
      }
    }
    sc = sc_n; sh = sh_n;
  }
}

// Columns [c_begin, BN) in steps of c_step (32 = this warp does every chunk; 64 with two warps per TMEM lane quarter).
template <int BN>
__device__ __forceinline__ void tc_epilogue(const TcParams& p, const TcTile& t, uint32_t tmem_base, int q, int lane,
                                            float* __restrict__ stage, int c_begin = 0, int c_step = 32) {
  const vbg_epilogue_t& ep = p.ep;
  const bool split_out = ep.out_mode == VBG_OUT_SPLIT_BF16;
  const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.N & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(ep.scale) & 15) == 0) && ((reinterpret_cast<uintptr_t>(ep.shift) & 15) == 0) &&
                      (!split_out || (ep.out_plane & 3) == 0) &&
                      (!ep.residual || (((reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0) && (ep.res_plane & 3) == 0 &&
                                        (ep.res_mode == VBG_RES_UP2 || (ep.ldr & 3) == 0)));
  if (!vec_ok) {
    tc_epilogue_generic<BN>(p, t, tmem_base, q, lane, stage, c_begin, c_step);
    return;
  }
  long long out_off[8], res_off[8];
  const uint32_t ok_mask = tc_row_offsets(p, t, q, lane >> 3, out_off, res_off);
  // warp-uniform choice: both variants execute tcgen05.ld.sync.aligned / __syncwarp, so the warp must not split here
  if (__all_sync(0xffffffffu, ok_mask == 0xffu)) tc_epilogue_fast<BN, true>(p, t, tmem_base, q, lane, stage, c_begin, c_step, out_off, res_off, ok_mask);
  else tc_epilogue_fast<BN, false>(p, t, tmem_base, q, lane, stage, c_begin, c_step, out_off, res_off, ok_mask);
}

// ---- host helpers (vbg_gemm_tc.cu)
bool tc_available();
// rank-N tiled map with SWIZZLE_128B; dtype_bf16 selects 2-byte elements
bool tc_encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, const cuuint32_t* elem_strides, bool dtype_bf16, bool swizzle64 = false);

}  // namespace vbg
