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

__device__ __forceinline__ TcTile ps_tile(const TcParams& p, int unit, int bn) {
  const int n_tiles = p.m_tiles * p.n_tiles;
  const int tile = unit % n_tiles, split = unit / n_tiles;       // the splits of a tile run in different waves' worth of CTAs
  const int xt = tile % p.m_tiles;             // m fastest: CTAs running together share the weight tile
  TcTile t{0, (tile / p.m_tiles) * bn, 0, 0, 0, 0, p.num_kb, 0};
  if (p.splits > 1) {
    t.kb0 = split * p.kb_per_split;
    t.kb1 = min(t.kb0 + p.kb_per_split, p.num_kb);
    t.row_shift = (long long)split * p.M;
  }
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
  const int n_tiles = p.m_tiles * p.n_tiles * (p.splits > 1 ? p.splits : 1);      // work units

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
        for (int kb = t.kb0; kb < t.kb1; ++kb, ++g) {
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
        const TcTile t = ps_tile(p, tile, BN);
        for (int kb = t.kb0; kb < t.kb1; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&full[s], (g / STAGES) & 1);
          tc_fence_after();
          const uint32_t base = smem_u32(ring + s * STAGE_BYTES);
          const uint64_t a1 = ps_desc<KB>(base), a2 = ps_desc<KB>(base + A_BYTES);
          const uint64_t w1 = ps_desc<KB>(base + 2 * A_BYTES), w2 = ps_desc<KB>(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {          // 16 bf16 = 32 B per MMA: +2 in the (addr>>4) field
            const uint64_t o = (uint64_t)(2 * k);
            umma_bf16(d, a1 + o, w1 + o, idesc, (kb != t.kb0) || (k != 0));
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

// ------------------------------------------------------------------ CTA-pair variant (tcgen05 cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) compute one 256 x BN tile: each loads ITS 128 rows of the A planes and
// HALF of the weight tile (BN/2 rows); the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads A from each
// CTA's own shared memory and the two B halves from both, and accumulates rows [0,128) into the leader's TMEM and rows
// [128,256) into the peer's.  Per CTA and K block that is 64 KB of L2 -> SM traffic instead of 96 KB and 8 KB instead of
// 12 KB of operand reads per MMA -- the single-CTA kernel is bound by exactly those two (shared-memory bandwidth:
// 96 B/clk operand reads + 64 B/clk TMA writes > 128 B/clk).
//   both CTAs   warp 0: TMA producer (own A rows + own B half; completion bytes land on the LEADER's full barrier)
//               warps 2-9: epilogue of the CTA's own 128 rows (own TMEM), then a remote arrive on the leader's tmem_empty
//   leader      warp 1: MMA issuer; tcgen05.commit multicasts to both CTAs' empty / tmem_full barriers
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(local_smem_addr), "r"(rank));
  return a;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: data into the executing CTA's smem, completion bytes on the barrier at cluster address `bar`
__device__ __forceinline__ void tma2_load_3d(const CUtensorMap* tm, uint32_t bar, void* dst, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma2_load_5d(const CUtensorMap* tm, uint32_t bar, void* dst, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once the MMAs issued so far complete) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

template <int BN, int KB, int STAGES>
__global__ void __launch_bounds__(kPsThreads, 1)
gemm_ps2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                const __grid_constant__ CUtensorMap tmW, const TcParams p) {
  constexpr int BH = BN / 2;                                          // weight rows held by each CTA
  constexpr uint32_t A_BYTES = BM * KB * 2;
  constexpr uint32_t B_BYTES = BH * KB * 2;
  constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;         // A1 | A2 | W1 half | W2 half
  constexpr uint32_t TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* epi = reinterpret_cast<float*>(ring + STAGES * STAGE_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + STAGES * STAGE_BYTES + kPsEpiBytes);   // used in the leader only
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;        // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2], used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int pm_tiles = (p.m_tiles + 1) >> 1;                          // pairs of 128-row tiles
  const int n_tiles = pm_tiles * p.n_tiles;
  if (threadIdx.x == 0) tc_stamp(p, 0);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA); prefetch_tmap(&tmW);
    if (p.kb_split < p.num_kb) prefetch_tmap(&tmA2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 512); }   // 2 CTAs x 256 epilogue threads
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();                          // both CTAs' barriers are initialised before any remote arrive / complete_tx
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) tc_stamp(p, 1);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) tc_stamp(p, 2);

  // this CTA's 128-row tile inside pair tile pt
  auto my_tile = [&](int pt) {
    const int pm = pt % pm_tiles, nt = pt / pm_tiles;
    const int xt = 2 * pm + (int)rank;         // may be == m_tiles for the odd tail: every load is then out of bounds (zero
    TcTile t{0, nt * BN, 0, 0, 0, 0, p.num_kb, 0};   // fill) and every epilogue row is masked
    if (p.conv) {
      int i = xt;
      t.w0 = (i % p.tiles_w) * p.tw; i /= p.tiles_w;
      t.h0 = (i % p.tiles_h) * p.th; i /= p.tiles_h;
      t.b0 = i * p.tb;
    } else {
      t.m0 = xt * BM;
    }
    return t;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs)
      const uint32_t a_rows = p.conv ? (uint32_t)(p.tw * p.th * p.tb) : (uint32_t)BM;
      const uint32_t tx_pair = 2u * (2u * a_rows * (uint32_t)(KB * 2) + 2u * B_BYTES);
      int g = 0;
      for (int pt = cluster_id; pt < n_tiles; pt += n_clusters) {
        const TcTile t = my_tile(pt);
        const int nb = t.n0 + (int)rank * BH;
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
          uint8_t* sa = ring + s * STAGE_BYTES;
          uint8_t* sb = sa + 2 * A_BYTES;
          if (rank == 0) mbar_expect_tx(&full[s], tx_pair);
          const uint32_t bar = mapa_rank(smem_u32(&full[s]), 0);
          if (p.conv) {
            const int tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
            const int fr = tap / p.kw, fs = tap - fr * p.kw;
            const int cw = t.w0 * p.sw + fs - p.pad_w, ch = t.h0 * p.sh + fr - p.pad_h;
            tma2_load_5d(&tmA, bar, sa, cb * KB, cw, ch, t.b0, 0);
            tma2_load_5d(&tmA, bar, sa + A_BYTES, cb * KB, cw, ch, t.b0, 1);
          } else if (kb < p.kb_split) {
            tma2_load_3d(&tmA, bar, sa, kb * KB, t.m0, 0);
            tma2_load_3d(&tmA, bar, sa + A_BYTES, kb * KB, t.m0, 1);
          } else {
            tma2_load_3d(&tmA2, bar, sa, (kb - p.kb_split) * KB, t.m0, 0);
            tma2_load_3d(&tmA2, bar, sa + A_BYTES, (kb - p.kb_split) * KB, t.m0, 1);
          }
          tma2_load_3d(&tmW, bar, sb, kb * KB, nb, 0);
          tma2_load_3d(&tmW, bar, sb + B_BYTES, kb * KB, nb, 1);
          if (g == 0) tc_stamp(p, 3);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader): per 16-wide K step  D += A1*W1 ; D += A2*W1 ; D += A1*W2, M = 256 over the pair
      constexpr uint32_t idesc = make_idesc(kFmtBF16, 2 * BM, BN);
      int g = 0, it = 0;
      for (int pt = cluster_id; pt < n_tiles; pt += n_clusters, ++it) {
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], ((it >> 1) & 1) ^ 1);             // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb, ++g) {
          const int s = g % STAGES;
          mbar_wait(&full[s], (g / STAGES) & 1);
          tc_fence_after();
          if (g == 0) tc_stamp(p, 4);
          const uint32_t base = smem_u32(ring + s * STAGE_BYTES);
          const uint64_t a1 = ps_desc<KB>(base), a2 = ps_desc<KB>(base + A_BYTES);
          const uint64_t w1 = ps_desc<KB>(base + 2 * A_BYTES), w2 = ps_desc<KB>(base + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            const uint64_t o = (uint64_t)(2 * k);
            umma2_bf16(d, a1 + o, w1 + o, idesc, (kb | k) != 0);
            umma2_bf16(d, a2 + o, w1 + o, idesc, 1);
            umma2_bf16(d, a1 + o, w2 + o, idesc, 1);
          }
          umma2_commit_both(&empty[s]);
        }
        umma2_commit_both(&tmem_full[acc]);
        if (it == 0) tc_stamp(p, 5);
      }
    }
  } else {
    // ===== epilogue warps (both CTAs): own 128 rows from own TMEM
    const int q = warp & 3, half = (warp - 2) >> 2;
    int it = 0;
    for (int pt = cluster_id; pt < n_tiles; pt += n_clusters, ++it) {
      const int acc = it & 1;
      const TcTile t = my_tile(pt);
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      if (it == 0 && threadIdx.x == 64) tc_stamp(p, 6);
      tc_epilogue<BN>(p, t, tmem_base + (uint32_t)(acc * BN), q, lane, epi + (warp - 2) * kEpiStageFloats, half * 32, 64);
      tc_fence_before();
      if (it == 0 && threadIdx.x == 64) tc_stamp(p, 7);
      mbar_arrive_cluster(mapa_rank(smem_u32(&tmem_empty[acc]), 0));
    }
  }

  tc_fence_before();
  if (threadIdx.x == 0) tc_stamp(p, 8);
  cluster_sync_all();                          // the peer's smem / TMEM / barriers stay alive until both CTAs are done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    if (lane == 0) tc_stamp(p, 9);
  }
}

// ------------------------------------------------------------------ split-K finish
// out[m, n] = epilogue( sum_s ws[s][m][n] ), splits added in a fixed order (deterministic, unlike fp32 atomics): the partial
// accumulators of an under-filled GEMM (few output tiles, long K) come from `splits` work units that ran on different SMs.
__global__ void __launch_bounds__(256)
splitk_finish_kernel(const float* __restrict__ ws, int splits, long long M, int N, vbg_epilogue_t ep, void* __restrict__ C, int ldc) {
  const int N4 = N >> 2;
  const long long total = M * N4, plane = M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / N4;
    const int n = (int)(i - m * N4) * 4;
    float4 a = __ldg(reinterpret_cast<const float4*>(ws + m * N + n));
    for (int s = 1; s < splits; ++s) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ws + s * plane + m * N + n));
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.scale) sc = __ldg(reinterpret_cast<const float4*>(ep.scale + n));
    if (ep.shift) sh = __ldg(reinterpret_cast<const float4*>(ep.shift + n));
    float4 o = make_float4(fmaf(a.x, sc.x, sh.x), fmaf(a.y, sc.y, sh.y), fmaf(a.z, sc.z, sh.z), fmaf(a.w, sc.w, sh.w));
    if (ep.residual) {
      long long r;
      if (ep.res_mode == VBG_RES_UP2) {
        const int wo = (int)(m % ep.out_w); const long long u = m / ep.out_w;
        const int ho = (int)(u % ep.out_h); const long long b = u / ep.out_h;
        r = ((b * (ep.out_h >> 1) + (ho >> 1)) * (ep.out_w >> 1) + (wo >> 1)) * (long long)N + n;
      } else {
        r = m * ep.ldr + n;
      }
      const float4 rv = ld4_fmt(ep.residual, ep.res_plane, (size_t)(r >> 2));
      o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
    }
    o.x = apply_act(o.x, ep.act); o.y = apply_act(o.y, ep.act); o.z = apply_act(o.z, ep.act); o.w = apply_act(o.w, ep.act);
    st4_fmt(C, ep.out_mode == VBG_OUT_SPLIT_BF16 ? ep.out_plane : 0, (size_t)((m * ldc + n) >> 2), o);
  }
}

int splitk_finish(const float* ws, int splits, long long M, int N, const vbg_epilogue_t& ep, void* C, int ldc, cudaStream_t s) {
  const long long total4 = M * (N / 4);
  int blocks = (int)((total4 + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  splitk_finish_kernel<<<blocks, 256, 0, s>>>(ws, splits, M, N, ep, C, ldc);
  return check_launch("split-K finish");
}

// ------------------------------------------------------------------ split -> fp32 (inspection / tests)
__global__ void merge_bf16_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo, long long n,
                                  float* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) out[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}

// ------------------------------------------------------------------ host side
// K elements per ring stage: a per-call choice (vbg_epilogue_t::tune) so the parity tests can exercise both variants.
static int ps_kb(int tune) { return (tune & VBG_TUNE_KB32) ? 32 : kPsDefaultKB; }

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
  const bool pdl = !(p.ep.tune & VBG_TUNE_NO_PDL);
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

static long long* g_timeline = nullptr;
long long* tc_debug_timeline() { return g_timeline; }
void tc_debug_set_timeline(long long* buf) { g_timeline = buf; }

template <int BN, int KB, int STAGES>
static int launch_ps2(const CUtensorMap& a, const CUtensorMap& a2, const CUtensorMap& w, const TcParams& p, cudaStream_t s) {
  constexpr size_t smem = (size_t)STAGES * (2 * BM * KB * 2 + 2 * (BN / 2) * KB * 2) + kPsEpiBytes + 1024 + 256;
  static_assert(smem <= 232448, "pre-split pair tile does not fit the 227 KB shared-memory limit");
  static_assert(8 * (2 * STAGES + 4) + 4 <= 256, "barrier block overflows its 256 bytes");
  static int max_clusters = 0;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kPsThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[1].val.programmaticStreamSerializationAllowed = 1;
  const bool pdl = !(p.ep.tune & VBG_TUNE_NO_PDL);
  cfg.attrs = at; cfg.numAttrs = pdl ? 2 : 1;
  if (max_clusters == 0) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ps2_kernel<BN, KB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gemm_ps2: smem opt-in failed: %s", cudaGetErrorString(e)); return VBG_ECUDA; }
    cfg.gridDim = dim3(kNumSMs);
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, gemm_ps2_kernel<BN, KB, STAGES>, &cfg);
    if (e != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = kNumSMs / 2; }
    max_clusters = n < kNumSMs / 2 ? n : kNumSMs / 2;
  }
  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  cfg.gridDim = dim3(2 * (pair_tiles < max_clusters ? pair_tiles : max_clusters));
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_ps2_kernel<BN, KB, STAGES>, a, a2, w, p);
  if (e != cudaSuccess) { set_error("vbg_gemm_ps (CTA pair) launch failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return VBG_ECUDA; }
  return check_launch("vbg_gemm_ps(tcgen05 bf16x3, pre-split, cta_group::2)");
}

// CTA-pair tiles pay off when the kernel is bound by L2 -> SM feed / shared-memory bandwidth (which the pair cuts by a third /
// halves for the weight tile): a grid that fills most of the chip, a K long enough that the mainloop -- not the epilogue --
// sets the tile time, and a weight tile wide enough to matter next to the A tile.  Measured on B200 (profiles/
// r1_gemm_presplit_pairs_k.log): +10..25 % on the BERT GEMMs and the 256-channel convs, -4..8 % on N <= 128 / short-K shapes.
// VBG_TUNE_PAIRS_OFF / VBG_TUNE_PAIRS_ON in the call's `tune` force the choice (tests run both).
static bool use_pairs(int m_tiles, int n_tiles_1cta, int N, int num_kb, int tune) {
  if (tune & (VBG_TUNE_PAIRS_OFF | VBG_TUNE_PAIRS_ON)) return (tune & VBG_TUNE_PAIRS_ON) && m_tiles >= 2;
  return m_tiles >= 2 && N >= 192 && num_kb >= 8 && (long long)m_tiles * n_tiles_1cta >= 96;
}

// Single-CTA tiles: N tile width and split-K factor chosen together.  An under-filled grid (layer-3/4 convs at 32x32 / 16x16,
// the ROI FC: 8 - 64 row tiles) used to shrink BN to get more CTAs, which quadruples the A traffic per FLOP; splitting K
// instead keeps wide tiles and fills the SMs with (tile, split) work units.  Needs a workspace of splits * M * N floats.
static void pick_bn_splits(int m_tiles, int N, int num_kb, bool allow_split, int tune, int* bn_out, int* splits_out) {
  // Opt-in (VBG_TUNE_SPLITK): measured on B200 at cfg2 the split form (single-CTA units + finishing kernel) is SLOWER than
  // the narrow CTA-pair tiles it replaces (6.67 vs 6.37 ms/step) -- kept, with its parity tests, as a tuning experiment.
  if (!(tune & VBG_TUNE_SPLITK)) allow_split = false;
  const int cand[4] = {256, 192, 128, 64};
  const int scand[6] = {1, 2, 3, 4, 6, 8};
  int best_bn = 64, best_s = 1; double best = -1.0;
  for (int i = 0; i < 4; ++i) {
    const int bn = cand[i];
    if (bn > 64 && N < bn - 32) continue;
    for (int j = 0; j < 6; ++j) {
      const int sp = scand[j];
      if (sp > 1 && (!allow_split || num_kb / sp < 8 || (N & 3))) break;
      const long long units = (long long)m_tiles * cdiv(N, bn) * sp;
      const double waves = (double)units / kNumSMs;
      const double wave_eff = waves / (double)((units + kNumSMs - 1) / kNumSMs);
      const double fill = (double)N / ((double)cdiv(N, bn) * bn);
      const double tile_eff = (double)bn / (bn + 96.0);
      const double kb_unit = (double)num_kb / sp;
      const double split_eff = sp == 1 ? 1.0 : kb_unit / (kb_unit + 6.0);          // per-unit ramp / drain + the finishing pass
      const double score = wave_eff * fill * tile_eff * split_eff;
      if (score > best) { best = score; best_bn = bn; best_s = sp; }
    }
  }
  *bn_out = best_bn; *splits_out = best_s;
}

static int dispatch_ps(const CUtensorMap& a, const CUtensorMap& a2, const void* w_hi, long long w_plane, int ldw, int K, TcParams& p,
                       int m_tiles, int kb, void* workspace, size_t ws_bytes, cudaStream_t s) {
  p.dbg = g_timeline;
  p.splits = 1; p.kb_per_split = p.num_kb;
  CUtensorMap w;
  p.m_tiles = m_tiles;
  int bn = 64, splits = 1;
  const bool ws_ok = workspace && aligned16(workspace);
  pick_bn_splits(m_tiles, p.N, p.num_kb, ws_ok, p.ep.tune, &bn, &splits);
  // a grid that stays under one wave even with the widest tile is a split-K case, not a pair case
  const bool underfilled = (long long)m_tiles * cdiv(p.N, 256) * 2 <= kNumSMs;
  if (!(underfilled && splits > 1)) {
  const int bn_pairs_probe = pick_bn3(m_tiles, p.N);
  if (kb == 64 && use_pairs(m_tiles, cdiv(p.N, bn_pairs_probe), p.N, p.num_kb, p.ep.tune)) {
    // pair tiles: same (wave efficiency x column fill x tile efficiency) score as pick_bn3, over clusters of two SMs
    const int pm_tiles = (m_tiles + 1) / 2, n_cl = kNumSMs / 2;
    const int cand[4] = {256, 192, 128, 64};
    int b2 = 64; double best = -1.0;
    for (int i = 0; i < 4; ++i) {
      const int bn_c = cand[i];
      if (bn_c > 64 && p.N < bn_c - 32) continue;
      const long long tiles = (long long)pm_tiles * cdiv(p.N, bn_c);
      const double waves = (double)tiles / n_cl;
      const double score = waves / (double)((tiles + n_cl - 1) / n_cl) * ((double)p.N / ((double)cdiv(p.N, bn_c) * bn_c)) *
                           ((double)bn_c / (bn_c + 64.0));
      if (score > best) { best = score; b2 = bn_c; }
    }
    if (!map_ps_2d(&w, w_hi, w_plane, p.N, K, ldw, b2 / 2, kb)) return VBG_EUNSUPPORTED;
    p.n_tiles = cdiv(p.N, b2);
    if (b2 == 256) return launch_ps2<256, 64, 3>(a, a2, w, p, s);
    if (b2 == 192) return launch_ps2<192, 64, 3>(a, a2, w, p, s);
    if (b2 == 128) return launch_ps2<128, 64, 4>(a, a2, w, p, s);
    return launch_ps2<64, 64, 4>(a, a2, w, p, s);
  }
  }
  if (splits > 1) {
    const int per = cdiv(p.num_kb, splits);
    splits = cdiv(p.num_kb, per);                                   // no empty split
    if (splits < 2 || (size_t)splits * (size_t)p.M * (size_t)p.N * 4 > ws_bytes) splits = 1;
    else p.kb_per_split = per;
  }
  if (splits == 1) pick_bn_splits(m_tiles, p.N, p.num_kb, false, p.ep.tune, &bn, &splits);
  if (!map_ps_2d(&w, w_hi, w_plane, p.N, K, ldw, bn, kb)) return VBG_EUNSUPPORTED;
  p.n_tiles = cdiv(p.N, bn);
  TcParams q = p;                                                   // what the GEMM kernel sees
  if (splits > 1) {
    q.splits = splits;
    q.C = reinterpret_cast<float*>(workspace); q.ldc = p.N;
    q.ep = vbg_epilogue_t{};                                        // raw fp32 partial accumulators
    q.ep.tune = p.ep.tune;
  }
  int rc;
  if (kb == 64) {
    if (bn == 256) rc = launch_ps<256, 64, 2>(a, a2, w, q, s);
    else if (bn == 192) rc = launch_ps<192, 64, 2>(a, a2, w, q, s);
    else if (bn == 128) rc = launch_ps<128, 64, 3>(a, a2, w, q, s);
    else rc = launch_ps<64, 64, 4>(a, a2, w, q, s);
  } else {
    if (bn == 256) rc = launch_ps<256, 32, 4>(a, a2, w, q, s);
    else if (bn == 192) rc = launch_ps<192, 32, 4>(a, a2, w, q, s);
    else if (bn == 128) rc = launch_ps<128, 32, 6>(a, a2, w, q, s);
    else rc = launch_ps<64, 32, 8>(a, a2, w, q, s);
  }
  if (rc != VBG_OK || splits == 1) return rc;
  const long long total4 = (long long)p.M * (p.N / 4);
  int blocks = (int)((total4 + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  splitk_finish_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float*>(workspace), splits, p.M, p.N, p.ep, p.C, p.ldc);
  return check_launch("vbg_gemm_ps(split-K finish)");
}

// Workspace a caller should provide so that under-filled shapes may split K (0: the shape never splits).
size_t ps_workspace_bytes(int m_tiles, long long M, int N, int K, int tune) {
  if (K % 64 || N < 64 || (N & 3)) return 0;
  const int num_kb = K / 64;
  int bn, splits;
  pick_bn_splits(m_tiles, N, num_kb, true, tune, &bn, &splits);
  const bool underfilled = (long long)m_tiles * cdiv(N, 256) * 2 <= kNumSMs;
  if (!(underfilled && splits > 1) && use_pairs(m_tiles, cdiv(N, pick_bn3(m_tiles, N)), N, num_kb, tune)) return 0;
  return splits > 1 ? (size_t)splits * (size_t)M * (size_t)N * 4 : 0;
}

static bool planes_ok(const void* p, long long plane) { return p && aligned16(p) && plane > 0 && (plane % 8) == 0; }

int gemm_ps(const void* A, long long a_plane, int lda, const void* A2, long long a2_plane, int lda2, int K1, const void* w_hi,
            long long w_plane, int ldw, void* C, int ldc, int M, int N, int K, const vbg_epilogue_t* ep, void* workspace,
            size_t ws_bytes, cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  // N < 64: one 64-wide tile with N live columns (weight rows beyond N are TMA zero fill) -- worth it only for tall problems
  if ((N < 64 && ((N & 3) || M < 8192)) || K % 64 || K1 % 64 || (lda & 7) || (ldw & 7) || !planes_ok(A, a_plane) || !planes_ok(w_hi, w_plane))
    return VBG_EUNSUPPORTED;
  if (K1 < K && ((lda2 & 7) || !planes_ok(A2, a2_plane))) return VBG_EUNSUPPORTED;
  const int kb = ps_kb(ep ? ep->tune : 0);
  CUtensorMap ta, ta2;
  if (!map_ps_2d(&ta, A, a_plane, M, K1, lda, BM, kb)) return VBG_EUNSUPPORTED;
  if (K1 < K) { if (!map_ps_2d(&ta2, A2, a2_plane, M, K - K1, lda2, BM, kb)) return VBG_EUNSUPPORTED; } else ta2 = ta;
  TcParams p{};
  p.C = reinterpret_cast<float*>(C); p.ldc = ldc; p.M = M; p.N = N; p.num_kb = K / kb; p.kb_split = K1 / kb; p.conv = 0;
  if (ep) p.ep = *ep;
  return dispatch_ps(ta, ta2, w_hi, w_plane, ldw, K, p, cdiv(M, BM), kb, workspace, ws_bytes, s);
}

bool tc_conv_geometry(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, TcParams& p,
                      cuuint64_t dims[4], cuuint64_t strides_b[3], cuuint32_t box[4], cuuint32_t estr[4]);

size_t conv_ps_workspace_bytes(int B, int H, int W, int Cin, int Cout, int kh, int kw, int stride, int pad, int tune) {
  TcParams p{};
  cuuint64_t d4[4], s4[3]; cuuint32_t b4[4], e4[4];
  if (Cin % 64 || !tc_conv_geometry(B, H, W, Cin, Cout, kh, kw, stride, pad, p, d4, s4, b4, e4)) return 0;
  return ps_workspace_bytes(p.tiles_w * p.tiles_h * cdiv(B, p.tb), p.M, Cout, kh * kw * Cin, tune);
}

int conv_ps(const void* x, long long x_plane, int B, int H, int W, int Cin, const void* w_hi, long long w_plane, int Cout, int kh,
            int kw, int stride, int pad, void* y, const vbg_epilogue_t* ep, void* workspace, size_t ws_bytes, cudaStream_t s) {
  if (!tc_available()) return VBG_EUNSUPPORTED;
  if (Cin % 64 || Cout < 64 || !planes_ok(x, x_plane) || !planes_ok(w_hi, w_plane)) return VBG_EUNSUPPORTED;
  const int kb = ps_kb(ep ? ep->tune : 0);
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
  return dispatch_ps(ta, ta, w_hi, w_plane, K, K, p, p.tiles_w * p.tiles_h * cdiv(B, p.tb), kb, workspace, ws_bytes, s);
}

int merge_bf16(const void* hi, const void* lo, long long n, float* out, cudaStream_t s) {
  if (n == 0) return VBG_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  merge_bf16_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(hi), reinterpret_cast<const __nv_bfloat16*>(lo), n, out);
  return check_launch("vbg_merge_bf16");
}

}  // namespace vbg
