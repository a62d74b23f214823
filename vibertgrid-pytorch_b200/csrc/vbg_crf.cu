// CRF negative log-likelihood (training loss of the `crf` field-type head) and its gradient:
// reference model/crf.py:47-93,148-152 driven per sample by model/field_type_classification_head.py:686-699.
// One warp per sample, lane == tag; the per-sequence arithmetic lives in vbg_crf_seq.h, which the CPU test-suite compiles
// with g++ and checks against the reference (tests/test_crf_host.py).  Latency-bound by construction (a chain of n
// dependent steps per document): emissions / alphas / tags are staged through shared memory 64 steps at a time so the
// chain never waits on HBM.
#include "vbg_common.cuh"
#include "vbg_crf_seq.h"

namespace vbg {

__global__ void __launch_bounds__(32) crf_nll_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ trans,
                                                         const int32_t* __restrict__ tags, const int32_t* __restrict__ seg_off,
                                                         int T, float* __restrict__ alpha, float* __restrict__ logz,
                                                         float* __restrict__ nll) {
  __shared__ CrfScratch S;
  const int b = blockIdx.x;
  const int s0 = seg_off[b], n = seg_off[b + 1] - s0;
  crf_nll_fwd_seq(&S, feats + (size_t)s0 * T, trans, tags + s0, n, T, alpha + (size_t)s0 * T, logz + b, nll + b);
}

__global__ void __launch_bounds__(32) crf_nll_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ trans,
                                                         const int32_t* __restrict__ tags, const int32_t* __restrict__ seg_off,
                                                         int T, const float* __restrict__ alpha, const float* __restrict__ dnll, float* __restrict__ dfeats,
                                                         float* __restrict__ dtrans_part) {
  __shared__ CrfScratch S;
  const int b = blockIdx.x;
  const int s0 = seg_off[b], n = seg_off[b + 1] - s0;
  crf_nll_bwd_seq(&S, feats + (size_t)s0 * T, trans, tags + s0, n, T, alpha + (size_t)s0 * T, dnll[b],
                  dfeats + (size_t)s0 * T, dtrans_part + (size_t)b * T * T);
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_crf_nll_fwd(const float* feats, const float* trans, const int32_t* tags, const int32_t* seg_off, int B, int K,
                               int T, float* alpha, float* logz, float* nll, vbg_stream_t stream) {
  VBG_REQUIRE(feats && trans && tags && seg_off && alpha && logz && nll && B > 0 && K >= 0 && T >= 3 && T <= VBG_CRF_TMAX,
              "vbg_crf_nll_fwd: bad arguments (3 <= T <= 32)");
  crf_nll_fwd_kernel<<<B, 32, 0, as_stream(stream)>>>(feats, trans, tags, seg_off, T, alpha, logz, nll);
  return check_launch("vbg_crf_nll_fwd");
}

extern "C" int vbg_crf_nll_bwd(const float* feats, const float* trans, const int32_t* tags, const int32_t* seg_off, int B, int K,
                               int T, const float* alpha, const float* dnll, float* dfeats, float* dtrans_part,
                               vbg_stream_t stream) {
  VBG_REQUIRE(feats && trans && tags && seg_off && alpha && dnll && dfeats && dtrans_part && B > 0 && K >= 0 && T >= 3 &&
                  T <= VBG_CRF_TMAX,
              "vbg_crf_nll_bwd: bad arguments (3 <= T <= 32)");
  crf_nll_bwd_kernel<<<B, 32, 0, as_stream(stream)>>>(feats, trans, tags, seg_off, T, alpha, dnll, dfeats, dtrans_part);
  return check_launch("vbg_crf_nll_bwd");
}
