// TEST-ONLY: compiles vibertgrid-pytorch_b200/csrc/vbg_crf_seq.h (the exact per-sequence source of the CUDA kernels
// crf_nll_{fwd,bwd}_kernel) for the host, so the CPU suite can hold it to the reference's CRF without a GPU.
#include "vbg_crf_seq.h"
#include <stdlib.h>

extern "C" int crf_host_nll_fwd(const float* feats, const float* trans, const int32_t* tags, const int32_t* seg_off, int B, int T,
                                float* alpha, float* logz, float* nll) {
  CrfScratch* S = (CrfScratch*)malloc(sizeof(CrfScratch));
  for (int b = 0; b < B; ++b) {
    int s0 = seg_off[b], n = seg_off[b + 1] - s0;
    crf_nll_fwd_seq(S, feats + (size_t)s0 * T, trans, tags + s0, n, T, alpha + (size_t)s0 * T, logz + b, nll + b);
  }
  free(S);
  return 0;
}

extern "C" int crf_host_nll_bwd(const float* feats, const float* trans, const int32_t* tags, const int32_t* seg_off, int B, int T,
                                const float* alpha, const float* dnll, float* dfeats, float* dtrans_part) {
  CrfScratch* S = (CrfScratch*)malloc(sizeof(CrfScratch));
  for (int b = 0; b < B; ++b) {
    int s0 = seg_off[b], n = seg_off[b + 1] - s0;
    crf_nll_bwd_seq(S, feats + (size_t)s0 * T, trans, tags + s0, n, T, alpha + (size_t)s0 * T, dnll[b],
                    dfeats + (size_t)s0 * T, dtrans_part + (size_t)b * T * T);
  }
  free(S);
  return 0;
}
