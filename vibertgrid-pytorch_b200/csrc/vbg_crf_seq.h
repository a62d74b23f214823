// Linear-chain CRF negative log-likelihood of ONE sequence and its gradient (reference model/crf.py:47-93, :148-152:
// `_forward_alg`, `_score_sentence`, `forward`; gradients = what torch autograd derives from them: posterior marginals
// minus gold counts).
//
// The same source is compiled twice:
//   * by nvcc into crf_nll_{fwd,bwd}_kernel (vbg_crf.cu): one warp per sequence, lane == tag (T <= 32), every
//     `VBG_CRF_FOR` body runs once on its own lane, phases separated by __syncwarp();
//   * by g++ into the CPU harness of tests/test_crf_host.py: `VBG_CRF_FOR` is a plain loop over the lanes and the
//     phase barrier is empty.
// So the arithmetic the GPU runs is checked against the reference on a box without a GPU.  The rule that makes both
// readings equal: inside one phase a lane writes only its own slots (index == lane, row == lane, or a strided
// cooperative copy) and reads only slots written in EARLIER phases.
//
// Numerics: the forward / backward variables of a long sequence grow like 3n, where one fp32 ulp would already be 1e-4 at
// n = 1000.  Both recursions are therefore kept NORMALISED (max subtracted after every step; the forward normalisers are
// summed in double into log Z) and the posteriors are normalised per step, so every exponent is O(1) and the gradient is
// accurate to fp32 round-off at any length (the reference's own fp32 autograd is not: it differentiates through the raw
// log-space variables).
//
// Tags: START = T-2, STOP = T-1 (field_type_classification_head.py:635-637).  trans[i*T + j] = score of moving TO i FROM j.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define VBG_CRF_HD __device__ __forceinline__
#define VBG_CRF_FOR(i, n) for (int i = (int)threadIdx.x; i < (n); i += 32)
#define VBG_CRF_SYNC() __syncwarp()
#else
#define VBG_CRF_HD static inline
#define VBG_CRF_FOR(i, n) for (int i = 0; i < (n); ++i)
#define VBG_CRF_SYNC() ((void)0)
#endif

#define VBG_CRF_TMAX 32
#define VBG_CRF_CHUNK 64

// scratch of one sequence (shared memory on the device)
struct CrfScratch {
  float tr[VBG_CRF_TMAX * VBG_CRF_TMAX];    // trans,   [to][from]
  float trT[VBG_CRF_TMAX * VBG_CRF_TMAX];   // trans^T, [from][to]
  float dtr[VBG_CRF_TMAX * VBG_CRF_TMAX];   // gradient accumulator (bwd)
  float cur[VBG_CRF_TMAX];                  // alpha_{t-1} (fwd) / beta_t (bwd)
  float nxt[VBG_CRF_TMAX];
  float v[VBG_CRF_TMAX];
  float f[VBG_CRF_CHUNK * VBG_CRF_TMAX];          // emissions of the current chunk of steps
  float a[(VBG_CRF_CHUNK + 1) * VBG_CRF_TMAX];    // alphas of the chunk, row 0 = the step before it (bwd)
  float gs[VBG_CRF_CHUNK];                        // gold-path terms of the chunk (fwd)
  int32_t tg[VBG_CRF_CHUNK + 1];                  // gold tags of the chunk, slot 0 = the tag before it
  float pw[VBG_CRF_TMAX * VBG_CRF_TMAX];           // un-normalised pairwise posteriors of one step (bwd)
  float rs[VBG_CRF_TMAX];                         // their row sums
  double gold;                                    // gold path score so far (fwd, lane 0)
  double shift;                                   // sum of the per-step normalisers (fwd, lane 0)
};

// log sum_j exp(v[j] + row[j]), stable form of crf.py:25-29
VBG_CRF_HD float crf_lse_row(const float* v, const float* row, int T) {
  float m = v[0] + row[0];
  for (int j = 1; j < T; ++j) m = fmaxf(m, v[j] + row[j]);
  float s = 0.f;
  for (int j = 0; j < T; ++j) s += expf(v[j] + row[j] - m);
  return m + logf(s);
}

VBG_CRF_HD void crf_load_trans(CrfScratch* S, const float* trans, int T) {
  VBG_CRF_FOR(i, T * T) {
    float x = trans[i];
    S->tr[i] = x;
    S->trT[(i % T) * T + (i / T)] = x;
  }
  VBG_CRF_SYNC();
}

// Forward algorithm + gold path score.  Writes ahat[n, T] = alpha_t - max_i alpha_t (kept for the gradient), *logz,
// *nll = (logZ - gold) / n.
VBG_CRF_HD void crf_nll_fwd_seq(CrfScratch* S, const float* feats, const float* trans, const int32_t* tags, int n, int T,
                                float* ahat, float* logz, float* nll) {
  const int start = T - 2, stop = T - 1;
  crf_load_trans(S, trans, T);
  VBG_CRF_FOR(i, T) S->cur[i] = (i == start) ? 0.f : -10000.f;          // crf.py:51-53
  VBG_CRF_FOR(i, 1) { S->gold = 0.0; S->shift = 0.0; }
  VBG_CRF_SYNC();
  for (int t0 = 0; t0 < n; t0 += VBG_CRF_CHUNK) {
    const int m = (n - t0 < VBG_CRF_CHUNK) ? n - t0 : VBG_CRF_CHUNK;
    VBG_CRF_FOR(i, m * T) S->f[i] = feats[(size_t)t0 * T + i];
    VBG_CRF_FOR(i, m + 1) S->tg[i] = (t0 + i > 0) ? tags[t0 + i - 1] : start;
    VBG_CRF_SYNC();
    VBG_CRF_FOR(t, m) S->gs[t] = S->tr[S->tg[t + 1] * T + S->tg[t]] + S->f[t * T + S->tg[t + 1]];   // crf.py:90-91
    VBG_CRF_SYNC();
    VBG_CRF_FOR(i, 1) {
      double acc = S->gold;
      for (int t = 0; t < m; ++t) acc += (double)S->gs[t];
      S->gold = acc;
    }
    for (int t = 0; t < m; ++t) {
      VBG_CRF_FOR(i, T) S->nxt[i] = crf_lse_row(S->cur, S->tr + i * T, T) + S->f[t * T + i];      // crf.py:59-75
      VBG_CRF_SYNC();
      VBG_CRF_FOR(i, T) {
        float c = S->nxt[0];
        for (int j = 1; j < T; ++j) c = fmaxf(c, S->nxt[j]);
        const float x = S->nxt[i] - c;
        S->cur[i] = x;                                                   // cur is not read in this phase
        ahat[(size_t)(t0 + t) * T + i] = x;
        if (i == 0) S->shift += (double)c;
      }
      VBG_CRF_SYNC();
    }
  }
  VBG_CRF_FOR(i, 1) {                                                    // lane 0: crf.py:76-78, :92
    const double z = S->shift + (double)crf_lse_row(S->cur, S->tr + stop * T, T);
    const double gold = S->gold + (double)S->tr[stop * T + ((n > 0) ? tags[n - 1] : start)];
    *logz = (float)z;
    *nll = (float)((z - gold) / (double)n);
  }
  VBG_CRF_SYNC();
}

// Gradient of  up * nll  w.r.t. the emissions (dfeats[n, T]) and the transitions (dtrans[T, T], this sequence's share).
VBG_CRF_HD void crf_nll_bwd_seq(CrfScratch* S, const float* feats, const float* trans, const int32_t* tags, int n, int T,
                                const float* ahat, float up, float* dfeats, float* dtrans) {
  const int start = T - 2, stop = T - 1;
  const float g = up / (float)n;
  crf_load_trans(S, trans, T);
  VBG_CRF_FOR(i, T * T) S->dtr[i] = 0.f;
  VBG_CRF_FOR(i, T) S->cur[i] = S->tr[stop * T + i];                    // beta_{n-1}[j] = trans[STOP, j]
  VBG_CRF_SYNC();
  if (n > 0) {
    // STOP row: posterior of ending in j = softmax_j(ahat_{n-1}[j] + trans[STOP, j])
    VBG_CRF_FOR(j, T) S->v[j] = ahat[(size_t)(n - 1) * T + j] + S->tr[stop * T + j];
    VBG_CRF_SYNC();
    VBG_CRF_FOR(j, T) {
      float mx = S->v[0];
      for (int q = 1; q < T; ++q) mx = fmaxf(mx, S->v[q]);
      float sum = 0.f;
      for (int q = 0; q < T; ++q) sum += expf(S->v[q] - mx);
      S->dtr[stop * T + j] = expf(S->v[j] - mx) / sum;
    }
    VBG_CRF_SYNC();
  }
  const int nchunk = (n + VBG_CRF_CHUNK - 1) / VBG_CRF_CHUNK;
  for (int c = nchunk - 1; c >= 0; --c) {
    const int t0 = c * VBG_CRF_CHUNK;
    const int m = (n - t0 < VBG_CRF_CHUNK) ? n - t0 : VBG_CRF_CHUNK;
    VBG_CRF_FOR(i, m * T) {
      S->f[i] = feats[(size_t)t0 * T + i];
      S->a[T + i] = ahat[(size_t)t0 * T + i];
    }
    VBG_CRF_FOR(i, m + 1) S->tg[i] = (t0 + i > 0) ? tags[t0 + i - 1] : start;
    VBG_CRF_FOR(i, T) S->a[i] = (t0 > 0) ? ahat[(size_t)(t0 - 1) * T + i] : ((i == start) ? 0.f : -10000.f);
    VBG_CRF_SYNC();
    for (int t = m - 1; t >= 0; --t) {
      const float* ap = S->a + t * T;                                    // (normalised) alpha of the previous step
      // un-normalised pairwise posterior of (tag_{t-1} = j, tag_t = i); cur = normalised beta_t, every exponent <= ~0
      VBG_CRF_FOR(i, T) {
        const float w = S->f[t * T + i] + S->cur[i];
        S->v[i] = w;
        float m2 = ap[0] + S->tr[i * T];
        for (int j = 1; j < T; ++j) m2 = fmaxf(m2, ap[j] + S->tr[i * T + j]);
        S->nxt[i] = m2 + w;                                              // row maximum
      }
      VBG_CRF_SYNC();
      VBG_CRF_FOR(i, T) {
        float mx = S->nxt[0];
        for (int q = 1; q < T; ++q) mx = fmaxf(mx, S->nxt[q]);
        const float w = S->v[i] - mx;
        float sum = 0.f;
        for (int j = 0; j < T; ++j) {
          const float e = expf(ap[j] + S->tr[i * T + j] + w);
          S->pw[i * T + j] = e;
          sum += e;
        }
        S->rs[i] = sum;
      }
      VBG_CRF_SYNC();
      VBG_CRF_FOR(i, T) {
        float tot = 0.f;
        for (int q = 0; q < T; ++q) tot += S->rs[q];
        const float inv = 1.f / tot;
        for (int j = 0; j < T; ++j) S->dtr[i * T + j] += S->pw[i * T + j] * inv;
        // P(tag_t = i) = row sum of the pairwise posterior
        dfeats[(size_t)(t0 + t) * T + i] = g * (S->rs[i] * inv - ((S->tg[t + 1] == i) ? 1.f : 0.f));
      }
      // beta_{t-1}[j] = lse_i(trans[i, j] + f_t[i] + beta_t[i]), re-normalised
      VBG_CRF_FOR(j, T) S->nxt[j] = crf_lse_row(S->v, S->trT + j * T, T);
      VBG_CRF_SYNC();
      VBG_CRF_FOR(j, T) {
        float mx = S->nxt[0];
        for (int q = 1; q < T; ++q) mx = fmaxf(mx, S->nxt[q]);
        S->cur[j] = S->nxt[j] - mx;
      }
      VBG_CRF_SYNC();
    }
    VBG_CRF_FOR(i, 1) {                                                  // lane 0: gold transition counts of the chunk
      for (int t = 0; t < m; ++t) S->dtr[S->tg[t + 1] * T + S->tg[t]] -= 1.f;
    }
    VBG_CRF_SYNC();
  }
  VBG_CRF_FOR(i, 1) S->dtr[stop * T + ((n > 0) ? tags[n - 1] : start)] -= 1.f;
  VBG_CRF_SYNC();
  VBG_CRF_FOR(i, T * T) dtrans[i] = g * S->dtr[i];
  VBG_CRF_SYNC();
}
