// Multi-tensor optimizer steps: ONE launch updates every parameter tensor of a param group.
//
// Replaces the per-step Python loops of torch.optim.SGD / torch.optim.AdamW that the reference's training loop runs
// (train_SROIE.py:217-235 builds them; pipeline/train_val_utils.py:272-284 steps them): SGD with momentum for the CNN / heads,
// AdamW for the parameters whose name contains "bert_model".  Same update rules, operation for operation (torch's
// single-tensor formulas), so optimizer state is interchangeable with torch's ("momentum_buffer", "exp_avg", "exp_avg_sq").
//
// HBM-bound: SGD reads p, g, buf and writes p, buf (20 B / element); AdamW reads p, g, m, v and writes p, m, v (28 B / element).
// The tensors of a group are described by a device table of 6 int64 per tensor {p, g, state1, state2, numel, first_chunk}; a CTA
// owns one chunk of kOptChunk elements, finds its tensor by binary search over first_chunk, and streams 128-bit accesses.
#include "vbg_common.cuh"

namespace vbg {

constexpr int kOptChunk = 8192;       // elements per CTA (256 threads x 8 float4)

struct OptTensor { float* p; const float* g; float* s1; float* s2; long long n; long long first_chunk; };

__device__ __forceinline__ int opt_find(const OptTensor* __restrict__ tab, int n_tensors, long long chunk) {
  int lo = 0, hi = n_tensors - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <bool kAdam>
__global__ void __launch_bounds__(256)
optim_step_kernel(const OptTensor* __restrict__ tab, int n_tensors, float lr, float a, float b, float eps, float wd, float c1, float c2,
                  int first_step, float grad_scale) {
  const int t = opt_find(tab, n_tensors, blockIdx.x);
  const OptTensor T = tab[t];
  const long long base = ((long long)blockIdx.x - T.first_chunk) * kOptChunk;
  const long long end = min(T.n, base + (long long)kOptChunk);
  auto upd = [&](float& p, float g, float& s1, float& s2) {
    g *= grad_scale;
    if (kAdam) {
      // torch.optim.AdamW (single-tensor form): decoupled decay, then the bias-corrected Adam update
      p = p * (1.0f - lr * wd);
      s1 = s1 + (g - s1) * (1.0f - a);                 // exp_avg.lerp_(grad, 1 - beta1)
      s2 = s2 * b + (1.0f - b) * g * g;                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = sqrtf(s2) / c2 + eps;        // c2 = sqrt(1 - beta2^t)
      p = p - (lr / c1) * (s1 / denom);                // c1 = 1 - beta1^t
    } else {
      // torch.optim.SGD: weight decay folded into the gradient, momentum buffer (dampening 0, no Nesterov)
      g = g + wd * p;
      if (a != 0.0f) { s1 = first_step ? g : a * s1 + g; g = s1; }
      p = p - lr * g;
    }
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(T.p) | reinterpret_cast<uintptr_t>(T.g) | reinterpret_cast<uintptr_t>(T.s1) |
                     (kAdam ? reinterpret_cast<uintptr_t>(T.s2) : 0)) & 15) == 0;
  if (vec) {
    const long long e4 = base + ((end - base) & ~3LL);
    for (long long i = base + 4LL * threadIdx.x; i < e4; i += 4LL * blockDim.x) {
      float4 p = *reinterpret_cast<float4*>(T.p + i);
      const float4 g = __ldg(reinterpret_cast<const float4*>(T.g + i));
      float4 s1 = (kAdam || a != 0.0f) ? *reinterpret_cast<float4*>(T.s1 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 s2 = kAdam ? *reinterpret_cast<float4*>(T.s2 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      upd(p.x, g.x, s1.x, s2.x); upd(p.y, g.y, s1.y, s2.y); upd(p.z, g.z, s1.z, s2.z); upd(p.w, g.w, s1.w, s2.w);
      *reinterpret_cast<float4*>(T.p + i) = p;
      if (kAdam || a != 0.0f) *reinterpret_cast<float4*>(T.s1 + i) = s1;
      if (kAdam) *reinterpret_cast<float4*>(T.s2 + i) = s2;
    }
    for (long long i = e4 + threadIdx.x; i < end; i += blockDim.x) {
      float p = T.p[i], s1 = (kAdam || a != 0.0f) ? T.s1[i] : 0.f, s2 = kAdam ? T.s2[i] : 0.f;
      upd(p, T.g[i], s1, s2);
      T.p[i] = p;
      if (kAdam || a != 0.0f) T.s1[i] = s1;
      if (kAdam) T.s2[i] = s2;
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += blockDim.x) {
      float p = T.p[i], s1 = (kAdam || a != 0.0f) ? T.s1[i] : 0.f, s2 = kAdam ? T.s2[i] : 0.f;
      upd(p, T.g[i], s1, s2);
      T.p[i] = p;
      if (kAdam || a != 0.0f) T.s1[i] = s1;
      if (kAdam) T.s2[i] = s2;
    }
  }
}

}  // namespace vbg

using namespace vbg;

extern "C" int vbg_optim_chunk(void) { return kOptChunk; }

extern "C" int vbg_sgd_step_mt(const void* table, int n_tensors, long long total_chunks, float lr, float momentum, float weight_decay,
                               int first_step, float grad_scale, vbg_stream_t stream) {
  VBG_REQUIRE(table && n_tensors > 0 && total_chunks > 0 && total_chunks < (1LL << 31), "vbg_sgd_step_mt: bad arguments");
  optim_step_kernel<false><<<(unsigned)total_chunks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const OptTensor*>(table), n_tensors, lr, momentum,
                                                                                 0.f, 0.f, weight_decay, 1.f, 1.f, first_step, grad_scale);
  return check_launch("vbg_sgd_step_mt");
}

extern "C" int vbg_adamw_step_mt(const void* table, int n_tensors, long long total_chunks, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, float bias_correction1, float sqrt_bias_correction2, float grad_scale,
                                 vbg_stream_t stream) {
  VBG_REQUIRE(table && n_tensors > 0 && total_chunks > 0 && total_chunks < (1LL << 31) && bias_correction1 > 0.f && sqrt_bias_correction2 > 0.f,
              "vbg_adamw_step_mt: bad arguments");
  optim_step_kernel<true><<<(unsigned)total_chunks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const OptTensor*>(table), n_tensors, lr, beta1,
                                                                                beta2, eps, weight_decay, bias_correction1, sqrt_bias_correction2,
                                                                                0, grad_scale);
  return check_launch("vbg_adamw_step_mt");
}
