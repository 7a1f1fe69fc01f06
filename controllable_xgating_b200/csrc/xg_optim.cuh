// Fused optimizer step that follows the hot path in every training iteration of the reference
// (starttrain.py:134-137): elementwise gradient clamp (myutils.clip_gradient, myutils.py:79-85) + Adam
// (optim.Adam(lr, weight_decay), starttrain.py:76) over all parameter tensors in ONE launch.
// HBM-streaming work: 4 reads + 3 writes of 4 bytes per element (28 B/element, 26.3 M elements at V=10k).
#pragma once
#include <math.h>

#include "xg_common.cuh"

namespace xg {

constexpr int ADAM_MAX_TENSORS = 64;
constexpr int ADAM_CHUNK = 4096;       // elements per CTA (256 threads x 4 x 4)

struct AdamArgs {
  float* param[ADAM_MAX_TENSORS];
  const float* grad[ADAM_MAX_TENSORS];
  float* m[ADAM_MAX_TENSORS];
  float* v[ADAM_MAX_TENSORS];
  long n[ADAM_MAX_TENSORS];
  int chunk_start[ADAM_MAX_TENSORS + 1];   // prefix sum of ceil(n / ADAM_CHUNK)
  int count;
  float lr, beta1, beta2, eps, weight_decay, grad_clip;
  float bias1, bias2_sqrt;                 // 1 - beta1^step,  sqrt(1 - beta2^step)
  int eps_mode;                            // 0: PyTorch 0.3.1 (denom = sqrt(v) + eps, step = lr sqrt(bc2)/bc1)
                                           // 1: PyTorch >= 1.0 (denom = sqrt(v)/sqrt(bc2) + eps, step = lr/bc1)
  int write_clamped_grad;                  // clip_gradient clamps .grad in place: keep that visible
};

__global__ void __launch_bounds__(256) adam_step_kernel(const __grid_constant__ AdamArgs a) {
  __shared__ int s_t;
  if (threadIdx.x == 0) {
    int lo = 0, hi = a.count - 1;
    while (lo < hi) {                                   // last tensor whose first chunk is <= blockIdx.x
      const int mid = (lo + hi + 1) >> 1;
      if (a.chunk_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    s_t = lo;
  }
  __syncthreads();
  const int t = s_t;
  const long base = (long)(blockIdx.x - a.chunk_start[t]) * ADAM_CHUNK;
  const long n = a.n[t];
  float* __restrict__ p = a.param[t];
  const float* __restrict__ g = a.grad[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  const float step0 = a.eps_mode == 0 ? a.lr * a.bias2_sqrt / a.bias1 : a.lr / a.bias1;
  const float inv_b2s = 1.f / a.bias2_sqrt;
#pragma unroll
  for (int q = 0; q < ADAM_CHUNK / 256; ++q) {
    const long i = base + q * 256 + threadIdx.x;
    if (i >= n) break;
    float gi = g[i];
    if (a.grad_clip > 0.f) gi = fminf(fmaxf(gi, -a.grad_clip), a.grad_clip);
    if (a.write_clamped_grad && a.grad_clip > 0.f) const_cast<float*>(g)[i] = gi;
    const float pi = p[i];
    if (a.weight_decay != 0.f) gi = fmaf(a.weight_decay, pi, gi);
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * gi;
    const float vi = a.beta2 * v[i] + (1.f - a.beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = a.eps_mode == 0 ? sqrtf(vi) + a.eps : sqrtf(vi) * inv_b2s + a.eps;
    p[i] = pi - step0 * (mi / denom);
  }
}

static int adam_step(ErrorSink& es, const xg_adam_tensor* tensors, int count, int step, float lr, float beta1, float beta2,
                     float eps, float weight_decay, float grad_clip, int eps_mode, int write_clamped_grad, cudaStream_t st) {
  XG_REQUIRE(es, tensors != nullptr, XG_ERR_NULL_POINTER, "xg_adam_step: null tensor table");
  XG_REQUIRE(es, count >= 1 && count <= ADAM_MAX_TENSORS, XG_ERR_BAD_ARG, "xg_adam_step: 1..64 tensors per call");
  XG_REQUIRE(es, step >= 1, XG_ERR_BAD_ARG, "xg_adam_step: step is 1-based");
  XG_REQUIRE(es, lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, XG_ERR_BAD_ARG,
             "xg_adam_step: bad hyper-parameters");
  AdamArgs a;
  a.count = count;
  int chunks = 0;
  for (int t = 0; t < count; ++t) {
    XG_REQUIRE(es, tensors[t].param && tensors[t].grad && tensors[t].exp_avg && tensors[t].exp_avg_sq, XG_ERR_NULL_POINTER,
               "xg_adam_step: null tensor pointer");
    XG_REQUIRE(es, tensors[t].n >= 0, XG_ERR_BAD_SHAPE, "xg_adam_step: negative element count");
    a.param[t] = tensors[t].param; a.grad[t] = tensors[t].grad; a.m[t] = tensors[t].exp_avg; a.v[t] = tensors[t].exp_avg_sq;
    a.n[t] = tensors[t].n;
    a.chunk_start[t] = chunks;
    chunks += (int)((tensors[t].n + ADAM_CHUNK - 1) / ADAM_CHUNK);
  }
  a.chunk_start[count] = chunks;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.grad_clip = grad_clip;
  a.bias1 = (float)(1.0 - pow((double)beta1, (double)step));
  a.bias2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  a.eps_mode = eps_mode;
  a.write_clamped_grad = write_clamped_grad;
  if (chunks == 0) return XG_OK;
  adam_step_kernel<<<chunks, 256, 0, st>>>(a);
  XG_LAUNCH_CHECK(es);
  return XG_OK;
}

}  // namespace xg
