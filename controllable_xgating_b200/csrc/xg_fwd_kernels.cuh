// Forward-path kernels that are not dense contractions: BatchNorm statistics/apply, the two LSTM
// cell updates, temporal soft attention, embedding gather, masked mean, row log-softmax, greedy pick.
// All of this is HBM/L2-bandwidth or latency bound integer/float streaming work: coalesced
// (lane == innermost index) accesses, warp-shuffle reductions, no atomics (deterministic sums).
#pragma once
#include "xg_common.cuh"

namespace xg {

// ------------------------------------------------------------------------------------
// BatchNorm1d over the (B*K) rows of the embedded stream (sub_modules.py:97-104,121,126)
// ------------------------------------------------------------------------------------
// partial column sums in double: grid (ceil(H/32), RS), block (32, 8)
__global__ void colstats_partial_kernel(const float* __restrict__ Y, int M, int H, double* __restrict__ part) {
  __shared__ double s1[8][33], s2[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int RS = gridDim.y;
  const int rows_per = (M + RS - 1) / RS;
  const int r0 = blockIdx.y * rows_per;
  const int r1 = min(M, r0 + rows_per);
  double a = 0.0, b = 0.0;
  if (j < H) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      float v = Y[(long)r * H + j];
      a += (double)v;
      b += (double)v * (double)v;
    }
  }
  s1[threadIdx.y][threadIdx.x] = a;
  s2[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.y == 0 && j < H) {
    for (int y = 1; y < 8; ++y) { a += s1[y][threadIdx.x]; b += s2[y][threadIdx.x]; }
    part[((long)blockIdx.y * H + j) * 2 + 0] = a;
    part[((long)blockIdx.y * H + j) * 2 + 1] = b;
  }
}

// train: batch statistics (+ running-stat update, unbiased variance, momentum);  eval: running stats.
// Outputs the folded affine  y = x*scale + shift  and (train) mean / invstd for the backward pass.
__global__ void bn_finalize_kernel(const double* __restrict__ part, int RS, int M, int H, int train,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ run_mean, float* __restrict__ run_var,
                                   float eps, float momentum, int update_running,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  float mean, var;
  if (train) {
    double a = 0.0, b = 0.0;
    for (int s = 0; s < RS; ++s) { a += part[((long)s * H + j) * 2]; b += part[((long)s * H + j) * 2 + 1]; }
    double mu = a / M;
    double vr = b / M - mu * mu;
    if (vr < 0.0) vr = 0.0;
    mean = (float)mu;
    var = (float)vr;
    if (update_running) {
      double unb = M > 1 ? vr * ((double)M / (double)(M - 1)) : vr;
      run_mean[j] = (1.f - momentum) * run_mean[j] + momentum * mean;
      run_var[j] = (1.f - momentum) * run_var[j] + momentum * (float)unb;
    }
  } else {
    mean = run_mean[j];
    var = run_var[j];
  }
  const float invstd = 1.0f / sqrtf(var + eps);
  const float sc = gamma[j] * invstd;
  scale[j] = sc;
  shift[j] = beta[j] - mean * sc;
  if (mean_out) mean_out[j] = mean;
  if (invstd_out) invstd_out[j] = invstd;
}

// E[(k*B+b), j] = relu(Y[(b*K+k), j]*scale + shift) * drop(b,k,j) * fmask[b,k]    (frame-major out)
__global__ void bn_apply_kernel(const float* __restrict__ Y, const float* __restrict__ scale,
                                const float* __restrict__ shift, const float* __restrict__ fmask,
                                int B, int K, int H, DropSpec drop, float* __restrict__ E) {
  const long n = (long)B * K * H;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % H);
    const long r = e / H;  // b*K + k
    const int k = (int)(r % K);
    const int b = (int)(r / K);
    float v = Y[e] * scale[j] + shift[j];
    v = v > 0.f ? v : 0.f;
    v *= drop.factor((uint64_t)e);
    v *= fmask[r];
    E[((long)k * B + b) * H + j] = v;
  }
}

// ------------------------------------------------------------------------------------
// encoder nn.LSTMCell pointwise part (gate order i,f,g,o) + zeroing frame mask
// (sub_modules.py:138-140,145-147).  Z (B,4H) holds x.W_ih^T + h.W_hh^T + biases and is
// overwritten with the activated gates (kept for the backward pass).
// ------------------------------------------------------------------------------------
__global__ void enc_cell_kernel(float* __restrict__ Z, const float* __restrict__ c_prev,
                                const float* __restrict__ fmask, int mask_stride, int t, int B, int H,
                                float* __restrict__ c_out, float* __restrict__ h_out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * H) return;
  const int b = e / H, j = e % H;
  float* z = Z + (long)b * 4 * H;
  const float i = sigmoid_f(z[j]);
  const float f = sigmoid_f(z[H + j]);
  const float g = tanhf(z[2 * H + j]);
  const float o = sigmoid_f(z[3 * H + j]);
  const float m = fmask[(long)b * mask_stride + t];
  const float cp = c_prev ? c_prev[e] : 0.f;
  const float c2 = f * cp + i * g;
  const float h = o * tanhf(c2) * m;      // h' = o*tanh(c2) ; h' *= mask   (:138-139)
  const float c = c2 * m;                 // c2 *= mask                     (:140)
  z[j] = i; z[H + j] = f; z[2 * H + j] = g; z[3 * H + j] = o;
  c_out[e] = c;
  h_out[e] = h;
}

// ------------------------------------------------------------------------------------
// decoder two_inputs_lstmcell pointwise part (gate order i,f,o,g; the mask CARRIES the state;
// dropout on the carried h)  (sub_modules.py:753-767)
// ------------------------------------------------------------------------------------
__global__ void dec_cell_kernel(float* __restrict__ Z, const float* __restrict__ c_prev,
                                const float* __restrict__ h_prev, long ld_hprev,
                                const float* __restrict__ mask, long mask_stride, int B, int H, DropSpec drop,
                                float* __restrict__ c_out, float* __restrict__ h_out, long ld_hout,
                                float* __restrict__ h_out2, long ld_hout2) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * H) return;
  const int b = e / H, j = e % H;
  float* z = Z + (long)b * 4 * H;
  const float i = sigmoid_f(z[j]);
  const float f = sigmoid_f(z[H + j]);
  const float o = sigmoid_f(z[2 * H + j]);
  const float g = tanhf(z[3 * H + j]);
  const float m = mask ? mask[(long)b * mask_stride] : 1.f;
  const float cp = c_prev[e];
  const float hp = h_prev[(long)b * ld_hprev + j];
  float c = f * cp + i * g;
  c = c * m + cp * (1.f - m);
  float h = o * tanhf(c);
  h = h * m + hp * (1.f - m);
  h *= drop.factor((uint64_t)e);
  z[j] = i; z[H + j] = f; z[2 * H + j] = o; z[3 * H + j] = g;
  c_out[e] = c;
  h_out[(long)b * ld_hout + j] = h;
  if (h_out2) h_out2[(long)b * ld_hout2 + j] = h;
}

// ------------------------------------------------------------------------------------
// embedding gather: out[r, :] = emb[tok(r), :],  r = i*B + b,  tok(r) = tokens[b*sb + i*si]
// ------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ emb, const int64_t* __restrict__ tokens,
                                   long sb, long si, int B, int rows, int E, int V, float* __restrict__ out) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int b = r % B, i = r / B;
  long t = tokens[(long)b * sb + (long)i * si];
  if (t < 0) t = 0;
  if (t >= V) t = V - 1;
  const float* src = emb + t * E;
  for (int j = threadIdx.x; j < E; j += blockDim.x) out[(long)r * E + j] = src[j];
}

// ------------------------------------------------------------------------------------
// init_hidden mean: mean[b,:] = sum_k V[b,k,:] / sum_k fmask[b,k]   (SAModel.py:59-61)
// ------------------------------------------------------------------------------------
__global__ void masked_mean_kernel(const float* __restrict__ V, const float* __restrict__ fmask, int B, int K, int H,
                                   float* __restrict__ mean) {
  const int b = blockIdx.x;
  float cnt = 0.f;
  for (int k = 0; k < K; ++k) cnt += fmask[(long)b * K + k];
  for (int j = blockIdx.y * blockDim.x + threadIdx.x; j < H; j += gridDim.y * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += V[((long)b * K + k) * H + j];
    mean[(long)b * H + j] = s / cnt;
  }
}

// The four init-state linears of SAModel.init_hidden (SAModel.py:62-65) in ONE launch: out_q[b, n] = mean[b, :] . W_q[n, :] + b_q[n].
// blockIdx.y = q; a CTA of 8 warps takes 16 rows of W_q (two per warp, a row in registers), mean (B x H) sits in shared
// memory in slabs of IS_CAPS captions.  H <= 32 * IS_KPL.
constexpr int IS_KPL = 16;       // K elements per lane
constexpr int IS_CAPS = 64;      // captions per shared-memory slab
struct InitStateArgs { const float* W[4]; const float* bias[4]; float* out[4]; long ld[4]; };
__global__ void __launch_bounds__(256) init_state_kernel(const float* __restrict__ mean, int B, int H, const InitStateArgs a) {
  extern __shared__ float is_mv[];                       // [IS_CAPS][H]
  const int q = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 16 + warp * 2;
  float w[2][IS_KPL];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < IS_KPL; ++i) {
      const int k = lane + 32 * i;
      w[r][i] = (n0 + r < H && k < H) ? __ldg(a.W[q] + (long)(n0 + r) * H + k) : 0.f;
    }
  const float b0 = n0 < H ? __ldg(a.bias[q] + n0) : 0.f, b1 = n0 + 1 < H ? __ldg(a.bias[q] + n0 + 1) : 0.f;
  for (int c0 = 0; c0 < B; c0 += IS_CAPS) {
    const int nc = min(IS_CAPS, B - c0);
    __syncthreads();
    if ((H & 3) == 0) {          // (rows of mean are 16-byte aligned: float4 staging, every load of a thread in flight)
      const float4* src = reinterpret_cast<const float4*>(mean + (long)c0 * H);
      float4* dst = reinterpret_cast<float4*>(is_mv);
      const int n4 = nc * H / 4;
#pragma unroll 4
      for (int e = threadIdx.x; e < n4; e += 256) dst[e] = __ldg(src + e);
    } else {
      for (int e = threadIdx.x; e < nc * H; e += 256) is_mv[e] = mean[(long)c0 * H + e];
    }
    __syncthreads();
    // lane c of a pass keeps the sums of caption cbase + c: 32 captions per pass, one reduction tree per caption and row
    for (int cbase = 0; cbase < nc; cbase += 32) {
      float keep0 = 0.f, keep1 = 0.f;
#pragma unroll 2
      for (int cc = 0; cc < 32 && cbase + cc < nc; ++cc) {
        const int c = cbase + cc;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < IS_KPL; ++i) {
          const int k = lane + 32 * i;
          const float m = k < H ? is_mv[c * H + k] : 0.f;
          s0 = fmaf(w[0][i], m, s0); s1 = fmaf(w[1][i], m, s1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        if (lane == cc) { keep0 = s0; keep1 = s1; }
      }
      const int c = cbase + lane;
      if (c < nc) {
        if (n0 < H) a.out[q][(long)(c0 + c) * a.ld[q] + n0] = keep0 + b0;
        if (n0 + 1 < H) a.out[q][(long)(c0 + c) * a.ld[q] + n0 + 1] = keep1 + b1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// temporal soft attention for one word step (sub_modules.py:677-680), one CTA per caption:
//   s_k = a2w . tanh(AH[b] + Uv[b,k,:]) + b ;  alpha = softmax_k(s) (NO frame mask) ;
//   af = sum_k alpha_k V[b,k,:]
// dynamic smem: A + K floats.
// ------------------------------------------------------------------------------------
__global__ void att_fwd_kernel(const float* __restrict__ AH, const float* __restrict__ Uv,
                               const float* __restrict__ V, const float* __restrict__ wa,
                               const float* __restrict__ ba, int K, int A, int H, int feat_div,
                               float* __restrict__ alpha_out, float* __restrict__ af_out) {
  // One CTA per state row.  Loads are issued in batches of 8 before their first use (the loops were one
  // dependent L2 round trip per element: 35 us per launch, latency only); tanh through ex2/rcp (~1e-7 abs).
  extern __shared__ float sm[];
  float* ah = sm;        // A
  float* sc = sm + A;    // K
  const int b = blockIdx.x;            // state row
  const int fb = b / feat_div;         // feature row (beam rows of one video share V / Uv)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int a = threadIdx.x; a < A; a += blockDim.x) ah[a] = AH[(long)b * A + a];
  __syncthreads();
  const float b0 = ba[0];
  for (int k = warp; k < K; k += nwarp) {
    const float* u = Uv + ((long)fb * K + k) * A;
    float p = 0.f;
    for (int a0 = lane; a0 < A; a0 += 32 * 8) {
      float uu[8], ww[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int a = min(a0 + 32 * q, A - 1);
        uu[q] = __ldg(u + a);
        ww[q] = (a0 + 32 * q < A) ? __ldg(wa + a) : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float x = ah[min(a0 + 32 * q, A - 1)] + uu[q];
        p += ww[q] * (1.f - __fdividef(2.f, 1.f + __expf(2.f * x)));
      }
    }
    p = warp_sum(p);
    if (lane == 0) sc[k] = p + b0;
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, sc[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < K; k += 32) { float e = expf(sc[k] - mx); sc[k] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int k = lane; k < K; k += 32) sc[k] *= inv;
  }
  __syncthreads();
  if (alpha_out)
    for (int k = threadIdx.x; k < K; k += blockDim.x) alpha_out[(long)b * K + k] = sc[k];
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const float* v = V + (long)fb * K * H + j;
    float s = 0.f;
    for (int k0 = 0; k0 < K; k0 += 8) {
      float vv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) vv[q] = __ldg(v + (long)min(k0 + q, K - 1) * H);
#pragma unroll
      for (int q = 0; q < 8; ++q) s += (k0 + q < K ? sc[k0 + q] : 0.f) * vv[q];
    }
    af_out[(long)b * H + j] = s;
  }
}

// ------------------------------------------------------------------------------------
// block reductions (value / value+index) for the vocabulary-wide row kernels
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < nw; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < nw; ++w) r += red[w];   // fixed order: deterministic
  return r;
}

// row log-softmax: out[orow(r), :] = x - max - log(sum exp(x - max));  one CTA per row.
// orow(r) = perm_rb ? (r % perm_rb) * perm_rs + r / perm_rb : r.   In-place safe when no permutation.
__global__ void logsoftmax_rows_kernel(const float* __restrict__ X, long ldx, int N, int perm_rb, int perm_rs,
                                       float* __restrict__ out, long ldo) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float* x = X + (long)r * ldx;
  const long orow = perm_rb ? (long)(r % perm_rb) * perm_rs + r / perm_rb : (long)r;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) mx = fmaxf(mx, x[j]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) s += expf(x[j] - mx);
  s = block_sum(s, red);
  const float lse = mx + logf(s);
  float* o = out + orow * ldo;
  for (int j = threadIdx.x; j < N; j += blockDim.x) o[j] = x[j] - lse;
}

// the same with the row held in registers (one read of X instead of three): 256 threads, N a multiple of 4 and at most
// 256 * 4 * LSM_V4 entries, 16-byte aligned rows
constexpr int LSM_V4 = 12;
__global__ void __launch_bounds__(256)
logsoftmax_rows_reg_kernel(const float* __restrict__ X, long ldx, int N, int perm_rb, int perm_rs, float* __restrict__ out, long ldo) {
  __shared__ float red[32];
  const int r = blockIdx.x, n4 = N >> 2;
  const float4* x = reinterpret_cast<const float4*>(X + (long)r * ldx);
  const long orow = perm_rb ? (long)(r % perm_rb) * perm_rs + r / perm_rb : (long)r;
  float4 v[LSM_V4];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < LSM_V4; ++i) {
    const int j = threadIdx.x + 256 * i;
    v[i] = j < n4 ? x[j] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  }
  mx = block_max(mx, red);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LSM_V4; ++i) s += (expf(v[i].x - mx) + expf(v[i].y - mx)) + (expf(v[i].z - mx) + expf(v[i].w - mx));
  s = block_sum(s, red);
  const float lse = mx + logf(s);
  float4* o = reinterpret_cast<float4*>(out + orow * ldo);
#pragma unroll
  for (int i = 0; i < LSM_V4; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < n4) o[j] = make_float4(v[i].x - lse, v[i].y - lse, v[i].z - lse, v[i].w - lse);
  }
}
static inline bool logsoftmax_reg_ok(const void* X, long ldx, int N, const void* out, long ldo) {
  return N % 4 == 0 && N <= 256 * 4 * LSM_V4 && ldx % 4 == 0 && ldo % 4 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)out & 15) == 0;
}

// ------------------------------------------------------------------------------------
// greedy / multinomial pick for SAModel.sample (SAModel.py:185-210), one CTA per caption row.
// Reads the step's logits (not log-probs: only max, argmax and logsumexp are needed for greedy),
// updates the bookkeeping the reference keeps on the host:
//   it = argmax (lowest index on ties, as torch.max) ; unfinished &= it > 0 ;
//   seq[b,t-1] = it * unfinished ; logp[b,t-1] = max logp ; next token = RAW it ; step mask = unfinished.
// flags[t-1] is set if any row is still unfinished (the reference's loop break at :206).
// ------------------------------------------------------------------------------------
__global__ void greedy_pick_kernel(const float* __restrict__ logits, int V, int t /* >= 1 */, int T,
                                   int sample_max, float inv_temperature, uint64_t seed,
                                   int64_t* __restrict__ seq, float* __restrict__ seqlogp,
                                   int64_t* __restrict__ next_tok, float* __restrict__ unfinished,
                                   int* __restrict__ flags) {
  __shared__ float red[32];
  __shared__ int redi[32];
  __shared__ float s_pick;
  const int b = blockIdx.x;
  const float* x = logits + (long)b * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // argmax with lowest-index tie break
  float best = -INFINITY; int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < V; j += blockDim.x) {
    float v = x[j];
    if (v > best) { best = v; bi = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) { red[warp] = best; redi[warp] = bi; }
  __syncthreads();
  best = red[0]; bi = redi[0];
  for (int w = 1; w < nw; ++w)
    if (red[w] > best || (red[w] == best && redi[w] < bi)) { best = red[w]; bi = redi[w]; }
  const float mx = best;
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += expf(x[j] - mx);
  s = block_sum(s, red);
  const float lse = mx + logf(s);
  int pick = bi;
  float pick_logp = mx - lse;
  if (!sample_max) {
    // multinomial over p ~ exp(logp / temperature): inverse-CDF with one Philox uniform per (row, step).
    // sequential scan by one warp in fixed order (deterministic for a given seed).
    float tot = 0.f;
    for (int j = threadIdx.x; j < V; j += blockDim.x) tot += expf((x[j] - lse) * inv_temperature);
    tot = block_sum(tot, red);
    if (threadIdx.x == 0) {
      const float u = philox_uniform(seed, 0x5a4d0000u + (uint32_t)t, (uint64_t)b) * tot;
      float acc = 0.f; int sel = V - 1;
      for (int j = 0; j < V; ++j) {
        acc += expf((x[j] - lse) * inv_temperature);
        if (acc > u) { sel = j; break; }
      }
      redi[0] = sel;
      s_pick = x[sel] - lse;
    }
    __syncthreads();
    pick = redi[0];
    pick_logp = s_pick;
  }
  if (threadIdx.x == 0) {
    float unf = (t == 1) ? 1.f : unfinished[b];
    unf = (unf != 0.f && pick > 0) ? 1.f : 0.f;
    unfinished[b] = unf;
    seq[(long)b * T + (t - 1)] = unf != 0.f ? (int64_t)pick : 0;
    seqlogp[(long)b * T + (t - 1)] = pick_logp;
    next_tok[b] = (int64_t)pick;
    if (unf != 0.f) flags[t - 1] = 1;   // benign race: every writer stores the same value
  }
}

// scheduled sampling (SAModel.py:89-99): with probability ss_prob the input token of step i is drawn from the
// previous step's word distribution exp(log_softmax(logits)), else it is the ground-truth token seq[b, i].
// One CTA per caption; one Philox uniform decides, a second one drives the inverse-CDF scan (fixed order).
__global__ void ss_pick_kernel(const float* __restrict__ logits, int V, int i, int L, float ss_prob, uint64_t seed,
                               const int64_t* __restrict__ seq, int64_t* __restrict__ tok, int64_t* __restrict__ used) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int64_t gt = seq[(long)b * L + i];
  const float coin = philox_uniform(seed, 0x53530000u + (uint32_t)i, (uint64_t)b);
  if (coin >= ss_prob) {           // uniform for the whole CTA
    if (threadIdx.x == 0) { tok[b] = gt; used[(long)b * L + i] = gt; }
    return;
  }
  const float* x = logits + (long)b * V;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < V; j += blockDim.x) mx = fmaxf(mx, x[j]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int j = threadIdx.x; j < V; j += blockDim.x) s += expf(x[j] - mx);
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float u = philox_uniform(seed, 0x53540000u + (uint32_t)i, (uint64_t)b) * s;
    float acc = 0.f; int sel = V - 1;
    for (int j = 0; j < V; ++j) {
      acc += expf(x[j] - mx);
      if (acc > u) { sel = j; break; }
    }
    tok[b] = (int64_t)sel;
    used[(long)b * L + i] = (int64_t)sel;
  }
}

__global__ void fill_kernel(float* __restrict__ p, long n, float v) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) p[e] = v;
}

}  // namespace xg
