// Backward-path kernels (hand-derived BPTT; the reference's backward is torch autograd,
// starttrain.py:134).  The derivation is restated in plain torch in oracle/manual_bptt.py and
// checked against autograd there; these kernels follow it step for step.
#pragma once
#include "xg_common.cuh"
#include "xg_fwd_kernels.cuh"

namespace xg {

// dlogits[r,:] = dlp[src(r),:] - exp(lp[src(r),:]) * sum_j dlp[src(r),j];  r = i*B + b (step-major),
// src(r) = b*Lp + i (the API's batch-major (B,Lp,N) layout).  One CTA per row.
__global__ void logsoftmax_bwd_rows_kernel(const float* __restrict__ lp, const float* __restrict__ dlp,
                                           int B, int Lp, int N, float* __restrict__ dlogits) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int b = r % B, i = r / B;
  const long src = (long)b * Lp + i;
  const float* d = dlp + src * N;
  const float* l = lp + src * N;
  float s = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) s += d[j];
  s = block_sum(s, red);
  float* o = dlogits + (long)r * N;
  for (int j = threadIdx.x; j < N; j += blockDim.x) o[j] = d[j] - expf(l[j]) * s;
}

// the same with both rows held in registers (one read of dlp instead of two); conditions of logsoftmax_reg_ok
__global__ void __launch_bounds__(256)
logsoftmax_bwd_rows_reg_kernel(const float* __restrict__ lp, const float* __restrict__ dlp, int B, int Lp, int N, float* __restrict__ dlogits) {
  __shared__ float red[32];
  const int r = blockIdx.x, n4 = N >> 2;
  const int b = r % B, i = r / B;
  const long src = (long)b * Lp + i;
  const float4* d = reinterpret_cast<const float4*>(dlp + src * N);
  const float4* l = reinterpret_cast<const float4*>(lp + src * N);
  float4 dv[LSM_V4], lv[LSM_V4];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < LSM_V4; ++q) {
    const int j = threadIdx.x + 256 * q;
    dv[q] = j < n4 ? d[j] : make_float4(0.f, 0.f, 0.f, 0.f);
    lv[q] = j < n4 ? l[j] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int q = 0; q < LSM_V4; ++q) s += (dv[q].x + dv[q].y) + (dv[q].z + dv[q].w);
  s = block_sum(s, red);
  float4* o = reinterpret_cast<float4*>(dlogits + (long)r * N);
#pragma unroll
  for (int q = 0; q < LSM_V4; ++q) {
    const int j = threadIdx.x + 256 * q;
    if (j < n4) o[j] = make_float4(dv[q].x - expf(lv[q].x) * s, dv[q].y - expf(lv[q].y) * s, dv[q].z - expf(lv[q].z) * s, dv[q].w - expf(lv[q].w) * s);
  }
}

// out_x[j] = beta*out_x[j] + sum_r X[r*ld + j]   (up to three identical outputs: the three LSTM biases
// of a cell share one column sum).  grid (ceil(N/32), S), block (32,8).  Narrow matrices would occupy a
// handful of SMs with one block per 32 columns, so the rows are cut into S slabs: every block leaves its
// partial sums in `part` [S][N] and the last block of a column group to finish (ticket counter, left at zero
// again) adds the S partials in slab order -- the result does not depend on which block that is.
__device__ __forceinline__ void colsum_body(const float* __restrict__ X, long ld, int R, int N, float beta,
                                            float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2,
                                            float* __restrict__ part, unsigned int* __restrict__ ctr, int bx, int by, int S) {
  __shared__ float sm[8][33];
  __shared__ unsigned int ticket;
  const int j = bx * 32 + threadIdx.x;
  const int chunk = (R + S - 1) / S;
  const int r0 = by * chunk, r1 = min(R, r0 + chunk);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (j < N) {
    int r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      a0 += X[(long)r * ld + j]; a1 += X[(long)(r + 8) * ld + j];
      a2 += X[(long)(r + 16) * ld + j]; a3 += X[(long)(r + 24) * ld + j];
    }
    for (; r < r1; r += 8) a0 += X[(long)r * ld + j];
  }
  float a = (a0 + a1) + (a2 + a3);
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int y = 1; y < 8; ++y) a += sm[y][threadIdx.x];
    if (S > 1 && j < N) __stcg(part + (long)by * N + j, a);
  }
  if (S > 1) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) ticket = atomicAdd(ctr + bx, 1u);
    __syncthreads();
    if (ticket != (unsigned int)(S - 1)) return;
    __threadfence();
    if (threadIdx.y == 0 && j < N) {
      a = 0.f;
      for (int s2 = 0; s2 < S; ++s2) a += __ldcg(part + (long)s2 * N + j);
    }
    if (threadIdx.x == 0 && threadIdx.y == 0) ctr[bx] = 0u;
  }
  if (threadIdx.y == 0 && j < N) {
    if (o0) o0[j] = (beta != 0.f ? beta * o0[j] : 0.f) + a;
    if (o1) o1[j] = (beta != 0.f ? beta * o1[j] : 0.f) + a;
    if (o2) o2[j] = (beta != 0.f ? beta * o2[j] : 0.f) + a;
  }
}

__global__ void colsum_kernel(const float* __restrict__ X, long ld, int R, int N, float beta,
                              float* __restrict__ o0, float* __restrict__ o1, float* __restrict__ o2,
                              float* __restrict__ part, unsigned int* __restrict__ ctr) {
  colsum_body(X, ld, R, N, beta, o0, o1, o2, part, ctr, blockIdx.x, blockIdx.y, gridDim.y);
}

// several column sums in ONE launch (the bias gradients of a backward pass): block -> (job, column group, row slab)
constexpr int COLSUM_MAX_JOBS = 16;
struct ColsumJobs {
  const float* X[COLSUM_MAX_JOBS]; long ld[COLSUM_MAX_JOBS]; int R[COLSUM_MAX_JOBS], N[COLSUM_MAX_JOBS], S[COLSUM_MAX_JOBS];
  float* o[COLSUM_MAX_JOBS][3];
  int first_block[COLSUM_MAX_JOBS + 1], part_off[COLSUM_MAX_JOBS], ctr_off[COLSUM_MAX_JOBS];
  int n; float beta;
};
__global__ void colsum_multi_kernel(const ColsumJobs J, float* __restrict__ part, unsigned int* __restrict__ ctr) {
  int q = 0;
  while (q + 1 < J.n && (int)blockIdx.x >= J.first_block[q + 1]) ++q;
  const int local = (int)blockIdx.x - J.first_block[q], gx = (J.N[q] + 31) / 32;
  colsum_body(J.X[q], J.ld[q], J.R[q], J.N[q], J.beta, J.o[q][0], J.o[q][1], J.o[q][2], part + J.part_off[q], ctr + J.ctr_off[q],
              local % gx, local / gx, J.S[q]);
}

// decoder cell backward (mirror of dec_cell_kernel).  G holds activated gates (i,f,o,g) and is
// overwritten with dz.  dh_out = dha + dhb (dhb optional).  dc (B,H) in: carried dc from the future,
// out: dc w.r.t. the previous cell state.  dh_direct (ld) out: the (1-m) pass-through part of dh.
__global__ void dec_cell_bwd_kernel(float* __restrict__ G, const float* __restrict__ c_new,
                                    const float* __restrict__ c_prev, const float* __restrict__ dha, long ld_dha,
                                    const float* __restrict__ dhb, long ld_dhb, float* __restrict__ dc,
                                    const float* __restrict__ mask, long mask_stride, int B, int H, DropSpec drop,
                                    float* __restrict__ dh_direct, long ld_dhd) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * H) return;
  const int b = e / H, j = e % H;
  float* g4 = G + (long)b * 4 * H;
  const float i = g4[j], f = g4[H + j], o = g4[2 * H + j], g = g4[3 * H + j];
  const float m = mask ? mask[(long)b * mask_stride] : 1.f;
  float dh = dha[(long)b * ld_dha + j];
  if (dhb) dh += dhb[(long)b * ld_dhb + j];
  const float dhd = dh * drop.factor((uint64_t)e);
  const float dh_t = dhd * m;
  const float tc = tanhf(c_new[e]);
  const float d_o = dh_t * tc;
  const float dcn = dc[e] + dh_t * o * (1.f - tc * tc);
  const float dc_t = dcn * m;
  const float cp = c_prev[e];
  dc[e] = dcn * (1.f - m) + dc_t * f;
  const float di = dc_t * g, dg = dc_t * i, df = dc_t * cp;
  g4[j] = di * i * (1.f - i);
  g4[H + j] = df * f * (1.f - f);
  g4[2 * H + j] = d_o * o * (1.f - o);
  g4[3 * H + j] = dg * (1.f - g * g);
  dh_direct[(long)b * ld_dhd + j] = dhd * (1.f - m);
}

// attention backward for one word step, one CTA per caption (see manual_bptt.backward):
//   dal_k = dAF . V_k ; dV_k += alpha_k dAF ; ds = alpha*(dal - sum alpha*dal) ;
//   dpre[k,a] = ds_k * wa[a] * (1 - tanh^2(AH[a] + Uv[k,a])) ; dUv += dpre ; dAH[a] = sum_k dpre ;
//   dwa_part[b,a] += sum_k ds_k * tanh(.) ; dba_part[b] += sum_k ds_k
// dynamic smem: A + 2K + H floats.
constexpr int ATT_BWD_SPLITS = 4;
__global__ void att_bwd_kernel(const float* __restrict__ dAF, const float* __restrict__ AH,
                               const float* __restrict__ Uv, const float* __restrict__ V,
                               const float* __restrict__ wa, const float* __restrict__ alpha,
                               int K, int A, int H, float* __restrict__ dV, float* __restrict__ dUv,
                               float* __restrict__ dAH, float* __restrict__ dwa_part, float* __restrict__ dba_part) {
  // grid (B, S): every CTA of a caption recomputes the K softmax-gradient terms (K dot products with V[b],
  // cheap), then takes 1/S of the frames for the dV update and 1/S of the attention units for the
  // dUv / dAH / dwa terms.  Loads are batched ahead of their first use (was 75 us per launch on 64 CTAs).
  extern __shared__ float sm[];
  float* al = sm;            // K
  float* ds = al + K;        // K
  float* daf = ds + K;       // H
  const int b = blockIdx.x, sp = blockIdx.y, S = gridDim.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int k = threadIdx.x; k < K; k += blockDim.x) al[k] = alpha[(long)b * K + k];
  for (int j = threadIdx.x; j < H; j += blockDim.x) daf[j] = dAF[(long)b * H + j];
  __syncthreads();
  for (int k = warp; k < K; k += nwarp) {
    const float* v = V + ((long)b * K + k) * H;
    float p = 0.f;
    for (int j0 = lane; j0 < H; j0 += 32 * 8) {
      float vv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) vv[q] = __ldg(v + min(j0 + 32 * q, H - 1));
#pragma unroll
      for (int q = 0; q < 8; ++q) p += (j0 + 32 * q < H ? daf[min(j0 + 32 * q, H - 1)] : 0.f) * vv[q];
    }
    p = warp_sum(p);
    if (lane == 0) ds[k] = p;   // dal_k for now
  }
  __syncthreads();
  {   // dV[b, k, :] += alpha_k * dAF[b, :]   for this CTA's frames
    const int kper = (K + S - 1) / S, kb0 = sp * kper, kb1 = min(K, kb0 + kper);
    const int n = max(0, kb1 - kb0) * H;
    float* dv = dV + ((long)b * K + kb0) * H;
    for (int e0 = threadIdx.x; e0 < n; e0 += blockDim.x * 8) {
      float old[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) old[q] = dv[min(e0 + (int)blockDim.x * q, n - 1)];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = e0 + (int)blockDim.x * q;
        if (e < n) dv[e] = old[q] + al[kb0 + e / H] * daf[e % H];
      }
    }
  }
  float dot = 0.f;
  for (int k = 0; k < K; ++k) dot += al[k] * ds[k];   // every thread computes the same fixed-order sum
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) ds[k] = al[k] * (ds[k] - dot);
  __syncthreads();
  const int aper = (A + S - 1) / S, a0 = sp * aper, a1 = min(A, a0 + aper);
  for (int a = a0 + threadIdx.x; a < a1; a += blockDim.x) {
    const float w = wa[a], h0 = AH[(long)b * A + a];
    float acc_ah = 0.f, acc_wa = 0.f;
    const long base = (long)b * K * A + a;
    for (int k0 = 0; k0 < K; k0 += 8) {
      float uu[8], dd[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const long idx = base + (long)min(k0 + q, K - 1) * A;
        uu[q] = __ldg(Uv + idx);
        dd[q] = dUv[idx];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        if (k0 + q < K) {
          const float th = tanhf(h0 + uu[q]);
          const float dp = ds[k0 + q] * w * (1.f - th * th);
          dUv[base + (long)(k0 + q) * A] = dd[q] + dp;
          acc_ah += dp;
          acc_wa += ds[k0 + q] * th;
        }
      }
    }
    dAH[(long)b * A + a] = acc_ah;
    dwa_part[(long)b * A + a] += acc_wa;
  }
  if (threadIdx.x == 0 && sp == 0) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += ds[k];
    dba_part[b] += s;
  }
}

// encoder nn.LSTMCell backward at frame t (mirror of enc_cell_kernel; binary frame mask).
// G (B,4H) activated gates (i,f,g,o) -> dz in place.  dh_in = dHs + dh_carry (dh_carry optional).
// dcc (B,H): in carried dc, out dc*f.
__global__ void enc_cell_bwd_kernel(float* __restrict__ G, const float* __restrict__ c_cur,
                                    const float* __restrict__ c_prev, const float* __restrict__ dHs,
                                    const float* __restrict__ dh_carry, float* __restrict__ dcc,
                                    const float* __restrict__ fmask, int mask_stride, int t, int B, int H) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * H) return;
  const int b = e / H, j = e % H;
  float* g4 = G + (long)b * 4 * H;
  const float i = g4[j], f = g4[H + j], g = g4[2 * H + j], o = g4[3 * H + j];
  const float m = fmask[(long)b * mask_stride + t];
  float dh = dHs[e];
  if (dh_carry) dh += dh_carry[e];
  dh *= m;
  const float tc = tanhf(c_cur[e]);
  const float d_o = dh * tc;
  const float dc = dcc[e] * m + dh * o * (1.f - tc * tc);
  const float cp = c_prev ? c_prev[e] : 0.f;
  const float di = dc * g, dg = dc * i, df = dc * cp;
  dcc[e] = dc * f;
  g4[j] = di * i * (1.f - i);
  g4[H + j] = df * f * (1.f - f);
  g4[2 * H + j] = dg * (1.f - g * g);
  g4[3 * H + j] = d_o * o * (1.f - o);
}

// cross-gate backward, rows (k,b): g = Htgt*(1+R), R = relu(.)*drop
//   dHtgt = dg*(1+R)   (overwrite) ;  dR = R>0 ? dg*Htgt*keep_scale : 0
__global__ void gate_bwd_kernel(const float* __restrict__ dG, long ld_dg, const float* __restrict__ R,
                                const float* __restrict__ Htgt, long n_rows, int H, float keep_scale,
                                float* __restrict__ dHtgt, float* __restrict__ dR) {
  const long n = n_rows * H;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const long r = e / H; const int j = (int)(e % H);
    const float dg = dG[r * ld_dg + j];
    const float rr = R[e];
    dHtgt[e] = dg * (1.f + rr);
    dR[e] = rr > 0.f ? dg * Htgt[e] * keep_scale : 0.f;
  }
}

// decoder POS gate backward, rows (i,b): GP = pos[b]*(1+RG):  dRG = RG>0 ? dGP*pos[b]*keep_scale : 0
__global__ void dgate_bwd_kernel(const float* __restrict__ dGP, const float* __restrict__ RG,
                                 const float* __restrict__ pos, int B, long n_rows, int H, float keep_scale,
                                 float* __restrict__ dRG) {
  const long n = n_rows * H;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const long r = e / H; const int j = (int)(e % H);
    const int b = (int)(r % B);
    dRG[e] = RG[e] > 0.f ? dGP[e] * pos[(long)b * H + j] * keep_scale : 0.f;
  }
}

// relu(+dropout) backward given the saved post-dropout activation: dx = y>0 ? dy*keep_scale : 0 (in place)
__global__ void relu_drop_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, long n, float keep_scale) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
    dy[e] = y[e] > 0.f ? dy[e] * keep_scale : 0.f;
}

// fusion backward, output rows (k,b): dF = dV[(b,k)] * drop * act'(pre) with the activation value
// recovered from the saved post-dropout V.
__global__ void fusion_bwd_kernel(const float* __restrict__ dV, const float* __restrict__ V, int B, int K, int H,
                                  int act, DropSpec drop, float* __restrict__ dF) {
  const long n = (long)B * K * H;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % H);
    const long r = e / H;            // k*B + b
    const int b = (int)(r % B), k = (int)(r / B);
    const long src = ((long)b * K + k) * H + j;
    const float d = drop.factor((uint64_t)e);
    const float v = V[src];
    float out = 0.f;
    if (d != 0.f) {
      const float y = v / d;          // activation value before dropout scaling
      float da;
      if (act == XG_ACT_RELU) da = y > 0.f ? 1.f : 0.f;
      else if (act == XG_ACT_TANH) da = 1.f - y * y;
      else if (act == XG_ACT_SIGMOID) da = y * (1.f - y);
      else da = 1.f;
      out = dV[src] * d * da;
    }
    dF[e] = out;
  }
}

// BN backward, step 1: dy[(b,k), j] = dE[(k,b), j] * fmask * drop * [bn > 0]   (rows back to (b,k))
__global__ void bn_bwd_prep_kernel(const float* __restrict__ dE, const float* __restrict__ Y,
                                   const float* __restrict__ scale, const float* __restrict__ shift,
                                   const float* __restrict__ fmask, int B, int K, int H, DropSpec drop,
                                   float* __restrict__ dy) {
  const long n = (long)B * K * H;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % H);
    const long r = e / H;            // b*K + k
    const int k = (int)(r % K), b = (int)(r / K);
    const float bn = Y[e] * scale[j] + shift[j];
    float v = 0.f;
    if (bn > 0.f) v = dE[((long)k * B + b) * H + j] * fmask[r] * drop.factor((uint64_t)e);
    dy[e] = v;
  }
}

// BN backward, step 2: partial column sums of dy and dy*xhat in double (same grid as colstats_partial)
__global__ void bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ Y,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    int M, int H, double* __restrict__ part) {
  __shared__ double s1[8][33], s2[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int RS = gridDim.y;
  const int rows_per = (M + RS - 1) / RS;
  const int r0 = blockIdx.y * rows_per;
  const int r1 = min(M, r0 + rows_per);
  double a = 0.0, b = 0.0;
  if (j < H) {
    const float mu = mean[j], is = invstd[j];
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const float d = dy[(long)r * H + j];
      const float xh = (Y[(long)r * H + j] - mu) * is;
      a += (double)d;
      b += (double)d * (double)xh;
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

// BN backward, step 3: reduce partials -> dbeta, dgamma (written with beta accumulate) and sums for step 4
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ part, int RS, int H, float beta,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ s_dy, float* __restrict__ s_dyx) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= H) return;
  double a = 0.0, b = 0.0;
  for (int s = 0; s < RS; ++s) { a += part[((long)s * H + j) * 2]; b += part[((long)s * H + j) * 2 + 1]; }
  s_dy[j] = (float)a;
  s_dyx[j] = (float)b;
  dbeta[j] = (beta != 0.f ? beta * dbeta[j] : 0.f) + (float)a;
  dgamma[j] = (beta != 0.f ? beta * dgamma[j] : 0.f) + (float)b;
}

// BN backward, step 4 (in place over dy): train: dY = gamma*invstd/M * (M*dy - s_dy - xhat*s_dyx);
// eval: dY = dy*gamma*invstd
__global__ void bn_bwd_apply_kernel(float* __restrict__ dy, const float* __restrict__ Y,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const float* __restrict__ s_dy,
                                    const float* __restrict__ s_dyx, int M, int H, int train) {
  const long n = (long)M * H;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % H);
    const float is = invstd[j];
    if (train) {
      const float xh = (Y[e] - mean[j]) * is;
      dy[e] = gamma[j] * is / (float)M * ((float)M * dy[e] - s_dy[j] - xh * s_dyx[j]);
    } else {
      dy[e] = dy[e] * gamma[j] * is;
    }
  }
}

// dense embedding gradient: dE[tok(r), :] += dXT[r, :]  (rows (i,b); tok(r) = seq[b*L + i]).
// fp32 atomics: the only non-deterministic summation order on the path (a handful of collisions per row).
__global__ void embed_scatter_add_kernel(const float* __restrict__ dXT, const int64_t* __restrict__ seq,
                                         int B, int L, int rows, int E, int V, float* __restrict__ dEmb) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int b = r % B, i = r / B;
  long t = seq[(long)b * L + i];
  if (t < 0) t = 0;
  if (t >= V) t = V - 1;
  for (int j = threadIdx.x; j < E; j += blockDim.x) atomicAdd(dEmb + t * E + j, dXT[(long)r * E + j]);
}

// LanguageModelCriterion / ClassiferCriterion (SAModel.py:225-253): per-row terms, then a fixed-order
// reduce.  tgt(b,i) = target[b, rotate ? (i+1) % Lp : i];  w = mask[b,i] (* class_mask[b,i])
__device__ __forceinline__ long nll_target(const int64_t* target, long ld, int rotate, int b, int i, int Lp, int N) {
  const int ti = rotate ? ((i + 1 == Lp) ? 0 : i + 1) : i;
  long t = target[(long)b * ld + ti];
  if (t < 0) t = 0;
  if (t >= N) t = N - 1;
  return t;
}
__global__ void nll_terms_kernel(const float* __restrict__ logp, int N, const int64_t* __restrict__ target,
                                 const float* __restrict__ mask, const float* __restrict__ cmask, long ld, int rotate,
                                 int B, int Lp, float* __restrict__ terms /* (B*Lp, 2): -logp*w, w */) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * Lp) return;
  const int b = r / Lp, i = r % Lp;
  const long t = nll_target(target, ld, rotate, b, i, Lp, N);
  const float m = mask[(long)b * ld + i];
  const float c = cmask ? cmask[(long)b * ld + i] : 1.f;
  // reference order of operations: (-1 * logp * mask) [* class_mask]; denominator mask [* class_mask]
  terms[2 * r + 0] = -1.f * logp[(long)r * N + t] * m * c;
  terms[2 * r + 1] = m * c;
}
__global__ void nll_reduce_kernel(const float* __restrict__ terms, int n, float* __restrict__ loss_out,
                                  float* __restrict__ denom_out) {
  __shared__ double sa[256], sb[256];
  double a = 0.0, b = 0.0;
  for (int r = threadIdx.x; r < n; r += blockDim.x) { a += terms[2 * r]; b += terms[2 * r + 1]; }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    loss_out[0] = (float)(sa[0] / sb[0]);
    denom_out[0] = (float)sb[0];
  }
}
// dlogp[b,i,tgt] = -w / denom * grad_out   (dlogp pre-zeroed)
__global__ void nll_grad_kernel(int N, const int64_t* __restrict__ target, const float* __restrict__ mask,
                                const float* __restrict__ cmask, long ld, int rotate, int B, int Lp,
                                const float* __restrict__ denom, const float* __restrict__ grad_out,
                                float* __restrict__ dlogp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= B * Lp) return;
  const int b = r / Lp, i = r % Lp;
  const long t = nll_target(target, ld, rotate, b, i, Lp, N);
  const float w = mask[(long)b * ld + i] * (cmask ? cmask[(long)b * ld + i] : 1.f);
  dlogp[(long)r * N + t] = -w / denom[0] * grad_out[0];
}

__global__ void seq_steps_kernel(const int64_t* __restrict__ seq, int B, int L, int* __restrict__ out) {
  // first i >= 1 whose column sums to zero (SAModel.py:103); L if none.  single CTA.
  __shared__ int first;
  if (threadIdx.x == 0) first = L;
  __syncthreads();
  for (int i = 1 + threadIdx.x; i < L; i += blockDim.x) {
    long s = 0;
    for (int b = 0; b < B; ++b) s += seq[(long)b * L + i];
    if (s == 0) atomicMin(&first, i);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[0] = first;
}

}  // namespace xg
