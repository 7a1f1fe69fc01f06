// fp32 SIMT GEMM with fused epilogues (bias / activation / dropout / cross-gate / row permutation /
// accumulate).  This is the general-shape engine: every dense contraction of the path can run on it
// at full fp32 accuracy (greedy token ids must be bit-exact against the fp32 reference, so bf16/tf32
// single-pass tensor-core math is not an option here; the tcgen05 engine in xg_gemm_tc.cuh uses a
// 3xTF32 split for the large shapes).
//
//   C[row(i), j] = epi( sum_r A(i,r) * B(r,j) )        i < M, j < N, r < K
//   A(i,r) = A[i*sa_i + r*sa_r],  B(r,j) = B[r*sb_r + j*sb_j]   (arbitrary strides, so the same
//   kernel serves  y = x W^T (NT),  dx = dy W (NN)  and  dW = dy^T x (TN)).
#pragma once
#include "xg_common.cuh"

namespace xg {

struct Epilogue {
  float alpha = 1.f;             // v = alpha * acc
  const float* bias0 = nullptr;  // v += bias0[j] + bias1[j] + bias2[j]
  const float* bias1 = nullptr;
  const float* bias2 = nullptr;
  int act = XG_ACT_NONE;         // v = act(v)
  DropSpec drop = {0, 0, 0.f, 1.f, 0};  // v *= mask(i*N + j)
  float* aux = nullptr;          // aux[i*ldaux + j] = v   (value after act+dropout, e.g. the gate R)
  long ldaux = 0;
  const float* tgt = nullptr;    // cross gate: v = tgt[ti*ldt + j] * (1 + v),
  long ldt = 0;                  //   ti = tgt_div ? i / tgt_div : (tgt_mod ? i % tgt_mod : i)
  int tgt_mod = 0;
  int tgt_div = 0;
  float beta = 0.f;              // v += beta * C_old
};

struct GemmP {
  const float* A; long sa_i, sa_r;
  const float* B; long sb_r, sb_j;
  float* C; long ldc;
  int M, N, K;
  int perm_rb = 0, perm_rs = 0;  // row(i) = perm_rb ? (i % perm_rb) * perm_rs + i / perm_rb : i
  Epilogue ep;
  // split-K (skinny products, M = batch): blockIdx.z owns k in [z*kchunk, (z+1)*kchunk); partial tiles go
  // to `ws`, the last CTA to arrive at a tile (counter `ctr`) adds them in split order (deterministic)
  // and runs the epilogue.  ksplit <= 1: plain single-pass GEMM.
  int ksplit = 1, kchunk = 0;
  float* ws = nullptr;
  unsigned int* ctr = nullptr;
};

struct SplitKScratch {           // owned by the handle (xg_context); one stream at a time
  float* ws = nullptr;
  size_t ws_floats = 0;
  unsigned int* ctr = nullptr;   // zero-initialised; every launch leaves it zero again
  int ctr_count = 0;
};

// element (i,j) of the logical output: bias, activation, dropout, aux store, cross gate, row
// permutation, accumulate.  Shared by the SIMT and the tcgen05 engines.
__device__ __forceinline__ float epilogue_bias(const Epilogue& ep, int j) {
  float b = 0.f;
  if (ep.bias0) b += ep.bias0[j];
  if (ep.bias1) b += ep.bias1[j];
  if (ep.bias2) b += ep.bias2[j];
  return b;
}
// v already holds alpha*acc + bias
__device__ __forceinline__ void epilogue_finish(const GemmP& p, int i, int j, float v) {
  const Epilogue& ep = p.ep;
  v = apply_act(v, ep.act);
  if (ep.drop.on()) v *= ep.drop.factor((uint64_t)i * (uint64_t)p.N + (uint64_t)j);
  if (ep.aux) ep.aux[(long)i * ep.ldaux + j] = v;
  if (ep.tgt) {
    const long trow = ep.tgt_div ? (i / ep.tgt_div) : (ep.tgt_mod ? (i % ep.tgt_mod) : i);
    v = ep.tgt[trow * ep.ldt + j] * (1.f + v);
  }
  const long orow = p.perm_rb ? (long)(i % p.perm_rb) * p.perm_rs + i / p.perm_rb : (long)i;
  float* c = p.C + orow * p.ldc + j;
  if (ep.beta != 0.f) v += ep.beta * (*c);
  *c = v;
}
__device__ __forceinline__ void epilogue_store(const GemmP& p, int i, int j, float acc) {
  const Epilogue& ep = p.ep;
  float v = ep.alpha * acc;
  if (ep.bias0) v += ep.bias0[j];
  if (ep.bias1) v += ep.bias1[j];
  if (ep.bias2) v += ep.bias2[j];
  v = apply_act(v, ep.act);
  if (ep.drop.on()) v *= ep.drop.factor((uint64_t)i * (uint64_t)p.N + (uint64_t)j);
  if (ep.aux) ep.aux[(long)i * ep.ldaux + j] = v;
  if (ep.tgt) {
    const long trow = ep.tgt_div ? (i / ep.tgt_div) : (ep.tgt_mod ? (i % ep.tgt_mod) : i);
    v = ep.tgt[trow * ep.ldt + j] * (1.f + v);
  }
  const long orow = p.perm_rb ? (long)(i % p.perm_rb) * p.perm_rs + i / p.perm_rb : (long)i;
  float* c = p.C + orow * p.ldc + j;
  if (ep.beta != 0.f) v += ep.beta * (*c);
  *c = v;
}

template <int BM, int BN, int BK, int TM, int TN, bool A_RC, bool B_RC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_simt_kernel(const GemmP p) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int VM = TM < 4 ? TM : 4;   // contiguous rows per chunk
  constexpr int VN = TN < 4 ? TN : 4;
  constexpr int CM = TM / VM;           // chunks
  constexpr int CN = TN / VN;
  constexpr int PAD = (VM == 4 && VN == 4) ? 4 : 1;
  constexpr int LA = (BM * BK) / NT;
  constexpr int LB = (BN * BK) / NT;
  static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "tile/threads mismatch");

  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;

  float acc[TM][TN];
#pragma unroll
  for (int m = 0; m < TM; ++m)
#pragma unroll
    for (int n = 0; n < TN; ++n) acc[m][n] = 0.f;

  float ra[LA], rb[LB];
  const int kbeg = p.ksplit > 1 ? blockIdx.z * p.kchunk : 0;
  const int kend = p.ksplit > 1 ? min(p.K, kbeg + p.kchunk) : p.K;

  auto load_tile = [&](int k0) {
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      int e = tid + l * NT;
      int rr = A_RC ? (e % BK) : (e / BM);
      int ii = A_RC ? (e / BK) : (e % BM);
      int gi = m0 + ii, gr = k0 + rr;
      ra[l] = (gi < p.M && gr < kend) ? __ldg(p.A + (long)gi * p.sa_i + (long)gr * p.sa_r) : 0.f;
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
      int e = tid + l * NT;
      int rr = B_RC ? (e % BK) : (e / BN);
      int jj = B_RC ? (e / BK) : (e % BN);
      int gj = n0 + jj, gr = k0 + rr;
      rb[l] = (gj < p.N && gr < kend) ? __ldg(p.B + (long)gr * p.sb_r + (long)gj * p.sb_j) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      int e = tid + l * NT;
      int rr = A_RC ? (e % BK) : (e / BM);
      int ii = A_RC ? (e / BK) : (e % BM);
      As[buf][rr][ii] = ra[l];
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
      int e = tid + l * NT;
      int rr = B_RC ? (e % BK) : (e / BN);
      int jj = B_RC ? (e / BK) : (e % BN);
      Bs[buf][rr][jj] = rb[l];
    }
  };

  const int ntiles = (kend - kbeg + BK - 1) / BK;
  load_tile(kbeg);
  store_tile(0);
  __syncthreads();
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) load_tile(kbeg + (t + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int c = 0; c < CM; ++c) {
        const int r0 = c * (BM / CM) + ty * VM;
        if constexpr (VM == 4 && PAD == 4) {
          float4 v = *reinterpret_cast<const float4*>(&As[buf][k][r0]);
          a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int u = 0; u < VM; ++u) a[c * VM + u] = As[buf][k][r0 + u];
        }
      }
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        const int c0 = c * (BN / CN) + tx * VN;
        if constexpr (VN == 4 && PAD == 4) {
          float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][c0]);
          b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
        } else {
#pragma unroll
          for (int u = 0; u < VN; ++u) b[c * VN + u] = Bs[buf][k][c0 + u];
        }
      }
#pragma unroll
      for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TN; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
    }
    if (t + 1 < ntiles) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- split-K: publish the partial tile; the last CTA of the tile reduces in split order ----
  if (p.ksplit > 1) {
    __shared__ int s_last;
    const long tile = (long)blockIdx.y * gridDim.x + blockIdx.x;
    float* mine = p.ws + ((long)blockIdx.z * gridDim.x * gridDim.y + tile) * (BM * BN);
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
      for (int n = 0; n < TN; ++n) __stcg(mine + (m * TN + n) * NT + tid, acc[m][n]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(p.ctr + tile, 1u);
      s_last = (prev == (unsigned)p.ksplit - 1u);
      if (s_last) p.ctr[tile] = 0u;          // ready for the next launch on this stream
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
      for (int n = 0; n < TN; ++n) acc[m][n] = 0.f;
    const long zstride = (long)gridDim.x * gridDim.y * (BM * BN);
    const float* base = p.ws + tile * (BM * BN);
    for (int z = 0; z < p.ksplit; ++z) {
      float v[TM * TN];
#pragma unroll
      for (int e = 0; e < TM * TN; ++e) v[e] = __ldcg(base + (long)z * zstride + e * NT + tid);
#pragma unroll
      for (int e = 0; e < TM * TN; ++e) acc[e / TN][e % TN] += v[e];
    }
  }

  // ---- fused epilogue ----
#pragma unroll
  for (int m = 0; m < TM; ++m) {
    const int i = m0 + (m / VM) * (BM / CM) + ty * VM + (m % VM);
    if (i >= p.M) continue;
#pragma unroll
    for (int n = 0; n < TN; ++n) {
      const int j = n0 + (n / VN) * (BN / CN) + tx * VN + (n % VN);
      if (j >= p.N) continue;
      epilogue_store(p, i, j, acc[m][n]);
    }
  }
}

template <int BM, int BN, int BK, int TM, int TN>
static int gemm_launch_cfg(ErrorSink& es, const GemmP& p, cudaStream_t st) {
  dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM), p.ksplit > 1 ? p.ksplit : 1);
  dim3 block((BM / TM) * (BN / TN));
  const bool arc = (p.sa_r == 1), brc = (p.sb_r == 1);
  if (arc && brc) gemm_simt_kernel<BM, BN, BK, TM, TN, true, true><<<grid, block, 0, st>>>(p);
  else if (arc && !brc) gemm_simt_kernel<BM, BN, BK, TM, TN, true, false><<<grid, block, 0, st>>>(p);
  else if (!arc && brc) gemm_simt_kernel<BM, BN, BK, TM, TN, false, true><<<grid, block, 0, st>>>(p);
  else gemm_simt_kernel<BM, BN, BK, TM, TN, false, false><<<grid, block, 0, st>>>(p);
  XG_LAUNCH_CHECK(es);
  return XG_OK;
}

// Tile choice: keep >= ~120 CTAs in flight on the 148 SMs where the shape allows it; skinny products
// (few output tiles, long K: the M = batch recurrent GEMMs) are cut along K instead.
static int gemm_simt(ErrorSink& es, const GemmP& p_in, cudaStream_t st, const SplitKScratch* sk = nullptr) {
  GemmP p = p_in;
  if (p.M <= 0 || p.N <= 0) return XG_OK;
  XG_REQUIRE(es, p.K >= 0 && p.A && p.B && p.C, XG_ERR_BAD_ARG, "gemm_simt: bad arguments");
  const long big = (long)ceil_div(p.M, 128) * ceil_div(p.N, 128);
  const long med = (long)ceil_div(p.M, 64) * ceil_div(p.N, 64);
  if (sk && sk->ws && med < 120 && p.K >= 128 && (double)p.M * p.N * p.K >= (double)(1 << 21)) {
    int S = (int)(296 / med);
    if (S > 16) S = 16;
    if (S > p.K / 64) S = p.K / 64;
    if (S >= 2) {
      const int kchunk = ceil_div(ceil_div(p.K, S), 32) * 32;
      S = ceil_div(p.K, kchunk);
      if (S >= 2 && med <= sk->ctr_count && (size_t)med * S * 64 * 64 <= sk->ws_floats) {
        p.ksplit = S; p.kchunk = kchunk; p.ws = sk->ws; p.ctr = sk->ctr;
        return gemm_launch_cfg<64, 64, 32, 4, 4>(es, p, st);
      }
    }
  }
  if (big >= 120) return gemm_launch_cfg<128, 128, 8, 8, 8>(es, p, st);
  if (med >= 120) return gemm_launch_cfg<64, 64, 16, 4, 4>(es, p, st);
  return gemm_launch_cfg<32, 32, 32, 2, 2>(es, p, st);
}

// ---- convenience builders (row-major operands) ----
// y (M,N) = x (M,K; ldx) . W (N,K; ldw)^T
static inline GemmP gemm_nt(const float* x, long ldx, const float* W, long ldw, float* y, long ldy, int M, int N, int K) {
  GemmP p{};
  p.A = x; p.sa_i = ldx; p.sa_r = 1;
  p.B = W; p.sb_r = 1; p.sb_j = ldw;
  p.C = y; p.ldc = ldy; p.M = M; p.N = N; p.K = K;
  return p;
}
// dx (M,Kout) = dy (M,N; lddy) . W (N,Kout; ldw)
static inline GemmP gemm_nn(const float* dy, long lddy, const float* W, long ldw, float* dx, long lddx, int M, int Kout, int N) {
  GemmP p{};
  p.A = dy; p.sa_i = lddy; p.sa_r = 1;
  p.B = W; p.sb_r = ldw; p.sb_j = 1;
  p.C = dx; p.ldc = lddx; p.M = M; p.N = Kout; p.K = N;
  return p;
}
// dW (N,Kin) = dy (R,N; lddy)^T . x (R,Kin; ldx)      (reduction over the R rows)
static inline GemmP gemm_tn(const float* dy, long lddy, const float* x, long ldx, float* dW, long lddw, int N, int Kin, int R) {
  GemmP p{};
  p.A = dy; p.sa_i = 1; p.sa_r = lddy;
  p.B = x; p.sb_r = ldx; p.sb_j = 1;
  p.C = dW; p.ldc = lddw; p.M = N; p.N = Kin; p.K = R;
  return p;
}

}  // namespace xg
