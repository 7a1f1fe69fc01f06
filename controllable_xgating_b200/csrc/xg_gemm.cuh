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
    for (int z0 = 0; z0 < p.ksplit; z0 += 2) {     // two splits' loads in flight, added in split order
      float v[2][TM * TN];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int e = 0; e < TM * TN; ++e) v[q][e] = (z0 + q < p.ksplit) ? __ldcg(base + (long)(z0 + q) * zstride + e * NT + tid) : 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int e = 0; e < TM * TN; ++e) acc[e / TN][e % TN] += v[q][e];
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

// ------------------------------------------------------------------------------------
// Skinny split-K kernel (M = batch rows, long K): the whole K-chunk of both operand tiles is fetched in
// ONE shot with 16-byte cp.async (zero-filled past the edges), so a CTA pays one memory round trip instead
// of one per k-tile (the tiled kernel above measured 29-34 us on 64x512x2048, all of it load latency).
//   A: (M,K) K-contiguous.   B_KMAJOR: B is (N,K) K-contiguous [y = x W^T];  else B is (K,N) N-contiguous
//   [dx = dy W].  Thread (tx, ty) of 16x16 owns rows ty + 16a and columns tx + 16b (K-major B, conflict-free
//   float4 reads along k) or 4tx + b (N-major B, one float4 per k).  Split-K publish / fixed-order reduce as above.
// ------------------------------------------------------------------------------------
constexpr int SK_KC = 128;                       // max K-chunk per CTA
constexpr int SK_PA = SK_KC + 4;                 // A / K-major B row pitch (floats): 33 x 16 B -> conflict-free
constexpr int SK_PB = 64 + 4;                    // N-major B row pitch
constexpr int SK_SMEM = (64 * SK_PA + (64 * SK_PA > SK_KC * SK_PB ? 64 * SK_PA : SK_KC * SK_PB)) * 4;

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}

template <bool B_KMAJOR>
__global__ void __launch_bounds__(256) gemm_skinny_kernel(const GemmP p) {
  extern __shared__ __align__(16) float sk_smem[];
  float* As = sk_smem;                 // [64][SK_PA]
  float* Bs = sk_smem + 64 * SK_PA;    // K-major: [64][SK_PA]   N-major: [SK_KC][SK_PB]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int kbeg0 = blockIdx.z * p.kchunk, kend = min(p.K, kbeg0 + p.kchunk);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  // the K range of this CTA in pieces of SK_KC (one piece when the product is split, K / SK_KC pieces otherwise)
  for (int kbeg = kbeg0; kbeg < kend; kbeg += SK_KC) {
    const int kc = min(SK_KC, kend - kbeg);           // multiple of 4 except possibly at the very end of K
    const int kc4 = (kc + 3) >> 2;
    const int kstop = kbeg + kc;
    if (kbeg != kbeg0) __syncthreads();               // the previous piece has been consumed
    // ---- one-shot fetch ----
    for (int e = tid; e < 64 * (SK_KC / 4); e += 256) {          // A: 64 rows x 32 chunks of 4 floats
      const int row = e / (SK_KC / 4), c4 = e % (SK_KC / 4);
      const int gi = m0 + row, gk = kbeg + c4 * 4;
      int nb = (gi < p.M && gk < kstop) ? min(16, (kstop - gk) * 4) : 0;
      const float* src = nb ? p.A + (long)gi * p.sa_i + gk : p.A;
      if (c4 < kc4) cp_async16(As + row * SK_PA + c4 * 4, src, nb);
    }
    if (B_KMAJOR) {
      for (int e = tid; e < 64 * (SK_KC / 4); e += 256) {
        const int row = e / (SK_KC / 4), c4 = e % (SK_KC / 4);
        const int gj = n0 + row, gk = kbeg + c4 * 4;
        int nb = (gj < p.N && gk < kstop) ? min(16, (kstop - gk) * 4) : 0;
        const float* src = nb ? p.B + (long)gj * p.sb_j + gk : p.B;
        if (c4 < kc4) cp_async16(Bs + row * SK_PA + c4 * 4, src, nb);
      }
    } else {
      for (int e = tid; e < SK_KC * 16; e += 256) {               // B: kc rows x 16 chunks of 4 columns
        const int kr = e >> 4, c4 = e & 15;
        const int gk = kbeg + kr, gj = n0 + c4 * 4;
        int nb = (gk < kstop && gj < p.N) ? min(16, (p.N - gj) * 4) : 0;
        const float* src = nb ? p.B + (long)gk * p.sb_r + gj : p.B;
        if (kr < kc4 * 4) cp_async16(Bs + kr * SK_PB + c4 * 4, src, nb);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
#pragma unroll 2
    for (int k4 = 0; k4 < kc4; ++k4) {
      float4 av[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = *reinterpret_cast<const float4*>(As + (ty + 16 * a) * SK_PA + k4 * 4);
      if (B_KMAJOR) {
        float4 bv[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) bv[b] = *reinterpret_cast<const float4*>(Bs + (tx + 16 * b) * SK_PA + k4 * 4);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            acc[a][b] = fmaf(av[a].x, bv[b].x, acc[a][b]); acc[a][b] = fmaf(av[a].y, bv[b].y, acc[a][b]);
            acc[a][b] = fmaf(av[a].z, bv[b].z, acc[a][b]); acc[a][b] = fmaf(av[a].w, bv[b].w, acc[a][b]);
          }
      } else {
        float4 bk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) bk[q] = *reinterpret_cast<const float4*>(Bs + (k4 * 4 + q) * SK_PB + tx * 4);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const float ak[4] = {av[a].x, av[a].y, av[a].z, av[a].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[a][0] = fmaf(ak[q], bk[q].x, acc[a][0]); acc[a][1] = fmaf(ak[q], bk[q].y, acc[a][1]);
            acc[a][2] = fmaf(ak[q], bk[q].z, acc[a][2]); acc[a][3] = fmaf(ak[q], bk[q].w, acc[a][3]);
          }
        }
      }
    }
  }

  // ---- split-K: publish the partial tile, wait until all S splits of the tile are published (the whole grid is
  //      co-resident: <= 2 CTAs per SM), then EVERY split CTA reduces its 1/S slice of the tile in split order
  //      and runs the epilogue on it — a single last-arriver reducing the whole tile cost a 20k-cycle tail ----
  if (p.ksplit > 1) {
    const int S = p.ksplit;
    const long tile = (long)blockIdx.y * gridDim.x + blockIdx.x;
    const long zstride = (long)gridDim.x * gridDim.y * 4096;
    float* mine = p.ws + (long)blockIdx.z * zstride + tile * 4096;
#pragma unroll
    for (int e = 0; e < 16; ++e) __stcg(mine + e * 256 + tid, acc[e >> 2][e & 3]);
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(p.ctr + tile, 1u);
      const long long t0 = clock64();
      while (true) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ctr + tile) : "memory");
        if (v >= (unsigned)S) break;
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
    __syncthreads();
    const float* base = p.ws + tile * 4096;
    for (int q = blockIdx.z * 256 + tid; q < 4096; q += S * 256) {
      float v[32];
#pragma unroll
      for (int z = 0; z < 32; ++z) v[z] = z < S ? __ldcg(base + (long)z * zstride + q) : 0.f;
      float sum = 0.f;
#pragma unroll
      for (int z = 0; z < 32; ++z) sum += v[z];
      const int e = q >> 8, t = q & 255, ea = e >> 2, eb = e & 3, qx = t & 15, qy = t >> 4;
      const int i = m0 + qy + 16 * ea;
      const int j = n0 + (B_KMAJOR ? qx + 16 * eb : qx * 4 + eb);
      if (i < p.M && j < p.N) epilogue_store(p, i, j, sum);
    }
    __syncthreads();
    if (tid == 0) {                                   // the last CTA to finish re-arms the tile's counters
      const unsigned prev = atomicAdd(p.ctr + 128 + tile, 1u);
      if (prev == (unsigned)S - 1u) { p.ctr[128 + tile] = 0u; __threadfence(); p.ctr[tile] = 0u; }
    }
    return;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = m0 + ty + 16 * a;
    if (i >= p.M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = n0 + (B_KMAJOR ? tx + 16 * b : tx * 4 + b);
      if (j < p.N) epilogue_store(p, i, j, acc[a][b]);
    }
  }
}

static inline bool skinny_ok(const GemmP& p) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (p.sa_r != 1 || (p.sa_i & 3) || !al16(p.A) || !al16(p.B)) return false;
  if (p.sb_r == 1) return (p.sb_j & 3) == 0;            // K-major B
  if (p.sb_j == 1) return (p.sb_r & 3) == 0;            // N-major B
  return false;
}

// The split CTAs of a tile wait for one another, so the whole grid must be co-resident: the kernel is launched
// cooperatively (all-or-nothing scheduling — e.g. next to an NCCL kernel that holds a few SMs) and only for
// grids the device can hold (checked against the occupancy calculator, >= 2 CTAs per SM expected).
static int gemm_skinny_launch(ErrorSink& es, const GemmP& p, cudaStream_t st, bool* too_large) {
  static int max_ctas[2] = {-1, -1};
  const bool kmaj = (p.sb_r == 1);
  *too_large = false;
  if (max_ctas[kmaj] < 0) {
    cudaError_t e = kmaj ? cudaFuncSetAttribute(gemm_skinny_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM)
                         : cudaFuncSetAttribute(gemm_skinny_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SK_SMEM);
    if (e != cudaSuccess) { es.set(__FILE__, __LINE__, "cudaFuncSetAttribute(gemm_skinny)", cudaGetErrorString(e)); return XG_ERR_CUDA; }
    int per_sm = 0, dev = 0, sms = 0;
    e = kmaj ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gemm_skinny_kernel<true>, 256, SK_SMEM)
             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gemm_skinny_kernel<false>, 256, SK_SMEM);
    if (e == cudaSuccess) e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { es.set(__FILE__, __LINE__, "occupancy query (gemm_skinny)", cudaGetErrorString(e)); return XG_ERR_CUDA; }
    max_ctas[kmaj] = per_sm * sms;
  }
  dim3 grid(ceil_div(p.N, 64), ceil_div(p.M, 64), p.ksplit > 1 ? p.ksplit : 1);
  if (p.ksplit <= 1) {            // unsplit: CTAs are independent, an ordinary launch (any grid size)
    if (kmaj) gemm_skinny_kernel<true><<<grid, 256, SK_SMEM, st>>>(p);
    else gemm_skinny_kernel<false><<<grid, 256, SK_SMEM, st>>>(p);
    XG_LAUNCH_CHECK(es);
    return XG_OK;
  }
  if ((long)grid.x * grid.y * grid.z > max_ctas[kmaj]) { *too_large = true; return XG_OK; }
  GemmP pc = p;
  void* args[1] = {(void*)&pc};
  cudaError_t e = cudaLaunchCooperativeKernel(kmaj ? (void*)gemm_skinny_kernel<true> : (void*)gemm_skinny_kernel<false>, grid,
                                              dim3(256), args, SK_SMEM, st);
  if (e != cudaSuccess) { es.set(__FILE__, __LINE__, "cudaLaunchCooperativeKernel(gemm_skinny)", cudaGetErrorString(e)); return XG_ERR_CUDA; }
  return XG_OK;
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
      int kchunk = ceil_div(ceil_div(p.K, S), 32) * 32;
      S = ceil_div(p.K, kchunk);
      if (S >= 2 && med <= sk->ctr_count && (size_t)med * S * 64 * 64 <= sk->ws_floats) {
        p.ksplit = S; p.kchunk = kchunk; p.ws = sk->ws; p.ctr = sk->ctr;
        if (skinny_ok(p) && med <= 128) {
          if (kchunk > SK_KC) {                       // one-shot fetch holds at most SK_KC of K per CTA
            kchunk = SK_KC; S = ceil_div(p.K, kchunk);
            if ((size_t)med * S * 64 * 64 > sk->ws_floats || S > 32 || med * S > 296) return gemm_launch_cfg<64, 64, 32, 4, 4>(es, p, st);
            p.ksplit = S; p.kchunk = kchunk;
          }
          bool too_large = false;
          const int rc = gemm_skinny_launch(es, p, st, &too_large);
          if (rc != XG_OK || !too_large) return rc;
        }
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
