// tcgen05 GEMM engine with fp32-grade accuracy (3xTF32 split), sm_100a.
//
//   D[p, q] = sum_k P[p,k] * Q[q,k]       p < Pn (128 TMEM lanes per tile), q < Qn (BN TMEM columns)
//
// Both operands are K-major fp32 matrices that were split beforehand into a tf32-exact high part and
// a tf32-rounded low part (x = hi + lo + O(2^-23 |x|)); the kernel accumulates  lo*hi + hi*lo + hi*hi
// in the fp32 TMEM accumulator (the dropped lo*lo term is O(2^-22)), which is what makes greedy token
// ids survive tensor-core math.  A plain tf32 or bf16 pass would flip ids.
//
// Pipeline (one 128 x BN output tile per CTA, 128 threads):
//   warp 0 / lane 0 : TMA producer  — cp.async.bulk.tensor 2D, SWIZZLE_128B boxes of 32 floats x rows,
//                     4 boxes (P_hi, P_lo, Q_hi, Q_lo) per stage, mbarrier complete_tx
//   warp 1 / lane 0 : MMA issuer    — tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BN, K=8, smem
//                     descriptors (K-major, 128B swizzle, SBO = 1024 B), tcgen05.commit frees the stage
//   warp 2          : TMEM allocator (BN fp32 columns)
//   all 4 warps     : epilogue — tcgen05.ld 32x32b (lane = p), fused epilogue_store(), coalesced stores
// No thread ever writes the operand tiles: they go global -> smem by TMA and smem -> tensor core by
// UMMA, both in the async proxy, so no proxy fences are needed on the operand path.
#pragma once
#include <cuda_fp16.h>
#include <cuda.h>

#include <map>
#include <tuple>
#include <unordered_map>

#include "xg_context.cuh"

namespace xg {

// ------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded spin: a mis-programmed pipeline traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();   // ~2 s at 1.9 GHz
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, %1;\n\t@P1 mov.s32 %0, 1;\n\t}" : "+r"(pred) : "r"(0xffffffffu));
  return pred != 0;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (8-row x 128-byte swizzle atoms, 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address            bits [0,14)
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset = 1024 B  [32,46)
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell) [46,48)
  d |= (uint64_t)2 << 61;                         // layout type SWIZZLE_128B      [61,64)
  return d;
}
// instruction descriptor: D fp32, A/B fp16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// fp16 operand pairs (3xFP16): hi = fp16(v), lo = fp16((v - hi) * 2^11); hi.hi + (lo.hi + hi.lo) / 2^11 carries the same
// 11 + 11 mantissa bits as a tf32 pair, the scale keeps lo out of the fp16 subnormals.  |v| must stay below 65504.
constexpr float TC16_SCALE = 2048.f;
// instruction descriptor: D fp32, A/B tf32, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------
// split pre-pass: hi = rna_tf32(x), lo = rna_tf32(x - hi), written K-major with the K extent padded
// to a multiple of 32 (zeros) so TMA boxes never leave the row.  src(r,k) = X[r*sr + k*sk] lets the
// same kernel transpose (the NN / TN layouts of the backward pass become K-major here).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__global__ void split_tf32_kernel(const float* __restrict__ X, long sr, long sk, int rows, int K, int Kp,
                                  float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const bool k_contig = (sk == 1);
  // load: threadIdx.x runs along the source-contiguous index
  for (int y = threadIdx.y; y < 32; y += 8) {
    int r = k_contig ? r0 + y : r0 + threadIdx.x;
    int k = k_contig ? k0 + threadIdx.x : k0 + y;
    float v = (r < rows && k < K) ? X[(long)r * sr + (long)k * sk] : 0.f;
    if (k_contig) tile[y][threadIdx.x] = v; else tile[threadIdx.x][y] = v;   // tile[r - r0][k - k0]
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int r = r0 + y, k = k0 + threadIdx.x;
    if (r < rows && k < Kp) {
      const float v = tile[y][threadIdx.x];
      const float h = tf32_rna(v);
      hi[(long)r * Kp + k] = h;
      lo[(long)r * Kp + k] = tf32_rna(v - h);
    }
  }
}

// Transposing split (the reduction index is the slow one of the source: X[r + k * sk], the operands of the weight-gradient
// products) on 64 x 64 tiles with 16-byte loads and stores: the 32 x 32 kernel above moves 4 scalars per thread and ran
// at 1.2 TB/s (8.8 us for a 512 x 1792 operand whose 11 MB of traffic are worth 2 us).
__global__ void __launch_bounds__(256)
split_tf32_t64_kernel(const float* __restrict__ X, long sk, int rows, int K, int Kp, float* __restrict__ hi, float* __restrict__ lo) {
  __shared__ float tile[64][65];                   // tile[k - k0][r - r0]
  const int r0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  float4 v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + i * 256, kl = idx >> 4, r = r0 + (idx & 15) * 4, k = k0 + kl;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) {
      const float* src = X + (long)k * sk + r;
      if (r + 3 < rows) v[i] = __ldg(reinterpret_cast<const float4*>(src));
      else {
        if (r < rows) v[i].x = __ldg(src);
        if (r + 1 < rows) v[i].y = __ldg(src + 1);
        if (r + 2 < rows) v[i].z = __ldg(src + 2);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + i * 256, kl = idx >> 4, rl = (idx & 15) * 4;
    tile[kl][rl] = v[i].x; tile[kl][rl + 1] = v[i].y; tile[kl][rl + 2] = v[i].z; tile[kl][rl + 3] = v[i].w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = threadIdx.x + i * 256, rl = idx >> 4, kl = (idx & 15) * 4, r = r0 + rl, k = k0 + kl;
    if (r < rows && k < Kp) {                        // (Kp is a multiple of 32, k of 4: the four columns are in range together)
      const float x[4] = {tile[kl][rl], tile[kl + 1][rl], tile[kl + 2][rl], tile[kl + 3][rl]};
      float4 h, l;
      h.x = tf32_rna(x[0]); h.y = tf32_rna(x[1]); h.z = tf32_rna(x[2]); h.w = tf32_rna(x[3]);
      l.x = tf32_rna(x[0] - h.x); l.y = tf32_rna(x[1] - h.y); l.z = tf32_rna(x[2] - h.z); l.w = tf32_rna(x[3] - h.w);
      *reinterpret_cast<float4*>(hi + (long)r * Kp + k) = h;
      *reinterpret_cast<float4*>(lo + (long)r * Kp + k) = l;
    }
  }
}

// K-contiguous sources with K == Kp (no padding column, 16-byte aligned rows of pitch sr): the split is elementwise,
// so it runs as a float4 stream with 4 independent 16-byte loads per thread in flight (the tiled kernel above moves
// 4 scalars per thread through shared memory: 10.6 -> 8.0 us per launch on the encoder activations).
__global__ void __launch_bounds__(256)
split_tf32_flat_kernel(const float4* __restrict__ X, long sr4, int k4, long n4, float4* __restrict__ hi, float4* __restrict__ lo) {
  const long stride = (long)gridDim.x * 256;
  const bool dense = sr4 == (long)k4;
  for (long i0 = (long)blockIdx.x * 256 + threadIdx.x; i0 < n4; i0 += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const long i = i0 + q * stride; if (i < n4) v[q] = __ldg(X + (dense ? i : (i / k4) * sr4 + (i % k4))); }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long i = i0 + q * stride;
      if (i < n4) {
        float4 h, l;
        h.x = tf32_rna(v[q].x); h.y = tf32_rna(v[q].y); h.z = tf32_rna(v[q].z); h.w = tf32_rna(v[q].w);
        l.x = tf32_rna(v[q].x - h.x); l.y = tf32_rna(v[q].y - h.y); l.z = tf32_rna(v[q].z - h.z); l.w = tf32_rna(v[q].w - h.w);
        hi[i] = h; lo[i] = l;
      }
    }
  }
}

// the same two passes for fp16 hi / lo pairs (K padded to a multiple of 64: 128-byte rows)
__global__ void split_f16_kernel(const float* __restrict__ X, long sr, long sk, int rows, int K, int Kp,
                                 __half* __restrict__ hi, __half* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const bool k_contig = (sk == 1);
  for (int y = threadIdx.y; y < 32; y += 8) {
    int r = k_contig ? r0 + y : r0 + threadIdx.x;
    int k = k_contig ? k0 + threadIdx.x : k0 + y;
    float v = (r < rows && k < K) ? X[(long)r * sr + (long)k * sk] : 0.f;
    if (k_contig) tile[y][threadIdx.x] = v; else tile[threadIdx.x][y] = v;   // tile[r - r0][k - k0]
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int r = r0 + y, k = k0 + threadIdx.x;
    if (r < rows && k < Kp) {
      const float v = tile[y][threadIdx.x];
      const __half h = __float2half_rn(v);
      hi[(long)r * Kp + k] = h;
      lo[(long)r * Kp + k] = __float2half_rn((v - __half2float(h)) * TC16_SCALE);
    }
  }
}
__global__ void __launch_bounds__(256)
split_f16_flat_kernel(const float4* __restrict__ X, long sr4, int k4, long n4, uint2* __restrict__ hi, uint2* __restrict__ lo) {
  const long stride = (long)gridDim.x * 256;
  const bool dense = sr4 == (long)k4;
  for (long i0 = (long)blockIdx.x * 256 + threadIdx.x; i0 < n4; i0 += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const long i = i0 + q * stride; if (i < n4) v[q] = __ldg(X + (dense ? i : (i / k4) * sr4 + (i % k4))); }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long i = i0 + q * stride;
      if (i < n4) {
        const float f[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
        __half h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { h[e] = __float2half_rn(f[e]); l[e] = __float2half_rn((f[e] - __half2float(h[e])) * TC16_SCALE); }
        uint2 ph, pl;
        ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
        ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
        pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
        pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
        hi[i] = ph; lo[i] = pl;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// the GEMM kernel
//
// Accuracy note (measured on B200): the tensor core adds into the fp32 TMEM accumulator with
// truncation, so a chain of n dependent MMAs carries a one-sided error of ~n/2 ulp of the running sum
// (3.6e-6 relative at K=512 with one accumulator).  Therefore
//   * the large hi*hi products are accumulated in SHORT chains: KCH k-blocks (4*KCH MMAs) into one of
//     two ping-pong TMEM accumulators, which the epilogue warps drain into fp32 REGISTER accumulators
//     with round-to-nearest adds while the tensor core fills the other one;
//   * the small cross terms lo*hi + hi*lo (2^-11 of the result) get their own TMEM accumulator for the
//     whole K extent: truncation there is 2^-11 smaller and invisible.
// Net error ~1e-7, the same class as an fp32 FFMA loop.
// ------------------------------------------------------------------------------------
struct TcParams {
  int Pn, Qn, K;       // K = padded reduction extent (multiple of 32)
  int swap_out;        // 1: logical (i,j) = (q,p)  [y = x W^T with W on the lane side]; 0: (i,j) = (p,q)
  int dbg;             // diagnostics (XG_TC_DEBUG): 1 skip MMA issue, 2 one long chain (no promotion), 4 skip lo loads
  GemmP g;             // output pointer / leading dimension / epilogue (M,N = logical output extents)
};

template <int BN>
struct TcCfg {
  static constexpr int kStageBytes = 2 * 128 * 128 + 2 * BN * 128;
  static constexpr int kStages = (BN == 64) ? 4 : 3;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kTmemCols = (BN == 64) ? 256 : 512;   // 2 ping-pong accumulator pairs (main | cross), power of 2
  static constexpr int kChunk = 2;                            // k-blocks (of 32) per accumulation chain
  static constexpr int kThreads = 192;                        // producer warp, MMA warp, 4 epilogue warps
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, bool F16 = false>      // F16: fp16 operand pairs, k-blocks of 64 (128-byte rows either way)
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmPh, const __grid_constant__ CUtensorMap tmPl,
               const __grid_constant__ CUtensorMap tmQh, const __grid_constant__ CUtensorMap tmQl, const TcParams prm) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the address space
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* acc_full = empty_bar + Cfg::kStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint64_t* small_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(small_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p0 = blockIdx.x * 128, q0 = blockIdx.y * BN;
  constexpr int KB = F16 ? 64 : 32;          // K elements of a k-block
  const int num_kb = prm.K / KB;
  const int kchunk = (prm.dbg & 2) ? (1 << 20) : Cfg::kChunk;
  const int num_chunks = (num_kb + kchunk - 1) / kchunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    mbar_init(small_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmPh); tma_prefetch_desc(&tmPl); tma_prefetch_desc(&tmQh); tma_prefetch_desc(&tmQl);
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // warp-uniform role dispatch (the whole warp walks the loops, one elected lane issues): inside a divergent
  // `lane == 0` region every UTCHMMA / UTMALDG compiles to an ELECT + R2UR.BROADCAST waterfall (~90 cycles each)
  const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
  if (uwarp == 0) {
    // ===== TMA producer =====
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % Cfg::kStages;
      const uint32_t ph = (kb / Cfg::kStages) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* st = smem + s * Cfg::kStageBytes;
      if (elect_one_sync()) {
        if (prm.dbg & 4) {
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes / 2);
          tma_load_2d(st, &tmPh, &full_bar[s], kb * KB, p0);
          tma_load_2d(st + 2 * 128 * 128, &tmQh, &full_bar[s], kb * KB, q0);
        } else {
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          tma_load_2d(st, &tmPh, &full_bar[s], kb * KB, p0);
          tma_load_2d(st + 128 * 128, &tmPl, &full_bar[s], kb * KB, p0);
          tma_load_2d(st + 2 * 128 * 128, &tmQh, &full_bar[s], kb * KB, q0);
          tma_load_2d(st + 2 * 128 * 128 + BN * 128, &tmQl, &full_bar[s], kb * KB, q0);
        }
      }
      __syncwarp();
    }
  } else if (uwarp == 1) {
    // ===== MMA issuer =====
    // Per k-step (K = 8) TWO instructions: P_hi . [Q_hi ; Q_lo] as one N = 2 BN product (the hi and lo tiles of Q lie back
    // to back in the stage: main term -> columns [0, BN), cross term hi.lo -> [BN, 2 BN) of the accumulator pair) and
    // P_lo . Q_hi (N = BN) onto the cross columns.  Three N = BN instructions per k-step cost ~130 cycles each here (ncu,
    // profiles/r1n_gemm_tc128_*): the issue rate of the MMA warp and the re-read of the P_hi tile paced the main loop.
    constexpr uint32_t idesc = F16 ? umma_idesc_f16(128, BN) : umma_idesc_tf32(128, BN);
    constexpr uint32_t idesc2 = F16 ? umma_idesc_f16(128, 2 * BN) : umma_idesc_tf32(128, 2 * BN);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem0 = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
    const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
    for (int c = 0; c < num_chunks; ++c) {
      const int b = c & 1;
      mbar_wait(&acc_empty[b], ((c >> 1) & 1) ^ 1);       // epilogue has drained this accumulator pair
      tc_fence_after();
      const uint32_t tmem_pair = tb + b * 2 * BN;
      for (int kk = 0; kk < kchunk; ++kk) {
        const int kb = c * kchunk + kk;
        if (kb >= num_kb) break;
        const int s = kb % Cfg::kStages;
        mbar_wait(&full_bar[s], (kb / Cfg::kStages) & 1);
        tc_fence_after();
        const uint32_t dlo = (((smem0 + s * Cfg::kStageBytes) & 0x3FFFFu) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {         // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte swizzle row
            const uint64_t ph_d = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2 + (1u << 16));
            const uint64_t pl_d = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2 + ((128 * 128) >> 4) + (1u << 16));
            const uint64_t qh_d = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2 + ((2 * 128 * 128) >> 4) + (1u << 16));      // 2 BN rows: Q_hi, Q_lo
            if (prm.dbg & 1) continue;
            if (F16) {
              umma_f16(tmem_pair, ph_d, qh_d, idesc2, (kk | k4) != 0);
              umma_f16(tmem_pair + BN, pl_d, qh_d, idesc, 1);
            } else {
              umma_tf32(tmem_pair, ph_d, qh_d, idesc2, (kk | k4) != 0);
              umma_tf32(tmem_pair + BN, pl_d, qh_d, idesc, 1);
            }
          }
          umma_commit(&empty_bar[s]);               // stage reusable once these MMAs have read it
          if (kk == kchunk - 1 || kb == num_kb - 1) umma_commit(&acc_full[b]);   // this chain is complete
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps: drain short chains (main + cross columns) into fp32 registers, then fused epilogue =====
    const int quad = warp & 3;                      // TMEM lanes [32*quad, 32*quad+32) belong to this warp
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int p = p0 + quad * 32 + lane;            // TMEM lane == row p of the tile
    float acc[BN];
#pragma unroll
    for (int u = 0; u < BN; ++u) acc[u] = 0.f;
    for (int c = 0; c < num_chunks; ++c) {
      const int b = c & 1;
      mbar_wait(&acc_full[b], (c >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < BN; cc += 32) {
        uint32_t r[32], r2[32];
        tmem_ld32(tmem_base + lane_base + (uint32_t)(b * 2 * BN + cc), r);
        tmem_ld32(tmem_base + lane_base + (uint32_t)(b * 2 * BN + BN + cc), r2);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u)
          acc[cc + u] += F16 ? fmaf(__uint_as_float(r2[u]), 1.f / TC16_SCALE, __uint_as_float(r[u])) : __uint_as_float(r[u]) + __uint_as_float(r2[u]);
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[b]);
    }
    const Epilogue& ep0 = prm.g.ep;
    const bool direct = prm.swap_out && !ep0.drop.on() && ep0.aux == nullptr && ep0.tgt == nullptr && ep0.beta == 0.f &&
                        (ep0.act == XG_ACT_NONE || ep0.act == XG_ACT_RELU);
    if (direct) {
      // Plain outputs (bias, optional ReLU, row permutation) — most products of the path — go straight from the
      // register accumulators to global memory, fully unrolled (~8 instructions per column).  The rolled loop below
      // re-read the staged tile through generic loads into ONE live register: ~195 cycles per column, and that tail
      // was 47 % of a CTA's lifetime on 128 x 128 tiles (ncu source page, profiles/r1n_gemm_tc128_*).
      if (p < prm.Pn) {
        const float alpha = ep0.alpha, bias_p = epilogue_bias(ep0, p);
        const bool relu = ep0.act == XG_ACT_RELU;
        const int rb = prm.g.perm_rb, rs = prm.g.perm_rs;
        const int qn = min(BN, prm.Qn - q0);
        int qm = rb ? q0 % rb : 0, qd = rb ? q0 / rb : 0;
        float* cbase = prm.g.C + p;
        const long ldc = prm.g.ldc;
#pragma unroll
        for (int u = 0; u < BN; ++u) {
          if (u < qn) {
            float v = fmaf(alpha, acc[u], bias_p);
            if (relu) v = fmaxf(v, 0.f);
            const long orow = rb ? (long)qm * rs + qd : (long)(q0 + u);
            cbase[orow * ldc] = v;
            if (rb) { if (++qm == rb) { qm = 0; ++qd; } }
          }
        }
      }
    } else {
      // Stage the register accumulators through the pipeline buffers (idle now: every MMA has retired)
      // so that ONE rolled copy of the generic epilogue serves all BN columns.  Unrolling it BN times
      // produced an 84k-instruction kernel that was instruction-fetch bound (ncu, profiles/r1b).
      float* stage_out = reinterpret_cast<float*>(smem);
      const int pl = quad * 32 + lane;
#pragma unroll
      for (int u = 0; u < BN; ++u) stage_out[u * 128 + pl] = acc[u];   // same thread reads it back: no barrier
      if (p < prm.Pn) {
        const Epilogue& ep = prm.g.ep;
        const float alpha = ep.alpha;
        const float bias_p = prm.swap_out ? epilogue_bias(ep, p) : 0.f;   // j == p: per-lane constant
        const int qn = min(BN, prm.Qn - q0);
        if (prm.swap_out) {
          // Everything else (dropout, auxiliary store, cross gate tgt.(1+v), accumulate, tanh / sigmoid): rounds of 8
          // columns, every load of a round (staged accumulator, gate target, old output) issued before its first store.
          // Written element by element the compiler may not move a load above the previous store (the pointers may
          // alias), so each column paid a full global-load latency.
          const GemmP& g = prm.g;
          const int rb = g.perm_rb, rs = g.perm_rs;
          const float beta = ep.beta;
          const bool has_drop = ep.drop.on();
#pragma unroll 1
          for (int u0 = 0; u0 < qn; u0 += 8) {
            float a[8], tg[8], cold[8];
            long coff[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int uu = min(u0 + k, qn - 1), q = q0 + uu;
              a[k] = stage_out[uu * 128 + pl];
              const long orow = rb ? (long)(q % rb) * rs + q / rb : (long)q;
              coff[k] = orow * g.ldc + p;
              tg[k] = 0.f; cold[k] = 0.f;
              if (ep.tgt) {
                const long trow = ep.tgt_div ? (q / ep.tgt_div) : (ep.tgt_mod ? (q % ep.tgt_mod) : q);
                tg[k] = ep.tgt[trow * ep.ldt + p];
              }
              if (beta != 0.f) cold[k] = g.C[coff[k]];
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (u0 + k < qn) {
                const int q = q0 + u0 + k;
                float v = apply_act(alpha * a[k] + bias_p, ep.act);
                if (has_drop) v *= ep.drop.factor((uint64_t)q * (uint64_t)g.N + (uint64_t)p);
                if (ep.aux) ep.aux[(long)q * ep.ldaux + p] = v;
                if (ep.tgt) v = tg[k] * (1.f + v);
                if (beta != 0.f) v += beta * cold[k];
                g.C[coff[k]] = v;
              }
            }
          }
        } else {
#pragma unroll 1
          for (int u = 0; u < qn; ++u) {
            const float a = stage_out[u * 128 + pl];
            epilogue_finish(prm.g, p, q0 + u, alpha * a + epilogue_bias(prm.g.ep, q0 + u));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  PFN_encodeTiled encode = nullptr;
  bool attr_set[4] = {false, false, false, false};
  // split copies of bound parameters: key = (param pointer, transposed?) -> {hi, lo, rows, Kp}
  struct Split { float* hi; float* lo; int rows; int K; int Kp; bool valid; };
  typedef std::tuple<const float*, long, long, int, int> SplitKey;   // (pointer, row stride, k stride, rows, K)
  std::map<SplitKey, Split> weight_cache;
  // scratch for on-the-fly operand splits (grown on demand)
  float* scratch[2] = {nullptr, nullptr};
  size_t scratch_floats[2] = {0, 0};
  // "hold" regions (tc_hold_begin / _end): the caller guarantees that no operand of the region is written after its
  // first use inside it, so an activation that feeds several GEMMs in the same layout (dz of a cell in its three weight
  // gradients, the embeddings in two products) is split once.  The buffers are kept from step to step.
  // fp16 operand pairs of the forward products (gemm_tc with ctx->tc_f16): same caching rules
  struct Split16 { __half* hi; __half* lo; int rows; int K; int Kp; bool valid; };
  std::map<SplitKey, Split16> weight_cache16;
  __half* scratch16[2] = {nullptr, nullptr};
  size_t scratch16_halfs[2] = {0, 0};
  struct Held { SplitKey key; float* buf; size_t floats; bool valid; };
  std::vector<Held> held;
  bool hold = false;
};

inline TcState*& tc_state(xg_context* ctx) {
  static std::unordered_map<xg_context*, TcState*> m;
  return m[ctx];
}

static int tc_init(xg_context* ctx, TcState*& ts) {
  ts = tc_state(ctx);
  if (ts) return XG_OK;
  ts = new TcState();
  tc_state(ctx) = ts;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  XG_CUDA_TRY(ctx->es, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  XG_REQUIRE(ctx->es, fn != nullptr && qres == cudaDriverEntryPointSuccess, XG_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  ts->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  return XG_OK;
}

static void tc_release(xg_context* ctx) {
  TcState* ts = tc_state(ctx);
  if (!ts) return;
  for (auto& kv : ts->weight_cache) { cudaFree(kv.second.hi); cudaFree(kv.second.lo); }
  for (auto& kv : ts->weight_cache16) { cudaFree(kv.second.hi); cudaFree(kv.second.lo); }
  for (int i = 0; i < 2; ++i) if (ts->scratch16[i]) cudaFree(ts->scratch16[i]);
  for (int i = 0; i < 2; ++i) if (ts->scratch[i]) cudaFree(ts->scratch[i]);
  for (auto& h : ts->held) if (h.buf) cudaFree(h.buf);
  delete ts;
  tc_state(ctx) = nullptr;
}

// the bound parameters changed (optimizer step, load_state_dict): the split copies are stale.  The buffers stay —
// freeing them would synchronise the device once per training step — and are refilled on next use.
static void tc_invalidate_weights(xg_context* ctx) {
  TcState* ts = tc_state(ctx);
  if (!ts) return;
  for (auto& kv : ts->weight_cache) kv.second.valid = false;
  for (auto& kv : ts->weight_cache16) kv.second.valid = false;
}

// 2-D map over a K-major fp32 matrix [rows][Kp] (pitch Kp floats): box = 32 floats x box_rows, 128B swizzle
static int tc_make_map(xg_context* ctx, TcState* ts, const float* base, int rows, int Kp, int box_rows, CUtensorMap* out) {
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * sizeof(float)};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = ts->encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->es.set(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed", nullptr);
    return XG_ERR_CUDA;
  }
  return XG_OK;
}

static int tc_split(xg_context* ctx, const float* X, long sr, long sk, int rows, int K, int Kp, float* hi, float* lo,
                    cudaStream_t st) {
  static const bool shapes = getenv("XG_PROF_SPLIT_SHAPES") != nullptr;      // per-shape rows in the profile report (diagnostics)
  char tag[64];
  if (shapes && ctx->prof_on) snprintf(tag, sizeof(tag), "split_tf32_%dx%d_%s", rows, K, sk == 1 ? "rowmajor" : "transposing");
  else snprintf(tag, sizeof(tag), "split_tf32");
  ProfScope ps(ctx, std::string(tag), st);
  if (sk == 1 && sr % 4 == 0 && sr >= (long)K && K == Kp && ((uintptr_t)X & 15) == 0 && ((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0) {
    const long n4 = (long)rows * Kp / 4;
    const long want = (n4 + 1023) / 1024;
    const int blocks = (int)std::max<long>(1, std::min<long>(want, (long)ctx->sm_count * 8));
    split_tf32_flat_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(X), sr / 4, Kp / 4, n4, reinterpret_cast<float4*>(hi),
                                                   reinterpret_cast<float4*>(lo));
    XG_LAUNCH_CHECK(ctx->es);
    return XG_OK;
  }
  if (sr == 1 && sk % 4 == 0 && ((uintptr_t)X & 15) == 0 && ((uintptr_t)hi & 15) == 0 && ((uintptr_t)lo & 15) == 0) {
    dim3 grid(ceil_div(Kp, 64), ceil_div(rows, 64));
    split_tf32_t64_kernel<<<grid, 256, 0, st>>>(X, sk, rows, K, Kp, hi, lo);
    XG_LAUNCH_CHECK(ctx->es);
    return XG_OK;
  }
  dim3 grid(Kp / 32, ceil_div(rows, 32));
  split_tf32_kernel<<<grid, dim3(32, 8), 0, st>>>(X, sr, sk, rows, K, Kp, hi, lo);
  XG_LAUNCH_CHECK(ctx->es);
  return XG_OK;
}

// operand = logical matrix O[r][k] = X[r*sr + k*sk], r < rows, k < K.  Returns split K-major copies.
// Bound parameters are split once and cached until xg_params_changed(); everything else goes to scratch.
static int tc_operand(xg_context* ctx, TcState* ts, int slot, const float* X, long sr, long sk, int rows, int K,
                      const float** hi, const float** lo, int* Kp_out, cudaStream_t st) {
  const int Kp = (K + 31) / 32 * 32;
  *Kp_out = Kp;
  bool is_param = false;
  for (int i = 0; i < XG_NUM_PARAMS && !is_param; ++i) {
    int pr, pc;
    param_shape(ctx->d, i, &pr, &pc);
    const float* b = ctx->P[i];
    if (b && X >= b && X < b + (long)pr * pc) is_param = true;
  }
  if (is_param) {
    // key: pointer, orientation and extents (sub-views such as W_h2a[:, H:] get their own entry)
    const TcState::SplitKey key(X, sr, sk, rows, K);
    auto it = ts->weight_cache.find(key);
    if (it == ts->weight_cache.end()) {
      TcState::Split sp{nullptr, nullptr, rows, K, Kp, false};
      XG_CUDA_TRY(ctx->es, cudaMalloc(&sp.hi, sizeof(float) * (size_t)rows * Kp));
      XG_CUDA_TRY(ctx->es, cudaMalloc(&sp.lo, sizeof(float) * (size_t)rows * Kp));
      it = ts->weight_cache.emplace(key, sp).first;
    }
    if (!it->second.valid) {
      XG_TRY(tc_split(ctx, X, sr, sk, rows, K, Kp, it->second.hi, it->second.lo, st));
      it->second.valid = true;
    }
    *hi = it->second.hi;
    *lo = it->second.lo;
    return XG_OK;
  }
  const size_t need = (size_t)rows * Kp * 2;
  if (ts->hold) {
    const TcState::SplitKey key(X, sr, sk, rows, K);
    TcState::Held* free_slot = nullptr;
    for (auto& h : ts->held) {
      if (h.valid && h.key == key) { *hi = h.buf; *lo = h.buf + (size_t)rows * Kp; return XG_OK; }
      if (!h.valid && (!free_slot || (h.floats >= need && (free_slot->floats < need || h.floats < free_slot->floats)))) free_slot = &h;
    }
    if (!free_slot) { ts->held.push_back(TcState::Held{key, nullptr, 0, false}); free_slot = &ts->held.back(); }
    if (free_slot->floats < need) {
      if (free_slot->buf) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(free_slot->buf); free_slot->buf = nullptr; free_slot->floats = 0; }
      XG_CUDA_TRY(ctx->es, cudaMalloc(&free_slot->buf, sizeof(float) * need));
      free_slot->floats = need;
    }
    XG_TRY(tc_split(ctx, X, sr, sk, rows, K, Kp, free_slot->buf, free_slot->buf + (size_t)rows * Kp, st));
    free_slot->key = key; free_slot->valid = true;
    *hi = free_slot->buf; *lo = free_slot->buf + (size_t)rows * Kp;
    return XG_OK;
  }
  if (ts->scratch_floats[slot] < need) {
    XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
    if (ts->scratch[slot]) cudaFree(ts->scratch[slot]);
    ts->scratch[slot] = nullptr;
    ts->scratch_floats[slot] = 0;
    XG_CUDA_TRY(ctx->es, cudaMalloc(&ts->scratch[slot], sizeof(float) * need));
    ts->scratch_floats[slot] = need;
  }
  float* h = ts->scratch[slot];
  float* l = h + (size_t)rows * Kp;
  XG_TRY(tc_split(ctx, X, sr, sk, rows, K, Kp, h, l, st));
  *hi = h;
  *lo = l;
  return XG_OK;
}

// ---- fp16-pair operands (forward products) ----
static int tc_make_map16(xg_context* ctx, TcState* ts, const __half* base, int rows, int Kp, int box_rows, CUtensorMap* out) {
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * sizeof(__half)};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = ts->encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->es.set(__FILE__, __LINE__, "cuTensorMapEncodeTiled (fp16) failed", nullptr);
    return XG_ERR_CUDA;
  }
  return XG_OK;
}
static int tc_split16(xg_context* ctx, const float* X, long sr, long sk, int rows, int K, int Kp, __half* hi, __half* lo, cudaStream_t st) {
  ProfScope ps(ctx, "split_f16", st);
  if (sk == 1 && sr % 4 == 0 && sr >= (long)K && K == Kp && ((uintptr_t)X & 15) == 0 && ((uintptr_t)hi & 7) == 0 && ((uintptr_t)lo & 7) == 0) {
    const long n4 = (long)rows * Kp / 4;
    const long want = (n4 + 1023) / 1024;
    const int blocks = (int)std::max<long>(1, std::min<long>(want, (long)ctx->sm_count * 8));
    split_f16_flat_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(X), sr / 4, Kp / 4, n4, reinterpret_cast<uint2*>(hi),
                                                  reinterpret_cast<uint2*>(lo));
    XG_LAUNCH_CHECK(ctx->es);
    return XG_OK;
  }
  dim3 grid(Kp / 32, ceil_div(rows, 32));
  split_f16_kernel<<<grid, dim3(32, 8), 0, st>>>(X, sr, sk, rows, K, Kp, hi, lo);
  XG_LAUNCH_CHECK(ctx->es);
  return XG_OK;
}
static int tc_operand16(xg_context* ctx, TcState* ts, int slot, const float* X, long sr, long sk, int rows, int K,
                        const __half** hi, const __half** lo, int* Kp_out, cudaStream_t st) {
  const int Kp = (K + 63) / 64 * 64;
  *Kp_out = Kp;
  bool is_param = false;
  for (int i = 0; i < XG_NUM_PARAMS && !is_param; ++i) {
    int pr, pc;
    param_shape(ctx->d, i, &pr, &pc);
    const float* b = ctx->P[i];
    if (b && X >= b && X < b + (long)pr * pc) is_param = true;
  }
  if (is_param) {
    const TcState::SplitKey key(X, sr, sk, rows, K);
    auto it = ts->weight_cache16.find(key);
    if (it == ts->weight_cache16.end()) {
      TcState::Split16 sp{nullptr, nullptr, rows, K, Kp, false};
      XG_CUDA_TRY(ctx->es, cudaMalloc(&sp.hi, sizeof(__half) * (size_t)rows * Kp));
      XG_CUDA_TRY(ctx->es, cudaMalloc(&sp.lo, sizeof(__half) * (size_t)rows * Kp));
      it = ts->weight_cache16.emplace(key, sp).first;
    }
    if (!it->second.valid) {
      XG_TRY(tc_split16(ctx, X, sr, sk, rows, K, Kp, it->second.hi, it->second.lo, st));
      it->second.valid = true;
    }
    *hi = it->second.hi;
    *lo = it->second.lo;
    return XG_OK;
  }
  const size_t need = (size_t)rows * Kp * 2;
  if (ts->scratch16_halfs[slot] < need) {
    XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
    if (ts->scratch16[slot]) cudaFree(ts->scratch16[slot]);
    ts->scratch16[slot] = nullptr;
    ts->scratch16_halfs[slot] = 0;
    XG_CUDA_TRY(ctx->es, cudaMalloc(&ts->scratch16[slot], sizeof(__half) * need));
    ts->scratch16_halfs[slot] = need;
  }
  __half* h = ts->scratch16[slot];
  __half* l = h + (size_t)rows * Kp;
  XG_TRY(tc_split16(ctx, X, sr, sk, rows, K, Kp, h, l, st));
  *hi = h;
  *lo = l;
  return XG_OK;
}

// hold regions, see TcState::held
static void tc_hold_begin(xg_context* ctx) {
  TcState* ts = tc_state(ctx);
  if (ts) { ts->hold = true; for (auto& h : ts->held) h.valid = false; }
}
static void tc_hold_end(xg_context* ctx) {
  TcState* ts = tc_state(ctx);
  if (ts) { ts->hold = false; for (auto& h : ts->held) h.valid = false; }
}
struct TcHold {          // scope guard
  xg_context* ctx;
  explicit TcHold(xg_context* c) : ctx(c) { tc_hold_begin(c); }
  ~TcHold() { tc_hold_end(ctx); }
};

template <int BN, bool F16>
static int tc_launch(xg_context* ctx, TcState* ts, int cfg_idx, const CUtensorMap& a, const CUtensorMap& b,
                     const CUtensorMap& c, const CUtensorMap& d, const TcParams& prm, cudaStream_t st) {
  if (!ts->attr_set[cfg_idx]) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(gemm_tc_kernel<BN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              TcCfg<BN>::kSmemBytes));
    ts->attr_set[cfg_idx] = true;
  }
  dim3 grid(ceil_div(prm.Pn, 128), ceil_div(prm.Qn, BN));
  gemm_tc_kernel<BN, F16><<<grid, TcCfg<BN>::kThreads, TcCfg<BN>::kSmemBytes, st>>>(a, b, c, d, prm);
  XG_LAUNCH_CHECK(ctx->es);
  return XG_OK;
}

// shape gate: where the tensor-core engine pays (one 128-lane tile must not be mostly padding)
// One CTA streams a whole 128-row x K operand slab, so the engine only pays when the output has enough
// 128 x 64 tiles to occupy a good part of the 148 SMs (measured: 64x2048x512 -> 16 tiles -> 77 us vs
// 23 us on the SIMT engine).  Skinny recurrent products stay on the SIMT engine until the fused
// split-K word-step kernel takes them over.
static bool tc_eligible(const GemmP& p) {
  const long tiles = (long)ceil_div(p.N, 128) * ceil_div(p.M, 64);
  // (32 tiles are enough when the reduction is long: the 512 x 512 x 1792 weight gradients of the cross gates took 58 us on
  //  the SIMT engine)
  return p.K >= 64 && (tiles >= 56 || (tiles >= 32 && p.K >= 1536)) && (long)p.M * p.N * p.K >= (1L << 24);
}

// C (M,N) = epi( A . B ) in the GemmP convention (A(i,r), B(r,j), arbitrary strides).
// Orientation: the side with more rows rides the 128 TMEM lanes unless it is the activation side of a
// forward Linear (weights on the lanes => coalesced row-major stores and per-lane bias).
static int gemm_tc(xg_context* ctx, const GemmP& p, cudaStream_t st) {
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  // operand "I" : rows i (M), reduction r ;  operand "J" : rows j (N), reduction r
  const int BNsel = (p.M <= 64 || (long)ceil_div(p.N, 128) * ceil_div(p.M, 128) < 120) ? 64 : 128;
  if (ctx->tc_f16) {
    // forward products: fp16 operand pairs (half the operand bytes of the tf32 pairs, k-blocks of 64)
    const __half *ih, *il, *jh, *jl;
    int Kp = 0, Kp2 = 0;
    XG_TRY(tc_operand16(ctx, ts, 0, p.A, p.sa_i, p.sa_r, p.M, p.K, &ih, &il, &Kp, st));
    XG_TRY(tc_operand16(ctx, ts, 1, p.B, p.sb_j, p.sb_r, p.N, p.K, &jh, &jl, &Kp2, st));
    TcParams prm;
    prm.K = Kp;
    prm.g = p;
    { const char* e = getenv("XG_TC_DEBUG"); prm.dbg = e ? atoi(e) : 0; }
    prm.Pn = p.N; prm.Qn = p.M; prm.swap_out = 1;
    CUtensorMap mPh, mPl, mQh, mQl;
    XG_TRY(tc_make_map16(ctx, ts, jh, p.N, Kp, 128, &mPh));
    XG_TRY(tc_make_map16(ctx, ts, jl, p.N, Kp, 128, &mPl));
    XG_TRY(tc_make_map16(ctx, ts, ih, p.M, Kp, BNsel, &mQh));
    XG_TRY(tc_make_map16(ctx, ts, il, p.M, Kp, BNsel, &mQl));
    char tag[96];
    if (ctx->prof_on) {
      const char* lay = (p.sa_r == 1) ? (p.sb_r == 1 ? "nt" : "nn") : "tn";
      snprintf(tag, sizeof(tag), "gemmtc16_%s_%dx%dx%d", lay, p.M, p.N, p.K);
    } else {
      tag[0] = 0;
    }
    ProfScope ps(ctx, std::string(tag), st);
    if (BNsel == 64) return tc_launch<64, true>(ctx, ts, 2, mPh, mPl, mQh, mQl, prm, st);
    return tc_launch<128, true>(ctx, ts, 3, mPh, mPl, mQh, mQl, prm, st);
  }
  const float *ih, *il, *jh, *jl;
  int Kp = 0, Kp2 = 0;
  XG_TRY(tc_operand(ctx, ts, 0, p.A, p.sa_i, p.sa_r, p.M, p.K, &ih, &il, &Kp, st));
  XG_TRY(tc_operand(ctx, ts, 1, p.B, p.sb_j, p.sb_r, p.N, p.K, &jh, &jl, &Kp2, st));
  TcParams prm;
  prm.K = Kp;
  prm.g = p;
  { const char* e = getenv("XG_TC_DEBUG"); prm.dbg = e ? atoi(e) : 0; }
  // lanes <- J (output columns) so that stores are coalesced along j; columns <- I
  prm.Pn = p.N;
  prm.Qn = p.M;
  prm.swap_out = 1;
  const int BN = BNsel;
  CUtensorMap mPh, mPl, mQh, mQl;
  XG_TRY(tc_make_map(ctx, ts, jh, p.N, Kp, 128, &mPh));
  XG_TRY(tc_make_map(ctx, ts, jl, p.N, Kp, 128, &mPl));
  XG_TRY(tc_make_map(ctx, ts, ih, p.M, Kp, BN, &mQh));
  XG_TRY(tc_make_map(ctx, ts, il, p.M, Kp, BN, &mQl));
  char tag[96];
  if (ctx->prof_on) {
    const char* lay = (p.sa_r == 1) ? (p.sb_r == 1 ? "nt" : "nn") : "tn";
    snprintf(tag, sizeof(tag), "gemmtc_%s_%dx%dx%d", lay, p.M, p.N, p.K);
  } else {
    tag[0] = 0;
  }
  ProfScope ps(ctx, std::string(tag), st);
  if (BN == 64) return tc_launch<64, false>(ctx, ts, 0, mPh, mPl, mQh, mQl, prm, st);
  return tc_launch<128, false>(ctx, ts, 1, mPh, mPl, mQh, mQl, prm, st);
}

}  // namespace xg
