// Persistent recurrent kernels: a whole serial loop of the path in ONE cooperative launch, one CTA per SM.
//
//   decode_persistent_kernel  — the greedy word loop of SAModel.sample (SAModel.py:182-219)
//   encode_persistent_kernel  — the frame recurrence of EncoderLstm_two_fc.forward (sub_modules.py:132-147)
//
// Shared machinery ("work items over all SMs"):
//  * every dense product of a phase is cut into (128 weight rows) x (64 caption columns) x (run of 32-wide
//    k-blocks) work items.  The host lays all k-blocks of a phase end to end and deals every CTA the same
//    number of them (+-1), so a CTA's share is 1-3 items of a few k-blocks each (PSched).  Split-K partial
//    tiles go to numbered global "slots"; the consuming pointwise phase adds them in slot order
//    (deterministic, no atomics).
//  * weights are streamed as they lie in the nn.Parameter storage (fp32, row-major = K-major): one TMA box
//    of 128 rows x 32 floats per k-block.  The tensor core reads the top 19 bits of each word (tf32 "hi"
//    by truncation); four split warps compute lo = rna_tf32(w - trunc(w)) into a second smem tile while
//    the next boxes are in flight, so the 3xTF32 product hi*hi + lo*hi + hi*lo costs 52.6 MB of weight
//    traffic per word step instead of 105 MB of pre-split copies — the whole working set stays in L2.
//  * activations (64 rows) are written pre-split (rna hi / lo) by the pointwise phases.
//  * MMA issue / accumulation discipline is the one of xg_gemm_tc.cuh (short hi*hi chains in ping-pong
//    TMEM accumulators promoted into fp32 registers, cross terms in their own accumulator).
//
// Warp roles (320 threads): 0 TMA producer, 1 MMA issuer, 2-5 epilogue (TMEM lane quads), 6-9 weight split.
#pragma once
#include <algorithm>
#include <vector>

#include "xg_gemm_tc.cuh"

namespace xg {

constexpr int PK_FALLBACK = -1;      // host: shape cannot be scheduled -> caller uses the unfused path
constexpr int PK_BN = 64;
constexpr int PK_STAGES = 4;
constexpr int PK_W_BYTES = 128 * 128;                            // 128 rows x 32 fp32
constexpr int PK_X_BYTES = PK_BN * 128;
constexpr int PK_STAGE_BYTES = 2 * PK_W_BYTES + 2 * PK_X_BYTES;  // W raw | W lo | X hi | X lo  (48 KB)
constexpr int PK_TX_BYTES = PK_W_BYTES + 2 * PK_X_BYTES;         // what TMA delivers per stage
constexpr int PK_CHUNK = 2;
constexpr int PK_THREADS = 320;
constexpr int PK_WARPS = PK_THREADS / 32;
constexpr int PK_SCRATCH_FLOATS = 2304;
constexpr int PK_SMEM_BYTES = PK_STAGES * PK_STAGE_BYTES + PK_SCRATCH_FLOATS * 4 + 1024 + 512;
constexpr int PK_MAX_ITEMS = 6;
constexpr int PK_MAX_SLOTS = 6;
constexpr int PK_MAX_DESCS = 8;
constexpr int PK_BULK_CHUNKS = 4;
constexpr int PK_STAMPS = 16;

struct PItem { short desc, slot, rt, cb, kb0, nkb; };
struct PSched { int n; PItem it[PK_MAX_ITEMS]; };

struct GDesc {                   // out[slot][r][n] = sum_k W[n,k] * X[r, 32*xkb0 + k]
  int w_map, x_hi, x_lo;         // indices into the tensor-map table
  int xkb0, n_rows, nkb;
  float* out;                    // [slots][R][n_rows]
  const unsigned char* nslots;   // [row tiles][column blocks]: slots written for that strip
};

struct MapTable { CUtensorMap m[16]; };

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// grid-wide barrier on a monotonically increasing counter (cooperative launch guarantees residency)
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, int G) {
  fence_proxy_async_global();          // generic-proxy global writes -> visible to later TMA reads
  __syncthreads();
  if (threadIdx.x == 0) {
    target += (unsigned)G;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    const long long t0 = clock64();
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((int)(v - target) >= 0) break;
      if (clock64() - t0 > 8000000000LL) __trap();
    }
  }
  __syncthreads();
  fence_proxy_async_global();
}

struct PipeState {       // running counters, identical in every thread by construction
  uint32_t kb_count;     // k-blocks issued so far (stage ring position)
  uint32_t chunk_count;  // accumulation chains issued so far (ping-pong accumulator position)
  uint32_t item_count;   // items processed so far (small-accumulator handshake)
};

struct SmemView {
  uint8_t* stages;
  float* scratch;
  uint64_t *full_bar, *empty_bar, *split_bar, *acc_full, *acc_empty, *small_full, *small_empty, *bulk_bar;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ SmemView carve_smem(uint8_t* smem_raw) {
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SmemView sv;
  sv.stages = smem;
  sv.scratch = reinterpret_cast<float*>(smem + PK_STAGES * PK_STAGE_BYTES);
  sv.full_bar = reinterpret_cast<uint64_t*>(smem + PK_STAGES * PK_STAGE_BYTES + PK_SCRATCH_FLOATS * 4);
  sv.empty_bar = sv.full_bar + PK_STAGES;
  sv.split_bar = sv.empty_bar + PK_STAGES;
  sv.acc_full = sv.split_bar + PK_STAGES;
  sv.acc_empty = sv.acc_full + 2;
  sv.small_full = sv.acc_empty + 2;
  sv.small_empty = sv.small_full + 1;
  sv.bulk_bar = sv.small_empty + 1;
  sv.tmem_slot = reinterpret_cast<uint32_t*>(sv.bulk_bar + PK_BULK_CHUNKS);
  return sv;
}

__device__ __forceinline__ uint32_t pipeline_setup(const SmemView& sv) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < PK_STAGES; ++s) {
      mbar_init(&sv.full_bar[s], 1); mbar_init(&sv.empty_bar[s], 1); mbar_init(&sv.split_bar[s], 128);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(&sv.acc_full[b], 1); mbar_init(&sv.acc_empty[b], 128); }
    mbar_init(sv.small_full, 1);
    mbar_init(sv.small_empty, 128);
    for (int c = 0; c < PK_BULK_CHUNKS; ++c) mbar_init(&sv.bulk_bar[c], 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<256>(sv.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sv.tmem_slot;
}

__device__ __forceinline__ void pipeline_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 2) tmem_dealloc<256>(tmem_base);
}

// all work items of this CTA for one GEMM phase
__device__ __noinline__ void gemm_phase(const GDesc* descs, const PSched* sc, const CUtensorMap* maps, int R, int hi_inplace,
                                        const SmemView& sv, uint32_t tmem_base, PipeState& ps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = sc->n;
  for (int ii = 0; ii < n_items; ++ii) {
    const PItem it = sc->it[ii];
    const GDesc& d = descs[it.desc];
    const int nkb = it.nkb, kb0 = it.kb0;
    const int n_chunks = (nkb + PK_CHUNK - 1) / PK_CHUNK;
    if (warp == 0) {
      if (lane == 0) {
        const CUtensorMap* mw = maps + d.w_map; const CUtensorMap* mxh = maps + d.x_hi; const CUtensorMap* mxl = maps + d.x_lo;
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t cnt = ps.kb_count + kb;
          const int s = cnt % PK_STAGES;
          mbar_wait(&sv.empty_bar[s], ((cnt / PK_STAGES) & 1) ^ 1);
          uint8_t* st = sv.stages + s * PK_STAGE_BYTES;
          mbar_expect_tx(&sv.full_bar[s], PK_TX_BYTES);
          tma_load_2d(st, mw, &sv.full_bar[s], (kb0 + kb) * 32, it.rt * 128);
          tma_load_2d(st + 2 * PK_W_BYTES, mxh, &sv.full_bar[s], (d.xkb0 + kb0 + kb) * 32, it.cb * PK_BN);
          tma_load_2d(st + 2 * PK_W_BYTES + PK_X_BYTES, mxl, &sv.full_bar[s], (d.xkb0 + kb0 + kb) * 32, it.cb * PK_BN);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_tf32(128, PK_BN);
        const uint32_t tmem_small = tmem_base + 2 * PK_BN;
        mbar_wait(sv.small_empty, (ps.item_count & 1) ^ 1);   // previous item's cross-term accumulator read out
        tc_fence_after();
        for (int c = 0; c < n_chunks; ++c) {
          const uint32_t cc = ps.chunk_count + c;
          const int b = cc & 1;
          mbar_wait(&sv.acc_empty[b], ((cc >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_main = tmem_base + b * PK_BN;
          for (int kk = 0; kk < PK_CHUNK; ++kk) {
            const int kb = c * PK_CHUNK + kk;
            if (kb >= nkb) break;
            const uint32_t cnt = ps.kb_count + kb;
            const int s = cnt % PK_STAGES;
            mbar_wait(&sv.split_bar[s], (cnt / PK_STAGES) & 1);     // TMA landed AND lo tile written
            tc_fence_after();
            const uint32_t base = smem_u32(sv.stages + s * PK_STAGE_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t wh = umma_desc_sw128(base + k4 * 32);
              const uint64_t wl = umma_desc_sw128(base + PK_W_BYTES + k4 * 32);
              const uint64_t xh = umma_desc_sw128(base + 2 * PK_W_BYTES + k4 * 32);
              const uint64_t xl = umma_desc_sw128(base + 2 * PK_W_BYTES + PK_X_BYTES + k4 * 32);
              umma_tf32(tmem_main, wh, xh, idesc, (kk | k4) != 0);
              umma_tf32(tmem_small, wl, xh, idesc, (kb | k4) != 0);
              umma_tf32(tmem_small, wh, xl, idesc, 1);
            }
            umma_commit(&sv.empty_bar[s]);
          }
          umma_commit(&sv.acc_full[b]);
        }
        umma_commit(sv.small_full);
      }
    } else if (warp < 6) {
      const int quad = warp & 3;
      const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
      float acc[PK_BN];
#pragma unroll
      for (int u = 0; u < PK_BN; ++u) acc[u] = 0.f;
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t cc = ps.chunk_count + c;
        const int b = cc & 1;
        mbar_wait(&sv.acc_full[b], (cc >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < PK_BN; q += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_base + (uint32_t)(b * PK_BN + q), r);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[q + u] += __uint_as_float(r[u]);
        }
        tc_fence_before();
        mbar_arrive(&sv.acc_empty[b]);
      }
      mbar_wait(sv.small_full, ps.item_count & 1);
      tc_fence_after();
#pragma unroll
      for (int q = 0; q < PK_BN; q += 32) {
        uint32_t r[32];
        tmem_ld32(tmem_base + 2 * PK_BN + lane_base + (uint32_t)q, r);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 32; ++u) acc[q + u] += __uint_as_float(r[u]);
      }
      tc_fence_before();
      mbar_arrive(sv.small_empty);
      const int n = it.rt * 128 + quad * 32 + lane;        // weight row held by this thread
      if (n < d.n_rows) {
        float* o = d.out + ((long)it.slot * R + it.cb * PK_BN) * d.n_rows + n;   // lanes -> consecutive n
#pragma unroll
        for (int u = 0; u < PK_BN; ++u) __stcg(o + (long)u * d.n_rows, acc[u]);
      }
    } else {
      // weight split: lo = rna_tf32(w - trunc_tf32(w)); the tensor core itself truncates the raw tile to hi
      const int t = threadIdx.x - 6 * 32;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t cnt = ps.kb_count + kb;
        const int s = cnt % PK_STAGES;
        mbar_wait(&sv.full_bar[s], (cnt / PK_STAGES) & 1);
        float4* src = reinterpret_cast<float4*>(sv.stages + s * PK_STAGE_BYTES);
        float4* dst = reinterpret_cast<float4*>(sv.stages + s * PK_STAGE_BYTES + PK_W_BYTES);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v = src[t + 128 * q];
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = tf32_rna(v.x - h.x);
          h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = tf32_rna(v.y - h.y);
          h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = tf32_rna(v.z - h.z);
          h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = tf32_rna(v.w - h.w);
          dst[t + 128 * q] = l;
          if (hi_inplace) src[t + 128 * q] = h;
        }
        fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor core
        mbar_arrive(&sv.split_bar[s]);
      }
    }
    ps.kb_count += nkb;
    ps.chunk_count += n_chunks;
    ps.item_count += 1;
  }
}

__device__ __forceinline__ void store_split(float* hi, float* lo, long idx, float v) {
  const float h = tf32_rna(v);
  hi[idx] = h;
  lo[idx] = tf32_rna(v - h);
}

// sum of the split-K partial slots of element (r, n), in slot order.  All loads are issued before the
// first add (one L2 round trip per element instead of one per slot).
__device__ __forceinline__ float zsum(const GDesc& d, int R, int r, int n) {
  const int ncb = R / PK_BN;
  const int ns = d.nslots[(n >> 7) * ncb + r / PK_BN];
  const float* p = d.out + (long)r * d.n_rows + n;
  const long sstr = (long)R * d.n_rows;
  float v[PK_MAX_SLOTS];
#pragma unroll
  for (int k = 0; k < PK_MAX_SLOTS; ++k) v[k] = k < ns ? __ldcg(p + k * sstr) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < PK_MAX_SLOTS; ++k) s += v[k];
  return s;
}

__device__ __forceinline__ unsigned f2ord(float f) {      // order-preserving float -> uint (for redux.sync)
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// tanh via one ex2.approx and one fast division: |error| ~1e-7 absolute (the attention scores only)
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// ====================================================================================
// decoder
// ====================================================================================
enum { DD_AH = 0, DD_Z1H, DD_Z2H, DD_Z1X, DD_Z1G, DD_Z2X, DD_Z2A, DD_LOGIT, DD_COUNT };

struct DecParams {
  GDesc d[PK_MAX_DESCS];
  const PSched* sched;          // [3][G]
  int B, R, K, H, E, Ep, A, V, T, hi_inplace;
  const float *b_h2a, *w_a2w, *b_a2w;
  const float *b1_i2h, *b1_a2h, *b1_h2h, *b2_i2h, *b2_a2h, *b2_h2h, *b_logit, *embed;
  const float* tgate;           // (V, H)  relu(embed . W_gate^T + b): the POS-gate pre-factor of every token
  const float *Vf, *Uv, *pos;   // (B,K,H), (B,K,A), (B,H)
  const float* state0[4];       // h1,c1,h2,c2 each (B,H)
  float *xt_hi, *xt_lo;         // [R][Ep]
  float *hh_hi, *hh_lo;         // [R][2H]   [h1 | h2]
  float *gp_hi, *gp_lo;         // [R][H]
  float *af_hi, *af_lo;         // [R][H]
  float *hx;                    // [R][2H] exact states
  float *c1, *c2;               // [R][H]
  float *stats;                 // [vocab blocks][R][4]  (max, argmax, sum-exp, -)
  float *unfinished;            // [R]
  int64_t *tok;                 // [R]
  int64_t* seq; float* seqlogp; int* flags;   // (B,T), (B,T), (T)
  unsigned int* sync_counter;
  long long* dbg_clock;         // [T][PK_STAMPS] SM-clock stamps of CTA 0 at phase boundaries, or NULL
};

constexpr int DEC_VBLOCK = 1024;   // vocabulary rows per statistics record

// lstm cell (decoder gate order i,f,o,g), elements (r, j) with j fastest
__device__ __forceinline__ void dec_cell_phase(const DecParams& P, int layer, bool use_mask, int part, int nparts) {
  const int H = P.H, R = P.R;
  const GDesc& da = layer == 0 ? P.d[DD_Z1X] : P.d[DD_Z2X];
  const GDesc& db = layer == 0 ? P.d[DD_Z1G] : P.d[DD_Z2A];
  const GDesc& dc = layer == 0 ? P.d[DD_Z1H] : P.d[DD_Z2H];
  const float* bi = layer == 0 ? P.b1_i2h : P.b2_i2h;
  const float* ba = layer == 0 ? P.b1_a2h : P.b2_a2h;
  const float* bh = layer == 0 ? P.b1_h2h : P.b2_h2h;
  float* cst = layer == 0 ? P.c1 : P.c2;
  for (int e = part * PK_THREADS + threadIdx.x; e < P.B * H; e += nparts * PK_THREADS) {
    const int r = e / H, j = e % H;
    float z[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = g * H + j;
      z[g] = zsum(da, R, r, n) + zsum(db, R, r, n) + zsum(dc, R, r, n) + (__ldg(bi + n) + __ldg(ba + n) + __ldg(bh + n));
    }
    const float ig = sigmoid_f(z[0]), fg = sigmoid_f(z[1]), og = sigmoid_f(z[2]), gg = tanhf(z[3]);
    const float m = use_mask ? __ldcg(P.unfinished + r) : 1.f;
    const float cp = cst[e];
    const float hp = P.hx[(long)r * 2 * H + layer * H + j];
    float c = fg * cp + ig * gg;
    c = c * m + cp * (1.f - m);
    float h = og * tanhf(c);
    h = h * m + hp * (1.f - m);
    cst[e] = c;
    P.hx[(long)r * 2 * H + layer * H + j] = h;
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + layer * H + j, h);
  }
}

// temporal attention of caption r on one CTA: Uv[r] arrives by four bulk copies into the (idle) pipeline
// stages while ah is assembled; scores -> softmax over ALL K frames -> context (sub_modules.py:677-680)
__device__ __forceinline__ void dec_attention(const DecParams& P, int r, const SmemView& sv, uint32_t& bulk_phase) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = P.K, A = P.A, H = P.H;
  float* ah = sv.scratch;                       // A floats
  float* sc = sv.scratch + A;                   // K floats
  float* uv = reinterpret_cast<float*>(sv.stages);
  const int fpc = (K + PK_BULK_CHUNKS - 1) / PK_BULK_CHUNKS;
  if (threadIdx.x == 0) {
    for (int c = 0; c < PK_BULK_CHUNKS; ++c) {
      const int k0 = c * fpc, k1 = min(K, k0 + fpc);
      if (k0 >= k1) break;
      const uint32_t nb = (uint32_t)(k1 - k0) * (uint32_t)A * 4u;
      mbar_expect_tx(&sv.bulk_bar[c], nb);
      bulk_g2s(uv + (long)k0 * A, P.Uv + ((long)r * K + k0) * A, nb, &sv.bulk_bar[c]);
    }
  }
  for (int a = threadIdx.x; a < A; a += PK_THREADS) ah[a] = zsum(P.d[DD_AH], P.R, r, a) + __ldg(P.b_h2a + a);
  __syncthreads();
  const float ba = __ldg(P.b_a2w);
  for (int c = 0; c < PK_BULK_CHUNKS; ++c) {
    const int k0 = c * fpc, k1 = min(K, k0 + fpc);
    if (k0 >= k1) break;
    mbar_wait(&sv.bulk_bar[c], bulk_phase & 1);
    for (int k = k0 + warp; k < k1; k += PK_WARPS) {
      const float* u = uv + (long)k * A;
      float p = 0.f;
#pragma unroll 4
      for (int a = lane; a < A; a += 32) p += __ldg(P.w_a2w + a) * tanh_fast(ah[a] + u[a]);
      p = warp_sum(p);
      if (lane == 0) sc[k] = p + ba;
    }
  }
  bulk_phase++;
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int k = lane; k < K; k += 32) mx = fmaxf(mx, sc[k]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < K; k += 32) { const float e = expf(sc[k] - mx); sc[k] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int k = lane; k < K; k += 32) sc[k] *= inv;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < H; j += PK_THREADS) {
    const float* v = P.Vf + (long)r * K * H + j;
    float a = 0.f;
    int k = 0;
    for (; k + 4 <= K; k += 4) {
      const float v0 = __ldg(v + (long)k * H), v1 = __ldg(v + (long)(k + 1) * H), v2 = __ldg(v + (long)(k + 2) * H),
                  v3 = __ldg(v + (long)(k + 3) * H);
      a += sc[k] * v0; a += sc[k + 1] * v1; a += sc[k + 2] * v2; a += sc[k + 3] * v3;
    }
    for (; k < K; ++k) a += sc[k] * __ldg(v + (long)k * H);
    store_split(P.af_hi, P.af_lo, (long)r * H + j, a);
  }
  __syncthreads();
}

// next-step inputs of caption r for token `tokv`: xt = embed[tok] and gp = pos * (1 + tgate[tok])
__device__ __forceinline__ void dec_token_inputs(const DecParams& P, int r, int tokv) {
  const float* src = P.embed + (long)tokv * P.E;
  for (int k = threadIdx.x; k < P.Ep; k += PK_THREADS)
    store_split(P.xt_hi, P.xt_lo, (long)r * P.Ep + k, k < P.E ? __ldg(src + k) : 0.f);
  const float* tg = P.tgate + (long)tokv * P.H;
  for (int j = threadIdx.x; j < P.H; j += PK_THREADS)
    store_split(P.gp_hi, P.gp_lo, (long)r * P.H + j, __ldg(P.pos + (long)r * P.H + j) * (1.f + __ldcg(tg + j)));
}

__global__ void __launch_bounds__(PK_THREADS, 1)
decode_persistent_kernel(const DecParams* __restrict__ Pp, const __grid_constant__ MapTable maps) {
  __shared__ DecParams Psm;
  for (int i = threadIdx.x; i < (int)(sizeof(DecParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Psm)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  __syncthreads();
  const DecParams& P = Psm;
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int H = P.H, R = P.R, B = P.B, T = P.T, V = P.V;
  const uint32_t tmem_base = pipeline_setup(sv);
  PipeState ps{0, 0, 0};
  unsigned int sync_target = 0;
  uint32_t bulk_phase = 0;
  const PSched* my_sched = P.sched + cta;

  // ---- prologue: states, <bos> inputs, bookkeeping ----
  for (int e = cta * PK_THREADS + threadIdx.x; e < R * H; e += G * PK_THREADS) {
    const int r = e / H, j = e % H;
    const float h1 = r < B ? P.state0[0][e] : 0.f, c1 = r < B ? P.state0[1][e] : 0.f;
    const float h2 = r < B ? P.state0[2][e] : 0.f, c2 = r < B ? P.state0[3][e] : 0.f;
    P.hx[(long)r * 2 * H + j] = h1; P.hx[(long)r * 2 * H + H + j] = h2;
    P.c1[e] = c1; P.c2[e] = c2;
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + j, h1);
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + H + j, h2);
  }
  for (int r = cta; r < R; r += G) {
    if (r < B) {
      dec_token_inputs(P, r, 0);                      // token 0 = <bos> (SAModel.py:184)
    } else {                                          // padding rows of the 64-wide operand tiles
      for (int k = threadIdx.x; k < P.Ep; k += PK_THREADS) { P.xt_hi[(long)r * P.Ep + k] = 0.f; P.xt_lo[(long)r * P.Ep + k] = 0.f; }
      for (int j = threadIdx.x; j < H; j += PK_THREADS) {
        P.gp_hi[(long)r * H + j] = 0.f; P.gp_lo[(long)r * H + j] = 0.f;
        P.af_hi[(long)r * H + j] = 0.f; P.af_lo[(long)r * H + j] = 0.f;
      }
    }
    if (threadIdx.x == 0) { P.unfinished[r] = 1.f; P.tok[r] = 0; }
  }
  grid_barrier(P.sync_counter, sync_target, G);

#define PK_STAMP(i) do { if (P.dbg_clock && cta == 0 && threadIdx.x == 0) P.dbg_clock[t * PK_STAMPS + (i)] = clock64(); } while (0)
  for (int t = 0; t < T; ++t) {
    PK_STAMP(0);
    // ===== G1: everything that needs only the previous state and the current token =====
    //   AH = W_h2a.[h1|h2]   Z1h = W_h2h1.h1   Z2h = W_h2h2.h2   Z1x = W_i2h1.xt   Z1g = W_a2h1.gp
    gemm_phase(P.d, my_sched, maps.m, R, P.hi_inplace, sv, tmem_base, ps);
    PK_STAMP(1);
    grid_barrier(P.sync_counter, sync_target, G);
    PK_STAMP(2);
    // ===== P1: attention (one CTA per caption)  ||  lstm_1 cell (the other CTAs) =====
    if (G > B) {
      if (cta < B) dec_attention(P, cta, sv, bulk_phase);
      else dec_cell_phase(P, 0, t > 0, cta - B, G - B);
    } else {
      for (int r = cta; r < B; r += G) dec_attention(P, r, sv, bulk_phase);
      dec_cell_phase(P, 0, t > 0, cta, G);
    }
    fence_proxy_async_smem();      // stages were read/written through the generic + bulk paths: order before TMA reuse
    PK_STAMP(3);
    grid_barrier(P.sync_counter, sync_target, G);
    PK_STAMP(4);
    // ===== G3: Z2x = W_i2h2.h1'   Z2a = W_a2h2.af =====
    gemm_phase(P.d, my_sched + G, maps.m, R, P.hi_inplace, sv, tmem_base, ps);
    PK_STAMP(5);
    grid_barrier(P.sync_counter, sync_target, G);
    PK_STAMP(6);
    // ===== P3: lstm_2 cell =====
    dec_cell_phase(P, 1, t > 0, cta, G);
    PK_STAMP(7);
    grid_barrier(P.sync_counter, sync_target, G);
    PK_STAMP(8);
    // ===== G4: logits (split-K partial tiles) =====
    gemm_phase(P.d, my_sched + 2 * G, maps.m, R, P.hi_inplace, sv, tmem_base, ps);
    PK_STAMP(9);
    grid_barrier(P.sync_counter, sync_target, G);
    PK_STAMP(10);
    // ===== P4a: per (caption, block of 1024 vocabulary rows): max / lowest argmax / sum-exp =====
    {
      const int nvb = (V + DEC_VBLOCK - 1) / DEC_VBLOCK;
      const GDesc& dl = P.d[DD_LOGIT];
      for (int item = cta * PK_WARPS + warp; item < B * nvb; item += G * PK_WARPS) {
        const int r = item / nvb, vb = item % nvb;
        float v[DEC_VBLOCK / 32];
#pragma unroll
        for (int i = 0; i < DEC_VBLOCK / 32; ++i) {
          const int n = vb * DEC_VBLOCK + i * 32 + lane;
          v[i] = n < V ? zsum(dl, R, r, n) + __ldg(P.b_logit + n) : -INFINITY;
        }
        float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < DEC_VBLOCK / 32; ++i)
          if (v[i] > best) { best = v[i]; bi = vb * DEC_VBLOCK + i * 32 + lane; }
        const unsigned mo = __reduce_max_sync(0xffffffffu, f2ord(best));
        const float wbest = ord2f(mo);
        const int wbi = (int)__reduce_min_sync(0xffffffffu, (f2ord(best) == mo && best != -INFINITY) ? (unsigned)bi : 0x7fffffffu);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < DEC_VBLOCK / 32; ++i) s += (v[i] == -INFINITY) ? 0.f : expf(v[i] - wbest);
        s = warp_sum(s);
        if (lane == 0) {
          float4 rec = make_float4(wbest, __int_as_float(wbi), s, 0.f);
          *reinterpret_cast<float4*>(P.stats + ((long)vb * R + r) * 4) = rec;
        }
      }
    }
    PK_STAMP(11);
    grid_barrier(P.sync_counter, sync_target, G);
    PK_STAMP(12);
    // ===== P4b: greedy bookkeeping + inputs of the next step (SAModel.py:185-210) =====
    {
      const int nvb = (V + DEC_VBLOCK - 1) / DEC_VBLOCK;
      for (int r = cta; r < B; r += G) {
        if (warp == 0) {
          float best = -INFINITY; int bi = 0x7fffffff;
          for (int q = lane; q < nvb; q += 32) {
            const float4 rec = __ldcg(reinterpret_cast<const float4*>(P.stats + ((long)q * R + r) * 4));
            const int vi = __float_as_int(rec.y);
            if (rec.x > best || (rec.x == best && vi < bi)) { best = rec.x; bi = vi; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
          }
          float s = 0.f;
          for (int q = lane; q < nvb; q += 32) {
            const float4 rec = __ldcg(reinterpret_cast<const float4*>(P.stats + ((long)q * R + r) * 4));
            if (rec.x != -INFINITY) s += rec.z * expf(rec.x - best);
          }
          s = warp_sum(s);
          if (lane == 0) {
            float unf = (t == 0) ? 1.f : __ldcg(P.unfinished + r);
            unf = (unf != 0.f && bi > 0) ? 1.f : 0.f;
            P.unfinished[r] = unf;
            P.seq[(long)r * T + t] = unf != 0.f ? (int64_t)bi : 0;
            P.seqlogp[(long)r * T + t] = -logf(s);
            P.tok[r] = bi;
            if (unf != 0.f) P.flags[t] = 1;
            reinterpret_cast<int*>(sv.scratch)[0] = bi;
          }
        }
        __syncthreads();
        const int tokv = reinterpret_cast<int*>(sv.scratch)[0];
        dec_token_inputs(P, r, tokv);
        __syncthreads();
      }
    }
    PK_STAMP(13);
    grid_barrier(P.sync_counter, sync_target, G);
    PK_STAMP(14);
    if (__ldcg(P.flags + t) == 0) break;     // every caption finished (SAModel.py:206)
  }
  pipeline_teardown(tmem_base);
}

// ====================================================================================
// encoder recurrence (both streams), t = 0..K-1:   z_t = XG_t + W_hh.h_{t-1}   ->  nn.LSTMCell (i,f,g,o)
// ====================================================================================
struct EncParams {
  GDesc d[2];                   // rgb, opfl recurrent products
  const PSched* sched;          // [G]
  int B, R, K, H, hi_inplace;
  float* Gt[2];                 // (K,B,4H) input projections + biases  ->  activated gates (in place)
  float* Hs[2];                 // (K,B,H)
  float* Cs[2];                 // (K,B,H)
  const float* fmask;           // (B,K)
  float *hh_hi, *hh_lo;         // [R][2H]  [h_rgb | h_opfl] of the previous frame
  unsigned int* sync_counter;
};

__global__ void __launch_bounds__(PK_THREADS, 1)
encode_persistent_kernel(const EncParams* __restrict__ Pp, const __grid_constant__ MapTable maps) {
  __shared__ EncParams Psm;
  for (int i = threadIdx.x; i < (int)(sizeof(EncParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Psm)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  __syncthreads();
  const EncParams& P = Psm;
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int cta = blockIdx.x, G = gridDim.x;
  const int H = P.H, R = P.R, B = P.B, K = P.K;
  const uint32_t tmem_base = pipeline_setup(sv);
  PipeState ps{0, 0, 0};
  unsigned int sync_target = 0;

  for (int e = cta * PK_THREADS + threadIdx.x; e < (R - B) * 2 * H; e += G * PK_THREADS) {   // padding rows
    P.hh_hi[(long)B * 2 * H + e] = 0.f; P.hh_lo[(long)B * 2 * H + e] = 0.f;
  }
  for (int t = 0; t < K; ++t) {
    if (t > 0) {
      gemm_phase(P.d, P.sched + cta, maps.m, R, P.hi_inplace, sv, tmem_base, ps);
      grid_barrier(P.sync_counter, sync_target, G);
    }
    for (int e = cta * PK_THREADS + threadIdx.x; e < 2 * B * H; e += G * PK_THREADS) {
      const int s = e / (B * H), b = (e / H) % B, j = e % H;
      float* z = P.Gt[s] + ((long)t * B + b) * 4 * H;
      float zz[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) zz[g] = z[g * H + j] + (t > 0 ? zsum(P.d[s], R, b, g * H + j) : 0.f);
      const float ig = sigmoid_f(zz[0]), fg = sigmoid_f(zz[1]), gg = tanhf(zz[2]), og = sigmoid_f(zz[3]);
      const float m = __ldg(P.fmask + (long)b * K + t);
      const long o = ((long)t * B + b) * H + j;
      const float cp = t > 0 ? P.Cs[s][o - (long)B * H] : 0.f;
      const float c2 = fg * cp + ig * gg;
      const float h = og * tanhf(c2) * m;      // h' *= mask (sub_modules.py:139,146)
      const float c = c2 * m;                  // c' *= mask (:140,147)
      z[j] = ig; z[H + j] = fg; z[2 * H + j] = gg; z[3 * H + j] = og;
      P.Cs[s][o] = c;
      P.Hs[s][o] = h;
      if (t + 1 < K) store_split(P.hh_hi, P.hh_lo, (long)b * 2 * H + s * H + j, h);
    }
    if (t + 1 < K) grid_barrier(P.sync_counter, sync_target, G);
  }
  pipeline_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct PhaseSchedule {
  std::vector<PSched> per_cta;                         // [G]
  std::vector<std::vector<unsigned char>> nslots;      // per desc (phase-local order): [rts * ncb]
  bool ok = true;
};

// lay the k-blocks of all strips (desc, row tile, column block) end to end; CTA c gets units [cU/G, (c+1)U/G)
static PhaseSchedule build_phase(const std::vector<int>& desc_ids, const GDesc* descs, int ncb, int G) {
  PhaseSchedule ph;
  ph.per_cta.assign(G, PSched{});
  struct Strip { int desc, rt, cb, nkb, slots; };
  std::vector<Strip> strips;
  long U = 0;
  for (int id : desc_ids) {
    const int rts = (descs[id].n_rows + 127) / 128;
    for (int rt = 0; rt < rts; ++rt)
      for (int cb = 0; cb < ncb; ++cb) { strips.push_back({id, rt, cb, descs[id].nkb, 0}); U += descs[id].nkb; }
  }
  size_t si = 0; int off = 0;
  for (int c = 0; c < G; ++c) {
    long need = (long)(c + 1) * U / G - (long)c * U / G;
    PSched& sc = ph.per_cta[c];
    while (need > 0 && si < strips.size()) {
      Strip& s = strips[si];
      const int take = (int)std::min<long>(need, s.nkb - off);
      if (sc.n >= PK_MAX_ITEMS || s.slots >= PK_MAX_SLOTS) { ph.ok = false; return ph; }
      sc.it[sc.n++] = PItem{(short)s.desc, (short)s.slots, (short)s.rt, (short)s.cb, (short)off, (short)take};
      s.slots++;
      off += take; need -= take;
      if (off == s.nkb) { ++si; off = 0; }
    }
  }
  ph.nslots.resize(desc_ids.size());
  size_t k = 0;
  for (size_t i = 0; i < desc_ids.size(); ++i) {
    const int rts = (descs[desc_ids[i]].n_rows + 127) / 128;
    ph.nslots[i].resize((size_t)rts * ncb);
    for (size_t q = 0; q < (size_t)rts * ncb; ++q) ph.nslots[i][q] = (unsigned char)strips[k++].slots;
  }
  return ph;
}

struct PersistState {
  // decoder
  int R = 0, K = 0;
  char* pool = nullptr;
  size_t pool_bytes = 0;
  DecParams hp;
  DecParams* d_params = nullptr;
  unsigned int* d_counter = nullptr;
  int* d_flags = nullptr;
  long long* d_dbg = nullptr;
  float* tgate = nullptr;
  unsigned long long tgate_epoch = ~0ull;
  bool attr_set = false;
  // encoder
  int eB = 0;
  char* epool = nullptr;
  size_t epool_bytes = 0;
  EncParams ep;
  EncParams* d_eparams = nullptr;
  unsigned int* d_ecounter = nullptr;
  bool eattr_set = false;
};

inline PersistState*& persist_state(xg_context* ctx) {
  static std::unordered_map<xg_context*, PersistState*> m;
  return m[ctx];
}
static void persist_release(xg_context* ctx) {
  PersistState* s = persist_state(ctx);
  if (!s) return;
  if (s->pool) cudaFree(s->pool);
  if (s->epool) cudaFree(s->epool);
  delete s;
  persist_state(ctx) = nullptr;
}

static inline int env_flag(const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; }

static bool persist_eligible(const xg_context* ctx, int B, int K) {
  const xg_dims& d = ctx->d;
  return ctx->persist_mode && d.rnn % 32 == 0 && d.embed % 4 == 0 && d.att % 4 == 0 && B <= 64 &&
         d.att + K + 8 <= PK_SCRATCH_FLOATS && (long)K * d.att * 4 <= (long)PK_STAGES * PK_STAGE_BYTES && d.vocab >= 2 &&
         d.vocab < 32000 && d.att < 32000 && d.rnn <= 4096;
}

// schedule of one kernel (phases laid one after the other, [phase][G]); false if a phase cannot be scheduled
static bool persist_plan(const std::vector<std::vector<int>>& phases, GDesc* descs, int ncb, int G, std::vector<PSched>& sched,
                         std::vector<std::vector<unsigned char>>& nslots_by_desc) {
  sched.clear();
  for (const auto& ids : phases) {
    PhaseSchedule ph = build_phase(ids, descs, ncb, G);
    if (!ph.ok) return false;
    sched.insert(sched.end(), ph.per_cta.begin(), ph.per_cta.end());
    for (size_t i = 0; i < ids.size(); ++i) nslots_by_desc[ids[i]] = ph.nslots[i];
  }
  return true;
}

static int gemm_run(xg_context* ctx, const GemmP& p, cudaStream_t st);   // xg_forward.cuh

static int persist_greedy(xg_context* ctx, const float* Vf, const float* Uv, const float* pos, const float* const* state0,
                          int B, int K, int T, int64_t* seq_out, float* logp_out, int* steps_out, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, A = d.att, V = d.vocab;
  const int R = 64, Ep = (E + 31) / 32 * 32, G = ctx->sm_count;
  if (T > 2048) return PK_FALLBACK;
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  PersistState*& S = persist_state(ctx);
  if (!S) S = new PersistState();
  DecParams& hp = S->hp;
  const int kbH = H / 32, kbE = Ep / 32;
  const int nvb = (V + DEC_VBLOCK - 1) / DEC_VBLOCK;

  // ---- products ----
  auto mk = [&](int id, int wmap, int xmap, int xkb0, int n_rows, int nkb) {
    GDesc& g = hp.d[id];
    g.w_map = wmap; g.x_hi = xmap; g.x_lo = xmap + 1; g.xkb0 = xkb0; g.n_rows = n_rows; g.nkb = nkb;
  };
  // maps: 0..7 raw weights (h2a, l1_h2h, l2_h2h, l1_i2h, l1_a2h, l2_i2h, l2_a2h, logit); xt 8,9  hh 10,11  gp 12,13  af 14,15
  mk(DD_AH, 0, 10, 0, A, 2 * kbH);
  mk(DD_Z1H, 1, 10, 0, 4 * H, kbH);
  mk(DD_Z2H, 2, 10, kbH, 4 * H, kbH);
  mk(DD_Z1X, 3, 8, 0, 4 * H, kbE);
  mk(DD_Z1G, 4, 12, 0, 4 * H, kbH);
  mk(DD_Z2X, 5, 10, 0, 4 * H, kbH);
  mk(DD_Z2A, 6, 14, 0, 4 * H, kbH);
  mk(DD_LOGIT, 7, 10, kbH, V, kbH);
  const std::vector<std::vector<int>> phases = {{DD_AH, DD_Z1H, DD_Z2H, DD_Z1X, DD_Z1G}, {DD_Z2X, DD_Z2A}, {DD_LOGIT}};
  std::vector<PSched> sched;
  std::vector<std::vector<unsigned char>> nslots(DD_COUNT);
  if (!persist_plan(phases, hp.d, R / PK_BN, G, sched, nslots)) return PK_FALLBACK;

  // ---- device pool ----
  if (S->R != R || S->K != K) {
    if (S->pool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->pool); S->pool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->pool, pass == 0 ? 0 : S->pool_bytes);
      S->d_params = a.take<DecParams>(1);
      S->d_counter = a.take<unsigned int>(64);
      S->d_flags = a.take<int>(2048);
      S->d_dbg = a.take<long long>(2048 * PK_STAMPS);
      hp.sched = a.take<PSched>(sched.size());
      for (int i = 0; i < DD_COUNT; ++i) {
        hp.d[i].nslots = a.take<unsigned char>(nslots[i].size());
        hp.d[i].out = a.take<float>((size_t)PK_MAX_SLOTS * R * hp.d[i].n_rows);
      }
      hp.xt_hi = a.take<float>((long)R * Ep); hp.xt_lo = a.take<float>((long)R * Ep);
      hp.hh_hi = a.take<float>((long)R * 2 * H); hp.hh_lo = a.take<float>((long)R * 2 * H);
      hp.gp_hi = a.take<float>((long)R * H); hp.gp_lo = a.take<float>((long)R * H);
      hp.af_hi = a.take<float>((long)R * H); hp.af_lo = a.take<float>((long)R * H);
      hp.hx = a.take<float>((long)R * 2 * H);
      hp.c1 = a.take<float>((long)R * H); hp.c2 = a.take<float>((long)R * H);
      hp.stats = a.take<float>((long)nvb * R * 4);
      hp.unfinished = a.take<float>(R);
      hp.tok = a.take<int64_t>(R);
      S->tgate = a.take<float>((long)V * H);
      if (pass == 0) {
        S->pool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->pool, S->pool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->pool, 0, S->pool_bytes, st));
      }
    }
    S->R = R; S->K = K;
    S->tgate_epoch = ~0ull;
  }
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<PSched*>(hp.sched), sched.data(), sizeof(PSched) * sched.size(),
                                       cudaMemcpyHostToDevice, st));
  for (int i = 0; i < DD_COUNT; ++i)
    XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<unsigned char*>(hp.d[i].nslots), nslots[i].data(), nslots[i].size(),
                                         cudaMemcpyHostToDevice, st));

  // ---- POS-gate table of every token: tgate = relu(embed . W_gate^T + b)  (sub_modules.py:29-32 applied to
  //      SAModel.py:198's embedding rows); rebuilt whenever the bound parameters change ----
  if (S->tgate_epoch != ctx->param_epoch) {
    GemmP g = gemm_nt(ctx->P[XG_P_EMBED_W], E, ctx->P[XG_P_DGATE_W], E, S->tgate, H, V, H, E);
    g.ep.bias0 = ctx->P[XG_P_DGATE_B];
    g.ep.act = XG_ACT_RELU;
    XG_TRY(gemm_run(ctx, g, st));
    S->tgate_epoch = ctx->param_epoch;
  }

  // ---- tensor maps ----
  MapTable mt;
  CUtensorMap* maps = mt.m;
  const int wpid[8] = {XG_P_H2A_W, XG_P_L1_H2H_W, XG_P_L2_H2H_W, XG_P_L1_I2H_W, XG_P_L1_A2H_W, XG_P_L2_I2H_W, XG_P_L2_A2H_W, XG_P_LOGIT_W};
  for (int i = 0; i < 8; ++i) {
    int rows, cols;
    param_shape(d, wpid[i], &rows, &cols);
    XG_TRY(tc_make_map(ctx, ts, ctx->P[wpid[i]], rows, cols, 128, &maps[i]));
  }
  XG_TRY(tc_make_map(ctx, ts, hp.xt_hi, R, Ep, PK_BN, &maps[8])); XG_TRY(tc_make_map(ctx, ts, hp.xt_lo, R, Ep, PK_BN, &maps[9]));
  XG_TRY(tc_make_map(ctx, ts, hp.hh_hi, R, 2 * H, PK_BN, &maps[10])); XG_TRY(tc_make_map(ctx, ts, hp.hh_lo, R, 2 * H, PK_BN, &maps[11]));
  XG_TRY(tc_make_map(ctx, ts, hp.gp_hi, R, H, PK_BN, &maps[12])); XG_TRY(tc_make_map(ctx, ts, hp.gp_lo, R, H, PK_BN, &maps[13]));
  XG_TRY(tc_make_map(ctx, ts, hp.af_hi, R, H, PK_BN, &maps[14])); XG_TRY(tc_make_map(ctx, ts, hp.af_lo, R, H, PK_BN, &maps[15]));

  hp.B = B; hp.R = R; hp.K = K; hp.H = H; hp.E = E; hp.Ep = Ep; hp.A = A; hp.V = V; hp.T = T;
  hp.hi_inplace = env_flag("XG_PERSIST_HI_INPLACE");
  hp.b_h2a = ctx->P[XG_P_H2A_B]; hp.w_a2w = ctx->P[XG_P_A2W_W]; hp.b_a2w = ctx->P[XG_P_A2W_B];
  hp.b1_i2h = ctx->P[XG_P_L1_I2H_B]; hp.b1_a2h = ctx->P[XG_P_L1_A2H_B]; hp.b1_h2h = ctx->P[XG_P_L1_H2H_B];
  hp.b2_i2h = ctx->P[XG_P_L2_I2H_B]; hp.b2_a2h = ctx->P[XG_P_L2_A2H_B]; hp.b2_h2h = ctx->P[XG_P_L2_H2H_B];
  hp.b_logit = ctx->P[XG_P_LOGIT_B]; hp.embed = ctx->P[XG_P_EMBED_W];
  hp.tgate = S->tgate;
  hp.Vf = Vf; hp.Uv = Uv; hp.pos = pos;
  for (int q = 0; q < 4; ++q) hp.state0[q] = state0[q];
  hp.seq = seq_out; hp.seqlogp = logp_out; hp.flags = S->d_flags;
  hp.sync_counter = S->d_counter;
  hp.dbg_clock = env_flag("XG_PERSIST_TRACE") ? S->d_dbg : nullptr;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_params, &hp, sizeof(DecParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_counter, 0, sizeof(unsigned int) * 64, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_flags, 0, sizeof(int) * (size_t)T, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(seq_out, 0, sizeof(int64_t) * (size_t)B * T, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(logp_out, 0, sizeof(float) * (size_t)B * T, st));
  if (hp.dbg_clock) XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_dbg, 0, sizeof(long long) * 2048 * PK_STAMPS, st));

  if (!S->attr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_persistent_kernel, PK_THREADS, PK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "persistent decoder does not fit on an SM");
    S->attr_set = true;
  }
  {
    ProfScope ps(ctx, "decode_persistent", st);
    const DecParams* dp = S->d_params;
    void* args[2] = {(void*)&dp, (void*)&mt};
    XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)decode_persistent_kernel, dim3(G), dim3(PK_THREADS), args,
                                                     PK_SMEM_BYTES, st));
  }
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(ctx->h_pinned, S->d_flags, sizeof(int) * (size_t)T, cudaMemcpyDeviceToHost, st));
  XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
  int steps = 0;
  while (steps < T && ctx->h_pinned[steps] != 0) ++steps;
  *steps_out = steps;
  if (hp.dbg_clock) {   // XG_PERSIST_TRACE=1: average SM cycles per phase (CTA 0), printed to stderr
    std::vector<long long> h((size_t)T * PK_STAMPS);
    cudaMemcpy(h.data(), S->d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    const char* names[7] = {"G1 (ah,z1h,z2h,z1x,z1g)", "P1 (attention || cell1)", "G3 (z2x,z2a)", "P3 (cell2)", "G4 (logits)",
                            "P4a (logit stats)", "P4b (pick, next inputs)"};
    double tot = 0;
    const int n = steps > 1 ? steps - 1 : 1;
    for (int i = 0; i < 7; ++i) {
      double w = 0, b = 0;
      for (int t = 1; t < std::max(steps, 2); ++t) {
        w += (double)(h[t * PK_STAMPS + 2 * i + 1] - h[t * PK_STAMPS + 2 * i]);
        b += (double)(h[t * PK_STAMPS + 2 * i + 2] - h[t * PK_STAMPS + 2 * i + 1]);
      }
      w /= n; b /= n;
      tot += w + b;
      fprintf(stderr, "[xg persist trace] %-26s own work %7.0f cycles   barrier wait %7.0f cycles\n", names[i], w, b);
    }
    fprintf(stderr, "[xg persist trace] step %.0f cycles\n", tot);
  }
  return XG_OK;
}

// frame recurrence of both encoder streams; eb.G holds the hoisted input projections (+ both biases)
static int persist_encode(xg_context* ctx, const float* fmask, int B, int K, EncBufs& eb, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, G = ctx->sm_count;
  if (!ctx->persist_mode || H % 32 != 0 || B < 1 || H > 4096 || K < 2) return PK_FALLBACK;
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN;
  if (R / PK_BN > 32) return PK_FALLBACK;
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  PersistState*& S = persist_state(ctx);
  if (!S) S = new PersistState();
  EncParams& ep = S->ep;
  const int kbH = H / 32;
  for (int s = 0; s < 2; ++s) {
    GDesc& g = ep.d[s];
    g.w_map = s; g.x_hi = 2; g.x_lo = 3; g.xkb0 = s * kbH; g.n_rows = 4 * H; g.nkb = kbH;
  }
  std::vector<PSched> sched;
  std::vector<std::vector<unsigned char>> nslots(2);
  if (!persist_plan({{0, 1}}, ep.d, R / PK_BN, G, sched, nslots)) return PK_FALLBACK;
  if (S->eB != B) {
    if (S->epool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->epool); S->epool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->epool, pass == 0 ? 0 : S->epool_bytes);
      S->d_eparams = a.take<EncParams>(1);
      S->d_ecounter = a.take<unsigned int>(64);
      ep.sched = a.take<PSched>(sched.size());
      for (int s = 0; s < 2; ++s) {
        ep.d[s].nslots = a.take<unsigned char>(nslots[s].size());
        ep.d[s].out = a.take<float>((size_t)PK_MAX_SLOTS * R * 4 * H);
      }
      ep.hh_hi = a.take<float>((long)R * 2 * H); ep.hh_lo = a.take<float>((long)R * 2 * H);
      if (pass == 0) {
        S->epool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->epool, S->epool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->epool, 0, S->epool_bytes, st));
      }
    }
    S->eB = B;
  }
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<PSched*>(ep.sched), sched.data(), sizeof(PSched) * sched.size(),
                                       cudaMemcpyHostToDevice, st));
  for (int s = 0; s < 2; ++s)
    XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<unsigned char*>(ep.d[s].nslots), nslots[s].data(), nslots[s].size(),
                                         cudaMemcpyHostToDevice, st));
  MapTable mt;
  XG_TRY(tc_make_map(ctx, ts, ctx->P[XG_P_LSTM_RGB_WHH], 4 * H, H, 128, &mt.m[0]));
  XG_TRY(tc_make_map(ctx, ts, ctx->P[XG_P_LSTM_OPFL_WHH], 4 * H, H, 128, &mt.m[1]));
  XG_TRY(tc_make_map(ctx, ts, ep.hh_hi, R, 2 * H, PK_BN, &mt.m[2]));
  XG_TRY(tc_make_map(ctx, ts, ep.hh_lo, R, 2 * H, PK_BN, &mt.m[3]));
  for (int i = 4; i < 16; ++i) mt.m[i] = mt.m[0];
  ep.B = B; ep.R = R; ep.K = K; ep.H = H; ep.hi_inplace = env_flag("XG_PERSIST_HI_INPLACE");
  for (int s = 0; s < 2; ++s) { ep.Gt[s] = eb.G[s]; ep.Hs[s] = eb.Hs[s]; ep.Cs[s] = eb.Cs[s]; }
  ep.fmask = fmask;
  ep.sync_counter = S->d_ecounter;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_eparams, &ep, sizeof(EncParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_ecounter, 0, sizeof(unsigned int) * 64, st));
  if (!S->eattr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(encode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, encode_persistent_kernel, PK_THREADS, PK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "persistent encoder does not fit on an SM");
    S->eattr_set = true;
  }
  ProfScope ps(ctx, "encode_persistent", st);
  const EncParams* dp = S->d_eparams;
  void* args[2] = {(void*)&dp, (void*)&mt};
  XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)encode_persistent_kernel, dim3(G), dim3(PK_THREADS), args, PK_SMEM_BYTES, st));
  return XG_OK;
}

}  // namespace xg
