// Persistent fused greedy decoder: the whole word loop of SAModel.sample (SAModel.py:182-219) in ONE
// cooperative kernel, one CTA per SM.
//
// Per word step the kernel walks eight grid-synchronised phases:
//   G1  tensor-core products that only need the previous state / current token:
//         AH  = [h1|h2] . W_h2a^T        GATE = xt . W_gate^T
//         Z1x = xt . W_i2h1^T            Z1h  = h1 . W_h2h1^T        Z2h = h2 . W_h2h2^T
//   P1  temporal attention (frame features V[b] staged into shared memory by a TMA bulk copy while
//       the scores are computed from Uv; softmax over all K frames; context) and the POS gate
//   G2  Z1g = gp . W_a2h1^T
//   P2  lstm_1 cell  (i,f,o,g; mask carries the state)
//   G3  Z2x = h1' . W_i2h2^T    Z2a = af . W_a2h2^T
//   P3  lstm_2 cell
//   G4  logits = h2' . W_logit^T + b, reduced in the epilogue to per-(tile, warp) max / argmax / sum-exp
//   P4  greedy bookkeeping (argmax, log-prob, unfinished mask, output ids) + embedding of the next token
//
// Every dense product is cut into work items (128 weight rows x one K chunk x 64 caption rows) that are
// dealt round-robin to the 148 CTAs, so each step's 105 MB of split weights streams from L2 through
// ALL SMs' TMA engines instead of through the 16 CTAs a one-tile-per-CTA GEMM would use.  Split-K
// partial sums go to global "slots" and are added in a fixed order by the consuming pointwise phase
// (deterministic; no atomics).  Inside a work item the pipeline is the one of xg_gemm_tc.cuh:
// TMA producer warp, MMA issuer warp (3xTF32, short hi*hi chains in ping-pong TMEM accumulators, cross
// terms in their own accumulator), four epilogue warps promoting into fp32 registers.
#pragma once
#include <cooperative_groups.h>

#include "xg_gemm_tc.cuh"

namespace xg {

struct GDesc {                 // one dense product  out[slot][n][r] = sum_k W[n,k] * X[r, xk0 + k]
  int w_hi, w_lo, x_hi, x_lo;  // indices into the tensor-map table
  int xkb0;                    // first k-block of X used by this product
  int n_rows;                  // weight rows
  int nkb;                     // k-blocks (of 32) of the product
  int kb_per_item;             // split-K granularity
  int slots;                   // ceil(nkb / kb_per_item)
  int mode;                    // 0: store partial tile; 1: logits statistics
  float* out;                  // [slots][n_rows][R]
};

struct MapTable { CUtensorMap m[26]; };

struct PersistParams {
  const CUtensorMap* maps;      // set on the device: points at the __grid_constant__ table
  GDesc ah, gate, z1x, z1h, z2h, z1g, z2x, z2a, logit;
  int B, R, K, H, E, Ep, A, V, T;
  // parameters (fp32, reference layouts)
  const float *b_h2a, *w_a2w, *b_a2w, *b_gate;
  const float *b1_i2h, *b1_a2h, *b1_h2h, *b2_i2h, *b2_a2h, *b2_h2h, *b_logit, *embed;
  // per-batch inputs
  const float *Vf, *Uv, *pos;           // (B,K,H), (B,K,A), (B,H)
  const float* state0[4];               // h1,c1,h2,c2 each (B,H)
  // activation operands of the tensor-core products, K-major, split hi/lo
  float *xt_hi, *xt_lo;                 // [R][Ep]
  float *hh_hi, *hh_lo;                 // [R][2H]   [h1 | h2]
  float *gp_hi, *gp_lo;                 // [R][H]
  float *af_hi, *af_lo;                 // [R][H]
  // exact states, unit-major for coalesced pointwise phases
  float *hx;                            // [2H][R]
  float *c1, *c2;                       // [H][R]
  float *stats;                         // [tiles*4][R][4]  (max, argmax, sum-exp, -)
  float *scores;                        // [R][K] attention scores of the current step
  float *unfinished;                    // [R]
  int64_t *tok;                         // [R]
  // outputs
  int64_t* seq; float* seqlogp; int* flags;   // (B,T), (B,T), (T)
  unsigned int* sync_counter;
  long long* dbg_clock;                 // [T][9] SM-clock stamps of CTA 0 at phase boundaries (diagnostics), or NULL
};

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
    __threadfence();
    atomicAdd(counter, 1u);
    target += (unsigned)G;
    const long long t0 = clock64();
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((int)(v - target) >= 0) break;
      if (clock64() - t0 > 8000000000LL) __trap();
    }
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async_global();
}

struct PipeState {       // running counters shared (by construction) between producer / MMA / epilogue roles
  uint32_t kb_count;     // k-blocks issued so far (stage ring position)
  uint32_t chunk_count;  // accumulation chains issued so far (ping-pong accumulator position)
  uint32_t item_count;   // items processed so far (small-accumulator handshake)
};

constexpr int PS_BN = 64;
constexpr int PS_STAGES = 4;
constexpr int PS_STAGE_BYTES = 2 * 128 * 128 + 2 * PS_BN * 128;   // 48 KB
constexpr int PS_CHUNK = 2;
constexpr int PS_SCRATCH_FLOATS = 2048;                           // attention scratch (A + K floats)
constexpr int PS_SMEM_BYTES = PS_STAGES * PS_STAGE_BYTES + PS_SCRATCH_FLOATS * 4 + 1024 + 512;
constexpr int PS_THREADS = 192;

struct SmemView {
  uint8_t* stages;
  float* scratch;
  uint64_t *full_bar, *empty_bar, *acc_full, *acc_empty, *small_full, *small_empty, *bulk_bar;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ unsigned f2ord(float f);
__device__ __forceinline__ float ord2f(unsigned o);

// run all work items of `nd` products; item i of the phase goes to CTA (i % G)
__device__ __noinline__ void gemm_phase(const PersistParams& P, const GDesc* const* descs, int nd, const SmemView& sv,
                                        uint32_t tmem_base, PipeState& ps, int cta, int G) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncb = P.R / PS_BN;
  int item_base = 0;
  for (int di = 0; di < nd; ++di) {
    const GDesc& d = *descs[di];
    const int rts = (d.n_rows + 127) / 128;
    const int n_items = rts * d.slots * ncb;
    // first item index of this product owned by this CTA
    int first = (cta - item_base % G + G) % G;
    for (int it = first; it < n_items; it += G) {
      const int cb = it % ncb;
      const int kc = (it / ncb) % d.slots;
      const int rt = it / (ncb * d.slots);
      const int kb0 = kc * d.kb_per_item;
      const int nkb = min(d.kb_per_item, d.nkb - kb0);
      const int n_chunks = (nkb + PS_CHUNK - 1) / PS_CHUNK;
      if (warp == 0) {
        if (lane == 0) {
          const CUtensorMap* mwh = P.maps + d.w_hi; const CUtensorMap* mwl = P.maps + d.w_lo;
          const CUtensorMap* mxh = P.maps + d.x_hi; const CUtensorMap* mxl = P.maps + d.x_lo;
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t cnt = ps.kb_count + kb;
            const int s = cnt % PS_STAGES;
            mbar_wait(&sv.empty_bar[s], ((cnt / PS_STAGES) & 1) ^ 1);
            uint8_t* st = sv.stages + s * PS_STAGE_BYTES;
            mbar_expect_tx(&sv.full_bar[s], PS_STAGE_BYTES);
            tma_load_2d(st, mwh, &sv.full_bar[s], (kb0 + kb) * 32, rt * 128);
            tma_load_2d(st + 128 * 128, mwl, &sv.full_bar[s], (kb0 + kb) * 32, rt * 128);
            tma_load_2d(st + 2 * 128 * 128, mxh, &sv.full_bar[s], (d.xkb0 + kb0 + kb) * 32, cb * PS_BN);
            tma_load_2d(st + 2 * 128 * 128 + PS_BN * 128, mxl, &sv.full_bar[s], (d.xkb0 + kb0 + kb) * 32, cb * PS_BN);
          }
        }
      } else if (warp == 1) {
        if (lane == 0) {
          constexpr uint32_t idesc = umma_idesc_tf32(128, PS_BN);
          const uint32_t tmem_small = tmem_base + 2 * PS_BN;
          // the previous item's small accumulator must have been read out
          mbar_wait(sv.small_empty, (ps.item_count & 1) ^ 1);
          tc_fence_after();
          for (int c = 0; c < n_chunks; ++c) {
            const uint32_t cc = ps.chunk_count + c;
            const int b = cc & 1;
            mbar_wait(&sv.acc_empty[b], ((cc >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_main = tmem_base + b * PS_BN;
            for (int kk = 0; kk < PS_CHUNK; ++kk) {
              const int kb = c * PS_CHUNK + kk;
              if (kb >= nkb) break;
              const uint32_t cnt = ps.kb_count + kb;
              const int s = cnt % PS_STAGES;
              mbar_wait(&sv.full_bar[s], (cnt / PS_STAGES) & 1);
              tc_fence_after();
              if (P.dbg_clock && cta == 0 && d.mode == 1 && kb < 16) P.dbg_clock[2048 * 17 + kb] = clock64();
              const uint32_t base = smem_u32(sv.stages + s * PS_STAGE_BYTES);
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t wh = umma_desc_sw128(base + k4 * 32);
                const uint64_t wl = umma_desc_sw128(base + 128 * 128 + k4 * 32);
                const uint64_t xh = umma_desc_sw128(base + 2 * 128 * 128 + k4 * 32);
                const uint64_t xl = umma_desc_sw128(base + 2 * 128 * 128 + PS_BN * 128 + k4 * 32);
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
      } else {
        const int quad = warp & 3;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        float acc[PS_BN];
#pragma unroll
        for (int u = 0; u < PS_BN; ++u) acc[u] = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          const uint32_t cc = ps.chunk_count + c;
          const int b = cc & 1;
          mbar_wait(&sv.acc_full[b], (cc >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int q = 0; q < PS_BN; q += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + (uint32_t)(b * PS_BN + q), r);
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
        for (int q = 0; q < PS_BN; q += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + 2 * PS_BN + lane_base + (uint32_t)q, r);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[q + u] += __uint_as_float(r[u]);
        }
        tc_fence_before();
        mbar_arrive(sv.small_empty);
        const int n = rt * 128 + quad * 32 + lane;        // weight row held by this thread
        if (d.mode == 0) {
          if (n < d.n_rows) {
            float4* o = reinterpret_cast<float4*>(d.out + ((long)kc * d.n_rows + n) * P.R + cb * PS_BN);
#pragma unroll
            for (int u = 0; u < PS_BN; u += 4) o[u / 4] = make_float4(acc[u], acc[u + 1], acc[u + 2], acc[u + 3]);
          }
        } else {
          // logits: per caption column, reduce (max, lowest argmax) over the 32 vocabulary rows of this
          // warp, then the sum of exp(x - max); one record per (row tile, warp, caption)
          const float bias = n < d.n_rows ? P.b_logit[n] : 0.f;
          float* srec = P.stats + ((long)(rt * 4 + quad) * P.R + cb * PS_BN) * 4;
#pragma unroll
          for (int u = 0; u < PS_BN; ++u) {
            const float v = n < d.n_rows ? acc[u] + bias : -INFINITY;
            const unsigned ov = f2ord(v);
            const unsigned mo = __reduce_max_sync(0xffffffffu, ov);
            const float best = ord2f(mo);
            const int bi = (int)__reduce_min_sync(0xffffffffu, ov == mo ? (unsigned)n : 0x7fffffffu);
            float e = (v == -INFINITY) ? 0.f : expf(v - best);
            e = warp_sum(e);
            if (lane == 0) {
              srec[u * 4 + 0] = best;
              srec[u * 4 + 1] = __int_as_float(bi);
              srec[u * 4 + 2] = e;
            }
          }
        }
      }
      ps.kb_count += nkb;
      ps.chunk_count += n_chunks;
      ps.item_count += 1;
    }
    item_base += n_items;
  }
}

__device__ __forceinline__ void store_split(float* hi, float* lo, long idx, float v) {
  const float h = tf32_rna(v);
  hi[idx] = h;
  lo[idx] = tf32_rna(v - h);
}

// Sum of split-K partial slots.  All loads are issued before the first add: with 6 warps per SM and an
// L1 that the grid barrier's fence leaves cold, a load->add->load chain costs one L2 round trip per slot
// (measured: 50k cycles per cell phase); batched, the whole element costs one.
template <int MAXS>
__device__ __forceinline__ float sum_slots(const float* base, int slots, long slot_stride, long idx) {
  float v[MAXS];
#pragma unroll
  for (int k = 0; k < MAXS; ++k) v[k] = k < slots ? __ldcg(base + (long)k * slot_stride + idx) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXS; ++k) s += v[k];   // fixed order (unused slots add +0)
  return s;
}
__device__ __forceinline__ unsigned f2ord(float f) {      // order-preserving float -> uint (for redux.sync)
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// lstm cell (decoder order i,f,o,g) for all (unit j, row r); z = sum of partial slots + three biases
__device__ __forceinline__ void cell_phase(const PersistParams& P, int layer, const float* mask_rows, bool use_mask,
                                           int gtid, int gthreads) {
  const int H = P.H, R = P.R;
  const GDesc& da = layer == 0 ? P.z1x : P.z2x;
  const GDesc& db = layer == 0 ? P.z1g : P.z2a;
  const GDesc& dc = layer == 0 ? P.z1h : P.z2h;
  const float* bi = layer == 0 ? P.b1_i2h : P.b2_i2h;
  const float* ba = layer == 0 ? P.b1_a2h : P.b2_a2h;
  const float* bh = layer == 0 ? P.b1_h2h : P.b2_h2h;
  float* cst = layer == 0 ? P.c1 : P.c2;
  const long sstr = (long)4 * H * R;
  for (int e = gtid; e < H * R; e += gthreads) {
    const int j = e / R, r = e % R;
    float z[4], za[4], zb[4], zc[4], zbias[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {            // issue every load of the element first
      const long idx = (long)(g * H + j) * R + r;
      za[g] = sum_slots<4>(da.out, da.slots, sstr, idx);
      zb[g] = sum_slots<4>(db.out, db.slots, sstr, idx);
      zc[g] = sum_slots<4>(dc.out, dc.slots, sstr, idx);
      zbias[g] = __ldg(bi + g * H + j) + __ldg(ba + g * H + j) + __ldg(bh + g * H + j);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) z[g] = za[g] + zb[g] + zc[g] + zbias[g];
    const float ig = sigmoid_f(z[0]), fg = sigmoid_f(z[1]), og = sigmoid_f(z[2]), gg = tanhf(z[3]);
    const float m = use_mask ? mask_rows[r] : 1.f;
    const float cp = cst[e];
    const float hp = P.hx[(long)(layer * H + j) * R + r];
    float c = fg * cp + ig * gg;
    c = c * m + cp * (1.f - m);
    float h = og * tanhf(c);
    h = h * m + hp * (1.f - m);
    cst[e] = c;
    P.hx[(long)(layer * H + j) * R + r] = h;
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + layer * H + j, h);
  }
}

__global__ void __launch_bounds__(PS_THREADS, 1)
decode_persistent_kernel(const PersistParams* __restrict__ Pp, const __grid_constant__ MapTable maps) {
  // parameter block -> shared memory (the inline-asm memory clobbers force re-reads; keep them on-chip)
  __shared__ PersistParams Psm;
  for (int i = threadIdx.x; i < (int)(sizeof(PersistParams) / 4); i += PS_THREADS)
    reinterpret_cast<uint32_t*>(&Psm)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  __syncthreads();
  if (threadIdx.x == 0) Psm.maps = maps.m;
  __syncthreads();
  const PersistParams& P = Psm;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SmemView sv;
  sv.stages = smem;
  sv.scratch = reinterpret_cast<float*>(smem + PS_STAGES * PS_STAGE_BYTES);
  sv.full_bar = reinterpret_cast<uint64_t*>(smem + PS_STAGES * PS_STAGE_BYTES + PS_SCRATCH_FLOATS * 4);
  sv.empty_bar = sv.full_bar + PS_STAGES;
  sv.acc_full = sv.empty_bar + PS_STAGES;
  sv.acc_empty = sv.acc_full + 2;
  sv.small_full = sv.acc_empty + 2;
  sv.small_empty = sv.small_full + 1;
  sv.bulk_bar = sv.small_empty + 1;
  sv.tmem_slot = reinterpret_cast<uint32_t*>(sv.bulk_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;
  const int gtid = cta * PS_THREADS + threadIdx.x, gthreads = G * PS_THREADS;
  const int H = P.H, R = P.R, B = P.B, K = P.K, A = P.A, E = P.E, Ep = P.Ep, T = P.T;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PS_STAGES; ++s) { mbar_init(&sv.full_bar[s], 1); mbar_init(&sv.empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sv.acc_full[b], 1); mbar_init(&sv.acc_empty[b], 128); }
    mbar_init(sv.small_full, 1);
    mbar_init(sv.small_empty, 128);
    mbar_init(sv.bulk_bar, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<256>(sv.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *sv.tmem_slot;
  PipeState ps{0, 0, 0};
  unsigned int sync_target = 0;
  uint32_t bulk_phase = 0;

  // ---- prologue: states, <bos> embedding, bookkeeping ----
  for (int e = gtid; e < H * R; e += gthreads) {
    const int j = e / R, r = e % R;
    const float h1 = r < B ? P.state0[0][(long)r * H + j] : 0.f;
    const float c1 = r < B ? P.state0[1][(long)r * H + j] : 0.f;
    const float h2 = r < B ? P.state0[2][(long)r * H + j] : 0.f;
    const float c2 = r < B ? P.state0[3][(long)r * H + j] : 0.f;
    P.hx[(long)j * R + r] = h1; P.hx[(long)(H + j) * R + r] = h2;
    P.c1[e] = c1; P.c2[e] = c2;
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + j, h1);
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + H + j, h2);
  }
  for (int e = gtid; e < R * Ep; e += gthreads) {
    const int r = e / Ep, k = e % Ep;
    store_split(P.xt_hi, P.xt_lo, e, (r < B && k < E) ? P.embed[k] : 0.f);   // token 0 = <bos> (SAModel.py:184)
  }
  for (int r = gtid; r < R; r += gthreads) { P.unfinished[r] = 1.f; P.tok[r] = 0; }
  grid_barrier(P.sync_counter, sync_target, G);

#define PS_STAMP(i) do { if (P.dbg_clock && cta == 0 && threadIdx.x == 0) P.dbg_clock[t * 17 + (i)] = clock64(); } while (0)
  for (int t = 0; t < T; ++t) {
    PS_STAMP(0);
    // ================= G1 =================
    {
      const GDesc* ds[5] = {&P.ah, &P.z1h, &P.z2h, &P.gate, &P.z1x};
      gemm_phase(P, ds, 5, sv, tmem_base, ps, cta, G);
    }
    PS_STAMP(1);
    grid_barrier(P.sync_counter, sync_target, G);
    PS_STAMP(2);
    // ================= P1: attention scores (all CTAs: caption x frame-group items) + POS gate =================
    {
      // s[r][k] = a2w . tanh(AH[r] + Uv[r,k,:]) + b      (sub_modules.py:677-678); softmax/context in P2
      const int ng = max(1, min(K, G / max(B, 1)));
      const int fpg = (K + ng - 1) / ng;            // frames per group
      const int nge = (K + fpg - 1) / fpg;
      const long sstr = (long)A * R;
      float* ah = sv.scratch;                       // A floats
      for (int it = cta; it < B * nge; it += G) {
        const int r = it / nge, k0 = (it % nge) * fpg, k1 = min(K, k0 + fpg);
        for (int a = threadIdx.x; a < A; a += PS_THREADS)
          ah[a] = sum_slots<4>(P.ah.out, P.ah.slots, sstr, (long)a * R + r) + __ldg(P.b_h2a + a);
        __syncthreads();
        for (int k = k0 + warp; k < k1; k += PS_THREADS / 32) {
          const float* u = P.Uv + ((long)r * K + k) * A;
          float p = 0.f;
#pragma unroll 4
          for (int a = lane; a < A; a += 32) p += __ldg(P.w_a2w + a) * tanhf(ah[a] + __ldg(u + a));
          p = warp_sum(p);
          if (lane == 0) P.scores[(long)r * K + k] = p + __ldg(P.b_a2w);
        }
        __syncthreads();
      }
      const long gstr = (long)H * R;
      for (int e = gtid; e < H * R; e += gthreads) {
        const int j = e / R, r = e % R;
        float g = sum_slots<4>(P.gate.out, P.gate.slots, gstr, (long)j * R + r) + __ldg(P.b_gate + j);
        g = g > 0.f ? g : 0.f;
        const float pv = r < B ? __ldg(P.pos + (long)r * H + j) : 0.f;
        store_split(P.gp_hi, P.gp_lo, (long)r * H + j, pv * (1.f + g));
      }
    }
    PS_STAMP(3);
    grid_barrier(P.sync_counter, sync_target, G);
    PS_STAMP(4);
    // ================= G2 =================
    {
      const GDesc* ds[1] = {&P.z1g};
      gemm_phase(P, ds, 1, sv, tmem_base, ps, cta, G);
    }
    PS_STAMP(5);
    grid_barrier(P.sync_counter, sync_target, G);
    PS_STAMP(6);
    // ================= P2: softmax over ALL K frames + context (V[r] staged by a TMA bulk copy), lstm_1 =================
    for (int r = cta; r < B; r += G) {
      float* sc = sv.scratch;                              // K floats
      float* vsm = reinterpret_cast<float*>(sv.stages);    // K*H floats: the frame-feature matrix of caption r
      if (threadIdx.x == 0) {
        mbar_expect_tx(sv.bulk_bar, (uint32_t)(K * H * 4));
        bulk_g2s(vsm, P.Vf + (long)r * K * H, (uint32_t)(K * H * 4), sv.bulk_bar);
      }
      if (warp == 1) {
        float mx = -INFINITY;
        for (int k = lane; k < K; k += 32) { const float v = __ldcg(P.scores + (long)r * K + k); sc[k] = v; mx = fmaxf(mx, v); }
        mx = warp_max(mx);
        __syncwarp();
        float sum = 0.f;
        for (int k = lane; k < K; k += 32) { const float e = expf(sc[k] - mx); sc[k] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int k = lane; k < K; k += 32) sc[k] *= inv;
      }
      mbar_wait(sv.bulk_bar, bulk_phase & 1);
      __syncthreads();
      for (int j = threadIdx.x; j < H; j += PS_THREADS) {
        float a = 0.f;
        for (int k = 0; k < K; ++k) a += sc[k] * vsm[k * H + j];
        store_split(P.af_hi, P.af_lo, (long)r * H + j, a);
      }
      bulk_phase++;
      __syncthreads();
    }
    fence_proxy_async_smem();      // vsm was read through the generic proxy: order before the next TMA writes
    cell_phase(P, 0, P.unfinished, t > 0, gtid, gthreads);
    PS_STAMP(7);
    grid_barrier(P.sync_counter, sync_target, G);
    PS_STAMP(8);
    // ================= G3 =================
    {
      const GDesc* ds[2] = {&P.z2x, &P.z2a};
      gemm_phase(P, ds, 2, sv, tmem_base, ps, cta, G);
    }
    PS_STAMP(9);
    grid_barrier(P.sync_counter, sync_target, G);
    PS_STAMP(10);
    // ================= P3: lstm_2 =================
    cell_phase(P, 1, P.unfinished, t > 0, gtid, gthreads);
    PS_STAMP(11);
    grid_barrier(P.sync_counter, sync_target, G);
    PS_STAMP(12);
    // ================= G4: logits statistics =================
    {
      const GDesc* ds[1] = {&P.logit};
      gemm_phase(P, ds, 1, sv, tmem_base, ps, cta, G);
    }
    PS_STAMP(13);
    grid_barrier(P.sync_counter, sync_target, G);
    PS_STAMP(14);
    // ================= P4: greedy bookkeeping + next embedding (SAModel.py:185-210) =================
    {
      const int nrec = ((P.V + 127) / 128) * 4;
      for (int r = cta; r < B; r += G) {
        if (warp == 0) {
          float best = -INFINITY; int bi = 0x7fffffff;
          for (int q = lane; q < nrec; q += 32) {
            const float* rec = P.stats + ((long)q * R + r) * 4;
            const float v = rec[0]; const int vi = __float_as_int(rec[1]);
            if (v > best || (v == best && vi < bi)) { best = v; bi = vi; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
          }
          float s = 0.f;
          for (int q = lane; q < nrec; q += 32) {
            const float* rec = P.stats + ((long)q * R + r) * 4;
            if (rec[0] != -INFINITY) s += rec[2] * expf(rec[0] - best);
          }
          s = warp_sum(s);
          if (lane == 0) {
            float unf = (t == 0) ? 1.f : P.unfinished[r];
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
        const float* src = P.embed + (long)tokv * E;
        for (int k = threadIdx.x; k < Ep; k += PS_THREADS)
          store_split(P.xt_hi, P.xt_lo, (long)r * Ep + k, k < E ? src[k] : 0.f);
        __syncthreads();
      }
    }
    PS_STAMP(15);
    grid_barrier(P.sync_counter, sync_target, G);
    PS_STAMP(16);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct PersistState {
  int R = 0, K = 0;
  char* pool = nullptr;          // one device allocation carved below
  size_t pool_bytes = 0;
  CUtensorMap* d_maps = nullptr;
  PersistParams* d_params = nullptr;
  unsigned int* d_counter = nullptr;
  int* d_flags = nullptr;
  long long* d_dbg = nullptr;
  PersistParams hp;              // host copy (pointers into the pool)
  bool attr_set = false;
};

inline PersistState*& persist_state(xg_context* ctx) {
  static std::unordered_map<xg_context*, PersistState*> m;
  return m[ctx];
}
static void persist_release(xg_context* ctx) {
  PersistState* s = persist_state(ctx);
  if (!s) return;
  if (s->pool) cudaFree(s->pool);
  delete s;
  persist_state(ctx) = nullptr;
}

static bool persist_eligible(const xg_context* ctx, int B, int K) {
  const xg_dims& d = ctx->d;
  return ctx->persist_mode && d.rnn % 32 == 0 && d.rnn <= 512 && d.embed <= 1024 && B <= 64 && d.att + K <= PS_SCRATCH_FLOATS &&
         (long)K * d.rnn * 4 <= (long)PS_STAGES * PS_STAGE_BYTES && ((long)K * d.rnn * 4) % 16 == 0 && d.vocab >= 2;
}

static int persist_greedy(xg_context* ctx, const float* Vf, const float* Uv, const float* pos, const float* const* state0,
                          int B, int K, int T, int64_t* seq_out, float* logp_out, int* steps_out, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, A = d.att, V = d.vocab;
  const int R = 64, Ep = (E + 31) / 32 * 32;
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  PersistState*& S = persist_state(ctx);
  if (!S) S = new PersistState();
  PersistParams& hp = S->hp;

  // ---- weights: cached tf32 hi/lo splits (K-major, K padded to 32) ----
  struct WSpec { int pid; int rows; int cols; };
  const WSpec wspec[9] = {{XG_P_H2A_W, A, 2 * H},  {XG_P_DGATE_W, H, E},       {XG_P_L1_I2H_W, 4 * H, E},
                          {XG_P_L1_A2H_W, 4 * H, H}, {XG_P_L1_H2H_W, 4 * H, H}, {XG_P_L2_I2H_W, 4 * H, H},
                          {XG_P_L2_A2H_W, 4 * H, H}, {XG_P_L2_H2H_W, 4 * H, H}, {XG_P_LOGIT_W, V, H}};
  const float* whi[9]; const float* wlo[9]; int wkp[9];
  for (int i = 0; i < 9; ++i)
    XG_TRY(tc_operand(ctx, ts, 0, ctx->P[wspec[i].pid], wspec[i].cols, 1, wspec[i].rows, wspec[i].cols, &whi[i], &wlo[i],
                      &wkp[i], st));

  // ---- device pool (activations, partial slots, maps, params) ----
  const int kbH = H / 32, kbE = Ep / 32;
  auto slots_of = [](int nkb, int per) { return (nkb + per - 1) / per; };
  const int per_g1 = 8, per_g2 = 4, per_g3 = 4;
  if (S->R != R || S->K != K) {
    if (S->pool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->pool); S->pool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->pool, pass == 0 ? 0 : S->pool_bytes);
      S->d_maps = a.take<CUtensorMap>(32);
      S->d_params = a.take<PersistParams>(1);
      S->d_counter = a.take<unsigned int>(64);
      S->d_flags = a.take<int>(2048);
      S->d_dbg = a.take<long long>(2048 * 17 + 64);
      hp.xt_hi = a.take<float>((long)R * Ep); hp.xt_lo = a.take<float>((long)R * Ep);
      hp.hh_hi = a.take<float>((long)R * 2 * H); hp.hh_lo = a.take<float>((long)R * 2 * H);
      hp.gp_hi = a.take<float>((long)R * H); hp.gp_lo = a.take<float>((long)R * H);
      hp.af_hi = a.take<float>((long)R * H); hp.af_lo = a.take<float>((long)R * H);
      hp.hx = a.take<float>((long)2 * H * R);
      hp.c1 = a.take<float>((long)H * R); hp.c2 = a.take<float>((long)H * R);
      hp.stats = a.take<float>((long)((V + 127) / 128) * 4 * R * 4);
      hp.unfinished = a.take<float>(R);
      hp.scores = a.take<float>((long)R * K);
      hp.tok = a.take<int64_t>(R);
      hp.ah.out = a.take<float>((long)slots_of(2 * kbH, per_g1) * A * R);
      hp.gate.out = a.take<float>((long)slots_of(kbE, per_g1) * H * R);
      hp.z1x.out = a.take<float>((long)slots_of(kbE, per_g1) * 4 * H * R);
      hp.z1h.out = a.take<float>((long)slots_of(kbH, per_g1) * 4 * H * R);
      hp.z2h.out = a.take<float>((long)slots_of(kbH, per_g1) * 4 * H * R);
      hp.z1g.out = a.take<float>((long)slots_of(kbH, per_g2) * 4 * H * R);
      hp.z2x.out = a.take<float>((long)slots_of(kbH, per_g3) * 4 * H * R);
      hp.z2a.out = a.take<float>((long)slots_of(kbH, per_g3) * 4 * H * R);
      hp.logit.out = nullptr;
      if (pass == 0) {
        S->pool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->pool, S->pool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->pool, 0, S->pool_bytes, st));
      }
    }
    S->R = R; S->K = K;
  }
  if (T > 2048) { ctx->es.msg = "persist_greedy: seq_length too large"; return XG_ERR_BAD_SHAPE; }

  // ---- tensor maps: 9 weights x (hi,lo) = 0..17 ; activations xt 18,19  hh 20,21  gp 22,23  af 24,25 ----
  MapTable mt;
  CUtensorMap* maps = mt.m;
  for (int i = 0; i < 9; ++i) {
    XG_TRY(tc_make_map(ctx, ts, whi[i], wspec[i].rows, wkp[i], 128, &maps[2 * i]));
    XG_TRY(tc_make_map(ctx, ts, wlo[i], wspec[i].rows, wkp[i], 128, &maps[2 * i + 1]));
  }
  XG_TRY(tc_make_map(ctx, ts, hp.xt_hi, R, Ep, PS_BN, &maps[18])); XG_TRY(tc_make_map(ctx, ts, hp.xt_lo, R, Ep, PS_BN, &maps[19]));
  XG_TRY(tc_make_map(ctx, ts, hp.hh_hi, R, 2 * H, PS_BN, &maps[20])); XG_TRY(tc_make_map(ctx, ts, hp.hh_lo, R, 2 * H, PS_BN, &maps[21]));
  XG_TRY(tc_make_map(ctx, ts, hp.gp_hi, R, H, PS_BN, &maps[22])); XG_TRY(tc_make_map(ctx, ts, hp.gp_lo, R, H, PS_BN, &maps[23]));
  XG_TRY(tc_make_map(ctx, ts, hp.af_hi, R, H, PS_BN, &maps[24])); XG_TRY(tc_make_map(ctx, ts, hp.af_lo, R, H, PS_BN, &maps[25]));

  auto mk = [&](GDesc& g, int widx, int xmap, int xkb0, int n_rows, int nkb, int per, int mode) {
    g.w_hi = 2 * widx; g.w_lo = 2 * widx + 1; g.x_hi = xmap; g.x_lo = xmap + 1;
    g.xkb0 = xkb0; g.n_rows = n_rows; g.nkb = nkb; g.kb_per_item = per; g.slots = slots_of(nkb, per); g.mode = mode;
  };
  mk(hp.ah, 0, 20, 0, A, 2 * kbH, per_g1, 0);
  mk(hp.gate, 1, 18, 0, H, kbE, per_g1, 0);
  mk(hp.z1x, 2, 18, 0, 4 * H, kbE, per_g1, 0);
  mk(hp.z1g, 3, 22, 0, 4 * H, kbH, per_g2, 0);
  mk(hp.z1h, 4, 20, 0, 4 * H, kbH, per_g1, 0);
  mk(hp.z2x, 5, 20, 0, 4 * H, kbH, per_g3, 0);
  mk(hp.z2a, 6, 24, 0, 4 * H, kbH, per_g3, 0);
  mk(hp.z2h, 7, 20, kbH, 4 * H, kbH, per_g1, 0);
  mk(hp.logit, 8, 20, kbH, V, kbH, kbH, 1);
  for (const GDesc* g : {&hp.ah, &hp.gate, &hp.z1x, &hp.z1g, &hp.z1h, &hp.z2x, &hp.z2a, &hp.z2h})
    XG_REQUIRE(ctx->es, g->slots <= 4, XG_ERR_UNSUPPORTED, "persistent decoder: more than 4 split-K slots");
  hp.maps = S->d_maps;
  hp.B = B; hp.R = R; hp.K = K; hp.H = H; hp.E = E; hp.Ep = Ep; hp.A = A; hp.V = V; hp.T = T;
  hp.b_h2a = ctx->P[XG_P_H2A_B]; hp.w_a2w = ctx->P[XG_P_A2W_W]; hp.b_a2w = ctx->P[XG_P_A2W_B]; hp.b_gate = ctx->P[XG_P_DGATE_B];
  hp.b1_i2h = ctx->P[XG_P_L1_I2H_B]; hp.b1_a2h = ctx->P[XG_P_L1_A2H_B]; hp.b1_h2h = ctx->P[XG_P_L1_H2H_B];
  hp.b2_i2h = ctx->P[XG_P_L2_I2H_B]; hp.b2_a2h = ctx->P[XG_P_L2_A2H_B]; hp.b2_h2h = ctx->P[XG_P_L2_H2H_B];
  hp.b_logit = ctx->P[XG_P_LOGIT_B]; hp.embed = ctx->P[XG_P_EMBED_W];
  hp.Vf = Vf; hp.Uv = Uv; hp.pos = pos;
  for (int q = 0; q < 4; ++q) hp.state0[q] = state0[q];
  hp.seq = seq_out; hp.seqlogp = logp_out; hp.flags = S->d_flags;
  hp.sync_counter = S->d_counter;
  hp.dbg_clock = getenv("XG_PERSIST_TRACE") ? S->d_dbg : nullptr;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_params, &hp, sizeof(PersistParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_counter, 0, sizeof(unsigned int) * 64, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_flags, 0, sizeof(int) * (size_t)T, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(seq_out, 0, sizeof(int64_t) * (size_t)B * T, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(logp_out, 0, sizeof(float) * (size_t)B * T, st));

  if (!S->attr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_persistent_kernel, PS_THREADS, PS_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "persistent decoder does not fit on an SM");
    S->attr_set = true;
  }
  {
    ProfScope ps(ctx, "decode_persistent", st);
    const PersistParams* dp = S->d_params;
    void* args[2] = {(void*)&dp, (void*)&mt};
    XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)decode_persistent_kernel, dim3(ctx->sm_count), dim3(PS_THREADS), args,
                                                     PS_SMEM_BYTES, st));
  }
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(ctx->h_pinned, S->d_flags, sizeof(int) * (size_t)T, cudaMemcpyDeviceToHost, st));
  XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
  int steps = 0;
  while (steps < T && ctx->h_pinned[steps] != 0) ++steps;
  *steps_out = steps;
  if (hp.dbg_clock) {   // XG_PERSIST_TRACE=1: average SM cycles per phase (CTA 0), printed to stderr
    std::vector<long long> h((size_t)T * 17);
    cudaMemcpy(h.data(), S->d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    const char* names[8] = {"G1", "P1", "G2", "P2", "G3", "P3", "G4", "P4"};
    double tot = 0;
    for (int i = 0; i < 8; ++i) {
      double w = 0, b = 0;
      for (int t = 1; t < T; ++t) {
        w += (double)(h[t * 17 + 2 * i + 1] - h[t * 17 + 2 * i]);
        b += (double)(h[t * 17 + 2 * i + 2] - h[t * 17 + 2 * i + 1]);
      }
      w /= (T > 1 ? T - 1 : 1); b /= (T > 1 ? T - 1 : 1);
      tot += w + b;
      fprintf(stderr, "[xg persist trace] %s  own work %7.0f cycles   barrier wait %7.0f cycles\n", names[i], w, b);
    }
    fprintf(stderr, "[xg persist trace] step %.0f cycles\n", tot);
    std::vector<long long> kbt(64);
    cudaMemcpy(kbt.data(), S->d_dbg + 2048 * 17, sizeof(long long) * 64, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[xg persist trace] logits item of CTA 0, cycles between consecutive 'stage full' events:");
    for (int i = 1; i < 16; ++i) fprintf(stderr, " %lld", kbt[i] - kbt[i - 1]);
    fprintf(stderr, "\n");
  }
  return XG_OK;
}

}  // namespace xg
