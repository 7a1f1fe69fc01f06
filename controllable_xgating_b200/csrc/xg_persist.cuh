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
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstring>
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
constexpr int PK_MAX_ITEMS = 20;
constexpr int PK_MAX_SLOTS = 6;
constexpr int PK_MAX_DESCS = 8;
constexpr int PK_BULK_CHUNKS = 8;     // Uv chunks (+1 barrier for V)
constexpr int PK_STAMPS = 16;

struct PItem { short desc, slot, rt, cb, kb0, nkb; };
struct PSched { short n, tot_kb, tot_chunks, pad; PItem it[PK_MAX_ITEMS]; };   // 80 bytes

struct GDesc {                   // out[slot][r][n] = sum_k W[n,k] * X[r, 32*xkb0 + k]
  int w_map, x_hi, x_lo;         // indices into the tensor-map table
  int xkb0, n_rows, nkb;
  float* out;                    // [slots][R][n_rows]; slots a strip never writes stay zero (pool is zero-filled
  int ns;                        //   once and the schedule is static), so every consumer adds `ns` slots
};

struct MapTable { CUtensorMap m[18]; };   // 0-15 GEMM operands, 16: V as [B*K][H] (box H/2 x K, no swizzle)

// CODE SIZE IS THE FIRST-ORDER COST HERE: every phase runs once per word step, so its instructions are
// fetched cold unless the whole step fits the 32 KB L1.5 instruction cache (measured: the first version,
// 12.3k instructions, ran at ~13 cycles per instruction in every phase).  Hence: rolled loops, shared
// __noinline__ helpers, ex2/rcp-based activations, descriptors advanced by adds.

__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// mbarrier wait: one try_wait inline (it suspends in hardware), the bounded spin lives out of line
__device__ __noinline__ void pk_wait_spin(uint32_t addr, uint32_t parity) {
  const long long t0 = clock64();
  while (true) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void pk_wait(uint32_t addr, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(addr), "r"(parity) : "memory");
  if (!done) pk_wait_spin(addr, parity);
}
__device__ __forceinline__ void pk_expect_tx_noarrive(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pk_arrive(uint32_t addr) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void pk_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pk_tma_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void pk_tma_prefetch_l2(const CUtensorMap* tm, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void pk_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// grid-wide barrier on a monotonically increasing counter (cooperative launch guarantees residency).
// (A variant with one release flag per CTA written by the last arriver measured slower: 2.2 vs 1.8 us.)
__device__ __noinline__ void grid_barrier(unsigned int* counter, unsigned int& target, int G) {
#ifdef PK_CG_BARRIER
  fence_proxy_async_global();
  cooperative_groups::this_grid().sync();
  fence_proxy_async_global();
  return;
#endif
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
  uint32_t npre;         // leading k-blocks of the NEXT GEMM phase whose weight tiles are already in flight
  uint32_t grp_b;        // accumulation chains drained so far by the second epilogue warpgroup (xg_grouped.cuh, GSched::dual)
};

// shared-memory map (32-bit shared-window addresses; the barriers are 8 bytes apart)
struct SmemView {
  uint8_t* stages;
  float* scratch;
  uint32_t stages_u32;
  uint32_t full_bar, empty_bar, split_bar, acc_full, acc_empty, small_full, small_empty, bulk_bar, acc_full2;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ SmemView carve_smem(uint8_t* smem_raw) {
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  SmemView sv;
  sv.stages = smem;
  sv.stages_u32 = smem_u32(smem);
  sv.scratch = reinterpret_cast<float*>(smem + PK_STAGES * PK_STAGE_BYTES);
  const uint32_t bars = smem_u32(smem + PK_STAGES * PK_STAGE_BYTES + PK_SCRATCH_FLOATS * 4);
  sv.full_bar = bars;
  sv.empty_bar = sv.full_bar + 8 * PK_STAGES;
  sv.split_bar = sv.empty_bar + 8 * PK_STAGES;
  sv.acc_full = sv.split_bar + 8 * PK_STAGES;
  sv.acc_empty = sv.acc_full + 16;
  sv.small_full = sv.acc_empty + 16;
  sv.small_empty = sv.small_full + 8;
  sv.bulk_bar = sv.small_empty + 8;
  sv.tmem_slot = reinterpret_cast<uint32_t*>(smem + PK_STAGES * PK_STAGE_BYTES + PK_SCRATCH_FLOATS * 4 + 8 * (3 * PK_STAGES + 6 + PK_BULK_CHUNKS + 1));
  sv.acc_full2 = bars + 8 * (3 * PK_STAGES + 6 + PK_BULK_CHUNKS + 2);      // (behind the TMEM slot; the barrier area is 512 bytes)
  return sv;
}

__device__ __forceinline__ void pk_mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count));
}

__device__ __noinline__ uint32_t pipeline_setup(const SmemView& sv) {
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < PK_STAGES; ++s) {
      pk_mbar_init(sv.full_bar + 8 * s, 1); pk_mbar_init(sv.empty_bar + 8 * s, 1); pk_mbar_init(sv.split_bar + 8 * s, 128);
    }
    for (int b = 0; b < 2; ++b) { pk_mbar_init(sv.acc_full + 8 * b, 1); pk_mbar_init(sv.acc_empty + 8 * b, 128); pk_mbar_init(sv.acc_full2 + 8 * b, 1); }
    pk_mbar_init(sv.small_full, 1);
    pk_mbar_init(sv.small_empty, 128);
    for (int c = 0; c <= PK_BULK_CHUNKS; ++c) pk_mbar_init(sv.bulk_bar + 8 * c, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<256>(sv.tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sv.tmem_slot;
}

__device__ __noinline__ void pipeline_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 2) tmem_dealloc<256>(tmem_base);
}

#ifdef PK_FINE_TRACE
#define PK_FINE(i) do { if (g_fine) g_fine[(i)] = clock64(); } while (0)
#else
#define PK_FINE(i) do { } while (0)
#endif

// lo part of the in-kernel weight split: residual of the truncation the tensor core applies to the raw word,
// rounded to tf32 (nearest, ties away) with integer ops — cvt.rna.tf32 runs on the 16-lane conversion pipe and
// made the split warps the slowest stage of the pipeline (1100 cycles per 16 KB tile)
__device__ __forceinline__ float tf32_lo(float w) {
  const float r = w - __uint_as_float(__float_as_uint(w) & 0xffffe000u);
  return __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
}

// all work items of this CTA for one GEMM phase (sc lives in shared memory)
__device__ __noinline__ void gemm_phase(const GDesc* descs, const PSched* sc, const PSched* sc_next, const CUtensorMap* maps, int R,
                                        const SmemView& sv, uint32_t tmem_base, PipeState& ps, long long* g_fine = nullptr) {
  // warp-uniform role dispatch: the broadcast makes the warp index (and every counter derived below)
  // provably uniform, so descriptors / coordinates live in uniform registers and each UTCHMMA / UTMALDG
  // issues directly instead of through a per-lane waterfall (ELECT + R2UR.BROADCAST loop, ~90 cycles each)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_items = __shfl_sync(0xffffffffu, (int)sc->n, 0);
  if (threadIdx.x == 0) PK_FINE(0);
  if (warp == 0) {            // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
    uint32_t cnt = __shfl_sync(0xffffffffu, ps.kb_count, 0);
    int npre = (int)__shfl_sync(0xffffffffu, ps.npre, 0);      // weight tiles of the first npre k-blocks were prefetched
    const uint32_t stages_u32 = __shfl_sync(0xffffffffu, sv.stages_u32, 0);
    const uint32_t full_bar = __shfl_sync(0xffffffffu, sv.full_bar, 0), empty_bar = __shfl_sync(0xffffffffu, sv.empty_bar, 0);
#pragma unroll 1
    for (int ii = 0; ii < n_items; ++ii) {
      const PItem it = sc->it[ii];
      const GDesc& d = descs[it.desc];
      const CUtensorMap* mw = maps + __shfl_sync(0xffffffffu, d.w_map, 0);
      const CUtensorMap* mxh = maps + __shfl_sync(0xffffffffu, d.x_hi, 0);
      const CUtensorMap* mxl = maps + __shfl_sync(0xffffffffu, d.x_lo, 0);
      int wk = __shfl_sync(0xffffffffu, it.kb0 * 32, 0), xk = __shfl_sync(0xffffffffu, (d.xkb0 + it.kb0) * 32, 0);
      const int row = __shfl_sync(0xffffffffu, it.rt * 128, 0), col = __shfl_sync(0xffffffffu, it.cb * PK_BN, 0);
      const int nkb = __shfl_sync(0xffffffffu, (int)it.nkb, 0);
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb, ++cnt, wk += 32, xk += 32) {
        const uint32_t s = cnt & (PK_STAGES - 1);
        const uint32_t st = stages_u32 + s * PK_STAGE_BYTES, fb = full_bar + 8 * s;
        if (npre > 0) {         // weights already on their way (gemm_prefetch): only the activation tiles remain
          --npre;
          if (elect_one_sync()) {
            pk_expect_tx(fb, 2 * PK_X_BYTES);
            pk_tma_2d(st + 2 * PK_W_BYTES, mxh, fb, xk, col);
            pk_tma_2d(st + 2 * PK_W_BYTES + PK_X_BYTES, mxl, fb, xk, col);
          }
        } else {
          pk_wait(empty_bar + 8 * s, ((cnt / PK_STAGES) & 1) ^ 1);
          if (elect_one_sync()) {
            pk_expect_tx(fb, PK_TX_BYTES);
            pk_tma_2d(st, mw, fb, wk, row);
            pk_tma_2d(st + 2 * PK_W_BYTES, mxh, fb, xk, col);
            pk_tma_2d(st + 2 * PK_W_BYTES + PK_X_BYTES, mxl, fb, xk, col);
            if (ii == 0 && kb < 4) PK_FINE(1 + kb);
          }
        }
        __syncwarp();
      }
    }
    // the weight tiles of the NEXT GEMM phase go to L2 now (HBM has slack; their TMA loads then see L2 latency)
    if (sc_next != nullptr) {
      const int nn = __shfl_sync(0xffffffffu, (int)sc_next->n, 0);
#pragma unroll 1
      for (int ii = 0; ii < nn; ++ii) {
        const PItem it = sc_next->it[ii];
        const CUtensorMap* mw = maps + __shfl_sync(0xffffffffu, descs[it.desc].w_map, 0);
        int wk = __shfl_sync(0xffffffffu, it.kb0 * 32, 0);
        const int row = __shfl_sync(0xffffffffu, it.rt * 128, 0);
        const int nkb = __shfl_sync(0xffffffffu, (int)it.nkb, 0);
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb, wk += 32) {
          if (elect_one_sync()) pk_tma_prefetch_l2(mw, wk, row);
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {     // ===== MMA issuer (same discipline) =====
    constexpr uint32_t idesc = umma_idesc_tf32(128, PK_BN);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t tmem_small = tb + 2 * PK_BN;
    uint32_t cnt = __shfl_sync(0xffffffffu, ps.kb_count, 0), cc = __shfl_sync(0xffffffffu, ps.chunk_count, 0),
             ic = __shfl_sync(0xffffffffu, ps.item_count, 0);
    const uint32_t split_bar = __shfl_sync(0xffffffffu, sv.split_bar, 0), empty_bar = __shfl_sync(0xffffffffu, sv.empty_bar, 0);
    const uint32_t acc_full = __shfl_sync(0xffffffffu, sv.acc_full, 0), acc_empty = __shfl_sync(0xffffffffu, sv.acc_empty, 0);
    const uint32_t small_full = __shfl_sync(0xffffffffu, sv.small_full, 0), small_empty = __shfl_sync(0xffffffffu, sv.small_empty, 0);
    // descriptor of byte offset 0 of the stage ring; everything else is an add on the address field
    const uint32_t desc_lo0 = __shfl_sync(0xffffffffu, (uint32_t)umma_desc_sw128(sv.stages_u32), 0);
    const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
#pragma unroll 1
    for (int ii = 0; ii < n_items; ++ii, ++ic) {
      const int nkb = __shfl_sync(0xffffffffu, (int)sc->it[ii].nkb, 0);
      pk_wait(small_empty, (ic & 1) ^ 1);      // previous item's cross-term accumulator read out
      tc_fence_after();
      uint32_t tmem_main = tb, b = 0;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb, ++cnt) {
        if ((kb & 1) == 0) {
          b = cc & 1;
          pk_wait(acc_empty + 8 * b, ((cc >> 1) & 1) ^ 1);
          tc_fence_after();
          tmem_main = tb + b * PK_BN;
        }
        const uint32_t s = cnt & (PK_STAGES - 1);
        pk_wait(split_bar + 8 * s, (cnt / PK_STAGES) & 1);     // TMA landed AND lo tile written
        tc_fence_after();
        const uint32_t dlo = desc_lo0 + ((s * PK_STAGE_BYTES) >> 4);
        if (elect_one_sync()) {
          if (ii == 0 && kb < 4) PK_FINE(13 + kb);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t wh = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2);
            const uint64_t wl = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2 + (PK_W_BYTES >> 4));
            const uint64_t xh = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2 + ((2 * PK_W_BYTES) >> 4));
            const uint64_t xl = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2 + ((2 * PK_W_BYTES + PK_X_BYTES) >> 4));
            umma_tf32(tmem_main, wh, xh, idesc, ((kb & 1) | k4) != 0);
            umma_tf32(tmem_small, wl, xh, idesc, (kb | k4) != 0);
            umma_tf32(tmem_small, wh, xl, idesc, 1);
          }
          pk_commit(empty_bar + 8 * s);
          if (ii == 0 && kb < 4) PK_FINE(17 + kb);
          if ((kb & 1) || kb == nkb - 1) pk_commit(acc_full + 8 * b);
          if (kb == nkb - 1) { pk_commit(small_full); if (ii == 0) PK_FINE(21); }
        }
        __syncwarp();
        if ((kb & 1) || kb == nkb - 1) ++cc;
      }
    }
  } else if (warp < 6) {      // ===== epilogue: promote short chains into fp32 registers, store the partial tile =====
    const int quad = warp & 3;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t cc = ps.chunk_count, ic = ps.item_count;
#pragma unroll 1
    for (int ii = 0; ii < n_items; ++ii, ++ic) {
      const PItem it = sc->it[ii];
      const GDesc& d = descs[it.desc];
      const int n_chunks = (it.nkb + PK_CHUNK - 1) / PK_CHUNK;
      float acc[PK_BN];
#pragma unroll
      for (int u = 0; u < PK_BN; ++u) acc[u] = 0.f;
#pragma unroll 1
      for (int c = 0; c <= n_chunks; ++c) {        // last round: the cross-term accumulator
        uint32_t col;
        if (c < n_chunks) {
          const uint32_t b = cc & 1;
          pk_wait(sv.acc_full + 8 * b, (cc >> 1) & 1);
          col = b * PK_BN;
        } else {
          pk_wait(sv.small_full, ic & 1);
          col = 2 * PK_BN;
        }
        tc_fence_after();
        if (ii == 0 && threadIdx.x == 64) { if (c < n_chunks) { if (c < 2) PK_FINE(22 + c); } else PK_FINE(24); }
        {
          uint32_t r[32];
          tmem_ld32(taddr + col, r);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[u] += __uint_as_float(r[u]);
          tmem_ld32(taddr + col + 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[32 + u] += __uint_as_float(r[u]);
        }
        tc_fence_before();
        if (c < n_chunks) { pk_arrive(sv.acc_empty + 8 * (cc & 1)); ++cc; }
        else pk_arrive(sv.small_empty);
      }
      const int n = it.rt * 128 + quad * 32 + lane;        // weight row held by this thread
      if (n < d.n_rows) {
        float* o = d.out + ((long)it.slot * R + it.cb * PK_BN) * d.n_rows + n;   // lanes -> consecutive n
        const long str = d.n_rows;
#pragma unroll
        for (int u = 0; u < PK_BN; ++u) { __stcg(o, acc[u]); o += str; }
      }
      if (ii == 0 && threadIdx.x == 64) PK_FINE(25);
    }
  } else {                    // ===== weight split: lo = rna_tf32(w - trunc_tf32(w)); the tensor core truncates the raw tile =====
    const int t = threadIdx.x - 6 * 32;
    uint32_t cnt = ps.kb_count;
    const int tot = sc->tot_kb;
#pragma unroll 1
    for (int q0 = 0; q0 < tot; ++q0, ++cnt) {
      const uint32_t s = cnt & (PK_STAGES - 1);
      pk_wait(sv.full_bar + 8 * s, (cnt / PK_STAGES) & 1);
      if (q0 < 4 && t == 0) PK_FINE(5 + q0);
      const float4* src = reinterpret_cast<const float4*>(sv.stages + s * PK_STAGE_BYTES) + t;
      float4* dst = reinterpret_cast<float4*>(sv.stages + s * PK_STAGE_BYTES + PK_W_BYTES) + t;
#pragma unroll 4
      for (int q = 0; q < 8; ++q) {
        const float4 v = src[128 * q];
        float4 l;
        l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
        dst[128 * q] = l;
      }
      fence_proxy_async_smem();            // generic-proxy smem writes -> visible to the tensor core
      pk_arrive(sv.split_bar + 8 * s);
      if (q0 < 4 && t == 0) PK_FINE(9 + q0);
    }
  }
  ps.kb_count += sc->tot_kb;
  ps.chunk_count += sc->tot_chunks;
  ps.item_count += n_items;
  ps.npre = 0;
}

// Issued at the END of a pointwise phase, before its grid barrier: the weight tiles of the first k-blocks of
// the next GEMM phase do not depend on anything the barrier protects, so their TMA latency hides behind it.
// (tx bytes are announced without an arrival; the producer's arrive.expect_tx for the activation tiles
// completes the stage.)
__device__ __noinline__ void gemm_prefetch(const GDesc* descs, const PSched* sc, const CUtensorMap* maps, const SmemView& sv,
                                           PipeState& ps) {
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int n_items = __shfl_sync(0xffffffffu, (int)sc->n, 0);
  int npre = 0;
  if (warp == 0) {
    uint32_t cnt = __shfl_sync(0xffffffffu, ps.kb_count, 0);
    const uint32_t stages_u32 = __shfl_sync(0xffffffffu, sv.stages_u32, 0);
    const uint32_t full_bar = __shfl_sync(0xffffffffu, sv.full_bar, 0), empty_bar = __shfl_sync(0xffffffffu, sv.empty_bar, 0);
#pragma unroll 1
    for (int ii = 0; ii < n_items && npre < PK_STAGES; ++ii) {
      const PItem it = sc->it[ii];
      const CUtensorMap* mw = maps + __shfl_sync(0xffffffffu, descs[it.desc].w_map, 0);
      int wk = __shfl_sync(0xffffffffu, it.kb0 * 32, 0);
      const int row = __shfl_sync(0xffffffffu, it.rt * 128, 0);
      const int nkb = __shfl_sync(0xffffffffu, (int)it.nkb, 0);
#pragma unroll 1
      for (int kb = 0; kb < nkb && npre < PK_STAGES; ++kb, ++cnt, wk += 32, ++npre) {
        const uint32_t s = cnt & (PK_STAGES - 1);
        pk_wait(empty_bar + 8 * s, ((cnt / PK_STAGES) & 1) ^ 1);
        if (elect_one_sync()) {
          pk_expect_tx_noarrive(full_bar + 8 * s, PK_W_BYTES);
          pk_tma_2d(stages_u32 + s * PK_STAGE_BYTES, mw, full_bar + 8 * s, wk, row);
        }
        __syncwarp();
      }
    }
  }
  // every thread must agree on npre: recompute it from the schedule
  const int tot = sc->tot_kb;
  ps.npre = (uint32_t)(tot < PK_STAGES ? tot : PK_STAGES);
}

// leaving the loop with prefetched weight tiles in flight: complete their stages and wait for the data
__device__ __noinline__ void gemm_prefetch_drain(const SmemView& sv, PipeState& ps) {
  if (threadIdx.x == 0) {
    for (uint32_t q = 0; q < ps.npre; ++q) {
      const uint32_t cnt = ps.kb_count + q, s = cnt & (PK_STAGES - 1);
      pk_arrive(sv.full_bar + 8 * s);
      pk_wait(sv.full_bar + 8 * s, (cnt / PK_STAGES) & 1);
    }
  }
  ps.npre = 0;
  __syncthreads();
}

__device__ __forceinline__ void store_split(float* hi, float* lo, long idx, float v) {
  const float h = tf32_rna(v);
  hi[idx] = h;
  lo[idx] = tf32_rna(v - h);
}

// fp16 operand pair of the 3xFP16 products (xg_grouped.cuh): hi = fp16(v), lo = fp16((v - hi) * 2^11); the 11 + 11
// mantissa bits equal what a tf32 hi / lo pair carries, the scale keeps lo out of the fp16 subnormals
constexpr float X16_SCALE = 2048.f;
__device__ __forceinline__ void store_split16(__half* hi, __half* lo, long idx, float v) {
  const __half h = __float2half_rn(v);
  hi[idx] = h;
  lo[idx] = __float2half_rn((v - __half2float(h)) * X16_SCALE);
}

// split-K partial slots of element (r, n): zload issues every load (one L2 round trip for the whole
// batch a caller builds), zadd adds them in slot order.
template <int MAXS>
__device__ __forceinline__ void zload(const GDesc& d, int R, int r, int n, float (&v)[MAXS]) {
  const float* p = d.out + (long)r * d.n_rows + n;
  const long sstr = (long)R * d.n_rows;
  const int ns = d.ns;
#pragma unroll
  for (int k = 0; k < MAXS; ++k) { v[k] = k < ns ? __ldcg(p) : 0.f; p += sstr; }
}
template <int MAXS>
__device__ __forceinline__ float zadd(const float (&v)[MAXS]) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXS; ++k) s += v[k];
  return s;
}
constexpr int ENCB_MAX_SLOTS = 24;     // encoder backward: 8 strips of 64 k-blocks over 148 CTAs

__device__ __forceinline__ unsigned f2ord(float f) {      // order-preserving float -> uint (for redux.sync)
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// ex2/rcp-based activations: |error| ~1e-7 absolute, a handful of instructions each
// 1 / x for x >= 1 (or +inf): the bare MUFU.RCP.  __fdividef(1.f, x) wraps it in a subnormal-input check (FSETP, FMUL, FSEL,
// FMUL: 7 instructions per attention element instead of 3), which the attention loops cannot need: x = 1 + e^{2ah} e^{2uv}.
__device__ __forceinline__ float rcp_ge1(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(-2.f, rcp_ge1(1.f + __expf(2.f * x)), 1.f); }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_ge1(1.f + __expf(-x)); }

// ====================================================================================
// decoder
// ====================================================================================
enum { DD_AH = 0, DD_Z1H, DD_Z2H, DD_Z1X, DD_Z1G, DD_Z2X, DD_Z2A, DD_LOGIT, DD_COUNT };

struct DecParams {
  GDesc d[PK_MAX_DESCS];
  const PSched* sched;          // [3][G]
  int B, R, K, H, E, Ep, A, V, T;
  const float *b_h2a, *w_a2w, *b_a2w;
  const float* bias[2][3];      // lstm_1 / lstm_2: i2h, a2h, h2h biases
  const float *b_logit, *embed;
  const float* tgate;           // (V, H)  relu(embed . W_gate^T + b): the POS-gate pre-factor of every token
  const float *Vf, *Uv, *pos;   // (B,K,H), (B,K,A), (B,H)
  float* EUv;                   // (B,K,A) exp(2 Uv), built in the prologue: tanh(ah + uv) = 1 - 2 / (e^{2ah} e^{2uv} + 1)
  const float* state0[4];       // h1,c1,h2,c2 each (B,H)
  float *xt_hi, *xt_lo;         // [R][Ep]
  float *hh_hi, *hh_lo;         // [R][2H]   [h1 | h2]
  float *gp_hi, *gp_lo;         // [R][H]
  float *af_hi, *af_lo;         // [R][H]
  float *hx;                    // [R][2H] exact states
  float *cx;                    // [2][R][H]
  float *unfinished;            // [R]
  int64_t *tok;                 // [R]
  int64_t* seq; float* seqlogp; int* flags;   // (B,T), (B,T), (T)
  unsigned int* sync_counter;
  long long* dbg_clock;         // diagnostics, or NULL
  // single-step mode (beam search: B rows = videos x beam, every feat_div consecutive rows share V / Uv / pos):
  int feat_div, build_euv;
  const int64_t* tokens_in;     // (B) input token of every row
  float* state_out[4];          // (B,H) h1,c1,h2,c2 after the step (may alias state0)
  float* logp_out;              // (B,V) log-softmax of the step, or NULL
  const int* parent_in;         // (B) row of state0 each row continues from (beam reordering), or NULL = identity
  float* ys_out; int* ix_out;   // (B,topk) per-row top-k of the log-probs with the UNK penalty (CaptionModel.py:94), or NULL
  int topk;
  // mode 1: teacher-forced training forward (SAModel.forward, SAModel.py:88-111) — no logit / pick phases; the
  // input parts of lstm_1 are hoisted (G1s holds them + all three biases on entry); every activation the
  // hand-written backward needs is stored in the step-major layouts of TrainSaved
  int mode, L;
  const float* seq_mask;        // (B, L)
  float *G1s, *G2s;             // (T, B, 4H)   pre-activations -> activated gates (i,f,o,g)
  float *C1s, *C2s;             // (T+1, B, H)
  float *H12s;                  // (T+1, B, 2H) [h1 | h2] entering each step (after dropout)
  float *AHs, *ALPHAs, *AFs;    // (T,B,A), (T,B,K), (T,B,H)
  DropSpec drop1, drop2;        // dropout on the carried h of lstm_1 / lstm_2 (index = t*B*H + b*H + j)
  int x16;                      // 1: xt / gp / af are fp16 hi / lo pairs (the grouped greedy kernel, xg_grouped.cuh)
};

constexpr int DEC_NA = 5;          // attention units per thread: att <= 5 * 320

// lstm cell (decoder gate order i,f,o,g), elements (r, j) with j fastest
template <int TRAIN>
__device__ __noinline__ void dec_cell_phase(const DecParams& P, int layer, int t, int part, int nparts, long long* g_fine = nullptr) {
  if (threadIdx.x == 0) PK_FINE(0);
  const int H = P.H, R = P.R, B = P.B;
  constexpr bool train = TRAIN != 0;
  const GDesc& da = P.d[layer == 0 ? DD_Z1X : DD_Z2X];
  const GDesc& db = P.d[layer == 0 ? DD_Z1G : DD_Z2A];
  const GDesc& dc = P.d[layer == 0 ? DD_Z1H : DD_Z2H];
  const float* bi = P.bias[layer][0]; const float* ba = P.bias[layer][1]; const float* bh = P.bias[layer][2];
  float* cst = P.cx + (long)layer * R * H;
  const bool hoisted = train && layer == 0;      // lstm_1 in training: x / gp parts + biases already sit in G1s
#pragma unroll 1
  for (int e = part * PK_THREADS + threadIdx.x; e < B * H; e += nparts * PK_THREADS) {
    const int r = e / H, j = e % H;
    float va[4][PK_MAX_SLOTS], vb[4][PK_MAX_SLOTS], vc[4][PK_MAX_SLOTS], bias[4];
    float* gsave = train ? (layer == 0 ? P.G1s : P.G2s) + ((long)t * B + r) * 4 * H + j : nullptr;
#pragma unroll
    for (int g = 0; g < 4; ++g) {              // every load of the element is in flight before the first add
      const int n = g * H + j;
      zload(dc, R, r, n, vc[g]);
      if (hoisted) {
        bias[g] = gsave[g * H];
      } else {
        zload(da, R, r, n, va[g]); zload(db, R, r, n, vb[g]);
        bias[g] = __ldg(bi + n) + __ldg(ba + n) + __ldg(bh + n);
      }
    }
    float m, cp, hp;
    float* hxp = P.hx + (long)r * 2 * H + layer * H + j;
    if (train) {
      m = __ldg(P.seq_mask + (long)r * P.L + t);
      cp = (layer == 0 ? P.C1s : P.C2s)[((long)t * B + r) * H + j];
      hp = P.H12s[((long)t * B + r) * 2 * H + layer * H + j];
    } else {
      m = t > 0 ? __ldcg(P.unfinished + r) : 1.f;
      cp = __ldcg(cst + e);
      hp = __ldcg(hxp);
    }
    float z[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) z[g] = hoisted ? bias[g] + zadd(vc[g]) : zadd(va[g]) + zadd(vb[g]) + zadd(vc[g]) + bias[g];
    if (threadIdx.x == 0 && z[0] != 12345.f) PK_FINE(1);
    const float ig = sigmoid_fast(z[0]), fg = sigmoid_fast(z[1]), og = sigmoid_fast(z[2]), gg = tanh_fast(z[3]);
    float c = fg * cp + ig * gg;
    c = c * m + cp * (1.f - m);
    float h = og * tanh_fast(c);
    h = h * m + hp * (1.f - m);
    if (train) {
      h *= (layer == 0 ? P.drop1 : P.drop2).factor((uint64_t)t * B * H + (uint64_t)e);
      gsave[0] = ig; gsave[H] = fg; gsave[2 * H] = og; gsave[3 * H] = gg;
      (layer == 0 ? P.C1s : P.C2s)[((long)(t + 1) * B + r) * H + j] = c;
      P.H12s[((long)(t + 1) * B + r) * 2 * H + layer * H + j] = h;
    } else {
      cst[e] = c;
      *hxp = h;
    }
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + layer * H + j, h);
    if (threadIdx.x == 0) PK_FINE(2);
  }
  if (threadIdx.x == 0) PK_FINE(3);
}

// temporal attention of caption r on one CTA (sub_modules.py:677-680).  Uv[r] arrives in 4-frame chunks by bulk
// copies into the (idle) pipeline stages; every thread owns the attention units a = tid + 320 i (ah[a], w[a]
// in registers) and runs the frames of a chunk with independent accumulators (their warp reductions
// interleave); frame scores are reduced warp -> CTA in a fixed order; softmax over ALL K frames; V[r] is
// fetched by TMA (two H/2-column panels) behind Uv when both fit, else over the first two consumed chunks.
// (A CTA pair per caption was tried: the score exchange costs what halving the frames saves.)
constexpr int DEC_FPC = 4;         // frames per Uv chunk (K <= 32): the first chunk (24 KB) lands before the query is ready
template <int TRAIN>
__device__ __noinline__ void dec_attention(const DecParams& P, const CUtensorMap* vmap, int r, int t, const SmemView& sv,
                                           uint32_t& bulk_phase, long long* g_fine = nullptr) {
  if (threadIdx.x == 0) PK_FINE(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = P.K, A = P.A, H = P.H, Hh = H / 2;
  const int fb = r / P.feat_div;                // feature row (beam rows of one video share V / Uv)
  float* red = sv.scratch;                      // [PK_WARPS][K] per-warp partial scores
  float* sc = sv.scratch + PK_WARPS * K;        // K floats
  const float* uv = reinterpret_cast<const float*>(sv.stages);
  constexpr int fpc = DEC_FPC;
  const bool v_behind = (long)K * (A + H) * 4 <= (long)PK_STAGES * PK_STAGE_BYTES;
  const uint32_t v_off = v_behind ? (uint32_t)K * A * 4u : 0u;
  const int vc = max(0, (K * H + fpc * A - 1) / (fpc * A) - 1);   // chunk after which the consumed chunks cover V[r]
  if (warp == 0) {              // lane c requests chunk c: the chunk copies are issued in one pass of the warp
    const int k0 = lane * fpc, k1 = min(K, k0 + fpc);
    if (lane < PK_BULK_CHUNKS && k0 < k1) {
      const uint32_t nb = (uint32_t)(k1 - k0) * (uint32_t)A * 4u;
      pk_expect_tx(sv.bulk_bar + 8 * lane, nb);
      bulk_g2s(sv.stages_u32 + (uint32_t)k0 * A * 4u, P.EUv + ((long)fb * K + k0) * A, nb, sv.bulk_bar + 8 * lane);
    }
    if (lane == 0 && v_behind) {
      pk_expect_tx(sv.bulk_bar + 8 * PK_BULK_CHUNKS, (uint32_t)K * H * 4u);
      pk_tma_2d(sv.stages_u32 + v_off, vmap, sv.bulk_bar + 8 * PK_BULK_CHUNKS, 0, fb * K);
      pk_tma_2d(sv.stages_u32 + v_off + (uint32_t)K * Hh * 4u, vmap, sv.bulk_bar + 8 * PK_BULK_CHUNKS, Hh, fb * K);
    }
    __syncwarp();
  }
  float ahr[DEC_NA], wr[DEC_NA];
  int aoff[DEC_NA];
  {
    float v[DEC_NA][PK_MAX_SLOTS];
#pragma unroll
    for (int i = 0; i < DEC_NA; ++i) {
      const int a = min(threadIdx.x + PK_THREADS * i, A - 1);
      aoff[i] = a;
      zload(P.d[DD_AH], P.R, r, a, v[i]);
      ahr[i] = __ldg(P.b_h2a + a);
      wr[i] = (threadIdx.x + PK_THREADS * i < A) ? __ldg(P.w_a2w + a) : 0.f;     // out-of-range units weigh 0
    }
#pragma unroll
    for (int i = 0; i < DEC_NA; ++i) ahr[i] += zadd(v[i]);
    if (TRAIN) {
#pragma unroll
      for (int i = 0; i < DEC_NA; ++i)
        if (threadIdx.x + PK_THREADS * i < A) P.AHs[((long)t * P.B + r) * A + aoff[i]] = ahr[i];
    }
  }
  // tanh(ah + uv) = 1 - 2 / (e^{2 ah} e^{2 uv} + 1): e^{2 uv} is step-invariant (EUv), so each of the 43k
  // evaluations costs one MUFU (rcp) instead of two; the constant sum of the weights is added once
  float wsum = 0.f;
#pragma unroll
  for (int i = 0; i < DEC_NA; ++i) {
    wsum += wr[i];
    ahr[i] = __expf(2.f * fminf(fmaxf(ahr[i], -40.f), 40.f));
    wr[i] *= -2.f;
  }
  if (threadIdx.x == 0) PK_FINE(1);
#pragma unroll 1
  for (int c = 0; c < PK_BULK_CHUNKS; ++c) {
    const int k0 = c * fpc, k1 = min(K, k0 + fpc);
    if (k0 >= k1) break;
    pk_wait(sv.bulk_bar + 8 * c, bulk_phase & 1);
    if (threadIdx.x == 0) PK_FINE(2 + 2 * c);
    float p[DEC_FPC];
#pragma unroll
    for (int f = 0; f < DEC_FPC; ++f) {
      p[f] = wsum;
      const float* u = uv + (long)min(k0 + f, K - 1) * A;
#pragma unroll
      for (int i = 0; i < DEC_NA; ++i) p[f] = fmaf(wr[i], rcp_ge1(fmaf(ahr[i], u[aoff[i]], 1.f)), p[f]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int f = 0; f < DEC_FPC; ++f) p[f] += __shfl_xor_sync(0xffffffffu, p[f], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int f = 0; f < DEC_FPC; ++f) if (k0 + f < k1) red[warp * K + k0 + f] = p[f];
    }
    if (threadIdx.x == 0) PK_FINE(3 + 2 * c);
    if (!v_behind && (c == vc || k1 == K)) {  // the first chunks are consumed by every warp: V[r] goes over them
      __syncthreads();
      if (threadIdx.x == 0 && (c == vc || (c < vc && k1 == K))) {
        fence_proxy_async_smem();
        pk_expect_tx(sv.bulk_bar + 8 * PK_BULK_CHUNKS, (uint32_t)K * H * 4u);
        pk_tma_2d(sv.stages_u32, vmap, sv.bulk_bar + 8 * PK_BULK_CHUNKS, 0, fb * K);
        pk_tma_2d(sv.stages_u32 + (uint32_t)K * Hh * 4u, vmap, sv.bulk_bar + 8 * PK_BULK_CHUNKS, Hh, fb * K);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) PK_FINE(20);
  if (warp == 0) {
    const float ba = __ldg(P.b_a2w);
    float mx = -INFINITY;
#pragma unroll 1
    for (int kk = lane; kk < K; kk += 32) {
      float q = 0.f;
#pragma unroll
      for (int w = 0; w < PK_WARPS; ++w) q += red[w * K + kk];
      q += ba;
      sc[kk] = q;
      mx = fmaxf(mx, q);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll 1
    for (int kk = lane; kk < K; kk += 32) { const float e = __expf(sc[kk] - mx); sc[kk] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll 1
    for (int kk = lane; kk < K; kk += 32) {
      sc[kk] *= inv;
      if (TRAIN) P.ALPHAs[((long)t * P.B + r) * K + kk] = sc[kk];
    }
  }
  pk_wait(sv.bulk_bar + 8 * PK_BULK_CHUNKS, bulk_phase & 1);
  bulk_phase++;
  __syncthreads();
  if (threadIdx.x == 0) PK_FINE(21);
  const float* vs = reinterpret_cast<const float*>(sv.stages + v_off);            // two panels [K][Hh]
#pragma unroll 1
  for (int j = threadIdx.x; j < H; j += PK_THREADS) {
    const float* vp = vs + (j >= Hh ? K * Hh + (j - Hh) : j);
    float a = 0.f;
#pragma unroll 4
    for (int k = 0; k < K; ++k) a += sc[k] * vp[k * Hh];
    if (TRAIN) P.AFs[((long)t * P.B + r) * H + j] = a;
    if (P.x16) store_split16(reinterpret_cast<__half*>(P.af_hi), reinterpret_cast<__half*>(P.af_lo), (long)r * H + j, a);
    else store_split(P.af_hi, P.af_lo, (long)r * H + j, a);
  }
  if (threadIdx.x == 0) PK_FINE(22);
  __syncthreads();
}

// next-step inputs of caption r for token `tokv`: xt = embed[tok] and gp = pos * (1 + tgate[tok])
// (all loads of a thread are issued before the first store: one L2 round trip)
constexpr int DEC_TI = 2;          // columns per thread: embed (padded) and rnn <= 2 * 320
__device__ __noinline__ void dec_token_inputs(const DecParams& P, int r, int tokv) {
  const float* src = P.embed + (long)tokv * P.E;
  const float* tg = P.tgate + (long)tokv * P.H;
  const float* ps = P.pos + (long)(r / P.feat_div) * P.H;
  float x[DEC_TI], g[DEC_TI], q[DEC_TI];
#pragma unroll
  for (int i = 0; i < DEC_TI; ++i) {
    const int k = threadIdx.x + i * PK_THREADS;
    x[i] = k < P.E ? __ldg(src + k) : 0.f;
    g[i] = k < P.H ? __ldcg(tg + k) : 0.f;
    q[i] = k < P.H ? __ldg(ps + k) : 0.f;
  }
#pragma unroll
  for (int i = 0; i < DEC_TI; ++i) {
    const int k = threadIdx.x + i * PK_THREADS;
    if (P.x16) {
      if (k < P.Ep) store_split16(reinterpret_cast<__half*>(P.xt_hi), reinterpret_cast<__half*>(P.xt_lo), (long)r * P.Ep + k, x[i]);
      if (k < P.H) store_split16(reinterpret_cast<__half*>(P.gp_hi), reinterpret_cast<__half*>(P.gp_lo), (long)r * P.H + k, q[i] * (1.f + g[i]));
    } else {
      if (k < P.Ep) store_split(P.xt_hi, P.xt_lo, (long)r * P.Ep + k, x[i]);
      if (k < P.H) store_split(P.gp_hi, P.gp_lo, (long)r * P.H + k, q[i] * (1.f + g[i]));
    }
  }
}

// greedy pick of caption r at step t on one CTA (SAModel.py:185-210): logits = sum of the split-K slots + bias,
// staged in shared memory (the pipeline stages are idle); max / lowest argmax, then log-sum-exp; returns the
// raw argmax token to every thread.
constexpr int DEC_PB = 32;         // logits per thread per batch of loads
__device__ __noinline__ int dec_pick(const DecParams& P, int r, int t, const SmemView& sv, long long* g_fine = nullptr) {
  if (threadIdx.x == 0) PK_FINE(0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = P.V, R = P.R;
  const GDesc& dl = P.d[DD_LOGIT];
  const long sstr = (long)R * dl.n_rows;
  float* lg = reinterpret_cast<float*>(sv.stages);            // V floats
  float* redf = sv.scratch;                                   // [PK_WARPS]
  int* redi = reinterpret_cast<int*>(sv.scratch + 16);        // [PK_WARPS]
  const float* lp0 = dl.out + (long)r * dl.n_rows;
  float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll 1
  for (int n0 = threadIdx.x; n0 < V; n0 += PK_THREADS * DEC_PB) {
    float v[DEC_PB];
#pragma unroll
    for (int i = 0; i < DEC_PB; ++i) v[i] = 0.f;
    const float* lp = lp0;
#pragma unroll 1
    for (int k = 0; k < dl.ns; ++k, lp += sstr) {        // slot-major: DEC_PB independent loads in flight per slot
      float tl[DEC_PB];
#pragma unroll
      for (int i = 0; i < DEC_PB; ++i) tl[i] = __ldcg(lp + min(n0 + i * PK_THREADS, V - 1));
#pragma unroll
      for (int i = 0; i < DEC_PB; ++i) v[i] += tl[i];
      if (threadIdx.x == 0 && v[0] != 12345.f && n0 < PK_THREADS) PK_FINE(1 + k);
    }
    float bl[DEC_PB];
#pragma unroll
    for (int i = 0; i < DEC_PB; ++i) bl[i] = __ldg(P.b_logit + min(n0 + i * PK_THREADS, V - 1));
#pragma unroll
    for (int i = 0; i < DEC_PB; ++i) {
      const int n = n0 + i * PK_THREADS;
      const float x = v[i] + bl[i];
      if (n < V) {
        lg[n] = x;
        if (x > best) { best = x; bi = n; }              // ascending n per thread: first maximum kept
      }
    }
  }
  {
    const unsigned mo = __reduce_max_sync(0xffffffffu, f2ord(best));
    const int wbi = (int)__reduce_min_sync(0xffffffffu, (f2ord(best) == mo) ? (unsigned)bi : 0x7fffffffu);
    if (lane == 0) { redf[warp] = ord2f(mo); redi[warp] = wbi; }
  }
  if (threadIdx.x == 0) PK_FINE(6);
  __syncthreads();
  best = redf[0]; bi = redi[0];
#pragma unroll
  for (int w = 1; w < PK_WARPS; ++w) {
    const float ov = redf[w]; const int oi = redi[w];
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  float s = 0.f;
#pragma unroll 4
  for (int n = threadIdx.x; n < V; n += PK_THREADS) s += __expf(lg[n] - best);
  s = warp_sum(s);
  __syncthreads();                                            // redf is reused
  if (lane == 0) redf[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < PK_WARPS; ++w) tot += redf[w];
    float unf = (t == 0) ? 1.f : __ldcg(P.unfinished + r);
    unf = (unf != 0.f && bi > 0) ? 1.f : 0.f;
    P.unfinished[r] = unf;
    P.seq[(long)r * P.T + t] = unf != 0.f ? (int64_t)bi : 0;
    P.seqlogp[(long)r * P.T + t] = -logf(tot);
    P.tok[r] = bi;
    if (unf != 0.f) P.flags[t] = 1;
    PK_FINE(7);
  }
  return bi;
}

// log-softmax of row r (single-step mode): logits = sum of the split-K slots + bias staged in shared memory
__device__ __noinline__ void dec_logsoftmax_row(const DecParams& P, int r, const SmemView& sv) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = P.V, R = P.R;
  const GDesc& dl = P.d[DD_LOGIT];
  const long sstr = (long)R * dl.n_rows;
  float* lg = reinterpret_cast<float*>(sv.stages);            // V floats
  float* redf = sv.scratch;                                   // [PK_WARPS]
  const float* lp0 = dl.out + (long)r * dl.n_rows;
  float best = -INFINITY;
#pragma unroll 1
  for (int n0 = threadIdx.x; n0 < V; n0 += PK_THREADS * DEC_PB) {
    float v[DEC_PB];
#pragma unroll
    for (int i = 0; i < DEC_PB; ++i) v[i] = 0.f;
    const float* lp = lp0;
#pragma unroll 1
    for (int k = 0; k < dl.ns; ++k, lp += sstr) {
      float tl[DEC_PB];
#pragma unroll
      for (int i = 0; i < DEC_PB; ++i) tl[i] = __ldcg(lp + min(n0 + i * PK_THREADS, V - 1));
#pragma unroll
      for (int i = 0; i < DEC_PB; ++i) v[i] += tl[i];
    }
    float bl[DEC_PB];
#pragma unroll
    for (int i = 0; i < DEC_PB; ++i) bl[i] = __ldg(P.b_logit + min(n0 + i * PK_THREADS, V - 1));
#pragma unroll
    for (int i = 0; i < DEC_PB; ++i) {
      const int n = n0 + i * PK_THREADS;
      const float x = v[i] + bl[i];
      if (n < V) { lg[n] = x; best = fmaxf(best, x); }
    }
  }
  best = warp_max(best);
  if (lane == 0) redf[warp] = best;
  __syncthreads();
  best = redf[0];
#pragma unroll
  for (int w = 1; w < PK_WARPS; ++w) best = fmaxf(best, redf[w]);
  float s = 0.f;
#pragma unroll 4
  for (int n = threadIdx.x; n < V; n += PK_THREADS) s += expf(lg[n] - best);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) redf[warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < PK_WARPS; ++w) tot += redf[w];
  const float lse = best + logf(tot);
  float* out = P.logp_out ? P.logp_out + (long)r * V : nullptr;
#pragma unroll 4
  for (int n = threadIdx.x; n < V; n += PK_THREADS) {
    const float y = lg[n] - lse;
    if (out) out[n] = y;
    lg[n] = n == 1 ? y - 1000.f : y;           // UNK penalty before the selection (CaptionModel.py:94)
  }
  __syncthreads();
  if (P.ys_out) {
    // per-row top-k for the beam step: value descending, lowest index first on exact ties (the order a stable
    // descending sort of the row gives, CaptionModel.py:40); a selected entry becomes NaN (never compares greater)
    // Every thread keeps the best of the entries it owns (n = tid mod 320); after a selection only the owner of
    // the selected entry rescans.
    int* redi = reinterpret_cast<int*>(sv.scratch + 16);
    auto scan_own = [&](float& bv, int& bi) {
      bv = -INFINITY; bi = 0x7fffffff;
#pragma unroll 4
      for (int n = threadIdx.x; n < V; n += PK_THREADS) {
        const float v = lg[n];
        if (v > bv || (v == bv && n < bi)) { bv = v; bi = n; }
      }
    };
    float bv; int bi;
    scan_own(bv, bi);
#pragma unroll 1
    for (int c = 0; c < P.topk; ++c) {
      float wv = bv; int wi = bi;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
        if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; }
      }
      if (lane == 0) { redf[warp] = wv; redi[warp] = wi; }
      __syncthreads();
      wv = redf[0]; wi = redi[0];
#pragma unroll
      for (int w = 1; w < PK_WARPS; ++w)
        if (redf[w] > wv || (redf[w] == wv && redi[w] < wi)) { wv = redf[w]; wi = redi[w]; }
      if (threadIdx.x == 0) {
        P.ys_out[(long)r * P.topk + c] = wv;
        P.ix_out[(long)r * P.topk + c] = wi;
      }
      if (wi < V && wi % PK_THREADS == (int)threadIdx.x) {
        lg[wi] = __int_as_float(0x7fc00000);
        scan_own(bv, bi);
      }
      __syncthreads();
    }
  }
}

// diagnostics (XG_PERSIST_TRACE): SM-clock stamps of CTA 0 for every step + globaltimer stamps of EVERY CTA for step 3
__device__ __noinline__ void pk_stamp(long long* dbg, int cta, int t, int i) {
  if (dbg == nullptr || threadIdx.x != 0) return;
  if (cta == 0) dbg[t * PK_STAMPS + i] = clock64();
  if (t == 3) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    dbg[(2048 + cta) * PK_STAMPS + i] = (long long)gt;
  }
}

template <int TRAIN>
__global__ void __launch_bounds__(PK_THREADS, 1)
decode_persistent_kernel(const DecParams* __restrict__ Pp, const __grid_constant__ MapTable maps) {
  __shared__ DecParams Psm;
  __shared__ PSched s_sched[3];
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(DecParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Psm)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  __syncthreads();
  const DecParams& P = Psm;
  for (int i = threadIdx.x; i < (int)(3 * sizeof(PSched) / 4); i += PK_THREADS) {
    const int ph = i / (int)(sizeof(PSched) / 4), w = i % (int)(sizeof(PSched) / 4);
    reinterpret_cast<uint32_t*>(&s_sched[ph])[w] = reinterpret_cast<const uint32_t*>(P.sched + (long)ph * G + cta)[w];
  }
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = P.H, R = P.R, B = P.B, T = P.T;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 17) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = 0;
  uint32_t bulk_phase = 0;

  constexpr bool train = TRAIN != 0;
  // ---- prologue: states, <bos> inputs, bookkeeping ----
  for (int e = cta * PK_THREADS + threadIdx.x; e < R * H; e += G * PK_THREADS) {
    const int r = e / H, j = e % H;
    float h1 = 0.f, h2 = 0.f;
    if (r < B) {
      if (train) {                                    // init_hidden wrote H12s[0] / C1s[0] / C2s[0]
        h1 = P.H12s[(long)r * 2 * H + j]; h2 = P.H12s[(long)r * 2 * H + H + j];
      } else {
        h1 = P.state0[0][e]; h2 = P.state0[2][e];
        P.cx[e] = P.state0[1][e]; P.cx[(long)R * H + e] = P.state0[3][e];
      }
    }
    P.hx[(long)r * 2 * H + j] = h1; P.hx[(long)r * 2 * H + H + j] = h2;
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + j, h1);
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + H + j, h2);
  }
  for (int r = cta; r < R; r += G) {
    if (r < B && !train) {
      dec_token_inputs(P, r, 0);                      // token 0 = <bos> (SAModel.py:184)
    } else if (r >= B) {                              // padding rows of the 64-wide operand tiles
      for (int k = threadIdx.x; k < P.Ep; k += PK_THREADS) { P.xt_hi[(long)r * P.Ep + k] = 0.f; P.xt_lo[(long)r * P.Ep + k] = 0.f; }
      for (int j = threadIdx.x; j < H; j += PK_THREADS) {
        P.gp_hi[(long)r * H + j] = 0.f; P.gp_lo[(long)r * H + j] = 0.f;
        P.af_hi[(long)r * H + j] = 0.f; P.af_lo[(long)r * H + j] = 0.f;
      }
    }
    if (threadIdx.x == 0) { P.unfinished[r] = 1.f; P.tok[r] = 0; }
  }
  {   // EUv = exp(2 Uv), clamped like the per-step factor
    const long n = (long)B * P.K * P.A;
    for (long e = (long)cta * PK_THREADS + threadIdx.x; e < n; e += (long)G * PK_THREADS)
      P.EUv[e] = __expf(2.f * fminf(fmaxf(__ldg(P.Uv + e), -40.f), 40.f));
  }
  if constexpr (train) {
    // ---------------- teacher-forced word loop: G1 {ah, z1h, z2h} -> P1 -> G3 -> P3 ----------------
    gemm_prefetch(P.d, &s_sched[0], maps.m, sv, ps);
    grid_barrier(P.sync_counter, sync_target, G);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      gemm_phase(P.d, &s_sched[0], &s_sched[1], maps.m, R, sv, tmem_base, ps);          // AH, Z1h, Z2h
      grid_barrier(P.sync_counter, sync_target, G);
      if (G >= B + 8) {              // enough SMs: captions and cell elements on disjoint CTAs
        if (cta < B) dec_attention<TRAIN>(P, &maps.m[16], cta, t, sv, bulk_phase);
        else dec_cell_phase<TRAIN>(P, 0, t, cta - B, G - B);
      } else {                       // large batches: every CTA walks its captions, then its cell elements
#pragma unroll 1
        for (int r = cta; r < B; r += G) dec_attention<TRAIN>(P, &maps.m[16], r, t, sv, bulk_phase);
        dec_cell_phase<TRAIN>(P, 0, t, cta, G);
      }
      fence_proxy_async_smem();      // stages were used through the generic + bulk paths: order before TMA reuse
      gemm_prefetch(P.d, &s_sched[1], maps.m, sv, ps);
      grid_barrier(P.sync_counter, sync_target, G);
      gemm_phase(P.d, &s_sched[1], &s_sched[0], maps.m, R, sv, tmem_base, ps);          // Z2x, Z2a
      grid_barrier(P.sync_counter, sync_target, G);
      dec_cell_phase<TRAIN>(P, 1, t, cta, G);
      if (t + 1 < T) gemm_prefetch(P.d, &s_sched[0], maps.m, sv, ps);
      grid_barrier(P.sync_counter, sync_target, G);
    }                                // the heads are batched over all steps after the loop (SAModel.py:109-110)
  } else {
    // ---------------- greedy word loop ----------------
    // The products that need only the states — AH, Z1h, Z2h of step t+1 — ride in the logits phase of step t, so
    // the attention of step t+1 runs NEXT TO the pick of step t (both are one-CTA-per-caption phases) and the
    // token-dependent phase G1 shrinks to Z1x, Z1g.  Before the loop the same schedule runs once on the initial
    // state (its logits are ignored; same schedule = same slot layout).
    gemm_prefetch(P.d, &s_sched[2], maps.m, sv, ps);
    grid_barrier(P.sync_counter, sync_target, G);
    gemm_phase(P.d, &s_sched[2], &s_sched[0], maps.m, R, sv, tmem_base, ps);
    gemm_prefetch(P.d, &s_sched[0], maps.m, sv, ps);
    grid_barrier(P.sync_counter, sync_target, G);
    const bool split_roles = G >= 2 * B;
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      pk_stamp(P.dbg_clock, cta, t, 0);
      // ===== G1: Z1x = W_i2h1.xt   Z1g = W_a2h1.gp =====
      gemm_phase(P.d, &s_sched[0], &s_sched[1], maps.m, R, sv, tmem_base, ps);
      pk_stamp(P.dbg_clock, cta, t, 1);
      grid_barrier(P.sync_counter, sync_target, G);
      pk_stamp(P.dbg_clock, cta, t, 2);
      // ===== P1: lstm_1 cell (step 0 also runs the attention of step 0) =====
      if (t == 0) {
        if (G >= B + 8) {
          if (cta < B) dec_attention<TRAIN>(P, &maps.m[16], cta, 0, sv, bulk_phase);
          else dec_cell_phase<TRAIN>(P, 0, 0, cta - B, G - B);
        } else {
#pragma unroll 1
          for (int r = cta; r < B; r += G) dec_attention<TRAIN>(P, &maps.m[16], r, 0, sv, bulk_phase);
          dec_cell_phase<TRAIN>(P, 0, 0, cta, G);
        }
        fence_proxy_async_smem();
      } else {
        dec_cell_phase<TRAIN>(P, 0, t, cta, G);
      }
      gemm_prefetch(P.d, &s_sched[1], maps.m, sv, ps);
      pk_stamp(P.dbg_clock, cta, t, 3);
      grid_barrier(P.sync_counter, sync_target, G);
      pk_stamp(P.dbg_clock, cta, t, 4);
      // ===== G3: Z2x = W_i2h2.h1'   Z2a = W_a2h2.af =====
      gemm_phase(P.d, &s_sched[1], &s_sched[2], maps.m, R, sv, tmem_base, ps);
      pk_stamp(P.dbg_clock, cta, t, 5);
      grid_barrier(P.sync_counter, sync_target, G);
      pk_stamp(P.dbg_clock, cta, t, 6);
      // ===== P3: lstm_2 cell =====
      dec_cell_phase<TRAIN>(P, 1, t, cta, G);
      gemm_prefetch(P.d, &s_sched[2], maps.m, sv, ps);
      pk_stamp(P.dbg_clock, cta, t, 7);
      grid_barrier(P.sync_counter, sync_target, G);
      pk_stamp(P.dbg_clock, cta, t, 8);
      // ===== G4: logits of step t  +  AH, Z1h, Z2h of step t+1 =====
      gemm_phase(P.d, &s_sched[2], &s_sched[0], maps.m, R, sv, tmem_base, ps);
      pk_stamp(P.dbg_clock, cta, t, 9);
      grid_barrier(P.sync_counter, sync_target, G);
      pk_stamp(P.dbg_clock, cta, t, 10);
      // ===== P4: greedy pick + next-step inputs (CTAs < B)  ||  attention of step t+1 (CTAs B..2B-1) =====
      if (split_roles) {
#ifdef PK_FINE_TRACE   // -DPK_FINE_TRACE + XG_PERSIST_TRACE=1: SM-clock stamps inside the pick (CTA 0) and the attention (CTA B) of step 3
        long long* gf = (P.dbg_clock && t == 3 && (cta == 0 || cta == B)) ? P.dbg_clock + (2048 + 256) * PK_STAMPS + (cta == 0 ? 0 : 32) : nullptr;
        if (gf && threadIdx.x == 0) gf[31] = clock64();
#else
        long long* gf = nullptr;
#endif
        if (cta < B) {
          const int tokv = dec_pick(P, cta, t, sv, gf);
          dec_token_inputs(P, cta, tokv);
          __syncthreads();
#ifdef PK_FINE_TRACE
          if (gf && threadIdx.x == 0) gf[30] = clock64();
#endif
        } else if (cta < 2 * B && t + 1 < T) {
          dec_attention<TRAIN>(P, &maps.m[16], cta - B, t + 1, sv, bulk_phase, gf);
        }
      } else {
#pragma unroll 1
        for (int r = cta; r < B; r += G) {
          const int tokv = dec_pick(P, r, t, sv);
          dec_token_inputs(P, r, tokv);
          __syncthreads();
        }
        if (t + 1 < T) {
          fence_proxy_async_smem();
#pragma unroll 1
          for (int r = cta; r < B; r += G) dec_attention<TRAIN>(P, &maps.m[16], r, t + 1, sv, bulk_phase);
        }
      }
      fence_proxy_async_smem();
      if (t + 1 < T) gemm_prefetch(P.d, &s_sched[0], maps.m, sv, ps);
      pk_stamp(P.dbg_clock, cta, t, 11);
      grid_barrier(P.sync_counter, sync_target, G);
      pk_stamp(P.dbg_clock, cta, t, 12);
      if (__ldcg(P.flags + t) == 0) break;     // every caption finished (SAModel.py:206)
    }
  }
  gemm_prefetch_drain(sv, ps);               // early exit with weight tiles in flight
  pipeline_teardown(tmem_base);
}

constexpr unsigned int PK_STEP_BARRIERS = 6;   // grid barriers of one decode_step_persistent_kernel launch
// ONE word step for arbitrary state rows (beam search, CaptionModel.py:121-125 -> SAModel.get_logprobs_state,
// SAModel.py:117-127): states and tokens in, states and the log-softmax of every row out.  Launched once per step;
// exp(2 Uv) and the POS-gate token table persist in the pool between launches.
__global__ void __launch_bounds__(PK_THREADS, 1)
decode_step_persistent_kernel(const DecParams* __restrict__ Pp, const __grid_constant__ MapTable maps, unsigned int sync_base) {
  __shared__ DecParams Psm;
  __shared__ PSched s_sched[3];
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(DecParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Psm)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  __syncthreads();
  const DecParams& P = Psm;
  for (int i = threadIdx.x; i < (int)(3 * sizeof(PSched) / 4); i += PK_THREADS) {
    const int ph = i / (int)(sizeof(PSched) / 4), w = i % (int)(sizeof(PSched) / 4);
    reinterpret_cast<uint32_t*>(&s_sched[ph])[w] = reinterpret_cast<const uint32_t*>(P.sched + (long)ph * G + cta)[w];
  }
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = P.H, R = P.R, B = P.B;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 17) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = sync_base;   // the barrier counter runs on from launch to launch (PK_STEP_BARRIERS per launch)
  uint32_t bulk_phase = 0;

  for (int e = cta * PK_THREADS + threadIdx.x; e < R * H; e += G * PK_THREADS) {
    const int r = e / H, j = e % H;
    float h1 = 0.f, h2 = 0.f;
    if (r < B) {
      const long src = P.parent_in ? (long)__ldg(P.parent_in + r) * H + j : (long)e;   // beam reordering (CaptionModel.py:62-64)
      h1 = P.state0[0][src]; h2 = P.state0[2][src];
      P.cx[e] = P.state0[1][src]; P.cx[(long)R * H + e] = P.state0[3][src];
    }
    P.hx[(long)r * 2 * H + j] = h1; P.hx[(long)r * 2 * H + H + j] = h2;
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + j, h1);
    store_split(P.hh_hi, P.hh_lo, (long)r * 2 * H + H + j, h2);
  }
  for (int r = cta; r < R; r += G) {
    if (r < B) {
      dec_token_inputs(P, r, (int)P.tokens_in[r]);
    } else {
      for (int k = threadIdx.x; k < P.Ep; k += PK_THREADS) { P.xt_hi[(long)r * P.Ep + k] = 0.f; P.xt_lo[(long)r * P.Ep + k] = 0.f; }
      for (int j = threadIdx.x; j < H; j += PK_THREADS) {
        P.gp_hi[(long)r * H + j] = 0.f; P.gp_lo[(long)r * H + j] = 0.f;
        P.af_hi[(long)r * H + j] = 0.f; P.af_lo[(long)r * H + j] = 0.f;
      }
    }
  }
  if (P.build_euv) {
    const long n = (long)((B + P.feat_div - 1) / P.feat_div) * P.K * P.A;
    for (long e = (long)cta * PK_THREADS + threadIdx.x; e < n; e += (long)G * PK_THREADS)
      P.EUv[e] = __expf(2.f * fminf(fmaxf(__ldg(P.Uv + e), -40.f), 40.f));
  }
  gemm_prefetch(P.d, &s_sched[0], maps.m, sv, ps);
  grid_barrier(P.sync_counter, sync_target, G);
  gemm_phase(P.d, &s_sched[0], &s_sched[1], maps.m, R, sv, tmem_base, ps);           // AH, Z1h, Z2h, Z1x, Z1g
  grid_barrier(P.sync_counter, sync_target, G);
#pragma unroll 1
  for (int r = cta; r < B; r += G) dec_attention<0>(P, &maps.m[16], r, 0, sv, bulk_phase);
  dec_cell_phase<0>(P, 0, 0, cta, G);
  fence_proxy_async_smem();
  gemm_prefetch(P.d, &s_sched[1], maps.m, sv, ps);
  grid_barrier(P.sync_counter, sync_target, G);
  gemm_phase(P.d, &s_sched[1], &s_sched[2], maps.m, R, sv, tmem_base, ps);           // Z2x, Z2a
  grid_barrier(P.sync_counter, sync_target, G);
  dec_cell_phase<0>(P, 1, 0, cta, G);
  gemm_prefetch(P.d, &s_sched[2], maps.m, sv, ps);
  grid_barrier(P.sync_counter, sync_target, G);
  gemm_phase(P.d, &s_sched[2], nullptr, maps.m, R, sv, tmem_base, ps);               // logits
  grid_barrier(P.sync_counter, sync_target, G);
#pragma unroll 1
  for (int r = cta; r < B; r += G) dec_logsoftmax_row(P, r, sv);
  for (int e = cta * PK_THREADS + threadIdx.x; e < B * H; e += G * PK_THREADS) {
    const int r = e / H, j = e % H;
    P.state_out[0][e] = __ldcg(P.hx + (long)r * 2 * H + j);
    P.state_out[2][e] = __ldcg(P.hx + (long)r * 2 * H + H + j);
    P.state_out[1][e] = __ldcg(P.cx + e);
    P.state_out[3][e] = __ldcg(P.cx + (long)R * H + e);
  }
  pipeline_teardown(tmem_base);
}

// ====================================================================================
// encoder recurrence (both streams), t = 0..K-1:   z_t = XG_t + W_hh.h_{t-1}   ->  nn.LSTMCell (i,f,g,o)
// ====================================================================================
struct EncParams {
  GDesc d[2];                   // rgb, opfl recurrent products
  const PSched* sched;          // [G]
  int B, R, K, H;
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
  __shared__ PSched s_sched;
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(EncParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Psm)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  __syncthreads();
  const EncParams& P = Psm;
  for (int i = threadIdx.x; i < (int)(sizeof(PSched) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&s_sched)[i] = reinterpret_cast<const uint32_t*>(P.sched + cta)[i];
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = P.H, R = P.R, B = P.B, K = P.K;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 4) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = 0;

  for (int e = cta * PK_THREADS + threadIdx.x; e < (R - B) * 2 * H; e += G * PK_THREADS) {   // padding rows
    P.hh_hi[(long)B * 2 * H + e] = 0.f; P.hh_lo[(long)B * 2 * H + e] = 0.f;
  }
#pragma unroll 1
  for (int t = 0; t < K; ++t) {
    if (t > 0) {
      gemm_phase(P.d, &s_sched, nullptr, maps.m, R, sv, tmem_base, ps);
      grid_barrier(P.sync_counter, sync_target, G);
    }
#pragma unroll 1
    for (int e = cta * PK_THREADS + threadIdx.x; e < 2 * B * H; e += G * PK_THREADS) {
      const int s = e / (B * H), b = (e / H) % B, j = e % H;
      float* z = P.Gt[s] + ((long)t * B + b) * 4 * H + j;
      float zz[4], vs[4][PK_MAX_SLOTS];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        zz[g] = z[g * H];
        if (t > 0) zload(P.d[s], R, b, g * H + j, vs[g]);
      }
      const float m = __ldg(P.fmask + (long)b * K + t);
      const long o = ((long)t * B + b) * H + j;
      const float cp = t > 0 ? P.Cs[s][o - (long)B * H] : 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) if (t > 0) zz[g] += zadd(vs[g]);
      const float ig = sigmoid_fast(zz[0]), fg = sigmoid_fast(zz[1]), gg = tanh_fast(zz[2]), og = sigmoid_fast(zz[3]);
      const float c2 = fg * cp + ig * gg;
      const float h = og * tanh_fast(c2) * m;  // h' *= mask (sub_modules.py:139,146)
      const float c = c2 * m;                  // c' *= mask (:140,147)
      z[0] = ig; z[H] = fg; z[2 * H] = gg; z[3 * H] = og;
      P.Cs[s][o] = c;
      P.Hs[s][o] = h;
      if (t + 1 < K) store_split(P.hh_hi, P.hh_lo, (long)b * 2 * H + s * H + j, h);
    }
    if (t + 1 < K) {
      gemm_prefetch(P.d, &s_sched, maps.m, sv, ps);
      grid_barrier(P.sync_counter, sync_target, G);
    }
  }
  pipeline_teardown(tmem_base);
}

// ====================================================================================
// encoder recurrence, BACKWARD (both streams), t = K-1..0 (autograd of sub_modules.py:132-147):
//   dh_t = (dH_t + W_hh^T-product of dz_{t+1}) * m_t ;  LSTMCell backward -> dz_t (in place over the saved gates)
//   dh carried to t-1:  dz_t . W_hh   [weights: the TRANSPOSED recurrent matrices, (H, 4H) K-major]
// ====================================================================================
struct EncBwdParams {
  GDesc d[2];                   // rgb, opfl:  dhc = W_hh^T . dz
  const PSched* sched;          // [G]
  int B, R, K, H;
  float* Gt[2];                 // (K,B,4H) activated gates (i,f,g,o) -> dz (in place)
  const float* Cs[2];           // (K,B,H)
  const float* dHs[2];          // (K,B,H) direct gradient of every h_t (cross gates)
  const float* fmask;           // (B,K)
  float* dcc;                   // [2][B][H] carried dc
  float *dz_hi, *dz_lo;         // [R][8H]  [dz_rgb | dz_opfl] of the frame just processed
  unsigned int* sync_counter;
};

__global__ void __launch_bounds__(PK_THREADS, 1)
encode_bwd_persistent_kernel(const EncBwdParams* __restrict__ Pp, const __grid_constant__ MapTable maps) {
  __shared__ EncBwdParams Psm;
  __shared__ PSched s_sched;
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(EncBwdParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Psm)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  __syncthreads();
  const EncBwdParams& P = Psm;
  for (int i = threadIdx.x; i < (int)(sizeof(PSched) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&s_sched)[i] = reinterpret_cast<const uint32_t*>(P.sched + cta)[i];
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = P.H, R = P.R, B = P.B, K = P.K;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 4) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = 0;

  for (int e = cta * PK_THREADS + threadIdx.x; e < (R - B) * 8 * H; e += G * PK_THREADS) {   // padding rows
    P.dz_hi[(long)B * 8 * H + e] = 0.f; P.dz_lo[(long)B * 8 * H + e] = 0.f;
  }
#pragma unroll 1
  for (int t = K - 1; t >= 0; --t) {
#pragma unroll 1
    for (int e = cta * PK_THREADS + threadIdx.x; e < 2 * B * H; e += G * PK_THREADS) {
      const int s = e / (B * H), b = (e / H) % B, j = e % H;
      float* g4 = P.Gt[s] + ((long)t * B + b) * 4 * H + j;
      float vs[ENCB_MAX_SLOTS];
      const bool carry = t < K - 1;
      if (carry) zload(P.d[s], R, b, j, vs);
      const float gi = g4[0], gf = g4[H], gg = g4[2 * H], go = g4[3 * H];
      const float m = __ldg(P.fmask + (long)b * K + t);
      const long o = ((long)t * B + b) * H + j;
      float dh = P.dHs[s][o];
      const float ccur = P.Cs[s][o];
      const float cp = t > 0 ? P.Cs[s][o - (long)B * H] : 0.f;
      float* dccp = P.dcc + (long)s * B * H + (long)b * H + j;
      const float dcin = carry ? *dccp : 0.f;
      if (carry) dh += zadd(vs);
      dh *= m;
      const float tc = tanhf(ccur);
      const float d_o = dh * tc;
      const float dc = dcin * m + dh * go * (1.f - tc * tc);
      const float di = dc * gg, dg = dc * gi, df = dc * cp;
      *dccp = dc * gf;
      const float z0 = di * gi * (1.f - gi), z1 = df * gf * (1.f - gf), z2 = dg * (1.f - gg * gg), z3 = d_o * go * (1.f - go);
      g4[0] = z0; g4[H] = z1; g4[2 * H] = z2; g4[3 * H] = z3;
      if (t > 0) {
        const long xb = (long)b * 8 * H + (long)s * 4 * H + j;
        store_split(P.dz_hi, P.dz_lo, xb, z0); store_split(P.dz_hi, P.dz_lo, xb + H, z1);
        store_split(P.dz_hi, P.dz_lo, xb + 2 * H, z2); store_split(P.dz_hi, P.dz_lo, xb + 3 * H, z3);
      }
    }
    if (t > 0) {
      gemm_prefetch(P.d, &s_sched, maps.m, sv, ps);
      grid_barrier(P.sync_counter, sync_target, G);
      gemm_phase(P.d, &s_sched, nullptr, maps.m, R, sv, tmem_base, ps);
      grid_barrier(P.sync_counter, sync_target, G);
    }
  }
  pipeline_teardown(tmem_base);
}

// WT[c][r] = W[r][c]   (rows x cols -> cols x rows), 32x32 tiles through shared memory
__global__ void transpose_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ WT) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int r = r0 + y, c = c0 + threadIdx.x;
    tile[y][threadIdx.x] = (r < rows && c < cols) ? W[(long)r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += 8) {
    const int c = c0 + y, r = r0 + threadIdx.x;
    if (r < rows && c < cols) WT[(long)c * rows + r] = tile[threadIdx.x][y];
  }
}

// ====================================================================================
// decoder word loop, BACKWARD (hand-derived BPTT of LSTMCore_two_layer_gate, oracle/manual_bptt.py), i = T-1..0:
//   Pa  lstm_2 cell backward:  dh2 = dOUT_i + carried  ->  dz2 (in place over the saved gates)
//   Gb  dh1' += dz2.W_i2h2    dAF = dz2.W_a2h2    dh2_prev += dz2.W_h2h2           [transposed weights, K = 4H]
//   Pc  attention backward on a CTA pair per caption (dV, dUv, dAH, dw_a2w, db_a2w)  ||  lstm_1 cell backward
//   Gd  dh1_prev += dz1.W_h2h1        d[h1|h2]_prev += dAH.W_h2a
// "carried" gradients are never materialised between steps: the cell phases add the split-K slots of the
// products that feed them (plus the masked pass-through part kept in dHcar); dHcar / dC1 / dC2 hold the
// gradients of the initial state when the loop ends.
// ====================================================================================
enum { DB_GB = 0, DB_GC, DB_GD, DB_GG, DB_GH, DB_COUNT };
constexpr int DECB_MAX_SLOTS = 20;
constexpr int DECB_SPARE_SHARES = 5;

struct DecBwdParams {
  GDesc d[DB_COUNT];
  const PSched* sched;          // [2][G]
  int B, R, K, H, A, T, L;
  const float* seq_mask;        // (B, L)
  const float* dOUT;            // (T, B, H)
  float *G1s, *G2s;             // (T, B, 4H) activated gates (i,f,o,g) -> dz in place
  const float *C1s, *C2s;       // (T+1, B, H)
  const float *AHs, *ALPHAs;    // (T,B,A), (T,B,K)
  const float *Uv, *Vf, *w_a2w;
  float* DAH;                   // (T, B, A)
  float *dV, *dUv;              // (B,K,H), (B,K,A)   accumulated over the steps
  float *dwa_part, *dba_part;   // (B,A), (B)
  float *dz2_hi, *dz2_lo, *dz1_hi, *dz1_lo;   // [R][4H]
  float *dah_hi, *dah_lo;       // [R][A]
  float* dHcar;                 // (B, 2H) masked pass-through parts during the loop; final d[h1|h2]_0 at the end
  float* dC[2];                 // (B, H) carried dc of lstm_1 / lstm_2
  DropSpec drop1, drop2;
  unsigned int* sync_counter;
  long long* dbg_clock;         // diagnostics (XG_PERSIST_TRACE), or NULL
};

// lstm cell backward of `layer` at step i for the elements of this partition
__device__ __noinline__ void decb_cell_phase(const DecBwdParams& P, int layer, int i, int part, int nparts) {
  const int H = P.H, R = P.R, B = P.B, T = P.T;
  float* gs = layer == 0 ? P.G1s : P.G2s;
  const float* cs = layer == 0 ? P.C1s : P.C2s;
  float* dcx = P.dC[layer];
  float* xhi = layer == 0 ? P.dz1_hi : P.dz2_hi;
  float* xlo = layer == 0 ? P.dz1_lo : P.dz2_lo;
  const bool carry = i < T - 1;
#pragma unroll 1
  for (int e = part * PK_THREADS + threadIdx.x; e < B * H; e += nparts * PK_THREADS) {
    const int b = e / H, j = e % H;
    float va[DECB_MAX_SLOTS], vb[DECB_MAX_SLOTS], vc[DECB_MAX_SLOTS];
    // lstm_2: carried = dz2_{i+1}.W_h2h2 + dAH_{i+1}.W_h2a[:, H:]        lstm_1: dz1_{i+1}.W_h2h1 + dAH_{i+1}.W_h2a[:, :H]
    if (carry) {
      zload(P.d[layer == 0 ? DB_GG : DB_GD], R, b, j, va);
      zload(P.d[DB_GH], R, b, layer * H + j, vb);
    }
    if (layer == 0) zload(P.d[DB_GB], R, b, j, vc);            // this step's dz2.W_i2h2
    float* g4 = gs + ((long)i * B + b) * 4 * H + j;
    const float gi = g4[0], gf = g4[H], go = g4[2 * H], gg = g4[3 * H];
    const float m = __ldg(P.seq_mask + (long)b * P.L + i);
    float* dhp = P.dHcar + (long)b * 2 * H + layer * H + j;
    float dh = *dhp;
    if (layer == 1) dh += P.dOUT[((long)i * B + b) * H + j];
    const float cn = cs[((long)(i + 1) * B + b) * H + j], cp = cs[((long)i * B + b) * H + j];
    const float dcin = dcx[e];
    if (carry) dh += zadd(va) + zadd(vb);
    if (layer == 0) dh += zadd(vc);
    const float dhd = dh * (layer == 0 ? P.drop1 : P.drop2).factor((uint64_t)i * B * H + (uint64_t)e);
    const float dh_t = dhd * m;
    const float tc = tanhf(cn);
    const float d_o = dh_t * tc;
    const float dcn = dcin + dh_t * go * (1.f - tc * tc);
    const float dc_t = dcn * m;
    dcx[e] = dcn * (1.f - m) + dc_t * gf;
    const float di = dc_t * gg, dg = dc_t * gi, df = dc_t * cp;
    const float z0 = di * gi * (1.f - gi), z1 = df * gf * (1.f - gf), z2 = d_o * go * (1.f - go), z3 = dg * (1.f - gg * gg);
    g4[0] = z0; g4[H] = z1; g4[2 * H] = z2; g4[3 * H] = z3;
    *dhp = dhd * (1.f - m);
    const long xb = (long)b * 4 * H + j;
    store_split(xhi, xlo, xb, z0); store_split(xhi, xlo, xb + H, z1);
    store_split(xhi, xlo, xb + 2 * H, z2); store_split(xhi, xlo, xb + 3 * H, z3);
  }
}

// shared-window accessors (generic loads from shared memory cost three address instructions each)
__device__ __forceinline__ float pk_lds(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void pk_sts(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
// fire-and-forget accumulation (RED.ADD): every element has ONE writer per step and the steps are separated by grid
// barriers, so the sums are formed in a fixed order
__device__ __forceinline__ void pk_red_add(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// attention backward of caption r at step i, split `sp` of 2 (attention units and frames are halved).
// V[r] (K x H) and this split's half of Uv[r] (K x A/2) are fetched into the idle pipeline stages by bulk copies at
// entry (one transfer instead of ~25 dependent L2 round trips: the phase took 30 of the 62 us of a step), the
// accumulations into dV / dUv / dw_a2w / db_a2w are RED.ADDs instead of read-modify-writes.
__device__ __noinline__ void decb_attention(const DecBwdParams& P, int r, int sp, int i, const SmemView& sv, uint32_t& bulk_phase) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = P.K, A = P.A, H = P.H, B = P.B;
  const int aper = (A + 1) / 2, a0 = sp * aper, a1 = min(A, a0 + aper), na = a1 - a0;
  const uint32_t v_s = sv.stages_u32;                                   // [K][H]
  const uint32_t uv_s = sv.stages_u32 + (uint32_t)K * H * 4u;           // [K][na]
  const uint32_t al_s = smem_u32(sv.scratch);                           // K
  const uint32_t ds_s = al_s + (uint32_t)K * 4u;                        // K
  const uint32_t daf_s = ds_s + (uint32_t)K * 4u;                       // H
  if (warp == 0) {
    if (lane == 0) {
      pk_expect_tx(sv.bulk_bar, (uint32_t)K * H * 4u);
      bulk_g2s(v_s, P.Vf + (long)r * K * H, (uint32_t)K * H * 4u, sv.bulk_bar);
      pk_expect_tx(sv.bulk_bar + 8, (uint32_t)K * na * 4u);
    }
    __syncwarp();
    for (int k = lane; k < K; k += 32)
      bulk_g2s(uv_s + (uint32_t)(k * na) * 4u, P.Uv + ((long)r * K + k) * A + a0, (uint32_t)na * 4u, sv.bulk_bar + 8);
  }
  for (int k = threadIdx.x; k < K; k += PK_THREADS) pk_sts(al_s + k * 4u, P.ALPHAs[((long)i * B + r) * K + k]);
  {   // dAF[r, :] = sum of the split-K slots of dz2.W_a2h2: float4 columns, every slot load of a thread in flight at once
    const GDesc& dc = P.d[DB_GC];
    const long sstr4 = (long)P.R * (dc.n_rows >> 2);
#pragma unroll 1
    for (int j4 = threadIdx.x; j4 < (H >> 2); j4 += PK_THREADS) {
      const float4* src = reinterpret_cast<const float4*>(dc.out + (long)r * dc.n_rows) + j4;
      float4 v[DECB_MAX_SLOTS];
#pragma unroll
      for (int k = 0; k < DECB_MAX_SLOTS; ++k) v[k] = k < dc.ns ? __ldcg(src + k * sstr4) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < DECB_MAX_SLOTS; ++k) { sum.x += v[k].x; sum.y += v[k].y; sum.z += v[k].z; sum.w += v[k].w; }
      pk_sts(daf_s + (j4 * 4) * 4u, sum.x); pk_sts(daf_s + (j4 * 4 + 1) * 4u, sum.y);
      pk_sts(daf_s + (j4 * 4 + 2) * 4u, sum.z); pk_sts(daf_s + (j4 * 4 + 3) * 4u, sum.w);
    }
  }
  __syncthreads();
  pk_wait(sv.bulk_bar, bulk_phase & 1);             // V[r] landed
#pragma unroll 1
  for (int k = warp; k < K; k += PK_WARPS) {        // ds[k] = dAF . V[r, k, :]
    float p = 0.f;
#pragma unroll 4
    for (int j = lane; j < H; j += 32) p = fmaf(pk_lds(daf_s + j * 4u), pk_lds(v_s + (uint32_t)(k * H + j) * 4u), p);
    p = warp_sum(p);
    if (lane == 0) pk_sts(ds_s + k * 4u, p);
  }
  {   // dV[r, k, :] += alpha_k * dAF[r, :]  for this split's frames
    const int kper = (K + 1) / 2, kb0 = sp * kper, kb1 = min(K, kb0 + kper);
    const int n = max(0, kb1 - kb0) * H;
    float* dv = P.dV + ((long)r * K + kb0) * H;
#pragma unroll 4
    for (int e = threadIdx.x; e < n; e += PK_THREADS) {
      const int kk = e / H, j = e - kk * H;
      pk_red_add(dv + e, pk_lds(al_s + (kb0 + kk) * 4u) * pk_lds(daf_s + j * 4u));
    }
  }
  __syncthreads();
  float dot = 0.f;
  for (int k = 0; k < K; ++k) dot += pk_lds(al_s + k * 4u) * pk_lds(ds_s + k * 4u);   // every thread computes the same fixed-order sum
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += PK_THREADS) pk_sts(ds_s + k * 4u, pk_lds(al_s + k * 4u) * (pk_lds(ds_s + k * 4u) - dot));
  __syncthreads();
  pk_wait(sv.bulk_bar + 8, bulk_phase & 1);         // this split's half of Uv[r] landed
  bulk_phase++;
#pragma unroll 1
  for (int al = threadIdx.x; al < na; al += PK_THREADS) {
    const int a = a0 + al;
    const float w = __ldg(P.w_a2w + a), h0 = P.AHs[((long)i * B + r) * A + a];
    float acc_ah = 0.f, acc_wa = 0.f;
    float* du = P.dUv + (long)r * K * A + a;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float th = tanh_fast(h0 + pk_lds(uv_s + (uint32_t)(k * na + al) * 4u));
      const float dsk = pk_lds(ds_s + k * 4u);
      const float dp = dsk * w * (1.f - th * th);
      pk_red_add(du + (long)k * A, dp);
      acc_ah += dp;
      acc_wa = fmaf(dsk, th, acc_wa);
    }
    P.DAH[((long)i * B + r) * A + a] = acc_ah;
    store_split(P.dah_hi, P.dah_lo, (long)r * A + a, acc_ah);
    pk_red_add(P.dwa_part + (long)r * A + a, acc_wa);
  }
  if (threadIdx.x == 0 && sp == 0) {
    float sdb = 0.f;
    for (int k = 0; k < K; ++k) sdb += pk_lds(ds_s + k * 4u);
    pk_red_add(P.dba_part + r, sdb);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(PK_THREADS, 1)
decode_bwd_persistent_kernel(const DecBwdParams* __restrict__ Pp, const __grid_constant__ MapTable maps) {
  __shared__ DecBwdParams Psm;
  __shared__ PSched s_sched[2];
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(DecBwdParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Psm)[i] = reinterpret_cast<const uint32_t*>(Pp)[i];
  __syncthreads();
  const DecBwdParams& P = Psm;
  for (int i = threadIdx.x; i < (int)(2 * sizeof(PSched) / 4); i += PK_THREADS) {
    const int ph = i / (int)(sizeof(PSched) / 4), w = i % (int)(sizeof(PSched) / 4);
    reinterpret_cast<uint32_t*>(&s_sched[ph])[w] = reinterpret_cast<const uint32_t*>(P.sched + (long)ph * G + cta)[w];
  }
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = P.H, R = P.R, B = P.B, T = P.T;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 11) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = 0;
  uint32_t bulk_phase = 0;

#pragma unroll 1
  for (int i = T - 1; i >= 0; --i) {
    const int ts = T - 1 - i;      // (stamps: step 3 of the loop)
    pk_stamp(P.dbg_clock, cta, ts, 0);
    decb_cell_phase(P, 1, i, cta, G);                                            // Pa
    gemm_prefetch(P.d, &s_sched[0], maps.m, sv, ps);
    pk_stamp(P.dbg_clock, cta, ts, 1);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, ts, 2);
    gemm_phase(P.d, &s_sched[0], &s_sched[1], maps.m, R, sv, tmem_base, ps);     // Gb
    pk_stamp(P.dbg_clock, cta, ts, 3);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, ts, 4);
    if (G >= 2 * B + 8) {          // Pc, enough SMs: caption halves and cell elements on disjoint CTAs
      // the cell elements are dealt in shares: one to every attention CTA (after its caption half), DECB_SPARE_SHARES to every
      // spare CTA (with all of them on the G - 2B spare CTAs those finished 9 us after the attention)
      const int nparts = 2 * B + (G - 2 * B) * DECB_SPARE_SHARES;
      if (cta < 2 * B) {
        decb_attention(P, cta >> 1, cta & 1, i, sv, bulk_phase);
        decb_cell_phase(P, 0, i, cta, nparts);
      } else {
#pragma unroll 1
        for (int w = 0; w < DECB_SPARE_SHARES; ++w) decb_cell_phase(P, 0, i, 2 * B + (cta - 2 * B) * DECB_SPARE_SHARES + w, nparts);
      }
    } else {                       // large batches: every CTA walks its caption halves, then its cell elements
#pragma unroll 1
      for (int u = cta; u < 2 * B; u += G) decb_attention(P, u >> 1, u & 1, i, sv, bulk_phase);
      decb_cell_phase(P, 0, i, cta, G);
    }
    fence_proxy_async_smem();      // the stages were read through the generic proxy: order before the TMA refill
    gemm_prefetch(P.d, &s_sched[1], maps.m, sv, ps);
    pk_stamp(P.dbg_clock, cta, ts, 5);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, ts, 6);
    gemm_phase(P.d, &s_sched[1], &s_sched[0], maps.m, R, sv, tmem_base, ps);     // Gd
    pk_stamp(P.dbg_clock, cta, ts, 7);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, ts, 8);
  }
  // gradients of the initial state: pass-through parts + the products of step 0
#pragma unroll 1
  for (int e = cta * PK_THREADS + threadIdx.x; e < B * 2 * H; e += G * PK_THREADS) {
    const int b = e / (2 * H), c = e % (2 * H), layer = c / H, j = c % H;
    float va[DECB_MAX_SLOTS], vb[DECB_MAX_SLOTS];
    zload(P.d[layer == 0 ? DB_GG : DB_GD], R, b, j, va);
    zload(P.d[DB_GH], R, b, c, vb);
    P.dHcar[e] += zadd(va) + zadd(vb);
  }
  pipeline_teardown(tmem_base);
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct PhaseSchedule {
  std::vector<PSched> per_cta;                         // [G]
  std::vector<int> ns;                                 // per desc (phase-local order): max slots of a strip
  bool ok = true;
};

// lay the k-blocks of all strips (desc, row tile, column block) end to end; CTA c gets units [cU/G, (c+1)U/G)
static PhaseSchedule build_phase(const std::vector<int>& desc_ids, const GDesc* descs, int ncb, int G, int max_slots) {
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
  // A strip of nkb k-blocks cut into runs of at least q k-blocks collects at most ceil(nkb / q) + 1 slots.  When a
  // phase holds less work than CTAs x q (small vocabularies, tiny batches: SURVEY config 1) the runs are kept at q
  // and the trailing CTAs idle, instead of one k-block per CTA and more partial tiles per strip than consumers add.
  int max_nkb = 1;
  for (const Strip& s : strips) max_nkb = std::max(max_nkb, s.nkb);
  const long q_min = max_slots > 1 ? (max_nkb + max_slots - 2) / (max_slots - 1) : max_nkb;
  size_t si = 0; int off = 0;
  for (int c = 0; c < G; ++c) {
    long need = std::max<long>((long)(c + 1) * U / G - (long)c * U / G, q_min);
    PSched& sc = ph.per_cta[c];
    while (need > 0 && si < strips.size()) {
      Strip& s = strips[si];
      const int take = (int)std::min<long>(need, s.nkb - off);
      if (sc.n >= PK_MAX_ITEMS || s.slots >= max_slots) { ph.ok = false; return ph; }
      sc.it[sc.n++] = PItem{(short)s.desc, (short)s.slots, (short)s.rt, (short)s.cb, (short)off, (short)take};
      sc.tot_kb += (short)take;
      sc.tot_chunks += (short)((take + PK_CHUNK - 1) / PK_CHUNK);
      s.slots++;
      off += take; need -= take;
      if (off == s.nkb) { ++si; off = 0; }
    }
  }
  ph.ns.assign(desc_ids.size(), 0);
  size_t k = 0;
  for (size_t i = 0; i < desc_ids.size(); ++i) {
    const int rts = (descs[desc_ids[i]].n_rows + 127) / 128;
    for (size_t q = 0; q < (size_t)rts * ncb; ++q) ph.ns[i] = std::max(ph.ns[i], strips[k++].slots);
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
  int sched_mode = -1;           // schedule the slot buffers were last written under (0 decode, 1 training, 2 single step)
  bool attr_set = false;
  // single-step mode: everything but the per-step pointers is reused while the key below holds
  bool step_valid = false;
  DecParams step_hp_dev;         // the parameter block as it lies on the device (later steps upload only when it differs)
  unsigned int step_sync_base = 0;
  int step_B = 0, step_fdiv = 0;
  const float *step_V = nullptr, *step_Uv = nullptr, *step_pos = nullptr;
  unsigned long long step_epoch = ~0ull;
  MapTable step_mt;
  // decoder backward
  int dbB = 0, dbK = 0;
  char* dbpool = nullptr;
  size_t dbpool_bytes = 0;
  DecBwdParams dbp;
  DecBwdParams* d_dbparams = nullptr;
  unsigned int* d_dbcounter = nullptr;
  long long* d_dbdbg = nullptr;
  float* dbwT[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  unsigned long long dbwT_epoch = ~0ull, wT_epoch = ~0ull;
  bool dbattr_set = false;
  // encoder backward
  int bB = 0;
  char* bpool = nullptr;
  size_t bpool_bytes = 0;
  EncBwdParams bp;
  EncBwdParams* d_bparams = nullptr;
  unsigned int* d_bcounter = nullptr;
  float* wT[2] = {nullptr, nullptr};
  bool battr_set = false;
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
  if (s->bpool) cudaFree(s->bpool);
  if (s->dbpool) cudaFree(s->dbpool);
  delete s;
  persist_state(ctx) = nullptr;
}

static inline int env_flag(const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; }

constexpr int PK_MAX_ROWS = 1024;    // caption rows (batch, or videos x beam) of one persistent launch: 16 column blocks

// A serial loop stays off its persistent kernel (shape outside the kernel's limits, or a mode it does not cover): the
// caller runs the per-step launches.  On a strict handle (xg_set_strict / XG_STRICT_PERSIST=1) that is an error instead
// of a silent slow path; with the persistent engine switched off by xg_set_engine it is what the caller asked for.
static int persist_refuse(xg_context* ctx, const char* loop, const char* why) {
  ctx->n_unfused++;
  if (ctx->strict_persist && ctx->persist_mode) {
    ctx->es.msg = std::string("strict mode: ") + loop + " cannot run on its persistent kernel (" + why + ")";
    return XG_ERR_UNSUPPORTED;
  }
  return PK_FALLBACK;
}

static bool persist_eligible(const xg_context* ctx, int B, int K, int max_rows = PK_MAX_ROWS) {
  const xg_dims& d = ctx->d;
  const bool v_behind = (long)K * (d.att + d.rnn) * 4 <= (long)PK_STAGES * PK_STAGE_BYTES;
  const bool v_over = d.rnn <= d.att;            // V[r] fits over the consumed exp(2Uv) chunks
  // bulk copies of exp(2Uv) need 16-byte rows; the V panels behind / over them must start on 128-byte boundaries
  const bool aligned = d.att % 4 == 0 && ((long)K * d.att * 4) % 128 == 0 && ((long)K * d.rnn * 2) % 128 == 0;
  return ctx->persist_mode && d.rnn % 32 == 0 && d.rnn <= 512 && d.embed <= DEC_TI * PK_THREADS && d.embed % 4 == 0 &&
         aligned && B <= max_rows && K >= 1 && K <= PK_BULK_CHUNKS * DEC_FPC && ctx->sm_count >= 16 && ctx->sm_count <= 256 &&
         (PK_WARPS + 1) * K + 8 <= PK_SCRATCH_FLOATS && d.att <= DEC_NA * PK_THREADS &&
         (long)K * d.att * 4 <= (long)PK_STAGES * PK_STAGE_BYTES && (v_behind || v_over) &&
         (long)d.vocab * 4 <= (long)PK_STAGES * PK_STAGE_BYTES && d.vocab >= 2 && d.vocab < 32000 && d.att < 32000;
}

// schedule of one kernel (phases laid one after the other, [phase][G]); false if a phase cannot be scheduled
static bool persist_plan(const std::vector<std::vector<int>>& phases, GDesc* descs, int ncb, int G, std::vector<PSched>& sched,
                         int max_slots = PK_MAX_SLOTS) {
  sched.clear();
  for (const auto& ids : phases) {
    PhaseSchedule ph = build_phase(ids, descs, ncb, G, max_slots);
    if (!ph.ok) return false;
    sched.insert(sched.end(), ph.per_cta.begin(), ph.per_cta.end());
    for (size_t i = 0; i < ids.size(); ++i) descs[ids[i]].ns = ph.ns[i];
  }
  return true;
}

static int gemm_run(xg_context* ctx, const GemmP& p, cudaStream_t st);   // xg_forward.cuh

// teacher-forced training forward: the step-major saved buffers of TrainSaved (xg_context.cuh)
struct PersistTrainIO {
  const float* seq_mask; int L;
  float *G1, *G2, *C1, *C2, *H12, *AH, *ALPHA, *AF;
  DropSpec drop1, drop2;
};

// single word step on arbitrary state rows (beam search): B rows, every feat_div of them share V / Uv / pos
struct PersistStepIO {
  const int64_t* tokens;   // (B)
  float* const* state;     // h1,c1,h2,c2 (B,H): read, then overwritten with the new state
  float* logp;             // (B,V) out, or NULL
  int feat_div;
  int first;               // first step of a call: (re)build exp(2 Uv)
  const int* parent;       // (B) or NULL: row r starts from state row parent[r] (read before any row is written)
  float* ys; int* ix;      // (B,topk) top-k of every row with the UNK penalty, or NULL
  int topk;
};

// mode 0 (tr == step == nullptr): greedy decoding, T = seq_length.  mode 1 (tr): teacher-forced forward over T = L'
// steps.  mode 2 (step): one word step, states and log-probs out (asynchronous).
static int persist_decode(xg_context* ctx, const float* Vf, const float* Uv, const float* pos, const float* const* state0,
                          int B, int K, int T, int64_t* seq_out, float* logp_out, int* steps_out, const PersistTrainIO* tr,
                          cudaStream_t st, const PersistStepIO* step = nullptr) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, A = d.att, V = d.vocab;
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN, Ep = (E + 31) / 32 * 32, G = ctx->sm_count;
  const char* loop_name = tr ? "the teacher-forced word loop" : (step ? "the beam-search word step" : "the greedy word loop");
  if (!persist_eligible(ctx, B, K)) return persist_refuse(ctx, loop_name, "dimensions outside the persistent decoder's limits, see xgating.h");
  if (T > 2048) return persist_refuse(ctx, loop_name, "more than 2048 word steps");
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  PersistState*& S = persist_state(ctx);
  if (!S) S = new PersistState();
  DecParams& hp = S->hp;
  const int kbH = H / 32, kbE = Ep / 32;
  auto step_io = [&]() {
    hp.build_euv = step->first;
    hp.tokens_in = step->tokens; hp.logp_out = step->logp;
    hp.parent_in = step->parent; hp.ys_out = step->ys; hp.ix_out = step->ix; hp.topk = step->topk;
    for (int q = 0; q < 4; ++q) { hp.state0[q] = step->state[q]; hp.state_out[q] = step->state[q]; }
  };
  if (step && S->step_valid && S->sched_mode == 2 && S->R == R && S->K == K && S->step_B == B && S->step_fdiv == step->feat_div &&
      S->step_V == Vf && S->step_Uv == Uv && S->step_pos == pos && S->step_epoch == ctx->param_epoch && !step->first) {
    // later steps of the same beam search: plans, tensor maps and the device schedule are those of the last launch
    // From the third step of a search on nothing in the block changes (the fused step gathers states in place), and
    // the barrier counter simply runs on: a step is then ONE stream operation instead of copy + memset + launch.
    step_io();
    if (memcmp(&hp, &S->step_hp_dev, sizeof(DecParams)) != 0) {
      XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_params, &hp, sizeof(DecParams), cudaMemcpyHostToDevice, st));
      memcpy(&S->step_hp_dev, &hp, sizeof(DecParams));
    }
    ProfScope ps(ctx, "decode_step_persistent", st);
    const DecParams* dp = S->d_params;
    unsigned int base = S->step_sync_base;
    void* args[3] = {(void*)&dp, (void*)&S->step_mt, (void*)&base};
    XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)decode_step_persistent_kernel, dim3(G), dim3(PK_THREADS), args, PK_SMEM_BYTES, st));
    ctx->n_fused++;
    S->step_sync_base += PK_STEP_BARRIERS * (unsigned int)G;
    return XG_OK;
  }
  S->step_valid = false;

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
  // every mode's schedule is planned: the slot buffers are shared and sized by the largest slot count
  const std::vector<std::vector<int>> phases_dec = {{DD_Z1X, DD_Z1G}, {DD_Z2X, DD_Z2A}, {DD_LOGIT, DD_AH, DD_Z1H, DD_Z2H}};
  const std::vector<std::vector<int>> phases_trn = {{DD_AH, DD_Z1H, DD_Z2H}, {DD_Z2X, DD_Z2A}, {}};
  const std::vector<std::vector<int>> phases_stp = {{DD_AH, DD_Z1H, DD_Z2H, DD_Z1X, DD_Z1G}, {DD_Z2X, DD_Z2A}, {DD_LOGIT}};
  const int mode = tr ? 1 : (step ? 2 : 0);
  const std::vector<std::vector<int>>* plans[3] = {&phases_dec, &phases_trn, &phases_stp};
  std::vector<PSched> sched;
  int ns_cap[DD_COUNT];
  for (int i = 0; i < DD_COUNT; ++i) ns_cap[i] = 0;
  for (int pm = 0; pm < 3; ++pm) {
    const int m = (mode + 1 + pm) % 3;            // the mode that runs is planned last (its slot counts stay in hp.d)
    for (int i = 0; i < DD_COUNT; ++i) hp.d[i].ns = 0;
    if (!persist_plan(*plans[m], hp.d, R / PK_BN, G, sched)) return persist_refuse(ctx, loop_name, "no work schedule for this shape");
    for (int i = 0; i < DD_COUNT; ++i) ns_cap[i] = std::max(ns_cap[i], hp.d[i].ns);
  }

  // ---- device pool ----
  if (S->R != R || S->K != K) {
    if (S->pool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->pool); S->pool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->pool, pass == 0 ? 0 : S->pool_bytes);
      S->d_params = a.take<DecParams>(1);
      S->d_counter = a.take<unsigned int>(32 * 258 + 256);
      S->d_flags = a.take<int>(2048);
      S->d_dbg = a.take<long long>((2048 + 256) * PK_STAMPS + 64);
      hp.sched = a.take<PSched>(sched.size());
      for (int i = 0; i < DD_COUNT; ++i) hp.d[i].out = a.take<float>((size_t)ns_cap[i] * R * hp.d[i].n_rows);
      hp.xt_hi = a.take<float>((long)R * Ep); hp.xt_lo = a.take<float>((long)R * Ep);
      hp.hh_hi = a.take<float>((long)R * 2 * H); hp.hh_lo = a.take<float>((long)R * 2 * H);
      hp.gp_hi = a.take<float>((long)R * H); hp.gp_lo = a.take<float>((long)R * H);
      hp.af_hi = a.take<float>((long)R * H); hp.af_lo = a.take<float>((long)R * H);
      hp.hx = a.take<float>((long)R * 2 * H);
      hp.cx = a.take<float>((long)2 * R * H);
      hp.unfinished = a.take<float>(R);
      hp.tok = a.take<int64_t>(R);
      S->tgate = a.take<float>((long)V * H);
      hp.EUv = a.take<float>((long)R * K * A);
      if (pass == 0) {
        S->pool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->pool, S->pool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->pool, 0, S->pool_bytes, st));
      }
    }
    S->R = R; S->K = K;
    S->tgate_epoch = ~0ull;
    S->sched_mode = -1;
  }
  // consumers add every slot up to the per-product maximum and rely on never-written slots being zero: that
  // holds per schedule, so the slot buffers are cleared when the schedule changes (decode <-> training)
  if (S->sched_mode != mode) {
    if (S->sched_mode != -1) {
      char* lo = reinterpret_cast<char*>(hp.d[0].out);
      char* hi = reinterpret_cast<char*>(hp.d[DD_COUNT - 1].out) + sizeof(float) * (size_t)ns_cap[DD_COUNT - 1] * R * hp.d[DD_COUNT - 1].n_rows;
      XG_CUDA_TRY(ctx->es, cudaMemsetAsync(lo, 0, (size_t)(hi - lo), st));
    }
    S->sched_mode = mode;
  }
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<PSched*>(hp.sched), sched.data(), sizeof(PSched) * sched.size(),
                                       cudaMemcpyHostToDevice, st));

  // ---- POS-gate table of every token: tgate = relu(embed . W_gate^T + b)  (sub_modules.py:29-32 applied to
  //      SAModel.py:198's embedding rows); rebuilt whenever the bound parameters change ----
  if (mode != 1 && S->tgate_epoch != ctx->param_epoch) {
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
  {   // V as [B*K][H]: one box = (H/2 columns) x (K frames) of a caption, dense in shared memory
    const int fdiv = step ? step->feat_div : 1;
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)((B + fdiv - 1) / fdiv) * K};
    cuuint64_t strides[1] = {(cuuint64_t)H * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)(H / 2), (cuuint32_t)K};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult cr = ts->encode(&maps[16], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Vf), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { ctx->es.set(__FILE__, __LINE__, "cuTensorMapEncodeTiled (V) failed", nullptr); return XG_ERR_CUDA; }
    maps[17] = maps[16];
  }

  hp.B = B; hp.R = R; hp.K = K; hp.H = H; hp.E = E; hp.Ep = Ep; hp.A = A; hp.V = V; hp.T = T;
  hp.b_h2a = ctx->P[XG_P_H2A_B]; hp.w_a2w = ctx->P[XG_P_A2W_W]; hp.b_a2w = ctx->P[XG_P_A2W_B];
  hp.bias[0][0] = ctx->P[XG_P_L1_I2H_B]; hp.bias[0][1] = ctx->P[XG_P_L1_A2H_B]; hp.bias[0][2] = ctx->P[XG_P_L1_H2H_B];
  hp.bias[1][0] = ctx->P[XG_P_L2_I2H_B]; hp.bias[1][1] = ctx->P[XG_P_L2_A2H_B]; hp.bias[1][2] = ctx->P[XG_P_L2_H2H_B];
  hp.b_logit = ctx->P[XG_P_LOGIT_B]; hp.embed = ctx->P[XG_P_EMBED_W];
  hp.tgate = S->tgate;
  hp.Vf = Vf; hp.Uv = Uv; hp.pos = pos;
  for (int q = 0; q < 4; ++q) hp.state0[q] = state0 ? state0[q] : nullptr;
  hp.mode = tr ? 1 : 0;
  hp.feat_div = step ? step->feat_div : 1;
  hp.build_euv = 1;
  if (step) step_io();
  if (tr) {
    hp.L = tr->L; hp.seq_mask = tr->seq_mask;
    hp.G1s = tr->G1; hp.G2s = tr->G2; hp.C1s = tr->C1; hp.C2s = tr->C2; hp.H12s = tr->H12;
    hp.AHs = tr->AH; hp.ALPHAs = tr->ALPHA; hp.AFs = tr->AF;
    hp.drop1 = tr->drop1; hp.drop2 = tr->drop2;
  }
  hp.seq = seq_out; hp.seqlogp = logp_out; hp.flags = S->d_flags;
  hp.sync_counter = S->d_counter;
  hp.dbg_clock = env_flag("XG_PERSIST_TRACE") ? S->d_dbg : nullptr;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_params, &hp, sizeof(DecParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_counter, 0, sizeof(unsigned int) * (32 * 258 + 256), st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_flags, 0, sizeof(int) * (size_t)T, st));
  if (mode == 0) {
    XG_CUDA_TRY(ctx->es, cudaMemsetAsync(seq_out, 0, sizeof(int64_t) * (size_t)B * T, st));
    XG_CUDA_TRY(ctx->es, cudaMemsetAsync(logp_out, 0, sizeof(float) * (size_t)B * T, st));
  }
  if (hp.dbg_clock) XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_dbg, 0, sizeof(long long) * ((2048 + 256) * PK_STAMPS + 64), st));

  if (!S->attr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_persistent_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_persistent_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    int nb = 0, nb1 = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_persistent_kernel<0>, PK_THREADS, PK_SMEM_BYTES));
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb1, decode_persistent_kernel<1>, PK_THREADS, PK_SMEM_BYTES));
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_step_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1 && nb1 >= 1, XG_ERR_CUDA, "persistent decoder does not fit on an SM");
    S->attr_set = true;
  }
  {
    ProfScope ps(ctx, mode == 1 ? "train_decode_persistent" : (mode == 2 ? "decode_step_persistent" : "decode_persistent"), st);
    const DecParams* dp = S->d_params;
    unsigned int base = 0;          // the counter was cleared above
    void* args[3] = {(void*)&dp, (void*)&mt, (void*)&base};      // the third argument exists in single-step mode only
    void* fn = mode == 1 ? (void*)decode_persistent_kernel<1> : (mode == 2 ? (void*)decode_step_persistent_kernel : (void*)decode_persistent_kernel<0>);
    XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(PK_THREADS), args, PK_SMEM_BYTES, st));
    ctx->n_fused++;
  }
  if (mode == 2) {
    memcpy(&S->step_hp_dev, &hp, sizeof(DecParams));
    S->step_sync_base = PK_STEP_BARRIERS * (unsigned int)G;
    S->step_valid = true; S->step_B = B; S->step_fdiv = step->feat_div; S->step_V = Vf; S->step_Uv = Uv; S->step_pos = pos;
    S->step_epoch = ctx->param_epoch; S->step_mt = mt;
  }
  if (mode != 0 || !steps_out) return XG_OK;     // asynchronous: the caller's next kernels follow on the same stream
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(ctx->h_pinned, S->d_flags, sizeof(int) * (size_t)T, cudaMemcpyDeviceToHost, st));
  XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
  int steps = 0;
  while (steps < T && ctx->h_pinned[steps] != 0) ++steps;
  *steps_out = steps;
  if (hp.dbg_clock) {   // XG_PERSIST_TRACE=1: average SM cycles per phase (CTA 0), printed to stderr
    std::vector<long long> h((size_t)T * PK_STAMPS);
    cudaMemcpy(h.data(), S->d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    const char* names[6] = {"G1 (z1x,z1g)", "P1 (cell1)", "G3 (z2x,z2a)", "P3 (cell2)", "G4 (logits,ah,z1h,z2h)",
                            "P4 (pick || attention t+1)"};
    double tot = 0;
    const int n = steps > 1 ? steps - 1 : 1;
    for (int i = 0; i < 6; ++i) {
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
#ifdef PK_FINE_TRACE
    {
      long long f[64];
      cudaMemcpy(f, S->d_dbg + (2048 + 256) * PK_STAMPS, sizeof(f), cudaMemcpyDeviceToHost);
      for (int who = 0; who < 2; ++who) {
        fprintf(stderr, "[xg persist trace] step 3 %s stamps (cycles after phase entry):", who == 0 ? "pick (cta 0)" : "attention (cta B)");
        for (int i = 0; i < 31; ++i) if (f[who * 32 + i]) fprintf(stderr, " [%d] %lld", i, f[who * 32 + i] - f[who * 32 + 31]);
        fprintf(stderr, "\n");
      }
    }
#endif
    if (steps > 3 && G <= 256) {   // step 3, all CTAs: when does each CTA finish its share of a phase (ns after the phase opened)?
      std::vector<long long> ga((size_t)G * PK_STAMPS);
      cudaMemcpy(ga.data(), S->d_dbg + 2048 * PK_STAMPS, sizeof(long long) * ga.size(), cudaMemcpyDeviceToHost);
      for (int i = 0; i < 6; ++i) {
        long long open = 0;
        for (int c = 0; c < G; ++c) open = std::max(open, ga[(size_t)c * PK_STAMPS + 2 * i]);   // last CTA through the previous barrier
        std::vector<long long> fin(G);
        for (int c = 0; c < G; ++c) fin[c] = ga[(size_t)c * PK_STAMPS + 2 * i + 1] - open;
        std::vector<long long> srt = fin;
        std::sort(srt.begin(), srt.end());
        long long close = 0;
        for (int c = 0; c < G; ++c) close = std::max(close, ga[(size_t)c * PK_STAMPS + 2 * i + 2]);
        int worst = (int)(std::max_element(fin.begin(), fin.end()) - fin.begin());
        fprintf(stderr, "[xg persist trace] step 3 %-26s work done after: min %6lld  median %6lld  p90 %6lld  max %6lld ns (cta %d)   barrier exit %6lld ns\n",
                names[i], srt[0], srt[G / 2], srt[G * 9 / 10], srt[G - 1], worst, close - open);
      }
    }
  }
  return XG_OK;
}

// frame recurrence of both encoder streams; eb.G holds the hoisted input projections (+ both biases)
static int persist_encode(xg_context* ctx, const float* fmask, int B, int K, EncBufs& eb, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, G = ctx->sm_count;
  if (K < 2) return PK_FALLBACK;          // a single frame has no recurrence to fuse
  if (!ctx->persist_mode || H % 32 != 0 || B < 1 || H > 4096 || G > 256)
    return persist_refuse(ctx, "the encoder frame recurrence", "rnn_size must be a multiple of 32, at most 4096");
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN;
  if (R > PK_MAX_ROWS) return persist_refuse(ctx, "the encoder frame recurrence", "more than 1024 captions per launch");
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
  if (!persist_plan({{0, 1}}, ep.d, R / PK_BN, G, sched)) return persist_refuse(ctx, "the encoder frame recurrence", "no work schedule for this shape");
  if (S->eB != B) {
    if (S->epool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->epool); S->epool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->epool, pass == 0 ? 0 : S->epool_bytes);
      S->d_eparams = a.take<EncParams>(1);
      S->d_ecounter = a.take<unsigned int>(32 * 258);
      ep.sched = a.take<PSched>(sched.size());
      for (int s = 0; s < 2; ++s) ep.d[s].out = a.take<float>((size_t)ep.d[s].ns * R * 4 * H);
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
  MapTable mt;
  XG_TRY(tc_make_map(ctx, ts, ctx->P[XG_P_LSTM_RGB_WHH], 4 * H, H, 128, &mt.m[0]));
  XG_TRY(tc_make_map(ctx, ts, ctx->P[XG_P_LSTM_OPFL_WHH], 4 * H, H, 128, &mt.m[1]));
  XG_TRY(tc_make_map(ctx, ts, ep.hh_hi, R, 2 * H, PK_BN, &mt.m[2]));
  XG_TRY(tc_make_map(ctx, ts, ep.hh_lo, R, 2 * H, PK_BN, &mt.m[3]));
  for (int i = 4; i < 18; ++i) mt.m[i] = mt.m[0];
  ep.B = B; ep.R = R; ep.K = K; ep.H = H;
  for (int s = 0; s < 2; ++s) { ep.Gt[s] = eb.G[s]; ep.Hs[s] = eb.Hs[s]; ep.Cs[s] = eb.Cs[s]; }
  ep.fmask = fmask;
  ep.sync_counter = S->d_ecounter;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_eparams, &ep, sizeof(EncParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_ecounter, 0, sizeof(unsigned int) * 32 * 258, st));
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
  ctx->n_fused++;
  return XG_OK;
}

// backward frame recurrence of both encoder streams: eb.G holds the activated gates (-> dz in place), dH the direct
// gradients of every h_t; dz of all frames is left in eb.G for the batched weight gradients that follow
static int persist_encode_bwd(xg_context* ctx, const float* fmask, int B, int K, const EncBufs& eb, float* const* dH, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, G = ctx->sm_count;
  if (K < 2) return PK_FALLBACK;
  if (!ctx->persist_mode || H % 32 != 0 || B < 1 || H > 4096 || G > 256)
    return persist_refuse(ctx, "the encoder backward recurrence", "rnn_size must be a multiple of 32, at most 4096");
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN;
  if (R > PK_MAX_ROWS) return persist_refuse(ctx, "the encoder backward recurrence", "more than 1024 captions per launch");
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  PersistState*& S = persist_state(ctx);
  if (!S) S = new PersistState();
  EncBwdParams& bp = S->bp;
  const int kb4H = 4 * H / 32;
  for (int s = 0; s < 2; ++s) {
    GDesc& g = bp.d[s];
    g.w_map = s; g.x_hi = 2; g.x_lo = 3; g.xkb0 = s * kb4H; g.n_rows = H; g.nkb = kb4H;
  }
  std::vector<PSched> sched;
  if (!persist_plan({{0, 1}}, bp.d, R / PK_BN, G, sched, ENCB_MAX_SLOTS))
    return persist_refuse(ctx, "the encoder backward recurrence", "no work schedule for this shape");
  if (S->bB != B) {
    if (S->bpool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->bpool); S->bpool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->bpool, pass == 0 ? 0 : S->bpool_bytes);
      S->d_bparams = a.take<EncBwdParams>(1);
      S->d_bcounter = a.take<unsigned int>(64);
      bp.sched = a.take<PSched>(sched.size());
      for (int s = 0; s < 2; ++s) bp.d[s].out = a.take<float>((size_t)bp.d[s].ns * R * H);
      bp.dz_hi = a.take<float>((long)R * 8 * H); bp.dz_lo = a.take<float>((long)R * 8 * H);
      bp.dcc = a.take<float>((long)2 * B * H);
      S->wT[0] = a.take<float>((long)4 * H * H); S->wT[1] = a.take<float>((long)4 * H * H);
      if (pass == 0) {
        S->bpool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->bpool, S->bpool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->bpool, 0, S->bpool_bytes, st));
      }
    }
    S->bB = B;
    S->wT_epoch = ~0ull;
  }
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<PSched*>(bp.sched), sched.data(), sizeof(PSched) * sched.size(),
                                       cudaMemcpyHostToDevice, st));
  const int whh[2] = {XG_P_LSTM_RGB_WHH, XG_P_LSTM_OPFL_WHH};
  // transposed copies of the recurrent matrices (2 x 4 MB): derived tables like the operand splits, rebuilt when the
  // parameter epoch moved (every optimizer step in training)
  for (int s = 0; s < 2 && S->wT_epoch != ctx->param_epoch; ++s) {
    ProfScope ps(ctx, "transpose", st);
    transpose_kernel<<<dim3(ceil_div(H, 32), ceil_div(4 * H, 32)), dim3(32, 8), 0, st>>>(ctx->P[whh[s]], 4 * H, H, S->wT[s]);
    XG_LAUNCH_CHECK(ctx->es);
  }
  S->wT_epoch = ctx->param_epoch;
  MapTable mt;
  XG_TRY(tc_make_map(ctx, ts, S->wT[0], H, 4 * H, 128, &mt.m[0]));
  XG_TRY(tc_make_map(ctx, ts, S->wT[1], H, 4 * H, 128, &mt.m[1]));
  XG_TRY(tc_make_map(ctx, ts, bp.dz_hi, R, 8 * H, PK_BN, &mt.m[2]));
  XG_TRY(tc_make_map(ctx, ts, bp.dz_lo, R, 8 * H, PK_BN, &mt.m[3]));
  for (int i = 4; i < 18; ++i) mt.m[i] = mt.m[0];
  bp.B = B; bp.R = R; bp.K = K; bp.H = H;
  for (int s = 0; s < 2; ++s) { bp.Gt[s] = eb.G[s]; bp.Cs[s] = eb.Cs[s]; bp.dHs[s] = dH[s]; }
  bp.fmask = fmask;
  bp.sync_counter = S->d_bcounter;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_bparams, &bp, sizeof(EncBwdParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_bcounter, 0, sizeof(unsigned int) * 64, st));
  if (!S->battr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(encode_bwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, encode_bwd_persistent_kernel, PK_THREADS, PK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "persistent encoder backward does not fit on an SM");
    S->battr_set = true;
  }
  ProfScope ps(ctx, "encode_bwd_persistent", st);
  const EncBwdParams* dp = S->d_bparams;
  void* args[2] = {(void*)&dp, (void*)&mt};
  XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)encode_bwd_persistent_kernel, dim3(G), dim3(PK_THREADS), args, PK_SMEM_BYTES, st));
  ctx->n_fused++;
  return XG_OK;
}

// backward word loop of the decoder; io = the saved step-major buffers (G1 / G2 become dz in place)
struct PersistBwdIO {
  const float* seq_mask; int L;
  const float* dOUT;
  float *G1, *G2; const float *C1, *C2, *AH, *ALPHA;
  float *DAH, *dV, *dUv, *dwa_part, *dba_part, *dHcar, *dC1, *dC2;
  DropSpec drop1, drop2;
};

static int persist_decode_bwd(xg_context* ctx, const float* Vf, const float* Uv, int B, int K, int T, const PersistBwdIO& io,
                              cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, A = d.att, G = ctx->sm_count;
  if (!ctx->persist_mode || H % 32 != 0 || A % 32 != 0 || B < 1 || B > PK_MAX_ROWS || G < 16 || G > 256 || T < 1 ||
      2 * K + H + 8 > PK_SCRATCH_FLOATS || (long)K * (H + A / 2) * 4 > (long)PK_STAGES * PK_STAGE_BYTES || (long)K * H * 4 >= (1L << 20))
    return persist_refuse(ctx, "the backward word loop", "rnn_size and att_size must be multiples of 32, at most 1024 captions, V[b] and half of Uv[b] within 192 KB");
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN;
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  PersistState*& S = persist_state(ctx);
  if (!S) S = new PersistState();
  DecBwdParams& bp = S->dbp;
  const int kb4H = 4 * H / 32, kbA = A / 32;
  auto mk = [&](int id, int wmap, int xmap, int n_rows, int nkb) {
    GDesc& g = bp.d[id];
    g.w_map = wmap; g.x_hi = xmap; g.x_lo = xmap + 1; g.xkb0 = 0; g.n_rows = n_rows; g.nkb = nkb;
  };
  // maps: 0..4 transposed weights (i2h2, a2h2, h2h2, h2h1: (H,4H); h2a: (2H,A)); dz2 5,6  dz1 7,8  dah 9,10
  mk(DB_GB, 0, 5, H, kb4H);
  mk(DB_GC, 1, 5, H, kb4H);
  mk(DB_GD, 2, 5, H, kb4H);
  mk(DB_GG, 3, 7, H, kb4H);
  mk(DB_GH, 4, 9, 2 * H, kbA);
  std::vector<PSched> sched;
  if (!persist_plan({{DB_GB, DB_GC, DB_GD}, {DB_GG, DB_GH}}, bp.d, R / PK_BN, G, sched, DECB_MAX_SLOTS))
    return persist_refuse(ctx, "the backward word loop", "no work schedule for this shape");
  if (S->dbB != R || S->dbK != K) {
    if (S->dbpool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->dbpool); S->dbpool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->dbpool, pass == 0 ? 0 : S->dbpool_bytes);
      S->d_dbparams = a.take<DecBwdParams>(1);
      S->d_dbcounter = a.take<unsigned int>(64);
      S->d_dbdbg = a.take<long long>((2048 + 256) * PK_STAMPS + 64);
      bp.sched = a.take<PSched>(sched.size());
      for (int i = 0; i < DB_COUNT; ++i) bp.d[i].out = a.take<float>((size_t)bp.d[i].ns * R * bp.d[i].n_rows);
      bp.dz2_hi = a.take<float>((long)R * 4 * H); bp.dz2_lo = a.take<float>((long)R * 4 * H);
      bp.dz1_hi = a.take<float>((long)R * 4 * H); bp.dz1_lo = a.take<float>((long)R * 4 * H);
      bp.dah_hi = a.take<float>((long)R * A); bp.dah_lo = a.take<float>((long)R * A);
      for (int w = 0; w < 4; ++w) S->dbwT[w] = a.take<float>((long)4 * H * H);
      S->dbwT[4] = a.take<float>((long)2 * H * A);
      if (pass == 0) {
        S->dbpool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->dbpool, S->dbpool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->dbpool, 0, S->dbpool_bytes, st));
      }
    }
    S->dbB = R; S->dbK = K;
    S->dbwT_epoch = ~0ull;
  }
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<PSched*>(bp.sched), sched.data(), sizeof(PSched) * sched.size(),
                                       cudaMemcpyHostToDevice, st));
  // transposed copies of the weights (4 x 4 MB + 6 MB): rebuilt when the parameter epoch moved
  const int wsrc[5] = {XG_P_L2_I2H_W, XG_P_L2_A2H_W, XG_P_L2_H2H_W, XG_P_L1_H2H_W, XG_P_H2A_W};
  for (int w = 0; w < 5 && S->dbwT_epoch != ctx->param_epoch; ++w) {
    int rows, cols;
    param_shape(d, wsrc[w], &rows, &cols);
    ProfScope ps(ctx, "transpose", st);
    transpose_kernel<<<dim3(ceil_div(cols, 32), ceil_div(rows, 32)), dim3(32, 8), 0, st>>>(ctx->P[wsrc[w]], rows, cols, S->dbwT[w]);
    XG_LAUNCH_CHECK(ctx->es);
  }
  S->dbwT_epoch = ctx->param_epoch;
  MapTable mt;
  for (int w = 0; w < 4; ++w) XG_TRY(tc_make_map(ctx, ts, S->dbwT[w], H, 4 * H, 128, &mt.m[w]));
  XG_TRY(tc_make_map(ctx, ts, S->dbwT[4], 2 * H, A, 128, &mt.m[4]));
  XG_TRY(tc_make_map(ctx, ts, bp.dz2_hi, R, 4 * H, PK_BN, &mt.m[5])); XG_TRY(tc_make_map(ctx, ts, bp.dz2_lo, R, 4 * H, PK_BN, &mt.m[6]));
  XG_TRY(tc_make_map(ctx, ts, bp.dz1_hi, R, 4 * H, PK_BN, &mt.m[7])); XG_TRY(tc_make_map(ctx, ts, bp.dz1_lo, R, 4 * H, PK_BN, &mt.m[8]));
  XG_TRY(tc_make_map(ctx, ts, bp.dah_hi, R, A, PK_BN, &mt.m[9])); XG_TRY(tc_make_map(ctx, ts, bp.dah_lo, R, A, PK_BN, &mt.m[10]));
  for (int i = 11; i < 18; ++i) mt.m[i] = mt.m[0];
  bp.B = B; bp.R = R; bp.K = K; bp.H = H; bp.A = A; bp.T = T; bp.L = io.L;
  bp.seq_mask = io.seq_mask; bp.dOUT = io.dOUT;
  bp.G1s = io.G1; bp.G2s = io.G2; bp.C1s = io.C1; bp.C2s = io.C2; bp.AHs = io.AH; bp.ALPHAs = io.ALPHA;
  bp.Uv = Uv; bp.Vf = Vf; bp.w_a2w = ctx->P[XG_P_A2W_W];
  bp.DAH = io.DAH; bp.dV = io.dV; bp.dUv = io.dUv; bp.dwa_part = io.dwa_part; bp.dba_part = io.dba_part;
  bp.dHcar = io.dHcar; bp.dC[0] = io.dC1; bp.dC[1] = io.dC2;
  bp.drop1 = io.drop1; bp.drop2 = io.drop2;
  bp.sync_counter = S->d_dbcounter;
  bp.dbg_clock = (env_flag("XG_PERSIST_TRACE") && T <= 2048) ? S->d_dbdbg : nullptr;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_dbparams, &bp, sizeof(DecBwdParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_dbcounter, 0, sizeof(unsigned int) * 64, st));
  if (!S->dbattr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_bwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_bwd_persistent_kernel, PK_THREADS, PK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "persistent decoder backward does not fit on an SM");
    S->dbattr_set = true;
  }
  ProfScope ps(ctx, "decode_bwd_persistent", st);
  const DecBwdParams* dp = S->d_dbparams;
  void* args[2] = {(void*)&dp, (void*)&mt};
  XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)decode_bwd_persistent_kernel, dim3(G), dim3(PK_THREADS), args, PK_SMEM_BYTES, st));
  ctx->n_fused++;
  if (bp.dbg_clock && T > 4 && G <= 256) {   // XG_PERSIST_TRACE=1: phase timeline of the fourth step, all CTAs
    XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
    std::vector<long long> ga((size_t)G * PK_STAMPS);
    cudaMemcpy(ga.data(), S->d_dbdbg + 2048 * PK_STAMPS, sizeof(long long) * ga.size(), cudaMemcpyDeviceToHost);
    const char* names[4] = {"Pa (lstm_2 cell bwd)", "Gb (dz2 products)", "Pc (attention bwd || lstm_1 cell bwd)", "Gd (dz1, dAH products)"};
    for (int i = 0; i < 4; ++i) {
      long long open = 0, close = 0;
      for (int c = 0; c < G; ++c) open = std::max(open, ga[(size_t)c * PK_STAMPS + 2 * i]);
      std::vector<long long> fin(G);
      for (int c = 0; c < G; ++c) fin[c] = ga[(size_t)c * PK_STAMPS + 2 * i + 1] - open;
      std::vector<long long> srt = fin;
      std::sort(srt.begin(), srt.end());
      for (int c = 0; c < G; ++c) close = std::max(close, ga[(size_t)c * PK_STAMPS + 2 * i + 2]);
      const int worst = (int)(std::max_element(fin.begin(), fin.end()) - fin.begin());
      fprintf(stderr, "[xg bwd trace] step 3 %-38s work done after: min %6lld  median %6lld  p90 %6lld  max %6lld ns (cta %d)   barrier exit %6lld ns\n",
              names[i], srt[0], srt[G / 2], srt[G * 9 / 10], srt[G - 1], worst, close - open);
    }
  }
  return XG_OK;
}

}  // namespace xg
