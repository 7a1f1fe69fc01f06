// The word loops, second generation: the LSTM cells run in the epilogue of their own products.
//
// Kernels of this file (all cooperative, one CTA per SM, fp16 operand pairs through TMA -> tcgen05 -> TMEM):
//   decode_grouped_kernel<0>      greedy word loop of SAModel.sample                      (SAModel.py:182-219)
//   decode_grouped_kernel<1>      the same loop sampling: multinomial draw / training dropout (SCST), and the
//                                 scheduled-sampling token pass                             (SAModel.py:188-196, 89-99)
//   train_grouped_kernel          teacher-forced loop of SAModel.forward                   (SAModel.py:88-111)
//   decode_step_grouped_kernel    beam search: all word steps of a search + the candidate merge of every position
//                                                                                          (SAModel.py:129-161, CaptionModel.py:22-128)
//   encode_grouped_kernel         frame recurrence of both encoder streams                 (sub_modules.py:132-147)
// The description below is the greedy loop's; the others reuse its phases (gphase, fused_cell_phase, dec_attention_rows).
//
// decode_persistent_kernel<0> (xg_persist.cuh) spends six grid barriers per word step: every LSTM layer is a GEMM
// phase (split-K partial tiles to L2), a grid barrier, a pointwise phase that reads the partial tiles back, and
// another grid barrier.  Here a layer is ONE phase:
//
//   * the 4H weight rows are tiled as (32 hidden units) x (4 gates): a 128-row operand tile is four 32-row TMA boxes
//     taken at rows g*H + 32*tile of the nn.Parameter storage (no permuted weight copy), so a tile holds everything
//     the cell of its 32 units needs;
//   * the K extent of ALL products of the layer (lstm_1: xt, gp, h1; lstm_2: h1', af, h2: 47 / 48 k-blocks of 32) is
//     laid end to end and cut over the `members` CTAs of the tile's GROUP (148 SMs / 16 tiles = 9): each member runs
//     one accumulation chain of 5-6 k-blocks through the same TMA -> split -> tcgen05 pipeline as before;
//   * members publish their 128 x 64 partial tile to L2 and arrive on the GROUP's counter (9 arrivals instead of a
//     148-CTA grid barrier), then each member adds the 9 partial tiles of its share of the captions (one warp per
//     caption, lane = hidden unit: every load is a 128-byte line) and runs the cell;
//   * [h1|h2] is double buffered (a cell overwrites h while other groups still stream the old one through TMA).
//
// Word step: F1 {lstm_1} | F3 {lstm_2} | G4 {logits + attention query of step t+1} | P4 {pick || attention}: four grid
// barriers instead of six, and the recurrent products Z1h / Z2h leave the logits phase.
// Hardware thread-block clusters + distributed shared memory were measured for the same job
// (scripts/microbench/cluster_probe.cu, profiles/r2_cluster_probe.txt): a B200 co-schedules only 15 clusters of 8 CTAs
// with this kernel's shared-memory footprint (120 CTAs, 16 are needed), and clusters of 4 need 64-row tiles whose
// longer K runs cost what the DSMEM reduction saves; the L2 version keeps all 148 SMs in every phase.
#pragma once
#include "xg_persist.cuh"

namespace xg {

constexpr int GK_MAX_ITEMS = 12;
constexpr int GK_MAX_MEMBERS = 9;      // CTAs of a group
constexpr int GK_MAX_SLOTS = 12;       // partial tiles a cell adds: the group's members + the "early" ones (below)
enum { GI_CONT_PREV = 1, GI_CONT_NEXT = 2, GI_FUSED = 4, GI_LOGITS = 8, GI_LAYER1 = 16 };
constexpr int GK_LT_STRIDE = 130;     // row stride of the transposed logits tile in shared memory (conflict-free both ways)
constexpr int GK_LH_STRIDE = 136;     // ... of the quarter tile of the single-step mode (eight threads per caption)
constexpr int GK_LQ_BYTES = 16 * GK_LH_STRIDE * 4;          // 16 captions x 128 vocabulary rows (+ padding), one per epilogue warpgroup
constexpr int GK_LT_BYTES = 2 * GK_LQ_BYTES;
constexpr int GK_SMEM_BYTES = PK_SMEM_BYTES + GK_LT_BYTES;  // the quarter tiles live behind the pipeline's shared memory
constexpr int GK_TOPK = 8;            // per-row candidates the logits epilogue can keep (beam search: beam_size <= 8)

struct GItem {            // one run of k-blocks (24 bytes)
  short w_map;            // tensor map of the weight matrix (standalone: 128-row boxes; fused: 32-row boxes)
  short x_map;            // hi map of the activation operand (lo = +1)
  short xsel;             // 0: x_map as is; 1: [h1|h2] entering the step (buffer t & 1); 2: the one being written
  short flags;            // GI_CONT_PREV: accumulate onto the previous item; GI_CONT_NEXT: the next item continues; GI_FUSED
  short wrow;             // standalone: first weight row of the tile; fused: first hidden unit (boxes at g*H + wrow)
  short wk0, xk0, nkb;    // first k-block in the weight / activation matrix, run length
  short desc, slot;       // standalone: product (DecParams.d[desc]) and split-K slot; fused: group, member
  short cb, pad;          // caption column block
};
// dual: the chains of this phase alternate between TWO epilogue warpgroups (warps 2-5 take the even chains, warps 6-9 the odd
// ones); chunks_b: accumulation chunks of the odd chains.  Reading an accumulator back costs ~1000 cycles per 64 KB pair
// (tensor memory reads run at 64 B/clk) and the end of a chain (partial-tile store, logits reduction) 1.5k - 11k more: with
// several chains per CTA (beam search: one per caption column block) one warpgroup was the serial resource of the phase.
struct GSched { short n, tot_kb, tot_chunks, n_chains, dual, chunks_b; GItem it[GK_MAX_ITEMS]; };   // 300 bytes

// Operands are fp16 hi / lo PAIRS (3xFP16: hi.hi + (lo.hi + hi.lo) / 2^11, the same 11 + 11 mantissa bits as the 3xTF32
// products of xg_persist.cuh): the weights as derived tables rebuilt when the bound parameters change (like the
// POS-gate table), the activations written that way by the pointwise phases.  A k-block is 64 wide (128-byte rows), a
// stage holds [W hi 16K | W lo 16K | X hi 8K | X lo 8K], all of it delivered by TMA: no in-kernel split pass, and
// 60 KB of shared-memory traffic per 32 K-elements instead of 136 KB (the GEMM phases were shared-memory bound).
constexpr int GK_KB = 64;             // K elements of a k-block
struct MapTable2 { CUtensorMap m[28]; };
// hi map (lo = +1) of: 0 h2a, 2 logit (128-row boxes); 4,6,8,10,12,14 l1_i2h, l1_a2h, l1_h2h, l2_i2h, l2_a2h, l2_h2h (32-row
// boxes); 16 xt; 18 / 20 the two [h1|h2] buffers; 22 gp; 24 af; 26: V as [B*K][H] fp32 (box H/2 x K, no swizzle)
constexpr int GM_H2A = 0, GM_LOGIT = 2, GM_W32 = 4, GM_XT = 16, GM_HH = 18, GM_GP = 22, GM_AF = 24, GM_V = 26;

// Candidate merge + bookkeeping of one beam-search position for video k (CaptionModel.beam_step and the done-beam
// harvest, CaptionModel.py:34-118), one warp.  The stable descending sort is a rank count (rank = candidates that score
// higher, or equal with a lower index).  Called by beam_merge_kernel (xg_beam.cuh) and by the fused step kernel below.
struct BeamMergeIO {
  const float* ys; const int* ix;          // (videos*beam, beam) per-row top-`beam` log-probs / ids (UNK penalty applied)
  int B, beam, T, t;                       // videos, beam size, seq_length, position being decided
  const int64_t* seq_in; const float* lps_in;
  int64_t* seq_out; float* lps_out;        // (videos, beam, T) ping-pong
  float* sum; int* parent; int64_t* tokens;
  int64_t* done_seq; float* done_lps; float* done_p; int* done_n;
};
constexpr int BM_MAX_BEAM = 16;
__device__ __forceinline__ void beam_merge_warp(const BeamMergeIO& M, int k, double* cp /* beam*beam */, int* sel /* beam */, int lane) {
  const int beam = M.beam, T = M.T, t = M.t;
  const int rows = (t == 0) ? 1 : beam;
  const int ncand = rows * beam;
  // enumerate column-major: c outer, q inner
  for (int n = lane; n < ncand; n += 32) {
    const int c = n / rows, q = n % rows;
    cp[n] = (double)__ldcg(M.sum + (long)k * beam + q) + (double)__ldcg(M.ys + ((long)k * beam + q) * beam + c);
  }
  if (lane < beam) sel[lane] = lane < ncand ? lane : 0;
  __syncwarp();
  for (int n = lane; n < ncand; n += 32) {
    const double pn = cp[n];
    int rank = 0;
    for (int m = 0; m < ncand; ++m) rank += (cp[m] > pn || (cp[m] == pn && m < n)) ? 1 : 0;
    if (rank < beam) sel[rank] = n;
  }
  __syncwarp();
  // lane v < beam owns the v-th survivor: its word, log-prob, sum, parent row and whether it ends here
  const int dn0 = __ldcg(M.done_n + k);
  int my_word = 0, my_q = 0; float my_wlp = 0.f, my_sum = 0.f; bool my_done = false;
  if (lane < beam) {
    const int cand = sel[lane];
    const int c = cand / rows; my_q = cand % rows;
    my_word = __ldcg(M.ix + ((long)k * beam + my_q) * beam + c);
    my_wlp = __ldcg(M.ys + ((long)k * beam + my_q) * beam + c);
    my_sum = (float)cp[cand];
    my_done = my_word == 0 || t == T - 1;
  }
  const unsigned dmask = __ballot_sync(0xffffffffu, my_done);
  // the survivors' histories, four at a time: every load of a batch is issued before its first store (the buffers may
  // alias as far as the compiler knows, and one survivor at a time meant one L2 round trip per survivor)
#pragma unroll 1
  for (int v0 = 0; v0 < beam; v0 += 4) {
    int word[4], q[4], dnv[4]; float wlp[4]; bool done[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int vix = min(v0 + b, beam - 1);
      word[b] = __shfl_sync(0xffffffffu, my_word, vix); q[b] = __shfl_sync(0xffffffffu, my_q, vix);
      wlp[b] = __shfl_sync(0xffffffffu, my_wlp, vix);
      done[b] = (dmask >> vix) & 1u;
      dnv[b] = dn0 + __popc(dmask & ((1u << vix) - 1u));
    }
#pragma unroll 1
    for (int u = lane; u < T; u += 32) {
      int64_t sv[4]; float lv[4];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const long src = ((long)k * beam + q[b]) * T;
        sv[b] = u < t ? __ldcg(M.seq_in + src + u) : (u == t ? (int64_t)word[b] : (int64_t)0);
        lv[b] = u < t ? __ldcg(M.lps_in + src + u) : (u == t ? wlp[b] : 0.f);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (v0 + b < beam) {
          const long dst = ((long)k * beam + v0 + b) * T;
          M.seq_out[dst + u] = sv[b]; M.lps_out[dst + u] = lv[b];
          if (done[b]) { const long dd = ((long)k * T * beam + dnv[b]) * T; M.done_seq[dd + u] = sv[b]; M.done_lps[dd + u] = lv[b]; }
        }
      }
    }
  }
  if (lane < beam) {      // (sum[] of this video was read for every candidate before it is rewritten here)
    if (my_done) M.done_p[(long)k * T * beam + dn0 + __popc(dmask & ((1u << lane) - 1u))] = my_sum;
    M.sum[(long)k * beam + lane] = my_done ? -1000.f : my_sum;
    M.parent[(long)k * beam + lane] = k * beam + my_q;
    M.tokens[(long)k * beam + lane] = my_word;
  }
  if (lane == 0) M.done_n[k] = dn0 + __popc(dmask);
}

struct GroupParams {
  DecParams dp;                  // the pick / attention / token-input phases of xg_persist.cuh read this part
  const GSched* gsched;          // [3][G]: F1, F3, G4
  float* fslots[2];              // per layer [groups][nslots][64 captions][128 rows] partial tiles of a fused cell phase
  unsigned int* group_ctr;       // [2][groups] arrival counters (monotonic over the steps of a launch)
  int members[2], groups, ncb;   // CTAs per group in F1 / F3; groups = (H/32) x caption column blocks
  int nslots[2];                 // members + early slots: the recurrent products W_h2h1.h1 / W_h2h2.h2 of step t+1 are
                                 //   computed next to the logits of step t (G4) into slots members..nslots-1
  int n_att;                     // the last n_att CTAs run the attention of step t+1 THROUGH the pick phase and F1
  unsigned int* pick_ctr;        // barrier counter of the pick phase (the attention CTAs are not part of it)
  __half* hh_hi[2]; __half* hh_lo[2]; // [R][2H] x 2: [h1|h2] entering the step / being written (fp16 hi / lo pairs)
  float4* lpart;                 // [R][ntv] per (caption, 128-row vocabulary tile): max logit, sum exp(x - max), arg-max
  int ntv;                       // vocabulary tiles
  int l2_hints;                  // 1: L2 eviction hints on the weight loads (XG_L2_HINT=0 switches them off)
  // single-step mode (beam search): the logits (+ bias) of every row are also stored, [R][ntv * 128] (-inf beyond V): the
  // row merge rescans the few vocabulary tiles that can hold a row's topk
  int topk;
  float* lraw;
  // sampling form of the word loop (SAModel.sample with sample_max = 0 and / or under model.train(): SAModel.py:188-196,
  // starttrain.py:131): multinomial draw from exp(logprob / temperature) in the pick phase, training dropout on the POS
  // gate and on both cells with the Philox sites / indices of xg_train_fwd (the teacher-forced replay reproduces them)
  int sample_max, step_drop;
  float inv_temp;
  unsigned long long sample_seed;
  DropSpec drop_gate, drop_h1, drop_h2;
  // scheduled-sampling token pass (SAModel.py:89-99): the input token of step i >= 1 is the ground truth seq[b, i] or,
  // with probability ss_prob, a draw from the word distribution of step i - 1; the step mask is seq_mask[b, i]
  int ss_mode, ss_L, ss_Lp;
  float ss_prob;
  unsigned long long ss_seed;
  BeamMergeIO mg;               // single-step mode: the candidate merge of the position runs at the end of the launch (mg_on)
  int mg_on;
  const int64_t* ss_seq;        // (B, L)
  const float* ss_mask;         // (B, L)
  int64_t* ss_used;             // (B, L) tokens actually fed (initialised with seq by the host)
};

// L2 residency: the recurrent weights + the attention operands (47 MB) are re-read every word step and fit one L2
// partition; the 20.5 MB logit matrix streamed behind them every step does not, and under the default policy it cycles
// everything out (ncu r1: 42 of the 52.6 MB of weights came from DRAM every step).  Weight tiles of the recurrent
// products are loaded evict_last, the logit stream evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_normal() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void pk_tma_2d_hint(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void pk_tma_prefetch_l2_hint(const CUtensorMap* tm, int c0, int c1, uint64_t pol) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile.L2::cache_hint [%0, {%1, %2}], %3;" ::"l"(tm), "r"(c0), "r"(c1), "l"(pol) : "memory");
}

// all work items of this CTA for one GEMM phase.  Same pipeline and accumulation discipline as gemm_phase
// (xg_persist.cuh); items may chain (several k-block runs, possibly of different products, into one accumulator
// set) and a fused chain leaves its partial tile in the group's slot buffer.
template <bool STEP, bool DUAL = false>      // DUAL: the schedule may ask for both epilogue warpgroups (GSched::dual; the beam-search step kernel)
__device__ __noinline__ void gphase(const GroupParams& C, const GSched* sc, const GSched* sc_next, const CUtensorMap* maps,
                                    int par, const SmemView& sv, uint32_t tmem_base, PipeState& ps) {
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n_items = __shfl_sync(0xffffffffu, (int)sc->n, 0);
  const bool dual = DUAL && __shfl_sync(0xffffffffu, (int)sc->dual, 0) != 0;
  const int R = C.dp.R, H = C.dp.H;
#ifdef GK_FINE
  long long* gw = (C.dp.dbg_clock && blockIdx.x == 0) ? C.dp.dbg_clock + (2048 + 256) * PK_STAMPS + 32 : nullptr;
  long long w_acc[4] = {0, 0, 0, 0};
  const long long t_in = clock64();
#define GKW(i, stmt) do { const long long t_ = clock64(); stmt; w_acc[i] += clock64() - t_; } while (0)
#else
#define GKW(i, stmt) do { stmt; } while (0)
#endif
  if (warp == 0) {            // ===== TMA producer =====
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first(), pol_norm = l2_policy_normal();
    uint32_t cnt = __shfl_sync(0xffffffffu, ps.kb_count, 0);
    int npre = (int)__shfl_sync(0xffffffffu, ps.npre, 0);
    const uint32_t stages_u32 = __shfl_sync(0xffffffffu, sv.stages_u32, 0);
    const uint32_t full_bar = __shfl_sync(0xffffffffu, sv.full_bar, 0), empty_bar = __shfl_sync(0xffffffffu, sv.empty_bar, 0);
#pragma unroll 1
    for (int ii = 0; ii < n_items; ++ii) {
      const GItem it = sc->it[ii];
      const int xm = it.x_map + (it.xsel ? 2 * ((par + it.xsel - 1) & 1) : 0);
      const CUtensorMap* mw = maps + __shfl_sync(0xffffffffu, (int)it.w_map, 0);
      const CUtensorMap* mwl = mw + 1;
      const CUtensorMap* mxh = maps + __shfl_sync(0xffffffffu, xm, 0);
      const CUtensorMap* mxl = mxh + 1;
      int wk = __shfl_sync(0xffffffffu, it.wk0 * GK_KB, 0), xk = __shfl_sync(0xffffffffu, it.xk0 * GK_KB, 0);
      const int row = __shfl_sync(0xffffffffu, (int)it.wrow, 0), col = __shfl_sync(0xffffffffu, it.cb * PK_BN, 0);
      const int nkb = __shfl_sync(0xffffffffu, (int)it.nkb, 0);
      const int fused = __shfl_sync(0xffffffffu, (int)(it.flags & GI_FUSED), 0);
      const uint64_t pol = !C.l2_hints ? pol_norm : (__shfl_sync(0xffffffffu, (int)(it.flags & GI_LOGITS), 0) ? pol_stream : pol_keep);
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb, ++cnt, wk += GK_KB, xk += GK_KB) {
        const uint32_t s = cnt & (PK_STAGES - 1);
        const uint32_t st = stages_u32 + s * PK_STAGE_BYTES, fb = full_bar + 8 * s;
        if (npre > 0) {         // weights already on their way (gprefetch): only the activation tiles remain
          --npre;
          if (elect_one_sync()) {
            pk_expect_tx(fb, 2 * PK_X_BYTES);
            pk_tma_2d(st + 2 * PK_W_BYTES, mxh, fb, xk, col);
            pk_tma_2d(st + 2 * PK_W_BYTES + PK_X_BYTES, mxl, fb, xk, col);
          }
        } else {
          GKW(0, pk_wait(empty_bar + 8 * s, ((cnt / PK_STAGES) & 1) ^ 1));
          if (elect_one_sync()) {
            pk_expect_tx(fb, PK_STAGE_BYTES);
            if (fused) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                pk_tma_2d_hint(st + g * (PK_W_BYTES / 4), mw, fb, wk, g * H + row, pol);
                pk_tma_2d_hint(st + PK_W_BYTES + g * (PK_W_BYTES / 4), mwl, fb, wk, g * H + row, pol);
              }
            } else {
              pk_tma_2d_hint(st, mw, fb, wk, row, pol);
              pk_tma_2d_hint(st + PK_W_BYTES, mwl, fb, wk, row, pol);
            }
            pk_tma_2d(st + 2 * PK_W_BYTES, mxh, fb, xk, col);
            pk_tma_2d(st + 2 * PK_W_BYTES + PK_X_BYTES, mxl, fb, xk, col);
          }
        }
        __syncwarp();
      }
    }
    if (sc_next != nullptr) {       // the weight tiles of the NEXT GEMM phase go to L2 now
      const int nn = __shfl_sync(0xffffffffu, (int)sc_next->n, 0);
#pragma unroll 1
      for (int ii = 0; ii < nn; ++ii) {
        const GItem it = sc_next->it[ii];
        const CUtensorMap* mw = maps + __shfl_sync(0xffffffffu, (int)it.w_map, 0);
        int wk = __shfl_sync(0xffffffffu, it.wk0 * GK_KB, 0);
        const int row = __shfl_sync(0xffffffffu, (int)it.wrow, 0);
        const int nkb = __shfl_sync(0xffffffffu, (int)it.nkb, 0);
        const int fused = __shfl_sync(0xffffffffu, (int)(it.flags & GI_FUSED), 0);
        const uint64_t pol = !C.l2_hints ? pol_norm : (__shfl_sync(0xffffffffu, (int)(it.flags & GI_LOGITS), 0) ? pol_stream : pol_keep);
#pragma unroll 1
        for (int kb = 0; kb < nkb; ++kb, wk += GK_KB) {
          if (elect_one_sync()) {
            if (fused) {
#pragma unroll
              for (int g = 0; g < 4; ++g) { pk_tma_prefetch_l2_hint(mw, wk, g * H + row, pol); pk_tma_prefetch_l2_hint(mw + 1, wk, g * H + row, pol); }
            } else {
              pk_tma_prefetch_l2_hint(mw, wk, row, pol); pk_tma_prefetch_l2_hint(mw + 1, wk, row, pol);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {     // ===== MMA issuer =====
    // Per 16-wide k-step TWO instructions: W_hi . [X_hi ; X_lo] as one N = 128 product (the hi and lo activation tiles lie
    // back to back in the stage: main term -> columns [0, 64), cross term hi.lo -> [64, 128) of the accumulator pair) and
    // W_lo . X_hi (N = 64) onto the cross columns.  Three N = 64 instructions per k-step measured ~80 cycles each (the
    // 128 x 16 weight tile is read from shared memory once per instruction): the MMA warp, not the operand traffic,
    // paced the GEMM phases.  The pair is drained into fp32 registers every PK_CHUNK k-blocks (truncating adds in TMEM).
    constexpr uint32_t idesc = umma_idesc_f16(128, PK_BN), idesc2 = umma_idesc_f16(128, 2 * PK_BN);
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t cnt = __shfl_sync(0xffffffffu, ps.kb_count, 0), cc = __shfl_sync(0xffffffffu, ps.chunk_count, 0);
    const uint32_t ready_bar = __shfl_sync(0xffffffffu, sv.full_bar, 0), empty_bar = __shfl_sync(0xffffffffu, sv.empty_bar, 0);
    const uint32_t acc_full = __shfl_sync(0xffffffffu, sv.acc_full, 0), acc_empty = __shfl_sync(0xffffffffu, sv.acc_empty, 0);
    const uint32_t acc_full2 = __shfl_sync(0xffffffffu, sv.acc_full2, 0);
    const uint32_t desc_lo0 = __shfl_sync(0xffffffffu, (uint32_t)umma_desc_sw128(sv.stages_u32), 0);
    const uint32_t desc_hi = (uint32_t)(umma_desc_sw128(0) >> 32);
    // the "accumulator full" barriers belong to the epilogue warpgroup that owns the chain: slot = that group's own chunk count
    uint32_t cg[2];
    cg[1] = DUAL ? __shfl_sync(0xffffffffu, ps.grp_b, 0) : 0u; cg[0] = cc - cg[1];
    uint32_t owner = 0, ch = 0;
#pragma unroll 1
    for (int ii = 0; ii < n_items; ++ii) {
      const int nkb = __shfl_sync(0xffffffffu, (int)sc->it[ii].nkb, 0);
      if (DUAL && !(__shfl_sync(0xffffffffu, (int)sc->it[ii].flags, 0) & GI_CONT_PREV)) { owner = dual ? (ch & 1u) : 0u; ++ch; }
      uint32_t tmem_pair = tb, b = 0;
#pragma unroll 1
      for (int kb = 0; kb < nkb; ++kb, ++cnt) {
        if ((kb & 1) == 0) {
          b = cc & 1;
          GKW(1, pk_wait(acc_empty + 8 * b, ((cc >> 1) & 1) ^ 1));
          tc_fence_after();
          tmem_pair = tb + b * 2 * PK_BN;
        }
        const uint32_t s = cnt & (PK_STAGES - 1);
        GKW(0, pk_wait(ready_bar + 8 * s, (cnt / PK_STAGES) & 1));     // all four tiles of the stage landed
        tc_fence_after();
        const uint32_t dlo = desc_lo0 + ((s * PK_STAGE_BYTES) >> 4);
        if (elect_one_sync()) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const uint64_t wh = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2);
            const uint64_t wl = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2 + (PK_W_BYTES >> 4));
            const uint64_t xh = ((uint64_t)desc_hi << 32) | (uint64_t)(dlo + k4 * 2 + ((2 * PK_W_BYTES) >> 4));      // 128 rows: X_hi, X_lo
            umma_f16(tmem_pair, wh, xh, idesc2, ((kb & 1) | k4) != 0);
            umma_f16(tmem_pair + PK_BN, wl, xh, idesc, 1);
          }
          pk_commit(empty_bar + 8 * s);
          if ((kb & 1) || kb == nkb - 1) pk_commit(DUAL ? (owner ? acc_full2 : acc_full) + 8 * (cg[owner] & 1u) : acc_full + 8 * b);
        }
        __syncwarp();
        if ((kb & 1) || kb == nkb - 1) { ++cc; if (DUAL) ++cg[owner]; }
      }
    }
  } else if (DUAL || warp < 6) {      // ===== epilogue: promote short chains into fp32 registers, store the partial tile =====
    // warps 2-5: warpgroup 0, warps 6-9: warpgroup 1 (a warp reads the 32 TMEM lanes of its index mod 4 either way)
    const int quad = warp & 3, grp = (DUAL && warp >= 6) ? 1 : 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t my_full = grp ? sv.acc_full2 : sv.acc_full;
    uint32_t cc = ps.chunk_count, cg = DUAL ? (grp ? ps.grp_b : ps.chunk_count - ps.grp_b) : ps.chunk_count, ch = 0;
    bool mine = !DUAL;
    float acc[PK_BN];
#pragma unroll 1
    for (int ii = 0; ii < n_items; ++ii) {
      const GItem it = sc->it[ii];
      const bool first = !(it.flags & GI_CONT_PREV), last = !(it.flags & GI_CONT_NEXT);
      const int n_chunks = (it.nkb + PK_CHUNK - 1) / PK_CHUNK;
      if (DUAL) {
        if (first) { mine = dual ? ((ch & 1u) == (uint32_t)grp) : (grp == 0); ++ch; }
        if (!mine) { cc += (uint32_t)n_chunks; continue; }
      }
      if (first) {
#pragma unroll
        for (int u = 0; u < PK_BN; ++u) acc[u] = 0.f;
      }
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c, ++cg) {
        const uint32_t b = cc & 1;
        GKW(0, pk_wait(my_full + 8 * (cg & 1u), (cg >> 1) & 1));
        tc_fence_after();
        const uint32_t col = b * 2 * PK_BN;
#ifdef GK_FINE
        const long long td0 = clock64();
#endif
        {
          uint32_t r[32], r2[32];
          tmem_ld32(taddr + col, r);
          tmem_ld32(taddr + col + PK_BN, r2);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[u] += __uint_as_float(r[u]);
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[u] = fmaf(__uint_as_float(r2[u]), 1.f / X16_SCALE, acc[u]);      // the cross terms carry the 2^11 of the lo parts
          tmem_ld32(taddr + col + 32, r);
          tmem_ld32(taddr + col + PK_BN + 32, r2);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[32 + u] += __uint_as_float(r[u]);
#pragma unroll
          for (int u = 0; u < 32; ++u) acc[32 + u] = fmaf(__uint_as_float(r2[u]), 1.f / X16_SCALE, acc[32 + u]);
        }
        tc_fence_before();
        pk_arrive(sv.acc_empty + 8 * b);
        ++cc;
#ifdef GK_FINE
        w_acc[1] += clock64() - td0;
#endif
      }
#ifdef GK_FINE
      const long long te0 = clock64();
#endif
      if (last && (it.flags & GI_LOGITS)) {
        if (!STEP) {                     // greedy decoding
          // The logits of this 128-row vocabulary tile never leave the SM: + bias, transposed through shared memory (the
          // pipeline stages are idle: a logits item is the only item of its CTA), then per caption the tile's max /
          // lowest arg-max / sum exp(x - max).  The pick phase combines the ntv partial results of a caption.
          const int V = C.dp.V;
          const int nl = quad * 32 + lane, n = it.wrow + nl;
          const float bl = n < V ? __ldg(C.dp.b_logit + n) : 0.f;
          float* T = reinterpret_cast<float*>(sv.stages);
#pragma unroll
          for (int u = 0; u < PK_BN; ++u) T[u * GK_LT_STRIDE + nl] = n < V ? acc[u] + bl : -INFINITY;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          const int e = (warp - 2 - 4 * grp) * 32 + lane, c = e >> 1, hh = e & 1;
          const float* row = T + c * GK_LT_STRIDE + hh;
          float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll 8
          for (int i = 0; i < 64; ++i) {
            const float x = row[2 * i];
            if (x > best) { best = x; bi = 2 * i + hh; }          // ascending rows: the first maximum is kept
          }
          {
            const float ob = __shfl_xor_sync(0xffffffffu, best, 1);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, 1);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
          }
          float sum = 0.f;
#pragma unroll 8
          for (int i = 0; i < 64; ++i) sum += __expf(row[2 * i] - best);
          sum += __shfl_xor_sync(0xffffffffu, sum, 1);
          if (hh == 0)
            C.lpart[(long)(it.cb * PK_BN + c) * C.ntv + it.wrow / 128] = make_float4(best, sum, __int_as_float(it.wrow + bi), 0.f);
          fence_proxy_async_smem();        // the stages go back to the TMA / bulk-copy engines
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        } else {                         // beam search: several items per CTA, logits stored for the row merge
          // The logits of this 128-row vocabulary tile: + bias, transposed through shared memory in four quarters of 16
          // captions (a buffer of its own: the pipeline stages may already hold the next item), then per caption the tile's
          // max / lowest arg-max / sum exp(x - max): eight threads per caption.  Greedy decoding needs nothing else (the
          // logits never leave the SM); beam search also stores them for the row merge.
          const int V = C.dp.V;
          const int nl = quad * 32 + lane, n = it.wrow + nl;
          const float bl = n < V ? __ldg(C.dp.b_logit + n) : 0.f;
          // (behind scratch + barriers: carve_smem keeps 1 KB of slack in front; one quarter-tile buffer per warpgroup)
          float* T = reinterpret_cast<float*>(sv.stages + PK_SMEM_BYTES - 1024 + grp * GK_LQ_BYTES);
          const int e = (warp - 2 - 4 * grp) * 32 + lane, cl = e >> 3, part = e & 7;
          const float invT = C.inv_temp;
          float* o = C.lraw ? C.lraw + (long)(it.cb * PK_BN) * (C.ntv * 128) + n : nullptr;
          const long str = (long)C.ntv * 128;
          // A LOOP over the quarters (the accumulators rotate down by 16 per pass, so the body always works on acc[0..16)):
          // unrolled, the four passes were ~1400 instructions of straight-line code executed once per item, and their
          // instruction fetches - not their arithmetic - set the pace of the pass (2500 cycles for ~350 instructions).
#pragma unroll 1
          for (int qt = 0; qt < 4; ++qt) {
#ifdef GK_FINE
            const long long q0 = clock64();
#endif
            if (o != nullptr) {
#pragma unroll
              for (int u = 0; u < 16; ++u) { __stcg(o, n < V ? acc[u] + bl : -INFINITY); o += str; }
            }
#pragma unroll
            for (int u = 0; u < 16; ++u) T[u * GK_LH_STRIDE + nl] = n < V ? acc[u] + bl : -INFINITY;
#ifdef GK_FINE
            const long long q1 = clock64();
#endif
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#ifdef GK_FINE
            const long long q2 = clock64();
#endif
            const float* row = T + cl * GK_LH_STRIDE + part;
            float xs[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) xs[i] = row[8 * i];
            float best = xs[0]; int bi = 0;
#pragma unroll
            for (int i = 1; i < 16; ++i)
              if (xs[i] > best) { best = xs[i]; bi = i; }                   // ascending ids: the first maximum is kept
            bi = it.wrow + 8 * bi + part;
#pragma unroll
            for (int sh = 1; sh <= 4; sh <<= 1) {
              const float ob = __shfl_xor_sync(0xffffffffu, best, sh);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, sh);
              if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            float sum = 0.f, sumT = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) { const float dx = xs[i] - best; sum += __expf(dx); sumT += __expf(dx * invT); }
#pragma unroll
            for (int sh = 1; sh <= 4; sh <<= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, sh); sumT += __shfl_xor_sync(0xffffffffu, sumT, sh); }
#ifdef GK_FINE
            const long long q3 = clock64();
#endif
            if (part == 0)      // (.w: the tile's mass at the sampling temperature, relative to its own max)
              C.lpart[(long)(it.cb * PK_BN + qt * 16 + cl) * C.ntv + it.wrow / 128] = make_float4(best, sum, __int_as_float(bi), sumT);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
#pragma unroll
            for (int u = 0; u < PK_BN - 16; ++u) acc[u] = acc[u + 16];
#ifdef GK_FINE
            if (gw && warp == 2 && lane == 0) { const long long q4 = clock64(); gw[6] += q1 - q0; gw[7] += q2 - q1; gw[14] += q3 - q2; gw[15] += q4 - q3; }
#endif
          }
        }
      } else if (last) {
        if (it.flags & GI_FUSED) {      // [group][member][caption][row]: lanes -> consecutive rows
          float* o = C.fslots[(it.flags & GI_LAYER1) ? 1 : 0] + ((long)(it.desc * it.pad + it.slot) * PK_BN) * 128 + quad * 32 + lane;   // pad = nslots
#pragma unroll
          for (int u = 0; u < PK_BN; ++u) { __stcg(o, acc[u]); o += 128; }
        } else {
          const GDesc& d = C.dp.d[it.desc];
          const int n = it.wrow + quad * 32 + lane;          // weight row held by this thread
          if (n < d.n_rows) {
            float* o = d.out + ((long)it.slot * R + it.cb * PK_BN) * d.n_rows + n;
            const long str = d.n_rows;
#pragma unroll
            for (int u = 0; u < PK_BN; ++u) { __stcg(o, acc[u]); o += str; }
          }
        }
      }
#ifdef GK_FINE
      if (last && (it.flags & GI_LOGITS)) w_acc[3] += clock64() - te0; else w_acc[2] += clock64() - te0;
#endif
    }
  }                           // (warps 6-9 have no role in the GEMM phases: the operands arrive split)
#ifdef GK_FINE
  if (gw && lane == 0 && warp <= 2) {      // per role: time in this phase, waits (producer: stage free | MMA: data, accumulator free,
    const long long dt = clock64() - t_in; //  cross accumulator free | epilogue warp 2: main accumulator full, cross accumulator full)
    gw[warp * 8 + 0] += dt; gw[warp * 8 + 1] += w_acc[0]; gw[warp * 8 + 2] += w_acc[1]; gw[warp * 8 + 3] += w_acc[2]; gw[warp * 8 + 4] += sc->tot_kb; gw[warp * 8 + 5] += w_acc[3];
  }
#endif
  ps.kb_count += sc->tot_kb;
  ps.chunk_count += sc->tot_chunks;
  ps.item_count += sc->n_chains;
  if (DUAL) ps.grp_b += (uint32_t)sc->chunks_b;
  ps.npre = 0;
}

// weight tiles of the first k-blocks of the next GEMM phase, issued before the grid barrier that precedes it
__device__ __noinline__ void gprefetch(const GroupParams& C, const GSched* sc, const CUtensorMap* maps, const SmemView& sv, PipeState& ps) {
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int n_items = __shfl_sync(0xffffffffu, (int)sc->n, 0);
  const int H = C.dp.H;
  if (warp == 0) {
    int npre = 0;
    uint32_t cnt = __shfl_sync(0xffffffffu, ps.kb_count, 0);
    const uint32_t stages_u32 = __shfl_sync(0xffffffffu, sv.stages_u32, 0);
    const uint32_t full_bar = __shfl_sync(0xffffffffu, sv.full_bar, 0), empty_bar = __shfl_sync(0xffffffffu, sv.empty_bar, 0);
#pragma unroll 1
    for (int ii = 0; ii < n_items && npre < PK_STAGES; ++ii) {
      const GItem it = sc->it[ii];
      const CUtensorMap* mw = maps + __shfl_sync(0xffffffffu, (int)it.w_map, 0);
      int wk = __shfl_sync(0xffffffffu, it.wk0 * GK_KB, 0);
      const int row = __shfl_sync(0xffffffffu, (int)it.wrow, 0);
      const int nkb = __shfl_sync(0xffffffffu, (int)it.nkb, 0);
      const int fused = __shfl_sync(0xffffffffu, (int)(it.flags & GI_FUSED), 0);
      const uint64_t pol = !C.l2_hints ? l2_policy_normal() : ((it.flags & GI_LOGITS) ? l2_policy_evict_first() : l2_policy_evict_last());
#pragma unroll 1
      for (int kb = 0; kb < nkb && npre < PK_STAGES; ++kb, ++cnt, wk += GK_KB, ++npre) {
        const uint32_t s = cnt & (PK_STAGES - 1);
        pk_wait(empty_bar + 8 * s, ((cnt / PK_STAGES) & 1) ^ 1);
        if (elect_one_sync()) {
          pk_expect_tx_noarrive(full_bar + 8 * s, 2 * PK_W_BYTES);
          const uint32_t st = stages_u32 + s * PK_STAGE_BYTES, fb = full_bar + 8 * s;
          if (fused) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              pk_tma_2d_hint(st + g * (PK_W_BYTES / 4), mw, fb, wk, g * H + row, pol);
              pk_tma_2d_hint(st + PK_W_BYTES + g * (PK_W_BYTES / 4), mw + 1, fb, wk, g * H + row, pol);
            }
          } else {
            pk_tma_2d_hint(st, mw, fb, wk, row, pol);
            pk_tma_2d_hint(st + PK_W_BYTES, mw + 1, fb, wk, row, pol);
          }
        }
        __syncwarp();
      }
    }
  }
  const int tot = sc->tot_kb;
  ps.npre = (uint32_t)(tot < PK_STAGES ? tot : PK_STAGES);
}

// cell of the captions [c0, c1) of a group: one warp per caption, lane = hidden unit; NCAP captions per pass with every
// partial-tile load of all of them in flight before the first add (MAXS bounds the slots a cell adds)
template <int MAXS, int NCAP, bool DROP = false>
__device__ __forceinline__ void group_cell(const GroupParams& C, int layer, int t, int grp, int c0, int c1) {
  const DecParams& P = C.dp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = P.H, R = P.R, B = P.B;
  const int tile = grp / C.ncb, cb = grp % C.ncb;
  const int ns = C.nslots[layer], par = t & 1;
  const int j = tile * 32 + lane;                              // hidden unit of this lane
  const float* bi = P.bias[layer][0]; const float* ba = P.bias[layer][1]; const float* bh = P.bias[layer][2];
  float* cst = P.cx + (long)layer * R * H;
  __half* hi_new = C.hh_hi[par ^ 1]; __half* lo_new = C.hh_lo[par ^ 1];
  const float* fs = C.fslots[layer] + ((long)(grp * ns) * PK_BN) * 128 + lane;
  float bias[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) { const int n = g * H + j; bias[g] = __ldg(bi + n) + __ldg(ba + n) + __ldg(bh + n); }
#pragma unroll 1
  for (int cA = c0 + warp; cA < c1; cA += NCAP * PK_WARPS) {
    float v[NCAP][4][MAXS], mk[NCAP], cp[NCAP], hp[NCAP];
#pragma unroll
    for (int q = 0; q < NCAP; ++q) {
      const int c = cA + q * PK_WARPS;
      const bool on = c < c1;
      const float* base = fs + (long)(on ? c : cA) * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int k = 0; k < MAXS; ++k) v[q][g][k] = (on && k < ns) ? __ldcg(base + (long)k * PK_BN * 128 + g * 32) : 0.f;
      const int r = cb * PK_BN + c;
      const bool live = on && r < B;
      mk[q] = (live && (t > 0 || (DROP && C.ss_mode))) ? __ldcg(P.unfinished + r) : 1.f;
      cp[q] = live ? __ldcg(cst + (long)r * H + j) : 0.f;
      hp[q] = live ? __ldcg(P.hx + (long)r * 2 * H + layer * H + j) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < NCAP; ++q) {
      const int c = cA + q * PK_WARPS, r = cb * PK_BN + c;
      if (c >= c1 || r >= B) continue;                         // padding rows of the 64-wide operand tiles stay zero
      float z[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < MAXS; ++k) sum += v[q][g][k];
        z[g] = sum + bias[g];
      }
      const float ig = sigmoid_fast(z[0]), fg = sigmoid_fast(z[1]), og = sigmoid_fast(z[2]), gg = tanh_fast(z[3]);
      float cn = fg * cp[q] + ig * gg;
      cn = cn * mk[q] + cp[q] * (1.f - mk[q]);
      float h = og * tanh_fast(cn);
      h = h * mk[q] + hp[q] * (1.f - mk[q]);
      if (DROP) h *= (layer == 0 ? C.drop_h1 : C.drop_h2).factor((uint64_t)t * B * H + (uint64_t)r * H + j);
      cst[(long)r * H + j] = cn;
      P.hx[(long)r * 2 * H + layer * H + j] = h;
      store_split16(hi_new, lo_new, (long)r * 2 * H + layer * H + j, h);
    }
  }
}

// cells of the captions [c0, c1) of every column block of a tile (group = tile * ncb + cb): one warp per (column block,
// caption) pair, lane = hidden unit; NCAP pairs per pass with every partial-tile load of all of them in flight before
// the first add (MAXS bounds the slots a cell adds)
template <int MAXS, int NCAP>
__device__ __noinline__ void group_cell_tile(const GroupParams& C, int layer, int t, int tile, int c0, int c1) {
  const DecParams& P = C.dp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = P.H, R = P.R, B = P.B;
  const int ncb = C.ncb, nc = c1 - c0, npairs = ncb * nc;
  const int ns = C.nslots[layer], par = t & 1;
  const int j = tile * 32 + lane;                              // hidden unit of this lane
  const float* bi = P.bias[layer][0]; const float* ba = P.bias[layer][1]; const float* bh = P.bias[layer][2];
  float* cst = P.cx + (long)layer * R * H;
  __half* hi_new = C.hh_hi[par ^ 1]; __half* lo_new = C.hh_lo[par ^ 1];
  const float* fs = C.fslots[layer] + ((long)(tile * ncb * ns) * PK_BN) * 128 + lane;
  float bias[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) { const int n = g * H + j; bias[g] = __ldg(bi + n) + __ldg(ba + n) + __ldg(bh + n); }
#pragma unroll 1
  for (int pA = warp; pA < npairs; pA += NCAP * PK_WARPS) {
    float v[NCAP][4][MAXS], mk[NCAP], cp[NCAP], hp[NCAP];
    int rr[NCAP];
#pragma unroll
    for (int q = 0; q < NCAP; ++q) {
      const int pr = pA + q * PK_WARPS;
      const bool on = pr < npairs;
      const int pe = on ? pr : pA;
      const int cb = pe / nc, c = c0 + pe % nc;
      const float* base = fs + ((long)(cb * ns) * PK_BN + c) * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int k = 0; k < MAXS; ++k) v[q][g][k] = (on && k < ns) ? __ldcg(base + (long)k * PK_BN * 128 + g * 32) : 0.f;
      const int r = cb * PK_BN + c;
      const bool live = on && r < B;
      rr[q] = live ? r : -1;
      mk[q] = (live && t > 0) ? __ldcg(P.unfinished + r) : 1.f;
      cp[q] = live ? __ldcg(cst + (long)r * H + j) : 0.f;
      hp[q] = live ? __ldcg(P.hx + (long)r * 2 * H + layer * H + j) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < NCAP; ++q) {
      const int r = rr[q];
      if (r < 0) continue;                                     // padding rows of the 64-wide operand tiles stay zero
      float z[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < MAXS; ++k) sum += v[q][g][k];
        z[g] = sum + bias[g];
      }
      const float ig = sigmoid_fast(z[0]), fg = sigmoid_fast(z[1]), og = sigmoid_fast(z[2]), gg = tanh_fast(z[3]);
      float cn = fg * cp[q] + ig * gg;
      cn = cn * mk[q] + cp[q] * (1.f - mk[q]);
      float h = og * tanh_fast(cn);
      h = h * mk[q] + hp[q] * (1.f - mk[q]);
      cst[(long)r * H + j] = cn;
      P.hx[(long)r * 2 * H + layer * H + j] = h;
      store_split16(hi_new, lo_new, (long)r * 2 * H + layer * H + j, h);
    }
  }
}

// cell of the captions [c0, c1) of a group in the teacher-forced training loop (SAModel.py:88-111): the mask is the
// caption's seq_mask at step t, the carried state and every activation the hand-written backward reads live in the
// step-major buffers of TrainSaved (gates i,f,o,g activated in place, c / [h1|h2] of step t+1, h after dropout);
// lstm_1's token-dependent parts and all three biases were hoisted into G1s by batched GEMMs.
template <int MAXS, int NCAP>
__device__ __forceinline__ void group_cell_train(const GroupParams& C, int layer, int t, int grp, int c0, int c1) {
  const DecParams& P = C.dp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = P.H, B = P.B;
  const int ns = C.nslots[layer], par = t & 1;
  const int j = grp * 32 + lane;                               // hidden unit of this lane (one column block: group = tile)
  __half* hi_new = C.hh_hi[par ^ 1]; __half* lo_new = C.hh_lo[par ^ 1];
  const float* fs = C.fslots[layer] + ((long)(grp * ns) * PK_BN) * 128 + lane;
  float* Gs = layer == 0 ? P.G1s : P.G2s;
  float* Cs = layer == 0 ? P.C1s : P.C2s;
  const DropSpec& drop = layer == 0 ? P.drop1 : P.drop2;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (layer == 1) {
#pragma unroll
    for (int g = 0; g < 4; ++g) { const int n = g * H + j; bias[g] = __ldg(P.bias[1][0] + n) + __ldg(P.bias[1][1] + n) + __ldg(P.bias[1][2] + n); }
  }
#pragma unroll 1
  for (int cA = c0 + warp; cA < c1; cA += NCAP * PK_WARPS) {
    float v[NCAP][4][MAXS], zb[NCAP][4], mk[NCAP], cp[NCAP], hp[NCAP];
#pragma unroll
    for (int q = 0; q < NCAP; ++q) {
      const int c = cA + q * PK_WARPS;
      const bool on = c < c1;
      const float* base = fs + (long)(on ? c : cA) * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int k = 0; k < MAXS; ++k) v[q][g][k] = (on && k < ns) ? __ldcg(base + (long)k * PK_BN * 128 + g * 32) : 0.f;
      const bool live = on && c < B;
      const long tb = (long)t * B + c;
      mk[q] = live ? __ldg(P.seq_mask + (long)c * P.L + t) : 0.f;
      cp[q] = live ? __ldcg(Cs + tb * H + j) : 0.f;
      hp[q] = live ? __ldcg(P.H12s + tb * 2 * H + layer * H + j) : 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) zb[q][g] = (layer == 0 && live) ? __ldcg(Gs + tb * 4 * H + g * H + j) : bias[g];
    }
#pragma unroll
    for (int q = 0; q < NCAP; ++q) {
      const int c = cA + q * PK_WARPS;
      if (c >= c1 || c >= B) continue;                         // padding rows of the 64-wide operand tiles stay zero
      float z[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < MAXS; ++k) sum += v[q][g][k];
        z[g] = sum + zb[q][g];
      }
      const float ig = sigmoid_fast(z[0]), fg = sigmoid_fast(z[1]), og = sigmoid_fast(z[2]), gg = tanh_fast(z[3]);
      float cn = fg * cp[q] + ig * gg;
      cn = cn * mk[q] + cp[q] * (1.f - mk[q]);
      float h = og * tanh_fast(cn);
      h = h * mk[q] + hp[q] * (1.f - mk[q]);
      h *= drop.factor((uint64_t)t * B * H + (uint64_t)c * H + j);
      const long tb = (long)t * B + c;
      float* gsave = Gs + tb * 4 * H + j;
      gsave[0] = ig; gsave[H] = fg; gsave[2 * H] = og; gsave[3 * H] = gg;
      Cs[(tb + B) * H + j] = cn;
      P.H12s[(tb + B) * 2 * H + layer * H + j] = h;
      store_split16(hi_new, lo_new, (long)c * 2 * H + layer * H + j, h);
    }
  }
}

// One LSTM layer of the word step (two_inputs_lstmcell, sub_modules.py:750-770): the products of the layer as one chain
// per group member, the group's partial tiles summed and the cell applied by the members themselves.
template <int MODE>      // 0: greedy loop, 1: single step (beam search), 2: teacher-forced training loop, 3: sampling loop
__device__ __noinline__ void fused_cell_phase(const GroupParams& C, const GSched* sc, const GSched* sc_next, const CUtensorMap* maps,
                                              int layer, int t, unsigned int sync_epoch, const SmemView& sv, uint32_t tmem_base, PipeState& ps) {
  // t: word step (buffer parity, state mask); sync_epoch: how many times this group counter has been used before
  const int par = t & 1;
#ifdef GK_FINE
  long long* gf = (C.dp.dbg_clock && (t == 3 || C.topk > 0) && blockIdx.x == 0) ? C.dp.dbg_clock + (2048 + 256) * PK_STAMPS + layer * 16 : nullptr;
  if (gf && threadIdx.x == 0) gf[0] = clock64();
#define GKF(i) do { if (gf && threadIdx.x == 0) gf[i] = clock64(); } while (0)
#define GKF_T(tid, i) do { if (gf && threadIdx.x == (tid)) gf[i] = clock64(); } while (0)
#else
#define GKF(i) do { } while (0)
#define GKF_T(tid, i) do { } while (0)
#endif
  gphase<MODE == 1 || MODE == 3, MODE == 1>(C, sc, sc_next, maps, par, sv, tmem_base, ps);
  GKF(1); GKF_T(64, 2);                  // producer done / first epilogue warp done
  // a CTA is member `mem` of the groups (tile, cb) of EVERY caption column block cb of its tile (one chain per column
  // block above); the members of those groups are the same CTAs, so one counter per tile covers them all
  const int cta = blockIdx.x, m = C.members[layer], ncb = C.ncb;
  if (cta >= (C.groups / ncb) * m) return;
  const int tile = cta / m, mem = cta % m;
  __syncthreads();                       // this member's partial tiles are written (all four epilogue warps)
  GKF(3);
  if (m == 1) {
    __threadfence();                     // a group of one: its own partial tile, read back through L2 below
  } else if (threadIdx.x == 0) {
    unsigned int* ctr = C.group_ctr + layer * C.groups + tile;
    const unsigned int target = (unsigned int)m * (sync_epoch + 1u);
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    const long long t0 = clock64();
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if ((int)(v - target) >= 0) break;
      if (clock64() - t0 > 8000000000LL) __trap();
    }
  }
  GKF(4);
  __syncthreads();
  GKF(5);
  const int c0 = mem * PK_BN / m, c1 = (mem + 1) * PK_BN / m;
  if (MODE == 2) {
    if (C.nslots[layer] <= 8 && c1 - c0 > PK_WARPS) group_cell_train<8, 2>(C, layer, t, tile, c0, c1);
    else if (C.nslots[layer] <= 8) group_cell_train<8, 1>(C, layer, t, tile, c0, c1);
    else group_cell_train<GK_MAX_SLOTS, 1>(C, layer, t, tile, c0, c1);
  } else if (MODE == 3) {
    if (C.nslots[layer] <= 8 && c1 - c0 > PK_WARPS) group_cell<8, 2, true>(C, layer, t, tile, c0, c1);
    else if (C.nslots[layer] <= 8) group_cell<8, 1, true>(C, layer, t, tile, c0, c1);
    else group_cell<GK_MAX_SLOTS, 1, true>(C, layer, t, tile, c0, c1);
  } else if (MODE == 0 || ncb == 1) {
    if (C.nslots[layer] <= 8 && c1 - c0 > PK_WARPS) group_cell<8, 2>(C, layer, t, tile, c0, c1);
    else if (C.nslots[layer] <= 8) group_cell<8, 1>(C, layer, t, tile, c0, c1);
    else group_cell<GK_MAX_SLOTS, 1>(C, layer, t, tile, c0, c1);
  } else {                                // single-step mode with several column blocks: (column block, caption) pairs
    if (C.nslots[layer] <= 8) group_cell_tile<8, 2>(C, layer, t, tile, c0, c1);
    else group_cell_tile<GK_MAX_SLOTS, 2>(C, layer, t, tile, c0, c1);
  }
  GKF(6);
}

// greedy pick of caption r at step t (SAModel.py:185-210) from the per-tile partial results of the logits phase:
// global max, lowest arg-max, log-sum-exp; one warp.  Returns the raw arg-max token to every thread.
__device__ __noinline__ int dec_pick_tiles(const GroupParams& C, int r, int t, const SmemView& sv) {
  const DecParams& P = C.dp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* redi = reinterpret_cast<int*>(sv.scratch + 16);
  if (warp == 0) {
    const float4* lp = C.lpart + (long)r * C.ntv;
    float4 q[8];
    float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int tile = lane + 32 * i;
      q[i] = tile < C.ntv ? __ldcg(lp + tile) : make_float4(-INFINITY, 0.f, __int_as_float(0x7fffffff), 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int idx = __float_as_int(q[i].z);
      if (q[i].x > best || (q[i].x == best && idx < bi)) { best = q[i].x; bi = idx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += q[i].y * __expf(q[i].x - best);       // empty tiles: 0 * exp(-inf) = 0
    tot = warp_sum(tot);
    if (lane == 0) {
      float unf = (t == 0) ? 1.f : __ldcg(P.unfinished + r);
      unf = (unf != 0.f && bi > 0) ? 1.f : 0.f;
      P.unfinished[r] = unf;
      P.seq[(long)r * P.T + t] = unf != 0.f ? (int64_t)bi : 0;
      P.seqlogp[(long)r * P.T + t] = -logf(tot);
      P.tok[r] = bi;
      if (unf != 0.f) P.flags[t] = 1;
      redi[0] = bi;
    }
  }
  __syncthreads();
  const int tok = redi[0];
  __syncthreads();
  return tok;
}

// inverse-CDF draw of one vocabulary entry from p ~ exp((x - best) * invT) (warp 0, all lanes): the tile from the
// per-tile masses in tile order, the entry from the stored logits of that tile (ids ascending).  u in [0, totT).
__device__ __forceinline__ int tile_draw(const GroupParams& C, int r, const float (&mass)[8], float best, float invT, float u, int lane,
                                         float* x_sel) {
  const int ntv = C.ntv, V = C.dp.V;
  float carry = 0.f, before = 0.f;
  int sel_tile = -1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float incl = mass[i];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    const unsigned hit = __ballot_sync(0xffffffffu, lane + 32 * i < ntv && carry + incl > u);
    if (sel_tile < 0 && hit) {
      const int first = __ffs(hit) - 1;
      sel_tile = first + 32 * i;
      before = carry + __shfl_sync(0xffffffffu, incl - mass[i], first);
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (sel_tile < 0) { sel_tile = ntv - 1; before = u; }          // rounding: past the end -> last entry below
  const float4 x4 = __ldcg(reinterpret_cast<const float4*>(C.lraw + (long)r * ntv * 128 + (long)sel_tile * 128) + lane);
  const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
  float e[4], loc = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) { e[k] = sel_tile * 128 + lane * 4 + k < V ? __expf((xs[k] - best) * invT) : 0.f; loc += e[k]; }
  float incl = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  const unsigned hit = __ballot_sync(0xffffffffu, before + incl > u);
  if (hit) {
    const int first = __ffs(hit) - 1;
    float acc = before + __shfl_sync(0xffffffffu, incl - loc, first);
    float ev[4], xv[4];      // (evaluated by every lane on lane `first`'s values)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) { ev[kk] = __shfl_sync(0xffffffffu, e[kk], first); xv[kk] = __shfl_sync(0xffffffffu, xs[kk], first); }
    int k = 0;
    for (; k < 3; ++k) { acc += ev[k]; if (acc > u) break; }
    *x_sel = xv[k];
    return sel_tile * 128 + first * 4 + k;
  }
  const int last = min(V - 1, sel_tile * 128 + 127);
  *x_sel = __shfl_sync(0xffffffffu, xs[(last & 127) & 3], (last & 127) >> 2);
  return last;
}

// pick of caption r at step t in the sampling form of the loop (SAModel.py:188-196): arg-max, or one multinomial draw
// from p ~ exp(logprob / temperature) (Philox stream of the per-step kernel greedy_pick_kernel).  In the scheduled-
// sampling token pass (SAModel.py:89-99; streams of ss_pick_kernel) the token fed to step t + 1 is the ground truth or,
// with probability ss_prob, a draw from this step's distribution; nothing but the tokens is recorded.
__device__ __noinline__ int dec_sample_tiles(const GroupParams& C, int r, int t, const SmemView& sv) {
  const DecParams& P = C.dp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* redi = reinterpret_cast<int*>(sv.scratch + 16);
  if (warp == 0) {
    const int ntv = C.ntv;
    const float4* lp = C.lpart + (long)r * ntv;
    float4 q[8];
    float best = -INFINITY; int bi = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int tile = lane + 32 * i;
      q[i] = tile < ntv ? __ldcg(lp + tile) : make_float4(-INFINITY, 0.f, __int_as_float(0x7fffffff), 0.f);
      const int idx = __float_as_int(q[i].z);
      if (q[i].x > best || (q[i].x == best && idx < bi)) { best = q[i].x; bi = idx; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    float tot = 0.f, totT = 0.f, mass[8];
    const float invT = C.inv_temp;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      tot += q[i].y * __expf(q[i].x - best);
      mass[i] = lane + 32 * i < ntv ? q[i].w * __expf((q[i].x - best) * invT) : 0.f;
      totT += mass[i];
    }
    tot = warp_sum(tot);
    totT = warp_sum(totT);
    if (C.ss_mode) {
      const int i = t + 1;                                            // the step this token feeds
      int tok = 0;
      if (i < C.ss_Lp) {
        tok = (int)__ldg(C.ss_seq + (long)r * C.ss_L + i);
        const float coin = philox_uniform(C.ss_seed, 0x53530000u + (uint32_t)i, (uint64_t)r);
        if (coin < C.ss_prob) {                                       // (warp-uniform: one caption per warp)
          float xsel;
          tok = tile_draw(C, r, mass, best, invT, philox_uniform(C.ss_seed, 0x53540000u + (uint32_t)i, (uint64_t)r) * totT, lane, &xsel);
        }
      }
      if (lane == 0) {
        if (i < C.ss_Lp) { C.ss_used[(long)r * C.ss_L + i] = tok; P.unfinished[r] = __ldg(C.ss_mask + (long)r * C.ss_L + i); }
        P.tok[r] = tok;
        P.flags[t] = 1;
        redi[0] = tok;
      }
    } else {
      const float lse = best + logf(tot);
      int pick = bi;
      float pick_logp = best - lse;
      if (!C.sample_max) {
        float xsel;
        pick = tile_draw(C, r, mass, best, invT, philox_uniform(C.sample_seed, 0x5a4d0000u + (uint32_t)(t + 1), (uint64_t)r) * totT, lane, &xsel);
        pick_logp = xsel - lse;
      }
      if (lane == 0) {
        float unf = (t == 0) ? 1.f : __ldcg(P.unfinished + r);
        unf = (unf != 0.f && pick > 0) ? 1.f : 0.f;
        P.unfinished[r] = unf;
        P.seq[(long)r * P.T + t] = unf != 0.f ? (int64_t)pick : 0;
        P.seqlogp[(long)r * P.T + t] = pick_logp;
        P.tok[r] = pick;
        if (unf != 0.f) P.flags[t] = 1;
        redi[0] = pick;
      }
    }
  }
  __syncthreads();
  const int tok = redi[0];
  __syncthreads();
  return tok;
}

// next-step inputs with the training dropout of the POS gate (index = step * B * H + caption * H + unit, the layout of
// the hoisted gate GEMM of xg_train_fwd); t_in = the step that consumes these inputs
__device__ __noinline__ void dec_token_inputs_drop(const GroupParams& C, int r, int tokv, int t_in) {
  const DecParams& P = C.dp;
  const float* src = P.embed + (long)tokv * P.E;
  const float* tg = P.tgate + (long)tokv * P.H;
  const float* ps = P.pos + (long)r * P.H;
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
    if (k < P.Ep) store_split16(reinterpret_cast<__half*>(P.xt_hi), reinterpret_cast<__half*>(P.xt_lo), (long)r * P.Ep + k, x[i]);
    if (k < P.H) {
      const float gd = g[i] * C.drop_gate.factor((uint64_t)t_in * P.B * P.H + (uint64_t)r * P.H + k);
      store_split16(reinterpret_cast<__half*>(P.gp_hi), reinterpret_cast<__half*>(P.gp_lo), (long)r * P.H + k, q[i] * (1.f + gd));
    }
  }
}

// wait (without arriving) until a counter barrier has completed: CTAs that are not part of a phase
__device__ __noinline__ void grid_barrier_observe(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if ((int)(v - target) >= 0) break;
      if (clock64() - t0 > 8000000000LL) __trap();
    }
  }
  __syncthreads();
}

// Temporal attention of the rows [r0, r0 + nrows) of ONE video (beam search: the beam rows of a video share exp(2Uv)
// and V, sub_modules.py:677-680).  exp(2Uv[fb]) is copied into shared memory once and every pass of NR rows reads each
// element once for all of its rows (dec_attention re-streams the 172 KB per row); the softmax weights of all rows wait
// in shared memory until V[fb] has been fetched (over the consumed exp(2Uv) when both do not fit), then the contexts.
// Shared memory is addressed through the shared window (ld.shared with 32-bit addresses: the generic loads of
// dec_attention cost three address instructions each), the split-K slots of the query are added by float4 loads of
// the whole CTA into a staging row behind exp(2Uv), and the NV = NR x 4 frame sums of a chunk are reduced over the
// warp by recursive halving (NV + 1 shuffles instead of 5 NV).
constexpr int ATT_NR = 3;            // rows per pass in beam search (NR = 1: one caption per CTA, the word loops)
constexpr int ATT_MAX_ROWS = 8;      // rows of a video one CTA takes
__device__ __forceinline__ float lds_f32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_f32x4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// sum over the warp of NV values per lane by recursive halving; on return v[0] of lane L is the total of value
// att_reduce_owner<NV>(L) (several lanes may own the same value)
template <int NV>
__device__ __forceinline__ void att_warp_reduce(float (&v)[NV], int lane) {
  static_assert(NV == 12 || NV == 4, "att_warp_reduce: 12 or 4 values");
  if (NV == 12) {
    {   // 12 -> 6 over lanes L ^ 16
      const bool hi = lane & 16;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float keep = hi ? v[i + 6] : v[i], send = hi ? v[i] : v[i + 6];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    {   // 6 -> 3 over L ^ 8
      const bool hi = lane & 8;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float keep = hi ? v[i + 3] : v[i], send = hi ? v[i] : v[i + 3];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    {   // 3 -> 2 over L ^ 4 (the third value is summed on both sides and kept by the high half as its second)
      const bool hi = lane & 4;
      const float k0 = hi ? v[1] : v[0], s0 = hi ? v[0] : v[1];
      const float t2 = v[2] + __shfl_xor_sync(0xffffffffu, v[2], 4);
      v[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, 4);
      v[1] = t2;
    }
    {   // low half of L ^ 4 holds (value 0, value 2), high half (value 1, value 2); L ^ 2: keep one of the two
      const bool hi = lane & 2;
      const float keep = hi ? v[1] : v[0], send = hi ? v[0] : v[1];
      v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  } else {
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float keep = hi ? v[i + 2] : v[i], send = hi ? v[i] : v[i + 2];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    {
      const bool hi = lane & 8;
      const float keep = hi ? v[1] : v[0], send = hi ? v[0] : v[1];
      v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
  }
}
// which of the NV values lane L ends up with
template <int NV>
__device__ __forceinline__ int att_reduce_owner(int lane) {
  if (NV == 12) {
    const int g6 = (lane & 16) ? 6 : 0, g3 = (lane & 8) ? 3 : 0;
    const int w = (lane & 2) ? 2 : ((lane & 4) ? 1 : 0);
    return g6 + g3 + w;
  }
  return ((lane & 16) ? 2 : 0) + ((lane & 8) ? 1 : 0);
}

template <int NR, int TRAIN>      // TRAIN: step t saves ah, alpha and the context in the step-major buffers of TrainSaved
__device__ __noinline__ void dec_attention_rows(const DecParams& P, const CUtensorMap* vmap, int fb, int r0, int nrows, int t, const SmemView& sv,
                                                uint32_t& bulk_phase) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = P.K, A = P.A, H = P.H, Hh = H / 2;
#ifdef GK_FINE
  long long* ga = (P.dbg_clock && blockIdx.x == 0 && threadIdx.x == 0) ? P.dbg_clock + (2048 + 256) * PK_STAMPS + 54 : nullptr;
#define GKA(i) do { if (ga) ga[i] = clock64(); } while (0)
#else
#define GKA(i) do { } while (0)
#endif
  GKA(0);
  const uint32_t red_s = smem_u32(sv.scratch);                                       // [NR][PK_WARPS][K] per-warp partial scores
  const uint32_t sc_s = red_s + (uint32_t)(NR * PK_WARPS * K) * 4u;              // [ATT_MAX_ROWS][K] softmax weights
  const uint32_t uv_s = sv.stages_u32;
  constexpr int fpc = DEC_FPC;
  const bool v_behind = (long)K * (A + H) * 4 <= (long)PK_STAGES * PK_STAGE_BYTES;
  const uint32_t v_off = v_behind ? (uint32_t)K * A * 4u : 0u;
  // query staging rows [NR][A] at the end of the stage area (the caller checked that exp(2Uv) (+ V) end before it)
  const uint32_t ah_s = sv.stages_u32 + (uint32_t)(PK_STAGES * PK_STAGE_BYTES) - (uint32_t)(NR * A) * 4u;
  float wr[DEC_NA], wsum = 0.f;
  uint32_t aoff4[DEC_NA];
#pragma unroll
  for (int i = 0; i < DEC_NA; ++i) {
    const int a = min(threadIdx.x + PK_THREADS * i, A - 1);
    aoff4[i] = (uint32_t)a * 4u;
    wr[i] = (threadIdx.x + PK_THREADS * i < A) ? __ldg(P.w_a2w + a) : 0.f;       // out-of-range units weigh 0
    wsum += wr[i];
    wr[i] *= -2.f;
  }
  GKA(6);
  const GDesc& dah = P.d[DD_AH];
  const int A4 = A >> 2;
  constexpr int QIT = (NR * DEC_NA * PK_THREADS / 4 + PK_THREADS - 1) / PK_THREADS;      // float4 columns of NR rows per thread
  // V[fb] goes over the consumed exp(2Uv) chunks as soon as they cover it (last pass only: earlier passes re-read them)
  const int vc = max(0, (K * H + fpc * A - 1) / (fpc * A) - 1);
  bool v_issued = v_behind;
#pragma unroll 1
  for (int g = 0; g < nrows; g += NR) {
    const int ng = min(NR, nrows - g);
    // ---- query of the ng rows: bias + split-K slots in slot order, e^{2 ah} into the staging rows.  Every load is issued
    //      before the exp(2Uv) copies below: behind 172 KB per SM of bulk traffic they took ~12k cycles ----
    {
      float4 v[QIT][PK_MAX_SLOTS];
      const long sstr4 = (long)P.R * A4;
#pragma unroll
      for (int it = 0; it < QIT; ++it) {
        const int idx = threadIdx.x + it * PK_THREADS;
        const bool on = idx < ng * A4;
        const int q = on ? idx / A4 : 0, a4 = on ? idx - q * A4 : 0;
        const float4* src = reinterpret_cast<const float4*>(dah.out + (long)(r0 + g + q) * A) + a4;
#pragma unroll
        for (int k = 0; k < PK_MAX_SLOTS; ++k) v[it][k] = (on && k < dah.ns) ? __ldcg(src + k * sstr4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      GKA(7);
      if (g == 0 && warp == 0) {
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
      GKA(8);
#pragma unroll
      for (int it = 0; it < QIT; ++it) {
        const int idx = threadIdx.x + it * PK_THREADS;
        if (idx < ng * A4) {
          const int q = idx / A4, a4 = idx - q * A4;
          float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < PK_MAX_SLOTS; ++k) { sum.x += v[it][k].x; sum.y += v[it][k].y; sum.z += v[it][k].z; sum.w += v[it][k].w; }
          const float4 bb = make_float4(__ldg(P.b_h2a + a4 * 4), __ldg(P.b_h2a + a4 * 4 + 1), __ldg(P.b_h2a + a4 * 4 + 2), __ldg(P.b_h2a + a4 * 4 + 3));
          const float4 ah = make_float4(bb.x + sum.x, bb.y + sum.y, bb.z + sum.z, bb.w + sum.w);
          if (TRAIN) {
            float* o = P.AHs + ((long)t * P.B + r0 + g + q) * A + a4 * 4;
            o[0] = ah.x; o[1] = ah.y; o[2] = ah.z; o[3] = ah.w;
          }
          float4 e;      // tanh(ah + uv) = 1 - 2 / (e^{2 ah} e^{2 uv} + 1)
          e.x = __expf(2.f * fminf(fmaxf(ah.x, -40.f), 40.f)); e.y = __expf(2.f * fminf(fmaxf(ah.y, -40.f), 40.f));
          e.z = __expf(2.f * fminf(fmaxf(ah.z, -40.f), 40.f)); e.w = __expf(2.f * fminf(fmaxf(ah.w, -40.f), 40.f));
          sts_f32x4(ah_s + (uint32_t)(q * A + a4 * 4) * 4u, e);
        }
      }
    }
    GKA(9);
    __syncthreads();
    float ahr[NR][DEC_NA];
#pragma unroll
    for (int q = 0; q < NR; ++q)
#pragma unroll
      for (int i = 0; i < DEC_NA; ++i) ahr[q][i] = lds_f32(ah_s + (uint32_t)(min(q, ng - 1) * A) * 4u + aoff4[i]);
    GKA(1);
#pragma unroll 1
    for (int c = 0; c < PK_BULK_CHUNKS; ++c) {
      const int k0 = c * fpc, k1 = min(K, k0 + fpc);
      if (k0 >= k1) break;
      pk_wait(sv.bulk_bar + 8 * c, bulk_phase & 1);            // (passes at once after the first pass)
      float p[NR * DEC_FPC];
#pragma unroll
      for (int u = 0; u < NR * DEC_FPC; ++u) p[u] = wsum;
#pragma unroll
      for (int f = 0; f < DEC_FPC; ++f) {
        const uint32_t urow = uv_s + (uint32_t)(min(k0 + f, K - 1) * A) * 4u;
#pragma unroll
        for (int i = 0; i < DEC_NA; ++i) {
          const float e = lds_f32(urow + aoff4[i]);
#pragma unroll
          for (int q = 0; q < NR; ++q) p[q * DEC_FPC + f] = fmaf(wr[i], rcp_ge1(fmaf(ahr[q][i], e, 1.f)), p[q * DEC_FPC + f]);
        }
      }
      att_warp_reduce<NR * DEC_FPC>(p, lane);
      {
        const int u = att_reduce_owner<NR * DEC_FPC>(lane), q = u / DEC_FPC, f = u % DEC_FPC;
        const bool writer = NR == 3 ? ((lane & 1) == 0 && !((lane & 4) && (lane & 2))) : (lane & 7) == 0;      // one writer per value
        if (writer && k0 + f < k1)
          sts_f32(red_s + (uint32_t)((q * PK_WARPS + warp) * K + k0 + f) * 4u, p[0]);
      }
      if (!v_issued && g + NR >= nrows && c == vc && (c + 1) * fpc < K) {      // (uniform: every thread takes the branch)
        __syncthreads();
        if (threadIdx.x == 0) {
          fence_proxy_async_smem();
          pk_expect_tx(sv.bulk_bar + 8 * PK_BULK_CHUNKS, (uint32_t)K * H * 4u);
          pk_tma_2d(sv.stages_u32, vmap, sv.bulk_bar + 8 * PK_BULK_CHUNKS, 0, fb * K);
          pk_tma_2d(sv.stages_u32 + (uint32_t)K * Hh * 4u, vmap, sv.bulk_bar + 8 * PK_BULK_CHUNKS, Hh, fb * K);
        }
        v_issued = true;
      }
    }
    GKA(2);
    __syncthreads();
    if (warp < ng) {                  // warp q: softmax over ALL K frames of row g + q
      const float ba = __ldg(P.b_a2w);
      const uint32_t s = sc_s + (uint32_t)((g + warp) * K) * 4u;
      float mx = -INFINITY;
#pragma unroll 1
      for (int kk = lane; kk < K; kk += 32) {
        float qv = 0.f;
#pragma unroll
        for (int w = 0; w < PK_WARPS; ++w) qv += lds_f32(red_s + (uint32_t)((warp * PK_WARPS + w) * K + kk) * 4u);
        qv += ba;
        sts_f32(s + kk * 4u, qv);
        mx = fmaxf(mx, qv);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll 1
      for (int kk = lane; kk < K; kk += 32) { const float e = __expf(lds_f32(s + kk * 4u) - mx); sts_f32(s + kk * 4u, e); sum += e; }
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
#pragma unroll 1
      for (int kk = lane; kk < K; kk += 32) {
        const float al = lds_f32(s + kk * 4u) * inv;
        sts_f32(s + kk * 4u, al);
        if (TRAIN) P.ALPHAs[((long)t * P.B + r0 + g + warp) * K + kk] = al;
      }
    }
    __syncthreads();
  }
  GKA(3);
  if (!v_issued && threadIdx.x == 0) {        // every score is computed: V[fb] goes over exp(2Uv)
    fence_proxy_async_smem();
    pk_expect_tx(sv.bulk_bar + 8 * PK_BULK_CHUNKS, (uint32_t)K * H * 4u);
    pk_tma_2d(sv.stages_u32, vmap, sv.bulk_bar + 8 * PK_BULK_CHUNKS, 0, fb * K);
    pk_tma_2d(sv.stages_u32 + (uint32_t)K * Hh * 4u, vmap, sv.bulk_bar + 8 * PK_BULK_CHUNKS, Hh, fb * K);
  }
  pk_wait(sv.bulk_bar + 8 * PK_BULK_CHUNKS, bulk_phase & 1);
  bulk_phase++;
  GKA(4);
  const uint32_t vs = sv.stages_u32 + v_off;            // two panels [K][Hh]
#pragma unroll 1
  for (int e = threadIdx.x; e < nrows * H; e += PK_THREADS) {
    const int q = e / H, j = e - q * H;
    const uint32_t vp = vs + (uint32_t)(j >= Hh ? K * Hh + (j - Hh) : j) * 4u;
    const uint32_t s = sc_s + (uint32_t)(q * K) * 4u;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int k = 0;
#pragma unroll 2
    for (; k + 4 <= K; k += 4) {
      a0 = fmaf(lds_f32(s + k * 4u), lds_f32(vp + (uint32_t)(k * Hh) * 4u), a0);
      a1 = fmaf(lds_f32(s + (k + 1) * 4u), lds_f32(vp + (uint32_t)((k + 1) * Hh) * 4u), a1);
      a2 = fmaf(lds_f32(s + (k + 2) * 4u), lds_f32(vp + (uint32_t)((k + 2) * Hh) * 4u), a2);
      a3 = fmaf(lds_f32(s + (k + 3) * 4u), lds_f32(vp + (uint32_t)((k + 3) * Hh) * 4u), a3);
    }
    for (; k < K; ++k) a0 = fmaf(lds_f32(s + k * 4u), lds_f32(vp + (uint32_t)(k * Hh) * 4u), a0);
    const float af = (a0 + a1) + (a2 + a3);
    if (TRAIN) P.AFs[((long)t * P.B + r0 + q) * H + j] = af;
    store_split16(reinterpret_cast<__half*>(P.af_hi), reinterpret_cast<__half*>(P.af_lo), (long)(r0 + q) * H + j, af);
  }
  GKA(5);
  __syncthreads();
}

// the greedy word loop of SAModel.sample (SAModel.py:182-219), grouped-cell form
template <int SAMPLE>      // 1: sampling form (multinomial draw and / or training dropout; the logits are stored for the draw)
__global__ void __launch_bounds__(PK_THREADS, 1)
decode_grouped_kernel(const GroupParams* __restrict__ Cp, const __grid_constant__ MapTable2 maps) {
  __shared__ GroupParams Csm;
  __shared__ GSched s_sched[3];
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(GroupParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Csm)[i] = reinterpret_cast<const uint32_t*>(Cp)[i];
  __syncthreads();
  const GroupParams& C = Csm;
  const DecParams& P = C.dp;
  for (int i = threadIdx.x; i < (int)(3 * sizeof(GSched) / 4); i += PK_THREADS) {
    const int ph = i / (int)(sizeof(GSched) / 4), w = i % (int)(sizeof(GSched) / 4);
    reinterpret_cast<uint32_t*>(&s_sched[ph])[w] = reinterpret_cast<const uint32_t*>(C.gsched + (long)ph * G + cta)[w];
  }
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = P.H, R = P.R, B = P.B, T = P.T;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 27) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = 0;
  uint32_t bulk_phase = 0;

  // ---- prologue: states into buffer 0, <bos> inputs, bookkeeping ----
  for (int e = cta * PK_THREADS + threadIdx.x; e < R * H; e += G * PK_THREADS) {
    const int r = e / H, j = e % H;
    float h1 = 0.f, h2 = 0.f;
    if (r < B) {
      h1 = P.state0[0][e]; h2 = P.state0[2][e];
      P.cx[e] = P.state0[1][e]; P.cx[(long)R * H + e] = P.state0[3][e];
    }
    P.hx[(long)r * 2 * H + j] = h1; P.hx[(long)r * 2 * H + H + j] = h2;
    store_split16(C.hh_hi[0], C.hh_lo[0], (long)r * 2 * H + j, h1);
    store_split16(C.hh_hi[0], C.hh_lo[0], (long)r * 2 * H + H + j, h2);
    if (r >= B) {                                                   // padding rows of the second buffer
      const __half z = __float2half_rn(0.f);
      C.hh_hi[1][(long)r * 2 * H + j] = z; C.hh_lo[1][(long)r * 2 * H + j] = z;
      C.hh_hi[1][(long)r * 2 * H + H + j] = z; C.hh_lo[1][(long)r * 2 * H + H + j] = z;
    }
  }
  for (int r = cta; r < R; r += G) {
    if (r < B) {
      if (SAMPLE) dec_token_inputs_drop(C, r, C.ss_mode ? (int)__ldg(C.ss_seq + (long)r * C.ss_L) : 0, 0);      // step 0 of the token pass: ground truth
      else dec_token_inputs(P, r, 0);                 // token 0 = <bos> (SAModel.py:184)
    } else {
      const __half z = __float2half_rn(0.f);
      __half* xh = reinterpret_cast<__half*>(P.xt_hi); __half* xl = reinterpret_cast<__half*>(P.xt_lo);
      __half* gh = reinterpret_cast<__half*>(P.gp_hi); __half* gl = reinterpret_cast<__half*>(P.gp_lo);
      __half* ah = reinterpret_cast<__half*>(P.af_hi); __half* al = reinterpret_cast<__half*>(P.af_lo);
      for (int k = threadIdx.x; k < P.Ep; k += PK_THREADS) { xh[(long)r * P.Ep + k] = z; xl[(long)r * P.Ep + k] = z; }
      for (int j = threadIdx.x; j < H; j += PK_THREADS) {
        gh[(long)r * H + j] = z; gl[(long)r * H + j] = z;
        ah[(long)r * H + j] = z; al[(long)r * H + j] = z;
      }
    }
    if (threadIdx.x == 0) { P.unfinished[r] = (SAMPLE && C.ss_mode && r < B) ? __ldg(C.ss_mask + (long)r * C.ss_L) : 1.f; P.tok[r] = 0; }
  }
  {   // EUv = exp(2 Uv), clamped like the per-step factor
    const long n = (long)B * P.K * P.A;
    for (long e = (long)cta * PK_THREADS + threadIdx.x; e < n; e += (long)G * PK_THREADS)
      P.EUv[e] = __expf(2.f * fminf(fmaxf(__ldg(P.Uv + e), -40.f), 40.f));
  }
  // attention query of step 0: the G4 schedule once on the initial state (its logits are ignored); the state sits in
  // buffer 0, which is "the buffer being written" of an odd step
  const int n_att = C.n_att;
  const bool is_att = cta >= G - n_att;             // attention CTA: caption cta - (G - n_att)
  unsigned int pick_target = 0;
  gprefetch(C, &s_sched[2], maps.m, sv, ps);
  grid_barrier(P.sync_counter, sync_target, G);
  gphase<SAMPLE != 0>(C, &s_sched[2], nullptr, maps.m, 1, sv, tmem_base, ps);
  grid_barrier(P.sync_counter, sync_target, G);
  if (is_att) dec_attention_rows<1, 0>(P, &maps.m[GM_V], cta - (G - n_att), cta - (G - n_att), 1, 0, sv, bulk_phase);
  fence_proxy_async_smem();
  gprefetch(C, &s_sched[0], maps.m, sv, ps);
  grid_barrier(P.sync_counter, sync_target, G);
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    pk_stamp(P.dbg_clock, cta, t, 0);
    // ===== F1: lstm_1 = cell(W_i2h1.xt + W_a2h1.gp + W_h2h1.h1)   [attention CTAs: still busy with step t's attention] =====
    fused_cell_phase<SAMPLE ? 3 : 0>(C, &s_sched[0], &s_sched[1], maps.m, 0, t, (unsigned int)t, sv, tmem_base, ps);
    gprefetch(C, &s_sched[1], maps.m, sv, ps);
    pk_stamp(P.dbg_clock, cta, t, 1);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, t, 2);
    // ===== F3: lstm_2 = cell(W_i2h2.h1' + W_a2h2.af + W_h2h2.h2) =====
    fused_cell_phase<SAMPLE ? 3 : 0>(C, &s_sched[1], &s_sched[2], maps.m, 1, t, (unsigned int)t, sv, tmem_base, ps);
    gprefetch(C, &s_sched[2], maps.m, sv, ps);
    pk_stamp(P.dbg_clock, cta, t, 3);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, t, 4);
    // ===== G4: logits of step t  +  attention query of step t+1 (both read the buffer just written) =====
    gphase<SAMPLE != 0>(C, &s_sched[2], &s_sched[0], maps.m, t & 1, sv, tmem_base, ps);
    pk_stamp(P.dbg_clock, cta, t, 5);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, t, 6);
    // ===== P4: greedy pick + next-step inputs on the first B CTAs.  The attention of step t+1 starts here on the last
    // n_att CTAs and runs on through F1 of step t+1 (it needs the query, not the token; its result feeds F3): those
    // CTAs are not part of the pick barrier, they only observe it (for the early-exit flag) when they are done. =====
    pick_target += (unsigned int)(G - n_att);
    if (is_att) {
      if (t + 1 < T) dec_attention_rows<1, 0>(P, &maps.m[GM_V], cta - (G - n_att), cta - (G - n_att), 1, t + 1, sv, bulk_phase);
      fence_proxy_async_smem();
      pk_stamp(P.dbg_clock, cta, t, 7);
      grid_barrier_observe(C.pick_ctr, pick_target);
    } else {
#pragma unroll 1
      for (int r = cta; r < B; r += G - n_att) {
        const int tokv = SAMPLE ? dec_sample_tiles(C, r, t, sv) : dec_pick_tiles(C, r, t, sv);
        if (SAMPLE) dec_token_inputs_drop(C, r, tokv, t + 1);
        else dec_token_inputs(P, r, tokv);
        __syncthreads();
      }
      if (t + 1 < T) gprefetch(C, &s_sched[0], maps.m, sv, ps);
      pk_stamp(P.dbg_clock, cta, t, 7);
      unsigned int tgt = pick_target - (unsigned int)(G - n_att);
      grid_barrier(C.pick_ctr, tgt, G - n_att);
    }
    pk_stamp(P.dbg_clock, cta, t, 8);
    if (__ldcg(P.flags + t) == 0) break;     // every caption finished (SAModel.py:206)
  }
  gemm_prefetch_drain(sv, ps);               // early exit with weight tiles in flight
  pipeline_teardown(tmem_base);
}

}  // namespace xg

namespace xg {

// ====================================================================================
// the teacher-forced word loop of SAModel.forward (SAModel.py:88-111), grouped-cell form.  The ground-truth tokens
// are known, so the token-dependent parts of lstm_1 are hoisted (batched GEMMs into G1s) and a step is THREE grid
// synchronised phases:
//   A {attention query W_h2a.[h1|h2]: split-K slots; the recurrent product W_h2h2.h2 of lstm_2 into its early slots}
//   B {temporal attention on the last B CTAs  ||  lstm_1 = cell(G1s[t] + W_h2h1.h1) as a fused group phase on the others}
//   C {lstm_2 = cell(W_i2h2.h1' + W_a2h2.af + W_h2h2.h2) as a fused group phase on all CTAs}
// (decode_persistent_kernel<1>, xg_persist.cuh: four phases, cells as separate pointwise phases, tf32 operand pairs
// split in the kernel.)  The logit / classifier heads run batched over all steps after the loop.
// ====================================================================================
__global__ void __launch_bounds__(PK_THREADS, 1)
train_grouped_kernel(const GroupParams* __restrict__ Cp, const __grid_constant__ MapTable2 maps) {
  __shared__ GroupParams Csm;
  __shared__ GSched s_sched[3];
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(GroupParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Csm)[i] = reinterpret_cast<const uint32_t*>(Cp)[i];
  __syncthreads();
  const GroupParams& C = Csm;
  const DecParams& P = C.dp;
  for (int i = threadIdx.x; i < (int)(3 * sizeof(GSched) / 4); i += PK_THREADS) {
    const int ph = i / (int)(sizeof(GSched) / 4), w = i % (int)(sizeof(GSched) / 4);
    reinterpret_cast<uint32_t*>(&s_sched[ph])[w] = reinterpret_cast<const uint32_t*>(C.gsched + (long)ph * G + cta)[w];
  }
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = P.H, R = P.R, B = P.B, T = P.T;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 27) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = 0;
  uint32_t bulk_phase = 0;
  const int n_att = C.n_att;
  const bool is_att = cta >= G - n_att;             // attention CTA: caption cta - (G - n_att)

  // ---- prologue: [h1|h2] of step 0 (init_hidden wrote H12s[0]) as fp16 pairs, zero padding rows, exp(2 Uv) ----
  for (int e = cta * PK_THREADS + threadIdx.x; e < R * 2 * H; e += G * PK_THREADS) {
    const int r = e / (2 * H);
    const float h = r < B ? P.H12s[e] : 0.f;
    store_split16(C.hh_hi[0], C.hh_lo[0], e, h);
    if (r >= B) { const __half z = __float2half_rn(0.f); C.hh_hi[1][e] = z; C.hh_lo[1][e] = z; }
  }
  for (int e = cta * PK_THREADS + threadIdx.x; e < (R - B) * H; e += G * PK_THREADS) {
    const __half z = __float2half_rn(0.f);
    reinterpret_cast<__half*>(P.af_hi)[(long)B * H + e] = z; reinterpret_cast<__half*>(P.af_lo)[(long)B * H + e] = z;
  }
  {
    const long n = (long)B * P.K * P.A;
    for (long e = (long)cta * PK_THREADS + threadIdx.x; e < n; e += (long)G * PK_THREADS)
      P.EUv[e] = __expf(2.f * fminf(fmaxf(__ldg(P.Uv + e), -40.f), 40.f));
  }
  gprefetch(C, &s_sched[0], maps.m, sv, ps);
  grid_barrier(P.sync_counter, sync_target, G);
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    pk_stamp(P.dbg_clock, cta, t, 0);
    // ===== A: attention query of step t =====
    gphase<false>(C, &s_sched[0], &s_sched[1], maps.m, t & 1, sv, tmem_base, ps);
    if (!is_att) gprefetch(C, &s_sched[1], maps.m, sv, ps);
    pk_stamp(P.dbg_clock, cta, t, 1);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, t, 2);
    // ===== B: attention || lstm_1 (+ the recurrent product of lstm_2) =====
    if (is_att) {
      dec_attention_rows<1, 1>(P, &maps.m[GM_V], cta - (G - n_att), cta - (G - n_att), 1, t, sv, bulk_phase);
      fence_proxy_async_smem();
    } else {
      fused_cell_phase<2>(C, &s_sched[1], &s_sched[2], maps.m, 0, t, (unsigned int)t, sv, tmem_base, ps);
    }
    gprefetch(C, &s_sched[2], maps.m, sv, ps);
    pk_stamp(P.dbg_clock, cta, t, 3);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, t, 4);
    // ===== C: lstm_2 =====
    fused_cell_phase<2>(C, &s_sched[2], &s_sched[0], maps.m, 1, t, (unsigned int)t, sv, tmem_base, ps);
    if (t + 1 < T) gprefetch(C, &s_sched[0], maps.m, sv, ps);
    pk_stamp(P.dbg_clock, cta, t, 5);
    grid_barrier(P.sync_counter, sync_target, G);
    pk_stamp(P.dbg_clock, cta, t, 6);
  }
  pipeline_teardown(tmem_base);
}

}  // namespace xg

namespace xg {

// ====================================================================================
// ONE word step for arbitrary state rows (beam search: CaptionModel.py:121-125 -> SAModel.get_logprobs_state,
// SAModel.py:117-127), grouped form.  Rows = videos x beam; every feat_div consecutive rows share V / exp(2Uv) / pos.
//   prologue {gather parent states, token inputs} | A {lstm_1 fused + attention query} | B {attention} | C {lstm_2 fused}
//   | D {logits tiles: max / sum-exp / topk per (row, tile) in the epilogue} | E {per row: log-sum-exp, topk merge, states out}
// The log-probs never exist as a (rows, V) matrix: a row's candidates are the topk of each of its vocabulary tiles.
// ====================================================================================
constexpr unsigned int GK_STEP_BARRIERS = 5;

// row r (one warp): log-sum-exp from the per-tile results, then the topk (log-prob, id) of the row with the UNK penalty of
// CaptionModel.py:94, value descending, lowest id first (the order of a stable descending sort, CaptionModel.py:39-40).
// An entry of the topk can only sit in one of the topk best vocabulary tiles (ordered by their max, lowest arg-max first)
// - or in tile 0, whose max does not know about the penalty on id 1 and which is therefore always scanned.
constexpr int GK_SURV = 64;           // survivors of the threshold test a warp ranks in shared memory (more: the serial selection)
__device__ __forceinline__ void step_row_merge(const GroupParams& C, int r, float2* surv /* [GK_SURV], this warp's */) {
  const DecParams& P = C.dp;
  const int lane = threadIdx.x & 31;
  const int ntv = C.ntv, topk = C.topk;
  const float4* lp = C.lpart + (long)r * ntv;
  float4 q[8];
  float best = -INFINITY;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int tile = lane + 32 * i;
    q[i] = tile < ntv ? __ldcg(lp + tile) : make_float4(-INFINITY, 0.f, __int_as_float(0x7fffffff), 0.f);
    best = fmaxf(best, q[i].x);
  }
  best = warp_max(best);
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += q[i].y * __expf(q[i].x - best);
  tot = warp_sum(tot);
  const float lse = best + logf(tot);
  // ---- the tiles to scan: tile 0 + the topk best of the others ----
  const float* lr = C.lraw + (long)r * ntv * 128;
  float4 cand[GK_TOPK + 1];
  int ctile[GK_TOPK + 1];
  cand[0] = __ldcg(reinterpret_cast<const float4*>(lr) + lane);
  ctile[0] = 0;
  if (lane == 0) cand[0].y -= 1000.f;                // id 1 = UNK
  unsigned used = lane == 0 ? 1u : 0u;               // bit i: tile lane + 32 i is taken (tile 0 from the start)
  float thr = -INFINITY; bool thr_ok = true;         // max of the topk-th best tile (when there are topk tiles besides tile 0)
#pragma unroll
  for (int c = 1; c <= GK_TOPK; ++c) {
    cand[c] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    ctile[c] = 0;
    if (c <= topk) {                                 // warp-uniform
      float bv = -INFINITY; int bid = 0x7fffffff, bsl = -1;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int id = __float_as_int(q[i].z);
        if (!((used >> i) & 1u) && lane + 32 * i < ntv && (q[i].x > bv || (q[i].x == bv && id < bid))) { bv = q[i].x; bid = id; bsl = i; }
      }
      float wv = bv; int wid = bid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, wid, o);
        if (ov > wv || (ov == wv && oi < wid)) { wv = ov; wid = oi; }
      }
      if (wid != 0x7fffffff) {                       // (ids are unique: exactly one lane owns the winner)
        if (bsl >= 0 && wid == bid) used |= 1u << bsl;
        ctile[c] = wid >> 7;
        cand[c] = __ldcg(reinterpret_cast<const float4*>(lr + (long)ctile[c] * 128) + lane);
        if (c == topk) thr = wv;
      } else {
        thr_ok = false;
      }
    }
  }
  // ---- topk of the scanned entries.  Each of the topk tiles holds an entry >= thr, so every entry of the row's topk is
  // >= thr: the few survivors of that test are compacted into shared memory and ranked against each other (value
  // descending, lowest id first).  One warp scanning 36 entries per lane five times, with ten shuffles per round, took
  // ~9 us of the 14 us this function cost. ----
  if (thr_ok) {
    int nsurv = 0;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int u = 0; u <= GK_TOPK; ++u) {
      if (u <= topk) {                                 // warp-uniform
        const float xs[4] = {cand[u].x, cand[u].y, cand[u].z, cand[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool ok = xs[e] >= thr;
          const unsigned bal = __ballot_sync(0xffffffffu, ok);
          const int pos = nsurv + __popc(bal & lt);
          if (ok && pos < GK_SURV) surv[pos] = make_float2(xs[e], __int_as_float(ctile[u] * 128 + lane * 4 + e));
          nsurv += __popc(bal);
        }
      }
    }
    __syncwarp();
    if (nsurv <= GK_SURV) {
#pragma unroll 1
      for (int n = lane; n < nsurv; n += 32) {
        const float2 me = surv[n];
        const int my_id = __float_as_int(me.y);
        int rank = 0;
#pragma unroll 1
        for (int m = 0; m < nsurv; ++m) {
          const float2 o = surv[m];
          rank += (o.x > me.x || (o.x == me.x && __float_as_int(o.y) < my_id)) ? 1 : 0;
        }
        if (rank < topk) {
          // the selection value carries the UNK penalty; the log-prob is value - lse (CaptionModel.py:94 subtracts 1000 from the log-prob)
          P.ys_out[(long)r * topk + rank] = me.x - lse;
          P.ix_out[(long)r * topk + rank] = my_id;
        }
      }
      __syncwarp();
      return;
    }
    __syncwarp();
  }
  // ---- the same selection, serially (vocabularies with fewer than topk + 1 tiles, or a flat row with more survivors) ----
  unsigned long long taken = 0ull;
#pragma unroll 1
  for (int c = 0; c < topk; ++c) {
    float bv = -INFINITY; int bid = 0x7fffffff, bpos = -1;
#pragma unroll
    for (int u = 0; u <= GK_TOPK; ++u) {
      const float xs[4] = {cand[u].x, cand[u].y, cand[u].z, cand[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int id = ctile[u] * 128 + lane * 4 + e;
        if (!((taken >> (u * 4 + e)) & 1ull) && (xs[e] > bv || (xs[e] == bv && id < bid))) { bv = xs[e]; bid = id; bpos = u * 4 + e; }
      }
    }
    float wv = bv; int wid = bid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wid, o);
      if (ov > wv || (ov == wv && oi < wid)) { wv = ov; wid = oi; }
    }
    if (bpos >= 0 && wid == bid && wv == bv) taken |= 1ull << bpos;
    if (lane == 0) {
      // the selection value carries the UNK penalty; the log-prob is value - lse (CaptionModel.py:94 subtracts 1000 from the log-prob)
      P.ys_out[(long)r * topk + c] = wv - lse;
      P.ix_out[(long)r * topk + c] = wid;
    }
  }
}

__global__ void __launch_bounds__(PK_THREADS, 1)
decode_step_grouped_kernel(const GroupParams* __restrict__ Cp, const __grid_constant__ MapTable2 maps, unsigned int sync_base,
                           unsigned int epoch0, int nsteps) {
  __shared__ GroupParams Csm;
  __shared__ GSched s_sched[3];
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(GroupParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Csm)[i] = reinterpret_cast<const uint32_t*>(Cp)[i];
  __syncthreads();
  const GroupParams& C = Csm;
  const DecParams& P = C.dp;
  for (int i = threadIdx.x; i < (int)(3 * sizeof(GSched) / 4); i += PK_THREADS) {
    const int ph = i / (int)(sizeof(GSched) / 4), w = i % (int)(sizeof(GSched) / 4);
    reinterpret_cast<uint32_t*>(&s_sched[ph])[w] = reinterpret_cast<const uint32_t*>(C.gsched + (long)ph * G + cta)[w];
  }
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = P.H, R = P.R, B = P.B;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 27) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = sync_base;
  uint32_t bulk_phase = 0;

  // nsteps > 1 (beam search with the candidate merge in the kernel): the word steps of a whole search in one launch -
  // everything a step hands to the next one (states, tokens, parents, beams) lives in global memory and is read
  // through L2 (__ldcg) after the grid barrier that closes the step
#pragma unroll 1
  for (int ks = 0; ks < nsteps; ++ks) {
  const unsigned int epoch = epoch0 + (unsigned int)ks;
  const int* parent_in = ks == 0 ? P.parent_in : C.mg.parent;
  pk_stamp(P.dbg_clock, cta, 3, 0);
  int tok_pre[4];                                   // (this CTA's first tokens: one round trip instead of one per row)
#pragma unroll
  for (int i = 0; i < 4; ++i) tok_pre[i] = cta + i * G < B ? (int)__ldcg(P.tokens_in + cta + i * G) : 0;
  // ---- prologue: parent states (beam reordering, CaptionModel.py:62-64) into the working buffers, token inputs ----
  // (four elements per thread and pass, the parent rows first, then every state load, then the stores: one element at a
  //  time cost two dependent L2 round trips per element - the stores may alias the loads as far as the compiler knows)
  for (int e0 = cta * PK_THREADS + threadIdx.x; e0 < R * H; e0 += 4 * G * PK_THREADS) {
    long src[4]; float h1[4], h2[4], c1[4], c2[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * G * PK_THREADS, r = e / H, j = e - r * H;
      src[u] = (e < R * H && r < B) ? (parent_in ? (long)__ldcg(parent_in + r) * H + j : (long)e) : -1L;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      h1[u] = h2[u] = c1[u] = c2[u] = 0.f;
      if (src[u] >= 0) {
        h1[u] = __ldcg(P.state0[0] + src[u]); h2[u] = __ldcg(P.state0[2] + src[u]);
        c1[u] = __ldcg(P.state0[1] + src[u]); c2[u] = __ldcg(P.state0[3] + src[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * G * PK_THREADS;
      if (e < R * H) {
        const int r = e / H, j = e - r * H;
        if (r < B) { P.cx[e] = c1[u]; P.cx[(long)R * H + e] = c2[u]; }
        P.hx[(long)r * 2 * H + j] = h1[u]; P.hx[(long)r * 2 * H + H + j] = h2[u];
        store_split16(C.hh_hi[0], C.hh_lo[0], (long)r * 2 * H + j, h1[u]);
        store_split16(C.hh_hi[0], C.hh_lo[0], (long)r * 2 * H + H + j, h2[u]);
        if (r >= B) {
          const __half z = __float2half_rn(0.f);
          C.hh_hi[1][(long)r * 2 * H + j] = z; C.hh_lo[1][(long)r * 2 * H + j] = z;
          C.hh_hi[1][(long)r * 2 * H + H + j] = z; C.hh_lo[1][(long)r * 2 * H + H + j] = z;
        }
      }
    }
  }
  for (int r = cta, ri = 0; r < R; r += G, ++ri) {
    if (r < B) {
      dec_token_inputs(P, r, ri == 0 ? tok_pre[0] : ri == 1 ? tok_pre[1] : ri == 2 ? tok_pre[2] : ri == 3 ? tok_pre[3] : (int)__ldcg(P.tokens_in + r));
    } else {
      const __half z = __float2half_rn(0.f);
      __half* xh = reinterpret_cast<__half*>(P.xt_hi); __half* xl = reinterpret_cast<__half*>(P.xt_lo);
      __half* gh = reinterpret_cast<__half*>(P.gp_hi); __half* gl = reinterpret_cast<__half*>(P.gp_lo);
      __half* ah = reinterpret_cast<__half*>(P.af_hi); __half* al = reinterpret_cast<__half*>(P.af_lo);
      for (int k = threadIdx.x; k < P.Ep; k += PK_THREADS) { xh[(long)r * P.Ep + k] = z; xl[(long)r * P.Ep + k] = z; }
      for (int j = threadIdx.x; j < H; j += PK_THREADS) {
        gh[(long)r * H + j] = z; gl[(long)r * H + j] = z;
        ah[(long)r * H + j] = z; al[(long)r * H + j] = z;
      }
    }
  }
  if (P.build_euv && ks == 0) {
    const long n = (long)((B + P.feat_div - 1) / P.feat_div) * P.K * P.A;
    for (long e = (long)cta * PK_THREADS + threadIdx.x; e < n; e += (long)G * PK_THREADS)
      P.EUv[e] = __expf(2.f * fminf(fmaxf(__ldg(P.Uv + e), -40.f), 40.f));
  }
  pk_stamp(P.dbg_clock, cta, 3, 1);
  grid_barrier(P.sync_counter, sync_target, G);
  pk_stamp(P.dbg_clock, cta, 3, 2);
  // ===== A: lstm_1 (x, gp, h1 chains, fused cell)  +  attention query W_h2a.[h1|h2] (split-K slots) =====
  fused_cell_phase<1>(C, &s_sched[0], &s_sched[1], maps.m, 0, 0, epoch, sv, tmem_base, ps);
  if (P.dbg_clock) __syncthreads();          // (trace runs: the stamp is the CTA's last warp, not the producer warp)
  pk_stamp(P.dbg_clock, cta, 3, 3);
  grid_barrier(P.sync_counter, sync_target, G);
  pk_stamp(P.dbg_clock, cta, 3, 4);
  // ===== B: temporal attention of every row =====
  {   // a video's rows go to cpv CTAs, each takes a run of them (rows of one video share exp(2Uv) / V)
    const int fdiv = P.feat_div, nvid = (B + fdiv - 1) / fdiv;
    int cpv = max(1, min(fdiv, G / nvid));
    while ((fdiv + cpv - 1) / cpv > ATT_MAX_ROWS) ++cpv;          // (fdiv <= XG_MAX_BEAM = 16: at most two CTAs more)
    const int rpc = (fdiv + cpv - 1) / cpv;
#pragma unroll 1
    for (int u = cta; u < nvid * cpv; u += G) {
      const int vid = u / cpv, part = u % cpv;
      const int r0 = vid * fdiv + part * rpc, r1 = min(min(r0 + rpc, (vid + 1) * fdiv), B);
      if (r0 < r1) dec_attention_rows<ATT_NR, 0>(P, &maps.m[GM_V], vid, r0, r1 - r0, 0, sv, bulk_phase);
    }
  }
  fence_proxy_async_smem();
  pk_stamp(P.dbg_clock, cta, 3, 5);
  grid_barrier(P.sync_counter, sync_target, G);
  pk_stamp(P.dbg_clock, cta, 3, 6);
  // ===== C: lstm_2 (h1', af, h2 chains, fused cell) =====
  fused_cell_phase<1>(C, &s_sched[1], &s_sched[2], maps.m, 1, 0, epoch, sv, tmem_base, ps);
  if (P.dbg_clock) __syncthreads();          // (trace runs: the stamp is the CTA's last warp, not the producer warp)
  pk_stamp(P.dbg_clock, cta, 3, 7);
  grid_barrier(P.sync_counter, sync_target, G);
  pk_stamp(P.dbg_clock, cta, 3, 8);
  // ===== D: logits tiles (reduced in the epilogue) =====
  gphase<true, true>(C, &s_sched[2], nullptr, maps.m, 0, sv, tmem_base, ps);
  if (P.dbg_clock) __syncthreads();          // (trace runs: the stamp is the CTA's last warp, not the producer warp)
  pk_stamp(P.dbg_clock, cta, 3, 9);
  grid_barrier(P.sync_counter, sync_target, G);
  pk_stamp(P.dbg_clock, cta, 3, 10);
  // ===== E: per row log-sum-exp + topk merge; new states out (every read of the old states happened in the prologue) =====
  float2* surv_all = reinterpret_cast<float2*>(sv.scratch + 768);      // [PK_WARPS][GK_SURV] (behind the candidate merge's cp / sel)
  if (C.mg_on) {
    // a CTA takes a video: one warp per beam row for the row merge, then warp 0 merges the video's candidates and does
    // the bookkeeping of the position (what used to be a kernel of its own after every word step)
    double* cp = reinterpret_cast<double*>(sv.scratch);
    int* sel = reinterpret_cast<int*>(sv.scratch + 2 * BM_MAX_BEAM * BM_MAX_BEAM);
    const int beam = C.mg.beam, wq = (int)(threadIdx.x >> 5);
#pragma unroll 1
    for (int k = cta; k < C.mg.B; k += G) {
      if (wq < beam && k * beam + wq < B) step_row_merge(C, k * beam + wq, surv_all + wq * GK_SURV);
      __syncthreads();            // the rows' candidates (ys_out / ix_out) are written
      pk_stamp(P.dbg_clock, cta, 3, 12);
      if (wq == 0) {
        BeamMergeIO m = C.mg;                       // position t + ks; the beam buffers alternate
        m.t += ks;
        if (ks & 1) { m.seq_in = C.mg.seq_out; m.lps_in = C.mg.lps_out; m.seq_out = const_cast<int64_t*>(C.mg.seq_in); m.lps_out = const_cast<float*>(C.mg.lps_in); }
        beam_merge_warp(m, k, cp, sel, (int)(threadIdx.x & 31));
      }
      __syncthreads();
      pk_stamp(P.dbg_clock, cta, 3, 13);
    }
  } else {
#pragma unroll 1
    for (int r = cta + G * (int)(threadIdx.x >> 5); r < B; r += G * PK_WARPS) step_row_merge(C, r, surv_all + (threadIdx.x >> 5) * GK_SURV);      // one warp per row
  }
  // the states go out on the CTAs that have no video to merge (when there are any: the merge is the long pole of this phase)
  const int nv_ctas = (C.mg_on && C.mg.B < G) ? C.mg.B : 0, ncopy = G - nv_ctas;
  for (int e0 = (cta - nv_ctas) * PK_THREADS + threadIdx.x; cta >= nv_ctas && e0 < B * H; e0 += 4 * ncopy * PK_THREADS) {      // (loads first, as in the prologue)
    float v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * ncopy * PK_THREADS;
      if (e < B * H) {
        const int r = e / H, j = e - r * H;
        v[u][0] = __ldcg(P.hx + (long)r * 2 * H + j); v[u][2] = __ldcg(P.hx + (long)r * 2 * H + H + j);
        v[u][1] = __ldcg(P.cx + e); v[u][3] = __ldcg(P.cx + (long)R * H + e);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * ncopy * PK_THREADS;
      if (e < B * H) {
#pragma unroll
        for (int q = 0; q < 4; ++q) P.state_out[q][e] = v[u][q];
      }
    }
  }
  pk_stamp(P.dbg_clock, cta, 3, 11);
  if (ks + 1 < nsteps) grid_barrier(P.sync_counter, sync_target, G);      // states, tokens, parents of the next step are out
  }
  pipeline_teardown(tmem_base);
}

}  // namespace xg

namespace xg {

// ====================================================================================
// encoder frame recurrence (both streams, sub_modules.py:132-147), grouped form with STATIONARY weights:
//   tiles = (stream) x (32 hidden units x 4 gates of nn.LSTMCell: i,f,g,o); a tile's K = H extent is cut over the members
//   of its group (148 / 32 = 4: two 64-wide k-blocks each), whose fp16 hi / lo weight tiles are loaded into the pipeline
//   stages ONCE per launch; a frame step only streams the new [h_rgb | h_opfl] operand tiles (16 KB per k-block), adds
//   the group's partial tiles and runs the cell.  One grid barrier per frame.
// ====================================================================================
// cell of nn.LSTMCell (gate order i,f,g,o) + frame mask for the captions [c0, c1) of group (stream s, unit tile)
template <int MAXS>
__device__ __forceinline__ void enc_group_cell(const GroupParams& C, const EncParams& E, int t, int grp, int ns, int c0, int c1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = E.H, B = E.B, K = E.K;
  const int ntile = H / 32;
  const int s = (grp / C.ncb) / ntile, tile = (grp / C.ncb) % ntile, cb = grp % C.ncb;
  const int j = tile * 32 + lane;
  const float* fs = C.fslots[0] + ((long)(grp * ns) * PK_BN) * 128 + lane;
  __half* hi_new = C.hh_hi[(t & 1) ^ 1]; __half* lo_new = C.hh_lo[(t & 1) ^ 1];
#pragma unroll 1
  for (int cA = c0 + warp; cA < c1; cA += 2 * PK_WARPS) {
    float v[2][4][MAXS], zz[2][4], mk[2], cp[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = cA + q * PK_WARPS, b = cb * PK_BN + c;
      const bool live = c < c1 && b < B;
      const float* base = fs + (long)(c < c1 ? c : cA) * 128;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int k = 0; k < MAXS; ++k) v[q][g][k] = (live && t > 0 && k < ns) ? __ldcg(base + (long)k * PK_BN * 128 + g * 32) : 0.f;
        zz[q][g] = live ? E.Gt[s][((long)t * B + b) * 4 * H + g * H + j] : 0.f;
      }
      mk[q] = live ? __ldg(E.fmask + (long)b * K + t) : 0.f;
      cp[q] = (live && t > 0) ? E.Cs[s][((long)(t - 1) * B + b) * H + j] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = cA + q * PK_WARPS, b = cb * PK_BN + c;
      if (c >= c1 || b >= B) continue;
      float z[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < MAXS; ++k) sum += v[q][g][k];
        z[g] = zz[q][g] + sum;
      }
      const float ig = sigmoid_fast(z[0]), fg = sigmoid_fast(z[1]), gg = tanh_fast(z[2]), og = sigmoid_fast(z[3]);
      const float c2 = fg * cp[q] + ig * gg;
      const float h = og * tanh_fast(c2) * mk[q];      // h' *= mask (sub_modules.py:139,146)
      const float cn = c2 * mk[q];                     // c' *= mask (:140,147)
      float* zo = E.Gt[s] + ((long)t * B + b) * 4 * H + j;
      zo[0] = ig; zo[H] = fg; zo[2 * H] = gg; zo[3 * H] = og;
      const long o = ((long)t * B + b) * H + j;
      E.Cs[s][o] = cn;
      E.Hs[s][o] = h;
      if (t + 1 < K) store_split16(hi_new, lo_new, (long)b * 2 * H + s * H + j, h);
    }
  }
}

__global__ void __launch_bounds__(PK_THREADS, 1)
encode_grouped_kernel(const GroupParams* __restrict__ Cp, const EncParams* __restrict__ Ep, const __grid_constant__ MapTable2 maps) {
  __shared__ GroupParams Csm;
  __shared__ EncParams Esm;
  __shared__ GSched s_sched;
  const int cta = blockIdx.x, G = gridDim.x;
  for (int i = threadIdx.x; i < (int)(sizeof(GroupParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Csm)[i] = reinterpret_cast<const uint32_t*>(Cp)[i];
  for (int i = threadIdx.x; i < (int)(sizeof(EncParams) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&Esm)[i] = reinterpret_cast<const uint32_t*>(Ep)[i];
  __syncthreads();
  const GroupParams& C = Csm;
  const EncParams& E = Esm;
  for (int i = threadIdx.x; i < (int)(sizeof(GSched) / 4); i += PK_THREADS)
    reinterpret_cast<uint32_t*>(&s_sched)[i] = reinterpret_cast<const uint32_t*>(C.gsched + cta)[i];
  extern __shared__ uint8_t smem_raw[];
  const SmemView sv = carve_smem(smem_raw);
  const int H = E.H, R = E.R, B = E.B, K = E.K;
  const uint32_t tmem_base = pipeline_setup(sv);
  if (threadIdx.x < 8) tma_prefetch_desc(&maps.m[threadIdx.x]);
  PipeState ps{0, 0, 0, 0};
  unsigned int sync_target = 0;
  const int m = C.members[0], ns = C.nslots[0];
  const bool member = cta < C.groups * m;
  const int grp = cta / m, mem = cta % m;
  const int tot_kb = s_sched.tot_kb;

  // padding rows of both operand buffers
  for (int e = cta * PK_THREADS + threadIdx.x; e < (R - B) * 2 * H; e += G * PK_THREADS) {
    const __half z = __float2half_rn(0.f);
    for (int q = 0; q < 2; ++q) { C.hh_hi[q][(long)B * 2 * H + e] = z; C.hh_lo[q][(long)B * 2 * H + e] = z; }
  }
  // stationary weights: stage st holds the hi / lo tiles of k-block (st mod tot_kb) of this member's chain for the whole
  // launch (a step walks the stage ring tot_kb stages at a time; tot_kb divides the ring)
  if (member && tot_kb > 0 && threadIdx.x == 0) {
    const uint64_t pol = l2_policy_evict_first();
    pk_expect_tx(sv.bulk_bar, (uint32_t)PK_STAGES * 2 * PK_W_BYTES);
    for (int st = 0; st < PK_STAGES; ++st) {
      int kk = st % tot_kb, ii = 0;
      while (kk >= s_sched.it[ii].nkb) { kk -= s_sched.it[ii].nkb; ++ii; }
      const GItem it = s_sched.it[ii];
      const uint32_t dst = sv.stages_u32 + st * PK_STAGE_BYTES;
      const int wk = (it.wk0 + kk) * GK_KB;
      for (int g = 0; g < 4; ++g) {
        pk_tma_2d_hint(dst + g * (PK_W_BYTES / 4), &maps.m[it.w_map], sv.bulk_bar, wk, g * H + it.wrow, pol);
        pk_tma_2d_hint(dst + PK_W_BYTES + g * (PK_W_BYTES / 4), &maps.m[it.w_map + 1], sv.bulk_bar, wk, g * H + it.wrow, pol);
      }
    }
    pk_wait(sv.bulk_bar, 0);
  }
  __syncthreads();
#pragma unroll 1
  for (int t = 0; t < K; ++t) {
    if (t > 0) {
      ps.npre = (uint32_t)tot_kb;        // "weights already in the stage": the producer only streams the operand tiles
      gphase<false>(C, &s_sched, nullptr, maps.m, t & 1, sv, tmem_base, ps);      // h of frame t-1 sits in buffer t & 1
    }
    if (member) {
      if (t > 0) {
        __syncthreads();
        if (threadIdx.x == 0) {
          unsigned int* ctr = C.group_ctr + grp;
          const unsigned int target = (unsigned int)m * (unsigned int)t;
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
          const long long t0 = clock64();
          while (true) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if ((int)(v - target) >= 0) break;
            if (clock64() - t0 > 8000000000LL) __trap();
          }
        }
        __syncthreads();
      }
      const int c0 = mem * PK_BN / m, c1 = (mem + 1) * PK_BN / m;
      if (ns <= 4) enc_group_cell<4>(C, E, t, grp, ns, c0, c1);
      else enc_group_cell<GK_MAX_MEMBERS>(C, E, t, grp, ns, c0, c1);
    }
    if (t + 1 < K) grid_barrier(E.sync_counter, sync_target, G);
  }
  pipeline_teardown(tmem_base);
}

}  // namespace xg

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
namespace xg {

// derived weight tables: W (rows x K fp32, row-major) -> fp16 hi / lo pairs (rows x Kp, Kp = K rounded up to 64, zero padded)
__global__ void split_weights_f16_kernel(const float* __restrict__ W, int rows, int K, int Kp, __half* __restrict__ hi, __half* __restrict__ lo) {
  const long n = (long)rows * Kp;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const int r = (int)(e / Kp), k = (int)(e % Kp);
    const float w = k < K ? __ldg(W + (long)r * K + k) : 0.f;
    const __half h = __float2half_rn(w);
    hi[e] = h;
    lo[e] = __float2half_rn((w - __half2float(h)) * X16_SCALE);
  }
}

// K-major fp16 operand map: boxes of 64 halves (128 bytes, SWIZZLE_128B) x box_rows rows
static int make_map16(xg_context* ctx, TcState* ts, const __half* base, int rows, int Kp, int box_rows, CUtensorMap* out) {
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)GK_KB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = ts->encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ctx->es.set(__FILE__, __LINE__, "cuTensorMapEncodeTiled (fp16 operand) failed", nullptr);
    return XG_ERR_CUDA;
  }
  return XG_OK;
}

// tables derived from the bound parameters, shared by the grouped word-loop kernels of a handle: the POS-gate factor of
// every token (tgate = relu(embed . W_gate^T + b), V x H) and the fp16 hi / lo pairs of the eight weight matrices of the
// word step.  Rebuilt (on the caller's stream) whenever the parameter epoch moved.
struct WordTables {
  float* tgate = nullptr;
  __half* w16[8][2] = {};        // h2a, logit, l1_i2h, l1_a2h, l1_h2h, l2_i2h, l2_a2h, l2_h2h
  unsigned long long tgate_epoch = ~0ull, w_epoch[8] = {~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull};
};
inline WordTables*& word_tables_slot(xg_context* ctx) {
  static std::unordered_map<xg_context*, WordTables*> m;
  return m[ctx];
}
static void word_tables_release(xg_context* ctx) {
  WordTables* t = word_tables_slot(ctx);
  if (!t) return;
  if (t->tgate) cudaFree(t->tgate);
  for (int i = 0; i < 8; ++i) for (int q = 0; q < 2; ++q) if (t->w16[i][q]) cudaFree(t->w16[i][q]);
  delete t;
  word_tables_slot(ctx) = nullptr;
}
static const int kWordParams[8] = {XG_P_H2A_W, XG_P_LOGIT_W, XG_P_L1_I2H_W, XG_P_L1_A2H_W, XG_P_L1_H2H_W, XG_P_L2_I2H_W, XG_P_L2_A2H_W, XG_P_L2_H2H_W};
constexpr unsigned WT_ALL = 0xffu, WT_TRAIN = 0x01u | 0x10u | 0x20u | 0x40u | 0x80u;    // the training loop streams h2a, l1_h2h and lstm_2
struct SplitJobs { const float* W[8]; __half* hi[8]; __half* lo[8]; int rows[8], K[8], Kp[8]; int n; };
// every stale table of a call in ONE launch: blockIdx.y = job
__global__ void split_weights_f16_multi_kernel(const SplitJobs J) {
  const int q = blockIdx.y;
  const float* W = J.W[q];
  __half* hi = J.hi[q]; __half* lo = J.lo[q];
  const int K = J.K[q], Kp = J.Kp[q];
  const long n = (long)J.rows[q] * Kp;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
    const int r = (int)(e / Kp), k = (int)(e % Kp);
    const float w = k < K ? __ldg(W + (long)r * K + k) : 0.f;
    const __half h = __float2half_rn(w);
    hi[e] = h;
    lo[e] = __float2half_rn((w - __half2float(h)) * X16_SCALE);
  }
}
static int word_tables(xg_context* ctx, cudaStream_t st, WordTables** out, bool need_tgate = true, unsigned need = WT_ALL) {
  const xg_dims& d = ctx->d;
  WordTables*& T = word_tables_slot(ctx);
  if (!T) {
    T = new WordTables();
    XG_CUDA_TRY(ctx->es, cudaMalloc(&T->tgate, sizeof(float) * (size_t)d.vocab * d.rnn));
    for (int i = 0; i < 8; ++i) {
      int rows, cols;
      param_shape(d, kWordParams[i], &rows, &cols);
      const size_t n = (size_t)rows * ((cols + GK_KB - 1) / GK_KB * GK_KB);
      for (int q = 0; q < 2; ++q) XG_CUDA_TRY(ctx->es, cudaMalloc(&T->w16[i][q], sizeof(__half) * n));
    }
  }
  if (need_tgate && T->tgate_epoch != ctx->param_epoch) {
    GemmP g = gemm_nt(ctx->P[XG_P_EMBED_W], d.embed, ctx->P[XG_P_DGATE_W], d.embed, T->tgate, d.rnn, d.vocab, d.rnn, d.embed);
    g.ep.bias0 = ctx->P[XG_P_DGATE_B];
    g.ep.act = XG_ACT_RELU;
    XG_TRY(gemm_run(ctx, g, st));
    T->tgate_epoch = ctx->param_epoch;
  }
  SplitJobs J;
  J.n = 0;
  for (int i = 0; i < 8; ++i) {
    if (!((need >> i) & 1u) || T->w_epoch[i] == ctx->param_epoch) continue;
    int rows, cols;
    param_shape(d, kWordParams[i], &rows, &cols);
    J.W[J.n] = ctx->P[kWordParams[i]]; J.hi[J.n] = T->w16[i][0]; J.lo[J.n] = T->w16[i][1];
    J.rows[J.n] = rows; J.K[J.n] = cols; J.Kp[J.n] = (cols + GK_KB - 1) / GK_KB * GK_KB;
    ++J.n;
    T->w_epoch[i] = ctx->param_epoch;
  }
  if (J.n > 0) {
    ProfScope ps(ctx, "split_weights_f16", st);
    split_weights_f16_multi_kernel<<<dim3(ctx->sm_count, J.n), 256, 0, st>>>(J);
    XG_LAUNCH_CHECK(ctx->es);
  }
  *out = T;
  return XG_OK;
}
// weight maps of the word step: 0 h2a, 2 logit (128-row boxes), GM_W32.. the six LSTM matrices (32-row boxes)
static int word_weight_maps(xg_context* ctx, TcState* ts, const WordTables* T, CUtensorMap* maps) {
  for (int i = 0; i < 8; ++i) {
    int rows, cols;
    param_shape(ctx->d, kWordParams[i], &rows, &cols);
    const int Kp = (cols + GK_KB - 1) / GK_KB * GK_KB;
    const int base = i < 2 ? 2 * i : GM_W32 + 2 * (i - 2);
    XG_TRY(make_map16(ctx, ts, T->w16[i][0], rows, Kp, i < 2 ? 128 : 32, &maps[base]));
    XG_TRY(make_map16(ctx, ts, T->w16[i][1], rows, Kp, i < 2 ? 128 : 32, &maps[base + 1]));
  }
  return XG_OK;
}

// dec_attention_rows stages the queries of NR rows behind exp(2Uv) (and V, when it is fetched behind it)
static bool att_rows_fit(const xg_dims& d, int K, int nr) {
  const long euv = (long)K * d.att * 4, vb = (long)K * (d.att + d.rnn) * 4 <= (long)PK_STAGES * PK_STAGE_BYTES ? (long)K * d.rnn * 4 : 0;
  return euv + vb + (long)nr * d.att * 4 <= (long)PK_STAGES * PK_STAGE_BYTES;
}

struct GroupedState {
  int R = 0, K = 0;
  float* lraw = nullptr;
  char* pool = nullptr;
  size_t pool_bytes = 0;
  GroupParams hp;
  GroupParams* d_params = nullptr;
  unsigned int* d_counter = nullptr;
  int* d_flags = nullptr;
  long long* d_dbg = nullptr;
  bool attr_set = false;
};
inline GroupedState*& grouped_state(xg_context* ctx) {
  static std::unordered_map<xg_context*, GroupedState*> m;
  return m[ctx];
}
static void grouped_release(xg_context* ctx) {
  GroupedState* s = grouped_state(ctx);
  if (!s) return;
  if (s->pool) cudaFree(s->pool);
  delete s;
  grouped_state(ctx) = nullptr;
}

// Greedy decoding on decode_grouped_kernel.  PK_FALLBACK: the shape is outside this kernel (the caller runs
// decode_persistent_kernel<0>, which covers every shape persist_eligible() accepts).
struct GroupedSampling {      // sampling form of the loop: multinomial draw (sample_max = 0) and / or the training dropout of the step
  int sample_max = 1; float temperature = 1.f; uint64_t seed = 0;
  bool step_drop = false; uint64_t drop_seed = 0;
  // scheduled-sampling token pass: T = Lp steps, tokens into ss_used (pre-filled with seq), nothing else recorded; asynchronous
  bool ss_mode = false; float ss_prob = 0.f; uint64_t ss_seed = 0;
  const int64_t* ss_seq = nullptr; const float* ss_mask = nullptr; int64_t* ss_used = nullptr; int ss_L = 0;
};
static int grouped_decode(xg_context* ctx, const float* Vf, const float* Uv, const float* pos, const float* const* state0,
                          int B, int K, int T, int64_t* seq_out, float* logp_out, int* steps_out, cudaStream_t st,
                          const GroupedSampling* smp = nullptr) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, A = d.att, V = d.vocab;
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN, Ep = (E + GK_KB - 1) / GK_KB * GK_KB, G = ctx->sm_count;
  if (env_flag("XG_NO_GROUPED") || !persist_eligible(ctx, B, K) || T > 2048 || H % GK_KB != 0 || Ep > DEC_TI * PK_THREADS) return PK_FALLBACK;
  const int kbH = H / GK_KB, kbE = Ep / GK_KB;
  const int ntiles = H / 32, ncb = R / PK_BN, groups = ntiles * ncb;
  if (groups > G || 4 * H > 32000 || !att_rows_fit(d, K, 1)) return PK_FALLBACK;
  // the attention of step t+1 keeps the last n_att CTAs through the pick phase and F1: F1's groups live on the others
  const int n_att = B;
  if (ncb != 1 || G - R < groups || G - R < B) return PK_FALLBACK;
  // (the schedule depends on the padded row count only, never on B: a caption's arithmetic - K split, summation order -
  //  is the same whatever batch it is decoded in)
  int members_l[2] = {std::max(1, std::min(std::min(GK_MAX_MEMBERS, (G - R) / groups), kbE + kbH)),
                      std::max(1, std::min(std::min(GK_MAX_MEMBERS, G / groups), 2 * kbH))};
  if (getenv("XG_M1")) members_l[0] = std::max(1, std::min(members_l[0], atoi(getenv("XG_M1"))));      // (experiments)
  if (getenv("XG_M2")) members_l[1] = std::max(1, std::min(members_l[1], atoi(getenv("XG_M2"))));
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  GroupedState*& S = grouped_state(ctx);
  if (!S) S = new GroupedState();
  GroupParams& hp = S->hp;
  DecParams& dp = hp.dp;

  // ---- products ----
  auto mk = [&](int id, int wmap, int xkb0, int n_rows, int nkb) {
    GDesc& g = dp.d[id];
    g.w_map = wmap; g.x_hi = GM_HH; g.x_lo = GM_HH + 1; g.xkb0 = xkb0; g.n_rows = n_rows; g.nkb = nkb; g.ns = 0;
  };
  for (int i = 0; i < PK_MAX_DESCS; ++i) { dp.d[i].ns = 0; dp.d[i].n_rows = 0; dp.d[i].nkb = 0; }
  mk(DD_AH, GM_H2A, 0, A, 2 * kbH);
  mk(DD_LOGIT, GM_LOGIT, kbH, V, kbH);      // (full-K tiles reduced in the epilogue: no slots)
  // logits: one CTA per (128-row vocabulary tile, column block), full K, reduced in the epilogue
  const int ntv = (V + 127) / 128, nlog = ntv * ncb;
  if (nlog > G - 8 || ntv > 256) return PK_FALLBACK;
  const int nside = G - nlog;

  std::vector<GSched> sched((size_t)3 * G);
  memset(sched.data(), 0, sizeof(GSched) * sched.size());
  // ---- fused cell phases: one chain per member over the layer's token / attention dependent K extent; the recurrent
  //      product of the layer (it needs only the state) is computed one phase-set earlier, next to the logits, by the
  //      CTAs that have no logits tile, into nv "early" slots of the same group ----
  struct Prod { int w_map, x_map, xsel, xkb0, nkb; };
  // in-phase products of a layer; the third one (recurrent: reads the [h1|h2] buffer entering the step) only when the layer
  // has no early slots
  const Prod layers[2][3] = {{{GM_W32 + 0, GM_XT, 0, 0, kbE}, {GM_W32 + 2, GM_GP, 0, 0, kbH}, {GM_W32 + 4, GM_HH, 1, 0, kbH}},
                             {{GM_W32 + 6, GM_HH, 2, 0, kbH}, {GM_W32 + 8, GM_AF, 0, 0, kbH}, {GM_W32 + 10, GM_HH, 1, kbH, kbH}}};
  const Prod early[2] = {{GM_W32 + 4, GM_HH, 2, 0, kbH}, {GM_W32 + 10, GM_HH, 2, kbH, kbH}};   // W_h2h1.h1', W_h2h2.h2' (the buffer just written)
  const int early_mask = getenv("XG_EARLY") ? atoi(getenv("XG_EARLY")) : 3;     // bit l: layer l's recurrent product runs early
  const int nv_l[2] = {(early_mask & 1) ? (kbH >= 2 ? 2 : 1) : 0, (early_mask & 2) ? (kbH >= 2 ? 2 : 1) : 0};
  const int nslots_l[2] = {members_l[0] + nv_l[0], members_l[1] + nv_l[1]};
  if (nslots_l[0] > GK_MAX_SLOTS || nslots_l[1] > GK_MAX_SLOTS) return PK_FALLBACK;
  for (int layer = 0; layer < 2; ++layer) {
    const int members = members_l[layer], ns = nslots_l[layer];
    const int nprod = nv_l[layer] ? 2 : 3;
    int Ktot = 0;
    for (int p = 0; p < nprod; ++p) Ktot += layers[layer][p].nkb;
    for (int grp = 0; grp < groups; ++grp) {
      const int tile = grp / ncb, cb = grp % ncb;
      for (int mem = 0; mem < members; ++mem) {
        GSched& sc = sched[(size_t)layer * G + grp * members + mem];
        int k0 = (int)((long)mem * Ktot / members), k1 = (int)((long)(mem + 1) * Ktot / members);
        int base = 0;
        for (int p = 0; p < nprod && k0 < k1; ++p) {
          const Prod& pr = layers[layer][p];
          const int lo = std::max(k0, base), hi = std::min(k1, base + pr.nkb);
          if (lo < hi) {
            GItem it{};
            it.w_map = (short)pr.w_map; it.x_map = (short)pr.x_map; it.xsel = (short)pr.xsel;
            it.flags = (short)(GI_FUSED | (layer ? GI_LAYER1 : 0) | (sc.n > 0 ? GI_CONT_PREV : 0));
            it.wrow = (short)(tile * 32); it.wk0 = (short)(lo - base); it.xk0 = (short)(pr.xkb0 + lo - base); it.nkb = (short)(hi - lo);
            it.desc = (short)grp; it.slot = (short)mem; it.cb = (short)cb; it.pad = (short)ns;
            if (sc.n > 0) sc.it[sc.n - 1].flags |= GI_CONT_NEXT;
            sc.it[sc.n++] = it;
            sc.tot_kb += (short)(hi - lo);
            sc.tot_chunks += (short)((hi - lo + PK_CHUNK - 1) / PK_CHUNK);
          }
          base += pr.nkb;
        }
        sc.n_chains = sc.n > 0 ? 1 : 0;
      }
    }
  }
  // ---- G4: logits tiles on the first nlog CTAs; the others share runs of: the attention query (split-K slots, consumed by
  //      the attention phase) and the early slots of both layers ----
  for (int c = 0; c < nlog; ++c) {
    GSched& sc = sched[(size_t)2 * G + c];
    GItem it{};
    it.w_map = (short)GM_LOGIT; it.x_map = (short)GM_HH; it.xsel = 2; it.flags = GI_LOGITS;
    it.wrow = (short)((c / ncb) * 128); it.wk0 = 0; it.xk0 = (short)kbH; it.nkb = (short)kbH;
    it.desc = DD_LOGIT; it.slot = 0; it.cb = (short)(c % ncb);
    sc.it[sc.n++] = it;
    sc.tot_kb = (short)kbH; sc.tot_chunks = (short)((kbH + PK_CHUNK - 1) / PK_CHUNK); sc.n_chains = 1;
  }
  {
    std::vector<GItem> runs;
    const int ah_slots = std::min(4, 2 * kbH), ah_run = (2 * kbH + ah_slots - 1) / ah_slots;
    dp.d[DD_AH].ns = (2 * kbH + ah_run - 1) / ah_run;
    for (int rt = 0; rt < (A + 127) / 128; ++rt)
      for (int cb = 0; cb < ncb; ++cb)
        for (int k0 = 0, sl = 0; k0 < 2 * kbH; k0 += ah_run, ++sl) {
          GItem it{};
          it.w_map = (short)GM_H2A; it.x_map = (short)GM_HH; it.xsel = 2; it.flags = 0;
          it.wrow = (short)(rt * 128); it.wk0 = (short)k0; it.xk0 = (short)k0; it.nkb = (short)std::min(ah_run, 2 * kbH - k0);
          it.desc = DD_AH; it.slot = (short)sl; it.cb = (short)cb;
          runs.push_back(it);
        }
    for (int layer = 0; layer < 2; ++layer)
      for (int grp = 0; grp < groups; ++grp)
        for (int v = 0; v < nv_l[layer]; ++v) {
          const Prod& pr = early[layer];
          const int nv = nv_l[layer];
          const int k0 = v * pr.nkb / nv, k1 = (v + 1) * pr.nkb / nv;
          GItem it{};
          it.w_map = (short)pr.w_map; it.x_map = (short)pr.x_map; it.xsel = (short)pr.xsel;
          it.flags = (short)(GI_FUSED | (layer ? GI_LAYER1 : 0));
          it.wrow = (short)((grp / ncb) * 32); it.wk0 = (short)k0; it.xk0 = (short)(pr.xkb0 + k0); it.nkb = (short)(k1 - k0);
          it.desc = (short)grp; it.slot = (short)(members_l[layer] + v); it.cb = (short)(grp % ncb); it.pad = (short)nslots_l[layer];
          runs.push_back(it);
        }
    std::stable_sort(runs.begin(), runs.end(), [](const GItem& a, const GItem& b) { return a.nkb > b.nkb; });
    std::vector<int> load(nside, 0);
    for (const GItem& it : runs) {
      const int c = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      GSched& sc = sched[(size_t)2 * G + nlog + c];
      if (sc.n >= GK_MAX_ITEMS) return PK_FALLBACK;
      sc.it[sc.n++] = it;
      sc.tot_kb += it.nkb; sc.tot_chunks += (short)((it.nkb + PK_CHUNK - 1) / PK_CHUNK); sc.n_chains++;
      load[c] += it.nkb;
    }
  }

  // ---- device pool ----
  if (S->R != R || S->K != K) {
    if (S->pool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->pool); S->pool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->pool, pass == 0 ? 0 : S->pool_bytes);
      S->d_params = a.take<GroupParams>(1);
      S->d_counter = a.take<unsigned int>(128 + 2 * groups);
      S->d_flags = a.take<int>(2048);
      S->d_dbg = a.take<long long>((2048 + 256) * PK_STAMPS + 64);
      hp.gsched = a.take<GSched>(sched.size());
      hp.lpart = a.take<float4>((size_t)R * ntv);
      S->lraw = a.take<float>((size_t)R * ntv * 128);
      dp.d[DD_AH].out = a.take<float>((size_t)PK_MAX_SLOTS * R * A);   // [slot][caption][row]
      for (int q = 0; q < 2; ++q) hp.fslots[q] = a.take<float>((size_t)groups * nslots_l[q] * PK_BN * 128);
      // fp16 hi / lo pairs (the DecParams fields are float*: the pointwise phases cast them back when x16 is set)
      dp.xt_hi = reinterpret_cast<float*>(a.take<__half>((long)R * Ep)); dp.xt_lo = reinterpret_cast<float*>(a.take<__half>((long)R * Ep));
      for (int q = 0; q < 2; ++q) { hp.hh_hi[q] = a.take<__half>((long)R * 2 * H); hp.hh_lo[q] = a.take<__half>((long)R * 2 * H); }
      dp.gp_hi = reinterpret_cast<float*>(a.take<__half>((long)R * H)); dp.gp_lo = reinterpret_cast<float*>(a.take<__half>((long)R * H));
      dp.af_hi = reinterpret_cast<float*>(a.take<__half>((long)R * H)); dp.af_lo = reinterpret_cast<float*>(a.take<__half>((long)R * H));

      dp.hx = a.take<float>((long)R * 2 * H);
      dp.cx = a.take<float>((long)2 * R * H);
      dp.unfinished = a.take<float>(R);
      dp.tok = a.take<int64_t>(R);
      dp.EUv = a.take<float>((long)R * K * A);
      if (pass == 0) {
        S->pool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->pool, S->pool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->pool, 0, S->pool_bytes, st));
      }
    }
    S->R = R; S->K = K;
  }
  dp.hh_hi = nullptr; dp.hh_lo = nullptr;
  hp.pick_ctr = S->d_counter + 64;
  hp.group_ctr = S->d_counter + 128;
  hp.members[0] = members_l[0]; hp.members[1] = members_l[1]; hp.groups = groups; hp.ncb = ncb; hp.ntv = ntv; hp.n_att = n_att;
  hp.nslots[0] = nslots_l[0]; hp.nslots[1] = nslots_l[1];
  hp.l2_hints = getenv("XG_L2_HINT") ? atoi(getenv("XG_L2_HINT")) : 1;
  hp.topk = 0; hp.lraw = smp ? S->lraw : nullptr;
  hp.sample_max = smp ? smp->sample_max : 1; hp.step_drop = smp && smp->step_drop ? 1 : 0;
  hp.inv_temp = smp && !smp->sample_max ? 1.f / smp->temperature : 1.f;
  hp.sample_seed = smp ? smp->seed : 0;
  hp.ss_mode = smp && smp->ss_mode ? 1 : 0; hp.ss_L = smp ? smp->ss_L : 0; hp.ss_Lp = T;
  hp.ss_prob = smp ? smp->ss_prob : 0.f; hp.ss_seed = smp ? smp->ss_seed : 0;
  hp.ss_seq = smp ? smp->ss_seq : nullptr; hp.ss_mask = smp ? smp->ss_mask : nullptr; hp.ss_used = smp ? smp->ss_used : nullptr;
  {
    const bool dr = smp && smp->step_drop;
    hp.drop_gate = make_drop(dr, d.drop_prob, smp ? smp->drop_seed : 0, XG_DROP_DEC_GATE);
    hp.drop_h1 = make_drop(dr, d.drop_prob, smp ? smp->drop_seed : 0, XG_DROP_DEC_H1);
    hp.drop_h2 = make_drop(dr, d.drop_prob, smp ? smp->drop_seed : 0, XG_DROP_DEC_H2);
  }
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<GSched*>(hp.gsched), sched.data(), sizeof(GSched) * sched.size(),
                                       cudaMemcpyHostToDevice, st));

  // ---- tables derived from the bound parameters + tensor maps ----
  WordTables* WT = nullptr;
  XG_TRY(word_tables(ctx, st, &WT));
  MapTable2 mt;
  CUtensorMap* maps = mt.m;
  XG_TRY(word_weight_maps(ctx, ts, WT, maps));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.xt_hi), R, Ep, PK_BN, &maps[GM_XT]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.xt_lo), R, Ep, PK_BN, &maps[GM_XT + 1]));
  for (int q = 0; q < 2; ++q) {
    XG_TRY(make_map16(ctx, ts, hp.hh_hi[q], R, 2 * H, PK_BN, &maps[GM_HH + 2 * q]));
    XG_TRY(make_map16(ctx, ts, hp.hh_lo[q], R, 2 * H, PK_BN, &maps[GM_HH + 2 * q + 1]));
  }
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.gp_hi), R, H, PK_BN, &maps[GM_GP]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.gp_lo), R, H, PK_BN, &maps[GM_GP + 1]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.af_hi), R, H, PK_BN, &maps[GM_AF]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.af_lo), R, H, PK_BN, &maps[GM_AF + 1]));
  {   // V as [B*K][H]: one box = (H/2 columns) x (K frames) of a caption, dense in shared memory
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)B * K};
    cuuint64_t strides[1] = {(cuuint64_t)H * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)(H / 2), (cuuint32_t)K};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult cr = ts->encode(&maps[GM_V], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Vf), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { ctx->es.set(__FILE__, __LINE__, "cuTensorMapEncodeTiled (V) failed", nullptr); return XG_ERR_CUDA; }
  }
  maps[27] = maps[0];

  dp.sched = nullptr;
  dp.B = B; dp.R = R; dp.K = K; dp.H = H; dp.E = E; dp.Ep = Ep; dp.A = A; dp.V = V; dp.T = T;
  dp.b_h2a = ctx->P[XG_P_H2A_B]; dp.w_a2w = ctx->P[XG_P_A2W_W]; dp.b_a2w = ctx->P[XG_P_A2W_B];
  dp.bias[0][0] = ctx->P[XG_P_L1_I2H_B]; dp.bias[0][1] = ctx->P[XG_P_L1_A2H_B]; dp.bias[0][2] = ctx->P[XG_P_L1_H2H_B];
  dp.bias[1][0] = ctx->P[XG_P_L2_I2H_B]; dp.bias[1][1] = ctx->P[XG_P_L2_A2H_B]; dp.bias[1][2] = ctx->P[XG_P_L2_H2H_B];
  dp.b_logit = ctx->P[XG_P_LOGIT_B]; dp.embed = ctx->P[XG_P_EMBED_W];
  dp.tgate = WT->tgate;
  dp.Vf = Vf; dp.Uv = Uv; dp.pos = pos;
  for (int q = 0; q < 4; ++q) dp.state0[q] = state0[q];
  dp.mode = 0; dp.feat_div = 1; dp.build_euv = 1; dp.x16 = 1;
  dp.seq = seq_out; dp.seqlogp = logp_out; dp.flags = S->d_flags;
  dp.sync_counter = S->d_counter;
  dp.dbg_clock = env_flag("XG_PERSIST_TRACE") ? S->d_dbg : nullptr;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_params, &hp, sizeof(GroupParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_counter, 0, sizeof(unsigned int) * (128 + 2 * groups), st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_flags, 0, sizeof(int) * (size_t)T, st));
  if (seq_out) XG_CUDA_TRY(ctx->es, cudaMemsetAsync(seq_out, 0, sizeof(int64_t) * (size_t)B * T, st));
  if (logp_out) XG_CUDA_TRY(ctx->es, cudaMemsetAsync(logp_out, 0, sizeof(float) * (size_t)B * T, st));
  if (dp.dbg_clock) XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_dbg, 0, sizeof(long long) * ((2048 + 256) * PK_STAMPS + 64), st));
  if (!S->attr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_grouped_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, GK_SMEM_BYTES));
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_grouped_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, GK_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_grouped_kernel<1>, PK_THREADS, GK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "grouped decoder does not fit on an SM");
    S->attr_set = true;
  }
  {
    ProfScope ps(ctx, "decode_persistent", st);
    const GroupParams* gp = S->d_params;
    void* args[2] = {(void*)&gp, (void*)&mt};
    void* fn = smp ? (void*)decode_grouped_kernel<1> : (void*)decode_grouped_kernel<0>;
    XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(PK_THREADS), args, GK_SMEM_BYTES, st));
    ctx->n_fused++;
  }
  if (hp.ss_mode || !steps_out) return XG_OK;      // the token pass has nothing to read back; steps_out == NULL: asynchronous call
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(ctx->h_pinned, S->d_flags, sizeof(int) * (size_t)T, cudaMemcpyDeviceToHost, st));
  XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
  int steps = 0;
  while (steps < T && ctx->h_pinned[steps] != 0) ++steps;
  *steps_out = steps;
  if (dp.dbg_clock) {   // XG_PERSIST_TRACE=1: average SM cycles per phase (CTA 0) + step 3 across all CTAs
    std::vector<long long> h((size_t)T * PK_STAMPS);
    cudaMemcpy(h.data(), S->d_dbg, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
    const char* names[4] = {"F1 (lstm_1 fused)", "F3 (lstm_2 fused)", "G4 (logits, ah)", "P4 (pick || attention t+1)"};
    double tot = 0;
    const int n = steps > 1 ? steps - 1 : 1;
    for (int i = 0; i < 4; ++i) {
      double w = 0, b = 0;
      for (int t = 1; t < std::max(steps, 2); ++t) {
        w += (double)(h[t * PK_STAMPS + 2 * i + 1] - h[t * PK_STAMPS + 2 * i]);
        b += (double)(h[t * PK_STAMPS + 2 * i + 2] - h[t * PK_STAMPS + 2 * i + 1]);
      }
      w /= n; b /= n;
      tot += w + b;
      fprintf(stderr, "[xg grouped trace] %-28s own work %7.0f cycles   barrier wait %7.0f cycles\n", names[i], w, b);
    }
    fprintf(stderr, "[xg grouped trace] step %.0f cycles\n", tot);
#ifdef GK_FINE
    {
      long long f[32];
      cudaMemcpy(f, S->d_dbg + (2048 + 256) * PK_STAMPS, sizeof(f), cudaMemcpyDeviceToHost);
      for (int l = 0; l < 2; ++l)
        fprintf(stderr, "[xg grouped trace] step 3 cta 0 fused layer %d (cycles after phase entry): producer done %lld, epilogue warp done %lld, "
                "cta synced %lld, group counter seen %lld, cta synced %lld, cell done %lld\n", l, f[l * 16 + 1] - f[l * 16], f[l * 16 + 2] - f[l * 16],
                f[l * 16 + 3] - f[l * 16], f[l * 16 + 4] - f[l * 16], f[l * 16 + 5] - f[l * 16], f[l * 16 + 6] - f[l * 16]);
    }
#endif
    if (steps > 3 && G <= 256) {
      std::vector<long long> ga((size_t)G * PK_STAMPS);
      cudaMemcpy(ga.data(), S->d_dbg + 2048 * PK_STAMPS, sizeof(long long) * ga.size(), cudaMemcpyDeviceToHost);
      for (int i = 0; i < 4; ++i) {
        long long open = 0, close = 0;
        for (int c = 0; c < G; ++c) open = std::max(open, ga[(size_t)c * PK_STAMPS + 2 * i]);
        std::vector<long long> fin(G);
        for (int c = 0; c < G; ++c) fin[c] = ga[(size_t)c * PK_STAMPS + 2 * i + 1] - open;
        std::vector<long long> srt = fin;
        std::sort(srt.begin(), srt.end());
        for (int c = 0; c < G; ++c) close = std::max(close, ga[(size_t)c * PK_STAMPS + 2 * i + 2]);
        const int worst = (int)(std::max_element(fin.begin(), fin.end()) - fin.begin());
        fprintf(stderr, "[xg grouped trace] step 3 %-28s work done after: min %6lld  median %6lld  p90 %6lld  max %6lld ns (cta %d)   barrier exit %6lld ns\n",
                names[i], srt[0], srt[G / 2], srt[G * 9 / 10], srt[G - 1], worst, close - open);
      }
    }
  }
  return XG_OK;
}

// ---- teacher-forced word loop on train_grouped_kernel; PK_FALLBACK: shape outside it (caller: decode_persistent_kernel<1>) ----
struct GroupedTrainState {
  int R = 0, K = 0;
  char* pool = nullptr;
  size_t pool_bytes = 0;
  GroupParams hp;
  GroupParams* d_params = nullptr;
  unsigned int* d_counter = nullptr;
  long long* d_dbg = nullptr;
  bool attr_set = false;
};
inline GroupedTrainState*& grouped_train_state(xg_context* ctx) {
  static std::unordered_map<xg_context*, GroupedTrainState*> m;
  return m[ctx];
}
static void grouped_train_release(xg_context* ctx) {
  GroupedTrainState* s = grouped_train_state(ctx);
  if (!s) return;
  if (s->pool) cudaFree(s->pool);
  delete s;
  grouped_train_state(ctx) = nullptr;
}

static int grouped_train(xg_context* ctx, const float* Vf, const float* Uv, int B, int K, int T, const PersistTrainIO& tr, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, A = d.att;
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN, G = ctx->sm_count;
  if (env_flag("XG_NO_GROUPED") || env_flag("XG_NO_GROUPED_TRAIN") || !persist_eligible(ctx, B, K) || T > 2048 || H % GK_KB != 0 || 4 * H > 32000)
    return PK_FALLBACK;
  const int kbH = H / GK_KB;
  const int groups = H / 32, ncb = R / PK_BN, nat = (A + 127) / 128;
  if (ncb != 1 || G - R < groups + 1 || !att_rows_fit(d, K, 1)) return PK_FALLBACK;
  // (the schedule depends on the padded row count only, never on B)
  const int avail = G - R;                       // CTAs of phase B that do not run the attention
  const int nv = kbH >= 2 ? 2 : 1;               // early slots of lstm_2: W_h2h2.h2 cut in nv runs
  // lstm_1's groups take every CTA phase B has outside the attention; the recurrent product of lstm_2 (it needs only
  // the state entering the step) rides in phase A next to the attention query
  int m1 = std::max(1, std::min(std::min(GK_MAX_MEMBERS, kbH), avail / groups));
  if (getenv("XG_TM1")) m1 = std::max(1, std::min(m1, atoi(getenv("XG_TM1"))));      // (experiments)
  const int m2 = std::max(1, std::min(std::min(GK_MAX_MEMBERS, G / groups), 2 * kbH));
  const int members_l[2] = {m1, m2}, nslots_l[2] = {m1, m2 + nv};
  if (nslots_l[1] > GK_MAX_SLOTS) return PK_FALLBACK;
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  GroupedTrainState*& S = grouped_train_state(ctx);
  if (!S) S = new GroupedTrainState();
  GroupParams& hp = S->hp;
  DecParams& dp = hp.dp;
  for (int i = 0; i < PK_MAX_DESCS; ++i) { dp.d[i].ns = 0; dp.d[i].n_rows = 0; dp.d[i].nkb = 0; }
  {
    GDesc& g = dp.d[DD_AH];
    g.w_map = GM_H2A; g.x_hi = GM_HH; g.x_lo = GM_HH + 1; g.xkb0 = 0; g.n_rows = A; g.nkb = 2 * kbH;
  }
  std::vector<GSched> sched((size_t)3 * G);
  memset(sched.data(), 0, sizeof(GSched) * sched.size());
  auto add = [&](GSched& sc, const GItem& it, bool new_chain) -> bool {
    if (sc.n >= GK_MAX_ITEMS) return false;
    sc.it[sc.n++] = it;
    sc.tot_kb += it.nkb; sc.tot_chunks += (short)((it.nkb + PK_CHUNK - 1) / PK_CHUNK);
    if (new_chain) sc.n_chains++;
    return true;
  };
  {   // phase A: attention query, (row tile) x (K run) items over all CTAs
    int ah_slots = std::min(PK_MAX_SLOTS, 2 * kbH);
    while (ah_slots > 1 && nat * ah_slots > G) --ah_slots;
    const int ah_run = (2 * kbH + ah_slots - 1) / ah_slots;
    dp.d[DD_AH].ns = (2 * kbH + ah_run - 1) / ah_run;
    int c = 0;
    for (int k0 = 0, sl = 0; k0 < 2 * kbH; k0 += ah_run, ++sl)
      for (int rt = 0; rt < nat; ++rt) {
        GItem it{};
        it.w_map = (short)GM_H2A; it.x_map = (short)GM_HH; it.xsel = 1; it.flags = 0;
        it.wrow = (short)(rt * 128); it.wk0 = (short)k0; it.xk0 = (short)k0; it.nkb = (short)std::min(ah_run, 2 * kbH - k0);
        it.desc = DD_AH; it.slot = (short)sl; it.cb = 0;
        if (!add(sched[(size_t)(c++ % G)], it, true)) return PK_FALLBACK;
      }
    for (int grp = 0; grp < groups; ++grp)      // W_h2h2.h2 into the early slots of lstm_2's groups
      for (int v = 0; v < nv; ++v) {
        const int k0 = v * kbH / nv, k1 = (v + 1) * kbH / nv;
        GItem it{};
        it.w_map = (short)(GM_W32 + 10); it.x_map = (short)GM_HH; it.xsel = 1; it.flags = (short)(GI_FUSED | GI_LAYER1);
        it.wrow = (short)(grp * 32); it.wk0 = (short)k0; it.xk0 = (short)(kbH + k0); it.nkb = (short)(k1 - k0);
        it.desc = (short)grp; it.slot = (short)(m2 + v); it.cb = 0; it.pad = (short)nslots_l[1];
        if (!add(sched[(size_t)(c++ % G)], it, true)) return PK_FALLBACK;
      }
  }
  // phase B: lstm_1 groups (recurrent product only: the rest sits in G1s) on the first groups x m1 CTAs, the recurrent product
  // of lstm_2 into its early slots on the side CTAs; phase C: lstm_2 groups (h1', af) on all CTAs
  for (int grp = 0; grp < groups; ++grp) {
    for (int mem = 0; mem < m1; ++mem) {
      const int k0 = mem * kbH / m1, k1 = (mem + 1) * kbH / m1;
      if (k0 >= k1) continue;
      GItem it{};
      it.w_map = (short)(GM_W32 + 4); it.x_map = (short)GM_HH; it.xsel = 1; it.flags = (short)GI_FUSED;
      it.wrow = (short)(grp * 32); it.wk0 = (short)k0; it.xk0 = (short)k0; it.nkb = (short)(k1 - k0);
      it.desc = (short)grp; it.slot = (short)mem; it.cb = 0; it.pad = (short)nslots_l[0];
      if (!add(sched[(size_t)G + grp * m1 + mem], it, true)) return PK_FALLBACK;
    }
    const int Ktot = 2 * kbH;
    for (int mem = 0; mem < m2; ++mem) {
      GSched& sc = sched[(size_t)2 * G + grp * m2 + mem];
      const int k0 = (int)((long)mem * Ktot / m2), k1 = (int)((long)(mem + 1) * Ktot / m2);
      for (int p = 0; p < 2; ++p) {
        const int base = p * kbH, lo = std::max(k0, base), hi = std::min(k1, base + kbH);
        if (lo >= hi) continue;
        GItem it{};
        it.w_map = (short)(p == 0 ? GM_W32 + 6 : GM_W32 + 8); it.x_map = (short)(p == 0 ? GM_HH : GM_AF); it.xsel = (short)(p == 0 ? 2 : 0);
        it.flags = (short)(GI_FUSED | GI_LAYER1 | (sc.n > 0 ? GI_CONT_PREV : 0));
        it.wrow = (short)(grp * 32); it.wk0 = (short)(lo - base); it.xk0 = (short)(lo - base); it.nkb = (short)(hi - lo);
        it.desc = (short)grp; it.slot = (short)mem; it.cb = 0; it.pad = (short)nslots_l[1];
        if (sc.n > 0) sc.it[sc.n - 1].flags |= GI_CONT_NEXT;
        if (!add(sc, it, sc.n == 0)) return PK_FALLBACK;
      }
    }
  }

  if (S->R != R || S->K != K) {
    if (S->pool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->pool); S->pool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->pool, pass == 0 ? 0 : S->pool_bytes);
      S->d_params = a.take<GroupParams>(1);
      S->d_counter = a.take<unsigned int>(128 + 2 * groups);
      S->d_dbg = a.take<long long>((2048 + 256) * PK_STAMPS + 64);
      hp.gsched = a.take<GSched>(sched.size());
      dp.d[DD_AH].out = a.take<float>((size_t)PK_MAX_SLOTS * R * A);
      for (int q = 0; q < 2; ++q) hp.fslots[q] = a.take<float>((size_t)groups * nslots_l[q] * PK_BN * 128);
      for (int q = 0; q < 2; ++q) { hp.hh_hi[q] = a.take<__half>((long)R * 2 * H); hp.hh_lo[q] = a.take<__half>((long)R * 2 * H); }
      dp.af_hi = reinterpret_cast<float*>(a.take<__half>((long)R * H)); dp.af_lo = reinterpret_cast<float*>(a.take<__half>((long)R * H));
      dp.EUv = a.take<float>((long)R * K * A);
      if (pass == 0) {
        S->pool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->pool, S->pool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->pool, 0, S->pool_bytes, st));
      }
    }
    S->R = R; S->K = K;
  }
  dp.hh_hi = nullptr; dp.hh_lo = nullptr; dp.xt_hi = dp.xt_lo = dp.gp_hi = dp.gp_lo = nullptr;
  dp.hx = nullptr; dp.cx = nullptr; dp.unfinished = nullptr; dp.tok = nullptr;
  hp.pick_ctr = S->d_counter + 64;
  hp.group_ctr = S->d_counter + 128;
  hp.members[0] = m1; hp.members[1] = m2; hp.groups = groups; hp.ncb = 1; hp.ntv = 0; hp.n_att = B;
  hp.nslots[0] = nslots_l[0]; hp.nslots[1] = nslots_l[1];
  hp.l2_hints = getenv("XG_L2_HINT") ? atoi(getenv("XG_L2_HINT")) : 1;
  hp.topk = 0; hp.lraw = nullptr; hp.lpart = nullptr;
  hp.sample_max = 1; hp.step_drop = 0; hp.inv_temp = 1.f; hp.sample_seed = 0; hp.ss_mode = 0; hp.mg_on = 0;
  hp.drop_gate = hp.drop_h1 = hp.drop_h2 = make_drop(false, 0.f, 0, 0);
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<GSched*>(hp.gsched), sched.data(), sizeof(GSched) * sched.size(),
                                       cudaMemcpyHostToDevice, st));
  WordTables* WT = nullptr;
  XG_TRY(word_tables(ctx, st, &WT, /*need_tgate=*/false, WT_TRAIN));
  MapTable2 mt;
  CUtensorMap* maps = mt.m;
  XG_TRY(word_weight_maps(ctx, ts, WT, maps));
  for (int q = 0; q < 2; ++q) {
    XG_TRY(make_map16(ctx, ts, hp.hh_hi[q], R, 2 * H, PK_BN, &maps[GM_HH + 2 * q]));
    XG_TRY(make_map16(ctx, ts, hp.hh_lo[q], R, 2 * H, PK_BN, &maps[GM_HH + 2 * q + 1]));
  }
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.af_hi), R, H, PK_BN, &maps[GM_AF]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.af_lo), R, H, PK_BN, &maps[GM_AF + 1]));
  maps[GM_XT] = maps[GM_XT + 1] = maps[GM_GP] = maps[GM_GP + 1] = maps[0];      // (hoisted: not streamed by this kernel)
  {
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)B * K};
    cuuint64_t strides[1] = {(cuuint64_t)H * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)(H / 2), (cuuint32_t)K};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult cr = ts->encode(&maps[GM_V], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Vf), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { ctx->es.set(__FILE__, __LINE__, "cuTensorMapEncodeTiled (V) failed", nullptr); return XG_ERR_CUDA; }
  }
  maps[27] = maps[0];

  dp.sched = nullptr;
  dp.B = B; dp.R = R; dp.K = K; dp.H = H; dp.E = d.embed; dp.Ep = 0; dp.A = A; dp.V = d.vocab; dp.T = T;
  dp.b_h2a = ctx->P[XG_P_H2A_B]; dp.w_a2w = ctx->P[XG_P_A2W_W]; dp.b_a2w = ctx->P[XG_P_A2W_B];
  dp.bias[0][0] = ctx->P[XG_P_L1_I2H_B]; dp.bias[0][1] = ctx->P[XG_P_L1_A2H_B]; dp.bias[0][2] = ctx->P[XG_P_L1_H2H_B];
  dp.bias[1][0] = ctx->P[XG_P_L2_I2H_B]; dp.bias[1][1] = ctx->P[XG_P_L2_A2H_B]; dp.bias[1][2] = ctx->P[XG_P_L2_H2H_B];
  dp.b_logit = nullptr; dp.embed = nullptr; dp.tgate = nullptr;
  dp.Vf = Vf; dp.Uv = Uv; dp.pos = nullptr;
  for (int q = 0; q < 4; ++q) dp.state0[q] = nullptr;
  dp.mode = 1; dp.feat_div = 1; dp.build_euv = 1; dp.x16 = 1;
  dp.L = tr.L; dp.seq_mask = tr.seq_mask;
  dp.G1s = tr.G1; dp.G2s = tr.G2; dp.C1s = tr.C1; dp.C2s = tr.C2; dp.H12s = tr.H12;
  dp.AHs = tr.AH; dp.ALPHAs = tr.ALPHA; dp.AFs = tr.AF;
  dp.drop1 = tr.drop1; dp.drop2 = tr.drop2;
  dp.seq = nullptr; dp.seqlogp = nullptr; dp.flags = nullptr;
  dp.sync_counter = S->d_counter;
  dp.dbg_clock = env_flag("XG_PERSIST_TRACE") ? S->d_dbg : nullptr;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_params, &hp, sizeof(GroupParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_counter, 0, sizeof(unsigned int) * (128 + 2 * groups), st));
  if (dp.dbg_clock) XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_dbg, 0, sizeof(long long) * ((2048 + 256) * PK_STAMPS + 64), st));
  if (!S->attr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(train_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, train_grouped_kernel, PK_THREADS, PK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "grouped training loop does not fit on an SM");
    S->attr_set = true;
  }
  {
    ProfScope ps(ctx, "train_decode_persistent", st);
    const GroupParams* gp = S->d_params;
    void* args[2] = {(void*)&gp, (void*)&mt};
    XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)train_grouped_kernel, dim3(G), dim3(PK_THREADS), args, PK_SMEM_BYTES, st));
    ctx->n_fused++;
  }
  if (dp.dbg_clock && T > 4 && G <= 256) {   // XG_PERSIST_TRACE=1: phase timeline of step 3, all CTAs
    XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
    std::vector<long long> ga((size_t)G * PK_STAMPS);
    cudaMemcpy(ga.data(), S->d_dbg + 2048 * PK_STAMPS, sizeof(long long) * ga.size(), cudaMemcpyDeviceToHost);
    const char* names[3] = {"A (attention query)", "B (attention || lstm_1)", "C (lstm_2 fused)"};
    for (int i = 0; i < 3; ++i) {
      long long open = 0, close = 0;
      for (int c = 0; c < G; ++c) open = std::max(open, ga[(size_t)c * PK_STAMPS + 2 * i]);
      std::vector<long long> fin(G);
      for (int c = 0; c < G; ++c) fin[c] = ga[(size_t)c * PK_STAMPS + 2 * i + 1] - open;
      std::vector<long long> srt = fin;
      std::sort(srt.begin(), srt.end());
      for (int c = 0; c < G; ++c) close = std::max(close, ga[(size_t)c * PK_STAMPS + 2 * i + 2]);
      const int worst = (int)(std::max_element(fin.begin(), fin.end()) - fin.begin());
      fprintf(stderr, "[xg grouped train trace] step 3 %-26s work done after: min %6lld  median %6lld  p90 %6lld  max %6lld ns (cta %d)   barrier exit %6lld ns\n",
              names[i], srt[0], srt[G / 2], srt[G * 9 / 10], srt[G - 1], worst, close - open);
    }
  }
  return XG_OK;
}

// chains of a phase schedule alternate between the two epilogue warpgroups (GSched::dual)
static void gsched_dual(GSched& sc) {
  int ch = 0, owner = 0, cb = 0;
  for (int i = 0; i < sc.n; ++i) {
    if (!(sc.it[i].flags & GI_CONT_PREV)) { owner = ch & 1; ++ch; }
    if (owner) cb += (sc.it[i].nkb + PK_CHUNK - 1) / PK_CHUNK;
  }
  sc.dual = 1; sc.chunks_b = (short)cb;
}

// ---- one word step on decode_step_grouped_kernel (beam search); PK_FALLBACK: shape / outputs outside it (caller:
//      decode_step_persistent_kernel through persist_decode) ----
struct GroupedStepState {
  int R = 0, K = 0;
  char* pool = nullptr;
  size_t pool_bytes = 0;
  GroupParams hp, hp_dev;
  GroupParams* d_params = nullptr;
  unsigned int* d_counter = nullptr;
  MapTable2 mt;
  bool valid = false, attr_set = false;
  int B = 0, fdiv = 0, topk = 0;
  const float *Vf = nullptr, *Uv = nullptr, *pos = nullptr;
  unsigned long long epoch = ~0ull;
  unsigned int sync_base = 0, launches = 0;
  long long* d_dbg = nullptr;
};
inline GroupedStepState*& grouped_step_state(xg_context* ctx) {
  static std::unordered_map<xg_context*, GroupedStepState*> m;
  return m[ctx];
}
static void grouped_step_release(xg_context* ctx) {
  GroupedStepState* s = grouped_step_state(ctx);
  if (!s) return;
  if (s->pool) cudaFree(s->pool);
  delete s;
  grouped_step_state(ctx) = nullptr;
}

static int grouped_step(xg_context* ctx, const float* Vf, const float* Uv, const float* pos, int B, int K, const PersistStepIO& io,
                        cudaStream_t st, const BeamMergeIO* mg = nullptr, int nsteps = 1) {
  // nsteps > 1: the word steps of positions mg->t .. mg->t + nsteps - 1 in ONE launch (needs the in-kernel merge)
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, A = d.att, V = d.vocab;
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN, Ep = (E + GK_KB - 1) / GK_KB * GK_KB, G = ctx->sm_count;
  if (env_flag("XG_NO_GROUPED") || env_flag("XG_NO_GROUPED_STEP") || !persist_eligible(ctx, B, K) || H % GK_KB != 0 ||
      Ep > DEC_TI * PK_THREADS || 4 * H > 32000)
    return PK_FALLBACK;
  if (io.logp != nullptr || io.ys == nullptr || io.ix == nullptr || io.topk < 1 || io.topk > GK_TOPK) return PK_FALLBACK;
  if (nsteps < 1 || (nsteps > 1 && !mg)) return PK_FALLBACK;
  if (mg && (mg->beam != io.topk || mg->beam != io.feat_div || mg->beam > PK_WARPS || mg->B * mg->beam != B)) return PK_FALLBACK;
  if (!att_rows_fit(d, K, ATT_NR)) return PK_FALLBACK;
  GroupedStepState*& S = grouped_step_state(ctx);
  if (!S) S = new GroupedStepState();
  GroupParams& hp = S->hp;
  DecParams& dp = hp.dp;
  auto step_io = [&]() {
    dp.build_euv = io.first;
    dp.tokens_in = io.tokens; dp.logp_out = nullptr;
    dp.parent_in = io.parent; dp.ys_out = io.ys; dp.ix_out = io.ix; dp.topk = io.topk;
    hp.topk = io.topk;
    for (int q = 0; q < 4; ++q) { dp.state0[q] = io.state[q]; dp.state_out[q] = io.state[q]; }
    hp.mg_on = mg ? 1 : 0;
    if (mg) hp.mg = *mg; else memset(&hp.mg, 0, sizeof(hp.mg));
  };
  auto launch_step = [&]() -> int {
    ProfScope ps(ctx, "decode_step_persistent", st);
    const GroupParams* gp = S->d_params;
    unsigned int base = S->sync_base, ep = S->launches;
    int ns = nsteps;
    void* args[5] = {(void*)&gp, (void*)&S->mt, (void*)&base, (void*)&ep, (void*)&ns};
    XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)decode_step_grouped_kernel, dim3(G), dim3(PK_THREADS), args, GK_SMEM_BYTES, st));
    ctx->n_fused++;
    S->sync_base += ((GK_STEP_BARRIERS + 1) * (unsigned int)nsteps - 1) * (unsigned int)G;
    S->launches += (unsigned int)nsteps;
    if (S->hp.dp.dbg_clock && S->launches == 6 && G <= 256) {   // XG_PERSIST_TRACE=1: phase timeline of the sixth step, all CTAs
      XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st));
      std::vector<long long> ga((size_t)G * PK_STAMPS);
      cudaMemcpy(ga.data(), S->d_dbg + 2048 * PK_STAMPS, sizeof(long long) * ga.size(), cudaMemcpyDeviceToHost);
      const char* names[6] = {"prologue", "A (lstm_1 fused, ah)", "B (attention)", "C (lstm_2 fused)", "D (logits tiles)", "E (row merge, states)"};
      for (int i = 0; i < 6; ++i) {
        long long open = 0, close = 0;
        for (int c = 0; c < G; ++c) open = std::max(open, ga[(size_t)c * PK_STAMPS + 2 * i]);
        std::vector<long long> fin(G);
        for (int c = 0; c < G; ++c) fin[c] = ga[(size_t)c * PK_STAMPS + 2 * i + 1] - open;
        std::vector<long long> srt = fin;
        std::sort(srt.begin(), srt.end());
        if (i < 5) for (int c = 0; c < G; ++c) close = std::max(close, ga[(size_t)c * PK_STAMPS + 2 * i + 2]);
        const int worst = (int)(std::max_element(fin.begin(), fin.end()) - fin.begin());
        fprintf(stderr, "[xg grouped step trace] %-26s work done after: min %6lld  median %6lld  p90 %6lld  max %6lld ns (cta %d)   barrier exit %6lld ns\n",
                names[i], srt[0], srt[G / 2], srt[G * 9 / 10], srt[G - 1], worst, i < 5 ? close - open : 0LL);
      }
      if (ga[12] && ga[13])      // CTA 0 merges video 0: row merges done / candidate merge done, after its entry into phase E
        fprintf(stderr, "[xg grouped step trace] cta 0, phase E: row merges done after %lld ns, candidate merge after %lld ns, states out after %lld ns\n",
                ga[12] - ga[10], ga[13] - ga[10], ga[11] - ga[10]);
#ifdef GK_FINE
      {
        long long f[32];
        cudaMemcpy(f, S->d_dbg + (2048 + 256) * PK_STAMPS, sizeof(f), cudaMemcpyDeviceToHost);
        for (int l = 0; l < 2; ++l)
          fprintf(stderr, "[xg grouped step trace] cta 0 fused layer %d (cycles after phase entry): producer done %lld, epilogue warp done %lld, "
                  "cta synced %lld, group counter seen %lld, cta synced %lld, cell done %lld\n", l, f[l * 16 + 1] - f[l * 16], f[l * 16 + 2] - f[l * 16],
                  f[l * 16 + 3] - f[l * 16], f[l * 16 + 4] - f[l * 16], f[l * 16 + 5] - f[l * 16], f[l * 16 + 6] - f[l * 16]);
        long long w[32];
        cudaMemcpy(w, S->d_dbg + (2048 + 256) * PK_STAMPS + 32, sizeof(w), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[xg grouped step trace] cta 0 attention (cycles after entry): weights loaded %lld, slot loads issued %lld, copies issued %lld, staged %lld, "
                "query ready %lld, chunks done %lld, softmax done %lld, V landed %lld, contexts stored %lld\n",
                w[28] - w[22], w[29] - w[22], w[30] - w[22], w[31] - w[22], w[23] - w[22], w[24] - w[22], w[25] - w[22], w[26] - w[22], w[27] - w[22]);
        fprintf(stderr, "[xg grouped step trace] cta 0, all GEMM phases of %u launches, %lld k-blocks per launch: producer %lld cycles in phases, %lld waiting for a free stage | "
                "MMA warp %lld in phases, waits: data %lld, accumulator free %lld, cross accumulator free %lld | epilogue warp %lld in phases, waits: accumulator full %lld, drains %lld, item ends %lld, logits item ends %lld\n",
                S->launches, w[4] / S->launches, w[0], w[1], w[8], w[9], w[10], w[11], w[16], w[17], w[18], w[19], w[21]);
        {
          fprintf(stderr, "[xg grouped step trace] logits quarter passes (warp 2): transposed stores %lld, barrier %lld, reduce %lld, result store + barrier %lld\n", w[6], w[7], w[14], w[15]);
        }
      }
#endif
    }
    return XG_OK;
  };
  if (S->valid && S->R == R && S->K == K && S->B == B && S->fdiv == io.feat_div && S->Vf == Vf && S->Uv == Uv && S->pos == pos &&
      S->epoch == ctx->param_epoch && !io.first) {
    // later steps of the same search: schedule, tensor maps and counters of the last launch (the barrier and group
    // counters run on from launch to launch)
    step_io();
    if (memcmp(&hp, &S->hp_dev, sizeof(GroupParams)) != 0) {
      XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_params, &hp, sizeof(GroupParams), cudaMemcpyHostToDevice, st));
      memcpy(&S->hp_dev, &hp, sizeof(GroupParams));
    }
    return launch_step();
  }
  S->valid = false;

  const int kbH = H / GK_KB, kbE = Ep / GK_KB;
  const int ntiles = H / 32, ncb = R / PK_BN, groups = ntiles * ncb;
  const int ntv = (V + 127) / 128, nlog = ntv * ncb, nat = (A + 127) / 128;
  if (ntiles >= G || ntv > 256 || (nlog + G - 1) / G > GK_MAX_ITEMS) return PK_FALLBACK;
  // A CTA of a fused phase is member `mem` of a 32-unit tile and runs that member's K slice once per caption column
  // block (fused_cell_phase).  Phase A shares the grid between the lstm_1 tiles and the attention query (split-K
  // slots on the remaining "side" CTAs): the member count that balances the two is taken.
  const int Ktot0 = kbE + 2 * kbH, Ktot1 = 3 * kbH;
  int members0 = 0, ah_slots = 1, best_load = 1 << 30;
  const int max_members = ncb > 1 ? 8 : GK_MAX_MEMBERS;      // (several column blocks: the cells take two (block, caption) pairs per pass, 8 slots)
  for (int m = 1; m <= max_members && ntiles * m < G; ++m) {
    const int nside = G - ntiles * m;
    for (int sl = 1; sl <= std::min(4, 2 * kbH); sl *= 2) {
      const int per = (nat * ncb * sl + nside - 1) / nside;
      if (per > GK_MAX_ITEMS || 2 * ncb > GK_MAX_ITEMS) continue;
      const int load = std::max(ncb * ((Ktot0 + m - 1) / m), per * ((2 * kbH + sl - 1) / sl));
      if (load < best_load) { best_load = load; members0 = m; ah_slots = sl; }
    }
  }
  if (members0 == 0 || ah_slots > PK_MAX_SLOTS) return PK_FALLBACK;
  const int nside = G - ntiles * members0;
  const int members_l[2] = {members0, std::max(1, std::min(std::min(max_members, G / ntiles), Ktot1))};
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));

  for (int i = 0; i < PK_MAX_DESCS; ++i) { dp.d[i].ns = 0; dp.d[i].n_rows = 0; dp.d[i].nkb = 0; }
  {
    GDesc& g = dp.d[DD_AH];
    g.w_map = GM_H2A; g.x_hi = GM_HH; g.x_lo = GM_HH + 1; g.xkb0 = 0; g.n_rows = A; g.nkb = 2 * kbH;
    GDesc& l = dp.d[DD_LOGIT];
    l.w_map = GM_LOGIT; l.x_hi = GM_HH; l.x_lo = GM_HH + 1; l.xkb0 = kbH; l.n_rows = V; l.nkb = kbH;
  }
  std::vector<GSched> sched((size_t)3 * G);
  memset(sched.data(), 0, sizeof(GSched) * sched.size());
  struct Prod { int w_map, x_map, xsel, xkb0, nkb; };
  // the state entering the step sits in buffer 0 (xsel 1 at parity 0), the cells write buffer 1 (xsel 2)
  const Prod layers[2][3] = {{{GM_W32 + 0, GM_XT, 0, 0, kbE}, {GM_W32 + 2, GM_GP, 0, 0, kbH}, {GM_W32 + 4, GM_HH, 1, 0, kbH}},
                             {{GM_W32 + 6, GM_HH, 2, 0, kbH}, {GM_W32 + 8, GM_AF, 0, 0, kbH}, {GM_W32 + 10, GM_HH, 1, kbH, kbH}}};
  for (int layer = 0; layer < 2; ++layer) {
    const int members = members_l[layer];
    int Ktot = 0;
    for (int p = 0; p < 3; ++p) Ktot += layers[layer][p].nkb;
    for (int tile = 0; tile < ntiles; ++tile)
      for (int mem = 0; mem < members; ++mem) {
        GSched& sc = sched[(size_t)layer * G + tile * members + mem];
        const int k0 = (int)((long)mem * Ktot / members), k1 = (int)((long)(mem + 1) * Ktot / members);
        for (int cb = 0; cb < ncb && k0 < k1; ++cb) {      // one chain per column block
          int base = 0, first = 1;
          for (int p = 0; p < 3; ++p) {
            const Prod& pr = layers[layer][p];
            const int lo = std::max(k0, base), hi = std::min(k1, base + pr.nkb);
            if (lo < hi) {
              if (sc.n >= GK_MAX_ITEMS) return PK_FALLBACK;
              GItem it{};
              it.w_map = (short)pr.w_map; it.x_map = (short)pr.x_map; it.xsel = (short)pr.xsel;
              it.flags = (short)(GI_FUSED | (layer ? GI_LAYER1 : 0) | (first ? 0 : GI_CONT_PREV));
              it.wrow = (short)(tile * 32); it.wk0 = (short)(lo - base); it.xk0 = (short)(pr.xkb0 + lo - base); it.nkb = (short)(hi - lo);
              it.desc = (short)(tile * ncb + cb); it.slot = (short)mem; it.cb = (short)cb; it.pad = (short)members;
              if (!first) sc.it[sc.n - 1].flags |= GI_CONT_NEXT;
              sc.it[sc.n++] = it;
              sc.tot_kb += (short)(hi - lo);
              sc.tot_chunks += (short)((hi - lo + PK_CHUNK - 1) / PK_CHUNK);
              first = 0;
            }
            base += pr.nkb;
          }
          sc.n_chains++;
        }
      }
  }
  {   // attention query W_h2a.[h1|h2] (state entering the step): split-K slots on the side CTAs of phase A
    const int ah_run = (2 * kbH + ah_slots - 1) / ah_slots;
    dp.d[DD_AH].ns = (2 * kbH + ah_run - 1) / ah_run;
    int c = 0;
    for (int rt = 0; rt < nat; ++rt)
      for (int cb = 0; cb < ncb; ++cb)
        for (int k0 = 0, sl = 0; k0 < 2 * kbH; k0 += ah_run, ++sl) {
          GItem it{};
          it.w_map = (short)GM_H2A; it.x_map = (short)GM_HH; it.xsel = 1; it.flags = 0;
          it.wrow = (short)(rt * 128); it.wk0 = (short)k0; it.xk0 = (short)k0; it.nkb = (short)std::min(ah_run, 2 * kbH - k0);
          it.desc = DD_AH; it.slot = (short)sl; it.cb = (short)cb;
          GSched& sc = sched[(size_t)ntiles * members0 + (c++ % nside)];
          if (sc.n >= GK_MAX_ITEMS) return PK_FALLBACK;
          sc.it[sc.n++] = it;
          sc.tot_kb += it.nkb; sc.tot_chunks += (short)((it.nkb + PK_CHUNK - 1) / PK_CHUNK); sc.n_chains++;
        }
  }
  for (int i = 0; i < nlog; ++i) {   // logits: (vocabulary tile, column block) items, a tile's column blocks side by side
    const int c = (int)((long)i * G / nlog);
    GSched& sc = sched[(size_t)2 * G + c];
    if (sc.n >= GK_MAX_ITEMS) return PK_FALLBACK;
    GItem it{};
    it.w_map = (short)GM_LOGIT; it.x_map = (short)GM_HH; it.xsel = 2; it.flags = GI_LOGITS;
    it.wrow = (short)((i / ncb) * 128); it.wk0 = 0; it.xk0 = (short)kbH; it.nkb = (short)kbH;
    it.desc = DD_LOGIT; it.slot = 0; it.cb = (short)(i % ncb);
    sc.it[sc.n++] = it;
    sc.tot_kb += (short)kbH; sc.tot_chunks += (short)((kbH + PK_CHUNK - 1) / PK_CHUNK); sc.n_chains++;
  }

  if (ncb > 1 && !env_flag("XG_NO_DUAL"))      // several chains per CTA: both epilogue warpgroups
    for (auto& sc : sched) gsched_dual(sc);
  if (S->R != R || S->K != K) {
    if (S->pool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->pool); S->pool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->pool, pass == 0 ? 0 : S->pool_bytes);
      S->d_params = a.take<GroupParams>(1);
      S->d_counter = a.take<unsigned int>(128 + 2 * groups);
      S->d_dbg = a.take<long long>((2048 + 256) * PK_STAMPS + 64);
      hp.gsched = a.take<GSched>(sched.size());
      hp.lpart = a.take<float4>((size_t)R * ntv);
      hp.lraw = a.take<float>((size_t)R * ntv * 128);
      dp.d[DD_AH].out = a.take<float>((size_t)PK_MAX_SLOTS * R * A);
      for (int q = 0; q < 2; ++q) hp.fslots[q] = a.take<float>((size_t)groups * members_l[q] * PK_BN * 128);   // (members depend on R only)
      dp.xt_hi = reinterpret_cast<float*>(a.take<__half>((long)R * Ep)); dp.xt_lo = reinterpret_cast<float*>(a.take<__half>((long)R * Ep));
      for (int q = 0; q < 2; ++q) { hp.hh_hi[q] = a.take<__half>((long)R * 2 * H); hp.hh_lo[q] = a.take<__half>((long)R * 2 * H); }
      dp.gp_hi = reinterpret_cast<float*>(a.take<__half>((long)R * H)); dp.gp_lo = reinterpret_cast<float*>(a.take<__half>((long)R * H));
      dp.af_hi = reinterpret_cast<float*>(a.take<__half>((long)R * H)); dp.af_lo = reinterpret_cast<float*>(a.take<__half>((long)R * H));
      dp.hx = a.take<float>((long)R * 2 * H);
      dp.cx = a.take<float>((long)2 * R * H);
      dp.unfinished = a.take<float>(R);
      dp.tok = a.take<int64_t>(R);
      dp.EUv = a.take<float>((long)R * K * A);
      if (pass == 0) {
        S->pool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->pool, S->pool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->pool, 0, S->pool_bytes, st));
      }
    }
    S->R = R; S->K = K;
  }
  dp.hh_hi = nullptr; dp.hh_lo = nullptr;
  hp.pick_ctr = S->d_counter + 64;
  hp.group_ctr = S->d_counter + 128;
  hp.members[0] = members_l[0]; hp.members[1] = members_l[1]; hp.groups = groups; hp.ncb = ncb; hp.ntv = ntv; hp.n_att = 0;
  hp.nslots[0] = members_l[0]; hp.nslots[1] = members_l[1];
  hp.l2_hints = getenv("XG_L2_HINT") ? atoi(getenv("XG_L2_HINT")) : 1;
  hp.sample_max = 1; hp.step_drop = 0; hp.inv_temp = 1.f; hp.sample_seed = 0; hp.ss_mode = 0; hp.mg_on = 0;
  hp.drop_gate = hp.drop_h1 = hp.drop_h2 = make_drop(false, 0.f, 0, 0);
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<GSched*>(hp.gsched), sched.data(), sizeof(GSched) * sched.size(),
                                       cudaMemcpyHostToDevice, st));
  WordTables* WT = nullptr;
  XG_TRY(word_tables(ctx, st, &WT));
  CUtensorMap* maps = S->mt.m;
  XG_TRY(word_weight_maps(ctx, ts, WT, maps));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.xt_hi), R, Ep, PK_BN, &maps[GM_XT]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.xt_lo), R, Ep, PK_BN, &maps[GM_XT + 1]));
  for (int q = 0; q < 2; ++q) {
    XG_TRY(make_map16(ctx, ts, hp.hh_hi[q], R, 2 * H, PK_BN, &maps[GM_HH + 2 * q]));
    XG_TRY(make_map16(ctx, ts, hp.hh_lo[q], R, 2 * H, PK_BN, &maps[GM_HH + 2 * q + 1]));
  }
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.gp_hi), R, H, PK_BN, &maps[GM_GP]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.gp_lo), R, H, PK_BN, &maps[GM_GP + 1]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.af_hi), R, H, PK_BN, &maps[GM_AF]));
  XG_TRY(make_map16(ctx, ts, reinterpret_cast<__half*>(dp.af_lo), R, H, PK_BN, &maps[GM_AF + 1]));
  {
    const int nvid = (B + io.feat_div - 1) / io.feat_div;
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)nvid * K};
    cuuint64_t strides[1] = {(cuuint64_t)H * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)(H / 2), (cuuint32_t)K};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult cr = ts->encode(&maps[GM_V], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Vf), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { ctx->es.set(__FILE__, __LINE__, "cuTensorMapEncodeTiled (V) failed", nullptr); return XG_ERR_CUDA; }
  }
  maps[27] = maps[0];

  dp.sched = nullptr;
  dp.B = B; dp.R = R; dp.K = K; dp.H = H; dp.E = E; dp.Ep = Ep; dp.A = A; dp.V = V; dp.T = 1;
  dp.b_h2a = ctx->P[XG_P_H2A_B]; dp.w_a2w = ctx->P[XG_P_A2W_W]; dp.b_a2w = ctx->P[XG_P_A2W_B];
  dp.bias[0][0] = ctx->P[XG_P_L1_I2H_B]; dp.bias[0][1] = ctx->P[XG_P_L1_A2H_B]; dp.bias[0][2] = ctx->P[XG_P_L1_H2H_B];
  dp.bias[1][0] = ctx->P[XG_P_L2_I2H_B]; dp.bias[1][1] = ctx->P[XG_P_L2_A2H_B]; dp.bias[1][2] = ctx->P[XG_P_L2_H2H_B];
  dp.b_logit = ctx->P[XG_P_LOGIT_B]; dp.embed = ctx->P[XG_P_EMBED_W];
  dp.tgate = WT->tgate;
  dp.Vf = Vf; dp.Uv = Uv; dp.pos = pos;
  dp.mode = 0; dp.feat_div = io.feat_div; dp.x16 = 1;
  dp.seq = nullptr; dp.seqlogp = nullptr; dp.flags = nullptr;
  dp.sync_counter = S->d_counter;
  dp.dbg_clock = env_flag("XG_PERSIST_TRACE") ? S->d_dbg : nullptr;
  step_io();
  dp.build_euv = 1;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_params, &hp, sizeof(GroupParams), cudaMemcpyHostToDevice, st));
  memcpy(&S->hp_dev, &hp, sizeof(GroupParams));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_counter, 0, sizeof(unsigned int) * (128 + 2 * groups), st));
  S->sync_base = 0; S->launches = 0;
  if (!S->attr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(decode_step_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GK_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, decode_step_grouped_kernel, PK_THREADS, GK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "grouped word step does not fit on an SM");
    S->attr_set = true;
  }
  XG_TRY(launch_step());
  S->valid = true; S->B = B; S->fdiv = io.feat_div; S->Vf = Vf; S->Uv = Uv; S->pos = pos; S->epoch = ctx->param_epoch;
  return XG_OK;
}

// ---- encoder frame recurrence on encode_grouped_kernel; PK_FALLBACK: shape outside it (caller: encode_persistent_kernel) ----
struct GroupedEncState {
  int R = 0;
  char* pool = nullptr;
  size_t pool_bytes = 0;
  GroupParams hp;
  EncParams ep;
  GroupParams* d_gparams = nullptr;
  EncParams* d_eparams = nullptr;
  unsigned int* d_counter = nullptr;
  __half* w16[2][2] = {};
  unsigned long long w_epoch = ~0ull;
  bool attr_set = false;
};
inline GroupedEncState*& grouped_enc_state(xg_context* ctx) {
  static std::unordered_map<xg_context*, GroupedEncState*> m;
  return m[ctx];
}
static void grouped_enc_release(xg_context* ctx) {
  GroupedEncState* s = grouped_enc_state(ctx);
  if (!s) return;
  if (s->pool) cudaFree(s->pool);
  delete s;
  grouped_enc_state(ctx) = nullptr;
}

static int grouped_encode(xg_context* ctx, const float* fmask, int B, int K, EncBufs& eb, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, G = ctx->sm_count;
  if (env_flag("XG_NO_GROUPED") || !ctx->persist_mode || H % GK_KB != 0 || B < 1 || K < 2 || G > 256 || 4 * H > 32000) return PK_FALLBACK;
  const int R = (B + PK_BN - 1) / PK_BN * PK_BN;
  const int kbH = H / GK_KB, ntile = H / 32, ncb = R / PK_BN, groups = 2 * ntile * ncb;
  if (groups > G) return PK_FALLBACK;
  int members = 0;
  for (int m = std::min(std::min(GK_MAX_MEMBERS, G / groups), kbH); m >= 1; --m) {
    const int q = kbH / m;
    if (kbH % m == 0 && (q == 1 || q == 2 || q == 4)) { members = m; break; }
  }
  if (members == 0) return PK_FALLBACK;      // the member's weight tiles must fit (and tile) the 4 pipeline stages
  TcState* ts = nullptr;
  XG_TRY(tc_init(ctx, ts));
  GroupedEncState*& S = grouped_enc_state(ctx);
  if (!S) S = new GroupedEncState();
  GroupParams& hp = S->hp;
  EncParams& ep = S->ep;
  std::vector<GSched> sched(G);
  memset(sched.data(), 0, sizeof(GSched) * sched.size());
  for (int grp = 0; grp < groups; ++grp) {
    const int s = (grp / ncb) / ntile, tile = (grp / ncb) % ntile, cb = grp % ncb;
    for (int mem = 0; mem < members; ++mem) {
      GSched& sc = sched[(size_t)grp * members + mem];
      const int k0 = mem * kbH / members, k1 = (mem + 1) * kbH / members;
      GItem it{};
      it.w_map = (short)(2 * s); it.x_map = 4; it.xsel = 1; it.flags = GI_FUSED;
      it.wrow = (short)(tile * 32); it.wk0 = (short)k0; it.xk0 = (short)(s * kbH + k0); it.nkb = (short)(k1 - k0);
      it.desc = (short)grp; it.slot = (short)mem; it.cb = (short)cb; it.pad = (short)members;
      sc.it[sc.n++] = it;
      sc.tot_kb = (short)(k1 - k0); sc.tot_chunks = (short)((k1 - k0 + PK_CHUNK - 1) / PK_CHUNK); sc.n_chains = 1;
    }
  }
  if (S->R != R) {
    if (S->pool) { XG_CUDA_TRY(ctx->es, cudaStreamSynchronize(st)); cudaFree(S->pool); S->pool = nullptr; }
    for (int pass = 0; pass < 2; ++pass) {
      Arena a(pass == 0 ? nullptr : S->pool, pass == 0 ? 0 : S->pool_bytes);
      S->d_gparams = a.take<GroupParams>(1);
      S->d_eparams = a.take<EncParams>(1);
      S->d_counter = a.take<unsigned int>(64 + groups);
      hp.gsched = a.take<GSched>(sched.size());
      hp.fslots[0] = a.take<float>((size_t)groups * members * PK_BN * 128);
      for (int q = 0; q < 2; ++q) { hp.hh_hi[q] = a.take<__half>((long)R * 2 * H); hp.hh_lo[q] = a.take<__half>((long)R * 2 * H); }
      for (int s = 0; s < 2; ++s) for (int q = 0; q < 2; ++q) S->w16[s][q] = a.take<__half>((long)4 * H * H);
      if (pass == 0) {
        S->pool_bytes = a.off + 1024;
        XG_CUDA_TRY(ctx->es, cudaMalloc(&S->pool, S->pool_bytes));
        XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->pool, 0, S->pool_bytes, st));
      }
    }
    S->R = R;
    S->w_epoch = ~0ull;
  }
  hp.fslots[1] = hp.fslots[0];
  hp.group_ctr = S->d_counter + 64;
  hp.members[0] = hp.members[1] = members; hp.nslots[0] = hp.nslots[1] = members;
  hp.groups = groups; hp.ncb = ncb; hp.ntv = 0; hp.n_att = 0; hp.l2_hints = 1; hp.lpart = nullptr; hp.pick_ctr = nullptr;
  hp.dp.R = R; hp.dp.H = H; hp.dp.B = B; hp.dp.V = 0;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(const_cast<GSched*>(hp.gsched), sched.data(), sizeof(GSched) * sched.size(), cudaMemcpyHostToDevice, st));
  const int whh[2] = {XG_P_LSTM_RGB_WHH, XG_P_LSTM_OPFL_WHH};
  if (S->w_epoch != ctx->param_epoch) {
    for (int s = 0; s < 2; ++s) {
      ProfScope ps(ctx, "split_weights_f16", st);
      split_weights_f16_kernel<<<ctx->sm_count * 2, 256, 0, st>>>(ctx->P[whh[s]], 4 * H, H, H, S->w16[s][0], S->w16[s][1]);
      XG_LAUNCH_CHECK(ctx->es);
    }
    S->w_epoch = ctx->param_epoch;
  }
  MapTable2 mt;
  for (int s = 0; s < 2; ++s)
    for (int q = 0; q < 2; ++q) XG_TRY(make_map16(ctx, ts, S->w16[s][q], 4 * H, H, 32, &mt.m[2 * s + q]));
  for (int q = 0; q < 2; ++q) {
    XG_TRY(make_map16(ctx, ts, hp.hh_hi[q], R, 2 * H, PK_BN, &mt.m[4 + 2 * q]));
    XG_TRY(make_map16(ctx, ts, hp.hh_lo[q], R, 2 * H, PK_BN, &mt.m[4 + 2 * q + 1]));
  }
  for (int i = 8; i < 28; ++i) mt.m[i] = mt.m[0];
  ep.B = B; ep.R = R; ep.K = K; ep.H = H;
  for (int s = 0; s < 2; ++s) { ep.Gt[s] = eb.G[s]; ep.Hs[s] = eb.Hs[s]; ep.Cs[s] = eb.Cs[s]; }
  ep.fmask = fmask;
  ep.sync_counter = S->d_counter;
  ep.sched = nullptr; ep.hh_hi = nullptr; ep.hh_lo = nullptr;
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_gparams, &hp, sizeof(GroupParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemcpyAsync(S->d_eparams, &ep, sizeof(EncParams), cudaMemcpyHostToDevice, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->d_counter, 0, sizeof(unsigned int) * (64 + groups), st));
  if (!S->attr_set) {
    XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(encode_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM_BYTES));
    int nb = 0;
    XG_CUDA_TRY(ctx->es, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, encode_grouped_kernel, PK_THREADS, PK_SMEM_BYTES));
    XG_REQUIRE(ctx->es, nb >= 1, XG_ERR_CUDA, "grouped encoder does not fit on an SM");
    S->attr_set = true;
  }
  ProfScope ps(ctx, "encode_persistent", st);
  const GroupParams* gp = S->d_gparams;
  const EncParams* epp = S->d_eparams;
  void* args[3] = {(void*)&gp, (void*)&epp, (void*)&mt};
  XG_CUDA_TRY(ctx->es, cudaLaunchCooperativeKernel((void*)encode_grouped_kernel, dim3(G), dim3(PK_THREADS), args, PK_SMEM_BYTES, st));
  ctx->n_fused++;
  return XG_OK;
}

}  // namespace xg
