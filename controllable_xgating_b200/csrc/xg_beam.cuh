// Batched on-device beam search: SAModel.sample_beam + CaptionModel.beam_search
// (SAModel.py:129-161, CaptionModel.py:22-128).  The reference walks the videos one at a time in
// Python and sorts the whole (beam, V) log-prob matrix on the host every step; here all B videos x
// `beam` rows advance together (row = video*beam + q) and only a per-row top-`beam` selection is made.
//
// Semantics kept bit-for-bit (PyTorch-0.3 scalar rules, see oracle/xgating_oracle.py):
//   - UNK (id 1) log-prob lowered by 1000 before selection (CaptionModel.py:94)
//   - candidates enumerated column-major (sorted position c outer, beam q inner), scored
//     double(sum[q]) + double(logp), ordered by a STABLE descending sort (:45-51)
//   - at t == 0 only beam 0 is expanded (:43-44)
//   - a beam that emits id 0 (or any beam at t == T-1) is appended to done_beams and its running sum
//     is set to -1000, but it stays in the beam (:108-118)
//   - result = stable sort of done_beams by score, first `beam` entries (:127)
#pragma once
#include "xg_forward.cuh"

namespace xg {

#define XG_MAX_BEAM 16

// per-row top-`beam` of the log-probs with the UNK penalty; one CTA per row.
// order: value descending, lowest index first on exact ties.
__global__ void beam_topk_kernel(const float* __restrict__ logp, int V, int beam, float* __restrict__ ys,
                                 int* __restrict__ ix) {
  __shared__ float red[32];
  __shared__ int redi[32];
  __shared__ int chosen[XG_MAX_BEAM];
  const int row = blockIdx.x;
  const float* x = logp + (long)row * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int c = 0; c < beam; ++c) {
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int j = threadIdx.x; j < V; j += blockDim.x) {
      bool skip = false;
      for (int u = 0; u < c; ++u) skip |= (chosen[u] == j);
      if (skip) continue;
      float v = x[j];
      if (j == 1) v -= 1000.f;
      if (v > best || (v == best && j < bi)) { best = v; bi = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { red[warp] = best; redi[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      best = red[0]; bi = redi[0];
      for (int w = 1; w < nw; ++w)
        if (red[w] > best || (red[w] == best && redi[w] < bi)) { best = red[w]; bi = redi[w]; }
      chosen[c] = bi;
      ys[(long)row * beam + c] = best;
      ix[(long)row * beam + c] = bi;
    }
    __syncthreads();
  }
}

struct BeamBufs {
  float* ys;          // (n, beam)
  int* ix;            // (n, beam)
  int64_t* seq[2];    // (B, beam, T)   ping-pong
  float* lps[2];      // (B, beam, T)
  float* sum;         // (B, beam)
  int* parent;        // (n)
  int64_t* tokens;    // (n)
  int64_t* done_seq;  // (B, T*beam, T)
  float* done_lps;    // (B, T*beam, T)
  float* done_p;      // (B, T*beam)
  int* done_n;        // (B)
  float* st[2][4];    // state ping-pong: h1,c1,h2,c2 each (n,H)
  float* st0[4];      // (B,H) init state per video
  float* meanV;       // (B,H)
  float* Uv;          // (B,K,A)
  float* logits;      // (n,V)
  StepBufs step;
};

inline void carve_beam(Arena& a, const xg_dims& d, int B, int K, int T, int beam, BeamBufs& w) {
  const long n = (long)B * beam, H = d.rnn;
  w.ys = a.take<float>(n * beam);
  w.ix = a.take<int>(n * beam);
  for (int s = 0; s < 2; ++s) { w.seq[s] = a.take<int64_t>(n * T); w.lps[s] = a.take<float>(n * T); }
  w.sum = a.take<float>(n);
  w.parent = a.take<int>(n);
  w.tokens = a.take<int64_t>(n);
  w.done_seq = a.take<int64_t>(n * T * T);
  w.done_lps = a.take<float>(n * T * T);
  w.done_p = a.take<float>(n * T);
  w.done_n = a.take<int>(B);
  for (int s = 0; s < 2; ++s)
    for (int q = 0; q < 4; ++q) w.st[s][q] = a.take<float>(n * H);
  for (int q = 0; q < 4; ++q) w.st0[q] = a.take<float>((long)B * H);
  w.meanV = a.take<float>((long)B * H);
  w.Uv = a.take<float>((long)B * K * d.att);
  w.logits = a.take<float>(n * d.vocab);
  carve_step(a, d, (int)n, w.step);
}

// one warp per video: candidate merge + bookkeeping of beam_step and the done-beam harvest (beam_merge_warp, xg_grouped.cuh:
// the grouped step kernel runs the same function at the end of its launch)
constexpr int BEAM_MERGE_WARPS = 4;
__global__ void beam_merge_kernel(const BeamMergeIO M) {
  __shared__ double s_cp[BEAM_MERGE_WARPS][XG_MAX_BEAM * XG_MAX_BEAM];
  __shared__ int s_sel[BEAM_MERGE_WARPS][XG_MAX_BEAM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * BEAM_MERGE_WARPS + warp;
  if (k >= M.B) return;
  beam_merge_warp(M, k, s_cp[warp], s_sel[warp], lane);
}

// dst[row,:] = src[parent[row],:] for the four state tensors
__global__ void beam_gather_state_kernel(const int* __restrict__ parent, int H, const float* __restrict__ s0,
                                         const float* __restrict__ s1, const float* __restrict__ s2,
                                         const float* __restrict__ s3, float* __restrict__ d0, float* __restrict__ d1,
                                         float* __restrict__ d2, float* __restrict__ d3) {
  const int row = blockIdx.x;
  const long src = (long)parent[row] * H, dst = (long)row * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    d0[dst + j] = s0[src + j]; d1[dst + j] = s1[src + j]; d2[dst + j] = s2[src + j]; d3[dst + j] = s3[src + j];
  }
}

// dst[(k*beam+q),:] = src[k,:]
__global__ void beam_expand_kernel(const float* __restrict__ src, int beam, int H, float* __restrict__ dst) {
  const int row = blockIdx.x;
  const long s = (long)(row / beam) * H, d = (long)row * H;
  for (int j = threadIdx.x; j < H; j += blockDim.x) dst[d + j] = src[s + j];
}

// final ranking: stable sort of the done list by descending score, top `beam` out; one thread per video
__global__ void beam_finalize_kernel(const int64_t* __restrict__ done_seq, const float* __restrict__ done_lps,
                                     const float* __restrict__ done_p, const int* __restrict__ done_n, int B, int beam,
                                     int T, int64_t* __restrict__ seq_out, float* __restrict__ logp_out,
                                     int64_t* __restrict__ out_seq, float* __restrict__ out_lps,
                                     float* __restrict__ out_p, int32_t* __restrict__ out_n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B) return;
  const int n = done_n[k];
  const float* p = done_p + (long)k * T * beam;
  const int keep = n < beam ? n : beam;
  // selection of the `keep` best, stable: among equal scores the earlier entry wins
  int picked[XG_MAX_BEAM];
  for (int r = 0; r < keep; ++r) {
    int best = -1;
    for (int e = 0; e < n; ++e) {
      bool used = false;
      for (int u = 0; u < r; ++u) used |= (picked[u] == e);
      if (used) continue;
      if (best < 0 || p[e] > p[best]) best = e;
    }
    picked[r] = best;
  }
  for (int r = 0; r < beam; ++r) {
    for (int u = 0; u < T; ++u) {
      const long o = ((long)k * beam + r) * T + u;
      if (r < keep) {
        const long s = ((long)k * T * beam + picked[r]) * T + u;
        if (out_seq) out_seq[o] = done_seq[s];
        if (out_lps) out_lps[o] = done_lps[s];
        if (r == 0) { seq_out[(long)k * T + u] = done_seq[s]; logp_out[(long)k * T + u] = done_lps[s]; }
      } else {
        if (out_seq) out_seq[o] = 0;
        if (out_lps) out_lps[o] = 0.f;
      }
    }
    if (out_p) out_p[(long)k * beam + r] = r < keep ? p[picked[r]] : 0.f;
  }
  if (out_n) out_n[k] = keep;
}

static int beam_core(xg_context* ctx, const float* V, const float* fmask, const float* pos, int B, int K, int T, int beam,
                     int64_t* seq_out, float* logp_out, int64_t* done_seq, float* done_lps, float* done_p,
                     int32_t* done_count, BeamBufs& w, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, Vn = d.vocab;
  const int n = B * beam;
  // v2a(V) once per batch; init state per video, expanded to the beam rows (SAModel.py:141-147)
  {
    GemmP g = gemm_nt(V, H, P_(ctx, XG_P_V2A_W), H, w.Uv, d.att, B * K, d.att, H);
    g.ep.bias0 = P_(ctx, XG_P_V2A_B);
    XG_TRY(gemm_run(ctx, g, st));
  }
  XG_TRY(init_hidden_core(ctx, V, fmask, B, K, w.meanV, w.st0, H, st));
  for (int q = 0; q < 4; ++q) {
    XG_TRY(launch(ctx, "beam_expand", beam_expand_kernel, n, 128, 0, st, w.st0[q], beam, H, w.st[0][q]));
  }
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(w.tokens, 0, sizeof(int64_t) * (size_t)n, st));   // <bos> (:150)
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(w.sum, 0, sizeof(float) * (size_t)n, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(w.done_n, 0, sizeof(int) * (size_t)B, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(w.seq[0], 0, sizeof(int64_t) * (size_t)n * T, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(w.lps[0], 0, sizeof(float) * (size_t)n * T, st));
  int cur = 0;     // state buffer holding the current states
  int sb = 0;      // seq buffer holding the current beams
  // the word step runs as one persistent cooperative launch when the shape allows it (xg_persist.cuh): it also
  // gathers the parent states on the way in and leaves the per-row top-`beam` on the way out
  bool fused = beam <= XG_MAX_BEAM;      // persist_decode checks the shape (and refuses loudly on a strict handle)
  bool merged = false;                   // the merge of position t ran inside the previous word-step launch
  auto merge_io = [&](int t) {
    BeamMergeIO m;
    m.ys = w.ys; m.ix = w.ix; m.B = B; m.beam = beam; m.T = T; m.t = t;
    m.seq_in = w.seq[sb]; m.lps_in = w.lps[sb]; m.seq_out = w.seq[sb ^ 1]; m.lps_out = w.lps[sb ^ 1];
    m.sum = w.sum; m.parent = w.parent; m.tokens = w.tokens;
    m.done_seq = w.done_seq; m.done_lps = w.done_lps; m.done_p = w.done_p; m.done_n = w.done_n;
    return m;
  };
  if (fused && !env_flag("XG_BEAM_STEPWISE")) {
    // the whole search in ONE launch of the grouped step kernel when the shape fits it: the <bos> step and T - 1 more,
    // each followed by the candidate merge of its position (a cooperative launch per word step left the GPU idle for
    // ~30 us between steps: 14 % of a batch)
    PersistStepIO io;
    io.tokens = w.tokens; io.state = w.st[0]; io.logp = nullptr; io.feat_div = beam; io.first = 1;
    io.parent = nullptr; io.ys = w.ys; io.ix = w.ix; io.topk = beam;
    const BeamMergeIO mg = merge_io(0);
    const int pst = grouped_step(ctx, V, w.Uv, pos, n, K, io, st, &mg, T);
    if (pst == XG_OK) {
      XG_TRY(launch(ctx, "beam_finalize", beam_finalize_kernel, ceil_div(B, 64), 64, 0, st, w.done_seq, w.done_lps, w.done_p, w.done_n, B, beam, T, seq_out,
                                                          logp_out, done_seq, done_lps, done_p, done_count));
      return XG_OK;
    }
    if (pst != PK_FALLBACK) return pst;
  }
  for (int t = -1; t < T; ++t) {
    if (t >= 0) {
      if (!merged) {
        if (!fused) XG_TRY(launch(ctx, "beam_topk", beam_topk_kernel, n, 256, 0, st, w.logits, Vn, beam, w.ys, w.ix));
        XG_TRY(launch(ctx, "beam_merge", beam_merge_kernel, ceil_div(B, BEAM_MERGE_WARPS), 32 * BEAM_MERGE_WARPS, 0, st, merge_io(t)));
      }
      sb ^= 1;
      if (t == T - 1) break;   // the reference's last get_logprobs_state result is never used
      if (!fused) {
        XG_TRY(launch(ctx, "beam_gather_state", beam_gather_state_kernel, n, 128, 0, st, w.parent, H, w.st[cur][0], w.st[cur][1], w.st[cur][2], w.st[cur][3],
                                                   w.st[cur ^ 1][0], w.st[cur ^ 1][1], w.st[cur ^ 1][2], w.st[cur ^ 1][3]));
        cur ^= 1;
      }
    }
    merged = false;
    // one word step on all rows (t == -1: the <bos> step at SAModel.py:150-154)
    if (fused) {
      PersistStepIO io;
      io.tokens = w.tokens; io.state = w.st[cur]; io.logp = nullptr; io.feat_div = beam; io.first = (t == -1);
      io.parent = t >= 0 ? w.parent : nullptr; io.ys = w.ys; io.ix = w.ix; io.topk = beam;
      // grouped-cell form (xg_grouped.cuh) when the shape fits: the merge of position t + 1 runs at the end of its launch
      const BeamMergeIO mg = merge_io(t + 1);
      int pst = grouped_step(ctx, V, w.Uv, pos, n, K, io, st, &mg);
      if (pst == XG_OK) { merged = true; continue; }
      if (pst == PK_FALLBACK) pst = persist_decode(ctx, V, w.Uv, pos, nullptr, n, K, 1, nullptr, nullptr, nullptr, nullptr, st, &io);
      if (pst != PK_FALLBACK) { XG_TRY(pst); continue; }
      XG_REQUIRE(ctx->es, t == -1, XG_ERR_CUDA, "persistent word step became unavailable inside a beam search");
      fused = false;
    }
    XG_TRY(launch(ctx, "gather_rows", gather_rows_kernel, n, 128, 0, st, P_(ctx, XG_P_EMBED_W), w.tokens, 1, 0, n, n, E, Vn, w.step.XT));
    StepState s{w.st[cur][0], H, w.st[cur][1], w.st[cur][2], H, w.st[cur][3],
                w.st[cur][0], H, w.st[cur][1], w.st[cur][2], H, w.st[cur][3]};
    XG_TRY(decode_step_core(ctx, w.step.XT, nullptr, 0, V, w.Uv, pos, s, w.step, nullptr, n, K, beam, st));
    XG_TRY(logits_core(ctx, w.st[cur][2], H, n, w.logits, st));
    XG_TRY(launch(ctx, "logsoftmax_rows", logsoftmax_rows_kernel, n, 256, 0, st, w.logits, Vn, Vn, 0, 0, w.logits, Vn));
  }
  XG_TRY(launch(ctx, "beam_finalize", beam_finalize_kernel, ceil_div(B, 64), 64, 0, st, w.done_seq, w.done_lps, w.done_p, w.done_n, B, beam, T, seq_out,
                                                      logp_out, done_seq, done_lps, done_p, done_count));
  return XG_OK;
}

}  // namespace xg
