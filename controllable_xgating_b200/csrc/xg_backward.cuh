// Host orchestration of the hand-written backward pass (BPTT) of xg_train_fwd.
// Mirrors oracle/manual_bptt.py::backward line by line.
#pragma once
#include "xg_bwd_kernels.cuh"
#include "xg_forward.cuh"

namespace xg {

struct BwdBufs {
  float* DLOGITS;  // (L*B, V)
  float* dOUT;     // (L*B, H)
  float* dCL;      // (L*B, C)
  float* dHc;      // (L*B, Q)
  float* dHcar;    // (B, 2H)  carried d[h1|h2]
  float* dC1;      // (B, H)
  float* dC2;      // (B, H)
  float* dAF;      // (B, H)
  float* DAH;      // (L, B, A)
  float* dV;       // (B, K, H)
  float* dUv;      // (B, K, A)
  float* dwa_part; // (B, A)
  float* dba_part; // (B)
  float* dGP;      // (L*B, H) (reused as dRG)
  float* dXT;      // (L*B, E)
  float* dF;       // (K*B, H)
  float* dGG;      // (K*B, 2H)
  float* dH[2];    // (K*B, H)
  float* dR[2];    // (K*B, H)
  float* dhc;      // (B, H)   encoder carried dh
  float* dcc;      // (B, H)   encoder carried dc
  float* dE;       // (K*B, H)
  float* dy;       // (B*K, H)
  float* s_dy;     // (H)
  float* s_dyx;    // (H)
  double* part;
};

inline void carve_bwd(Arena& a, const xg_dims& d, int B, int K, int L, BwdBufs& w) {
  const long H = d.rnn, LB = (long)L * B, KB = (long)K * B;
  w.DLOGITS = a.take<float>(LB * d.vocab);
  w.dOUT = a.take<float>(LB * H);
  w.dCL = a.take<float>(LB * d.categories);
  w.dHc = a.take<float>(LB * d.cls_hidden);
  w.dHcar = a.take<float>((long)B * 2 * H);
  w.dC1 = a.take<float>((long)B * H);
  w.dC2 = a.take<float>((long)B * H);
  w.dAF = a.take<float>((long)B * H);
  w.DAH = a.take<float>(LB * d.att);
  w.dV = a.take<float>(KB * H);
  w.dUv = a.take<float>(KB * d.att);
  w.dwa_part = a.take<float>((long)B * d.att);
  w.dba_part = a.take<float>(B);
  w.dGP = a.take<float>(LB * H);
  w.dXT = a.take<float>(LB * d.embed);
  w.dF = a.take<float>(KB * H);
  w.dGG = a.take<float>(KB * 2 * H);
  for (int s = 0; s < 2; ++s) { w.dH[s] = a.take<float>(KB * H); w.dR[s] = a.take<float>(KB * H); }
  w.dhc = a.take<float>((long)B * H);
  w.dcc = a.take<float>((long)B * H);
  w.dE = a.take<float>(KB * H);
  w.dy = a.take<float>(KB * H);
  w.s_dy = a.take<float>(H);
  w.s_dyx = a.take<float>(H);
  w.part = a.take<double>((long)bn_row_splits((int)KB) * H * 2);
}

static int colsum_run(xg_context* ctx, const float* X, long ld, int R, int N, float beta, float* o0, float* o1,
                      float* o2, cudaStream_t st) {
  // row slabs so that about two blocks per SM run (scratch and ticket counters of the handle; one stream at a time)
  const int gx = ceil_div(N, 32);
  int S = 1;
  if (gx <= ctx->splitk.ctr_count && gx < 2 * ctx->sm_count) {
    S = std::min(std::min(ceil_div(2 * ctx->sm_count, gx), ceil_div(R, 32)), 64);
    if ((size_t)S * N > ctx->splitk.ws_floats) S = 1;
  }
  XG_TRY(launch(ctx, "colsum", colsum_kernel, dim3(gx, std::max(S, 1)), dim3(32, 8), 0, st, X, ld, R, N, beta, o0, o1, o2,
                ctx->splitk.ws, ctx->splitk.ctr));
  return XG_OK;
}

// The bias gradients of a backward pass are column sums of matrices that stay intact until the pass ends: they are
// collected and run as ONE launch per flush (decoder side / first encoder stream / second encoder stream) instead of 19.
struct ColsumScratch { float* part = nullptr; unsigned int* ctr = nullptr; };
constexpr size_t COLSUM_PART_FLOATS = (size_t)2 << 20;
constexpr int COLSUM_CTRS = 4096;
inline ColsumScratch*& colsum_scratch(xg_context* ctx) {
  static std::unordered_map<xg_context*, ColsumScratch*> m;
  return m[ctx];
}
static void colsum_release(xg_context* ctx) {
  ColsumScratch* s = colsum_scratch(ctx);
  if (!s) return;
  if (s->part) cudaFree(s->part);
  if (s->ctr) cudaFree(s->ctr);
  delete s;
  colsum_scratch(ctx) = nullptr;
}
struct ColsumBatch {
  ColsumJobs J;
  ColsumBatch() { J.n = 0; }
  int add(xg_context* ctx, const float* X, long ld, int R, int N, float beta, float* o0, float* o1, float* o2, cudaStream_t st) {
    if (J.n == COLSUM_MAX_JOBS) XG_TRY(flush(ctx, beta, st));
    const int q = J.n++;
    J.X[q] = X; J.ld[q] = ld; J.R[q] = R; J.N[q] = N; J.o[q][0] = o0; J.o[q][1] = o1; J.o[q][2] = o2;
    return XG_OK;
  }
  int flush(xg_context* ctx, float beta, cudaStream_t st) {
    if (J.n == 0) return XG_OK;
    ColsumScratch*& S = colsum_scratch(ctx);
    if (!S) {
      S = new ColsumScratch();
      XG_CUDA_TRY(ctx->es, cudaMalloc(&S->part, sizeof(float) * COLSUM_PART_FLOATS));
      XG_CUDA_TRY(ctx->es, cudaMalloc(&S->ctr, sizeof(unsigned int) * COLSUM_CTRS));
      XG_CUDA_TRY(ctx->es, cudaMemsetAsync(S->ctr, 0, sizeof(unsigned int) * COLSUM_CTRS, st));
    }
    // row slabs: every job is cut so that one block sums ~256 rows of its 32 columns (at least one slab)
    long blocks = 0, part = 0, ctrs = 0;
    for (int q = 0; q < J.n; ++q) {
      const int gx = ceil_div(J.N[q], 32);
      int s = std::max(1, std::min(64, ceil_div(J.R[q], 256)));
      if (part + (long)s * J.N[q] > (long)COLSUM_PART_FLOATS || ctrs + gx > COLSUM_CTRS) {      // scratch exhausted: run what is there
        const int keep = J.n;
        J.n = q;
        XG_TRY(launch_jobs(ctx, beta, (int)blocks, S, st));
        for (int i = q; i < keep; ++i) move(i, i - q);
        J.n = keep - q;
        return flush(ctx, beta, st);
      }
      J.S[q] = s; J.first_block[q] = (int)blocks; J.part_off[q] = (int)part; J.ctr_off[q] = (int)ctrs;
      blocks += (long)gx * s; part += (long)s * J.N[q]; ctrs += gx;
    }
    J.first_block[J.n] = (int)blocks;
    XG_TRY(launch_jobs(ctx, beta, (int)blocks, S, st));
    J.n = 0;
    return XG_OK;
  }
 private:
  void move(int from, int to) {
    J.X[to] = J.X[from]; J.ld[to] = J.ld[from]; J.R[to] = J.R[from]; J.N[to] = J.N[from];
    for (int k = 0; k < 3; ++k) J.o[to][k] = J.o[from][k];
  }
  int launch_jobs(xg_context* ctx, float beta, int blocks, ColsumScratch* S, cudaStream_t st) {
    if (J.n == 0 || blocks == 0) return XG_OK;
    J.beta = beta;
    J.first_block[J.n] = blocks;
    XG_TRY(launch(ctx, "colsum", colsum_multi_kernel, blocks, dim3(32, 8), 0, st, J, S->part, S->ctr));
    return XG_OK;
  }
};

// dW (+)= dy^T . x
static int wgrad(xg_context* ctx, const float* dy, long lddy, const float* x, long ldx, float* dW, int N, int Kin, int R,
                 float beta, cudaStream_t st) {
  GemmP g = gemm_tn(dy, lddy, x, ldx, dW, Kin, N, Kin, R);
  g.ep.beta = beta;
  return gemm_run(ctx, g, st);
}

static int train_bwd_core(xg_context* ctx, const float* rgb, const float* opfl, const float* fmask, const float* pos,
                          const int64_t* seq, const float* seq_mask, int B, int K, int L, int Lp, int train,
                          uint64_t seed, const float* logp, const float* cat, const float* dlogp, const float* dcat,
                          const TrainSaved& S, BwdBufs& W, float* const* G, float beta, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, A = d.att, V = d.vocab, C = d.categories, Q = d.cls_hidden;
  const int LB = Lp * B, KB = K * B;
  const float keep = (train && d.drop_prob > 0.f) ? 1.f / (1.f - d.drop_prob) : 1.f;
  const float* OUT = S.H12 + (long)B * 2 * H + H;   // (LB, H) ld 2H
  const long ldo = 2 * H;
  ColsumBatch cs;

  // ---------------- heads ----------------
  if (dlogp) {
    if (logsoftmax_reg_ok(logp, V, V, W.DLOGITS, V) && ((uintptr_t)dlogp & 15) == 0) {
      XG_TRY(launch(ctx, "logsoftmax_bwd_rows", logsoftmax_bwd_rows_reg_kernel, LB, 256, 0, st, logp, dlogp, B, Lp, V, W.DLOGITS));
    } else {
      XG_TRY(launch(ctx, "logsoftmax_bwd_rows", logsoftmax_bwd_rows_kernel, LB, 256, 0, st, logp, dlogp, B, Lp, V, W.DLOGITS));
    }
    XG_TRY(wgrad(ctx, W.DLOGITS, V, OUT, ldo, G[XG_P_LOGIT_W], V, H, LB, beta, st));
    XG_TRY(cs.add(ctx, W.DLOGITS, V, LB, V, beta, G[XG_P_LOGIT_B], nullptr, nullptr, st));
    GemmP g = gemm_nn(W.DLOGITS, V, P_(ctx, XG_P_LOGIT_W), H, W.dOUT, H, LB, H, V);
    XG_TRY(gemm_run(ctx, g, st));
  } else {
    XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dOUT, 0, sizeof(float) * (size_t)LB * H, st));
    if (beta == 0.f) {
      XG_CUDA_TRY(ctx->es, cudaMemsetAsync(G[XG_P_LOGIT_W], 0, sizeof(float) * (size_t)V * H, st));
      XG_CUDA_TRY(ctx->es, cudaMemsetAsync(G[XG_P_LOGIT_B], 0, sizeof(float) * (size_t)V, st));
    }
  }
  if (dcat) {
    XG_TRY(launch(ctx, "logsoftmax_bwd_rows", logsoftmax_bwd_rows_kernel, LB, 128, 0, st, cat, dcat, B, Lp, C, W.dCL));
    XG_TRY(wgrad(ctx, W.dCL, C, S.Hc, Q, G[XG_P_CLS3_W], C, Q, LB, beta, st));
    XG_TRY(cs.add(ctx, W.dCL, C, LB, C, beta, G[XG_P_CLS3_B], nullptr, nullptr, st));
    GemmP g = gemm_nn(W.dCL, C, P_(ctx, XG_P_CLS3_W), Q, W.dHc, Q, LB, Q, C);
    XG_TRY(gemm_run(ctx, g, st));
    XG_TRY(launch(ctx, "relu_drop_bwd", relu_drop_bwd_kernel, ew_grid((long)LB * Q), 256, 0, st, W.dHc, S.Hc, (long)LB * Q, keep));
    XG_TRY(wgrad(ctx, W.dHc, Q, OUT, ldo, G[XG_P_CLS0_W], Q, H, LB, beta, st));
    XG_TRY(cs.add(ctx, W.dHc, Q, LB, Q, beta, G[XG_P_CLS0_B], nullptr, nullptr, st));
    GemmP g2 = gemm_nn(W.dHc, Q, P_(ctx, XG_P_CLS0_W), H, W.dOUT, H, LB, H, Q);
    g2.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, g2, st));
  } else if (beta == 0.f) {
    XG_CUDA_TRY(ctx->es, cudaMemsetAsync(G[XG_P_CLS3_W], 0, sizeof(float) * (size_t)C * Q, st));
    XG_CUDA_TRY(ctx->es, cudaMemsetAsync(G[XG_P_CLS3_B], 0, sizeof(float) * (size_t)C, st));
    XG_CUDA_TRY(ctx->es, cudaMemsetAsync(G[XG_P_CLS0_W], 0, sizeof(float) * (size_t)Q * H, st));
    XG_CUDA_TRY(ctx->es, cudaMemsetAsync(G[XG_P_CLS0_B], 0, sizeof(float) * (size_t)Q, st));
  }

  // ---------------- decoder BPTT ----------------
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dHcar, 0, sizeof(float) * (size_t)B * 2 * H, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dC1, 0, sizeof(float) * (size_t)B * H, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dC2, 0, sizeof(float) * (size_t)B * H, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dV, 0, sizeof(float) * (size_t)KB * H, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dUv, 0, sizeof(float) * (size_t)KB * A, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dwa_part, 0, sizeof(float) * (size_t)B * A, st));
  XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dba_part, 0, sizeof(float) * (size_t)B, st));
  float* G1 = S.G1;   // gates are overwritten by dz in place (the saved block is consumed by backward)
  float* G2 = S.G2;
  const size_t att_smem = (size_t)(A + 2 * K + H) * sizeof(float);
  // the word loop: one persistent cooperative kernel for all L' steps when the shape allows it (xg_persist.cuh)
  int pdb = PK_FALLBACK;
  {
    PersistBwdIO io;
    io.seq_mask = seq_mask; io.L = L; io.dOUT = W.dOUT;
    io.G1 = G1; io.G2 = G2; io.C1 = S.C1; io.C2 = S.C2; io.AH = S.AH; io.ALPHA = S.ALPHA;
    io.DAH = W.DAH; io.dV = W.dV; io.dUv = W.dUv; io.dwa_part = W.dwa_part; io.dba_part = W.dba_part;
    io.dHcar = W.dHcar; io.dC1 = W.dC1; io.dC2 = W.dC2;
    io.drop1 = make_drop(train, d.drop_prob, seed, XG_DROP_DEC_H1);
    io.drop2 = make_drop(train, d.drop_prob, seed, XG_DROP_DEC_H2);
    pdb = persist_decode_bwd(ctx, S.V, S.Uv, B, K, Lp, io, st);
    if (pdb != PK_FALLBACK) XG_TRY(pdb);
  }
  for (int i = Lp - 1; i >= 0 && pdb == PK_FALLBACK; --i) {
    const float* m = seq_mask + i;
    float* dz2 = G2 + (long)i * B * 4 * H;
    float* dz1 = G1 + (long)i * B * 4 * H;
    // (a) lstm_2 cell backward: dh_out = dOUT[i] + carried dh2
    XG_TRY(launch(ctx, "dec_cell_bwd", dec_cell_bwd_kernel, ceil_div(B * H, 256), 256, 0, st, 
        dz2, S.C2 + (long)(i + 1) * B * H, S.C2 + (long)i * B * H, W.dOUT + (long)i * B * H, H, W.dHcar + H, 2 * H, W.dC2,
        m, L, B, H, make_drop(train, d.drop_prob, seed, XG_DROP_DEC_H2, (uint64_t)i * B * H), W.dHcar + H, 2 * H));
    // (b) dh1_new += dz2 . W_i2h ; (c) dAF = dz2 . W_a2h ; (d) dh2_prev += dz2 . W_h2h
    GemmP gb = gemm_nn(dz2, 4 * H, P_(ctx, XG_P_L2_I2H_W), H, W.dHcar, 2 * H, B, H, 4 * H);
    gb.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gb, st));
    GemmP gc = gemm_nn(dz2, 4 * H, P_(ctx, XG_P_L2_A2H_W), H, W.dAF, H, B, H, 4 * H);
    XG_TRY(gemm_run(ctx, gc, st));
    GemmP gd = gemm_nn(dz2, 4 * H, P_(ctx, XG_P_L2_H2H_W), H, W.dHcar + H, 2 * H, B, H, 4 * H);
    gd.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gd, st));
    // (e) attention backward
    XG_TRY(launch(ctx, "att_bwd", att_bwd_kernel, dim3(B, ATT_BWD_SPLITS), 256, att_smem, st, W.dAF, S.AH + (long)i * B * A, S.Uv, S.V, P_(ctx, XG_P_A2W_W),
                                            S.ALPHA + (long)i * B * K, K, A, H, W.dV, W.dUv,
                                            W.DAH + (long)i * B * A, W.dwa_part, W.dba_part));
    // (f) lstm_1 cell backward on dh1_new
    XG_TRY(launch(ctx, "dec_cell_bwd", dec_cell_bwd_kernel, ceil_div(B * H, 256), 256, 0, st, 
        dz1, S.C1 + (long)(i + 1) * B * H, S.C1 + (long)i * B * H, W.dHcar, 2 * H, nullptr, 0, W.dC1, m, L, B, H,
        make_drop(train, d.drop_prob, seed, XG_DROP_DEC_H1, (uint64_t)i * B * H), W.dHcar, 2 * H));
    // (g) dh1_prev += dz1 . W_h2h ; (h) d[h1|h2]_prev += dAH . W_h2a
    GemmP gg = gemm_nn(dz1, 4 * H, P_(ctx, XG_P_L1_H2H_W), H, W.dHcar, 2 * H, B, H, 4 * H);
    gg.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gg, st));
    GemmP gh = gemm_nn(W.DAH + (long)i * B * A, A, P_(ctx, XG_P_H2A_W), 2 * H, W.dHcar, 2 * H, B, 2 * H, A);
    gh.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gh, st));
  }
  // init-state linears (the mean is detached from the encoder: SAModel.py:59-62)
  {
    const int iw[4] = {XG_P_INIT_H1_W, XG_P_INIT_C1_W, XG_P_INIT_H2_W, XG_P_INIT_C2_W};
    const float* dsrc[4] = {W.dHcar, W.dC1, W.dHcar + H, W.dC2};
    const long ldd[4] = {2 * H, H, 2 * H, H};
    for (int q = 0; q < 4; ++q) {
      XG_TRY(wgrad(ctx, dsrc[q], ldd[q], S.enc.meanV, H, G[iw[q]], H, H, B, beta, st));
      XG_TRY(cs.add(ctx, dsrc[q], ldd[q], B, H, beta, G[iw[q] + 1], nullptr, nullptr, st));
    }
  }
  // batched weight gradients over all steps
  {
    // Nothing this block reads as a GEMM operand is written after its first use in it (W.dGP and W.dXT are outputs before
    // they are operands): dz2 / dz1 feed three weight gradients each in the same layout, dz1 two input gradients, the
    // embeddings two weight gradients -> one operand split each.
    TcHold hold(ctx);
    const float* Hprev = S.H12;                       // rows (i,b), i in [0,Lp)
    const float* Hnew = S.H12 + (long)B * 2 * H;      // i in [1,Lp]
    XG_TRY(wgrad(ctx, G2, 4 * H, Hnew, 2 * H, G[XG_P_L2_I2H_W], 4 * H, H, LB, beta, st));
    XG_TRY(wgrad(ctx, G2, 4 * H, S.AF, H, G[XG_P_L2_A2H_W], 4 * H, H, LB, beta, st));
    XG_TRY(wgrad(ctx, G2, 4 * H, Hprev + H, 2 * H, G[XG_P_L2_H2H_W], 4 * H, H, LB, beta, st));
    XG_TRY(cs.add(ctx, G2, 4 * H, LB, 4 * H, beta, G[XG_P_L2_I2H_B], G[XG_P_L2_A2H_B], G[XG_P_L2_H2H_B], st));
    XG_TRY(wgrad(ctx, G1, 4 * H, S.XT, E, G[XG_P_L1_I2H_W], 4 * H, E, LB, beta, st));
    XG_TRY(wgrad(ctx, G1, 4 * H, S.GP, H, G[XG_P_L1_A2H_W], 4 * H, H, LB, beta, st));
    XG_TRY(wgrad(ctx, G1, 4 * H, Hprev, 2 * H, G[XG_P_L1_H2H_W], 4 * H, H, LB, beta, st));
    XG_TRY(cs.add(ctx, G1, 4 * H, LB, 4 * H, beta, G[XG_P_L1_I2H_B], G[XG_P_L1_A2H_B], G[XG_P_L1_H2H_B], st));
    XG_TRY(wgrad(ctx, W.DAH, A, Hprev, 2 * H, G[XG_P_H2A_W], A, 2 * H, LB, beta, st));
    XG_TRY(cs.add(ctx, W.DAH, A, LB, A, beta, G[XG_P_H2A_B], nullptr, nullptr, st));
    XG_TRY(cs.add(ctx, W.dwa_part, A, B, A, beta, G[XG_P_A2W_W], nullptr, nullptr, st));
    XG_TRY(cs.add(ctx, W.dba_part, 1, B, 1, beta, G[XG_P_A2W_B], nullptr, nullptr, st));
    // POS gate + embedding
    GemmP gp = gemm_nn(G1, 4 * H, P_(ctx, XG_P_L1_A2H_W), H, W.dGP, H, LB, H, 4 * H);
    XG_TRY(gemm_run(ctx, gp, st));
    XG_TRY(launch(ctx, "dgate_bwd", dgate_bwd_kernel, ew_grid((long)LB * H), 256, 0, st, W.dGP, S.RG, pos, B, LB, H, keep, W.dGP));
    XG_TRY(wgrad(ctx, W.dGP, H, S.XT, E, G[XG_P_DGATE_W], H, E, LB, beta, st));
    XG_TRY(cs.add(ctx, W.dGP, H, LB, H, beta, G[XG_P_DGATE_B], nullptr, nullptr, st));
    GemmP gx = gemm_nn(G1, 4 * H, P_(ctx, XG_P_L1_I2H_W), E, W.dXT, E, LB, E, 4 * H);
    XG_TRY(gemm_run(ctx, gx, st));
    GemmP gx2 = gemm_nn(W.dGP, H, P_(ctx, XG_P_DGATE_W), E, W.dXT, E, LB, E, H);
    gx2.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gx2, st));
    if (beta == 0.f) XG_CUDA_TRY(ctx->es, cudaMemsetAsync(G[XG_P_EMBED_W], 0, sizeof(float) * (size_t)V * E, st));
    XG_TRY(launch(ctx, "embed_scatter_add", embed_scatter_add_kernel, LB, 128, 0, st, W.dXT, seq, B, L, LB, E, V, G[XG_P_EMBED_W]));
    // v2a
    XG_TRY(wgrad(ctx, W.dUv, A, S.V, H, G[XG_P_V2A_W], A, H, KB, beta, st));
    XG_TRY(cs.add(ctx, W.dUv, A, KB, A, beta, G[XG_P_V2A_B], nullptr, nullptr, st));
    GemmP gv = gemm_nn(W.dUv, A, P_(ctx, XG_P_V2A_W), H, W.dV, H, KB, H, A);
    gv.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gv, st));
  }

  // Every gradient of the decoder side (init-state linears, POS gate, both LSTM cells, attention, embedding, logit and
  // classifier heads: parameters XG_P_INIT_H1_W .. XG_P_CLS3_B, the tail of the flat gradient buffer) is final here; what
  // follows only writes the encoder's.  Data-parallel callers start the all-reduce of that tail on another stream now.
  // The event itself is recorded further down, behind the cooperative frame-recurrence kernel: an NCCL kernel and a
  // cooperative launch that wants every SM exclude each other, the batched GEMMs after it do not.
  XG_TRY(cs.flush(ctx, beta, st));
  // ---------------- encoder backward (rows (k,b)) ----------------
  const EncBufs& eb = S.enc;
  XG_TRY(launch(ctx, "fusion_bwd", fusion_bwd_kernel, ew_grid((long)KB * H), 256, 0, st, W.dV, S.V, B, K, H, d.fusion_act,
                                                          make_drop(train, d.drop_prob, seed, XG_DROP_ENC_FUSION), W.dF));
  XG_TRY(wgrad(ctx, W.dF, H, eb.GG, 2 * H, G[XG_P_FUSION_W], H, 2 * H, KB, beta, st));
  XG_TRY(cs.add(ctx, W.dF, H, KB, H, beta, G[XG_P_FUSION_B], nullptr, nullptr, st));
  {
    GemmP g = gemm_nn(W.dF, H, P_(ctx, XG_P_FUSION_W), 2 * H, W.dGG, 2 * H, KB, 2 * H, H);
    XG_TRY(gemm_run(ctx, g, st));
  }
  const int gw[2] = {XG_P_GATE_RGB_W, XG_P_GATE_OPFL_W};
  for (int s = 0; s < 2; ++s) {
    XG_TRY(launch(ctx, "gate_bwd", gate_bwd_kernel, ew_grid((long)KB * H), 256, 0, st, W.dGG + (long)s * H, 2 * H, eb.R[s], eb.Hs[s], KB, H, keep,
                                                          W.dH[s], W.dR[s]));
  }
  for (int s = 0; s < 2; ++s) {
    const int src = 1 - s;
    XG_TRY(wgrad(ctx, W.dR[s], H, eb.Hs[src], H, G[gw[s]], H, H, KB, beta, st));
    XG_TRY(cs.add(ctx, W.dR[s], H, KB, H, beta, G[gw[s] + 1], nullptr, nullptr, st));
    GemmP g = gemm_nn(W.dR[s], H, P_(ctx, gw[s]), H, W.dH[src], H, KB, H, H);
    g.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, g, st));
  }
  const float* X[2] = {rgb, opfl};
  const int din[2] = {d.feat_rgb, d.feat_opfl};
  const int pw[2] = {XG_P_EMB_RGB_W, XG_P_EMB_OPFL_W};
  const int plstm[2] = {XG_P_LSTM_RGB_WIH, XG_P_LSTM_OPFL_WIH};
  const uint32_t site_emb[2] = {XG_DROP_ENC_EMB_RGB, XG_DROP_ENC_EMB_OPFL};
  const int RS = bn_row_splits(KB);
  // frame recurrence of both streams: one persistent cooperative kernel when the shape allows it (xg_persist.cuh)
  const int pbw = persist_encode_bwd(ctx, fmask, B, K, eb, W.dH, st);
  if (pbw != PK_FALLBACK) XG_TRY(pbw);
  if (ctx->bwd_split_event) XG_CUDA_TRY(ctx->es, cudaEventRecord(ctx->bwd_split_event, st));      // decoder-side gradients final (see above)
  for (int s = 0; s < 2; ++s) {
    float* DZ = eb.G[s];   // gates -> dz in place
    if (pbw == PK_FALLBACK) {
      XG_CUDA_TRY(ctx->es, cudaMemsetAsync(W.dcc, 0, sizeof(float) * (size_t)B * H, st));
      for (int t = K - 1; t >= 0; --t) {
        float* dzt = DZ + (long)t * B * 4 * H;
        XG_TRY(launch(ctx, "enc_cell_bwd", enc_cell_bwd_kernel, ceil_div(B * H, 256), 256, 0, st,
            dzt, eb.Cs[s] + (long)t * B * H, t > 0 ? eb.Cs[s] + (long)(t - 1) * B * H : nullptr,
            W.dH[s] + (long)t * B * H, t < K - 1 ? W.dhc : nullptr, W.dcc, fmask, K, t, B, H));
        if (t > 0) {
          GemmP g = gemm_nn(dzt, 4 * H, P_(ctx, plstm[s] + 1), H, W.dhc, H, B, H, 4 * H);
          XG_TRY(gemm_run(ctx, g, st));
        }
      }
    }
    // weight_hh: rows t>=1 of dz pair with h[t-1]
    if (K > 1) {
      XG_TRY(wgrad(ctx, DZ + (long)B * 4 * H, 4 * H, eb.Hs[s], H, G[plstm[s] + 1], 4 * H, H, (K - 1) * B, beta, st));
    } else if (beta == 0.f) {
      XG_CUDA_TRY(ctx->es, cudaMemsetAsync(G[plstm[s] + 1], 0, sizeof(float) * (size_t)4 * H * H, st));
    }
    XG_TRY(wgrad(ctx, DZ, 4 * H, eb.E[s], H, G[plstm[s]], 4 * H, H, KB, beta, st));
    XG_TRY(cs.add(ctx, DZ, 4 * H, KB, 4 * H, beta, G[plstm[s] + 2], G[plstm[s] + 3], nullptr, st));
    {
      GemmP g = gemm_nn(DZ, 4 * H, P_(ctx, plstm[s]), H, W.dE, H, KB, H, 4 * H);
      XG_TRY(gemm_run(ctx, g, st));
    }
    // through mask, dropout, ReLU and BatchNorm
    XG_TRY(launch(ctx, "bn_bwd_prep", bn_bwd_prep_kernel, ew_grid((long)KB * H), 256, 0, st, W.dE, eb.Y[s], eb.scale[s], eb.shift[s], fmask, B, K, H,
                                                             make_drop(train, d.drop_prob, seed, site_emb[s]), W.dy));
    XG_TRY(launch(ctx, "bn_bwd_stats", bn_bwd_stats_kernel, dim3(ceil_div(H, 32), RS), dim3(32, 8), 0, st, W.dy, eb.Y[s], eb.mean[s], eb.invstd[s], KB, H,
                                                                          W.part));
    XG_TRY(launch(ctx, "bn_bwd_finalize", bn_bwd_finalize_kernel, ceil_div(H, 128), 128, 0, st, W.part, RS, H, beta, G[pw[s] + 2], G[pw[s] + 3], W.s_dy,
                                                            W.s_dyx));
    XG_TRY(launch(ctx, "bn_bwd_apply", bn_bwd_apply_kernel, ew_grid((long)KB * H), 256, 0, st, W.dy, eb.Y[s], eb.mean[s], eb.invstd[s],
                                                              P_(ctx, pw[s] + 2), W.s_dy, W.s_dyx, KB, H, train));
    XG_TRY(wgrad(ctx, W.dy, H, X[s], din[s], G[pw[s]], H, din[s], KB, beta, st));
    XG_TRY(cs.add(ctx, W.dy, H, KB, H, beta, G[pw[s] + 1], nullptr, nullptr, st));
    XG_TRY(cs.flush(ctx, beta, st));      // (W.dy is reused by the second stream)
  }
  return XG_OK;
}

}  // namespace xg
