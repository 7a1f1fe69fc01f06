// Host orchestration of the forward path: encoder (Cross-Gating block), init state, one word step,
// greedy decode loop, teacher-forced training forward.
#pragma once
#include "xg_context.cuh"
#include "xg_fwd_kernels.cuh"
#include "xg_gemm_tc.cuh"
#include "xg_persist.cuh"
#include "xg_grouped.cuh"

namespace xg {

#define P_(ctx, id) ((ctx)->P[id])

static inline int ew_grid(long n, int block = 256) {
  long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > 148L * 16) g = 148L * 16;
  return (int)g;
}

static int gemm_run(xg_context* ctx, const GemmP& p, cudaStream_t st) {
  if (ctx->tc_mode && tc_eligible(p)) return gemm_tc(ctx, p, st);
  if (ctx->prof_on) {
    char tag[96];
    const char* lay = (p.sa_r == 1) ? (p.sb_r == 1 ? "nt" : "nn") : "tn";
    snprintf(tag, sizeof(tag), "gemm_%s_%dx%dx%d", lay, p.M, p.N, p.K);
    ProfScope ps(ctx, std::string(tag), st);
    return gemm_simt(ctx->es, p, st, &ctx->splitk);
  }
  return gemm_simt(ctx->es, p, st, &ctx->splitk);
}

// ------------------------------------------------------------------------------------
// EncoderLstm_two_fc.forward (sub_modules.py:118-159)
// ------------------------------------------------------------------------------------
static int init_hidden_core(xg_context* ctx, const float* V, const float* fmask, int B, int K, float* meanV,
                            float* const* state_out, long ld_out, cudaStream_t st);

static int encode_core(xg_context* ctx, const float* rgb, const float* opfl, const float* fmask, int B, int K,
                       int train_flags, uint64_t seed, EncBufs& eb, float* V_out, float* Uv_out,
                       float* const* state_out, cudaStream_t st) {
  // train_flags: bit 0 = training mode (batch statistics, dropout); bit 1 = do NOT update the BatchNorm running
  // statistics (second pass over the same batch: the teacher-forced pass of the self-critical path)
  const int train = train_flags & 1, bn_update = (train_flags & 2) ? 0 : 1;
  const xg_dims& d = ctx->d;
  const int H = d.rnn;
  const int BK = B * K;
  const float* X[2] = {rgb, opfl};
  const int din[2] = {d.feat_rgb, d.feat_opfl};
  const int pw[2] = {XG_P_EMB_RGB_W, XG_P_EMB_OPFL_W};
  const int plstm[2] = {XG_P_LSTM_RGB_WIH, XG_P_LSTM_OPFL_WIH};
  const uint32_t site_emb[2] = {XG_DROP_ENC_EMB_RGB, XG_DROP_ENC_EMB_OPFL};
  const int RS = bn_row_splits(BK);

  for (int s = 0; s < 2; ++s) {
    // visual_emb_*.0 : Linear over the flattened (B*K) rows, padded rows included (:121,126)
    GemmP g = gemm_nt(X[s], din[s], P_(ctx, pw[s]), din[s], eb.Y[s], H, BK, H, din[s]);
    g.ep.bias0 = P_(ctx, pw[s] + 1);
    XG_TRY(gemm_run(ctx, g, st));
    // visual_emb_*.1 : BatchNorm1d
    if (train) {
      XG_TRY(launch(ctx, "colstats_partial", colstats_partial_kernel, dim3(ceil_div(H, 32), RS), dim3(32, 8), 0, st, eb.Y[s], BK, H, eb.part));
    }
    XG_TRY(launch(ctx, "bn_finalize", bn_finalize_kernel, ceil_div(H, 128), 128, 0, st, eb.part, RS, BK, H, train, P_(ctx, pw[s] + 2),
                                                        P_(ctx, pw[s] + 3), ctx->bn[2 * s], ctx->bn[2 * s + 1],
                                                        d.bn_eps, d.bn_momentum, (train && bn_update) ? 1 : 0, eb.scale[s],
                                                        eb.shift[s], eb.mean[s], eb.invstd[s]));
    // ReLU, dropout, frame mask; output frame-major (:123,128)
    XG_TRY(launch(ctx, "bn_apply", bn_apply_kernel, ew_grid((long)BK * H), 256, 0, st, eb.Y[s], eb.scale[s], eb.shift[s], fmask, B, K, H,
                                                          make_drop(train, d.drop_prob, seed, site_emb[s]), eb.E[s]));
    // nn.LSTMCell input projection hoisted over all frames: XG = E . W_ih^T + b_ih + b_hh
    GemmP gi = gemm_nt(eb.E[s], H, P_(ctx, plstm[s]), H, eb.G[s], 4 * H, BK, 4 * H, H);
    gi.ep.bias0 = P_(ctx, plstm[s] + 2);
    gi.ep.bias1 = P_(ctx, plstm[s] + 3);
    XG_TRY(gemm_run(ctx, gi, st));
  }
  // recurrence over frames (:132-147); the two streams are independent: one persistent cooperative
  // kernel for all frames of both streams when the shape allows it (xg_persist.cuh)
  int pst = grouped_encode(ctx, fmask, B, K, eb, st);      // stationary-weight grouped kernel (xg_grouped.cuh) when the shape fits
  if (pst == PK_FALLBACK) pst = persist_encode(ctx, fmask, B, K, eb, st);
  if (pst != PK_FALLBACK) XG_TRY(pst);
  for (int t = 0; t < K && pst == PK_FALLBACK; ++t) {
    for (int s = 0; s < 2; ++s) {
      float* Zt = eb.G[s] + (long)t * B * 4 * H;
      if (t > 0) {
        GemmP gh = gemm_nt(eb.Hs[s] + (long)(t - 1) * B * H, H, P_(ctx, plstm[s] + 1), H, Zt, 4 * H, B, 4 * H, H);
        gh.ep.beta = 1.f;
        XG_TRY(gemm_run(ctx, gh, st));
      }
      XG_TRY(launch(ctx, "enc_cell", enc_cell_kernel, ceil_div(B * H, 256), 256, 0, st, Zt, t > 0 ? eb.Cs[s] + (long)(t - 1) * B * H : nullptr,
                                                           fmask, K, t, B, H, eb.Cs[s] + (long)t * B * H,
                                                           eb.Hs[s] + (long)t * B * H));
    }
  }
  // cross gates batched over all frames (:151-152): g_tgt = h_tgt * (1 + dropout(relu(Linear(h_src))))
  {
    const int gw[2] = {XG_P_GATE_RGB_W, XG_P_GATE_OPFL_W};
    const uint32_t site_gate[2] = {XG_DROP_ENC_GATE_RGB, XG_DROP_ENC_GATE_OPFL};
    for (int s = 0; s < 2; ++s) {
      const int src = 1 - s;
      GemmP g = gemm_nt(eb.Hs[src], H, P_(ctx, gw[s]), H, eb.GG + (long)s * H, 2 * H, BK, H, H);
      g.ep.bias0 = P_(ctx, gw[s] + 1);
      g.ep.act = XG_ACT_RELU;
      g.ep.drop = make_drop(train, d.drop_prob, seed, site_gate[s]);
      g.ep.aux = eb.R[s];
      g.ep.ldaux = H;
      g.ep.tgt = eb.Hs[s];
      g.ep.ldt = H;
      XG_TRY(gemm_run(ctx, g, st));
    }
  }
  // late fusion (:158), rows back to batch-major
  {
    GemmP g = gemm_nt(eb.GG, 2 * H, P_(ctx, XG_P_FUSION_W), 2 * H, V_out, H, BK, H, 2 * H);
    g.ep.bias0 = P_(ctx, XG_P_FUSION_B);
    g.ep.act = d.fusion_act;
    g.ep.drop = make_drop(train, d.drop_prob, seed, XG_DROP_ENC_FUSION);
    g.perm_rb = B;
    g.perm_rs = K;
    XG_TRY(gemm_run(ctx, g, st));
  }
  if (Uv_out) {  // v2a(V): loop invariant of the word loop (sub_modules.py:677)
    GemmP g = gemm_nt(V_out, H, P_(ctx, XG_P_V2A_W), H, Uv_out, d.att, BK, d.att, H);
    g.ep.bias0 = P_(ctx, XG_P_V2A_B);
    XG_TRY(gemm_run(ctx, g, st));
  }
  if (state_out) XG_TRY(init_hidden_core(ctx, V_out, fmask, B, K, eb.meanV, state_out, H, st));      // SAModel.init_hidden (SAModel.py:58-65)
  return XG_OK;
}

static int init_hidden_core(xg_context* ctx, const float* V, const float* fmask, int B, int K, float* meanV,
                            float* const* state_out, long ld_out, cudaStream_t st) {
  const int H = ctx->d.rnn;
  XG_TRY(launch(ctx, "masked_mean", masked_mean_kernel, dim3(B, ceil_div(H, 128)), 128, 0, st, V, fmask, B, K, H, meanV));
  const int iw[4] = {XG_P_INIT_H1_W, XG_P_INIT_C1_W, XG_P_INIT_H2_W, XG_P_INIT_C2_W};
  if (H <= 32 * IS_KPL && ctx->tc_mode != 0) {           // the four linears in one launch (engine mode 0 keeps the plain GEMM path)
    InitStateArgs a;
    for (int q = 0; q < 4; ++q) { a.W[q] = P_(ctx, iw[q]); a.bias[q] = P_(ctx, iw[q] + 1); a.out[q] = state_out[q]; a.ld[q] = q % 2 == 0 ? ld_out : H; }
    const size_t smem = sizeof(float) * (size_t)IS_CAPS * H;
    static bool attr_set = false;
    if (!attr_set) { XG_CUDA_TRY(ctx->es, cudaFuncSetAttribute(init_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * IS_CAPS * 32 * IS_KPL))); attr_set = true; }
    XG_TRY(launch(ctx, "init_state", init_state_kernel, dim3(ceil_div(H, 16), 4), 256, smem, st, (const float*)meanV, B, H, a));
    return XG_OK;
  }
  for (int q = 0; q < 4; ++q) {
    GemmP g = gemm_nt(meanV, H, P_(ctx, iw[q]), H, state_out[q], q % 2 == 0 ? ld_out : H, B, H, H);
    g.ep.bias0 = P_(ctx, iw[q] + 1);
    XG_TRY(gemm_run(ctx, g, st));
  }
  return XG_OK;
}

// ------------------------------------------------------------------------------------
// LSTMCore_two_layer_gate.forward (sub_modules.py:671-687), inference form (nothing hoisted)
// ------------------------------------------------------------------------------------
struct StepState {
  const float* h1p; long ld_h1p; const float* c1p; const float* h2p; long ld_h2p; const float* c2p;
  float* h1n; long ld_h1n; float* c1n; float* h2n; long ld_h2n; float* c2n;
};

static int decode_step_core(xg_context* ctx, const float* xt, const float* mask, long mask_stride, const float* V,
                            const float* Uv, const float* pos, const StepState& s, StepBufs& sb, float* alpha_out,
                            int B, int K, int feat_div, cudaStream_t st, const DropSpec* drops = nullptr) {
  // drops (optional): training dropout of the step {POS gate relu, lstm_1 h, lstm_2 h} with per-step index bases
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, A = d.att;
  // attention on the PREVIOUS states
  {
    GemmP g = gemm_nt(s.h1p, s.ld_h1p, P_(ctx, XG_P_H2A_W), 2 * H, sb.AH, A, B, A, H);
    g.ep.bias0 = P_(ctx, XG_P_H2A_B);
    XG_TRY(gemm_run(ctx, g, st));
    GemmP g2 = gemm_nt(s.h2p, s.ld_h2p, P_(ctx, XG_P_H2A_W) + H, 2 * H, sb.AH, A, B, A, H);
    g2.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, g2, st));
    XG_TRY(launch(ctx, "att_fwd", att_fwd_kernel, B, 256, (A + K) * sizeof(float), st, sb.AH, Uv, V, P_(ctx, XG_P_A2W_W), P_(ctx, XG_P_A2W_B), K, A,
                                                           H, feat_div, alpha_out, sb.AF));
  }
  // POS gate: gp = pos * (1 + relu(Linear(xt)))
  {
    GemmP g = gemm_nt(xt, E, P_(ctx, XG_P_DGATE_W), E, sb.GP, H, B, H, E);
    g.ep.bias0 = P_(ctx, XG_P_DGATE_B);
    g.ep.act = XG_ACT_RELU;
    g.ep.tgt = pos;
    g.ep.ldt = H;
    g.ep.tgt_div = feat_div > 1 ? feat_div : 0;
    if (drops) g.ep.drop = drops[0];
    XG_TRY(gemm_run(ctx, g, st));
  }
  // lstm_1
  {
    GemmP g = gemm_nt(xt, E, P_(ctx, XG_P_L1_I2H_W), E, sb.Z1, 4 * H, B, 4 * H, E);
    g.ep.bias0 = P_(ctx, XG_P_L1_I2H_B); g.ep.bias1 = P_(ctx, XG_P_L1_A2H_B); g.ep.bias2 = P_(ctx, XG_P_L1_H2H_B);
    XG_TRY(gemm_run(ctx, g, st));
    GemmP ga = gemm_nt(sb.GP, H, P_(ctx, XG_P_L1_A2H_W), H, sb.Z1, 4 * H, B, 4 * H, H);
    ga.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, ga, st));
    GemmP gh = gemm_nt(s.h1p, s.ld_h1p, P_(ctx, XG_P_L1_H2H_W), H, sb.Z1, 4 * H, B, 4 * H, H);
    gh.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gh, st));
    XG_TRY(launch(ctx, "dec_cell", dec_cell_kernel, ceil_div(B * H, 256), 256, 0, st, sb.Z1, s.c1p, s.h1p, s.ld_h1p, mask, mask_stride, B, H,
                                                         drops ? drops[1] : make_drop(false, 0.f, 0, 0), s.c1n, s.h1n, s.ld_h1n, nullptr, 0));
  }
  // lstm_2
  {
    GemmP g = gemm_nt(s.h1n, s.ld_h1n, P_(ctx, XG_P_L2_I2H_W), H, sb.Z2, 4 * H, B, 4 * H, H);
    g.ep.bias0 = P_(ctx, XG_P_L2_I2H_B); g.ep.bias1 = P_(ctx, XG_P_L2_A2H_B); g.ep.bias2 = P_(ctx, XG_P_L2_H2H_B);
    XG_TRY(gemm_run(ctx, g, st));
    GemmP ga = gemm_nt(sb.AF, H, P_(ctx, XG_P_L2_A2H_W), H, sb.Z2, 4 * H, B, 4 * H, H);
    ga.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, ga, st));
    GemmP gh = gemm_nt(s.h2p, s.ld_h2p, P_(ctx, XG_P_L2_H2H_W), H, sb.Z2, 4 * H, B, 4 * H, H);
    gh.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gh, st));
    XG_TRY(launch(ctx, "dec_cell", dec_cell_kernel, ceil_div(B * H, 256), 256, 0, st, sb.Z2, s.c2p, s.h2p, s.ld_h2p, mask, mask_stride, B, H,
                                                         drops ? drops[2] : make_drop(false, 0.f, 0, 0), s.c2n, s.h2n, s.ld_h2n, nullptr, 0));
  }
  return XG_OK;
}

// logits (rows, V) = out (rows, H; ld) . W_logit^T + b
static int logits_core(xg_context* ctx, const float* out, long ld_out, int rows, float* logits, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  GemmP g = gemm_nt(out, ld_out, P_(ctx, XG_P_LOGIT_W), d.rnn, logits, d.vocab, rows, d.vocab, d.rnn);
  g.ep.bias0 = P_(ctx, XG_P_LOGIT_B);
  return gemm_run(ctx, g, st);
}

// ------------------------------------------------------------------------------------
// SAModel.forward (SAModel.py:67-115), teacher forced; everything that depends only on the
// ground-truth tokens is hoisted out of the word loop and batched over all L' steps.
// ------------------------------------------------------------------------------------
static int train_fwd_core(xg_context* ctx, const float* rgb, const float* opfl, const float* fmask, const float* pos,
                          const int64_t* seq, const float* seq_mask, int B, int K, int L, int Lp, int train,
                          uint64_t seed, float* logp, float* cat, TrainSaved& S, float* logits_scratch,
                          float* cls_scratch, cudaStream_t st) {
  const xg_dims& d = ctx->d;
  const int H = d.rnn, E = d.embed, A = d.att, V = d.vocab, C = d.categories, Q = d.cls_hidden;
  const long LB = (long)Lp * B;
  float* st0[4] = {S.H12, S.C1, S.H12 + H, S.C2};   // h1 | h2 interleaved in H12[0] with ld 2H
  // encoder + Uv; the init state goes straight into H12[0] / C1[0] / C2[0]
  XG_TRY(encode_core(ctx, rgb, opfl, fmask, B, K, train, seed, S.enc, S.V, S.Uv, nullptr, st));
  train &= 1;      // bit 1 (no running-statistics update) only concerns the encoder's BatchNorm
  XG_TRY(init_hidden_core(ctx, S.V, fmask, B, K, S.enc.meanV, st0, 2 * H, st));
  // hoisted over all steps: embedding, POS gate, input parts of lstm_1
  XG_TRY(launch(ctx, "gather_rows", gather_rows_kernel, (int)LB, 128, 0, st, P_(ctx, XG_P_EMBED_W), seq, L, 1, B, (int)LB, E, V, S.XT));
  {
    TcHold hold(ctx);      // S.XT feeds the POS gate and lstm_1 in the same layout: one operand split
    GemmP g = gemm_nt(S.XT, E, P_(ctx, XG_P_DGATE_W), E, S.GP, H, (int)LB, H, E);
    g.ep.bias0 = P_(ctx, XG_P_DGATE_B);
    g.ep.act = XG_ACT_RELU;
    g.ep.drop = make_drop(train, d.drop_prob, seed, XG_DROP_DEC_GATE);
    g.ep.aux = S.RG; g.ep.ldaux = H;
    g.ep.tgt = pos; g.ep.ldt = H; g.ep.tgt_mod = B;
    XG_TRY(gemm_run(ctx, g, st));
    GemmP gi = gemm_nt(S.XT, E, P_(ctx, XG_P_L1_I2H_W), E, S.G1, 4 * H, (int)LB, 4 * H, E);
    gi.ep.bias0 = P_(ctx, XG_P_L1_I2H_B); gi.ep.bias1 = P_(ctx, XG_P_L1_A2H_B); gi.ep.bias2 = P_(ctx, XG_P_L1_H2H_B);
    XG_TRY(gemm_run(ctx, gi, st));
    GemmP ga = gemm_nt(S.GP, H, P_(ctx, XG_P_L1_A2H_W), H, S.G1, 4 * H, (int)LB, 4 * H, H);
    ga.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, ga, st));
  }
  // the word loop: one persistent cooperative kernel for all L' steps when the shape allows it (xg_persist.cuh)
  int pst = PK_FALLBACK;
  if (Lp >= 1) {
    PersistTrainIO io;
    io.seq_mask = seq_mask; io.L = L;
    io.G1 = S.G1; io.G2 = S.G2; io.C1 = S.C1; io.C2 = S.C2; io.H12 = S.H12; io.AH = S.AH; io.ALPHA = S.ALPHA; io.AF = S.AF;
    io.drop1 = make_drop(train, d.drop_prob, seed, XG_DROP_DEC_H1);
    io.drop2 = make_drop(train, d.drop_prob, seed, XG_DROP_DEC_H2);
    pst = grouped_train(ctx, S.V, S.Uv, B, K, Lp, io, st);      // grouped-cell form (xg_grouped.cuh) when the shape fits
    if (pst == PK_FALLBACK) pst = persist_decode(ctx, S.V, S.Uv, pos, nullptr, B, K, Lp, nullptr, nullptr, nullptr, &io, st);
    if (pst != PK_FALLBACK) XG_TRY(pst);
  }
  for (int i = 0; i < Lp && pst == PK_FALLBACK; ++i) {
    const float* h12p = S.H12 + (long)i * B * 2 * H;
    float* h12n = S.H12 + (long)(i + 1) * B * 2 * H;
    float* AHi = S.AH + (long)i * B * A;
    float* AFi = S.AF + (long)i * B * H;
    float* Z1 = S.G1 + (long)i * B * 4 * H;
    float* Z2 = S.G2 + (long)i * B * 4 * H;
    const float* m = seq_mask + i;   // (B,) stride L
    // attention on previous [h1,h2]
    GemmP g = gemm_nt(h12p, 2 * H, P_(ctx, XG_P_H2A_W), 2 * H, AHi, A, B, A, 2 * H);
    g.ep.bias0 = P_(ctx, XG_P_H2A_B);
    XG_TRY(gemm_run(ctx, g, st));
    XG_TRY(launch(ctx, "att_fwd", att_fwd_kernel, B, 256, (A + K) * sizeof(float), st, AHi, S.Uv, S.V, P_(ctx, XG_P_A2W_W), P_(ctx, XG_P_A2W_B), K,
                                                           A, H, 1, S.ALPHA + (long)i * B * K, AFi));
    // lstm_1: recurrent part only
    GemmP gh = gemm_nt(h12p, 2 * H, P_(ctx, XG_P_L1_H2H_W), H, Z1, 4 * H, B, 4 * H, H);
    gh.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, gh, st));
    XG_TRY(launch(ctx, "dec_cell", dec_cell_kernel, ceil_div(B * H, 256), 256, 0, st, 
        Z1, S.C1 + (long)i * B * H, h12p, 2 * H, m, L, B, H,
        make_drop(train, d.drop_prob, seed, XG_DROP_DEC_H1, (uint64_t)i * B * H), S.C1 + (long)(i + 1) * B * H, h12n,
        2 * H, nullptr, 0));
    // lstm_2
    GemmP g2 = gemm_nt(h12n, 2 * H, P_(ctx, XG_P_L2_I2H_W), H, Z2, 4 * H, B, 4 * H, H);
    g2.ep.bias0 = P_(ctx, XG_P_L2_I2H_B); g2.ep.bias1 = P_(ctx, XG_P_L2_A2H_B); g2.ep.bias2 = P_(ctx, XG_P_L2_H2H_B);
    XG_TRY(gemm_run(ctx, g2, st));
    GemmP g2a = gemm_nt(AFi, H, P_(ctx, XG_P_L2_A2H_W), H, Z2, 4 * H, B, 4 * H, H);
    g2a.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, g2a, st));
    GemmP g2h = gemm_nt(h12p + H, 2 * H, P_(ctx, XG_P_L2_H2H_W), H, Z2, 4 * H, B, 4 * H, H);
    g2h.ep.beta = 1.f;
    XG_TRY(gemm_run(ctx, g2h, st));
    XG_TRY(launch(ctx, "dec_cell", dec_cell_kernel, ceil_div(B * H, 256), 256, 0, st, 
        Z2, S.C2 + (long)i * B * H, h12p + H, 2 * H, m, L, B, H,
        make_drop(train, d.drop_prob, seed, XG_DROP_DEC_H2, (uint64_t)i * B * H), S.C2 + (long)(i + 1) * B * H,
        h12n + H, 2 * H, nullptr, 0));
  }
  // heads, batched over all steps: OUT = H12[1..Lp][:, H:2H] is a (L'B, H) matrix with ld 2H
  const float* OUT = S.H12 + (long)B * 2 * H + H;
  if (logp) {
    // logits land directly in the (B,L',V) output buffer (row (i,b) -> b*L'+i), log-softmax in place
    GemmP g = gemm_nt(OUT, 2 * H, P_(ctx, XG_P_LOGIT_W), H, logp, V, (int)LB, V, H);
    g.ep.bias0 = P_(ctx, XG_P_LOGIT_B);
    g.perm_rb = B; g.perm_rs = Lp;
    XG_TRY(gemm_run(ctx, g, st));
    if (logsoftmax_reg_ok(logp, V, V, logp, V)) {
      XG_TRY(launch(ctx, "logsoftmax_rows", logsoftmax_rows_reg_kernel, (int)LB, 256, 0, st, logp, V, V, 0, 0, logp, V));
    } else {
      XG_TRY(launch(ctx, "logsoftmax_rows", logsoftmax_rows_kernel, (int)LB, 256, 0, st, logp, V, V, 0, 0, logp, V));
    }
  }
  if (cat) {
    GemmP g = gemm_nt(OUT, 2 * H, P_(ctx, XG_P_CLS0_W), H, S.Hc, Q, (int)LB, Q, H);
    g.ep.bias0 = P_(ctx, XG_P_CLS0_B);
    g.ep.act = XG_ACT_RELU;
    g.ep.drop = make_drop(train, d.drop_prob, seed, XG_DROP_CLS);
    XG_TRY(gemm_run(ctx, g, st));
    GemmP g3 = gemm_nt(S.Hc, Q, P_(ctx, XG_P_CLS3_W), Q, cat, C, (int)LB, C, Q);
    g3.ep.bias0 = P_(ctx, XG_P_CLS3_B);
    g3.perm_rb = B; g3.perm_rs = Lp;
    XG_TRY(gemm_run(ctx, g3, st));
    XG_TRY(launch(ctx, "logsoftmax_rows", logsoftmax_rows_kernel, (int)LB, 128, 0, st, cat, C, C, 0, 0, cat, C));
  }
  (void)logits_scratch; (void)cls_scratch;
  return XG_OK;
}

}  // namespace xg
