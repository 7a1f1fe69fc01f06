// Handle, parameter table, buffer carving shared by the entry points.
#pragma once
#include <map>
#include <utility>
#include <vector>

#include "xg_common.cuh"
#include "xg_gemm.cuh"

struct xg_context {
  xg_dims d;
  int device = 0;
  int sm_count = 148;
  const float* P[XG_NUM_PARAMS];
  bool bound = false;
  unsigned long long param_epoch = 0;   // bumped by xg_bind_params / xg_params_changed: derived tables are stale
  float* bn[4] = {nullptr, nullptr, nullptr, nullptr};  // rm_rgb, rv_rgb, rm_opfl, rv_opfl
  bool bn_bound = false;
  xg::ErrorSink es;
  int* h_pinned = nullptr;   // small pinned staging area for SYNC read-backs
  int* d_small = nullptr;    // matching device words
  xg::SplitKScratch splitk;  // partial tiles + tile counters of the split-K SIMT GEMM
  // optional per-kernel timing with CUDA events on the launching stream (xg_profile_enable/_report)
  int tc_mode = 1;           // 1: dense contractions above the size gate run on the tcgen05 3xTF32 engine
  int persist_mode = 1;      // 1: greedy decoding runs in the fused persistent word-step kernel when eligible
  int strict_persist = 0;    // xg_set_strict / XG_STRICT_PERSIST=1: a serial loop that cannot run on its persistent kernel is an ERROR
  unsigned long long n_fused = 0, n_unfused = 0;   // serial loops run by a persistent kernel / by per-step launches (xg_path_counters)
  bool tc_f16 = false;                    // batched tcgen05 products on fp16 operand pairs (set by the forward entry points, TcF16Scope)
  cudaEvent_t bwd_split_event = nullptr;   // xg_set_bwd_split_event: recorded by xg_train_bwd once every decoder-side gradient is final
  int dec_drop_on = 0;       // xg_set_decode_dropout: xg_sample_greedy applies the TRAINING dropout of the word step
  unsigned long long dec_drop_seed = 0;   //   (same Philox sites / indices as xg_train_fwd with this seed)
  bool prof_on = false;
  struct ProfRec { std::string name; cudaEvent_t e0, e1; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  static constexpr int kPinnedInts = 1024;
  static constexpr int kSmallInts = 1 << 18;   // device scratch words (criterion terms etc.)
};

namespace xg {

// RAII timing scope around one launch (no-op unless profiling is enabled on the handle)
struct ProfScope {
  xg_context* c; cudaStream_t st; cudaEvent_t e0 = nullptr, e1 = nullptr; const char* name; std::string dyn;
  static cudaEvent_t get(xg_context* c) {
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  ProfScope(xg_context* ctx, const char* nm, cudaStream_t s) : c(ctx), st(s), name(nm) {
    if (c && c->prof_on) { e0 = get(c); e1 = get(c); cudaEventRecord(e0, st); }
  }
  ProfScope(xg_context* ctx, std::string nm, cudaStream_t s) : c(ctx), st(s), name(nullptr), dyn(std::move(nm)) {
    if (c && c->prof_on) { e0 = get(c); e1 = get(c); cudaEventRecord(e0, st); }
  }
  ~ProfScope() {
    if (e0) { cudaEventRecord(e1, st); c->prof_recs.push_back({name ? std::string(name) : dyn, e0, e1}); }
  }
};

// every non-GEMM kernel goes through this launcher (launch check + optional timing)
template <typename... KArgs, typename... Args>
static int launch(xg_context* ctx, const char* name, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                  cudaStream_t st, Args&&... args) {
  ProfScope ps(ctx, name, st);
  kernel<<<grid, block, smem, st>>>(std::forward<Args>(args)...);
  XG_LAUNCH_CHECK(ctx->es);
  return XG_OK;
}

inline void param_shape(const xg_dims& d, int idx, int* rows, int* cols) {
  const int H = d.rnn, R = d.feat_rgb, F = d.feat_opfl, E = d.embed, A = d.att, V = d.vocab, C = d.categories,
            Q = d.cls_hidden;
  int r = 0, c = 1;
  switch (idx) {
    case XG_P_EMB_RGB_W: r = H; c = R; break;
    case XG_P_EMB_OPFL_W: r = H; c = F; break;
    case XG_P_EMB_RGB_B: case XG_P_BN_RGB_G: case XG_P_BN_RGB_B:
    case XG_P_EMB_OPFL_B: case XG_P_BN_OPFL_G: case XG_P_BN_OPFL_B: r = H; break;
    case XG_P_LSTM_RGB_WIH: case XG_P_LSTM_RGB_WHH: case XG_P_LSTM_OPFL_WIH: case XG_P_LSTM_OPFL_WHH: r = 4 * H; c = H; break;
    case XG_P_LSTM_RGB_BIH: case XG_P_LSTM_RGB_BHH: case XG_P_LSTM_OPFL_BIH: case XG_P_LSTM_OPFL_BHH: r = 4 * H; break;
    case XG_P_GATE_RGB_W: case XG_P_GATE_OPFL_W: r = H; c = H; break;
    case XG_P_GATE_RGB_B: case XG_P_GATE_OPFL_B: r = H; break;
    case XG_P_FUSION_W: r = H; c = 2 * H; break;
    case XG_P_FUSION_B: r = H; break;
    case XG_P_INIT_H1_W: case XG_P_INIT_C1_W: case XG_P_INIT_H2_W: case XG_P_INIT_C2_W: r = H; c = H; break;
    case XG_P_INIT_H1_B: case XG_P_INIT_C1_B: case XG_P_INIT_H2_B: case XG_P_INIT_C2_B: r = H; break;
    case XG_P_DGATE_W: r = H; c = E; break;
    case XG_P_DGATE_B: r = H; break;
    case XG_P_L1_I2H_W: r = 4 * H; c = E; break;
    case XG_P_L1_A2H_W: case XG_P_L1_H2H_W: case XG_P_L2_I2H_W: case XG_P_L2_A2H_W: case XG_P_L2_H2H_W: r = 4 * H; c = H; break;
    case XG_P_L1_I2H_B: case XG_P_L1_A2H_B: case XG_P_L1_H2H_B:
    case XG_P_L2_I2H_B: case XG_P_L2_A2H_B: case XG_P_L2_H2H_B: r = 4 * H; break;
    case XG_P_V2A_W: r = A; c = H; break;
    case XG_P_V2A_B: r = A; break;
    case XG_P_H2A_W: r = A; c = 2 * H; break;
    case XG_P_H2A_B: r = A; break;
    case XG_P_A2W_W: r = 1; c = A; break;
    case XG_P_A2W_B: r = 1; break;
    case XG_P_EMBED_W: r = V; c = E; break;
    case XG_P_LOGIT_W: r = V; c = H; break;
    case XG_P_LOGIT_B: r = V; break;
    case XG_P_CLS0_W: r = Q; c = H; break;
    case XG_P_CLS0_B: r = Q; break;
    case XG_P_CLS3_W: r = C; c = Q; break;
    case XG_P_CLS3_B: r = C; break;
    default: break;
  }
  *rows = r;
  *cols = c;
}

// ---- encoder buffers (frame-major recurrent buffers) ----
struct EncBufs {
  float* Y[2];       // (B*K, H) pre-BN linear output, rows (b,k)
  float* mean[2];    // (H)
  float* invstd[2];  // (H)
  float* scale[2];   // (H)
  float* shift[2];   // (H)
  float* E[2];       // (K*B, H) embedded stream, rows (k,b)
  float* G[2];       // (K, B, 4H) gate pre-activations -> activated gates
  float* Hs[2];      // (K, B, H)
  float* Cs[2];      // (K, B, H)
  float* R[2];       // (K*B, H) cross-gate relu (post dropout)
  float* GG;         // (K*B, 2H) gated hidden states [rgb | opfl]
  float* meanV;      // (B, H)
  double* part;      // BN partial sums
};

inline int bn_row_splits(int M) {
  int rs = (M + 63) / 64;
  return rs < 1 ? 1 : (rs > 32 ? 32 : rs);
}

inline void carve_enc(Arena& a, const xg_dims& d, int B, int K, EncBufs& e) {
  const long H = d.rnn, BK = (long)B * K;
  for (int s = 0; s < 2; ++s) {
    e.Y[s] = a.take<float>(BK * H);
    e.mean[s] = a.take<float>(H);
    e.invstd[s] = a.take<float>(H);
    e.scale[s] = a.take<float>(H);
    e.shift[s] = a.take<float>(H);
    e.E[s] = a.take<float>(BK * H);
    e.G[s] = a.take<float>(BK * 4 * H);
    e.Hs[s] = a.take<float>(BK * H);
    e.Cs[s] = a.take<float>(BK * H);
    e.R[s] = a.take<float>(BK * H);
  }
  e.GG = a.take<float>(BK * 2 * H);
  e.meanV = a.take<float>((long)B * H);
  e.part = a.take<double>((long)bn_row_splits((int)BK) * H * 2);
}

// ---- decoder per-step scratch ----
struct StepBufs {
  float* XT;   // (B, E)
  float* AH;   // (B, A)
  float* AF;   // (B, H)
  float* GP;   // (B, H)
  float* Z1;   // (B, 4H)
  float* Z2;   // (B, 4H)
};
inline void carve_step(Arena& a, const xg_dims& d, int B, StepBufs& s) {
  s.XT = a.take<float>((long)B * d.embed);
  s.AH = a.take<float>((long)B * d.att);
  s.AF = a.take<float>((long)B * d.rnn);
  s.GP = a.take<float>((long)B * d.rnn);
  s.Z1 = a.take<float>((long)B * 4 * d.rnn);
  s.Z2 = a.take<float>((long)B * 4 * d.rnn);
}

// ---- activations kept between xg_train_fwd and xg_train_bwd ----
struct TrainSaved {
  EncBufs enc;
  float* V;      // (B,K,H)
  float* Uv;     // (B,K,A)
  float* XT;     // (L*B, E) rows (i,b)
  float* RG;     // (L*B, H)
  float* GP;     // (L*B, H)
  float* G1;     // (L, B, 4H)
  float* G2;     // (L, B, 4H)
  float* H12;    // (L+1, B, 2H)   [h1 | h2] entering step i
  float* C1;     // (L+1, B, H)
  float* C2;     // (L+1, B, H)
  float* AH;     // (L, B, A)
  float* ALPHA;  // (L, B, K)
  float* AF;     // (L, B, H)
  float* Hc;     // (L*B, cls_hidden)
};
inline void carve_saved(Arena& a, const xg_dims& d, int B, int K, int L, TrainSaved& s) {
  const long H = d.rnn, LB = (long)L * B;
  carve_enc(a, d, B, K, s.enc);
  s.V = a.take<float>((long)B * K * H);
  s.Uv = a.take<float>((long)B * K * d.att);
  s.XT = a.take<float>(LB * d.embed);
  s.RG = a.take<float>(LB * H);
  s.GP = a.take<float>(LB * H);
  s.G1 = a.take<float>(LB * 4 * H);
  s.G2 = a.take<float>(LB * 4 * H);
  s.H12 = a.take<float>((long)(L + 1) * B * 2 * H);
  s.C1 = a.take<float>((long)(L + 1) * B * H);
  s.C2 = a.take<float>((long)(L + 1) * B * H);
  s.AH = a.take<float>(LB * d.att);
  s.ALPHA = a.take<float>(LB * K);
  s.AF = a.take<float>(LB * H);
  s.Hc = a.take<float>(LB * d.cls_hidden);
}

}  // namespace xg
