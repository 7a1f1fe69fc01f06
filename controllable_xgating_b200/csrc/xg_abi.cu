// libxgating.so — C ABI entry points (include/xgating.h).  Single translation unit, sm_100a.
#include <cstring>
#include <new>

#include "xg_backward.cuh"
#include "xg_beam.cuh"
#include "xg_forward.cuh"
#include "xg_optim.cuh"
#include "xg_persist.cuh"
#include "xg_grouped.cuh"

using namespace xg;

namespace {

thread_local std::string g_last_error;   // errors raised before a handle exists

int fail(xg_context* ctx, int code, const char* text) {
  if (ctx) ctx->es.msg = text; else g_last_error = text;
  return code;
}

#define CHECK_HANDLE(h)                                                        \
  do {                                                                         \
    if (!(h)) return fail(nullptr, XG_ERR_NULL_POINTER, "null handle");        \
    (h)->es.msg.clear();                                                       \
  } while (0)
#define CHECK_BOUND(h)                                                                         \
  do {                                                                                         \
    if (!(h)->bound) return fail((h), XG_ERR_NOT_BOUND, "parameters not bound (xg_bind_params)"); \
  } while (0)
#define CHECK_PTR(h, p)                                                        \
  do {                                                                         \
    if (!(p)) return fail((h), XG_ERR_NULL_POINTER, "null pointer: " #p);      \
  } while (0)
#define CHECK_POS(h, v)                                                        \
  do {                                                                         \
    if ((v) <= 0) return fail((h), XG_ERR_BAD_SHAPE, "non-positive size: " #v);  \
  } while (0)

int set_device(xg_context* ctx) {
  XG_CUDA_TRY(ctx->es, cudaSetDevice(ctx->device));
  return XG_OK;
}

struct GreedyBufs {
  StepBufs step;
  float* st[4];
  float* logits;
  int64_t* tok;
  float* unfinished;
  int* flags;
  float* Uv;
};
void carve_greedy(Arena& a, const xg_dims& d, int B, int K, int T, GreedyBufs& g) {
  carve_step(a, d, B, g.step);
  for (int q = 0; q < 4; ++q) g.st[q] = a.take<float>((long)B * d.rnn);
  g.logits = a.take<float>((long)B * d.vocab);
  g.tok = a.take<int64_t>(B);
  g.unfinished = a.take<float>(B);
  g.flags = a.take<int>(T + 1);
  g.Uv = a.take<float>((long)B * K * d.att);
}

struct DecStepBufs {
  StepBufs step;
  float* Uv;
  float* logits;
};
void carve_decstep(Arena& a, const xg_dims& d, int B, int K, DecStepBufs& s) {
  carve_step(a, d, B, s.step);
  s.Uv = a.take<float>((long)B * K * d.att);
  s.logits = a.take<float>((long)B * d.vocab);
}

}  // namespace

extern "C" {

int xg_abi_version(void) { return XG_ABI_VERSION; }

const char* xg_status_string(int s) {
  switch (s) {
    case XG_OK: return "ok";
    case XG_ERR_BAD_ARG: return "bad argument";
    case XG_ERR_BAD_SHAPE: return "bad shape";
    case XG_ERR_NULL_POINTER: return "null pointer";
    case XG_ERR_CUDA: return "CUDA error";
    case XG_ERR_NOT_BOUND: return "parameters / buffers not bound";
    case XG_ERR_WORKSPACE: return "workspace too small";
    case XG_ERR_UNSUPPORTED: return "unsupported";
    default: return "unknown status";
  }
}

const char* xg_last_error(xg_handle h) { return h ? h->es.msg.c_str() : g_last_error.c_str(); }

int xg_create(const xg_dims* dims, int device, xg_handle* out) {
  if (!dims || !out) return fail(nullptr, XG_ERR_NULL_POINTER, "xg_create: null argument");
  const xg_dims& d = *dims;
  if (d.feat_rgb <= 0 || d.feat_opfl <= 0 || d.rnn <= 0 || d.embed <= 0 || d.att <= 0 || d.vocab <= 0 ||
      d.categories <= 0 || d.cls_hidden <= 0)
    return fail(nullptr, XG_ERR_BAD_SHAPE, "xg_create: dimensions must be positive");
  if (d.fusion_act < XG_ACT_RELU || d.fusion_act > XG_ACT_SIGMOID)
    return fail(nullptr, XG_ERR_BAD_ARG, "xg_create: fusion_act must be ReLU/Tanh/Sigmoid");
  if (!(d.drop_prob >= 0.f && d.drop_prob < 1.f))
    return fail(nullptr, XG_ERR_BAD_ARG, "xg_create: drop_prob must be in [0,1)");   // myopts.py:79
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    g_last_error = std::string("xg_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
    return XG_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) return fail(nullptr, XG_ERR_BAD_ARG, "xg_create: bad device ordinal");
  xg_context* ctx = new (std::nothrow) xg_context();
  if (!ctx) return fail(nullptr, XG_ERR_BAD_ARG, "xg_create: out of host memory");
  ctx->d = d;
  ctx->device = device;
  for (int i = 0; i < XG_NUM_PARAMS; ++i) ctx->P[i] = nullptr;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, XG_ERR_CUDA, "xg_create: cannot query device");
  }
  if (prop.major != 10) {
    delete ctx;
    return fail(nullptr, XG_ERR_UNSUPPORTED, "xg_create: this library is built for sm_100a (B200) only");
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->strict_persist = env_flag("XG_STRICT_PERSIST") ? 1 : 0;
  if (cudaMallocHost(&ctx->h_pinned, sizeof(int) * xg_context::kPinnedInts) != cudaSuccess ||
      cudaMalloc(&ctx->d_small, sizeof(int) * xg_context::kSmallInts) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, XG_ERR_CUDA, "xg_create: cannot allocate staging words");
  }
  ctx->splitk.ws_floats = (size_t)320 * 64 * 64;   // 296 CTAs x one 64x64 partial tile (5 MB)
  ctx->splitk.ctr_count = 128;     // + 128 completion counters of the skinny kernel
  if (cudaMalloc(&ctx->splitk.ws, sizeof(float) * ctx->splitk.ws_floats) != cudaSuccess ||
      cudaMalloc(&ctx->splitk.ctr, sizeof(unsigned int) * 2 * ctx->splitk.ctr_count) != cudaSuccess ||
      cudaMemset(ctx->splitk.ctr, 0, sizeof(unsigned int) * 2 * ctx->splitk.ctr_count) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, XG_ERR_CUDA, "xg_create: cannot allocate split-K scratch");
  }
  *out = ctx;
  return XG_OK;
}

int xg_destroy(xg_handle h) {
  if (!h) return XG_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  persist_release(h);
  grouped_release(h);
  grouped_enc_release(h);
  grouped_step_release(h);
  grouped_train_release(h);
  colsum_release(h);
  word_tables_release(h);
  tc_release(h);
  for (auto& r : h->prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto e : h->prof_pool) cudaEventDestroy(e);
  if (h->h_pinned) cudaFreeHost(h->h_pinned);
  if (h->d_small) cudaFree(h->d_small);
  if (h->splitk.ws) cudaFree(h->splitk.ws);
  if (h->splitk.ctr) cudaFree(h->splitk.ctr);
  delete h;
  return XG_OK;
}

int xg_param_shape(xg_handle h, int index, int* rows, int* cols) {
  CHECK_HANDLE(h);
  if (index < 0 || index >= XG_NUM_PARAMS || !rows || !cols) return fail(h, XG_ERR_BAD_ARG, "xg_param_shape: bad index");
  param_shape(h->d, index, rows, cols);
  return XG_OK;
}

int xg_bind_params(xg_handle h, const float* const* params, int count) {
  CHECK_HANDLE(h);
  CHECK_PTR(h, params);
  if (count != XG_NUM_PARAMS) return fail(h, XG_ERR_BAD_ARG, "xg_bind_params: expected XG_NUM_PARAMS pointers");
  for (int i = 0; i < count; ++i)
    if (!params[i]) return fail(h, XG_ERR_NULL_POINTER, "xg_bind_params: null parameter pointer");
  for (int i = 0; i < count; ++i) h->P[i] = params[i];
  h->bound = true;
  h->param_epoch++;
  return XG_OK;
}

int xg_bind_bn_buffers(xg_handle h, float* rm_rgb, float* rv_rgb, float* rm_opfl, float* rv_opfl) {
  CHECK_HANDLE(h);
  if (!rm_rgb || !rv_rgb || !rm_opfl || !rv_opfl) return fail(h, XG_ERR_NULL_POINTER, "xg_bind_bn_buffers: null buffer");
  h->bn[0] = rm_rgb; h->bn[1] = rv_rgb; h->bn[2] = rm_opfl; h->bn[3] = rv_opfl;
  h->bn_bound = true;
  return XG_OK;
}

int xg_params_changed(xg_handle h) {
  CHECK_HANDLE(h);
  // no device synchronisation: the derived copies keep their buffers and are refilled on the stream of the next call
  // that uses them, i.e. after everything already queued on it (the update itself must be ordered on that stream too)
  tc_invalidate_weights(h);   // tf32 hi/lo splits of the bound parameters
  h->param_epoch++;           // POS-gate token table of the persistent decoder
  return XG_OK;
}

int xg_set_decode_dropout(xg_handle h, int on, uint64_t seed) {
  CHECK_HANDLE(h);
  h->dec_drop_on = on ? 1 : 0;
  h->dec_drop_seed = seed;
  return XG_OK;
}

int xg_set_strict(xg_handle h, int on) {
  CHECK_HANDLE(h);
  h->strict_persist = on ? 1 : 0;
  return XG_OK;
}

int xg_path_counters(xg_handle h, uint64_t* fused, uint64_t* unfused) {
  CHECK_HANDLE(h);
  if (fused) *fused = h->n_fused;
  if (unfused) *unfused = h->n_unfused;
  return XG_OK;
}

int xg_set_bwd_split_event(xg_handle h, void* cuda_event) {
  CHECK_HANDLE(h);
  h->bwd_split_event = (cudaEvent_t)cuda_event;
  return XG_OK;
}

int xg_set_engine(xg_handle h, int tensor_cores) {
  CHECK_HANDLE(h);
  h->tc_mode = tensor_cores >= 1 ? 1 : 0;
  h->persist_mode = tensor_cores >= 2 ? 1 : 0;
  return XG_OK;
}

size_t xg_workspace_bytes(xg_handle h, int kind, int B, int K, int L_or_T, int beam) {
  if (!h || B <= 0 || K <= 0) return 0;
  Arena a(nullptr, 0);
  const xg_dims& d = h->d;
  switch (kind) {
    case XG_WS_ENCODE: { EncBufs e; carve_enc(a, d, B, K, e); break; }
    case XG_WS_DECODE_STEP: { DecStepBufs s; carve_decstep(a, d, B, K, s); break; }
    case XG_WS_GREEDY: { GreedyBufs g; carve_greedy(a, d, B, K, L_or_T > 0 ? L_or_T : 1, g); break; }
    case XG_WS_BEAM: {
      if (beam <= 0 || beam > XG_MAX_BEAM || L_or_T <= 0) return 0;
      BeamBufs w; carve_beam(a, d, B, K, L_or_T, beam, w); break;
    }
    case XG_WS_TRAIN_SAVED: { if (L_or_T <= 0) return 0; TrainSaved s; carve_saved(a, d, B, K, L_or_T, s); break; }
    case XG_WS_TRAIN_FWD: { a.take<float>(64); break; }
    case XG_WS_TRAIN_BWD: { if (L_or_T <= 0) return 0; BwdBufs w; carve_bwd(a, d, B, K, L_or_T, w); break; }
    default: return 0;
  }
  return a.off + 256;
}

// Forward entry points run their batched tcgen05 products on fp16 operand pairs (xg_gemm_tc.cuh, 3xFP16: the accuracy of
// the tf32 pairs at half the operand bytes; |operand| < 65504).  The backward pass keeps tf32 pairs: gradients do not
// fit the fp16 range without a scale.  XG_NO_TC16=1 keeps tf32 pairs everywhere.
struct TcF16Scope {
  xg_context* c; bool old;
  explicit TcF16Scope(xg_context* c_) : c(c_), old(c_->tc_f16) { c->tc_f16 = !env_flag("XG_NO_TC16"); }
  ~TcF16Scope() { c->tc_f16 = old; }
};

int xg_encode_fwd(xg_handle h, const float* rgb, const float* opfl, const float* feat_mask, int B, int K, int train,
                  uint64_t seed, float* V_out, float* Uv_out, float* const* state_out, void* ws, size_t ws_bytes,
                  void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, rgb); CHECK_PTR(h, opfl); CHECK_PTR(h, feat_mask); CHECK_PTR(h, V_out); CHECK_PTR(h, ws);
  CHECK_POS(h, B); CHECK_POS(h, K);
  if (!h->bn_bound) return fail(h, XG_ERR_NOT_BOUND, "BatchNorm buffers not bound (xg_bind_bn_buffers)");
  if (state_out) for (int q = 0; q < 4; ++q) CHECK_PTR(h, state_out[q]);
  XG_TRY(set_device(h));
  TcF16Scope f16(h);
  Arena a(ws, ws_bytes);
  EncBufs eb; carve_enc(a, h->d, B, K, eb);
  if (a.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_encode_fwd: workspace too small");
  return encode_core(h, rgb, opfl, feat_mask, B, K, train, seed, eb, V_out, Uv_out, state_out, (cudaStream_t)stream);
}

int xg_init_hidden(xg_handle h, const float* V, const float* feat_mask, int B, int K, float* const* state_out, void* ws,
                   size_t ws_bytes, void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, V); CHECK_PTR(h, feat_mask); CHECK_PTR(h, state_out); CHECK_PTR(h, ws);
  CHECK_POS(h, B); CHECK_POS(h, K);
  for (int q = 0; q < 4; ++q) CHECK_PTR(h, state_out[q]);
  XG_TRY(set_device(h));
  TcF16Scope f16(h);
  Arena a(ws, ws_bytes);
  float* meanV = a.take<float>((long)B * h->d.rnn);
  if (a.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_init_hidden: workspace too small");
  return init_hidden_core(h, V, feat_mask, B, K, meanV, state_out, h->d.rnn, (cudaStream_t)stream);
}

int xg_attend_precompute(xg_handle h, const float* V, int B, int K, float* Uv_out, void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, V); CHECK_PTR(h, Uv_out); CHECK_POS(h, B); CHECK_POS(h, K);
  XG_TRY(set_device(h));
  TcF16Scope f16(h);
  GemmP g = gemm_nt(V, h->d.rnn, h->P[XG_P_V2A_W], h->d.rnn, Uv_out, h->d.att, B * K, h->d.att, h->d.rnn);
  g.ep.bias0 = h->P[XG_P_V2A_B];
  return gemm_run(h, g, (cudaStream_t)stream);
}

int xg_decode_step(xg_handle h, const int64_t* tokens, const float* xt, const float* xt_mask, const float* V, const float* Uv,
                   const float* pos, const float* const* state_in, float* const* state_out, float* out, float* logp,
                   int B, int K, void* ws, size_t ws_bytes, void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, V); CHECK_PTR(h, pos); CHECK_PTR(h, state_in); CHECK_PTR(h, state_out); CHECK_PTR(h, ws);
  CHECK_POS(h, B); CHECK_POS(h, K);
  if ((tokens == nullptr) == (xt == nullptr)) return fail(h, XG_ERR_BAD_ARG, "xg_decode_step: pass exactly one of tokens / xt");
  for (int q = 0; q < 4; ++q) { CHECK_PTR(h, state_in[q]); CHECK_PTR(h, state_out[q]); }
  XG_TRY(set_device(h));
  TcF16Scope f16(h);
  cudaStream_t st = (cudaStream_t)stream;
  const xg_dims& d = h->d;
  const int H = d.rnn;
  Arena a(ws, ws_bytes);
  DecStepBufs sb; carve_decstep(a, d, B, K, sb);
  if (a.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_decode_step: workspace too small");
  if (!Uv) {
    XG_TRY(xg_attend_precompute(h, V, B, K, sb.Uv, stream));
    Uv = sb.Uv;
  }
  if (tokens) {
    XG_TRY(launch(h, "gather_rows", gather_rows_kernel, B, 128, 0, st, h->P[XG_P_EMBED_W], tokens, 1, 0, B, B, d.embed, d.vocab, sb.step.XT));
    xt = sb.step.XT;
  }
  StepState s{state_in[0], H, state_in[1], state_in[2], H, state_in[3],
              state_out[0], H, state_out[1], state_out[2], H, state_out[3]};
  XG_TRY(decode_step_core(h, xt, xt_mask, 1, V, Uv, pos, s, sb.step, nullptr, B, K, 1, st));
  if (out) XG_CUDA_TRY(h->es, cudaMemcpyAsync(out, state_out[2], sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
  if (logp) {
    XG_TRY(logits_core(h, state_out[2], H, B, logp, st));
    XG_TRY(launch(h, "logsoftmax_rows", logsoftmax_rows_kernel, B, 256, 0, st, logp, d.vocab, d.vocab, 0, 0, logp, d.vocab));
  }
  return XG_OK;
}

int xg_sample_greedy(xg_handle h, const float* V, const float* Uv, const float* pos, const float* const* state0, int B,
                     int K, int T, int sample_max, float temperature, uint64_t seed, int64_t* seq_out, float* logp_out,
                     int* steps_out, void* ws, size_t ws_bytes, void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, V); CHECK_PTR(h, pos); CHECK_PTR(h, state0); CHECK_PTR(h, seq_out); CHECK_PTR(h, logp_out);
  CHECK_PTR(h, ws);      // (steps_out may be NULL: asynchronous call)
  CHECK_POS(h, B); CHECK_POS(h, K); CHECK_POS(h, T);
  for (int q = 0; q < 4; ++q) CHECK_PTR(h, state0[q]);
  if (T + 1 > xg_context::kPinnedInts) return fail(h, XG_ERR_BAD_SHAPE, "xg_sample_greedy: seq_length too large");
  if (!sample_max && !(temperature > 0.f)) return fail(h, XG_ERR_BAD_ARG, "xg_sample_greedy: temperature must be > 0");
  XG_TRY(set_device(h));
  TcF16Scope f16(h);
  cudaStream_t st = (cudaStream_t)stream;
  const xg_dims& d = h->d;
  const int H = d.rnn;
  Arena a(ws, ws_bytes);
  GreedyBufs g; carve_greedy(a, d, B, K, T, g);
  if (a.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_sample_greedy: workspace too small");
  if (!Uv) {
    XG_TRY(xg_attend_precompute(h, V, B, K, g.Uv, stream));
    Uv = g.Uv;
  }
  const bool step_drop = h->dec_drop_on && d.drop_prob > 0.f;     // training-mode sampling (self-critical path)
  if (sample_max && !step_drop) {   // fused persistent word loop (xg_persist.cuh)
    // grouped-cell kernel (xg_grouped.cuh) when the shape fits it, else the six-phase kernel (xg_persist.cuh)
    int ps = grouped_decode(h, V, Uv, pos, state0, B, K, T, seq_out, logp_out, steps_out, st);
    if (ps == PK_FALLBACK) ps = persist_decode(h, V, Uv, pos, state0, B, K, T, seq_out, logp_out, steps_out, nullptr, st);
    if (ps != PK_FALLBACK) return ps;
  } else {          // sampling form of the grouped kernel: multinomial draw in the pick phase, training dropout in the cells
    GroupedSampling smp;
    smp.sample_max = sample_max; smp.temperature = temperature; smp.seed = seed;
    smp.step_drop = step_drop; smp.drop_seed = h->dec_drop_seed;
    int ps = grouped_decode(h, V, Uv, pos, state0, B, K, T, seq_out, logp_out, steps_out, st, &smp);
    if (ps == PK_FALLBACK) ps = persist_refuse(h, "the sampling word loop", "shape outside the grouped decoder: multinomial draws and training-mode dropout run on per-step launches");
    if (ps != PK_FALLBACK) return ps;
  }
  for (int q = 0; q < 4; ++q)
    XG_CUDA_TRY(h->es, cudaMemcpyAsync(g.st[q], state0[q], sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
  XG_CUDA_TRY(h->es, cudaMemsetAsync(g.tok, 0, sizeof(int64_t) * (size_t)B, st));          // <bos> (SAModel.py:184)
  XG_CUDA_TRY(h->es, cudaMemsetAsync(g.flags, 0, sizeof(int) * (size_t)(T + 1), st));
  XG_CUDA_TRY(h->es, cudaMemsetAsync(seq_out, 0, sizeof(int64_t) * (size_t)B * T, st));
  XG_CUDA_TRY(h->es, cudaMemsetAsync(logp_out, 0, sizeof(float) * (size_t)B * T, st));
  StepState s{g.st[0], H, g.st[1], g.st[2], H, g.st[3], g.st[0], H, g.st[1], g.st[2], H, g.st[3]};
  for (int t = 0; t <= T; ++t) {
    if (t >= 1) {
      XG_TRY(launch(h, "greedy_pick", greedy_pick_kernel, B, 256, 0, st, g.logits, d.vocab, t, T, sample_max, sample_max ? 1.f : 1.f / temperature, seed,
                                           seq_out, logp_out, g.tok, g.unfinished, g.flags));
    }
    if (t == T) break;   // the reference runs one more (unused) word step here (SAModel.py:216-217)
    XG_TRY(launch(h, "gather_rows", gather_rows_kernel, B, 128, 0, st, h->P[XG_P_EMBED_W], g.tok, 1, 0, B, B, d.embed, d.vocab, g.step.XT));
    DropSpec drops[3];
    if (step_drop) {
      const uint64_t base = (uint64_t)t * B * H;
      drops[0] = make_drop(true, d.drop_prob, h->dec_drop_seed, XG_DROP_DEC_GATE, base);
      drops[1] = make_drop(true, d.drop_prob, h->dec_drop_seed, XG_DROP_DEC_H1, base);
      drops[2] = make_drop(true, d.drop_prob, h->dec_drop_seed, XG_DROP_DEC_H2, base);
    }
    XG_TRY(decode_step_core(h, g.step.XT, t == 0 ? nullptr : g.unfinished, 1, V, Uv, pos, s, g.step, nullptr, B, K, 1, st,
                            step_drop ? drops : nullptr));
    XG_TRY(logits_core(h, g.st[2], H, B, g.logits, st));
  }
  if (!steps_out) return XG_OK;      // asynchronous call
  XG_CUDA_TRY(h->es, cudaMemcpyAsync(h->h_pinned, g.flags, sizeof(int) * (size_t)T, cudaMemcpyDeviceToHost, st));
  XG_CUDA_TRY(h->es, cudaStreamSynchronize(st));
  int steps = 0;
  while (steps < T && h->h_pinned[steps] != 0) ++steps;
  *steps_out = steps;
  return XG_OK;
}

int xg_scheduled_tokens(xg_handle h, const float* V, const float* Uv, const float* pos, const float* const* state0,
                        const int64_t* seq, const float* seq_mask, int B, int K, int L, int Lp, float ss_prob,
                        uint64_t ss_seed, int64_t* tokens_out, void* ws, size_t ws_bytes, void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, V); CHECK_PTR(h, pos); CHECK_PTR(h, state0); CHECK_PTR(h, seq); CHECK_PTR(h, seq_mask);
  CHECK_PTR(h, tokens_out); CHECK_PTR(h, ws);
  CHECK_POS(h, B); CHECK_POS(h, K); CHECK_POS(h, L); CHECK_POS(h, Lp);
  if (Lp > L) return fail(h, XG_ERR_BAD_SHAPE, "xg_scheduled_tokens: Lp > L");
  if (!(ss_prob >= 0.f && ss_prob <= 1.f)) return fail(h, XG_ERR_BAD_ARG, "xg_scheduled_tokens: ss_prob must be in [0,1]");
  for (int q = 0; q < 4; ++q) CHECK_PTR(h, state0[q]);
  XG_TRY(set_device(h));
  TcF16Scope f16(h);
  cudaStream_t st = (cudaStream_t)stream;
  const xg_dims& d = h->d;
  const int H = d.rnn;
  Arena a(ws, ws_bytes);
  GreedyBufs g; carve_greedy(a, d, B, K, L, g);
  if (a.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_scheduled_tokens: workspace too small");
  if (!Uv) {
    XG_TRY(xg_attend_precompute(h, V, B, K, g.Uv, stream));
    Uv = g.Uv;
  }
  const bool step_drop = h->dec_drop_on && d.drop_prob > 0.f;
  if (ss_prob > 0.f) {      // one launch of the sampling form of the grouped kernel (xg_grouped.cuh) when the shape fits
    XG_CUDA_TRY(h->es, cudaMemcpyAsync(tokens_out, seq, sizeof(int64_t) * (size_t)B * L, cudaMemcpyDeviceToDevice, st));
    GroupedSampling smp;
    smp.sample_max = 0; smp.temperature = 1.f; smp.step_drop = step_drop; smp.drop_seed = h->dec_drop_seed;
    smp.ss_mode = true; smp.ss_prob = ss_prob; smp.ss_seed = ss_seed; smp.ss_seq = seq; smp.ss_mask = seq_mask; smp.ss_used = tokens_out; smp.ss_L = L;
    int ps = grouped_decode(h, V, Uv, pos, state0, B, K, Lp, nullptr, nullptr, nullptr, st, &smp);
    if (ps == PK_FALLBACK) ps = persist_refuse(h, "the scheduled-sampling token pass", "shape outside the grouped decoder: it runs on per-step launches");
    if (ps != PK_FALLBACK) return ps;
  }
  for (int q = 0; q < 4; ++q)
    XG_CUDA_TRY(h->es, cudaMemcpyAsync(g.st[q], state0[q], sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
  XG_CUDA_TRY(h->es, cudaMemcpyAsync(tokens_out, seq, sizeof(int64_t) * (size_t)B * L, cudaMemcpyDeviceToDevice, st));
  StepState s{g.st[0], H, g.st[1], g.st[2], H, g.st[3], g.st[0], H, g.st[1], g.st[2], H, g.st[3]};
  for (int i = 0; i < Lp; ++i) {
    if (i == 0 || ss_prob <= 0.f) {      // step 0 always takes the ground truth (SAModel.py:89: i >= 1)
      XG_TRY(launch(h, "gather_rows", gather_rows_kernel, B, 128, 0, st, h->P[XG_P_EMBED_W], seq + i, L, 0, B, B, d.embed, d.vocab, g.step.XT));
    } else {
      XG_TRY(launch(h, "ss_pick", ss_pick_kernel, B, 256, 0, st, g.logits, d.vocab, i, L, ss_prob, ss_seed, seq, g.tok, tokens_out));
      XG_TRY(launch(h, "gather_rows", gather_rows_kernel, B, 128, 0, st, h->P[XG_P_EMBED_W], g.tok, 1, 0, B, B, d.embed, d.vocab, g.step.XT));
    }
    DropSpec drops[3];
    if (step_drop) {
      const uint64_t base = (uint64_t)i * B * H;
      drops[0] = make_drop(true, d.drop_prob, h->dec_drop_seed, XG_DROP_DEC_GATE, base);
      drops[1] = make_drop(true, d.drop_prob, h->dec_drop_seed, XG_DROP_DEC_H1, base);
      drops[2] = make_drop(true, d.drop_prob, h->dec_drop_seed, XG_DROP_DEC_H2, base);
    }
    XG_TRY(decode_step_core(h, g.step.XT, seq_mask + i, L, V, Uv, pos, s, g.step, nullptr, B, K, 1, st, step_drop ? drops : nullptr));
    if (i + 1 < Lp && ss_prob > 0.f) XG_TRY(logits_core(h, g.st[2], H, B, g.logits, st));
  }
  return XG_OK;
}

int xg_sample_beam(xg_handle h, const float* V, const float* feat_mask, const float* pos, int B, int K, int T, int beam,
                   int64_t* seq_out, float* logp_out, int64_t* done_seq, float* done_logps, float* done_p,
                   int32_t* done_count, void* ws, size_t ws_bytes, void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, V); CHECK_PTR(h, feat_mask); CHECK_PTR(h, pos); CHECK_PTR(h, seq_out); CHECK_PTR(h, logp_out); CHECK_PTR(h, ws);
  CHECK_POS(h, B); CHECK_POS(h, K); CHECK_POS(h, T); CHECK_POS(h, beam);
  if (beam > h->d.vocab) return fail(h, XG_ERR_BAD_SHAPE, "beam_size <= vocab_size required (SAModel.py:134)");
  if (beam > XG_MAX_BEAM) return fail(h, XG_ERR_UNSUPPORTED, "xg_sample_beam: beam_size > 16 not supported");
  XG_TRY(set_device(h));
  TcF16Scope f16(h);
  Arena a(ws, ws_bytes);
  BeamBufs w; carve_beam(a, h->d, B, K, T, beam, w);
  if (a.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_sample_beam: workspace too small");
  return beam_core(h, V, feat_mask, pos, B, K, T, beam, seq_out, logp_out, done_seq, done_logps, done_p, done_count, w,
                   (cudaStream_t)stream);
}

int xg_seq_steps(xg_handle h, const int64_t* seq, int B, int L, int* steps_out, void* stream) {
  CHECK_HANDLE(h);
  CHECK_PTR(h, seq); CHECK_PTR(h, steps_out); CHECK_POS(h, B); CHECK_POS(h, L);
  XG_TRY(set_device(h));
  cudaStream_t st = (cudaStream_t)stream;
  XG_TRY(launch(h, "seq_steps", seq_steps_kernel, 1, 128, 0, st, seq, B, L, h->d_small));
  XG_CUDA_TRY(h->es, cudaMemcpyAsync(h->h_pinned, h->d_small, sizeof(int), cudaMemcpyDeviceToHost, st));
  XG_CUDA_TRY(h->es, cudaStreamSynchronize(st));
  *steps_out = h->h_pinned[0];
  return XG_OK;
}

int xg_train_fwd(xg_handle h, const float* rgb, const float* opfl, const float* feat_mask, const float* pos,
                 const int64_t* seq, const float* seq_mask, int B, int K, int L, int Lp, int train, uint64_t seed,
                 float* logp, float* cat, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, rgb); CHECK_PTR(h, opfl); CHECK_PTR(h, feat_mask); CHECK_PTR(h, pos); CHECK_PTR(h, seq); CHECK_PTR(h, seq_mask);
  CHECK_PTR(h, saved);
  CHECK_POS(h, B); CHECK_POS(h, K); CHECK_POS(h, L); CHECK_POS(h, Lp);
  if (Lp > L) return fail(h, XG_ERR_BAD_SHAPE, "xg_train_fwd: Lp > L");
  if (!h->bn_bound) return fail(h, XG_ERR_NOT_BOUND, "BatchNorm buffers not bound (xg_bind_bn_buffers)");
  XG_TRY(set_device(h));
  TcF16Scope f16(h);
  Arena a(saved, saved_bytes);
  TrainSaved S; carve_saved(a, h->d, B, K, L, S);
  if (a.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_train_fwd: saved-activation block too small");
  (void)ws; (void)ws_bytes;
  return train_fwd_core(h, rgb, opfl, feat_mask, pos, seq, seq_mask, B, K, L, Lp, train, seed, logp, cat, S, nullptr, nullptr,
                        (cudaStream_t)stream);
}

int xg_train_bwd(xg_handle h, const float* rgb, const float* opfl, const float* feat_mask, const float* pos,
                 const int64_t* seq, const float* seq_mask, int B, int K, int L, int Lp, int train, uint64_t seed,
                 const float* logp, const float* cat, const float* dlogp, const float* dcat, const void* saved,
                 size_t saved_bytes, float* const* grads, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  CHECK_HANDLE(h); CHECK_BOUND(h);
  CHECK_PTR(h, rgb); CHECK_PTR(h, opfl); CHECK_PTR(h, feat_mask); CHECK_PTR(h, pos); CHECK_PTR(h, seq); CHECK_PTR(h, seq_mask);
  CHECK_PTR(h, saved); CHECK_PTR(h, grads); CHECK_PTR(h, ws);
  CHECK_POS(h, B); CHECK_POS(h, K); CHECK_POS(h, L); CHECK_POS(h, Lp);
  if (Lp > L) return fail(h, XG_ERR_BAD_SHAPE, "xg_train_bwd: Lp > L");
  if (dlogp) CHECK_PTR(h, logp);
  if (dcat) CHECK_PTR(h, cat);
  for (int i = 0; i < XG_NUM_PARAMS; ++i)
    if (!grads[i]) return fail(h, XG_ERR_NULL_POINTER, "xg_train_bwd: null gradient pointer");
  XG_TRY(set_device(h));
  Arena a(const_cast<void*>(saved), saved_bytes);
  TrainSaved S; carve_saved(a, h->d, B, K, L, S);
  if (a.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_train_bwd: saved-activation block too small");
  Arena w(ws, ws_bytes);
  BwdBufs W; carve_bwd(w, h->d, B, K, L, W);
  if (w.overflow) return fail(h, XG_ERR_WORKSPACE, "xg_train_bwd: workspace too small");
  return train_bwd_core(h, rgb, opfl, feat_mask, pos, seq, seq_mask, B, K, L, Lp, train, seed, logp, cat, dlogp, dcat, S, W,
                        grads, accumulate ? 1.f : 0.f, (cudaStream_t)stream);
}

int xg_nll_criterion_fwd(const float* logp, int N, const int64_t* target, const float* mask, const float* class_mask,
                         int ld, int rotate, int B, int Lp, float* loss_out, float* denom_out, float* scratch,
                         void* stream) {
  if (!logp || !target || !mask || !loss_out || !denom_out || !scratch)
    return fail(nullptr, XG_ERR_NULL_POINTER, "xg_nll_criterion_fwd: null pointer");
  if (N <= 0 || B <= 0 || Lp <= 0 || ld < Lp) return fail(nullptr, XG_ERR_BAD_SHAPE, "xg_nll_criterion_fwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = B * Lp;
  nll_terms_kernel<<<ceil_div(n, 128), 128, 0, st>>>(logp, N, target, mask, class_mask, ld, rotate, B, Lp, scratch);
  nll_reduce_kernel<<<1, 256, 0, st>>>(scratch, n, loss_out, denom_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return XG_ERR_CUDA; }
  return XG_OK;
}

int xg_nll_criterion_bwd(int N, const int64_t* target, const float* mask, const float* class_mask, int ld, int rotate,
                         int B, int Lp, const float* denom, const float* grad_out, float* dlogp, void* stream) {
  if (!target || !mask || !denom || !grad_out || !dlogp)
    return fail(nullptr, XG_ERR_NULL_POINTER, "xg_nll_criterion_bwd: null pointer");
  if (N <= 0 || B <= 0 || Lp <= 0 || ld < Lp) return fail(nullptr, XG_ERR_BAD_SHAPE, "xg_nll_criterion_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int n = B * Lp;
  cudaError_t e = cudaMemsetAsync(dlogp, 0, sizeof(float) * (size_t)n * N, st);
  if (e == cudaSuccess) {
    nll_grad_kernel<<<ceil_div(n, 128), 128, 0, st>>>(N, target, mask, class_mask, ld, rotate, B, Lp, denom, grad_out, dlogp);
    e = cudaGetLastError();
  }
  if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return XG_ERR_CUDA; }
  return XG_OK;
}

// ---- profiling ------------------------------------------------------------------------
int xg_profile_enable(xg_handle h, int on) {
  CHECK_HANDLE(h);
  h->prof_on = on != 0;
  return XG_OK;
}

int xg_profile_report(xg_handle h, char* buf, size_t buf_bytes) {
  CHECK_HANDLE(h);
  CHECK_PTR(h, buf);
  XG_TRY(set_device(h));
  XG_CUDA_TRY(h->es, cudaDeviceSynchronize());
  std::map<std::string, std::pair<long, double>> agg;
  for (auto& r : h->prof_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    auto& a = agg[r.name];
    a.first += 1;
    a.second += ms;
    h->prof_pool.push_back(r.e0);
    h->prof_pool.push_back(r.e1);
  }
  h->prof_recs.clear();
  std::string out = "[";
  bool first = true;
  for (auto& kv : agg) {
    char line[256];
    snprintf(line, sizeof(line), "%s{\"name\": \"%s\", \"launches\": %ld, \"ms\": %.6f}", first ? "" : ", ",
             kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
    first = false;
  }
  out += "]";
  if (out.size() + 1 > buf_bytes) return fail(h, XG_ERR_WORKSPACE, "xg_profile_report: buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return XG_OK;
}

// ---- diagnostics ---------------------------------------------------------------------
__global__ void dropout_mask_kernel(uint64_t seed, uint32_t site, size_t n, float p, float keep, float* out) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
    out[e] = p > 0.f ? drop_factor(seed, site, e, p, keep) : 1.f;
}

int xg_adam_step(const xg_adam_tensor* tensors, int count, int step, float lr, float beta1, float beta2, float eps,
                 float weight_decay, float grad_clip, int eps_mode, int write_clamped_grad, void* stream) {
  ErrorSink es;
  const int s = adam_step(es, tensors, count, step, lr, beta1, beta2, eps, weight_decay, grad_clip, eps_mode,
                          write_clamped_grad, (cudaStream_t)stream);
  if (s != XG_OK) g_last_error = es.msg;
  return s;
}

int xg_debug_dropout_mask(uint64_t seed, int site, size_t n, float p, float* out, void* stream) {
  if (!out) return fail(nullptr, XG_ERR_NULL_POINTER, "xg_debug_dropout_mask: null output");
  if (!(p >= 0.f && p < 1.f)) return fail(nullptr, XG_ERR_BAD_ARG, "xg_debug_dropout_mask: p must be in [0,1)");
  if (n == 0) return XG_OK;
  dropout_mask_kernel<<<ew_grid((long)n), 256, 0, (cudaStream_t)stream>>>(seed, (uint32_t)site, n, p, 1.f / (1.f - p), out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { g_last_error = cudaGetErrorString(e); return XG_ERR_CUDA; }
  return XG_OK;
}

int xg_debug_gemm(int layout, int engine, const float* A, const float* B, float* C, int M, int N, int K, void* stream) {
  if (!A || !B || !C) return fail(nullptr, XG_ERR_NULL_POINTER, "xg_debug_gemm: null pointer");
  if (M <= 0 || N <= 0 || K <= 0) return fail(nullptr, XG_ERR_BAD_SHAPE, "xg_debug_gemm: non-positive size");
  GemmP p;
  if (layout == 0) p = gemm_nt(A, K, B, K, C, N, M, N, K);
  else if (layout == 1) p = gemm_nn(A, K, B, N, C, N, M, N, K);
  else if (layout == 2) p = gemm_tn(A, M, B, N, C, N, M, N, K);
  else return fail(nullptr, XG_ERR_BAD_ARG, "xg_debug_gemm: layout must be 0 (NT), 1 (NN) or 2 (TN)");
  if (engine == 2 || engine == 4) {
    // tcgen05 engine on a parameter-less scratch context (operands are split on the fly): 2 = 3xTF32, 4 = 3xFP16 pairs
    static xg_context* dbg = nullptr;
    if (!dbg) {
      dbg = new xg_context();
      dbg->d = xg_dims{1, 1, 1, 1, 1, 1, 1, 1, XG_ACT_RELU, 0.f, 1e-5f, 0.1f};
      for (int i = 0; i < XG_NUM_PARAMS; ++i) dbg->P[i] = nullptr;
    }
    cudaGetDevice(&dbg->device);
    dbg->tc_f16 = engine == 4;
    int s = gemm_tc(dbg, p, (cudaStream_t)stream);
    if (s != XG_OK) g_last_error = dbg->es.msg;
    return s;
  }
  ErrorSink es;
  static SplitKScratch dbg_sk;     // engine 3: SIMT engine with split-K enabled
  if (engine == 3 && !dbg_sk.ws) {
    dbg_sk.ws_floats = (size_t)320 * 64 * 64; dbg_sk.ctr_count = 128;
    if (cudaMalloc(&dbg_sk.ws, sizeof(float) * dbg_sk.ws_floats) != cudaSuccess ||
        cudaMalloc(&dbg_sk.ctr, sizeof(unsigned int) * 2 * dbg_sk.ctr_count) != cudaSuccess ||
        cudaMemset(dbg_sk.ctr, 0, sizeof(unsigned int) * 2 * dbg_sk.ctr_count) != cudaSuccess)
      return fail(nullptr, XG_ERR_CUDA, "xg_debug_gemm: cannot allocate split-K scratch");
  }
  int s = gemm_simt(es, p, (cudaStream_t)stream, engine == 3 ? &dbg_sk : nullptr);
  if (s != XG_OK) g_last_error = es.msg;
  return s;
}

}  // extern "C"
