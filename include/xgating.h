/*
 * xgating.h — C ABI of the B200-native gated-fusion caption decoder (libxgating.so).
 *
 * This is the drop-in boundary for ONE path of vsislab/Controllable_XGating: the model
 * code in caption_src/SAModel.py + sub_modules.py + CaptionModel.py.  The reference has
 * no FFI of its own (it is pure Python on torch ops), so each entry point below names the
 * reference *function* (file:line under /root/reference/caption_src) whose arithmetic it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.  No torch types cross this boundary: plain device pointers, sizes, a stream.
 *
 * Conventions
 *   - every entry returns int: 0 = XG_OK, otherwise an xg_status code; no exceptions
 *     cross the boundary; xg_last_error() gives a human-readable message per handle.
 *   - all tensors are device pointers to row-major contiguous fp32 unless stated;
 *     token ids are int64.  The caller owns every buffer; the library allocates nothing
 *     the caller frees: scratch comes from a caller-provided workspace whose size is
 *     returned by xg_workspace_bytes().
 *   - `stream` is a cudaStream_t passed as void*; entries are asynchronous w.r.t. the
 *     host unless documented ("SYNC").  A handle is re-entrant across streams only if
 *     calls use distinct workspaces; it is not thread-safe.
 *   - sm_100a only.  There is no CPU fallback.
 *
 * Notation: B batch rows, K frames, H=rnn, E=embed, A=att, R=feat_rgb, F=feat_opfl,
 * V=vocab, C=categories, T=seq_length, L=T+1 teacher-forced steps.
 */
#ifndef XGATING_H_
#define XGATING_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XG_ABI_VERSION 1
#define XG_NUM_PARAMS 57 /* learnable tensors, reference state_dict order */

typedef struct xg_context* xg_handle;

typedef enum xg_status {
  XG_OK = 0,
  XG_ERR_BAD_ARG = 1,      /* reference: Python assert / TypeError */
  XG_ERR_BAD_SHAPE = 2,    /* reference: assert at sub_modules.py:69,673 ; SAModel.py:134 */
  XG_ERR_NULL_POINTER = 3,
  XG_ERR_CUDA = 4,         /* cudaGetLastError() captured in xg_last_error() */
  XG_ERR_NOT_BOUND = 5,    /* parameters / BN buffers not bound */
  XG_ERR_WORKSPACE = 6,    /* workspace too small */
  XG_ERR_UNSUPPORTED = 7
} xg_status;

/* fusion_activity (myopts.py:29, sub_modules.py:64) */
enum { XG_ACT_NONE = 0, XG_ACT_RELU = 1, XG_ACT_TANH = 2, XG_ACT_SIGMOID = 3 };

/* model dimensions = the opt fields SAModel.__init__ reads (SAModel.py:14-49) */
typedef struct xg_dims {
  int feat_rgb;    /* opt.feat_size            R (1536) */
  int feat_opfl;   /* opt.feat_size2           F (1024) */
  int rnn;         /* opt.rnn_size             H (512)  */
  int embed;       /* opt.input_encoding_size  E (468)  */
  int att;         /* opt.att_size             A (1536) */
  int vocab;       /* opt.vocab_size           V        */
  int categories;  /* opt.category_size        C        */
  int cls_hidden;  /* 128, SAModel.py:46 */
  int fusion_act;  /* XG_ACT_* */
  float drop_prob; /* opt.drop_prob_lm */
  float bn_eps;      /* 1e-5  (nn.BatchNorm1d default) */
  float bn_momentum; /* 0.1 */
} xg_dims;

/* Parameter table, in the reference's state_dict order (SAModel.py:14-49 registration
 * order).  Index i of the pointer tables passed to xg_bind_params / xg_train_bwd. */
typedef enum xg_param {
  XG_P_EMB_RGB_W = 0, XG_P_EMB_RGB_B, XG_P_BN_RGB_G, XG_P_BN_RGB_B,       /* two_spatial_encoder.visual_emb_rgb.{0,1}  */
  XG_P_EMB_OPFL_W, XG_P_EMB_OPFL_B, XG_P_BN_OPFL_G, XG_P_BN_OPFL_B,      /* ...visual_emb_opfl.{0,1}                  */
  XG_P_LSTM_RGB_WIH, XG_P_LSTM_RGB_WHH, XG_P_LSTM_RGB_BIH, XG_P_LSTM_RGB_BHH,   /* ...lstmcell_rgb (i,f,g,o)          */
  XG_P_LSTM_OPFL_WIH, XG_P_LSTM_OPFL_WHH, XG_P_LSTM_OPFL_BIH, XG_P_LSTM_OPFL_BHH,
  XG_P_GATE_RGB_W, XG_P_GATE_RGB_B, XG_P_GATE_OPFL_W, XG_P_GATE_OPFL_B,  /* ...gate_{rgb,opfl}.gate.0                 */
  XG_P_FUSION_W, XG_P_FUSION_B,                                           /* ...fusion.late_fusion.0                   */
  XG_P_INIT_H1_W, XG_P_INIT_H1_B, XG_P_INIT_C1_W, XG_P_INIT_C1_B,        /* img_embed_{h,c}_{1,2}                     */
  XG_P_INIT_H2_W, XG_P_INIT_H2_B, XG_P_INIT_C2_W, XG_P_INIT_C2_B,
  XG_P_DGATE_W, XG_P_DGATE_B,                                             /* lstmcore.gate.gate.0                      */
  XG_P_L1_I2H_W, XG_P_L1_I2H_B, XG_P_L1_A2H_W, XG_P_L1_A2H_B, XG_P_L1_H2H_W, XG_P_L1_H2H_B,   /* lstmcore.lstm_1 (i,f,o,g) */
  XG_P_L2_I2H_W, XG_P_L2_I2H_B, XG_P_L2_A2H_W, XG_P_L2_A2H_B, XG_P_L2_H2H_W, XG_P_L2_H2H_B,   /* lstmcore.lstm_2 */
  XG_P_V2A_W, XG_P_V2A_B, XG_P_H2A_W, XG_P_H2A_B, XG_P_A2W_W, XG_P_A2W_B, /* lstmcore.{v2a,h2a,a2w}                    */
  XG_P_EMBED_W,                                                           /* embed.weight                              */
  XG_P_LOGIT_W, XG_P_LOGIT_B,                                             /* logit                                     */
  XG_P_CLS0_W, XG_P_CLS0_B, XG_P_CLS3_W, XG_P_CLS3_B                      /* classifer.{0,3}                           */
} xg_param;

/* workspace kinds for xg_workspace_bytes() */
typedef enum xg_ws_kind {
  XG_WS_ENCODE = 0,      /* xg_encode_fwd             (B,K)        */
  XG_WS_DECODE_STEP = 1, /* xg_decode_step            (B,K)        */
  XG_WS_GREEDY = 2,      /* xg_sample_greedy          (B,K,T)      */
  XG_WS_BEAM = 3,        /* xg_sample_beam            (B,K,T,beam) */
  XG_WS_TRAIN_SAVED = 4, /* activations kept between xg_train_fwd and xg_train_bwd (B,K,L) */
  XG_WS_TRAIN_FWD = 5,   /* scratch of xg_train_fwd   (B,K,L)      */
  XG_WS_TRAIN_BWD = 6    /* scratch of xg_train_bwd   (B,K,L)      */
} xg_ws_kind;

/* dropout sites (logical mask layouts in DESIGN.md); for xg_debug_dropout_mask */
typedef enum xg_drop_site {
  XG_DROP_ENC_EMB_RGB = 1, XG_DROP_ENC_EMB_OPFL = 2,   /* (B,K,H) */
  XG_DROP_ENC_GATE_RGB = 3, XG_DROP_ENC_GATE_OPFL = 4, /* (K,B,H) */
  XG_DROP_ENC_FUSION = 5,                              /* (K,B,H) */
  XG_DROP_DEC_GATE = 6, XG_DROP_DEC_H1 = 7, XG_DROP_DEC_H2 = 8, /* (L,B,H) */
  XG_DROP_CLS = 9                                      /* (L,B,cls_hidden) */
} xg_drop_site;

/* ---- library / handle ------------------------------------------------------------- */
int xg_abi_version(void);
const char* xg_status_string(int status);
const char* xg_last_error(xg_handle h);

/* SAModel.__init__ (SAModel.py:14-50): fixes dimensions; device = CUDA ordinal. */
int xg_create(const xg_dims* dims, int device, xg_handle* out);
int xg_destroy(xg_handle h);
/* rows/cols of parameter `index` (cols = 1 for vectors): what load_state_dict(strict=True) checks. */
int xg_param_shape(xg_handle h, int index, int* rows, int* cols);
/* nn.Parameter storage is read in place: table of XG_NUM_PARAMS device pointers (xg_param order). */
int xg_bind_params(xg_handle h, const float* const* params, int count);
/* BatchNorm1d running_mean / running_var of both streams (sub_modules.py:98,102); updated in place
 * by xg_train_fwd / xg_encode_fwd when train != 0. */
int xg_bind_bn_buffers(xg_handle h, float* rm_rgb, float* rv_rgb, float* rm_opfl, float* rv_opfl);
/* call after the optimizer (or load_state_dict) rewrites parameter storage: marks the derived copies (tf32 splits,
 * POS-gate token table) stale.  Does not synchronise; the copies are rebuilt on the stream of the next call that needs
 * them, so the parameter update has to be ordered before that call on the same stream (or by an event). */
int xg_params_changed(xg_handle h);
/* mode 0: everything on the SIMT fp32 engine; 1: dense contractions above a size gate run on the
 * tcgen05 engine with two-term operands (fp32-grade accuracy: fp16 pairs in the forward entry points, whose operands -
 * features, activations, weights - must stay below 65504 in magnitude; tf32 pairs in xg_train_bwd; XG_NO_TC16=1 in
 * the environment keeps tf32 pairs everywhere); 2 (default): 1 + greedy decoding in the fused persistent
 * word-step kernel.  Results agree to ~1e-6 relative. */
int xg_set_engine(xg_handle h, int mode);
/* Fused-path policy.  Every serial loop of the path (encoder frame recurrence and its backward, greedy word loop,
 * teacher-forced word loop and its backward, beam-search word step) runs as ONE persistent kernel when the shape is
 * inside that kernel's limits: rnn_size a multiple of 32 and <= 512 (encoder: <= 4096), att_size a multiple of 4
 * (backward: of 32) and <= 1600, input_encoding_size a multiple of 4 and <= 640, frames <= 32, vocab < 32000,
 * at most 1024 caption rows (batch, or videos x beam) per call; the persistent backward loops at most 256 captions
 * per call.  Inside those limits the second-generation kernels (xg_grouped.cuh: fp16 operand pairs, LSTM cells fused
 * behind their products) take the greedy / sampling (multinomial draw, training dropout) / scheduled-sampling /
 * teacher-forced word loops up to 64 captions per call, the beam-search step up to 9 x 64 rows and the encoder recurrence when rnn_size is a
 * multiple of 64; larger calls run the first-generation persistent kernels (same results to ~1e-6).  Outside the
 * limits, and for multinomial sampling / the scheduled-sampling token pass above 64 captions, the same arithmetic
 * runs as per-step launches.
 * strict = 1 (or XG_STRICT_PERSIST=1 in the environment at xg_create) turns every such downgrade into
 * XG_ERR_UNSUPPORTED, unless mode < 2 was asked for with xg_set_engine. */
int xg_set_strict(xg_handle h, int strict);
/* how many serial loops this handle has run on a persistent kernel / on per-step launches so far */
int xg_path_counters(xg_handle h, uint64_t* fused, uint64_t* unfused);

size_t xg_workspace_bytes(xg_handle h, int kind, int B, int K, int L_or_T, int beam);

/* ---- encoder + init state ---------------------------------------------------------- */
/* EncoderLstm_two_fc.forward (sub_modules.py:118-159) + v2a(V) hoisted out of the word loop
 * (sub_modules.py:677) + SAModel.init_hidden (SAModel.py:58-65).
 *   rgb (B,K,R), opfl (B,K,F), feat_mask (B,K)
 *   V_out (B,K,H); Uv_out (B,K,A) or NULL; state_out[4] = h1,c1,h2,c2 each (B,H), or NULL.
 *   train != 0: BatchNorm batch statistics (+ running-stat update) and dropout with `seed`. */
int xg_encode_fwd(xg_handle h, const float* rgb, const float* opfl, const float* feat_mask,
                  int B, int K, int train, uint64_t seed,
                  float* V_out, float* Uv_out, float* const* state_out,
                  void* ws, size_t ws_bytes, void* stream);

/* SAModel.init_hidden alone (SAModel.py:58-65): feats (B,K,H) already fused. */
int xg_init_hidden(xg_handle h, const float* V, const float* feat_mask, int B, int K,
                   float* const* state_out, void* ws, size_t ws_bytes, void* stream);

/* v2a(V): Uv (B,K,A) = V (B,K,H) . W_v2a^T + b   (sub_modules.py:677, loop-invariant part) */
int xg_attend_precompute(xg_handle h, const float* V, int B, int K, float* Uv_out, void* stream);

/* ---- one word step ----------------------------------------------------------------- */
/* embed + LSTMCore_two_layer_gate.forward (sub_modules.py:671-687) [+ logit + log_softmax]
 * = SAModel.get_logprobs_state (SAModel.py:117-127) when xt_mask == NULL (mask of ones).
 *   tokens (B,) int64 OR xt (B,E) already embedded (exactly one non-NULL: lstmcore.forward takes xt,
 *   get_logprobs_state takes ids); xt_mask (B,) or NULL; V (B,K,H); Uv (B,K,A) or NULL (computed inside);
 *   pos (B,H); state_in[4] / state_out[4]: h1,c1,h2,c2 each (B,H) (may alias pairwise);
 *   out (B,H) or NULL; logp (B,V) or NULL.  Eval mode (no dropout). */
int xg_decode_step(xg_handle h, const int64_t* tokens, const float* xt, const float* xt_mask,
                   const float* V, const float* Uv, const float* pos,
                   const float* const* state_in, float* const* state_out,
                   float* out, float* logp, int B, int K,
                   void* ws, size_t ws_bytes, void* stream);

/* Training-mode sampling (the self-critical path, starttrain.py:131: model.sample() under model.train()):
 * while `on`, xg_sample_greedy applies the dropout of the training word step (POS gate, lstm_1 / lstm_2 hidden
 * states) with the Philox sites and indices xg_train_fwd uses for `seed`, so that a teacher-forced xg_train_fwd on
 * the sampled tokens with the same seed reproduces the sampled log-probs and provides their gradients.
 * For xg_train_fwd / xg_encode_fwd, bit 1 of `train` (train = 3) keeps training-mode statistics and dropout but
 * does not update the BatchNorm running statistics (second pass over the same batch). */
int xg_set_decode_dropout(xg_handle h, int on, uint64_t seed);

/* ---- decoding ---------------------------------------------------------------------- */
/* SAModel.sample, beam_size == 1 (SAModel.py:176-219), word loop only (encoder via xg_encode_fwd).
 *   sample_max != 0: greedy (torch.max, :186).  sample_max == 0: multinomial with `temperature`
 *   and a Philox stream seeded by `seed` (:189-196).
 *   seq_out (B,T) int64, logp_out (B,T); columns >= *steps_out are zero.
 *   steps_out (HOST int*): number of columns the reference would have returned (loop break at :206).
 *   SYNC: returns after the stream has been synchronised (steps_out is a host value).
 *   steps_out == NULL: ASYNCHRONOUS - the call returns once the work is queued on `stream` (a serving loop queues the
 *   next batch before it reads this one back); the number of columns is then the number of leading columns of seq_out
 *   with a non-zero entry (an unfinished caption always emits a non-zero id, SAModel.py:199-210). */
int xg_sample_greedy(xg_handle h, const float* V, const float* Uv, const float* pos,
                     const float* const* state0, int B, int K, int T,
                     int sample_max, float temperature, uint64_t seed,
                     int64_t* seq_out, float* logp_out, int* steps_out,
                     void* ws, size_t ws_bytes, void* stream);

/* Scheduled sampling, token pass (SAModel.py:89-99 inside SAModel.forward): walks the teacher-forced word loop
 * WITHOUT gradients and returns the input token of every step in tokens_out (B,L) int64: column 0 and, with
 * probability 1 - ss_prob per caption and step, column i are the ground truth seq[b,i]; otherwise the token is drawn
 * from the previous step's word distribution (Philox stream ss_seed).  State masks are seq_mask[:, i]; the training
 * dropout of the step is applied when xg_set_decode_dropout is on.  A teacher-forced xg_train_fwd on tokens_out with
 * the same dropout seed then reproduces the activations and provides log-probs with gradients.  Workspace:
 * XG_WS_GREEDY with T = L. */
int xg_scheduled_tokens(xg_handle h, const float* V, const float* Uv, const float* pos, const float* const* state0,
                        const int64_t* seq, const float* seq_mask, int B, int K, int L, int Lp, float ss_prob,
                        uint64_t ss_seed, int64_t* tokens_out, void* ws, size_t ws_bytes, void* stream);

/* SAModel.sample_beam + CaptionModel.beam_search (SAModel.py:129-161, CaptionModel.py:22-128),
 * all B videos x `beam` rows advanced together on the device.
 *   V (B,K,H) fused feats, feat_mask (B,K), pos (B,H)
 *   seq_out (B,T) int64 / logp_out (B,T): best finished beam per video (done_beams[k][0])
 *   done_seq (B,beam,T) int64, done_logps (B,beam,T), done_p (B,beam), done_count (B,) int32:
 *   the top-`beam` finished beams per video sorted by score (self.done_beams); any may be NULL. */
int xg_sample_beam(xg_handle h, const float* V, const float* feat_mask, const float* pos,
                   int B, int K, int T, int beam,
                   int64_t* seq_out, float* logp_out,
                   int64_t* done_seq, float* done_logps, float* done_p, int32_t* done_count,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- training ---------------------------------------------------------------------- */
/* number of teacher-forced steps SAModel.forward executes before its early exit
 * (`seq[:, i].sum() == 0`, SAModel.py:103).  SYNC (one 4-byte D2H). */
int xg_seq_steps(xg_handle h, const int64_t* seq, int B, int L, int* steps_out, void* stream);

/* SAModel.forward (SAModel.py:67-115), ss_prob == 0.
 *   seq (B,L) int64, seq_mask (B,L); Lp = steps from xg_seq_steps;
 *   logp (B,Lp,V), cat (B,Lp,C); `saved` keeps activations for xg_train_bwd (may be NULL when
 *   no backward will follow). */
int xg_train_fwd(xg_handle h, const float* rgb, const float* opfl, const float* feat_mask,
                 const float* pos, const int64_t* seq, const float* seq_mask,
                 int B, int K, int L, int Lp, int train, uint64_t seed,
                 float* logp, float* cat, void* saved, size_t saved_bytes,
                 void* ws, size_t ws_bytes, void* stream);

/* backward of xg_train_fwd (reference: torch autograd, starttrain.py:134).
 *   dlogp (B,Lp,V), dcat (B,Lp,C) (NULL = zero); logp/cat: the forward outputs.
 *   grads: XG_NUM_PARAMS device pointers (xg_param order); accumulate != 0 adds into them,
 *   otherwise they are overwritten.  No gradient flows to rgb/opfl/pos/masks; the init state
 *   path is detached from the encoder (SAModel.py:59-62). */
int xg_train_bwd(xg_handle h, const float* rgb, const float* opfl, const float* feat_mask,
                 const float* pos, const int64_t* seq, const float* seq_mask,
                 int B, int K, int L, int Lp, int train, uint64_t seed,
                 const float* logp, const float* cat, const float* dlogp, const float* dcat,
                 const void* saved, size_t saved_bytes, float* const* grads, int accumulate,
                 void* ws, size_t ws_bytes, void* stream);

/* Data-parallel overlap hook ("overlapped with backward in 2-3 buckets").  When a cudaEvent_t is set,
 * xg_train_bwd records it on its stream at the point where every decoder-side gradient (parameters XG_P_INIT_H1_W and
 * after: the contiguous tail of a flat gradient buffer laid out in xg_param order) is final and only the encoder's are
 * still being written, so that a caller can all-reduce the tail on another stream under the encoder backward.
 * NULL (default) disables it. */
int xg_set_bwd_split_event(xg_handle h, void* cuda_event);

/* LanguageModelCriterion (SAModel.py:225-234; rotate != 0: target rotated left by one) and
 * ClassiferCriterion (SAModel.py:241-253; rotate == 0, optional class_mask):
 *     loss = -sum_r logp[r, tgt(r)] * w(r) / sum_r w(r),   w = mask (* class_mask),  r = (b,i), i < Lp
 *   logp (B,Lp,N); target (B, ld) int64; mask / class_mask (B, ld) fp32 (class_mask may be NULL);
 *   loss_out (1,), denom_out (1,) device scalars; scratch: 2*B*Lp floats. */
int xg_nll_criterion_fwd(const float* logp, int N, const int64_t* target, const float* mask,
                         const float* class_mask, int ld, int rotate, int B, int Lp,
                         float* loss_out, float* denom_out, float* scratch, void* stream);
/* d(loss)/d(logp): dlogp (B,Lp,N) is overwritten (zeros except one entry per row);
 * grad_out: device scalar d(objective)/d(loss); denom: from the forward call. */
int xg_nll_criterion_bwd(int N, const int64_t* target, const float* mask, const float* class_mask,
                         int ld, int rotate, int B, int Lp, const float* denom,
                         const float* grad_out, float* dlogp, void* stream);

/* ---- profiling ------------------------------------------------------------------------- */
/* Per-kernel timing with CUDA events recorded on the launching stream around every launch the
 * handle makes (GEMMs are tagged by layout and shape).  Costs two event records per launch, so
 * throughput numbers are never taken with it enabled.  xg_profile_report is SYNC: it writes a JSON
 * array [{"name","launches","ms"}...] into buf and clears the records. */
int xg_profile_enable(xg_handle h, int on);
int xg_profile_report(xg_handle h, char* buf, size_t buf_bytes);

/* ---- optimizer step (the step right after the hot path: starttrain.py:134-137) --------- */
/* One launch over all parameter tensors: elementwise clamp of the gradient to +-grad_clip
 * (myutils.clip_gradient, myutils.py:79-85; grad_clip <= 0 disables; write_clamped_grad != 0 stores the
 * clamped value back like clamp_ does) followed by Adam (optim.Adam(model.parameters(), lr, weight_decay),
 * starttrain.py:76: L2 weight decay added to the gradient, no amsgrad).  eps_mode 0 = the PyTorch 0.3.1
 * formula the reference pins (denom = sqrt(v) + eps), 1 = PyTorch >= 1.0 (denom = sqrt(v)/sqrt(1-beta2^t) + eps).
 * step is 1-based.  Up to 64 tensors per call; exp_avg / exp_avg_sq are caller-owned fp32 state buffers.
 * Callers that go on using a handle bound to these parameters must call xg_params_changed(). */
typedef struct xg_adam_tensor {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t n;
} xg_adam_tensor;
int xg_adam_step(const xg_adam_tensor* tensors, int count, int step, float lr, float beta1, float beta2, float eps,
                 float weight_decay, float grad_clip, int eps_mode, int write_clamped_grad, void* stream);

/* ---- test / diagnostics hooks -------------------------------------------------------- */
/* the dropout mask (0 or 1/(1-p)) the kernels apply at `site` for logical element indices
 * [0,n) under `seed`, so a train-mode run can be replayed in the CPU oracle. */
int xg_debug_dropout_mask(uint64_t seed, int site, size_t n, float p, float* out, void* stream);

/* C (M,N) = A . B with the library's GEMM kernels; layout: 0 = NT (A (M,K), B (N,K)),
 * 1 = NN (A (M,K), B (K,N)), 2 = TN (A (K,M), B (K,N)).  engine: 0 = auto, 1 = SIMT fp32,
 * 2 = tcgen05 3xTF32 (XG_ERR_UNSUPPORTED if the shape is not eligible), 3 = SIMT fp32 with the
 * deterministic split-K path enabled for skinny shapes (what the handle-bound path uses), 4 = tcgen05 on fp16 operand
 * pairs (3xFP16: what the forward entry points use for their batched products; |operand| < 65504). */
int xg_debug_gemm(int layout, int engine, const float* A, const float* B, float* C,
                  int M, int N, int K, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XGATING_H_ */
