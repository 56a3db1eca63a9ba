/*
 * rn_b200.h -- C ABI of librn_b200.so: the B200 (sm_100a) Relation-Network hot path.
 *
 * The reference (mesnico/RelationNetworks-CLEVR) is pure Python/PyTorch and has no FFI of
 * its own; the interface each entry point replaces is a span of reference `model.py`,
 * cited per function below.  The host side (relationnetworks_clevr_b200/ops.py) binds these
 * with ctypes -- see INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - Every function returns 0 on success, a negative RN_ERR_* code otherwise; a human
 *     readable message for the calling thread is available from rn_last_error().
 *   - All pointers are DEVICE pointers unless the name starts with `h_`.  The caller owns all
 *     memory: inputs, outputs, `saved` (activations kept from forward for backward) and
 *     `scratch` (dead after the call returns its work to the stream).  The library never
 *     allocates or frees device memory and keeps no pointer after a call returns.
 *   - All tensors are contiguous, row-major, fp32 unless stated.  `stream` is a cudaStream_t
 *     passed as void*.  Calls are asynchronous on that stream, never synchronise the device,
 *     and are CUDA-graph capturable.
 *   - Buffers must be 16-byte aligned (256-byte for `saved` and `scratch`).
 */
#ifndef RN_B200_H_
#define RN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_ABI_VERSION 3

/* error codes */
#define RN_OK 0
#define RN_ERR_INVALID_ARG (-1)   /* bad shape / null pointer / misalignment */
#define RN_ERR_UNSUPPORTED (-2)   /* shape or mode not supported by the requested kernel family */
#define RN_ERR_CUDA (-3)          /* a CUDA runtime call or launch failed */
#define RN_ERR_ARCH (-4)          /* device is not compute capability 10.x */

/* arithmetic of the g-MLP layers 1..L-1 (layer 0 is always fp32) */
#define RN_PRECISION_FP32 0    /* fp32 SIMT kernels, any shape; on-device yardstick                 */
#define RN_PRECISION_PARITY 1  /* tcgen05, fp32-level: (A_hi + A_lo) x (W_hi + W_lo) fp16 splits, fp32 accum.
                                * Training forward: 3 MMA passes per K-step (A_hi W_hi + A_lo W_hi + A_hi W_lo, A_lo
                                * read from tensor memory) so the ReLU masks the backward uses agree with an fp32
                                * evaluation; eval forward: 2 passes (A_hi only, measured 2e-6..9e-5 on log-probs). */
#define RN_PRECISION_FAST 2    /* tcgen05: fp16 activations x fp16 weights, one pass, fp32 accum    */

/* rn_relation_cfg.flags (0 = defaults).  Forward and backward of one step must use the same flags. */
#define RN_REL_FLAG_FWD_2PASS 1u    /* PARITY training forward with fp16 activations only (the round-1 kernel) */
#define RN_REL_FLAG_DGRAD_2PASS 2u  /* data gradient with W_hi + W_lo instead of W_hi                          */

#define RN_MAX_G_LAYERS 8

/* Shape of one relation-layer call.  Mirrors RelationalLayer.__init__ (model.py:82-102) and
 * the sizes RelationalLayer.forward reads from its inputs (model.py:108-109). */
typedef struct rn_relation_cfg {
  int32_t B;          /* samples in this call                                              */
  int32_t n;          /* objects per sample (d*d cells, or 12 state-description rows)      */
  int32_t k;          /* features per object (26 from pixels incl. coords, 7 for SD)       */
  int32_t Q;          /* question embedding width (hyp["lstm_hidden"])                     */
  int32_t G;          /* width of every g layer (hyp["g_layers"], all equal)               */
  int32_t L;          /* number of g layers                                                */
  int32_t qinj;       /* hyp["question_injection_position"], 0 <= qinj < L                 */
  int32_t precision;  /* RN_PRECISION_*                                                    */
  int32_t training;   /* 1: keep what rn_relation_bwd needs in `saved`                     */
  uint32_t flags;     /* RN_REL_FLAG_*                                                     */
} rn_relation_cfg;

int rn_abi_version(void);
const char* rn_last_error(void);

/* Number of kernels this library has launched in the calling process (monotonic; for bench accounting). */
unsigned long long rn_launch_count(void);
/* sizeof of the seven structs of this header as the library was compiled, in declaration order (rn_relation_cfg, rn_f_cfg,
 * rn_conv_cfg, rn_conv_layer, rn_conv_grads, rn_lstm_cfg, rn_adam_cfg): a binding checks its own mirrors against them once at
 * load time (a stale mirror would otherwise be read as garbage fields).  Returns the number of entries written (<= n). */
int rn_abi_struct_sizes(int32_t* out, int n);

/* Returns RN_OK when `device` is a compute-capability 10.x GPU, RN_ERR_ARCH otherwise. */
int rn_device_check(int device);

/* 1 when the tcgen05 kernel family supports this shape (G == 256, n*n % 128 == 0, L == 4 ...). */
int rn_relation_tc_supported(const rn_relation_cfg* cfg);

/* Bytes of `saved` and `scratch` the forward/backward pair needs for this cfg. */
int rn_relation_workspace(const rn_relation_cfg* cfg, size_t* saved_bytes, size_t* scratch_bytes);

/*
 * g-MLP over all n*n ordered pairs + pair-sum.  Replaces RelationalLayer.forward up to and
 * including the `.sum(1)` (model.py:104-152):
 *     row p = a*n + c of sample b is [x[b,c] | x[b,a] | q[b] at layer qinj]
 *     x_g[b] = sum_p relu(W_{L-1} ... relu(W_0 row + b_0) ... + b_{L-1})
 *   x    [B, n, k]      q  [B, Q]
 *   g_w  h_ array of L device pointers, g_w[l] is [G, fan_in_l] exactly as
 *        rl.g_layers.l.weight (fan_in_0 = 2k, +Q at l == qinj; fan_in_l = G otherwise)
 *   g_b  h_ array of L device pointers to [G]
 *   xg   [B, G] out
 */
int rn_relation_fwd(const rn_relation_cfg* cfg, const float* x, const float* q,
                    const float* const* h_g_w, const float* const* h_g_b, float* xg,
                    void* saved, void* scratch, void* stream);

/*
 * Backward of rn_relation_fwd (what autograd derives for model.py:104-152).
 *   dxg [B, G] in;  dx [B, n, k], dq [B, Q] out (overwritten)
 *   dg_w / dg_b: h_ arrays of L device pointers, same shapes as g_w / g_b, OVERWRITTEN with the
 *   gradient summed over the batch.
 */
int rn_relation_bwd(const rn_relation_cfg* cfg, const float* dxg, const float* x, const float* q,
                    const float* const* h_g_w, const void* saved, float* dx, float* dq,
                    float* const* h_dg_w, float* const* h_dg_b, void* scratch, void* stream);

/* ---- feature extraction (extract.py:63-74) ---- */
/* maxf / avgf [B, W]: max and mean over the P rows of each sample of z / max(||z||_2, 1e-12), z [B * P, ld] fp32 using the
 * first W columns (W <= 512).  scratch: floats, B * 32 * 2 * W. */
int rn_extract_stats(const float* z, int B, long long P, int ld, int W, float* maxf, float* avgf, float* scratch,
                     void* stream);
/* *out = device pointer of H_{l+1} = relu(g layer l) [B * n * n, G] inside the `saved` buffer of an RN_PRECISION_FP32
 * training-mode rn_relation_fwd: the tensor a forward hook on rl.g_layers[l + 1] receives as its input (extract.py:43-47). */
int rn_relation_activation(const rn_relation_cfg* cfg, const void* saved, int l, const float** out);

/* ---- f-MLP head: fc1 -> ReLU -> fc2 -> Dropout -> ReLU -> fc3 -> log_softmax (model.py:155-162) ---- */
typedef struct rn_f_cfg {
  int32_t B;        /* samples */
  int32_t G;        /* input width  (g_layers[-1]) */
  int32_t F1;       /* hyp["f_fc1"] */
  int32_t F2;       /* hyp["f_fc2"] */
  int32_t A;        /* answers (adict_size) */
  float keep_scale; /* 1/(1-p) applied where drop_mask != 0; ignored when drop_mask == NULL */
} rn_f_cfg;

/* saved: floats, B*(F1 + F2) (h1 and the post-dropout post-ReLU h2).  drop_mask: [B,F2] uint8 (0 = dropped)
 * or NULL (eval / p == 0); the mask comes from the caller so the RNG stays the framework's (torch). */
int rn_f_fwd(const rn_f_cfg* cfg, const float* xg, const float* w1, const float* b1, const float* w2,
             const float* b2, const float* w3, const float* b3, const uint8_t* drop_mask, float* logp,
             float* saved, void* stream);

/* dlogp [B,A] in.  Outputs overwritten.  scratch: floats, B*(A + F2 + F1).  drop_mask only tells whether
 * dropout was active (its contents are implied by saved h2 > 0). */
int rn_f_bwd(const rn_f_cfg* cfg, const float* dlogp, const float* logp, const float* xg, const float* w1,
             const float* w2, const float* w3, const uint8_t* drop_mask, const float* saved, float* dxg,
             float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3, float* scratch,
             void* stream);

/* ---- conv feature extractor: 4 x [conv3x3 s2 p1 -> BatchNorm -> ReLU] + coords (model.py:22-36,192-201) ---- */
#define RN_CONV_LAYERS 4
#define RN_CONV_CH 24

typedef struct rn_conv_cfg {
  int32_t B;         /* images */
  int32_t side;      /* input height == width, multiple of 16 */
  int32_t training;  /* 1: batch statistics (+ running-stat update), 0: running statistics */
  float eps;         /* BatchNorm eps (1e-5) */
  float momentum;    /* BatchNorm momentum (0.1) */
  int32_t img_u8;    /* 1: `img` is uint8 [B,3,S,S] (raw pixels); the first layer computes x = u / 255 (torchvision ToTensor,
                      * train.py:182-188) while staging -- a quarter of the host-to-device bytes.  0: fp32 in [0,1] */
  int32_t flags;     /* RN_CONV_FLAG_*; 0 = default kernels */
} rn_conv_cfg;

/* side % 64 == 0: the BACKWARD convolutions (weight and data gradients) run on the tensor cores (mma.sync, TF32 operands,
 * error-compensated 3-pass split: 2^-21 relative per product).  The FORWARD convolutions stay on the fp32 SIMT kernels by
 * default: their rounding decides the ReLU masks of the whole stack, and a 2^-21 evaluation flips ~10x more near-zero
 * pre-activations than a 2^-24 one (measured on the trained checkpoint: conv2.weight gradient 6e-3 instead of < 1e-3).
 * RN_CONV_FLAG_SIMT forces the SIMT kernels everywhere (they also serve every other side); RN_CONV_FLAG_TC_FWD opts the
 * forward into the tensor-core kernels (20 % faster forward, reduced mask fidelity). */
#define RN_CONV_FLAG_SIMT 1
#define RN_CONV_FLAG_TC_FWD 2

/* Per-layer parameter block, host array of RN_CONV_LAYERS entries (device pointers inside). */
typedef struct rn_conv_layer {
  const float* w;        /* [24, cin, 3, 3], cin = 3 for layer 0 else 24 */
  const float* bias;     /* [24] */
  const float* gamma;    /* [24] BatchNorm weight */
  const float* beta;     /* [24] BatchNorm bias */
  float* running_mean;   /* [24] read in eval; updated in place in training */
  float* running_var;    /* [24] */
} rn_conv_layer;

/* saved floats: raw conv outputs of all layers + 2*24 batch stats per layer; see rn_conv_workspace.
 * objects out: [B, d*d, 26] with d = side/16 -- channels 0..23 the layer-4 activations, 24/25 the
 * x/y coordinates linspace(-d/2, d/2, d) (model.py:208-213). */
int rn_conv_workspace(const rn_conv_cfg* cfg, size_t* saved_floats, size_t* scratch_floats);
int rn_conv_fwd(const rn_conv_cfg* cfg, const void* img, const rn_conv_layer* h_layers, float* objects,
                float* saved, float* scratch, void* stream);

typedef struct rn_conv_grads {
  float* dw;      /* [24, cin, 3, 3] */
  float* dbias;   /* [24] */
  float* dgamma;  /* [24] */
  float* dbeta;   /* [24] */
} rn_conv_grads;

/* dobjects [B, d*d, 26] in (coordinate columns ignored).  Gradients overwritten.  No image gradient
 * (utils.py:135,143: images never require grad). */
int rn_conv_bwd(const rn_conv_cfg* cfg, const void* img, const float* dobjects, const rn_conv_layer* h_layers,
                const float* saved, const rn_conv_grads* h_grads, float* scratch, void* stream);

/* ---- question encoder: Embedding -> 1-layer LSTM (zero initial state) -> last hidden state (model.py:39-58) ---- */
typedef struct rn_lstm_cfg {
  int32_t B;         /* questions */
  int32_t T;         /* tokens per question (<= 64) */
  int32_t V;         /* embedding rows (qdict_size + 1); token ids are clamped to [0, V) */
  int32_t E;         /* embedding width (hyp["lstm_word_emb"]) */
  int32_t H;         /* hidden size (hyp["lstm_hidden"]); the kernels exist for H == 128 */
  int32_t training;  /* 1: keep what rn_lstm_bwd needs in `saved` */
} rn_lstm_cfg;

/* 1 when the kernels support this shape (H == 128, T <= 64); the host keeps torch's nn.LSTM otherwise. */
int rn_lstm_supported(const rn_lstm_cfg* cfg);
int rn_lstm_workspace(const rn_lstm_cfg* cfg, size_t* saved_floats, size_t* scratch_floats);

/* tokens [B, T] int64 (as the reference feeds nn.Embedding); emb [V, E]; w_ih [4H, E], w_hh [4H, H], b_ih / b_hh [4H]
 * exactly as text.lstm.{weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0} (gate order i, f, g, o).  q [B, H] out. */
int rn_lstm_fwd(const rn_lstm_cfg* cfg, const int64_t* tokens, const float* emb, const float* w_ih, const float* w_hh,
                const float* b_ih, const float* b_hh, float* q, float* saved, void* stream);

/* dq [B, H] in; all five gradients are OVERWRITTEN (embedding rows of unused tokens receive zeros). */
int rn_lstm_bwd(const rn_lstm_cfg* cfg, const int64_t* tokens, const float* emb, const float* w_ih, const float* w_hh,
                const float* dq, const float* saved, float* demb, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                float* scratch, void* stream);

/* ---- optimiser tail: clip_grad_norm + Adam with L2 weight decay (train.py:45-48,330) on flat buffers ---- */
typedef struct rn_adam_cfg {
  int64_t n;          /* elements in the flat parameter / gradient buffers */
  float lr, beta1, beta2, eps, weight_decay;
  float clip_norm;    /* <= 0 disables clipping */
  float grad_scale;   /* gradients are multiplied by this first (1/world_size after an allreduce-sum) */
  int32_t step;       /* 1-based Adam step */
} rn_adam_cfg;

/* norm_scratch: >= 1024 floats + 1.  total_norm_out (device, 1 float) receives the pre-clip norm.
 * d_step / d_lr (device, optional): when non-NULL the kernels increment *d_step and use it as the Adam step, and read the
 * learning rate from *d_lr, instead of cfg->step / cfg->lr -- so a CUDA graph captured once replays correct steps. */
int rn_clip_adam(const rn_adam_cfg* cfg, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                 float* norm_scratch, float* total_norm_out, int32_t* d_step, const float* d_lr, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RN_B200_H_ */
