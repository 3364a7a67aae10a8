/*
 * mamba_tts_b200.h -- C ABI of the B200 (sm_100a) MambaTTSDecoder hot-path library.
 *
 * This is the drop-in boundary.  The reference (whcorkran/mamba-TTS-project) reaches all of its SSM
 * arithmetic through ONE third-party Python call, `Mamba(d_model)` (mamba_decoder.py:4,29,61,63).
 * Underneath, `mamba_ssm` binds two pybind11 extensions that take `at::Tensor`s
 * (`selective_scan_cuda.{fwd,bwd}`, `causal_conv1d_cuda.{causal_conv1d_fwd,causal_conv1d_bwd,
 * causal_conv1d_update}`) plus a Triton kernel (`selective_state_update`).  Every entry point
 * below replaces one of those bindings with a plain-C call: POD parameter struct of device
 * pointers / sizes / element strides + a CUDA stream.  No torch types cross this line.
 *
 * Rules:
 *   - the caller owns every buffer (inputs, outputs, saved-for-backward, workspaces); the library
 *     never allocates, frees or retains a pointer past the call;
 *   - work is enqueued on `stream`, asynchronously, with no host synchronisation and no allocation,
 *     so every call is CUDA-graph capturable;
 *   - return value 0 = success; >0 = MTTS_ERR_* argument error (nothing was launched);
 *     <0 = -(cudaError_t) reported by the launch.  `mtts_error_string` explains either;
 *   - in-place mutation happens only where stated (conv_state / ssm_state / accumulate-into grads);
 *   - there is NO CPU path: without a CUDA device every compute call fails.
 *
 * Layout conventions (upstream's): activations are channel-major `(batch, dim, seqlen)` with unit
 * stride along seqlen; B/C are `(batch, dstate, seqlen)`; A is `(dim, dstate)` fp32 = -exp(A_log);
 * D, delta_bias, conv weight `(dim, width)` and conv bias are fp32.  "io dtype" is the element type
 * of the activation tensors of one call (all the same): MTTS_F32 or MTTS_BF16.
 * All strides are in ELEMENTS.
 */
#ifndef MAMBA_TTS_B200_H_
#define MAMBA_TTS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mtts_stream_t; /* a cudaStream_t */

enum { MTTS_F32 = 0, MTTS_BF16 = 1 };

enum {
  MTTS_OK = 0,
  MTTS_ERR_NULL = 1,      /* a required pointer is NULL                         */
  MTTS_ERR_SHAPE = 2,     /* a dimension is out of the supported range          */
  MTTS_ERR_DTYPE = 3,     /* unknown io dtype                                   */
  MTTS_ERR_ALIGN = 4,     /* pointer/stride alignment the kernel cannot handle  */
  MTTS_ERR_UNSUPPORTED = 5
};

/* Timesteps between two saved scan states ("checkpoints"): selective_scan_fwd writes the state at
 * the START of every chunk of this many timesteps; selective_scan_bwd recomputes inside a chunk. */
#define MTTS_SCAN_CHUNK 32
#define MTTS_MAX_DSTATE 256
#define MTTS_MAX_CONV_WIDTH 4

const char* mtts_error_string(int code);
int mtts_abi_version(void);
/* Compiled-for architecture, e.g. 100 for sm_100a. */
int mtts_target_sm(void);
/* Which selective-scan kernel family serves a shape: 0 = the library's choice (default), 1 = time-sequential,
 * 2 = time-parallel.  Process-wide; a test / measurement hook (both families are parity-tested on the same inputs). */
int mtts_set_scan_impl(int impl);
/* sizeof() of the n-th parameter struct below (declaration order, from 0); -1 past the end.
 * Lets a foreign binding verify its mirror of the layout when it loads the library. */
int mtts_sizeof_params(int which);

/* ---------------------------------------------------------------------------------------------
 * causal_conv1d_fwd  -- replaces causal_conv1d_cuda.causal_conv1d_fwd (called by
 * mamba_ssm Mamba.forward, reached from mamba_decoder.py:61).
 *   out[b,d,t] = act( bias[d] + sum_{k<width} weight[d,k] * x[b,d,t-(width-1)+k] ),  x[t<0] taken
 *   from initial_states[b,d,(width-1)+t] or 0.  act = SiLU when silu != 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, dim, seqlen, width; /* width in 2..4 */
  int32_t io_dtype;
  int32_t silu;
  const void* x;
  int64_t x_batch_stride, x_dim_stride;
  const float* weight; /* (dim, width) contiguous */
  const float* bias;   /* (dim) or NULL */
  const void* initial_states; /* (batch, dim, width-1) io dtype, or NULL */
  int64_t init_batch_stride, init_dim_stride;
  void* out;
  int64_t out_batch_stride, out_dim_stride;
} mtts_conv1d_fwd_params;
int mtts_causal_conv1d_fwd(const mtts_conv1d_fwd_params* p, mtts_stream_t stream);

/* causal_conv1d_bwd -- replaces causal_conv1d_cuda.causal_conv1d_bwd.
 * dweight (dim,width) and dbias (dim) are fp32 and ACCUMULATED INTO (caller zero-fills). */
typedef struct {
  int32_t batch, dim, seqlen, width;
  int32_t io_dtype;
  int32_t silu;
  const void* x;
  int64_t x_batch_stride, x_dim_stride;
  const float* weight;
  const float* bias;
  const void* initial_states;
  int64_t init_batch_stride, init_dim_stride;
  const void* dout;
  int64_t dout_batch_stride, dout_dim_stride;
  void* dx;
  int64_t dx_batch_stride, dx_dim_stride;
  float* dweight;
  float* dbias; /* may be NULL */
} mtts_conv1d_bwd_params;
int mtts_causal_conv1d_bwd(const mtts_conv1d_bwd_params* p, mtts_stream_t stream);

/* causal_conv1d_update -- replaces causal_conv1d_cuda.causal_conv1d_update (Mamba.step).
 * conv_state (batch, dim, width) io dtype is rolled left by one IN PLACE, x written last. */
typedef struct {
  int32_t batch, dim, width;
  int32_t io_dtype;
  int32_t silu;
  const void* x; /* (batch, dim) */
  int64_t x_batch_stride;
  void* conv_state; /* contiguous (batch, dim, width) */
  const float* weight;
  const float* bias;
  void* out; /* (batch, dim) */
  int64_t out_batch_stride;
} mtts_conv1d_update_params;
int mtts_causal_conv1d_update(const mtts_conv1d_update_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * selective_scan_fwd -- replaces selective_scan_cuda.fwd (mamba_ssm selective_scan_fn /
 * mamba_inner_fn, reached from mamba_decoder.py:61).
 *   dl = delta + delta_bias; dl = softplus(dl) if delta_softplus
 *   h_t = exp(dl_t * A) * h_{t-1} + dl_t * B_t * u_t ;  y_t = <C_t, h_t> + D * u_t
 *   out_t = y_t * silu(z_t)   (z optional)
 * checkpoints (batch, dim, nchunks, dstate) fp32, nchunks = ceil(seqlen / MTTS_SCAN_CHUNK):
 *   state at the start of each chunk (chunk 0 = initial state); NULL when no backward follows.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, dim, seqlen, dstate;
  int32_t io_dtype;
  int32_t delta_softplus;
  const void* u;
  int64_t u_batch_stride, u_dim_stride;
  const void* delta;
  int64_t delta_batch_stride, delta_dim_stride;
  const float* A; /* (dim, dstate) contiguous */
  const void* B;
  int64_t B_batch_stride, B_state_stride;
  const void* C;
  int64_t C_batch_stride, C_state_stride;
  const float* D;          /* (dim) or NULL */
  const float* delta_bias; /* (dim) or NULL */
  const void* z;           /* or NULL */
  int64_t z_batch_stride, z_dim_stride;
  const float* initial_state; /* (batch, dim, dstate) contiguous or NULL */
  void* out;
  int64_t out_batch_stride, out_dim_stride;
  float* last_state;  /* (batch, dim, dstate) contiguous or NULL */
  float* checkpoints; /* see above, or NULL */
  /* optional: y = <C, h> + D u BEFORE the silu(z) gate, io dtype (batch, dim, seqlen).  Upstream saves the
   * same tensor (`out`, next to `out_z`) for its backward; handing it to selective_scan_bwd spares the
   * backward the <C, h> recompute it otherwise needs for dz. */
  void* y_pre;
  int64_t y_batch_stride, y_dim_stride;
} mtts_scan_fwd_params;
int mtts_selective_scan_fwd(const mtts_scan_fwd_params* p, mtts_stream_t stream);

/* selective_scan_bwd -- replaces selective_scan_cuda.bwd.  Recomputes the states inside each chunk
 * from `checkpoints`.  dB, dC (batch, dstate, seqlen) fp32 contiguous, dA (dim, dstate), dD (dim),
 * ddelta_bias (dim) fp32 are ACCUMULATED INTO (caller zero-fills).  du, ddelta, dz are io dtype. */
typedef struct {
  int32_t batch, dim, seqlen, dstate;
  int32_t io_dtype;
  int32_t delta_softplus;
  const void* u;
  int64_t u_batch_stride, u_dim_stride;
  const void* delta;
  int64_t delta_batch_stride, delta_dim_stride;
  const float* A;
  const void* B;
  int64_t B_batch_stride, B_state_stride;
  const void* C;
  int64_t C_batch_stride, C_state_stride;
  const float* D;
  const float* delta_bias;
  const void* z;
  int64_t z_batch_stride, z_dim_stride;
  const void* dout;
  int64_t dout_batch_stride, dout_dim_stride;
  const float* checkpoints;
  void* du;
  int64_t du_batch_stride, du_dim_stride;
  void* ddelta;
  int64_t ddelta_batch_stride, ddelta_dim_stride;
  void* dz; /* required iff z != NULL */
  int64_t dz_batch_stride, dz_dim_stride;
  float* dA;
  float* dB;
  float* dC;
  float* dD;          /* required iff D != NULL */
  float* ddelta_bias; /* required iff delta_bias != NULL */
  const void* y_pre;  /* the forward's y_pre, or NULL (then y is recomputed) */
  int64_t y_batch_stride, y_dim_stride;
} mtts_scan_bwd_params;
int mtts_selective_scan_bwd(const mtts_scan_bwd_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * selective_state_update -- replaces mamba_ssm.ops.triton.selective_state_update (Mamba.step).
 * state (batch, dim, dstate) fp32 contiguous, updated IN PLACE.  x, dt, z, out (batch, dim);
 * B, C (batch, dstate): io dtype, rows contiguous.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, dim, dstate;
  int32_t io_dtype;
  int32_t dt_softplus;
  float* state;
  const void* x;
  int64_t x_batch_stride;
  const void* dt;
  int64_t dt_batch_stride;
  const float* A;
  const void* B;
  int64_t B_batch_stride;
  const void* C;
  int64_t C_batch_stride;
  const float* D;       /* or NULL */
  const void* z;        /* or NULL */
  int64_t z_batch_stride;
  const float* dt_bias; /* or NULL */
  void* out;
  int64_t out_batch_stride;
} mtts_state_update_params;
int mtts_selective_state_update(const mtts_state_update_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * mamba_decode_step -- the whole inner part of Mamba.step in ONE launch (mamba_decoder.py:63 ->
 * Mamba.step: causal_conv1d_update -> x_proj -> dt_proj -> selective_state_update -> gate):
 *   x, z = xz[:, :dim], xz[:, dim:]
 *   conv_state <- roll(conv_state); conv_state[..., -1] = x ; xc = silu(<conv_state, w> + b)
 *   (dt_low, B, C) = x_proj_w @ xc ;  dt = softplus(dt_proj_w @ dt_low + dt_bias)
 *   ssm_state <- ssm_state * exp(dt A) + dt B xc ;  y = <ssm_state, C> + D xc ;  y *= silu(z)
 * xz, y, conv_state, x_proj_w (dt_rank + 2 dstate, dim), dt_proj_w (dim, dt_rank) are io dtype and
 * contiguous; ssm_state fp32.  Both states are updated IN PLACE.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, dim, dstate, dt_rank, width;
  int32_t io_dtype;
  const void* xz; /* (batch, 2*dim) */
  int64_t xz_batch_stride;
  void* conv_state; /* (batch, dim, width) */
  float* ssm_state; /* (batch, dim, dstate) */
  const float* conv_weight; /* (dim, width) */
  const float* conv_bias;   /* (dim) or NULL */
  const void* x_proj_w;
  const void* dt_proj_w;
  const float* dt_bias; /* (dim) */
  const float* A;       /* (dim, dstate) */
  const float* D;       /* (dim) */
  void* y;              /* (batch, dim) */
  int64_t y_batch_stride;
  /* Optional L2 prefetch for the kernels that follow in the step: prefetch_bytes bytes per batch element
   * starting at prefetch_a + b * prefetch_bytes (and likewise prefetch_b) are pulled from HBM into L2 once the
   * kernel's own loads are done -- the decoder passes the layer's cached K and V, which the cross-attention
   * two launches later then finds in L2 (this kernel itself moves 10 MB and leaves HBM idle).  NULL = none. */
  const void* prefetch_a;
  const void* prefetch_b;
  int64_t prefetch_bytes;
} mtts_decode_step_params;
int mtts_mamba_decode_step(const mtts_decode_step_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * cross_attn_decode -- single-query attention of decode_step against the cached K/V of
 * [ref || text] (mamba_decoder.py:72-77 with T_q = 1; K/V projected once per generation, D9).
 *   q (batch, heads*head_dim) already projected (bias included, NOT yet scaled);
 *   k, v (batch, t_kv, heads*head_dim) contiguous; mask (batch, t_kv) uint8, 1 = attend, or NULL;
 *   out (batch, heads*head_dim) = softmax(q k^T / sqrt(head_dim) + mask) v, fp32 softmax.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, heads, head_dim, t_kv;
  int32_t io_dtype;
  const void* q;
  const void* k;
  const void* v;
  const uint8_t* mask;
  void* out;
} mtts_cross_attn_decode_params;
int mtts_cross_attn_decode(const mtts_cross_attn_decode_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * cross_attn_block_decode -- the whole cross-attention branch of decode_step for one new token in ONE
 * launch (mamba_decoder.py:67-81 with T_q = 1; what the eager path runs as add_layernorm + q Linear +
 * cross_attn_decode + out_proj Linear + add_layernorm):
 *   x1 = x + delta;  hq = LN(x1; lnq);  q = wq hq + bq;  a = softmax(q k^T / sqrt(head_dim) + mask) v;
 *   o = wo a + bo;   x_out = x1 + o;    out = film_gamma * LN(x_out; lno) + film_beta
 *   x, x_out (batch, d) fp32 residual stream (x_out may alias x only when wo is given: the cluster
 *   synchronises before the write; the front-half variant's CTAs are independent); delta (batch, d) io dtype or NULL;
 *   wq, wo (d, d) row-major [out][in] and bq, bo (d) in the io dtype (nn.MultiheadAttention's
 *   in_proj_weight[0:d] / out_proj); k, v (batch, t_kv, d); mask as in cross_attn_decode;
 *   LN weights / biases and film_gamma / film_beta (batch, d) fp32 (FiLM optional); out (batch, d) io dtype.
 * wo == NULL selects the front half only: x_out = x1, out = a (the caller runs the out projection and
 * the second add_layernorm itself); bo / lno / film are then ignored.
 * One thread-block cluster of `heads` CTAs per batch element.  Supported: heads = 8, head_dim = 64,
 * t_kv <= 256; anything else returns MTTS_ERR_UNSUPPORTED and the caller uses the separate ops.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, heads, head_dim, t_kv;
  int32_t io_dtype;
  float eps_q, eps_o;
  const float* x;
  const void* delta; /* or NULL */
  float* x_out;      /* or NULL */
  const float* lnq_weight;
  const float* lnq_bias;
  const void* wq;
  const void* bq;
  const void* k;
  const void* v;
  const uint8_t* mask; /* or NULL */
  const void* wo;
  const void* bo;
  const float* lno_weight;
  const float* lno_bias;
  const float* film_gamma; /* both or neither */
  const float* film_beta;
  void* out;
} mtts_cross_attn_block_params;
int mtts_cross_attn_block_decode(const mtts_cross_attn_block_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * decode_embed / decode_greedy -- the token plumbing around one decode step of a greedy generation loop
 * (the missing caller of decode_step; embedding sum of mamba_decoder.py:218-221, argmax over
 * the logits of :254-256), with every counter resident on the device so a step is CUDA-graph replayable:
 *   embed:  x[b, :] = tok_embed[tok[b], :] + pos_embed[*pos, :]  (fp32);  then *step += 1
 *   greedy: tok[b] = argmax_v logits[b, v] (lowest index on ties); out[b, *step] = tok[b];  then *pos += 1
 *           (rows that already produced eos_id get pad_id instead, see the struct)
 * tok (batch) int64 ids (caller guarantees 0 <= id < vocab rows of tok_embed); pos, step: device int64
 * scalars (step = column of `out` being generated, -1 before the first embed).  dim % 4 == 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, dim;
  const int64_t* tok;
  const int64_t* pos;
  const float* tok_embed;
  const float* pos_embed;
  float* x;      /* (batch, dim) */
  int64_t* step; /* incremented after the gather, or NULL */
} mtts_decode_embed_params;
int mtts_decode_embed(const mtts_decode_embed_params* p, mtts_stream_t stream);

typedef struct {
  int32_t batch, vocab;
  int32_t io_dtype;
  int32_t reserved;
  const void* logits; /* (batch, vocab) contiguous, io dtype */
  int64_t* tok;       /* (batch) */
  int64_t* out;       /* (batch, out_stride) or NULL */
  int64_t out_stride;
  const int64_t* step;
  int64_t* pos;       /* incremented after the argmax, or NULL */
  /* end of sequence (optional): once a row has produced eos_id it is finished -- the eos itself is kept,
   * every later token of that row is pad_id, and lengths[b] holds the number of tokens up to and
   * including the eos.  lengths (batch) int64 must be initialised to -1 (= still running). */
  int64_t eos_id;     /* < 0: no end-of-sequence handling */
  int64_t pad_id;
  int64_t* lengths;   /* required when eos_id >= 0 */
} mtts_decode_greedy_params;
int mtts_decode_greedy(const mtts_decode_greedy_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * add_layernorm_fwd / add_layernorm_bwd -- residual add + LayerNorm (eps, affine) + optional FiLM:
 * the memory-bound glue of mamba_decoder.py:59,64,67,78,81-86,89 in one pass per tensor.
 *   x_out = x + delta                 (x, x_out: fp32 residual stream; delta: io dtype, optional)
 *   out   = LN(x_out) * w + b ; out = gamma_b * out + beta_b   (gamma/beta (batch, dim) fp32, optional)
 * rows = batch * rows_per_batch; all (rows, dim) tensors contiguous; delta_bias (dim) fp32 optional (the bias of the Linear that produced
 * delta: x_out = x + delta + delta_bias); dim % 4 == 0, dim <= 2048
 * (backward: dim <= 1024).  x_out may alias x.  mean / rstd (rows) fp32 are saved for the backward.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t rows, dim, rows_per_batch;
  int32_t io_dtype;
  float eps;
  const float* x;
  const void* delta; /* or NULL */
  const float* delta_bias; /* (dim) fp32 bias added together with delta, or NULL */
  float* x_out;      /* or NULL (then x_out == x is implied and nothing is written) */
  const float* ln_weight;
  const float* ln_bias;
  const float* film_gamma; /* both or neither */
  const float* film_beta;
  void* out;
  float* mean; /* or NULL */
  float* rstd; /* or NULL */
} mtts_add_layernorm_fwd_params;
int mtts_add_layernorm_fwd(const mtts_add_layernorm_fwd_params* p, mtts_stream_t stream);

/* Backward of the above.  dout (io dtype) = grad of `out`; dx_out (fp32, optional) = grad that
 * reaches x_out from the rest of the residual stream.  Writes dx (fp32) = grad of x, and ddelta
 * (io dtype, optional) = the same values for the branch.  colsum (batch, 3, dim) fp32 is ACCUMULATED
 * INTO with S1 = sum_t dout * xhat, S2 = sum_t dout and S3 = sum_t dx (= grad of delta_bias) per batch element; the parameter gradients
 * are linear combinations of S1/S2 the caller forms on (batch, dim)-sized tensors:
 *   dw = sum_b gamma_b S1_b, db = sum_b gamma_b S2_b, dgamma_b = w S1_b + bias S2_b, dbeta_b = S2_b. */
typedef struct {
  int32_t rows, dim, rows_per_batch;
  int32_t io_dtype;
  const float* x_out;
  const float* mean;
  const float* rstd;
  const float* ln_weight;
  const float* film_gamma; /* or NULL */
  const void* dout;
  const float* dx_out; /* or NULL */
  float* dx;
  void* ddelta; /* or NULL */
  float* colsum;
} mtts_add_layernorm_bwd_params;
int mtts_add_layernorm_bwd(const mtts_add_layernorm_bwd_params* p, mtts_stream_t stream);

/* The (batch, dim)-sized finishing step of a FiLM'd add_layernorm_bwd (norm_ff + gamma / beta, mamba_decoder.py:81-86)
 * in one launch: from colsum (batch, 3, dim)
 *   dweight = sum_b gamma_b S1_b,  dbias = sum_b gamma_b S2_b,  dgamma_b = w S1_b + bias S2_b,  dbeta_b = S2_b,
 *   ddelta_bias = sum_b S3_b (optional).  All fp32, all overwritten. */
typedef struct {
  int32_t batch, dim;
  const float* colsum;
  const float* film_gamma;
  const float* ln_weight;
  const float* ln_bias;
  float* dweight;
  float* dbias;
  float* dgamma;
  float* dbeta;
  float* ddelta_bias; /* or NULL */
} mtts_add_layernorm_finish_params;
int mtts_add_layernorm_bwd_finish(const mtts_add_layernorm_finish_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * length_regulate_fwd / length_regulate_bwd -- LengthRegulator.forward of the style pipeline next to the
 * decoder (style_cross_attention.py:144-198): phoneme rows repeated by their durations.
 *   dur[b, t] = max(round_half_even(durations[b, t]), 0);  output_lengths[b] = sum_t dur[b, t]
 *   expanded[b, f, :] = hidden[b, t(f), :] with t(f) the phoneme whose frame range contains f, for
 *   f < min(output_lengths[b], max_len); zeros beyond.  Every element of expanded is written.
 *   frame_index (batch, max_len) int32, optional: t(f), -1 for padding.
 * hidden (batch, t_text, dim), expanded (batch, max_len, dim) contiguous in the io dtype; durations fp32.
 * Backward: dhidden[b, t, :] = sum over the frames of phoneme t (inside max_len) of dexpanded[b, f, :];
 * every element of dhidden is written.  t_text <= 8192.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, t_text, dim, max_len;
  int32_t io_dtype;
  int32_t reserved;
  const void* hidden;
  const float* durations;
  void* expanded;
  int64_t* output_lengths; /* (batch) or NULL */
  int32_t* frame_index;    /* (batch, max_len) or NULL */
} mtts_length_regulate_fwd_params;
int mtts_length_regulate_fwd(const mtts_length_regulate_fwd_params* p, mtts_stream_t stream);

typedef struct {
  int32_t batch, t_text, dim, max_len;
  int32_t io_dtype;
  int32_t reserved;
  const float* durations;
  const void* dexpanded;
  void* dhidden;
} mtts_length_regulate_bwd_params;
int mtts_length_regulate_bwd(const mtts_length_regulate_bwd_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * skinny_linear -- decode_step's projections for m <= 64 rows in one launch:
 *     out = act( A @ W^T + bias ),  A = a                                   (ln_mode == 0)
 *                                   A = FiLM(LN(x + delta))                 (ln_mode == 1)
 * replaces nn.LayerNorm + residual add + nn.Linear (+ nn.GELU) of mamba_decoder.py:59-89 when the
 * sequence length is 1.  a (m, k) io dtype; x / x_out (m, k) fp32 residual stream (x_out receives
 * x + delta and must NOT alias x); delta (m, k) io dtype or NULL; film_gamma/beta (m, k) fp32 per-row
 * terms or NULL; w (n, k) and bias (n) io dtype; out (m, n) io dtype.  io dtype must be MTTS_BF16;
 * k % 16 == 0, and k <= 1024 when ln_mode.  gelu != 0 applies the exact (erf) GELU.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t m, n, k;
  int32_t io_dtype;
  int32_t ln_mode;
  int32_t gelu;
  float eps;
  const void* a;
  const float* x;
  const void* delta;
  float* x_out;
  const float* ln_weight;
  const float* ln_bias;
  const float* film_gamma;
  const float* film_beta;
  const void* w;
  const void* bias;
  void* out;
} mtts_skinny_linear_params;
int mtts_skinny_linear(const mtts_skinny_linear_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * gemm_bf16 -- the FFN contractions of mamba_decoder.py:39-43,88 on the tcgen05 tensor cores:
 *     out[m, n] = act( sum_k a[m, k] * w[n, k] + bias[n] )      act = exact-erf GELU when gelu != 0
 * a (m, k) and w (n, k) bf16, K contiguous (nn.Linear layout), leading dimensions lda / ldw / ldo in
 * elements (multiples of 8); bias (n) fp32 or NULL; out (m, n) bf16; pre_out (m, n) bf16, optional:
 * the pre-activation a backward pass needs.  fp32 accumulation in tensor memory.  k % 8 == 0, n % 8 == 0.
 * Replaces nn.Linear (+ nn.GELU) = a cuBLASLt GEMM plus an elementwise kernel.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t m, n, k;
  int32_t gelu;
  const void* a;
  int64_t lda;
  const void* w;
  int64_t ldw;
  const float* bias;
  void* out;
  int64_t ldo;
  void* pre_out;
} mtts_gemm_bf16_params;
int mtts_gemm_bf16(const mtts_gemm_bf16_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * gemm -- every dense contraction of the teacher-forced path on the tcgen05 tensor cores (bf16 operands,
 * fp32 accumulation in tensor memory), forward and backward:
 *     C[bo][bi][m, n] = epilogue( sum_k A[bo][bi][m, k] * B[bo][bi][n, k] )
 * Reference sites: Mamba in/x/dt/out projections (mamba_decoder.py:29,61 -> mamba_ssm nn.Linear /
 * mamba_inner_fn), nn.MultiheadAttention (:32-36,72-77: packed q / kv projections, QK^T, softmax, PV, out
 * projection), FFN (:39-43,88), head (:118,185) and their autograd backward (train.py:231).
 * Replaces cuBLASLt GEMMs, cuDNN SDPA and aten::gelu / softmax.
 *
 * Operand storage, element strides (lda / ldb and every batch stride multiples of 8, bases 16-byte aligned):
 *   a_major = 0 (K-major):  A[m, k] at a + m * lda + k          (nn.Linear weights, token-major activations)
 *   a_major = 1 (M-major):  A[m, k] at a + k * lda + m          (channel-major activations, transposes)
 *   likewise b_major / ldb for B[n, k].  Batch offsets: bo * x_bo_stride + bi * x_bi_stride; a stride of 0
 *   shares the operand across that batch dimension (weights).  batch_inner exists for (batch, head) views.
 * k_batches > 1 (weight gradients): the contraction also runs over kb in [0, k_batches), A and B advancing by
 *   their bo strides; batch_outer must be 1.  split_k > 1 (or -1 = automatic) splits the contraction across
 *   CTAs, which add into the fp32 output with vector REDs (the library zeroes it first unless accumulate).
 * Output: C row-major [m, n], leading dimension ldc, batch strides c_bo_stride / c_bi_stride; out_dtype
 *   MTTS_BF16 or MTTS_F32 (MTTS_EPI_STORE only).  accumulate != 0: C += result.
 * Epilogues:
 *   MTTS_EPI_STORE     C = acc + bias_n[n] + bias_m[m]                      (either bias may be NULL)
 *   MTTS_EPI_GELU      aux = pre = acc + bias_n (bf16, optional);  C = gelu(pre)        exact-erf GELU
 *                      with flag MTTS_GEMM_AUX_GELU_GRAD: aux = gelu'(pre) instead (one tanh serves both)
 *   MTTS_EPI_GELU_BWD  C = acc * gelu'(aux)                                 aux = the forward's pre (bf16)
 *   MTTS_EPI_MUL_AUX   C = acc * aux                                        aux = the forward's gelu'(pre) (bf16)
 *   MTTS_EPI_SOFTMAX   C = softmax_n(scale * acc + key mask)                n <= 256; mask (batch_outer, n)
 *                      uint8, 1 = attend, NULL = all; a fully masked row gives zeros
 *   MTTS_EPI_DSOFTMAX  C = scale * P o (acc - sum_n P o acc)                aux = P (bf16), n <= 256
 * aux is addressed like C with ld_aux / aux_bo_stride / aux_bi_stride.
 * ------------------------------------------------------------------------------------------- */
enum { MTTS_GEMM_SINGLE_CTA = 1, MTTS_GEMM_AUX_GELU_GRAD = 2 };
enum { MTTS_EPI_STORE = 0, MTTS_EPI_GELU = 1, MTTS_EPI_GELU_BWD = 2, MTTS_EPI_SOFTMAX = 3, MTTS_EPI_DSOFTMAX = 4,
       MTTS_EPI_MUL_AUX = 5 };
typedef struct {
  int32_t m, n, k;
  int32_t batch_outer, batch_inner;
  int32_t k_batches;
  int32_t a_major, b_major;
  int32_t out_dtype;
  int32_t epilogue;
  int32_t accumulate;
  int32_t split_k;
  const void* a;
  int64_t lda, a_bo_stride, a_bi_stride;
  const void* b;
  int64_t ldb, b_bo_stride, b_bi_stride;
  void* out;
  int64_t ldc, c_bo_stride, c_bi_stride;
  const float* bias_n;
  const float* bias_m;
  void* aux;
  int64_t ld_aux, aux_bo_stride, aux_bi_stride;
  const uint8_t* mask;
  int64_t mask_bo_stride;
  float scale;
  int32_t flags;         /* MTTS_GEMM_SINGLE_CTA: never pair CTAs (cta_group::1 tiles of 128 rows; measurements);
                            MTTS_GEMM_AUX_GELU_GRAD: see MTTS_EPI_GELU */
  float* row_stat;       /* MTTS_EPI_SOFTMAX, optional: (batch_outer, batch_inner, m) fp32, the base-2 log-sum-exp of
                            every row, lse2 = log2(sum_n exp(scale * acc + mask)): P = exp2(scale * log2(e) * acc - lse2);
                            +inf for a fully masked row.  What the fused attention backward needs instead of P. */
} mtts_gemm_params;
int mtts_gemm(const mtts_gemm_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * film_ffn_fwd / film_ffn_bwd -- the FiLM'd feed-forward branch of a decoder layer as ONE call each
 * (mamba_decoder.py:38-48,81-89: gamma/beta FiLM on LayerNorm(x), Linear -> GELU -> Linear), bf16 tensor-core path:
 *   fwd:  [x_out = x + delta (+ delta_bias);  h = FiLM(LN(x_out))]        when ln.x != NULL (mtts_add_layernorm_fwd)
 *         act = gelu(h W1^T + b1), gprime = gelu'(.)                      one mtts_gemm, both from its epilogue
 *         f   = act W2^T                                                  (W2's bias rides into the next add_layernorm)
 *   bwd:  dpre = (df W2) o gprime;  dW2 = df^T act;  dW1 = dpre^T h;  db1 = colsum(dpre);  dh = dpre W1  (dh optional)
 * tokens = rows of every activation; h (tokens, d_model), act / gprime / dpre (tokens, d_ff), f / df / dh
 * (tokens, d_model): bf16, contiguous; w1 (d_ff, d_model), w2 (d_model, d_ff): bf16; b1 fp32 or NULL;
 * dw1 / dw2 fp32 (overwritten), db1 fp32 (d_ff) or NULL, ACCUMULATED INTO (zero it first).  d_model, d_ff multiples of 8.
 * The same struct serves both directions: the forward reads h (or produces it through ln) and writes
 * act / gprime / f; the backward reads h / act / gprime / df and writes dpre / dw1 / db1 / dw2 / dh.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t tokens;
  int32_t d_model, d_ff;
  mtts_add_layernorm_fwd_params ln; /* ln.x == NULL: h is given; else ln.out must be h and ln.io_dtype MTTS_BF16 */
  const void* h;
  const void* w1;
  const float* b1;
  const void* w2;
  void* act;
  void* gprime;
  void* f;
  const void* df;
  void* dpre;
  float* dw1;
  float* db1;
  float* dw2;
  void* dh; /* or NULL */
} mtts_film_ffn_params;
int mtts_film_ffn_fwd(const mtts_film_ffn_params* p, mtts_stream_t stream);
int mtts_film_ffn_bwd(const mtts_film_ffn_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * cross_attn_fwd / cross_attn_bwd -- nn.MultiheadAttention(batch_first=True) of the teacher-forced path
 * (mamba_decoder.py:32-36,72-77; key_padding_mask = ~mask) without its out-projection bias, as ONE call each; every
 * contraction is an mtts_gemm launch (q / kv projections, QK^T with scale + key mask + softmax in the epilogue,
 * PV, out projection; backward: dO, dW_o, dV = P^T dO, dS = dsoftmax(dO V^T), dQ = dS K, dK = dS^T Q, dW_in,
 * db_in, d query, d memory).  t_kv <= 256, head_dim and d_model multiples of 8.
 *   query (batch, t_q, E), memory (batch, t_kv, E): bf16 contiguous;  w_in (3E, E) = [Wq; Wk; Wv], w_out (E, E): bf16;
 *   b_in (3E) fp32;  mask (batch, t_kv) uint8, 1 = attend, or NULL.
 *   saved by the forward (caller's buffers): q (batch, t_q, E), kv (batch, t_kv, 2E), p (batch, heads, t_q, t_kvp)
 *   with t_kvp = t_kv rounded up to 8, o (batch, t_q, E);  out (batch, t_q, E).
 *   backward: dout (batch, t_q, E); workspaces d_o (batch, t_q, E), ds (like p), dq (like q), dkv (like kv);
 *   results dw_in (3E, E), dw_out (E, E) fp32 overwritten, db_in (3E) fp32 ACCUMULATED INTO, dquery / dmemory
 *   (like query / memory; either may be NULL).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, t_q, t_kv, d_model, heads;
  int32_t reserved;
  const void* query;
  const void* memory;
  const void* w_in;
  const float* b_in;
  const void* w_out;
  const uint8_t* mask;
  void* q;
  void* kv;
  void* p;
  void* o;
  void* out;
  float* lse2; /* (batch, heads, t_q) fp32 or NULL: written by the forward; when given to the backward (and head_dim is
                  64) the fused mtts_attn_core_bwd replaces the GEMM-by-GEMM core and `p` / `ds` are not touched; a
                  forward with lse2 and p == NULL (head_dim 64) runs the fused mtts_attn_core_fwd */
  const void* dout;
  void* d_o;
  void* ds;
  void* dq;
  void* dkv;
  float* dw_in;
  float* db_in;
  float* dw_out;
  void* dquery;
  void* dmemory;
} mtts_cross_attn_params;
int mtts_cross_attn_fwd(const mtts_cross_attn_params* p, mtts_stream_t stream);
int mtts_cross_attn_bwd(const mtts_cross_attn_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * attn_core_bwd -- the backward of softmax(scale q k^T + mask) v (the inside of nn.MultiheadAttention,
 * mamba_decoder.py:32-36,72-77, under loss.backward(), train.py:231) for all heads in ONE launch, with the score-sized
 * tensors (P, dP, dS) kept on the SM (tensor memory / shared memory): given the forward's per-row base-2
 * log-sum-exp (mtts_gemm's row_stat), o and d_o, it produces dq and dk | dv.  head_dim = 64 (d_model = 64 heads),
 * t_kv <= 256; q, o, d_o, dq (batch, t_q, d_model), kv, dkv (batch, t_kv, 2 d_model) bf16 contiguous; lse2
 * (batch, heads, t_q) fp32; mask (batch, t_kv) uint8, 1 = attend, or NULL.  Used by mtts_cross_attn_bwd when the
 * shapes allow (replaces its dV / dsoftmax / dQ / dK GEMMs).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, heads, t_q, t_kv, d_model;
  float scale;
  const void* q;
  const void* kv;
  const void* o;
  const void* d_o;
  const float* lse2;
  const uint8_t* mask;
  void* dq;
  void* dkv;
} mtts_attn_core_bwd_params;
int mtts_attn_core_bwd(const mtts_attn_core_bwd_params* p, mtts_stream_t stream);

/* The forward of the same core (mamba_decoder.py:72-77) in one launch: o = softmax(scale q k^T + mask) v (bf16, (batch, t_q, d_model)) and
 * lse2 (batch, heads, t_q) fp32, the scores and probabilities kept in tensor / shared memory.  Same shape limits.
 * Used by mtts_cross_attn_fwd when the caller gives lse2 and no probability buffer. */
typedef struct {
  int32_t batch, heads, t_q, t_kv, d_model;
  float scale;
  const void* q;
  const void* kv;
  const uint8_t* mask;
  void* o;
  float* lse2;
} mtts_attn_core_fwd_params;
int mtts_attn_core_fwd(const mtts_attn_core_fwd_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * FFN / projection glue (mamba_decoder.py:39-43,86-88: Linear -> GELU -> Linear, and the bias gradients
 * of every biased Linear on the path).  Row-major (rows, cols) tensors of the io dtype with leading
 * dimension ld (elements); cols, ld multiples of the 16-byte vector (8 bf16 / 4 fp32).
 *   bias_gelu_fwd: out = gelu(x + bias)                       exact erf GELU; bias (cols) fp32 or NULL
 *   bias_gelu_bwd: out = dout * gelu'(x + bias);  colsum[c] += sum_r out[r, c]   (colsum fp32, ACCUMULATED)
 *   colsum:        colsum[c] += sum_r x[r, c]                  (dout, out, bias ignored)
 * Replace aten::gelu / gelu_backward / sum(0) [upstream: nn.GELU, nn.Linear's bias gradient].
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t rows, cols;
  int32_t io_dtype;
  int32_t reserved;
  int64_t ld;
  const void* x;
  const float* bias;
  const void* dout;
  void* out;
  float* colsum;
} mtts_bias_gelu_params;
int mtts_bias_gelu_fwd(const mtts_bias_gelu_params* p, mtts_stream_t stream);
int mtts_bias_gelu_bwd(const mtts_bias_gelu_params* p, mtts_stream_t stream);
int mtts_colsum(const mtts_bias_gelu_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The caller's side of the training step (train.py:31-42,115-131,152-159,230-235).
 *
 * embed_sum_fwd: x[b, l, :] = token_embed[tokens[b, l]] + pos_embed[pos_ids[l]] (+ quant_embed[quant_ids[l]])
 *   -- mamba_decoder.py:167-171 (teacher-forced embedding sum) and train.py:115-131 (`embed_codec_tokens`).
 *   tokens (batch, seqlen) int64; pos_ids, quant_ids (seqlen) int64 (quant_ids NULL: no quantizer term);
 *   tables fp32 row-major (rows, dim); x (batch, seqlen, dim) fp32; dim % 4 == 0.
 * embed_sum_bwd: the same struct with x = d loss / d x and the three table pointers = their (pre-zeroed or
 *   accumulating) fp32 gradients; a NULL table pointer skips that gradient.  Replaces three aten::embedding /
 *   embedding_dense_backward launches and two adds.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t batch, seqlen, dim;
  int32_t reserved;
  const int64_t* tokens;
  const int64_t* pos_ids;
  const int64_t* quant_ids;
  float* token_embed;
  float* pos_embed;
  float* quant_embed;
  float* x;
} mtts_embed_sum_params;
int mtts_embed_sum_fwd(const mtts_embed_sum_params* p, mtts_stream_t stream);
int mtts_embed_sum_bwd(const mtts_embed_sum_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * ce_loss -- F.cross_entropy(logits, targets, ignore_index) of train.py:31-42 in one pass:
 *   loss_sum += sum over rows with targets != ignore_index of (logsumexp(logits[r]) - logits[r, target])
 *   dlogits[r, :] = grad_scale / n_valid * (softmax(logits[r]) - onehot(target))   (zeros for ignored rows)
 * logits (rows, vocab) io dtype, leading dimension ld; targets (rows) int64; n_valid: device pointer to the
 * number of non-ignored targets (of the GLOBAL batch under data parallelism); loss_sum fp32 device scalar,
 * ACCUMULATED (zero it first; the mean loss is loss_sum / n_valid); row_loss (rows) fp32 optional; dlogits NULL
 * skips the gradient.  vocab, ld multiples of the 16-byte vector.  Replaces aten log_softmax + nll_loss forward
 * and backward and the fp32 copy of the logits.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t rows;
  int32_t vocab;
  int32_t io_dtype;
  int64_t ld;
  int64_t ignore_index;
  const void* logits;
  const int64_t* targets;
  const float* n_valid;
  float grad_scale;
  int32_t reserved;
  float* loss_sum;
  float* row_loss;
  void* dlogits;
} mtts_ce_loss_params;
int mtts_ce_loss(const mtts_ce_loss_params* p, mtts_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * grad_sumsq + adam_step -- torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam.step()
 * (train.py:152-159,233-234) over a DEVICE-resident table of fp32 tensors:
 *   grad_sumsq:  *grad_sumsq = sum over all tensors of sum(grad^2)           (zeroed by the call)
 *   adam_step:   g = grad * min(1, max_norm / (sqrt(*grad_sumsq) + 1e-6))    (max_norm <= 0: no clipping)
 *                exp_avg = beta1 exp_avg + (1 - beta1) g;  exp_avg_sq = beta2 exp_avg_sq + (1 - beta2) g^2
 *                param -= step_size * exp_avg / (sqrt(exp_avg_sq) / bias_correction2_sqrt + eps)
 *   with step_size = lr / (1 - beta1^t), bias_correction2_sqrt = sqrt(1 - beta2^t) computed by the caller.
 * tensors: device array of mtts_adam_tensor; chunks: device array of (tensor index, chunk index) int32 pairs,
 * chunk = mtts_adam_chunk_elems() consecutive elements.  The gradients are not modified.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} mtts_adam_tensor;
typedef struct {
  const mtts_adam_tensor* tensors;
  const int32_t* chunks;
  int32_t num_chunks;
  float max_norm;
  float* grad_sumsq;
  float step_size, beta1, beta2, eps, bias_correction2_sqrt;
  int32_t reserved;
} mtts_adam_params;
int mtts_grad_sumsq(const mtts_adam_params* p, mtts_stream_t stream);
int mtts_adam_step(const mtts_adam_params* p, mtts_stream_t stream);
int mtts_adam_chunk_elems(void);

#ifdef __cplusplus
}
#endif
#endif /* MAMBA_TTS_B200_H_ */
