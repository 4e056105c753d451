/* osb200.h — C ABI of libosb200.so, the B200 (sm_100a) implementation of the OptiSpeech
 * synthesis / training hot path.
 *
 * The reference (mush42/optispeech @ 3bdde20) is pure Python/PyTorch and has NO native
 * interface; every entry point below therefore names the reference *Python* function it
 * replaces (file:line relative to the reference checkout).  The host side
 * (optispeech_b200/, Python) binds these with ctypes — see INTEGRATION.md for the stub a
 * reference maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named h_*;
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never
 *     allocates or frees device memory and never synchronises the stream;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the
 *     functions are CUDA-graph-capture safe;
 *   - return value: 0 = OSB_OK, negative = osb_status, positive = cudaError_t of the launch;
 *   - activations are channels-last: a (B, T, C) tensor is B*T rows of C contiguous values;
 *   - "h16" buffers hold IEEE fp16 (the tensor-core operand type), "f32" buffers fp32.
 */
#ifndef OSB200_H_
#define OSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum osb_status {
  OSB_OK = 0,
  OSB_ERR_SHAPE = -1,     /* unsupported / inconsistent shape            */
  OSB_ERR_ALIGN = -2,     /* pointer or leading dimension not aligned    */
  OSB_ERR_ARCH = -3,      /* device is not sm_100                         */
  OSB_ERR_DRIVER = -4,    /* cuTensorMapEncodeTiled unavailable / failed  */
  OSB_ERR_WORKSPACE = -5, /* workspace too small                          */
  OSB_ERR_ARG = -6        /* null pointer / bad enum                      */
} osb_status;

/* library bookkeeping ------------------------------------------------------------------- */
int osb_version(void);                        /* ABI version, bumps on any signature change */
const char* osb_strerror(int status);         /* static string for a negative osb_status    */
unsigned long long osb_launch_count(void);    /* kernels launched by this library so far    */
int osb_check_device(int device);             /* OSB_OK iff `device` is compute capability 10.x */

/* ---------------------------------------------------------------------------------------
 * Tensor-core contraction with fused epilogue (tcgen05.mma, TMA-staged operands, TMEM
 * accumulators).  Computes, for every batch b, row t and output channel n,
 *
 *     acc[b,t,n] = sum_{tap<taps} sum_{k<K} A[b, t + tap - pad, k] * W[tap, n, k]
 *
 * (rows outside [0,T) read as zero, i.e. Conv1d zero padding), then applies `epi`.
 * With taps = 1 this is nn.Linear; with taps = k it is nn.Conv1d(kernel_size=k,
 * padding=pad) on channels-last data.
 *
 * Replaces: nn.Linear / nn.Conv1d call sites of the reference —
 *   ConvNeXtBlock.pwconv1/pwconv2        optispeech/model/generator/modules/convnext.py:24-26,39-41
 *   VariancePredictor.conv[i][0]/.linear optispeech/model/generator/modules/core.py:66-81,92-96
 *   AlignmentModule.{t,f}_conv*          optispeech/model/generator/alignments.py:34-40,55-64
 *   WaveNeXt.embed, WaveNeXtHead.linear_* optispeech/model/vocoder/wavenext/__init__.py:24-25,43-47,67,83
 * ------------------------------------------------------------------------------------- */
typedef enum osb_epilogue {
  OSB_EPI_BIAS = 0,     /* out_f32 = acc + bias                     (flags: CLIP, KEEPMASK, OUT_H16)  */
  OSB_EPI_GELU = 1,     /* out_h16 = gelu_erf(acc + bias)           (flags: SAVE_PRE -> aux_h16)     */
  OSB_EPI_RESID = 2,    /* out_f32 = (resid + gamma*(acc+bias)*row_scale[b]) * keep[b,t]  (SAVE_PRE: aux = fp16(acc+bias)) */
  OSB_EPI_RELU_LN = 3,  /* y = LN(relu(acc+bias)) -> out_h16; needs BN == N (flags: DOT, SAVE_PRE)    */
  OSB_EPI_BIAS_LN = 4,  /* out_f32 = LN(acc + bias); needs BN == N  (flags: OUT_H16)                  */
  OSB_EPI_RELU = 5,     /* out_h16 = relu(acc + bias)               (flags: none)                     */
  /* backward epilogues (acc is a gradient; aux_in_h16 is an activation saved by the forward pass) */
  OSB_EPI_GELU_BWD = 6,    /* out_h16 = acc * gelu_erf'(aux_in)                                          */
  OSB_EPI_LN_BWD = 7,      /* out_f32 = LayerNorm backward of acc wrt its input, given xhat = aux_in and
                              rstd = row_stat (no affine: it is folded into the GEMM weights); BN == N   */
  OSB_EPI_RELU_LN_BWD = 8, /* acc = grad of a RELU_LN output y; aux_in = r = relu(conv) saved by SAVE_PRE;
                              out_h16 = grad wrt the conv output = LN_bwd(acc * ln_w; r) * [r > 0];
                              with OUT_H16 also aux_h16 = fp16(acc) (for the LN parameter gradients); BN == N */
  OSB_EPI_RELU_BWD = 9,    /* out_h16 = acc * [aux_in > 0]                                               */
  /* attention of the alignment module (AlignmentModule.forward, generator/alignments.py:66-81), BN == padded N:
   *   score[t,n] = -sqrt(max(row_stat[t] + bias[b*N+n] - 2*acc, 0))   (row_stat = |f_t|^2, bias = |e_n|^2)
   *   out_f32[t,n] = score - logsumexp_{n < col_len[b]}(score) + resid[t,n]  (resid = beta-binomial log-prior), -inf for n >= col_len[b]
   *   out_dot[t] = that logsumexp                                                                           */
  OSB_EPI_ATTN_LOGP = 10,
  OSB_EPI_AXPY = 11        /* out_f32 = acc + row_stat[row] * resid[row, n]                                 */
} osb_epilogue;

enum {
  OSB_FLAG_CLIP = 1,      /* clamp result to [-1, 1]                     (WaveNeXtHead, wavenext/__init__.py:47) */
  OSB_FLAG_KEEPMASK = 2,  /* multiply row by (1 - pad_mask[b,t])                                               */
  OSB_FLAG_OUT_H16 = 4,   /* also write an fp16 copy of the result to aux_h16                                  */
  OSB_FLAG_SAVE_PRE = 8,  /* write the pre-activation / pre-LN value as fp16 to aux_h16 (for backward)         */
  OSB_FLAG_DOT = 16,      /* RELU_LN only: out_dot[b,t] = <LN row, dot_w> + dot_b, 0 at padded rows            */
  /* Split precision ("fp16x3"): a value v is carried as hi = fp16(v), lo = fp16(v - hi) and the product
   * is accumulated as a_hi*w_hi + a_lo*w_hi + a_hi*w_lo in the fp32 TMEM accumulator (~2^-21 relative
   * error instead of 2^-11).  Needed to hold the 1e-3 waveform tolerance at full-scale amplitude. */
  OSB_FLAG_SPLIT_IN = 32, /* a is (B,T,[hi K | lo K]) with lda >= 2K; w is (2, taps, N, ldw): [0] = hi parts, [1] = lo parts */
  OSB_FLAG_SPLIT_OUT = 64, /* fp16 outputs are written as rows [hi ldo | lo ldo] (row stride 2*ldo)                */
  OSB_FLAG_RELU = 128,     /* EPI_BIAS only: out = relu(acc + bias)                                                */
  OSB_FLAG_NO_F32 = 256,   /* EPI_BIAS with OUT_H16: write only the fp16 copy (out may be NULL)                    */
  OSB_FLAG_COLSUM = 512,   /* GELU_BWD / RELU_BWD / RELU_LN_BWD: out_colsum[n] += sum_rows out_h16[row, n] — the bias
                              gradient of the layer this dgrad belongs to, taken from the tile while it is on chip   */
  /* Data gradient on the FORWARD weight pack (no transposed copy): w is (taps, K, ldw >= N) with the output index n
   * contiguous — i.e. the forward pack (taps, N_fwd, K_fwd) of the layer read as (taps, K = N_fwd, N = K_fwd) — and is fed
   * to tcgen05.mma as an MN-major B operand:  acc[b,t,n] = sum_tap sum_k a[b,t+tap-pad,k] * w[tap', k, n],
   * tap' = tap, or taps-1-tap with TAP_REVERSE (the transposed convolution of a Conv1d dgrad).  N % 64 == 0. */
  OSB_FLAG_W_MN = 1024,
  OSB_FLAG_TAP_REVERSE = 2048,
  OSB_FLAG_LRELU = 4096    /* EPI_BIAS only: out = leaky_relu(acc + bias, lrelu_slope)                              */
};

typedef struct osb_gemm_desc {
  /* operands */
  const void* a;        /* fp16 (B, T, lda) activations, K valid columns                           */
  const void* w;        /* fp16 (taps, N, ldw) packed weights, K valid columns                     */
  int64_t lda, ldw;     /* leading dimensions in elements, multiples of 8                          */
  int32_t B, T, N, K;   /* batch, rows per batch, output channels, contraction length per tap      */
  int32_t taps, pad;    /* kernel size and left zero padding                                       */
  /* epilogue */
  int32_t epi;          /* osb_epilogue                                                            */
  int32_t flags;        /* OSB_FLAG_*                                                              */
  void* out;            /* fp32 or fp16 (B*T, ldo) primary output (see osb_epilogue)               */
  void* aux_h16;        /* optional fp16 (B*T, ldo) secondary output                               */
  int64_t ldo;          /* leading dimension of out / aux / resid, elements                        */
  const float* bias;    /* (N) or NULL                                                             */
  const float* resid;   /* RESID: fp32 (B*T, ldo)                                                  */
  const float* gamma;   /* RESID: (N) layer scale                                                  */
  const float* row_scale; /* RESID: optional (B) DropPath scale per sample (NULL = 1)              */
  const uint8_t* pad_mask; /* optional (B*T) bytes, 1 = padded position                            */
  const float* ln_w;    /* *_LN: (N) LayerNorm weight                                              */
  const float* ln_b;    /* *_LN: (N) LayerNorm bias                                                */
  float ln_eps;
  const float* dot_w;   /* DOT: (N) weight of the trailing Linear(N -> 1)                          */
  const float* dot_b;   /* DOT: (1) its bias (device pointer, may be NULL = 0)                     */
  float* out_dot;       /* DOT: (B*T) fp32                                                         */
  const void* aux_in_h16; /* *_BWD: fp16 (B*T, ldo) saved activation                                */
  const float* row_stat;  /* LN_BWD: (B*T) rstd saved by osb_dwconv_ln                              */
  /* Dropout after the LayerNorm of RELU_LN (VariancePredictor, core.py:78), element (row, n) keyed by
   * dropout_seed: forward scales the LN output, RELU_LN_BWD scales the incoming gradient by the same mask.
   * Also honoured by EPI_RELU (after the ReLU), EPI_RESID (on acc+bias, before the residual add) and EPI_RELU_BWD
   * (multiplies by 1/(1-p); its aux_in must be the post-dropout ReLU output) — the Transformer layers' dropouts. */
  float dropout_p;        /* 0 disables                                                              */
  uint64_t dropout_seed;
  const uint64_t* dropout_seed_dev; /* optional device scalar added to dropout_seed (lets a captured CUDA graph draw a new
                                       mask on every replay: the host bumps / a captured kernel increments the scalar)      */
  /* per-batch B operand (batched matmul): w is (B, N, ldw) — with SPLIT_IN (B, N, [hi K | lo K]) — and taps == 1 */
  int32_t w_batched;
  const int64_t* col_len; /* ATTN_LOGP: (B) number of valid columns (text length)                    */
  float* out_colsum;      /* COLSUM: (N) fp32, accumulated with atomics (caller zeroes it)            */
  /* Strided convolution (row_stride > 1; groundwork for the period discriminators' (5,1)/stride-3 convs,
   * vocoder/wavenext/disc/_discriminators.py:52-60): output row t reads input rows t*row_stride + tap - pad of an input with
   * T_in rows per batch (a is (B, T_in, lda)); T is the number of OUTPUT rows.  The stride is a TMA traversal stride
   * (elementStrides) — no im2col copy.  0 / 1 = dense rows (T_in ignored). */
  int32_t row_stride;
  int32_t T_in;
  float lrelu_slope;      /* OSB_FLAG_LRELU: negative slope                                          */
  /* OSB_FLAG_KEEPMASK without a pad_mask: rows are grouped in sequences of seq_pitch rows whose first seq_valid rows are
   * kept and the rest zeroed (the flat sequence layout of the period discriminators: the zero tail of one sequence is the
   * convolution padding of the next).  0 = unused. */
  int32_t seq_pitch;
  int32_t seq_valid;
} osb_gemm_desc;

int osb_gemm(const osb_gemm_desc* desc, void* stream);

/* Weight-gradient contraction (split over rows, fp32 atomic accumulation into dw):
 *     dw[tap, n, k] += sum_{b,t} dy[b, t, n] * a[b, t + tap - pad, k]
 * dy: fp16 (B, T, ldy) ; a: fp16 (B, T, lda) ; dw: fp32 (taps, N, K) (must be zeroed or hold the
 * running accumulation; 16-byte aligned, K a multiple of 4: the tile is flushed with vector reds,
 * a warp per 512-byte row segment).  Both operands are consumed MN-major straight from their
 * channels-last layout (no transposes in HBM).
 * Replaces autograd's Conv1d/Linear weight gradient for the layers listed above. */
int osb_gemm_wgrad(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T,
                   int32_t N, int32_t K, int32_t taps, int32_t pad, void* stream);

/* ---------------------------------------------------------------------------------------
 * ONE kernel per ConvNeXt block, forward (osb_convnext.cu): depthwise-conv7 + LayerNorm prologue on CUDA cores writing the
 * fp16 A operand into swizzled shared memory, pwconv1 (tcgen05) -> bias + erf-GELU epilogue into shared memory -> pwconv2
 * (tcgen05, accumulating over 64-column chunks of the intermediate dimension in TMEM) -> bias, layer scale, DropPath scale,
 * residual and pad mask.  Weights are TMA-streamed fp16: w1f (I, C) = pwconv1.weight * norm.weight (LN affine folded),
 * b1f = pwconv1.bias + pwconv1.weight @ norm.bias, w2 (C, I).  (C, I) in {(256, 1024), (384, 1152)}.
 * Replaces ConvNeXtBlock.forward + the mask of ConvNeXtBackbone.forward (modules/convnext.py:34-47,98-101). */
int osb_convnext_block_fwd(const float* x, const float* dw_w /*(C,7)*/, const float* dw_b, const void* w1f_h16, const float* b1f,
                           const void* w2_h16, const float* b2, const float* gamma, const float* row_scale /*(B) or NULL*/,
                           const uint8_t* pad_mask /*(B*T) or NULL*/, float* out, int32_t B, int32_t T, int32_t C, int32_t I, float eps,
                           void* stream);

/* ---------------------------------------------------------------------------------------
 * HBM-bound kernels of the synthesis path (osb_pointwise.cu).  Rows are channels-last.
 * ------------------------------------------------------------------------------------- */

/* out[b,t,:] = sqrt(dim)*table[ids[b,t]] + scale[0]*[sin(t*inv_freq) | cos(t*inv_freq)]   (fp32)
 * Replaces TextEmbedding.forward (optispeech/model/generator/modules/core.py:25-31) and
 * ScaledSinusoidalEmbedding.forward (modules/layers.py:59-71); dropout is the caller's job. */
int osb_embed_text(const int64_t* ids, const float* table, const float* inv_freq, const float* scale, float* out,
                   int32_t B, int32_t T, int32_t dim, int32_t n_vocab, void* stream);

/* xhat[b,t,:] = normalise_C( bias + sum_j w[:,j] * x[b,t+j-3,:] )  as fp16 (no affine: the
 * LayerNorm weight/bias are folded into the pointwise-1 GEMM weights by the host), rstd optional.
 * Replaces ConvNeXtBlock.dwconv + .norm (modules/convnext.py:22-23,36-38).  C in {128,256,384,512}. */
int osb_dwconv_ln(const float* x, const float* w /*(C,7)*/, const float* bias, void* xhat_h16, float* rstd /*(B*T) or NULL*/,
                  int32_t B, int32_t T, int32_t C, float eps, int32_t split /* rows [hi C | lo C] */, void* stream);

/* Row LayerNorm with affine; fp32 and/or fp16 output.  Replaces ConvNeXtBackbone.final_layer_norm
 * (modules/convnext.py:84,102) and WaveNeXt.norm (vocoder/wavenext/__init__.py:68,84). */
int osb_layernorm(const float* x, const float* w, const float* b, float* out_f32, void* out_h16, int64_t rows, int32_t C,
                  float eps, int32_t split, void* stream);

/* y = LayerNorm(relu(x)) * w + b per row -> fp16 (plain / split rows [hi C | lo C]) and / or out_dot[row] =
 * pad_mask[row] ? 0 : y . dot_w + dot_b[0].  The stand-alone form of OSB_EPI_RELU_LN (+ OSB_FLAG_DOT): a VariancePredictor
 * layer's ReLU -> LayerNorm (-> Linear(C,1) -> masked_fill) (modules/core.py:73-96) behind a plain-bias osb_gemm in narrow
 * tiles, used when the problem is a few row tiles (synthesis of one utterance).  x already holds conv + bias. */
int osb_relu_layernorm(const float* x, const float* w, const float* b, void* out_h16 /* or NULL */, int64_t rows, int32_t C, float eps,
                       int32_t split, const float* dot_w /* or NULL */, const float* dot_b /* or NULL */,
                       const uint8_t* pad_mask /* (rows) or NULL */, float* out_dot /* (rows) or NULL */, void* stream);

/* out = (x + emb_scale * (bias + Conv1d(1->C, k, same)(val))) * (1 - pad_mask); emb_scale (B,T,C) is the optional
 * dropout mask/(1-p) of the embedding branch (NULL = 1).  Replaces PitchPredictor.forward/infer's embed + add + mask
 * (modules/core.py:143-176). */
int osb_variance_embed(const float* x, const float* val /*(B,T)*/, const float* w /*(C,k)*/, const float* bias,
                       const uint8_t* pad_mask, const float* emb_scale, float* out_f32, void* out_h16, int32_t B, int32_t T, int32_t C, int32_t ksize,
                       int32_t split, void* stream);

/* dur = clamp(ceil((exp(log_d) - clip_val) * factor), 0) as int64, 0 at pads; lengths[b] = sum_t dur.
 * Replaces DurationPredictor.infer (modules/core.py:126-133) and y_lengths (generator/__init__.py:258). */
int osb_durations(const float* log_d, const uint8_t* pad_mask, int64_t* dur, int64_t* lengths, int32_t B, int32_t T,
                  float factor, float clip_val, void* stream);

/* centres = cumsum(dur) - dur/2 (fp32) and csum = inclusive cumsum (int64, optional).  dur is int64 or fp32.
 * Replaces GaussianUpsampling's `c` (generator/alignments.py:167) and expand_by_duration's cumsum (:287). */
int osb_centres(const void* dur, int32_t dur_is_i64, float* centres, int64_t* csum, int32_t B, int32_t T, void* stream);

/* y[b,t,:] = softmax_i(-delta*(t*[t<y_len] - c_i)^2 over i < x_len) @ hs[b].  Replaces
 * GaussianUpsampling.forward (generator/alignments.py:159-173) without materialising p_attn in HBM. */
int osb_gaussian_upsample(const float* hs, const float* centres, const int64_t* x_len, const int64_t* y_len, float* out_f32,
                          void* out_h16, int32_t B, int32_t Tx, int32_t Tm, int32_t C, float delta, void* stream);

/* The same rows for a WINDOW of frames per sample: output row j of sample b is frame win_start[b] - halo + j of the Tm-frame
 * sequence, j < W; rows outside [0, Tm) are zero (what the zero padding of the decoder's convolutions reads there).  The
 * training step only consumes `segment_size` decoder frames per sample (get_random_segments, generator/__init__.py:146-152) and
 * the ConvNeXt decoder is local (3 frames of context per block), so upsampler + decoder run on segment + 2 * halo frames. */
int osb_gaussian_upsample_window(const float* hs, const float* centres, const int64_t* x_len, const int64_t* y_len,
                                 const int64_t* win_start /*(B)*/, float* out_f32, void* out_h16, int32_t B, int32_t Tx, int32_t Tm,
                                 int32_t W, int32_t halo, int32_t C, float delta, void* stream);

/* Hard length regulator: out[b,t,:] = x[b, i(t), :] with csum[i-1] <= t < csum[i]; zero past the length.
 * index_out (optional, int32, -1 past the length) is the bit-exact indexing target.
 * Replaces expand_by_duration (generator/alignments.py:283-297). */
int osb_expand_gather(const float* x, const int64_t* csum, float* out, int32_t* index_out, int32_t B, int32_t Tx, int32_t Tm,
                      int32_t C, void* stream);

/* v = src[r*src_ld + c*src_cs] * col_scale[c];  dst[r*dst_rs + c] = hi = fp16(v) for c < cols, 0 for cols <= c < dst_cols;
 * if dst_lo != NULL also dst_lo[r*dst_rs + c] = fp16(v - hi).
 * Weight/activation packing for the tensor-core operands (no reference counterpart: the
 * reference's `16-mixed` autocast does this implicitly, configs/trainer/default.yaml:11). */
int osb_pack_h16(const float* src, int64_t src_ld, int64_t src_cs, const float* col_scale, void* dst, void* dst_lo, int64_t dst_rs,
                 int32_t dst_cols, int64_t rows, int32_t cols, void* stream);

/* Conv1d weight (N, Cin, k) fp32 -> fp16 operand for all taps in one launch: forward form (k, N, Kp) or, with
 * transpose_reverse, the dgrad form (k, Cin, N) with the taps reversed; dst_lo (optional) = rounding residual. */
int osb_pack_conv_h16(const float* w, void* dst, void* dst_lo, int32_t N, int32_t Cin, int32_t k, int32_t Kp,
                      int32_t transpose_reverse, void* stream);

/* ---------------------------------------------------------------------------------------
 * HBM-bound backward kernels (osb_backward.cu).  They replace what torch.autograd derives for the
 * reference modules named at each entry; parameter gradients are ACCUMULATED (+=) into caller-zeroed
 * fp32 buffers.  Incoming gradients may carry the host's static loss scale (everything is linear).
 * ------------------------------------------------------------------------------------- */

/* Backward of the residual epilogue out = (x + gamma * z * row_scale[b]) * keep  (ConvNeXtBlock, convnext.py:42-46,
 * + the mask of ConvNeXtBackbone.forward :100-101):  dyg = fp16(dout*keep*rs*gamma), dgamma += sum dout*keep*rs*z,
 * db2 += sum dout*keep*rs*gamma.  z = pwconv2 output (fp16, saved by OSB_EPI_RESID + OSB_FLAG_SAVE_PRE). */
int osb_resid_bwd_prep(const float* dout, const void* z_h16, const float* gamma, const uint8_t* pad_mask, const float* row_scale,
                       void* dyg_h16, float* dgamma, float* db2, int64_t rows, int32_t T, int32_t C, void* stream);

/* out[n] += sum_rows x[row, n] for an fp16 (rows, N) matrix: bias gradients of Linear / Conv1d layers. */
int osb_colsum_h16(const void* x_h16, float* out, int64_t rows, int32_t N, void* stream);

/* Undo the LayerNorm-affine folding of pwconv1 (W1f = W1 diag(ln_w), b1f = b1 + W1 ln_b):
 * dw1 (in: dW1f, out: dW1[i,c] = dW1f[i,c] ln_w[c] + db1[i] ln_b[c]); dln_w[c] += sum_i dW1f[i,c] W1[i,c];
 * dln_b[c] += sum_i db1[i] W1[i,c]. */
int osb_ln_fold_bwd(float* dw1, const float* w1, const float* ln_w, const float* ln_b, const float* db1, float* dln_w, float* dln_b,
                    int32_t I, int32_t C, void* stream);

/* Depthwise Conv1d(k=7) backward plus the residual path of the block:
 * dx = dout*keep + corr(dd, w);  ddw[c,j] += sum dd[b,t,c] x[b,t+j-3,c];  ddb[c] += sum dd.   (convnext.py:22,36) */
int osb_dwconv_bwd(const float* dd, const float* dout, const float* x, const float* w /*(C,7)*/, const uint8_t* pad_mask, float* dx,
                   float* ddw, float* ddb, int32_t B, int32_t T, int32_t C, void* stream);

/* LayerNorm(affine) backward, statistics recomputed from x (final_layer_norm, convnext.py:84,102). */
int osb_layernorm_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db, int64_t rows, int32_t C,
                      float eps, void* stream);

/* Backward of a VariancePredictor's tail (core.py:92-96): last LayerNorm + Linear(->1) + masked_fill.
 * g_conv = fp16 gradient wrt the last Conv1d output (ReLU gate applied); accumulates dlin_w, dlin_b, dln_w, dln_b.
 * r = relu(conv) of the last layer, fp16, saved by OSB_EPI_RELU_LN + OSB_FLAG_SAVE_PRE. */
int osb_predictor_tail_bwd(const float* d_out /*(rows)*/, const uint8_t* pad_mask, const void* r_h16, const float* ln_w,
                           const float* ln_b, const float* lin_w, void* g_conv_h16, float* dlin_w, float* dlin_b, float* dln_w,
                           float* dln_b, int64_t rows, int32_t C, float eps, float dropout_p, uint64_t dropout_seed,
                           const uint64_t* dropout_seed_dev, void* stream);

/* LayerNorm parameter gradients of an inner predictor layer: dln_w += sum gy*xhat(r), dln_b += sum gy, with
 * gy = fp16 gradient wrt the layer output (aux of OSB_EPI_RELU_LN_BWD + OSB_FLAG_OUT_H16). */
int osb_ln_param_grad(const void* gy_h16, const void* r_h16, const float* ln_w, float* dln_w, float* dln_b, int64_t rows, int32_t C,
                      float eps, void* stream);

/* Backward of osb_variance_embed: dx = dout*keep (optional), dw[c,j] += sum dout*keep*emb_scale*val[b,t+j-h],
 * db[c] += sum dout*keep*emb_scale. */
int osb_variance_embed_bwd(const float* dout, const float* val, const uint8_t* pad_mask, const float* emb_scale, float* dx, float* dw,
                           float* db, int32_t B, int32_t T, int32_t C, int32_t ksize, void* stream);

/* Backward of osb_embed_text: dtable[id] += sqrt(dim)*dout (padding row untouched), dscale += sum dout*pe. */
int osb_embed_text_bwd(const float* dout, const int64_t* ids, const float* inv_freq, float* dtable, float* dscale, int32_t B,
                       int32_t T, int32_t dim, int32_t n_vocab, int32_t padding_idx, void* stream);

/* osb_gemm_wgrad for a strided convolution (stride <= 4): dw[tap,n,k] += sum_{b,t<T} dy[b,t,n] * a[b, t*stride + tap - pad, k]
 * with a fp16 (B, T_in, lda).  The strided rows are a TMA traversal stride (no im2col copy).  Weight gradient of the period
 * discriminators' (5,1)/(3,1) convolutions (vocoder/wavenext/disc/_discriminators.py:52-58). */
int osb_gemm_wgrad_strided(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T, int32_t T_in,
                           int32_t N, int32_t K, int32_t taps, int32_t pad, int32_t stride, void* stream);

/* Batched form of osb_gemm_wgrad: dw[b, n, k] += sum_t dy[b, t, n] * a[b, t, k]  (A^T B per batch; dw is (B, N, K)). */
int osb_gemm_wgrad_batched(const void* dy, int64_t ldy, const void* a, int64_t lda, float* dw, int32_t B, int32_t T, int32_t N,
                           int32_t K, void* stream);

/* ---------------------------------------------------------------------------------------
 * Alignment-learning kernels (osb_align.cu) — the reference runs these on the host CPU with numba.
 * ------------------------------------------------------------------------------------- */

/* Monotonic alignment search per sample over log_p_attn[b, :m_len, :x_len] (fp32 (B,Tm,Tx)): float64 DP with the
 * reference's float32 first-row running sum and '>=' tie-break, back-track, bincount.
 * path (B,Tm) int32 = token of each frame (-1 past m_len), durations (B,Tx) fp32.  Bit-exact vs the reference.
 * Replaces _monotonic_alignment_search + the bincount of viterbi_decode (generator/alignments.py:177-235). */
int osb_mas(const float* log_p_attn, const int64_t* x_len, const int64_t* m_len, int32_t* path, float* durations, int32_t B,
            int32_t Tm, int32_t Tx, void* stream);

/* out[b,n] = mean(xs[b, start_n:end_n]) over the duration span of token n (0 for empty spans / pads).
 * Replaces average_by_duration (generator/alignments.py:242-280). */
int osb_average_by_duration(const float* ds, const float* xs, const int64_t* x_len, const int64_t* m_len, float* out, int32_t B,
                            int32_t Tm, int32_t Tx, void* stream);

/* ---------------------------------------------------------------------------------------
 * Optimizer over one flat fp32 bucket (osb_optim.cu).  Replaces, for the generator / discriminator parameter
 * groups, Lightning's clip_gradients(norm) + torch.optim.AdamW.step of BaseLightningModule.training_step
 * (optispeech/model/base_lightning_module.py:99-105,119-125; configs/model/optimizer/adamw.yaml).
 * ------------------------------------------------------------------------------------- */

/* stats[0] += sum g^2, stats[1] = 1 if a non-finite value was seen.  stats must be zeroed by the caller. */
int osb_grad_sumsq(const float* g, int64_t n, float* stats /*(2)*/, void* stream);

/* Fused: unscale by inv_scale, clip by global norm (max_norm <= 0 disables), decoupled-weight-decay Adam update with
 * bias correction for `step` (1-based).  Skipped entirely when stats[1] != 0 (non-finite gradient). */
/* Multi-tensor gather of one step's gradients into the flat bucket (what torch's AccumulateGrad + a flatten would do in ~100
 * launches).  table: device int64 [n_tensors][3] = {source pointer (0 = no gradient: zeros), destination offset in floats
 * (multiple of 4), numel}; chunks: device int32 [n_chunks][2] = {tensor index, chunk index}, 2048 floats per chunk.
 * stats (optional): stats[0] += sum of squares, stats[1] = 1 on a non-finite value — as osb_grad_sumsq. */
int osb_grad_gather(const int64_t* table, const int32_t* chunks, int64_t n_chunks, float* flat_g, float* stats, void* stream);

int osb_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, const float* stats, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int64_t step, float max_norm, float inv_scale, void* stream);

/* Same update with the per-step scalars read from device memory — hyper = [lr, 1 - beta1^step, sqrt(1 - beta2^step)] — so that the
 * launch can live inside a captured CUDA graph while the host keeps driving the LR schedule. */
int osb_adamw_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* stats, const float* hyper, float beta1,
                       float beta2, float eps, float weight_decay, float max_norm, float inv_scale, void* stream);

/* Beta-binomial alignment prior on the device: out[b,t,n] = log BetaBinomial(n; N_b, t+1, T_b-t) for t < T_b, n < N_b, -inf
 * elsewhere; log_factorial[m] = log(m!) in float64 for m <= Tm + Tx.  Replaces AlignmentModule._generate_prior
 * (generator/alignments.py:85-123: scipy.stats.betabinom per sample on the host). */
int osb_beta_binomial_prior(const double* log_factorial, int64_t table_len, const int64_t* x_len, const int64_t* m_len, float* out,
                            int32_t B, int32_t Tm, int32_t Tx, void* stream);

/* Forward-sum alignment loss and its gradient in one launch: per sample, CTC over the frames t < m_len of
 * log_softmax([blank_logit | log_p_attn[b,t,:x_len]]) with target 1..x_len, nll / x_len ('mean' reduction), 0 when infinite
 * (zero_infinity).  loss (B) holds the per-sample values (ForwardSumLoss = sum / B); grad (B,Tm,Tx) = d(sum/B)/d(log_p_attn);
 * alpha_ws is an fp32 workspace of 2*B*Tm*Tx + B*Tm + B values (token-state alpha and beta, per-frame normalisers, nll).  Replaces ForwardSumLoss.forward (generator/loss.py:150-194: a Python loop of
 * F.ctc_loss calls) and its autograd. */
int osb_forward_sum(const float* log_p_attn, const int64_t* x_len, const int64_t* m_len, float blank_logit, float* alpha_ws,
                    float* loss, float* grad, int32_t B, int32_t Tm, int32_t Tx, void* stream);

/* out[row] = sum_c x[row,c]^2 : the |f|^2 / |e|^2 terms of the pairwise distance (alignments.py:66-67). */
int osb_rownorm_sq(const float* x, float* out, int64_t rows, int32_t C, void* stream);

/* Backward of OSB_EPI_ATTN_LOGP with respect to the distance, as a matrix for two contractions:
 * Wn[b,t,n] = -(dscore/score), dscore = G - softmax*rowsum(G) (fp16, row stride ldw, zero padded);
 * neg_rsn[b,t] = -sum_n Wn; csn[b,n] += sum_t Wn.  Then dF = neg_rsn*F + Wn@E (OSB_EPI_AXPY) and
 * dE = -csn*E + Wn^T@F (osb_scale_rows + osb_gemm_wgrad_batched).  Replaces autograd through
 * torch.norm(feats - text) / log_softmax (alignments.py:66-74). */
int osb_attn_bwd_prep(const float* G, const float* lp, const float* prior, const float* lse, const int64_t* x_len, const int64_t* m_len,
                      void* wn_h16, float* neg_rsn, float* csn, int32_t B, int32_t Tm, int32_t Tx, int32_t ldw, void* stream);

/* out[b,c,t] = fp16(x[b,t,c]) (t < T), 0 for T <= t < Tp: per-batch transposed tensor-core operand. */
int osb_transpose_pack_h16(const float* x, void* out_h16, int32_t B, int32_t T, int32_t C, int32_t Tp, void* stream);

/* out[row,:] = sign * row_scale[row] * x[row,:] */
int osb_scale_rows(const float* x, const float* row_scale, float* out, int64_t rows, int32_t C, float sign, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused STFT-domain reconstruction losses (osb_spectral.cu): one CTA per frame, shared-memory FFT of the
 * prediction and the target together, loss partial sums; the backward variant adds d(loss)/d(x_hat).
 * ------------------------------------------------------------------------------------- */

/* One resolution of MultiResolutionSTFTLoss (disc/loss.py:123-142,197-270): centre/reflect framing, `window` (win values,
 * centred in n_fft), magnitudes sqrt(clamp(re^2+im^2, clamp_min)).  n_fft in {512, 1024, 2048}.
 *   forward  (dx_hat == NULL): stats[0] += sum (Ym-Xm)^2, stats[1] += sum Ym^2, stats[2] += sum |log Ym - log Xm|
 *   backward (dx_hat != NULL): dx_hat += coef[0] * d(sqrt(S1))*sqrt(S1).. i.e. with coef[0] = dL/dSC / (sqrt(S1) sqrt(S2)) and
 *                              coef[1] = dL/dMAG / count the kernel accumulates dL/dx_hat (fp32 atomics; caller zeroes dx_hat). */
int osb_stft_loss(const float* x_hat, const float* y, const float* window, int32_t B, int32_t L, int32_t n_fft, int32_t hop,
                  int32_t win, float clamp_min, double* stats, const float* coef, float* dx_hat, void* stream);

/* MelSpecReconstructionLoss (disc/loss.py:88-120): |STFT| (no clamp) -> fb (n_fft/2+1, n_mels) -> log(clip(., clamp_min)) -> L1.
 * klo/khi (n_mels): first/last non-zero bin of each filter; jlo/jhi (n_fft/2+1): first/last filter covering each bin.
 *   forward: stats[2] += sum |log mel_y - log mel_x| ; backward: coef[1] = dL/dMEL / count. */
int osb_mel_loss(const float* x_hat, const float* y, const float* window, const float* fb, const int32_t* klo, const int32_t* khi,
                 const int32_t* jlo, const int32_t* jhi, int32_t n_mels, int32_t B, int32_t L, int32_t n_fft, int32_t hop, int32_t win,
                 float clamp_min, double* stats, const float* coef, float* dx_hat, void* stream);

/* FastSpeech2Loss (optispeech/model/generator/loss.py:83-140) as the reference evaluates it (its masks broadcast inside
 * masked_select: see osb_loss.cu) and its gradients in one launch.  All inputs fp32 (B, Tx); x_len (B) int64.
 *   losses[0] = duration MSE in the log domain, losses[1] = pitch SmoothL1, losses[2] = energy SmoothL1 (weighted means)
 *   g_d / g_p / g_e (B, Tx) = d losses[0] / d d_hat, d losses[1] / d p_hat, d losses[2] / d e_hat */
int osb_fs2_losses(const float* d_hat, const float* p_hat, const float* e_hat, const float* ds, const float* p_tgt, const float* e_tgt,
                   const int64_t* x_len, float* losses, float* g_d, float* g_p, float* g_e, int32_t B, int32_t Tx, void* stream);

/* align_loss = forward-sum loss + bin loss (generator/__init__.py:174-175; bin loss: alignments.py:236-238) from the outputs
 * of osb_forward_sum (per-sample losses, gradient) and osb_mas (path):  out[0] = align_loss, out[1] = forward-sum part,
 * out[2] = bin part = -(1/B) sum_b mean_t log_p_attn[b,t,path[b,t]];  fs_grad (B,Tm,Tx) receives the bin-loss gradient
 * in place, so it becomes d(align_loss)/d(log_p_attn).  workspace: B + 1 fp32 words; the last word is a counter that has to
 * be ZERO on entry and is left zero (one CTA per sample; the last one to finish adds the partial sums in sample order). */
int osb_align_loss_fold(const float* log_p_attn, const int32_t* path, const int64_t* m_len, const float* per_sample_fs, float* fs_grad,
                        float* out /*(3)*/, float* workspace /*(B+1)*/, int32_t B, int32_t Tm, int32_t Tx, void* stream);

/* Every fp32 -> fp16 weight pack of a training step in ONE launch (the weights change every step, so the packs are
 * per-step work: 44 small launches otherwise).  jobs_dev: device array, sorted by first_elem (prefix sums of the
 * destination element counts, first_elem[0] = 0); total_elems = sum of destination elements.
 *   kind 0: src (rows, cols) fp32 row-major, optionally scaled per column -> dst (rows, dst_cols) fp16, zero padded
 *   kind 1: src Conv1d weight (rows = N, cols = Cin, k) -> dst (k, N, dst_cols) fp16 (one K-major matrix per tap)
 *   kind 2: dst (rows) fp32 = aux (rows) + src (rows, cols) @ col_scale (cols)   (counts as 8 * rows destination elements):
 *           the pwconv1 bias with the LayerNorm bias folded in, b1 + W1 @ ln_b (modules/convnext.py:38-39)
 * dst_cols and every first_elem are multiples of 8 (a thread writes 8 destination elements with one 16-byte store). */
typedef struct osb_pack_job {
  const void* src;
  void* dst;
  const void* col_scale;   /* kind 0: (cols) fp32 or NULL; kind 2: the vector */
  const void* aux;         /* kind 2: (rows) fp32 addend or NULL */
  int64_t first_elem;
  int32_t kind, rows, cols, k, dst_cols, reserved;
} osb_pack_job;
int osb_pack_multi(const osb_pack_job* jobs_dev, int32_t n_jobs, int64_t total_elems, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-head self-attention of the Transformer backbone (osb_mha.cu), d_k = 128.
 * q, k, v: fp16 (B, T, ld_qkv) with head h in columns [h*128, +128) (typically three column slices of one fused
 * (B, T, 3*H*128) projection output).  kv_len (B) or NULL: keys >= kv_len[b] are masked (their probability is exactly 0);
 * query rows are never masked.  Counter-based dropout on the probabilities (element ((b*H+h)*T+t)*T+key of `seed`).
 * Replaces MultiHeadedAttention.forward / forward_attention up to (not including) linear_out
 * (optispeech/model/generator/modules/_transformer/attention.py:84-125).
 * ------------------------------------------------------------------------------------- */

/* ctx[b,t,h*128+d] = sum_key softmax_key(q.k * scale)[key] * D[key] * v[key,d]  as fp16 (ld_ctx); ctx_lo_off > 0 also writes the
 * fp16 rounding residual at column offset ctx_lo_off (split-precision operand for linear_out).  row_max / row_inv_l (B,H,T)
 * fp32, optional (both or neither): raw row maximum and 1/sum(exp) for the backward pass. */
int osb_mha_fwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const int64_t* kv_len, void* ctx, int64_t ld_ctx,
                int64_t ctx_lo_off, float* row_max, float* row_inv_l, int32_t B, int32_t T, int32_t H, int32_t d_k, float scale,
                float dropout_p, uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* stream);

/* Backward of osb_mha_fwd with respect to q (complete) and, as fp16 matrices for two further contractions, to k and v:
 *   dq[b,t,h*128+d]       = sum_key dS[t,key] k[key,d]                       (fp16, ld_dq)
 *   ds_out[b,t,h*Tp+key]  = dS = P o (dP o D - delta) * scale,  dP = d_ctx v^T, delta = <d_ctx_row, ctx_row>
 *   pd_out[b,t,h*Tp+key]  = P o D                                            (both fp16, ld_p >= H*Tp, Tp >= T, Tp % 8 == 0;
 *                                                                              columns [T, Tp) are zero)
 * then dk_h = ds_out_h^T q_h and dv_h = pd_out_h^T d_ctx_h through osb_gemm_wgrad_batched. */
int osb_mha_bwd(const void* q, const void* k, const void* v, int64_t ld_qkv, const int64_t* kv_len, const void* ctx, int64_t ld_ctx,
                const void* d_ctx, int64_t ld_dctx, const float* row_max, const float* row_inv_l, void* dq, int64_t ld_dq,
                void* ds_out, void* pd_out, int64_t ld_p, int32_t Tp, int32_t B, int32_t T, int32_t H, int32_t d_k, float scale,
                float dropout_p, uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* stream);

/* dst[b,t, h*128 + d] = fp16(src[h,b,t,d]): gathers per-head fp32 (H,B,T,128) gradients into a channels-last fp16 operand. */
int osb_mha_pack_heads(const float* src, void* dst_h16, int64_t ld_dst, int32_t B, int32_t T, int32_t H, void* stream);

/* dst[i] = fp16(src[i] * D(i)), i over rows*N elements, D the counter-based dropout scale of `seed` (p = 0: plain cast).
 * The gradient entering a branch whose forward output was dropped out by OSB_EPI_RESID (same seed, same element order). */
int osb_dropout_pack_h16(const float* src, void* dst_h16, int64_t rows, int32_t N, float dropout_p, uint64_t dropout_seed,
                         const uint64_t* dropout_seed_dev, void* stream);

/* out[b,t,:] = (x[b,t,:] + alpha[0] * pe[t,:]) * D : ScaledPositionalEncoding.forward and its dropout
 * (modules/_transformer/embedding.py:111-124); pe (T, C) fp32 table, alpha device scalar.  pe = alpha = NULL applies the
 * dropout mask only (the backward pass: dx = dout * D). */
int osb_add_posenc(const float* x, const float* pe, const float* alpha, float* out, int32_t B, int32_t T, int32_t C, float dropout_p,
                   uint64_t dropout_seed, const uint64_t* dropout_seed_dev, void* stream);

/* ---------------------------------------------------------------------------------------
 * Index / mask glue (osb_glue.cu)
 * ------------------------------------------------------------------------------------- */
/* valid[b,t] = t < lengths[b], pad[b,t] = !valid (bytes; either output may be NULL).
 * Replaces sequence_mask / make_non_pad_mask / make_pad_mask (optispeech/utils/model.py:12-21). */
int osb_sequence_mask(const int64_t* lengths, uint8_t* valid, uint8_t* pad, int32_t B, int32_t T, void* stream);

/* start[b] = (int64) (rand[b] * max((float)(lengths[b] - margin) - S, 0)): the segment start draw of
 * get_random_segments (optispeech/utils/segments.py:29-35; margin = 4 at generator/__init__.py:147-153). */
int osb_segment_starts(const float* rand, const int64_t* lengths, int64_t* start, int32_t B, int32_t margin, int32_t S, void* stream);

/* out[b,s,:] = x[b, start[b]*scale + s, :] for s < S (zero outside [0,T)); x (B,T,C) fp32 channels-last, C = 1 for waveforms.
 * Replaces get_segments / get_segments_numpy (optispeech/utils/segments.py:38-72; scale = hop_length for the ground-truth
 * waveform crop of base_lightning_module.py:38-43). */
int osb_gather_segments(const float* x, const int64_t* start, float* out, int32_t B, int64_t T, int32_t C, int32_t S, int32_t scale,
                        void* stream);

/* ---------------------------------------------------------------------------------------
 * ConvNeXt block, training path (osb_convnext.cu / osb_convnext_bwd.cu)
 * ------------------------------------------------------------------------------------- */
/* osb_convnext_block_fwd that also emits what the backward needs while it is on chip: xhat (fp16 normalised dwconv output,
 * the A operand of pwconv1), rstd (fp32 per position), pre (fp16 GELU argument) and h (fp16 GELU output).  Same arguments and
 * result otherwise.  Replaces ConvNeXtBlock.forward under autograd (modules/convnext.py:34-47). */
int osb_convnext_block_fwd_train(const float* x, const float* dw_w, const float* dw_b, const void* w1f_h16, const float* b1f,
                                 const void* w2_h16, const float* b2, const float* gamma, const float* row_scale,
                                 const uint8_t* pad_mask, float* out, void* xhat_h16, float* rstd, void* pre_h16, void* h_h16, int32_t B,
                                 int32_t T, int32_t C, int32_t I, float eps, void* stream);

/* Both data-gradient contractions of a ConvNeXt block in one tcgen05 kernel (the autograd of modules/convnext.py:39-46):
 *   dyg   = fp16(dout * gamma * keep * row_scale[b])                     (B,T,C)  dy of the pwconv2 weight gradient
 *   dh    = fp16((dyg . W2) * gelu_erf'(pre))                            (B,T,I)  dy of the pwconv1 weight gradient
 *   dxhat = dh . W1f   (fp32)                                      (parts,B,T,C)  gradient wrt the normalised dwconv output, as
 *           `parts` = osb_convnext_block_bwd_parts(B, T, I) partial sums (the intermediate dimension is split over CTAs when
 *           there are few row tiles; osb_ln_dwconv_bwd adds them up)
 * w2_h16 (C, I) and w1f_h16 (I, C) are the FORWARD fp16 packs (read as MN-major B operands).  (C, I) = (256,1024) or (384,1152). */
int osb_convnext_block_bwd_parts(int32_t B, int32_t T, int32_t I);
int osb_convnext_block_bwd(const float* dout, const float* gamma, const float* row_scale, const uint8_t* pad_mask, const void* pre_h16,
                           const void* w2_h16, const void* w1f_h16, void* dyg_h16, void* dh_h16, float* dxhat, int32_t B, int32_t T,
                           int32_t C, int32_t I, void* stream);

/* LayerNorm backward (affine folded away) + depthwise conv7 backward + residual path + depthwise parameter gradients in one pass:
 *   dd = LN_bwd(dxhat; xhat, rstd);  dx[t] = dout[t] * keep[t] + sum_j w[:,j] * dd[t-j+3]
 *   dparam (8, C) fp32 += [ddw[:,0] | ... | ddw[:,6] | ddb]   (tap-major; accumulated: zero it first)
 * Autograd of nn.Conv1d(groups=C) + nn.LayerNorm (modules/convnext.py:36-38). */
int osb_ln_dwconv_bwd(const float* dxhat /*(nparts,B,T,C)*/, int32_t nparts, const void* xhat_h16, const float* rstd, const float* dout,
                      const float* x, const float* dw_w, const uint8_t* pad_mask, float* dx, float* dparam, int32_t B, int32_t T,
                      int32_t C, void* stream);

/* Layer-scale and pwconv2-bias gradients from the block's input x and output out (gamma * z * rs = out - x on unmasked rows):
 *   dgamma += sum_rows dout * keep * (out - x) / gamma ;  db2 += sum_rows dout * keep * rs * gamma   (modules/convnext.py:42-46) */
int osb_resid_param_grad(const float* dout, const float* out, const float* x, const float* gamma, const uint8_t* pad_mask,
                         const float* row_scale, float* dgamma, float* db2, int64_t rows, int32_t T, int32_t C, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-period discriminator (osb_disc.cu) — reference vocoder/wavenext/disc/_discriminators.py:41-97
 *
 * Layout.  A period discriminator sees the waveform as `period` interleaved sequences x_j[l] = wav[l*period + j] (after the
 * reflect padding of the tail to a multiple of the period, :66-70).  All NSEQ = NS*period sequences of a layer share one flat
 * fp16 matrix (NSEQ * P rows, C columns): a sequence owns P consecutive rows, its L valid rows first, zeros after.  With
 * P_5 = P_4 = L_4 + 2 and P_i = 3 * P_(i+1) the zero tail doubles as convolution padding and a layer is one osb_gemm call with
 * row_stride 3 (or 1), OSB_FLAG_LRELU and a keep mask on the gap rows.
 * fp16 gradient tensors carry the caller's loss scale times an extra power of two; entry points producing fp32 gradients take
 * `inv_scale` to remove it.
 * ------------------------------------------------------------------------------------- */
/* Layer 1 (Conv2d(1, 32, (5,1), (stride,1), padding (2,0)) + LeakyReLU, :52): out (NS*period*P1, CP) fp16, columns >= 32 zero. */
int osb_mpd_first_fwd(const float* wav /*(NS,T)*/, const float* w /*(32,5)*/, const float* bias, void* out_h16, int32_t NS, int32_t T,
                      int32_t period, int32_t L1, int32_t P1, int32_t CP, int32_t stride, float slope, void* stream);
/* Its backward from the gated gradient g (same layout as out): dwav (NS,T) += (reflected positions fold back), dw (32,5) +=,
 * db (32) +=; any of dwav / (dw, db) may be NULL. */
int osb_mpd_first_bwd(const void* g_h16, const float* wav, const float* w, float* dwav, float* dw, float* db, int32_t NS, int32_t T,
                      int32_t period, int32_t L1, int32_t P1, int32_t CP, int32_t stride, float inv_scale, void* stream);
/* conv_post (Conv2d(1024, 1, (3,1), padding (1,0)), :62): score (NS, L*period) fp32 in the reference's flatten order l*period + j. */
int osb_mpd_post_fwd(const void* x_h16, const float* w /*(C,3)*/, const float* bias, float* score, int32_t NSEQ, int32_t period, int32_t L,
                     int32_t P, int32_t C, void* stream);
/* dx (NSEQ*P, C) fp16 = scale * conv_post^T(dscore); dw (C,3) +=, db (1) += (unscaled; NULL to skip either part). */
int osb_mpd_post_bwd(const float* dscore, const void* x_h16, const float* w, void* dx_h16, float* dw, float* db, int32_t NSEQ,
                     int32_t period, int32_t L, int32_t P, int32_t C, float scale, void* stream);
/* g = dy * (y > 0 ? 1 : slope) on the valid rows (row % P < L), 0 on the gap rows — LeakyReLU backward from the saved output.
 * colsum (optional, (C) fp32, accumulated): += colsum_scale * column sums of g = the layer's bias gradient. */
int osb_lrelu_bwd_h16(const void* dy, const void* y, void* g, int64_t rows, int32_t C, int32_t P, int32_t L, float slope, float* colsum,
                      float colsum_scale, void* stream);
/* Data gradient of a strided convolution from the per-tap products of one GEMM: col (rows_out, taps*C) -> dx (rows_in, C),
 * dx[r] = sum_{tap: (r + pad - tap) % stride == 0} col[(r + pad - tap) / stride, j(tap)*C : +C], j(tap) = taps-1-tap if `reversed`. */
int osb_col2im_h16(const void* col, void* dx, int64_t rows_in, int64_t rows_out, int32_t C, int32_t taps, int32_t pad, int32_t stride,
                   int32_t reversed, void* stream);
/* Feature-matching term (disc/loss.py:67-85): out_sum[0] += sum |a - b| over n fp16 elements; and its gradient with respect
 * to b: db = coef[0] * scale * sign(b - a). */
int osb_l1_pair_fwd(const void* a_h16, const void* b_h16, float* out_sum, int64_t n, void* stream);
int osb_l1_pair_bwd(const void* a_h16, const void* b_h16, const float* coef, float scale, void* db_h16, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Multi-resolution discriminator (osb_disc.cu) — reference vocoder/wavenext/disc/_discriminators.py:139-216
 *
 * Conv2d stacks over magnitude spectrograms (NS, F, W).  Same flat fp16 layout as the period discriminators with the
 * (signal, frame) pairs as sequences and frequency along the rows: row = (n*W_i + w)*P_i + h, h < H_i valid, 64 channels;
 * P_5 = H_5 + 2, P_i = 2*P_(i+1).  The kw taps along the frame axis are gathered into the contraction (osb_wim2col_h16,
 * K = kw*64), the kh taps along frequency are implicit-GEMM taps of osb_gemm with row_stride 2.
 * ------------------------------------------------------------------------------------- */
/* Layer 1 (Conv2d(1, 64, (7,5), (2,2), (3,2)) + LeakyReLU, :154): spec (NS, F, W) fp32 -> (NS*W1*P1, 64) fp16. */
int osb_mrd_first_fwd(const float* spec, const float* w /*(64,35)*/, const float* bias, void* out_h16, int32_t NS, int32_t F, int32_t W,
                      int32_t H1, int32_t W1, int32_t P1, float slope, void* stream);
/* Backward from the gated gradient: dspec (NS,F,W) written; dw (64,35) +=, db (64) +=; either part may be NULL. */
int osb_mrd_first_bwd(const void* g_h16, const float* spec, const float* w, float* dspec, float* dw, float* db, int32_t NS, int32_t F,
                      int32_t W, int32_t H1, int32_t W1, int32_t P1, float inv_scale, void* stream);
/* Layer 1 as a GEMM (the path the host takes): xcol (NS*W1*P1, 64) fp16 with xcol[(n, w1, h1), kh*5 + kw] =
 * spec[n, 2 h1 + kh - 3, 2 w1 + kw - 2] (taps 35..63, out-of-range positions and gap rows zero); and the adjoint gather
 * dspec[n, f, t] = inv_scale * sum col[(n, w1, h1), kh*5 + kw] over the taps that touch (f, t), col = g . W1 (fp16). */
int osb_spec_im2col_h16(const float* spec, void* xcol, int32_t NS, int32_t F, int32_t W, int32_t H1, int32_t W1, int32_t P1, void* stream);
int osb_spec_col2im(const void* col_h16, float* dspec, int32_t NS, int32_t F, int32_t W, int32_t H1, int32_t W1, int32_t P1,
                    float inv_scale, void* stream);
/* xcol[(n, wo, h), kw*C + c] = x[(n, wo*sw + kw - pw, h), c] (zero outside [0, W_in)), and its adjoint. */
int osb_wim2col_h16(const void* x, void* xcol, int32_t NS, int32_t W_in, int32_t W_out, int32_t P, int32_t C, int32_t KW, int32_t pw,
                    int32_t sw, void* stream);
int osb_wcol2im_h16(const void* dxcol, void* dx, int32_t NS, int32_t W_in, int32_t W_out, int32_t P, int32_t C, int32_t KW, int32_t pw,
                    int32_t sw, void* stream);
/* conv_post (Conv2d(64, 1, (3,3), padding 1), :163): score (NS, H*W) fp32 in the reference's flatten order h*W + w; backward:
 * dx (NS*W*P, 64) fp16 = scale * conv_post^T(dscore); dw (64,9) +=, db (1) += (unscaled). */
int osb_mrd_post_fwd(const void* x_h16, const float* w /*(64,9)*/, const float* bias, float* score, int32_t NS, int32_t W, int32_t H,
                     int32_t P, void* stream);
int osb_mrd_post_bwd(const float* dscore, const void* x_h16, const float* w, void* dx_h16, float* dw, float* db, int32_t NS, int32_t W,
                     int32_t H, int32_t P, float scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * Feature extraction of the data path (osb_spectral.cu) — reference dataset/feature_extractors/__init__.py:110-200
 * ------------------------------------------------------------------------------------- */
/* log-mel spectrogram and frame energy of a ragged batch of waveforms, one launch: reflect padding by (n_fft - hop)/2 with
 * every utterance's OWN length, no centring, Hann `window`, mag = sqrt(re^2 + im^2 + mag_eps);
 *   mel[b, j, f] = log(max(sum_k fb[j, k] mag[f, k], clip_val)),  energy[b, f] = ||mag[f, :]||_2,  f < lengths[b] / hop, else 0.
 * fb is (n_mels, n_fft/2+1) row-major with the non-zero bin range [klo[j], khi[j]] of every filter.  n_fft in {512, 1024, 2048}.
 * Replaces CommonFeatureExtractor.get_mel (:151-200) and FeatureExtractor.get_energy (:114-146), which run per utterance. */
int osb_mel_energy(const float* wav, const int64_t* lengths, const float* window, const float* fb, const int32_t* klo, const int32_t* khi,
                   float* mel, float* energy, int32_t B, int32_t Lmax, int32_t Fmax, int32_t n_mels, int32_t n_fft, int32_t hop,
                   int32_t win, float mag_eps, float clip_val, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSB200_H_ */
