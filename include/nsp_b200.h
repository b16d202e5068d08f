/*
 * nsp_b200 -- C ABI of the B200-native (sm_100a) speech-encoder + CTC / RNN-T hot path of
 * hirofumi0810/neural_sp.  Plain pointers and sizes only: no torch types cross this boundary.
 *
 * The reference has no FFI of its own (it is pure PyTorch); the narrowest cut-lines its code
 * offers for this path are Python call sites, and every entry point below names the reference
 * call (file:line under /root/reference) it replaces.  INTEGRATION.md shows the ctypes stubs a
 * maintainer of the reference would add at those call sites.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is
 *    enqueued asynchronously on it, nothing synchronises unless stated;
 *  - tensors are dense row-major unless explicit element strides are given;
 *  - return value: NSP_OK or an error code; nsp_last_error() returns a thread-local message;
 *  - inputs are borrowed for the duration of the enqueued work, outputs/workspaces are
 *    caller-allocated (query sizes with the *_workspace_bytes functions).
 */
#ifndef NSP_B200_H_
#define NSP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    NSP_OK = 0,
    NSP_ERR_INVALID = 1,     /* bad argument (shape, alignment, null pointer) */
    NSP_ERR_CUDA = 2,        /* a CUDA runtime/driver call failed */
    NSP_ERR_UNSUPPORTED = 3, /* valid request outside what the sm_100a kernels cover */
    NSP_ERR_NO_DEVICE = 4    /* no sm_100 device is visible */
} nsp_status;

/* Library ABI version (major*100 + minor). */
int nsp_version(void);
/* Thread-local message describing the last non-OK status returned on this thread. */
const char* nsp_last_error(void);
/* NSP_OK iff the current CUDA device is compute capability 10.x (B200); fills sm_count. */
nsp_status nsp_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * CTC loss, forward + backward fused (HBM-bound; fp32).
 *
 * Replaces   CTC.loss_fn            neural_sp/models/seq2seq/decoders/ctc.py:139-150
 *            (logits.log_softmax(2) -> nn.CTCLoss(reduction="sum", zero_infinity=True) -> / B)
 *   and      kldiv_lsm_ctc          neural_sp/models/criterion.py:110-127
 *            (mixed as loss*(1-lsm) + kl*lsm at ctc.py:128-129)
 *   including their autograd backward: grad = d(loss)/d(logits).
 *
 * logits : fp32 [B, T, V] addressed as logits[b*stride_b + t*stride_t + v]  (the reference
 *          passes the [T,B,V] transpose view of a batch-major tensor; both layouts work).
 * labels : int32 [B, Lmax] padded (padding value ignored); ylens int32 [B]; elens int32 [B]
 *          (valid frames, <= T).  blank is the blank id (0 in the reference, speech2text.py:66).
 * nll    : fp32 [B] out, -log p(y_b | x_b), 0 for infeasible utterances (zero_infinity).
 * loss   : fp32 [1] out, (1-lsm) * sum_b nll_b / B  +  lsm * KL.
 * grad   : fp32 [B, T, V] dense out, d loss / d logits (rows t >= elens[b] are zero).
 * ------------------------------------------------------------------------------------------ */
size_t nsp_ctc_loss_workspace_bytes(int B, int T, int Lmax);
nsp_status nsp_ctc_loss_fwd_bwd(const float* logits, int64_t stride_b, int64_t stride_t,
                                int B, int T, int V,
                                const int32_t* labels, int Lmax,
                                const int32_t* elens, const int32_t* ylens,
                                int blank, float lsm_prob,
                                float* nll, float* loss, float* grad,
                                void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * CTC forced alignment (bit-exact integer output).
 *
 * Replaces   CTCForcedAligner.__call__ / align   decoders/ctc.py:632-753
 * logits fp32 [B, T, V] dense (not modified; the reference masks a clone, ctc.py:131,651).
 * trigger_points int32 [B, Lmax+1] out (zero-initialised by the call).
 * ------------------------------------------------------------------------------------------ */
size_t nsp_ctc_align_workspace_bytes(int B, int T, int Lmax);
nsp_status nsp_ctc_forced_align(const float* logits, int B, int T, int V,
                                const int32_t* labels, int Lmax,
                                const int32_t* elens, const int32_t* ylens, int blank,
                                int32_t* trigger_points,
                                void* workspace, size_t workspace_bytes, void* stream);


/* ------------------------------------------------------------------------------------------
 * Dense projection on the tcgen05 tensor cores with a fused epilogue (tensor-pipe bound):
 *     out = residual + alpha * act(x[M,K] * w[N,K]^T + bias)
 *
 * Replaces every nn.Linear / kernel-size-1 nn.Conv1d of the encoder path:
 *   PositionwiseFeedForward.w_1/w_2            modules/positionwise_feed_forward.py:47-48,89
 *   RelMHA w_key/w_value/w_query/w_pos/w_out   modules/relative_multihead_attention.py:57-60,169-176,217
 *   ConformerConvBlock.pointwise_conv1 (+F.glu) / pointwise_conv2   modules/conformer_convolution.py:44-69,110-126
 *   ConvEncoder.bridge encoders/conv.py:87,193 ; CTC.output decoders/ctc.py:81-91,124
 *
 * prec  NSP_PREC_BF16: x, w are bf16.  NSP_PREC_TF32: x, w are fp32, one tf32 pass (~1e-3).
 *       NSP_PREC_FP32: x/x_lo and w/w_lo are the hi/lo halves produced by nsp_split_tf32; three
 *       tf32 passes accumulate hi*hi + lo*hi + hi*lo in TMEM (fp32-level accuracy; parity mode).
 * x [M,K] row pitch ldx, w [N,K] row pitch ldw (elements; pitches must be 16-byte multiples).
 * glu=1: w holds value rows [0,N/2) and gate rows [N/2,N); out[:, j] = v_j * sigmoid(g_j), width N/2.
 * act: 0 none, 1 relu, 2 swish (x*sigmoid(x)), 3 gelu (erf), 4 gelu (tanh approximation); bias fp32 [N] or NULL; residual fp32 [M,ldr] or NULL.
 * out: fp32 (out_bf16=0) or bf16 (out_bf16=1) with row pitch ldo; out2 (optional, may be NULL) receives a
 * bf16 copy of an fp32 result with pitch ldo2.  out may alias residual.
 * ------------------------------------------------------------------------------------------ */
typedef enum { NSP_PREC_BF16 = 0, NSP_PREC_TF32 = 1, NSP_PREC_FP32 = 2 } nsp_precision;
nsp_status nsp_linear_fwd(int prec, const void* x, const void* x_lo, int64_t ldx,
                          const void* w, const void* w_lo, int64_t ldw,
                          int M, int N, int K, int glu, int act,
                          const float* bias, const float* residual, int64_t ldr, float alpha,
                          void* out, int64_t ldo, int out_bf16, void* out2, int64_t ldo2, void* stream);
/* hi = round-to-nearest tf32(x), lo = x - hi (both fp32, n elements). */
nsp_status nsp_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
/* fp32 -> bf16 (round to nearest even), n elements. */
nsp_status nsp_cast_f32_to_bf16(const float* x, void* y_bf16, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------
 * LayerNorm over the last dimension (HBM-bound).  y = (x*in_scale - mean) * rstd * gamma + beta.
 * Replaces nn.LayerNorm at encoders/conformer_block.py:53,58,70,76,80 and encoders/transformer.py:600.
 * x fp32 [M, D] (pitch ldx); y fp32 [M, D] and/or y_bf16 bf16 [M, D] (either may be NULL, not both).
 * ------------------------------------------------------------------------------------------ */
nsp_status nsp_layernorm_fwd(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps,
                             float in_scale, float* y, int64_t ldy, void* y_bf16, int64_t ldyb, int M, int D,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * Relative-position multi-head self-attention, flash-style (scores never reach HBM).
 * Replaces RelativeMultiheadAttentionMechanism.forward  modules/relative_multihead_attention.py:146-220
 *   (between the input projections and w_out), _rel_shift :112-144, MultiheadAttentionMechanism.forward
 *   modules/multihead_attention.py:93-157 (r == NULL), and the masks of encoders/transformer.py:633-686.
 * q [B*Tq, >=H*dk], k,v [B*Tk, ...] element (b*T + t, h*dk + c), row pitches ldq/ldk/ldv; all bf16
 * (is_bf16=1) or all fp32.  r: projected position table [rlen, H*dk] (row = relative distance) or NULL.
 * e[b,h,i,j] = ((q_i+u_h).k_j + (q_i+v_h).r[min(|Tk-Tq+i-j|, clamp_len)]) / sqrt(dk); keys j >= klens[b]
 * (and causal / chunk-wise exclusions) get finfo.min before the softmax; out[b*Tq+i, h*dk+c] = sum_j aw v.
 * causal: key visible iff j <= Tk-Tq+i+lookahead.  chunk_c>0: keys restricted to
 * [chunk_start-chunk_l, chunk_start+chunk_c) of the query's chunk (make_chunkwise_san_mask).
 * ------------------------------------------------------------------------------------------ */
nsp_status nsp_relpos_attention_fwd(int is_bf16, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                    const void* v, int64_t ldv, const void* r, int64_t ldr, int rlen,
                                    const float* u_bias, const float* v_bias, const int32_t* klens,
                                    void* out, int64_t ldo, int B, int H, int Tq, int Tk, int dk,
                                    int clamp_len, int causal, int lookahead, int chunk_c, int chunk_l,
                                    void* stream);

/* ------------------------------------------------------------------------------------------
 * Conformer convolution module core: y = Swish(Norm(depthwise_conv1d_k(x) + bias)), x,y [B,T,d].
 * Replaces ConformerConvBlock.forward  modules/conformer_convolution.py:113-124 (depthwise Conv1d,
 * causal trim, LayerNorm | BatchNorm1d(eval) | GroupNorm(d/2 groups), Swish).  w: taps TRANSPOSED to [k,d].
 * norm_mode 0 LayerNorm(eps) over d, 1 BatchNorm with running stats, 2 GroupNorm (2 channels/group).
 * ------------------------------------------------------------------------------------------ */
nsp_status nsp_conformer_conv_fwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias,
                                  int norm_mode, const float* norm_w, const float* norm_b,
                                  const float* run_mean, const float* run_var, float eps,
                                  void* y, int64_t ldy, int B, int T, int d, int k, int causal, void* stream);

/* Training of the convolution module with BatchNorm (the reference's default `conformer_normalization`,
 * modules/conformer_convolution.py:118-124: statistics over all B*T frames):
 *   nsp_dwconv_stats_fwd: z = depthwise_conv1d_k(x) + bias (same dtype as x, [B,T,d]); stats fp32 [2,d] = (sum z, sum z^2);
 *     the normalised output then comes from nsp_conformer_conv_fwd(norm_mode 1) with run_mean / run_var = the batch statistics;
 *   nsp_bn_swish_bwd: dy = gradient w.r.t. Swish(gamma zh + beta), zh = (z - mean) rsqrt(var + eps):
 *     sums fp32 [2,d] = (sum du = d beta, sum du zh = d gamma), dz = gamma rstd (du - sums0 / M - zh sums1 / M), M = B*T rows;
 *   nsp_dwconv_bwd: dx, dw [k,d] (+=), dbias [d] (+=) from dz (the depthwise half of nsp_conformer_conv_bwd). */
nsp_status nsp_dwconv_stats_fwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias, void* z,
                                int64_t ldz, float* stats, int B, int T, int d, int k, int causal, void* stream);
nsp_status nsp_bn_swish_bwd(int is_bf16, const void* z, int64_t ldz, const void* dy, int64_t lddy, const float* mean,
                            const float* var, const float* gamma, const float* beta, float eps, float* sums,
                            void* dz, int64_t lddz, int64_t M, int d, void* stream);
/* The same backward WITHOUT an activation inside (du = gradient w.r.t. gamma zh + beta): nn.BatchNorm2d of the CNN front-end's
 * Conv2dBlock in training (encoders/conv.py:362-394) on channels-last activations, M = B*T*F rows, d = channels; the ReLU /
 * pooling mask is applied before (nsp_relu_mask / nsp_maxpool2d_relu_bwd), the statistics come from nsp_dwconv_stats_fwd
 * with k = 1.  sums fp32 [2,d] = (d beta, d gamma). */
nsp_status nsp_bn_bwd(int is_bf16, const void* z, int64_t ldz, const void* du, int64_t lddu, const float* mean,
                      const float* var, const float* gamma, float eps, float* sums, void* dz, int64_t lddz,
                      int64_t M, int d, void* stream);
/* GroupNorm(d/2 groups = channel pairs) variant of the same module (conformer_convolution.py:47-48): per-frame pair statistics;
 * dz = gradient w.r.t. z, dgamma / dbeta fp32 [d] are accumulated (+=). */
nsp_status nsp_gn2_swish_bwd(int is_bf16, const void* z, int64_t ldz, const void* dy, int64_t lddy, const float* gamma,
                             const float* beta, float eps, void* dz, int64_t lddz, float* dgamma, float* dbeta,
                             int64_t M, int d, void* stream);
nsp_status nsp_dwconv_bwd(int is_bf16, const void* x, int64_t ldx, const float* w, const void* dz, int64_t lddz,
                          void* dx, int64_t lddx, float* dw, float* dbias, int B, int T, int d, int k, int causal,
                          void* stream);

/* x *= a in place (LayerDrop's eval-time 1/(1-p) rescale, encoders/conformer_block.py:122-126). */
nsp_status nsp_scale_inplace(float* x, float a, int64_t n, void* stream);
/* x[b,t,:] = x[b,t,:] * a + pe[t,:] in place: PositionalEncoding.forward (pe_type='add')
 * modules/positional_embedding.py:82-90; x fp32 [B,T,D], pe fp32 [T,D] = rows offset..offset+T of the sinusoid buffer. */
nsp_status nsp_add_pos_enc(float* x, const float* pe, float a, int B, int T, int D, void* stream);
/* SpecAugment masking in place (frontends/spec_augment.py:112-140 `xs[:, :, f0:f1] = 0`, `xs[:, t0:t1] = 0`): x fp32 [B,T,F];
 * freq_rects / time_rects are HOST arrays of (begin, end) int32 pairs (at most 32 each), passed to the kernel by value. */
nsp_status nsp_mask_rects(float* x, int B, int T, int F, const int32_t* freq_rects, int n_freq,
                          const int32_t* time_rects, int n_time, void* stream);
/* y[n] = sum_m x[m, n] for a dense fp32 [M, N] matrix (bias gradients). */
nsp_status nsp_colsum(const float* x, float* y, int M, int N, void* stream);
/* TransformerXL sinusoid table: XLPositionalEmbedding.forward modules/positional_embedding.py:135-138.
 * table fp32 [rows, d]; row r is position -(r+1) = relative distance r: cat(sin(pos*inv_freq), cos(pos*inv_freq)). */
nsp_status nsp_xl_pos_table(const float* inv_freq, float* table, int rows, int d, void* stream);

/* ------------------------------------------------------------------------------------------
 * Convolutional front-end on channels-last activations [B, T, F, C].
 * Replaces Conv2dBlock.forward encoders/conv.py:347-396 and the view/transposes of ConvEncoder.forward :181-189.
 * nsp_conv3x3_relu_fwd: y = relu?(conv3x3(x, w, pad 1, stride 1) + bias); w is nn.Conv2d weight [CO, CI, 3, 3];
 *   in_chmajor=1 reads the raw feature layout [B, T, CI, F] (conv.py:183).  in/out fp32 or bf16.
 * nsp_maxpool2d_fwd: max-pool kernel=stride=(pool_t, pool_f), ceil_mode (conv.py:330-337); out_chmajor=1 writes
 *   [B, T', C*F'] with index c*F'+f, the flatten of conv.py:189; f_keep>0 keeps only the first f_keep bins.
 * nsp_maxpool_time_fwd: MaxPoolSubsampler  encoders/subsampling.py:175-209 on [B, T, D] -> [B, ceil(T/f), D].
 * ------------------------------------------------------------------------------------------ */
nsp_status nsp_conv3x3_relu_fwd(int in_bf16, int out_bf16, const void* x, int in_chmajor, const float* w,
                                const float* bias, void* y, int B, int T, int F, int CI, int CO, int relu,
                                void* stream);
nsp_status nsp_maxpool2d_fwd(int in_bf16, int out_bf16, const void* x, void* y, int B, int T, int F, int C,
                             int pool_t, int pool_f, int f_keep, int out_chmajor, void* stream);
nsp_status nsp_maxpool_time_fwd(int is_bf16, const void* x, void* y, int B, int T, int D, int factor, void* stream);
/* Same with mode: 0 max, 1 mean of the in-range frames (MeanPoolSubsampler subsampling.py:212-246), 2 first frame
 * (DropSubsampler :97-126), 3 sum (AddSubsampler :129-172). */
nsp_status nsp_pool_time_fwd(int is_bf16, const void* x, void* y, int B, int T, int D, int factor, int mode, void* stream);

/* ------------------------------------------------------------------------------------------
 * RNN-Transducer loss, forward + backward (HBM-bound; fp32), blank-first lattice.
 * Replaces  warp_rnnt.rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, average_frames=False,
 *           reduction='mean', gather=False)     decoders/rnn_transducer.py:248-252
 *           (and warprnnt_pytorch.RNNTLoss, :254-256) including its gradient w.r.t. log_probs.
 * log_probs fp32 [B, T, U1, V] dense (U1 = Umax + 1); labels int32 [B, U1-1]; flens, ylens int32 [B].
 * nll fp32 [B]; loss fp32 [1] = mean_b nll_b; grad fp32 [B, T, U1, V] = d loss / d log_probs (may be NULL).
 * ------------------------------------------------------------------------------------------ */
size_t nsp_rnnt_loss_workspace_bytes(int B, int T, int U1);
nsp_status nsp_rnnt_loss_fwd_bwd(const float* log_probs, int B, int T, int U1, int V,
                                 const int32_t* labels, const int32_t* flens, const int32_t* ylens, int blank,
                                 float* nll, float* loss, float* grad,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* Training from logits: d loss / d logits of the same loss in ONE pass, after nsp_rnnt_loss_fwd_bwd(..., grad = NULL, ...) has
 * run on the SAME workspace (it holds the gathered emissions and the alpha / beta lattices) with log_probs = log_softmax(logits):
 *   dz[b,t,u,v] = g * ( gb [v == blank] + gl [v == y_{u+1}] - exp(log_probs[b,t,u,v]) (gb + gl) ),
 * gb / gl = the two non-zero entries of d loss / d log_probs of the cell; g = optional device scalar (upstream gradient).
 * dz fp32 (may alias log_probs: in place) or bf16 [B,T,U1,V].  Replaces torch autograd through
 * `torch.log_softmax` + `warp_rnnt.rnnt_loss` (rnn_transducer.py:242-252) without materialising d loss / d log_probs. */
nsp_status nsp_rnnt_grad_logits(const float* log_probs, int B, int T, int U1, int V, const int32_t* labels,
                                const int32_t* flens, const int32_t* ylens, int blank, const float* nll,
                                const void* workspace, size_t workspace_bytes, const float* gscale,
                                void* dz, int dz_bf16, void* stream);

/* 3x3 conv, C_in = C_out = 32, as an implicit GEMM on tcgen05 (bf16 in/out, fp32 accumulate), fused bias + ReLU
 * and optional fused 2x2 ceil-mode max-pool.  Replaces nn.Conv2d(32,32,3,pad 1)+ReLU(+MaxPool2d) of
 * Conv2dBlock.forward encoders/conv.py:362-394.  x bf16 [B,T,F,32]; w_taps bf16 [32, 288] with column
 * (ky*3+kx)*32 + ci; y bf16 [B,T,F,32] or [B,ceil(T/2),ceil(F/2),32] when pool2x2. */
nsp_status nsp_conv3x3_c32_tc_fwd(const void* x, const void* w_taps, const float* bias, void* y,
                                  int B, int T, int F, int relu, int pool2x2, void* stream);

/* Row-wise softmax (log_mode=0) or log-softmax (log_mode=1) of x / temperature, fp32 [rows, V]; y may alias x.
 * Replaces CTC.probs / CTC.scores decoders/ctc.py:197-217 and torch.log_softmax at decoders/rnn_transducer.py:242. */
nsp_status nsp_softmax_rows(const float* x, float* y, int64_t rows, int V, int log_mode, float temperature, void* stream);
/* Greedy CTC path: best[b,t] = argmax_v logits[b,t,v] (first max); hyp[b,:hyp_lens[b]] = collapsed non-blank labels of
 * frames t < elens[b]; trigger[b,n] = first frame of the n-th non-blank run.  All int32; hyp/trigger are [B,T].
 * Replaces CTC.greedy decoders/ctc.py:219-243 and CTC.trigger_points decoders/ctc.py:152-195. */
nsp_status nsp_ctc_greedy(const float* logits, int B, int T, int V, const int32_t* elens, int blank,
                          int32_t* best, int32_t* hyp, int32_t* hyp_lens, int32_t* trigger, void* stream);
/* RNN-T joint pre-activation out[b,t,u,:] = tanh(enc[b,t,:] + dec[b,u,:]) (decoders/rnn_transducer.py:272-275);
 * enc fp32 [B,T,J], dec fp32 [B,U1,J], out fp32 or bf16 [B,T,U1,J]; J % 4 == 0. */
nsp_status nsp_rnnt_joint_tanh(const float* enc, const float* dec, void* out, int out_bf16, int B, int T, int U1,
                               int J, void* stream);

/* Training of the joint network (the reference: torch autograd through rnn_transducer.py:242 and :273).
 * nsp_log_softmax_bwd: dz = g * (dlp - exp(lp) * rowsum(dlp)) IN PLACE on dlp; lp, dlp fp32 [rows, V]; g = optional device
 *   scalar (upstream gradient of the loss), NULL = 1.
 * nsp_rnnt_joint_tanh_bwd: p = dh * (1 - h^2); de[b,t,:] = sum_u p, dd[b,u,:] = sum_t p; h, dh fp32 or bf16 [B,T,U1,J],
 *   de fp32 [B,T,J], dd fp32 [B,U1,J]. */
nsp_status nsp_log_softmax_bwd(const float* lp, float* dlp, int64_t rows, int V, const float* gscale, void* stream);
nsp_status nsp_rnnt_joint_tanh_bwd(int is_bf16, const void* h, const void* dh, float* de, float* dd, int B, int T, int U1,
                                   int J, void* stream);

/* ------------------------------------------------------------------------------------------
 * (Bi)LSTM layer recurrence over length-masked sequences (persistent cooperative kernel, fp32 math).
 * Replaces Padding.forward encoders/rnn.py:534-546 (pack_padded_sequence -> nn.LSTM -> pad_packed_sequence) for one
 * layer; the input projection gates_x = x W_ih^T + b_ih + b_hh [B, T, ndir*4H] (PyTorch gate order i,f,g,o; direction-
 * major) is computed beforehand with nsp_linear_fwd.  w_hh fp32 [ndir, 4H, H]; lens int32 [B]; y fp32 [B, T, ndir*H]
 * (zeros for t >= lens[b], like pad_packed_sequence).  H % 8 == 0, B <= 128.
 * ------------------------------------------------------------------------------------------ */
size_t nsp_lstm_workspace_bytes(int B, int H, int ndir);
nsp_status nsp_lstm_seq_fwd(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                            int B, int T, int H, int ndir, void* workspace, size_t workspace_bytes, void* stream);

/* Streaming variant (state carry-over across chunks, encoders/rnn.py:343-346 `rnn(xs, hx=prev_state)` and the LC-BLSTM
 * chunk loop :454-498): h0, c0 fp32 [ndir, B, H] = nn.LSTM's hx (null: zeros); hN, cN fp32 [ndir, B, H] receive the
 * state after each utterance's last valid frame (both null: not wanted). */
nsp_status nsp_lstm_seq_fwd_state(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                  int B, int T, int H, int ndir, const float* h0, const float* c0,
                                  float* hN, float* cN, void* workspace, size_t workspace_bytes, void* stream);

/* Training variants of the LSTM layer recurrence (reference: autograd through nn.LSTM, encoders/rnn.py:534-546).
 * _fwd_save additionally stores, for every valid (b, t, dir): the gate activations i,f,g,o  acts [B,T,ndir,4H], the cell
 * state entering the step  cprev [B,T,ndir,H]  and the hidden state entering it  hprev [B,T,ndir,H]  (buffers must be
 * zero-initialised by the caller: padded frames are not written).
 * _bwd runs backpropagation through time: dy [B,T,ndir*H] -> dgates [B,T,ndir*4H] = d loss / d gate pre-activations
 * (layout of gates_x; zero on padded frames).  The weight / input gradients are GEMMs on dgates (nsp_linear_wgrad, nsp_linear_fwd).
 * workspace: >= 256 bytes. */
nsp_status nsp_lstm_seq_fwd_save(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                 int B, int T, int H, int ndir, float* acts, float* cprev, float* hprev,
                                 void* workspace, size_t workspace_bytes, void* stream);
nsp_status nsp_lstm_seq_bwd(const float* dy, const float* acts, const float* cprev, const float* w_hh,
                            const int32_t* lens, float* dgates, int B, int T, int H, int ndir,
                            void* workspace, size_t workspace_bytes, void* stream);
/* Training with a state carried across chunks (latency-controlled BLSTM, rnn.py:454-498: the forward LSTM of chunk i+1 starts
 * from the state chunk i ended in, and autograd back-propagates through it): _fwd_save_state = _fwd_save + the initial / final
 * state arguments of _fwd_state; _bwd_state additionally takes the gradient w.r.t. the final state (dhN, dcN; null = none)
 * and returns the gradient w.r.t. the initial state (dh0, dc0; null = not wanted), all fp32 [ndir, B, H]. */
nsp_status nsp_lstm_seq_fwd_save_state(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                       int B, int T, int H, int ndir, float* acts, float* cprev, float* hprev,
                                       const float* h0, const float* c0, float* hN, float* cN,
                                       void* workspace, size_t workspace_bytes, void* stream);
nsp_status nsp_lstm_seq_bwd_state(const float* dy, const float* acts, const float* cprev, const float* w_hh,
                                  const int32_t* lens, float* dgates, int B, int T, int H, int ndir,
                                  const float* dhN, const float* dcN, float* dh0, float* dc0,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* bf16-mode recurrence on the tensor cores (lstm_tc.cu): the same layer, same arguments and outputs, with the per-step
 * products h_{t-1} W_hh^T (forward) and dG_t W_hh (backward) issued as tcgen05.mma on bf16 operands with fp32 accumulation;
 * cell state, gate math, saved activations and every output stay fp32.  This is what the reference's AMP mode does through
 * cuDNN (encoders/rnn.py:534-546 under torch.cuda.amp.autocast, models/seq2seq/speech2text.py:207-214).
 * One entry point per direction of time covers all the variants above: acts / cprev / hprev (all or none), h0 / c0, hN / cN,
 * dhN / dcN, dh0 / dc0 are optional (null).  _supported: 1 when (B, H, ndir) is covered (B <= 128, H % 64 == 0, H / 8 CTAs
 * co-resident), else the caller uses the fp32 entry points.  Workspace: _tc_workspace_bytes(B, H, ndir, backward). */
int nsp_lstm_tc_supported(int B, int H, int ndir);
size_t nsp_lstm_tc_workspace_bytes(int B, int H, int ndir, int backward);
nsp_status nsp_lstm_seq_fwd_tc(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                               int B, int T, int H, int ndir, float* acts, float* cprev, float* hprev,
                               const float* h0, const float* c0, float* hN, float* cN,
                               void* workspace, size_t workspace_bytes, void* stream);
nsp_status nsp_lstm_seq_bwd_tc(const float* dy, const float* acts, const float* cprev, const float* w_hh,
                               const int32_t* lens, float* dgates, int B, int T, int H, int ndir,
                               const float* dhN, const float* dcN, float* dh0, float* dc0,
                               void* workspace, size_t workspace_bytes, void* stream);


/* ==========================================================================================
 * Backward pass of the encoder path (training).  The reference obtains all of these from torch autograd over
 * the modules cited at the forward entry points above; each function below is the hand-written gradient of
 * one forward entry point.  Gradients of parameters are ACCUMULATED into caller-provided fp32 buffers
 * (red.global.add / atomicAdd), matching autograd's .grad accumulation.
 * ========================================================================================== */

/* nsp_linear_fwd that additionally saves the pre-activation (acc + bias) for the backward pass.
 * pre: [M, N] in the operand dtype (bf16 for NSP_PREC_BF16, fp32 otherwise), pitch ldpre; for glu=1 columns
 * [0,N/2) hold the value half and [N/2,N) the gate half.  pre may be NULL (then identical to nsp_linear_fwd). */
nsp_status nsp_linear_fwd_save(int prec, const void* x, const void* x_lo, int64_t ldx,
                               const void* w, const void* w_lo, int64_t ldw,
                               int M, int N, int K, int glu, int act,
                               const float* bias, const float* residual, int64_t ldr, float alpha,
                               void* out, int64_t ldo, int out_bf16, void* out2, int64_t ldo2,
                               void* pre, int64_t ldpre, void* stream);

/* Process-wide tuning switch for nsp_linear_fwd / nsp_linear_fwd_save with NSP_PREC_BF16 (no reference counterpart; same
 * results in every mode):
 *   0 = every epilogue thread writes its output row with vector stores (default);
 *   1 = outputs (and the residual operand) move through swizzled shared-memory tiles and TMA bulk tensor copies
 *       (gemm_tma_epi.cu) for the shapes inside that kernel's envelope, mode 0 elsewhere;
 *   2 = mode 1, and problems with enough 256-row tiles run on CTA pairs (clusters of 2, tcgen05 cta_group::2: each CTA
 *       stages half of the W tile, halving the L2->shared-memory operand traffic per FLOP).
 * Not thread-safe against concurrent GEMM calls; set it once at start-up. */
nsp_status nsp_set_gemm_epilogue(int mode);
int nsp_get_gemm_epilogue(void);
/* number of GEMM calls of this process that ran the mode-1/2 kernel, and how many of those as CTA pairs (diagnostics / tests) */
long long nsp_gemm_tma_epilogue_launches(void);
long long nsp_gemm_cta_pair_launches(void);
/* modes >= 1 also switch nsp_linear_wgrad from per-thread red.global.add to staged cp.reduce.async.bulk.tensor adds */
long long nsp_wgrad_tma_epilogue_launches(void);

/* Weight gradient of out = x w^T (nn.Linear / 1x1 Conv1d):  dw[N,K] (+)= alpha * dy[M,N]^T x[M,K]  on tcgen05
 * with MN-major operands (no transposed copies) and split-K reduction by red.global.add.  NSP_PREC_BF16 only
 * (returns NSP_ERR_UNSUPPORTED otherwise: tf32 MN-major operands need another swizzle atom; parity-mode weight
 * gradients go through nsp_linear_fwd on transposed operands).  accumulate=0 zeroes dw first.
 * (The input gradient dx = dy w is nsp_linear_fwd with the transposed weight.) */
nsp_status nsp_linear_wgrad(int prec, const void* dy, const void* dy_lo, int64_t lddy,
                            const void* x, const void* x_lo, int64_t ldx, int M, int N, int K,
                            float alpha, float* dw, int64_t lddw, int accumulate, void* stream);

/* LayerNorm backward with the residual-branch gradient fused in:
 *   dx = dres + d/dx [ LN(x * in_scale) ] . dy ;  dgamma += sum_rows dy*xhat ;  dbeta += sum_rows dy.
 * dy, x fp32 [M,D]; dres fp32 [M,D] or NULL; dx fp32 and/or dx_bf16 outputs; dgamma/dbeta fp32 [D] or NULL.
 * dcol (optional) += dcol_alpha * column sums of dx: the bias gradient of the residual-branch GEMM that receives dx as dy. */
nsp_status nsp_layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* gamma,
                             float eps, float in_scale, const float* dres, int64_t lddr,
                             float* dx, int64_t lddx, void* dx_bf16, int64_t lddxb,
                             float* dgamma, float* dbeta, float* dcol, float dcol_alpha, int M, int D, void* stream);

/* dz = dh * act'(z) elementwise (act codes of nsp_linear_fwd); all bf16 (is_bf16=1) or all fp32. */
nsp_status nsp_act_bwd(int is_bf16, int act, const void* dh, const void* z, void* dz, int64_t n, void* stream);
/* nsp_act_bwd (mode 0, [M,N]) / nsp_glu_bwd (mode 1, dh = dg [M,N/2], z = pre, dz = dpre [M,N]) for bf16 tensors with the
 * bias gradient of that pre-activation fused: dbias fp32 [N] += column sums of dz. */
nsp_status nsp_act_bwd_bias(int mode, int act, const void* dh, const void* z, void* dz, float* dbias, int M, int N, void* stream);
/* GLU backward: pre = [a | b] ([M, 2d]), out = a*sigmoid(b); dpre = [dg*s | dg*a*s*(1-s)]. */
nsp_status nsp_glu_bwd(int is_bf16, const void* dg, const void* pre, void* dpre, int64_t M, int d, void* stream);
/* y[n] += alpha * sum_m x[m,n]  (bias gradients); x bf16 or fp32 [M,N] with pitch ldx. */
nsp_status nsp_colsum_acc(int is_bf16, const void* x, int64_t ldx, int M, int N, float alpha, float* y, void* stream);
/* MaxPoolSubsampler backward (encoders/subsampling.py:175-209): x, dx fp32 [B,T,D], dy fp32 [B,ceil(T/f),D]. */
nsp_status nsp_maxpool_time_bwd(const float* x, const float* dy, float* dx, int B, int T, int D, int factor, void* stream);
/* Backward of nsp_pool_time_fwd for mode 1 mean (MeanPoolSubsampler subsampling.py:212-246), 2 drop (DropSubsampler
 * :97-126), 3 add (AddSubsampler :129-172): dy fp32 [B,ceil(T/f),D] -> dx fp32 [B,T,D]. */
nsp_status nsp_pool_time_bwd(const float* dy, float* dx, int B, int T, int D, int factor, int mode, void* stream);
/* Dropout of the training path (nn.Dropout in the reference: conformer_block.py:133-180, transformer_block.py:128-139,
 * positionwise_feed_forward.py:83, positional_embedding.py:139, ctc.py:87).  Masks are never stored: element i of call site
 * `stream_id` is kept iff philox4x32_10(counter = (i/4, stream_id, offset), key = seed)[i%4] >= p * 2^32, with
 * rng_state = {seed, offset} (two uint64 in device memory), so the backward regenerates the forward's mask.
 *   nsp_dropout:     y = keep ? x * scale / (1-p) : 0            x, y fp32 or bf16; in place allowed for equal dtypes
 *   nsp_dropout_add: out = res + (keep ? t * alpha / (1-p) : 0)  t fp32 or bf16; res, out fp32 (out may alias res)
 *   nsp_rng_advance: offset += 1 (one call per training step; needed only so that CUDA-graph replays draw new masks) */
nsp_status nsp_dropout(int in_bf16, int out_bf16, const void* x, void* y, int64_t n, float p, float scale,
                       const uint64_t* rng_state, uint32_t stream_id, void* stream);
nsp_status nsp_dropout_add(int t_bf16, const void* t, const float* res, float* out, int64_t n, float p, float alpha,
                           const uint64_t* rng_state, uint32_t stream_id, void* stream);
nsp_status nsp_rng_advance(uint64_t* rng_state, void* stream);
/* ReLU backward through the saved post-activation: dz = a > 0 ? dx : 0. */
nsp_status nsp_relu_mask(int is_bf16, const void* dx, const void* a, void* dz, int64_t n, void* stream);
/* ReLU + MaxPool2d(ceil_mode) backward on channels-last [B,T,F,C] (encoders/conv.py:362-394): a = saved post-ReLU
 * activation, dy = gradient of the pooled tensor ([B,T',F',C], or [B,T',C*F'] when in_chmajor=1). */
nsp_status nsp_maxpool2d_relu_bwd(int is_bf16, int dy_bf16, const void* a, const void* dy, void* dz, int B, int T, int F,
                                  int C, int pool_t, int pool_f, int in_chmajor, void* stream);
/* 3x3 conv weight/bias gradient: dw[CO,CI,3,3] += sum dz[b,t,f,co] * a[b,t+ky-1,f+kx-1,ci]; dbias[CO] += sum dz. */
nsp_status nsp_conv3x3_wgrad(int a_bf16, int dz_bf16, const void* a, int in_chmajor, const void* dz, float* dw,
                             float* dbias, int B, int T, int F, int CI, int CO, void* stream);

/* Same weight gradient for the 32 -> 32 channel layers as an implicit GEMM on tcgen05 (bf16 a, dz [B,T,F,32];
 * dw fp32 [32,32,3,3] accumulated; the bias gradient is nsp_colsum_acc over dz viewed as [B*T*F, 32]). */
nsp_status nsp_conv3x3_c32_wgrad_tc(const void* a, const void* dz, float* dw, int B, int T, int F, void* stream);

/* nsp_relpos_attention_fwd that also returns the softmax statistics the tensor-core backward needs:
 * stats fp32 [B,H,Tq,2] = (row maximum in the log2 domain, 1 / row sum).  *stats_written_host (HOST int) is set to 1
 * iff the tcgen05 kernel ran and filled them (bf16, d_k = 64, clamped or absent relative term, no XL biases). */
nsp_status nsp_relpos_attention_fwd_stats(int is_bf16, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                          const void* v, int64_t ldv, const void* r, int64_t ldr, int rlen,
                                          const float* u_bias, const float* v_bias, const int32_t* klens,
                                          void* out, int64_t ldo, int B, int H, int Tq, int Tk, int dk,
                                          int clamp_len, int causal, int lookahead, int chunk_c, int chunk_l,
                                          float* stats, int* stats_written_host, void* stream);

/* Backward of nsp_relpos_attention_fwd (same argument meaning).  out is the forward result, dout its gradient;
 * stats (optional) = the statistics of nsp_relpos_attention_fwd_stats: when given and the shape is inside the
 * tensor-core envelope the five GEMMs of the backward run on tcgen05, otherwise exact fp32 CUDA-core kernels.
 * dq/dk/dv receive the gradients in the I/O dtype (may point into one fused [B*T, 3*H*dk] buffer);
 * dr fp32 [rlen, H*dk] (pitch lddr), du / dvb fp32 [H*dk] are ACCUMULATED (may be NULL). */
size_t nsp_relpos_attention_bwd_workspace_bytes(int B, int H, int Tq, int rlen, int clamp_len, int has_r);
nsp_status nsp_relpos_attention_bwd(int is_bf16, const void* q, int64_t ldq, const void* k, int64_t ldk,
                                    const void* v, int64_t ldv, const void* r, int64_t ldr, int rlen,
                                    const float* u_bias, const float* v_bias, const int32_t* klens,
                                    const void* out, int64_t ldo, const void* dout, int64_t lddo,
                                    void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv,
                                    float* dr, int64_t lddr, float* du, float* dvb, const float* stats,
                                    int B, int H, int Tq, int Tk, int dk_dim, int clamp_len, int causal, int lookahead,
                                    int chunk_c, int chunk_l, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of nsp_conformer_conv_fwd (LayerNorm variant only): dy = gradient of y; dz_ws = scratch [B*T, d] in the
 * I/O dtype; dx = gradient of x; dw [k,d], dbias [d], dnorm_w [d], dnorm_b [d] fp32 are ACCUMULATED.  workspace
 * (nsp_conformer_conv_bwd_workspace_bytes) holds the per-CTA partial sums of the four parameter gradients, which a last
 * kernel adds up: no atomics, deterministic.  Replaces autograd over modules/conformer_convolution.py:113-124. */
size_t nsp_conformer_conv_bwd_workspace_bytes(int B, int T, int d, int k);
nsp_status nsp_conformer_conv_bwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias,
                                  int norm_mode, const float* norm_w, const float* norm_b, float eps,
                                  const void* dy, int64_t lddy, void* dz_ws, int64_t lddz, void* dx, int64_t lddx,
                                  float* dw, float* dbias, float* dnorm_w, float* dnorm_b,
                                  int B, int T, int d, int k, int causal, void* workspace, size_t workspace_bytes,
                                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSP_B200_H_ */
