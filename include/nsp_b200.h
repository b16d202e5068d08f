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

#ifdef __cplusplus
}
#endif
#endif /* NSP_B200_H_ */
