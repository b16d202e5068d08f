// Streaming kernels of the Conformer convolution core (conformer_conv_stream.cu), called first by the C-ABI entry points in
// conformer_conv.cu / conformer_conv_bwd.cu; NSP_ERR_UNSUPPORTED (without an error message) = shape not covered, use the
// general kernels.
#pragma once
#include "common.cuh"

namespace nsp {
size_t conv_stream_bwd_workspace_bytes(int B, int T, int d, int k);
nsp_status conv_stream_fwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias, const float* g,
                           const float* bta, float eps, void* y, int64_t ldy, int B, int T, int d, int k, int causal,
                           cudaStream_t st);
nsp_status conv_stream_bwd(int is_bf16, const void* x, int64_t ldx, const float* w, const float* bias, const float* g,
                           const float* bta, float eps, const void* dy, int64_t lddy, void* dz, int64_t lddz, void* dx,
                           int64_t lddx, float* dw, float* dbias, float* dg, float* db, int B, int T, int d, int k, int causal,
                           void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace nsp
