// extern "C" wrappers that need no kernels of their own.
#include "common.cuh"

namespace nsp {
nsp_status gemm_dispatch(int precision, const void* a, const void* a_lo, int64_t lda, const void* w, const void* w_lo,
                         int64_t ldw, int M, int N, int K, int glu, int act, const float* bias,
                         const float* residual, int64_t ldr, float alpha, void* out, int64_t ldo, int out_bf16,
                         void* out2, int64_t ldo2, void* pre, int64_t ldpre, cudaStream_t st);
void set_gemm_epilogue_mode(int mode);
int gemm_epilogue_mode();
long long gemm_ts_launch_count();
long long gemm_pair_launch_count();
long long wgrad_tma_launch_count();
}

extern "C" nsp_status nsp_set_gemm_epilogue(int mode) {
    NSP_CHECK_ARG(mode >= 0 && mode <= 2, "nsp_set_gemm_epilogue: mode=%d (0 = direct stores, 1 = TMA-store epilogue, 2 = 1 + CTA pairs)", mode);
    nsp::set_gemm_epilogue_mode(mode);
    return NSP_OK;
}

extern "C" int nsp_get_gemm_epilogue(void) { return nsp::gemm_epilogue_mode(); }

extern "C" long long nsp_gemm_tma_epilogue_launches(void) { return nsp::gemm_ts_launch_count(); }
extern "C" long long nsp_gemm_cta_pair_launches(void) { return nsp::gemm_pair_launch_count(); }
extern "C" long long nsp_wgrad_tma_epilogue_launches(void) { return nsp::wgrad_tma_launch_count(); }

extern "C" nsp_status nsp_linear_fwd(int prec, const void* x, const void* x_lo, int64_t ldx,
                                     const void* w, const void* w_lo, int64_t ldw,
                                     int M, int N, int K, int glu, int act,
                                     const float* bias, const float* residual, int64_t ldr, float alpha,
                                     void* out, int64_t ldo, int out_bf16, void* out2, int64_t ldo2, void* stream) {
    return nsp::gemm_dispatch(prec, x, x_lo, ldx, w, w_lo, ldw, M, N, K, glu, act, bias, residual, ldr, alpha,
                              out, ldo, out_bf16, out2, ldo2, nullptr, 0, (cudaStream_t)stream);
}

extern "C" nsp_status nsp_linear_fwd_save(int prec, const void* x, const void* x_lo, int64_t ldx,
                                          const void* w, const void* w_lo, int64_t ldw,
                                          int M, int N, int K, int glu, int act,
                                          const float* bias, const float* residual, int64_t ldr, float alpha,
                                          void* out, int64_t ldo, int out_bf16, void* out2, int64_t ldo2,
                                          void* pre, int64_t ldpre, void* stream) {
    return nsp::gemm_dispatch(prec, x, x_lo, ldx, w, w_lo, ldw, M, N, K, glu, act, bias, residual, ldr, alpha,
                              out, ldo, out_bf16, out2, ldo2, pre, ldpre, (cudaStream_t)stream);
}
