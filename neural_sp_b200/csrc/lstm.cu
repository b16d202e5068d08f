// (Bi)LSTM layer recurrence over packed (length-masked) sequences, persistent cooperative kernel, fp32 math.
//
// Replaces  Padding.forward  encoders/rnn.py:534-546  (pack_padded_sequence -> nn.LSTM (cuDNN) -> pad_packed_sequence)
//           for one layer; the input projection x W_ih^T + b_ih + b_hh of all frames is a tcgen05 GEMM done before.
// PyTorch parameter layout: weight_hh_l0[_reverse] [4H, H], gate order i, f, g, o.
// Semantics kept: per-utterance lengths -- state freezes and outputs are zero for t >= len_b; the reverse direction
// of utterance b starts at its own last frame len_b - 1.
//
// Grid = n_dir x (H / 8) CTAs, all co-resident (cooperative launch).  A CTA owns 8 hidden units = 32 gate rows of W_hh,
// resident in shared memory for the whole sequence (fp32, padded rows: conflict-free).  Per time step every CTA reads
// h_{t-1} [B, H] from a double-buffered global array (L2, ld.global.cg), computes its 32 x B pre-activations with a
// 2-row x 4-batch register tile per thread, applies the cell update for its units (c and h live in registers) and
// publishes h_t; one device-scope barrier per direction per step orders the exchange.
#include <cooperative_groups.h>
#include "common.cuh"

namespace nsp {
namespace {

constexpr int UPC = 8;            // hidden units per CTA
constexpr int ROWS = 4 * UPC;     // gate rows per CTA
constexpr int BT = 32;            // batch tile
constexpr int KC = 128;           // k chunk staged per iteration

struct LstmParams {
    const float* gx;        // [B, T, ndir*4H]  x W_ih^T + b_ih + b_hh
    const float* whh;       // [ndir, 4H, H]
    const int32_t* lens;    // [B]
    float* y;               // [B, T, ndir*H]   (pre-zeroed)
    float* hbuf;            // [ndir, 2, B, H]  (pre-zeroed: h_0 = 0)
    unsigned int* bar;      // [ndir] (pre-zeroed)
    int B, T, H, ndir;
    int dir0;               // first direction handled by this launch (directions may be launched one at a time)
    // optional (training): what the backward recurrence needs, index ((b*T + t)*ndir + dir)*{4H|H} + {gate*H +} unit
    float* acts;            // [B, T, ndir, 4H]  gate activations i, f, g, o
    float* cprev;           // [B, T, ndir, H]   cell state entering the step
    float* hprev;           // [B, T, ndir, H]   hidden state entering the step (operand of the W_hh weight gradient)
    // optional (streaming, encoders/rnn.py:343-346): state carried across chunks, nn.LSTM's (h_n, c_n) layout [ndir, B, H]
    const float* h0;        // initial hidden state (null: zeros); the launcher also copies it into hbuf slot 0
    const float* c0;        // initial cell state (null: zeros)
    float* hN;              // final hidden state = state after each utterance's last valid frame (null: not wanted)
    float* cN;              // final cell state
};

__device__ __forceinline__ float sigmoid_exact(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256, 1) lstm_seq_kernel(LstmParams p) {
    pdl_entry();
    extern __shared__ float sm[];
    const int H = p.H;
    const int WP = H + 4;                              // padded row pitch (16-byte aligned, bank-staggered)
    float* Ws = sm;                                    // [ROWS][WP]   row r = gate * UPC + u
    float* hs = Ws + ROWS * WP;                        // [BT][KC + 4]
    float* pre = hs + BT * (KC + 4);                   // [2][ROWS][BT + 1] partial sums of the two k halves
    const int ctas_per_dir = H / UPC;
    const int dir = p.dir0 + blockIdx.x / ctas_per_dir;
    const int j0 = (blockIdx.x % ctas_per_dir) * UPC;
    const int tid = threadIdx.x;
    const int G = p.ndir * 4 * H;

    // resident W_hh slice
    const float* wg = p.whh + (size_t)dir * 4 * H * H;
    for (int e = tid; e < ROWS * H; e += 256) {
        const int r = e / H, k = e % H;
        const int gate = r / UPC, u = r % UPC;
        Ws[r * WP + k] = __ldg(wg + (size_t)(gate * H + j0 + u) * H + k);
    }
    __syncthreads();

    // GEMV-like tile: thread = (k half, row pair, batch quad)
    const int kh = tid >> 7, rp = (tid & 127) >> 3, bq = tid & 7;
    // cell-update mapping: thread = (unit, batch)
    const int cu = tid >> 5, cb = tid & 31;
    const int nbt = (p.B + BT - 1) / BT;
    float c_state[4], h_state[4];                      // up to 4 batch tiles (B <= 128) kept in registers
#pragma unroll
    for (int i = 0; i < 4; ++i) { c_state[i] = 0.f; h_state[i] = 0.f; }
    if (p.h0 || p.c0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = i * BT + cb;
            if (b < p.B) {
                const size_t si = ((size_t)dir * p.B + b) * H + j0 + cu;
                if (p.c0) c_state[i] = __ldg(p.c0 + si);
                if (p.h0) h_state[i] = __ldg(p.h0 + si);
            }
        }
    }

    for (int s = 0; s < p.T; ++s) {
        const float* hprev = p.hbuf + ((size_t)(dir * 2 + (s & 1)) * p.B) * H;
        float* hnext = p.hbuf + ((size_t)(dir * 2 + ((s + 1) & 1)) * p.B) * H;
#pragma unroll 1
        for (int bt = 0; bt < nbt; ++bt) {
            const int b0 = bt * BT;
            float acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            for (int k0 = 0; k0 < H; k0 += KC) {
                __syncthreads();
                for (int e = tid; e < BT * (KC / 4); e += 256) {       // stage h_{t-1}[b0:b0+32, k0:k0+KC]
                    const int b = e / (KC / 4), k4 = e % (KC / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b0 + b < p.B && k0 + k4 * 4 < H) v = __ldcg(reinterpret_cast<const float4*>(hprev + (size_t)(b0 + b) * H + k0) + k4);
                    *reinterpret_cast<float4*>(hs + b * (KC + 4) + k4 * 4) = v;
                }
                __syncthreads();
                const int kbeg = kh * (KC / 2), kend = kbeg + KC / 2;
                const float* w0 = Ws + (2 * rp) * WP + k0;
                const float* w1 = w0 + WP;
                const float* hb = hs + (bq * 4) * (KC + 4);
#pragma unroll 4
                for (int k = kbeg; k < kend; k += 4) {
                    if (k0 + k >= H) break;
                    const float4 a0 = *reinterpret_cast<const float4*>(w0 + k);
                    const float4 a1 = *reinterpret_cast<const float4*>(w1 + k);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 hv = *reinterpret_cast<const float4*>(hb + j * (KC + 4) + k);
                        acc[0][j] = fmaf(a0.x, hv.x, fmaf(a0.y, hv.y, fmaf(a0.z, hv.z, fmaf(a0.w, hv.w, acc[0][j]))));
                        acc[1][j] = fmaf(a1.x, hv.x, fmaf(a1.y, hv.y, fmaf(a1.z, hv.z, fmaf(a1.w, hv.w, acc[1][j]))));
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) pre[(kh * ROWS + 2 * rp + i) * (BT + 1) + bq * 4 + j] = acc[i][j];
            __syncthreads();
            // cell update for (unit cu, batch b0 + cb)
            const int b = b0 + cb;
            if (b < p.B) {
                const int len = min(max(p.lens[b], 0), p.T);
                if (s < len) {
                    const int t = dir == 0 ? s : (len - 1 - s);
                    const float* gxr = p.gx + ((size_t)b * p.T + t) * G + (size_t)dir * 4 * H + j0 + cu;
                    float g4[4];
#pragma unroll
                    for (int gate = 0; gate < 4; ++gate) {
                        const int r = gate * UPC + cu;
                        g4[gate] = pre[r * (BT + 1) + cb] + pre[(ROWS + r) * (BT + 1) + cb] + __ldg(gxr + (size_t)gate * H);
                    }
                    const float ig = sigmoid_exact(g4[0]), fg = sigmoid_exact(g4[1]), gg = tanhf(g4[2]), og = sigmoid_exact(g4[3]);
                    float c = c_state[bt], h;
                    if (p.acts) {
                        const size_t cell = ((size_t)b * p.T + t) * p.ndir + dir;
                        float* ar = p.acts + cell * 4 * H + j0 + cu;
                        ar[0] = ig; ar[H] = fg; ar[2 * (size_t)H] = gg; ar[3 * (size_t)H] = og;
                        p.cprev[cell * H + j0 + cu] = c;
                        p.hprev[cell * H + j0 + cu] = h_state[bt];
                    }
                    c = fg * c + ig * gg;
                    h = og * tanhf(c);
                    c_state[bt] = c; h_state[bt] = h;
                    p.y[((size_t)b * p.T + t) * (p.ndir * H) + (size_t)dir * H + j0 + cu] = h;
                }
                __stcg(hnext + (size_t)b * H + j0 + cu, h_state[bt]);   // frozen state keeps being republished
            }
        }
        // device-scope barrier among the CTAs of this direction
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicAdd(p.bar + dir, 1u);
            const unsigned target = (unsigned)(s + 1) * (unsigned)ctas_per_dir;
            while (atomicAdd(p.bar + dir, 0u) < target) { __nanosleep(20); }
        }
        __syncthreads();
    }
    if (p.hN) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = i * BT + cb;
            if (b < p.B) {
                const size_t si = ((size_t)dir * p.B + b) * H + j0 + cu;
                p.hN[si] = h_state[i];
                p.cN[si] = c_state[i];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Backward through time of the same recurrence (the reference: autograd through nn.LSTM).  Same decomposition: a CTA
// owns 8 hidden units; W_hh^T rows of those units ([8][4H]) stay in shared memory.  Per step (in reverse step order):
//   a) cell backward for the CTA's units from dy_t + dh_rec (registers) and the saved gate activations / c_{t-1}:
//        dG_t[b, 4 gates x own units]  ->  global (this IS the output: d loss / d gate pre-activations);
//   b) device-scope barrier of the direction;
//   c) dh_rec[b, own units] = sum_r dG_t[b, r] * W_hh[r, unit]  over all 4H gate rows (read back from global / L2).
// dG then feeds the tensor-core GEMMs for dW_ih, dx (same layout as gates_x) and dW_hh (against the saved h_{t-1}).
// ---------------------------------------------------------------------------------------------
struct LstmBwdParams {
    const float* dy;        // [B, T, ndir*H]
    const float* acts;      // [B, T, ndir, 4H]
    const float* cprev;     // [B, T, ndir, H]
    const float* whh;       // [ndir, 4H, H]
    const int32_t* lens;
    float* dg;              // [B, T, ndir*4H]  (pre-zeroed)
    unsigned int* bar;      // [ndir] (pre-zeroed)
    int B, T, H, ndir, dir0;
    // optional (state carried across chunks in training, rnn.py:470-478): gradient w.r.t. the FINAL state flowing in from the
    // next chunk, and the gradient w.r.t. the INITIAL state flowing out to the previous one; all [ndir, B, H]
    const float* dhN; const float* dcN;
    float* dh0; float* dc0;
};

constexpr int KS = 8;             // k splits of the 4H-long reduction in phase c

__global__ void __launch_bounds__(256, 1) lstm_seq_bwd_kernel(LstmBwdParams p) {
    pdl_entry();
    extern __shared__ float sm[];
    const int H = p.H, H4 = 4 * p.H;
    const int WP = H4 + 4;
    float* Wt = sm;                                    // [UPC][WP]   Wt[u][r] = W_hh[r][j0 + u]
    float* xs = Wt + UPC * WP;                         // [BT][KC + 4]
    float* pre = xs + BT * (KC + 4);                   // [KS][UPC][BT + 1]
    const int ctas_per_dir = H / UPC;
    const int dir = p.dir0 + blockIdx.x / ctas_per_dir;
    const int j0 = (blockIdx.x % ctas_per_dir) * UPC;
    const int tid = threadIdx.x;
    const int G = p.ndir * H4;

    const float* wg = p.whh + (size_t)dir * H4 * H;
    for (int e = tid; e < UPC * H4; e += 256) {
        const int u = e % UPC, r = e / UPC;            // consecutive threads: consecutive units of one row (32-byte segments)
        Wt[u * WP + r] = __ldg(wg + (size_t)r * H + j0 + u);
    }
    __syncthreads();

    const int ks = tid >> 5, rp = (tid & 31) >> 3, bq = tid & 7;     // phase c tile: (k split, unit pair, batch quad)
    const int cu = tid >> 5, cb = tid & 31;                          // cell mapping: (unit, batch)
    const int nbt = (p.B + BT - 1) / BT;
    float dc_state[4], dh_rec[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { dc_state[i] = 0.f; dh_rec[i] = 0.f; }

    for (int it = 0; it < p.T; ++it) {
        const int s = p.T - 1 - it;                    // forward step index being undone
        // ---- a) cell backward ----
#pragma unroll 1
        for (int bt = 0; bt < nbt; ++bt) {
            const int b = bt * BT + cb;
            if (b < p.B) {
                const int len = min(max(p.lens[b], 0), p.T);
                if (s < len) {
                    const int t = dir == 0 ? s : (len - 1 - s);
                    const size_t cell = ((size_t)b * p.T + t) * p.ndir + dir;
                    const float* ar = p.acts + cell * H4 + j0 + cu;
                    const float ig = __ldg(ar), fg = __ldg(ar + H), gg = __ldg(ar + 2 * (size_t)H), og = __ldg(ar + 3 * (size_t)H);
                    const float cp = __ldg(p.cprev + cell * H + j0 + cu);
                    const float c = fg * cp + ig * gg;
                    const float tc = tanhf(c);
                    float dh = __ldg(p.dy + ((size_t)b * p.T + t) * (p.ndir * H) + (size_t)dir * H + j0 + cu) + dh_rec[bt];
                    float dc_in = dc_state[bt];
                    if (p.dhN && s == len - 1) {               // the final state (h_N, c_N) is the state after the last valid step
                        const size_t si = ((size_t)dir * p.B + b) * H + j0 + cu;
                        dh += __ldg(p.dhN + si);
                        dc_in += __ldg(p.dcN + si);
                    }
                    const float dc = dc_in + dh * og * (1.f - tc * tc);
                    dc_state[bt] = dc * fg;
                    float* gr = p.dg + ((size_t)b * p.T + t) * G + (size_t)dir * H4 + j0 + cu;
                    __stcg(gr, dc * gg * ig * (1.f - ig));
                    __stcg(gr + H, dc * cp * fg * (1.f - fg));
                    __stcg(gr + 2 * (size_t)H, dc * ig * (1.f - gg * gg));
                    __stcg(gr + 3 * (size_t)H, dh * tc * og * (1.f - og));
                }
            }
        }
        if (s == 0 && !p.dh0) break;                   // nothing flows into a zero initial state
        // ---- b) device-scope barrier among the CTAs of this direction ----
        __threadfence();
        __syncthreads();
        if (tid == 0) {
            atomicAdd(p.bar + dir, 1u);
            const unsigned target = (unsigned)(it + 1) * (unsigned)ctas_per_dir;
            const long long deadline = clock64() + 4000000000LL;       // ~2 s: a lost CTA becomes an error, not a hang
            while (atomicAdd(p.bar + dir, 0u) < target) {
                __nanosleep(20);
                if (clock64() > deadline) { printf("lstm_seq_bwd: barrier timeout (block %d, step %d)\n", blockIdx.x, s); __trap(); }
            }
        }
        __syncthreads();
        // ---- c) dh_rec for the own units: [B, 4H] x [4H, 8] ----
#pragma unroll 1
        for (int bt = 0; bt < nbt; ++bt) {
            const int b0 = bt * BT;
            float acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
            for (int k0 = 0; k0 < H4; k0 += KC) {
                __syncthreads();
                for (int e = tid; e < BT * (KC / 4); e += 256) {       // stage dG_t[b0:b0+32, k0:k0+KC] (zero for finished rows)
                    const int bb = e / (KC / 4), k4 = e % (KC / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int b = b0 + bb;
                    if (b < p.B && k0 + k4 * 4 < H4) {
                        const int len = min(max(p.lens[b], 0), p.T);
                        if (s < len) {
                            const int t = dir == 0 ? s : (len - 1 - s);
                            v = __ldcg(reinterpret_cast<const float4*>(p.dg + ((size_t)b * p.T + t) * G + (size_t)dir * H4 + k0) + k4);
                        }
                    }
                    *reinterpret_cast<float4*>(xs + bb * (KC + 4) + k4 * 4) = v;
                }
                __syncthreads();
                const int kbeg = ks * (KC / KS), kend = kbeg + KC / KS;
                const float* w0 = Wt + (2 * rp) * WP + k0;
                const float* w1 = w0 + WP;
                const float* xb = xs + bq * (KC + 4);            // batch rows bq, bq + 8, bq + 16, bq + 24: conflict-free float4 reads
#pragma unroll
                for (int k = kbeg; k < kend; k += 4) {
                    if (k0 + k >= H4) break;
                    const float4 a0 = *reinterpret_cast<const float4*>(w0 + k);
                    const float4 a1 = *reinterpret_cast<const float4*>(w1 + k);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 xv = *reinterpret_cast<const float4*>(xb + j * 8 * (KC + 4) + k);
                        acc[0][j] = fmaf(a0.x, xv.x, fmaf(a0.y, xv.y, fmaf(a0.z, xv.z, fmaf(a0.w, xv.w, acc[0][j]))));
                        acc[1][j] = fmaf(a1.x, xv.x, fmaf(a1.y, xv.y, fmaf(a1.z, xv.z, fmaf(a1.w, xv.w, acc[1][j]))));
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) pre[(ks * UPC + 2 * rp + i) * (BT + 1) + bq + 8 * j] = acc[i][j];
            __syncthreads();
            float r = 0.f;
#pragma unroll
            for (int q = 0; q < KS; ++q) r += pre[(q * UPC + cu) * (BT + 1) + cb];
            dh_rec[bt] = r;
        }
    }
    if (p.dh0) {                                       // gradient w.r.t. (h_0, c_0): what the recurrence passed below step 0
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = i * BT + cb;
            if (b < p.B) {
                const size_t si = ((size_t)dir * p.B + b) * H + j0 + cu;
                p.dh0[si] = dh_rec[i];
                p.dc0[si] = dc_state[i];
            }
        }
    }
}

}  // namespace
}  // namespace nsp

using namespace nsp;

extern "C" size_t nsp_lstm_workspace_bytes(int B, int H, int ndir) {
    if (B <= 0 || H <= 0 || ndir <= 0) return 0;
    return align_up((size_t)ndir * 2 * B * H * sizeof(float), 256) + 256;
}

static nsp_status lstm_fwd_impl(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                int B, int T, int H, int ndir, float* acts, float* cprev, float* hprev,
                                void* workspace, size_t workspace_bytes, void* stream,
                                const float* h0 = nullptr, const float* c0 = nullptr, float* hN = nullptr, float* cN = nullptr) {
    NSP_CHECK_ARG(gates_x && w_hh && lens && y && workspace, "lstm_seq: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && H > 0 && (ndir == 1 || ndir == 2), "lstm_seq: bad shape");
    if (H % UPC != 0 || H % 4 != 0) { set_error("lstm_seq: H=%d must be a multiple of 8", H); return NSP_ERR_UNSUPPORTED; }
    if (B > 4 * BT) { set_error("lstm_seq: B=%d unsupported (max 128)", B); return NSP_ERR_UNSUPPORTED; }
    NSP_CHECK_ARG(workspace_bytes >= nsp_lstm_workspace_bytes(B, H, ndir), "lstm_seq: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * ((size_t)ROWS * (H + 4) + (size_t)BT * (KC + 4) + (size_t)2 * ROWS * (BT + 1));
    if (smem > 226 * 1024) { set_error("lstm_seq: H=%d needs %zu B of shared memory", H, smem); return NSP_ERR_UNSUPPORTED; }
    NSP_CUDA_OK(cudaFuncSetAttribute(lstm_seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NSP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_seq_kernel, 256, smem));
    const int capacity = per_sm * num_sms();
    const int per_dir = H / UPC;
    if (capacity < per_dir) {
        set_error("lstm_seq: %d CTAs per direction cannot be co-resident (%d per SM x %d SMs)", per_dir, per_sm, num_sms());
        return NSP_ERR_UNSUPPORTED;
    }
    LstmParams p;
    p.gx = gates_x; p.whh = w_hh; p.lens = lens; p.y = y; p.B = B; p.T = T; p.H = H; p.ndir = ndir;
    p.acts = acts; p.cprev = cprev; p.hprev = hprev;
    p.h0 = h0; p.c0 = c0; p.hN = hN; p.cN = cN;
    const size_t hbytes = align_up((size_t)ndir * 2 * B * H * sizeof(float), 256);
    p.hbuf = (float*)workspace;
    p.bar = (unsigned int*)((char*)workspace + hbytes);
    NSP_CUDA_OK(cudaMemsetAsync(workspace, 0, hbytes + 256, st));
    NSP_CUDA_OK(cudaMemsetAsync(y, 0, (size_t)B * T * ndir * H * sizeof(float), st));
    if (h0)     // h_{-1} as seen by the other CTAs at step 0: slot 0 of each direction's double buffer
        for (int d = 0; d < ndir; ++d)
            NSP_CUDA_OK(cudaMemcpyAsync(p.hbuf + (size_t)d * 2 * B * H, h0 + (size_t)d * B * H, (size_t)B * H * sizeof(float),
                                        cudaMemcpyDeviceToDevice, st));
    // both directions in one cooperative launch when they fit together, else one launch per direction
    const int dirs_per_launch = (ndir * per_dir <= capacity) ? ndir : 1;
    for (int d0 = 0; d0 < ndir; d0 += dirs_per_launch) {
        p.dir0 = d0;
        void* args[] = {&p};
        NSP_CUDA_OK(cudaLaunchCooperativeKernel((void*)lstm_seq_kernel, dim3(dirs_per_launch * per_dir), dim3(256), args, smem, st));
    }
    return NSP_OK;
}

extern "C" nsp_status nsp_lstm_seq_fwd(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                       int B, int T, int H, int ndir, void* workspace, size_t workspace_bytes, void* stream) {
    return lstm_fwd_impl(gates_x, w_hh, lens, y, B, T, H, ndir, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" nsp_status nsp_lstm_seq_fwd_state(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                             int B, int T, int H, int ndir, const float* h0, const float* c0,
                                             float* hN, float* cN, void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG((hN == nullptr) == (cN == nullptr), "lstm_seq_fwd_state: hN and cN go together");
    return lstm_fwd_impl(gates_x, w_hh, lens, y, B, T, H, ndir, nullptr, nullptr, nullptr, workspace, workspace_bytes, stream,
                         h0, c0, hN, cN);
}

extern "C" nsp_status nsp_lstm_seq_fwd_save(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                            int B, int T, int H, int ndir, float* acts, float* cprev, float* hprev,
                                            void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(acts && cprev && hprev, "lstm_seq_fwd_save: null save buffer");
    return lstm_fwd_impl(gates_x, w_hh, lens, y, B, T, H, ndir, acts, cprev, hprev, workspace, workspace_bytes, stream);
}

extern "C" nsp_status nsp_lstm_seq_fwd_save_state(const float* gates_x, const float* w_hh, const int32_t* lens, float* y,
                                                  int B, int T, int H, int ndir, float* acts, float* cprev, float* hprev,
                                                  const float* h0, const float* c0, float* hN, float* cN,
                                                  void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(acts && cprev && hprev, "lstm_seq_fwd_save_state: null save buffer");
    NSP_CHECK_ARG((hN == nullptr) == (cN == nullptr), "lstm_seq_fwd_save_state: hN and cN go together");
    return lstm_fwd_impl(gates_x, w_hh, lens, y, B, T, H, ndir, acts, cprev, hprev, workspace, workspace_bytes, stream, h0, c0, hN, cN);
}

static nsp_status lstm_bwd_impl(const float* dy, const float* acts, const float* cprev, const float* w_hh,
                                const int32_t* lens, float* dgates, int B, int T, int H, int ndir,
                                const float* dhN, const float* dcN, float* dh0, float* dc0,
                                void* workspace, size_t workspace_bytes, void* stream);

extern "C" nsp_status nsp_lstm_seq_bwd(const float* dy, const float* acts, const float* cprev, const float* w_hh,
                                       const int32_t* lens, float* dgates, int B, int T, int H, int ndir,
                                       void* workspace, size_t workspace_bytes, void* stream) {
    return lstm_bwd_impl(dy, acts, cprev, w_hh, lens, dgates, B, T, H, ndir, nullptr, nullptr, nullptr, nullptr, workspace,
                         workspace_bytes, stream);
}

extern "C" nsp_status nsp_lstm_seq_bwd_state(const float* dy, const float* acts, const float* cprev, const float* w_hh,
                                             const int32_t* lens, float* dgates, int B, int T, int H, int ndir,
                                             const float* dhN, const float* dcN, float* dh0, float* dc0,
                                             void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG((dhN == nullptr) == (dcN == nullptr) && (dh0 == nullptr) == (dc0 == nullptr), "lstm_seq_bwd_state: state gradients come in pairs");
    return lstm_bwd_impl(dy, acts, cprev, w_hh, lens, dgates, B, T, H, ndir, dhN, dcN, dh0, dc0, workspace, workspace_bytes, stream);
}

static nsp_status lstm_bwd_impl(const float* dy, const float* acts, const float* cprev, const float* w_hh,
                                const int32_t* lens, float* dgates, int B, int T, int H, int ndir,
                                const float* dhN, const float* dcN, float* dh0, float* dc0,
                                void* workspace, size_t workspace_bytes, void* stream) {
    NSP_CHECK_ARG(dy && acts && cprev && w_hh && lens && dgates && workspace, "lstm_seq_bwd: null pointer");
    NSP_CHECK_ARG(B > 0 && T > 0 && H > 0 && (ndir == 1 || ndir == 2), "lstm_seq_bwd: bad shape");
    if (H % UPC != 0 || H % 4 != 0) { set_error("lstm_seq_bwd: H=%d must be a multiple of 8", H); return NSP_ERR_UNSUPPORTED; }
    if (B > 4 * BT) { set_error("lstm_seq_bwd: B=%d unsupported (max 128)", B); return NSP_ERR_UNSUPPORTED; }
    NSP_CHECK_ARG(workspace_bytes >= 256, "lstm_seq_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(float) * ((size_t)UPC * (4 * H + 4) + (size_t)BT * (KC + 4) + (size_t)KS * UPC * (BT + 1));
    if (smem > 226 * 1024) { set_error("lstm_seq_bwd: H=%d needs %zu B of shared memory", H, smem); return NSP_ERR_UNSUPPORTED; }
    NSP_CUDA_OK(cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    NSP_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lstm_seq_bwd_kernel, 256, smem));
    const int capacity = per_sm * num_sms();
    const int per_dir = H / UPC;
    if (capacity < per_dir) {
        set_error("lstm_seq_bwd: %d CTAs per direction cannot be co-resident (%d per SM x %d SMs)", per_dir, per_sm, num_sms());
        return NSP_ERR_UNSUPPORTED;
    }
    LstmBwdParams p;
    p.dy = dy; p.acts = acts; p.cprev = cprev; p.whh = w_hh; p.lens = lens; p.dg = dgates;
    p.B = B; p.T = T; p.H = H; p.ndir = ndir;
    p.dhN = dhN; p.dcN = dcN; p.dh0 = dh0; p.dc0 = dc0;
    p.bar = (unsigned int*)workspace;
    NSP_CUDA_OK(cudaMemsetAsync(workspace, 0, 256, st));
    NSP_CUDA_OK(cudaMemsetAsync(dgates, 0, (size_t)B * T * ndir * 4 * H * sizeof(float), st));
    const int dirs_per_launch = (ndir * per_dir <= capacity) ? ndir : 1;
    for (int d0 = 0; d0 < ndir; d0 += dirs_per_launch) {
        p.dir0 = d0;
        void* args[] = {&p};
        NSP_CUDA_OK(cudaLaunchCooperativeKernel((void*)lstm_seq_bwd_kernel, dim3(dirs_per_launch * per_dir), dim3(256), args, smem, st));
    }
    return NSP_OK;
}
