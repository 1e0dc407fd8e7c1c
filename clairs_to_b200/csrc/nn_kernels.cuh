// Kernel launchers for the AFF (CvT) / NEG (BiGRU) forward passes and the posterior combine.
// All activations are channels-last fp32: [candidate, position, channel].
#pragma once
#include "common.cuh"

namespace cto {

enum Act { ACT_NONE = 0, ACT_GELU = 1, ACT_SELU = 2 };

// A-operand addressing for gemm_nt: plain row-major rows, or the rows of a 3-tap / stride-2 /
// pad-1 convolution over a channels-last [B, Win, Cin] tensor (clairs/model.py:195 on H=1 maps).
struct AView {
    const float* ptr;
    int64_t lda;      // plain: row stride
    int conv;         // 0 plain, 1 conv rows
    int win, wout, cin;
};

static inline AView plain_a(const float* p, int64_t lda) { return AView{p, lda, 0, 0, 0, 0}; }
static inline AView conv_a(const float* p, int win, int wout, int cin) { return AView{p, 0, 1, win, wout, cin}; }

// C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual)
int launch_gemm_nt(const AView& a, const float* w, const float* bias, const float* residual, int64_t ldr,
                   float* c, int64_t ldc, int64_t m, int n, int k, int act, cudaStream_t s);

// tcgen05 TF32 path (gemm_tc.cu); plain row-major A only
bool gemm_tc_supported(const float* a, int64_t lda, const float* w, int64_t m, int n, int k, const float* c, int64_t ldc,
                       const float* residual, int64_t ldr);
// w_hi / w_lo: the weight matrix split into TF32 hi + lo parts (launch_split_tf32)
int launch_gemm_tc(const float* a, int64_t lda, const float* w_hi, const float* w_lo, const float* bias,
                   const float* residual, int64_t ldr, float* c, int64_t ldc, int64_t m, int n, int k, int act,
                   cudaStream_t s);
int launch_split_tf32(const float* w, float* hi, float* lo, int64_t n, cudaStream_t s);

// ld_out >= 34: row stride of the fp32 output (extra columns are zero-filled)
int launch_rescale(const int16_t* x, const int32_t* depth, int64_t n, float* out, int ld_out, cudaStream_t s);
int launch_pad_rows(const float* x, int64_t rows, int cols, float* out, int ld_out, cudaStream_t s);
int launch_channel_ln(const float* x, const float* g, const float* b, float* y, int64_t rows, int c, cudaStream_t s);
int launch_dwconv3(const float* y, const float* taps, float* out, int64_t batch, int win, int wout, int stride,
                   int c, cudaStream_t s);
int launch_ln_dwconv(const float* x, const float* g, const float* b, const float* taps_q, const float* taps_kv, float* dq,
                     float* dkv, int64_t batch, int w, int wkv, int c, cudaStream_t s);
int launch_attention(const float* q, const float* kv, float* out, int64_t batch, int w, int wkv, int heads,
                     cudaStream_t s);
int launch_gru_recurrent(const float* xproj, const float* whh_t, const float* bhn, float* out, int64_t batch,
                         int hidden, cudaStream_t s);
// tensor-core recurrence (gru_tc.cu): w_hi / w_lo = W_hh regrouped per 32-unit block, TF32 hi / lo
int launch_gru_tc(const float* xproj, const float* w_hi, const float* w_lo, const float* bhn, float* out, int64_t batch,
                  int hidden, cudaStream_t s);
// CTA-pair recurrence (gru_tc2.cu, tcgen05 cta_group::2): W_hh regrouped per (32-unit block, 16-unit half, gate)
int launch_gru_pair(const float* xproj, const float* w_hi, const float* w_lo, const float* bhn, float* out, int64_t batch,
                    int hidden, cudaStream_t s);
int launch_head_fc3(const float* y, const float* w3, const float* b3, float* logits, int64_t batch, int n_heads,
                    cudaStream_t s);
int launch_softmax_posterior(const float* logits_aff, const float* logits_neg, int64_t n, int n_heads,
                             const double* tables, float* probs, double* post, int32_t* call, cudaStream_t s);
int launch_posterior_from_probs(const double* pa, const double* pn, int64_t n, int n_heads, const double* tables,
                                double* post, int32_t* call, cudaStream_t s);
int launch_strand_counts(const int16_t* x_aff, int64_t n, int32_t* fwd, int32_t* rev, cudaStream_t s);

}  // namespace cto
