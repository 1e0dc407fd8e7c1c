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

// tcgen05 bf16x3 path (gemm_tc.cu); plain row-major fp32 A only
bool gemm_tc_supported(const float* a, int64_t lda, const void* w, int64_t m, int n, int k, const float* c, int64_t ldc,
                       const float* residual, int64_t ldr);
// w_hi / w_mid: the weight matrix [n, k] split into bf16 hi + mid parts (launch_split_bf16)
int launch_gemm_tc(const float* a, int64_t lda, const uint16_t* w_hi, const uint16_t* w_mid, const float* bias,
                   const float* residual, int64_t ldr, float* c, int64_t ldc, int64_t m, int n, int k, int act,
                   cudaStream_t s);
int launch_split_bf16(const float* w, uint16_t* hi, uint16_t* mid, int64_t n, cudaStream_t s);
int launch_join_bf16(const uint16_t* hi, const uint16_t* mid, float* out, int64_t n, cudaStream_t s);
// the general form: A either fp32 (split by the kernel's converter warps) or already split into bf16 hi / mid
// planes; W always pre-split, row stride ldw
// GEMM_WIDE_N (with GEMM_A_PRESPLIT, many column tiles): 128 x 256 tiles in two 96 KB stages instead of 128 x 128 in three 64 KB
enum GemmFlags { GEMM_A_PRESPLIT = 1, GEMM_BIAS_PER_ROW = 2, GEMM_TILES_N_MAJOR = 4, GEMM_OUT_SPLIT = 8, GEMM_WIDE_N = 16 };
struct GemmTc {
    const float* a = nullptr;                          // fp32 A [m, k], row stride lda        (flags & A_PRESPLIT == 0)
    const uint16_t *a_hi = nullptr, *a_mid = nullptr;  // bf16 planes of A [m, k], row stride lda (flags & A_PRESPLIT)
    int64_t lda = 0;
    const uint16_t *w_hi = nullptr, *w_mid = nullptr;  // bf16 planes of W [n, k], row stride ldw
    int64_t ldw = 0;
    const float* bias = nullptr;                       // [n], or [m] with GEMM_BIAS_PER_ROW
    const float* residual = nullptr;
    int64_t ldr = 0;
    float* c = nullptr;                                // fp32 C [m, n], row stride ldc          (flags & OUT_SPLIT == 0)
    uint16_t *c_hi = nullptr, *c_mid = nullptr;        // bf16 planes of C, row stride ldc       (flags & OUT_SPLIT)
    int64_t ldc = 0, m = 0;
    int n = 0, k = 0, act = ACT_NONE, flags = 0;
};
int launch_gemm_tc_ex(const GemmTc& g, cudaStream_t s);
// the transposed GRU input projection on CTA pairs with A resident in shared memory (gemm_pair.cu)
bool gemm_pair_supported(const GemmTc& g);
int launch_gemm_pair(const GemmTc& g, cudaStream_t s);

// ld_out >= 34: row stride of the fp32 output (extra columns are zero-filled)
int launch_rescale(const int16_t* x, const int32_t* depth, int64_t n, float* out, int ld_out, cudaStream_t s);
// fp32 [n, win, cin] -> rows of the 3-tap / stride-2 / pad-1 convolution as bf16 hi / mid planes [n * wout, 3 * cin]
int launch_im2col3_split(const float* x, int64_t n, int win, int wout, int cin, uint16_t* hi, uint16_t* mid, cudaStream_t s);
constexpr int NEG_PLANE_LD = 40;   // row stride of the NEG input planes (34 channels + zeros; = engine.cuh NEG_IN_LD)
// int16 tensor -> rescaled NEG input written directly as time-major bf16 hi / mid planes [33, bp, 40]
int launch_rescale_split_time_major(const int16_t* x, const int32_t* depth, int64_t n, int64_t bp, uint16_t* hi, uint16_t* mid,
                                    cudaStream_t s, int one_col = -1);
int launch_pad_rows(const float* x, int64_t rows, int cols, float* out, int ld_out, cudaStream_t s);
// y_hi / y_mid, dq_hi .., out_hi ..: when given, the result is written as bf16 hi / mid planes (the pre-split A
// operand of the next GEMM) instead of fp32
int launch_channel_ln(const float* x, const float* g, const float* b, float* y, int64_t rows, int c, cudaStream_t s,
                      uint16_t* y_hi = nullptr, uint16_t* y_mid = nullptr);
int launch_ln_dwconv(const float* x, const float* g, const float* b, const float* taps_q, const float* taps_kv, float* dq,
                     float* dkv, int64_t batch, int w, int wkv, int c, cudaStream_t s, uint16_t* dq_hi = nullptr,
                     uint16_t* dq_mid = nullptr, uint16_t* dkv_hi = nullptr, uint16_t* dkv_mid = nullptr);
int launch_attention(const float* q, const float* kv, float* out, int64_t batch, int w, int wkv, int heads,
                     cudaStream_t s, uint16_t* out_hi = nullptr, uint16_t* out_mid = nullptr);
int launch_gru_recurrent(const float* xproj, const float* whh_t, const float* bhn, float* out, int64_t batch,
                         int hidden, cudaStream_t s);
// tensor-core recurrence (gru_tc3.cu): CTA pair (tcgen05 cta_group::2) + bf16x3, reads the TRANSPOSED projection xproj[6H][t * bp + b], writes
// h_t as bf16 hi / mid planes at row (b * osb + t * ost) of [.., 2H]
int launch_gru3(const float* xproj, int64_t ldx, int64_t bp, const uint16_t* w_hi, const uint16_t* w_mid, const float* bhn,
                uint16_t* out_hi, uint16_t* out_mid, int64_t osb, int64_t ost, int64_t batch, int hidden, cudaStream_t s);
// the same recurrence with TWO chains per CTA pair (gru_tc4.cu, H = 192): the gate math of one chain overlaps the MMAs of the
// other; dbg = device int[8] watchdog record
int launch_gru4(const float* xproj, int64_t ldx, int64_t bp, const uint16_t* w_hi, const uint16_t* w_mid, const float* bhn,
                uint16_t* out_hi, uint16_t* out_mid, int64_t osb, int64_t ost, int64_t batch, int hidden, int* dbg, cudaStream_t s);
// fp32 rows [n, t_len, ld_in] -> time-major bf16 hi / mid planes [t_len, bp, ld_in] (rows b >= n are left untouched)
// one_col >= 0: that (padding) column is set to 1.0 -- the bias column of the fused first GRU layer
int launch_split_time_major(const float* x, int64_t n, int t_len, int ld_in, int64_t bp, uint16_t* hi, uint16_t* mid,
                            cudaStream_t s, int one_col = -1);
// first GRU layer with the input projection fused in (gru_in_tc.cu)
int launch_gru1_fused(const uint16_t* x_hi, const uint16_t* x_mid, int ldx, int64_t bp, const uint16_t* win_hi,
                      const uint16_t* win_mid, const uint16_t* whh_hi, const uint16_t* whh_mid, const float* bhn,
                      uint16_t* out_hi, uint16_t* out_mid, int64_t osb, int64_t ost, int64_t batch, int hidden,
                      cudaStream_t s);
int launch_head_fc3(const float* y, const float* w3, const float* b3, float* logits, int64_t batch, int n_heads,
                    cudaStream_t s);
// tables: n_heads x 122 likelihood doubles at [0, 6 * 122), then the three QUAL thresholds at [6 * 122, 6 * 122 + 3)
int launch_softmax_posterior(const float* logits_aff, const float* logits_neg, int64_t n, int n_heads,
                             const double* tables, float* probs, double* post, int32_t* call, cudaStream_t s,
                             double* qual = nullptr, int32_t* flt = nullptr);
int launch_posterior_from_probs(const double* pa, const double* pn, int64_t n, int n_heads, const double* tables,
                                double* post, int32_t* call, cudaStream_t s);
int launch_strand_counts(const int16_t* x_aff, int64_t n, int32_t* fwd, int32_t* rev, cudaStream_t s);

}  // namespace cto
