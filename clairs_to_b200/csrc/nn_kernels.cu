// fp32 CUDA-core kernels of the AFF / NEG forward passes (exact-parity path) and the fp64
// posterior combine.  The dense contractions that dominate (GRU projections, fc1, 1x1 convs)
// are additionally served by the tcgen05 kernels in gemm_tc.cu; this file holds everything
// that is not a plain GEMM plus the CUDA-core GEMM used for odd shapes.
//
// Reference: clairs/model.py (M:line), clairs/predict.py (P:line), clairs/call_variants.py (CV:line).
#include "nn_kernels.cuh"
#include <float.h>

namespace cto {

// ------------------------------------------------------------------------------------------
// activations
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {           // nn.GELU() default (M:83)
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float selu(float x) {               // nn.SELU (M:226)
    const float alpha = 1.6732632423543772848170429916717f;
    const float scale = 1.0507009873554804934193349852946f;
    return scale * (x > 0.0f ? x : alpha * expm1f(x));
}
__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_GELU) return gelu_erf(v);
    if (act == ACT_SELU) return selu(v);
    return v;
}

// ------------------------------------------------------------------------------------------
// int16 tensor -> fp32 network input with the depth rescale of P:179-197, 207
// ------------------------------------------------------------------------------------------
__global__ void rescale_kernel(const int16_t* __restrict__ x, const int32_t* __restrict__ depth, int64_t rows, int ld_out,
                               float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld_out) return;
    const int64_t row = i / ld_out;                    // (candidate, position)
    const int ch = (int)(i - row * ld_out);
    if (ch >= N_CH) { out[i] = 0.0f; return; }
    const int d = depth[row / N_POS];
    const double v = (double)x[row * N_CH + ch];
    // python: float(item) * (50.0 / depth) in double, then numpy float32 cast
    out[i] = d > MIN_RESCALE_COV ? __double2float_rn(__dmul_rn(v, __ddiv_rn(50.0, (double)d))) : (float)v;
}

// One thread per (candidate, position) row of 34 counts: the scale factor is computed once per row and the
// row is read as 17 aligned 32-bit words.  SPLIT = false: fp32 rows [n, 33, 34] (AFF input).  SPLIT = true: the
// NEG input as time-major bf16 hi / mid planes [33, bp, 40] (operand of the transposed input projection); threads
// are ordered position-major there so that a warp writes 32 consecutive output rows.
template <bool SPLIT>
__global__ void rescale_rows_kernel(const int16_t* __restrict__ x, const int32_t* __restrict__ depth, int64_t n, int64_t bp,
                                    float* __restrict__ out, uint16_t* __restrict__ hi, uint16_t* __restrict__ mid, int one_col) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * N_POS) return;
    const int64_t b = SPLIT ? i % n : i / N_POS;
    const int t = (int)(SPLIT ? i / n : i - b * N_POS);
    const int d = depth[b];
    const bool scale = d > MIN_RESCALE_COV;
    const double f = __ddiv_rn(50.0, (double)(scale ? d : 50));
    const uint32_t* src = reinterpret_cast<const uint32_t*>(x + (b * N_POS + t) * N_CH);      // 68-byte rows: 4-byte aligned
    float v[40];
    #pragma unroll
    for (int k = 0; k < N_CH / 2; ++k) {
        const uint32_t wd = __ldg(src + k);
        const int lo = (int)(int16_t)(wd & 0xFFFFu), hi16 = (int)(int16_t)(wd >> 16);
        // python: float(item) * (50.0 / depth) in double, then numpy float32 cast (P:179-197)
        v[2 * k] = scale ? __double2float_rn(__dmul_rn((double)lo, f)) : (float)lo;
        v[2 * k + 1] = scale ? __double2float_rn(__dmul_rn((double)hi16, f)) : (float)hi16;
    }
    if (SPLIT) {
        #pragma unroll
        for (int k = N_CH; k < 40; ++k) v[k] = k == one_col ? 1.0f : 0.0f;   // optional constant-one column (bias rides in W_ih)
        const int64_t o = ((int64_t)t * bp + b) * 40;
        #pragma unroll
        for (int q = 0; q < 5; ++q) {
            uint32_t h[4], m[4];
            #pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a0 = v[8 * q + 2 * e], a1 = v[8 * q + 2 * e + 1];
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[e]) : "f"(a1), "f"(a0));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m[e]) : "f"(a1 - __uint_as_float(h[e] & 0xFFFF0000u)), "f"(a0 - __uint_as_float(h[e] << 16)));
            }
            *reinterpret_cast<uint4*>(hi + o + 8 * q) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(mid + o + 8 * q) = make_uint4(m[0], m[1], m[2], m[3]);
        }
    } else {
        float2* dst = reinterpret_cast<float2*>(out + i * N_CH);                               // 136-byte rows: 8-byte aligned
        #pragma unroll
        for (int k = 0; k < N_CH / 2; ++k) dst[k] = make_float2(v[2 * k], v[2 * k + 1]);
    }
}

int launch_rescale(const int16_t* x, const int32_t* depth, int64_t n, float* out, int ld_out, cudaStream_t s) {
    if (n <= 0) return 0;
    CTO_REQUIRE(ld_out >= N_CH, "rescale: ld_out %d < %d", ld_out, N_CH);
    if (ld_out == N_CH && (reinterpret_cast<uintptr_t>(x) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
        rescale_rows_kernel<false><<<ceil_div(n * N_POS, 128), 128, 0, s>>>(x, depth, n, 0, out, nullptr, nullptr, -1);
    } else {
        const int64_t total = n * N_POS * ld_out;
        rescale_kernel<<<ceil_div(total, 256), 256, 0, s>>>(x, depth, n * N_POS, ld_out, out);
    }
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_rescale_split_time_major(const int16_t* x, const int32_t* depth, int64_t n, int64_t bp, uint16_t* hi, uint16_t* mid,
                                    cudaStream_t s, int one_col) {
    if (n <= 0) return 0;
    CTO_REQUIRE((reinterpret_cast<uintptr_t>(x) & 3) == 0 && bp >= n, "rescale_split: unaligned input or bp < n");
    static_assert(NEG_PLANE_LD == 40, "rescale_rows_kernel writes 40-element plane rows");
    CTO_REQUIRE(one_col < 0 || (one_col >= N_CH && one_col < NEG_PLANE_LD), "rescale_split: constant column %d", one_col);
    rescale_rows_kernel<true><<<ceil_div(n * N_POS, 128), 128, 0, s>>>(x, depth, n, bp, nullptr, hi, mid, one_col);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

__global__ void pad_rows_kernel(const float* __restrict__ x, int64_t rows, int cols, int ld_out, float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * ld_out) return;
    const int64_t row = i / ld_out;
    const int ch = (int)(i - row * ld_out);
    out[i] = ch < cols ? x[row * cols + ch] : 0.0f;
}

int launch_pad_rows(const float* x, int64_t rows, int cols, float* out, int ld_out, cudaStream_t s) {
    if (rows <= 0) return 0;
    pad_rows_kernel<<<ceil_div(rows * ld_out, 256), 256, 0, s>>>(x, rows, cols, ld_out, out);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// fp32 rows [n, t_len, ld] -> time-major bf16 hi / mid planes [t_len, bp, ld]; one thread per 8-element chunk
__global__ void split_time_major_kernel(const float* __restrict__ x, int64_t n, int t_len, int ld, int64_t bp,
                                        uint16_t* __restrict__ hi, uint16_t* __restrict__ mid, int one_col) {
    const int chunks = ld / 8;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * t_len * chunks) return;
    const int c = (int)(i % chunks);
    const int64_t row = i / chunks;                    // b * t_len + t
    const int64_t b = row / t_len;
    const int t = (int)(row - b * t_len);
    const float4 v0 = *reinterpret_cast<const float4*>(x + row * ld + c * 8);
    const float4 v1 = *reinterpret_cast<const float4*>(x + row * ld + c * 8 + 4);
    float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    if (one_col >= c * 8 && one_col < c * 8 + 8) v[one_col - c * 8] = 1.0f;      // optional constant-one column
    uint32_t h[4], m[4];
    #pragma unroll
    for (int e = 0; e < 4; ++e) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[e]) : "f"(v[2 * e + 1]), "f"(v[2 * e]));
        const float r0 = v[2 * e] - __uint_as_float(h[e] << 16), r1 = v[2 * e + 1] - __uint_as_float(h[e] & 0xFFFF0000u);
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m[e]) : "f"(r1), "f"(r0));
    }
    const int64_t o = ((int64_t)t * bp + b) * ld + c * 8;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(mid + o) = make_uint4(m[0], m[1], m[2], m[3]);
}

int launch_split_time_major(const float* x, int64_t n, int t_len, int ld_in, int64_t bp, uint16_t* hi, uint16_t* mid,
                            cudaStream_t s, int one_col) {
    if (n <= 0) return 0;
    CTO_REQUIRE(ld_in % 8 == 0 && bp >= n, "split_time_major: ld %d must be a multiple of 8 and bp >= n", ld_in);
    const int64_t total = n * t_len * (ld_in / 8);
    split_time_major_kernel<<<ceil_div(total, 256), 256, 0, s>>>(x, n, t_len, ld_in, bp, hi, mid, one_col);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// Rows of the 3-tap / stride-2 / pad-1 embed convolution (M:195 on H=1 maps) as an explicit matrix, written as bf16
// hi / mid planes [n * wout, 3 * cin] so that the convolution runs as a pre-split tensor-core GEMM.  One thread per
// four channels of one tap.
__global__ void im2col3_split_kernel(const float* __restrict__ x, int64_t n, int win, int wout, int cin,
                                     uint16_t* __restrict__ hi, uint16_t* __restrict__ mid) {
    const int q4 = cin / 4;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * wout * 3 * q4) return;
    const int c4 = (int)(i % q4);
    const int64_t r3 = i / q4;                          // (b * wout + w) * 3 + tap
    const int tap = (int)(r3 % 3);
    const int64_t row = r3 / 3;
    const int w = (int)(row % wout);
    const int64_t b = row / wout;
    const int pos = 2 * w - 1 + tap;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos >= 0 && pos < win) v = *reinterpret_cast<const float4*>(x + (b * win + pos) * cin + 4 * c4);
    uint32_t h0, h1, m0, m1;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(v.y), "f"(v.x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(v.w), "f"(v.z));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m0) : "f"(v.y - __uint_as_float(h0 & 0xFFFF0000u)), "f"(v.x - __uint_as_float(h0 << 16)));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m1) : "f"(v.w - __uint_as_float(h1 & 0xFFFF0000u)), "f"(v.z - __uint_as_float(h1 << 16)));
    const int64_t o = r3 * cin + 4 * c4;                // = row * 3 * cin + tap * cin + 4 * c4
    *reinterpret_cast<uint2*>(hi + o) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(mid + o) = make_uint2(m0, m1);
}

int launch_im2col3_split(const float* x, int64_t n, int win, int wout, int cin, uint16_t* hi, uint16_t* mid, cudaStream_t s) {
    if (n <= 0) return 0;
    CTO_REQUIRE(cin % 4 == 0 && wout == (win + 1) / 2, "im2col3: cin %d / win %d / wout %d", cin, win, wout);
    const int64_t total = n * wout * 3 * (cin / 4);
    im2col3_split_kernel<<<ceil_div(total, 256), 256, 0, s>>>(x, n, win, wout, cin, hi, mid);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// strand-count recovery (P:626-642): centre row, forward cols 0:4, reverse cols 9:13
__global__ void strand_counts_kernel(const int16_t* __restrict__ x, int64_t n, int32_t* __restrict__ fwd,
                                     int32_t* __restrict__ rev) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 2) return;
    const int64_t cand = i >> 1;
    const int which = (int)(i & 1);
    const int16_t* row = x + (cand * N_POS + CENTER) * N_CH + (which ? 9 : 0);
    int v[4], sum = 0;
    #pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = row[k]; sum += v[k]; }
    int32_t* dst = (which ? rev : fwd) + cand * 4;
    #pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = v[k] < 0 ? -sum : v[k];
}

int launch_strand_counts(const int16_t* x_aff, int64_t n, int32_t* fwd, int32_t* rev, cudaStream_t s) {
    if (n <= 0) return 0;
    strand_counts_kernel<<<ceil_div(n * 2, 128), 128, 0, s>>>(x_aff, n, fwd, rev);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// CUDA-core GEMM: 128x64 tile, 8x4 micro-tile, 256 threads
// ------------------------------------------------------------------------------------------
constexpr int G_BM = 128, G_BN = 64, G_BK = 16;

__device__ __forceinline__ void a_row_info(const AView& a, int64_t m, int k_total, const float*& base, int& k_lo,
                                           int& k_hi) {
    if (!a.conv) {
        base = a.ptr + m * a.lda;
        k_lo = 0;
        k_hi = k_total;
    } else {
        const int64_t b = m / a.wout;
        const int w = (int)(m - b * a.wout);
        const int first = 2 * w - 1;                       // stride 2, pad 1 (M:195)
        base = a.ptr + (b * a.win + first) * (int64_t)a.cin;
        k_lo = first < 0 ? a.cin : 0;
        const int avail = (a.win - first) * a.cin;
        k_hi = avail < k_total ? avail : k_total;
    }
}

__global__ void __launch_bounds__(256)
gemm_nt_kernel(AView a, const float* __restrict__ w, const float* __restrict__ bias,
               const float* residual, int64_t ldr, float* c, int64_t ldc, int64_t m_total,
               int n_total, int k_total, int act, int vec_ok) {
    __shared__ __align__(16) float As[G_BK][G_BM + 4];
    __shared__ __align__(16) float Ws[G_BK][G_BN + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * G_BM;
    const int n0 = blockIdx.y * G_BN;
    const int ty = tid >> 4, tx = tid & 15;

    // loader roles
    const int ar = tid >> 1, ak = (tid & 1) * 8;           // A: row ar, k offset ak..ak+7
    const int wr = tid >> 2, wk = (tid & 3) * 4;           // W: row wr, k offset wk..wk+3
    const float* a_base = nullptr;
    int a_lo = 0, a_hi = 0;
    const bool a_valid = (m0 + ar) < m_total;
    if (a_valid) a_row_info(a, m0 + ar, k_total, a_base, a_lo, a_hi);
    const bool w_valid = (n0 + wr) < n_total;
    const float* w_base = w + (int64_t)(n0 + wr) * k_total;

    float acc[8][4];
    #pragma unroll
    for (int i = 0; i < 8; ++i)
        #pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int k0 = 0; k0 < k_total; k0 += G_BK) {
        float av[8], wv[4];
        if (vec_ok && a_valid && k0 + ak + 8 <= k_total) {
            const float4 v0 = *reinterpret_cast<const float4*>(a_base + k0 + ak);
            const float4 v1 = *reinterpret_cast<const float4*>(a_base + k0 + ak + 4);
            av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
            av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
        } else {
            #pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = k0 + ak + i;
                av[i] = (a_valid && k >= a_lo && k < a_hi) ? a_base[k] : 0.0f;
            }
        }
        if (vec_ok && w_valid && k0 + wk + 4 <= k_total) {
            const float4 v = *reinterpret_cast<const float4*>(w_base + k0 + wk);
            wv[0] = v.x; wv[1] = v.y; wv[2] = v.z; wv[3] = v.w;
        } else {
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + wk + i;
                wv[i] = (w_valid && k < k_total) ? w_base[k] : 0.0f;
            }
        }
        __syncthreads();
        #pragma unroll
        for (int i = 0; i < 8; ++i) As[ak + i][ar] = av[i];
        #pragma unroll
        for (int i = 0; i < 4; ++i) Ws[wk + i][wr] = wv[i];
        __syncthreads();
        #pragma unroll
        for (int kk = 0; kk < G_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float ar8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float br4[4] = {b0.x, b0.y, b0.z, b0.w};
            #pragma unroll
            for (int i = 0; i < 8; ++i)
                #pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar8[i], br4[j], acc[i][j]);
        }
    }

    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t m = m0 + ty * 8 + i;
        if (m >= m_total) continue;
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= n_total) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.0f);
            v = apply_act(v, act);
            if (residual) v += residual[m * ldr + n];
            c[m * ldc + n] = v;
        }
    }
}

int launch_gemm_nt(const AView& a, const float* w, const float* bias, const float* residual, int64_t ldr, float* c,
                   int64_t ldc, int64_t m, int n, int k, int act, cudaStream_t s) {
    if (m <= 0 || n <= 0) return 0;
    const int vec_ok = (!a.conv && (a.lda % 4 == 0) && (k % 4 == 0) &&
                        ((reinterpret_cast<uintptr_t>(a.ptr) & 15) == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0))
                           ? 1 : 0;
    dim3 grid(ceil_div(m, G_BM), ceil_div(n, G_BN));
    gemm_nt_kernel<<<grid, 256, 0, s>>>(a, w, bias, residual, ldr, c, ldc, m, n, k, act, vec_ok);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// channel LayerNorm (M:57-67): population std, eps added to the std; one warp per (b, w) row
// ------------------------------------------------------------------------------------------
// Split-plane outputs: a tensor that is only read as the A operand of a bf16x3 GEMM is written as two bf16
// planes (hi = bf16(v), mid = bf16(v - hi)) instead of fp32 -- same bytes, and the GEMM needs no converter pass.

// Each lane owns four consecutive channels (one float4), C/4 lanes form a row group, 128/C rows per warp; the
// scalar version of this kernel was instruction-bound (ncu: 85 % issue utilisation at 20 % of the HBM rate).
template <bool SPLIT, int LPR>                          // LPR = lanes per row = C / 4
__global__ void channel_ln_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                  const float* __restrict__ b, float* __restrict__ y, uint16_t* __restrict__ y_hi,
                                  uint16_t* __restrict__ y_mid, int64_t rows) {
    constexpr int C = 4 * LPR, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    const int ch = (lane % LPR) * 4;
    const bool ok = row < rows;
    const float4 v = ok ? *reinterpret_cast<const float4*>(x + row * C + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
    float sum = (v.x + v.y) + (v.z + v.w);
    #pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)C;
    const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
    float sq = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    #pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float denom = sqrtf(sq / (float)C) + 1e-5f;
    if (!ok) return;
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g + ch)), bv = __ldg(reinterpret_cast<const float4*>(b + ch));
    const float o0 = d0 / denom * gv.x + bv.x, o1 = d1 / denom * gv.y + bv.y, o2 = d2 / denom * gv.z + bv.z,
                o3 = d3 / denom * gv.w + bv.w;
    if (SPLIT) {
        uint32_t h0, h1, m0, m1;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(o1), "f"(o0));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(o3), "f"(o2));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m0) : "f"(o1 - __uint_as_float(h0 & 0xFFFF0000u)), "f"(o0 - __uint_as_float(h0 << 16)));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m1) : "f"(o3 - __uint_as_float(h1 & 0xFFFF0000u)), "f"(o2 - __uint_as_float(h1 << 16)));
        *reinterpret_cast<uint2*>(y_hi + row * C + ch) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(y_mid + row * C + ch) = make_uint2(m0, m1);
    } else {
        *reinterpret_cast<float4*>(y + row * C + ch) = make_float4(o0, o1, o2, o3);
    }
}

template <bool SPLIT>
static int launch_channel_ln_t(const float* x, const float* g, const float* b, float* y, uint16_t* y_hi, uint16_t* y_mid,
                               int64_t rows, int c, cudaStream_t s) {
    const int rpw = 128 / c;                            // rows per warp
    const int grid = ceil_div(rows, 8 * rpw);
    switch (c) {
        case 16: channel_ln_kernel<SPLIT, 4><<<grid, 256, 0, s>>>(x, g, b, y, y_hi, y_mid, rows); break;
        case 32: channel_ln_kernel<SPLIT, 8><<<grid, 256, 0, s>>>(x, g, b, y, y_hi, y_mid, rows); break;
        case 64: channel_ln_kernel<SPLIT, 16><<<grid, 256, 0, s>>>(x, g, b, y, y_hi, y_mid, rows); break;
        case 128: channel_ln_kernel<SPLIT, 32><<<grid, 256, 0, s>>>(x, g, b, y, y_hi, y_mid, rows); break;
        default: CTO_REQUIRE(false, "channel_ln: C=%d not built (16, 32, 64, 128 are)", c);
    }
    return 0;
}

int launch_channel_ln(const float* x, const float* g, const float* b, float* y, int64_t rows, int c, cudaStream_t s,
                      uint16_t* y_hi, uint16_t* y_mid) {
    if (rows <= 0) return 0;
    CTO_REQUIRE(y_hi ? y_mid != nullptr : y != nullptr, "channel_ln: no output buffer");
    if (int rc = y_hi ? launch_channel_ln_t<true>(x, g, b, nullptr, y_hi, y_mid, rows, c, s)
                      : launch_channel_ln_t<false>(x, g, b, y, nullptr, nullptr, rows, c, s))
        return rc;
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// fused PreNorm + both depth-wise 3-tap convolutions (pad 1) of the attention block (M:57-67, 91-97, 112-113); the
// eval-mode BatchNorm scale is folded into the taps on the host, its shift into the following 1x1 conv's bias:
//   y = channel_LN(x);  dq = dw3_stride1(y) * BN-scale;  dkv = dw3_stride2(y) * BN-scale
// One warp per candidate; the candidate's <= 17 x 128 activations stay in shared memory, so x is read
// once and y never touches HBM.
// ------------------------------------------------------------------------------------------
constexpr int LND_WARPS = 4, LND_MAXW = 17;

__device__ __forceinline__ float4 fma4(const float4 a, const float4 b, const float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
template <bool SPLIT>
__device__ __forceinline__ void store4(float* __restrict__ o, uint16_t* __restrict__ hi, uint16_t* __restrict__ mid, int64_t idx,
                                       const float4 v) {
    if (SPLIT) {
        uint32_t h0, h1, m0, m1;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(v.y), "f"(v.x));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(v.w), "f"(v.z));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m0) : "f"(v.y - __uint_as_float(h0 & 0xFFFF0000u)), "f"(v.x - __uint_as_float(h0 << 16)));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(m1) : "f"(v.w - __uint_as_float(h1 & 0xFFFF0000u)), "f"(v.z - __uint_as_float(h1 << 16)));
        *reinterpret_cast<uint2*>(hi + idx) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(mid + idx) = make_uint2(m0, m1);
    } else {
        *reinterpret_cast<float4*>(o + idx) = v;
    }
}

// Each lane owns four consecutive channels; C/4 lanes form a row group and the 32/LPR groups of the warp take
// alternate rows (the scalar version was instruction-bound).
template <bool SPLIT, int LPR>
__global__ void __launch_bounds__(LND_WARPS * 32)
ln_dwconv_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                 const float* __restrict__ taps_q, const float* __restrict__ taps_kv, float* __restrict__ dq,
                 float* __restrict__ dkv, uint16_t* __restrict__ dq_hi, uint16_t* __restrict__ dq_mid,
                 uint16_t* __restrict__ dkv_hi, uint16_t* __restrict__ dkv_mid, int64_t batch, int w, int wkv) {
    constexpr int C = 4 * LPR, RPW = 32 / LPR;
    extern __shared__ __align__(16) float lnd_smem[];    // [LND_WARPS][w][C]: sized for the actual stage, not the maxima
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t cand = (int64_t)blockIdx.x * LND_WARPS + wib;
    if (cand >= batch) return;
    float* sy = lnd_smem + wib * w * C;
    const int grp = lane / LPR, ch = (lane % LPR) * 4;
    const float4 gv = __ldg(reinterpret_cast<const float4*>(g + ch)), bv = __ldg(reinterpret_cast<const float4*>(b + ch));
    for (int r0 = 0; r0 < w; r0 += RPW) {
        const int r = r0 + grp;
        const bool ok = r < w;
        const float4 v = ok ? *reinterpret_cast<const float4*>(x + (cand * w + r) * C + ch) : make_float4(0.f, 0.f, 0.f, 0.f);
        float sum = (v.x + v.y) + (v.z + v.w);
        #pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / (float)C;
        const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        float sq = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        #pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float denom = sqrtf(sq / (float)C) + 1e-5f;
        if (ok)
            *reinterpret_cast<float4*>(sy + r * C + ch) = make_float4(d0 / denom * gv.x + bv.x, d1 / denom * gv.y + bv.y,
                                                                      d2 / denom * gv.z + bv.z, d3 / denom * gv.w + bv.w);
    }
    __syncwarp();
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(taps_q + ch)), q1 = __ldg(reinterpret_cast<const float4*>(taps_q + C + ch)),
                 q2 = __ldg(reinterpret_cast<const float4*>(taps_q + 2 * C + ch));
    const float4 k0 = __ldg(reinterpret_cast<const float4*>(taps_kv + ch)), k1 = __ldg(reinterpret_cast<const float4*>(taps_kv + C + ch)),
                 k2 = __ldg(reinterpret_cast<const float4*>(taps_kv + 2 * C + ch));
    #define SY4(r) (*reinterpret_cast<const float4*>(sy + (r) * C + ch))
    for (int r = grp; r < w; r += RPW) {                    // stride 1, pad 1; taps in position order, BN scale folded on the host
        float4 acc = zero;
        if (r - 1 >= 0) acc = fma4(SY4(r - 1), q0, acc);
        acc = fma4(SY4(r), q1, acc);
        if (r + 1 < w) acc = fma4(SY4(r + 1), q2, acc);
        store4<SPLIT>(dq, dq_hi, dq_mid, (cand * w + r) * C + ch, acc);
    }
    for (int r = grp; r < wkv; r += RPW) {                  // stride 2, pad 1
        const int s0 = 2 * r - 1;
        float4 acc = zero;
        if (s0 >= 0) acc = fma4(SY4(s0), k0, acc);
        acc = fma4(SY4(s0 + 1), k1, acc);
        if (s0 + 2 < w) acc = fma4(SY4(s0 + 2), k2, acc);
        store4<SPLIT>(dkv, dkv_hi, dkv_mid, (cand * wkv + r) * C + ch, acc);
    }
    #undef SY4
}

template <bool SPLIT>
static int launch_ln_dwconv_t(const float* x, const float* g, const float* b, const float* taps_q, const float* taps_kv, float* dq,
                              float* dkv, uint16_t* dq_hi, uint16_t* dq_mid, uint16_t* dkv_hi, uint16_t* dkv_mid, int64_t batch,
                              int w, int wkv, int c, cudaStream_t s) {
    const int grid = ceil_div(batch, LND_WARPS);
    const size_t smem = sizeof(float) * LND_WARPS * w * c;
#define CTO_LND(LPR) ln_dwconv_kernel<SPLIT, LPR><<<grid, LND_WARPS * 32, smem, s>>>(x, g, b, taps_q, taps_kv, dq, dkv, dq_hi, dq_mid, \
                                                                                  dkv_hi, dkv_mid, batch, w, wkv)
    switch (c) {
        case 16: CTO_LND(4); break;
        case 32: CTO_LND(8); break;
        case 64: CTO_LND(16); break;
        case 128: CTO_LND(32); break;
        default: CTO_REQUIRE(false, "ln_dwconv: C=%d not built (16, 32, 64, 128 are)", c);
    }
#undef CTO_LND
    return 0;
}

int launch_ln_dwconv(const float* x, const float* g, const float* b, const float* taps_q, const float* taps_kv, float* dq,
                     float* dkv, int64_t batch, int w, int wkv, int c, cudaStream_t s, uint16_t* dq_hi, uint16_t* dq_mid,
                     uint16_t* dkv_hi, uint16_t* dkv_mid) {
    if (batch <= 0) return 0;
    CTO_REQUIRE(w <= LND_MAXW, "ln_dwconv: W=%d exceeds %d", w, LND_MAXW);
    CTO_REQUIRE(dq_hi ? (dq_mid && dkv_hi && dkv_mid) : (dq && dkv), "ln_dwconv: no output buffer");
    if (int rc = dq_hi ? launch_ln_dwconv_t<true>(x, g, b, taps_q, taps_kv, nullptr, nullptr, dq_hi, dq_mid, dkv_hi, dkv_mid, batch, w, wkv, c, s)
                       : launch_ln_dwconv_t<false>(x, g, b, taps_q, taps_kv, dq, dkv, nullptr, nullptr, nullptr, nullptr, batch, w, wkv, c, s))
        return rc;
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// attention over <= 17 query and <= 9 key positions, dim_head 64 (M:120-132); one warp per
// (candidate, head).  The 64^-0.5 scale is folded into the q projection on the host.
// ------------------------------------------------------------------------------------------
constexpr int ATT_MAXW = 17, ATT_MAXKV = 9, ATT_D = 64, ATT_WARPS = 4;

template <bool SPLIT>
__global__ void __launch_bounds__(ATT_WARPS * 32)
attention_kernel(const float* __restrict__ q, const float* __restrict__ kv, float* __restrict__ out, uint16_t* __restrict__ out_hi,
                 uint16_t* __restrict__ out_mid, int64_t n_bh, int w, int wkv, int heads) {
    // rows padded to 68 floats: 16-byte aligned for float4 reads, and 8 consecutive rows start 4 banks apart.
    // Dynamic shared memory sized for the actual (w, wkv) of the stage: twice the occupancy of arrays sized for 17 / 9.
    extern __shared__ __align__(16) float att_smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int LD = ATT_D + 4;
    const int per_warp = (w + wkv) * LD + ((w * (wkv + 1) + 3) & ~3);
    float* sq = att_smem + wib * per_warp;               // [w][68]
    float* sk = sq + w * LD;                             // [wkv][68]
    float* sp = sk + wkv * LD;                           // [w][wkv + 1]
    const int ldp = wkv + 1;
    const int64_t bh = (int64_t)blockIdx.x * ATT_WARPS + wib;
    if (bh >= n_bh) return;
    const int64_t b = bh / heads;
    const int h = (int)(bh - b * heads);
    const int inner = heads * ATT_D;
    // each lane owns head dimensions 2*lane, 2*lane+1: 8-byte global accesses, v stays in registers
    float2 vreg[ATT_MAXKV];
    for (int i = 0; i < w; ++i)
        *reinterpret_cast<float2*>(sq + i * LD + 2 * lane) = *reinterpret_cast<const float2*>(q + (b * w + i) * inner + h * ATT_D + 2 * lane);
    #pragma unroll
    for (int j = 0; j < ATT_MAXKV; ++j) {
        vreg[j] = make_float2(0.f, 0.f);
        if (j < wkv) {
            const float* src = kv + (b * wkv + j) * (2 * inner) + h * ATT_D + 2 * lane;
            *reinterpret_cast<float2*>(sk + j * LD + 2 * lane) = *reinterpret_cast<const float2*>(src);
            vreg[j] = *reinterpret_cast<const float2*>(src + inner);
        }
    }
    __syncwarp();
    for (int p = lane; p < w * wkv; p += 32) {
        const int i = p / wkv, j = p - i * wkv;
        const float4* qr = reinterpret_cast<const float4*>(sq + i * LD);
        const float4* kr = reinterpret_cast<const float4*>(sk + j * LD);
        float acc = 0.0f;
        #pragma unroll
        for (int d = 0; d < ATT_D / 4; ++d) {              // same summation order as the scalar loop
            const float4 a = qr[d], b = kr[d];
            acc = fmaf(a.x, b.x, acc);
            acc = fmaf(a.y, b.y, acc);
            acc = fmaf(a.z, b.z, acc);
            acc = fmaf(a.w, b.w, acc);
        }
        sp[i * ldp + j] = acc;
    }
    __syncwarp();
    if (lane < w) {
        float mx = -FLT_MAX;
        for (int j = 0; j < wkv; ++j) mx = fmaxf(mx, sp[lane * ldp + j]);
        float sum = 0.0f;
        for (int j = 0; j < wkv; ++j) {
            const float e = expf(sp[lane * ldp + j] - mx);
            sp[lane * ldp + j] = e;
            sum += e;
        }
        const float inv = 1.0f / sum;
        for (int j = 0; j < wkv; ++j) sp[lane * ldp + j] *= inv;
    }
    __syncwarp();
    for (int i = 0; i < w; ++i) {
        float o0 = 0.0f, o1 = 0.0f;
        #pragma unroll
        for (int j = 0; j < ATT_MAXKV; ++j) {
            if (j < wkv) {
                const float pij = sp[i * ldp + j];
                o0 = fmaf(pij, vreg[j].x, o0);
                o1 = fmaf(pij, vreg[j].y, o1);
            }
        }
        const int64_t o = (b * w + i) * inner + h * ATT_D + 2 * lane;
        if (SPLIT) {
            uint32_t hh, mm;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hh) : "f"(o1), "f"(o0));
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mm) : "f"(o1 - __uint_as_float(hh & 0xFFFF0000u)), "f"(o0 - __uint_as_float(hh << 16)));
            *reinterpret_cast<uint32_t*>(out_hi + o) = hh;
            *reinterpret_cast<uint32_t*>(out_mid + o) = mm;
        } else {
            *reinterpret_cast<float2*>(out + o) = make_float2(o0, o1);
        }
    }
}

int launch_attention(const float* q, const float* kv, float* out, int64_t batch, int w, int wkv, int heads,
                     cudaStream_t s, uint16_t* out_hi, uint16_t* out_mid) {
    const int64_t n_bh = batch * heads;
    if (n_bh <= 0) return 0;
    CTO_REQUIRE(w <= ATT_MAXW && wkv <= ATT_MAXKV, "attention: W=%d Wkv=%d exceed %d/%d", w, wkv, ATT_MAXW, ATT_MAXKV);
    const size_t smem = sizeof(float) * ATT_WARPS * ((w + wkv) * (ATT_D + 4) + ((w * (wkv + 1) + 3) & ~3));
    if (out_hi) attention_kernel<true><<<ceil_div(n_bh, ATT_WARPS), ATT_WARPS * 32, smem, s>>>(q, kv, nullptr, out_hi, out_mid, n_bh, w, wkv, heads);
    else attention_kernel<false><<<ceil_div(n_bh, ATT_WARPS), ATT_WARPS * 32, smem, s>>>(q, kv, out, nullptr, nullptr, n_bh, w, wkv, heads);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// GRU recurrence (torch.nn.GRU semantics, gate order r|z|n; M:412-417, 442-443).
//   xproj [B, T, 2*3H]  = W_ih x + b_ih (+ b_hh for the r,z gates), direction-major columns
//   whh_t [2][H][3H]    = W_hh transposed (k-major) per direction
//   bhn   [2][H]        = b_hn (stays inside r * (.))
//   out   [B, T, 2H]    = h_t, forward in columns 0..H-1, backward in H..2H-1
// One CTA = BM candidates x one direction for all 33 steps; h lives in shared memory (k-major),
// each thread owns one hidden unit for TM candidates and keeps its three gate accumulators in
// registers.  blockIdx.y = direction.
// ------------------------------------------------------------------------------------------
template <int H, int BM, int TM>
__global__ void __launch_bounds__(H * (BM / TM), 1)
gru_recurrent_kernel(const float* __restrict__ xproj, const float* __restrict__ whh_t,
                     const float* __restrict__ bhn, float* __restrict__ out, int64_t batch) {
    static_assert(TM == 8, "micro-tile is two float4");
    __shared__ __align__(16) float hT[H][BM + 4];      // +4: spreads the per-unit row stores over banks
    const int tid = threadIdx.x;
    const int j = tid % H;
    const int mg = tid / H;
    const int dir = blockIdx.y;
    const int64_t b0 = (int64_t)blockIdx.x * BM + mg * TM;
    const float* wt = whh_t + (int64_t)dir * H * 3 * H;
    const float b_hn = bhn[dir * H + j];
    const int64_t xstride = (int64_t)N_POS * 6 * H;        // per candidate
    const int64_t ostride = (int64_t)N_POS * 2 * H;

    float hreg[TM];                                    // this thread's h_{t-1}[m, j]
    #pragma unroll
    for (int i = 0; i < TM; ++i) { hreg[i] = 0.0f; hT[j][mg * TM + i] = 0.0f; }
    __syncthreads();

    for (int step = 0; step < N_POS; ++step) {
        const int t = dir ? (N_POS - 1 - step) : step;
        float xr[TM], xz[TM], xn[TM];
        #pragma unroll
        for (int i = 0; i < TM; ++i) {
            int64_t b = b0 + i;
            if (b >= batch) b = batch - 1;
            const float* xp = xproj + b * xstride + (int64_t)t * 6 * H + dir * 3 * H + j;
            xr[i] = xp[0];
            xz[i] = xp[H];
            xn[i] = xp[2 * H];
        }
        float ar[TM], az[TM], an[TM];
        #pragma unroll
        for (int i = 0; i < TM; ++i) { ar[i] = 0.0f; az[i] = 0.0f; an[i] = 0.0f; }
        #pragma unroll 4
        for (int k = 0; k < H; ++k) {
            const float wr = wt[k * 3 * H + j];
            const float wz = wt[k * 3 * H + H + j];
            const float wn = wt[k * 3 * H + 2 * H + j];
            const float4 h0 = *reinterpret_cast<const float4*>(&hT[k][mg * TM]);
            const float4 h1 = *reinterpret_cast<const float4*>(&hT[k][mg * TM + 4]);
            const float hv[TM] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            #pragma unroll
            for (int i = 0; i < TM; ++i) {
                ar[i] = fmaf(hv[i], wr, ar[i]);
                az[i] = fmaf(hv[i], wz, az[i]);
                an[i] = fmaf(hv[i], wn, an[i]);
            }
        }
        #pragma unroll
        for (int i = 0; i < TM; ++i) {
            const float r = 1.0f / (1.0f + expf(-(xr[i] + ar[i])));
            const float z = 1.0f / (1.0f + expf(-(xz[i] + az[i])));
            const float n = tanhf(xn[i] + r * (an[i] + b_hn));
            hreg[i] = (1.0f - z) * n + z * hreg[i];
        }
        __syncthreads();                               // every thread is done reading h_{t-1}
        *reinterpret_cast<float4*>(&hT[j][mg * TM]) = make_float4(hreg[0], hreg[1], hreg[2], hreg[3]);
        *reinterpret_cast<float4*>(&hT[j][mg * TM + 4]) = make_float4(hreg[4], hreg[5], hreg[6], hreg[7]);
        #pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int64_t b = b0 + i;
            if (b < batch) out[b * ostride + (int64_t)t * 2 * H + dir * H + j] = hreg[i];
        }
        __syncthreads();
    }
}

int launch_gru_recurrent(const float* xproj, const float* whh_t, const float* bhn, float* out, int64_t batch,
                         int hidden, cudaStream_t s) {
    if (batch <= 0) return 0;
    constexpr int BM = 32, TM = 8;
    dim3 grid(ceil_div(batch, BM), 2);
    if (hidden == 128) {
        gru_recurrent_kernel<128, BM, TM><<<grid, 128 * (BM / TM), 0, s>>>(xproj, whh_t, bhn, out, batch);
    } else if (hidden == 192) {
        gru_recurrent_kernel<192, BM, TM><<<grid, 192 * (BM / TM), 0, s>>>(xproj, whh_t, bhn, out, batch);
    } else {
        CTO_REQUIRE(false, "gru_recurrent: hidden size %d not built (128 and 192 are, M:403-404)", hidden);
    }
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// per-head fc3 (128 -> 2) + SELU (M:250-253): one warp per candidate
// ------------------------------------------------------------------------------------------
__global__ void head_fc3_kernel(const float* __restrict__ y, const float* __restrict__ w3, const float* __restrict__ b3,
                                float* __restrict__ logits, int64_t batch, int n_heads) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    for (int h = 0; h < n_heads; ++h) {
        const float4 yv = *reinterpret_cast<const float4*>(y + (b * n_heads + h) * 128 + lane * 4);
        #pragma unroll
        for (int o = 0; o < 2; ++o) {
            const float4 wv = *reinterpret_cast<const float4*>(w3 + (h * 2 + o) * 128 + lane * 4);
            float acc = yv.x * wv.x;
            acc = fmaf(yv.y, wv.y, acc);
            acc = fmaf(yv.z, wv.z, acc);
            acc = fmaf(yv.w, wv.w, acc);
            #pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, sft);
            if (lane == 0) logits[(b * n_heads + h) * 2 + o] = selu(acc + b3[h * 2 + o]);
        }
    }
}

int launch_head_fc3(const float* y, const float* w3, const float* b3, float* logits, int64_t batch, int n_heads,
                    cudaStream_t s) {
    if (batch <= 0) return 0;
    head_fc3_kernel<<<ceil_div(batch, 8), 256, 0, s>>>(y, w3, b3, logits, batch, n_heads);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// ------------------------------------------------------------------------------------------
// softmax (P:574, 659-684) + likelihood-matrix Bayes combine (CV:154-224 / 226-304) in fp64.
// tables: per head 100 matrix entries (row = AFF bin), 11 AFF edges, 11 NEG edges.
// The reference re-parses probabilities printed with 8 decimals (P:121-132, CV:803-829); the
// round trip is reproduced exactly: float32 -> double, x 1e8 (exact), rint, / 1e8.
// call[n]: bits 0-7 argmax head, bit 8 = some bin index was clamped (reference would raise,
// SURVEY.md 9.12).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int digitize11(double x, const double* edges) {   // np.digitize(x, edges) - 1
    int c = 0;
    #pragma unroll
    for (int i = 0; i < 11; ++i) c += (edges[i] <= x) ? 1 : 0;
    return c - 1;
}

// QUAL of clairs/call_variants.py:81-88 from the winning posterior, and the QUAL -> FILTER thresholds
// (call_variants.py:67-76 `--qual`, src/postprocess_vcf.py:61-82 phaseable / unphaseable cuts; shared/param.py:35-40):
// flt bit 0 = QUAL >= thr[0] (PASS of call_variants), bit 1 = QUAL >= thr[1] (phaseable), bit 2 = QUAL >= thr[2]
// (unphaseable).  The RefCall decision needs the alt_info strings and stays on the host.
__device__ __forceinline__ void qual_filter(double p, const double* thr, double* qual, int32_t* flt, int64_t i) {
    // Phred_Trans * log(((1 - p) + 1e-10) / (p + 1e-10)) + 2, floored at 0, rounded to 4 decimals
    const double phred_trans = -4.3429448190325175;                     // -10 * log(e, 10)
    const double ratio = __ddiv_rn(__dadd_rn(__dsub_rn(1.0, p), 1e-10), __dadd_rn(p, 1e-10));
    double t = __dadd_rn(__dmul_rn(phred_trans, log(ratio)), 2.0);
    t = t > 0.0 ? t : 0.0;
    const double q = __ddiv_rn(rint(__dmul_rn(t, 1e4)), 1e4);
    if (qual) qual[i] = q;
    if (flt) flt[i] = (q >= thr[0] ? 1 : 0) | (q >= thr[1] ? 2 : 0) | (q >= thr[2] ? 4 : 0);
}

__global__ void softmax_posterior_kernel(const float* __restrict__ la, const float* __restrict__ ln, int64_t n,
                                         int n_heads, const double* __restrict__ tables, float* __restrict__ probs,
                                         double* __restrict__ post, int32_t* __restrict__ call, double* __restrict__ qual,
                                         int32_t* __restrict__ flt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double best = -1.0;
    int best_h = 0, clamped = 0;
    for (int h = 0; h < n_heads; ++h) {
        float pa[2], pn[2];
        {
            const float x0 = la[(i * n_heads + h) * 2], x1 = la[(i * n_heads + h) * 2 + 1];
            const float mx = fmaxf(x0, x1);
            const float e0 = expf(x0 - mx), e1 = expf(x1 - mx);
            const float s = e0 + e1;
            pa[0] = e0 / s; pa[1] = e1 / s;
        }
        {
            const float x0 = ln[(i * n_heads + h) * 2], x1 = ln[(i * n_heads + h) * 2 + 1];
            const float mx = fmaxf(x0, x1);
            const float e0 = expf(x0 - mx), e1 = expf(x1 - mx);
            const float s = e0 + e1;
            pn[0] = e0 / s; pn[1] = e1 / s;
        }
        if (probs) {
            float* pr = probs + i * (4 * n_heads);
            pr[h * 2] = pa[0]; pr[h * 2 + 1] = pa[1];
            pr[(n_heads + h) * 2] = pn[0]; pr[(n_heads + h) * 2 + 1] = pn[1];
        }
        if (!tables) continue;
        const double p = __ddiv_rn(rint(__dmul_rn((double)pa[1], 1e8)), 1e8);
        const double q = __ddiv_rn(rint(__dmul_rn((double)pn[1], 1e8)), 1e8);
        const double* t = tables + h * 122;
        const double one_minus_q = __dsub_rn(1.0, q);
        int bi = digitize11(p, t + 100);
        int bj = digitize11(one_minus_q, t + 111);
        if (bi < 0 || bi > 9 || bj < 0 || bj > 9) clamped = 1;
        bi = min(max(bi, 0), 9);
        bj = min(max(bj, 0), 9);
        const double wgt = __dadd_rn(t[bi * 10 + bj], DBL_EPSILON);
        const double num = __dmul_rn(__dmul_rn(p, one_minus_q), wgt);
        const double alt = __dmul_rn(__dmul_rn(__dsub_rn(1.0, p), q), __dsub_rn(1.0, wgt));
        const double ps = __ddiv_rn(num, __dadd_rn(num, alt));
        if (post) post[i * n_heads + h] = ps;
        if (ps > best) { best = ps; best_h = h; }          // first maximum wins like np.argmax
    }
    if (call && tables) call[i] = best_h | (clamped << 8);
    if (tables && (qual || flt)) qual_filter(best, tables + 6 * 122, qual, flt, i);   // thresholds follow the tables
}

// Bayes combine on probabilities that were already parsed from a predict file (clairs/call_variants.py:798-829):
// pa / pn = P(positive class) per head as doubles, exactly the python floats the reference combines.
__global__ void posterior_from_probs_kernel(const double* __restrict__ pa, const double* __restrict__ pn, int64_t n,
                                            int n_heads, const double* __restrict__ tables, double* __restrict__ post,
                                            int32_t* __restrict__ call) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double best = -1.0;
    int best_h = 0, clamped = 0;
    for (int h = 0; h < n_heads; ++h) {
        const double p = pa[i * n_heads + h], q = pn[i * n_heads + h];
        const double* t = tables + h * 122;
        const double one_minus_q = __dsub_rn(1.0, q);
        int bi = digitize11(p, t + 100);
        int bj = digitize11(one_minus_q, t + 111);
        if (bi < 0 || bi > 9 || bj < 0 || bj > 9) clamped = 1;
        bi = min(max(bi, 0), 9);
        bj = min(max(bj, 0), 9);
        const double wgt = __dadd_rn(t[bi * 10 + bj], DBL_EPSILON);
        const double num = __dmul_rn(__dmul_rn(p, one_minus_q), wgt);
        const double alt = __dmul_rn(__dmul_rn(__dsub_rn(1.0, p), q), __dsub_rn(1.0, wgt));
        const double ps = __ddiv_rn(num, __dadd_rn(num, alt));
        post[i * n_heads + h] = ps;
        if (ps > best) { best = ps; best_h = h; }
    }
    call[i] = best_h | (clamped << 8);
}

int launch_posterior_from_probs(const double* pa, const double* pn, int64_t n, int n_heads, const double* tables,
                                double* post, int32_t* call, cudaStream_t s) {
    if (n <= 0) return 0;
    posterior_from_probs_kernel<<<ceil_div(n, 128), 128, 0, s>>>(pa, pn, n, n_heads, tables, post, call);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_softmax_posterior(const float* logits_aff, const float* logits_neg, int64_t n, int n_heads,
                             const double* tables, float* probs, double* post, int32_t* call, cudaStream_t s, double* qual,
                             int32_t* flt) {
    if (n <= 0) return 0;
    softmax_posterior_kernel<<<ceil_div(n, 128), 128, 0, s>>>(logits_aff, logits_neg, n, n_heads, tables, probs, post,
                                                              call, qual, flt);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
