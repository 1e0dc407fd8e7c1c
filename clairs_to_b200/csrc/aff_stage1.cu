// Whole first CvT stage of the AFF network in ONE kernel (clairs/model.py:194-198 with the predict.py:520-553
// hyper-parameters: C=16, one head, depth 1):
//
//   embed conv (3-tap, stride 2, pad 1, 34 -> 16 channels) -> channel LN
//   -> x += Attention(LN(x))   (depth-wise 3-tap q / kv projections, 17 queries x 9 keys, out-projection)
//   -> x += FeedForward(LN(x)) (16 -> 64, exact-erf GELU, 64 -> 16)
//
// The stage is only 0.27 MFLOP per candidate but was ten launches whose activations ([.,17,16] .. [.,17,64]) went
// through L2 / HBM nine times; per candidate everything fits in 15 KB of shared memory and the weights in 38 KB.
// One warp owns one candidate from the rescaled input to the stage output; the small dense products run on the
// CUDA cores as warp-level GEMMs (lanes over output columns, rows in registers, float4 shared-memory operands),
// in fp32 with fmaf -- the same arithmetic as the per-op kernels, so parity with the oracle is unchanged.
#include "engine.cuh"

namespace cto {

namespace s1 {

constexpr int C = 16, CIN = 34, WIN = 33, W = 17, WKV = 9, INNER = 64, FF = 64;
constexpr int WARPS = 12;
constexpr int LDXP = 36;                 // input row stride (34 channels + 2 zeros): 3 rows = one 108-float conv window
constexpr int KE = 3 * LDXP;             // embed conv K (with the zero columns)
// weight row strides in shared memory: K rounded up to 4*odd floats so that float4 reads of 8 consecutive rows
// land in 8 different 16-byte bank groups
constexpr int LDW_E = 108, LDW_16 = 20, LDW_64 = 68;
constexpr int LDA_16 = 20, LDA_64 = 68, LDA_KV = 132;

// shared-memory weight block (floats)
constexpr int O_EMB = 0;                             // [16][108]
constexpr int O_QPW = O_EMB + C * LDW_E;             // [64][20]
constexpr int O_KVPW = O_QPW + INNER * LDW_16;       // [128][20]
constexpr int O_OUT = O_KVPW + 2 * INNER * LDW_16;   // [16][68]
constexpr int O_FF1 = O_OUT + C * LDW_64;            // [64][20]
constexpr int O_FF2 = O_FF1 + FF * LDW_16;           // [16][68]
constexpr int O_VEC = O_FF2 + C * LDW_64;            // small vectors, see V_*
constexpr int V_EMB_B = 0, V_LN_G = 16, V_LN_B = 32, V_LN1_G = 48, V_LN1_B = 64, V_QDW = 80, V_KVDW = 128, V_QB = 176,
              V_KVB = 240, V_OUT_B = 368, V_LN2_G = 384, V_LN2_B = 400, V_FF1_B = 416, V_FF2_B = 480, V_END = 496;
constexpr int W_FLOATS = O_VEC + V_END;

// per-warp activation block (floats)
constexpr int A_X = 0;                               // [35][36] zero-framed input; later reused: q / o / ff [17][68]
constexpr int A_KV = A_X + (WIN + 2) * LDXP;         // [9][132]
constexpr int A_XS = A_KV + WKV * LDA_KV;            // [17][20] residual stream
constexpr int A_Y = A_XS + W * LDA_16;               // [19][20] LN output with a zero row above and below
constexpr int A_DQ = A_Y + (W + 2) * LDA_16;         // [17][20]
constexpr int A_DKV = A_DQ + W * LDA_16;             // [9][20]
constexpr int A_P = A_DKV + WKV * LDA_16;            // [17][12] attention probabilities
constexpr int A_FLOATS = A_P + W * 12;
constexpr int SMEM_BYTES = (W_FLOATS + WARPS * A_FLOATS) * 4;

struct Weights {
    const float *embed_w, *embed_b, *ln_g, *ln_b, *ln1_g, *ln1_b, *q_dw, *q_pw, *q_bias, *kv_dw, *kv_pw, *kv_bias, *out_w,
        *out_b, *ln2_g, *ln2_b, *ff1_w, *ff1_b, *ff2_w, *ff2_b;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}

enum { EPI_STORE = 0, EPI_ADD = 1, EPI_GELU = 2 };

// out[r][n] (op)= bias[n] + sum_k a[r][k] * w[n][k]   for r < ROWS, n < N; N a multiple of 32: lane owns N/32 columns
template <int N, int K, int LDW, int ROWS, int EPI>
__device__ __forceinline__ void wgemm_wide(const float* a, int lda, const float* w, const float* bias, float* out, int ldo, int lane) {
    constexpr int NC = N / 32;
    float acc[NC][ROWS];
    #pragma unroll
    for (int j = 0; j < NC; ++j) {
        const float b = bias[lane + 32 * j];
        #pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[j][r] = b;
    }
    #pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        float4 wv[NC];
        #pragma unroll
        for (int j = 0; j < NC; ++j) wv[j] = *reinterpret_cast<const float4*>(w + (lane + 32 * j) * LDW + k);
        #pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float4 av = *reinterpret_cast<const float4*>(a + r * lda + k);       // broadcast
            #pragma unroll
            for (int j = 0; j < NC; ++j) acc[j][r] = dot4(av, wv[j], acc[j][r]);
        }
    }
    #pragma unroll
    for (int j = 0; j < NC; ++j)
        #pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float v = acc[j][r];
            if (EPI == EPI_GELU) v = gelu_erf(v);
            out[r * ldo + lane + 32 * j] = v;
        }
}

// N == 16: lanes 0-15 take rows [0, 9), lanes 16-31 rows [9, 17)
template <int K, int LDW, int EPI>
__device__ __forceinline__ void wgemm_16(const float* a, int lda, const float* w, const float* bias, float* out, int ldo, int lane) {
    const int n = lane & 15, r0 = (lane >> 4) * 9;
    float acc[9];
    #pragma unroll
    for (int r = 0; r < 9; ++r) acc[r] = bias[n];
    #pragma unroll 2
    for (int k = 0; k < K; k += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(w + n * LDW + k);
        #pragma unroll
        for (int r = 0; r < 9; ++r) {
            const int rr = r0 + r < W ? r0 + r : W - 1;                                  // row 17 of the upper half: recomputed, not stored
            acc[r] = dot4(*reinterpret_cast<const float4*>(a + rr * lda + k), wv, acc[r]);
        }
    }
    #pragma unroll
    for (int r = 0; r < 9; ++r) {
        if (r0 + r >= W) break;
        float* o = out + (r0 + r) * ldo + n;
        *o = EPI == EPI_ADD ? *o + acc[r] : acc[r];
    }
}

// channel LN over 16 channels (M:57-67): population std, eps added to the std.  Lanes 0..16 take one row each.
__device__ __forceinline__ void ln16(const float* x, int ldx, const float* g, const float* b, float* y, int ldy, int lane) {
    if (lane < W) {
        float v[C], sum = 0.0f;
        #pragma unroll
        for (int q = 0; q < C / 4; ++q) {
            const float4 t = *reinterpret_cast<const float4*>(x + lane * ldx + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
        #pragma unroll
        for (int c = 0; c < C; ++c) sum += v[c];
        const float mean = sum / (float)C;
        float sq = 0.0f;
        #pragma unroll
        for (int c = 0; c < C; ++c) { const float d = v[c] - mean; sq += d * d; }
        const float denom = sqrtf(sq / (float)C) + 1e-5f;
        #pragma unroll
        for (int c = 0; c < C; ++c) y[lane * ldy + c] = (v[c] - mean) / denom * g[c] + b[c];
    }
}

__global__ void __launch_bounds__(WARPS * 32, 1)
aff_stage1_kernel(const float* __restrict__ x, Weights wt, float* __restrict__ out, int64_t batch) {
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* act = smem + W_FLOATS + warp * A_FLOATS;

    // ---- weights -> shared memory (once per CTA; the grid is persistent) ----
    for (int i = threadIdx.x; i < W_FLOATS; i += WARPS * 32) sw[i] = 0.0f;
    __syncthreads();
    for (int i = threadIdx.x; i < C * 3 * CIN; i += WARPS * 32) {                       // [c][tap*34 + ci] -> [c][tap*36 + ci]
        const int c = i / (3 * CIN), k = i - c * 3 * CIN, t = k / CIN, ci = k - t * CIN;
        sw[O_EMB + c * LDW_E + t * LDXP + ci] = wt.embed_w[i];
    }
    for (int i = threadIdx.x; i < INNER * C; i += WARPS * 32) sw[O_QPW + (i / C) * LDW_16 + (i % C)] = wt.q_pw[i];
    for (int i = threadIdx.x; i < 2 * INNER * C; i += WARPS * 32) sw[O_KVPW + (i / C) * LDW_16 + (i % C)] = wt.kv_pw[i];
    for (int i = threadIdx.x; i < C * INNER; i += WARPS * 32) sw[O_OUT + (i / INNER) * LDW_64 + (i % INNER)] = wt.out_w[i];
    for (int i = threadIdx.x; i < FF * C; i += WARPS * 32) sw[O_FF1 + (i / C) * LDW_16 + (i % C)] = wt.ff1_w[i];
    for (int i = threadIdx.x; i < C * FF; i += WARPS * 32) sw[O_FF2 + (i / FF) * LDW_64 + (i % FF)] = wt.ff2_w[i];
    {
        float* v = sw + O_VEC;
        const int i = threadIdx.x;
        if (i < C) {
            v[V_EMB_B + i] = wt.embed_b[i]; v[V_LN_G + i] = wt.ln_g[i]; v[V_LN_B + i] = wt.ln_b[i];
            v[V_LN1_G + i] = wt.ln1_g[i]; v[V_LN1_B + i] = wt.ln1_b[i]; v[V_OUT_B + i] = wt.out_b[i];
            v[V_LN2_G + i] = wt.ln2_g[i]; v[V_LN2_B + i] = wt.ln2_b[i]; v[V_FF2_B + i] = wt.ff2_b[i];
        }
        if (i < 3 * C) { v[V_QDW + i] = wt.q_dw[i]; v[V_KVDW + i] = wt.kv_dw[i]; }
        if (i < INNER) { v[V_QB + i] = wt.q_bias[i]; v[V_FF1_B + i] = wt.ff1_b[i]; }
        if (i < 2 * INNER) v[V_KVB + i] = wt.kv_bias[i];
    }
    __syncthreads();
    const float* vec = sw + O_VEC;

    float* sx = act + A_X;        // framed input, later q / o / ff
    float* skv = act + A_KV;
    float* sxs = act + A_XS;
    float* sy = act + A_Y;        // row 0 and row W+1 stay zero
    float* sdq = act + A_DQ;
    float* sdkv = act + A_DKV;
    float* sp = act + A_P;
    for (int i = lane; i < (W + 2) * LDA_16; i += 32) sy[i] = 0.0f;

    for (int64_t cand = (int64_t)blockIdx.x * WARPS + warp; cand < batch; cand += (int64_t)gridDim.x * WARPS) {
        // ---- input [33][34] -> zero-framed [35][36] ----
        for (int i = lane; i < (WIN + 2) * LDXP; i += 32) sx[i] = 0.0f;
        __syncwarp();
        const float* xc = x + cand * (WIN * CIN);
        for (int i = lane; i < WIN * CIN; i += 32) {
            const int r = i / CIN, ci = i - r * CIN;
            sx[(r + 1) * LDXP + ci] = xc[i];
        }
        __syncwarp();
        // ---- embed conv: output row w reads framed rows 2w .. 2w+2 = 108 contiguous floats ----
        wgemm_16<KE, LDW_E, EPI_STORE>(sx, 2 * LDXP, sw + O_EMB, vec + V_EMB_B, sdq, LDA_16, lane);
        __syncwarp();
        ln16(sdq, LDA_16, vec + V_LN_G, vec + V_LN_B, sxs, LDA_16, lane);
        __syncwarp();
        // ---- attention block ----
        ln16(sxs, LDA_16, vec + V_LN1_G, vec + V_LN1_B, sy + LDA_16, LDA_16, lane);
        __syncwarp();
        for (int o = lane; o < W * C; o += 32) {                                     // depth-wise 3-tap, stride 1 (BN scale folded)
            const int w = o >> 4, c = o & 15;
            float acc = 0.0f;
            if (w - 1 >= 0) acc = fmaf(sy[w * LDA_16 + c], vec[V_QDW + c], acc);
            acc = fmaf(sy[(w + 1) * LDA_16 + c], vec[V_QDW + C + c], acc);
            if (w + 1 < W) acc = fmaf(sy[(w + 2) * LDA_16 + c], vec[V_QDW + 2 * C + c], acc);
            sdq[w * LDA_16 + c] = acc;
        }
        for (int o = lane; o < WKV * C; o += 32) {                                   // stride 2
            const int r = o >> 4, c = o & 15, s0 = 2 * r - 1;
            float acc = 0.0f;
            if (s0 >= 0) acc = fmaf(sy[(s0 + 1) * LDA_16 + c], vec[V_KVDW + c], acc);
            acc = fmaf(sy[(s0 + 2) * LDA_16 + c], vec[V_KVDW + C + c], acc);
            if (s0 + 2 < W) acc = fmaf(sy[(s0 + 3) * LDA_16 + c], vec[V_KVDW + 2 * C + c], acc);
            sdkv[r * LDA_16 + c] = acc;
        }
        __syncwarp();
        float* sq = sx;                                                              // [17][68]; the input is no longer needed
        wgemm_wide<INNER, C, LDW_16, W, EPI_STORE>(sdq, LDA_16, sw + O_QPW, vec + V_QB, sq, LDA_64, lane);
        wgemm_wide<2 * INNER, C, LDW_16, WKV, EPI_STORE>(sdkv, LDA_16, sw + O_KVPW, vec + V_KVB, skv, LDA_KV, lane);
        __syncwarp();
        for (int p = lane; p < W * WKV; p += 32) {                                   // q k^T (the 64^-0.5 scale is folded into q)
            const int i = p / WKV, j = p - i * WKV;
            float acc = 0.0f;
            #pragma unroll
            for (int d = 0; d < INNER; d += 4)
                acc = dot4(*reinterpret_cast<const float4*>(sq + i * LDA_64 + d), *reinterpret_cast<const float4*>(skv + j * LDA_KV + d), acc);
            sp[i * 12 + j] = acc;
        }
        __syncwarp();
        if (lane < W) {
            float mx = -3.402823466e38f;
            #pragma unroll
            for (int j = 0; j < WKV; ++j) mx = fmaxf(mx, sp[lane * 12 + j]);
            float e[WKV], sum = 0.0f;
            #pragma unroll
            for (int j = 0; j < WKV; ++j) { e[j] = expf(sp[lane * 12 + j] - mx); sum += e[j]; }
            const float inv = 1.0f / sum;
            #pragma unroll
            for (int j = 0; j < WKV; ++j) sp[lane * 12 + j] = e[j] * inv;
        }
        __syncwarp();
        {                                                                            // o = p v, overwrites q (lane owns columns)
            float v0[WKV], v1[WKV];
            #pragma unroll
            for (int j = 0; j < WKV; ++j) { v0[j] = skv[j * LDA_KV + INNER + lane]; v1[j] = skv[j * LDA_KV + INNER + lane + 32]; }
            __syncwarp();
            #pragma unroll
            for (int i = 0; i < W; ++i) {
                float o0 = 0.0f, o1 = 0.0f;
                #pragma unroll
                for (int j = 0; j < WKV; ++j) {
                    const float pij = sp[i * 12 + j];
                    o0 = fmaf(pij, v0[j], o0);
                    o1 = fmaf(pij, v1[j], o1);
                }
                sq[i * LDA_64 + lane] = o0;
                sq[i * LDA_64 + lane + 32] = o1;
            }
        }
        __syncwarp();
        wgemm_16<INNER, LDW_64, EPI_ADD>(sq, LDA_64, sw + O_OUT, vec + V_OUT_B, sxs, LDA_16, lane);
        __syncwarp();
        // ---- feed-forward block ----
        ln16(sxs, LDA_16, vec + V_LN2_G, vec + V_LN2_B, sdq, LDA_16, lane);
        __syncwarp();
        wgemm_wide<FF, C, LDW_16, W, EPI_GELU>(sdq, LDA_16, sw + O_FF1, vec + V_FF1_B, sq, LDA_64, lane);
        __syncwarp();
        wgemm_16<FF, LDW_64, EPI_ADD>(sq, LDA_64, sw + O_FF2, vec + V_FF2_B, sxs, LDA_16, lane);
        __syncwarp();
        float* oc = out + cand * (W * C);
        for (int i = lane; i < W * C; i += 32) oc[i] = sxs[(i >> 4) * LDA_16 + (i & 15)];
        __syncwarp();
    }
}

}  // namespace s1

bool aff_stage1_fused_supported(const CvtStage& st) {
    return st.c == s1::C && st.cin == s1::CIN && st.win == s1::WIN && st.heads == 1 && st.depth == 1;
}

// x: fp32 [n, 33, 34] (rescaled input); out: fp32 [n, 17, 16] = the stage output (input of the second embed conv)
int launch_aff_stage1(const CvtStage& st, const float* x, float* out, int64_t n, cudaStream_t s) {
    if (n <= 0) return 0;
    CTO_REQUIRE(aff_stage1_fused_supported(st), "aff_stage1: stage shape C=%d Cin=%d W=%d heads=%d depth=%d is not the fused one",
                st.c, st.cin, st.win, st.heads, st.depth);
    const int sm_count = device_sm_count();
    CTO_REQUIRE(sm_count > 0, "aff_stage1: no CUDA device");
    CTO_CHECK(set_max_dynamic_smem(s1::aff_stage1_kernel, s1::SMEM_BYTES));
    const CvtLayer& L = st.layers[0];
    s1::Weights w{st.embed_w, st.embed_b, st.ln_g, st.ln_b, L.ln1_g, L.ln1_b, L.q_dw, L.q_pw, L.q_bias, L.kv_dw, L.kv_pw,
                  L.kv_bias, L.out_w, L.out_b, L.ln2_g, L.ln2_b, L.ff1_w, L.ff1_b, L.ff2_w, L.ff2_b};
    const int64_t want = (n + s1::WARPS - 1) / s1::WARPS;
    const int grid = (int)(want < sm_count ? want : sm_count);
    s1::aff_stage1_kernel<<<grid, s1::WARPS * 32, s1::SMEM_BYTES, s>>>(x, w, out, n);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
