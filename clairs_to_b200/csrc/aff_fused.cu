// Fused CvT transformer layers of the AFF network on tcgen05 tensor cores (sm_100a).
//
// Reference arithmetic: clairs/model.py:134-147 (Transformer), 102-132 (Attention), 91-100 (DepthWiseConv2d),
// 78-89 (FeedForward), 57-67 (channel LayerNorm).  One launch runs ALL layers of one CvT stage on the stage's
// residual stream x[n, W, C] (fp32, channels-last, in place):
//
//   for every layer:   x += Attention(LN1(x));   x += FeedForward(LN2(x))
//
// Round 1 ran each of these as 8 kernels per layer (LN+dwconv, q GEMM, kv GEMM, attention, out-proj GEMM, LN,
// FF1 GEMM, FF2 GEMM), every intermediate making an HBM round trip: ~60 KB of traffic per candidate and layer for
// 2.5 KB of state, 41 % of the whole step at 5 % of the tensor peak.  Here the tile's state never
// leaves the SM:
//
//   * a CTA owns 128 rows (= TC candidates x W positions; a candidate never straddles a warp) and keeps the
//     residual stream x IN TENSOR MEMORY: columns [0, C) of the 128 TMEM lanes.  The out-projection and the second
//     feed-forward GEMM accumulate straight onto it (D += A B), so "+ x" costs nothing; their biases are carried in
//     cumulative-bias vectors precomputed at load time and added whenever x is read (x_true = x_tmem + cb);
//   * dense products are bf16x3 split products (hi*hi + mid*hi + hi*mid, fp32 accumulate; see gemm_tc.cu), M = 128,
//     N = 64 per MMA, A operand = bf16 hi / mid planes in shared memory written by the compute warps in the UMMA
//     K-major 128-byte-swizzle layout, B operand = pre-swizzled 16 KB weight tiles streamed from L2 by bulk async
//     copies (cp.async.bulk) through a ring, in exactly the order the MMA warp consumes them (prepacked at load time);
//   * 128 compute threads own one row each (TMEM lane = row): LN over the channels of the row in registers,
//     depth-wise 3-tap convolutions and the attention's key/value exchange by warp shuffles (the stride-2 conv of
//     to_kv is the stride-1 conv evaluated at even positions, so k / v of key j live in the lane of row 2j of the
//     same candidate), softmax, GELU (Abramowitz-Stegun 7.1.26 erf, |err| < 2e-7); a second group of 128 threads
//     takes every other feed-forward chunk so that the GELU epilogue keeps up with the tensor pipe;
//   * per head: q | k | v accumulators (192 TMEM columns, double buffered), attention output -> A planes of the
//     out-projection; feed-forward hidden units in chunks of 64 (six 64-column accumulators), GELU -> A planes of FF2.
//
// Warp roles: 0 weight producer, 1 MMA issuer (+ TMEM allocation), 2-5 compute group A (everything), 6-9 group B
// (odd feed-forward chunks).  Every mbarrier wait is bounded: on a timeout the kernel records which barrier stalled
// in p.dbg and unwinds instead of hanging the GPU.
#include "engine.cuh"
#include "gru_ptx.cuh"
#include <cstring>

namespace cto {

namespace fz {

using namespace tc;

constexpr int ROWS = 128;
constexpr int TILE_A = ROWS * 128;            // 16 KB: one [128 rows x 64 k] bf16 plane k-block
constexpr int WTILE = 64 * 128;               // 8 KB:  one [64 n x 64 k] bf16 weight plane tile
constexpr int WSTAGE = 2 * WTILE;             // hi | mid
constexpr int THREADS = 320;                  // 10 warps: 204 registers per thread
constexpr int NB = 6;                         // feed-forward hidden accumulators (64 TMEM columns each)
constexpr int LAG = 2;                        // FF1 MMAs run one pair of hidden chunks ahead of FF2
constexpr uint32_t SPIN_LIMIT = 1u << 22;

template <int C>
struct Cfg {
    static constexpr int CP = C < 64 ? 64 : C;                // channels padded to the 64-wide k-block
    static constexpr int KB = CP / 64;                        // k-blocks of a C-wide contraction
    static constexpr int NT = CP / 64;                        // 64-column output tiles of a C-wide product
    static constexpr int NH = (4 * C) / 64;                   // feed-forward hidden chunks
    static constexpr int NPC = C >= 128 ? 1 : 2;              // attention-output / hidden plane buffers
    static constexpr int WST = C >= 128 ? 3 : 5;              // weight ring stages
    static constexpr int VEC = 16 * 1024;                     // the layer's small vectors (LN, taps, biases) in shared memory
    static constexpr int PLANE = 2 * KB * TILE_A;             // hi k-blocks | mid k-blocks
    static constexpr int PCBUF = 2 * TILE_A;                  // hi | mid, one k-block
    static constexpr int NFF = NPC + PLANE / PCBUF;           // plane buffers of the feed-forward phase: pc[] + the (dead) dkv planes
    static constexpr int SMEM = 2 * PLANE + NPC * PCBUF + WST * WSTAGE + VEC + 1024 /*align*/ + 1024 /*barriers*/;
    // float offsets inside the layer's vector block (host prepack and kernel agree); + 16 C: bq [inner], bkv [2 inner]
    static constexpr int O_LN1G = 0, O_LN1B = C, O_TQ = 2 * C, O_TK = 5 * C, O_LN2G = 8 * C, O_LN2B = 9 * C, O_CB1 = 10 * C,
                         O_CB2 = 11 * C, O_B1 = 12 * C, O_BQ = 16 * C;
    static_assert(C == 32 || C == 64 || C == 128, "stage width");
    static_assert((4 * C) % 64 == 0 && NH >= LAG, "hidden chunking");
};

struct Params {
    float* x;                     // [n * W, C] fp32, in place
    const uint8_t* wstream;       // prepacked weight tiles, layer l at l * layer_bytes
    const float* vblocks;         // per layer one block of Cfg::VEC bytes: the layer's small vectors (Cfg::O_*)
    int vec_bytes;                // live bytes of a block (multiple of 16)
    const float* cb_final;        // [C] sum of every out-projection / FF2 bias of the stage (added when x leaves TMEM)
    long long layer_bytes;
    long long n;
    int W, WKV, heads, depth;
    int tiles, cpw;               // tiles of 4 * cpw candidates; cpw = candidates per warp = 32 / W
    int* dbg;                     // [8] abort flag + diagnostics
    long long* timing;            // optional [32] per-phase clock64() sums of CTA 0 (profiles/phase_timing_aff.py); nullptr in production
};

// ---- PTX helpers not in gru_ptx.cuh ------------------------------------------------------------------------
__device__ __forceinline__ void f_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(g_smem_u32(dst)), "l"(src), "r"(bytes), "r"(g_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool f_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(g_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: returns false when the kernel is being aborted (another thread timed out, or this one just did)
__device__ __noinline__ bool f_wait_slow(uint64_t* bar, uint32_t parity, int code, int* dbg) {
    for (uint32_t tries = 0; tries < SPIN_LIMIT; ++tries) {
        if (f_try_wait(bar, parity)) return true;
        if ((tries & 255u) == 255u && *reinterpret_cast<volatile int*>(dbg) != 0) return false;
    }
    if (atomicCAS(dbg, 0, code) == 0) {
        dbg[1] = (int)blockIdx.x;
        dbg[2] = (int)threadIdx.x;
        dbg[3] = (int)parity;
    }
    return false;
}
__device__ __forceinline__ bool f_wait(uint64_t* bar, uint32_t parity, int code, int* dbg) {
    if (f_try_wait(bar, parity)) return true;
    return f_wait_slow(bar, parity, code, dbg);
}
__device__ __forceinline__ void f_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void f_tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    #pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void f_tmem_st32(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
}
__device__ __forceinline__ void f_tmem_ld8f(uint32_t taddr, float* v) {
    uint32_t r[8];
    g_tmem_ld8(taddr, r);
    #pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// eight consecutive floats of a (shared-memory) vector: every lane reads the same address, a broadcast
__device__ __forceinline__ void f_lds8(const float* src, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void f_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void f_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void f_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t f_pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void f_split2(float x0, float x1, uint32_t& hi, uint32_t& mid) {
    hi = f_pack(x0, x1);
    mid = f_pack(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xFFFF0000u));
}
// eight consecutive k of one row -> one 16-byte chunk of the hi plane and of the mid plane (UMMA K-major, 128-byte swizzle)
__device__ __forceinline__ void f_store8(uint8_t* hi_tile, uint8_t* mid_tile, int row, int chunk, const float* v) {
    uint4 h, m;
    f_split2(v[0], v[1], h.x, m.x);
    f_split2(v[2], v[3], h.y, m.y);
    f_split2(v[4], v[5], h.z, m.z);
    f_split2(v[6], v[7], h.w, m.w);
    const uint32_t off = (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4));
    *reinterpret_cast<uint4*>(hi_tile + off) = h;
    *reinterpret_cast<uint4*>(mid_tile + off) = m;
}
__device__ __forceinline__ float f_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float f_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exact-erf GELU (nn.GELU default, clairs/model.py:83) with erf from Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7)
__device__ __forceinline__ float f_gelu(float x) {
    const float ax = fabsf(x) * 0.70710678118654752440f;
    const float t = f_rcp(fmaf(0.3275911f, ax, 1.0f));
    float p = fmaf(t, 1.061405429f, -1.453152027f);
    p = fmaf(t, p, 1.421413741f);
    p = fmaf(t, p, -0.284496736f);
    p = fmaf(t, p, 0.254829592f);
    p *= t;
    const float e = f_ex2(-ax * ax * 1.4426950408889634f);
    const float erf_abs = fmaf(-p, e, 1.0f);
    const float erf = copysignf(erf_abs, x);
    return 0.5f * x * (1.0f + erf);
}
// eight elements at once, stage by stage, so that the eight MUFU / FMA chains interleave
__device__ __forceinline__ void f_gelu8(float* v, const float* bias) {
    float x[8], ax[8], t[8], pl[8], e[8];
    #pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = v[i] + bias[i]; ax[i] = fabsf(x[i]) * 0.70710678118654752440f; }
    #pragma unroll
    for (int i = 0; i < 8; ++i) t[i] = f_rcp(fmaf(0.3275911f, ax[i], 1.0f));
    #pragma unroll
    for (int i = 0; i < 8; ++i) e[i] = f_ex2(-ax[i] * ax[i] * 1.4426950408889634f);
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        float q = fmaf(t[i], 1.061405429f, -1.453152027f);
        q = fmaf(t[i], q, 1.421413741f);
        q = fmaf(t[i], q, -0.284496736f);
        q = fmaf(t[i], q, 0.254829592f);
        pl[i] = q * t[i];
    }
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float erf = copysignf(fmaf(-pl[i], e[i], 1.0f), x[i]);
        v[i] = 0.5f * x[i] * (1.0f + erf);
    }
}
// D=f32, A=B=bf16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t f_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// barrier codes reported in dbg[0] on a timeout
enum { B_WFULL = 1, B_WEMPTY, B_AB, B_QKVFULL, B_QKVFREE, B_PCREADY, B_PCFREE, B_XREADY, B_Y2, B_HIDFULL, B_HIDFREE, B_LAYER, B_XLOADED, B_VEC };

// channel LayerNorm (clairs/model.py:57-67): population std over the C channels of a row, eps added to the std.
// The row lives in tensor memory (x_true = x_tmem + cb); statistics with the shifted-data formulas (shift = first channel).
template <int C>
__device__ __forceinline__ void f_row_stats(uint32_t trow, const float* __restrict__ cb, float& mean, float& inv) {
    float s1 = 0.0f, s2 = 0.0f, shift = 0.0f;
    #pragma unroll 1
    for (int c0 = 0; c0 < C; c0 += 32) {
        float v[32];
        f_tmem_ld32(trow + (uint32_t)c0, v);
        f_wait_ld();
        #pragma unroll
        for (int c4 = 0; c4 < 32; c4 += 4) {
            const float4 t = *reinterpret_cast<const float4*>(cb + c0 + c4);
            v[c4] += t.x; v[c4 + 1] += t.y; v[c4 + 2] += t.z; v[c4 + 3] += t.w;
        }
        if (c0 == 0) shift = v[0];
        float a1 = 0.0f, a2 = 0.0f, b1 = 0.0f, b2 = 0.0f;
        #pragma unroll
        for (int c = 0; c < 32; c += 2) {
            const float d0 = v[c] - shift, d1 = v[c + 1] - shift;
            a1 += d0; a2 = fmaf(d0, d0, a2);
            b1 += d1; b2 = fmaf(d1, d1, b2);
        }
        s1 += a1 + b1;
        s2 += a2 + b2;
    }
    const float m = s1 * (1.0f / (float)C);
    mean = shift + m;
    const float var = fmaxf(s2 * (1.0f / (float)C) - m * m, 0.0f);
    inv = 1.0f / (sqrtf(var) + 1e-5f);
}
// 8 channels of the row: (x_tmem + cb - mean) * inv * g + b
__device__ __forceinline__ void f_normalise8(float* v, const float* __restrict__ cb, const float* __restrict__ g,
                                             const float* __restrict__ b, float mean, float inv) {
    float t[8], gv[8], bv[8];
    f_lds8(cb, t); f_lds8(g, gv); f_lds8(b, bv);
    #pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf((v[e] + t[e] - mean) * inv, gv[e], bv[e]);
}
// 32 channels of the row: (x_tmem + cb - mean) * inv * g + b
__device__ __forceinline__ void f_normalise32(float* v, const float* __restrict__ cb, const float* __restrict__ g,
                                              const float* __restrict__ b, float mean, float inv) {
    #pragma unroll
    for (int c4 = 0; c4 < 32; c4 += 4) {
        const float4 t = *reinterpret_cast<const float4*>(cb + c4);
        const float4 gv = *reinterpret_cast<const float4*>(g + c4), bv = *reinterpret_cast<const float4*>(b + c4);
        v[c4] = fmaf((v[c4] + t.x - mean) * inv, gv.x, bv.x);
        v[c4 + 1] = fmaf((v[c4 + 1] + t.y - mean) * inv, gv.y, bv.y);
        v[c4 + 2] = fmaf((v[c4 + 2] + t.z - mean) * inv, gv.z, bv.z);
        v[c4 + 3] = fmaf((v[c4 + 3] + t.w - mean) * inv, gv.w, bv.w);
    }
}

template <int C, int W>
__global__ void __launch_bounds__(THREADS, 1) aff_layers_kernel(const Params p) {
    using K = Cfg<C>;
    constexpr int WKV = (W + 1) / 2;                  // keys per candidate (stride-2 projection)
    constexpr int CPW = 32 / W;                       // candidates per warp
    constexpr int CP = K::CP, KB = K::KB, NT = K::NT, NH = K::NH, NPC = K::NPC, WST = K::WST;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (g_smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* pa = base;                               // dq planes, later the LN2 output planes
    uint8_t* pb = pa + K::PLANE;                      // dkv planes
    uint8_t* pc = pb + K::PLANE;                      // attention output / hidden chunk planes [NPC]
    uint8_t* wring_hi = pc + NPC * K::PCBUF;          // [WST] hi tiles; consecutive stages are contiguous: two of them
    uint8_t* wring_mid = wring_hi + WST * WTILE;      // [WST] mid tiles   form the B operand of one N = 128 MMA
    float* sv = reinterpret_cast<float*>(wring_mid + WST * WTILE);    // the current layer's vector block
    uint64_t* bars = reinterpret_cast<uint64_t*>(wring_mid + WST * WTILE + K::VEC);
    uint64_t* w_full = bars;                          // [WST]
    uint64_t* w_empty = w_full + WST;                 // [WST]
    uint64_t* ab_ready = w_empty + WST;
    uint64_t* qkv_full = ab_ready + 1;                // [2]
    uint64_t* qkv_free = qkv_full + 2;                // [2]
    uint64_t* pc_ready = qkv_free + 2;                // [2]
    uint64_t* pc_free = pc_ready + 2;                 // [2 writer groups][2 buffers]
    uint64_t* x_ready = pc_free + 4;
    uint64_t* y2_ready = x_ready + 1;
    uint64_t* hid_full = y2_ready + 1;                // [NB]
    uint64_t* hid_free = hid_full + NB;               // [NB]
    uint64_t* layer_done = hid_free + NB;
    uint64_t* x_loaded = layer_done + 1;
    uint64_t* vec_full = x_loaded + 1;
    uint64_t* ff_ready = vec_full + 1;                // [3]
    uint64_t* ff_free = ff_ready + 3;                 // [2 writer groups][3 buffers]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ff_free + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* dbg = p.dbg;
    const bool tim = p.timing != nullptr && blockIdx.x == 0;
    long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = tim ? clock64() : 0;
    #define FTOC(i) do { if (tim) { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } } while (0)
    const int heads = p.heads, depth = p.depth;
    const int tiles_per_layer = heads * (3 * KB + NT) + NH * (KB + NT);
    if (threadIdx.x == 0) {
        for (int s = 0; s < WST; ++s) { g_mbar_init(&w_full[s], 1); g_mbar_init(&w_empty[s], 1); }
        g_mbar_init(ab_ready, 8);
        for (int b = 0; b < 2; ++b) {
            g_mbar_init(&qkv_full[b], 1); g_mbar_init(&qkv_free[b], 4);
            g_mbar_init(&pc_ready[b], 4);
        }
        for (int b = 0; b < 4; ++b) g_mbar_init(&pc_free[b], 1);
        g_mbar_init(x_ready, 1);
        g_mbar_init(y2_ready, 8);
        for (int b = 0; b < NB; ++b) { g_mbar_init(&hid_full[b], 1); g_mbar_init(&hid_free[b], 4); }
        g_mbar_init(layer_done, 1);
        g_mbar_init(x_loaded, 4);
        g_mbar_init(vec_full, 1);
        for (int b = 0; b < 3; ++b) g_mbar_init(&ff_ready[b], 4);
        for (int b = 0; b < 6; ++b) g_mbar_init(&ff_free[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    f_fence_before();
    __syncthreads();
    f_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- weight producer: the layer's tiles in consumption order through the ring ----
        uint32_t wit = 0, lit = 0;
        bool alive = true;
        for (int tile = blockIdx.x; tile < p.tiles && alive; tile += gridDim.x) {
            for (int l = 0; l < depth && alive; ++l, ++lit) {
                const uint8_t* src = p.wstream + (long long)l * p.layer_bytes;
                for (int i = 0; i < tiles_per_layer; ++i, ++wit) {
                    if (i == (tiles_per_layer < WST ? tiles_per_layer - 1 : WST - 1)) {
                        // The layer's vector block, after the first ring-full of weight tiles (those only wait for slots the
                        // previous layer frees): the previous layer's vectors are dead once its last MMA has retired.
                        if (lit >= 1 && !f_wait(layer_done, (lit - 1) & 1, B_LAYER, dbg)) { alive = false; break; }
                        if (g_elect_one()) {
                            g_mbar_expect_tx(vec_full, (uint32_t)p.vec_bytes);
                            f_bulk_load(sv, reinterpret_cast<const uint8_t*>(p.vblocks) + (long long)l * K::VEC, (uint32_t)p.vec_bytes, vec_full);
                        }
                        __syncwarp();
                    }
                    const int s = wit % WST;
                    if (!f_wait(&w_empty[s], ((wit / WST) & 1) ^ 1, B_WEMPTY, dbg)) { alive = false; break; }
                    if (g_elect_one()) {
                        g_mbar_expect_tx(&w_full[s], WSTAGE);
                        f_bulk_load(wring_hi + s * WTILE, src + (long long)i * WSTAGE, WTILE, &w_full[s]);
                        f_bulk_load(wring_mid + s * WTILE, src + (long long)i * WSTAGE + WTILE, WTILE, &w_full[s]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (warp-uniform control flow, one elected lane issues) ----
        const uint32_t idesc64 = f_idesc(64), idesc128 = f_idesc(128);
        uint32_t wit = 0, lit = 0;
        bool alive = true;
        // D[128 x 64 (x 2)] (+)= A[128 x 64 nkb] * W^T with nkb (x 2) weight tiles from the ring.  pair: two consecutive
        // tiles (output columns d_col .. +64 and +64 .. +128) feed ONE N = 128 MMA per k-step when their ring stages are
        // adjacent (the M = 128, N = 64 MMA re-reads the 4 KB A operand for 2 KB of B: shared-memory-bandwidth bound at
        // 48 cycles; N = 128 moves 8 KB per 64 cycles), two N = 64 MMAs when the pair wraps around the ring.
        auto product = [&](uint32_t d_col, const uint8_t* a_hi, const uint8_t* a_mid, int nkb, bool acc_first, bool pair) -> bool {
            for (int kb = 0; kb < nkb; ++kb) {
                const int s0 = wit % WST, s1 = (wit + 1) % WST;
                FTOC(0);
                if (!f_wait(&w_full[s0], (wit / WST) & 1, B_WFULL, dbg)) return false;
                if (pair && !f_wait(&w_full[s1], ((wit + 1) / WST) & 1, B_WFULL, dbg)) return false;
                FTOC(1);
                f_fence_after();
                const uint64_t d_ahi = g_desc_k_sw128(g_smem_u32(a_hi + kb * TILE_A));
                const uint64_t d_amid = g_desc_k_sw128(g_smem_u32(a_mid + kb * TILE_A));
                const uint64_t d_whi = g_desc_k_sw128(g_smem_u32(wring_hi + s0 * WTILE));
                const uint64_t d_wmid = g_desc_k_sw128(g_smem_u32(wring_mid + s0 * WTILE));
                const bool fused = pair && s1 == s0 + 1;
                const uint32_t idesc = fused ? idesc128 : idesc64;
                const uint32_t first = (acc_first || kb) ? 1u : 0u;
                if (g_elect_one()) {
                    #pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t o = (uint64_t)(k * 2);          // 16 bf16 = 32 bytes along the swizzle row
                        f_mma(d_col, d_ahi + o, d_whi + o, idesc, (first || k) ? 1u : 0u);
                        f_mma(d_col, d_amid + o, d_whi + o, idesc, 1u);
                        f_mma(d_col, d_ahi + o, d_wmid + o, idesc, 1u);
                    }
                    if (pair && !fused) {                              // the pair wraps: second tile from stage 0
                        const uint64_t e_whi = g_desc_k_sw128(g_smem_u32(wring_hi + s1 * WTILE));
                        const uint64_t e_wmid = g_desc_k_sw128(g_smem_u32(wring_mid + s1 * WTILE));
                        #pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t o = (uint64_t)(k * 2);
                            f_mma(d_col + 64, d_ahi + o, e_whi + o, idesc64, (first || k) ? 1u : 0u);
                            f_mma(d_col + 64, d_amid + o, e_whi + o, idesc64, 1u);
                            f_mma(d_col + 64, d_ahi + o, e_wmid + o, idesc64, 1u);
                        }
                    }
                    g_commit(&w_empty[s0]);
                    if (pair) g_commit(&w_empty[s1]);
                }
                __syncwarp();
                wit += pair ? 2 : 1;
            }
            return true;
        };
        // x[:, 0 .. C) += planes * W^T (out-projection of a head / FF2 of a hidden chunk).  Plane buffers are handed over
        // with one "ready" barrier per buffer (writer group -> this warp) and one "free" barrier per (writer group,
        // buffer) (this warp -> the group that writes the buffer next: a barrier that two groups wait on in turns would
        // let the group that skipped a phase alias an old parity).  All buffers are free at the start of every layer.
        uint32_t pc_uses[2] = {0, 0}, ff_uses[3] = {0, 0, 0};
        auto onto_x = [&](const uint8_t* a, uint64_t* ready, uint32_t& uses, uint64_t* free_next) -> bool {
            FTOC(0);
            if (!f_wait(ready, uses & 1, B_PCREADY, dbg)) return false;
            ++uses;
            FTOC(2);
            f_fence_after();
            if (!product(tmem_base, a, a + TILE_A, 1, true, NT == 2)) return false;
            if (free_next != nullptr && g_elect_one()) g_commit(free_next);
            __syncwarp();
            return true;
        };
        auto ff_buf = [&](int fb) -> const uint8_t* { return fb < NPC ? pc + fb * K::PCBUF : pb + (fb - NPC) * K::PCBUF; };
        for (int tile = blockIdx.x; tile < p.tiles && alive; tile += gridDim.x) {
            for (int l = 0; l < depth && alive; ++l, ++lit) {
                FTOC(0);
                if (!f_wait(ab_ready, lit & 1, B_AB, dbg)) { alive = false; break; }
                FTOC(3);
                f_fence_after();
                // out-projection of head h: buffer h % NPC, next written by head h + NPC (group = its accumulator buffer)
                auto out_proj = [&](int h) -> bool {
                    const int pbuf = h % NPC;
                    const bool more = h + NPC < heads;
                    const int next_grp = (int)((lit * (uint32_t)heads + h + NPC) & 1u);
                    return onto_x(pc + pbuf * K::PCBUF, &pc_ready[pbuf], pc_uses[pbuf], more ? &pc_free[next_grp * 2 + pbuf] : nullptr);
                };
                for (int h = 0; h < heads && alive; ++h) {
                    const uint32_t hit = lit * (uint32_t)heads + h, b = hit & 1;
                    FTOC(0);
                    if (!f_wait(&qkv_free[b], ((hit >> 1) & 1) ^ 1, B_QKVFREE, dbg)) { alive = false; break; }
                    FTOC(4);
                    f_fence_after();
                    const uint32_t acc = tmem_base + (uint32_t)(CP + b * 192);
                    alive = product(acc, pa, pa + KB * TILE_A, KB, false, false) &&          // q_h
                            product(acc + 64, pb, pb + KB * TILE_A, KB, false, true);         // k_h | v_h
                    if (!alive) break;
                    if (g_elect_one()) g_commit(&qkv_full[b]);
                    __syncwarp();
                    if (h >= 1) alive = out_proj(h - 1);
                }
                if (!alive) break;
                if (!out_proj(heads - 1)) { alive = false; break; }
                if (g_elect_one()) g_commit(x_ready);
                __syncwarp();
                FTOC(0);
                if (!f_wait(y2_ready, lit & 1, B_Y2, dbg)) { alive = false; break; }
                FTOC(5);
                f_fence_after();
                // FF2 of hidden chunk j: buffer j % NFF, next written by chunk j + NFF (group = chunk parity)
                auto ff2 = [&](int j) -> bool {
                    const int fb = j % K::NFF;
                    const bool more = j + K::NFF < NH;
                    return onto_x(ff_buf(fb), &ff_ready[fb], ff_uses[fb], more ? &ff_free[((j + K::NFF) & 1) * 3 + fb] : nullptr);
                };
                for (int j = 0; j < NH && alive; j += 2) {                 // hidden chunks in pairs: one N = 128 product
                    const uint32_t cit = lit * (uint32_t)NH + j, hb = cit % NB;
                    FTOC(0);
                    if (!f_wait(&hid_free[hb], ((cit / NB) & 1) ^ 1, B_HIDFREE, dbg) ||
                        !f_wait(&hid_free[hb + 1], (((cit + 1) / NB) & 1) ^ 1, B_HIDFREE, dbg)) { alive = false; break; }
                    FTOC(6);
                    f_fence_after();
                    if (!product(tmem_base + (uint32_t)(CP + hb * 64), pa, pa + KB * TILE_A, KB, false, true)) { alive = false; break; }
                    if (g_elect_one()) { g_commit(&hid_full[hb]); g_commit(&hid_full[hb + 1]); }
                    __syncwarp();
                    if (j >= 2) alive = ff2(j - 2) && ff2(j - 1);
                }
                if (!alive) break;
                alive = ff2(NH - 2) && ff2(NH - 1);
                if (!alive) break;
                if (g_elect_one()) g_commit(layer_done);
                __syncwarp();
            }
        }
        FTOC(0);
        if (tim && lane == 0) { for (int i = 0; i < 7; ++i) p.timing[i] = tacc[i]; p.timing[7] = (long long)lit; }
    } else {
        // ---- compute groups 0 (warps 2-5) and 1 (warps 6-9): one thread per row in each (TMEM lane = row).  The two
        // groups split every phase: LayerNorm / convolution channel chunks, attention heads, hidden chunks.
        // Every loop over columns is ROLLED and works on 8 columns read from tensor memory per iteration: the first,
        // fully unrolled version of this code was 360 KB of SASS and spent a third of its issue slots waiting for
        // instruction fetch (ncu: stall_no_inst 33 %).
        const int grp = (warp - 2) >> 2;
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const int cl = lane / W, w = lane - cl * W;                    // candidate inside the warp, position
        const bool lane_ok = lane < CPW * W;
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
        const bool has_up = lane_ok && w > 0, has_dn = lane_ok && w < W - 1, kv_row = lane_ok && !(w & 1);
        const int src0 = cl * W;                                       // lane of the candidate's first row: key j lives in lane src0 + 2j
        uint8_t* pa_mid = pa + KB * TILE_A;
        uint8_t* pb_mid = pb + KB * TILE_A;
        const bool tg = tim && grp == 0;                               // group 0 reports its phase times
        #define GTOC(i) do { if (tg) { const long long _t = clock64(); tacc[i] += _t - tprev; tprev = _t; } } while (0)
        uint32_t lit = 0, tit = 0;
        uint32_t pcf_cnt[2] = {0, 0}, fff_cnt[3] = {0, 0, 0};          // signalled waits so far on pc_free / ff_free [grp][buffer]
        bool alive = true;
        for (int tile = blockIdx.x; tile < p.tiles && alive; tile += gridDim.x, ++tit) {
            const long long cand = ((long long)tile * 4 + quad) * CPW + cl;
            const bool ok = lane_ok && cand < p.n;
            float* xg = p.x + (((long long)tile * 4 + quad) * CPW * W + lane) * (long long)C;
            if (grp == 0) {
                // x tile: global -> registers -> TMEM columns [0, CP)
                #pragma unroll 1
                for (int c0 = 0; c0 < CP; c0 += 32) {
                    float v[32];
                    #pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok && c0 < C) t = *reinterpret_cast<const float4*>(xg + c0 + 4 * q);
                        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                    }
                    f_tmem_st32(trow + (uint32_t)c0, v);
                }
                f_wait_st();
                f_fence_before();
                __syncwarp();
                if (lane == 0) g_mbar_arrive(x_loaded);
                GTOC(0);
            } else {
                if (!f_wait(x_loaded, tit & 1, B_XLOADED, dbg)) { alive = false; break; }
                f_fence_after();
            }
            for (int l = 0; l < depth && alive; ++l, ++lit) {
                if (!f_wait(vec_full, lit & 1, B_VEC, dbg)) { alive = false; break; }
                // ---------- LN1 + depth-wise convolutions -> dq (PA), dkv (PB) ----------
                // The row is read from tensor memory twice: statistics (both groups, whole row), then normalise + convolve
                // + store, 8 channels per iteration, the 8-channel chunks alternating between the groups.
                {
                    float mean, inv;
                    f_row_stats<C>(trow, sv + K::O_CB1, mean, inv);
                    #pragma unroll 1
                    for (int c8 = grp * 8; c8 < CP; c8 += 16) {
                        float dq[8], dk[8];
                        if (c8 < C) {
                            float v[8];
                            f_tmem_ld8f(trow + (uint32_t)c8, v);
                            f_wait_ld();
                            f_normalise8(v, sv + K::O_CB1 + c8, sv + K::O_LN1G + c8, sv + K::O_LN1B + c8, mean, inv);
                            // neighbour rows by warp shuffles, issued as a batch (a shuffle feeding its own FMA serialises
                            // on the shuffle latency: measured 10 cycles per instruction)
                            float up[8], dn[8];
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) up[e] = __shfl_up_sync(0xffffffffu, v[e], 1);
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) dn[e] = __shfl_down_sync(0xffffffffu, v[e], 1);
                            float t0[8], t1[8], t2[8], k0[8], k1[8], k2[8];
                            f_lds8(sv + K::O_TQ + c8, t0); f_lds8(sv + K::O_TQ + C + c8, t1); f_lds8(sv + K::O_TQ + 2 * C + c8, t2);
                            f_lds8(sv + K::O_TK + c8, k0); f_lds8(sv + K::O_TK + C + c8, k1); f_lds8(sv + K::O_TK + 2 * C + c8, k2);
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float u = has_up ? up[e] : 0.0f, dd = has_dn ? dn[e] : 0.0f;
                                // taps in position order, pad 1 (zero rows outside the candidate); BN scale folded on the host
                                const float qv = fmaf(dd, t2[e], fmaf(v[e], t1[e], u * t0[e]));
                                const float kvv = fmaf(dd, k2[e], fmaf(v[e], k1[e], u * k0[e]));
                                dq[e] = lane_ok ? qv : 0.0f;
                                dk[e] = kv_row ? kvv : 0.0f;           // stride-2 conv = stride-1 conv at even positions
                            }
                        } else {
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) { dq[e] = 0.0f; dk[e] = 0.0f; }
                        }
                        const int kb = c8 >> 6, ch = (c8 & 63) >> 3;
                        f_store8(pa + kb * TILE_A, pa_mid + kb * TILE_A, row, ch, dq);
                        f_store8(pb + kb * TILE_A, pb_mid + kb * TILE_A, row, ch, dk);
                    }
                }
                f_fence_async();
                f_fence_before();
                __syncwarp();
                if (lane == 0) g_mbar_arrive(ab_ready);
                GTOC(1);
                // ---------- attention: the heads whose accumulator buffer belongs to this group ----------
                for (int h = 0; h < heads && alive; ++h) {
                    const uint32_t hit = lit * (uint32_t)heads + h, b = hit & 1;
                    if ((int)b != grp) continue;
                    if (!f_wait(&qkv_full[b], (hit >> 1) & 1, B_QKVFULL, dbg)) { alive = false; break; }
                    GTOC(2);
                    f_fence_after();
                    const uint32_t acc = trow + (uint32_t)(CP + b * 192);
                    const float* bq = sv + K::O_BQ + h * 64;
                    const float* bk = sv + K::O_BQ + heads * 64 + h * 64;
                    const float* bv = sv + K::O_BQ + 2 * heads * 64 + h * 64;
                    // scores against the keys of the own candidate (the 64^-0.5 scale is folded into q on the host)
                    float s[WKV];
                    #pragma unroll
                    for (int j = 0; j < WKV; ++j) s[j] = 0.0f;
                    #pragma unroll 1
                    for (int d0 = 0; d0 < 64; d0 += 8) {
                        float q[8], k[8], b0[8], b1[8];
                        f_tmem_ld8f(acc + (uint32_t)d0, q);
                        f_tmem_ld8f(acc + (uint32_t)(64 + d0), k);
                        f_lds8(bq + d0, b0);
                        f_lds8(bk + d0, b1);
                        f_wait_ld();
                        #pragma unroll
                        for (int e = 0; e < 8; ++e) { q[e] += b0[e]; k[e] += b1[e]; }
                        #pragma unroll
                        for (int j = 0; j < WKV; ++j) {
                            float t[8];
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) t[e] = __shfl_sync(0xffffffffu, k[e], src0 + 2 * j);   // batch of shuffles first
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) s[j] = fmaf(q[e], t[e], s[j]);
                        }
                    }
                    float mx = s[0];
                    #pragma unroll
                    for (int j = 1; j < WKV; ++j) mx = fmaxf(mx, s[j]);
                    float sum = 0.0f;
                    #pragma unroll
                    for (int j = 0; j < WKV; ++j) {
                        s[j] = expf(s[j] - mx);
                        sum += s[j];
                    }
                    const float inv = 1.0f / sum;
                    #pragma unroll
                    for (int j = 0; j < WKV; ++j) s[j] *= inv;
                    GTOC(3);
                    // the plane buffer of this head: free once the out-projection of head h - NPC has retired
                    const int pbuf = h % NPC;
                    if (h >= NPC) {
                        if (!f_wait(&pc_free[grp * 2 + pbuf], pcf_cnt[pbuf] & 1, B_PCFREE, dbg)) { alive = false; break; }
                        ++pcf_cnt[pbuf];
                    }
                    GTOC(4);
                    uint8_t* pch = pc + pbuf * K::PCBUF;
                    #pragma unroll 1
                    for (int d0 = 0; d0 < 64; d0 += 8) {
                        float v[8], b2[8], o[8];
                        f_tmem_ld8f(acc + (uint32_t)(128 + d0), v);
                        f_lds8(bv + d0, b2);
                        f_wait_ld();
                        #pragma unroll
                        for (int e = 0; e < 8; ++e) { v[e] += b2[e]; o[e] = 0.0f; }
                        #pragma unroll
                        for (int j = 0; j < WKV; ++j) {
                            float t[8];
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) t[e] = __shfl_sync(0xffffffffu, v[e], src0 + 2 * j);
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) o[e] = fmaf(s[j], t[e], o[e]);
                        }
                        f_store8(pch, pch + TILE_A, row, d0 >> 3, o);
                    }
                    f_fence_async();
                    f_fence_before();
                    __syncwarp();
                    if (lane == 0) { g_mbar_arrive(&qkv_free[b]); g_mbar_arrive(&pc_ready[pbuf]); }   // TMEM reads done, planes written
                    GTOC(3);
                }
                if (!alive) break;
                // ---------- LN2 -> PA ----------
                if (!f_wait(x_ready, lit & 1, B_XREADY, dbg)) { alive = false; break; }
                GTOC(5);
                f_fence_after();
                {
                    float mean, inv;
                    f_row_stats<C>(trow, sv + K::O_CB2, mean, inv);
                    #pragma unroll 1
                    for (int c8 = grp * 8; c8 < CP; c8 += 16) {
                        float v[8];
                        if (c8 < C) {
                            f_tmem_ld8f(trow + (uint32_t)c8, v);
                            f_wait_ld();
                            f_normalise8(v, sv + K::O_CB2 + c8, sv + K::O_LN2G + c8, sv + K::O_LN2B + c8, mean, inv);
                        }
                        #pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = (c8 < C && lane_ok) ? v[e] : 0.0f;
                        const int kb = c8 >> 6, ch = (c8 & 63) >> 3;
                        f_store8(pa + kb * TILE_A, pa_mid + kb * TILE_A, row, ch, v);
                    }
                }
                f_fence_async();
                f_fence_before();
                __syncwarp();
                if (lane == 0) g_mbar_arrive(y2_ready);
                GTOC(6);
                // ---------- feed-forward hidden chunks: even ones in group 0, odd ones in group 1 ----------
                for (int j = grp; j < NH && alive; j += 2) {
                    const uint32_t cit = lit * (uint32_t)NH + j, hb = cit % NB;
                    const int fb = j % K::NFF;
                    if (j >= K::NFF) {                                  // FF2 of chunk j - NFF has retired
                        if (!f_wait(&ff_free[grp * 3 + fb], fff_cnt[fb] & 1, B_PCFREE, dbg)) { alive = false; break; }
                        ++fff_cnt[fb];
                    }
                    GTOC(4);
                    if (!f_wait(&hid_full[hb], (cit / NB) & 1, B_HIDFULL, dbg)) { alive = false; break; }
                    GTOC(7);
                    f_fence_after();
                    uint8_t* fbuf = fb < NPC ? pc + fb * K::PCBUF : pb + (fb - NPC) * K::PCBUF;
                    const float* b1 = sv + K::O_B1 + j * 64;
                    #pragma unroll 1
                    for (int d8 = 0; d8 < 64; d8 += 8) {
                        float hdn[8], bb[8];
                        f_tmem_ld8f(trow + (uint32_t)(CP + hb * 64 + d8), hdn);
                        f_lds8(b1 + d8, bb);
                        f_wait_ld();
                        f_gelu8(hdn, bb);
                        f_store8(fbuf, fbuf + TILE_A, row, d8 >> 3, hdn);
                    }
                    f_fence_async();
                    f_fence_before();
                    __syncwarp();
                    if (lane == 0) { g_mbar_arrive(&hid_free[hb]); g_mbar_arrive(&ff_ready[fb]); }
                    GTOC(8);
                }
                if (!alive) break;
                if (!f_wait(layer_done, lit & 1, B_LAYER, dbg)) { alive = false; break; }
                GTOC(9);
                f_fence_after();
            }
            if (!alive) break;
            if (grp == 0) {
                // x tile: TMEM + cumulative bias -> global
                #pragma unroll 1
                for (int c0 = 0; c0 < C; c0 += 32) {
                    float v[32];
                    f_tmem_ld32(trow + (uint32_t)c0, v);
                    f_wait_ld();
                    if (ok) {
                        #pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(p.cb_final + c0 + 4 * q));
                            *reinterpret_cast<float4*>(xg + c0 + 4 * q) =
                                make_float4(v[4 * q] + t.x, v[4 * q + 1] + t.y, v[4 * q + 2] + t.z, v[4 * q + 3] + t.w);
                        }
                    }
                }
                f_fence_before();
                GTOC(10);
            }
        }
        if (tg && warp == 2 && lane == 0) { for (int i = 0; i < 11; ++i) p.timing[8 + i] = tacc[i]; }
        #undef GTOC
    }
    #undef FTOC
    f_fence_before();
    __syncthreads();
    if (warp == 1) {
        f_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- host side: weight stream prepack --------------------------------------------------------------------
static inline uint16_t h_bf16(float f) {          // round to nearest even, like cvt.rn.bf16.f32
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static inline float h_bf16_f(uint16_t b) {
    const uint32_t u = (uint32_t)b << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// one [64 n x 64 k] tile of W[n_total, k_total] (row-major, fp32) at (n0, k0) -> hi image | mid image, zero padded
static void emit_tile(std::vector<uint8_t>& out, const float* w, int n_total, int k_total, int n0, int k0) {
    const size_t at = out.size();
    out.resize(at + WSTAGE, 0);
    uint8_t* hi = out.data() + at;
    uint8_t* mid = hi + WTILE;
    for (int r = 0; r < 64; ++r) {
        if (n0 + r >= n_total) break;
        for (int k = 0; k < 64; ++k) {
            if (k0 + k >= k_total) break;
            const float v = w[(size_t)(n0 + r) * k_total + k0 + k];
            const uint16_t h = h_bf16(v), m = h_bf16(v - h_bf16_f(h));
            const size_t off = (size_t)(r >> 3) * 1024 + (r & 7) * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
            std::memcpy(hi + off, &h, 2);
            std::memcpy(mid + off, &m, 2);
        }
    }
}

}  // namespace fz

long long* g_fused_timing = nullptr;     // cto_debug_timing_fused(): device buffer [32], debug only

bool aff_layers_fused_supported(const CvtStage& st) {
    if (!(st.c == 32 || st.c == 64 || st.c == 128)) return false;
    if (st.depth < 1 || st.heads < 1 || !(st.wout == 17 || st.wout == 9 || st.wout == 5) || st.wkv != (st.wout + 1) / 2) return false;
    return true;
}

// Builds the weight stream + per-layer vector table of one stage.  host_blob / dev_blob: the stage's weights live at
// the same offsets in both (CvtLayer pointers point into dev_blob).
int aff_fused_prepare(CvtStage& st, const float* host_blob, const float* dev_blob) {
    using namespace fz;
    const int c = st.c, cp = c < 64 ? 64 : c, kb = cp / 64, nt = cp / 64, nh = 4 * c / 64, inner = st.heads * DIM_HEAD;
    std::vector<uint8_t> stream;
    const int vfloats = 16 * c + 3 * inner;
    CTO_REQUIRE(vfloats * 4 <= 16 * 1024, "aff_fused: %d heads at width %d do not fit the vector block", st.heads, c);
    std::vector<float> vblocks((size_t)st.depth * 4096, 0.0f), run(c, 0.0f);
    auto host = [&](const float* dev_ptr) { return host_blob + (dev_ptr - dev_blob); };
    for (int d = 0; d < st.depth; ++d) {
        const CvtLayer& L = st.layers[d];
        const size_t before = stream.size();
        const float *wq = host(L.q_pw), *wkv = host(L.kv_pw), *wo = host(L.out_w), *w1 = host(L.ff1_w), *w2 = host(L.ff2_w);
        auto out_tiles = [&](int h) { for (int t = 0; t < nt; ++t) emit_tile(stream, wo, c, inner, t * 64, h * 64); };
        auto ff2_tiles = [&](int j) { for (int t = 0; t < nt; ++t) emit_tile(stream, w2, c, 4 * c, t * 64, j * 64); };
        // consumption order of the MMA warp; tiles that feed one N = 128 MMA (k_h | v_h, the two halves of a 128-wide
        // output, hidden chunks j | j + 1) are adjacent
        for (int h = 0; h < st.heads; ++h) {
            for (int k = 0; k < kb; ++k) emit_tile(stream, wq, inner, c, h * 64, k * 64);
            for (int k = 0; k < kb; ++k) {
                emit_tile(stream, wkv, 2 * inner, c, h * 64, k * 64);
                emit_tile(stream, wkv, 2 * inner, c, inner + h * 64, k * 64);
            }
            if (h >= 1) out_tiles(h - 1);
        }
        out_tiles(st.heads - 1);
        for (int j = 0; j < nh; j += 2) {
            for (int k = 0; k < kb; ++k) {
                emit_tile(stream, w1, 4 * c, c, j * 64, k * 64);
                emit_tile(stream, w1, 4 * c, c, (j + 1) * 64, k * 64);
            }
            if (j >= 2) { ff2_tiles(j - 2); ff2_tiles(j - 1); }
        }
        ff2_tiles(nh - 2);
        ff2_tiles(nh - 1);
        const size_t bytes = stream.size() - before;
        if (d == 0) st.fused_layer_bytes = (long long)bytes;
        CTO_REQUIRE((long long)bytes == st.fused_layer_bytes &&
                        bytes == (size_t)(st.heads * (3 * kb + nt) + nh * (kb + nt)) * WSTAGE,
                    "aff_fused: weight stream of layer %d has %zu bytes", d, bytes);
        // the layer's vector block (fz::Cfg<C>::O_*); cb1 / cb2 = the out-projection and FF2 biases accumulated so far
        float* vb = vblocks.data() + (size_t)d * 4096;
        auto put = [&](int off, const float* dev_ptr, int n) { std::memcpy(vb + off, host(dev_ptr), sizeof(float) * n); };
        put(0, L.ln1_g, c); put(c, L.ln1_b, c); put(2 * c, L.q_dw, 3 * c); put(5 * c, L.kv_dw, 3 * c);
        put(8 * c, L.ln2_g, c); put(9 * c, L.ln2_b, c); put(12 * c, L.ff1_b, 4 * c);
        put(16 * c, L.q_bias, inner); put(16 * c + inner, L.kv_bias, 2 * inner);
        const float *bo = host(L.out_b), *b2 = host(L.ff2_b);
        for (int i = 0; i < c; ++i) {
            vb[10 * c + i] = run[i];
            run[i] += bo[i];
            vb[11 * c + i] = run[i];
            run[i] += b2[i];
        }
    }
    st.fused_vec_bytes = (vfloats * 4 + 15) / 16 * 16;
    CTO_CHECK(cudaMalloc(&st.fused_vecs, sizeof(float) * vblocks.size()));
    CTO_CHECK(cudaMemcpy(st.fused_vecs, vblocks.data(), sizeof(float) * vblocks.size(), cudaMemcpyHostToDevice));
    CTO_CHECK(cudaMalloc(&st.fused_cb, sizeof(float) * c));
    CTO_CHECK(cudaMemcpy(st.fused_cb, run.data(), sizeof(float) * c, cudaMemcpyHostToDevice));
    CTO_CHECK(cudaMalloc(&st.fused_stream, stream.size()));
    CTO_CHECK(cudaMemcpy(st.fused_stream, stream.data(), stream.size(), cudaMemcpyHostToDevice));
    return 0;
}

void aff_fused_release(CvtStage& st) {
    if (st.fused_stream) cudaFree(st.fused_stream);
    if (st.fused_vecs) cudaFree(st.fused_vecs);
    if (st.fused_cb) cudaFree(st.fused_cb);
    st.fused_stream = nullptr;
    st.fused_vecs = nullptr;
    st.fused_cb = nullptr;
}

template <int C, int W>
static int launch_layers_t(const fz::Params& p, int grid, cudaStream_t s) {
    CTO_CHECK(set_max_dynamic_smem(fz::aff_layers_kernel<C, W>, fz::Cfg<C>::SMEM));
    fz::aff_layers_kernel<C, W><<<grid, fz::THREADS, fz::Cfg<C>::SMEM, s>>>(p);
    return 0;
}
template <int C>
static int launch_layers_w(const fz::Params& p, int grid, cudaStream_t s) {
    // the stage widths of a 33-position input: 17 -> 9 -> 5 (3-tap, stride 2, pad 1)
    if (p.W == 17) return launch_layers_t<C, 17>(p, grid, s);
    if (p.W == 9) return launch_layers_t<C, 9>(p, grid, s);
    return launch_layers_t<C, 5>(p, grid, s);
}

// x: fp32 [n, W, C], the stage's residual stream after the embed convolution + LN; all `depth` layers in place.
// dbg: device int[8], zeroed by the caller once; non-zero dbg[0] after the launch = a barrier timed out (see fz::B_*)
int launch_aff_layers(const CvtStage& st, float* x, int64_t n, int* dbg, cudaStream_t s) {
    if (n <= 0) return 0;
    CTO_REQUIRE(aff_layers_fused_supported(st) && st.fused_stream && st.fused_vecs, "aff_layers: stage C=%d heads=%d depth=%d W=%d not prepared",
                st.c, st.heads, st.depth, st.wout);
    CTO_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "aff_layers: x must be 16-byte aligned");
    fz::Params p;
    p.x = x;
    p.wstream = st.fused_stream;
    p.vblocks = st.fused_vecs;
    p.vec_bytes = st.fused_vec_bytes;
    p.cb_final = st.fused_cb;
    p.layer_bytes = st.fused_layer_bytes;
    p.n = n;
    p.W = st.wout;
    p.WKV = st.wkv;
    p.heads = st.heads;
    p.depth = st.depth;
    p.cpw = 32 / st.wout;
    const int64_t per_tile = 4 * p.cpw;
    const int64_t tiles = (n + per_tile - 1) / per_tile;
    CTO_REQUIRE(tiles < (1ll << 31), "aff_layers: too many tiles");
    p.tiles = (int)tiles;
    p.dbg = dbg;
    p.timing = g_fused_timing;
    const int sms = device_sm_count();
    CTO_REQUIRE(sms > 0, "aff_layers: no device");
    const int grid = (int)(tiles < sms ? tiles : sms);
    int rc;
    if (st.c == 32) rc = launch_layers_w<32>(p, grid, s);
    else if (st.c == 64) rc = launch_layers_w<64>(p, grid, s);
    else rc = launch_layers_w<128>(p, grid, s);
    if (rc) return rc;
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
