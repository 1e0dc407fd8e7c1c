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
constexpr int LAG = 2;                        // FF1 MMAs run this many chunks ahead of FF2
constexpr uint32_t SPIN_LIMIT = 1u << 22;

template <int C>
struct Cfg {
    static constexpr int CP = C < 64 ? 64 : C;                // channels padded to the 64-wide k-block
    static constexpr int KB = CP / 64;                        // k-blocks of a C-wide contraction
    static constexpr int NT = CP / 64;                        // 64-column output tiles of a C-wide product
    static constexpr int NH = (4 * C) / 64;                   // feed-forward hidden chunks
    static constexpr int NPC = C >= 128 ? 1 : 2;              // attention-output / hidden plane buffers
    static constexpr int WST = C >= 128 ? 4 : 6;              // weight ring stages
    static constexpr int PLANE = 2 * KB * TILE_A;             // hi k-blocks | mid k-blocks
    static constexpr int PCBUF = 2 * TILE_A;                  // hi | mid, one k-block
    static constexpr int SMEM = 2 * PLANE + NPC * PCBUF + WST * WSTAGE + 1024 /*align*/ + 1024 /*barriers*/;
    static_assert(C == 32 || C == 64 || C == 128, "stage width");
    static_assert((4 * C) % 64 == 0 && NH >= LAG, "hidden chunking");
};

struct Params {
    float* x;                     // [n * W, C] fp32, in place
    const uint8_t* wstream;       // prepacked weight tiles, layer l at l * layer_bytes
    const FusedLayerVecs* vecs;   // device array [depth]
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
// D=f32, A=B=bf16, both K-major, M=128, N=64
__device__ __forceinline__ uint32_t f_idesc64() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// barrier codes reported in dbg[0] on a timeout
enum { B_WFULL = 1, B_WEMPTY, B_AB, B_QKVFULL, B_QKVFREE, B_PCREADY, B_PCFREE, B_XREADY, B_Y2, B_HIDFULL, B_HIDFREE, B_LAYER };

// channel LayerNorm (clairs/model.py:57-67): population std over the C channels of a row, eps added to the std.
// The row lives in tensor memory (x_true = x_tmem + cb); statistics with the shifted-data formulas (shift = first channel).
template <int C>
__device__ __forceinline__ void f_row_stats(uint32_t trow, const float* __restrict__ cb, float& mean, float& inv) {
    float s1 = 0.0f, s2 = 0.0f, shift = 0.0f;
    #pragma unroll
    for (int c0 = 0; c0 < C; c0 += 32) {
        float v[32];
        f_tmem_ld32(trow + (uint32_t)c0, v);
        f_wait_ld();
        #pragma unroll
        for (int c4 = 0; c4 < 32; c4 += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(cb + c0 + c4));
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
// 32 channels of the row: (x_tmem + cb - mean) * inv * g + b
__device__ __forceinline__ void f_normalise32(float* v, const float* __restrict__ cb, const float* __restrict__ g,
                                              const float* __restrict__ b, float mean, float inv) {
    #pragma unroll
    for (int c4 = 0; c4 < 32; c4 += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(cb + c4));
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + c4)), bv = __ldg(reinterpret_cast<const float4*>(b + c4));
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
    uint8_t* wring = pc + NPC * K::PCBUF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(wring + WST * WSTAGE);
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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(layer_done + 1);

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
        g_mbar_init(ab_ready, 4);
        for (int b = 0; b < 2; ++b) {
            g_mbar_init(&qkv_full[b], 1); g_mbar_init(&qkv_free[b], 4);
            g_mbar_init(&pc_ready[b], 4);
        }
        for (int b = 0; b < 4; ++b) g_mbar_init(&pc_free[b], 1);
        g_mbar_init(x_ready, 1);
        g_mbar_init(y2_ready, 4);
        for (int b = 0; b < NB; ++b) { g_mbar_init(&hid_full[b], 1); g_mbar_init(&hid_free[b], 4); }
        g_mbar_init(layer_done, 1);
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
        uint32_t wit = 0;
        bool alive = true;
        for (int tile = blockIdx.x; tile < p.tiles && alive; tile += gridDim.x) {
            for (int l = 0; l < depth && alive; ++l) {
                const uint8_t* src = p.wstream + (long long)l * p.layer_bytes;
                for (int i = 0; i < tiles_per_layer; ++i, ++wit) {
                    const int s = wit % WST;
                    if (!f_wait(&w_empty[s], ((wit / WST) & 1) ^ 1, B_WEMPTY, dbg)) { alive = false; break; }
                    if (g_elect_one()) {
                        g_mbar_expect_tx(&w_full[s], WSTAGE);
                        f_bulk_load(wring + s * WSTAGE, src + (long long)i * WSTAGE, WSTAGE, &w_full[s]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (warp-uniform control flow, one elected lane issues) ----
        const uint32_t idesc = f_idesc64();
        uint32_t wit = 0, lit = 0;
        bool alive = true;
        // one [128 x 64*nkb] x [64 x 64*nkb]^T product: nkb weight tiles from the ring, A k-blocks at a_hi / a_mid
        auto product = [&](uint32_t d_col, const uint8_t* a_hi, const uint8_t* a_mid, int nkb, bool acc_first) -> bool {
            for (int kb = 0; kb < nkb; ++kb, ++wit) {
                const int s = wit % WST;
                FTOC(0);
                if (!f_wait(&w_full[s], (wit / WST) & 1, B_WFULL, dbg)) return false;
                FTOC(1);
                f_fence_after();
                const uint64_t d_ahi = g_desc_k_sw128(g_smem_u32(a_hi + kb * TILE_A));
                const uint64_t d_amid = g_desc_k_sw128(g_smem_u32(a_mid + kb * TILE_A));
                const uint32_t w_addr = g_smem_u32(wring + s * WSTAGE);
                const uint64_t d_whi = g_desc_k_sw128(w_addr), d_wmid = g_desc_k_sw128(w_addr + WTILE);
                if (g_elect_one()) {
                    #pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t o = (uint64_t)(k * 2);          // 16 bf16 = 32 bytes along the swizzle row
                        f_mma(d_col, d_ahi + o, d_whi + o, idesc, (acc_first || kb || k) ? 1u : 0u);
                        f_mma(d_col, d_amid + o, d_whi + o, idesc, 1u);
                        f_mma(d_col, d_ahi + o, d_wmid + o, idesc, 1u);
                    }
                    g_commit(&w_empty[s]);
                }
                __syncwarp();
            }
            return true;
        };
        // x[:, nt*64 .. +64) += PC[p] * W_tile(nt)^T  for the plane buffer of use `pit`
        auto onto_x = [&](uint32_t pit) -> bool {
            const int pbuf = pit % NPC;
            FTOC(0);
            if (!f_wait(&pc_ready[pbuf], (pit / NPC) & 1, B_PCREADY, dbg)) return false;
            FTOC(2);
            f_fence_after();
            const uint8_t* a = pc + pbuf * K::PCBUF;
            for (int nt = 0; nt < NT; ++nt)
                if (!product(tmem_base + (uint32_t)(nt * 64), a, a + TILE_A, 1, true)) return false;
            // the buffer goes back to the group that writes use pit + NPC: group A (0) for attention heads and even
            // hidden chunks, group B (1) for odd hidden chunks.  One barrier per (writer group, buffer): a barrier that
            // two groups wait on in turns would let the group that skipped a phase alias an old parity
            const uint32_t r = (pit + NPC) % (uint32_t)(heads + NH);
            const int next_grp = (r >= (uint32_t)heads && ((r - heads) & 1u)) ? 1 : 0;
            if (g_elect_one()) g_commit(&pc_free[next_grp * 2 + pbuf]);
            __syncwarp();
            return true;
        };
        for (int tile = blockIdx.x; tile < p.tiles && alive; tile += gridDim.x) {
            for (int l = 0; l < depth && alive; ++l, ++lit) {
                const uint32_t pit0 = lit * (uint32_t)(heads + NH);
                FTOC(0);
                if (!f_wait(ab_ready, lit & 1, B_AB, dbg)) { alive = false; break; }
                FTOC(3);
                f_fence_after();
                for (int h = 0; h < heads && alive; ++h) {
                    const uint32_t hit = lit * (uint32_t)heads + h, b = hit & 1;
                    FTOC(0);
                    if (!f_wait(&qkv_free[b], ((hit >> 1) & 1) ^ 1, B_QKVFREE, dbg)) { alive = false; break; }
                    FTOC(4);
                    f_fence_after();
                    const uint32_t acc = tmem_base + (uint32_t)(CP + b * 192);
                    alive = product(acc, pa, pa + KB * TILE_A, KB, false) &&
                            product(acc + 64, pb, pb + KB * TILE_A, KB, false) &&
                            product(acc + 128, pb, pb + KB * TILE_A, KB, false);
                    if (!alive) break;
                    if (g_elect_one()) g_commit(&qkv_full[b]);
                    __syncwarp();
                    if (h >= 1) alive = onto_x(pit0 + h - 1);
                }
                if (!alive) break;
                if (!onto_x(pit0 + heads - 1)) { alive = false; break; }
                if (g_elect_one()) g_commit(x_ready);
                __syncwarp();
                FTOC(0);
                if (!f_wait(y2_ready, lit & 1, B_Y2, dbg)) { alive = false; break; }
                FTOC(5);
                f_fence_after();
                for (int j = 0; j < NH && alive; ++j) {
                    const uint32_t cit = lit * (uint32_t)NH + j, hb = cit % NB;
                    FTOC(0);
                    if (!f_wait(&hid_free[hb], ((cit / NB) & 1) ^ 1, B_HIDFREE, dbg)) { alive = false; break; }
                    FTOC(6);
                    f_fence_after();
                    if (!product(tmem_base + (uint32_t)(CP + hb * 64), pa, pa + KB * TILE_A, KB, false)) { alive = false; break; }
                    if (g_elect_one()) g_commit(&hid_full[hb]);
                    __syncwarp();
                    if (j >= LAG) alive = onto_x(pit0 + heads + j - LAG);
                }
                for (int j = NH - LAG; j < NH && alive; ++j) alive = onto_x(pit0 + heads + j);
                if (!alive) break;
                if (g_elect_one()) g_commit(layer_done);
                __syncwarp();
            }
        }
        FTOC(0);
        if (tim && lane == 0) { for (int i = 0; i < 7; ++i) p.timing[i] = tacc[i]; p.timing[7] = (long long)lit; }
    } else if (warp < 6) {
        // ---- compute group A: one thread per row (TMEM lane = row) ----
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        constexpr int cpw = CPW;
        const int cl = lane / W, w = lane - cl * W;                    // candidate inside the warp, position
        const bool lane_ok = lane < cpw * W;
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
        const bool has_up = lane_ok && w > 0, has_dn = lane_ok && w < W - 1, kv_row = lane_ok && !(w & 1);
        uint8_t* pa_mid = pa + KB * TILE_A;
        uint8_t* pb_mid = pb + KB * TILE_A;
        uint32_t lit = 0;
        uint32_t fcnt[2] = {0, 0};                                     // signalled waits so far on pc_free[group A][buffer]
        bool alive = true;
        for (int tile = blockIdx.x; tile < p.tiles && alive; tile += gridDim.x) {
            const long long cand = ((long long)tile * 4 + quad) * cpw + cl;
            const bool ok = lane_ok && cand < p.n;
            float* xg = p.x + (((long long)tile * 4 + quad) * cpw * W + lane) * (long long)C;
            // x tile: global -> registers -> TMEM columns [0, CP)
            #pragma unroll
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
            FTOC(0);
            for (int l = 0; l < depth && alive; ++l, ++lit) {
                const FusedLayerVecs lv = p.vecs[l];
                const uint32_t pit0 = lit * (uint32_t)(heads + NH);
                // ---------- LN1 + depth-wise convolutions -> dq (PA), dkv (PB) ----------
                // The row is read from tensor memory twice, 32 columns at a time: statistics, then normalise + convolve +
                // store (the whole 128-channel row in registers left no room for batched shuffles).
                {
                    float mean, inv;
                    f_row_stats<C>(trow, lv.cb1, mean, inv);
                    #pragma unroll
                    for (int c0 = 0; c0 < CP; c0 += 32) {
                        float v[32];
                        if (c0 < C) {
                            f_tmem_ld32(trow + (uint32_t)c0, v);
                            f_wait_ld();
                            f_normalise32(v, lv.cb1 + c0, lv.ln1_g + c0, lv.ln1_b + c0, mean, inv);
                        }
                        #pragma unroll
                        for (int c8 = 0; c8 < 32; c8 += 8) {
                            float dq[8], dk[8];
                            if (c0 < C) {
                                // neighbour rows by warp shuffles, issued as a batch (a shuffle feeding its own FMA serialises
                                // on the shuffle latency: measured 10 cycles per instruction)
                                float up[8], dn[8];
                                #pragma unroll
                                for (int e = 0; e < 8; ++e) up[e] = __shfl_up_sync(0xffffffffu, v[c8 + e], 1);
                                #pragma unroll
                                for (int e = 0; e < 8; ++e) dn[e] = __shfl_down_sync(0xffffffffu, v[c8 + e], 1);
                                #pragma unroll
                                for (int h4 = 0; h4 < 8; h4 += 4) {
                                    const int c = c0 + c8 + h4;
                                    const float4 a0 = __ldg(reinterpret_cast<const float4*>(lv.tq + c));
                                    const float4 a1 = __ldg(reinterpret_cast<const float4*>(lv.tq + C + c));
                                    const float4 a2 = __ldg(reinterpret_cast<const float4*>(lv.tq + 2 * C + c));
                                    const float4 b0 = __ldg(reinterpret_cast<const float4*>(lv.tk + c));
                                    const float4 b1 = __ldg(reinterpret_cast<const float4*>(lv.tk + C + c));
                                    const float4 b2 = __ldg(reinterpret_cast<const float4*>(lv.tk + 2 * C + c));
                                    const float t0[4] = {a0.x, a0.y, a0.z, a0.w}, t1[4] = {a1.x, a1.y, a1.z, a1.w}, t2[4] = {a2.x, a2.y, a2.z, a2.w};
                                    const float k0[4] = {b0.x, b0.y, b0.z, b0.w}, k1[4] = {b1.x, b1.y, b1.z, b1.w}, k2[4] = {b2.x, b2.y, b2.z, b2.w};
                                    #pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float y = v[c8 + h4 + e];
                                        const float u = has_up ? up[h4 + e] : 0.0f, dd = has_dn ? dn[h4 + e] : 0.0f;
                                        // taps in position order, pad 1 (zero rows outside the candidate); BN scale folded on the host
                                        const float qv = fmaf(dd, t2[e], fmaf(y, t1[e], u * t0[e]));
                                        const float kvv = fmaf(dd, k2[e], fmaf(y, k1[e], u * k0[e]));
                                        dq[h4 + e] = lane_ok ? qv : 0.0f;
                                        dk[h4 + e] = kv_row ? kvv : 0.0f;   // stride-2 conv = stride-1 conv at even positions
                                    }
                                }
                            } else {
                                #pragma unroll
                                for (int e = 0; e < 8; ++e) { dq[e] = 0.0f; dk[e] = 0.0f; }
                            }
                            const int cc = c0 + c8, kb = cc >> 6, ch = (cc & 63) >> 3;
                            f_store8(pa + kb * TILE_A, pa_mid + kb * TILE_A, row, ch, dq);
                            f_store8(pb + kb * TILE_A, pb_mid + kb * TILE_A, row, ch, dk);
                        }
                    }
                }
                f_fence_async();
                f_fence_before();
                __syncwarp();
                if (lane == 0) g_mbar_arrive(ab_ready);
                FTOC(1);
                // ---------- attention, head by head ----------
                for (int h = 0; h < heads && alive; ++h) {
                    const uint32_t hit = lit * (uint32_t)heads + h, b = hit & 1;
                    if (!f_wait(&qkv_full[b], (hit >> 1) & 1, B_QKVFULL, dbg)) { alive = false; break; }
                    FTOC(2);
                    f_fence_after();
                    const uint32_t acc = trow + (uint32_t)(CP + b * 192);
                    float q[64], kv[64];
                    f_tmem_ld32(acc, q);
                    f_tmem_ld32(acc + 32, q + 32);
                    f_tmem_ld32(acc + 64, kv);
                    f_tmem_ld32(acc + 96, kv + 32);
                    f_wait_ld();
                    const float* bq = lv.bq + h * 64;
                    const float* bk = lv.bkv + h * 64;
                    const float* bv = lv.bkv + heads * 64 + h * 64;
                    #pragma unroll
                    for (int d4 = 0; d4 < 64; d4 += 4) {
                        const float4 a = __ldg(reinterpret_cast<const float4*>(bq + d4)), c = __ldg(reinterpret_cast<const float4*>(bk + d4));
                        q[d4] += a.x; q[d4 + 1] += a.y; q[d4 + 2] += a.z; q[d4 + 3] += a.w;
                        kv[d4] += c.x; kv[d4 + 1] += c.y; kv[d4 + 2] += c.z; kv[d4 + 3] += c.w;
                    }
                    // scores against the keys of the own candidate: key j lives in the lane of row (cl, 2j)
                    float s[9];
                    const int src0 = cl * W;
                    #pragma unroll
                    for (int j = 0; j < 9; ++j) s[j] = 0.0f;
                    #pragma unroll
                    for (int d0 = 0; d0 < 64; d0 += 8) {
                        #pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            if (j < WKV) {
                                const int src = src0 + 2 * j;
                                float t[8];
                                #pragma unroll
                                for (int e = 0; e < 8; ++e) t[e] = __shfl_sync(0xffffffffu, kv[d0 + e], src);   // batch of shuffles first
                                #pragma unroll
                                for (int e = 0; e < 8; ++e) s[j] = fmaf(q[d0 + e], t[e], s[j]);
                            }
                        }
                    }
                    float mx = s[0];
                    #pragma unroll
                    for (int j = 1; j < 9; ++j) if (j < WKV) mx = fmaxf(mx, s[j]);
                    float sum = 0.0f;
                    #pragma unroll
                    for (int j = 0; j < 9; ++j) {
                        s[j] = j < WKV ? expf(s[j] - mx) : 0.0f;
                        sum += s[j];
                    }
                    const float inv = 1.0f / sum;
                    // values: reuse the key registers
                    f_tmem_ld32(acc + 128, kv);
                    f_tmem_ld32(acc + 160, kv + 32);
                    f_wait_ld();
                    f_fence_before();
                    #pragma unroll
                    for (int d4 = 0; d4 < 64; d4 += 4) {
                        const float4 c = __ldg(reinterpret_cast<const float4*>(bv + d4));
                        kv[d4] += c.x; kv[d4 + 1] += c.y; kv[d4 + 2] += c.z; kv[d4 + 3] += c.w;
                    }
                    #pragma unroll
                    for (int j = 0; j < 9; ++j) s[j] *= inv;
                    #pragma unroll
                    for (int d0 = 0; d0 < 64; d0 += 8) {
                        float o[8];
                        #pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = 0.0f;
                        #pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            if (j < WKV) {
                                const int src = src0 + 2 * j;
                                float t[8];
                                #pragma unroll
                                for (int e = 0; e < 8; ++e) t[e] = __shfl_sync(0xffffffffu, kv[d0 + e], src);
                                #pragma unroll
                                for (int e = 0; e < 8; ++e) o[e] = fmaf(s[j], t[e], o[e]);
                            }
                        }
                        #pragma unroll
                        for (int e = 0; e < 8; ++e) q[d0 + e] = o[e];
                    }
                    __syncwarp();
                    if (lane == 0) g_mbar_arrive(&qkv_free[b]);                // every TMEM read of this buffer has completed
                    const uint32_t pit = pit0 + h;
                    const int pbuf = pit % NPC;
                    FTOC(3);
                    if (pit >= (uint32_t)NPC) {                                // the MMAs that read the buffer's previous use have retired
                        if (!f_wait(&pc_free[pbuf], fcnt[pbuf] & 1, B_PCFREE, dbg)) { alive = false; break; }
                        ++fcnt[pbuf];
                    }
                    FTOC(4);
                    uint8_t* pch = pc + pbuf * K::PCBUF;
                    #pragma unroll
                    for (int ch = 0; ch < 8; ++ch) f_store8(pch, pch + TILE_A, row, ch, q + 8 * ch);
                    f_fence_async();
                    __syncwarp();
                    if (lane == 0) g_mbar_arrive(&pc_ready[pbuf]);
                }
                if (!alive) break;
                // ---------- LN2 -> PA ----------
                FTOC(3);
                if (!f_wait(x_ready, lit & 1, B_XREADY, dbg)) { alive = false; break; }
                FTOC(5);
                f_fence_after();
                {
                    float mean, inv;
                    f_row_stats<C>(trow, lv.cb2, mean, inv);
                    #pragma unroll
                    for (int c0 = 0; c0 < CP; c0 += 32) {
                        float v[32];
                        if (c0 < C) {
                            f_tmem_ld32(trow + (uint32_t)c0, v);
                            f_wait_ld();
                            f_normalise32(v, lv.cb2 + c0, lv.ln2_g + c0, lv.ln2_b + c0, mean, inv);
                        }
                        #pragma unroll
                        for (int c8 = 0; c8 < 32; c8 += 8) {
                            float z[8];
                            #pragma unroll
                            for (int e = 0; e < 8; ++e) z[e] = (c0 < C && lane_ok) ? v[c8 + e] : 0.0f;
                            const int cc = c0 + c8, kb = cc >> 6, ch = (cc & 63) >> 3;
                            f_store8(pa + kb * TILE_A, pa_mid + kb * TILE_A, row, ch, z);
                        }
                    }
                }
                f_fence_async();
                f_fence_before();
                __syncwarp();
                if (lane == 0) g_mbar_arrive(y2_ready);
                FTOC(6);
                // ---------- feed-forward hidden chunks (even ones; group B takes the odd ones) ----------
                for (int j = 0; j < NH && alive; j += 2) {
                    const uint32_t cit = lit * (uint32_t)NH + j, hb = cit % NB;
                    if (!f_wait(&hid_full[hb], (cit / NB) & 1, B_HIDFULL, dbg)) { alive = false; break; }
                    FTOC(7);
                    f_fence_after();
                    float hdn[64];
                    f_tmem_ld32(trow + (uint32_t)(CP + hb * 64), hdn);
                    f_tmem_ld32(trow + (uint32_t)(CP + hb * 64 + 32), hdn + 32);
                    f_wait_ld();
                    f_fence_before();
                    __syncwarp();
                    if (lane == 0) g_mbar_arrive(&hid_free[hb]);
                    const float* b1 = lv.b1 + j * 64;
                    #pragma unroll
                    for (int d4 = 0; d4 < 64; d4 += 4) {
                        const float4 c = __ldg(reinterpret_cast<const float4*>(b1 + d4));
                        hdn[d4] = f_gelu(hdn[d4] + c.x); hdn[d4 + 1] = f_gelu(hdn[d4 + 1] + c.y);
                        hdn[d4 + 2] = f_gelu(hdn[d4 + 2] + c.z); hdn[d4 + 3] = f_gelu(hdn[d4 + 3] + c.w);
                    }
                    const uint32_t pit = pit0 + heads + j;
                    const int pbuf = pit % NPC;
                    FTOC(8);
                    if (pit >= (uint32_t)NPC) {
                        if (!f_wait(&pc_free[pbuf], fcnt[pbuf] & 1, B_PCFREE, dbg)) { alive = false; break; }
                        ++fcnt[pbuf];
                    }
                    FTOC(4);
                    uint8_t* pch = pc + pbuf * K::PCBUF;
                    #pragma unroll
                    for (int ch = 0; ch < 8; ++ch) f_store8(pch, pch + TILE_A, row, ch, hdn + 8 * ch);
                    f_fence_async();
                    __syncwarp();
                    if (lane == 0) g_mbar_arrive(&pc_ready[pbuf]);
                    FTOC(8);
                }
                if (!alive) break;
                if (!f_wait(layer_done, lit & 1, B_LAYER, dbg)) { alive = false; break; }
                FTOC(9);
                f_fence_after();
            }
            if (!alive) break;
            // x tile: TMEM + cumulative bias -> global
            #pragma unroll
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
            FTOC(10);
        }
        if (tim && warp == 4 && lane == 0) { for (int i = 0; i < 11; ++i) p.timing[8 + i] = tacc[i]; }
    } else {
        // ---- compute group B: the odd feed-forward chunks ----
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t lit = 0;
        uint32_t fcnt[2] = {0, 0};                                     // signalled waits so far on pc_free[group B][buffer]
        bool alive = true;
        for (int tile = blockIdx.x; tile < p.tiles && alive; tile += gridDim.x) {
            for (int l = 0; l < depth && alive; ++l, ++lit) {
                const float* b1l = p.vecs[l].b1;
                const uint32_t pit0 = lit * (uint32_t)(heads + NH);
                for (int j = 1; j < NH && alive; j += 2) {
                    const uint32_t cit = lit * (uint32_t)NH + j, hb = cit % NB;
                    if (!f_wait(&hid_full[hb], (cit / NB) & 1, B_HIDFULL, dbg)) { alive = false; break; }
                    f_fence_after();
                    float hdn[64];
                    f_tmem_ld32(trow + (uint32_t)(CP + hb * 64), hdn);
                    f_tmem_ld32(trow + (uint32_t)(CP + hb * 64 + 32), hdn + 32);
                    f_wait_ld();
                    f_fence_before();
                    __syncwarp();
                    if (lane == 0) g_mbar_arrive(&hid_free[hb]);
                    const float* b1 = b1l + j * 64;
                    #pragma unroll
                    for (int d4 = 0; d4 < 64; d4 += 4) {
                        const float4 c = __ldg(reinterpret_cast<const float4*>(b1 + d4));
                        hdn[d4] = f_gelu(hdn[d4] + c.x); hdn[d4 + 1] = f_gelu(hdn[d4 + 1] + c.y);
                        hdn[d4 + 2] = f_gelu(hdn[d4 + 2] + c.z); hdn[d4 + 3] = f_gelu(hdn[d4 + 3] + c.w);
                    }
                    const uint32_t pit = pit0 + heads + j;
                    const int pbuf = pit % NPC;
                    if (pit >= (uint32_t)NPC) {
                        if (!f_wait(&pc_free[2 + pbuf], fcnt[pbuf] & 1, B_PCFREE, dbg)) { alive = false; break; }
                        ++fcnt[pbuf];
                    }
                    uint8_t* pch = pc + pbuf * K::PCBUF;
                    #pragma unroll
                    for (int ch = 0; ch < 8; ++ch) f_store8(pch, pch + TILE_A, row, ch, hdn + 8 * ch);
                    f_fence_async();
                    __syncwarp();
                    if (lane == 0) g_mbar_arrive(&pc_ready[pbuf]);
                }
            }
        }
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
    std::vector<FusedLayerVecs> vecs(st.depth);
    // cumulative biases: cb1[l] = every out-projection / FF2 bias of the layers before l, cb2[l] = cb1[l] + b_out[l]
    std::vector<float> cbs((size_t)(2 * st.depth + 1) * c, 0.0f), run(c, 0.0f);
    CTO_CHECK(cudaMalloc(&st.fused_cb, sizeof(float) * cbs.size()));
    auto host = [&](const float* dev_ptr) { return host_blob + (dev_ptr - dev_blob); };
    for (int d = 0; d < st.depth; ++d) {
        const CvtLayer& L = st.layers[d];
        const size_t before = stream.size();
        const float *wq = host(L.q_pw), *wkv = host(L.kv_pw), *wo = host(L.out_w), *w1 = host(L.ff1_w), *w2 = host(L.ff2_w);
        auto out_tiles = [&](int h) { for (int t = 0; t < nt; ++t) emit_tile(stream, wo, c, inner, t * 64, h * 64); };
        auto ff2_tiles = [&](int j) { for (int t = 0; t < nt; ++t) emit_tile(stream, w2, c, 4 * c, t * 64, j * 64); };
        for (int h = 0; h < st.heads; ++h) {
            for (int k = 0; k < kb; ++k) emit_tile(stream, wq, inner, c, h * 64, k * 64);
            for (int k = 0; k < kb; ++k) emit_tile(stream, wkv, 2 * inner, c, h * 64, k * 64);
            for (int k = 0; k < kb; ++k) emit_tile(stream, wkv, 2 * inner, c, inner + h * 64, k * 64);
            if (h >= 1) out_tiles(h - 1);
        }
        out_tiles(st.heads - 1);
        for (int j = 0; j < nh; ++j) {
            for (int k = 0; k < kb; ++k) emit_tile(stream, w1, 4 * c, c, j * 64, k * 64);
            if (j >= LAG) ff2_tiles(j - LAG);
        }
        for (int j = nh - LAG; j < nh; ++j) ff2_tiles(j);
        const size_t bytes = stream.size() - before;
        if (d == 0) st.fused_layer_bytes = (long long)bytes;
        CTO_REQUIRE((long long)bytes == st.fused_layer_bytes &&
                        bytes == (size_t)(st.heads * (3 * kb + nt) + nh * (kb + nt)) * WSTAGE,
                    "aff_fused: weight stream of layer %d has %zu bytes", d, bytes);
        FusedLayerVecs& v = vecs[d];
        v.ln1_g = L.ln1_g; v.ln1_b = L.ln1_b; v.tq = L.q_dw; v.tk = L.kv_dw; v.bq = L.q_bias; v.bkv = L.kv_bias;
        v.ln2_g = L.ln2_g; v.ln2_b = L.ln2_b; v.b1 = L.ff1_b;
        v.cb1 = st.fused_cb + (size_t)(2 * d) * c;
        v.cb2 = st.fused_cb + (size_t)(2 * d + 1) * c;
        const float *bo = host(L.out_b), *b2 = host(L.ff2_b);
        for (int i = 0; i < c; ++i) {
            cbs[(size_t)(2 * d) * c + i] = run[i];
            run[i] += bo[i];
            cbs[(size_t)(2 * d + 1) * c + i] = run[i];
            run[i] += b2[i];
        }
    }
    for (int i = 0; i < c; ++i) cbs[(size_t)(2 * st.depth) * c + i] = run[i];
    CTO_CHECK(cudaMemcpy(st.fused_cb, cbs.data(), sizeof(float) * cbs.size(), cudaMemcpyHostToDevice));
    CTO_CHECK(cudaMalloc(&st.fused_stream, stream.size()));
    CTO_CHECK(cudaMemcpy(st.fused_stream, stream.data(), stream.size(), cudaMemcpyHostToDevice));
    CTO_CHECK(cudaMalloc(&st.fused_vecs, sizeof(FusedLayerVecs) * st.depth));
    CTO_CHECK(cudaMemcpy(st.fused_vecs, vecs.data(), sizeof(FusedLayerVecs) * st.depth, cudaMemcpyHostToDevice));
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
    p.vecs = st.fused_vecs;
    p.cb_final = st.fused_cb + (size_t)(2 * st.depth) * st.c;
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
