// tcgen05 (5th-gen tensor core) GEMM for sm_100a with fp32-grade accuracy ("bf16x3"):
//
//     C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual)
//
// A single reduced-precision pass (TF32 or bf16) moves the AFF/NEG logits by up to 1e-2, ten times the
// parity contract.  Every fp32 operand is therefore split into two bf16 parts, hi = bf16(x) and
// mid = bf16(x - hi) (x - hi - mid is below 2^-18 |x|), and three kind::f16 MMAs are accumulated per
// k-step in fp32 TMEM:  hi*hi + mid*hi + hi*mid  (the dropped mid*mid term is 2^-18 relative too).
// Compared with the 3xTF32 split this kernel used first, the bf16 MMAs run at twice the tensor rate
// and the split operands take half the shared memory; measured max |d logit| stays below 1e-4.
// Weights are split once at load time (W_hi, W_mid as bf16 in HBM); activations on the fly.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer: A as two 128x32 fp32 boxes, W_hi, W_mid (BN x 64 bf16) -> 3-stage smem
//               ring, 128-byte swizzle as the UMMA descriptors expect
//   warps 2-5   converter: fp32 A boxes -> bf16 A_hi / A_mid tiles (128 x 64, in place: row r of the two
//               boxes is exactly the storage of row r of the two tiles), fence to the async proxy
//   warp 1      MMA issuer (one elected thread): tcgen05.mma.kind::f16 (bf16), M=128, N=BN, K=16;
//               tcgen05.commit releases smem stages and publishes the accumulator
//   warps 6-13  epilogue (two groups of four warps, alternating 32-column slabs): tcgen05.ld (one TMEM
//               lane = one output row), bias / GELU / SELU / residual in fp32, swizzled staging in smem,
//               TMA store (full 128-byte lines, M tail clipped by the tensor map)
// The accumulator is double buffered in TMEM (2 x 128 columns) so the epilogue of tile i overlaps
// the main loop of tile i+1.  K and M tails are zero-filled by TMA; N must be a multiple of 64, K of 8.
#include "nn_kernels.cuh"
#include <cuda.h>

namespace cto {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;                       // K elements per stage = one 128-byte swizzle row of bf16
constexpr int MAX_BN = 256;                  // TMEM columns of one accumulator buffer (two buffers = all 512 columns)
constexpr int TILE_BYTES = BM * 128;         // 16 KB: one fp32 A box (128 x 32) on arrival, one bf16 tile (128 x 64) after
                                             // the conversion; W tiles use BN*128 bytes of theirs
constexpr int STAGE_BYTES = 4 * TILE_BYTES;  // A box 0 -> A_hi | A box 1 -> A_mid | W_hi | W_mid          (BN <= 128)
constexpr int STAGES = 3;
// BN = 256 ("wide" tiles of the transposed GRU projections): a stage is A_hi | A_mid | W_hi (32 KB) | W_mid (32 KB) = 96 KB and
// there are two of them in the same 192 KB.  One MMA then reads 4 KB of A + 8 KB of W per 128 clk = 96 B/clk of shared
// memory instead of 8 KB per 64 clk = the full 128 B/clk, and a tile needs 25 % fewer L2 bytes per output element.
constexpr int WIDE_STAGE_BYTES = 2 * TILE_BYTES + 2 * 256 * 128;
constexpr int WIDE_STAGES = 2;
constexpr int SLAB = 32;                     // epilogue works on 128 x 32 fp32 slabs
constexpr int STAGING_BYTES = BM * SLAB * 4; // 16 KB, two of them
constexpr int THREADS = 448;                 // warp 0 TMA, 1 MMA, 2-5 converter, 6-13 two epilogue groups
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGING_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// two fp32 -> packed bf16x2 (round to nearest even); `lo` lands in the low half = the lower address
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// x[0..1] -> hi pair, mid pair:  hi = bf16(x), mid = bf16(x - hi)  (x - hi is exact in fp32)
__device__ __forceinline__ void split2_bf16(float x0, float x1, uint32_t& hi, uint32_t& mid) {
    hi = pack_bf16x2(x0, x1);
    const float r0 = x0 - __uint_as_float(hi << 16);
    const float r1 = x1 - __uint_as_float(hi & 0xFFFF0000u);
    mid = pack_bf16x2(r0, r1);
}
// one lane of a converged warp; the surrounding loop stays warp-uniform so that descriptors live in uniform
// registers (inside an `if (lane == 0)` region ptxas wraps every UTCHMMA / UTMALDG in an ELECT /
// R2UR.BROADCAST / BRA.U.ANY loop, ~70 cycles per MMA)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}" : "+r"(pred));
    return pred != 0;
}

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes, 8-row groups
// 1024 bytes apart (SBO), LBO unused (=1), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                               // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                               // version = 1
    d |= (uint64_t)2 << 61;                               // layout type: SWIZZLE_128B
    return d;
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=bn
__device__ __forceinline__ uint32_t make_idesc_bf16(int bn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float selu(float x) {
    const float alpha = 1.6732632423543772848170429916717f;
    const float scale = 1.0507009873554804934193349852946f;
    return scale * (x > 0.0f ? x : alpha * expm1f(x));
}

template <int ACT>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_amid,
                   const __grid_constant__ CUtensorMap tma_whi, const __grid_constant__ CUtensorMap tma_wmid,
                   const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_cmid,
                   const float* __restrict__ bias, const float* residual,
                   int64_t ldr, int64_t m_total, int n_total, int k_total, int bn, int flags, int dbg, long long* timing) {
    const bool presplit = flags & GEMM_A_PRESPLIT;       // A arrives as bf16 hi / mid planes: no converter pass
    const bool bias_row = flags & GEMM_BIAS_PER_ROW;     // bias indexed by the output row (transposed products)
    const bool n_major = flags & GEMM_TILES_N_MAJOR;     // consecutive CTAs share the W tile instead of the A tile
    const bool out_split = flags & GEMM_OUT_SPLIT;       // C leaves as bf16 hi / mid planes (tma_c / tma_cmid)
    extern __shared__ uint8_t smem_raw[];
    const bool tim = dbg && blockIdx.x == 0 && timing != nullptr;    // per-phase cycle counters (profiles/)
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    #define TIC long long _t0 = tim ? clock64() : 0
    #define TOC(i) do { if (tim) { long long _t1 = clock64(); tacc[i] += _t1 - _t0; _t0 = _t1; } } while (0)
    // 1024-byte alignment by POINTER arithmetic on the __shared__ array: rounding through uintptr_t makes the
    // compiler lose the address space and emit generic LD.E / ST.E for every shared-memory access
    uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* staging = base + STAGES * STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(staging + 2 * STAGING_BYTES);
    uint64_t* conv = full + STAGES;
    uint64_t* empty = conv + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (k_total + BK - 1) / BK;
    const int n_tiles = (n_total + bn - 1) / bn;     // a ragged last column tile is zero-filled / clipped by TMA
    const int64_t m_tiles = (m_total + BM - 1) / BM;
    const int64_t num_tiles = m_tiles * n_tiles;
    const int nstages = bn > 128 ? WIDE_STAGES : STAGES;
    const int stage_bytes = bn > 128 ? WIDE_STAGE_BYTES : STAGE_BYTES;
    const int wmid_off = 2 * TILE_BYTES + (bn > 128 ? bn * 128 : TILE_BYTES);      // W_mid behind W_hi

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], 128); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {                                                  // ---- TMA producer (warp-uniform loop, elected lane issues) ----
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (n_major ? (int)(t % m_tiles) : (int)(t / n_tiles)) * BM;
                const int n0 = (n_major ? (int)(t / m_tiles) : (int)(t % n_tiles)) * bn;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % nstages;
                    const uint32_t ph = (it / nstages) & 1;
                    TIC;
                    mbar_wait(&empty[s], ph ^ 1);
                    TOC(0);
                    uint8_t* st = base + s * stage_bytes;
                    // fp32 A: the second 128 x 32 box is skipped when it holds no live column; pre-split A: hi and mid tile
                    const bool two = presplit || k_total - kb * BK > 32;
                    if (elect_one()) {
                        mbar_expect_tx(&full[s], (uint32_t)((two ? 2 : 1) * TILE_BYTES + 2 * bn * 128));
                        tma_load_2d(&tma_a, &full[s], st, kb * BK, m0);
                        if (presplit) tma_load_2d(&tma_amid, &full[s], st + TILE_BYTES, kb * BK, m0);
                        else if (two) tma_load_2d(&tma_a, &full[s], st + TILE_BYTES, kb * BK + 32, m0);
                        tma_load_2d(&tma_whi, &full[s], st + 2 * TILE_BYTES, kb * BK, n0);
                        tma_load_2d(&tma_wmid, &full[s], st + wmid_off, kb * BK, n0);
                    }
                    __syncwarp();
                    TOC(1);
                }
            }
            if (tim && lane == 0) { timing[0] = tacc[0]; timing[1] = tacc[1]; }
        }
    } else if (warp == 1) {
        {                                                  // ---- MMA issuer (warp-uniform loop, elected lane issues) ----
            const uint32_t idesc = make_idesc_bf16(bn);
            uint32_t it = 0, acc_it = 0;
            for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++acc_it) {
                const uint32_t ab = acc_it & 1, aph = (acc_it >> 1) & 1;
                TIC;
                mbar_wait(&acc_empty[ab], aph ^ 1);        // epilogue has drained this accumulator
                TOC(0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + ab * MAX_BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % nstages;
                    const uint32_t ph = (it / nstages) & 1;
                    mbar_wait(&full[s], ph);               // W tiles landed
                    TOC(1);
                    if (!presplit) mbar_wait(&conv[s], ph);  // A split into bf16 hi / mid
                    TOC(2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(base + s * stage_bytes);
                    const uint64_t d_ahi = make_desc_k_sw128(a_addr);
                    const uint64_t d_amid = make_desc_k_sw128(a_addr + TILE_BYTES);
                    const uint64_t d_whi = make_desc_k_sw128(a_addr + 2 * TILE_BYTES);
                    const uint64_t d_wmid = make_desc_k_sw128(a_addr + wmid_off);
                    const int ksteps = min(BK / 16, (k_total - kb * BK + 15) >> 4);   // K tail: skip all-zero k-steps
                    if (elect_one()) {
                        #pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            if (k >= ksteps) break;
                            // 16 bf16 = 32 bytes along the swizzle row: +2 in the (>>4) start-address field
                            const uint64_t o = (uint64_t)(k * 2);
                            mma_bf16(acc, d_ahi + o, d_whi + o, idesc, (kb | k) ? 1u : 0u);
                            mma_bf16(acc, d_amid + o, d_whi + o, idesc, 1u);
                            mma_bf16(acc, d_ahi + o, d_wmid + o, idesc, 1u);
                        }
                        tcgen05_commit(&empty[s]);             // stage reusable once these MMAs retire
                        if (kb == num_kb - 1) tcgen05_commit(&acc_full[ab]);   // accumulator complete
                    }
                    __syncwarp();
                    TOC(3);
                }
            }
            if (tim && lane == 0) { for (int i = 0; i < 4; ++i) timing[4 + i] = tacc[i]; }
        }
    } else if (warp < 6) {                                 // ---- converter: warps 2..5 ----
        const int ct = threadIdx.x - 64;                   // 0..127
        const int cw = warp - 2;
        const int c = lane & 7;                            // 16-byte output chunk = 8 consecutive k
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < num_tiles && !presplit; t += gridDim.x) {
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % nstages;
                const uint32_t ph = (it / nstages) & 1;
                TIC;
                mbar_wait(&full[s], ph);
                TOC(0);
                uint8_t* st = base + s * stage_bytes;
                const int nchunks = min(8, ((k_total - kb * BK + 15) >> 4) * 2);   // chunks the MMAs will read
                // fp32 element (r, k) sits in box k/32 at 16-byte unit ((k%32)/4) ^ (r%8) of row r; bf16 element
                // (r, k) goes to chunk (k/8) ^ (r%8) of row r.  Eight lanes own one row per pass, so the in-place
                // rewrite only needs the warp to finish its loads before its stores.
                #pragma unroll
                for (int half = 0; half < 2; ++half) {             // 4 rows per lane in flight: loads, then stores
                    float4 v[4][2];
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = (half * 4 + j) * 16 + cw * 4 + (lane >> 3);
                        const uint8_t* src = st + (c >> 2) * TILE_BYTES + (r >> 3) * 1024 + (r & 7) * 128;
                        const int u0 = 2 * (c & 3);
                        v[j][0] = v[j][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (c < nchunks) {
                            v[j][0] = *reinterpret_cast<const float4*>(src + ((u0 ^ (r & 7)) << 4));
                            v[j][1] = *reinterpret_cast<const float4*>(src + (((u0 + 1) ^ (r & 7)) << 4));
                        }
                    }
                    __syncwarp();
                    if (c < nchunks) {
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int r = (half * 4 + j) * 16 + cw * 4 + (lane >> 3);
                            uint4 hi, mid;
                            split2_bf16(v[j][0].x, v[j][0].y, hi.x, mid.x);
                            split2_bf16(v[j][0].z, v[j][0].w, hi.y, mid.y);
                            split2_bf16(v[j][1].x, v[j][1].y, hi.z, mid.z);
                            split2_bf16(v[j][1].z, v[j][1].w, hi.w, mid.w);
                            const uint32_t dst = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
                            *reinterpret_cast<uint4*>(st + dst) = hi;
                            *reinterpret_cast<uint4*>(st + TILE_BYTES + dst) = mid;
                        }
                    }
                    __syncwarp();
                }
                TOC(1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> UMMA reads
                TOC(2);
                mbar_arrive(&conv[s]);
            }
        }
        if (tim && ct == 0) { for (int i = 0; i < 3; ++i) timing[8 + i] = tacc[i]; }
    } else {                                               // ---- epilogue: warps 6..13, two groups of four ----
        // both groups cover all four TMEM lane quadrants; group g takes the 32-column slabs with index % 2 == g
        const int quad = warp & 3;                         // TMEM lanes [32*quad, 32*quad+32)
        const int grp = (warp - 6) >> 2;
        const int et = (threadIdx.x - 192) & 127;          // 0..127 inside the group
        const int r_in_tile = quad * 32 + lane;
        uint8_t* stg = staging + grp * STAGING_BYTES;      // one staging buffer per group
        uint32_t acc_it = 0;
        for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++acc_it) {
            const int m0 = (n_major ? (int)(t % m_tiles) : (int)(t / n_tiles)) * BM;
            const int n0 = (n_major ? (int)(t / m_tiles) : (int)(t % n_tiles)) * bn;
            const uint32_t ab = acc_it & 1, aph = (acc_it >> 1) & 1;
            TIC;
            mbar_wait(&acc_full[ab], aph);
            TOC(0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t row = (int64_t)m0 + r_in_tile;
            const bool row_ok = row < m_total;
            const float rbias = (bias_row && bias && row_ok) ? __ldg(bias + row) : 0.0f;
            for (int cb = grp * SLAB; cb < bn; cb += 2 * SLAB) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ab * MAX_BN + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                      "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                TOC(1);
                if (cb + 2 * SLAB >= bn) {                 // this group's last slab of the tile: hand the accumulator back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&acc_empty[ab]);
                }
                TOC(2);
                const float* rrow = (residual && row_ok) ? residual + row * ldr + n0 + cb : nullptr;
                float4 o[8];
                #pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias_row) bv = make_float4(rbias, rbias, rbias, rbias);
                    else if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + n0 + cb) + q);
                    float v[4] = {__uint_as_float(r[q * 4]) + bv.x, __uint_as_float(r[q * 4 + 1]) + bv.y,
                                  __uint_as_float(r[q * 4 + 2]) + bv.z, __uint_as_float(r[q * 4 + 3]) + bv.w};
                    if (ACT == ACT_GELU) {
                        #pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = gelu_erf(v[e]);
                    } else if (ACT == ACT_SELU) {
                        #pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = selu(v[e]);
                    }
                    if (rrow) {
                        const float4 rv = *reinterpret_cast<const float4*>(rrow + q * 4);
                        v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
                    }
                    o[q] = make_float4(v[0], v[1], v[2], v[3]);
                }
                TOC(3);
                // this warp's previous TMA store must have finished reading its 32-row staging strip
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                uint8_t* wstg = stg + quad * (32 * 128);       // this warp's 4 KB strip of the staging buffer
                if (out_split) {
                    // two un-swizzled 32 x 32 bf16 boxes (64-byte rows): hi plane, then mid plane
                    #pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 hi, mid;
                        split2_bf16(o[2 * q].x, o[2 * q].y, hi.x, mid.x);
                        split2_bf16(o[2 * q].z, o[2 * q].w, hi.y, mid.y);
                        split2_bf16(o[2 * q + 1].x, o[2 * q + 1].y, hi.z, mid.z);
                        split2_bf16(o[2 * q + 1].z, o[2 * q + 1].w, hi.w, mid.w);
                        *reinterpret_cast<uint4*>(wstg + lane * 64 + q * 16) = hi;
                        *reinterpret_cast<uint4*>(wstg + 2048 + lane * 64 + q * 16) = mid;
                    }
                } else {
                    #pragma unroll
                    for (int q = 0; q < 8; ++q)   // 128-byte swizzle: 16-byte chunk q of row r lives at chunk (q ^ (r % 8))
                        *reinterpret_cast<float4*>(stg + r_in_tile * 128 + ((q ^ (r_in_tile & 7)) << 4)) = o[q];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                TOC(4);
                __syncwarp();
                TOC(5);
                if (lane == 0) {                               // one 32 x 32 box per warp: no cross-warp barrier in the epilogue
                    tma_store_2d(&tma_c, wstg, n0 + cb, m0 + quad * 32);
                    if (out_split) tma_store_2d(&tma_cmid, wstg + 2048, n0 + cb, m0 + quad * 32);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (tim && et == 0 && grp == 0) { for (int i = 0; i < 6; ++i) timing[12 + i] = tacc[i]; timing[20] = (long long)num_tiles; timing[21] = num_kb; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// fp32 -> bf16 hi + bf16 mid (both round to nearest even)
__global__ void split_bf16_kernel(const float* __restrict__ w, uint16_t* __restrict__ hi, uint16_t* __restrict__ mid, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h, m;
    split2_bf16(w[i], 0.0f, h, m);
    hi[i] = (uint16_t)(h & 0xFFFFu);
    mid[i] = (uint16_t)(m & 0xFFFFu);
}

// bf16 hi + mid planes -> fp32 (the inverse of the split up to 2^-17 relative; used by tests and debugging)
__global__ void join_bf16_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ mid, float* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __uint_as_float((uint32_t)hi[i] << 16) + __uint_as_float((uint32_t)mid[i] << 16);
}

// 2-D tiled tensor map with 128-byte swizzle; the box is (128 bytes of the inner dimension) x box_rows
// cuTensorMapEncodeTiled, resolved through the runtime so that the library has no link-time dependency on libcuda
typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int get_encode(encode_fn* out) {
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CTO_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        CTO_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
        encode = reinterpret_cast<encode_fn>(fn);
    }
    *out = encode;
    return 0;
}

// 2-D tiled tensor map with 128-byte swizzle; the box is (128 bytes of the inner dimension) x box_rows
static int make_map_any(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* ptr, int64_t rows, int64_t cols,
                        int64_t ld, int box_rows, int box_cols = 0, bool swizzle = true) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(box_cols ? box_cols : 128 / elem_bytes), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    encode_fn encode;
    if (get_encode(&encode)) return 1;
    CUresult r = encode(map, dtype, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return 1;
    }
    return 0;
}

// un-swizzled fp32 map of rank 2..5 over a row-major [rows][ld] array: dimension 0 = columns, dimension i >= 1 walks rows with
// a stride of row_stride[i - 1] rows (the views may overlap).  One TMA instruction then gathers a box of several row groups.
int make_map_rows_nd(CUtensorMap* map, const float* ptr, int rank, int64_t cols, int64_t ld, const int64_t* dim_size,
                     const int64_t* row_stride, const int* box) {
    CTO_REQUIRE(rank >= 2 && rank <= 5, "make_map_rows_nd: rank %d", rank);
    cuuint64_t dims[5], strides[4];
    cuuint32_t bx[5], estr[5] = {1, 1, 1, 1, 1};
    dims[0] = (cuuint64_t)cols; bx[0] = (cuuint32_t)box[0];
    for (int i = 1; i < rank; ++i) {
        dims[i] = (cuuint64_t)dim_size[i - 1];
        strides[i - 1] = (cuuint64_t)row_stride[i - 1] * (cuuint64_t)ld * 4;
        bx[i] = (cuuint32_t)box[i];
    }
    encode_fn encode;
    if (get_encode(&encode)) return 1;
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(ptr), dims, strides, bx, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (rank %d) failed (CUresult %d) cols=%lld ld=%lld", rank, (int)r, (long long)cols, (long long)ld);
        return 1;
    }
    return 0;
}

int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    return make_map_any(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, rows, cols, ld, box_rows);
}
// un-swizzled fp32 box (box_cols x box_rows) for staging that is read by ordinary LDS
int make_map_plain(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols) {
    return make_map_any(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, ptr, rows, cols, ld, box_rows, box_cols, false);
}
int make_map_bf16(CUtensorMap* map, const uint16_t* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    return make_map_any(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, rows, cols, ld, box_rows);
}

}  // namespace tc

long long* g_gemm_timing = nullptr;   // device buffer [32] filled by CTA 0 when (dbg & 16)
int g_gemm_debug = 0;       // non-zero: CTA 0 records per-phase cycle counters into g_gemm_timing (cto_debug_set)

int launch_split_bf16(const float* w, uint16_t* hi, uint16_t* mid, int64_t n, cudaStream_t s) {
    if (n <= 0) return 0;
    tc::split_bf16_kernel<<<ceil_div(n, 256), 256, 0, s>>>(w, hi, mid, n);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_join_bf16(const uint16_t* hi, const uint16_t* mid, float* out, int64_t n, cudaStream_t s) {
    if (n <= 0) return 0;
    tc::join_bf16_kernel<<<ceil_div(n, 256), 256, 0, s>>>(hi, mid, out, n);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

bool gemm_tc_supported(const float* a, int64_t lda, const void* w, int64_t m, int n, int k, const float* c, int64_t ldc,
                       const float* residual, int64_t ldr) {
    if (m <= 0 || n < 64 || n % 64 != 0 || k < 8 || k % 8 != 0) return false;     // bf16 W rows: 16-byte strides
    if (lda % 4 != 0 || ldc % 4 != 0 || (residual && ldr % 4 != 0)) return false;
    if (m >= (1ll << 31) - tc::BM) return false;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) |
                           reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(residual);
    return (bits & 15) == 0;
}

int launch_gemm_tc_ex(const GemmTc& g, cudaStream_t s) {
    const bool presplit = g.flags & GEMM_A_PRESPLIT, out_split = g.flags & GEMM_OUT_SPLIT;
    const void* c_any = out_split ? (const void*)g.c_hi : (const void*)g.c;
    const void* a_any = presplit ? (const void*)g.a_hi : (const void*)g.a;
    const int64_t a_align = presplit ? 8 : 4;            // elements per 16 bytes
    const uintptr_t bits = reinterpret_cast<uintptr_t>(a_any) | reinterpret_cast<uintptr_t>(g.a_mid) |
                           reinterpret_cast<uintptr_t>(g.w_hi) | reinterpret_cast<uintptr_t>(g.w_mid) |
                           reinterpret_cast<uintptr_t>(c_any) | reinterpret_cast<uintptr_t>(g.c_mid) |
                           reinterpret_cast<uintptr_t>(g.residual);
    CTO_REQUIRE(a_any && g.w_hi && g.w_mid && c_any && (!presplit || g.a_mid) && (!out_split || (g.c_mid && g.ldc % 8 == 0)) &&
                    (bits & 15) == 0 && g.m > 0 &&
                    g.m < (1ll << 31) - tc::BM && g.n >= 64 && g.n % 64 == 0 && g.k >= 8 && g.k % 8 == 0 &&
                    g.lda % a_align == 0 && g.ldw % 8 == 0 && g.ldc % 4 == 0 && (!g.residual || g.ldr % 4 == 0),
                "gemm_tc: unsupported shape/alignment m=%lld n=%d k=%d lda=%lld ldw=%lld ldc=%lld flags=%d", (long long)g.m,
                g.n, g.k, (long long)g.lda, (long long)g.ldw, (long long)g.ldc, g.flags);
    const int sm_count = device_sm_count();
    CTO_REQUIRE(sm_count > 0, "gemm_tc: no CUDA device");
    if (g.act == ACT_GELU) CTO_CHECK(set_max_dynamic_smem(tc::gemm_bf16x3_kernel<ACT_GELU>, tc::SMEM_BYTES));
    else if (g.act == ACT_SELU) CTO_CHECK(set_max_dynamic_smem(tc::gemm_bf16x3_kernel<ACT_SELU>, tc::SMEM_BYTES));
    else CTO_CHECK(set_max_dynamic_smem(tc::gemm_bf16x3_kernel<ACT_NONE>, tc::SMEM_BYTES));
    // 128-wide tiles unless that leaves SMs without a tile (long-K, few-row GEMMs such as the NEG fc1)
    int bn = (g.n % 128 == 0 && (int64_t)ceil_div(g.m, tc::BM) * (g.n / 128) >= sm_count) ? 128 : 64;
    // wide tiles for the transposed GRU projections (few rows = the layer's 6H gate units, very many columns)
    if ((g.flags & GEMM_WIDE_N) && presplit && bn == 128 && g.n >= 256 * 8 &&
        (int64_t)ceil_div(g.m, tc::BM) * ceil_div(g.n, 256) >= 2 * sm_count)
        bn = 256;
    CUtensorMap map_a, map_amid, map_whi, map_wmid, map_c, map_cmid;
    if (presplit) {
        if (tc::make_map_bf16(&map_a, g.a_hi, g.m, g.k, g.lda, tc::BM)) return 1;
        if (tc::make_map_bf16(&map_amid, g.a_mid, g.m, g.k, g.lda, tc::BM)) return 1;
    } else {
        if (tc::make_map(&map_a, g.a, g.m, g.k, g.lda, tc::BM)) return 1;
        map_amid = map_a;
    }
    if (tc::make_map_bf16(&map_whi, g.w_hi, g.n, g.k, g.ldw, bn)) return 1;
    if (tc::make_map_bf16(&map_wmid, g.w_mid, g.n, g.k, g.ldw, bn)) return 1;
    // every epilogue warp stores its own 32-row strip: one swizzled 32 x 32 fp32 box, or two plain 32 x 32 bf16 boxes
    if (out_split) {
        if (tc::make_map_any(&map_c, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.c_hi, g.m, g.n, g.ldc, 32, 32, false)) return 1;
        if (tc::make_map_any(&map_cmid, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.c_mid, g.m, g.n, g.ldc, 32, 32, false)) return 1;
    } else {
        if (tc::make_map(&map_c, g.c, g.m, g.n, g.ldc, 32)) return 1;
        map_cmid = map_c;
    }
    const int64_t tiles = (int64_t)ceil_div(g.m, tc::BM) * ceil_div(g.n, bn);
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
#define CTO_LAUNCH_GEMM(A)                                                                                             \
    tc::gemm_bf16x3_kernel<A><<<grid, tc::THREADS, tc::SMEM_BYTES, s>>>(map_a, map_amid, map_whi, map_wmid, map_c, map_cmid, g.bias, \
                                                                       g.residual, g.ldr, g.m, g.n, g.k, bn, g.flags,   \
                                                                       g_gemm_debug, g_gemm_timing)
    if (g.act == ACT_GELU) CTO_LAUNCH_GEMM(ACT_GELU);
    else if (g.act == ACT_SELU) CTO_LAUNCH_GEMM(ACT_SELU);
    else CTO_LAUNCH_GEMM(ACT_NONE);
#undef CTO_LAUNCH_GEMM
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

int launch_gemm_tc(const float* a, int64_t lda, const uint16_t* w_hi, const uint16_t* w_mid, const float* bias,
                   const float* residual, int64_t ldr, float* c, int64_t ldc, int64_t m, int n, int k, int act,
                   cudaStream_t s) {
    GemmTc g;
    g.a = a; g.lda = lda; g.w_hi = w_hi; g.w_mid = w_mid; g.ldw = k; g.bias = bias; g.residual = residual; g.ldr = ldr;
    g.c = c; g.ldc = ldc; g.m = m; g.n = n; g.k = k; g.act = act;
    return launch_gemm_tc_ex(g, s);
}

}  // namespace cto
