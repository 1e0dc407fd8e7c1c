// tcgen05 (5th-gen tensor core) TF32 GEMM for sm_100a:  C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual)
//
// Both operands are K-major fp32 in global memory and are consumed as TF32 (fp32 accumulate in TMEM):
//   - a TMA producer warp streams 128 x 32 (A) and BN x 32 (W) fp32 boxes into a multi-stage
//     shared-memory ring with the 128-byte swizzle the UMMA descriptors expect;
//   - one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) four times per stage and
//     releases the stage with tcgen05.commit;
//   - four epilogue warps read the accumulator with tcgen05.ld (one TMEM lane = one output row),
//     apply bias / GELU / SELU / residual in fp32 and store rows with 16-byte stores.
// Used for the dense contractions of the hot path: GRU input projections and fc1 (clairs/model.py:
// 412-420), the CvT 1x1 convolutions and heads (ibid. 78-118, 214-224).  K and M tails are
// zero-filled by TMA; N must be a multiple of BN (16, 64 or 128).
#include "nn_kernels.cuh"
#include <cuda.h>

namespace cto {

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;                       // fp32 elements = one 128-byte swizzle row
constexpr int MAX_BN = 128;
constexpr int A_BYTES = BM * BK * 4;         // 16 KB
constexpr int W_BYTES = MAX_BN * BK * 4;     // 16 KB
constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
constexpr int THREADS = 192;                 // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-5 epilogue

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
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

// instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N=bn
__device__ __forceinline__ uint32_t make_idesc_tf32(int bn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float selu(float x) {
    const float alpha = 1.6732632423543772848170429916717f;
    const float scale = 1.0507009873554804934193349852946f;
    return scale * (x > 0.0f ? x : alpha * expm1f(x));
}

template <int STAGES>
__global__ void __launch_bounds__(THREADS)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w,
                 const float* __restrict__ bias, const float* residual, int64_t ldr, float* c, int64_t ldc,
                 int64_t m_total, int k_total, int bn, int act) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(base + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * bn;
    const int num_kb = (k_total + BK - 1) / BK;
    const uint32_t tmem_cols = bn <= 32 ? 32 : (bn <= 64 ? 64 : 128);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer ----
            const uint32_t tx = (uint32_t)(BM + bn) * BK * 4;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], tx);
                uint8_t* st = base + s * STAGE_BYTES;
                tma_load_2d(&tma_a, &full[s], st, kb * BK, (int)m0);
                tma_load_2d(&tma_w, &full[s], st + A_BYTES, kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer ----
            const uint32_t idesc = make_idesc_tf32(bn);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(base + s * STAGE_BYTES);
                const uint64_t da = make_desc_k_sw128(a_addr);
                const uint64_t dw = make_desc_k_sw128(a_addr + A_BYTES);
                #pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    // advance 8 tf32 = 32 bytes inside the swizzle row: +2 in the (>>4) address field
                    mma_tf32(tmem_acc, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                }
                tcgen05_commit(&empty[s]);                 // stage reusable once these MMAs retire
            }
            tcgen05_commit(acc_full);                      // accumulator complete
        }
    } else {                                               // ---- epilogue: warps 2..5 ----
        const int quad = warp & 3;                         // TMEM lanes [32*quad, 32*quad+32)
        const int64_t row = m0 + quad * 32 + lane;
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const bool row_ok = row < m_total;
        float* crow = c + row * ldc + n0;
        const float* rrow = residual ? residual + row * ldr + n0 : nullptr;
        for (int cb = 0; cb < bn; cb += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row_ok) {
                #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v[4];
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = __uint_as_float(r[q * 4 + e]);
                        if (bias) x += __ldg(bias + n0 + cb + q * 4 + e);
                        if (act == ACT_GELU) x = gelu_erf(x);
                        else if (act == ACT_SELU) x = selu(x);
                        v[e] = x;
                    }
                    if (rrow) {
                        const float4 rv = *reinterpret_cast<const float4*>(rrow + cb + q * 4);
                        v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
                    }
                    *reinterpret_cast<float4*>(crow + cb + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(tmem_cols));
    }
}

constexpr int STAGES = 3;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;

int make_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    // resolved through the runtime so that the library has no link-time dependency on libcuda
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CTO_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        CTO_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
        encode = reinterpret_cast<encode_fn>(fn);
    }
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides,
                                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%lld ld=%lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return 1;
    }
    return 0;
}

}  // namespace tc

bool gemm_tc_supported(const float* a, int64_t lda, const float* w, int64_t m, int n, int k, const float* c, int64_t ldc,
                       const float* residual, int64_t ldr) {
    if (m <= 0 || n < 16 || n % 16 != 0 || k < 8) return false;
    if (n > 16 && n % 64 != 0) return false;
    if (lda % 4 != 0 || k % 4 != 0 || ldc % 4 != 0 || (residual && ldr % 4 != 0)) return false;
    const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) |
                           reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(residual);
    return (bits & 15) == 0;
}

int launch_gemm_tc(const float* a, int64_t lda, const float* w, const float* bias, const float* residual, int64_t ldr,
                   float* c, int64_t ldc, int64_t m, int n, int k, int act, cudaStream_t s) {
    CTO_REQUIRE(gemm_tc_supported(a, lda, w, m, n, k, c, ldc, residual, ldr),
                "gemm_tc: unsupported shape/alignment m=%lld n=%d k=%d lda=%lld ldc=%lld", (long long)m, n, k,
                (long long)lda, (long long)ldc);
    const int bn = (n % 128 == 0) ? 128 : (n % 64 == 0 ? 64 : 16);
    CUtensorMap map_a, map_w;
    if (tc::make_map(&map_a, a, m, k, lda, tc::BM)) return 1;
    if (tc::make_map(&map_w, w, n, k, k, bn)) return 1;
    static bool attr_set = false;
    if (!attr_set) {
        CTO_CHECK(cudaFuncSetAttribute(tc::gemm_tf32_kernel<tc::STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       tc::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(ceil_div(m, tc::BM), n / bn);
    tc::gemm_tf32_kernel<tc::STAGES><<<grid, tc::THREADS, tc::SMEM_BYTES, s>>>(map_a, map_w, bias, residual, ldr, c, ldc,
                                                                              m, k, bn, act);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
