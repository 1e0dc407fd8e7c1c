// Transposed input projection of a GRU layer on CTA pairs, with the weights resident in shared memory:
//
//     X^T[M, N] = A[M, K] * W[N, K]^T + bias[M]        M = 6H gate units (1152), K = layer input width (<= 256),
//                                                      N = 33 * bp candidate-positions (~1.25 M per engine chunk)
//
// (torch.nn.GRU's W_ih x_t + b_ih for every t at once, clairs/model.py:412-417; the recurrence kernels consume X^T.)
// Same arithmetic as gemm_tc.cu ("bf16x3": hi*hi + mid*hi + hi*mid per k-step, fp32 accumulation in TMEM, same order),
// different data movement.  What the phase counters of gemm_tc.cu said on this shape (profiles/r2_gemm_proj2_phase.txt):
// the epilogue is never waited for, the MMA warp waits for operands -- every 128 x 128 x 256 tile pulls 256 KB through
// L2 (both operands, hi + mid planes) in the ~3 100 cycles its 48 MMAs need, 83 B/clk/SM against the ~42 B/clk/SM the L2
// delivers when all SMs pull.  Two changes take that to 64 KB per tile:
//   * A = the layer's W_ih is tiny (9 row blocks): every CTA keeps the four k-blocks of ITS row block resident
//     (128 KB, loaded once per launch) and walks column tiles;
//   * two CTAs form a pair (tcgen05 cta_group::2, M = 256 = two row blocks): they share the column tile, each loads
//     HALF of it (64 candidate-positions x 64 k, hi + mid = 16 KB per stage) and the tensor cores of both SMs read
//     both halves.  Four 16 KB stages keep 48 KB in flight per SM.
// Row blocks: ceil(M / 256) pairs of row blocks; an odd count (1152 = 4.5 x 256) leaves the last pair's second CTA on rows
// beyond M: its A is zero-filled by TMA, its stores are clipped -- 10 % idle tensor work, accepted.
// Warp roles (320 threads): 0 TMA producer, 1 MMA issuer (leader CTA of the pair only), 2-9 epilogue (two groups of four
// warps, alternating 32-column slabs; bias per row; swizzled staging; one TMA store of a 32 x 32 fp32 box per warp).
#include "gru_ptx.cuh"

namespace cto {
extern long long* g_gemm_timing;
extern int g_gemm_debug;

namespace tc {

constexpr int P_BM = 128;                    // rows per CTA (256 per pair)
constexpr int P_BN = 128;                    // columns per tile
constexpr int P_BK = 64;
constexpr int P_MAXKB = 4;                   // K <= 256
constexpr int P_ATILE = P_BM * 128;          // 16 KB: 128 rows x 64 k bf16
constexpr int P_WHALF = (P_BN / 2) * 128;    // 8 KB: this CTA's 64 of the tile's 128 columns x 64 k bf16
constexpr int P_STAGE = 2 * P_WHALF;         // W_hi | W_mid halves
constexpr int P_STAGES = 4;
constexpr int P_SLAB = 32;
constexpr int P_STAGING = P_BM * P_SLAB * 4; // 16 KB per epilogue group
constexpr int P_THREADS = 320;
constexpr int P_SMEM = P_MAXKB * 2 * P_ATILE + P_STAGES * P_STAGE + 2 * P_STAGING + 1024 + 256;
constexpr uint32_t P_PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ void p_tma_load_2sm(const CUtensorMap* map, uint64_t* leader_bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(g_smem_u32(dst)), "l"(map), "r"(g_smem_u32(leader_bar) & P_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_mma_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void p_commit_2sm(uint64_t* bar) {       // arrives on the same barrier in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(g_smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void p_arrive_leader(uint64_t* bar) {    // default .release.cta semantics (no MEMBAR.GPU)
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(g_smem_u32(bar)));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void p_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "PW_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PW_DONE;\n\t"
        "bra PW_LOOP;\n\t"
        "PW_DONE:\n\t"
        "}" ::"r"(g_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void p_tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(g_smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_prefetch_map(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__global__ void __launch_bounds__(P_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_ahi, const __grid_constant__ CUtensorMap tma_amid,
                 const __grid_constant__ CUtensorMap tma_whi, const __grid_constant__ CUtensorMap tma_wmid,
                 const __grid_constant__ CUtensorMap tma_c, const float* __restrict__ bias, int64_t m_total, int64_t n_total,
                 int k_total, int row_pairs, long long* timing) {
    const bool tim = timing != nullptr && blockIdx.x == 0;
    long long tacc[4] = {0, 0, 0, 0};
    #define PTIC long long _t0 = tim ? clock64() : 0
    #define PTOC(i) do { if (tim) { long long _t1 = clock64(); tacc[i] += _t1 - _t0; _t0 = _t1; } } while (0)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (g_smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_res = base;                                           // [kb][A_hi | A_mid]
    uint8_t* ring = base + P_MAXKB * 2 * P_ATILE;                    // [stage][W_hi half | W_mid half]
    uint8_t* staging = ring + P_STAGES * P_STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(staging + 2 * P_STAGING);   // leader: both halves of a stage have landed
    uint64_t* empty = full + P_STAGES;                               // both CTAs: the stage's MMAs have retired
    uint64_t* acc_full = empty + P_STAGES;                           // [2] both CTAs
    uint64_t* acc_empty = acc_full + 2;                              // [2] leader: drained by the epilogue warps of BOTH CTAs
    uint64_t* a_full = acc_empty + 2;                                // this CTA's resident A rows have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = g_cluster_rank();
    const bool leader = crank == 0;
    const int pair = (int)(blockIdx.x >> 1), pairs = (int)(gridDim.x >> 1);
    const int num_kb = (k_total + P_BK - 1) / P_BK;
    const int64_t n_tiles = (n_total + P_BN - 1) / P_BN;
    const int m0 = ((pair % row_pairs) * 2 + (int)crank) * P_BM;     // this CTA's row block (may lie beyond M: zero A, clipped C)
    const int64_t t_begin = pair / row_pairs, t_step = pairs / row_pairs;

    if (threadIdx.x == 0) {
        for (int s = 0; s < P_STAGES; ++s) { g_mbar_init(&full[s], 1); g_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { g_mbar_init(&acc_full[b], 1); g_mbar_init(&acc_empty[b], 16); }   // 8 warps x 2 CTAs
        g_mbar_init(a_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        p_prefetch_map(&tma_ahi); p_prefetch_map(&tma_amid); p_prefetch_map(&tma_whi); p_prefetch_map(&tma_wmid); p_prefetch_map(&tma_c);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {                                                 // resident A: this CTA's row block, all k-blocks
        if (g_elect_one()) {
            g_mbar_expect_tx(a_full, (uint32_t)(num_kb * 2 * P_ATILE));
            for (int kb = 0; kb < num_kb; ++kb) {
                g_tma_load_2d(&tma_ahi, a_full, a_res + kb * 2 * P_ATILE, kb * P_BK, m0);
                g_tma_load_2d(&tma_amid, a_full, a_res + (kb * 2 + 1) * P_ATILE, kb * P_BK, m0);
            }
        }
        __syncwarp();
        g_mbar_wait(a_full, 0);
    }
    __syncthreads();
    g_cluster_sync();                                                // both CTAs: barriers initialised, TMEM allocated, A resident
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- W producer: this CTA's 64 columns of every tile, hi and mid planes, onto the LEADER's full barrier ----
        uint32_t it = 0;
        for (int64_t t = t_begin; t < n_tiles; t += t_step) {
            const int n0 = (int)(t * P_BN) + (int)crank * (P_BN / 2);
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % P_STAGES;
                PTIC;
                g_mbar_wait(&empty[s], ((it / P_STAGES) & 1) ^ 1);
                PTOC(0);
                uint8_t* st = ring + s * P_STAGE;
                if (g_elect_one()) {
                    if (leader) g_mbar_expect_tx(&full[s], 2 * P_STAGE);      // both CTAs' halves land on the leader's barrier
                    p_tma_load_2sm(&tma_whi, &full[s], st, kb * P_BK, n0);
                    p_tma_load_2sm(&tma_wmid, &full[s], st + P_WHALF, kb * P_BK, n0);
                }
                __syncwarp();
                PTOC(1);
            }
        }
        if (tim && lane == 0) { timing[0] = tacc[0]; timing[1] = tacc[1]; }
    } else if (warp == 1) {
        if (leader) {
            // ---- MMA issuer: D = f32, A = B = bf16, K-major, M = 256 (128 rows per CTA), N = 128 ----
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(P_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
            uint32_t it = 0, acc_it = 0;
            for (int64_t t = t_begin; t < n_tiles; t += t_step, ++acc_it) {
                const uint32_t ab = acc_it & 1, aph = (acc_it >> 1) & 1;
                PTIC;
                p_wait_cluster(&acc_empty[ab], aph ^ 1);             // drained in BOTH CTAs
                PTOC(0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + ab * P_BN;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % P_STAGES;
                    p_wait_cluster(&full[s], (it / P_STAGES) & 1);
                    PTOC(1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = g_smem_u32(a_res + kb * 2 * P_ATILE);
                    const uint32_t w_addr = g_smem_u32(ring + s * P_STAGE);
                    const uint64_t d_ahi = g_desc_k_sw128(a_addr), d_amid = g_desc_k_sw128(a_addr + P_ATILE);
                    const uint64_t d_whi = g_desc_k_sw128(w_addr), d_wmid = g_desc_k_sw128(w_addr + P_WHALF);
                    const int ksteps = min(P_BK / 16, (k_total - kb * P_BK + 15) >> 4);
                    if (g_elect_one()) {
                        #pragma unroll
                        for (int k = 0; k < P_BK / 16; ++k) {
                            if (k >= ksteps) break;
                            const uint64_t o = (uint64_t)(k * 2);
                            p_mma_2sm(acc, d_ahi + o, d_whi + o, idesc, (kb | k) ? 1u : 0u);
                            p_mma_2sm(acc, d_amid + o, d_whi + o, idesc, 1u);
                            p_mma_2sm(acc, d_ahi + o, d_wmid + o, idesc, 1u);
                        }
                        p_commit_2sm(&empty[s]);
                        if (kb == num_kb - 1) p_commit_2sm(&acc_full[ab]);
                    }
                    __syncwarp();
                    PTOC(2);
                }
            }
            if (tim && lane == 0) { timing[4] = tacc[0]; timing[5] = tacc[1]; timing[7] = tacc[2]; }
        }
    } else {
        // ---- epilogue: warps 2..9, two groups of four; group g takes the 32-column slabs with index % 2 == g ----
        const int quad = warp & 3;                         // TMEM lanes [32*quad, +32)
        const int grp = (warp - 2) >> 2;
        const int r_in_tile = quad * 32 + lane;
        uint8_t* stg = staging + grp * P_STAGING;
        const int64_t row = (int64_t)m0 + r_in_tile;
        const float rbias = (bias && row < m_total) ? __ldg(bias + row) : 0.0f;
        uint32_t acc_it = 0;
        for (int64_t t = t_begin; t < n_tiles; t += t_step, ++acc_it) {
            const int n0 = (int)(t * P_BN);
            const uint32_t ab = acc_it & 1, aph = (acc_it >> 1) & 1;
            PTIC;
            g_mbar_wait(&acc_full[ab], aph);
            PTOC(0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int cb = grp * P_SLAB; cb < P_BN; cb += 2 * P_SLAB) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ab * P_BN + ((uint32_t)(quad * 32) << 16) + (uint32_t)cb;
                g_tmem_ld16(taddr, r);
                g_tmem_ld16(taddr + 16, r + 16);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cb + 2 * P_SLAB >= P_BN) {             // this group's last slab of the tile: hand the accumulator back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) p_arrive_leader(&acc_empty[ab]);
                }
                PTOC(1);
                // this warp's previous TMA store must have finished reading its 32-row staging strip
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                uint8_t* wstg = stg + quad * (32 * 128);
                #pragma unroll
                for (int q = 0; q < 8; ++q) {              // 128-byte swizzle: 16-byte chunk q of row r lives at chunk (q ^ (r % 8))
                    const float4 o = make_float4(__uint_as_float(r[q * 4]) + rbias, __uint_as_float(r[q * 4 + 1]) + rbias,
                                                 __uint_as_float(r[q * 4 + 2]) + rbias, __uint_as_float(r[q * 4 + 3]) + rbias);
                    *reinterpret_cast<float4*>(stg + r_in_tile * 128 + ((q ^ (r_in_tile & 7)) << 4)) = o;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    p_tma_store_2d(&tma_c, wstg, n0 + cb, m0 + quad * 32);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                PTOC(2);
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (tim && warp == 2 && lane == 0) { timing[12] = tacc[0]; timing[13] = tacc[1]; timing[15] = tacc[2];
                                             timing[20] = (long long)((n_tiles - t_begin + t_step - 1) / t_step) * 148; timing[21] = num_kb; }
    }
    #undef PTIC
    #undef PTOC
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    g_cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

}  // namespace tc

bool gemm_pair_supported(const GemmTc& g) {
    const int need = GEMM_A_PRESPLIT | GEMM_BIAS_PER_ROW;
    if ((g.flags & need) != need || (g.flags & GEMM_OUT_SPLIT) || g.residual || g.act != ACT_NONE) return false;
    if (g.k < 8 || g.k % 8 != 0 || g.k > tc::P_MAXKB * tc::P_BK || g.m <= 0 || g.n < 64 || g.n % 4 != 0) return false;
    const int sm = device_sm_count();
    const int row_pairs = (int)ceil_div(g.m, 2 * tc::P_BM);
    const int pairs = sm / 2 / row_pairs * row_pairs;
    return pairs >= row_pairs && ceil_div(g.n, tc::P_BN) >= 4 * (pairs / row_pairs);       // enough column tiles per pair
}

// X^T = A W^T + bias[row] on CTA pairs with A resident (see the file comment).  Same operands as launch_gemm_tc_ex with
// GEMM_A_PRESPLIT | GEMM_BIAS_PER_ROW; call gemm_pair_supported first.
int launch_gemm_pair(const GemmTc& g, cudaStream_t s) {
    CTO_REQUIRE(gemm_pair_supported(g), "gemm_pair: unsupported problem m=%lld n=%d k=%d flags=%d", (long long)g.m, g.n, g.k, g.flags);
    const uintptr_t bits = reinterpret_cast<uintptr_t>(g.a_hi) | reinterpret_cast<uintptr_t>(g.a_mid) | reinterpret_cast<uintptr_t>(g.w_hi) |
                           reinterpret_cast<uintptr_t>(g.w_mid) | reinterpret_cast<uintptr_t>(g.c);
    CTO_REQUIRE(g.a_hi && g.a_mid && g.w_hi && g.w_mid && g.c && (bits & 15) == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0 && g.ldc % 4 == 0,
                "gemm_pair: bad buffers / strides");
    CUtensorMap map_ahi, map_amid, map_whi, map_wmid, map_c;
    if (tc::make_map_bf16(&map_ahi, g.a_hi, g.m, g.k, g.lda, tc::P_BM)) return 1;
    if (tc::make_map_bf16(&map_amid, g.a_mid, g.m, g.k, g.lda, tc::P_BM)) return 1;
    if (tc::make_map_bf16(&map_whi, g.w_hi, g.n, g.k, g.ldw, tc::P_BN / 2)) return 1;
    if (tc::make_map_bf16(&map_wmid, g.w_mid, g.n, g.k, g.ldw, tc::P_BN / 2)) return 1;
    if (tc::make_map(&map_c, g.c, g.m, g.n, g.ldc, 32)) return 1;
    CTO_CHECK(set_max_dynamic_smem(tc::gemm_pair_kernel, tc::P_SMEM));
    const int row_pairs = (int)ceil_div(g.m, 2 * tc::P_BM);
    const int pairs = device_sm_count() / 2 / row_pairs * row_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
    cfg.blockDim = dim3(tc::P_THREADS, 1, 1);
    cfg.dynamicSmemBytes = tc::P_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CTO_CHECK(cudaLaunchKernelEx(&cfg, tc::gemm_pair_kernel, map_ahi, map_amid, map_whi, map_wmid, map_c, g.bias, g.m, (int64_t)g.n, g.k,
                                 row_pairs, (g_gemm_debug && g_gemm_timing) ? g_gemm_timing : nullptr));
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
