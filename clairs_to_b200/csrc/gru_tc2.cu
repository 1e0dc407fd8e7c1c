// GRU recurrence on a CTA PAIR (tcgen05 cta_group::2): two SMs cooperate on one 128-candidate MMA.
//
// Same arithmetic as gru_tc.cu (3xTF32, fp32 gate math, torch.nn.GRU semantics, clairs/model.py:412-417),
// different mapping: with cta_group::1 an M=64 MMA occupies the tensor core as long as an M=128 one, so the
// single-CTA kernel runs at half the tensor rate.  Here the pair issues M=128 MMAs: each CTA owns 64
// candidates (its h tiles are the A rows) and HALF of every W_hh block (48 of the 96 gate rows are the B
// rows it streams from L2), which also halves W traffic and W shared memory per SM.
//   * leader CTA (cluster rank 0): one elected thread issues tcgen05.mma.cta_group::2.kind::tf32 (M=128, N=96,
//     K=8), three per k-step; tcgen05.commit multicasts stage-free / accumulator-ready to both CTAs;
//   * both CTAs: TMA producer warp for their half of W (2SM TMA, completion on the leader's barrier),
//     eight gate-math warps on their own TMEM (row m, 16 units per block: accumulator column n < 48 lives in
//     lane m, n >= 48 in lane 64+m, hence the (r,z,n) x 16-unit row order of the W blocks);
//   * h_t stays in registers until all MMAs of the step retired, then goes back to shared memory as TF32
//     hi/lo tiles; the peer's gate-math threads arrive on the leader's barrier through DSMEM.
#include "gru_ptx.cuh"

namespace cto {

namespace tc {

constexpr int P_M = 64;                      // candidates per CTA
constexpr int P_BLK = 32;                    // hidden units per W block
constexpr int P_N = 3 * P_BLK;               // 96 gate columns per block (MMA N)
constexpr int P_HALF = P_N / 2;              // B rows / accumulator columns per CTA
constexpr int P_K = 32;
constexpr int P_HTILE = P_M * 128;           // 8 KB
constexpr int P_WTILE = P_HALF * 128;        // 6 KB
constexpr int P_STAGE = 2 * P_WTILE;         // hi | lo
constexpr int P_STAGES = 8;
constexpr int P_THREADS = 320;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

template <int H>
struct PairSmem {
    static constexpr int KB = H / P_K;
    static constexpr int TOTAL = 2 * KB * P_HTILE + P_STAGES * P_STAGE + 1024 + 256 + H * 4;
};

__device__ __forceinline__ void p_tma_load_2sm(const CUtensorMap* map, uint64_t* leader_bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(g_smem_u32(dst)), "l"(map), "r"(g_smem_u32(leader_bar) & PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p_mma_tf32_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void p_commit_2sm(uint64_t* bar) {       // arrives on the same barrier in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(g_smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void p_arrive_leader(uint64_t* bar) {   // DSMEM arrive on the leader CTA's barrier
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(g_smem_u32(bar)));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void p_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
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

template <int H>
__global__ void __launch_bounds__(P_THREADS, 1)
gru_pair_kernel(const __grid_constant__ CUtensorMap tma_whi, const __grid_constant__ CUtensorMap tma_wlo,
                const float* __restrict__ xproj, const float* __restrict__ bhn, float* __restrict__ out, int64_t batch) {
    constexpr int KB = H / P_K;
    constexpr int NB = H / P_BLK;
    constexpr int PAIRS = NB / 2;
    constexpr uint32_t TMEM_COLS = NB * P_HALF <= 256 ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* h_hi = base;
    uint8_t* h_lo = base + KB * P_HTILE;
    uint8_t* wring = base + 2 * KB * P_HTILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(wring + P_STAGES * P_STAGE);
    uint64_t* empty = full + P_STAGES;
    uint64_t* acc_full = empty + P_STAGES;
    uint64_t* h_ready = acc_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);
    float* s_bhn = reinterpret_cast<float*>(tmem_slot + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const uint32_t crank = g_cluster_rank();
    const bool leader = crank == 0;
    for (int i = threadIdx.x; i < H; i += P_THREADS) s_bhn[i] = bhn[dir * H + i];

    if (threadIdx.x == 0) {
        for (int s = 0; s < P_STAGES; ++s) { g_mbar_init(&full[s], 1); g_mbar_init(&empty[s], 1); }
        for (int p = 0; p < 3; ++p) g_mbar_init(&acc_full[p], 1);
        g_mbar_init(h_ready, 512);                           // gate-math threads of BOTH CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    g_cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {                                                  // ---- TMA producer: this CTA's half of every W block ----
            uint32_t it = 0;
            for (int step = 0; step < N_POS; ++step) {
                for (int blk = 0; blk < NB; ++blk) {
                    for (int kb = 0; kb < KB; ++kb, ++it) {
                        const int s = it % P_STAGES;
                        const uint32_t ph = (it / P_STAGES) & 1;
                        g_mbar_wait(&empty[s], ph ^ 1);
                        uint8_t* st = wring + s * P_STAGE;
                        const int row = dir * 3 * H + blk * P_N + (int)crank * P_HALF;
                        if (!g_elect_one()) continue;
                        if (leader) g_mbar_expect_tx(&full[s], 2 * P_STAGE);      // both halves land on the leader's barrier
                        p_tma_load_2sm(&tma_whi, &full[s], st, kb * P_K, row);
                        p_tma_load_2sm(&tma_wlo, &full[s], st + P_WTILE, kb * P_K, row);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {                                      // ---- MMA issuer (leader CTA only; warp-uniform loop) ----
            // D=f32, A=B=tf32, K-major, M=128 (64 rows per CTA), N=96
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t it = 0;
            for (int step = 0; step < N_POS; ++step) {
                p_wait_cluster(h_ready, step & 1);         // h_{t-1} is in BOTH shared memories, accumulators drained
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int blk = 0; blk < NB; ++blk) {
                    const uint32_t acc = tmem_base + (uint32_t)(blk * P_HALF);
                    for (int kb = 0; kb < KB; ++kb, ++it) {
                        const int s = it % P_STAGES;
                        const uint32_t ph = (it / P_STAGES) & 1;
                        g_mbar_wait(&full[s], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t d_hhi = g_desc_k_sw128(g_smem_u32(h_hi + kb * P_HTILE));
                        const uint64_t d_hlo = g_desc_k_sw128(g_smem_u32(h_lo + kb * P_HTILE));
                        const uint32_t w_addr = g_smem_u32(wring + s * P_STAGE);
                        const uint64_t d_whi = g_desc_k_sw128(w_addr);
                        const uint64_t d_wlo = g_desc_k_sw128(w_addr + P_WTILE);
                        if (g_elect_one()) {
                            #pragma unroll
                            for (int k = 0; k < P_K / 8; ++k) {
                                const uint64_t o = (uint64_t)(k * 2);
                                p_mma_tf32_2sm(acc, d_hhi + o, d_whi + o, idesc, (kb | k) ? 1u : 0u);
                                p_mma_tf32_2sm(acc, d_hlo + o, d_whi + o, idesc, 1u);
                                p_mma_tf32_2sm(acc, d_hhi + o, d_wlo + o, idesc, 1u);
                            }
                            p_commit_2sm(&empty[s]);                   // stage free in both CTAs
                            if ((blk & 1) && kb == KB - 1) p_commit_2sm(&acc_full[blk >> 1]);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {                                               // ---- gate math: warps 2..9 in both CTAs ----
        const int quad = warp & 3;                         // TMEM lanes [32*quad, +32)
        const int wsel = (warp - 2) >> 2;                  // this warp takes blocks wsel, wsel+2, wsel+4
        const int tl = quad * 32 + lane;                   // TMEM lane
        const int m = tl & 63;                             // candidate row inside the CTA
        const int uhalf = tl >> 6;                         // units [16*uhalf, +16) of each block
        const int64_t b_raw = ((int64_t)blockIdx.x) * P_M + m;
        const bool b_ok = b_raw < batch;
        const int64_t b = b_ok ? b_raw : batch - 1;
        const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
        for (int blk = wsel; blk < NB; blk += 2) {         // h_0 = 0
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t chunk = (uint32_t)(((uhalf * 4 + j) ^ (m & 7)) << 4);
                *reinterpret_cast<float4*>(h_hi + blk * P_HTILE + row_off + chunk) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(h_lo + blk * P_HTILE + row_off + chunk) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        p_arrive_leader(h_ready);

        constexpr int ITERS = PAIRS * 2;                   // (block of this warp) x (8-unit half)
        float4 xq[6];
        auto prefetch = [&](int step, int i) {
            const int t = dir ? (N_POS - 1 - step) : step;
            const int uu = (2 * (i >> 1) + wsel) * P_BLK + uhalf * 16 + (i & 1) * 8;
            const float* xp = xproj + (b * N_POS + t) * (int64_t)(6 * H) + dir * 3 * H + uu;
            xq[0] = *reinterpret_cast<const float4*>(xp);
            xq[1] = *reinterpret_cast<const float4*>(xp + 4);
            xq[2] = *reinterpret_cast<const float4*>(xp + H);
            xq[3] = *reinterpret_cast<const float4*>(xp + H + 4);
            xq[4] = *reinterpret_cast<const float4*>(xp + 2 * H);
            xq[5] = *reinterpret_cast<const float4*>(xp + 2 * H + 4);
        };
        prefetch(0, 0);
        for (int step = 0; step < N_POS; ++step) {
            const int t = dir ? (N_POS - 1 - step) : step;
            float* op = out + (b * N_POS + t) * (int64_t)(2 * H) + dir * H;
            if (step + 1 < N_POS) {                        // next step's xproj lines -> L2
                const int tn = dir ? (N_POS - 2 - step) : step + 1;
                const float* xn = xproj + (b * N_POS + tn) * (int64_t)(6 * H) + dir * 3 * H + wsel * P_BLK + uhalf * 16;
                #pragma unroll
                for (int p = 0; p < PAIRS; ++p)
                    #pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + 2 * p * P_BLK + g * H));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + 2 * p * P_BLK + g * H + 8));
                    }
            }
            float hk[PAIRS * 16];
            #pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                const int p = i >> 1;
                const int blk = 2 * p + wsel;
                const int uu = blk * P_BLK + uhalf * 16 + (i & 1) * 8;
                // block blk: accumulator columns [48*blk, +48) = r(16) z(16) n(16) of this lane's 16 units
                const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(blk * P_HALF + (i & 1) * 8);
                if ((i & 1) == 0) {
                    g_mbar_wait(&acc_full[p], step & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                uint32_t ar[8], az[8], an[8];
                g_tmem_ld8(tcol, ar);
                g_tmem_ld8(tcol + 16, az);
                g_tmem_ld8(tcol + 32, an);
                float xv[24];
                #pragma unroll
                for (int q = 0; q < 6; ++q) { xv[q * 4] = xq[q].x; xv[q * 4 + 1] = xq[q].y; xv[q * 4 + 2] = xq[q].z; xv[q * 4 + 3] = xq[q].w; }
                if (i + 1 < ITERS) prefetch(step, i + 1);
                else if (step + 1 < N_POS) prefetch(step + 1, 0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 bn = *reinterpret_cast<const float4*>(s_bhn + uu + q * 4);
                    const uint32_t chunk = (uint32_t)(((uhalf * 4 + (i & 1) * 2 + q) ^ (m & 7)) << 4);
                    const float4 ohi = *reinterpret_cast<const float4*>(h_hi + blk * P_HTILE + row_off + chunk);
                    const float4 olo = *reinterpret_cast<const float4*>(h_lo + blk * P_HTILE + row_off + chunk);
                    const float hp[4] = {ohi.x + olo.x, ohi.y + olo.y, ohi.z + olo.z, ohi.w + olo.w};
                    const float bnv[4] = {bn.x, bn.y, bn.z, bn.w};
                    float hn[4];
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = q * 4 + e;
                        const float r = g_sigmoid(xv[c] + __uint_as_float(ar[c]));
                        const float z = g_sigmoid(xv[8 + c] + __uint_as_float(az[c]));
                        const float n = g_tanh(xv[16 + c] + r * (__uint_as_float(an[c]) + bnv[e]));
                        hn[e] = (1.0f - z) * n + z * hp[e];
                        hk[i * 8 + c] = hn[e];
                    }
                    if (b_ok) *reinterpret_cast<float4*>(op + uu + q * 4) = make_float4(hn[0], hn[1], hn[2], hn[3]);
                }
            }
            // acc_full of the last pair => every MMA of the step (in both CTAs) has read h_{t-1}
            #pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                const int blk = 2 * (i >> 1) + wsel;
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t chunk = (uint32_t)(((uhalf * 4 + (i & 1) * 2 + q) ^ (m & 7)) << 4);
                    float hh[4], hl[4];
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float v = hk[i * 8 + q * 4 + e];
                        hh[e] = g_round_tf32(v);
                        hl[e] = v - hh[e];
                    }
                    *reinterpret_cast<float4*>(h_hi + blk * P_HTILE + row_off + chunk) = make_float4(hh[0], hh[1], hh[2], hh[3]);
                    *reinterpret_cast<float4*>(h_lo + blk * P_HTILE + row_off + chunk) = make_float4(hl[0], hl[1], hl[2], hl[3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            p_arrive_leader(h_ready);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    g_cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

template <int H>
static int launch_pair(const CUtensorMap& map_hi, const CUtensorMap& map_lo, const float* xproj, const float* bhn, float* out,
                       int64_t batch, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        CTO_CHECK(cudaFuncSetAttribute(gru_pair_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairSmem<H>::TOTAL));
        attr = true;
    }
    const int ctas = ceil_div(batch, P_M);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((ctas + 1) / 2 * 2), 2, 1);
    cfg.blockDim = dim3(P_THREADS, 1, 1);
    cfg.dynamicSmemBytes = PairSmem<H>::TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CTO_CHECK(cudaLaunchKernelEx(&cfg, gru_pair_kernel<H>, map_hi, map_lo, xproj, bhn, out, batch));
    return 0;
}

}  // namespace tc

// w_hi / w_lo: [2 directions][3H rows regrouped as (32-unit block, 16-unit half, gate, unit)][H] (TF32 hi / lo)
int launch_gru_pair(const float* xproj, const float* w_hi, const float* w_lo, const float* bhn, float* out, int64_t batch,
                    int hidden, cudaStream_t s) {
    if (batch <= 0) return 0;
    CTO_REQUIRE(hidden == 128 || hidden == 192, "gru_pair: hidden size %d not built", hidden);
    CUtensorMap map_hi, map_lo;
    if (tc::make_map(&map_hi, w_hi, 6 * hidden, hidden, hidden, tc::P_HALF)) return 1;
    if (tc::make_map(&map_lo, w_lo, 6 * hidden, hidden, hidden, tc::P_HALF)) return 1;
    const int rc = hidden == 128 ? tc::launch_pair<128>(map_hi, map_lo, xproj, bhn, out, batch, s)
                                 : tc::launch_pair<192>(map_hi, map_lo, xproj, bhn, out, batch, s);
    if (rc) return rc;
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
