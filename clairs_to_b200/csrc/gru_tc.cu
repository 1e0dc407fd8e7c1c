// GRU recurrence on tcgen05 tensor cores (sm_100a), fp32-grade accuracy via the 3xTF32 split.
//
// torch.nn.GRU semantics (clairs/model.py:412-417, 442-443), gate order r|z|n:
//     r = sigmoid(xr + W_hr h)   z = sigmoid(xz + W_hz h)   n = tanh(xn + r * (W_hn h + b_hn))
//     h' = (1 - z) * n + z * h
// xproj = W_ih x + b_ih (+ b_hh for r,z) comes from the projection GEMM (gemm_tc.cu).
//
// One CTA = 64 candidates x one direction for all 33 steps:
//   * h_{t-1} lives in shared memory as two K-major 128-byte-swizzled fp32 tiles (TF32 hi + lo), the
//     A operand of the MMAs (M = 64);
//   * W_hh (hi + lo, rows regrouped as blocks of 32 units x {r,z,n} = 96 rows) is streamed from L2
//     by a TMA producer warp through a 5-stage ring every step;
//   * one elected thread issues tcgen05.mma.kind::tf32 (M=64, N=96, K=8), three per k-step
//     (hi*hi + lo*hi + hi*lo); the accumulators of ALL gate columns of the step stay in TMEM:
//     unit block 2p sits in lanes 0-15 and block 2p+1 in lanes 16-31 of columns [96p, 96p+96);
//   * eight gate-math warps (one TMEM lane = one candidate x one unit block, two warps per lane
//     quadrant) read the gate pre-activations with tcgen05.ld, add the software-prefetched xproj,
//     apply the gate math in fp32, write h_t to HBM and the TF32 hi/lo split of h_t back into the
//     shared-memory tiles for the next step.
#include "gru_ptx.cuh"

namespace cto {

namespace tc {
constexpr int GM = 64;                       // candidates per CTA (MMA M)
constexpr int GBLK = 32;                     // hidden units per block
constexpr int GN = 3 * GBLK;                 // 96 gate columns per block (MMA N)
constexpr int GK = 32;                       // fp32 per swizzle row
constexpr int G_HTILE = GM * 128;            // 8 KB: 64 rows x 128 bytes
constexpr int G_WTILE = GN * 128;            // 12 KB
constexpr int G_STAGE = 2 * G_WTILE;         // hi | lo
constexpr int G_STAGES = 5;
constexpr int G_THREADS = 320;               // warp 0 TMA, warp 1 MMA, warps 2-9 gate math

template <int H>
struct GruSmem {
    static constexpr int KB = H / GK;        // k-blocks (and also unit blocks)
    static constexpr int H_BYTES = 2 * KB * G_HTILE;
    static constexpr int TOTAL = H_BYTES + G_STAGES * G_STAGE + 1024 + 256 + H * 4;
};

template <int H, int C>
__global__ void __launch_bounds__(G_THREADS, 1)
gru_tc_kernel(const __grid_constant__ CUtensorMap tma_whi, const __grid_constant__ CUtensorMap tma_wlo,
              const float* __restrict__ xproj, const float* __restrict__ bhn, float* __restrict__ out, int64_t batch,
              long long* timing) {
    const bool tim = timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    #define GTIC long long _t0 = tim ? clock64() : 0
    #define GTOC(i) do { if (tim) { long long _t1 = clock64(); tacc[i] += _t1 - _t0; _t0 = _t1; } } while (0)
    constexpr int KB = H / GK;               // 4 (H=128) or 6 (H=192)
    constexpr int NB = H / GBLK;             // unit blocks, == KB
    constexpr int PAIRS = NB / 2;
    constexpr uint32_t TMEM_COLS = PAIRS * GN <= 256 ? 256 : 512;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* h_hi = base;                                    // KB tiles of 8 KB
    uint8_t* h_lo = base + KB * G_HTILE;
    uint8_t* wring = base + 2 * KB * G_HTILE;
    uint64_t* full = reinterpret_cast<uint64_t*>(wring + G_STAGES * G_STAGE);
    uint64_t* empty = full + G_STAGES;
    uint64_t* acc_full = empty + G_STAGES;                 // one per unit-block pair
    uint64_t* h_ready = acc_full + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 1);
    float* s_bhn = reinterpret_cast<float*>(tmem_slot + 4);      // b_hn of this direction (L1 is ~empty: smem is full)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    for (int i = threadIdx.x; i < H; i += G_THREADS) s_bhn[i] = bhn[dir * H + i];
    // cluster of C CTAs (same direction, different candidates): every W_hh stage is fetched from L2 once per
    // cluster - each CTA loads 1/C of the rows and multicasts them into all C shared memories
    const uint32_t crank = C > 1 ? g_cluster_rank() : 0;
    constexpr uint16_t CMASK = (uint16_t)((1u << C) - 1);
    constexpr int SLICE_ROWS = GN / C;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_STAGES; ++s) { g_mbar_init(&full[s], 1); g_mbar_init(&empty[s], C); }
        for (int p = 0; p < 3; ++p) g_mbar_init(&acc_full[p], 1);
        g_mbar_init(h_ready, 256);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (C > 1) g_cluster_sync();                           // every CTA's barriers exist before any remote arrive
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        {                                                  // ---- TMA producer: W_hh blocks, every step (warp-uniform) ----
            uint32_t it = 0;
            for (int step = 0; step < N_POS; ++step) {
                for (int blk = 0; blk < NB; ++blk) {
                    for (int kb = 0; kb < KB; ++kb, ++it) {
                        const int s = it % G_STAGES;
                        const uint32_t ph = (it / G_STAGES) & 1;
                        g_mbar_wait(&empty[s], ph ^ 1);
                        uint8_t* st = wring + s * G_STAGE;
                        const int row = dir * 3 * H + blk * GN;
                        if (!g_elect_one()) continue;
                        g_mbar_expect_tx(&full[s], G_STAGE);
                        if (C == 1) {
                            g_tma_load_2d(&tma_whi, &full[s], st, kb * GK, row);
                            g_tma_load_2d(&tma_wlo, &full[s], st + G_WTILE, kb * GK, row);
                        } else {
                            const int off = (int)crank * SLICE_ROWS;
                            g_tma_load_2d_mc(&tma_whi, &full[s], st + off * 128, kb * GK, row + off, CMASK);
                            g_tma_load_2d_mc(&tma_wlo, &full[s], st + G_WTILE + off * 128, kb * GK, row + off, CMASK);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        {                                                  // ---- MMA issuer: whole warp runs the loop, one elected lane issues ----
            // D=f32, A=B=tf32, K-major, M=64, N=96
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GN >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
            uint32_t it = 0;
            for (int step = 0; step < N_POS; ++step) {
                GTIC;
                g_mbar_wait(h_ready, step & 1);            // h_{t-1} (hi/lo) is in smem, accumulators drained
                GTOC(0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int blk = 0; blk < NB; ++blk) {
                    const uint32_t acc = tmem_base + ((uint32_t)((blk & 1) * 16) << 16) + (uint32_t)((blk >> 1) * GN);
                    for (int kb = 0; kb < KB; ++kb, ++it) {
                        const int s = it % G_STAGES;
                        const uint32_t ph = (it / G_STAGES) & 1;
                        g_mbar_wait(&full[s], ph);
                        GTOC(1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t d_hhi = g_desc_k_sw128(g_smem_u32(h_hi + kb * G_HTILE));
                        const uint64_t d_hlo = g_desc_k_sw128(g_smem_u32(h_lo + kb * G_HTILE));
                        const uint32_t w_addr = g_smem_u32(wring + s * G_STAGE);
                        const uint64_t d_whi = g_desc_k_sw128(w_addr);
                        const uint64_t d_wlo = g_desc_k_sw128(w_addr + G_WTILE);
                        if (g_elect_one()) {
                            #pragma unroll
                            for (int k = 0; k < GK / 8; ++k) {
                                const uint64_t o = (uint64_t)(k * 2);
                                g_mma_tf32(acc, d_hhi + o, d_whi + o, idesc, (kb | k) ? 1u : 0u);
                                g_mma_tf32(acc, d_hlo + o, d_whi + o, idesc, 1u);
                                g_mma_tf32(acc, d_hhi + o, d_wlo + o, idesc, 1u);
                            }
                            if (C == 1) g_commit(&empty[s]);
                            else g_commit_mc(&empty[s], CMASK);     // frees the stage in every CTA of the cluster
                            if ((blk & 1) && kb == KB - 1) g_commit(&acc_full[blk >> 1]);   // the pair is accumulated
                        }
                        __syncwarp();
                        GTOC(2);
                    }
                }
            }
            if (tim && lane == 0) { timing[0] = tacc[0]; timing[1] = tacc[1]; timing[2] = tacc[2]; }
        }
    } else {                                               // ---- gate math: warps 2..9 ----
        // two warps per TMEM lane quadrant; thread = (candidate row m, unit block parity `sub`, 16-unit half `part`)
        const int quad = warp & 3;
        const int part = (warp - 2) >> 2;
        const int m = quad * 16 + (lane & 15);             // candidate row inside the CTA
        const int sub = lane >> 4;                         // which unit block of each pair
        const int64_t b_raw = (int64_t)blockIdx.x * GM + m;
        const bool b_ok = b_raw < batch;
        const int64_t b = b_ok ? b_raw : batch - 1;
        const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
        // h_0 = 0: this thread owns chunks [4*part, 4*part+4) of row m in the tiles of its blocks
        for (int kb = sub; kb < KB; kb += 2) {
            #pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t chunk = (uint32_t)(((part * 4 + j) ^ (m & 7)) << 4);
                *reinterpret_cast<float4*>(h_hi + kb * G_HTILE + row_off + chunk) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(h_lo + kb * G_HTILE + row_off + chunk) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        g_mbar_arrive(h_ready);

        const float* bhn_d = s_bhn;
        constexpr int ITERS = PAIRS * 2;                   // 8 units per iteration
        float4 xq[6];                                      // prefetched xproj: r(2) z(2) n(2) float4
        auto prefetch = [&](int step, int i) {
            const int t = dir ? (N_POS - 1 - step) : step;
            const int uu = (2 * (i >> 1) + sub) * GBLK + part * 16 + (i & 1) * 8;
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
            // pull the xproj lines of the NEXT step into L2 now (the register prefetch is one iteration deep)
            if (step + 1 < N_POS) {
                const int tn = dir ? (N_POS - 2 - step) : step + 1;
                const float* xn = xproj + (b * N_POS + tn) * (int64_t)(6 * H) + dir * 3 * H + sub * GBLK + part * 16;
                #pragma unroll
                for (int p = 0; p < PAIRS; ++p)
                    #pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + 2 * p * GBLK + g * H));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(xn + 2 * p * GBLK + g * H + 8));
                    }
            }
            // h_t of this thread's units stays in registers until every MMA of the step has read h_{t-1}:
            // the gate math of unit-block pair p overlaps the MMAs of the pairs behind it
            float hk[PAIRS * 16];
            GTIC;
            #pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                const int p = i >> 1;
                const int blk = 2 * p + sub;
                const int uu = blk * GBLK + part * 16 + (i & 1) * 8;       // first of this iteration's 8 units
                const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(p * GN + part * 16 + (i & 1) * 8);
                if ((i & 1) == 0) {
                    g_mbar_wait(&acc_full[p], step & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (i == 0) GTOC(3);
                }
                uint32_t ar[8], az[8], an[8];
                g_tmem_ld8(tcol, ar);
                g_tmem_ld8(tcol + GBLK, az);
                g_tmem_ld8(tcol + 2 * GBLK, an);
                float xv[24];
                #pragma unroll
                for (int q = 0; q < 6; ++q) { xv[q * 4] = xq[q].x; xv[q * 4 + 1] = xq[q].y; xv[q * 4 + 2] = xq[q].z; xv[q * 4 + 3] = xq[q].w; }
                // next iteration's xproj (possibly of the next step) flies while this one computes
                if (i + 1 < ITERS) prefetch(step, i + 1);
                else if (step + 1 < N_POS) prefetch(step + 1, 0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 bn = *reinterpret_cast<const float4*>(bhn_d + uu + q * 4);
                    const uint32_t chunk = (uint32_t)(((part * 4 + (i & 1) * 2 + q) ^ (m & 7)) << 4);
                    const float4 ohi = *reinterpret_cast<const float4*>(h_hi + blk * G_HTILE + row_off + chunk);
                    const float4 olo = *reinterpret_cast<const float4*>(h_lo + blk * G_HTILE + row_off + chunk);
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
            GTOC(4);
            // the last pair's barrier implies all MMAs of this step retired: h_{t-1} may now be overwritten
            #pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                const int blk = 2 * (i >> 1) + sub;
                #pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const uint32_t chunk = (uint32_t)(((part * 4 + (i & 1) * 2 + q) ^ (m & 7)) << 4);
                    float hh[4], hl[4];
                    #pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float v = hk[i * 8 + q * 4 + e];
                        hh[e] = g_round_tf32(v);
                        hl[e] = v - hh[e];
                    }
                    *reinterpret_cast<float4*>(h_hi + blk * G_HTILE + row_off + chunk) = make_float4(hh[0], hh[1], hh[2], hh[3]);
                    *reinterpret_cast<float4*>(h_lo + blk * G_HTILE + row_off + chunk) = make_float4(hl[0], hl[1], hl[2], hl[3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            g_mbar_arrive(h_ready);
            GTOC(5);
        }
        if (tim && threadIdx.x == 64) { timing[3] = tacc[3]; timing[4] = tacc[4]; timing[5] = tacc[5]; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (C > 1) g_cluster_sync();                           // nobody exits while peers may still multicast to it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

}  // namespace tc

extern int g_gemm_debug;
long long* g_gru_timing = nullptr;   // device buffer [8] (debug): per-phase cycles of CTA (0,0)

template <int H, int C>
static int launch_gru_tc_t(const CUtensorMap& map_hi, const CUtensorMap& map_lo, const float* xproj, const float* bhn,
                           float* out, int64_t batch, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        CTO_CHECK(cudaFuncSetAttribute(tc::gru_tc_kernel<H, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       tc::GruSmem<H>::TOTAL));
        attr = true;
    }
    const int ctas = ceil_div(batch, tc::GM);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((ctas + C - 1) / C * C), 2, 1);
    cfg.blockDim = dim3(tc::G_THREADS, 1, 1);
    cfg.dynamicSmemBytes = tc::GruSmem<H>::TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CTO_CHECK(cudaLaunchKernelEx(&cfg, tc::gru_tc_kernel<H, C>, map_hi, map_lo, xproj, bhn, out, batch, g_gru_timing));
    return 0;
}

int g_gru_cluster = 2;      // CTAs per cluster sharing one W_hh stream (1, 2 or 4)

// w_hi / w_lo: [2 directions][3H rows regrouped as (unit block, gate, unit)][H] fp32 (TF32 hi / lo)
int launch_gru_tc(const float* xproj, const float* w_hi, const float* w_lo, const float* bhn, float* out, int64_t batch,
                  int hidden, cudaStream_t s) {
    if (batch <= 0) return 0;
    CTO_REQUIRE(hidden == 128 || hidden == 192, "gru_tc: hidden size %d not built (128 and 192 are, clairs/model.py:403-404)",
                hidden);
    const int c = g_gru_cluster;
    CUtensorMap map_hi, map_lo;
    if (tc::make_map(&map_hi, w_hi, 6 * hidden, hidden, hidden, tc::GN / c)) return 1;
    if (tc::make_map(&map_lo, w_lo, 6 * hidden, hidden, hidden, tc::GN / c)) return 1;
    int rc;
    if (hidden == 128) {
        rc = c == 4 ? launch_gru_tc_t<128, 4>(map_hi, map_lo, xproj, bhn, out, batch, s)
           : c == 2 ? launch_gru_tc_t<128, 2>(map_hi, map_lo, xproj, bhn, out, batch, s)
                    : launch_gru_tc_t<128, 1>(map_hi, map_lo, xproj, bhn, out, batch, s);
    } else {
        rc = c == 4 ? launch_gru_tc_t<192, 4>(map_hi, map_lo, xproj, bhn, out, batch, s)
           : c == 2 ? launch_gru_tc_t<192, 2>(map_hi, map_lo, xproj, bhn, out, batch, s)
                    : launch_gru_tc_t<192, 1>(map_hi, map_lo, xproj, bhn, out, batch, s);
    }
    if (rc) return rc;
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
