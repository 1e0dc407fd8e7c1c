// GRU recurrence, third generation: CTA pair (tcgen05 cta_group::2) + bf16x3 operands, register-resident state.
//
// Arithmetic: torch.nn.GRU semantics (clairs/model.py:412-417), gate order r|z|n,
//     n = tanh(W_in x + b_in + r * (W_hn h + b_hn)),   h' = (1 - z) * n + z * h,
// with the recurrent product h * W_hh^T on the tensor cores as three bf16 MMAs per k-step
// (h_hi*W_hi + h_mid*W_hi + h_hi*W_mid, fp32 accumulation in TMEM; hi = bf16(x), mid = bf16(x - hi)) and every
// non-linearity in fp32.  The fp32 state h itself never leaves the registers of the thread that owns it.
//
// What the ncu capture of the previous kernel (gru_tc2.cu, profiles/r1_ncu_gru_pair_stalls.txt) showed and what
// changed here:
//   * the gate-math warps were latency-bound (17 % issue utilisation): every LDG.128 of the row-major input projection
//     touched 32 lines -> transposed projection, staged through shared memory by a TMA producer warp (conflict-free LDS)
//     (more gate-math warps were measured too: 12 / 16 warps are slower than 8);
//   * every gate thread re-read its old h from shared memory and kept the new h in registers until all MMAs of
//     the step had retired -> h tiles are double buffered (bf16 halves the tile size), h_t is written at once,
//     h_{t-1} comes from registers;
//   * 512 cluster-scope release arrives per step (MEMBAR + ERRBAR each) -> one arrive per warp (a CTA-local barrier
//     relayed to the leader by an idle warp was measured too: no gain);
//   * W_hh as TF32 hi/lo streamed 885 KB per step and pair from L2 -> bf16 hi/mid halves that and the
//     shared-memory operand traffic of the MMAs (the M=64-per-CTA MMA is smem-bandwidth bound).
//
// Mapping (unchanged from gru_tc2.cu): each CTA owns 64 candidates (its h tiles are the A rows of the M=128 pair
// MMA) and streams HALF of every 96-row W block (48 rows: r,z,n x 16 units) through a TMA ring; accumulator
// column n < 48 of a block lives in TMEM lane m, n >= 48 in lane 64+m.
#include "gru_ptx.cuh"

namespace cto {

namespace tc {

constexpr int Q_M = 64;                      // candidates per CTA
constexpr int Q_BLK = 32;                    // hidden units per W block
constexpr int Q_N = 3 * Q_BLK;               // 96 gate columns per block (MMA N)
constexpr int Q_HALF = Q_N / 2;              // B rows / accumulator columns per CTA
constexpr int Q_K = 64;                      // bf16 elements per 128-byte swizzle row
constexpr int Q_HTILE = Q_M * 128;           // 8 KB: 64 rows x 64 k (bf16)
constexpr int Q_WTILE = Q_HALF * 128;        // 6 KB
constexpr int Q_STAGE = 2 * Q_WTILE;         // hi | mid
constexpr int Q_STAGES = 6;
constexpr int Q_XSTAGES = 2;                 // input-projection ring: one stage = 3 gates x 32 units x 64 candidates fp32
constexpr int Q_XSTAGE = 3 * Q_BLK * Q_M * 4; // 24 KB
// GW = gate-math warps per TMEM lane quadrant; the block is warp 0 TMA, warp 1 MMA, 4*GW gate-math warps
}  // namespace tc
extern long long* g_gemm_timing;      // cto_debug_timing(): device buffer [32]
extern int g_gemm_debug;
namespace tc {
constexpr uint32_t Q_PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> leader CTA

template <int H>
struct Gru3Smem {
    static constexpr int KB = H / Q_K;
    static constexpr int HBUF = 2 * KB * Q_HTILE;             // hi | mid tiles of one h_t
    static constexpr int TOTAL = 2 * HBUF + Q_STAGES * Q_STAGE + Q_XSTAGES * Q_XSTAGE + 1024 + 256 + H * 4;
};

__device__ __forceinline__ void q_tma_load_2sm(const CUtensorMap* map, uint64_t* leader_bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(g_smem_u32(dst)), "l"(map), "r"(g_smem_u32(leader_bar) & Q_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void q_mma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void q_commit_2sm(uint64_t* bar) {       // arrives on the same barrier in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(g_smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void q_arrive_leader(uint64_t* bar) {   // DSMEM arrive on the leader CTA's barrier
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(g_smem_u32(bar)));
    // default semantics (.release.cta), as CUTLASS signals a peer CTA: the .release.cluster form compiles to MEMBAR.ALL.GPU,
    // which stalls the warp until every output store it has in flight is acknowledged by L2
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void q_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "QW_LOOP:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra QW_DONE;\n\t"
        "bra QW_LOOP;\n\t"
        "QW_DONE:\n\t"
        "}" ::"r"(g_smem_u32(bar)), "r"(parity) : "memory");
}
// two fp32 -> packed bf16x2 (round to nearest even); `lo` lands in the low half = the lower address
__device__ __forceinline__ uint32_t q_pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void q_split2(float x0, float x1, uint32_t& hi, uint32_t& mid) {
    hi = q_pack_bf16x2(x0, x1);
    mid = q_pack_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xFFFF0000u));
}

template <int H, int GW, int VAR>
__global__ void __launch_bounds__(96 + 128 * GW, 1)
gru3_kernel(const __grid_constant__ CUtensorMap tma_whi, const __grid_constant__ CUtensorMap tma_wmid,
            const __grid_constant__ CUtensorMap tma_x, int64_t bp, const float* __restrict__ bhn,
            uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_mid, int64_t osb, int64_t ost, int64_t batch,
            long long* timing) {
    // per-phase clock64() counters of cluster 0, direction 0 (profiles/phase_timing_gru.py); timing == nullptr in production
    const bool tim = timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    #define GTIC long long _t0 = tim ? clock64() : 0
    #define GTOC(i) do { if (tim) { long long _t1 = clock64(); tacc[i] += _t1 - _t0; _t0 = _t1; } } while (0)
    constexpr int KB = H / Q_K;
    constexpr int NB = H / Q_BLK;
    constexpr int PAIRS = NB / 2;
    constexpr int Q_GW = GW;
    constexpr int Q_THREADS = 96 + 128 * GW;              // warp 0 W producer, warp 1 MMA, 4*GW gate warps, last warp x producer
    constexpr int ITERS = 2 * NB / Q_GW;                      // 8-unit half-blocks per gate thread
    constexpr int HBUF = Gru3Smem<H>::HBUF;
    constexpr uint32_t TMEM_COLS = NB * Q_HALF <= 256 ? 256 : 512;
    static_assert(H % Q_K == 0 && (2 * NB) % Q_GW == 0 && NB % 2 == 0, "hidden size not supported by the gate-math mapping");
    extern __shared__ uint8_t smem_raw[];
    // pointer arithmetic (not uintptr_t rounding) keeps the shared address space: LDS / STS instead of generic LD / ST
    uint8_t* base = smem_raw + ((1024u - (g_smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* hbuf = base;                                     // [2 buffers][hi | mid][KB][64 x 128 B]
    uint8_t* wring = base + 2 * HBUF;
    float* xring = reinterpret_cast<float*>(wring + Q_STAGES * Q_STAGE);
    uint64_t* full = reinterpret_cast<uint64_t*>(wring + Q_STAGES * Q_STAGE + Q_XSTAGES * Q_XSTAGE);
    uint64_t* empty = full + Q_STAGES;
    uint64_t* xfull = empty + Q_STAGES;
    uint64_t* xempty = xfull + Q_XSTAGES;
    uint64_t* acc_full = xempty + Q_XSTAGES;
    uint64_t* h_ready = acc_full + 6;                     // leader only: h_t complete in both CTAs, accumulators drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 2);
    float* s_bhn = reinterpret_cast<float*>(tmem_slot + 4);   // barriers take 192 + 16 bytes: keep the float4 reads of s_bhn 16-byte aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const uint32_t crank = g_cluster_rank();
    const bool leader = crank == 0;
    for (int i = threadIdx.x; i < H; i += Q_THREADS) s_bhn[i] = bhn[dir * H + i];

    if (threadIdx.x == 0) {
        for (int s = 0; s < Q_STAGES; ++s) { g_mbar_init(&full[s], 1); g_mbar_init(&empty[s], 1); }
        for (int p = 0; p < 6; ++p) g_mbar_init(&acc_full[p], 1);     // one per 32-unit block (round 1: one per block pair)
        for (int s = 0; s < Q_XSTAGES; ++s) { g_mbar_init(&xfull[s], 1); g_mbar_init(&xempty[s], 4 * Q_GW); }
        g_mbar_init(h_ready, 2 * 4 * Q_GW);                  // one arrive per gate-math warp of BOTH CTAs
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
                        const int s = it % Q_STAGES;
                        const uint32_t ph = (it / Q_STAGES) & 1;
                        g_mbar_wait(&empty[s], ph ^ 1);
                        uint8_t* st = wring + s * Q_STAGE;
                        const int row = dir * 3 * H + blk * Q_N + (int)crank * Q_HALF;
                        if (!g_elect_one()) continue;
                        if (leader) g_mbar_expect_tx(&full[s], 2 * Q_STAGE);      // both halves land on the leader's barrier
                        q_tma_load_2sm(&tma_whi, &full[s], st, kb * Q_K, row);
                        q_tma_load_2sm(&tma_wmid, &full[s], st + Q_WTILE, kb * Q_K, row);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {                                      // ---- MMA issuer (leader CTA only; warp-uniform loop) ----
            // D=f32, A=B=bf16, K-major, M=128 (64 rows per CTA), N=96
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Q_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t it = 0;
            for (int step = 0; step < N_POS; ++step) {
                GTIC;
                q_wait_cluster(h_ready, step & 1);         // h_{t-1} is in BOTH shared memories, accumulators drained
                GTOC(0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint8_t* hb = hbuf + (step & 1) * HBUF;
                for (int blk = 0; blk < NB; ++blk) {
                    const uint32_t acc = tmem_base + (uint32_t)(blk * Q_HALF);
                    for (int kb = 0; kb < KB; ++kb, ++it) {
                        const int s = it % Q_STAGES;
                        const uint32_t ph = (it / Q_STAGES) & 1;
                        g_mbar_wait(&full[s], ph);
                        GTOC(1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t d_hhi = g_desc_k_sw128(g_smem_u32(hb + kb * Q_HTILE));
                        const uint64_t d_hmid = g_desc_k_sw128(g_smem_u32(hb + (KB + kb) * Q_HTILE));
                        const uint32_t w_addr = g_smem_u32(wring + s * Q_STAGE);
                        const uint64_t d_whi = g_desc_k_sw128(w_addr);
                        const uint64_t d_wmid = g_desc_k_sw128(w_addr + Q_WTILE);
                        if (g_elect_one()) {
                            #pragma unroll
                            for (int k = 0; k < Q_K / 16; ++k) {
                                const uint64_t o = (uint64_t)(k * 2);          // 16 bf16 = 32 bytes along the swizzle row
                                q_mma_bf16_2sm(acc, d_hhi + o, d_whi + o, idesc, (kb | k) ? 1u : 0u);
                                q_mma_bf16_2sm(acc, d_hmid + o, d_whi + o, idesc, 1u);
                                q_mma_bf16_2sm(acc, d_hhi + o, d_wmid + o, idesc, 1u);
                            }
                            q_commit_2sm(&empty[s]);                   // stage free in both CTAs
                            if (kb == KB - 1) q_commit_2sm(&acc_full[blk]);   // the gate math of this block can start
                        }
                        __syncwarp();
                        GTOC(2);
                    }
                }
            }
            if (tim && lane == 0) { timing[0] = tacc[0]; timing[1] = tacc[1]; timing[2] = tacc[2]; }
        }
    } else if (warp == 2 + 4 * Q_GW) {
        // ---- x producer: the input projection xproj^T[(dir, gate, unit)][t * bp + b] of this CTA's 64 candidates, one
        // 32-unit block (= one half-block per gate warp) per stage, by TMA, up to two half-blocks ahead of the gate math
        // and across step boundaries.  (Per-thread global loads with one half-block of register look-ahead cost a
        // third of the kernel in long-scoreboard stalls; deeper register look-ahead spilled.  An extra L2 tensor
        // prefetch one step ahead was measured too: no gain, the gate warps wait < 4 % of a step for this ring.)
        static_assert(Q_GW == 2, "one x stage = block j: the half-block order of two gate warps per quadrant");
        const int b0 = (int)blockIdx.x * Q_M;
        uint32_t q = 0;
        for (int step = 0; step < N_POS; ++step) {
            const int t = dir ? (N_POS - 1 - step) : step;
            for (int j = 0; j < NB; ++j, ++q) {
                const int s = q % Q_XSTAGES;
                g_mbar_wait(&xempty[s], ((q / Q_XSTAGES) & 1) ^ 1);
                if (g_elect_one()) {
                    g_mbar_expect_tx(&xfull[s], Q_XSTAGE);
                    #pragma unroll
                    for (int g = 0; g < 3; ++g)
                        g_tma_load_2d(&tma_x, &xfull[s], xring + (s * 3 + g) * (Q_BLK * Q_M), (int)(t * bp) + b0,
                                      dir * 3 * H + g * H + j * Q_BLK);
                }
                __syncwarp();
            }
        }
    } else {                                               // ---- gate math: warps 2..9 in both CTAs ----
        const int quad = warp & 3;                         // TMEM lanes [32*quad, +32)
        const int wsel = (warp - 2) >> 2;                  // 0..GW-1: half-blocks wsel, wsel+GW, wsel+2*GW, ...
        const int tl = quad * 32 + lane;                   // TMEM lane
        const int m = tl & 63;                             // candidate row inside the CTA
        const int uhalf = tl >> 6;                         // units [16*uhalf, +16) of each block
        const int64_t b = ((int64_t)blockIdx.x) * Q_M + m;   // < bp: the transposed projection has columns for padded rows
        const bool b_ok = b < batch;
        const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
        // half-block j of this thread: block blk_j, 8 units starting at uu_j; its bf16 values are 16-byte chunk
        // ((uu % 64) / 8) of row m in k-block uu / 64
        auto unit0 = [&](int j) { const int hb = wsel + Q_GW * j; return (hb >> 1) * Q_BLK + uhalf * 16 + (hb & 1) * 8; };
        auto tile_off = [&](int j) {
            const int uu = unit0(j);
            return (uint32_t)((uu / Q_K) * Q_HTILE) + row_off + (uint32_t)(((((uu % Q_K) >> 3)) ^ (m & 7)) << 4);
        };
        float hk[ITERS * 8];                               // fp32 state of this thread's units, lives in registers
        #pragma unroll
        for (int j = 0; j < ITERS; ++j) {                  // h_0 = 0
            #pragma unroll
            for (int c = 0; c < 8; ++c) hk[j * 8 + c] = 0.0f;
            *reinterpret_cast<uint4*>(hbuf + tile_off(j)) = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(hbuf + KB * Q_HTILE + tile_off(j)) = make_uint4(0u, 0u, 0u, 0u);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) q_arrive_leader(h_ready);

        // The input projection is stored TRANSPOSED, xproj[gate unit][t * bp + b], and staged by the x producer as
        // [gate][unit of the block][candidate]: lanes are consecutive candidates, so the 24 reads below are conflict-free
        // LDS (the row-major layout of round 1 cost 32 L1 lines per load instruction).
        uint32_t xq_it = 0;
        for (int step = 0; step < N_POS; ++step) {
            const int t = dir ? (N_POS - 1 - step) : step;
            const int64_t orow = (b * osb + t * ost) * (int64_t)(2 * H) + dir * H;
            uint8_t* hnext = hbuf + ((step + 1) & 1) * HBUF;   // read by the MMAs of step + 1; those of step - 1 have retired
            #pragma unroll
            for (int j = 0; j < ITERS; ++j) {
                const int hb = wsel + Q_GW * j;
                const int blk = hb >> 1;
                const int uu = unit0(j);
                // block blk: accumulator columns [48*blk, +48) = r(16) z(16) n(16) of this lane's 16 units
                const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(blk * Q_HALF + (hb & 1) * 8);
                float xv[24];
                GTIC;
                {
                    const int xs = xq_it % Q_XSTAGES;
                    g_mbar_wait(&xfull[xs], (xq_it / Q_XSTAGES) & 1);
                    GTOC(5);
                    const float* xp = xring + xs * (3 * Q_BLK * Q_M) + (uhalf * 16 + (hb & 1) * 8) * Q_M + m;
                    #pragma unroll
                    for (int g = 0; g < 3; ++g)
                        #pragma unroll
                        for (int c = 0; c < 8; ++c) xv[g * 8 + c] = (VAR & 2) ? 0.25f : xp[(g * Q_BLK + c) * Q_M];
                    ++xq_it;
                }
                g_mbar_wait(&acc_full[blk], step & 1);
                GTOC(j < 2 ? 0 : (j < ITERS - 2 ? 1 : 2));     // wait for the first / middle / last accumulators of the step
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t ar[8], az[8], an[8];
                g_tmem_ld8(tcol, ar);
                g_tmem_ld8(tcol + 16, az);
                g_tmem_ld8(tcol + 32, an);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                GTOC(3);
                const float4 bn0 = *reinterpret_cast<const float4*>(s_bhn + uu);
                const float4 bn1 = *reinterpret_cast<const float4*>(s_bhn + uu + 4);
                const float bnv[8] = {bn0.x, bn0.y, bn0.z, bn0.w, bn1.x, bn1.y, bn1.z, bn1.w};
                #pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float r, z, n;
                    if (VAR & 8) {                             // timing experiment only: no MUFU
                        r = 0.5f + 0.25f * (xv[c] + __uint_as_float(ar[c]));
                        z = 0.5f + 0.25f * (xv[8 + c] + __uint_as_float(az[c]));
                        n = 0.5f * (xv[16 + c] + r * (__uint_as_float(an[c]) + bnv[c]));
                    } else {
                        hk[j * 8 + c] = g_gru_cell(xv[c] + __uint_as_float(ar[c]), xv[8 + c] + __uint_as_float(az[c]),
                                                   __uint_as_float(an[c]) + bnv[c], xv[16 + c], hk[j * 8 + c]);
                        continue;
                    }
                    hk[j * 8 + c] = (1.0f - z) * n + z * hk[j * 8 + c];
                }
                const float* hv = hk + j * 8;
                uint4 hi, mid;
                q_split2(hv[0], hv[1], hi.x, mid.x);
                q_split2(hv[2], hv[3], hi.y, mid.y);
                q_split2(hv[4], hv[5], hi.z, mid.z);
                q_split2(hv[6], hv[7], hi.w, mid.w);
                const uint32_t to = tile_off(j);
                *reinterpret_cast<uint4*>(hnext + to) = hi;
                *reinterpret_cast<uint4*>(hnext + KB * Q_HTILE + to) = mid;
                // The projection stage goes back to the producer only HERE, behind stores of values computed from all 24 loads:
                // an arrive issued right behind the loads does not wait for their data, and a refill that hits in L2 can
                // overwrite the stage under them (profiles/r2_gru4_race.txt).
                __syncwarp();
                if (lane == 0) g_mbar_arrive(&xempty[(xq_it - 1) % Q_XSTAGES]);
                if (b_ok && !(VAR & 4)) {                      // pre-split output: an operand of the next GEMM
                    *reinterpret_cast<uint4*>(out_hi + orow + uu) = hi;
                    *reinterpret_cast<uint4*>(out_mid + orow + uu) = mid;
                }
                GTOC(4);
            }
            // every tcgen05.ld of the step has completed (wait::ld above) and h_t is in this CTA's shared memory
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) q_arrive_leader(h_ready);
        }
        if (tim && warp == 2 && lane == 0) { for (int i = 0; i < 6; ++i) timing[8 + i] = tacc[i]; timing[15] = clock64(); }
    }
    #undef GTIC
    #undef GTOC
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    g_cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

template <int H, int GW, int VAR>
static int launch_gru3_t(const CUtensorMap& map_hi, const CUtensorMap& map_mid, const CUtensorMap& map_x, int64_t bp,
                         const float* bhn, uint16_t* out_hi, uint16_t* out_mid, int64_t osb, int64_t ost, int64_t batch,
                         cudaStream_t s) {
    CTO_CHECK(set_max_dynamic_smem(gru3_kernel<H, GW, VAR>, Gru3Smem<H>::TOTAL));
    const int ctas = ceil_div(batch, Q_M);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((ctas + 1) / 2 * 2), 2, 1);
    cfg.blockDim = dim3(96 + 128 * GW, 1, 1);
    cfg.dynamicSmemBytes = Gru3Smem<H>::TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CTO_CHECK(cudaLaunchKernelEx(&cfg, gru3_kernel<H, GW, VAR>, map_hi, map_mid, map_x, bp, bhn, out_hi, out_mid, osb, ost, batch,
                                 (g_gemm_debug && g_gemm_timing) ? g_gemm_timing + 32 : nullptr));   // GEMM kernels own [0, 32)
    return 0;
}

}  // namespace tc

// Eight gate-math warps (GW = 2 per TMEM lane quadrant).  12 and 16 warps were built and measured in round 1: slower
// (register spills at <= 112 registers per thread, and the gate phase is not short of warps but of load latency).

// w_hi / w_mid: [2 directions][3H rows regrouped as (32-unit block, 16-unit half, gate, unit)][H] as bf16 hi / mid.
// xproj: TRANSPOSED input projection, fp32 [6H (dir, gate, unit)][ldx], column t * bp + b; bp >= batch rounded up to 128.
// out_hi / out_mid: bf16 split of h_t, row (b * osb + t * ost) of [.., 2H]: time-major (osb 1, ost bp) when the next
// consumer is another transposed projection, batch-major (osb 33, ost 1) for the flattening fc1.
int launch_gru3(const float* xproj, int64_t ldx, int64_t bp, const uint16_t* w_hi, const uint16_t* w_mid, const float* bhn,
                uint16_t* out_hi, uint16_t* out_mid, int64_t osb, int64_t ost, int64_t batch, int hidden, cudaStream_t s) {
    if (batch <= 0) return 0;
    CTO_REQUIRE(hidden == 128 || hidden == 192, "gru3: hidden size %d not built", hidden);
    CTO_REQUIRE(xproj && out_hi && out_mid && bp % 128 == 0 && bp >= batch && ldx >= N_POS * bp, "gru3: bad buffers / padding");
    CTO_REQUIRE(ldx % 4 == 0 && ldx < (1ll << 31), "gru3: projection row stride %lld", (long long)ldx);
    CUtensorMap map_hi, map_mid, map_x;
    if (tc::make_map_bf16(&map_hi, w_hi, 6 * hidden, hidden, hidden, tc::Q_HALF)) return 1;
    if (tc::make_map_bf16(&map_mid, w_mid, 6 * hidden, hidden, hidden, tc::Q_HALF)) return 1;
    if (tc::make_map_plain(&map_x, xproj, 6 * hidden, ldx, ldx, tc::Q_BLK, tc::Q_M)) return 1;   // box: 32 units x 64 candidates
#define CTO_GRU3(HH, GG) tc::launch_gru3_t<HH, GG, 0>(map_hi, map_mid, map_x, bp, bhn, out_hi, out_mid, osb, ost, batch, s)
    int rc;
    if (hidden == 128) rc = CTO_GRU3(128, 2);
#ifdef CTO_DEBUG_KNOBS
    // timing-attribution builds (WRONG results: no projection loads / no output stores / no MUFU); only compiled into the
    // debug library (python -m clairs_to_b200.build --debug), never into the release libcto_b200.so
#define CTO_GRU3V(VV) tc::launch_gru3_t<192, 2, VV>(map_hi, map_mid, map_x, bp, bhn, out_hi, out_mid, osb, ost, batch, s)
    else if ((g_gemm_debug & 14) == 2) rc = CTO_GRU3V(2);
    else if ((g_gemm_debug & 14) == 4) rc = CTO_GRU3V(4);
    else if ((g_gemm_debug & 14) == 8) rc = CTO_GRU3V(8);
    else if ((g_gemm_debug & 14) == 14) rc = CTO_GRU3V(14);
#undef CTO_GRU3V
#endif
    else rc = CTO_GRU3(192, 2);
#undef CTO_GRU3
    if (rc) return rc;
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
