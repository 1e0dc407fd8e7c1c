// GRU recurrence, fourth generation: TWO independent recurrence chains per CTA pair, so that the gate math of one
// chain runs while the tensor cores work on the other.
//
// What the phase counters of gru_tc3.cu showed (profiles/r2_gru4_phase.txt, H = 192, one chain per CTA pair): per step the
// MMA warp is busy ~7 000 cycles and then WAITS ~9 000 cycles for the gate math, because h_t must be complete before
// the products of step t+1 can start -- the tensor pipe idles 56 % of the time.  The dependency is inherent to ONE chain.
//
// Chain c of a CTA pair owns 128 candidates (64 per CTA; [pair * 256 + c * 128, +128)) of direction blockIdx.y.
//   * MMA order: per step and 32-unit block, chain 0's products, then chain 1's, on the SAME W stages (the weights belong
//     to the step, not to the chain): W passes through L2 once per step, and the chains stay one block apart.
//   * tensor memory: every chain has THREE accumulator regions of 48 columns (one block each), used cyclically for blocks
//     (0, 3), (1, 4), (2, 5); a region goes back to the MMA warp as soon as the chain's gate warps have loaded it.  That
//     leaves 192 columns for the fp32 state h itself (96 per chain: each thread owns one TMEM lane and reads / writes its
//     96 units with tcgen05.ld / st) instead of 96 registers per thread;
//   * shared memory: h_t tiles (bf16 hi / mid, the A operand) are SINGLE buffered per chain (2 x 48 KB): the tiles of the
//     last block are written directly (all MMAs of the chain-step have retired by then), the others in a short second
//     pass from the state in tensor memory.  W ring 6 x 12 KB, projection ring 2 x 12 KB per chain, filled by one 4-d TMA
//     box per half-block (a TMA instruction costs its issuing warp ~135 cycles);
//   * the projection stage goes back to its producer with an arrive that is DATA DEPENDENT on the loaded values
//     (profiles/r2_gru4_race.txt).
// Measured: the kernel is bound by shared-memory bandwidth (per step, in 128-byte wavefronts: MMA operands 12 k, W stages
// 1.7 k, projection in 2.3 k + out 2.3 k, h tiles 0.8 k against a step of ~20 k cycles).  Reading the projection with
// per-thread global loads instead (no shared memory at all) was built and is slower: one half-block of register
// look-ahead does not cover the latency, two do not fit in 168 registers.
// Arithmetic, the weight / projection layouts, the thread <-> (candidate, unit) mapping and the output planes are those
// of gru_tc3.cu (torch.nn.GRU semantics, clairs/model.py:412-417); the two kernels are bit-identical (tests/test_gpu_gru.py).
#include "gru_ptx.cuh"

namespace cto {

extern long long* g_gemm_timing;      // cto_debug_timing(): device buffer [64]; [32, 64) is the GRU kernel's
extern int g_gemm_debug;
namespace tc {

constexpr int R_M = 64;                      // candidates per CTA and chain
constexpr int R_BLK = 32;                    // hidden units per W block
constexpr int R_N = 3 * R_BLK;               // 96 gate columns per block (MMA N)
constexpr int R_HALF = R_N / 2;              // B rows / accumulator columns per CTA
constexpr int R_K = 64;                      // bf16 elements per 128-byte swizzle row
constexpr int R_HTILE = R_M * 128;           // 8 KB: 64 rows x 64 k (bf16)
constexpr int R_WTILE = R_HALF * 128;        // 6 KB
constexpr int R_STAGE = 2 * R_WTILE;         // hi | mid
// W ring: WST stages; projection ring: XST stages per chain, one stage = one half-block: 3 gates x 2 unit groups x 8 units x 64 candidates fp32
constexpr int R_XSTAGE = 3 * 2 * 8 * R_M * 4;   // 12 KB
constexpr int R_THREADS = 384;               // warp 0 W producer, 1 MMA, 2-5 gates chain 0, 6-9 gates chain 1, 10 / 11 x producers
constexpr uint32_t R_PEER_MASK = 0xFEFFFFFFu;
constexpr uint32_t R_SPIN = 1u << 24;

template <int H, int XST, int WST>                            // XST projection stages per chain, WST W stages
struct Gru4Smem {
    static constexpr int KB = H / R_K;
    static constexpr int HBUF = 2 * KB * R_HTILE;             // hi | mid tiles of one chain's h_t
    static constexpr int TOTAL = 2 * HBUF + WST * R_STAGE + 2 * XST * R_XSTAGE + 1024 + 512 + H * 4;
    static_assert(TOTAL <= 232448, "shared memory");
};

__device__ __forceinline__ void r_tma_load_2sm(const CUtensorMap* map, uint64_t* leader_bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(g_smem_u32(dst)), "l"(map), "r"(g_smem_u32(leader_bar) & R_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void r_mma_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void r_commit_2sm(uint64_t* bar) {       // arrives on the same barrier in both CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(g_smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void r_arrive_leader(uint64_t* bar) {   // DSMEM arrive on the leader CTA's barrier
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote) : "r"(g_smem_u32(bar)));
    // default semantics (.release.cta), as CUTLASS signals a peer CTA: the .release.cluster form compiles to MEMBAR.ALL.GPU,
    // which stalls the warp until every output store it has in flight is acknowledged by L2
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// bounded waits: a protocol error records (code, block, thread) in dbg and unwinds instead of hanging the GPU
__device__ __forceinline__ bool r_try(uint64_t* bar, uint32_t parity, bool cluster) {
    uint32_t ok;
    if (cluster)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(g_smem_u32(bar)), "r"(parity) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(g_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __noinline__ bool r_wait_slow(uint64_t* bar, uint32_t parity, bool cluster, int code, int* dbg) {
    for (uint32_t tries = 0; tries < R_SPIN; ++tries) {
        if (r_try(bar, parity, cluster)) return true;
        if ((tries & 255u) == 255u && *reinterpret_cast<volatile int*>(dbg) != 0) return false;
    }
    if (atomicCAS(dbg, 0, code) == 0) {
        dbg[1] = (int)(blockIdx.x + 10000 * blockIdx.y);
        dbg[2] = (int)threadIdx.x;
        dbg[3] = (int)parity;
    }
    return false;
}
__device__ __forceinline__ bool r_wait(uint64_t* bar, uint32_t parity, bool cluster, int code, int* dbg) {
    if (r_try(bar, parity, cluster)) return true;
    return r_wait_slow(bar, parity, cluster, code, dbg);
}
__device__ __forceinline__ uint32_t r_pack(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void r_split2(float x0, float x1, uint32_t& hi, uint32_t& mid) {
    hi = r_pack(x0, x1);
    mid = r_pack(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xFFFF0000u));
}
__device__ __forceinline__ void r_tmem_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                   "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                   "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}

__device__ __forceinline__ void r_tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void r_tmem_ld32(uint32_t taddr, uint32_t* r) {
    r_tmem_ld16(taddr, r);
    r_tmem_ld16(taddr + 16, r + 16);
}

__device__ __forceinline__ void r_st_global_256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

enum { RB_WFULL = 101, RB_WEMPTY, RB_HREADY, RB_SLOTFREE, RB_SLOTFULL, RB_XFULL, RB_XEMPTY };

template <int H, int XST, int WST>
__global__ void __launch_bounds__(R_THREADS, 1)
gru4_kernel(const __grid_constant__ CUtensorMap tma_whi, const __grid_constant__ CUtensorMap tma_wmid,
            const __grid_constant__ CUtensorMap tma_x, int64_t bp,
            const float* __restrict__ bhn, uint16_t* __restrict__ out_hi, uint16_t* __restrict__ out_mid, int64_t osb,
            int64_t ost, int64_t batch, int* dbg,
            long long* timing, int zero) {
    // per-phase clock64() counters of cluster 0, direction 0 (profiles/phase_timing_gru.py); timing == nullptr in production
    const bool tim = timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    #define RTIC long long _t0 = tim ? clock64() : 0
    #define RTOC(i) do { if (tim) { long long _t1 = clock64(); tacc[i] += _t1 - _t0; _t0 = _t1; } } while (0)
    constexpr int KB = H / R_K;
    constexpr int NB = H / R_BLK;                              // 6 blocks of 32 units
    constexpr int NREG = 3;                                    // accumulator regions (one 32-unit block each) per chain
    constexpr int NHB = 2 * NB;                                // 12 half-blocks (8 units per thread each)
    constexpr int HBUF = Gru4Smem<H, XST, WST>::HBUF;
    constexpr uint32_t STATE_COL = 2 * NREG * R_HALF;          // 288: fp32 state, 96 columns per chain
    static_assert(H == 192, "the slot / state layout below is written for H = 192 (288 + 2 * 96 <= 512 TMEM columns)");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = smem_raw + ((1024u - (g_smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* hbuf = base;                                     // [chain][hi | mid][KB][64 x 128 B]
    uint8_t* wring = base + 2 * HBUF;
    float* xring = reinterpret_cast<float*>(wring + WST * R_STAGE);      // [chain][stage][gate 3][group 2][unit 8][candidate 64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(wring + WST * R_STAGE + 2 * XST * R_XSTAGE);
    uint64_t* full = bars;                                    // [6]  W ring (leader: both halves land here)
    uint64_t* empty = full + WST;                        // [6]
    uint64_t* xfull = empty + WST;                       // [chain][2]
    uint64_t* xempty = xfull + 2 * XST;                 // [chain][2]
    uint64_t* acc_full = xempty + 2 * XST;              // [chain][3]  accumulators of a block complete (both CTAs)
    uint64_t* acc_free = acc_full + 2 * NREG;                 // [chain][3]  leader: loaded by the gate warps of BOTH CTAs
    uint64_t* h_ready = acc_free + 2 * NREG;                  // [chain]     leader: h_t tiles written in BOTH CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(h_ready + 2);
    float* s_bhn = reinterpret_cast<float*>(bars + 48);       // 384 bytes of barriers

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const uint32_t crank = g_cluster_rank();
    const bool leader = crank == 0;
    const int pair = (int)(blockIdx.x >> 1);
    for (int i = threadIdx.x; i < H; i += R_THREADS) s_bhn[i] = bhn[dir * H + i];

    if (threadIdx.x == 0) {
        for (int s = 0; s < WST; ++s) { g_mbar_init(&full[s], 1); g_mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2 * XST; ++s) { g_mbar_init(&xfull[s], 1); g_mbar_init(&xempty[s], 4); }
        for (int s = 0; s < 2 * NREG; ++s) { g_mbar_init(&acc_full[s], 1); g_mbar_init(&acc_free[s], 8); }
        for (int c = 0; c < 2; ++c) g_mbar_init(&h_ready[c], 8);           // one arrive per gate warp of the chain in BOTH CTAs
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        g_prefetch_map(&tma_whi); g_prefetch_map(&tma_wmid); g_prefetch_map(&tma_x);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    g_cluster_sync();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---- W producer: this CTA's half of every W block, ONCE PER STEP: both chains use the stage ----
        uint32_t it = 0;
        bool alive = true;
        for (int cs = 0; cs < N_POS && alive; ++cs) {
            for (int blk = 0; blk < NB && alive; ++blk) {
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const int s = it % WST;
                    if (!r_wait(&empty[s], ((it / WST) & 1) ^ 1, false, RB_WEMPTY, dbg)) { alive = false; break; }
                    uint8_t* st = wring + s * R_STAGE;
                    const int row = dir * 3 * H + blk * R_N + (int)crank * R_HALF;
                    if (g_elect_one()) {
                        if (leader) g_mbar_expect_tx(&full[s], 2 * R_STAGE);      // both halves land on the leader's barrier
                        r_tma_load_2sm(&tma_whi, &full[s], st, kb * R_K, row);
                        r_tma_load_2sm(&tma_wmid, &full[s], st + R_WTILE, kb * R_K, row);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ---- MMA issuer (leader CTA only).  Per step and 32-unit block: chain 0's products, then chain 1's, on the SAME
            // W stages (the weights are those of the step, not of the chain): W goes through L2 once per step instead of
            // once per chain-step, and the chains stay one block apart, so each one's gate math runs under the other's MMAs.
            // D=f32, A=B=bf16, K-major, M=128 (64 rows per CTA), N=96
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(R_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            uint32_t it = 0;
            bool alive = true;
            RTIC;
            for (int step = 0; step < N_POS && alive; ++step) {
                for (int blk = 0; blk < NB && alive; ++blk, it += KB) {
                    const int r = blk % NREG, u = step * 2 + blk / NREG;       // accumulator region of the chain, and its use count
                    for (int c = 0; c < 2 && alive; ++c) {
                        if (blk == 0) {                                        // h_{t-1} of chain c in BOTH CTAs
                            if (!r_wait(&h_ready[c], step & 1, true, RB_HREADY, dbg)) { alive = false; break; }
                            RTOC(0);
                        }
                        if (u >= 1) {                                          // the chain's gate warps have loaded the region's previous block
                            if (!r_wait(&acc_free[c * NREG + r], (u - 1) & 1, true, RB_SLOTFREE, dbg)) { alive = false; break; }
                            RTOC(1);
                        }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint8_t* hb = hbuf + c * HBUF;
                        const uint32_t acc = tmem_base + (uint32_t)((c * NREG + r) * R_HALF);
                        for (int kb = 0; kb < KB; ++kb) {
                            const uint32_t wi = it + kb;
                            const int ws = wi % WST;
                            if (c == 0) {
                                if (!r_wait(&full[ws], (wi / WST) & 1, false, RB_WFULL, dbg)) { alive = false; break; }
                                RTOC(2);
                                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            }
                            const uint64_t d_hhi = g_desc_k_sw128(g_smem_u32(hb + kb * R_HTILE));
                            const uint64_t d_hmid = g_desc_k_sw128(g_smem_u32(hb + (KB + kb) * R_HTILE));
                            const uint32_t w_addr = g_smem_u32(wring + ws * R_STAGE);
                            const uint64_t d_whi = g_desc_k_sw128(w_addr);
                            const uint64_t d_wmid = g_desc_k_sw128(w_addr + R_WTILE);
                            if (g_elect_one()) {
                                #pragma unroll
                                for (int k = 0; k < R_K / 16; ++k) {
                                    const uint64_t o = (uint64_t)(k * 2);          // 16 bf16 = 32 bytes along the swizzle row
                                    r_mma_2sm(acc, d_hhi + o, d_whi + o, idesc, (kb | k) ? 1u : 0u);
                                    r_mma_2sm(acc, d_hmid + o, d_whi + o, idesc, 1u);
                                    r_mma_2sm(acc, d_hhi + o, d_wmid + o, idesc, 1u);
                                }
                                if (c == 1) r_commit_2sm(&empty[ws]);      // both chains are through: stage free in both CTAs
                                if (kb == KB - 1) r_commit_2sm(&acc_full[c * NREG + r]);
                            }
                            __syncwarp();
                            RTOC(3);
                        }
                    }
                }
            }
            if (tim && lane == 0) { for (int i = 0; i < 4; ++i) timing[i] = tacc[i]; }
        }
    } else if (warp >= 10) {
        // ---- x producers (one per chain): the input projection xproj^T[(dir, gate, unit)][t * bp + b] of the chain's 64
        // candidates in this CTA, one half-block (the 8 + 8 units of one gate-math iteration) per stage ----
        const int c = warp - 10;
        {
        const int b0 = pair * 256 + c * 128 + (int)crank * R_M;
        float* ring = xring + c * (XST * R_XSTAGE / 4);
        uint32_t q = 0;
        bool alive = true;
        for (int step = 0; step < N_POS && alive; ++step) {
            const int t = dir ? (N_POS - 1 - step) : step;
            for (int hb = 0; hb < NHB; ++hb, ++q) {
                const int s = q % XST;
                if (!r_wait(&xempty[c * XST + s], ((q / XST) & 1) ^ 1, false, RB_XEMPTY, dbg)) { alive = false; break; }
                if (g_elect_one()) {
                    // ONE 4-d box = (64 candidates, 8 units, 2 unit groups 16 rows apart, 3 gates H rows apart): issuing a TMA
                    // instruction costs the producer ~135 cycles, six 2-d boxes per half-block made it the pace setter
                    g_mbar_expect_tx(&xfull[c * XST + s], R_XSTAGE);
                    g_tma_load_4d(&tma_x, &xfull[c * XST + s], ring + s * (R_XSTAGE / 4), (int)(t * bp) + b0,
                                  dir * 3 * H + (hb >> 1) * R_BLK + (hb & 1) * 8, 0, 0);
                }
                __syncwarp();
            }
        }
        }
    } else {
        // ---- gate math: warps 2-5 chain 0, warps 6-9 chain 1, in both CTAs ----
        const int c = (warp - 2) >> 2;
        const int quad = warp & 3;                         // TMEM lanes [32*quad, +32)
        const int tl = quad * 32 + lane;                   // TMEM lane
        const int m = tl & 63;                             // candidate row inside the CTA
        const int uhalf = tl >> 6;                         // units [16*uhalf, +16) of each block
        const int64_t b = (int64_t)pair * 256 + c * 128 + (int64_t)crank * R_M + m;
        const bool b_ok = b < batch;
        const uint32_t lanes = (uint32_t)(quad * 32) << 16;
        const uint32_t row_off = (uint32_t)((m >> 3) * 1024 + (m & 7) * 128);
        uint8_t* hb_hi = hbuf + c * HBUF;
        uint8_t* hb_mid = hb_hi + KB * R_HTILE;
        const float* ring = xring + c * (XST * R_XSTAGE / 4);
        const uint32_t state = tmem_base + lanes + STATE_COL + (uint32_t)(c * 96);
        // half-block (blk, sub): this thread's 8 units start at uu = 32 blk + 16 uhalf + 8 sub; their bf16 values are
        // 16-byte chunk ((uu % 64) / 8) of row m in k-block uu / 64; their fp32 state is TMEM columns state + 16 blk + 8 sub
        auto tile_off = [&](int uu) {
            return (uint32_t)((uu / R_K) * R_HTILE) + row_off + (uint32_t)(((((uu % R_K) >> 3)) ^ (m & 7)) << 4);
        };
        {   // h_0 = 0: tiles and state
            float z8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            #pragma unroll 1
            for (int hb = 0; hb < NHB; ++hb) {
                const int uu = (hb >> 1) * R_BLK + uhalf * 16 + (hb & 1) * 8;
                *reinterpret_cast<uint4*>(hb_hi + tile_off(uu)) = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(hb_mid + tile_off(uu)) = make_uint4(0u, 0u, 0u, 0u);
                r_tmem_st8(state + (uint32_t)(hb * 8), z8);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) r_arrive_leader(&h_ready[c]);
        }
        uint32_t xq = 0;
        bool alive = true;
        RTIC;
        for (int step = 0; step < N_POS && alive; ++step) {
            const int t = dir ? (N_POS - 1 - step) : step;
            const int64_t orow = (b * osb + t * ost) * (int64_t)(2 * H) + dir * H;
            #pragma unroll 1
            for (int blk = 0; blk < NB; ++blk) {
                const int r = blk % NREG;
                // the block's accumulators: second use of the region in this step for blk >= 3
                if (!r_wait(&acc_full[c * NREG + r], (uint32_t)(step * 2 + blk / NREG) & 1, false, RB_SLOTFULL, dbg)) { alive = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                RTOC(1);
                // region r of this chain: columns [48 * (3c + r), +48) = r(16) z(16) n(16) of this lane's 16 units; and their state
                const uint32_t tcol = tmem_base + lanes + (uint32_t)((c * NREG + r) * R_HALF);
                uint32_t acc[48], hs[16];
                r_tmem_ld32(tcol, acc);
                r_tmem_ld16(tcol + 32, acc + 32);
                r_tmem_ld16(state + (uint32_t)(blk * 16), hs);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the region goes back to the MMA warp
                __syncwarp();
                if (lane == 0) r_arrive_leader(&acc_free[c * NREG + r]);
                RTOC(2);
                uint4 ohi[2], omid[2];
                #pragma unroll
                for (int sub = 0; sub < 2; ++sub, ++xq) {
                    const int uu = blk * R_BLK + uhalf * 16 + sub * 8;
                    // the projection of this half-block: [gate][unit group = uhalf][unit][candidate]: lanes are consecutive
                    // candidates, conflict-free LDS
                    const int xs = xq % XST;
                    float xv[24];
                    if (!r_wait(&xfull[c * XST + xs], (xq / XST) & 1, false, RB_XFULL, dbg)) { alive = false; break; }
                    const float* xp = ring + xs * (R_XSTAGE / 4) + uhalf * (8 * R_M) + m;
                    #pragma unroll
                    for (int g = 0; g < 3; ++g)
                        #pragma unroll
                        for (int e = 0; e < 8; ++e) xv[g * 8 + e] = xp[(g * 2) * (8 * R_M) + e * R_M];
                    // The stage goes back to the producer as soon as the 24 loads have RETURNED.  An arrive issued right behind
                    // the loads does not wait for their data, and a refill that hits in L2 then overwrites the stage under
                    // them (seen as sporadic wrong candidates in groups of 4..32 when the projection had just been written,
                    // profiles/r2_gru4_race.txt).  So the arrive is made data dependent on all 24 values: its address adds
                    // (OR of their bits) & zero, where zero is a kernel argument that is always 0.
                    uint32_t dep = 0;
                    #pragma unroll
                    for (int i = 0; i < 24; ++i) dep |= __float_as_uint(xv[i]);
                    __syncwarp();
                    if (lane == 0) g_mbar_arrive(&xempty[c * XST + xs] + (dep & (uint32_t)zero));
                    RTOC(0);
                    const float4 bn0 = *reinterpret_cast<const float4*>(s_bhn + uu);
                    const float4 bn1 = *reinterpret_cast<const float4*>(s_bhn + uu + 4);
                    const float bnv[8] = {bn0.x, bn0.y, bn0.z, bn0.w, bn1.x, bn1.y, bn1.z, bn1.w};
                    float hv[8];
                    #pragma unroll
                    for (int e = 0; e < 8; ++e)
                        hv[e] = g_gru_cell(xv[e] + __uint_as_float(acc[sub * 8 + e]), xv[8 + e] + __uint_as_float(acc[16 + sub * 8 + e]),
                                           __uint_as_float(acc[32 + sub * 8 + e]) + bnv[e], xv[16 + e], __uint_as_float(hs[sub * 8 + e]));
                    r_tmem_st8(state + (uint32_t)(blk * 16 + sub * 8), hv);
                    RTOC(3);
                    uint4& hi = ohi[sub];
                    uint4& mid = omid[sub];
                    r_split2(hv[0], hv[1], hi.x, mid.x);
                    r_split2(hv[2], hv[3], hi.y, mid.y);
                    r_split2(hv[4], hv[5], hi.z, mid.z);
                    r_split2(hv[6], hv[7], hi.w, mid.w);
                    if (blk == NB - 1) {
                        // last block complete = every MMA of this chain-step has retired: these tiles may be overwritten now
                        const uint32_t to = tile_off(uu);
                        *reinterpret_cast<uint4*>(hb_hi + to) = hi;
                        *reinterpret_cast<uint4*>(hb_mid + to) = mid;
                    }
                    RTOC(4);
                }
                if (!alive) break;
                if (b_ok) {
                    // pre-split output, an operand of the next GEMM: the thread's 16 units of the block are 32 contiguous bytes
                    // per plane = one full sector per store (two 16-byte stores per sector cost 14 % of the kernel)
                    const int64_t o = orow + blk * R_BLK + uhalf * 16;
                    r_st_global_256(out_hi + o, ohi[0], ohi[1]);
                    r_st_global_256(out_mid + o, omid[0], omid[1]);
                }
                RTOC(4);
            }
            if (!alive) break;
            // the h tiles of the other blocks: their gate math ran while later MMAs of the step were still reading
            // h_{t-1}; rebuilt now from the state in tensor memory
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            #pragma unroll 1
            for (int blk = 0; blk < NB - 1; ++blk) {
                uint32_t hs[16];
                r_tmem_ld16(state + (uint32_t)(blk * 16), hs);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                #pragma unroll
                for (int sub = 0; sub < 2; ++sub) {
                    uint4 hi, mid;
                    r_split2(__uint_as_float(hs[sub * 8 + 0]), __uint_as_float(hs[sub * 8 + 1]), hi.x, mid.x);
                    r_split2(__uint_as_float(hs[sub * 8 + 2]), __uint_as_float(hs[sub * 8 + 3]), hi.y, mid.y);
                    r_split2(__uint_as_float(hs[sub * 8 + 4]), __uint_as_float(hs[sub * 8 + 5]), hi.z, mid.z);
                    r_split2(__uint_as_float(hs[sub * 8 + 6]), __uint_as_float(hs[sub * 8 + 7]), hi.w, mid.w);
                    const uint32_t to = tile_off(blk * R_BLK + uhalf * 16 + sub * 8);
                    *reinterpret_cast<uint4*>(hb_hi + to) = hi;
                    *reinterpret_cast<uint4*>(hb_mid + to) = mid;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) r_arrive_leader(&h_ready[c]);
            RTOC(5);
        }
        if (tim && (warp == 2 || warp == 6) && lane == 0) { for (int i = 0; i < 6; ++i) timing[8 + c * 8 + i] = tacc[i]; }
    }
    #undef RTIC
    #undef RTOC
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    g_cluster_sync();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace tc

// Layer-2 recurrence with two chains per CTA pair (H = 192).  Same arguments as launch_gru3, plus dbg (device int[8],
// zeroed once by the caller: a non-zero dbg[0] after the launch = an in-kernel barrier wait timed out).
// bp must cover the 256-candidate granularity of a CTA pair: columns beyond the projection's width are zero-filled by TMA.
int launch_gru4(const float* xproj, int64_t ldx, int64_t bp, const uint16_t* w_hi, const uint16_t* w_mid, const float* bhn,
                uint16_t* out_hi, uint16_t* out_mid, int64_t osb, int64_t ost, int64_t batch, int hidden, int* dbg, cudaStream_t s) {
    if (batch <= 0) return 0;
    CTO_REQUIRE(hidden == 192, "gru4: hidden size %d not built (192 is)", hidden);
    CTO_REQUIRE(xproj && out_hi && out_mid && dbg && bp % 128 == 0 && bp >= batch && ldx >= N_POS * bp, "gru4: bad buffers / padding");
    CTO_REQUIRE(ldx % 4 == 0 && ldx < (1ll << 31), "gru4: projection row stride %lld", (long long)ldx);
    CUtensorMap map_hi, map_mid, map_x;
    if (tc::make_map_bf16(&map_hi, w_hi, 6 * hidden, hidden, hidden, tc::R_HALF)) return 1;
    if (tc::make_map_bf16(&map_mid, w_mid, 6 * hidden, hidden, hidden, tc::R_HALF)) return 1;
    {   // projection box: 64 candidates x (8 units, 2 unit groups 16 rows apart, 3 gates `hidden` rows apart) in one instruction
        const int64_t dim_size[3] = {6 * hidden, 2, 3}, row_stride[3] = {1, 16, hidden};
        const int box[4] = {tc::R_M, 8, 2, 3};
        if (tc::make_map_rows_nd(&map_x, xproj, 4, ldx, ldx, dim_size, row_stride, box)) return 1;
    }
    const int pairs = ceil_div(batch, 256);
    long long* timing = (g_gemm_debug && g_gemm_timing) ? g_gemm_timing + 32 : nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(pairs * 2), 2, 1);
    cfg.blockDim = dim3(tc::R_THREADS, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    // ring depths: 2 projection stages per chain + 6 W stages (a block's three stages serve both chains, three more in flight)
    CTO_CHECK(set_max_dynamic_smem(tc::gru4_kernel<192, 2, 6>, tc::Gru4Smem<192, 2, 6>::TOTAL));
    cfg.dynamicSmemBytes = tc::Gru4Smem<192, 2, 6>::TOTAL;
    CTO_CHECK(cudaLaunchKernelEx(&cfg, tc::gru4_kernel<192, 2, 6>, map_hi, map_mid, map_x, bp, bhn, out_hi, out_mid, osb, ost, batch, dbg, timing, 0));
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
