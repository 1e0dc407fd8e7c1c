// Pileup tensor encoder: per-site read arrays -> int16 [N, 33, 34].
//
// Replaces decode_pileup_bases() + window assembly of the reference
// (src/create_tensor_pileup_calling.py:146-229, 461, 513-516, 537-543; cited as CT).
//
// HBM-bound integer work.  One CTA encodes a group of 4 candidates = 132 (candidate, flank slot)
// pairs, one THREAD per slot:
//   1. the group's rows cover one contiguous span of the read arrays (rows are position-sorted), so the
//      three byte streams (code, bq, mq) of the span are fetched with three bulk async copies
//      (cp.async.bulk, completion on an mbarrier) into shared memory - full-line HBM reads, no LSU work;
//   2. every thread walks the reads of its own row in shared memory and counts in REGISTERS: the three
//      groups of eight base fields (MQ >= 20, MQ < 20, low BQ) are three 64-bit registers of 8-bit lanes,
//      one shifted increment per read, flushed into 32-bit totals every 255 reads (round 1 kept 16-bit
//      counters in shared memory: a load-add-store chain per read, 45 % issue utilisation);
//   3. the rare indel-carrying reads come from a sparse side list and are resolved exactly (per-allele
//      maximum, CT:184-187, 201-204) with a K^2 scan that has no table-size limit;
//   4. the 34 int16 of each slot are staged in shared memory and the group's 8.8 KB output block is
//      written with 16-byte coalesced stores.
// Groups whose span does not fit the staging buffers (very deep pileups, scattered rows) take the
// same code path with the pointers left in global memory.
#include "common.cuh"

namespace cto {

namespace enc {

constexpr int GROUP = 4;                          // candidates per CTA (4 CTAs per SM hide the bulk-copy latency)
constexpr int SLOTS = GROUP * N_POS;              // 132
constexpr int THREADS = 160;                      // 5 warps
constexpr int STAGE_CAP = 10 * 1024;              // bytes per staged array
constexpr int FLUSH = 255;                        // reads per pass of the 8-bit lane counters
constexpr int OUT_BYTES = SLOTS * N_CH * 2;       // 8976

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// set-A field -> channel (CT:55-58)
__device__ __forceinline__ int channel_of_a(int f) {
    if (f < 4) return f;                // A C G T
    if (f < 8) return f + 5;            // a c g t  -> 9..12
    if (f == 8) return 8;               // '*'
    if (f == 9) return 17;              // '#'
    return f + 8;                       // LMQ      -> 18..25
}

// Reads [lo, hi) of one row -> the 26 counters (CT:146-149, 160-221).  The same code for the staged (shared
// memory) and the unstaged (global) arrays; the pointer type keeps the address space, so the staged instance
// compiles to LDS.  One read = one shifted 64-bit increment `1 << 8*b8` added to up to two of the three lane
// registers under a predicate; indel-carrying reads (bit 4 of the code) count nowhere here (CT:160-204).
template <typename P>
__device__ __forceinline__ void count_reads(P p_code, P p_bq, P p_mq, int lo, int hi, int low_bq_cut, int (&cnt)[26]) {
    for (int base = lo; base < hi; base += FLUSH) {
        const int end = min(hi, base + FLUSH);
        unsigned long long a = 0, l = 0, b = 0;           // 8 x 8-bit lanes each: MQ >= 20 | MQ < 20 | low BQ
        unsigned int sh = 0;                              // '*' in bits 0-15, '#' in bits 16-31
        for (int i = base; i < end; ++i) {
            const unsigned int c = p_code[i], m = p_mq[i], q = p_bq[i];
            const unsigned int sym = c & 0xF;
            const bool plain = !(c & 0x10);
            const bool base8 = plain && (sym < 4 || (sym - 5u) < 4u);                  // A C G T a c g t
            const unsigned int b8 = sym - (sym > 4 ? 1u : 0u);
            const bool mq_hi = m >= (unsigned)MIN_MQ && m != (unsigned)QUAL_ABSENT;
            const unsigned long long inc = base8 ? (1ull << (8 * b8)) : 0ull;
            a += mq_hi ? inc : 0ull;
            l += m < (unsigned)MIN_MQ ? inc : 0ull;                                    // CT:215-217
            b += (q != (unsigned)QUAL_ABSENT && (int)q < low_bq_cut) ? inc : 0ull;     // CT:149, 219-221
            sh += (plain && mq_hi && sym == 10) ? 1u : 0u;
            sh += (plain && mq_hi && sym == 11) ? 0x10000u : 0u;
        }
        #pragma unroll
        for (int f = 0; f < 8; ++f) {
            cnt[f] += (int)((a >> (8 * f)) & 0xFF);
            cnt[10 + f] += (int)((l >> (8 * f)) & 0xFF);
            cnt[18 + f] += (int)((b >> (8 * f)) & 0xFF);
        }
        cnt[8] += (int)(sh & 0xFFFF);
        cnt[9] += (int)(sh >> 16);
    }
}

__global__ void __launch_bounds__(THREADS)
encode_pileup_kernel(const uint8_t* __restrict__ code, const uint8_t* __restrict__ bq,
                     const uint8_t* __restrict__ mq, const int32_t* __restrict__ pos_off,
                     const uint8_t* __restrict__ ref_code, const int32_t* __restrict__ ind_off,
                     const uint32_t* __restrict__ ind_entry, const int32_t* __restrict__ win_pos,
                     int64_t n_slots, int low_bq_cut, int16_t* __restrict__ tensor,
                     int32_t* __restrict__ depth_out) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_code = smem;
    uint8_t* s_bq = smem + STAGE_CAP;
    uint8_t* s_mq = smem + 2 * STAGE_CAP;
    int16_t* s_out = reinterpret_cast<int16_t*>(smem + 3 * STAGE_CAP);   // [SLOTS][34]
    __shared__ uint64_t s_bar;
    __shared__ int s_min_row, s_max_row;

    const int tid = threadIdx.x;
    const int64_t slot0 = (int64_t)blockIdx.x * SLOTS;
    const int64_t slot = slot0 + tid;
    const bool active = tid < SLOTS && slot < n_slots;
    const int row = active ? win_pos[slot] : -1;

    if (tid == 0) {
        s_min_row = 0x7fffffff;
        s_max_row = -1;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (row >= 0) {
        // warp-aggregated min / max of the rows this group touches
        const unsigned m = __activemask();
        const int lo = __reduce_min_sync(m, row), hi = __reduce_max_sync(m, row);
        if ((tid & 31) == __ffs(m) - 1) {
            atomicMin(&s_min_row, lo);
            atomicMax(&s_max_row, hi);
        }
    }
    __syncthreads();
    const int min_row = s_min_row, max_row = s_max_row;
    int64_t span_lo = 0, span_hi = 0;
    if (max_row >= 0) {
        span_lo = pos_off[min_row];
        span_hi = pos_off[max_row + 1];
    }
    const int64_t a_lo = span_lo & ~int64_t(15);                      // 16-byte aligned superset of the span
    const int64_t a_hi = (span_hi + 15) & ~int64_t(15);
    const bool staged = (a_hi - a_lo) <= STAGE_CAP && a_hi > a_lo;
    if (staged && tid == 0) {
        const uint32_t bytes = (uint32_t)(a_hi - a_lo);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(3 * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(s_code)), "l"(code + a_lo), "r"(bytes), "r"(smem_u32(&s_bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(s_bq)), "l"(bq + a_lo), "r"(bytes), "r"(smem_u32(&s_bar)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(s_mq)), "l"(mq + a_lo), "r"(bytes), "r"(smem_u32(&s_bar)) : "memory");
    }

    // per-slot metadata and the sparse indel list are read while the bulk copies are in flight
    int lo = 0, hi = 0, ref = 0;
    int tot[4] = {0, 0, 0, 0}, best[4] = {0, 0, 0, 0};
    if (row >= 0) {
        lo = pos_off[row];
        hi = pos_off[row + 1];
        ref = ref_code[row];
        const int ilo = ind_off[row], ihi = ind_off[row + 1];
        for (int i = ilo; i < ihi; ++i) {
            const uint32_t e = ind_entry[i];
            const uint32_t em = (e >> 16) & 0xFF;
            if ((e & (1u << 26)) || em < MIN_MQ || em == QUAL_ABSENT) continue;     // CT:147, 174-176, 189-191
            const uint32_t key = e & 0x0300FFFFu;                                   // allele id + class bits
            int same = 0;
            for (int j = ilo; j < ihi; ++j) {
                const uint32_t o = ind_entry[j];
                const uint32_t om = (o >> 16) & 0xFF;
                same += (!(o & (1u << 26)) && om >= MIN_MQ && om != QUAL_ABSENT && (o & 0x0300FFFFu) == key) ? 1 : 0;
            }
            const int cls = (e >> 24) & 3;                                          // bit0 deletion, bit1 reverse
            #pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                if (cls == c4) {
                    tot[c4] += 1;
                    best[c4] = max(best[c4], same);
                }
            }
        }
    }

    if (staged) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "ENC_WAIT:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
            "@p bra ENC_DONE;\n\t"
            "bra ENC_WAIT;\n\t"
            "ENC_DONE:\n\t"
            "}" ::"r"(smem_u32(&s_bar)) : "memory");
    }
    // counts of this slot: cnt[0..7] = A C G T a c g t with MQ >= 20, cnt[8] = '*', cnt[9] = '#',
    // cnt[10..17] = the eight bases with MQ < 20, cnt[18..25] = the eight bases with BQ < low_bq_cut
    int cnt[26];
    #pragma unroll
    for (int f = 0; f < 26; ++f) cnt[f] = 0;
    if (staged) count_reads(s_code, s_bq, s_mq, lo - (int)a_lo, hi - (int)a_lo, low_bq_cut, cnt);
    else count_reads(code, bq, mq, lo, hi, low_bq_cut, cnt);

    if (tid < SLOTS) {
        int16_t* o = s_out + tid * N_CH;
        int v[N_CH];
        #pragma unroll
        for (int ch = 0; ch < N_CH; ++ch) v[ch] = 0;
        if (row >= 0) {
            #pragma unroll
            for (int f = 0; f < 18; ++f) v[channel_of_a(f)] = cnt[f];
            #pragma unroll
            for (int f = 0; f < 8; ++f) v[26 + f] = cnt[18 + f];
            v[4] = tot[0]; v[6] = tot[1]; v[13] = tot[2]; v[15] = tot[3];           // I D i d
            v[5] = best[0]; v[7] = best[1]; v[14] = best[2]; v[16] = best[3];       // I1 D1 i1 d1 (CT:210-213)
            if (depth_out && (slot % N_POS) == CENTER) {                            // depth of alt_info (CT:208)
                int d = 0;
                #pragma unroll
                for (int ch = 0; ch < 18; ++ch)
                    if (ch != 5 && ch != 7 && ch != 14 && ch != 16) d += v[ch];
                depth_out[slot / N_POS] = d;
            }
            // reference channel := -(sum of the four base counts) in each of the six groups (CT:223-228)
            #pragma unroll
            for (int g = 0; g < 6; ++g) {
                const int gb = g == 0 ? 0 : (g == 1 ? 9 : 18 + (g - 2) * 4);
                const int sum = v[gb] + v[gb + 1] + v[gb + 2] + v[gb + 3];
                #pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k == ref) v[gb + k] = -sum;
            }
        } else if (active && depth_out && (slot % N_POS) == CENTER) {
            depth_out[slot / N_POS] = 0;                                            // CT:461: no pileup row
        }
        #pragma unroll
        for (int ch = 0; ch < N_CH; ch += 2)
            *reinterpret_cast<uint32_t*>(o + ch) = (uint32_t)(uint16_t)v[ch] | ((uint32_t)(uint16_t)v[ch + 1] << 16);
    }
    __syncthreads();

    // coalesced 16-byte stores of the group's contiguous output block
    const int64_t out_byte0 = slot0 * N_CH * 2;
    const int64_t total_bytes = n_slots * N_CH * 2;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(s_out);
    uint8_t* dst = reinterpret_cast<uint8_t*>(tensor) + out_byte0;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
    const int64_t remain = total_bytes - out_byte0;
    const int n_bytes = (int)(remain < OUT_BYTES ? remain : OUT_BYTES);
    if (vec_ok) {
        for (int b = tid * 16; b + 16 <= n_bytes; b += THREADS * 16)
            *reinterpret_cast<uint4*>(dst + b) = *reinterpret_cast<const uint4*>(src + b);
        const int tail = n_bytes & ~15;
        if (tid < n_bytes - tail) dst[tail + tid] = src[tail + tid];
    } else {
        for (int b = tid * 2; b < n_bytes; b += THREADS * 2)
            *reinterpret_cast<uint16_t*>(dst + b) = *reinterpret_cast<const uint16_t*>(src + b);
    }
}

constexpr int SMEM_BYTES = 3 * STAGE_CAP + SLOTS * N_CH * 2 + 16;

}  // namespace enc

int launch_encode_pileup(const uint8_t* code, const uint8_t* bq, const uint8_t* mq,
                         const int32_t* pos_off, const uint8_t* ref_code, const int32_t* ind_off,
                         const uint32_t* ind_entry, const int32_t* win_pos, int64_t n_candidates,
                         int low_bq_cut, int16_t* tensor, int32_t* depth, cudaStream_t stream) {
    if (n_candidates <= 0) return 0;
    const int64_t n_slots = n_candidates * N_POS;
    const int64_t grid = (n_candidates + enc::GROUP - 1) / enc::GROUP;
    CTO_REQUIRE(grid < (1ll << 31), "encode_pileup: too many candidates in one launch (%lld)", (long long)n_candidates);
    static bool attr = false;
    if (!attr) {
        CTO_CHECK(cudaFuncSetAttribute(enc::encode_pileup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, enc::SMEM_BYTES));
        attr = true;
    }
    enc::encode_pileup_kernel<<<(unsigned)grid, enc::THREADS, enc::SMEM_BYTES, stream>>>(
        code, bq, mq, pos_off, ref_code, ind_off, ind_entry, win_pos, n_slots, low_bq_cut, tensor, depth);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
