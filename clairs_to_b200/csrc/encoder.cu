// Pileup tensor encoder: per-site read arrays -> int16 [N, 33, 34].
//
// Replaces decode_pileup_bases() + window assembly of the reference
// (src/create_tensor_pileup_calling.py:146-229, 461, 513-516, 537-543; cited as CT).
// HBM-bound integer work: one warp per (candidate, flank slot); reads are fetched with
// coalesced byte loads, classified into at most two packed 8-bit counters per read and
// summed across the warp with REDUX (no atomics, no shared-memory contention).  The rare
// indel-carrying reads come from a sparse side list and are resolved exactly (per-allele
// maximum, CT:184-187, 201-204) with an O(K^2/32) scan that has no table-size limit.
#include "common.cuh"

namespace cto {

// fields of the two packed counter sets
//   set A (exclusive by MQ): 0-3 ACGT, 4-7 acgt, 8 '*', 9 '#', 10-13 ACGT-LMQ, 14-17 acgt-LMQ
//   set B (BQ < cut):        0-3 ACGT-LBQ, 4-7 acgt-LBQ
__device__ __forceinline__ int field_to_channel_a(int f) {
    if (f < 4) return f;                // A C G T
    if (f < 8) return f + 5;            // a c g t  -> 9..12
    if (f == 8) return 8;               // '*'
    if (f == 9) return 17;              // '#'
    return f + 8;                       // LMQ      -> 18..25
}

constexpr int WARPS_PER_CTA = 8;

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
encode_pileup_kernel(const uint8_t* __restrict__ code, const uint8_t* __restrict__ bq,
                     const uint8_t* __restrict__ mq, const int32_t* __restrict__ pos_off,
                     const uint8_t* __restrict__ ref_code, const int32_t* __restrict__ ind_off,
                     const uint32_t* __restrict__ ind_entry, const int32_t* __restrict__ win_pos,
                     int64_t n_slots, int low_bq_cut, int16_t* __restrict__ tensor,
                     int32_t* __restrict__ depth_out) {
    __shared__ int s_cnt[WARPS_PER_CTA][N_CH + 2];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t slot = (int64_t)blockIdx.x * WARPS_PER_CTA + wib;
    if (slot >= n_slots) return;
    const int row = win_pos[slot];
    int16_t* out = tensor + slot * N_CH;
    const bool is_center = (slot % N_POS) == CENTER;
    if (row < 0) {                                     // CT:461: no pileup row -> zeros
        out[lane] = 0;
        if (lane < N_CH - 32) out[32 + lane] = 0;
        if (is_center && lane == 0 && depth_out) depth_out[slot / N_POS] = 0;
        return;
    }
    const int lo = pos_off[row], hi = pos_off[row + 1];
    const int ref = ref_code[row];

    // 16-bit-field accumulators, identical in every lane: a_lo/a_hi hold even/odd bytes of
    // the five set-A words, b_* of the two set-B words.
    uint32_t a_lo[5] = {0, 0, 0, 0, 0}, a_hi[5] = {0, 0, 0, 0, 0};
    uint32_t b_lo[2] = {0, 0}, b_hi[2] = {0, 0};

    for (int base = lo; base < hi; base += 32) {
        const int i = base + lane;
        int fa = -1, fb = -1;
        if (i < hi) {
            const uint32_t c = code[i];
            const uint32_t q = bq[i];
            const uint32_t m = mq[i];
            const uint32_t sym = c & 0xF;
            if (!(c & 0x10)) {                                  // plain read (CT:160-171)
                int b8 = -1;
                if (sym < 4) b8 = sym;
                else if (sym >= 5 && sym <= 8) b8 = sym - 1;
                if (m != QUAL_ABSENT) {
                    if (m >= MIN_MQ) {
                        if (b8 >= 0) fa = b8;
                        else if (sym == 10) fa = 8;
                        else if (sym == 11) fa = 9;
                    } else if (b8 >= 0) {
                        fa = 10 + b8;                           // CT:215-217
                    }
                }
                if (b8 >= 0 && q != QUAL_ABSENT && (int)q < low_bq_cut) fb = b8;   // CT:149, 219-221
            }
        }
        const uint32_t one_a = fa >= 0 ? (1u << ((fa & 3) * 8)) : 0u;
        const int wa = fa >> 2;
        #pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint32_t s = __reduce_add_sync(0xffffffffu, wa == k ? one_a : 0u);
            a_lo[k] += s & 0x00FF00FFu;
            a_hi[k] += (s >> 8) & 0x00FF00FFu;
        }
        const uint32_t one_b = fb >= 0 ? (1u << ((fb & 3) * 8)) : 0u;
        const int wb = fb >> 2;
        #pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t s = __reduce_add_sync(0xffffffffu, wb == k ? one_b : 0u);
            b_lo[k] += s & 0x00FF00FFu;
            b_hi[k] += (s >> 8) & 0x00FF00FFu;
        }
    }

    // sparse indel list: totals per class (ins/del x fwd/rev) and per-allele maxima
    const int ilo = ind_off[row], ihi = ind_off[row + 1];
    int tot[4] = {0, 0, 0, 0}, best[4] = {0, 0, 0, 0};
    for (int base = ilo; base < ihi; base += 32) {
        const int i = base + lane;
        uint32_t e = 0;
        bool ok = false;
        if (i < ihi) {
            e = ind_entry[i];
            const uint32_t m = (e >> 16) & 0xFF;
            ok = !(e & (1u << 26)) && m >= MIN_MQ && m != QUAL_ABSENT;    // CT:147, 174-176, 189-191
        }
        const int cls = (e >> 24) & 3;                  // bit0 deletion, bit1 reverse
        const uint32_t key = e & 0x0300FFFFu;
        int same = 0;
        for (int j = ilo; j < ihi; ++j) {               // uniform (broadcast) loads
            const uint32_t o = ind_entry[j];
            const uint32_t om = (o >> 16) & 0xFF;
            const bool ook = !(o & (1u << 26)) && om >= MIN_MQ && om != QUAL_ABSENT;
            same += (ook && (o & 0x0300FFFFu) == key) ? 1 : 0;
        }
        #pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
            const bool mine = ok && cls == c4;
            tot[c4] += __popc(__ballot_sync(0xffffffffu, mine));
            best[c4] = max(best[c4], (int)__reduce_max_sync(0xffffffffu, mine ? (unsigned)same : 0u));
        }
    }

    int* cnt = s_cnt[wib];
    // unpack: lane f < 18 owns set-A field f, lanes 18..25 own set-B field f-18
    {
        int f = lane, v = 0, ch = -1;
        if (lane < 18) {
            uint32_t lo16 = 0, hi16 = 0;
            #pragma unroll
            for (int k = 0; k < 5; ++k) if ((f >> 2) == k) { lo16 = a_lo[k]; hi16 = a_hi[k]; }
            const uint32_t w = (f & 1) ? hi16 : lo16;
            v = (f & 2) ? (w >> 16) : (w & 0xFFFF);
            ch = field_to_channel_a(f);
        } else if (lane < 26) {
            f = lane - 18;
            uint32_t lo16 = 0, hi16 = 0;
            #pragma unroll
            for (int k = 0; k < 2; ++k) if ((f >> 2) == k) { lo16 = b_lo[k]; hi16 = b_hi[k]; }
            const uint32_t w = (f & 1) ? hi16 : lo16;
            v = (f & 2) ? (w >> 16) : (w & 0xFFFF);
            ch = 26 + f;
        } else if (lane < 30) {                         // I, D, i, d totals
            const int c4 = lane - 26;                   // 0 ins fwd, 1 del fwd, 2 ins rev, 3 del rev
            v = c4 == 0 ? tot[0] : c4 == 1 ? tot[1] : c4 == 2 ? tot[2] : tot[3];
            ch = (c4 & 2 ? 13 : 4) + (c4 & 1 ? 2 : 0);
        }
        if (ch >= 0) cnt[ch] = v;
        if (lane < 4) {                                 // I1, D1, i1, d1 (CT:210-213)
            const int ch1 = (lane & 2 ? 14 : 5) + (lane & 1 ? 2 : 0);
            cnt[ch1] = lane == 0 ? best[0] : lane == 1 ? best[1] : lane == 2 ? best[2] : best[3];
        }
    }
    __syncwarp();

    if (is_center && depth_out) {                       // depth of alt_info (CT:167-195, 208)
        int d = 0;
        if (lane < 18) {
            const int ch = lane;
            const bool counted = ch != 5 && ch != 7 && ch != 14 && ch != 16;
            d = counted ? cnt[ch] : 0;
        }
        d = __reduce_add_sync(0xffffffffu, d);
        if (lane == 0) depth_out[slot / N_POS] = d;
    }

    // reference channel := -(sum of the four base counts) in each of the six groups (CT:223-228)
    #pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        const int ch = lane + rep * 32;
        if (ch < N_CH) {
            int v = cnt[ch];
            int g = -1;
            if (ch < 4) g = 0;
            else if (ch >= 9 && ch < 13) g = 9;
            else if (ch >= 18) g = 18 + ((ch - 18) & ~3);
            if (g >= 0 && ch - g == ref) v = -(cnt[g] + cnt[g + 1] + cnt[g + 2] + cnt[g + 3]);
            out[ch] = (int16_t)v;
        }
    }
}

int launch_encode_pileup(const uint8_t* code, const uint8_t* bq, const uint8_t* mq,
                         const int32_t* pos_off, const uint8_t* ref_code, const int32_t* ind_off,
                         const uint32_t* ind_entry, const int32_t* win_pos, int64_t n_candidates,
                         int low_bq_cut, int16_t* tensor, int32_t* depth, cudaStream_t stream) {
    if (n_candidates <= 0) return 0;
    const int64_t n_slots = n_candidates * N_POS;
    const int64_t grid = (n_slots + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    CTO_REQUIRE(grid < (1ll << 31), "encode_pileup: too many candidates in one launch (%lld)", (long long)n_candidates);
    encode_pileup_kernel<<<(unsigned)grid, WARPS_PER_CTA * 32, 0, stream>>>(
        code, bq, mq, pos_off, ref_code, ind_off, ind_entry, win_pos, n_slots, low_bq_cut, tensor, depth);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
