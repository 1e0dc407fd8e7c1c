// Pileup tensor encoder: packed per-site read planes -> int16 [N, 33, 34].
//
// Replaces decode_pileup_bases() + window assembly of the reference
// (src/create_tensor_pileup_calling.py:146-229, 461, 513-516, 537-543; cited as CT).
//
// HBM-bound integer work.  Round 1 walked the reads of a row one byte at a time (three byte arrays: symbol, base
// quality, mapping quality) and was instruction-bound at 19 % of the HBM rate (~30 instructions per read).  Round 2
// counts BIT-SLICED: the host packs every read into ONE byte (clairs_to_b200/pileup_format.py: symbol nibble, "plain"
// flag, MQ >= 20 flag, MQ < 20 flag, low-BQ flag -- everything decode_pileup_bases() looks at) and stores each group of
// eight reads as eight bit-plane bytes.  A thread that owns a pileup row loads four groups (32 reads) with four 8-byte
// loads, regroups them into eight 32-bit plane words with byte permutes, forms the 26 class masks (8 bases x {MQ >= 20,
// MQ < 20, low BQ} + '*' + '#') with one LOP3 each and adds their population counts: ~4 instructions per read
// instead of ~30, and 1 byte per read over PCIe / HBM instead of 3.
//
// One CTA encodes 4 candidates = 132 (candidate, flank slot) pairs, one THREAD per slot:
//   1. the group's rows cover one contiguous span of the plane array (rows are position-sorted): one bulk async copy
//      (cp.async.bulk, completion on an mbarrier) stages it in shared memory -- full-line HBM reads, no LSU work;
//   2. every thread counts its own row from shared memory as above;
//   3. the rare indel-carrying reads come from a sparse side list and are resolved exactly (per-allele maximum,
//      CT:184-187, 201-204) with a K^2 scan that has no table-size limit;
//   4. the 34 int16 of each slot are staged in shared memory and the group's 8.8 KB output block is written with
//      16-byte coalesced stores.
// Groups whose span does not fit the staging buffer (very deep pileups, scattered rows), or whose plane array is not
// 16-byte aligned, take the same code path with the pointer left in global memory.
#include "common.cuh"

namespace cto {

namespace enc {

constexpr int GROUP = 4;                          // candidates per CTA
constexpr int SLOTS = GROUP * N_POS;              // 132
constexpr int THREADS = 160;                      // 5 warps
constexpr int STAGE_CAP = 16 * 1024;              // bytes of staged planes
constexpr int OUT_BYTES = SLOTS * N_CH * 2;       // 8976

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// the 24 base counters + '*' + '#' of reads [8 * g_lo, 8 * g_hi) of one row.  planes: 8 bytes per group of eight reads,
// byte j = bit j of the eight packed read bytes (pileup_format.py):  bits 0-3 symbol (0-3 ACGT, 4-7 acgt, 8 '*', 9 '#',
// 10 N, 11 n), bit 4 plain (a real read without an indel suffix; CT:160-204 counts indel reads only toward I/D),
// bit 5 MQ >= 20 (CT:147), bit 6 MQ < 20 (CT:148), bit 7 BQ < low-BQ cut (CT:149).  A zero byte is a null read.
template <typename P>
__device__ __forceinline__ void count_planes(P planes, int64_t byte_bias, int64_t g_lo, int64_t g_hi, int (&cnt)[26]) {
    for (int64_t g = g_lo; g < g_hi; g += 4) {
        uint2 grp[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            grp[k] = make_uint2(0u, 0u);
            if (g + k < g_hi) grp[k] = *reinterpret_cast<const uint2*>(planes + ((g + k) * 8 - byte_bias));
        }
        // plane word j = byte j of the four groups: 32 reads per word
        uint32_t pl[8];
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t sel = (uint32_t)j | ((uint32_t)(4 + j) << 4);
            pl[j] = __byte_perm(__byte_perm(grp[0].x, grp[1].x, sel), __byte_perm(grp[2].x, grp[3].x, sel), 0x5410);
            pl[4 + j] = __byte_perm(__byte_perm(grp[0].y, grp[1].y, sel), __byte_perm(grp[2].y, grp[3].y, sel), 0x5410);
        }
        const uint32_t f_hi = pl[4] & pl[5];                  // plain, MQ >= 20   -> A C G T a c g t * #
        const uint32_t f_lo = pl[4] & pl[6];                  // plain, MQ < 20    -> LMQ channels (CT:215-217)
        const uint32_t f_bq = pl[4] & pl[7];                  // plain, low BQ     -> LBQ channels (CT:219-221)
        const uint32_t base = ~pl[3];                         // symbols 0..7
        const uint32_t e0 = base & ~pl[1] & ~pl[0], e1 = base & ~pl[1] & pl[0], e2 = base & pl[1] & ~pl[0], e3 = base & pl[1] & pl[0];
        const uint32_t fw = ~pl[2], rv = pl[2];
        const uint32_t eb[4] = {e0, e1, e2, e3};
        #pragma unroll
        for (int b = 0; b < 4; ++b) {
            cnt[b] += __popc(f_hi & fw & eb[b]);
            cnt[4 + b] += __popc(f_hi & rv & eb[b]);
            cnt[10 + b] += __popc(f_lo & fw & eb[b]);
            cnt[14 + b] += __popc(f_lo & rv & eb[b]);
            cnt[18 + b] += __popc(f_bq & fw & eb[b]);
            cnt[22 + b] += __popc(f_bq & rv & eb[b]);
        }
        const uint32_t star_hash = f_hi & pl[3] & ~pl[2] & ~pl[1];
        cnt[8] += __popc(star_hash & ~pl[0]);
        cnt[9] += __popc(star_hash & pl[0]);
    }
}

// set-A field -> channel (CT:55-58)
__device__ __forceinline__ int channel_of_a(int f) {
    if (f < 4) return f;                // A C G T
    if (f < 8) return f + 5;            // a c g t  -> 9..12
    if (f == 8) return 8;               // '*'
    if (f == 9) return 17;              // '#'
    return f + 8;                       // LMQ      -> 18..25
}

__global__ void __launch_bounds__(THREADS)
encode_pileup_kernel(const uint8_t* __restrict__ planes, const int32_t* __restrict__ grp_off,
                     const uint8_t* __restrict__ ref_code, const int32_t* __restrict__ ind_off,
                     const uint32_t* __restrict__ ind_entry, const int32_t* __restrict__ win_pos,
                     int64_t n_slots, int64_t plane_bytes, int stage_ok, int16_t* __restrict__ tensor,
                     int32_t* __restrict__ depth_out) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_planes = smem;
    int16_t* s_out = reinterpret_cast<int16_t*>(smem + STAGE_CAP);   // [SLOTS][34]
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
        span_lo = (int64_t)grp_off[min_row] * 8;
        span_hi = (int64_t)grp_off[max_row + 1] * 8;
    }
    const int64_t a_lo = span_lo & ~int64_t(15);                      // 16-byte aligned superset of the span,
    int64_t a_hi = (span_hi + 15) & ~int64_t(15);                     // clamped to the (16-byte padded) array
    if (a_hi > plane_bytes) a_hi = plane_bytes;
    const bool staged = stage_ok && (a_hi - a_lo) <= STAGE_CAP && a_hi > a_lo && span_hi <= a_hi;
    if (staged && tid == 0) {
        const uint32_t bytes = (uint32_t)(a_hi - a_lo);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&s_bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(s_planes)), "l"(planes + a_lo), "r"(bytes), "r"(smem_u32(&s_bar)) : "memory");
    }

    // per-slot metadata and the sparse indel list are read while the bulk copy is in flight
    int64_t g_lo = 0, g_hi = 0;
    int ref = 0;
    int tot[4] = {0, 0, 0, 0}, best[4] = {0, 0, 0, 0};
    if (row >= 0) {
        g_lo = grp_off[row];
        g_hi = grp_off[row + 1];
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
    if (staged) count_planes(s_planes, a_lo, g_lo, g_hi, cnt);        // the pointer type keeps the address space: LDS
    else count_planes(planes, (int64_t)0, g_lo, g_hi, cnt);

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

constexpr int SMEM_BYTES = STAGE_CAP + SLOTS * N_CH * 2 + 16;

}  // namespace enc

// planes: 8 * n_groups bytes of bit planes; the ALLOCATION must extend to the next multiple of 16 bytes (the staged path
// copies 16-byte aligned supersets of a span) -- clairs_to_b200.pileup_format.pack_stream and the host call guarantee it.
// A plane array that is not 16-byte aligned (an offset view) is read with ordinary loads instead of bulk copies.
int launch_encode_pileup(const uint8_t* planes, const int32_t* grp_off, const uint8_t* ref_code, const int32_t* ind_off,
                         const uint32_t* ind_entry, const int32_t* win_pos, int64_t n_candidates, int64_t n_groups,
                         int16_t* tensor, int32_t* depth, cudaStream_t stream) {
    if (n_candidates <= 0) return 0;
    const int64_t n_slots = n_candidates * N_POS;
    const int64_t grid = (n_candidates + enc::GROUP - 1) / enc::GROUP;
    CTO_REQUIRE(grid < (1ll << 31), "encode_pileup: too many candidates in one launch (%lld)", (long long)n_candidates);
    CTO_REQUIRE(n_groups >= 0 && n_groups < (1ll << 31), "encode_pileup: group count %lld out of range", (long long)n_groups);
    CTO_REQUIRE((reinterpret_cast<uintptr_t>(planes) & 7) == 0, "encode_pileup: the plane array must be 8-byte aligned");
    const int stage_ok = (reinterpret_cast<uintptr_t>(planes) & 15) == 0 ? 1 : 0;
    const int64_t plane_bytes = (n_groups * 8 + 15) & ~int64_t(15);
    CTO_CHECK(set_max_dynamic_smem(enc::encode_pileup_kernel, enc::SMEM_BYTES));
    enc::encode_pileup_kernel<<<(unsigned)grid, enc::THREADS, enc::SMEM_BYTES, stream>>>(
        planes, grp_off, ref_code, ind_off, ind_entry, win_pos, n_slots, plane_bytes, stage_ok, tensor, depth);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // namespace cto
