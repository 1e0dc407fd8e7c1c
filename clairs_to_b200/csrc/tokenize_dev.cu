// mpileup text -> the encoder's packed input, on the GPU.
//
// Device twin of the host tokenizer + packer (host_codec.cpp: cto_tokenize_mpileup + cto_pack_reads), which restate the row
// split and the tokenizer of the reference, src/create_tensor_pileup_calling.py:472-497 and 120-144 (cited as CT).  The host
// pair runs at ~1 GB/s per box and bounded every "from mpileup text" number; here the text is copied to the device as it came
// out of samtools and never touched by the host:
//
//   cto_index_rows            row offsets (candidates.cu)
//   tok_rows_kernel<COUNT>    one thread per row: columns, position, reads per row, indel-carrying reads per row
//   cub::DeviceScan           groups-of-eight and indel offsets (exclusive sums)
//   tok_rows_kernel<WRITE>    one thread per row again: one packed byte per read (symbol, plain / indel, MQ class, low BQ),
//                             eight reads transposed into eight bit-plane bytes in registers, one 8-byte store per group;
//                             the side list of indel-carrying reads with per-row allele ids (alleles compared on the text)
//   window_table_kernel       candidate position -> the 33 row indices of its window (binary search; -1 = no pileup row)
//
// The state machine is the host's, hence the reference's: `+N<seq>` / `-N<seq>` attach to (and overwrite on) the previous read,
// `^x` skips two characters, characters outside ACGTNacgtn*# produce no read, read k takes the k-th character of the BQ / MQ
// columns (positional zip, CT:147-149: a short quality column leaves the later reads without quality).
#include "../../include/clairs_to_b200.h"
#include "common.cuh"
#include <cub/device/device_scan.cuh>

namespace cto {
namespace tokd {

constexpr int TB = 128;                  // rows per block
constexpr int MAX_STAGE = 96 * 1024;
constexpr int TABLE = 24;                // distinct indel alleles per row in the first pass
constexpr int BIG_TABLE = 8192;          // ... for the rows that overflow it (global memory)
constexpr uint8_t QUAL_ABSENT = 254;
constexpr uint32_t IND_DEL = 1u << 24, IND_REV = 1u << 25, IND_LONG = 1u << 26;
enum { E_COLUMNS = 1, E_REF = 2, E_EMPTY = 4, E_OVERFLOW = 8 };

struct Allele { uint32_t off, len, hash; uint8_t sym, sign; };

struct GlobalText {
    const uint8_t* p;
    __device__ __forceinline__ uint8_t operator[](int64_t i) const { return __ldg(p + i); }
};
struct StagedText {
    const uint8_t* s;
    int64_t bias;
    __device__ __forceinline__ uint8_t operator[](int64_t i) const { return s[i - bias]; }
};

__device__ __forceinline__ int symbol_code(uint8_t c) {       // packed order: ACGT acgt * # N n (host_codec.cpp RECODE); c is a read symbol
    // arithmetic, not a chain of `c == 'A' ? 0 : ...`: the compiler turns such a chain into one branch per base, which splits the
    // lanes of a warp by base identity (measured on scan_kernel)
    const uint32_t lc = c | 0x20u, k5 = lc - 'a';
    const uint32_t is_base = k5 < 26u ? (0x80045u >> k5) & 1u : 0u;              // a, c, g, t
    const uint32_t code = (lc >> 1) & 3u;                                        // a 0, c 1, t 2, g 3
    const int base = (int)((code ^ (code >> 1)) + ((c & 0x20u) >> 3));           // A 0 C 1 G 2 T 3, lower case + 4
    const int other = c == '*' ? 8 : c == '#' ? 9 : c == 'N' ? 10 : 11;
    return is_base ? base : other;
}
__device__ __forceinline__ bool is_symbol(uint8_t c) {         // one of ACGTNacgtn*# (CT:140)
    const uint32_t k = (c | 0x20u) - 'a';                      // a = 0, c = 2, g = 6, n = 13, t = 19
    return (k < 26u && ((0x82045u >> k) & 1u)) || c == '*' || c == '#';
}
__device__ __forceinline__ uint64_t transpose8x8(uint64_t x) {
    uint64_t t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAULL;  x = x ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCULL; x = x ^ t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ULL; x = x ^ t ^ (t << 28);
    return x;
}

struct Out {
    int32_t* row_pos;        // [n_rows]
    uint8_t* ref_code;       // [n_rows]
    int32_t* grp_off;        // [n_rows + 1]: COUNT writes groups per row, the scan turns them into offsets
    int32_t* ind_off;        // [n_rows + 1]: likewise for indel-carrying reads
    uint8_t* planes;         // WRITE
    uint32_t* ind_entry;     // WRITE
    int32_t* error;          // [2]: flag bits, first failing row
};

// Id of the indel allele (symbol, sign, exact sequence) within its row, in order of first appearance (host_codec.cpp:
// allele_ids).  Alleles are compared on the text itself.  Rare, hence out of line: the hot loop stays small and converged.
template <typename TextPtr>
__device__ __noinline__ int allele_id(TextPtr t, Allele* table, int cap, int& n_alleles, int& err, uint8_t cur, uint8_t sign, int64_t seq_off,
                                      int64_t seq_len) {
    uint32_t h = 2166136261u ^ cur;
    h = (h ^ sign) * 16777619u;
    for (int64_t z = 0; z < seq_len; ++z) h = (h ^ t[seq_off + z]) * 16777619u;
    int e = 0;
    for (; e < n_alleles; ++e) {
        const Allele& a = table[e];
        if (a.hash != h || a.len != (uint32_t)seq_len || a.sym != cur || a.sign != sign) continue;
        int64_t z = 0;
        while (z < seq_len && t[(int64_t)a.off + z] == t[seq_off + z]) ++z;
        if (z == seq_len) break;
    }
    if (e == n_alleles) {
        if (n_alleles < cap) {
            Allele& a = table[n_alleles++];
            a.off = (uint32_t)seq_off; a.len = (uint32_t)seq_len; a.hash = h; a.sym = cur; a.sign = sign;
        } else {
            err |= E_OVERFLOW;
        }
    }
    return e;
}

// One row.  WRITE = false: counts only.  Returns error bits.
template <bool WRITE, typename TextPtr>
__device__ int tok_row(TextPtr t, int64_t lo, int64_t hi, int64_t r, const uint8_t* __restrict__ ref, int64_t ref_start, int64_t ref_len,
                       int low_bq_cut, int max_indel_length, Allele* table, int cap, const Out& o) {
    int64_t eol = hi;
    if (eol > lo && t[eol - 1] == '\n') --eol;
    int64_t c_lo[8], c_len[8];
    int ncol = 0;
    for (int64_t q = lo; q <= eol && ncol < 8;) {             // host: memchr for tabs, at most 8 columns
        int64_t tab = q;
        while (tab < eol && t[tab] != '\t') ++tab;
        c_lo[ncol] = q; c_len[ncol] = tab - q; ++ncol;
        q = tab + 1;
    }
    if (ncol == 0 || (ncol == 1 && c_len[0] == 0)) return E_EMPTY;
    if (ncol < 7) return E_COLUMNS;
    while (c_len[6] > 0 && (t[c_lo[6] + c_len[6] - 1] == '\r' || t[c_lo[6] + c_len[6] - 1] == ' ')) --c_len[6];
    int64_t pos = 0;
    for (int64_t k = c_lo[1]; k < c_lo[1] + c_len[1] && t[k] >= '0' && t[k] <= '9'; ++k) pos = pos * 10 + (t[k] - '0');
    const int64_t roff = pos - ref_start;
    if (roff < 0 || roff >= ref_len) return E_REF;
    if (!WRITE) {
        o.row_pos[r] = (int32_t)pos;
        uint8_t rc = ref[roff];
        if (rc >= 'a' && rc <= 'z') rc -= 32;
        o.ref_code[r] = rc == 'C' ? 1 : rc == 'G' ? 2 : rc == 'T' ? 3 : 0;     // evc_base_from: everything else becomes A (CT:82-92)
    }
    const int64_t b0 = c_lo[4], nb = c_len[4];
    const int64_t n_bq = c_len[5], n_mq = c_len[6];
    int64_t k = -1;                                             // index of the read being assembled
    int n_ind = 0, n_alleles = 0, err = 0;
    uint8_t cur = 0, sign = 0;
    int cur_code = 0;
    int64_t seq_off = 0, seq_len = 0;
    uint64_t x = 0;
    const int64_t g0 = WRITE ? o.grp_off[r] : 0;
    const int64_t i0 = WRITE ? o.ind_off[r] : 0;
    // Indel-carrying reads are rare per row but frequent per warp; handled where they occur they run on ONE lane.  They are
    // queued (in read order: allele ids count first appearances) and turned into side-list entries behind the loop, where all
    // lanes work on their queues at the same time; a deep row drains its queue whenever it is full.
    constexpr int QUEUE = 8;
    uint32_t q_off[QUEUE], q_len[QUEUE], q_meta[QUEUE];
    int n_q = 0;
    auto drain = [&]() {
        for (int e = 0; e < n_q; ++e) {
            const uint8_t sym = (uint8_t)(q_meta[e] & 0xff), sg = (uint8_t)((q_meta[e] >> 8) & 0xff), m = (uint8_t)(q_meta[e] >> 16);
            const int64_t off = lo + q_off[e], len = q_len[e];
            const int id = allele_id(t, table, cap, n_alleles, err, sym, sg, off, len);
            uint32_t ent = ((uint32_t)id & 0xFFFFu) | ((uint32_t)m << 16);
            const bool is_del = sg == '-';
            if (is_del) ent |= IND_DEL;
            const bool fwd = sym == 'A' || sym == 'C' || sym == 'G' || sym == 'T' || sym == 'N' || sym == '*';   // CT:182, 199
            if (!fwd) ent |= IND_REV;
            if ((is_del ? len + 1 : len) > max_indel_length) ent |= IND_LONG;                                    // CT:174, 189
            o.ind_entry[i0 + n_ind] = ent;
            ++n_ind;
        }
        n_q = 0;
    };
    auto finish = [&]() {                                       // the read `k` is complete: its suffix cannot change any more
        if (k < 0) return;
        if (!WRITE) { n_ind += sign != 0; return; }
        const uint8_t m = k < n_mq ? (uint8_t)(t[c_lo[6] + k] - 33) : QUAL_ABSENT;
        const uint8_t q = k < n_bq ? (uint8_t)(t[c_lo[5] + k] - 33) : QUAL_ABSENT;
        uint8_t b = (uint8_t)cur_code;
        if (!sign) b |= 0x10;
        if (m != QUAL_ABSENT) b |= (m >= 20) ? 0x20 : 0x40;     // CT:147-148
        if (q != QUAL_ABSENT && (int)q < low_bq_cut) b |= 0x80;  // CT:149
        x |= (uint64_t)b << (8 * (k & 7));
        if ((k & 7) == 7) {
            *reinterpret_cast<uint64_t*>(o.planes + (g0 + (k >> 3)) * 8) = transpose8x8(x);
            x = 0;
        }
        if (sign) {                                             // side list entry: queued, written behind the loop (see below)
            if (n_q == QUEUE) drain();
            q_off[n_q] = (uint32_t)(seq_off - lo); q_len[n_q] = (uint32_t)seq_len; q_meta[n_q] = (uint32_t)cur | ((uint32_t)sign << 8) | ((uint32_t)m << 16);
            ++n_q;
        }
    };
    // CT:120-144.  One more iteration than the column has characters: the virtual character behind it completes the last read,
    // so that finish() is inlined ONCE and the loop body stays small.
    for (int64_t i = 0; i <= nb;) {
        const bool at_end = i == nb;
        const uint8_t ch = at_end ? (uint8_t)'A' : t[b0 + i];
        if ((ch == '+' || ch == '-') && !at_end) {
            int64_t j = i + 1, len = 0;
            while (j < nb && t[b0 + j] >= '0' && t[b0 + j] <= '9') { len = len * 10 + (t[b0 + j] - '0'); ++j; }
            if (k >= 0) {
                const int64_t avail = nb - j > 0 ? nb - j : 0;
                sign = ch; seq_off = b0 + j; seq_len = len < avail ? len : avail;
            }
            i = j + len;
            if (i > nb) i = nb;                                 // a length that runs past the column: the column ends here
            continue;
        }
        if (is_symbol(ch)) {
            finish();
            if (at_end) break;
            ++k; cur = ch; sign = 0;
            if (WRITE) cur_code = symbol_code(ch);
        } else if (ch == '^') {
            ++i;
            if (i >= nb) i = nb - 1;                            // `^` as the last character: nothing behind it to skip
        }
        ++i;
    }
    if (WRITE) drain();
    const int64_t n = k + 1;
    if (!WRITE) {
        o.grp_off[r] = (int32_t)((n + 7) >> 3);
        o.ind_off[r] = n_ind;
    } else if (n & 7) {
        *reinterpret_cast<uint64_t*>(o.planes + (g0 + (n >> 3)) * 8) = transpose8x8(x);
    }
    return err;
}

template <bool WRITE>
__global__ void __launch_bounds__(TB)
tok_rows_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ row_off, int64_t n_rows, const uint8_t* __restrict__ ref,
                int64_t ref_start, int64_t ref_len, int low_bq_cut, int max_indel_length, int stage_bytes, Out o,
                int32_t* __restrict__ overflow_rows, int32_t* __restrict__ overflow_count) {
    extern __shared__ __align__(16) uint8_t stage[];
    const int64_t r0 = (int64_t)blockIdx.x * TB;
    const int64_t r1 = r0 + TB < n_rows ? r0 + TB : n_rows;
    const int64_t span_lo = row_off[r0] & ~int64_t(15), span_hi = row_off[r1];
    const bool staged = span_hi - span_lo <= stage_bytes && (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    if (staged) {
        const uint4* src = reinterpret_cast<const uint4*>(text + span_lo);
        const int64_t n16 = (span_hi - span_lo + 15) >> 4;
        for (int64_t k = threadIdx.x; k < n16; k += TB) reinterpret_cast<uint4*>(stage)[k] = __ldg(src + k);
    }
    __syncthreads();
    const int64_t r = r0 + threadIdx.x;
    if (r >= r1) return;
    Allele table[TABLE];
    int err;
    if (staged)
        err = tok_row<WRITE>(StagedText{stage, span_lo}, row_off[r], row_off[r + 1], r, ref, ref_start, ref_len, low_bq_cut, max_indel_length,
                             table, TABLE, o);
    else
        err = tok_row<WRITE>(GlobalText{text}, row_off[r], row_off[r + 1], r, ref, ref_start, ref_len, low_bq_cut, max_indel_length, table,
                             TABLE, o);
    if (WRITE && (err & E_OVERFLOW)) {
        overflow_rows[atomicAdd(overflow_count, 1)] = (int32_t)r;
        err &= ~E_OVERFLOW;
    }
    if (err) {
        atomicOr(o.error, err);
        atomicMin(o.error + 1, (int32_t)r);
    }
}

// rows with more than TABLE distinct indel alleles: again, with the table in global memory
__global__ void __launch_bounds__(TB)
tok_big_rows_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ row_off, const int32_t* __restrict__ rows, int n_list,
                    const uint8_t* __restrict__ ref, int64_t ref_start, int64_t ref_len, int low_bq_cut, int max_indel_length,
                    Allele* __restrict__ scratch, Out o) {
    const int k = blockIdx.x * TB + threadIdx.x;
    if (k >= n_list) return;
    const int64_t r = rows[k];
    const int err = tok_row<true>(GlobalText{text}, row_off[r], row_off[r + 1], r, ref, ref_start, ref_len, low_bq_cut, max_indel_length,
                                  scratch + (int64_t)k * BIG_TABLE, BIG_TABLE, o);
    if (err) {
        atomicOr(o.error, err);
        atomicMin(o.error + 1, (int32_t)r);
    }
}

// win_pos[c][s] = index of the pileup row at position cand[c] - 16 + s, or -1 (CT:461: rows samtools did not print are zero)
__global__ void __launch_bounds__(256)
window_table_kernel(const int32_t* __restrict__ row_pos, int64_t n_rows, const int64_t* __restrict__ cand, int64_t n_cand,
                    int32_t* __restrict__ win_pos) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_cand * N_POS) return;
    const int64_t want = cand[idx / N_POS] - CENTER + (idx % N_POS);
    int64_t lo = 0, hi = n_rows;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (row_pos[mid] < want) lo = mid + 1; else hi = mid;
    }
    win_pos[idx] = (lo < n_rows && row_pos[lo] == want) ? (int32_t)lo : -1;
}

static int stage_bytes_for(int64_t text_len, int64_t n_rows) {
    int64_t stage = (text_len / n_rows + 1) * TB * 3 / 2 + 1024;
    stage = (stage + 1023) & ~int64_t(1023);
    if (stage < 8 * 1024) stage = 8 * 1024;
    if (stage > MAX_STAGE) stage = MAX_STAGE;
    return (int)stage;
}

static int report(const int32_t* err_host, const char* what) {
    if (!err_host[0]) return 0;
    const char* why = (err_host[0] & E_COLUMNS) ? "a row with fewer than 7 columns (need chr pos ref depth bases BQ MQ)"
                      : (err_host[0] & E_REF)   ? "a position outside the reference window"
                      : (err_host[0] & E_EMPTY) ? "an empty line"
                                                : "more than 8192 distinct indel alleles in one row";
    set_error("%s: %s (first at row %d)", what, why, err_host[1]);
    return 2;
}

}  // namespace tokd
}  // namespace cto

using namespace cto;

extern "C" {

int cto_tokenize_count(const uint8_t* text_dev, int64_t text_len, const int64_t* row_off_dev, int64_t n_rows, const uint8_t* ref_dev,
                       int64_t ref_start, int64_t ref_len, int32_t* row_pos_dev, uint8_t* ref_code_dev, int32_t* grp_off_dev,
                       int32_t* ind_off_dev, int64_t* n_groups, int64_t* n_ind, void* stream) {
    CTO_REQUIRE(n_groups && n_ind, "tokenize_count: NULL size output");
    *n_groups = *n_ind = 0;
    CTO_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "tokenize_count: %lld rows", (long long)n_rows);
    CTO_REQUIRE(grp_off_dev && ind_off_dev, "tokenize_count: NULL offset array");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_rows == 0) {
        CTO_CHECK(cudaMemsetAsync(grp_off_dev, 0, sizeof(int32_t), s));
        CTO_CHECK(cudaMemsetAsync(ind_off_dev, 0, sizeof(int32_t), s));
        return 0;
    }
    CTO_REQUIRE(text_dev && row_off_dev && ref_dev && row_pos_dev && ref_code_dev, "tokenize_count: NULL array");
    CTO_REQUIRE(text_len < (1ll << 32), "tokenize_count: %lld bytes of text in one call (limit 4 GiB)", (long long)text_len);
    if (cto_device_check(nullptr)) return 3;
    int32_t* err = nullptr;
    size_t tmp_bytes = 0;
    CTO_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, grp_off_dev, grp_off_dev, (int)(n_rows + 1), s));
    uint8_t* scratch = nullptr;
    CTO_CHECK(scratch_alloc((void**)&scratch, tmp_bytes + 256, s));
    err = reinterpret_cast<int32_t*>(scratch);
    const int32_t err_init[2] = {0, INT32_MAX};
    CTO_CHECK(cudaMemcpyAsync(err, err_init, sizeof(err_init), cudaMemcpyHostToDevice, s));
    CTO_CHECK(cudaMemsetAsync(grp_off_dev + n_rows, 0, sizeof(int32_t), s));
    CTO_CHECK(cudaMemsetAsync(ind_off_dev + n_rows, 0, sizeof(int32_t), s));
    const int stage = tokd::stage_bytes_for(text_len, n_rows);
    CTO_CHECK(set_max_dynamic_smem(tokd::tok_rows_kernel<false>, tokd::MAX_STAGE));
    tokd::Out o{row_pos_dev, ref_code_dev, grp_off_dev, ind_off_dev, nullptr, nullptr, err};
    tokd::tok_rows_kernel<false><<<(unsigned)ceil_div(n_rows, tokd::TB), tokd::TB, stage, s>>>(text_dev, row_off_dev, n_rows, ref_dev, ref_start,
                                                                                               ref_len, 0, 0, stage, o, nullptr, nullptr);
    CTO_CHECK(cudaGetLastError());
    CTO_CHECK(cub::DeviceScan::ExclusiveSum(scratch + 256, tmp_bytes, grp_off_dev, grp_off_dev, (int)(n_rows + 1), s));
    CTO_CHECK(cub::DeviceScan::ExclusiveSum(scratch + 256, tmp_bytes, ind_off_dev, ind_off_dev, (int)(n_rows + 1), s));
    count_launch(3);
    int32_t err_host[2] = {0, 0}, totals[2] = {0, 0};
    CTO_CHECK(cudaMemcpyAsync(err_host, err, sizeof(err_host), cudaMemcpyDeviceToHost, s));
    CTO_CHECK(cudaMemcpyAsync(&totals[0], grp_off_dev + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CTO_CHECK(cudaMemcpyAsync(&totals[1], ind_off_dev + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CTO_CHECK(cudaFreeAsync(scratch, s));
    CTO_CHECK(cudaStreamSynchronize(s));
    if (int rc = tokd::report(err_host, "tokenize_count")) return rc;
    *n_groups = totals[0];
    *n_ind = totals[1];
    return 0;
}

int cto_tokenize_write(const uint8_t* text_dev, int64_t text_len, const int64_t* row_off_dev, int64_t n_rows, const uint8_t* ref_dev,
                       int64_t ref_start, int64_t ref_len, int low_bq_cut, int max_indel_length, const int32_t* grp_off_dev,
                       const int32_t* ind_off_dev, uint8_t* planes_dev, uint32_t* ind_entry_dev, void* stream) {
    CTO_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "tokenize_write: %lld rows", (long long)n_rows);
    if (n_rows == 0) return 0;
    CTO_REQUIRE(text_dev && row_off_dev && ref_dev && grp_off_dev && ind_off_dev && planes_dev && ind_entry_dev, "tokenize_write: NULL array");
    CTO_REQUIRE((reinterpret_cast<uintptr_t>(planes_dev) & 7) == 0, "tokenize_write: planes must be 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    int32_t* over = nullptr;                                     // [0] counter, [1..2] error words, [4..] overflow row list
    CTO_CHECK(scratch_alloc((void**)&over, sizeof(int32_t) * (size_t)(n_rows + 4), s));
    const int32_t init[3] = {0, 0, INT32_MAX};
    CTO_CHECK(cudaMemcpyAsync(over, init, sizeof(init), cudaMemcpyHostToDevice, s));
    const int stage = tokd::stage_bytes_for(text_len, n_rows);
    CTO_CHECK(set_max_dynamic_smem(tokd::tok_rows_kernel<true>, tokd::MAX_STAGE));
    tokd::Out o{nullptr, nullptr, const_cast<int32_t*>(grp_off_dev), const_cast<int32_t*>(ind_off_dev), planes_dev, ind_entry_dev, over + 1};
    tokd::tok_rows_kernel<true><<<(unsigned)ceil_div(n_rows, tokd::TB), tokd::TB, stage, s>>>(text_dev, row_off_dev, n_rows, ref_dev, ref_start,
                                                                                              ref_len, low_bq_cut, max_indel_length, stage, o,
                                                                                              over + 4, over);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    int32_t head[3] = {0, 0, 0};
    CTO_CHECK(cudaMemcpyAsync(head, over, sizeof(head), cudaMemcpyDeviceToHost, s));
    CTO_CHECK(cudaStreamSynchronize(s));
    int rc = tokd::report(head + 1, "tokenize_write");
    if (!rc && head[0] > 0) {
        const int n_over = head[0];
        const int batch = n_over < 256 ? n_over : 256;
        tokd::Allele* scratch = nullptr;
        CTO_CHECK(scratch_alloc((void**)&scratch, sizeof(tokd::Allele) * (size_t)tokd::BIG_TABLE * batch, s));
        for (int b0 = 0; b0 < n_over; b0 += batch) {
            const int nb = n_over - b0 < batch ? n_over - b0 : batch;
            tokd::tok_big_rows_kernel<<<ceil_div(nb, tokd::TB), tokd::TB, 0, s>>>(text_dev, row_off_dev, over + 4 + b0, nb, ref_dev, ref_start,
                                                                                  ref_len, low_bq_cut, max_indel_length, scratch, o);
            CTO_CHECK(cudaGetLastError());
            count_launch();
        }
        CTO_CHECK(cudaMemcpyAsync(head, over, sizeof(head), cudaMemcpyDeviceToHost, s));
        CTO_CHECK(cudaFreeAsync(scratch, s));
        CTO_CHECK(cudaStreamSynchronize(s));
        rc = tokd::report(head + 1, "tokenize_write");
    }
    cudaFreeAsync(over, s);
    return rc;
}

int cto_window_table(const int32_t* row_pos_dev, int64_t n_rows, const int64_t* cand_pos_dev, int64_t n_candidates, int32_t* win_pos_dev,
                     void* stream) {
    CTO_REQUIRE(n_candidates >= 0 && n_rows >= 0, "window_table: negative size");
    if (n_candidates == 0) return 0;
    CTO_REQUIRE(cand_pos_dev && win_pos_dev && (n_rows == 0 || row_pos_dev), "window_table: NULL array");
    tokd::window_table_kernel<<<(unsigned)ceil_div(n_candidates * N_POS, 256), 256, 0, (cudaStream_t)stream>>>(row_pos_dev, n_rows, cand_pos_dev,
                                                                                                               n_candidates, win_pos_dev);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

}  // extern "C"
