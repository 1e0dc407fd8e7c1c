// Candidate scan: `samtools mpileup` text -> per-position candidate decision, on the GPU.
//
// Replaces the per-row work of STEP 1 of the reference, src/extract_candidates_calling.py (cited as EC): the tokenizer
// of decode_pileup_bases (EC:73-93), its counters (EC:95-120), the allele-frequency / coverage tests (EC:122-144) and the
// candidate-set rules of extract_pair_candidates (EC:335-377).  The reference runs this Python loop over EVERY genomic
// position of a chunk; it is byte work with a few integer counters per row, so here it is one THREAD per mpileup row:
//
//   * the text of a chunk lies in device memory as it came out of samtools; row r is text[row_off[r], row_off[r + 1]);
//   * every thread walks its row in place (byte loads through L1; a shared-memory stage was measured slower, see scan_kernel):
//     column 2 = position, column 5 = the bases string; read symbols, `+N<seq>` / `-N<seq>` suffixes (attached to the previous
//     read, EC:76-87), `^x` (two characters) and everything else skipped -- exactly the reference's state machine, including
//     its corner cases (a `+0` suffix, a second suffix replacing the first, an N read carrying an indel); one character per
//     loop iteration, character classes by arithmetic, so that the 32 rows of a warp advance together;
//   * indel-carrying reads are queued per thread and counted behind the loop (all lanes at once instead of one at a time);
//   * with --select_indel_candidates the reference counts every distinct indel allele ('I' + base + sequence in upper case,
//     'D' + length, EC:111-120): a per-thread table of 24 alleles, compared on the text itself; a row with more distinct
//     alleles is flagged and re-run by a second launch whose tables live in global memory (8192 alleles per row);
//   * the frequency tests are the reference's double-precision comparisons (count / depth >= min_af).
// Output per row: position, depth, flag bits (valid reference base, pass_af, SNV candidate, indel candidate).
#include "../../include/clairs_to_b200.h"
#include "common.cuh"
#include <string.h>
#include <vector>

namespace cto {

namespace cand {

constexpr int TB = 128;                  // rows per block
constexpr int TABLE = 24;                // distinct indel alleles per row in the first pass
constexpr int BIG_TABLE = 8192;          // ... in the second pass (global memory)

enum { F_VALID = 1, F_PASS_AF = 2, F_SNV = 4, F_INDEL = 8, F_MALFORMED = 32, F_BAD_REF = 64, F_OVERFLOW = 128 };

struct Params {
    double min_coverage, snv_min_af, indel_min_af;
    int alt_num;                         // alternative_base_num; < 0 = None (no candidate can pass, EC:134-137)
    int select_indel;
};

struct Allele {
    uint32_t off;                        // text offset of the first occurrence's sequence
    uint32_t len;                        // sequence length
    uint32_t hash;
    int32_t count;
    uint8_t del, sym;                    // deletion?  upper-case read symbol (insertions only)
};

__device__ __forceinline__ uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
__device__ __forceinline__ int base_index(uint8_t up) { return up == 'A' ? 0 : up == 'C' ? 1 : up == 'G' ? 2 : up == 'T' ? 3 : -1; }

// EC:110-120: one counter per distinct indel allele of the row ('I' + base + sequence in upper case, 'D' + length).  Rare
// (1-2 % of the reads), so it is kept out of the hot loop.
template <typename TextPtr>
__device__ __noinline__ void note_allele(TextPtr t, Allele* table, int cap, int& n_alleles, bool& overflow, uint8_t cur, uint8_t sign,
                                         int64_t seq_off, uint32_t seq_len) {
    const uint8_t up = upper(cur);
    const bool del = sign == '-';
    uint32_t h = del ? 0x9E3779B9u * (seq_len + 1) : 2166136261u ^ up;
    if (!del)
        for (uint32_t k = 0; k < seq_len; ++k) h = (h ^ upper(t[seq_off + k])) * 16777619u;
    int e = 0;
    for (; e < n_alleles; ++e) {
        const Allele& a = table[e];
        if (a.hash != h || a.del != (uint8_t)del || a.len != seq_len) continue;
        if (del) break;                                         // 'D' + 'N' * length: the length is the key
        if (a.sym != up) continue;
        uint32_t k = 0;
        while (k < seq_len && upper(t[a.off + k]) == upper(t[seq_off + k])) ++k;
        if (k == seq_len) break;
    }
    if (e < n_alleles) ++table[e].count;
    else if (n_alleles < cap) {
        Allele& a = table[n_alleles++];
        a.off = (uint32_t)seq_off; a.len = seq_len; a.hash = h; a.count = 1; a.del = (uint8_t)del; a.sym = up;
    } else overflow = true;
}

// One row.  `t` indexes the chunk text (shared-memory copy or global memory, same offsets).
template <typename TextPtr>
__device__ void scan_row(TextPtr t, int64_t lo, int64_t hi, const uint8_t* __restrict__ ref, int64_t ref_start, int64_t ref_len,
                         const Params& p, Allele* table, int cap, int32_t* pos_out, int32_t* depth_out, uint8_t* flag_out) {
    int64_t i = lo;
    while (i < hi && t[i] != '\t') ++i;                         // column 1: contig
    ++i;
    int64_t pos = 0;
    while (i < hi && t[i] >= '0' && t[i] <= '9') { pos = pos * 10 + (t[i] - '0'); ++i; }   // column 2
    for (int col = 0; col < 3 && i < hi; ++col) {               // to the start of column 5
        while (i < hi && t[i] != '\t') ++i;
        ++i;
    }
    *pos_out = (int32_t)pos;
    *depth_out = 0;
    if (i > hi) { *flag_out = F_MALFORMED; return; }            // fewer than five columns: the reference raises IndexError
    const int64_t rp = pos - ref_start;
    if (rp < 0 || rp >= ref_len) { *flag_out = F_BAD_REF; return; }
    const uint8_t ref_up = upper(ref[rp]);
    const int ref_b = base_index(ref_up);
    if (ref_b < 0) { *flag_out = 0; return; }                   // EC:341-342: reference base not in ACGT, row skipped

    // Hot loop: one byte per iteration, the common characters (read symbols) handled by straight-line code so that the lanes
    // of a warp stay together; the rare indel suffix is parsed in a branch and its allele bookkeeping is OUT OF LINE
    // (note_allele).  A first version kept the four base counters in an indexed array (local memory) and inlined the allele
    // code at both flush sites: ncu showed the symbol path executing with 6 of 32 lanes (profiles/r2_ncu_metrics_scan.txt).
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    int depth = 0, n_alleles = 0;
    bool plain_alt = false, overflow = false;
    // the read being assembled: its symbol, whether it is a plain alt base, and the LAST indel suffix attached to it (EC:86 overwrites)
    uint8_t cur = 0, sign = 0;
    bool cur_alt = false;
    int64_t seq_off = 0;
    uint32_t seq_len = 0;
    // One character per iteration for every lane, whatever it is: `skip` swallows the characters of an indel sequence and the
    // one behind `^`, `digits` collects the length behind a sign.  No inner loops and no `continue`, so the rows of a warp
    // advance in lockstep and only the per-character work is predicated.
    // Indel-carrying reads are rare per row (1-2 %) but a warp of 32 rows meets one every few characters, and each used to be
    // handled where it occurred, by ONE lane (31 % of the issued instructions ran with 0-3 lanes).  They are queued instead and
    // counted behind the loop, where every lane works on its own queue at the same time.
    constexpr int QUEUE = 8;
    uint32_t q_off[QUEUE], q_len[QUEUE];
    uint16_t q_meta[QUEUE];
    int n_q = 0;
    auto queue_allele = [&]() {
        if (n_q < QUEUE) {
            q_off[n_q] = (uint32_t)(seq_off - lo); q_len[n_q] = seq_len; q_meta[n_q] = (uint16_t)(cur | (sign << 8));
            ++n_q;
        } else {
            note_allele(t, table, cap, n_alleles, overflow, cur, sign, seq_off, seq_len);      // a deep row: count it right away
        }
    };
    uint32_t skip = 0, adv = 0;
    bool digits = false, done = false;
    uint8_t pend = 0;                                           // the sign whose length is being read
    while (i < hi && !done) {
        const uint8_t c = t[i];
        bool normal = true;
        if (skip) {
            --skip;
            normal = false;
        } else if (digits) {                                     // EC:76-87
            if (c >= '0' && c <= '9') {
                adv = adv * 10 + (c - '0');
                normal = false;
            } else {
                digits = false;
                if (cur) { sign = pend; seq_off = i; seq_len = adv < (uint32_t)(hi - i) ? adv : (uint32_t)(hi - i); }
                if (adv) { skip = adv - 1; normal = false; }     // this character is the first of the sequence; adv == 0 re-reads it
            }
        }
        if (normal) {
            // character class by arithmetic only: a `lc == 'a' ? 0 : lc == 'c' ? 1 : ...` chain is compiled into a branch tree with
            // one path per base, which splits the lanes of a warp by base identity (6 of 32 lanes per path, ncu source page)
            const uint32_t lc = c | 0x20u;                       // 'A' and 'a' -> 'a' (no other byte maps onto a letter)
            const uint32_t k5 = lc - 'a';
            const uint32_t is_base = k5 < 26u ? (0x80045u >> k5) & 1u : 0u;      // a, c, g, t = bits 0, 2, 6, 19
            const uint32_t code = (lc >> 1) & 3u;                // a 0, c 1, t 2, g 3
            const int b = is_base ? (int)(code ^ (code >> 1)) : -1;             // -> A 0, C 1, G 2, T 3
            const bool gap = c == '#' || c == '*';
            if (is_base || lc == 'n' || gap) {                   // EC:89-90: a new read; the previous one is complete
                if (sign) {
                    if (p.select_indel) queue_allele();   // EC:110-120
                } else {
                    plain_alt |= cur_alt;                        // an alt_list key that is a single base (EC:369, 146-148)
                }
                c0 += is_base & (uint32_t)(b == 0); c1 += is_base & (uint32_t)(b == 1);
                c2 += is_base & (uint32_t)(b == 2); c3 += is_base & (uint32_t)(b == 3);
                depth += is_base | (uint32_t)gap;                // EC:105-109 (reads carrying an indel count too)
                cur = c; cur_alt = is_base && b != ref_b; sign = 0;
            } else if (c == '+' || c == '-') {
                digits = true; adv = 0; pend = c;
            } else if (c == '^') {                               // EC:91-92
                skip = 1;
            } else if (c == '\t' || c == '\n' || c == '\r') {    // end of column 5 (row.strip().split('\t'))
                done = true;
            }
        }
        ++i;
    }
    if (digits && cur) { sign = pend; seq_off = hi; seq_len = 0; }   // a length that runs into the end of the row
    if (cur) {
        if (sign) {
            if (p.select_indel) queue_allele();
        } else {
            plain_alt |= cur_alt;
        }
    }
    for (int e = 0; e < n_q; ++e)
        note_allele(t, table, cap, n_alleles, overflow, (uint8_t)(q_meta[e] & 0xff), (uint8_t)(q_meta[e] >> 8), lo + q_off[e], q_len[e]);
    const int cnt[4] = {c0, c1, c2, c3};

    const double denom = depth > 0 ? (double)depth : 1.0;      // EC:121
    const bool pass_depth = (double)depth > p.min_coverage;     // EC:127
    bool pass_snv = false, pass_indel = false;
    if (p.alt_num >= 0) {
        for (int b = 0; b < 4; ++b)                             // EC:128-137
            if (b != ref_b && cnt[b] > 0 && (double)cnt[b] / denom >= p.snv_min_af && cnt[b] >= p.alt_num) pass_snv = true;
        if (p.select_indel)
            for (int e = 0; e < n_alleles; ++e)
                if ((double)table[e].count / denom >= p.indel_min_af && table[e].count >= p.alt_num) pass_indel = true;
    }
    const bool pass_af = (pass_snv || pass_indel) && pass_depth;   // EC:144
    uint8_t f = F_VALID;
    if (pass_af) f |= F_PASS_AF;
    if (pass_af && pass_snv && plain_alt) f |= F_SNV;             // EC:366-371
    if (p.select_indel && pass_af && pass_indel) f |= F_INDEL;    // EC:372-377 (an indel read exists whenever pass_indel holds)
    if (overflow) f |= F_OVERFLOW;
    *depth_out = depth;
    *flag_out = f;
}

struct GlobalText {
    const uint8_t* p;
    __device__ __forceinline__ uint8_t operator[](int64_t i) const { return __ldg(p + i); }
};

// The text is read in place (byte loads through L1): staging a block's span in shared memory first was measured 10 % SLOWER
// (1.16 against 1.06 ms per 572 MB) -- two thirds of a row are the quality columns, which this scan never looks at, and the
// stage limited the SM to 6 blocks.  (The tokenizer kernels, which read every column, keep their stage.)
__global__ void __launch_bounds__(TB)
scan_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ row_off, int64_t n_rows, const uint8_t* __restrict__ ref,
            int64_t ref_start, int64_t ref_len, Params p, int32_t* __restrict__ pos_out, int32_t* __restrict__ depth_out,
            uint8_t* __restrict__ flags_out, int32_t* __restrict__ overflow_rows, int32_t* __restrict__ overflow_count) {
    const int64_t r = (int64_t)blockIdx.x * TB + threadIdx.x;
    if (r >= n_rows) return;
    Allele table[TABLE];
    uint8_t flag;
    scan_row(GlobalText{text}, row_off[r], row_off[r + 1], ref, ref_start, ref_len, p, table, TABLE, pos_out + r, depth_out + r, &flag);
    flags_out[r] = flag;
    if (flag & F_OVERFLOW) overflow_rows[atomicAdd(overflow_count, 1)] = (int32_t)r;
}

// second pass: the rows whose allele table overflowed, tables in global memory
__global__ void __launch_bounds__(TB)
scan_big_kernel(const uint8_t* __restrict__ text, const int64_t* __restrict__ row_off, const int32_t* __restrict__ rows, int n_list,
                const uint8_t* __restrict__ ref, int64_t ref_start, int64_t ref_len, Params p, Allele* __restrict__ scratch,
                int32_t* __restrict__ pos_out, int32_t* __restrict__ depth_out, uint8_t* __restrict__ flags_out) {
    const int k = blockIdx.x * TB + threadIdx.x;
    if (k >= n_list) return;
    const int64_t r = rows[k];
    scan_row(GlobalText{text}, row_off[r], row_off[r + 1], ref, ref_start, ref_len, p, scratch + (int64_t)k * BIG_TABLE, BIG_TABLE,
             pos_out + r, depth_out + r, flags_out + r);
}

// ---- row index: byte offsets of the rows of a text ('\n' terminated; a last row without '\n' counts) ----------------------
constexpr int IDX_THREADS = 256;
constexpr int IDX_BYTES = 32;                                   // bytes per thread: two 16-byte loads
constexpr int IDX_TILE = IDX_THREADS * IDX_BYTES;

// 16 bytes of the text at offset `off` (a multiple of 16), bytes at or beyond `len` read as zero
__device__ __forceinline__ uint4 load16(const uint8_t* __restrict__ text, int64_t off, int64_t len, bool aligned) {
    if (aligned && off + 16 <= len) return __ldg(reinterpret_cast<const uint4*>(text + off));
    uint32_t w[4] = {0, 0, 0, 0};
    for (int k = 0; k < 16; ++k)
        if (off + k < len) w[k >> 2] |= (uint32_t)__ldg(text + off + k) << (8 * (k & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint32_t newline_bits(uint32_t w) {  // bit 8k+7 set where byte k == '\n'
    return __vcmpeq4(w, 0x0A0A0A0Au) & 0x80808080u;
}
__device__ __forceinline__ int count16(uint4 v) {
    return __popc(newline_bits(v.x)) + __popc(newline_bits(v.y)) + __popc(newline_bits(v.z)) + __popc(newline_bits(v.w));
}

__global__ void __launch_bounds__(IDX_THREADS)
count_newlines_kernel(const uint8_t* __restrict__ text, int64_t len, int32_t* __restrict__ tile_count) {
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    const int64_t off = (int64_t)blockIdx.x * IDX_TILE + threadIdx.x * IDX_BYTES;
    int n = 0;
    if (off < len) n = count16(load16(text, off, len, aligned)) + count16(load16(text, off + 16, len, aligned));
    __shared__ int warp_sum[IDX_THREADS / 32];
    for (int d = 16; d; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < IDX_THREADS / 32; ++w) t += warp_sum[w];
        tile_count[blockIdx.x] = t;
    }
}

// exclusive prefix sum of the tile counts, in place (one block); total[0] = rows, counting an unterminated last row
__global__ void __launch_bounds__(1024)
scan_tiles_kernel(int32_t* __restrict__ tile_count, int n_tiles, const uint8_t* __restrict__ text, int64_t len, int64_t* __restrict__ total) {
    __shared__ int warp_sum[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int k = base + threadIdx.x;
        const int v = k < n_tiles ? tile_count[k] : 0;
        int inc = v;
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += o;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = warp_sum[threadIdx.x];
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, w, d);
                if (threadIdx.x >= d) w += o;
            }
            warp_sum[threadIdx.x] = w;                          // inclusive over warps
        }
        __syncthreads();
        const int carry = carry_s;
        const int warp_base = (threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0;
        if (k < n_tiles) tile_count[k] = carry + warp_base + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = carry_s + ((len > 0 && __ldg(text + len - 1) != '\n') ? 1 : 0);
}

// row_off[0] = 0; row_off[k + 1] = offset behind the k-th '\n'; an unterminated last row ends at len
__global__ void __launch_bounds__(IDX_THREADS)
write_row_offsets_kernel(const uint8_t* __restrict__ text, int64_t len, const int32_t* __restrict__ tile_base,
                         int64_t* __restrict__ row_off, int64_t cap_rows) {
    const bool aligned = (reinterpret_cast<uintptr_t>(text) & 15) == 0;
    const int64_t off = (int64_t)blockIdx.x * IDX_TILE + threadIdx.x * IDX_BYTES;
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (off < len) {
        const uint4 a = load16(text, off, len, aligned), b = load16(text, off + 16, len, aligned);
        w[0] = newline_bits(a.x); w[1] = newline_bits(a.y); w[2] = newline_bits(a.z); w[3] = newline_bits(a.w);
        w[4] = newline_bits(b.x); w[5] = newline_bits(b.y); w[6] = newline_bits(b.z); w[7] = newline_bits(b.w);
    }
    int n = 0;
    for (int k = 0; k < 8; ++k) n += __popc(w[k]);
    int inc = n;
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if ((threadIdx.x & 31) >= d) inc += o;
    }
    __shared__ int warp_sum[IDX_THREADS / 32];
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = inc;
    __syncthreads();
    int64_t at = tile_base[blockIdx.x] + inc - n;
    for (int q = 0; q < (int)(threadIdx.x >> 5); ++q) at += warp_sum[q];
    if (blockIdx.x == 0 && threadIdx.x == 0) row_off[0] = 0;
    for (int k = 0; k < 8; ++k) {
        uint32_t m = w[k];
        while (m) {
            const int bit = __ffs(m) - 1;                       // 8 * byte + 7
            m &= m - 1;
            if (at + 1 <= cap_rows) row_off[at + 1] = off + 4 * k + (bit >> 3) + 1;
            ++at;
        }
    }
    if (off <= len - 1 && len - 1 < off + IDX_BYTES && __ldg(text + len - 1) != '\n' && at + 1 <= cap_rows) row_off[at + 1] = len;
}

}  // namespace cand

// Row offsets of a '\n'-separated text in device memory.  Phase 1 counts (synchronises the stream to return the row count),
// phase 2 writes row_off[0 .. n_rows] (int64).  tile_dev: int32 [ceil(len / 8192) + 1] scratch shared by the phases.
int launch_count_rows(const uint8_t* text_dev, int64_t len, int32_t* tile_dev, int64_t* total_dev, int64_t* n_rows, cudaStream_t s) {
    *n_rows = 0;
    if (len <= 0) return 0;
    const int n_tiles = ceil_div(len, cand::IDX_TILE);
    cand::count_newlines_kernel<<<n_tiles, cand::IDX_THREADS, 0, s>>>(text_dev, len, tile_dev);
    CTO_CHECK(cudaGetLastError());
    cand::scan_tiles_kernel<<<1, 1024, 0, s>>>(tile_dev, n_tiles, text_dev, len, total_dev);
    CTO_CHECK(cudaGetLastError());
    count_launch(2);
    CTO_CHECK(cudaMemcpyAsync(n_rows, total_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CTO_CHECK(cudaStreamSynchronize(s));
    return 0;
}
int launch_write_row_offsets(const uint8_t* text_dev, int64_t len, const int32_t* tile_dev, int64_t* row_off_dev, int64_t cap_rows,
                             cudaStream_t s) {
    if (len <= 0) return 0;
    cand::write_row_offsets_kernel<<<ceil_div(len, cand::IDX_TILE), cand::IDX_THREADS, 0, s>>>(text_dev, len, tile_dev, row_off_dev,
                                                                                               cap_rows);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}

// text_dev: the chunk's mpileup text, ALLOCATED up to the next multiple of 16 bytes; row_off_dev: int64 [n_rows + 1] byte offsets
// of the rows (row_off[n_rows] = text length); ref_dev: reference bases (ASCII) of positions ref_start .. ref_start + ref_len - 1.
// overflow_dev: int32 [n_rows + 1] scratch (entry 0 = counter).  Writes the number of rows that needed the second pass to
// *n_overflow (host) and synchronises the stream.
int launch_scan_candidates(const uint8_t* text_dev, const int64_t* row_off_dev, int64_t n_rows, int64_t text_len, const uint8_t* ref_dev,
                           int64_t ref_start, int64_t ref_len, double min_coverage, double snv_min_af, double indel_min_af,
                           int alt_num, int select_indel, int32_t* pos_dev, int32_t* depth_dev, uint8_t* flags_dev,
                           int32_t* overflow_dev, int* n_overflow, cudaStream_t s) {
    if (n_overflow) *n_overflow = 0;
    if (n_rows <= 0) return 0;
    CTO_REQUIRE(n_rows < (1ll << 31), "scan_candidates: %lld rows in one call", (long long)n_rows);
    cand::Params p{min_coverage, snv_min_af, indel_min_af, alt_num, select_indel};
    CTO_CHECK(cudaMemsetAsync(overflow_dev, 0, sizeof(int32_t), s));
    const unsigned blocks = (unsigned)ceil_div(n_rows, cand::TB);
    cand::scan_kernel<<<blocks, cand::TB, 0, s>>>(text_dev, row_off_dev, n_rows, ref_dev, ref_start, ref_len, p, pos_dev, depth_dev, flags_dev,
                                                  overflow_dev + 1, overflow_dev);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    int32_t n_over = 0;
    CTO_CHECK(cudaMemcpyAsync(&n_over, overflow_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CTO_CHECK(cudaStreamSynchronize(s));
    if (n_over > 0) {                                           // in batches: 160 KB of allele table per row
        const int batch = n_over < 256 ? n_over : 256;
        cand::Allele* scratch = nullptr;
        CTO_CHECK(scratch_alloc((void**)&scratch, sizeof(cand::Allele) * (size_t)cand::BIG_TABLE * batch, s));
        for (int b0 = 0; b0 < n_over; b0 += batch) {
            const int nb = n_over - b0 < batch ? n_over - b0 : batch;
            cand::scan_big_kernel<<<ceil_div(nb, cand::TB), cand::TB, 0, s>>>(text_dev, row_off_dev, overflow_dev + 1 + b0, nb, ref_dev,
                                                                              ref_start, ref_len, p, scratch, pos_dev, depth_dev, flags_dev);
            CTO_CHECK(cudaGetLastError());
            count_launch();
        }
        CTO_CHECK(cudaFreeAsync(scratch, s));
        CTO_CHECK(cudaStreamSynchronize(s));
    }
    if (n_overflow) *n_overflow = n_over;
    return 0;
}

}  // namespace cto

using namespace cto;

extern "C" {

int cto_index_rows(const uint8_t* text_dev, int64_t text_len, int64_t* row_off_dev, int64_t cap_rows, int64_t* n_rows, void* stream) {
    CTO_REQUIRE(n_rows, "index_rows: NULL n_rows");
    CTO_REQUIRE(text_len >= 0 && (text_len == 0 || text_dev), "index_rows: bad text");
    cudaStream_t s = (cudaStream_t)stream;
    *n_rows = 0;
    if (text_len == 0) {
        if (row_off_dev) CTO_CHECK(cudaMemsetAsync(row_off_dev, 0, sizeof(int64_t), s));
        return 0;
    }
    const int n_tiles = ceil_div(text_len, cand::IDX_TILE);
    int32_t* tiles = nullptr;
    CTO_CHECK(scratch_alloc((void**)&tiles, sizeof(int32_t) * (size_t)n_tiles + 16, s));
    int64_t* total = reinterpret_cast<int64_t*>(tiles + ((n_tiles + 1) & ~1));
    int rc = launch_count_rows(text_dev, text_len, tiles, total, n_rows, s);
    if (!rc && row_off_dev) {
        if (*n_rows > cap_rows) {
            set_error("index_rows: %lld rows, room for %lld", (long long)*n_rows, (long long)cap_rows);
            rc = 2;
        } else {
            rc = launch_write_row_offsets(text_dev, text_len, tiles, row_off_dev, cap_rows, s);
        }
    }
    cudaFreeAsync(tiles, s);
    return rc;
}

int cto_scan_candidates(const uint8_t* text_dev, int64_t text_len, const int64_t* row_off_dev, int64_t n_rows, const uint8_t* ref_dev,
                        int64_t ref_start, int64_t ref_len, double min_coverage, double snv_min_af, double indel_min_af,
                        int alternative_base_num, int select_indel_candidates, int32_t* pos_dev, int32_t* depth_dev,
                        uint8_t* flags_dev, int32_t* n_overflow, void* stream) {
    CTO_REQUIRE(n_rows >= 0, "scan_candidates: negative row count");
    if (n_overflow) *n_overflow = 0;
    if (n_rows == 0) return 0;
    CTO_REQUIRE(text_dev && row_off_dev && ref_dev && pos_dev && depth_dev && flags_dev, "scan_candidates: NULL array");
    CTO_REQUIRE(text_len < (1ll << 32), "scan_candidates: %lld bytes of text in one call (limit 4 GiB)", (long long)text_len);
    cudaStream_t s = (cudaStream_t)stream;
    int32_t* over = nullptr;
    CTO_CHECK(scratch_alloc((void**)&over, sizeof(int32_t) * (size_t)(n_rows + 1), s));
    int n_over = 0;
    const int rc = launch_scan_candidates(text_dev, row_off_dev, n_rows, text_len, ref_dev, ref_start, ref_len, min_coverage, snv_min_af,
                                          indel_min_af, alternative_base_num, select_indel_candidates, pos_dev, depth_dev, flags_dev, over,
                                          &n_over, s);
    cudaFreeAsync(over, s);
    if (n_overflow) *n_overflow = n_over;
    return rc;
}

int cto_scan_candidates_host(const char* text, int64_t text_len, const char* ref, int64_t ref_start, int64_t ref_len,
                             double min_coverage, double snv_min_af, double indel_min_af, int alternative_base_num,
                             int select_indel_candidates, int64_t cap_rows, int32_t* pos, int32_t* depth, uint8_t* flags,
                             int64_t* n_rows, int64_t* n_overflow, void* stream) {
    CTO_REQUIRE(n_rows, "scan_candidates_host: NULL n_rows");
    *n_rows = 0;
    if (n_overflow) *n_overflow = 0;
    CTO_REQUIRE(text_len >= 0 && ref_len > 0 && ref, "scan_candidates_host: bad argument");
    if (text_len == 0) return 0;
    CTO_REQUIRE(text && pos && depth && flags, "scan_candidates_host: NULL array");
    if (cto_device_check(nullptr)) return 3;
    cudaStream_t s = (cudaStream_t)stream;
    constexpr int64_t PIECE = 32ll << 20;                      // text bytes per pipeline stage
    constexpr int64_t PIECE_ROWS = PIECE / 8;                  // a valid row has at least 10 bytes
    // piece boundaries on row ends
    std::vector<int64_t> cut{0};
    while (cut.back() < text_len) {
        int64_t e = cut.back() + PIECE;
        if (e >= text_len) e = text_len;
        else {
            const void* nl = memrchr(text + cut.back(), '\n', (size_t)(e - cut.back()));
            CTO_REQUIRE(nl, "scan_candidates_host: a row longer than %lld bytes", (long long)PIECE);
            e = (const char*)nl - text + 1;
        }
        cut.push_back(e);
    }
    const int n_pieces = (int)cut.size() - 1;
    const int n_slots = n_pieces > 1 ? 2 : 1;
    auto up = [](int64_t b) { return (b + 255) & ~int64_t(255); };
    const int64_t first = cut[1] - cut[0];
    const int64_t text_b = up((n_pieces > 1 ? PIECE : first) + 16);
    const int64_t rows_cap = n_pieces > 1 ? PIECE_ROWS : first / 8 + 2;
    const int64_t tiles_b = up(sizeof(int32_t) * (size_t)(text_b / cand::IDX_TILE + 2) + 16);
    const int64_t off_b = up(sizeof(int64_t) * (size_t)(rows_cap + 1));
    const int64_t i32_b = up(sizeof(int32_t) * (size_t)(rows_cap + 1));
    const int64_t u8_b = up(rows_cap);
    const int64_t slot_b = text_b + tiles_b + off_b + 3 * i32_b + u8_b;
    uint8_t* arena = nullptr;
    CTO_CHECK(scratch_alloc((void**)&arena, (size_t)(slot_b * n_slots + up(ref_len)), s));
    uint8_t* ref_dev = arena + slot_b * n_slots;
    cudaStream_t cs = nullptr;
    cudaEvent_t copied[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    int rc = 0;
    auto fail = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && !rc) {
            set_error("scan_candidates_host: %s: %s", what, cudaGetErrorString(e));
            rc = 1;
        }
        return e != cudaSuccess;
    };
    fail(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking), "stream");
    for (int k = 0; k < n_slots && !rc; ++k) {
        fail(cudaEventCreateWithFlags(&copied[k], cudaEventDisableTiming), "event");
        fail(cudaEventCreateWithFlags(&freed[k], cudaEventDisableTiming), "event");
    }
    if (!rc) fail(cudaMemcpyAsync(ref_dev, ref, (size_t)ref_len, cudaMemcpyHostToDevice, s), "reference copy");
    cudaEvent_t arena_ready = nullptr;
    if (!rc) fail(cudaEventCreateWithFlags(&arena_ready, cudaEventDisableTiming), "event");
    if (!rc) fail(cudaEventRecord(arena_ready, s), "event");
    if (!rc) fail(cudaStreamWaitEvent(cs, arena_ready, 0), "wait");
    auto issue_copy = [&](int k) {                             // piece k -> its slot, on the copy stream
        const int slot = k & 1;
        uint8_t* dst = arena + slot_b * slot;
        if (k >= 2 && fail(cudaStreamWaitEvent(cs, freed[slot], 0), "wait")) return;
        if (fail(cudaMemcpyAsync(dst, text + cut[k], (size_t)(cut[k + 1] - cut[k]), cudaMemcpyHostToDevice, cs), "text copy")) return;
        fail(cudaEventRecord(copied[slot], cs), "event");
    };
    int64_t row_base = 0, over_total = 0;
    if (!rc) issue_copy(0);
    for (int k = 0; k < n_pieces && !rc; ++k) {
        if (k + 1 < n_pieces) issue_copy(k + 1);
        if (rc) break;
        const int slot = k & 1;
        uint8_t* base = arena + slot_b * slot;
        uint8_t* text_dev = base;
        int32_t* tiles = reinterpret_cast<int32_t*>(base + text_b);
        int64_t* total = reinterpret_cast<int64_t*>(base + text_b + tiles_b - 16);
        int64_t* row_off = reinterpret_cast<int64_t*>(base + text_b + tiles_b);
        int32_t* pos_d = reinterpret_cast<int32_t*>(base + text_b + tiles_b + off_b);
        int32_t* depth_d = reinterpret_cast<int32_t*>(base + text_b + tiles_b + off_b + i32_b);
        int32_t* over_d = reinterpret_cast<int32_t*>(base + text_b + tiles_b + off_b + 2 * i32_b);
        uint8_t* flags_d = base + text_b + tiles_b + off_b + 3 * i32_b;
        const int64_t len = cut[k + 1] - cut[k];
        if (fail(cudaStreamWaitEvent(s, copied[slot], 0), "wait")) break;
        int64_t n = 0;
        rc = launch_count_rows(text_dev, len, tiles, total, &n, s);
        if (rc) break;
        if (n > rows_cap || row_base + n > cap_rows) {
            set_error("scan_candidates_host: %lld rows in piece %d (room for %lld; caller gave room for %lld in total)", (long long)n, k,
                      (long long)rows_cap, (long long)cap_rows);
            rc = 2;
            break;
        }
        rc = launch_write_row_offsets(text_dev, len, tiles, row_off, rows_cap, s);
        int n_over = 0;
        if (!rc)
            rc = launch_scan_candidates(text_dev, row_off, n, len, ref_dev, ref_start, ref_len, min_coverage, snv_min_af, indel_min_af,
                                        alternative_base_num, select_indel_candidates, pos_d, depth_d, flags_d, over_d, &n_over, s);
        if (rc) break;
        over_total += n_over;
        fail(cudaMemcpyAsync(pos + row_base, pos_d, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
        fail(cudaMemcpyAsync(depth + row_base, depth_d, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
        fail(cudaMemcpyAsync(flags + row_base, flags_d, (size_t)n, cudaMemcpyDeviceToHost, s), "D2H");
        fail(cudaEventRecord(freed[slot], s), "event");
        row_base += n;
    }
    if (cs) cudaStreamSynchronize(cs);
    cudaFreeAsync(arena, s);
    const cudaError_t se = cudaStreamSynchronize(s);
    for (int k = 0; k < 2; ++k) {
        if (copied[k]) cudaEventDestroy(copied[k]);
        if (freed[k]) cudaEventDestroy(freed[k]);
    }
    if (arena_ready) cudaEventDestroy(arena_ready);
    if (cs) cudaStreamDestroy(cs);
    if (!rc && se != cudaSuccess) {
        set_error("scan_candidates_host: %s", cudaGetErrorString(se));
        rc = 1;
    }
    if (!rc) {
        *n_rows = row_base;
        if (n_overflow) *n_overflow = over_total;
    }
    return rc;
}

}  // extern "C"
