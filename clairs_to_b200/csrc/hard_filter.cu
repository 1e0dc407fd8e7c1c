// Per-site hard filters on the GPU.
//
// Replaces _haplotype_build_state_and_line + _haplotype_finalize_line (src/haplotype_filtering.py:570-703, 344-565, cited
// as HF) and the post-filter twins (src/postfilter_variants.py:368-446, 278-365, PV).  The reference keeps, per site, Python
// dicts keyed by read name for the 2 * flanking + 1 pileup rows around the site and runs set algebra on them.  Here the host
// tokenizer (hard_filter_host.cpp) has turned every string into a dense id, and ONE THREAD BLOCK evaluates one site:
//
//   * the per-read state of the site (haplotype 0/1/2, "carries the alt allele", "starts or ends nearby") is ONE BYTE per read,
//     indexed by read id - the window's smallest read id: 16 KB of shared memory cover 16 384 reads (larger windows use a global
//     scratch table); membership tests of the reference's sets are byte loads, insertions are atomicOr on the containing word;
//   * rows are walked by warps (one row per warp at a time, lanes over the row's entries, 14 bytes per entry, coalesced);
//   * the variant-cluster test needs the most frequent non-reference token among the alt reads of a row, but only if that
//     token is carried by MORE THAN HALF of the alt reads (HF:421-424 with eps = 0.5): a Boyer-Moore majority vote per lane,
//     merged across the warp, then one counting pass -- no per-row hash map;
//   * every ratio test of the reference (x / y <= c with small integers) is evaluated as the equivalent integer inequality;
//   * Fisher's exact test (HF:60-98): the reference divides exact big integers (one correctly rounded quotient) and then walks
//     the neighbouring tables by multiply / divide in double precision, comparing each term with `<=` -- ties between mirror
//     tables depend on the last bit.  The quotient is formed here in double-double arithmetic (106 bits, binary exponent
//     carried separately) and rounded once, the walk uses the same IEEE operations in the same order;
//   * the sequence entropy (HF:101-151) adds and subtracts table entries in the reference's order; the table (e * log e) comes
//     from the caller's libm.
#include "../../include/clairs_to_b200.h"
#include "common.cuh"

namespace cto {
namespace hf {

constexpr int TB = 128;
constexpr int WARPS = TB / 32;
constexpr int SMEM_WORDS = CTO_HF_SMEM_READS / 4;
enum : uint32_t { S_HAP = 3u, S_ALT = 4u, S_RSE = 8u };
enum { ALL_N = 0, ALL_F = 3, ALL_R = 6, ALT_N = 9, ALT_F = 12, ALT_R = 15, BQ_SUM = 18, MQ_SUM = 19, N_MATCH = 20, RSE_HITS = 21,
       MATCH_COUNT = 22, INS_LEN = 23, PASS_HET = 24, PASS_HOM = 25, N_ACC = 26 };

struct Params {
    cto_hf_chunk_arrays c;
    cto_hf_site_arrays s;
    int mode, flanking, disable_rse, max_co;
    double tab[35];
    double mul, thr;
    uint32_t* scratch;
    uint32_t* out_flags;
    double* out_p;
    int32_t* out_counts;
};

struct State {
    uint32_t* w;
    int32_t base;
    __device__ __forceinline__ uint32_t get(int32_t rid) const {
        const int32_t k = rid - base;
        return (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
    }
    __device__ __forceinline__ uint32_t fetch_or(int32_t rid, uint32_t bits) {
        const int32_t k = rid - base;
        const int sh = 8 * (k & 3);
        return (atomicOr(&w[k >> 2], bits << sh) >> sh) & 0xffu;
    }
    // HF:623-625: the first haplotype tag seen for a read key stays
    __device__ __forceinline__ void set_hap_once(int32_t rid, uint32_t hap) {
        const int32_t k = rid - base;
        const int sh = 8 * (k & 3);
        uint32_t* p = &w[k >> 2];
        uint32_t cur = *reinterpret_cast<volatile uint32_t*>(p);
        while (((cur >> sh) & S_HAP) == 0) {
            const uint32_t prev = atomicCAS(p, cur, cur | (hap << sh));
            if (prev == cur) break;
            cur = prev;
        }
    }
};

// ---- double-double helpers (value = hi + lo, |lo| <= ulp(hi) / 2) -----------------------------------------------------------
struct dd { double hi, lo; };
__device__ __forceinline__ dd dd_mul(dd a, double b) {
    const double p = __dmul_rn(a.hi, b);
    double e = __fma_rn(a.hi, b, -p);
    e = __fma_rn(a.lo, b, e);
    const double s = __dadd_rn(p, e);
    return dd{s, __dsub_rn(e, __dsub_rn(s, p))};
}
__device__ __forceinline__ dd dd_div(dd a, double b) {
    const double q1 = __ddiv_rn(a.hi, b);
    const double p = __dmul_rn(q1, b);
    const double e = __fma_rn(q1, b, -p);
    const double r = __dadd_rn(__dsub_rn(__dsub_rn(a.hi, p), e), a.lo);
    const double q2 = __ddiv_rn(r, b);
    const double s = __dadd_rn(q1, q2);
    return dd{s, __dsub_rn(q2, __dsub_rn(s, q1))};
}
struct Scaled { dd v; int exp2; };
__device__ __forceinline__ void rescale(Scaled& x) {
    if (fabs(x.v.hi) > 0x1p400) { x.v.hi *= 0x1p-400; x.v.lo *= 0x1p-400; x.exp2 += 400; }
    else if (fabs(x.v.hi) < 0x1p-400 && x.v.hi != 0.0) { x.v.hi *= 0x1p400; x.v.lo *= 0x1p400; x.exp2 -= 400; }
}
// x *= C(n, k) (up = true) or x /= C(n, k): the product of (n - k' + i) / i, i = 1 .. k', k' = min(k, n - k) (HF:44-57)
__device__ void times_binomial(Scaled& x, int64_t n, int64_t k, bool up) {
    if (k > n - k) k = n - k;
    for (int64_t i = 1; i <= k; ++i) {
        const double num = (double)(n - k + i), den = (double)i;
        x.v = up ? dd_div(dd_mul(x.v, num), den) : dd_div(dd_mul(x.v, den), num);
        rescale(x);
    }
}

// HF:60-98
__device__ double fisher_exact(int64_t a, int64_t b, int64_t c, int64_t d) {
    if (a == b && b == c && c == d) return 1.0;
    Scaled x{dd{1.0, 0.0}, 0};
    times_binomial(x, a + b, a, true);
    times_binomial(x, c + d, c, true);
    times_binomial(x, a + b + c + d, a + c, false);
    const double t = scalbn(x.v.hi, x.exp2);                      // hi is the correctly rounded value of hi + lo
    double p = t;
    {
        int64_t w = a, xx = b, y = c, z = d;
        double cur = t, side = 0.0;
        while (w > 0 && z > 0) {
            cur = __dmul_rn(cur, (double)(w * z));
            --w; ++xx; ++y; --z;
            cur = __ddiv_rn(cur, (double)(xx * y));
            if (cur <= t) side = __dadd_rn(side, cur);
        }
        p = __dadd_rn(p, side);
    }
    {
        int64_t w = a, xx = b, y = c, z = d;
        double cur = t, side = 0.0;
        while (xx > 0 && y > 0) {
            cur = __dmul_rn(cur, (double)(xx * y));
            ++w; --xx; --y; ++z;
            cur = __ddiv_rn(cur, (double)(w * z));
            if (cur <= t) side = __dadd_rn(side, cur);
        }
        p = __dadd_rn(p, side);
    }
    return p;
}

// HF:101-151 for a sequence no longer than the window (33): only the "suffix" branch of the loop runs
__device__ double sequence_entropy(const uint8_t* __restrict__ seq, int len, const double* tab, double mul) {
    int kmers[33], cnt[33];
    int nk = 0;
    uint32_t kmer = 0;
    double total = 0.0;
    for (int i = 0; i < len; ++i) {
        kmer = ((kmer << 2) | seq[i]) & 1023u;
        int j = 0;
        while (j < nk && kmers[j] != (int)kmer) ++j;
        if (j == nk) { kmers[nk] = (int)kmer; cnt[nk] = 0; ++nk; }
        total = __dsub_rn(total, tab[cnt[j]]);
        ++cnt[j];
        total = __dadd_rn(total, tab[cnt[j]]);
    }
    return __dmul_rn(total, mul);
}

__device__ __forceinline__ int warp_sum(int v) {
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(TB)
hard_filter_kernel(const Params P) {
    __shared__ uint32_t table[SMEM_WORDS];
    __shared__ int acc[N_ACC];
    __shared__ double tab[35];
    const int s = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const cto_hf_chunk_arrays& C = P.c;
    const cto_hf_site_arrays& S = P.s;
    const int row_lo = S.row_lo[s], row_hi = S.row_hi[s], centre = S.centre_row[s];
    const int kind = S.kind[s], alt_tok = S.alt_tok[s], del_len = S.del_len[s];
    const int span = S.rid_span[s];
    const bool hap_mode = P.mode == 1;
    State st{span <= CTO_HF_SMEM_READS ? table : P.scratch + S.scratch_off[s], S.rid_min[s]};
    for (int k = tid; k < (span + 3) / 4; k += TB) st.w[k] = 0;
    if (tid < N_ACC) acc[tid] = (tid == PASS_HET || tid == PASS_HOM) ? 1 : 0;
    if (tid < 35) tab[tid] = P.tab[tid];
    __syncthreads();

    // ---- haplotype of every read: HP tags of the site's row and of the heterozygous germline rows, in row order (HF:621-625)
    if (hap_mode) {
        for (int j = S.ph_off[s]; j < S.ph_off[s + 1]; ++j) {
            const int r = S.ph_row[j];
            for (int e = C.row_off[r] + tid; e < C.row_off[r + 1]; e += TB) {
                const uint32_t hp = C.info[e] & CTO_HF_HAP_MASK;
                if (hp) st.set_hap_once(C.rid[e], hp);
            }
            __syncthreads();
        }
    }

    // ---- the site's own row: depth by haplotype and strand, the alt read set and its counts (HF:632-688)
    if (centre >= 0) {
        for (int e = C.row_off[centre] + tid; e < C.row_off[centre + 1]; e += TB) {
            const uint32_t inf = C.info[e];
            const int32_t rid = C.rid[e];
            const int h = st.get(rid) & S_HAP;
            const bool rev = inf & CTO_HF_REV;
            atomicAdd(&acc[ALL_N + h], 1);
            atomicAdd(&acc[(rev ? ALL_R : ALL_F) + h], 1);
            bool m;
            if (kind <= 1) m = alt_tok >= 0 && C.tok[e] == alt_tok;
            else if (kind == 2) m = (inf & CTO_HF_MINUS) && (int)((inf >> CTO_HF_LEN_SHIFT) & 0xffffu) == del_len;
            else m = false;
            if (m) {
                const uint32_t q = C.qual[e];
                atomicAdd(&acc[BQ_SUM], (int)(q & 0xffu));
                atomicAdd(&acc[MQ_SUM], (int)(q >> 8));
                atomicAdd(&acc[N_MATCH], 1);
                if (!(st.fetch_or(rid, S_ALT) & S_ALT)) {
                    atomicAdd(&acc[ALT_N + h], 1);
                    atomicAdd(&acc[(rev ? ALT_R : ALT_F) + h], 1);
                }
            }
        }
    }
    __syncthreads();
    const int n_alt = acc[ALT_N] + acc[ALT_N + 1] + acc[ALT_N + 2];

    // ---- reads that start or end in a row where at least a fifth of the reads do (HF:627-628), among the alt reads
    if (!P.disable_rse) {
        for (int r = row_lo + tid; r < row_hi; r += TB) {
            if (!(C.row_flags[r] & CTO_HF_ROW_RSE)) continue;
            for (int j = C.rse_off[r]; j < C.rse_off[r + 1]; ++j) {
                const uint32_t old = st.fetch_or(C.rid[C.rse_ent[j]], S_RSE);
                if (!(old & S_RSE) && (old & S_ALT)) atomicAdd(&acc[RSE_HITS], 1);
            }
        }
    }
    // (the start/end bit is not read below, so no barrier is needed before the next phase)

    // ---- variant cluster: rows where most alt reads carry one and the same other allele that few other reads carry (HF:394-434)
    for (int r = row_lo + warp; r < row_hi; r += WARPS) {
        const uint32_t rf = C.row_flags[r];
        if (r == centre || (hap_mode && !(rf & CTO_HF_ROW_REF_OK))) continue;
        const int e0 = C.row_off[r], e1 = C.row_off[r + 1];
        int ins = 0, n_car = 0, cand = -1, votes = 0;
        for (int e = e0 + lane; e < e1; e += 32) {
            const uint32_t inf = C.info[e];
            if (inf & CTO_HF_SHADOW) continue;
            const int len = (inf >> CTO_HF_LEN_SHIFT) & 0xffffu;
            if ((inf & CTO_HF_PLUS) && len > 3) ins += min(len - 1, 2 * P.flanking);
            if ((st.get(C.rid[e]) & S_ALT) && !(inf & (CTO_HF_IS_REF | CTO_HF_STAR))) {
                const int t = C.tok[e];
                ++n_car;
                if (votes == 0) { cand = t; votes = 1; }
                else if (t == cand) ++votes;
                else --votes;
            }
        }
        ins = warp_sum(ins);
        n_car = warp_sum(n_car);
        if (lane == 0 && ins) atomicAdd(&acc[INS_LEN], ins);
        if (n_car == 0) continue;
        for (int d = 16; d; d >>= 1) {                            // merge the majority votes
            const int oc = __shfl_xor_sync(0xffffffffu, cand, d), ov = __shfl_xor_sync(0xffffffffu, votes, d);
            if (oc == cand) votes += ov;
            else if (ov > votes) { cand = oc; votes = ov - votes; }
            else votes -= ov;
        }
        cand = __shfl_sync(0xffffffffu, cand, 0);
        int top = 0, all = 0;
        for (int e = e0 + lane; e < e1; e += 32) {
            if (C.tok[e] != cand) continue;
            ++all;                                                // the row's Counter counts every entry (HF:184)
            const uint32_t inf = C.info[e];
            if (!(inf & CTO_HF_SHADOW) && (st.get(C.rid[e]) & S_ALT) && !(inf & (CTO_HF_IS_REF | CTO_HF_STAR))) ++top;
        }
        top = warp_sum(top);
        all = warp_sum(all);
        if (lane == 0) {
            const bool out_of_bounds = 2 * top >= 3 * n_alt || 2 * top <= n_alt;          // HF:423-424, eps = 0.5
            if (!out_of_bounds && (rf & CTO_HF_ROW_COUNTER) && !(2 * all >= 3 * top)) atomicAdd(&acc[MATCH_COUNT], 1);
        }
    }

    // ---- consistency with nearby germline variants (HF:436-523)
    if (hap_mode) {
        const int h1 = acc[ALT_N + 1], h2 = acc[ALT_N + 2];
        const int hi = max(h1, h2), lo = min(h1, h2);
        const bool phasable = (int64_t)h1 * h2 == 0 || (hi >= 5 * lo && (h1 > P.max_co || h2 > P.max_co));   // HF:387
        const uint32_t hap_index = !phasable ? 0u : (h1 > h2 ? 1u : 2u);
        if (hap_index > 0) {
            for (int j = S.het_off[s] + warp; j < S.het_off[s + 1]; j += WARPS) {
                const int g = S.het_idx[j], r = S.g_row[g];
                if (r < row_lo || r >= row_hi || !(C.row_flags[r] & CTO_HF_ROW_REF_OK)) continue;
                const int e0 = C.row_off[r], n = C.row_off[r + 1] - e0;
                const uint8_t* gm = S.g_match + S.g_off[g];
                int overlap = 0, phased = 0, inter = 0;
                for (int k = lane; k < n; k += 32) {
                    if ((C.info[e0 + k] & CTO_HF_SHADOW) || !(gm[k] & 1)) continue;
                    ++overlap;
                    const uint32_t v = st.get(C.rid[e0 + k]);
                    if ((v & S_HAP) == hap_index) { ++phased; if (v & S_ALT) ++inter; }
                }
                overlap = warp_sum(overlap); phased = warp_sum(phased); inter = warp_sum(inter);
                if (lane == 0 && !(phased == 0 || 2 * phased < overlap) && inter == 0) acc[PASS_HET] = 0;
            }
        }
        for (int j = S.hom_off[s] + warp; j < S.hom_off[s + 1]; j += WARPS) {
            const int g = S.hom_idx[j], r = S.g_row[g];
            if (r < row_lo || r >= row_hi || !(C.row_flags[r] & CTO_HF_ROW_REF_OK)) continue;
            const int e0 = C.row_off[r], n = C.row_off[r + 1] - e0;
            const uint8_t* gm = S.g_match + S.g_off[g];
            int a1 = 0, a2 = 0, a_all = 0, c0 = 0, c1 = 0, c2 = 0, inter = 0, both = 0;
            for (int k = lane; k < n; k += 32) {
                if (C.info[e0 + k] & CTO_HF_SHADOW) continue;
                const uint32_t v = st.get(C.rid[e0 + k]);
                const uint32_t h = v & S_HAP;
                const bool m = gm[k] & 2;
                ++a_all; a1 += h == 1; a2 += h == 2;
                if (m) { c0 += h == 0; c1 += h == 1; c2 += h == 2; }
                if (v & S_ALT) { ++inter; both += m; }
            }
            a1 = warp_sum(a1); a2 = warp_sum(a2); a_all = warp_sum(a_all); c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
            inter = warp_sum(inter); both = warp_sum(both);
            if (lane == 0) {
                const int carriers = c0 + c1 + c2;
                const bool af_low = a_all > 0 ? 4 * carriers < 3 * a_all : true;                  // af_g < 0.75 (HF:490)
                bool allele_phasable = true;                                                       // HF:492-503
                if ((int64_t)a1 * a2 == 0) allele_phasable = false;
                else if ((int64_t)c1 * c2 > 0 && max(c1, c2) <= 10 * min(c1, c2)) allele_phasable = false;
                if (!(af_low || allele_phasable) && inter > 0 && (both == 0 || 2 * both < inter)) acc[PASS_HOM] = 0;
            }
        }
    }
    __syncthreads();
    if (tid != 0) return;

    // ---- verdict (HF:525-565 / PV:343-365)
    const int all_n[3] = {acc[ALL_N], acc[ALL_N + 1], acc[ALL_N + 2]};
    const int alt_n[3] = {acc[ALT_N], acc[ALT_N + 1], acc[ALT_N + 2]};
    const bool is_snp = kind == 0;
    bool pass_rse = true;
    if (!P.disable_rse && (n_alt > 0 || !hap_mode))
        if ((double)acc[RSE_HITS] >= __dmul_rn(0.3, (double)n_alt)) pass_rse = false;               // HF:371-373 / PV:294-296
    bool pass_bq = true, pass_mq = true, pass_both = true;
    if (hap_mode) {
        if (acc[N_MATCH] > 0 && acc[BQ_SUM] <= 20 * acc[N_MATCH]) pass_bq = false;                   // HF:659-660 (mean <= 20)
        if (acc[N_MATCH] > 0 && acc[MQ_SUM] <= 20 * acc[N_MATCH]) pass_mq = false;
        const int h1 = alt_n[1], h2 = alt_n[2], hi = max(h1, h2), lo = min(h1, h2);
        if (S.low_af[s] && (int64_t)h1 * h2 > 0 && (lo > P.max_co || hi <= 10 * lo)) pass_both = false;   // HF:380-385
    }
    const int depth = all_n[0] + all_n[1] + all_n[2] > 0 ? all_n[0] + all_n[1] + all_n[2] : 1;
    const bool pass_co_exist = !(acc[MATCH_COUNT] >= P.max_co || (int64_t)acc[INS_LEN] > 3ll * depth);   // HF:527-528
    const bool phaseable = (int64_t)all_n[1] * all_n[2] > 0 && (int64_t)alt_n[1] * alt_n[2] == 0 &&
                           (alt_n[1] > P.max_co || alt_n[2] > P.max_co);                              // HF:532-533
    const int64_t a0 = acc[ALT_F] + acc[ALT_F + 1] + acc[ALT_F + 2], a1 = acc[ALT_R] + acc[ALT_R + 1] + acc[ALT_R + 2];
    const int64_t r0 = acc[ALL_F] + acc[ALL_F + 1] + acc[ALL_F + 2] - a0, r1 = acc[ALL_R] + acc[ALL_R + 1] + acc[ALL_R + 2] - a1;
    const double p_value = fisher_exact(a0, r0, a1, r1);
    bool pass_sb;
    if (hap_mode) pass_sb = !(p_value < (is_snp ? 0.001 : 0.01) || a0 == 0 || a1 == 0);             // HF:545-548
    else pass_sb = !(p_value < 0.001);                                                               // PV:353-354
    bool pass_entropy = true;
    if (!is_snp && sequence_entropy(S.seq + S.seq_off[s], S.seq_len[s], tab, P.mul) < P.thr) pass_entropy = false;
    const bool pass_het = acc[PASS_HET], pass_hom = acc[PASS_HOM];
    const bool verdict = pass_het && pass_hom && pass_both && pass_rse && pass_bq && pass_mq && pass_co_exist && pass_sb && pass_entropy;
    uint32_t f = 0;
    if (verdict) f |= CTO_HFO_VERDICT;
    if (phaseable) f |= CTO_HFO_PHASEABLE;
    if (pass_het) f |= CTO_HFO_HETERO;
    if (pass_hom) f |= CTO_HFO_HOMO;
    if (pass_rse) f |= CTO_HFO_READ_START_END;
    if (pass_bq) f |= CTO_HFO_BQ;
    if (pass_mq) f |= CTO_HFO_MQ;
    if (pass_co_exist) f |= CTO_HFO_CO_EXIST;
    if (pass_both) f |= CTO_HFO_BOTH_SIDE;
    if (pass_sb) f |= CTO_HFO_STRAND_BIAS;
    if (pass_entropy) f |= CTO_HFO_ENTROPY;
    P.out_flags[s] = f;
    P.out_p[s] = p_value;
    if (P.out_counts) {
        int32_t* o = P.out_counts + 8ll * s;
        o[0] = (int32_t)a0; o[1] = (int32_t)r0; o[2] = (int32_t)a1; o[3] = (int32_t)r1;
        o[4] = acc[MATCH_COUNT]; o[5] = acc[INS_LEN]; o[6] = depth; o[7] = n_alt;
    }
}

}  // namespace hf
}  // namespace cto

using namespace cto;

extern "C" int cto_hard_filter_sites(const cto_hf_chunk_arrays* chunk, const cto_hf_site_arrays* sites, int mode, int flanking,
                                     int disable_read_start_end_filtering, int max_co_exist_read_num, const double* entropy_tab,
                                     double entropy_mul, double entropy_threshold, uint32_t* scratch, uint32_t* out_flags,
                                     double* out_p, int32_t* out_counts, void* stream) {
    CTO_REQUIRE(chunk && sites && entropy_tab, "hard_filter_sites: NULL argument");
    CTO_REQUIRE(mode == 0 || mode == 1, "hard_filter_sites: mode %d (0 = post filter, 1 = haplotype filter)", mode);
    CTO_REQUIRE(sites->n_sites >= 0 && sites->n_sites < (1ll << 31), "hard_filter_sites: %lld sites", (long long)sites->n_sites);
    if (sites->n_sites == 0) return 0;
    CTO_REQUIRE(out_flags && out_p, "hard_filter_sites: NULL output");
    CTO_REQUIRE(flanking >= 0 && flanking < (1 << 20), "hard_filter_sites: flanking %d", flanking);
    if (cto_device_check(nullptr)) return 3;
    hf::Params p;
    p.c = *chunk;
    p.s = *sites;
    p.mode = mode; p.flanking = flanking; p.disable_rse = disable_read_start_end_filtering; p.max_co = max_co_exist_read_num;
    for (int i = 0; i < 35; ++i) p.tab[i] = entropy_tab[i];
    p.mul = entropy_mul; p.thr = entropy_threshold;
    p.scratch = scratch; p.out_flags = out_flags; p.out_p = out_p; p.out_counts = out_counts;
    hf::hard_filter_kernel<<<(unsigned)sites->n_sites, hf::TB, 0, (cudaStream_t)stream>>>(p);
    CTO_CHECK(cudaGetLastError());
    count_launch();
    return 0;
}
