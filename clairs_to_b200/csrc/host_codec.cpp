// Host side of the C ABI: mpileup tokenizer and chunk-file text codec (no CUDA).
//
// Tokenizer semantics follow src/create_tensor_pileup_calling.py:120-144 (token loop), 147-149
// (positional zip with the quality strings), 158-209 (alt_info), 472-497 (row split); cited as CT.
#include "../../include/clairs_to_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <string>
#include <thread>
#include <chrono>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace cto {
void set_error(const char* fmt, ...);
}

namespace {

constexpr uint8_t HAS_INDEL = 0x10;
constexpr uint8_t QUAL_ABSENT = 254;
constexpr uint32_t IND_DEL = 1u << 24, IND_REV = 1u << 25, IND_LONG = 1u << 26;

struct SymbolTable {
    int8_t idx[256];
    SymbolTable() {
        for (int k = 0; k < 256; ++k) idx[k] = -1;
        const char* order = "ACGTNacgtn*#";
        for (int k = 0; order[k]; ++k) idx[(uint8_t)order[k]] = (int8_t)k;
    }
};
const SymbolTable SYMBOLS;
inline int symbol_index(char c) { return SYMBOLS.idx[(uint8_t)c]; }

inline char upper(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }

inline int coerce_ref(char c) {            // evc_base_from + upper (CT:82-92, 485)
    switch (upper(c)) {
        case 'C': return 1; case 'G': return 2; case 'T': return 3;
        default: return 0;                 // A, N and every IUPAC code become A
    }
}

struct Entry {
    char sym;
    char sign;                             // '+', '-' or 0 when the read carries no indel
    int32_t seq_begin, seq_len;            // indel sequence as a slice of the row's bases string
};

}  // namespace

struct cto_tokens {
    std::vector<uint8_t> code, bq, mq, ref_code;
    std::vector<int32_t> pos_off, ind_off;
    std::vector<uint32_t> ind_entry;
    std::vector<int64_t> row_pos;
    std::string alt_info;
    std::vector<int64_t> alt_off;
};

namespace {

// one contiguous range of whole rows -> its own token block (rows are independent)
int tokenize_range(const char* text, int64_t text_len, const char* ref_seq, int64_t ref_len, int64_t ref_start,
                   const std::unordered_set<int64_t>& cand, int max_indel_length, cto_tokens* t, std::string& err) {
    t->pos_off.push_back(0);
    t->ind_off.push_back(0);
    t->alt_off.push_back(0);
    {   // a row of depth d takes about 3 d + 30 characters: reserve once instead of growing by doubling
        const size_t reads = (size_t)text_len / 3 + 64, rows = (size_t)text_len / 64 + 64;
        t->code.reserve(reads); t->bq.reserve(reads); t->mq.reserve(reads);
        t->ref_code.reserve(rows); t->row_pos.reserve(rows); t->pos_off.reserve(rows); t->ind_off.reserve(rows); t->alt_off.reserve(rows);
    }
    std::vector<Entry> entries;
    std::string key;
    std::unordered_map<std::string, uint32_t> allele_ids;
    std::vector<std::pair<std::string, int>> main_counts;          // first-occurrence ordered Counter
    std::unordered_map<std::string, size_t> main_index;
    std::vector<std::pair<std::string, int>> alt;                  // ordered alt_info_dict
    std::unordered_map<std::string, size_t> alt_index;

    const char* p = text;
    const char* end = text + text_len;
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', end - p);
        if (!eol) eol = end;
        const char* col[8];
        int col_len[8];
        int ncol = 0;
        const char* q = p;
        while (q <= eol && ncol < 8) {
            const char* tab = (const char*)memchr(q, '\t', eol - q);
            if (!tab) tab = eol;
            col[ncol] = q;
            col_len[ncol] = (int)(tab - q);
            ++ncol;
            q = tab + 1;
        }
        const char* next = eol + 1;
        if (ncol == 0 || (ncol == 1 && col_len[0] == 0)) { p = next; continue; }
        if (ncol < 7) {
            err = "tokenize_mpileup: row with " + std::to_string(ncol) + " columns (need 7: chr pos ref depth bases BQ MQ)";
            return 2;
        }
        while (col_len[6] > 0 && (col[6][col_len[6] - 1] == '\r' || col[6][col_len[6] - 1] == ' ')) --col_len[6];   // row.strip()
        const int64_t pos = strtoll(col[1], nullptr, 10);
        const int64_t roff = pos - ref_start;
        if (roff < 0 || roff >= ref_len) {
            err = "tokenize_mpileup: position " + std::to_string((long long)pos) + " outside the reference window";
            return 2;
        }
        const int ref = coerce_ref(ref_seq[roff]);
        const char* bases = col[4];
        const int nb = col_len[4];

        entries.clear();
        for (int i = 0; i < nb;) {                                   // CT:120-144
            const char ch = bases[i];
            if (ch == '+' || ch == '-') {
                int j = i + 1, len = 0;
                while (j < nb && bases[j] >= '0' && bases[j] <= '9') { len = len * 10 + (bases[j] - '0'); ++j; }
                if (!entries.empty()) {                              // attaches to (overwrites on) the previous entry
                    const int avail = nb - j > 0 ? nb - j : 0;
                    entries.back().sign = ch;
                    entries.back().seq_begin = j;
                    entries.back().seq_len = len < avail ? len : avail;
                }
                i = j + len;
                continue;
            }
            if (symbol_index(ch) >= 0) entries.push_back(Entry{ch, 0, 0, 0});
            else if (ch == '^') ++i;
            ++i;
        }

        const int n_mq = col_len[6], n_bq = col_len[5];
        if (!allele_ids.empty()) allele_ids.clear();             // clear() walks the whole bucket array even when the map is empty
        main_counts.clear();
        if (!main_index.empty()) main_index.clear();
        const bool is_cand = cand.count(pos) != 0;
        const size_t read_base = t->code.size();
        t->code.resize(read_base + entries.size());
        t->mq.resize(read_base + entries.size());
        t->bq.resize(read_base + entries.size());
        uint8_t* out_code = t->code.data() + read_base;
        uint8_t* out_mq = t->mq.data() + read_base;
        uint8_t* out_bq = t->bq.data() + read_base;
        for (size_t k = 0; k < entries.size(); ++k) {
            const Entry& e = entries[k];
            const uint8_t mqv = (int)k < n_mq ? (uint8_t)(col[6][k] - 33) : QUAL_ABSENT;
            const uint8_t bqv = (int)k < n_bq ? (uint8_t)(col[5][k] - 33) : QUAL_ABSENT;
            uint8_t c = (uint8_t)symbol_index(e.sym);
            if (e.sign || is_cand) key.assign(1, e.sym);
            if (e.sign) {
                const bool is_del = e.sign == '-';
                const int seq_len = e.seq_len;
                key.push_back(e.sign);
                key.append(bases + e.seq_begin, seq_len);
                c |= HAS_INDEL;
                auto it = allele_ids.find(key);
                uint32_t id;
                if (it == allele_ids.end()) {
                    id = (uint32_t)allele_ids.size();
                    allele_ids.emplace(key, id);
                } else {
                    id = it->second;
                }
                uint32_t ent = (id & 0xFFFF) | ((uint32_t)mqv << 16);
                if (is_del) ent |= IND_DEL;
                const char s = e.sym;
                const bool fwd = s == 'A' || s == 'C' || s == 'G' || s == 'T' || s == 'N' || s == '*';   // CT:182, 199
                if (!fwd) ent |= IND_REV;
                const int span = is_del ? seq_len + 1 : seq_len;     // CT:174, 189 (deletion counts the sign)
                if (span > max_indel_length) ent |= IND_LONG;
                t->ind_entry.push_back(ent);
            }
            out_code[k] = c;
            out_mq[k] = mqv;
            out_bq[k] = bqv;
            if (is_cand && mqv != QUAL_ABSENT && mqv >= 20) {        // base_counter, CT:147
                auto it = main_index.find(key);
                if (it == main_index.end()) {
                    main_index.emplace(key, main_counts.size());
                    main_counts.emplace_back(key, 1);
                } else {
                    ++main_counts[it->second].second;
                }
            }
        }
        t->pos_off.push_back((int32_t)t->code.size());
        t->ind_off.push_back((int32_t)t->ind_entry.size());
        t->ref_code.push_back((uint8_t)ref);
        t->row_pos.push_back(pos);

        if (is_cand) {                                               // CT:153-209
            alt.clear();
            alt_index.clear();
            auto bump = [&](const std::string& k, int n) {
                auto it = alt_index.find(k);
                if (it == alt_index.end()) {
                    alt_index.emplace(k, alt.size());
                    alt.emplace_back(k, n);
                } else {
                    alt[it->second].second += n;
                }
            };
            const char ref_char = "ACGT"[ref];
            int depth = 0, ref_count = 0;
            for (const auto& kc : main_counts) {
                const std::string& key = kc.first;
                const int count = kc.second;
                if (key.size() == 1) {
                    const char u = upper(key[0]);
                    if (u == 'A' || u == 'C' || u == 'G' || u == 'T') {
                        if (u != ref_char) bump(std::string("X") + u, count);
                        else ref_count += count;
                        depth += count;
                    } else if (key[0] == '#' || key[0] == '*') {
                        depth += count;
                    }
                } else if (key[1] == '+') {
                    if ((int)key.size() - 2 > max_indel_length) continue;
                    depth += count;
                    std::string name = "I";
                    name.push_back(upper(key[0]));
                    for (size_t z = 2; z < key.size(); ++z) name.push_back(upper(key[z]));
                    bump(name, count);
                } else {
                    const int span = (int)key.size() - 1;
                    if (span > max_indel_length) continue;
                    depth += count;
                    std::string name = "D";
                    for (int z = 0; z < span && z < max_indel_length && roff + z < ref_len; ++z) name.push_back(upper(ref_seq[roff + z]));
                    bump(name, count);
                }
            }
            if (ref_count > 0) {
                std::string name = "R";
                name.push_back(ref_char);
                auto it = alt_index.find(name);
                if (it == alt_index.end()) alt.emplace_back(name, ref_count);
                else alt[it->second].second = ref_count;
            }
            char buf[32];
            snprintf(buf, sizeof(buf), "%d-", depth);
            t->alt_info += buf;
            for (size_t z = 0; z < alt.size(); ++z) {
                if (z) t->alt_info.push_back(' ');
                t->alt_info += alt[z].first;
                snprintf(buf, sizeof(buf), " %d", alt[z].second);
                t->alt_info += buf;
            }
            t->alt_info.push_back('-');
        }
        t->alt_off.push_back((int64_t)t->alt_info.size());
        p = next;
    }
    return 0;
}

// 8 packed read bytes (read i in byte i) -> 8 bit-plane bytes (plane j in byte j; bit i of plane j = bit j of read i)
inline uint64_t transpose8x8(uint64_t x) {
    uint64_t t;
    t = (x ^ (x >> 7)) & 0x00AA00AA00AA00AAULL;  x = x ^ t ^ (t << 7);
    t = (x ^ (x >> 14)) & 0x0000CCCC0000CCCCULL; x = x ^ t ^ (t << 14);
    t = (x ^ (x >> 28)) & 0x00000000F0F0F0F0ULL; x = x ^ t ^ (t << 28);
    return x;
}

int resolve_threads(int n_threads, int64_t work_items, int64_t min_items_per_thread) {
    int n = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (n < 1) n = 1;
    if (n > 64) n = 64;
    const int64_t cap = work_items / (min_items_per_thread > 0 ? min_items_per_thread : 1);
    if (cap < n) n = cap < 1 ? 1 : (int)cap;
    return n;
}

}  // namespace

extern "C" {

int cto_tokenize_mpileup(const char* text, int64_t text_len, const char* ref_seq, int64_t ref_len, int64_t ref_start,
                         const int64_t* candidate_pos, int64_t n_candidates, int max_indel_length, int n_threads,
                         cto_tokens** out) {
    if (!text || !ref_seq || !out) {
        cto::set_error("tokenize_mpileup: NULL argument");
        return 2;
    }
    std::unordered_set<int64_t> cand;
    for (int64_t i = 0; i < n_candidates; ++i) cand.insert(candidate_pos[i]);
    const int nt = resolve_threads(n_threads, text_len, 1 << 20);
    // split at row boundaries
    std::vector<int64_t> cut(nt + 1, 0);
    cut[nt] = text_len;
    for (int k = 1; k < nt; ++k) {
        int64_t c = text_len * k / nt;
        if (c < cut[k - 1]) c = cut[k - 1];
        const char* nl = (const char*)memchr(text + c, '\n', (size_t)(text_len - c));
        cut[k] = nl ? (int64_t)(nl - text) + 1 : text_len;
    }
    const bool timing = getenv("CTO_TOKENIZE_TIMING") != nullptr;
    const auto t_start = std::chrono::steady_clock::now();
    std::vector<cto_tokens> parts(nt);
    std::vector<std::string> errs(nt);
    std::vector<int> rcs(nt, 0);
    auto work = [&](int k) {
        const auto w0 = std::chrono::steady_clock::now();
        rcs[k] = tokenize_range(text + cut[k], cut[k + 1] - cut[k], ref_seq, ref_len, ref_start, cand, max_indel_length, &parts[k], errs[k]);
        if (timing) fprintf(stderr, "[tokenize]   part %d: %lld bytes in %.3f s\n", k, (long long)(cut[k + 1] - cut[k]),
                            std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count());
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back(work, k);
        for (auto& x : th) x.join();
    }
    for (int k = 0; k < nt; ++k)
        if (rcs[k]) {
            cto::set_error("%s", errs[k].c_str());
            return rcs[k];
        }
    const auto t_tok = std::chrono::steady_clock::now();
    cto_tokens* t = new cto_tokens();
    if (nt == 1) {
        *t = std::move(parts[0]);
    } else {
        size_t n_reads = 0, n_rows = 0, n_ind = 0, n_alt = 0;
        for (auto& p : parts) { n_reads += p.code.size(); n_rows += p.ref_code.size(); n_ind += p.ind_entry.size(); n_alt += p.alt_info.size(); }
        if (n_reads >= (size_t)INT32_MAX) {
            cto::set_error("tokenize_mpileup: %zu reads in one call (limit 2^31)", n_reads);
            delete t;
            return 2;
        }
        t->code.reserve(n_reads); t->bq.reserve(n_reads); t->mq.reserve(n_reads);
        t->ref_code.reserve(n_rows); t->row_pos.reserve(n_rows); t->ind_entry.reserve(n_ind);
        t->pos_off.reserve(n_rows + 1); t->ind_off.reserve(n_rows + 1); t->alt_off.reserve(n_rows + 1);
        t->alt_info.reserve(n_alt);
        t->pos_off.push_back(0); t->ind_off.push_back(0); t->alt_off.push_back(0);
        for (auto& p : parts) {
            const int32_t r0 = (int32_t)t->code.size(), i0 = (int32_t)t->ind_entry.size();
            const int64_t a0 = (int64_t)t->alt_info.size();
            t->code.insert(t->code.end(), p.code.begin(), p.code.end());
            t->bq.insert(t->bq.end(), p.bq.begin(), p.bq.end());
            t->mq.insert(t->mq.end(), p.mq.begin(), p.mq.end());
            t->ref_code.insert(t->ref_code.end(), p.ref_code.begin(), p.ref_code.end());
            t->row_pos.insert(t->row_pos.end(), p.row_pos.begin(), p.row_pos.end());
            t->ind_entry.insert(t->ind_entry.end(), p.ind_entry.begin(), p.ind_entry.end());
            t->alt_info += p.alt_info;
            for (size_t r = 1; r < p.pos_off.size(); ++r) {
                t->pos_off.push_back(p.pos_off[r] + r0);
                t->ind_off.push_back(p.ind_off[r] + i0);
                t->alt_off.push_back(p.alt_off[r] + a0);
            }
        }
    }
    if (timing)
        fprintf(stderr, "[tokenize] %d threads: tokenize %.3f s, merge %.3f s\n", nt,
                std::chrono::duration<double>(t_tok - t_start).count(),
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_tok).count());
    *out = t;
    return 0;
}

// A decompressed predict chunk file (rows of clairs/predict.py:114-152) in one pass: replaces the row loop of
// clairs/call_variants.py:798-829.  Per row: fields[r][k] = (offset, length) of chrom, pos, ref, alt_info and the two strand
// list-reprs; p_aff / p_neg [r][h] = the SECOND number of each "p0 p1" probability field parsed with strtod (correctly
// rounded, i.e. the same double python's float() gives the reference).  Rows with fewer than 6 + 2 * n_heads fields are
// an error (the reference would raise on them too).
int cto_parse_predict_file(const char* text, int64_t len, int n_heads, int64_t max_rows, double* p_aff, double* p_neg,
                           int64_t* fields, int64_t* n_rows) {
    if (!text || !p_aff || !p_neg || !fields || !n_rows || (n_heads != 4 && n_heads != 6)) {
        cto::set_error("parse_predict_file: bad argument");
        return 2;
    }
    int64_t r = 0;
    const char* p = text;
    const char* end = text + len;
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        if (eol > p) {
            if (r >= max_rows) {
                cto::set_error("parse_predict_file: more than %lld rows", (long long)max_rows);
                return 2;
            }
            const char* q = p;
            const char* stop = eol;
            while (stop > q && (stop[-1] == '\r' || stop[-1] == ' ' || stop[-1] == '\t')) --stop;   // line.rstrip()
            for (int k = 0; k < 6 + 2 * n_heads; ++k) {
                if (q > stop) {
                    cto::set_error("parse_predict_file: row %lld has %d fields, expected %d", (long long)r, k, 6 + 2 * n_heads);
                    return 2;
                }
                const char* tab = (const char*)memchr(q, '\t', (size_t)(stop - q));
                if (!tab) tab = stop;
                if (k < 6) {
                    fields[(r * 6 + k) * 2] = (int64_t)(q - text);
                    fields[(r * 6 + k) * 2 + 1] = (int64_t)(tab - q);
                } else {
                    char* e1 = nullptr;
                    strtod(q, &e1);                                   // p0
                    char* e2 = nullptr;
                    const double p1 = strtod(e1, &e2);
                    if (e1 == q || e2 == e1 || e2 > tab) {
                        cto::set_error("parse_predict_file: row %lld field %d is not 'p0 p1'", (long long)r, k);
                        return 2;
                    }
                    const int h = k - 6;
                    if (h < n_heads) p_aff[r * n_heads + h] = p1;
                    else p_neg[r * n_heads + (h - n_heads)] = p1;
                }
                q = tab + 1;
            }
            ++r;
        }
        p = eol + 1;
    }
    *n_rows = r;
    return 0;
}

// Synthetic `samtools mpileup` text for a read-array stream (the inverse of the tokenizer; bench.py's text-in leg and the
// tests render the same sites as text): rows "ctg \t pos \t N \t depth \t bases \t BQ \t MQ \n", position = first_pos + row.
// ind_len / ind_seq: per indel-carrying read (in read order) the indel length and, for insertions, the inserted bases
// packed two bits each (deletions print N / n like samtools without -f).  Returns bytes written, -1 if cap is too small.
int64_t cto_render_mpileup(const uint8_t* code, const uint8_t* bq, const uint8_t* mq, const int32_t* pos_off, int64_t n_rows,
                           const uint32_t* ind_entry, const int64_t* ind_len, const int64_t* ind_seq, const char* ctg,
                           int64_t first_pos, char* out, int64_t cap) {
    static const char SYM[] = "ACGTNacgtn*#";
    int64_t w = 0, k = 0;
    const size_t ctg_len = strlen(ctg);
    for (int64_t r = 0; r < n_rows; ++r) {
        const int32_t lo = pos_off[r], hi = pos_off[r + 1];
        // worst case per read: symbol + sign + 3 digits + 60 bases, + 2 quality characters
        if (w + (int64_t)ctg_len + 64 + (int64_t)(hi - lo) * 70 > cap) return -1;
        memcpy(out + w, ctg, ctg_len); w += ctg_len;
        w += snprintf(out + w, 48, "\t%lld\tN\t%d\t", (long long)(first_pos + r), hi - lo);
        if (hi == lo) {
            memcpy(out + w, "*\t*\t*\n", 6); w += 6;
            continue;
        }
        for (int32_t i = lo; i < hi; ++i) {
            const uint8_t c = code[i];
            out[w++] = SYM[c & 0xF];
            if (c & HAS_INDEL) {
                const uint32_t e = ind_entry[k];
                const bool is_del = e & IND_DEL, rev = e & IND_REV;
                const int64_t len = ind_len[k];
                int64_t seq = ind_seq[k];
                ++k;
                w += snprintf(out + w, 16, "%c%lld", is_del ? '-' : '+', (long long)len);
                for (int64_t z = 0; z < len; ++z) {
                    char b = is_del ? 'N' : "ACGT"[seq & 3];
                    seq >>= 2;
                    out[w++] = rev ? (char)(b + 32) : b;
                }
            }
        }
        out[w++] = '\t';
        for (int32_t i = lo; i < hi; ++i) out[w++] = (char)(33 + bq[i]);
        out[w++] = '\t';
        for (int32_t i = lo; i < hi; ++i) out[w++] = (char)(33 + mq[i]);
        out[w++] = '\n';
    }
    return w;
}

int cto_pack_reads(const uint8_t* code, const uint8_t* bq, const uint8_t* mq, const int32_t* pos_off, int64_t n_rows,
                   int low_bq_cut, const int32_t* grp_off, uint8_t* planes_out, int n_threads) {
    if (n_rows < 0 || !pos_off || !grp_off || (n_rows > 0 && grp_off[n_rows] > 0 && (!code || !bq || !mq || !planes_out))) {
        cto::set_error("pack_reads: NULL / negative argument");
        return 2;
    }
    // old symbol order "ACGTNacgtn*#" -> packed order ACGT acgt * # N n
    static const uint8_t RECODE[16] = {0, 1, 2, 3, 10, 4, 5, 6, 7, 11, 8, 9, 15, 15, 15, 15};
    const int64_t total = (int64_t)grp_off[n_rows] * 8;
    const int64_t padded = (total + 15) & ~int64_t(15);
    for (int64_t i = total; i < padded; ++i) planes_out[i] = 0;
    const int nt = resolve_threads(n_threads, n_rows, 4096);
    int bad = 0;
    auto work = [&](int k) {
        const int64_t r_lo = n_rows * k / nt, r_hi = n_rows * (k + 1) / nt;
        for (int64_t r = r_lo; r < r_hi; ++r) {
            const int32_t lo = pos_off[r], hi = pos_off[r + 1];
            const int64_t g0 = grp_off[r], g1 = grp_off[r + 1];
            if (hi < lo || (int64_t)(hi - lo + 7) / 8 != g1 - g0) { __atomic_store_n(&bad, 1, __ATOMIC_RELAXED); return; }
            for (int64_t g = g0; g < g1; ++g) {
                uint64_t x = 0;
                const int32_t base = lo + (int32_t)(g - g0) * 8;
                for (int i = 0; i < 8; ++i) {
                    const int32_t idx = base + i;
                    if (idx >= hi) break;
                    const uint8_t c = code[idx], m = mq[idx], q = bq[idx];
                    uint8_t b = RECODE[c & 0xF];
                    if (!(c & HAS_INDEL)) b |= 0x10;                                   // plain read
                    if (m != QUAL_ABSENT) b |= (m >= 20) ? 0x20 : 0x40;                // CT:147-148
                    if (q != QUAL_ABSENT && (int)q < low_bq_cut) b |= 0x80;            // CT:149
                    x |= (uint64_t)b << (8 * i);
                }
                x = transpose8x8(x);
                memcpy(planes_out + g * 8, &x, 8);
            }
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; ++k) th.emplace_back(work, k);
        for (auto& x : th) x.join();
    }
    if (bad) {
        cto::set_error("pack_reads: grp_off is not the prefix sum of ceil(row depth / 8)");
        return 2;
    }
    return 0;
}

int cto_tokens_sizes(const cto_tokens* t, int64_t* n_reads, int64_t* n_rows, int64_t* n_ind, int64_t* alt_info_bytes) {
    if (!t) return 2;
    if (n_reads) *n_reads = (int64_t)t->code.size();
    if (n_rows) *n_rows = (int64_t)t->ref_code.size();
    if (n_ind) *n_ind = (int64_t)t->ind_entry.size();
    if (alt_info_bytes) *alt_info_bytes = (int64_t)t->alt_info.size();
    return 0;
}

int cto_tokens_export(const cto_tokens* t, uint8_t* code, uint8_t* bq, uint8_t* mq, int32_t* pos_off, uint8_t* ref_code,
                      int32_t* ind_off, uint32_t* ind_entry, int64_t* row_pos, char* alt_info, int64_t* alt_info_off) {
    if (!t) return 2;
    auto cp = [](void* dst, const void* src, size_t n) { if (dst && n) memcpy(dst, src, n); };
    cp(code, t->code.data(), t->code.size());
    cp(bq, t->bq.data(), t->bq.size());
    cp(mq, t->mq.data(), t->mq.size());
    cp(pos_off, t->pos_off.data(), t->pos_off.size() * sizeof(int32_t));
    cp(ref_code, t->ref_code.data(), t->ref_code.size());
    cp(ind_off, t->ind_off.data(), t->ind_off.size() * sizeof(int32_t));
    cp(ind_entry, t->ind_entry.data(), t->ind_entry.size() * sizeof(uint32_t));
    cp(row_pos, t->row_pos.data(), t->row_pos.size() * sizeof(int64_t));
    cp(alt_info, t->alt_info.data(), t->alt_info.size());
    cp(alt_info_off, t->alt_off.data(), t->alt_off.size() * sizeof(int64_t));
    return 0;
}

void cto_tokens_destroy(cto_tokens* t) { delete t; }

// ---- chunk-file text codec ------------------------------------------------------------------

static inline char* put_int(char* o, int v) {
    if (v < 0) { *o++ = '-'; v = -v; }
    char tmp[12];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *o++ = tmp[--n];
    return o;
}

int64_t cto_format_tensor_row(const int16_t* row, char* out, int64_t cap) {
    const int n = CTO_N_POS * CTO_N_CH;
    if (cap < (int64_t)n * 7 + 1) return -1;                 // "-32768 " is 7 bytes
    char* o = out;
    for (int i = 0; i < n; ++i) {
        if (i) *o++ = ' ';
        o = put_int(o, row[i]);
    }
    *o = 0;
    return (int64_t)(o - out);
}

int64_t cto_format_prob_fields(const float* probs, int n_pairs, char* out, int64_t cap) {
    char* o = out;
    for (int i = 0; i < n_pairs; ++i) {
        if (cap - (o - out) < 64) return -1;
        if (i) *o++ = '\t';
        // "{:0.8f}".format(np.float32) formats the exact double value of the float (clairs/predict.py:121)
        o += snprintf(o, 32, "%.8f", (double)probs[2 * i]);
        *o++ = ' ';
        o += snprintf(o, 32, "%.8f", (double)probs[2 * i + 1]);
    }
    *o = 0;
    return (int64_t)(o - out);
}

int cto_parse_tensor_row(const char* text, int64_t len, int16_t* row) {
    const int n = CTO_N_POS * CTO_N_CH;
    const char* p = text;
    const char* end = text + len;
    for (int i = 0; i < n; ++i) {
        while (p < end && (*p == ' ' || *p == '\t')) ++p;
        if (p >= end) {
            cto::set_error("parse_tensor_row: %d values found, %d expected", i, n);
            return 2;
        }
        bool negv = false;
        if (*p == '-') { negv = true; ++p; }
        int v = 0;
        bool any = false;
        while (p < end && *p >= '0' && *p <= '9') { v = v * 10 + (*p - '0'); ++p; any = true; }
        if (!any) {
            cto::set_error("parse_tensor_row: non-integer token at value %d", i);
            return 2;
        }
        row[i] = (int16_t)(negv ? -v : v);
    }
    return 0;
}

// ---- whole chunk files at once ------------------------------------------
// After the GPU offload the per-row Python work of the predict sub-command (split, parse, format) is what a chunk
// costs; these two calls take a whole decompressed tensor_can file and produce a whole predict file body.

static inline const char* find_char(const char* p, const char* end, char c) {
    const void* q = memchr(p, c, (size_t)(end - p));
    return q ? (const char*)q : end;
}

int cto_parse_tensor_file(const char* text, int64_t len, int64_t max_rows, int16_t* tensor, int32_t* depth, int64_t* fields,
                          int64_t* n_rows) {
    if (!text || !tensor || !depth || !fields || !n_rows) { cto::set_error("parse_tensor_file: NULL argument"); return 2; }
    const int n_val = CTO_N_POS * CTO_N_CH;
    const char* p = text;
    const char* const end = text + len;
    int64_t kept = 0, line_no = 0;
    // pass 1 (serial, memchr-bound): split rows and fields, apply the two row filters, parse the depth
    while (p < end) {
        const char* eol = find_char(p, end, '\n');
        ++line_no;
        // the first seven tab-separated fields (clairs/predict.py:172-175); shorter rows are skipped
        const char* f0[7];
        const char* f1[7];
        const char* q = p;
        int nf = 0;
        while (nf < 7 && q <= eol) {
            const char* t = find_char(q, eol, '\t');
            f0[nf] = q;
            f1[nf] = t;
            ++nf;
            if (t >= eol) break;
            q = t + 1;
        }
        const char* next = eol < end ? eol + 1 : end;
        if (nf < 7) { p = next; continue; }
        // rows whose centre reference base is not ACGT are dropped (clairs/predict.py:219-220)
        if (f1[2] - f0[2] <= 16) { cto::set_error("parse_tensor_file: line %lld: reference context shorter than 17 bases", (long long)line_no); return 2; }
        const char cb = f0[2][16];
        if (cb != 'A' && cb != 'C' && cb != 'G' && cb != 'T') { p = next; continue; }
        if (kept >= max_rows) { cto::set_error("parse_tensor_file: more than %lld rows", (long long)max_rows); return 2; }
        // depth = float(alt_info.split('-')[0]) (clairs/predict.py:179); written by "%d" in the encoder
        {
            const char* a = f0[4];
            const char* dash = find_char(a, f1[4], '-');
            char tmp[32];
            const size_t dl = (size_t)(dash - a) < sizeof(tmp) - 1 ? (size_t)(dash - a) : sizeof(tmp) - 1;
            memcpy(tmp, a, dl);
            tmp[dl] = 0;
            char* stop = nullptr;
            const double dv = strtod(tmp, &stop);
            if (stop == tmp) { cto::set_error("parse_tensor_file: line %lld: alt_info does not start with a depth", (long long)line_no); return 2; }
            depth[kept] = (int32_t)dv;
        }
        // the last kept field loses a trailing CR like str.strip() would (ref_center is stripped in predict.py)
        const char* e6 = f1[6];
        while (e6 > f0[6] && (e6[-1] == '\r' || e6[-1] == ' ')) --e6;
        f1[6] = e6;
        for (int k = 0; k < 7; ++k) {
            fields[(kept * 7 + k) * 2] = (int64_t)(f0[k] - text);
            fields[(kept * 7 + k) * 2 + 1] = (int64_t)(f1[k] - f0[k]);
        }
        ++kept;
        p = next;
    }
    // pass 2 (parallel over rows): the 1122 integers of every kept row
    // up to 8 host threads (CTO_HOST_THREADS overrides; the reference predict pins itself to one, clairs/predict.py:475)
    unsigned hw = std::thread::hardware_concurrency();
    int n_thr = (int)(hw ? (hw < 8 ? hw : 8) : 1);
    if (const char* ev = getenv("CTO_HOST_THREADS")) n_thr = atoi(ev) > 0 ? atoi(ev) : 1;
    if (n_thr > 64) n_thr = 64;
    if (kept < 256) n_thr = 1;
    std::vector<int> bad(n_thr, 0);
    auto work = [&](int tid) {
        const int64_t r0 = kept * tid / n_thr, r1 = kept * (tid + 1) / n_thr;
        for (int64_t r = r0; r < r1; ++r) {
            const int64_t* f = fields + (r * 7 + 3) * 2;
            if (cto_parse_tensor_row(text + f[0], f[1], tensor + r * n_val)) { bad[tid] = 1; return; }
        }
    };
    if (n_thr == 1) {
        work(0);
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_thr; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    for (int t = 0; t < n_thr; ++t)
        if (bad[t]) { cto::set_error("parse_tensor_file: a tensor field does not hold %d integers", n_val); return 2; }
    *n_rows = kept;
    return 0;
}

int64_t cto_format_tensor_can_rows(const char* ctg, int64_t ctg_len, int64_t n, const int64_t* pos, const char* ref33,
                                   const int16_t* tensor, const char* blob, const int64_t* alt_off, const int64_t* type_off,
                                   char* out, int64_t cap) {
    if (!ctg || !pos || !ref33 || !tensor || !blob || !alt_off || !type_off || !out) return -2;
    const int n_val = CTO_N_POS * CTO_N_CH;
    char* o = out;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t al = alt_off[2 * r + 1], tl = type_off[2 * r + 1];
        if (cap - (o - out) < ctg_len + 24 + CTO_N_POS + (int64_t)n_val * 7 + al + tl + 16) return -1;
        memcpy(o, ctg, (size_t)ctg_len); o += ctg_len; *o++ = '\t';
        {   // "%d" of the (positive) genomic position
            char tmp[24];
            int k = 0;
            long long v = (long long)pos[r];
            if (v < 0) { *o++ = '-'; v = -v; }
            do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
            while (k) *o++ = tmp[--k];
        }
        *o++ = '\t';
        memcpy(o, ref33 + r * CTO_N_POS, CTO_N_POS); o += CTO_N_POS; *o++ = '\t';
        o += cto_format_tensor_row(tensor + r * n_val, o, cap - (o - out)); *o++ = '\t';
        memcpy(o, blob + alt_off[2 * r], (size_t)al); o += al; *o++ = '\t';
        memcpy(o, blob + type_off[2 * r], (size_t)tl); o += tl; *o++ = '\t';
        *o++ = ref33[r * CTO_N_POS + CTO_N_POS / 2];
        *o++ = '\n';
    }
    return (int64_t)(o - out);
}

static inline char* put_count_list(char* o, const int32_t* v) {           // str([float(a), ...]) of integer-valued floats
    *o++ = '[';
    for (int k = 0; k < 4; ++k) {
        if (k) { *o++ = ','; *o++ = ' '; }
        o = put_int(o, v[k]);
        *o++ = '.';
        *o++ = '0';
    }
    *o++ = ']';
    return o;
}

int64_t cto_format_predict_rows(const char* text, const int64_t* fields, int64_t n, const int32_t* fwd, const int32_t* rev,
                                const float* probs, int n_heads, char* out, int64_t cap) {
    if (!text || !fields || !fwd || !rev || !probs || !out || (n_heads != 4 && n_heads != 6)) return -2;
    char* o = out;
    for (int64_t r = 0; r < n; ++r) {
        const int64_t* f = fields + r * 14;
        const int64_t need = f[1] + f[3] + f[9] + 2 * 64 + (int64_t)n_heads * 2 * 32 + 64;
        if (cap - (o - out) < need) return -1;
        memcpy(o, text + f[0], (size_t)f[1]); o += f[1]; *o++ = '\t';               // contig
        memcpy(o, text + f[2], (size_t)f[3]); o += f[3]; *o++ = '\t';               // position
        const char c = text[f[4] + 16];                                            // seq[16].upper()
        *o++ = (char)((c >= 'a' && c <= 'z') ? c - 32 : c); *o++ = '\t';
        memcpy(o, text + f[8], (size_t)f[9]); o += f[9]; *o++ = '\t';               // alt_info
        o = put_count_list(o, fwd + r * 4); *o++ = '\t';
        o = put_count_list(o, rev + r * 4); *o++ = '\t';
        const int64_t w = cto_format_prob_fields(probs + r * n_heads * 4, 2 * n_heads, o, cap - (o - out));
        if (w < 0) return -1;
        o += w;
        if (n_heads == 4) *o++ = '\t';                                             // the SNV row ends with an empty field
        *o++ = '\n';
    }
    return (int64_t)(o - out);
}

}  // extern "C"
