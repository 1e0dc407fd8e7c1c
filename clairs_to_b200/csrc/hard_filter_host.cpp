// Host tokenizer of the per-site hard filters: `samtools mpileup --output-MQ --output-QNAME
// [--output-extra HP]` text -> integer arrays the device kernel (hard_filter.cu) works on.
//
// Replaces get_base_list + _parse_mpileup_to_chunk_dict of the reference, src/haplotype_filtering.py:154-185, 246-275 (cited
// as HF) and their twins in src/postfilter_variants.py:144-175, 237-259 (PV).  Everything that is a STRING in the reference
// (read keys = QNAME + strand suffix, upper-cased tokens = symbol + indel suffix, raw suffixes) is interned into dense ids
// here, so that the set logic of the filters becomes integer work:
//   rid   read key id, in order of first appearance in the chunk
//   tok   id of (symbol + suffix).upper() -- the key of the reference's per-row Counter (HF:184)
//   sfx   id of the raw suffix ('' = 0): the germline insertion test (HF:447-448) is case sensitive on it
//   info  bit field below; qual = bq | mq << 8
// Quirks kept: `^` marks the read BEFORE it (HF:180-181; before the first read that is index -1 = the row's last read by
// Python indexing) and `$` the read it follows, the larger of the two index sets is the row's start/end set (HF:183); without
// the HP column the last QNAME of a row keeps its line feed (PV:243), which makes it a different key.
#include "../../include/clairs_to_b200.h"
#include "common.cuh"
#include <string.h>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace {

struct Interner {
    std::unordered_map<std::string, int32_t> ids;
    std::vector<int32_t> off{0};
    std::string blob;
    int32_t get(const std::string& s) {
        auto it = ids.find(s);
        if (it != ids.end()) return it->second;
        const int32_t id = (int32_t)ids.size();
        ids.emplace(s, id);
        blob += s;
        off.push_back((int32_t)blob.size());
        return id;
    }
};

inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; }
inline bool is_symbol(char c) {
    switch (c) {
        case 'A': case 'C': case 'G': case 'T': case 'N': case 'a': case 'c': case 'g': case 't': case 'n': case '#': case '*': return true;
        default: return false;
    }
}

}  // namespace

struct cto_hf_chunk {
    std::vector<int32_t> row_pos, row_off{0}, rse_off{0}, rse_ent, rid, tok, sfx;
    std::vector<uint8_t> row_flags;
    std::vector<uint32_t> info;
    std::vector<uint16_t> qual;
    Interner reads, toks, sfxs;
};

using cto::set_error;

namespace {

// Rows of text[begin, len) (whole rows) -> `ck`, with ids local to this range.  On failure: `err` = what, *err_line = the
// 1-based row within the range.  *first_pos / *last_pos: positions of the first and last row taken (-1: none).
int parse_range(const char* text, int64_t begin, int64_t len, int with_phasing, const char* ref, int64_t ref_len, int64_t region_lo,
                cto_hf_chunk* ck, std::string& err, int64_t* err_line, int64_t* first_pos, int64_t* last_pos_out) {
    ck->sfxs.get(std::string());                                 // suffix id 0 = no suffix
    *first_pos = *last_pos_out = -1;
    std::vector<int32_t> last_row_of_read, last_entry_of_read;
    // Consecutive pileup rows list mostly the same reads in the same order: the read keys of the previous row (text offset,
    // length, strand, id) are tried first, the hash map only for reads that are new or moved.  Single-character tokens
    // (no indel suffix) bypass the token map.
    struct Seen { int64_t off; int32_t len; int32_t rid; bool rev; };
    std::vector<Seen> prev_row, this_row;
    int32_t single_tok[256];
    for (int k = 0; k < 256; ++k) single_tok[k] = -1;
    struct Tok { char sym; int64_t s_off; int64_t s_len; char sign; };
    std::vector<Tok> row_toks;
    std::string key, token, suffix;
    auto fail = [&](const char* what, int64_t line) {
        err = what;
        *err_line = line;
        return 2;
    };
    int64_t i = begin, line_no = 0;
    int64_t last_pos = -1;
    while (i < len) {
        const char* nl = (const char*)memchr(text + i, '\n', (size_t)(len - i));
        const int64_t line_end = nl ? (nl - text) + 1 : len;     // the python rows keep their '\n'
        ++line_no;
        // columns (split('\t') of the whole line, line feed included in the last one)
        int64_t c_lo[10], c_hi[10];
        int n_col = 0;
        for (int64_t a = i; n_col < 10;) {
            const char* tab = a < line_end ? (const char*)memchr(text + a, '\t', (size_t)(line_end - a)) : nullptr;
            c_lo[n_col] = a;
            c_hi[n_col] = tab ? tab - text : line_end;
            ++n_col;
            if (!tab) break;
            a = (tab - text) + 1;
        }
        const int64_t next = line_end;
        if (n_col < (with_phasing ? 9 : 8)) { i = next; continue; }            // HF:250-251 / PV:241-242
        int64_t pos = 0;
        if (c_hi[1] == c_lo[1]) return fail("empty position column", line_no);
        for (int64_t k = c_lo[1]; k < c_hi[1]; ++k) {
            if (text[k] < '0' || text[k] > '9') return fail("position is not a number", line_no);
            pos = pos * 10 + (text[k] - '0');
        }
        if (pos <= last_pos) return fail("rows are not in strictly increasing position order", line_no);
        if (pos >= (1ll << 31)) return fail("position beyond 2^31", line_no);
        last_pos = pos;
        if (*first_pos < 0) *first_pos = pos;
        // the bases column: HF:154-185
        row_toks.clear();
        std::set<int32_t> starts, ends;
        const int64_t b_hi = c_hi[4];
        for (int64_t k = c_lo[4]; k < b_hi;) {
            const char c = text[k];
            if (c == '+' || c == '-') {
                ++k;
                int64_t adv = 0;
                for (;;) {
                    if (k >= b_hi) return fail("indel length runs into the end of the bases column", line_no);
                    if (text[k] < '0' || text[k] > '9') break;
                    adv = adv * 10 + (text[k] - '0');
                    ++k;
                }
                if (row_toks.empty()) return fail("indel suffix before the first read", line_no);
                Tok& t = row_toks.back();
                t.sign = c; t.s_off = k; t.s_len = adv < b_hi - k ? adv : b_hi - k;
                k += adv - 1;
            } else if (is_symbol(c)) {
                row_toks.push_back(Tok{c, 0, 0, 0});
            } else if (c == '^') {
                ++k;
                starts.insert((int32_t)row_toks.size() - 1);
            }
            if (c == '$') ends.insert((int32_t)row_toks.size() - 1);
            ++k;
        }
        const int32_t n = (int32_t)row_toks.size();
        if (c_hi[5] - c_lo[5] < n || c_hi[6] - c_lo[6] < n) return fail("fewer base / mapping qualities than reads", line_no);
        const int32_t r = (int32_t)ck->row_pos.size();
        const int32_t e0 = (int32_t)ck->rid.size();
        const int64_t rp = pos - region_lo;
        const bool ref_ok = rp >= 0 && rp < ref_len;
        const char centre = ref_ok ? ref[rp] : 0;
        // names, haplotype tags
        int64_t nm = c_lo[7], hp = with_phasing ? c_lo[8] : 0;
        int64_t hp_hi = 0;
        if (with_phasing) {
            hp_hi = c_hi[8];
            while (hp_hi > hp && text[hp_hi - 1] == '\n') --hp_hi;            // .strip('\n')
            while (hp < hp_hi && text[hp] == '\n') ++hp;
        }
        int32_t first_tok = -1;
        bool single = true;
        size_t hint = 0;
        this_row.clear();
        for (int32_t e = 0; e < n; ++e) {
            if (nm > c_hi[7]) return fail("fewer read names than reads", line_no);
            const Tok& t = row_toks[e];
            const bool rev = t.sym == '#' || (t.sym >= 'a' && t.sym <= 'z');
            int64_t q = nm;
            int32_t rid = -1;
            if (hint < prev_row.size()) {                         // the same read as in the previous row, at the expected place?
                const Seen& p = prev_row[hint];
                const int64_t end = nm + p.len;
                if (p.rev == rev && end <= c_hi[7] && (end == c_hi[7] || text[end] == ',') &&
                    memcmp(text + p.off, text + nm, (size_t)p.len) == 0 && !memchr(text + nm, ',', (size_t)p.len)) {
                    rid = p.rid;
                    q = end;
                    ++hint;
                }
            }
            if (rid < 0)
                while (q < c_hi[7] && text[q] != ',') ++q;
            const int64_t name_off = nm;
            const int32_t name_len = (int32_t)(q - nm);
            nm = q + 1;
            uint32_t hap = 0;
            if (with_phasing) {
                if (hp > hp_hi) return fail("fewer haplotype tags than reads", line_no);
                int64_t h = hp;
                while (h < hp_hi && text[h] != ',') ++h;
                if (h - hp == 1 && (text[hp] == '1' || text[hp] == '2')) hap = (uint32_t)(text[hp] - '0');
                else if (h == hp || (h - hp == 2 && text[hp] == '1' && text[hp + 1] == '2'))
                    return fail("empty or '12' haplotype tag (the reference's `hap in '12'` raises on it)", line_no);
                hp = h + 1;
            }
            for (size_t j = hint, tries = 0; rid < 0 && j < prev_row.size() && tries < 4; ++j, ++tries) {
                const Seen& p = prev_row[j];
                if (p.len == name_len && p.rev == rev && memcmp(text + p.off, text + name_off, (size_t)name_len) == 0) {
                    rid = p.rid;
                    hint = j + 1;
                    break;
                }
            }
            if (rid < 0) {
                key.assign(text + name_off, (size_t)name_len);
                key += rev ? "_1" : "_0";
                rid = ck->reads.get(key);
            }
            this_row.push_back(Seen{name_off, name_len, rid, rev});
            suffix.clear();
            int32_t tok, sfx = 0;
            const char sym_up = up(t.sym);
            if (!t.sign) {
                int32_t& cached = single_tok[(uint8_t)sym_up];
                if (cached < 0) { token.assign(1, sym_up); cached = ck->toks.get(token); }
                tok = cached;
                token.assign(1, sym_up);
            } else {
                suffix += t.sign;
                suffix.append(text + t.s_off, (size_t)t.s_len);
                token.assign(1, sym_up);
                for (char ch : suffix) token += up(ch);
                tok = ck->toks.get(token);
                sfx = ck->sfxs.get(suffix);
            }
            if ((size_t)rid >= last_row_of_read.size()) { last_row_of_read.resize(rid + 1, -1); last_entry_of_read.resize(rid + 1, -1); }
            uint32_t info = hap | (rev ? CTO_HF_REV : 0);
            if (token.size() == 1 && (token[0] == '#' || token[0] == '*')) info |= CTO_HF_STAR;
            if (ref_ok && token.size() == 1 && token[0] == centre) info |= CTO_HF_IS_REF;
            if (t.sign == '+') info |= CTO_HF_PLUS;
            if (suffix.find('-') != std::string::npos) info |= CTO_HF_MINUS;
            const size_t sl = suffix.size() > 65535 ? 65535 : suffix.size();
            info |= (uint32_t)sl << CTO_HF_LEN_SHIFT;
            if (last_row_of_read[rid] == r) ck->info[last_entry_of_read[rid]] |= CTO_HF_SHADOW;   // dict(zip(...)): the last token wins
            last_row_of_read[rid] = r;
            last_entry_of_read[rid] = e0 + e;
            if (first_tok < 0) first_tok = tok;
            else if (tok != first_tok) single = false;
            ck->rid.push_back(rid); ck->tok.push_back(tok); ck->sfx.push_back(sfx); ck->info.push_back(info);
            ck->qual.push_back((uint16_t)((uint8_t)(text[c_lo[5] + e] - 33) | ((uint16_t)(uint8_t)(text[c_lo[6] + e] - 33) << 8)));
        }
        if (n > 0) {                                              // every name must be used up: names == reads
            if (nm <= c_hi[7]) return fail("more read names than reads", line_no);
        }
        // row flags: HF:183, 627 (start/end set), HF:690-696 (the Counter kept for the variant-cluster test)
        const std::set<int32_t>& rse = starts.size() > ends.size() ? starts : ends;
        uint8_t flags = ref_ok ? CTO_HF_ROW_REF_OK : 0;
        if ((double)rse.size() >= (double)n * 0.2) flags |= CTO_HF_ROW_RSE;
        const bool only_ref = n > 0 && single && ref_ok && ck->toks.off[first_tok + 1] - ck->toks.off[first_tok] == 1 &&
                              ck->toks.blob[ck->toks.off[first_tok]] == centre;
        if (ref_ok && !only_ref) flags |= CTO_HF_ROW_COUNTER;
        for (int32_t idx : rse) {
            if (n == 0) return fail("read start / end marker in a row without reads", line_no);
            ck->rse_ent.push_back(e0 + (idx < 0 ? n - 1 : idx));
        }
        prev_row.swap(this_row);
        ck->rse_off.push_back((int32_t)ck->rse_ent.size());
        ck->row_pos.push_back((int32_t)pos);
        ck->row_off.push_back((int32_t)ck->rid.size());
        ck->row_flags.push_back(flags);
        if (ck->rid.size() >= (1ull << 31)) return fail("more than 2^31 pileup entries in one chunk", line_no);
        i = next;
    }
    *last_pos_out = last_pos;
    return 0;
}

// ids of `part`'s strings in `all` (interned in the part's own order, which keeps "order of first appearance" overall)
std::vector<int32_t> remap(Interner& all, const Interner& part) {
    std::vector<int32_t> m(part.ids.size());
    for (size_t k = 0; k < m.size(); ++k) m[k] = all.get(part.blob.substr((size_t)part.off[k], (size_t)(part.off[k + 1] - part.off[k])));
    return m;
}

}  // namespace

extern "C" {

// n_threads: 0 = as many as the machine has, at most 8, and one for texts under 8 MB.  The text is cut at row ends into
// n_threads ranges parsed side by side with local ids; the ranges are then appended in order and their ids translated, which
// gives the very arrays the single-threaded parse gives (tests/test_hard_filter_host.py).
int cto_hf_parse_mt(const char* text, int64_t len, int with_phasing, const char* ref, int64_t ref_len, int64_t region_lo, int n_threads,
                    cto_hf_chunk** out) {
    CTO_REQUIRE(out, "hf_parse: NULL out");
    *out = nullptr;
    CTO_REQUIRE(len >= 0 && (len == 0 || text), "hf_parse: bad text");
    CTO_REQUIRE(ref_len >= 0 && (ref_len == 0 || ref), "hf_parse: bad reference");
    int nt = n_threads;
    if (nt <= 0) {
        nt = (int)std::thread::hardware_concurrency();
        if (nt > 8) nt = 8;
        if (len < (8ll << 20)) nt = 1;
    }
    if (nt < 1) nt = 1;
    std::vector<int64_t> cut{0};                                 // range boundaries on row starts
    for (int k = 1; k < nt; ++k) {
        int64_t at = len * k / nt;
        if (at <= cut.back()) continue;
        const void* nl = memchr(text + at, '\n', (size_t)(len - at));
        at = nl ? (const char*)nl - text + 1 : len;
        if (at > cut.back() && at < len) cut.push_back(at);
    }
    cut.push_back(len);
    const int n_parts = (int)cut.size() - 1;
    std::vector<cto_hf_chunk*> parts(n_parts, nullptr);
    std::vector<std::string> errs(n_parts);
    std::vector<int64_t> err_line(n_parts, 0), first(n_parts, -1), last(n_parts, -1);
    std::vector<int> rcs(n_parts, 0);
    auto work = [&](int k) {
        parts[k] = new cto_hf_chunk();
        rcs[k] = parse_range(text, cut[k], cut[k + 1], with_phasing, ref, ref_len, region_lo, parts[k], errs[k], &err_line[k], &first[k], &last[k]);
    };
    if (n_parts == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < n_parts; ++k) th.emplace_back(work, k);
        for (auto& t : th) t.join();
    }
    auto cleanup = [&]() { for (auto* p : parts) delete p; };
    int64_t prev_last = -1;
    for (int k = 0; k < n_parts; ++k) {
        if (!rcs[k] && first[k] >= 0 && first[k] <= prev_last) {
            rcs[k] = 2; errs[k] = "rows are not in strictly increasing position order"; err_line[k] = 1;
        }
        if (rcs[k]) {
            int64_t before = 0;                                  // rows in front of this range
            for (int64_t b = 0; b < cut[k];) {
                const void* nl = memchr(text + b, '\n', (size_t)(cut[k] - b));
                if (!nl) break;
                ++before;
                b = (const char*)nl - text + 1;
            }
            set_error("hf_parse: %s (row %lld)", errs[k].c_str(), (long long)(before + err_line[k]));
            cleanup();
            return 2;
        }
        if (last[k] >= 0) prev_last = last[k];
    }
    if (n_parts == 1) {
        *out = parts[0];
        return 0;
    }
    cto_hf_chunk* ck = new cto_hf_chunk();
    ck->sfxs.get(std::string());
    size_t n_rows = 0, n_ent = 0, n_rse = 0;
    for (auto* p : parts) { n_rows += p->row_pos.size(); n_ent += p->rid.size(); n_rse += p->rse_ent.size(); }
    if (n_ent >= (1ull << 31)) {
        set_error("hf_parse: more than 2^31 pileup entries in one chunk");
        cleanup();
        delete ck;
        return 2;
    }
    ck->row_pos.resize(n_rows); ck->row_off.resize(n_rows + 1); ck->rse_off.resize(n_rows + 1); ck->row_flags.resize(n_rows);
    ck->rse_ent.resize(n_rse);
    ck->rid.resize(n_ent); ck->tok.resize(n_ent); ck->sfx.resize(n_ent); ck->info.resize(n_ent); ck->qual.resize(n_ent);
    ck->row_off[0] = 0; ck->rse_off[0] = 0;
    // the id translation tables are built in range order (that IS the order of first appearance); the arrays are then filled
    // by the ranges side by side
    std::vector<std::vector<int32_t>> mr(n_parts), mt(n_parts), ms(n_parts);
    std::vector<size_t> e_base(n_parts), r_base(n_parts), s_base(n_parts);
    size_t eb = 0, rb = 0, sb = 0;
    for (int k = 0; k < n_parts; ++k) {
        mr[k] = remap(ck->reads, parts[k]->reads); mt[k] = remap(ck->toks, parts[k]->toks); ms[k] = remap(ck->sfxs, parts[k]->sfxs);
        e_base[k] = eb; r_base[k] = rb; s_base[k] = sb;
        eb += parts[k]->rid.size(); rb += parts[k]->row_pos.size(); sb += parts[k]->rse_ent.size();
    }
    auto fill = [&](int k) {
        const cto_hf_chunk* p = parts[k];
        const size_t e0 = e_base[k], r0 = r_base[k], s0 = s_base[k];
        for (size_t e = 0; e < p->rid.size(); ++e) {
            ck->rid[e0 + e] = mr[k][p->rid[e]];
            ck->tok[e0 + e] = mt[k][p->tok[e]];
            ck->sfx[e0 + e] = ms[k][p->sfx[e]];
        }
        if (!p->info.empty()) {
            memcpy(ck->info.data() + e0, p->info.data(), p->info.size() * sizeof(uint32_t));
            memcpy(ck->qual.data() + e0, p->qual.data(), p->qual.size() * sizeof(uint16_t));
        }
        for (size_t j = 0; j < p->rse_ent.size(); ++j) ck->rse_ent[s0 + j] = p->rse_ent[j] + (int32_t)e0;
        for (size_t r = 0; r < p->row_pos.size(); ++r) {
            ck->row_pos[r0 + r] = p->row_pos[r];
            ck->row_flags[r0 + r] = p->row_flags[r];
            ck->row_off[r0 + r + 1] = p->row_off[r + 1] + (int32_t)e0;
            ck->rse_off[r0 + r + 1] = p->rse_off[r + 1] + (int32_t)s0;
        }
    };
    {
        std::vector<std::thread> th;
        for (int k = 0; k < n_parts; ++k) th.emplace_back(fill, k);
        for (auto& t : th) t.join();
    }
    cleanup();
    *out = ck;
    return 0;
}

int cto_hf_parse(const char* text, int64_t len, int with_phasing, const char* ref, int64_t ref_len, int64_t region_lo, cto_hf_chunk** out) {
    return cto_hf_parse_mt(text, len, with_phasing, ref, ref_len, region_lo, 0, out);
}

int cto_hf_sizes(const cto_hf_chunk* ck, int64_t* sizes) {
    CTO_REQUIRE(ck && sizes, "hf_sizes: NULL argument");
    sizes[0] = (int64_t)ck->row_pos.size();
    sizes[1] = (int64_t)ck->rid.size();
    sizes[2] = (int64_t)ck->rse_ent.size();
    sizes[3] = (int64_t)ck->reads.ids.size();
    sizes[4] = (int64_t)ck->toks.ids.size();
    sizes[5] = (int64_t)ck->toks.blob.size();
    sizes[6] = (int64_t)ck->sfxs.ids.size();
    sizes[7] = (int64_t)ck->sfxs.blob.size();
    return 0;
}

int cto_hf_export(const cto_hf_chunk* ck, int32_t* row_pos, int32_t* row_off, uint8_t* row_flags, int32_t* rse_off, int32_t* rse_ent,
                  int32_t* rid, int32_t* tok, int32_t* sfx, uint32_t* info, uint16_t* qual, int32_t* tok_off, char* tok_blob,
                  int32_t* sfx_off, char* sfx_blob) {
    CTO_REQUIRE(ck, "hf_export: NULL chunk");
#define CTO_HF_COPY(dst, src)                                                                 \
    if (dst && !(src).empty()) memcpy(dst, (src).data(), (src).size() * sizeof((src)[0]))
    CTO_HF_COPY(row_pos, ck->row_pos); CTO_HF_COPY(row_off, ck->row_off); CTO_HF_COPY(row_flags, ck->row_flags);
    CTO_HF_COPY(rse_off, ck->rse_off); CTO_HF_COPY(rse_ent, ck->rse_ent);
    CTO_HF_COPY(rid, ck->rid); CTO_HF_COPY(tok, ck->tok); CTO_HF_COPY(sfx, ck->sfx); CTO_HF_COPY(info, ck->info); CTO_HF_COPY(qual, ck->qual);
    CTO_HF_COPY(tok_off, ck->toks.off); CTO_HF_COPY(sfx_off, ck->sfxs.off);
    if (tok_blob && !ck->toks.blob.empty()) memcpy(tok_blob, ck->toks.blob.data(), ck->toks.blob.size());
    if (sfx_blob && !ck->sfxs.blob.empty()) memcpy(sfx_blob, ck->sfxs.blob.data(), ck->sfxs.blob.size());
#undef CTO_HF_COPY
    return 0;
}

void cto_hf_free(cto_hf_chunk* ck) { delete ck; }

}  // extern "C"
