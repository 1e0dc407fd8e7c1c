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

extern "C" {

int cto_hf_parse(const char* text, int64_t len, int with_phasing, const char* ref, int64_t ref_len, int64_t region_lo, cto_hf_chunk** out) {
    CTO_REQUIRE(out, "hf_parse: NULL out");
    *out = nullptr;
    CTO_REQUIRE(len >= 0 && (len == 0 || text), "hf_parse: bad text");
    CTO_REQUIRE(ref_len >= 0 && (ref_len == 0 || ref), "hf_parse: bad reference");
    cto_hf_chunk* ck = new cto_hf_chunk();
    ck->sfxs.get(std::string());                                 // suffix id 0 = no suffix
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
        set_error("hf_parse: %s (row %lld)", what, (long long)line);
        delete ck;
        return 2;
    };
    int64_t i = 0, line_no = 0;
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
        CTO_REQUIRE(pos < (1ll << 31), "hf_parse: position %lld", (long long)pos);
        last_pos = pos;
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
        CTO_REQUIRE(ck->rid.size() < (1ull << 31), "hf_parse: more than 2^31 pileup entries in one chunk");
        i = next;
    }
    *out = ck;
    return 0;
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
