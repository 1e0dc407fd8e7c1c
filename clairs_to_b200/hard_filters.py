"""Per-site hard filters on the GPU: host mirror of the per-chunk inner loop of the reference's
``src/haplotype_filtering.py`` (HF; long reads, phased) and ``src/postfilter_variants.py`` (PV; short reads).

The reference parses one ``samtools mpileup --output-MQ --output-QNAME [--output-extra HP]`` stream per chunk of sites
(``_run_mpileup_chunk_dict`` HF:298-341, ``_run_mpileup_postfilter_chunk_dict`` PV:262-275) and then calls, per site,
``_haplotype_build_state_and_line`` (HF:570-703) / ``_postfilter_build_state_and_line`` (PV:368-446), each returning one text
line of pass/fail fields.  Here:

* ``parse_chunk``  -- the chunk's mpileup text -> integer arrays (``cto_hf_parse``, host C++);
* ``haplotype_filter_chunk`` / ``postfilter_chunk`` -- every site of the chunk in ONE kernel launch (``cto_hard_filter_sites``),
  returning the same lines, in the order of ``sites``.

There is no CPU fallback: a missing library or GPU is an error.  Deviations from the reference, all on inputs samtools does
not produce: rows must come in increasing position order; a row must have as many read names / HP tags / qualities as
reads; an HP tag that is empty or ``12`` is rejected (the reference's ``hap in '12'`` accepts it and then raises or
mis-indexes); a read that carries different HP tags in different rows keeps an unspecified one of them.
"""

from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib

LOW_AF_SNV, LOW_AF_INDEL = 0.1, 0.3              # HF:22-23
SEQUENCE_ENTROPY_THRESHOLD = 0.9                 # HF:28, PV:21
ENTROPY_WINDOW, ENTROPY_CENTRE, ENTROPY_FLANK = 33, 100, 16   # shared/param.py:60-61; HF:29 (module constant `flanking`)
SMEM_READS = 16384                               # CTO_HF_SMEM_READS
IUPAC = dict(zip("ACGTURYSWKMBDHVN", (0, 1, 2, 3, 3, 0, 1, 1, 0, 2, 0, 1, 0, 0, 0, 0)))   # shared/utils.py:18-21

ROW_REF_OK, ROW_RSE, ROW_COUNTER = 1, 2, 4
(O_VERDICT, O_PHASEABLE, O_HETERO, O_HOMO, O_RSE, O_BQ, O_MQ, O_CO_EXIST, O_BOTH, O_SB, O_ENTROPY) = (1 << k for k in range(11))


class ChunkArrays(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_entries", C.c_int64)] + [(n, C.c_void_p) for n in (
        "row_pos", "row_off", "rse_off", "rse_ent", "rid", "tok", "row_flags", "info", "qual")]


class SiteArrays(C.Structure):
    _fields_ = [("n_sites", C.c_int64)] + [(n, C.c_void_p) for n in (
        "row_lo", "row_hi", "centre_row", "alt_tok", "del_len", "rid_min", "rid_span", "seq_off", "ph_off", "ph_row", "het_off",
        "het_idx", "hom_off", "hom_idx", "kind", "low_af", "seq_len", "seq", "scratch_off")] + [("n_germline", C.c_int64)] + [
        (n, C.c_void_p) for n in ("g_row", "g_off", "g_match")]


@dataclass
class PhasedChunk:
    """One chunk's pileup in integer form (layout: include/clairs_to_b200.h, cto_hf_export)."""
    with_phasing: bool
    chunk_ref: str
    region_lo: int
    row_pos: np.ndarray
    row_off: np.ndarray
    row_flags: np.ndarray
    rse_off: np.ndarray
    rse_ent: np.ndarray
    rid: np.ndarray
    tok: np.ndarray
    sfx: np.ndarray
    info: np.ndarray
    qual: np.ndarray
    tokens: list
    suffixes: list
    n_reads: int
    tok_ids: dict = field(default_factory=dict)
    _dev: dict = field(default_factory=dict)

    @property
    def n_rows(self):
        return len(self.row_pos)

    @property
    def n_entries(self):
        return len(self.rid)


def entropy_table():
    """HF:106-110: e * log(e) for e = i / 33 and the final multiplier, by the interpreter's libm like the reference."""
    tab = [0.0] * (ENTROPY_WINDOW + 2)
    for i in range(1, ENTROPY_WINDOW + 2):
        e = 1.0 / ENTROPY_WINDOW * i
        tab[i] = e * math.log(e)
    return (C.c_double * 35)(*tab), -1 / math.log(ENTROPY_WINDOW)


def parse_chunk(text: bytes, with_phasing: bool, chunk_ref: str, region_lo: int, n_threads: int = 0) -> PhasedChunk:
    """HF:246-275 / PV:237-259 over the whole chunk: ``text`` is what samtools wrote.  ``n_threads``: host threads of the
    tokenizer (0 = up to 8 for texts of 8 MB and more; pass 1 when chunks already run on a thread pool)."""
    lib = _lib.lib()
    chunk_ref = chunk_ref or ""
    refb = chunk_ref.encode()
    handle = C.c_void_p()
    _lib.check(lib.cto_hf_parse_mt(text, len(text), int(bool(with_phasing)), refb, len(refb), int(region_lo), int(n_threads), C.byref(handle)),
               "cto_hf_parse")
    try:
        sizes = (C.c_int64 * 8)()
        _lib.check(lib.cto_hf_sizes(handle, sizes), "cto_hf_sizes")
        rows, ents, n_rse, n_reads, n_tok, tok_bytes, n_sfx, sfx_bytes = (int(x) for x in sizes)
        a = dict(row_pos=np.empty(rows, np.int32), row_off=np.empty(rows + 1, np.int32), row_flags=np.empty(rows, np.uint8),
                 rse_off=np.empty(rows + 1, np.int32), rse_ent=np.empty(n_rse, np.int32), rid=np.empty(ents, np.int32),
                 tok=np.empty(ents, np.int32), sfx=np.empty(ents, np.int32), info=np.empty(ents, np.uint32), qual=np.empty(ents, np.uint16))
        tok_off, sfx_off = np.empty(n_tok + 1, np.int32), np.empty(n_sfx + 1, np.int32)
        tok_blob, sfx_blob = C.create_string_buffer(max(tok_bytes, 1)), C.create_string_buffer(max(sfx_bytes, 1))
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        _lib.check(lib.cto_hf_export(handle, p(a["row_pos"]), p(a["row_off"]), p(a["row_flags"]), p(a["rse_off"]), p(a["rse_ent"]),
                                     p(a["rid"]), p(a["tok"]), p(a["sfx"]), p(a["info"]), p(a["qual"]), p(tok_off), tok_blob,
                                     p(sfx_off), sfx_blob), "cto_hf_export")
    finally:
        lib.cto_hf_free(handle)
    tb, sb = tok_blob.raw[:tok_bytes].decode("latin-1"), sfx_blob.raw[:sfx_bytes].decode("latin-1")
    tokens = [tb[tok_off[k]:tok_off[k + 1]] for k in range(n_tok)]
    suffixes = [sb[sfx_off[k]:sfx_off[k + 1]] for k in range(n_sfx)]
    return PhasedChunk(bool(with_phasing), chunk_ref, int(region_lo), tokens=tokens, suffixes=suffixes, n_reads=n_reads,
                       tok_ids={t: k for k, t in enumerate(tokens)}, **a)


def _split_germline(info):
    """HF:586-589: 'pos-alt,pos-alt' -> set of (pos, alt)."""
    out = set()
    if info:
        for item in info.split(','):
            parts = tuple(item.split('-'))
            if len(parts) != 2:
                raise ValueError("germline entry %r is not '<pos>-<alt>'" % item)
            out.add((int(parts[0]), parts[1]))
    return out


def _site_tables(chunk: PhasedChunk, mode: int, sites, flanking: int):
    """Everything of a site that is a string in the reference, resolved to ids / row indices (host side)."""
    n = len(sites)
    i32 = lambda: np.zeros(n, np.int32)
    t = dict(row_lo=i32(), row_hi=i32(), centre_row=i32(), alt_tok=i32(), del_len=i32(), rid_min=i32(), rid_span=i32(), seq_off=i32(),
             kind=np.zeros(n, np.uint8), low_af=np.zeros(n, np.uint8), seq_len=np.zeros(n, np.uint8),
             scratch_off=np.full(n, -1, np.int64))
    ph_off, ph_row, het_off, het_idx, hom_off, hom_idx, seq = [0], [], [0], [], [0], [], []
    germline = {}                                               # (pos, alt) -> record index
    g_row, g_off, g_match = [], [0], []
    row_pos, row_off = chunk.row_pos, chunk.row_off
    nonempty = np.flatnonzero(row_off[1:] > row_off[:-1])
    row_min = np.full(chunk.n_rows, np.iinfo(np.int32).max, np.int64)
    row_max = np.full(chunk.n_rows, -1, np.int64)
    if len(nonempty):
        row_min[nonempty] = np.minimum.reduceat(chunk.rid, row_off[nonempty])
        row_max[nonempty] = np.maximum.reduceat(chunk.rid, row_off[nonempty])
    ins_by_text = {}                                            # token text without '+' -> ids of the insertion tokens spelling it
    for k, tk in enumerate(chunk.tokens):
        if '+' in tk:
            ins_by_text.setdefault(tk.replace('+', ''), []).append(k)
    bodies = [x[1:] if len(x) > 1 else None for x in chunk.suffixes]     # `value[1][1:]` of HF:447, only for len(value[1]) > 1
    key_hits = {}                                               # substring -> uint8 [suffixes]: the substring occurs in the body

    def suffix_hits(key):
        if key not in key_hits:
            key_hits[key] = np.fromiter((b is not None and key in b for b in bodies), np.uint8, len(bodies))
        return key_hits[key]
    scratch_words = 0

    def record(gp, ab):
        key = (gp, ab)
        if key in germline:
            return germline[key]
        r = int(np.searchsorted(row_pos, gp))
        present = r < chunk.n_rows and row_pos[r] == gp
        g_row.append(r if present else -1)
        if present:
            lo, hi = row_off[r], row_off[r + 1]
            if len(ab) == 1:                                    # HF:444-445 / 473-474: symbol + raw suffix == alt
                m = (chunk.tok[lo:hi] == chunk.tok_ids.get(ab, -1)).astype(np.uint8) * 3
            elif len(ab) > 1:                                   # HF:446-448 (ab[:2]) / 475-477 (ab[1:2]): substring of the raw suffix
                m = (suffix_hits(ab[:2]) | (suffix_hits(ab[1:2]) << 1))[chunk.sfx[lo:hi]]
            else:
                m = np.zeros(hi - lo, np.uint8)
            g_match.append(m)
            g_off.append(g_off[-1] + int(hi - lo))
        else:
            g_off.append(g_off[-1])
        germline[key] = len(g_row) - 1
        return germline[key]

    site_pos = np.array([int(site[0]) for site in sites], np.int64)
    anchors = np.maximum(site_pos - flanking, 1)
    lo_all = np.searchsorted(row_pos, anchors, 'left')
    hi_all = np.searchsorted(row_pos, site_pos + flanking, 'right')
    c_all = np.searchsorted(row_pos, site_pos)
    for s, site in enumerate(sites):
        pos, ref_base, alt_base = int(site[0]), site[1], site[2]
        af = site[3] if len(site) > 3 and site[3] is not None else 1.0
        het = _split_germline(site[4]) if mode == 1 and len(site) > 4 else set()
        hom = _split_germline(site[5]) if mode == 1 and len(site) > 5 else set()
        anchor = max(pos - flanking, 1)
        if anchor < chunk.region_lo:
            raise ValueError("site %d: its window starts before the chunk's reference (%d)" % (pos, chunk.region_lo))
        lo, hi, c = int(lo_all[s]), int(hi_all[s]), int(c_all[s])
        t["row_lo"][s], t["row_hi"][s] = lo, hi
        t["centre_row"][s] = c if c < chunk.n_rows and row_pos[c] == pos else -1
        is_snp = len(ref_base) == 1 and len(alt_base) == 1
        kind = 0 if is_snp else 1 if len(ref_base) == 1 and len(alt_base) > 1 else 2 if len(ref_base) > 1 and len(alt_base) == 1 else 3
        t["kind"][s] = kind
        if kind == 0:                                           # HF:639-642: symbol + suffix == alt
            t["alt_tok"][s] = chunk.tok_ids.get(alt_base, -1)
        elif kind == 1:                                         # HF:643-647: '+' in it and the upper-cased text without '+' == alt
            hits = ins_by_text.get(alt_base, [])
            if len(hits) > 1:
                raise ValueError("site %d: several pileup tokens spell the insertion %r" % (pos, alt_base))
            t["alt_tok"][s] = hits[0] if hits else -1
        else:
            t["alt_tok"][s] = -1
        t["del_len"][s] = len(ref_base)
        t["low_af"][s] = af < (LOW_AF_SNV if is_snp else LOW_AF_INDEL)
        if mode == 0 and hi > lo and not (chunk.row_flags[lo:hi] & ROW_REF_OK).all():
            raise IndexError("site %d: pileup rows beyond the chunk's reference (PV:404)" % pos)
        if hi > lo and row_max[lo:hi].max() >= 0:
            t["rid_min"][s] = row_min[lo:hi].min()
            t["rid_span"][s] = row_max[lo:hi].max() - t["rid_min"][s] + 1
        if t["rid_span"][s] > SMEM_READS:
            t["scratch_off"][s] = scratch_words
            scratch_words += (int(t["rid_span"][s]) + 3) // 4
        site_ref = chunk.chunk_ref[anchor - chunk.region_lo: pos + flanking + 2 - chunk.region_lo]
        if kind != 0:                                           # HF:147-148 with the module constant 100
            sq = site_ref[ENTROPY_CENTRE - ENTROPY_FLANK: ENTROPY_CENTRE + ENTROPY_FLANK + 1]
            t["seq_off"][s], t["seq_len"][s] = len(seq), len(sq)
            seq.extend(IUPAC[ch] for ch in sq)
        if mode == 1:
            want = sorted(set(gp for gp, _ in het) | {pos})
            idx = np.searchsorted(row_pos, want)
            ph_row.extend(int(r) for r, gp in zip(idx, want) if lo <= r < hi and row_pos[r] == gp)
            het_idx.extend(record(gp, ab) for gp, ab in sorted(het))
            hom_idx.extend(record(gp, ab) for gp, ab in sorted(hom))
        ph_off.append(len(ph_row)); het_off.append(len(het_idx)); hom_off.append(len(hom_idx))
    t.update(ph_off=np.array(ph_off, np.int32), ph_row=np.array(ph_row, np.int32), het_off=np.array(het_off, np.int32),
             het_idx=np.array(het_idx, np.int32), hom_off=np.array(hom_off, np.int32), hom_idx=np.array(hom_idx, np.int32),
             seq=np.array(seq, np.uint8), g_row=np.array(g_row, np.int32), g_off=np.array(g_off, np.int32),
             g_match=np.concatenate(g_match) if g_match else np.zeros(0, np.uint8))
    return t, scratch_words


CHUNK_FIELDS = ("row_pos", "row_off", "rse_off", "rse_ent", "rid", "tok", "row_flags", "info", "qual")
SITE_FIELDS = ("row_lo", "row_hi", "centre_row", "alt_tok", "del_len", "rid_min", "rid_span", "seq_off", "ph_off", "ph_row", "het_off",
               "het_idx", "hom_off", "hom_idx", "kind", "low_af", "seq_len", "seq", "scratch_off", "g_row", "g_off", "g_match")


def _to_device(arr, device):
    import torch
    a = np.ascontiguousarray(arr)
    view = {np.dtype(np.uint32): np.int32, np.dtype(np.uint16): np.int16}.get(a.dtype)
    tt = torch.from_numpy(a.view(view) if view else a)
    if tt.numel() == 0:
        return torch.zeros(1, dtype=tt.dtype, device=device)    # a valid pointer for an empty array
    return tt.to(device, non_blocking=False)


def chunk_to_device(chunk: PhasedChunk, device="cuda"):
    if str(device) not in chunk._dev:
        chunk._dev[str(device)] = {k: _to_device(getattr(chunk, k), device) for k in CHUNK_FIELDS}
    return chunk._dev[str(device)]


def run_sites(chunk: PhasedChunk, mode: int, sites, flanking=100, disable_read_start_end_filtering=False, max_co_exist_read_num=3,
              device="cuda", return_counts=False):
    """Flags (CTO_HFO_* bits), Fisher p-values (and the debug counters) of every site: one kernel launch."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("clairs_to_b200.hard_filters needs a CUDA device (there is no CPU path)")
    lib = _lib.lib()
    n = len(sites)
    if n == 0:
        out = np.zeros(0, np.uint32), np.zeros(0, np.float64), np.zeros((0, 8), np.int32)
        return out if return_counts else out[:2]
    tables, scratch_words = _site_tables(chunk, mode, sites, flanking)
    dev = chunk_to_device(chunk, device)
    sd = {k: _to_device(tables[k], device) for k in SITE_FIELDS}
    ca = ChunkArrays(chunk.n_rows, chunk.n_entries, *[C.c_void_p(dev[k].data_ptr()) for k in CHUNK_FIELDS])
    sa = SiteArrays(n, *[C.c_void_p(sd[k].data_ptr()) for k in SITE_FIELDS[:19]], len(tables["g_row"]),
                    *[C.c_void_p(sd[k].data_ptr()) for k in SITE_FIELDS[19:]])
    scratch = torch.empty(max(scratch_words, 1), dtype=torch.int32, device=device)
    flags = torch.empty(n, dtype=torch.int32, device=device)
    pval = torch.empty(n, dtype=torch.float64, device=device)
    counts = torch.empty((n, 8), dtype=torch.int32, device=device)
    tab, mul = entropy_table()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.cto_hard_filter_sites(C.byref(ca), C.byref(sa), int(mode), int(flanking), int(bool(disable_read_start_end_filtering)),
                                         int(max_co_exist_read_num), tab, mul, SEQUENCE_ENTROPY_THRESHOLD,
                                         C.c_void_p(scratch.data_ptr()), C.c_void_p(flags.data_ptr()), C.c_void_p(pval.data_ptr()),
                                         C.c_void_p(counts.data_ptr()), stream), "cto_hard_filter_sites")
    out = flags.cpu().numpy().view(np.uint32), pval.cpu().numpy(), counts.cpu().numpy()
    return out if return_counts else out[:2]


def format_lines(mode, ctg_name, sites, flags, pval):
    """HF:560-565 / PV:362-365: the result line of every site."""
    lines = []
    for site, f, p in zip(sites, flags, pval):
        b = lambda bit: str(bool(f & bit))
        if mode == 1:
            fields = [b(O_VERDICT), b(O_PHASEABLE), b(O_HETERO), b(O_HOMO), b(O_RSE), b(O_BQ), b(O_MQ), b(O_CO_EXIST), b(O_BOTH), b(O_SB),
                      str(round(float(p), 5)), b(O_ENTROPY)]
        else:
            fields = [b(O_VERDICT), b(O_RSE), b(O_CO_EXIST), b(O_SB), str(round(float(p), 5)), b(O_ENTROPY)]
        lines.append(' '.join([ctg_name, str(int(site[0]))] + fields))
    return lines


def haplotype_filter_chunk(ctg_name, sites, mpileup_text, chunk_ref, region_lo, flanking=100, disable_read_start_end_filtering=False,
                           max_co_exist_read_num=3, device="cuda", n_threads=0):
    """The site loop of ``_run_haplotype_chunk`` (HF:1078-1125): ``sites`` = [(pos, ref_base, alt_base, af, hetero_info,
    homo_info)] as in HAP_INFO (HF:1023-1030); returns the lines ``_haplotype_build_state_and_line`` would."""
    chunk = parse_chunk(mpileup_text, True, chunk_ref, region_lo, n_threads)
    flags, pval = run_sites(chunk, 1, sites, flanking, disable_read_start_end_filtering, max_co_exist_read_num, device)
    return format_lines(1, ctg_name, sites, flags, pval)


def postfilter_chunk(ctg_name, sites, mpileup_text, chunk_ref, region_lo, flanking=100, disable_read_start_end_filtering=False,
                     max_co_exist_read_num=3, device="cuda", n_threads=0):
    """The site loop of the post-filter chunk mode: ``sites`` = [(pos, ref_base, alt_base)]; returns the lines
    ``_postfilter_build_state_and_line`` (PV:368-446) would."""
    chunk = parse_chunk(mpileup_text, False, chunk_ref, region_lo, n_threads)
    flags, pval = run_sites(chunk, 0, sites, flanking, disable_read_start_end_filtering, max_co_exist_read_num, device)
    return format_lines(0, ctg_name, sites, flags, pval)
