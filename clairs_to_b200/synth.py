"""Seeded synthetic candidate sites in the encoder's input layout.

There is no network for BAMs, so every workload is synthetic: per candidate a window of
33 pileup rows, each row a Poisson-depth stack of reads with the error / indel / quality
model of SURVEY.md §8d.  The NEG stream keeps every read (``--min-BQ 0``,
run_clairs_to:1264); the AFF stream drops reads below the platform's ``--min-BQ``
(run_clairs_to:1237, shared/param.py:34) exactly as ``samtools mpileup`` would.

``render_mpileup`` turns a (small) stream back into mpileup text rows so that the same
sites can be fed to the reference-shaped tokenizer and to the oracle.
"""

from __future__ import annotations

import numpy as np

from .pileup_format import (HAS_INDEL, IND_DEL, IND_LONG, IND_REV, N_POS, SYMBOLS, PileupStream)

PLATFORM_MODEL = {
    # name: (depth mean, mismatch, star, ins, del, N, mq60, bq sampler id, aff min_bq)
    'ont': dict(depth=50, mismatch=0.015, star=0.01, ins=0.007, dele=0.007, n=0.001, mq60=0.9, bq='normal', min_bq=20),
    'ilmn': dict(depth=50, mismatch=0.002, star=0.001, ins=0.00025, dele=0.00025, n=0.0005, mq60=0.92, bq='ilmn', min_bq=0),
    'hifi': dict(depth=50, mismatch=0.002, star=0.002, ins=0.0015, dele=0.0015, n=0.0002, mq60=0.95, bq='hifi', min_bq=0),
}


def _sample_bq(rng, n, kind):
    if kind == 'normal':
        return np.clip(np.rint(rng.normal(25.0, 8.0, n)), 1, 50).astype(np.uint8)
    if kind == 'ilmn':
        return rng.choice(np.array([37, 25, 11, 2], dtype=np.uint8), size=n, p=[0.85, 0.08, 0.05, 0.02])
    return np.clip(np.rint(rng.normal(40.0, 6.0, n)), 2, 93).astype(np.uint8)


def synth_stream(n_candidates, seed, platform='ont', depth_lo=8, depth_hi=120, depth_mean=None,
                 max_indel_len=8):
    """Generate the NEG-style (unfiltered) stream for ``n_candidates`` disjoint windows."""
    m = PLATFORM_MODEL[platform]
    rng = np.random.default_rng(seed)
    n_rows = n_candidates * N_POS
    mean = m['depth'] if depth_mean is None else depth_mean
    depth = np.clip(rng.poisson(mean, n_rows), depth_lo, depth_hi).astype(np.int32)
    pos_off = np.zeros(n_rows + 1, dtype=np.int32)
    np.cumsum(depth, out=pos_off[1:])
    n_reads = int(pos_off[-1])
    ref_code = rng.integers(0, 4, n_rows, dtype=np.uint8)
    row_of = np.repeat(np.arange(n_rows, dtype=np.int32), depth)
    ref_r = ref_code[row_of]

    u = rng.random(n_reads, dtype=np.float32)
    shift = rng.integers(1, 4, n_reads, dtype=np.uint8)
    base = np.where(u < m['mismatch'], (ref_r + shift) & 3, ref_r).astype(np.uint8)

    # centre-row somatic-like allele in half of the sites, AF ~ U(.05, .5)
    has_alt = rng.random(n_candidates) < 0.5
    af = rng.uniform(0.05, 0.5, n_candidates).astype(np.float32)
    alt_shift = rng.integers(1, 4, n_candidates, dtype=np.uint8)
    centre_rows = np.arange(n_candidates, dtype=np.int64) * N_POS + N_POS // 2
    is_centre = (row_of % N_POS) == (N_POS // 2)
    cand_of = row_of[is_centre] // N_POS
    flip = has_alt[cand_of] & (rng.random(cand_of.size, dtype=np.float32) < af[cand_of])
    cb = base[is_centre]
    cb = np.where(flip, (ref_code[centre_rows][cand_of] + alt_shift[cand_of]) & 3, cb)
    base[is_centre] = cb

    reverse = rng.integers(0, 2, n_reads, dtype=np.uint8)
    sym = base + 5 * reverse                                   # ACGT -> 0..3, acgt -> 5..8
    v = rng.random(n_reads, dtype=np.float32)
    is_n = v < m['n']
    sym[is_n] = 4 + 5 * reverse[is_n]
    is_star = (v >= m['n']) & (v < m['n'] + m['star'])
    sym[is_star] = 10 + reverse[is_star]                       # '*' forward, '#' reverse (--reverse-del)

    w = rng.random(n_reads, dtype=np.float32)
    is_ins = w < m['ins']
    is_del = (w >= m['ins']) & (w < m['ins'] + m['dele'])
    indel_idx = np.flatnonzero(is_ins | is_del)
    k = indel_idx.size
    ind_len = np.minimum(rng.geometric(0.5, k), max_indel_len).astype(np.int64)
    ind_seq = rng.integers(0, 4 ** max_indel_len, k, dtype=np.int64) % (4 ** ind_len)   # base-4 packed
    ind_is_del = is_del[indel_idx]
    ind_seq[ind_is_del] = 0                                    # no -f: deleted bases print as N
    code = sym.astype(np.uint8)
    code[indel_idx] |= HAS_INDEL

    mq = np.where(rng.random(n_reads, dtype=np.float32) < m['mq60'], 60,
                  rng.integers(0, 60, n_reads, dtype=np.uint8)).astype(np.uint8)
    bq = _sample_bq(rng, n_reads, m['bq'])

    # allele ids: distinct (symbol, sign, length, sequence) keys inside a row
    ind_row = row_of[indel_idx].astype(np.int64)
    key = (((sym[indel_idx].astype(np.int64) * 2 + ind_is_del) * 16 + ind_len) << 20) | ind_seq
    order = np.lexsort((key, ind_row))
    srow, skey = ind_row[order], key[order]
    new_key = np.ones(k, dtype=bool)
    new_key[1:] = (srow[1:] != srow[:-1]) | (skey[1:] != skey[:-1])
    new_row = np.ones(k, dtype=bool)
    new_row[1:] = srow[1:] != srow[:-1]
    run = np.cumsum(new_key) - 1
    row_first = np.maximum.accumulate(np.where(new_row, run, 0))
    allele = np.empty(k, dtype=np.int64)
    allele[order] = run - row_first
    sym_i = sym[indel_idx]
    is_rev = ~np.isin(sym_i, np.array([0, 1, 2, 3, 4, 10]))
    entry = (allele.astype(np.uint32) & 0xFFFF) | (mq[indel_idx].astype(np.uint32) << 16)
    entry |= np.where(ind_is_del, IND_DEL, 0).astype(np.uint32)
    entry |= np.where(is_rev, IND_REV, 0).astype(np.uint32)
    ind_off = np.zeros(n_rows + 1, dtype=np.int32)
    np.cumsum(np.bincount(ind_row, minlength=n_rows), out=ind_off[1:])
    win_pos = np.arange(n_rows, dtype=np.int32)
    stream = PileupStream(code, bq, mq, pos_off, ref_code, ind_off, entry.astype(np.uint32), win_pos)
    aux = dict(ind_len=ind_len, ind_seq=ind_seq, indel_idx=indel_idx)
    return stream, aux


def filter_min_bq(stream: PileupStream, min_bq: int, aux=None):
    """What ``samtools mpileup --min-BQ`` does to a row: reads below the cut vanish (with
    their indel suffix)."""
    if min_bq <= 0:
        return stream, aux
    keep = stream.bq >= min_bq
    n_rows = stream.n_rows
    csum = np.zeros(stream.n_reads + 1, dtype=np.int64)
    np.cumsum(keep, out=csum[1:])
    pos_off = csum[stream.pos_off].astype(np.int32)
    has_ind = (stream.code & HAS_INDEL) != 0
    ind_keep = keep[has_ind]
    isum = np.zeros(ind_keep.size + 1, dtype=np.int64)
    np.cumsum(ind_keep, out=isum[1:])
    ind_off = isum[stream.ind_off].astype(np.int32)
    out = PileupStream(stream.code[keep], stream.bq[keep], stream.mq[keep], pos_off, stream.ref_code,
                       ind_off, stream.ind_entry[ind_keep], stream.win_pos)
    new_aux = None
    if aux is not None:
        new_aux = dict(ind_len=aux['ind_len'][ind_keep], ind_seq=aux['ind_seq'][ind_keep],
                       indel_idx=(csum[aux['indel_idx']][ind_keep]))
    assert out.n_rows == n_rows
    return out, new_aux


def synth_pair(n_candidates, seed, platform='ont', **kw):
    """(AFF stream, NEG stream) for one batch of candidates; Illumina/HiFi share one stream
    (run_clairs_to:1248-1252 symlinks NEG to AFF when both use --min-BQ 0)."""
    neg, aux = synth_stream(n_candidates, seed, platform, **kw)
    min_bq = PLATFORM_MODEL[platform]['min_bq']
    aff, aff_aux = filter_min_bq(neg, min_bq, aux)
    return (aff, aff_aux), (neg, aux)


def concat_streams(streams):
    """Concatenate disjoint PileupStreams (offsets and window indices are rebased)."""
    code = np.concatenate([s.code for s in streams])
    bq = np.concatenate([s.bq for s in streams])
    mq = np.concatenate([s.mq for s in streams])
    ref_code = np.concatenate([s.ref_code for s in streams])
    ind_entry = np.concatenate([s.ind_entry for s in streams])
    pos_off, ind_off, win_pos = [np.zeros(1, np.int32)], [np.zeros(1, np.int32)], []
    reads = inds = rows = 0
    for s in streams:
        pos_off.append((s.pos_off[1:].astype(np.int64) + reads).astype(np.int32))
        ind_off.append((s.ind_off[1:].astype(np.int64) + inds).astype(np.int32))
        win_pos.append(np.where(s.win_pos >= 0, s.win_pos + rows, -1).astype(np.int32))
        reads += s.n_reads
        inds += len(s.ind_entry)
        rows += s.n_rows
    assert reads < 2 ** 31, "one batch holds at most 2^31 reads"
    return PileupStream(code, bq, mq, np.concatenate(pos_off), ref_code, np.concatenate(ind_off), ind_entry,
                        np.concatenate(win_pos))


def synth_pair_large(n_candidates, seed, platform='ont', piece=20000, **kw):
    """synth_pair in bounded-memory pieces (bench-scale batches)."""
    affs, negs = [], []
    for k, lo in enumerate(range(0, n_candidates, piece)):
        (a, _), (n, _) = synth_pair(min(piece, n_candidates - lo), seed * 1000 + k, platform, **kw)
        affs.append(a)
        negs.append(n)
    return concat_streams(affs), concat_streams(negs)


_BASES = "ACGT"


def _indel_text(sym_char, is_del, length, seq_packed, reverse):
    if is_del:
        body = ('n' if reverse else 'N') * length
        return "-%d%s" % (length, body)
    letters = []
    for _ in range(length):
        letters.append(_BASES[seq_packed & 3])
        seq_packed >>= 2
    body = ''.join(letters)
    return "+%d%s" % (length, body.lower() if reverse else body)


def render_mpileup(stream: PileupStream, aux, ctg='chr1', first_pos=1001, decorate_seed=None):
    """Stream -> list of mpileup text rows ``ctg pos N depth bases BQ MQ`` (one per pileup row,
    positions ``first_pos + row``).  With ``decorate_seed`` some reads get ``^<mq>`` / ``$``
    decorations, which the tokenizer must skip (create_tensor_pileup_calling.py:142-144)."""
    rng = np.random.default_rng(decorate_seed) if decorate_seed is not None else None
    rows = []
    k = 0
    for r in range(stream.n_rows):
        lo, hi = int(stream.pos_off[r]), int(stream.pos_off[r + 1])
        parts = []
        for i in range(lo, hi):
            c = int(stream.code[i])
            ch = SYMBOLS[c & 0xF]
            tok = ch
            if rng is not None and rng.random() < 0.03:
                tok = '^' + chr(33 + int(rng.integers(0, 60))) + tok
            if c & HAS_INDEL:
                e = int(stream.ind_entry[k])
                tok += _indel_text(ch, bool(e & IND_DEL), int(aux['ind_len'][k]), int(aux['ind_seq'][k]),
                                   bool(e & IND_REV))
                k += 1
            if rng is not None and rng.random() < 0.03:
                tok += '$'
            parts.append(tok)
        bqs = ''.join(chr(33 + int(q)) for q in stream.bq[lo:hi])
        mqs = ''.join(chr(33 + int(q)) for q in stream.mq[lo:hi])
        bases = ''.join(parts)
        if not parts:                      # samtools prints "0 * * *" for an empty column
            bases = bqs = mqs = '*'
        rows.append("%s\t%d\tN\t%d\t%s\t%s\t%s\n" % (ctg, first_pos + r, hi - lo, bases, bqs, mqs))
    return rows


def render_mpileup_text(stream: PileupStream, aux, ctg='chr1', first_pos=1001) -> bytes:
    """The rows of ``render_mpileup`` (without decorations) as one bytes object, by the native renderer
    (``cto_render_mpileup``): bench-scale streams cannot go through a Python loop per read."""
    import ctypes as C
    from . import _lib
    lib = _lib.lib()
    n_ind = len(stream.ind_entry)
    ind_len = np.ascontiguousarray(aux['ind_len'], dtype=np.int64) if n_ind else np.zeros(1, np.int64)
    ind_seq = np.ascontiguousarray(aux['ind_seq'], dtype=np.int64) if n_ind else np.zeros(1, np.int64)
    cap = int(stream.n_reads) * 3 + int(ind_len.sum()) + 8 * n_ind + stream.n_rows * (len(ctg) + 48) + 1024
    buf = np.empty(cap, dtype=np.uint8)
    arrs = [np.ascontiguousarray(a) for a in (stream.code, stream.bq, stream.mq, stream.pos_off, stream.ind_entry)]
    if arrs[4].size == 0:
        arrs[4] = np.zeros(1, np.uint32)
    p = lambda a: C.c_void_p(a.ctypes.data)
    w = lib.cto_render_mpileup(p(arrs[0]), p(arrs[1]), p(arrs[2]), p(arrs[3]), stream.n_rows, p(arrs[4]), p(ind_len), p(ind_seq),
                               ctg.encode(), int(first_pos), p(buf), cap)
    if w < 0:
        raise RuntimeError("cto_render_mpileup: buffer too small")
    return buf[:w].tobytes()


def take_candidates(stream: PileupStream, n: int) -> PileupStream:
    """The first ``n`` candidates of a stream whose windows are disjoint and in row order (what synth_stream makes)."""
    rows = n * N_POS
    reads, inds = int(stream.pos_off[rows]), int(stream.ind_off[rows])
    return PileupStream(stream.code[:reads], stream.bq[:reads], stream.mq[:reads], stream.pos_off[:rows + 1],
                        stream.ref_code[:rows], stream.ind_off[:rows + 1], stream.ind_entry[:inds], stream.win_pos[:rows])


def synth_pair_tiled(n_candidates, seed, platform='ont', base=50000, **kw):
    """Bench-scale batches: ``base`` distinct candidates generated once and tiled to ``n_candidates`` (separate copies in
    memory, so the bytes moved are real).  Returns ((aff, aff_aux), (neg, neg_aux)) like synth_pair when
    n_candidates <= base, else ((aff, None), (neg, None))."""
    if n_candidates <= base:
        if n_candidates <= 20000:
            return synth_pair(n_candidates, seed, platform, **kw)
        affs, negs, a_aux, n_aux = [], [], [], []                      # bounded-memory pieces, aux arrays concatenated
        for k, lo in enumerate(range(0, n_candidates, 20000)):
            (a, aa), (g, ga) = synth_pair(min(20000, n_candidates - lo), seed * 1000 + k, platform, **kw)
            affs.append(a); negs.append(g); a_aux.append(aa); n_aux.append(ga)
        cat = lambda auxs: dict(ind_len=np.concatenate([x['ind_len'] for x in auxs]), ind_seq=np.concatenate([x['ind_seq'] for x in auxs]))
        return (concat_streams(affs), cat(a_aux)), (concat_streams(negs), cat(n_aux))
    (a, _), (g, _) = synth_pair(base, seed, platform, **kw)
    reps = -(-n_candidates // base)
    aff = take_candidates(concat_streams([a] * reps), n_candidates)
    neg = take_candidates(concat_streams([g] * reps), n_candidates)
    return (aff, None), (neg, None)


def scan_rows_text(n_rows, seed, ctg='chr20', first_pos=1001, depth_mean=40, weird=0.0):
    """Whole-chunk mpileup rows for the candidate scan (STEP 1, ``samtools mpileup`` WITHOUT ``--output-MQ``: six columns
    ``ctg pos N depth bases BQ``), with planted SNVs / insertions / deletions at a range of allele fractions around the
    thresholds, strand case, ``*`` / ``#`` / N reads, ``^<q>`` read starts whose quality character is a structural
    character (``+ - ^ $`` digits, letters), ``$`` read ends and ``< >`` reference skips.  ``weird`` > 0 adds the
    tokenizer's corner cases with that probability per row (``+0``, a second suffix on one read, a suffix on ``*``/N).
    Returns (rows, reference) where reference[i] is the base of position first_pos + i (some are N / lower case).
    Small inputs only (a Python loop per read)."""
    rng = np.random.default_rng(seed)
    ref = rng.choice(list("ACGT"), size=n_rows)
    ref[rng.random(n_rows) < 0.01] = 'N'
    lower = rng.random(n_rows) < 0.05
    reference = ''.join(b.lower() if l else b for b, l in zip(ref, lower))
    quals_special = "+-^$0123456789ACGTNacgtn*#<>!~"
    rows = []
    for r in range(n_rows):
        mode = rng.random()
        depth = int(rng.poisson(depth_mean)) if rng.random() > 0.08 else int(rng.integers(0, 7))
        rb = ref[r] if ref[r] in "ACGT" else 'A'
        alt = "ACGT".replace(rb, "")[int(rng.integers(0, 3))]
        af = float(rng.choice([0.0, 0.02, 0.045, 0.05, 0.055, 0.08, 0.2, 0.5, 1.0]))
        n_alleles = int(rng.integers(1, 4))
        ins_alleles = [''.join(rng.choice(list("ACGT"), size=int(rng.integers(1, 7)))) for _ in range(n_alleles)]
        del_alleles = [int(rng.integers(1, 12)) for _ in range(n_alleles)]
        parts = []
        for read in range(depth):
            rev = rng.random() < 0.5
            u = rng.random()
            base = rb
            if u < 0.01:
                base = "ACGT"[int(rng.integers(0, 4))]
            elif u < 0.02:
                base = '*' if not rev else '#'
            elif u < 0.023:
                base = 'N'
            suffix = ""
            if mode < 0.10 and rng.random() < af:
                base = alt
            elif 0.10 <= mode < 0.17 and rng.random() < af:
                seq = ins_alleles[int(rng.integers(0, n_alleles))]
                suffix = "+%d%s" % (len(seq), seq.lower() if rev else seq)
            elif 0.17 <= mode < 0.24 and rng.random() < af:
                ln = del_alleles[int(rng.integers(0, n_alleles))]
                suffix = "-%d%s" % (ln, ('n' if rev else 'N') * ln)
            elif rng.random() < 0.004:
                seq = ''.join(rng.choice(list("ACGT"), size=int(rng.integers(1, 4))))
                suffix = "+%d%s" % (len(seq), seq.lower() if rev else seq)
            tok = (base.lower() if rev and base in "ACGTN" else base) + suffix
            if weird and read + 1 < depth and rng.random() < weird:     # never last: a trailing '+0' raises in the reference
                k = int(rng.integers(0, 4))
                if k == 0:
                    tok += "+0"
                elif k == 1:
                    tok += "-2nn"                               # a second suffix replaces the first (EC:86)
                elif k == 2:
                    tok = "*+1A"
                else:
                    tok = "N-1n"
            if rng.random() < 0.04:
                tok = '^' + quals_special[int(rng.integers(0, len(quals_special)))] + tok
            if rng.random() < 0.04:
                tok += '$'
            if rng.random() < 0.005:
                tok += '<' if rev else '>'
            parts.append(tok)
        bases = ''.join(parts) if parts else '*'
        quals = ''.join(chr(33 + int(q)) for q in rng.integers(1, 50, size=max(depth, 1)))
        rows.append("%s\t%d\tN\t%d\t%s\t%s\n" % (ctg, first_pos + r, depth, bases, quals))
    return rows, reference


def hard_filter_chunk(n_sites, seed, ctg='chr20', region_lo=5001, depth=40, read_len=(120, 900), flanking=100, spacing=90,
                      with_phasing=True):
    """A phased tumour pileup chunk for the per-site hard filters: mpileup rows as printed by
    ``samtools mpileup --output-MQ --output-QNAME [--output-extra HP]`` (src/haplotype_filtering.py:303-311,
    src/postfilter_variants.py:262-268) over one region holding ``n_sites`` called variants.

    Reads are simulated (start, length, strand, haplotype tag, mapping quality); planted on them are heterozygous and
    homozygous germline variants (SNVs and insertions), the called somatic variants themselves (SNV / insertion / deletion, some
    strand biased, some only on reads that start or end nearby, some with poor base quality, some inside low-complexity
    sequence), passenger variants carried by the same reads (variant cluster) and sequencing noise.
    Returns (rows, chunk_ref, region_lo, sites) with sites = [(pos, ref_base, alt_base, af, hetero_info, homo_info)]."""
    rng = np.random.default_rng(seed)
    span = max(n_sites * spacing, 1) + 2 * flanking + 40
    ref = rng.integers(0, 4, span + 80)
    for _ in range(max(1, span // 400)):                              # low-complexity stretches (sequence entropy filter)
        at = int(rng.integers(0, span))
        unit = rng.integers(0, 4, int(rng.integers(1, 4)))
        n = int(rng.integers(20, 60))
        ref[at:at + n] = np.resize(unit, n)[:len(ref[at:at + n])]
    ref_s = ''.join("ACGT"[b] for b in ref)
    region_hi = region_lo + span - 1

    # variants: position (absolute) -> kind, payload
    site_pos = sorted(set(int(region_lo + flanking + 10 + k * spacing + rng.integers(0, spacing // 3)) for k in range(n_sites)))
    taken = set(site_pos)
    variants = {}                                                      # pos -> dict(kind, alt, length, who)

    def other_base(p):
        return "ACGT"[(ref[p - region_lo] + int(rng.integers(1, 4))) % 4]

    def rand_seq(n):
        return ''.join("ACGT"[b] for b in rng.integers(0, 4, n))

    germline = []
    for _ in range(max(2, n_sites)):
        p = int(rng.integers(region_lo + 5, region_hi - 5))
        if any(abs(p - q) < 3 for q in taken):
            continue
        taken.add(p)
        zyg = 'hom' if rng.random() < 0.35 else 'het'
        kind = 'ins' if rng.random() < 0.2 else 'snv'
        hap = int(rng.integers(1, 3))
        alt = other_base(p) if kind == 'snv' else ref_s[p - region_lo] + rand_seq(int(rng.integers(1, 5)))
        variants[p] = dict(kind=kind, alt=alt, zyg=zyg, hap=hap, origin='germline')
        germline.append((p, alt, zyg))
    for p in site_pos:
        u = rng.random()
        kind = 'snv' if u < 0.6 else 'ins' if u < 0.8 else 'del'
        length = int(rng.integers(1, 6))
        alt = other_base(p) if kind == 'snv' else ref_s[p - region_lo] + rand_seq(length) if kind == 'ins' else ref_s[p - region_lo]
        variants[p] = dict(kind=kind, alt=alt, length=length, origin='somatic', af=float(rng.uniform(0.04, 0.6)),
                           hap=int(rng.integers(0, 4)),              # 0: any read, 1/2: one haplotype, 3: both tagged haplotypes
                           style=('strand', 'ends', 'lowbq', 'lowmq', 'cluster', 'paralog', 'plain', 'plain', 'plain')[int(rng.integers(0, 9))])
        if variants[p]['style'] == 'cluster':                          # passenger variants travelling with the alt reads
            for d in (-7, 9, 23):
                q = p + d
                if q not in taken and region_lo < q < region_hi:
                    taken.add(q)
                    variants[q] = dict(kind='snv', alt=other_base(q), origin='passenger', of=p)

    # reads
    n_reads = int(depth * span / ((read_len[0] + read_len[1]) / 2.0)) + 4
    starts = rng.integers(region_lo - read_len[1] // 2, region_hi, n_reads)
    lens = rng.integers(read_len[0], read_len[1], n_reads)
    snap = rng.random(n_reads)                                         # piles of read starts / ends (read start/end filter)
    ends = starts + lens - 1
    starts = np.where(snap < 0.25, starts // 61 * 61, starts)
    ends = np.where(snap > 0.75, ends // 67 * 67, ends)
    lens = np.maximum(ends - starts + 1, 30)
    bundle_of = np.zeros(n_reads, np.int64)
    for p in site_pos:                                                 # a pile of reads that all begin just before the site
        if variants[p]['style'] == 'ends':
            k = max(3, int(depth * 0.4))
            starts = np.concatenate([starts, np.full(k, p - 2)])
            lens = np.concatenate([lens, rng.integers(read_len[0], read_len[1], k)])
            bundle_of = np.concatenate([bundle_of, np.full(k, p)])
    n_reads = len(starts)
    order = np.argsort(starts, kind='stable')
    starts, lens, bundle_of = starts[order], lens[order], bundle_of[order]
    reads = []
    for k in range(n_reads):
        reads.append(dict(name="read%05d/%d" % (k, seed % 97), lo=int(starts[k]), hi=int(starts[k] + lens[k] - 1),
                          rev=bool(rng.random() < 0.5), hap=int(rng.choice([0, 1, 2], p=[0.25, 0.4, 0.35])),
                          mq=int(60 if rng.random() < 0.85 else rng.integers(0, 60)), carries={}, bundle=int(bundle_of[k])))
    longest = int(lens.max()) if n_reads else 0

    def covering(p):                                                   # reads are sorted by start
        for r in reads[int(np.searchsorted(starts, p - longest)):int(np.searchsorted(starts, p, 'right'))]:
            if r['lo'] <= p <= r['hi']:
                yield r

    for p, v in variants.items():
        for r in covering(p):
            if v['origin'] == 'germline':
                hap = r['hap'] if r['hap'] else int(rng.integers(1, 3))
                on = v['zyg'] == 'hom' or hap == v['hap']
                on = on and rng.random() < 0.97
            elif v['origin'] == 'somatic':
                ok_hap = v['hap'] == 0 or (v['hap'] == 3 and r['hap'] > 0) or r['hap'] == v['hap']
                on = ok_hap and rng.random() < v['af'] * (1.6 if v['hap'] in (1, 2) else 1.0)
                if v['style'] == 'strand':
                    on = on and not r['rev']
                if v['style'] == 'ends':
                    on = rng.random() < (0.9 if r['bundle'] == p else 0.03)
                if v['style'] == 'lowmq':
                    on = ok_hap and rng.random() < (0.9 if r['mq'] < 30 else 0.01)
            else:
                continue
            if on:
                r['carries'][p] = v
    for p, v in variants.items():
        if v['origin'] == 'somatic' and v['style'] == 'paralog':       # alt reads from elsewhere: they lack the germline alleles
            for r in covering(p):
                if p in r['carries']:
                    r['carries'] = {q: w for q, w in r['carries'].items() if w['origin'] != 'germline'}
    for p, v in variants.items():
        if v['origin'] == 'passenger':
            for r in covering(p):
                if v['of'] in r['carries'] and rng.random() < 0.95:
                    r['carries'][p] = v

    sites = []
    for p in site_pos:
        v = variants[p]
        near = [(g, alt, z) for g, alt, z in germline if p - flanking < g <= p + flanking and g != p]
        het = ','.join("%d-%s" % (g, alt) for g, alt, z in near if z == 'het')
        hom = ','.join("%d-%s" % (g, alt) for g, alt, z in near if z == 'hom')
        rb = ref_s[p - region_lo] if v['kind'] != 'del' else ref_s[p - region_lo: p - region_lo + 1 + v['length']]
        sites.append((p, rb, v['alt'], round(v['af'] * float(rng.uniform(0.5, 1.2)), 4), het, hom))

    # rows
    rows = []
    deleted_until = {}                                                 # read index -> last deleted position
    active = []
    nxt = 0
    for p in range(region_lo, region_hi + 1):
        while nxt < n_reads and reads[nxt]['lo'] <= p:
            active.append(nxt)
            nxt += 1
        active = [k for k in active if reads[k]['hi'] >= p]
        if not active:
            continue
        bases, bqs, mqs, names, hps = [], [], [], [], []
        rb = ref_s[p - region_lo]
        for k in active:
            r = reads[k]
            tok = ''
            if r['lo'] == p:
                tok += '^' + chr(min(r['mq'], 60) + 33)
            bq = int(np.clip(np.rint(rng.normal(28, 9)), 1, 60))
            if deleted_until.get(k, 0) >= p:
                sym = '*'
            else:
                v = r['carries'].get(p)
                sym = rb
                if v is not None and v['kind'] == 'snv':
                    sym = v['alt']
                    if v.get('style') == 'lowbq':
                        bq = int(rng.integers(3, 19))
                elif rng.random() < 0.008:
                    sym = "ACGTN"[int(rng.integers(0, 5))]
                suffix = ''
                if v is not None and v['kind'] == 'ins':
                    suffix = '+%d%s' % (len(v['alt']) - 1, v['alt'][1:])
                elif v is not None and v['kind'] == 'del' and p + v['length'] <= r['hi']:
                    suffix = '-%d%s' % (v['length'], ref_s[p - region_lo + 1: p - region_lo + 1 + v['length']])
                    deleted_until[k] = p + v['length']
                elif rng.random() < 0.004 and p + 12 <= r['hi']:
                    n = int(rng.integers(1, 9))
                    if rng.random() < 0.5:
                        suffix = '+%d%s' % (n, rand_seq(n))
                    else:
                        suffix = '-%d%s' % (n, ref_s[p - region_lo + 1: p - region_lo + 1 + n])
                        deleted_until[k] = p + n
                sym += suffix
            if r['rev']:
                sym = sym.lower().replace('*', '#')
            tok += sym
            if r['hi'] == p:
                tok += '$'
            bases.append(tok)
            bqs.append(chr(bq + 33))
            mqs.append(chr(min(r['mq'], 60) + 33))
            names.append(r['name'])
            hps.append(str(r['hap']) if r['hap'] else ('0' if rng.random() < 0.5 else '*'))
        cols = [ctg, str(p), 'N', str(len(active)), ''.join(bases), ''.join(bqs), ''.join(mqs), ','.join(names)]
        if with_phasing:
            cols.append(','.join(hps))
        rows.append('\t'.join(cols) + '\n')
    return rows, ref_s[:span], region_lo, sites
