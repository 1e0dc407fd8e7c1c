"""ORACLE (test infrastructure, not product code): CPU restatement of ClairS-TO's per-site hard filters, SURVEY.md
section 8 row f4.  HF = ``src/haplotype_filtering.py`` (long reads, phased), PV = ``src/postfilter_variants.py`` (short reads).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

Pinned against the unmodified reference by ``tests/test_oracle_vs_reference.py`` (live fuzz of ``site_line`` against
``_haplotype_build_state_and_line`` / ``_postfilter_build_state_and_line`` imported from /root/reference) and by
``tests/golden/hard_filter`` (lines the reference functions returned; generator ``tests/golden/make_golden.py``).

One function serves both files: the PV variant is the HF variant without haplotypes, germline sets and BQ/MQ tests, with
three differences that are kept (``mode``): the read-start/end test is not guarded against an empty alt read set (PV:294-296
vs HF:369-373), the strand-bias rule (PV:353-354 vs HF:545-548) and missing reference bases raise instead of skipping.
"""

import math
from collections import Counter

LOW_AF_SNV, LOW_AF_INDEL = 0.1, 0.3          # HF:22-23
MIN_HOM_GERMLINE_AF = 0.75                   # HF:24
EPS, EPS_RSE = 0.5, 0.2                      # HF:26-27, PV:19-20
ENTROPY_THRESHOLD = 0.9                      # HF:28
ENTROPY_WINDOW, ENTROPY_FLANK = 33, 16       # shared/param.py:60-61
ENTROPY_CENTRE = 100                         # HF:29: the slice uses the MODULE constant `flanking`, not the --flanking argument
MIN_ALT_BQ, MIN_ALT_MQ = 20, 20              # shared/param.py:17-18 (ont_min_bq, min_mq)
BASE2NUM = dict(zip("ACGTURYSWKMBDHVN", (0, 1, 2, 3, 3, 0, 1, 1, 0, 2, 0, 1, 0, 0, 0, 0)))   # shared/utils.py:18-21


def binomial(n, k):
    """HF:44-57: exact integer C(n, k)."""
    return math.comb(n, k) if 0 <= k <= n else 0


def fisher_exact(a, b, c, d):
    """HF:60-98: two-sided p of the 2x2 table [[a, b], [c, d]]: the table's own probability (exact integers, one correctly
    rounded division) plus every more extreme table's probability <= it, walked by multiply / divide in double precision."""
    if a == b == c == d:
        return 1.0
    p = t = binomial(a + b, a) * binomial(c + d, c) / binomial(a + b + c + d, a + c)
    for step in (-1, 1):
        w, x, y, z = a, b, c, d
        cur = float(t)
        side = 0.0
        while (w > 0 and z > 0) if step < 0 else (x > 0 and y > 0):
            if step < 0:
                cur *= w * z
                w, x, y, z = w - 1, x + 1, y + 1, z - 1
                cur /= x * y
            else:
                cur *= x * y
                w, x, y, z = w + 1, x - 1, y - 1, z + 1
                cur /= w * z
            if cur <= t:
                side += cur
        p += side
    return p


def entropy_table(window=ENTROPY_WINDOW):
    """HF:106-109: e * log(e) for e = i / window, and the final multiplier."""
    tab = [0.0] * (window + 2)
    for i in range(1, window + 2):
        e = 1.0 / window * i
        tab[i] = e * math.log(e)
    return tab, -1 / math.log(window)


def sequence_entropy(site_ref):
    """HF:144-151: the entropy of the 33 bases around offset 100 of the site's reference window."""
    return entropy_of(site_ref[ENTROPY_CENTRE - ENTROPY_FLANK: ENTROPY_CENTRE + ENTROPY_FLANK + 1])


def entropy_of(seq):
    """HF:101-142: 5-mer entropy of a sequence over a window of 33; the running sum is updated in the reference's order
    (subtract the old term, add the new one) because the result is compared in floating point."""
    tab, mul = entropy_table()
    counts = {}
    kmer, total = 0, 0.0
    i, i2 = 0, -ENTROPY_WINDOW
    prefix = 0
    while i2 < len(seq):
        if i < len(seq):
            kmer = ((kmer << 2) | BASE2NUM[seq[i]]) & 1023
            c = counts.get(kmer, 0)
            total -= tab[c]
            counts[kmer] = c + 1
            total += tab[c + 1]
        if i2 >= 0 and i < len(seq):                      # only for sequences longer than the window (never for 33 bases)
            prefix = ((prefix << 2) | BASE2NUM[seq[i2]]) & 1023
            c = counts.get(prefix, 0)
            total -= tab[c]
            counts[prefix] = c - 1
            total += tab[c - 1]
        i += 1
        i2 += 1
    return total * mul


class Row:
    __slots__ = ("names", "toks", "counter", "rse", "phasing", "bq", "mq")


def parse_row(columns, with_phasing):
    """HF:154-185 + 246-275 (PV:144-175 + 237-259): one mpileup row with --output-MQ --output-QNAME [--output-extra HP].
    Read keys are QNAME + '_1' (reverse strand: lower-case symbol or '#') / '_0'; without the HP column the row's last
    QNAME still carries the line feed, so that read has a different key in rows where it is not last.  Quirks kept: ``^`` marks the read BEFORE
    it (index -1 = the row's last read, by Python indexing), ``$`` the read it follows; the larger of the two sets is used."""
    s = columns[4]
    toks, starts, ends = [], set(), set()
    i = 0
    while i < len(s):
        ch = s[i]
        if ch in '+-':
            i += 1
            n = 0
            while s[i].isdigit():
                n = n * 10 + int(s[i])
                i += 1
            toks[-1][1] = ch + s[i:i + n]
            i += n - 1
        elif ch in "ACGTNacgtn#*":
            toks.append([ch, ""])
        elif ch == '^':
            i += 1
            starts.add(len(toks) - 1)
        if ch == '$':
            ends.add(len(toks) - 1)
        i += 1
    r = Row()
    r.rse = starts if len(starts) > len(ends) else ends
    r.counter = Counter((a + b).upper() for a, b in toks)
    names = columns[7].split(',')            # PV:243: the row's last name keeps its '\n' (there is no ninth column), kept
    r.names = [nm + ('_1' if (t[0] == '#' or 'a' <= t[0] <= 'z') else '_0') for nm, t in zip(names, toks)]
    r.toks = [(a.upper(), b) for a, b in toks]
    r.phasing = columns[8].strip('\n').split(',') if with_phasing else None
    r.bq = [ord(q) - 33 for q in columns[5]]
    r.mq = [ord(q) - 33 for q in columns[6]]
    return r


def parse_chunk(lines, with_phasing):
    rows = {}
    for line in lines:
        columns = line.split('\t')
        if len(columns) < (9 if with_phasing else 8):
            continue
        rows[int(columns[1])] = parse_row(columns, with_phasing)
    return rows


def _alt_match(kind, ref_base, alt_base):
    if kind == 'snp':
        return lambda t: t[0] + t[1] == alt_base
    if kind == 'ins':
        return lambda t: '+' in t[0] + t[1] and (t[0] + t[1]).replace('+', '').upper() == alt_base
    if kind == 'del':
        return lambda t: len(ref_base) == len(t[1]) and '-' in t[1]
    return lambda t: False


def _germline_match(rb, ab, second):
    """HF:444-451 (heterozygous, ``ab[:2]``) / HF:473-481 (homozygous, ``ab[1:2]``): which reads of a row carry the germline
    allele.  The insertion test is a case-sensitive substring test on the raw suffix, kept as it is."""
    if len(rb) == 1 and len(ab) == 1:
        return lambda t: t[0] + t[1] == ab
    if len(rb) == 1 and len(ab) > 1:
        key = ab[1:2] if second else ab[:2]
        return lambda t: len(t[1]) > 1 and key in t[1][1:]
    if len(rb) > 1 and len(ab) == 1:
        return lambda t: '-' in t[0] + t[1]
    return lambda t: False


def site_line(mode, ctg_name, pos, ref_base, alt_base, flanking, rows, chunk_ref, region_lo, hetero_info=None, homo_info=None,
              disable_read_start_end_filtering=False, max_co_exist_read_num=3, af=None):
    """HF:570-703 + 344-565 (``mode='haplotype'``) / PV:368-446 + 278-365 (``mode='postfilter'``): the result line of one site."""
    hap_mode = mode == 'haplotype'
    kind = 'snp' if len(ref_base) == 1 and len(alt_base) == 1 else 'ins' if len(ref_base) == 1 else 'del' if len(alt_base) == 1 else 'other'
    is_snp = kind == 'snp'
    anchor = max(pos - flanking, 1)
    site_ref = (chunk_ref or '')[anchor - region_lo: pos + flanking + 1 - region_lo + 1]
    hetero = set(tuple(x.split('-')) for x in hetero_info.split(',')) if hap_mode and hetero_info else set()
    homo = set(tuple(x.split('-')) for x in homo_info.split(',')) if hap_mode and homo_info else set()
    hetero_pos = set(int(x[0]) for x in hetero)
    match = _alt_match(kind, ref_base, alt_base)

    hap = {}                                   # read key -> 1 / 2 (absent = 0)
    by_pos, counters = {}, {}
    rse_reads, alt_reads = set(), set()
    all_n, alt_n = [0, 0, 0], [0, 0, 0]
    all_f, all_r, alt_f, alt_r = [0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]
    pass_bq = pass_mq = True
    for p in range(anchor, pos + flanking + 1):
        row = rows.get(p)
        if row is None:
            continue
        if hap_mode and (p in hetero_pos or p == pos):                         # HF:621-625: first haplotype seen wins
            for k, h in enumerate(row.phasing):
                if h in '12' and row.names[k] not in hap:
                    hap[row.names[k]] = int(h)
        if len(row.rse) >= len(row.toks) * EPS_RSE:                            # HF:627-628
            rse_reads.update(row.names[k] for k in row.rse)
        by_pos[p] = dict(zip(row.names, row.toks))
        if p == pos:
            if hap_mode:                                                       # HF:632-662: mean BQ / MQ of the alt reads
                bqs = [q for _, t, q in zip(row.names, row.toks, row.bq) if match(t)]
                mqs = [q for _, t, q in zip(row.names, row.toks, row.mq) if match(t)]
                if bqs and sum(bqs) / len(bqs) <= MIN_ALT_BQ:
                    pass_bq = False
                if mqs and sum(mqs) / len(mqs) <= MIN_ALT_MQ:
                    pass_mq = False
            for nm in row.names:                                               # HF:664-669
                h = hap.get(nm, 0)
                all_n[h] += 1
                (all_f if nm.endswith('0') else all_r)[h] += 1
            alt_reads = set(nm for nm, t in zip(row.names, row.toks) if match(t))
            for nm in alt_reads:                                               # HF:683-688
                h = hap.get(nm, 0)
                alt_n[h] += 1
                (alt_f if nm.endswith('0') else alt_r)[h] += 1
        ci = p - region_lo
        if hap_mode and (not chunk_ref or ci < 0 or ci >= len(chunk_ref)):     # HF:690-692
            continue
        centre = chunk_ref[ci] if hap_mode else site_ref[p - anchor]          # PV:404 raises past the reference
        if len(row.counter) == 1 and row.counter[centre] > 0:
            continue
        counters[p] = row.counter

    pass_rse = True
    if not disable_read_start_end_filtering and (alt_reads or not hap_mode):   # HF:369-373 / PV:294-296
        if len(rse_reads & alt_reads) >= 0.3 * len(alt_reads):
            pass_rse = False

    pass_both = True
    h1, h2 = alt_n[1], alt_n[2]
    hi, lo = max(h1, h2), min(h1, h2)
    af = 1.0 if af is None else af
    if hap_mode and af < (LOW_AF_SNV if is_snp else LOW_AF_INDEL):            # HF:380-385
        if h1 * h2 > 0 and (lo > max_co_exist_read_num or hi / lo <= 10):
            pass_both = False
    phasable = h1 * h2 == 0 or (hi / lo >= 5 and (h1 > max_co_exist_read_num or h2 > max_co_exist_read_num))   # HF:387
    hap_index = 0 if not (hap_mode and phasable) else (1 if h1 > h2 else 2)

    match_count = ins_length = 0
    for p, reads in by_pos.items():                                            # HF:394-434 / PV:304-341: variant cluster
        ri = p - anchor
        if hap_mode and (ri < 0 or ri >= len(site_ref)):
            continue
        rb = site_ref[ri]
        if p == pos:
            continue
        ins_length += sum(min(len(t[1]) - 1, flanking * 2) for t in reads.values() if len(t[1]) > 3 and t[1][0] == '+')
        carried = Counter()
        for nm in reads.keys() & alt_reads:
            tok = (reads[nm][0] + reads[nm][1]).upper()
            if tok != rb and tok not in '#*':
                carried[tok] += 1
        if not carried:
            continue
        top_tok, top = max(carried.items(), key=lambda kv: kv[1])              # a tie cannot pass the bounds below (EPS = 0.5)
        if top >= len(alt_reads) * (1 + EPS) or top <= len(alt_reads) * (1 - EPS):
            continue
        if p not in counters or counters[p][top_tok] >= top * (1 + EPS):
            continue
        match_count += 1

    pass_hetero = pass_homo = True
    if hap_index > 0:                                                          # HF:436-468
        for gp, ab in hetero:
            gp = int(gp)
            ri = gp - anchor
            if gp not in by_pos or ri < 0 or ri >= len(site_ref):
                continue
            reads = by_pos[gp]
            carries = _germline_match(site_ref[ri], ab, second=False)
            overlap = set(nm for nm, t in reads.items() if carries(t))
            phased = set(nm for nm in overlap if hap.get(nm, 0) == hap_index)
            if not phased or len(phased) * 2 < float(len(overlap)):
                continue
            if not (set(nm for nm in alt_reads if hap.get(nm, 0) == hap_index) & phased):
                pass_hetero = False
                break
    for gp, ab in homo:                                                        # HF:470-523
        gp = int(gp)
        ri = gp - anchor
        if gp not in by_pos or ri < 0 or ri >= len(site_ref):
            continue
        reads = by_pos[gp]
        carries = _germline_match(site_ref[ri], ab, second=True)
        carriers = [nm for nm, t in reads.items() if carries(t)]
        c = [0, 0, 0]
        for nm in carriers:
            c[hap.get(nm, 0)] += 1
        a = [0, 0, 0]
        for nm in reads:
            a[hap.get(nm, 0)] += 1
        af_g = sum(c) / float(sum(a)) if sum(a) > 0 else 0.0
        unphasable_allele = a[1] * a[2] == 0 or (c[1] * c[2] > 0 and max(c[1], c[2]) / min(c[1], c[2]) <= 10)   # HF:492-503, negated
        if af_g < MIN_HOM_GERMLINE_AF or not unphasable_allele:
            continue
        inter = reads.keys() & alt_reads
        if not inter:
            continue
        both = [nm for nm in carriers if nm in inter]
        if not both or len(both) / len(inter) < EPS:
            pass_homo = False
            break

    depth = sum(all_n) if sum(all_n) > 0 else 1
    pass_co_exist = not (match_count >= max_co_exist_read_num or ins_length / depth > 3)
    phaseable = all_n[1] * all_n[2] > 0 and alt_n[1] * alt_n[2] == 0 and (alt_n[1] > max_co_exist_read_num or alt_n[2] > max_co_exist_read_num)
    a0, a1 = sum(alt_f), sum(alt_r)
    r0, r1 = sum(all_f) - a0, sum(all_r) - a1
    p_value = fisher_exact(a0, r0, a1, r1)
    if hap_mode:                                                               # HF:545-548: `and` binds tighter than `or`
        limit = 0.001 if is_snp else 0.01
        pass_sb = not (p_value < limit or a0 == 0 or a1 == 0)
    else:
        pass_sb = not p_value < 0.001                                          # PV:353-354
    pass_entropy = True
    if not is_snp and sequence_entropy(site_ref) < ENTROPY_THRESHOLD:
        pass_entropy = False

    if hap_mode:
        verdict = (pass_hetero and pass_homo and pass_both and pass_rse and pass_bq and pass_mq and pass_co_exist and pass_sb
                   and pass_entropy)
        fields = [verdict, phaseable, pass_hetero, pass_homo, pass_rse, pass_bq, pass_mq, pass_co_exist, pass_both, pass_sb,
                  round(p_value, 5), pass_entropy]
    else:
        verdict = pass_rse and pass_co_exist and pass_sb and pass_entropy
        fields = [verdict, pass_rse, pass_co_exist, pass_sb, round(p_value, 5), pass_entropy]
    return ' '.join([ctg_name, str(pos)] + [str(x) for x in fields])
