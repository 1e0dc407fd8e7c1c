"""ORACLE (test infrastructure, not product code): CPU restatement of the per-position candidate decision of ClairS-TO's
STEP 1, ``src/extract_candidates_calling.py`` (cited as EC), SURVEY.md section 8 row f3.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

Pinned against the unmodified reference two ways: ``tests/test_oracle_vs_reference.py`` fuzzes ``site_decision`` against
``decode_pileup_bases`` imported from /root/reference (when present), and ``tests/golden/candidates`` holds the output
files of the reference sub-command run through ``tests/fake_samtools.py`` (generator: ``tests/golden/make_golden.py``).
"""

from collections import Counter


def tokenize(pileup_bases):
    """EC:73-93: ``[symbol, indel]`` per read.  ``+N<seq>`` / ``-N<seq>`` attach to the previous read entry, ``^x`` skips
    two characters, everything else (``$``, ``<``, ``>`` ...) is skipped."""
    base_list = []
    i, n = 0, len(pileup_bases)
    while i < n:
        base = pileup_bases[i]
        if base == '+' or base == '-':
            i += 1
            advance = 0
            while True:
                num = pileup_bases[i]
                if num.isdigit():
                    advance = advance * 10 + int(num)
                    i += 1
                else:
                    break
            base_list[-1][1] = base + pileup_bases[i: i + advance]
            i += advance - 1
        elif base in "ACGTNacgtn#*":
            base_list.append([base, ""])
        elif base == '^':
            i += 1
        i += 1
    return base_list


def site_decision(pileup_bases, reference_base, min_coverage, snv_min_af, indel_min_af, alternative_base_num,
                  select_indel_candidates):
    """EC:95-147 + the set logic of EC:366-377.  Returns (depth, pass_af, is_snv_candidate, is_indel_candidate)."""
    base_list = tokenize(pileup_bases)
    pileup = {}
    base_counter = Counter(''.join(item) for item in base_list)
    alt_keys = set(''.join(item).upper() for item in base_list)
    depth = 0
    for key, count in base_counter.items():                                    # EC:104-120
        up = key[0].upper()
        if up in 'ACGT':
            pileup[up] = pileup.get(up, 0) + count
            depth += count
        elif key[0] in '#*':
            depth += count
        if len(key) > 1 and key[1] == '+':
            k = ('I' + up + key[2:].upper()) if select_indel_candidates else 'I'
            pileup[k] = pileup.get(k, 0) + count
        elif len(key) > 1 and key[1] == '-':
            k = ('D' + len(key[2:]) * 'N') if select_indel_candidates else 'D'
            pileup[k] = pileup.get(k, 0) + count
    denominator = depth if depth > 0 else 1
    pass_snv_af = pass_indel_af = False
    pass_depth = depth > min_coverage                                          # EC:127
    for item, count in pileup.items():                                         # EC:128-137 (order does not matter for an OR)
        if item == reference_base:
            continue
        if item[0] in 'ID':
            if select_indel_candidates:
                pass_indel_af = pass_indel_af or (float(count) / denominator >= indel_min_af and
                                                  (alternative_base_num is not None and count >= alternative_base_num))
            continue
        pass_snv_af = pass_snv_af or (float(count) / denominator >= snv_min_af) and (
            alternative_base_num is not None and count >= alternative_base_num)
    pass_af = (pass_snv_af or pass_indel_af) and pass_depth                    # EC:144
    alt = [k for k in alt_keys if k.upper() != reference_base]                 # EC:146-148
    snv = pass_af and pass_snv_af and any(k in "ACGT" for k in alt)            # EC:366-371 (substring test of the whole key)
    indel = bool(select_indel_candidates) and pass_af and pass_indel_af and any('+' in k or '-' in k for k in alt)   # EC:372-377
    return depth, bool(pass_af), bool(snv), bool(indel)


def scan_rows(rows, reference_sequence, reference_start, **kw):
    """EC:335-377 over mpileup rows (text lines): {pos: (depth, pass_af, snv, indel)} for rows whose reference base is ACGT."""
    out = {}
    for row in rows:
        columns = row.strip().split('\t')
        pos = int(columns[1])
        reference_base = reference_sequence[pos - reference_start].upper()
        if reference_base not in "ACGT":
            continue
        out[pos] = site_decision(columns[4], reference_base, **kw)
    return out


def chunk_range(chunk_id, chunk_num, contig_length=None, bed_range=None):
    """EC:236-262: (ctg_start, ctg_end) of 0-based chunk ``chunk_id`` of ``chunk_num``; with a confident BED the split is
    over (bed_start, bed_end) instead of the contig length."""
    if bed_range is not None:
        bed_start, bed_end = bed_range
        span = bed_end - bed_start
        chunk_size = span // chunk_num + 1 if span % chunk_num else span // chunk_num
        ctg_start = bed_start + 1 + chunk_size * chunk_id
    else:
        chunk_size = contig_length // chunk_num + 1 if contig_length % chunk_num else contig_length // chunk_num
        ctg_start = chunk_size * chunk_id
    return ctg_start, ctg_start + chunk_size


def reads_region(ctg_start, ctg_end, flanking=16):
    """EC:271-275: the 1-based inclusive range handed to ``samtools mpileup -r``."""
    n = 2 * flanking + 1
    return max(ctg_start - n, 1), ctg_end + n


def candidate_lists(rows, reference_sequence, reference_start, indel_intervals=None, **kw):
    """EC:335-404: (every candidate, SNV candidates, indel candidates) as sorted position lists.  ``indel_intervals``:
    half-open [start, end) intervals of --call_indels_only_in_these_regions (None / empty = no filter, EC:397-403)."""
    sites = scan_rows(rows, reference_sequence, reference_start, **kw)
    every = sorted(p for p, v in sites.items() if v[1])
    snv = sorted(p for p, v in sites.items() if v[2])
    indel = sorted(p for p, v in sites.items() if v[3])
    if indel_intervals:
        indel = [p for p in indel if any(s < p and e > p - 1 for s, e in indel_intervals)]
    return every, snv, indel


def region_rows(ctg_name, positions, flanking=16):
    """EC:458-461: the rows of a region file."""
    return ['\t'.join([ctg_name, str(max(x - flanking - 1, 1)), str(x + flanking + 1)]) for x in positions]
