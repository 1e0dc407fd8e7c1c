"""Oracle for the pileup tensor encoder.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure Python on purpose: it
restates the reference's per-read string logic one read at a time so that every
quirk is visible, and it is only run on small cases.

Reference: src/create_tensor_pileup_calling.py (cited below as CT:line).
"""

from __future__ import annotations

# channel order, CT:55-58
CHANNELS = ['A', 'C', 'G', 'T', 'I', 'I1', 'D', 'D1', '*', 'a', 'c', 'g', 't', 'i', 'i1', 'd', 'd1', '#',
            'ALMQ', 'CLMQ', 'GLMQ', 'TLMQ', 'aLMQ', 'cLMQ', 'gLMQ', 'tLMQ',
            'ALBQ', 'CLBQ', 'GLBQ', 'TLBQ', 'aLBQ', 'cLBQ', 'gLBQ', 'tLBQ']
CH = {name: i for i, name in enumerate(CHANNELS)}
N_CHANNELS = len(CHANNELS)          # 34
FLANK = 16                          # shared/param.py:59
N_POS = 2 * FLANK + 1               # shared/param.py:60
READ_SYMBOLS = "ACGTNacgtn#*"       # CT:140
MAX_INDEL_LENGTH = 60               # shared/param.py:101
MIN_MQ = 20                         # literal in CT:147-148


def coerce_ref_base(base: str) -> str:
    """CT:82-92 (evc_base_from) followed by the .upper() at CT:485."""
    if base in 'ACGTacgt':
        return base.upper()
    return 'A'


def tokenize(pileup_bases: str):
    """CT:120-144.  Returns the list of [symbol, indel] read entries.

    '+'/'-' attach "<sign><seq>" to the PREVIOUS entry (overwriting an earlier
    indel on it), '^' swallows the following mapping-quality character, every
    other character ('$', '<', '>', ...) is dropped without producing an entry.
    """
    entries = []
    i, n = 0, len(pileup_bases)
    while i < n:
        ch = pileup_bases[i]
        if ch == '+' or ch == '-':
            j = i + 1
            length = 0
            while pileup_bases[j].isdigit():
                length = length * 10 + int(pileup_bases[j])
                j += 1
            entries[-1][1] = ch + pileup_bases[j:j + length]
            i = j + length
            continue
        if ch in READ_SYMBOLS:
            entries.append([ch, ""])
        elif ch == '^':
            i += 1
        i += 1
    return entries


def _ordered_counts(keys):
    """collections.Counter semantics as used at CT:146-149: first-occurrence order."""
    out = {}
    for k in keys:
        out[k] = out.get(k, 0) + 1
    return out


def position_vector(pileup_bases, mapping_quality, base_quality, reference_base,
                    is_candidate=False, chunk_ref_seq="", platform="ont",
                    max_indel_length=MAX_INDEL_LENGTH):
    """One genomic position -> (int[34], alt_info or None).  CT:95-233.

    mapping_quality / base_quality are lists of ints (phred), zipped positionally
    against the tokenised entries exactly like the reference (CT:147-149), so a
    shorter quality list silently truncates.
    """
    entries = tokenize(pileup_bases)
    keys = [e[0] + e[1] for e in entries]
    low_bq_cut = 30 if platform == 'ont' else 10                       # CT:149
    main = _ordered_counts(k for k, mq in zip(keys, mapping_quality) if mq >= MIN_MQ)
    low_mq = _ordered_counts(k for k, mq in zip(keys, mapping_quality) if mq < MIN_MQ)
    low_bq = _ordered_counts(k for k, bq in zip(keys, base_quality) if bq < low_bq_cut)

    vec = [0] * N_CHANNELS
    depth = 0
    best = {'I': 0, 'i': 0, 'D': 0, 'd': 0}
    alt = {}
    ref_count = 0
    for key, count in main.items():
        if len(key) == 1:
            up = key.upper()
            if up in 'ACGT':                                            # CT:160-168
                if is_candidate:
                    if up != reference_base:
                        alt['X' + up] = alt.get('X' + up, 0) + count
                    else:
                        ref_count += count
                depth += count
                vec[CH[key]] += count
            elif key in '#*':                                           # CT:169-171
                vec[CH[key]] += count
                depth += count
            continue
        forward = key[0] in 'ACGTN*'                                    # CT:182, 199
        if key[1] == '+':                                               # CT:172-187
            if len(key) - 2 > max_indel_length:
                continue
            depth += count
            if is_candidate:
                name = 'I' + key[0].upper() + key[2:].upper()
                alt[name] = alt.get(name, 0) + count
            ch = 'I' if forward else 'i'
        else:                                                           # CT:188-204
            span = len(key) - 1          # includes the '-' sign
            if span > max_indel_length:
                continue
            depth += count
            if is_candidate:
                name = 'D' + chunk_ref_seq[:span]
                alt[name] = alt.get(name, 0) + count
            ch = 'D' if forward else 'd'
        vec[CH[ch]] += count
        best[ch] = max(best[ch], count)
    if is_candidate and ref_count > 0:                                  # CT:205-206
        alt['R' + reference_base] = ref_count

    vec[CH['I1']], vec[CH['i1']] = best['I'], best['i']                 # CT:210-213
    vec[CH['D1']], vec[CH['d1']] = best['D'], best['d']

    for counts, suffix in ((low_mq, 'LMQ'), (low_bq, 'LBQ')):           # CT:215-221
        for key, count in counts.items():
            if len(key) == 1 and key.upper() in 'ACGT':
                vec[CH[key + suffix]] += count

    for group, names in (('', 'ACGT'), ('', 'acgt'), ('LMQ', 'ACGT'), ('LMQ', 'acgt'),
                         ('LBQ', 'ACGT'), ('LBQ', 'acgt')):             # CT:223-228
        total = sum(vec[CH[b + group]] for b in names)
        ref = reference_base if names == 'ACGT' else reference_base.lower()
        vec[CH[ref + group]] = -total

    alt_info = None
    if is_candidate:                                                    # CT:208-209
        alt_info = "%d-%s-" % (depth, ' '.join("%s %d" % kv for kv in alt.items()))
    return vec, alt_info


def parse_mpileup_row(row: str):
    """CT:472-490: columns chr, pos, ref, depth, bases, BQ, MQ (phred+33)."""
    cols = row.rstrip('\n').split('\t')
    pos = int(cols[1])
    bq = [ord(c) - 33 for c in cols[5]]
    mq = [ord(c) - 33 for c in cols[6]]
    return pos, cols[4], bq, mq


def encode_windows(mpileup_rows, candidates, reference_sequence, reference_start,
                   extend_start, extend_end, ctg_name, platform="ont",
                   candidate_types=None, max_indel_length=MAX_INDEL_LENGTH):
    """Window assembly, CT:461, 465-570.

    mpileup_rows: iterable of text rows; candidates: iterable of 1-based centre
    positions.  Returns a list of (pos, ref33, int[33][34], alt_info, variant_type,
    ref_centre) in the order the reference writes them (sorted by position), and
    the exact text rows of the tensor_can file.
    """
    candidate_types = candidate_types or {}
    cand = sorted(set(candidates))
    cand_set = set(cand)
    zero = [0] * N_CHANNELS
    table = [zero] * (extend_end - extend_start + FLANK)                # CT:461
    alt_infos = {}
    for row in mpileup_rows:
        pos, bases, bq, mq = parse_mpileup_row(row)
        ref_base = coerce_ref_base(reference_sequence[pos - reference_start])
        off = pos - reference_start
        chunk_ref = reference_sequence[off:off + max_indel_length].upper()   # CT:497
        # create_tensor() never forwards its --platform to decode_pileup_bases (CT:499-511), so the
        # callee's default platform="ont" applies and the low-BQ cut is ALWAYS 30 in the pipeline
        vec, alt_info = position_vector(bases, mq, bq, ref_base, is_candidate=pos in cand_set,
                                        chunk_ref_seq=chunk_ref, platform="ont",
                                        max_indel_length=max_indel_length)
        table[pos - extend_start] = vec                                 # CT:513-514
        if pos in cand_set:
            alt_infos[pos] = alt_info
    out, text = [], []
    for pos in cand:
        start = pos - FLANK - extend_start
        end = start + N_POS
        if start < 0 or end >= extend_end - extend_start:               # CT:542-543
            continue
        if pos not in alt_infos:                                        # CT:552-553
            continue
        off = pos - reference_start
        ref33 = reference_sequence[off - FLANK: off + FLANK + 1].upper()
        window = [list(v) for v in table[start:end]]
        vtype = candidate_types.get(pos, 'unknown')
        centre = ref33[FLANK]
        out.append((pos, ref33, window, alt_infos[pos], vtype, centre))
        flat = " ".join(" ".join("%d" % x for x in v) for v in window)  # CT:551
        text.append("%s\t%d\t%s\t%s\t%s\t%s\t%s\n" % (ctg_name, pos, ref33, flat, alt_infos[pos], vtype, centre))
    return out, text
