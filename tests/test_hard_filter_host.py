"""CPU tests of the hard-filter host tokenizer: ``cto_hf_parse`` against the oracle's row parser
(oracle/hard_filter_oracle.py, itself pinned to the reference by tests/test_oracle_vs_reference.py and tests/golden/hard_filter)."""

import numpy as np
import pytest

from clairs_to_b200 import hard_filters as hf
from clairs_to_b200 import _lib, synth
from oracle import hard_filter_oracle as ho


def rebuild_rows(chunk):
    """{pos: (keys, tokens, rse indices resolved, hp, bq, mq)} from the integer arrays; read keys come back as ids."""
    out = {}
    for r, pos in enumerate(chunk.row_pos):
        lo, hi = chunk.row_off[r], chunk.row_off[r + 1]
        out[int(pos)] = dict(rid=chunk.rid[lo:hi], tok=[chunk.tokens[k] for k in chunk.tok[lo:hi]],
                             sfx=[chunk.suffixes[k] for k in chunk.sfx[lo:hi]], info=chunk.info[lo:hi], qual=chunk.qual[lo:hi],
                             rse=sorted(int(e - lo) for e in chunk.rse_ent[chunk.rse_off[r]:chunk.rse_off[r + 1]]),
                             flags=int(chunk.row_flags[r]))
    return out


@pytest.mark.parametrize("phased", [True, False])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_parse_matches_oracle(seed, phased):
    rows, ref, lo, sites = synth.hard_filter_chunk(6, seed, with_phasing=phased, depth=(20, 45)[seed % 2], read_len=((120, 900), (60, 300))[seed % 2])
    want = ho.parse_chunk(rows, phased)
    chunk = hf.parse_chunk("".join(rows).encode(), phased, ref, lo)
    got = rebuild_rows(chunk)
    assert sorted(got) == sorted(want)
    key_of = {}
    for pos, w in want.items():
        g = got[pos]
        n = len(w.toks)
        assert len(g["rid"]) == n
        assert g["tok"] == [(a + b).upper() for a, b in w.toks]
        assert g["sfx"] == [b for _, b in w.toks]
        for k in range(n):                                         # read key <-> id is one to one over the whole chunk
            assert key_of.setdefault(int(g["rid"][k]), w.names[k]) == w.names[k]
        info = g["info"]
        assert [bool(x & 4) for x in info] == [nm.endswith('1') for nm in w.names]
        assert [int(x >> 8) & 0xffff for x in info] == [len(b) for _, b in w.toks]
        assert [bool(x & 32) for x in info] == [b[:1] == '+' for _, b in w.toks]
        assert [bool(x & 64) for x in info] == ['-' in b for _, b in w.toks]
        assert [bool(x & 8) for x in info] == [a + b in ('#', '*') for a, b in w.toks]
        rb = ref[pos - lo]
        assert [bool(x & 16) for x in info] == [a + b == rb for a, b in w.toks]
        if phased:
            assert [int(x & 3) for x in info] == [int(h) if h in ('1', '2') else 0 for h in w.phasing]
        assert [int(q & 0xff) for q in g["qual"]] == w.bq[:n] and [int(q >> 8) for q in g["qual"]] == w.mq[:n]
        assert g["rse"] == sorted((k if k >= 0 else n - 1) for k in w.rse)
        assert bool(g["flags"] & 2) == (len(w.rse) >= n * ho.EPS_RSE)
        assert bool(g["flags"] & 4) == (not (len(w.counter) == 1 and w.counter[rb] > 0))
        last = {}
        for k, nm in enumerate(w.names):
            last[nm] = k
        assert [bool(x & 128) for x in info] == [last[nm] != k for k, nm in enumerate(w.names)]
    assert len(set(key_of.values())) == len(key_of) == chunk.n_reads


def test_parse_quirks_and_errors():
    ref = "ACGTACGTAC"
    # '^' before the first read marks index -1 (= the last read); two start markers against one end marker; a duplicate name
    row = "c\t3\tN\t4\t^]G$A+2acT^]c\tIIII\t]]]]\tr1,r2,r1,r4\n"
    chunk = hf.parse_chunk(row.encode(), False, ref, 1)
    g = rebuild_rows(chunk)[3]
    assert g["tok"] == ["G", "A+AC", "T", "C"] and g["sfx"] == ["", "+ac", "", ""]      # the length digits are not part of it
    assert g["rse"] == [2, 3]                                      # starts {-1, 2} -> entries 3 and 2; one '$' set is smaller
    assert [bool(x & 128) for x in g["info"]] == [True, False, False, False]       # r1_0 occurs again: the first is shadowed
    assert chunk.n_reads == 3
    want = ho.parse_chunk([row], False)[3]
    assert want.names[3] == "r4\n_1" and sorted(want.rse) == [-1, 2]
    # the row's last name keeps its line feed without the HP column, not with it
    two = "c\t3\tN\t1\tG\tI\t]\tr9\nc\t4\tN\t2\tGG\tII\t]]\tr9,r8\n"
    assert hf.parse_chunk(two.encode(), False, ref, 1).n_reads == 3
    two_hp = "c\t3\tN\t1\tG\tI\t]\tr9\t1\nc\t4\tN\t2\tGG\tII\t]]\tr9,r8\t1,*\n"
    ck = hf.parse_chunk(two_hp.encode(), True, ref, 1)
    assert ck.n_reads == 2 and [int(x & 3) for x in ck.info] == [1, 1, 0]
    assert hf.parse_chunk(b"", True, ref, 1).n_rows == 0
    assert hf.parse_chunk(b"c\t3\tN\t1\tG\tI\t]\n", False, ref, 1).n_rows == 0            # fewer than 8 columns: skipped (PV:241)
    for bad in (b"c\t4\tN\t1\tG\tI\t]\tr\nc\t3\tN\t1\tG\tI\t]\tr\n",                      # not increasing
                b"c\t3\tN\t2\tGG\tII\t]]\tr1\n",                                         # fewer names than reads
                b"c\t3\tN\t1\tG\tI\t]\tr1,r2\n",                                         # more names than reads
                b"c\t3\tN\t1\t+2ac\tI\t]\tr1\n",                                         # suffix before any read
                b"c\t3\tN\t1\tG\tI\t]\tr1\t12\n"):                                       # HP tag '12'
        with pytest.raises(_lib.CtoError):
            hf.parse_chunk(bad, bad.split(b"\n")[0].count(b"\t") >= 8, ref, 1)


def test_site_tables_resolve_strings():
    rows, ref, lo, sites = synth.hard_filter_chunk(10, 5, with_phasing=True)
    chunk = hf.parse_chunk("".join(rows).encode(), True, ref, lo)
    t, scratch = hf._site_tables(chunk, 1, sites, 100)
    assert scratch == 0 and (t["rid_span"] <= hf.SMEM_READS).all()
    want = ho.parse_chunk(rows, True)
    for s, (pos, rb, ab, af, het, hom) in enumerate(sites):
        assert chunk.row_pos[t["centre_row"][s]] == pos
        assert chunk.row_pos[t["row_lo"][s]] >= pos - 100 and chunk.row_pos[t["row_hi"][s] - 1] <= pos + 100
        row = want[pos]
        match = ho._alt_match('snp' if len(rb) == len(ab) == 1 else 'ins' if len(rb) == 1 else 'del', rb, ab)
        lo_e, hi_e = chunk.row_off[t["centre_row"][s]], chunk.row_off[t["centre_row"][s] + 1]
        if t["kind"][s] < 2:
            got = chunk.tok[lo_e:hi_e] == t["alt_tok"][s]
        else:
            inf = chunk.info[lo_e:hi_e]
            got = ((inf & 64) != 0) & (((inf >> 8) & 0xffff) == t["del_len"][s])
        assert list(got) == [bool(match(tk)) for tk in row.toks]
        rids = chunk.rid[chunk.row_off[t["row_lo"][s]]:chunk.row_off[t["row_hi"][s]]]
        assert rids.min() == t["rid_min"][s] and rids.max() == t["rid_min"][s] + t["rid_span"][s] - 1
    # germline records: the per-entry match bytes equal the oracle's predicates
    n_checked = 0
    for s, (pos, rb, ab, af, het, hom) in enumerate(sites):
        for zyg, off, idx, bit in (("het", t["het_off"], t["het_idx"], 1), ("hom", t["hom_off"], t["hom_idx"], 2)):
            entries = sorted(hf._split_germline(het if zyg == "het" else hom))
            for (gp, gab), g in zip(entries, idx[off[s]:off[s + 1]]):
                r = t["g_row"][g]
                assert chunk.row_pos[r] == gp
                carries = ho._germline_match(ref[gp - lo], gab, second=(zyg == "hom"))
                m = t["g_match"][t["g_off"][g]:t["g_off"][g + 1]]
                assert [bool(x & bit) for x in m] == [bool(carries(tk)) for tk in want[gp].toks]
                n_checked += 1
    assert n_checked > 5


# ---- the oracle against the lines / numbers the unmodified reference produced (tests/golden/make_golden.py) ---------------------
import gzip
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hard_filter")


def load_golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


def golden_rows(name):
    with gzip.open(os.path.join(GOLDEN, name + ".mpileup.gz"), "rt") as f:
        return f.readlines()


@pytest.mark.parametrize("name", ["phased_long", "phased_short", "unphased_short"])
def test_oracle_lines_equal_reference_golden(name):
    g = load_golden()[name]
    rows = ho.parse_chunk(golden_rows(name), g["phased"])
    n_fail = 0
    for key, want in g["lines"].items():
        disable, max_co = (int(x.split("=")[1]) for x in key.split(","))
        for (pos, rb, ab, af, het, hom), line in zip(g["sites"], want):
            got = ho.site_line('haplotype' if g["phased"] else 'postfilter', "chr20", pos, rb, ab, 100, rows, g["ref"], g["region_lo"],
                               het, hom, bool(disable), max_co, af)
            assert got == line
            n_fail += line.split()[2] == "False"
    assert n_fail > 10                                                       # the fixture exercises failing filters


def test_oracle_fisher_and_entropy_equal_reference_golden():
    g = load_golden()
    for (a, b, c, d), want in g["fisher"]:
        assert repr(ho.fisher_exact(a, b, c, d)) == want
    for seq, want in g["entropy"]:
        assert repr(ho.entropy_of(seq)) == want


def test_row_bisection_of_the_text_path():
    """engine._row_start_at_or_after (host side of Engine.run_sites_text): byte offset of the first mpileup row at or behind a
    position, by bisection on byte offsets of position-sorted rows."""
    import torch
    from clairs_to_b200.engine import _row_start_at_or_after
    positions = list(range(100, 200)) + list(range(250, 300)) + [100000, 100001]
    rows = [b"chr1\t%d\tN\t%d\t%s\tI\t]\n" % (p, 1 + p % 3, b"A" * (1 + 37 * (p % 5))) for p in positions]
    text = b"".join(rows)
    t = torch.frombuffer(bytearray(text), dtype=torch.uint8)
    starts = np.cumsum([0] + [len(r) for r in rows])
    for q in (0, 100, 101, 150, 199, 200, 220, 250, 299, 300, 99999, 100001, 100002):
        want = next((int(starts[i]) for i, p in enumerate(positions) if p >= q), len(text))
        assert _row_start_at_or_after(t, len(text), q) == want, q
    assert _row_start_at_or_after(t[:0], 0, 5) == 0
    one = torch.frombuffer(bytearray(rows[0][:-1]), dtype=torch.uint8)            # a single row without a line feed
    assert _row_start_at_or_after(one, one.numel(), 100) == 0 and _row_start_at_or_after(one, one.numel(), 101) == one.numel()


@pytest.mark.parametrize("phased", [True, False])
def test_threaded_parse_equals_single_threaded(phased):
    """cto_hf_parse_mt: row ranges parsed on several threads with local ids, translated when appended -> the same arrays and
    the same interned strings in the same order as one thread gives; errors name the same row."""
    rows, ref, lo, sites = synth.hard_filter_chunk(30, 77, with_phasing=phased, depth=35, read_len=(60, 400))
    text = "".join(rows).encode()
    one = hf.parse_chunk(text, phased, ref, lo, n_threads=1)
    for nt in (2, 3, 7, 64):
        many = hf.parse_chunk(text, phased, ref, lo, n_threads=nt)
        for name in ("row_pos", "row_off", "row_flags", "rse_off", "rse_ent", "rid", "tok", "sfx", "info", "qual"):
            assert np.array_equal(getattr(one, name), getattr(many, name)), (nt, name)
        assert one.tokens == many.tokens and one.suffixes == many.suffixes and one.n_reads == many.n_reads
    bad = rows[:400] + [rows[100]] + rows[400:]                               # a position out of order deep inside the text
    for nt in (1, 4):
        with pytest.raises(_lib.CtoError, match="row 401"):
            hf.parse_chunk("".join(bad).encode(), phased, ref, lo, n_threads=nt)
    cut = rows[:250] + ["\t".join(rows[250].split("\t")[:7] + ["a,b"] + (["0"] if phased else [])) + "\n"] + rows[251:]
    msgs = []
    for nt in (1, 5):
        with pytest.raises(_lib.CtoError) as e:
            hf.parse_chunk("".join(cut).encode(), phased, ref, lo, n_threads=nt)
        msgs.append(str(e.value))
    assert msgs[0] == msgs[1] and "row 251" in msgs[0]


def test_result_lines_and_no_cpu_fallback():
    """format_lines spells the reference's result line (HF:560-565 / PV:362-365) from the flag bits; without a CUDA device
    the filters refuse to run instead of falling back to anything (here: the build container has no GPU)."""
    import torch
    f = hf.O_VERDICT | hf.O_HETERO | hf.O_HOMO | hf.O_RSE | hf.O_BQ | hf.O_MQ | hf.O_CO_EXIST | hf.O_BOTH | hf.O_SB | hf.O_ENTROPY
    lines = hf.format_lines(1, "chr7", [(140753336, "A", "T")], np.array([f], np.uint32), np.array([0.123456789]))
    assert lines == ["chr7 140753336 True False True True True True True True True True 0.12346 True"]
    lines = hf.format_lines(1, "chr7", [(5,)], np.array([hf.O_PHASEABLE], np.uint32), np.array([1e-9]))
    assert lines == ["chr7 5 False True False False False False False False False False 0.0 False"]
    lines = hf.format_lines(0, "chr7", [(5,)], np.array([hf.O_VERDICT | hf.O_RSE | hf.O_CO_EXIST | hf.O_SB | hf.O_ENTROPY], np.uint32), np.array([1.0]))
    assert lines == ["chr7 5 True True True True 1.0 True"]
    assert hf._split_germline("12-A,15-ACG,12-A") == {(12, "A"), (15, "ACG")} and hf._split_germline("") == set() and hf._split_germline(None) == set()
    with pytest.raises(ValueError):
        hf._split_germline("12-A-C")
    if not torch.cuda.is_available():
        chunk = hf.parse_chunk(b"c\t3\tN\t1\tG\tI\t]\tr9\t1\n", True, "ACGTACGT", 1)
        with pytest.raises(RuntimeError, match="no CPU path"):
            hf.run_sites(chunk, 1, [(3, "A", "G", 0.5, "", "")])
