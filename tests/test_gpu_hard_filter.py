"""GPU parity of the per-site hard filters through the C ABI (``cto_hf_parse`` +
``cto_hard_filter_sites``, host mirror clairs_to_b200/hard_filters.py): result lines equal to the lines the UNMODIFIED
reference returned (tests/golden/hard_filter, generator tests/golden/make_golden.py) and to the oracle on fresh seeds;
Fisher p-values bit for bit.  Flags are boolean and the p-value is the reference's double: every comparison is exact."""

import gzip
import json
import os

import numpy as np
import pytest

from clairs_to_b200 import hard_filters as hf
from clairs_to_b200 import synth
from oracle import hard_filter_oracle as ho

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hard_filter")


def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["phased_long", "phased_short", "unphased_short"])
def test_lines_equal_reference_golden(name):
    g = golden()[name]
    with gzip.open(os.path.join(GOLDEN, name + ".mpileup.gz"), "rb") as f:
        text = f.read()
    sites = [tuple(s) for s in g["sites"]]
    fn = hf.haplotype_filter_chunk if g["phased"] else hf.postfilter_chunk
    for key, want in g["lines"].items():
        disable, max_co = (int(x.split("=")[1]) for x in key.split(","))
        got = fn("chr20", sites if g["phased"] else [s[:3] for s in sites], text, g["ref"], g["region_lo"], 100, bool(disable), max_co)
        assert got == want, [(a, b) for a, b in zip(got, want) if a != b][:3]


@pytest.mark.parametrize("phased", [True, False])
@pytest.mark.parametrize("seed", range(5))
def test_lines_equal_oracle_fuzz(seed, phased):
    rows, ref, lo, sites = synth.hard_filter_chunk(30, 300 + seed, with_phasing=phased, depth=(15, 40, 90)[seed % 3],
                                                   read_len=((120, 900), (60, 300))[seed % 2], flanking=(100, 40)[seed % 2])
    flanking = (100, 40)[seed % 2]
    mine = ho.parse_chunk(rows, phased)
    mode = 'haplotype' if phased else 'postfilter'
    want = [ho.site_line(mode, "chr20", p, rb, ab, flanking, mine, ref, lo, het, hom, False, 3, af) for p, rb, ab, af, het, hom in sites]
    fn = hf.haplotype_filter_chunk if phased else hf.postfilter_chunk
    got = fn("chr20", sites, "".join(rows).encode(), ref, lo, flanking, False, 3)
    assert got == want, [(a, b) for a, b in zip(got, want) if a != b][:3]
    assert sum(w.split()[2] == "False" for w in want) >= 3


def one_row_chunk(a0, r0, a1, r1):
    """A chunk of one pileup row at position 150 with a0 / a1 forward / reverse alt reads (T) and r0 / r1 reference reads (A)."""
    bases = "T" * a0 + "A" * r0 + "t" * a1 + "a" * r1
    n = len(bases)
    names = ",".join("q%d" % k for k in range(n))
    return ("c\t150\tN\t%d\t%s\t%s\t%s\t%s\t%s\n" % (n, bases, "I" * n, "]" * n, names, ",".join("0" * n))).encode()


def test_fisher_p_values_bit_exact():
    """HF:60-98 on the device: exact quotient (double-double) + the multiply / divide walk.  Mirror tables make `curP <= t`
    a tie that only identical arithmetic resolves identically; big tables overflow a plain double product."""
    ref = "A" * 400
    g = golden()["fisher"]
    for (a0, r0, a1, r1), want in g:
        if a0 + r0 + a1 + r1 == 0:
            continue
        chunk = hf.parse_chunk(one_row_chunk(a0, r0, a1, r1), True, ref, 1)
        flags, p, counts = hf.run_sites(chunk, 1, [(150, "A", "T", 0.5, "", "")], return_counts=True)
        assert list(counts[0][:4]) == [a0, r0, a1, r1]
        assert repr(float(p[0])) == want, ((a0, r0, a1, r1), float(p[0]), want)
        sb = not (float(want) < 0.001 or a0 == 0 or a1 == 0)
        assert bool(flags[0] & hf.O_SB) == sb


def test_entropy_flags_equal_oracle():
    rng = np.random.default_rng(9)
    seqs = ["".join("ACGT"[b] for b in rng.integers(0, 4, 233)) for _ in range(12)]
    seqs += ["A" * 233, "AC" * 116 + "A", ("ACGTTGCA" * 30)[:233], ("A" * 90 + "".join("ACGT"[b] for b in rng.integers(0, 4, 143)))]
    seqs += [s[:100] + "AAAAAAAAAAAAAAAACCCCCCCCCCCCCCCCC"[:17] + s[117:] for s in seqs[:4]]
    for ref in seqs:
        row = ("c\t101\tN\t2\tA+2GGa\tII\t]]\tq1,q2\t0,0\n").encode()
        chunk = hf.parse_chunk(row, True, ref, 1)
        flags, p = hf.run_sites(chunk, 1, [(101, ref[100], ref[100] + "GG", 0.5, "", "")])
        want = ho.sequence_entropy(ref[0:202]) >= ho.ENTROPY_THRESHOLD
        assert bool(flags[0] & hf.O_ENTROPY) == want, ref[84:117]


def test_wide_read_id_span_uses_the_global_table():
    """A read present in every row keeps the window's smallest read id small: spans beyond the 16 384-read shared-memory
    table take the global scratch table.  Same lines as the oracle."""
    rows, ref, lo, sites = synth.hard_filter_chunk(400, 41, with_phasing=True, depth=70, read_len=(30, 70), spacing=40)
    long_rows = []
    for r in rows:
        c = r.rstrip("\n").split("\t")
        c[3] = str(int(c[3]) + 1); c[4] += "A"; c[5] += "I"; c[6] += "]"; c[7] += ",ultralong"; c[8] += ",1"
        long_rows.append("\t".join(c) + "\n")
    chunk = hf.parse_chunk("".join(long_rows).encode(), True, ref, lo)
    tables, scratch = hf._site_tables(chunk, 1, sites, 100)
    assert scratch > 0 and (tables["rid_span"] > hf.SMEM_READS).sum() > 10
    flags, p = hf.run_sites(chunk, 1, sites)
    got = hf.format_lines(1, "chr20", sites, flags, p)
    mine = ho.parse_chunk(long_rows, True)
    pick = list(range(0, len(sites), 9)) + [len(sites) - 1]
    for k in pick:
        pos, rb, ab, af, het, hom = sites[k]
        assert got[k] == ho.site_line('haplotype', "chr20", pos, rb, ab, 100, mine, ref, lo, het, hom, False, 3, af)


def test_empty_and_absent_rows():
    ref = "ACGT" * 100
    chunk = hf.parse_chunk(b"", True, ref, 1)
    assert hf.haplotype_filter_chunk("c", [], b"", ref, 1) == []
    flags, p = hf.run_sites(chunk, 1, [(150, "A", "T", None, "", "120-G")])          # no pileup at all
    mine = ho.parse_chunk([], True)
    assert hf.format_lines(1, "c", [(150,)], flags, p) == [ho.site_line('haplotype', "c", 150, "A", "T", 100, mine, ref, 1, "", "120-G")]
    flags, p = hf.run_sites(hf.parse_chunk(b"", False, ref, 1), 0, [(150, "A", "T")])
    assert hf.format_lines(0, "c", [(150,)], flags, p) == [ho.site_line('postfilter', "c", 150, "A", "T", 100, ho.parse_chunk([], False), ref, 1)]
