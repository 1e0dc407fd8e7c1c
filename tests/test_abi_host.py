"""CPU: the C-ABI library loads, exports every declared symbol, and its HOST-side entry points
(mpileup tokenizer, chunk-file text codec) agree with the oracle.  No compute call needs a GPU."""

import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from clairs_to_b200 import _lib, synth
from clairs_to_b200.host import tokenize_mpileup, format_tensor_rows, parse_tensor_row, format_prob_fields
from clairs_to_b200.pileup_format import HAS_INDEL, SYMBOLS
from oracle import pileup_oracle, posterior_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "clairs_to_b200.h")).read()
    declared = set(re.findall(r"\b(cto_[a-z0-9_]+)\s*\(", header))
    declared -= {"cto_engine", "cto_tokens", "cto_host_stream"}
    assert len(declared) >= 20
    handle = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), "missing export " + name
    assert set(_lib.SIGNATURES) == declared
    assert _lib.lib().cto_abi_version() == 4


def test_tokenizer_matches_oracle_on_golden_rows(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "encoder_golden.json")))
    ref_window = "ACGT"
    for c in cases:
        if c["platform"] != "ont":          # tokenisation is platform independent; one pass is enough
            continue
        chunk_ref = c["chunk_ref"] if c["chunk_ref"] else c["ref"]
        # a reference window whose first base is the row's reference base and continues with chunk_ref
        ref_seq = (chunk_ref if chunk_ref[0] == c["ref"] else c["ref"] + chunk_ref[1:]) + "A" * 64
        row = "chr1\t100\tN\t0\t%s\t%s\t%s\n" % (c["bases"], c["bq"], c["mq"])
        tok = tokenize_mpileup(row, ref_seq, 100, [100] if c["candidate"] else [], 60)
        entries = pileup_oracle.tokenize(c["bases"])
        assert tok.stream.n_reads == len(entries)
        for i, (sym, indel) in enumerate(entries):
            code = int(tok.stream.code[i])
            assert SYMBOLS[code & 0xF] == sym
            assert bool(code & HAS_INDEL) == bool(indel)
        if c["candidate"] and chunk_ref[0] == c["ref"]:
            assert tok.alt_info[0] == c["alt_info"], c["bases"]


def test_tokenizer_multi_row_layout():
    stream, aux = synth.synth_stream(3, 21, 'ont', depth_lo=0, depth_hi=40, depth_mean=12)
    rows = synth.render_mpileup(stream, aux, decorate_seed=3)
    ref_seq = ''.join("ACGT"[int(r)] for r in stream.ref_code)
    tok = tokenize_mpileup(''.join(rows), ref_seq, 1001, [1001 + 16, 1001 + 49], 60)
    s = tok.stream
    assert s.n_rows == stream.n_rows
    assert np.array_equal(tok.row_pos, 1001 + np.arange(stream.n_rows))
    # rows that were empty in the generator come back as samtools' single '*' placeholder read
    empty = np.diff(stream.pos_off) == 0
    expect_depth = np.where(empty, 1, np.diff(stream.pos_off))
    assert np.array_equal(np.diff(s.pos_off), expect_depth)
    assert np.array_equal(s.ref_code, stream.ref_code)
    keep = np.repeat(~empty, expect_depth)
    assert np.array_equal(s.code[keep], stream.code)
    assert np.array_equal(s.mq[keep], stream.mq)
    assert np.array_equal(s.bq[keep], stream.bq)
    assert np.array_equal(np.diff(s.ind_off), np.diff(stream.ind_off))
    # allele ids may be numbered differently; everything else in the sparse entries is identical
    assert np.array_equal(s.ind_entry & 0xFFFF0000, stream.ind_entry & 0xFFFF0000)
    assert [bool(a) for a in tok.alt_info] == [(p in (1017, 1050)) for p in tok.row_pos]


def test_tensor_text_codec_roundtrip():
    rng = np.random.default_rng(3)
    t = rng.integers(-300, 300, size=(5, 33, 34)).astype(np.int16)
    t[0, 0, 0] = -32768
    t[0, 0, 1] = 32767
    texts = format_tensor_rows(t)
    for k in range(5):
        expect = " ".join(" ".join("%d" % x for x in row) for row in t[k])      # CT:551
        assert texts[k] == expect
        assert np.array_equal(parse_tensor_row(texts[k]), t[k])
    with pytest.raises(_lib.CtoError):
        parse_tensor_row("1 2 3")


def test_prob_field_format_matches_python():
    rng = np.random.default_rng(4)
    p = rng.random((6, 8, 2)).astype(np.float32)
    p[0, 0] = [0.0, 1.0]
    p[0, 1] = [np.float32(0.123456785), np.float32(1e-9)]
    for k in range(6):
        fields = format_prob_fields(p[k]).split("\t")
        expect = posterior_oracle.format_predict_row("c", 1, "A", "x", [0.0], [0.0], list(p[k])).split("\t")[6:14]
        assert fields == expect


def _python_rows(text):
    """the row-wise Python path the whole-file codec replaces (clairs_to_b200/predict.py:read_tensor_rows)"""
    from clairs_to_b200 import host
    rows = []
    for row in text.decode().splitlines(True):
        cols = row.split("\t")[:7]
        if len(cols) < 7 or cols[2][16] not in "ACGT":
            continue
        rows.append((cols[0], cols[1], cols[2], host.parse_tensor_row(cols[3]), cols[4], cols[5], cols[6].strip()))
    return rows


def _golden_tensor_texts():
    import glob
    import gzip
    here = os.path.dirname(os.path.abspath(__file__))
    for path in sorted(glob.glob(os.path.join(here, "golden", "pipeline", "tensor_can_*")) +
                       glob.glob(os.path.join(here, "golden", "create_tensor", "tensor_can_*"))):
        raw = open(path, "rb").read()
        yield path, (gzip.decompress(raw) if raw[:2] == b"\x1f\x8b" else raw)


def test_parse_tensor_file_matches_rowwise_parser_on_golden_files():
    from clairs_to_b200 import host
    seen = 0
    for path, text in _golden_tensor_texts():
        # add the cases the reference reader filters: a short row, a row with a non-ACGT centre base, CRLF
        lines = text.split(b"\n")
        extra = b"chrX\t5\tshort row\n"
        first = lines[0].split(b"\t")
        bad = b"\t".join([first[0], first[1], first[2][:16] + b"N" + first[2][17:]] + first[3:]) + b"\n"
        text2 = extra + bad + text.replace(b"\n", b"\r\n", 1)
        want = _python_rows(text2)
        tf = host.TensorFile(text2)
        assert tf.n == len(want) == len(_python_rows(text)), path
        for r, w in enumerate(want):
            assert (tf.field(r, 0), tf.field(r, 1), tf.field(r, 2), tf.field(r, 4), tf.field(r, 5), tf.field(r, 6)) == \
                   (w[0], w[1], w[2], w[4], w[5], w[6])
            assert np.array_equal(tf.tensor[r], w[3])
            assert tf.depth[r] == int(float(w[4].split('-')[0]))
        seen += tf.n
    assert seen > 0


@pytest.mark.parametrize("n_heads", [4, 6])
def test_format_predict_rows_matches_python_formatter(n_heads):
    from clairs_to_b200 import host
    from clairs_to_b200.predict import format_rows
    path, text = next(_golden_tensor_texts())
    tf = host.TensorFile(text)
    n = tf.n
    rng = np.random.default_rng(n_heads)
    fwd = rng.integers(0, 200, size=(n, 4)).astype(np.int32)
    rev = rng.integers(0, 200, size=(n, 4)).astype(np.int32)
    probs = rng.random(size=(n, 2 * n_heads, 2)).astype(np.float32)
    probs[0, 0] = (1.0, 0.0)
    probs[0, 1] = (0.999999996, 4e-9)                      # rounds to 1.00000000 / 0.00000000 at 8 decimals
    meta = [tuple(tf.field(r, k) if k != 3 else None for k in range(7)) for r in range(n)]
    want = "".join(format_rows(meta, fwd, rev, probs, n_heads)).encode()
    assert host.format_predict_rows(tf, 0, n, fwd, rev, probs, n_heads) == want
    half = n // 2
    assert host.format_predict_rows(tf, half, n - half, fwd[half:], rev[half:], probs[half:], n_heads) == \
        "".join(format_rows(meta[half:], fwd[half:], rev[half:], probs[half:], n_heads)).encode()


def test_format_tensor_can_rows_matches_python_formatting():
    from clairs_to_b200 import host
    rng = np.random.default_rng(3)
    n = 57
    tensors = rng.integers(-300, 300, size=(n, 33, 34)).astype(np.int16)
    tensors[0, 0, 0], tensors[0, 0, 1] = -32768, 32767
    pos = np.sort(rng.integers(1, 2_000_000_000, size=n))
    ref33 = ["".join(rng.choice(list("ACGTN"), size=33)) for _ in range(n)]
    alts = ["%d-X%s %d R%s %d-" % (rng.integers(1, 99), "ACGT"[k % 4], k, "ACGT"[(k + 1) % 4], 2 * k) for k in range(n)]
    alts[3] = ""
    types = [("snv", "indel", "unknown")[k % 3] for k in range(n)]
    flat = format_tensor_rows(tensors)
    want = "".join("%s\t%d\t%s\t%s\t%s\t%s\t%s\n" % ("chr20", pos[k], ref33[k], flat[k], alts[k], types[k], ref33[k][16])
                   for k in range(n)).encode()
    assert host.format_tensor_can_rows("chr20", pos, ref33, tensors, alts, types) == want
    assert host.format_tensor_can_rows("chr20", pos[:0], [], tensors[:0], [], []) == b""


def test_parse_tensor_file_edge_cases():
    from clairs_to_b200 import host
    assert host.TensorFile(b"").n == 0
    assert host.TensorFile(b"\n\n").n == 0
    ints = " ".join(str(v) for v in range(-561, 561))
    row = "chr1\t77\t%s\t%s\t12-XA 2-\tsnv\tA" % ("ACGT" * 8 + "A", ints)
    # no trailing newline, extra columns after the seventh, a lower-case centre base (dropped: not in "ACGT")
    lower = row.replace("ACGT" * 8 + "A", "ACGT" * 4 + "a" + "CGT" + "ACGT" * 3 + "A")
    tf = host.TensorFile((row + "\textra\tcols\n" + lower + "\n" + row).encode())
    assert tf.n == 2
    assert tf.field(0, 6) == "A" and tf.field(1, 6) == "A" and tf.field(0, 1) == "77"
    assert tf.tensor[1].reshape(-1).tolist() == list(range(-561, 561))
    assert tf.depth.tolist() == [12, 12]
    with pytest.raises(_lib.CtoError):                       # a tensor field that is too short
        host.TensorFile(("chr1\t77\t%s\t1 2 3\t12-\tsnv\tA\n" % ("ACGT" * 8 + "A")).encode())


def test_bit_plane_packer_matches_numpy_statement():
    """cto_pack_reads (native, multi-threaded) == the numpy statement of the packed layout, for both low-BQ literals."""
    from clairs_to_b200 import synth
    from clairs_to_b200.pileup_format import pack_stream, pack_stream_numpy
    (aff, _), (neg, _) = synth.synth_pair(700, 3, 'ont', depth_lo=0, depth_hi=150)
    for s in (aff, neg):
        for cut in (10, 30):
            for threads in (1, 3):
                a, b = pack_stream(s, cut, n_threads=threads), pack_stream_numpy(s, cut)
                assert a.n_groups == b.n_groups and np.array_equal(a.grp_off, b.grp_off)
                assert np.array_equal(a.planes[:8 * a.n_groups], b.planes[:8 * b.n_groups])
                assert not a.planes[8 * a.n_groups:].any()            # the 16-byte padding is zero-filled
    # rows are padded to whole groups with NULL reads: one byte per read + at most 7 per row
    assert 8 * a.n_groups <= neg.n_reads + 7 * neg.n_rows


def test_native_renderer_and_threaded_tokenizer_round_trip():
    """cto_render_mpileup == the python renderer; the tokenizer gives the same arrays with 1 or 4 threads and
    reproduces the generated read arrays."""
    from clairs_to_b200 import synth
    from clairs_to_b200.host import tokenize_mpileup
    (aff, aa), (neg, na) = synth.synth_pair(400, 11, 'ont', depth_lo=0, depth_hi=120)
    for s, a in ((aff, aa), (neg, na)):
        text = synth.render_mpileup_text(s, a)
        assert text == ''.join(synth.render_mpileup(s, a)).encode()
        ref = ''.join("ACGT"[c] for c in s.ref_code)
        cands = list(range(1001 + 16, 1001 + s.n_rows, 33))
        t1 = tokenize_mpileup(text, ref, 1001, cands, 60, n_threads=1)
        t4 = tokenize_mpileup(text * 1, ref, 1001, cands, 60, n_threads=4)
        for name in ("code", "bq", "mq", "pos_off", "ref_code", "ind_off", "ind_entry"):
            assert np.array_equal(getattr(t1.stream, name), getattr(t4.stream, name)), name
        assert t1.alt_info == t4.alt_info and np.array_equal(t1.row_pos, t4.row_pos)
        assert np.array_equal(t1.stream.code, s.code) and np.array_equal(t1.stream.bq, s.bq) and np.array_equal(t1.stream.mq, s.mq)
        assert np.array_equal(t1.stream.pos_off, s.pos_off) and np.array_equal(t1.stream.ind_off, s.ind_off)


def test_predict_file_parser_matches_python_row_parse(golden_dir):
    """cto_parse_predict_file == the reference's row loop (clairs/call_variants.py:798-829: split on tabs, float() every
    probability) on the predict files the unmodified reference wrote."""
    import gzip
    from clairs_to_b200.host import PredictFile
    for tag, n_heads in (("snv", 4), ("indel", 6)):
        text = gzip.open(os.path.join(golden_dir, "pipeline", "predict_" + tag), "rb").read()
        pf = PredictFile(text, n_heads)
        rows = [r.rstrip().split("\t") for r in text.decode().splitlines() if r.strip()]
        assert pf.n == len(rows) == 39
        for k, cols in enumerate(rows):
            assert [pf.field(k, f) for f in range(6)] == cols[:6]
            probs = [[float(v) for v in f.split()] for f in cols[6:6 + 2 * n_heads]]
            assert pf.p_aff[k].tolist() == [p[1] for p in probs[:n_heads]]
            assert pf.p_neg[k].tolist() == [p[1] for p in probs[n_heads:]]
    with pytest.raises(Exception):
        PredictFile(b"chr1\t5\tA\t3-XT 1-\t[0.0]\t[0.0]\t0.5 0.5\n", 4)          # too few fields
