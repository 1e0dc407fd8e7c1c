"""GPU parity of the device tokenizer (csrc/tokenize_dev.cu) against the host tokenizer + packer
(``cto_tokenize_mpileup`` + ``cto_pack_reads``, themselves pinned to the reference's ``decode_pileup_bases`` by
tests/test_abi_host.py and the create_tensor goldens): every output array byte for byte, and the encoded tensors."""

import numpy as np
import pytest
import torch

from clairs_to_b200 import _lib, synth
from clairs_to_b200.device_tokenizer import text_to_device, tokenize_text_device
from clairs_to_b200.engine import encode_pileup, packed_to_device
from clairs_to_b200.host import tokenize_mpileup
from clairs_to_b200.pileup_format import pack_stream

pytestmark = pytest.mark.gpu


def host_packed(text, ref, ref_start, cands, cut):
    tok = tokenize_mpileup(text, ref, ref_start, cands, 60, n_threads=2)
    pos = tok.row_pos
    want = (np.asarray(cands, np.int64)[:, None] - 16 + np.arange(33)[None, :]).ravel()
    idx = np.searchsorted(pos, want)
    ok = (idx < len(pos)) & (pos[np.minimum(idx, len(pos) - 1)] == want) if len(pos) else np.zeros(len(want), bool)
    tok.stream.win_pos = np.where(ok, idx, -1).astype(np.int32)
    return pack_stream(tok.stream, cut), pos


def compare(text, ref, ref_start, cands, cut, misalign=0):
    want, pos = host_packed(text, ref, ref_start, cands, cut)
    keep = torch.zeros(misalign + len(text) + 64, dtype=torch.uint8, device="cuda")
    view = keep[misalign:]
    if len(text):
        view[:len(text)] = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
    ref_dev = torch.frombuffer(bytearray(ref.encode()), dtype=torch.uint8).cuda()
    cand_dev = torch.as_tensor(np.asarray(cands, np.int64)).cuda()
    got, row_pos = tokenize_text_device(view, len(text), ref_dev, ref_start, cut, cand_dev)
    torch.cuda.synchronize()
    assert np.array_equal(row_pos.cpu().numpy(), pos.astype(np.int32))
    assert got.n_groups == want.n_groups
    assert np.array_equal(got.grp_off.cpu().numpy(), want.grp_off)
    assert np.array_equal(got.ind_off.cpu().numpy(), want.ind_off)
    assert np.array_equal(got.ref_code.cpu().numpy(), want.ref_code)
    assert np.array_equal(got.planes.cpu().numpy()[:want.n_groups * 8], np.asarray(want.planes)[:want.n_groups * 8])
    assert np.array_equal(got.ind_entry.cpu().numpy().view(np.uint32), np.asarray(want.ind_entry, np.uint32))
    assert np.array_equal(got.win_pos.cpu().numpy(), want.win_pos)
    if len(cands):
        ta, da = encode_pileup(got)
        tb, db = encode_pileup(packed_to_device(want, "cuda"))
        assert torch.equal(ta, tb) and torch.equal(da, db)
    return want


@pytest.mark.parametrize("platform,cut", [("ont", 30), ("ont", 10), ("ilmn", 10), ("hifi", 10)])
@pytest.mark.parametrize("misalign", [0, 7])
def test_synthetic_sites_equal_host_tokenizer(platform, cut, misalign):
    (aff, aff_aux), (neg, neg_aux) = synth.synth_pair(300, 11, platform, depth_mean=45, depth_hi=150)
    for s, aux in ((aff, aff_aux), (neg, neg_aux)):
        text = synth.render_mpileup_text(s, aux)
        ref = ''.join("ACGT"[c] for c in s.ref_code)
        cands = np.arange(1001 + 16, 1001 + s.n_rows, 33, dtype=np.int64)
        want = compare(text, ref, 1001, cands, cut, misalign)
        assert len(want.ind_entry) > 50


def test_weird_rows_gaps_and_short_quality_columns():
    """The tokenizer's corner cases (synth.scan_rows_text(weird=...): `^` + any character, `$`, `<` `>` reference skips that
    shift the quality zip, `+0`, two suffixes on one read, N reads carrying indels), positions with no pileup row inside
    windows, IUPAC reference bases, quality columns shorter than the read list, extra columns, CRLF."""
    rows, reference = synth.scan_rows_text(3000, 21, first_pos=501, depth_mean=30, weird=0.03)
    rows = [r for k, r in enumerate(rows) if k % 97 != 5]                      # gaps: windows with absent rows
    out = []
    rng = np.random.default_rng(8)
    for k, r in enumerate(rows):
        c = r.rstrip("\n").split("\t")
        c.append("".join(chr(33 + int(q)) for q in rng.integers(0, 61, len(c[5]))))     # the STEP-1 rows have no MQ column
        if k % 50 == 3:
            c[5] = c[5][:len(c[5]) // 2]                                      # short BQ column
        if k % 50 == 7:
            c[6] = c[6][:max(0, len(c[6]) - 3)]                               # short MQ column
        if k % 50 == 11:
            c.append("extra,column")                                          # an eighth column is ignored
        if k % 50 == 13:
            c[6] += " \r"                                                     # row.strip()
        out.append("\t".join(c) + "\n")
    text = "".join(out).encode()
    cands = np.arange(501 + 20, 501 + 2900, 41, dtype=np.int64)
    for cut in (30, 10):
        compare(text, reference, 501, cands, cut)
    compare(text[:-1], reference, 501, cands, 10, misalign=3)                   # no trailing newline


def test_deep_rows_and_allele_table_overflow():
    rng = np.random.default_rng(3)
    rows, ref = [], []
    for r in range(120):
        ref.append("ACGT"[r % 4])
        n_alleles = (1, 20, 24, 25, 40, 300)[r % 6]
        depth = 40 if r % 40 else 5000
        parts = []
        for k in range(depth):
            sym = "ACGTacgt*#Nn"[int(rng.integers(0, 12))]
            if k % 3 == 0:
                a = int(rng.integers(0, n_alleles))
                seq = "".join("ACGT"[(a >> (2 * z)) & 3] for z in range(1 + a % 7 + (60 if a % 11 == 0 else 0)))
                sym += ("+%d%s" if a & 1 else "-%d%s") % (len(seq), seq if sym.isupper() or sym == '*' else seq.lower())
            parts.append(sym)
        rows.append("chr1\t%d\tN\t%d\t%s\t%s\t%s\n" % (1001 + r, depth, "".join(parts), "".join(chr(33 + int(q)) for q in rng.integers(0, 60, depth)),
                                                      "".join(chr(33 + int(q)) for q in rng.integers(0, 61, depth))))
    compare("".join(rows).encode(), "".join(ref), 1001, np.array([1001 + 16, 1001 + 60, 1001 + 100], np.int64), 10)


def test_errors_and_empty_text():
    ref_dev = torch.frombuffer(bytearray(b"ACGT" * 10), dtype=torch.uint8).cuda()
    buf, n = text_to_device(b"", "cuda")
    ps, pos = tokenize_text_device(buf, 0, ref_dev, 1, 10, torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert ps.n_groups == 0 and pos.numel() == 0
    for bad in (b"chr1\t5\tN\t1\tA\tI\n", b"chr1\t500\tN\t1\tA\tI\t]\n", b"chr1\t5\tN\t1\tA\tI\t]\n\nchr1\t6\tN\t1\tA\tI\t]\n"):
        buf, n = text_to_device(bad, "cuda")
        with pytest.raises(_lib.CtoError):
            tokenize_text_device(buf, n, ref_dev, 1, 10)


def test_run_sites_text_equals_host_tokenized_call():
    """Engine.run_sites_text (text copied in pieces, tokenized on the device) against Engine.run_sites_host on the host-tokenized,
    host-packed streams of the same sites: probabilities, posteriors, calls, QUAL and FILTER bit-identical, for one piece and
    for pieces whose row spans overlap."""
    from clairs_to_b200 import synth_weights as sw
    from clairs_to_b200.engine import Engine
    aff_sd = sw.synth_state_dict(sw.aff_state_dict_shapes(4), 104)
    neg_sd = sw.synth_state_dict(sw.neg_state_dict_shapes(4), 204)
    rng = np.random.default_rng(3)
    likelihood = np.concatenate([rng.uniform(0.05, 0.95, size=(40, 10)), np.sort(rng.uniform(0.02, 0.98, size=(8, 10)), axis=1)])
    eng = Engine(aff_sd, neg_sd, max_batch=256, likelihood=likelihood)
    (aff, aff_aux), (neg, neg_aux) = synth.synth_pair(700, 31, 'ont')
    texts = [synth.render_mpileup_text(s, a) for s, a in ((aff, aff_aux), (neg, neg_aux))]
    ref = ''.join("ACGT"[c] for c in neg.ref_code)
    cands = np.arange(1001 + 16, 1001 + neg.n_rows, 33, dtype=np.int64)
    packed = [host_packed(t, ref, 1001, cands, 30)[0] for t in texts]
    want = eng.run_sites_host(packed[0], packed[1])
    pinned = [torch.frombuffer(bytearray(t), dtype=torch.uint8).pin_memory() for t in texts]
    for pieces in (1, 3, 7, None):
        got = eng.run_sites_text(pinned[0], pinned[1], ref.encode(), 1001, cands, 30, pieces=pieces)
        for k in ("probs", "post", "call", "qual", "filter"):
            assert torch.equal(got[k], want[k]), (pieces, k)
    one = eng.run_sites_text(pinned[1], None, ref.encode(), 1001, cands, 30, pieces=2)     # one stream feeds both networks
    assert torch.equal(one["probs"], eng.run_sites_host(packed[1], None)["probs"])
    eng.close()
