"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Bars: integer pileup tensor, depths, strand counts and the fp64 posterior (given identical
probabilities) are BIT-EXACT; AFF/NEG logits, probabilities and end-to-end posteriors are within
1e-3 absolute (BASELINE.json north_star)."""

import gzip
import json
import os

import numpy as np
import pytest
import torch

from clairs_to_b200 import synth
from clairs_to_b200.host import tokenize_mpileup
from clairs_to_b200.pileup_format import N_CH, N_POS, PileupStream
from oracle import nn_oracle, pileup_oracle, posterior_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _engine(n_heads=4, max_batch=4096, gain=1.0, likelihood=None, tensor_cores=True):
    from clairs_to_b200.engine import Engine
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads), 100 + n_heads, gain)
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 200 + n_heads, gain)
    eng = Engine(aff_sd, neg_sd, max_batch=max_batch, likelihood=likelihood)
    eng.set_tensor_cores(tensor_cores)
    return eng, aff_sd, neg_sd


@pytest.fixture(scope="module")
def eng4():
    e, a, n = _engine(4)
    yield e, a, n
    e.close()


def _oracle_tensor(stream, aux, platform_literal):
    """Render the stream to mpileup text and push every row through the oracle."""
    rows = synth.render_mpileup(stream, aux, decorate_seed=1)
    vecs, depths = [], []
    for i, text in enumerate(rows):
        pos, bases, bq, mq = pileup_oracle.parse_mpileup_row(text)
        ref = "ACGT"[int(stream.ref_code[i])]
        vec, alt = pileup_oracle.position_vector(bases, mq, bq, ref, is_candidate=True, chunk_ref_seq=ref,
                                                 platform=platform_literal)
        vecs.append(vec)
        depths.append(int(alt.split('-')[0]))
    t = np.array(vecs, dtype=np.int16).reshape(-1, N_POS, N_CH)
    d = np.array(depths, dtype=np.int32).reshape(-1, N_POS)[:, N_POS // 2]
    return t, d


@pytest.mark.parametrize("platform,literal", [("ont", "ont"), ("ont", "ont_r10_dorado_sup_5khz"),
                                              ("ilmn", "ilmn_ssrs"), ("hifi", "hifi_revio")])
def test_encoder_bit_exact_vs_oracle(eng4, platform, literal):
    from clairs_to_b200.engine import low_bq_cut_for, stream_to_device
    eng = eng4[0]
    (aff, aff_aux), (neg, neg_aux) = synth.synth_pair(20, 31, platform, depth_lo=0, depth_hi=150)
    for stream, aux in ((aff, aff_aux), (neg, neg_aux)):
        want_t, want_d = _oracle_tensor(stream, aux, literal)
        got_t, got_d = eng.encode(stream_to_device(stream, eng.device), low_bq_cut_for(literal))
        assert np.array_equal(got_t.cpu().numpy(), want_t)
        assert np.array_equal(got_d.cpu().numpy(), want_d)


def test_encoder_golden_rows_through_tokenizer(eng4, golden_dir):
    """reference decode_pileup_bases outputs (golden) == tokenizer (host C++) + encoder (CUDA)."""
    from clairs_to_b200.engine import low_bq_cut_for, stream_to_device
    eng = eng4[0]
    cases = json.load(open(os.path.join(golden_dir, "encoder_golden.json")))
    for literal in sorted({c["platform"] for c in cases}):
        group = [c for c in cases if c["platform"] == literal]
        text = ''.join("chr1\t%d\tN\t0\t%s\t%s\t%s\n" % (100 + i, c["bases"], c["bq"], c["mq"]) for i, c in enumerate(group))
        ref_seq = ''.join(c["ref"] for c in group)
        tok = tokenize_mpileup(text, ref_seq, 100, [], 60)
        s = tok.stream
        assert s.n_rows == len(group)
        win = np.full((len(group), N_POS), -1, dtype=np.int32)
        win[:, N_POS // 2] = np.arange(len(group))
        s.win_pos = win.reshape(-1)
        got, depth = eng.encode(stream_to_device(s, eng.device), low_bq_cut_for(literal))
        got = got.cpu().numpy()
        want = np.array([c["vec"] for c in group], dtype=np.int16)
        assert np.array_equal(got[:, N_POS // 2, :], want)
        assert not got[:, :N_POS // 2].any() and not got[:, N_POS // 2 + 1:].any()     # absent rows are zero rows
        for c, d in zip(group, depth.cpu().numpy()):
            if c["alt_info"] is not None:
                assert int(c["alt_info"].split('-')[0]) == d


def test_encoder_edge_cases(eng4):
    from clairs_to_b200.engine import stream_to_device
    eng = eng4[0]
    # empty batch
    empty = PileupStream(*(np.zeros(0, dt) for dt in (np.uint8, np.uint8, np.uint8)), np.zeros(1, np.int32),
                         np.zeros(0, np.uint8), np.zeros(1, np.int32), np.zeros(0, np.uint32), np.zeros(0, np.int32))
    t, d = eng.encode(stream_to_device(empty, eng.device), 10)
    assert t.shape == (0, N_POS, N_CH) and d.shape == (0,)
    # one very deep row with many distinct indel alleles (exercises the multi-chunk allele scan)
    rng = np.random.default_rng(9)
    reads = []
    for k in range(3000):
        sym = "ACGTacgt"[rng.integers(0, 8)]
        r = rng.random()
        if r < 0.2:
            L = int(rng.integers(1, 5))
            seq = ''.join("ACGT"[rng.integers(0, 2)] for _ in range(L))
            reads.append(sym + "+%d%s" % (L, seq if sym.isupper() else seq.lower()))
        elif r < 0.35:
            L = int(rng.integers(1, 4))
            reads.append(sym + "-%d%s" % (L, ('N' if sym.isupper() else 'n') * L))
        else:
            reads.append(sym)
    mq = ''.join(chr(33 + int(v)) for v in rng.choice([60, 60, 60, 5], size=3000))
    bq = ''.join(chr(33 + int(v)) for v in rng.integers(1, 50, size=3000))
    text = "chr1\t500\tN\t3000\t%s\t%s\t%s\n" % (''.join(reads), bq, mq)
    tok = tokenize_mpileup(text, "G" * 60, 500, [500], 60)
    s = tok.stream
    win = np.full(N_POS, -1, dtype=np.int32)
    win[N_POS // 2] = 0
    s.win_pos = win
    got, depth = eng.encode(stream_to_device(s, eng.device), 10)
    want, alt = pileup_oracle.position_vector(''.join(reads), [ord(c) - 33 for c in mq], [ord(c) - 33 for c in bq], "G",
                                              is_candidate=True, chunk_ref_seq="G" * 60, platform="x")
    assert got.cpu().numpy()[0, N_POS // 2].tolist() == want
    assert int(depth.cpu()[0]) == int(alt.split('-')[0])
    assert tok.alt_info[0] == alt


def test_encoder_unaligned_plane_array_takes_the_unstaged_path(eng4):
    """an offset view of the read array (not 16-byte aligned) must not fault or read out of bounds: the launch
    detects it and reads with ordinary loads instead of 16-byte bulk copies.  Same tensor either way."""
    import ctypes as C
    from clairs_to_b200 import _lib
    from clairs_to_b200.engine import packed_to_device
    from clairs_to_b200.pileup_format import pack_stream
    eng = eng4[0]
    (aff, _), (neg, _) = synth.synth_pair(37, 19, 'ont', depth_lo=0, depth_hi=90)
    ps = packed_to_device(pack_stream(neg, 30), eng.device)
    want, want_d = eng.encode(ps)
    shifted = torch.zeros(ps.planes.numel() + 8, dtype=torch.uint8, device=eng.device)
    shifted[8:] = ps.planes                                       # 8-byte aligned, not 16
    view = shifted[8:]
    assert view.data_ptr() % 16 == 8
    t = torch.empty_like(want)
    d = torch.empty_like(want_d)
    p = lambda x: C.c_void_p(x.data_ptr())
    _lib.check(eng.lib.cto_encode_pileup(p(view), p(ps.grp_off), p(ps.ref_code), p(ps.ind_off), p(ps.ind_entry), p(ps.win_pos),
                                         want.shape[0], ps.n_groups, p(t), p(d), None), "encode")
    torch.cuda.synchronize()
    assert torch.equal(t, want) and torch.equal(d, want_d)


@pytest.mark.parametrize("tensor_cores,tol", [(True, TOL), (False, 5e-5)])
@pytest.mark.parametrize("n_heads", [4, 6])
def test_forward_matches_reference_golden(golden_dir, n_heads, tensor_cores, tol):
    """Reference logits (golden) vs the engine: bf16x3 tensor-core path within the 1e-3 contract,
    fp32 CUDA-core path at fp32 rounding level."""
    eng, _, _ = _engine(n_heads, max_batch=16, tensor_cores=tensor_cores)   # 24 candidates -> two internal chunks
    g = np.load(os.path.join(golden_dir, "nn_golden.npz"))
    x = torch.from_numpy(g["x_%d" % n_heads])
    la = eng.forward_aff(x).cpu().numpy()
    ln = eng.forward_neg(x).cpu().numpy()
    err_a = np.abs(la - g["aff_logits_%d" % n_heads]).max()
    err_n = np.abs(ln - g["neg_logits_%d" % n_heads]).max()
    print("tensor_cores=%s: max |logit err| vs reference: AFF %.3g NEG %.3g" % (tensor_cores, err_a, err_n))
    assert err_a < tol and err_n < tol
    eng.close()


@pytest.mark.parametrize("tensor_cores", [True, False])
@pytest.mark.parametrize("gain", [0.5, 1.0, 1.5])
def test_forward_vs_oracle_weight_scales(gain, tensor_cores):
    """Trained weights are unavailable offline: hold the contract across weight scales.
    * logits: 1e-3 ABSOLUTE for gain <= 1 (|logit| up to ~7, the regime the contract is stated for);
    * class probabilities (what predict writes and the posterior consumes): 1e-3 absolute at EVERY gain;
    * gain 1.5 drives |logit| past 30 (post-SELU, far outside any trained model), where fp32 itself is no longer good to
      1e-3: the fp32 torch oracle differs from its own fp64 evaluation by > 1e-4 there (asserted below, so this branch
      cannot silently widen), and the logit bar becomes 1e-3 per 10 units of |logit| (a 1e-4 relative bar)."""
    eng, aff_sd, neg_sd = _engine(4, max_batch=128, gain=gain, tensor_cores=tensor_cores)
    (aff, _), (neg, _) = synth.synth_pair(300, 77, 'ont', depth_mean=70, depth_hi=200)
    from clairs_to_b200.engine import stream_to_device
    xa, da = eng.encode(stream_to_device(aff, eng.device), 10)
    xn, dn = eng.encode(stream_to_device(neg, eng.device), 10)
    fa = eng.rescale(xa, da)
    fn = eng.rescale(xn, dn)
    # rescale is bit-exact against python-double arithmetic (predict.py:179-197)
    want = np.stack([posterior_oracle.rescale_tensor(t, d) for t, d in zip(xa.cpu().numpy(), da.cpu().numpy())])
    assert np.array_equal(fa.cpu().numpy(), want)
    la = eng.forward_aff(fa).cpu().numpy()
    ln = eng.forward_neg(fn).cpu().numpy()
    oa = nn_oracle.aff_forward(fa.cpu().numpy(), aff_sd).numpy()
    on = nn_oracle.neg_forward(fn.cpu().numpy(), neg_sd).numpy()
    err_a, err_n = np.abs(la - oa).max(), np.abs(ln - on).max()
    print("gain %.1f tc=%s: max |logit err| AFF %.3g (max|logit| %.2f) NEG %.3g (max|logit| %.2f)"
          % (gain, tensor_cores, err_a, np.abs(oa).max(), err_n, np.abs(on).max()))
    sm = lambda z: nn_oracle.softmax_heads(z).numpy()
    assert np.abs(sm(la) - sm(oa)).max() < TOL and np.abs(sm(ln) - sm(on)).max() < TOL
    assert err_n < TOL
    if gain <= 1.0:
        assert err_a < TOL
    else:
        o64 = nn_oracle.aff_forward(fa.cpu().numpy(), aff_sd, dtype=torch.float64).numpy()
        assert np.abs(oa).max() > 20 and np.abs(oa - o64).max() > 1e-4      # fp32 itself is off by > 1e-4 in this regime
        assert err_a < TOL * np.abs(oa).max() / 10
    eng.close()


@pytest.mark.parametrize("tag,n_heads", [("snv", 4), ("indel", 6)])
def test_predict_posterior_against_reference_files(golden_dir, tag, n_heads):
    """tensor_can chunk files -> CUDA predict -> probabilities vs the reference's predict file;
    posterior kernel bit-exact vs the oracle on identical probabilities."""
    pdir = os.path.join(golden_dir, "pipeline")
    eng, _, _ = _engine(n_heads, max_batch=64, likelihood=os.path.join(pdir, "likelihood_%s.txt" % tag))

    def load(path):
        rows = [r.rstrip("\n").split("\t") for r in gzip.open(path, "rt")]
        rows = [r for r in rows if r[2][16] in "ACGT"]
        t = np.array([[int(v) for v in r[3].split()] for r in rows], dtype=np.int16).reshape(-1, N_POS, N_CH)
        d = np.array([int(r[4].split('-')[0]) for r in rows], dtype=np.int32)
        return rows, torch.from_numpy(t).cuda(), torch.from_numpy(d).cuda()

    rows, xa, da = load(os.path.join(pdir, "tensor_can_aff_" + tag))
    _, xn, dn = load(os.path.join(pdir, "tensor_can_neg_" + tag))
    out = eng.predict(xa, da, xn, dn)
    ref_rows = [r.rstrip("\n").split("\t") for r in gzip.open(os.path.join(pdir, "predict_" + tag), "rt")]
    probs = out['probs'].cpu().numpy()
    ref_probs = np.array([[[float(v) for v in f.split()] for f in r[6:6 + 2 * n_heads]] for r in ref_rows])
    assert np.abs(probs - ref_probs).max() < TOL
    # strand counts: exact ("[0.0, 1.0, 0.0, 35.0]" list reprs in the predict file)
    fwd = out['fwd'].cpu().numpy().astype(np.float64).tolist()
    rev = out['rev'].cpu().numpy().astype(np.float64).tolist()
    for k, r in enumerate(ref_rows):
        assert str(fwd[k]) == r[4] and str(rev[k]) == r[5]
    # posterior: identical inputs (the kernel's own probabilities, 8-decimal round trip) -> identical doubles
    mats, ea, en = posterior_oracle.load_likelihood(os.path.join(pdir, "likelihood_%s.txt" % tag), n_heads)
    post = out['post'].cpu().numpy()
    call = out['call'].cpu().numpy()
    for k in range(len(rows)):
        p8 = [float("{:0.8f}".format(v)) for v in probs[k, :, 1]]
        want = posterior_oracle.posterior(p8[:n_heads], p8[n_heads:], mats, ea, en)
        assert np.array_equal(post[k], want), (k, post[k], want)
        assert (call[k] & 0xFF) == int(np.argmax(want)) and (call[k] >> 8) == 0
    eng.close()


def test_run_sites_host_equals_device_path(eng4):
    from clairs_to_b200.engine import stream_to_device
    eng = eng4[0]
    (aff, _), (neg, _) = synth.synth_pair(150, 5, 'ont')
    dev = eng.run_sites(stream_to_device(aff, eng.device), stream_to_device(neg, eng.device), 10, posterior=False)
    host = eng.run_sites_host(aff, neg, 10, want_tensors=True)
    assert np.array_equal(host['tensor_aff'].numpy(), dev['tensor_aff'].cpu().numpy())
    assert np.array_equal(host['tensor_neg'].numpy(), dev['tensor_neg'].cpu().numpy())
    assert np.array_equal(host['probs'].numpy(), dev['probs'].cpu().numpy())
    # Illumina-style single stream (NEG symlinked to AFF, run_clairs_to:1248-1252)
    one = eng.run_sites_host(neg, None, 10)
    two = eng.run_sites_host(neg, neg, 10)
    assert np.array_equal(one['probs'].numpy(), two['probs'].numpy())


def test_full_size_properties(eng4):
    """BASELINE config-2 scale slice: size-independent checks without running the python oracle."""
    from clairs_to_b200.engine import stream_to_device
    eng = eng4[0]
    n = 20000
    (aff, _), (neg, _) = synth.synth_pair(n, 123, 'ont')
    s = neg
    t, d = eng.encode(stream_to_device(s, eng.device), 10)
    t = t.cpu().numpy().astype(np.int64)
    rows = np.repeat(np.arange(s.n_rows), np.diff(s.pos_off))
    plain = (s.code & 0x10) == 0
    hi_mq = s.mq >= 20
    sym = s.code & 0xF
    flat = t.reshape(-1, N_CH)
    # checksum of checksums: '*' and '#' channels, and LMQ/LBQ group sums, against numpy bincounts
    for symbol, ch in ((10, 8), (11, 17)):
        want = np.bincount(rows[plain & hi_mq & (sym == symbol)], minlength=s.n_rows)
        assert np.array_equal(flat[:, ch], want)
    is_base = plain & (((sym < 4)) | ((sym >= 5) & (sym <= 8)))
    fwd_base = plain & (sym < 4)
    want_ref_fwd = -np.bincount(rows[fwd_base & hi_mq], minlength=s.n_rows)
    ref = s.ref_code.astype(np.int64)
    assert np.array_equal(flat[np.arange(s.n_rows), ref], want_ref_fwd)          # ref channel = -(A+C+G+T)
    lbq = np.bincount(rows[is_base & (s.bq < 10)], minlength=s.n_rows)
    got_lbq = -(flat[np.arange(s.n_rows), 26 + ref] + flat[np.arange(s.n_rows), 30 + ref])
    assert np.array_equal(got_lbq, lbq)
    # permutation equivariance of the whole device path
    perm = np.random.default_rng(0).permutation(n)[:4096]
    win = s.win_pos.reshape(n, N_POS)[perm].reshape(-1).copy()
    sp = PileupStream(s.code, s.bq, s.mq, s.pos_off, s.ref_code, s.ind_off, s.ind_entry, win)
    a = eng.run_sites(stream_to_device(sp, eng.device), None, 10, posterior=False)
    b = eng.run_sites(stream_to_device(s, eng.device), None, 10, posterior=False)
    assert np.array_equal(a['tensor_aff'].cpu().numpy(), b['tensor_aff'].cpu().numpy()[perm])
    assert np.abs(a['probs'].cpu().numpy() - b['probs'].cpu().numpy()[perm]).max() < 1e-5
    assert np.isfinite(b['probs'].cpu().numpy()).all()
    assert np.allclose(b['probs'].cpu().numpy().sum(-1), 1.0, atol=1e-5)


@pytest.mark.parametrize("n", [200, 129, 128, 64, 1])
def test_gru_pair_kernel_ragged_batches(golden_dir, n):
    """The CTA-pair GRU kernel works on 128-candidate pairs: batches that leave the last pair partly, half or
    almost entirely empty must give the same logits as the oracle."""
    eng, _, neg_sd = _engine(4, max_batch=512)
    try:
        rng = np.random.default_rng(1000 + n)
        x = torch.from_numpy(rng.integers(-50, 51, size=(n, 33, 34)).astype(np.float32))
        got = eng.forward_neg(x).cpu().numpy()
        want = nn_oracle.neg_forward(x.numpy(), neg_sd).numpy()
        assert np.abs(got - want).max() < TOL
    finally:
        eng.close()


def test_run_sites_host_pipelined_chunks_and_sparse_windows():
    """cto_run_sites_host copies the read arrays span by span while earlier chunks compute: many small
    chunks (max_batch 64), windows with missing rows (-1) and windows that overlap the previous candidate."""
    from clairs_to_b200.engine import stream_to_device
    eng, _, _ = _engine(4, max_batch=64)
    (aff, _), (neg, _) = synth.synth_pair(300, 17, 'ont', depth_lo=0, depth_hi=90)
    rng = np.random.default_rng(3)
    for s in (aff, neg):
        win = s.win_pos.reshape(-1, N_POS).copy()
        win[rng.random(win.shape) < 0.05] = -1                    # rows without pileup output (CT:461)
        win[5] = win[4] + 7                                        # candidate 5 re-uses rows of candidate 4 (overlap)
        win[5][win[4] < 0] = -1
        s.win_pos = win.reshape(-1)
    neg.win_pos = aff.win_pos.copy()
    dev = eng.run_sites(stream_to_device(aff, eng.device), stream_to_device(neg, eng.device), 30, posterior=False)
    host = eng.run_sites_host(aff, neg, 30, want_tensors=True)
    assert np.array_equal(host['tensor_aff'].numpy(), dev['tensor_aff'].cpu().numpy())
    assert np.array_equal(host['tensor_neg'].numpy(), dev['tensor_neg'].cpu().numpy())
    assert np.array_equal(host['probs'].numpy(), dev['probs'].cpu().numpy())
    t = host['tensor_aff'].numpy()
    assert not t.reshape(-1, N_CH)[aff.win_pos < 0].any()         # absent rows are all-zero rows
    eng.close()


@pytest.mark.parametrize("platform,n_heads,single_stream", [("ont", 6, False), ("ilmn", 4, True), ("hifi", 6, False)])
def test_baseline_config_shapes_end_to_end(platform, n_heads, single_stream):
    """BASELINE.json configs 3-5 in miniature: indel (6-head) models, the Illumina single-stream case
    (NEG symlinked to AFF, run_clairs_to:1248-1252) and HiFi, host arrays in -> posteriors out, against the oracle."""
    lk = np.concatenate([np.random.default_rng(1).uniform(0.05, 0.95, size=(10 * n_heads, 10)),
                         np.sort(np.random.default_rng(2).uniform(0.02, 0.98, size=(2 * n_heads, 10)), axis=1)])
    eng, aff_sd, neg_sd = _engine(n_heads, max_batch=128, likelihood=lk)
    eng.set_qual_thresholds(8.0, 8.0, 12.0)                        # shared/param.py:35-40 (ont / hifi)
    (aff, _), (neg, _) = synth.synth_pair(200, 41, platform, depth_mean=70, depth_hi=180)
    out = eng.run_sites_host(aff, None if single_stream else neg, 30, want_tensors=True)
    ta, tn = out['tensor_aff'].numpy(), out['tensor_neg'].numpy()
    if single_stream:
        assert np.array_equal(ta, tn)
    from clairs_to_b200.engine import stream_to_device
    da = eng.encode(stream_to_device(aff, eng.device), 30)[1].cpu().numpy()
    dn = eng.encode(stream_to_device(aff if single_stream else neg, eng.device), 30)[1].cpu().numpy()
    xa = np.stack([posterior_oracle.rescale_tensor(t, d) for t, d in zip(ta, da)])
    xn = np.stack([posterior_oracle.rescale_tensor(t, d) for t, d in zip(tn, dn)])
    pa = nn_oracle.softmax_heads(nn_oracle.aff_forward(xa, aff_sd)).numpy()
    pn = nn_oracle.softmax_heads(nn_oracle.neg_forward(xn, neg_sd)).numpy()
    probs = out['probs'].numpy()
    assert np.abs(probs[:, :n_heads] - pa).max() < TOL and np.abs(probs[:, n_heads:] - pn).max() < TOL
    mats, ea, en = posterior_oracle.load_likelihood(lk, n_heads)
    post = out['post'].numpy()
    qual, flt = out['qual'].numpy(), out['filter'].numpy()
    far = 0
    for k in range(200):
        p8 = [float("{:0.8f}".format(v)) for v in probs[k, :, 1]]
        want = posterior_oracle.posterior(p8[:n_heads], p8[n_heads:], mats, ea, en)
        assert np.array_equal(post[k], want)                       # fp64 combine is bit-exact on identical inputs
        # QUAL of the winning posterior (call_variants.py:81-88) and the three threshold bits, evaluated on the device
        q = posterior_oracle.quality_score(want.max())
        assert abs(qual[k] - q) <= 1.0001e-4
        assert flt[k] == (1 if qual[k] >= 8.0 else 0) | (2 if qual[k] >= 8.0 else 0) | (4 if qual[k] >= 12.0 else 0)
        r8 = [float("{:0.8f}".format(v)) for v in np.concatenate([pa[k, :, 1], pn[k, :, 1]])]
        ref = posterior_oracle.posterior(r8[:n_heads], r8[n_heads:], mats, ea, en)
        # End to end against the oracle's own probabilities: the posterior is within 1e-3 absolute UNLESS the difference
        # is explained by (a) a probability within 1e-3 of a likelihood bin edge (the two sides may fall into different
        # bins: the reference's formula is discontinuous there), or (b) an ill-conditioned head: both class probabilities
        # so small that num + alt < 0.02, where 1e-3 on a probability moves the ratio by more than 1e-3.
        for h in range(n_heads):
            if abs(ref[h] - want[h]) <= TOL:
                continue
            p, q1 = r8[h], 1.0 - r8[n_heads + h]
            near_edge = np.abs(ea[h] - p).min() <= TOL or np.abs(en[h] - q1).min() <= TOL
            i = min(max(int(np.digitize(p, ea[h])) - 1, 0), 9)
            j = min(max(int(np.digitize(q1, en[h])) - 1, 0), 9)
            w = mats[h][i][j]
            ill = p * q1 * w + (1 - p) * (1 - q1) * (1 - w) < 0.02
            assert near_edge or ill, (k, h, ref[h], want[h], p, q1)
            far += 1
    assert far <= 8                                                 # and such heads are rare (<= 1 % of 200 x H)
    eng.close()
