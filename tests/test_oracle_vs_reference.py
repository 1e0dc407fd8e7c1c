"""CPU, build container only: the oracle against the LIVE reference (imported read-only from /root/reference), beyond
the committed golden vectors -- randomised fuzz of the three pieces the oracle restates, and the reference's own
checkpoint format (whole pickled modules, clairs/predict.py:513-517) through clairs_to_b200.weights.

Skipped wherever /root/reference does not exist (the GPU box); the golden-vector tests cover that case."""

import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from clairs_to_b200 import synth
from oracle import nn_oracle, pileup_oracle, posterior_oracle

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "clairs")), reason="reference checkout not present")

CLASS_DEFAULT_CVT = dict(s1=(32, 1, 1), s2=(64, 3, 2), s3=(128, 6, 10))


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, REF)
    try:
        import clairs.model as model
        import clairs.call_variants as cv
        import src.create_tensor_pileup_calling as ct
        yield dict(model=model, cv=cv, ct=ct)
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split('.')[0] in ("clairs", "shared", "src")]:
            del sys.modules[k]


def test_encoder_oracle_fuzz(ref):
    """decode_pileup_bases (src/create_tensor_pileup_calling.py:95-233) on 600 random synthetic rows, all platforms."""
    ct = ref["ct"]
    n = 0
    for seed, platform, literal in ((1, 'ont', 'ont'), (2, 'ont', 'ont_r10_dorado_sup_5khz'), (3, 'ilmn', 'ilmn_ssrs'), (4, 'hifi', 'hifi_revio')):
        stream, aux = synth.synth_stream(5, 400 + seed, platform, depth_lo=0, depth_hi=120)
        for i, text in enumerate(synth.render_mpileup(stream, aux, decorate_seed=seed)):
            cols = text.rstrip("\n").split("\t")
            r = "ACGT"[int(stream.ref_code[i])]
            mq = [ord(c) - 33 for c in cols[6]]
            bq = [ord(c) - 33 for c in cols[5]]
            want_vec, want_alt = _decode(ct, cols[4], r, mq, bq, literal)
            vec, alt = pileup_oracle.position_vector(cols[4], mq, bq, r, is_candidate=True, chunk_ref_seq=(r + "ACGTTGCA" * 8)[:60],
                                                     platform=literal)
            assert vec == [int(v) for v in want_vec] and alt == want_alt
            n += 1
    assert n >= 600


def _decode(ct, bases, ref_base, mq, bq, platform):
    from argparse import Namespace
    vec, _, _, _, _, alt = ct.decode_pileup_bases(
        args=Namespace(max_indel_length=60), pos=100, pileup_bases=bases, reference_base=ref_base,
        minimum_snp_af_for_candidate=0.05, minimum_indel_af_for_candidate=0.05, has_pileup_candidates=True,
        candidates_type_dict={100: 'snv'}, is_tumor=True, mapping_quality=mq, base_quality=bq, phasing_info=None,
        chunk_ref_seq=(ref_base + "ACGTTGCA" * 8)[:60], platform=platform)
    return vec, alt


@pytest.mark.parametrize("n_heads,cfg", [(4, None), (6, None), (4, CLASS_DEFAULT_CVT), (6, CLASS_DEFAULT_CVT)])
def test_nn_oracle_vs_reference_modules(ref, n_heads, cfg):
    """CvT / CvT_Indel with the predict.py hyper-parameters and with the class defaults, BiGRU_NACGT / _Indel: the
    reference modules on random count-like input vs the oracle on the same state_dict."""
    m = ref["model"]
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    kw = {} if cfg else make_golden.CVT_KW
    aff = (m.CvT if n_heads == 4 else m.CvT_Indel)(**kw).eval()
    neg = (m.BiGRU_NACGT if n_heads == 4 else m.BiGRU_NACGT_Indel)(apply_softmax=False, num_classes=2, channel_size=34, model_type="nacgt").eval()
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads, cfg), 500 + n_heads, 0.8)
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 600 + n_heads, 0.8)
    aff.load_state_dict(aff_sd, strict=False)
    neg.load_state_dict(neg_sd, strict=False)
    rng = np.random.default_rng(n_heads)
    x = rng.integers(-50, 51, size=(16, 33, 34)).astype(np.float32)
    x[rng.random(x.shape) < 0.5] = 0
    with torch.no_grad():
        la = torch.stack(aff(torch.from_numpy(x)), 1).numpy()
        ln = torch.stack(neg(torch.from_numpy(x)), 1).numpy()
    assert np.abs(nn_oracle.aff_forward(x, aff_sd).numpy() - la).max() < 3e-5
    assert np.abs(nn_oracle.neg_forward(x, neg_sd).numpy() - ln).max() < 3e-5


def test_quality_score_matches_reference(ref):
    cv = ref["cv"]
    rng = np.random.default_rng(0)
    for p in list(rng.random(2000)) + [0.0, 1.0, 0.5, 1e-12, 1 - 1e-12]:
        assert posterior_oracle.quality_score(p) == cv.quality_score_from(p)


def test_reference_pickled_checkpoints_load_without_the_reference_package(ref, tmp_path):
    """The reference checkpoint format: ``torch.save({'model_acgt': <CvT module>})`` (clairs/predict.py:513-517).  Loaded
    in a fresh interpreter WITHOUT /root/reference on sys.path, so the tolerant unpickler branch of
    clairs_to_b200.weights.state_dict_from_checkpoint is the one that runs; the result must equal module.state_dict()
    bit for bit and export to an engine blob."""
    m = ref["model"]
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    aff = m.CvT(**make_golden.CVT_KW).eval()
    neg = m.BiGRU_NACGT(apply_softmax=False, num_classes=2, channel_size=34, model_type="nacgt").eval()
    aff.load_state_dict(nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104), strict=False)
    neg.load_state_dict(nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 204), strict=False)
    pa, pn = str(tmp_path / "pileup_affirmative.pkl"), str(tmp_path / "pileup_negational.pkl")
    torch.save({'model_acgt': aff}, pa)
    torch.save({'model_nacgt': neg}, pn)
    torch.save({k: v for k, v in aff.state_dict().items()}, str(tmp_path / "aff_sd.pt"))
    torch.save({k: v for k, v in neg.state_dict().items()}, str(tmp_path / "neg_sd.pt"))
    code = """
import sys, torch
assert not any('reference' in p for p in sys.path), sys.path
from clairs_to_b200.weights import state_dict_from_checkpoint, export_aff, export_neg
try:
    import clairs.model
    raise SystemExit('clairs.model must not be importable in this process')
except ImportError:
    pass
for path, key, sd_path, export in ((%r, 'model_acgt', %r, export_aff), (%r, 'model_nacgt', %r, export_neg)):
    sd = state_dict_from_checkpoint(path, key)
    want = torch.load(sd_path)
    assert set(sd) == set(want), (set(sd) ^ set(want))
    assert all(torch.equal(sd[k], want[k]) for k in want)
    blob, cfg = export(sd)
    print(key, len(blob), cfg.tolist())
""" % (pa, str(tmp_path / "aff_sd.pt"), pn, str(tmp_path / "neg_sd.pt"))
    env = dict(os.environ, PYTHONPATH=ROOT)
    out = subprocess.run([sys.executable, "-c", code], env=env, cwd=str(tmp_path), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "model_acgt" in out.stdout and "[4, 3, 16, 1, 1, 64, 3, 2, 128, 4, 3]" in out.stdout
    assert "model_nacgt" in out.stdout and "[4, 34, 128, 192]" in out.stdout


def test_candidates_oracle_fuzz():
    """oracle/candidates_oracle.site_decision against the reference's decode_pileup_bases of STEP 1
    (src/extract_candidates_calling.py:55-169) + the set rules of ibid. 355-377, on 4000 synthetic rows incl. the
    tokenizer's corner cases, for both --select_indel_candidates settings and thresholds on and off the planted AFs."""
    sys.path.insert(0, REF)
    try:
        import src.extract_candidates_calling as ec
        from oracle import candidates_oracle as co
        rows, reference = synth.scan_rows_text(4000, 77, first_pos=1, depth_mean=30, weird=0.02)
        checked = 0
        for k, row in enumerate(rows):
            cols = row.strip().split('\t')
            rb = reference[int(cols[1]) - 1].upper()
            if rb not in "ACGT":
                continue
            select = bool(k & 1)
            kw = dict(min_coverage=(4, 0, 10)[k % 3], snv_min_af=(0.05, 0.08, 0.0)[k % 3], indel_min_af=(0.05, 0.1, 1.0)[(k // 2) % 3],
                      alternative_base_num=(3, None, 1)[(k // 3) % 3], select_indel_candidates=select)
            got = co.site_decision(cols[4], rb, **kw)
            base_list, depth, pass_af, af, af_infos, pileup_infos, tumor_infos, alt_list, pass_snv_af, pass_indel_af, pileup_list = \
                ec.decode_pileup_bases(pileup_bases=cols[4], reference_base=rb, min_coverage=kw["min_coverage"],
                                       minimum_snv_af_for_candidate=kw["snv_min_af"], minimum_indel_af_for_candidate=kw["indel_min_af"],
                                       alternative_base_num=kw["alternative_base_num"], has_pileup_candidates=False, read_name_list=[],
                                       is_tumor=False, select_indel_candidates=select)
            snv = bool(pass_af and pass_snv_af and len([i for i in alt_list if i[0] in "ACGT"]) > 0)          # EC:366-371
            indel = bool(select and pass_af and pass_indel_af and len([i for i in alt_list if '+' in i[0] or '-' in i[0]]) > 0)
            assert got == (depth, bool(pass_af), snv, indel), (k, cols[4], kw)
            checked += 1
        assert checked > 3500
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k.split('.')[0] in ("clairs", "shared", "src")]:
            del sys.modules[k]


def test_hard_filter_oracle_fuzz():
    """oracle/hard_filter_oracle.site_line against the reference's per-site functions:
    _haplotype_build_state_and_line (src/haplotype_filtering.py:570-703) on phased chunks and _postfilter_build_state_and_line
    (src/postfilter_variants.py:368-446) on unphased ones, plus fisher_exact and calculate_sequence_entropy on their own."""
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "src"))
    try:
        import src.haplotype_filtering as HF
        import src.postfilter_variants as PV
        from oracle import hard_filter_oracle as ho
        checked = failing = 0
        for seed in range(4):
            for phased in (True, False):
                rows, ref, lo, sites = synth.hard_filter_chunk(10, 100 + seed, with_phasing=phased, depth=(20, 40, 80)[seed % 3],
                                                               read_len=((120, 900), (60, 300))[seed % 2])
                theirs = HF._parse_mpileup_to_chunk_dict(rows) if phased else PV._parse_mpileup_postfilter_chunk_dict(rows)
                mine = ho.parse_chunk(rows, phased)
                for pos, rb, ab, af, het, hom in sites:
                    for disable, max_co in ((False, 3), (True, 2)):
                        if phased:
                            want = HF._haplotype_build_state_and_line("chr20", pos, rb, ab, 100, theirs, ref, lo, het, hom, disable, max_co, af, 50.0)
                            got = ho.site_line('haplotype', "chr20", pos, rb, ab, 100, mine, ref, lo, het, hom, disable, max_co, af)
                        else:
                            want = PV._postfilter_build_state_and_line("chr20", pos, rb, ab, 100, theirs, ref, lo, disable, max_co)
                            got = ho.site_line('postfilter', "chr20", pos, rb, ab, 100, mine, ref, lo, None, None, disable, max_co, None)
                        assert got == want, (seed, phased, pos, rb, ab)
                        checked += 1
                        failing += want.split()[2] == "False"
        assert checked >= 150 and failing > 20
        rng = np.random.default_rng(5)
        for t in [tuple(int(x) for x in rng.integers(0, 60, 4)) for _ in range(300)] + [(k, k + 1, k + 1, k) for k in range(40)]:
            assert ho.fisher_exact(*t) == HF.fisher_exact([[t[0], t[1]], [t[2], t[3]]]), t
        for _ in range(50):
            q = ''.join("ACGT"[b] for b in rng.integers(0, 4, 33))
            assert ho.entropy_of(q) == PV.calculate_sequence_entropy(q, entropy_window=33)
    finally:
        sys.path.remove(REF)
        sys.path.remove(os.path.join(REF, "src"))
        for k in [k for k in sys.modules if k.split('.')[0] in ("clairs", "shared", "src", "haplotype_filtering", "postfilter_variants")]:
            del sys.modules[k]
