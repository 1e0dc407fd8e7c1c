"""CPU: pin the oracle to the fixtures the reference itself produced (tests/golden/make_golden.py)."""

import gzip
import json
import os

import numpy as np
import pytest
import torch

from oracle import nn_oracle, pileup_oracle, posterior_oracle


def test_encoder_oracle_matches_reference_decode(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "encoder_golden.json")))
    assert len(cases) > 400
    for c in cases:
        vec, alt_info = pileup_oracle.position_vector(
            c["bases"], [ord(ch) - 33 for ch in c["mq"]], [ord(ch) - 33 for ch in c["bq"]], c["ref"],
            is_candidate=c["candidate"], chunk_ref_seq=c["chunk_ref"], platform=c["platform"])
        assert vec == c["vec"], c["bases"]
        assert alt_info == c["alt_info"], c["bases"]


@pytest.mark.parametrize("n_heads", [4, 6])
def test_nn_oracle_matches_reference_forward(golden_dir, n_heads):
    g = np.load(os.path.join(golden_dir, "nn_golden.npz"))
    x = g["x_%d" % n_heads]
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads), 100 + n_heads)
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 200 + n_heads)
    la = nn_oracle.aff_forward(x, aff_sd).numpy()
    ln = nn_oracle.neg_forward(x, neg_sd).numpy()
    # same arithmetic (torch CPU fp32) in a different op order: 2e-5 absolute
    assert np.abs(la - g["aff_logits_%d" % n_heads]).max() < 2e-5
    assert np.abs(ln - g["neg_logits_%d" % n_heads]).max() < 2e-5
    assert np.abs(g["aff_logits_%d" % n_heads]).max() > 0.1      # the comparison is not vacuous


@pytest.mark.parametrize("n_heads", [4, 6])
def test_nn_oracle_matches_reference_class_default_cvt(golden_dir, n_heads):
    """CvT() / CvT_Indel() with the CLASS-DEFAULT hyper-parameters (clairs/model.py:153-176: stage-1 width 32, stage 3 with
    6 heads and depth 10), golden logits from the unmodified reference (make_golden.nn_golden_default)."""
    g = np.load(os.path.join(golden_dir, "nn_golden_default.npz"))
    cfg = dict(s1=(32, 1, 1), s2=(64, 3, 2), s3=(128, 6, 10))
    sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads, cfg), 300 + n_heads, 0.7)
    la = nn_oracle.aff_forward(g["x_%d" % n_heads], sd).numpy()
    assert np.abs(la - g["aff_logits_%d" % n_heads]).max() < 2e-5
    assert np.abs(g["aff_logits_%d" % n_heads]).max() > 0.1


def _read_rows(path):
    with gzip.open(path, "rt") as f:
        return [r.rstrip("\n").split("\t") for r in f]


@pytest.mark.parametrize("tag,n_heads", [("snv", 4), ("indel", 6)])
def test_predict_rows_match_reference_predict(golden_dir, tag, n_heads):
    """oracle rescale + forward + softmax + strand counts + row format vs the reference's predict file."""
    pdir = os.path.join(golden_dir, "pipeline")
    aff_rows = _read_rows(os.path.join(pdir, "tensor_can_aff_" + tag))
    neg_rows = _read_rows(os.path.join(pdir, "tensor_can_neg_" + tag))
    ref_rows = _read_rows(os.path.join(pdir, "predict_" + tag))
    keep = [i for i, r in enumerate(aff_rows) if r[2][16] in "ACGT"]          # predict.py:219-220
    assert len(keep) == len(ref_rows) == len(aff_rows) - 1
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads), 100 + n_heads)
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 200 + n_heads)

    def tensors(rows):
        raw = np.array([[int(v) for v in rows[i][3].split()] for i in keep], dtype=np.int32).reshape(-1, 33, 34)
        scaled = np.stack([posterior_oracle.rescale_tensor(raw[k], posterior_oracle.depth_from_alt_info(rows[i][4]))
                           for k, i in enumerate(keep)])
        return raw, scaled

    raw_a, xa = tensors(aff_rows)
    _, xn = tensors(neg_rows)
    pa = nn_oracle.softmax_heads(nn_oracle.aff_forward(xa, aff_sd)).numpy()
    pn = nn_oracle.softmax_heads(nn_oracle.neg_forward(xn, neg_sd)).numpy()
    fwd, rev = posterior_oracle.strand_counts(raw_a.astype(np.float32))
    for k, i in enumerate(keep):
        got = posterior_oracle.format_predict_row(aff_rows[i][0], aff_rows[i][1], aff_rows[i][2][16], aff_rows[i][4],
                                                  fwd[k].tolist(), rev[k].tolist(),
                                                  list(pa[k]) + list(pn[k])).rstrip("\n").split("\t")
        ref = ref_rows[k]
        assert got[:6] == ref[:6]
        assert len(got) == len(ref) == 6 + 2 * n_heads + (1 if n_heads == 4 else 0)
        for a, b in zip(got[6:6 + 2 * n_heads], ref[6:6 + 2 * n_heads]):
            assert np.allclose([float(v) for v in a.split()], [float(v) for v in b.split()], atol=2e-6)


def _vcf_records(path):
    if not os.path.exists(path):
        return []
    return [r.rstrip("\n").split("\t") for r in open(path) if not r.startswith("#")]


@pytest.mark.parametrize("tag,n_heads", [("snv", 4), ("indel", 6)])
def test_posterior_oracle_matches_reference_calls(golden_dir, tag, n_heads):
    """posterior + QUAL + RefCall/variant decision vs the reference call_variants --show_ref VCF."""
    pdir = os.path.join(golden_dir, "pipeline")
    mats, ea, en = posterior_oracle.load_likelihood(os.path.join(pdir, "likelihood_%s.txt" % tag), n_heads)
    recs = {r[1]: r for r in _vcf_records(os.path.join(pdir, "call_%s_showref.vcf" % tag))}
    assert recs
    checked = 0
    for row in _read_rows(os.path.join(pdir, "predict_" + tag)):
        pos, ref_base = row[1], row[2]
        probs = [[float(v) for v in f.split()] for f in row[6:6 + 2 * n_heads]]
        post = posterior_oracle.posterior([p[1] for p in probs[:n_heads]], [p[1] for p in probs[n_heads:]], mats, ea, en)
        k, pmax, is_variant = posterior_oracle.decide(post, ref_base, snv_mode=(n_heads == 4))
        if pos not in recs:
            continue
        rec = recs[pos]
        # QUAL always comes from the arg-max posterior (CV:417-586), for RefCall and variant rows alike
        assert abs(float(rec[5]) - posterior_oracle.quality_score(pmax)) < 1e-4
        if rec[6] != "RefCall":
            assert is_variant
        elif n_heads == 6:
            assert not is_variant
        checked += 1
    assert checked > 5


# ---- STEP 1 candidate extraction ----------------------------------------------------------------
def _candidates_case(golden_dir, name):
    """Inputs of one golden run of the reference `extract_candidates_calling` (tests/golden/make_golden.py:candidates_golden)."""
    import fake_samtools
    work = os.path.join(golden_dir, "candidates")
    argv = json.load(open(os.path.join(work, name, "args.json")))
    opt = {argv[i][2:]: argv[i + 1] for i in range(0, len(argv), 2)}
    ctg = "chr20"
    seq = fake_samtools.read_fasta(os.path.join(work, "ref.fa"))[ctg]
    contig_length = int(open(os.path.join(work, "ref.fa.fai")).read().split("\t")[1])
    return work, opt, ctg, seq, contig_length


def _bed(path, ctg):
    out = []
    for row in open(path):
        if row[0] == '#':
            continue
        c = row.split()
        if c[0] == ctg:
            out.append((int(c[1]), int(c[2]) + (1 if c[1] == c[2] else 0)))
    return out


@pytest.mark.parametrize("name", ["snv_indel", "snv_only", "bed"])
def test_candidates_oracle_matches_reference_files(golden_dir, name):
    """oracle/candidates_oracle.py against the files the unmodified reference wrote (bed, region files, [INFO] line)."""
    from oracle import candidates_oracle as co
    work, opt, ctg, seq, contig_length = _candidates_case(golden_dir, name)
    chunk_id, chunk_num = int(opt["chunk_id"]) - 1, int(opt["chunk_num"])
    bed_range = None
    if "bed_fn" in opt:
        iv = _bed(os.path.join(work, opt["bed_fn"]), ctg)
        bed_range = (min(s for s, _ in iv), max(e for _, e in iv))
    ctg_start, ctg_end = co.chunk_range(chunk_id, chunk_num, contig_length, bed_range)
    lo, hi = co.reads_region(ctg_start, ctg_end)
    rows = [r for r in open(os.path.join(work, "tumor.bam.minbq20.mpileup")) if lo <= int(r.split("\t", 2)[1]) <= hi]
    reference_start = max(ctg_start - 1000, 1)
    select_indel = opt.get("select_indel_candidates") == "True"
    intervals = None
    if select_indel and opt["bed_fn_source"] == "None":
        intervals = _bed(os.path.join(work, opt["call_indels_only_in_these_regions"]), ctg)
    every, snv, indel = co.candidate_lists(
        rows, seq[reference_start - 1:ctg_end + 1000], reference_start, indel_intervals=intervals, min_coverage=4.0,
        snv_min_af=float(opt["snv_min_af"]), indel_min_af=float(opt["indel_min_af"]), alternative_base_num=3,
        select_indel_candidates=select_indel)
    folder = os.path.join(work, name)
    bed_rows = open(os.path.join(folder, "bed", "%s_%d.bed" % (ctg, chunk_id))).read().splitlines()
    assert bed_rows == ["%s\t%d\t%d" % (ctg, p - 1, p) for p in every]
    assert open(os.path.join(folder, "%s.%d_0_1_snv" % (ctg, chunk_id))).read().splitlines() == co.region_rows(ctg, snv)
    if select_indel:
        assert open(os.path.join(folder, "%s.%d_0_1_indel" % (ctg, chunk_id))).read().splitlines() == co.region_rows(ctg, indel)
        assert "Total SNV candidates found: %d, total Indel candidates found: %d" % (len(snv), len(indel)) in \
            open(os.path.join(folder, "stdout.txt")).read()
    assert len(snv) > 20
