"""Drop-in sub-commands against files produced by the UNMODIFIED reference CLIs (tests/golden):

  create_tensor_pileup_calling  -> tensor_can chunk file, byte-identical text
  predict                       -> predict chunk file, identical fields, probabilities within 1e-3
  call_variants                 -> per-chunk VCF, identical text

plus (CPU) the oracle's window assembly against the reference's tensor_can file."""

import gzip
import os
import sys

import numpy as np
import pytest
import torch

from oracle import nn_oracle, pileup_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "fake_samtools.py")


def _ct_inputs(golden_dir):
    work = os.path.join(golden_dir, "create_tensor")
    seq = "".join(l.strip() for l in open(os.path.join(work, "ref.fa")) if not l.startswith(">"))
    cand = {}
    for row in open(os.path.join(work, "chr20.0_0_9_snv")):
        c = row.rstrip().split("\t")
        position, end = int(c[1]) + 1, int(c[2]) + 1
        cand[position + (end - position) // 2 - 1] = 'unknown'
    return work, seq, cand


@pytest.mark.parametrize("name,min_bq", [("aff", 20), ("neg", 0)])
def test_oracle_window_assembly_matches_reference_create_tensor(golden_dir, name, min_bq):
    work, seq, cand = _ct_inputs(golden_dir)
    rows = open(os.path.join(work, "tumor.bam.minbq%d.mpileup" % min_bq)).readlines()
    ctg_start = min(cand) - 16
    ctg_start = 2                        # first region row is (1, 27): position = 2 (create_tensor...:356-360)
    ctg_end = max(cand) + 18
    reference_start = max(1, ctg_start - 1000)
    extend_start, extend_end = max(1, ctg_start - 33), ctg_end + 33
    ref = seq[reference_start - 1: ctg_end + 1000].upper()
    _, text = pileup_oracle.encode_windows(rows, cand, ref, reference_start, extend_start, extend_end, "chr20",
                                           platform="ont_r10_dorado_sup_5khz", candidate_types=cand)
    want = gzip.open(os.path.join(work, "tensor_can_" + name), "rt").readlines()
    assert text == want


@pytest.mark.gpu
@pytest.mark.parametrize("name,min_bq", [("aff", 20), ("neg", 0)])
def test_create_tensor_cli_byte_identical(golden_dir, tmp_path, name, min_bq):
    from clairs_to_b200 import create_tensor_pileup_calling as ct
    work = os.path.join(golden_dir, "create_tensor")
    out = str(tmp_path / ("tensor_can_" + name))
    ct.main(["--tumor_bam_fn", os.path.join(work, "tumor.bam"), "--ref_fn", os.path.join(work, "ref.fa"),
             "--ctg_name", "chr20", "--samtools", SHIM, "--min_bq", str(min_bq),
             "--candidates_bed_regions", os.path.join(work, "chr20.0_0_9_snv"), "--tensor_can_fn", out,
             "--platform", "ont_r10_dorado_sup_5khz"])
    got = gzip.open(out, "rt").read()
    want = gzip.open(os.path.join(work, "tensor_can_" + name), "rt").read()
    assert got == want


def _save_checkpoints(tmp_path, n_heads):
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads), 100 + n_heads)
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 200 + n_heads)
    a, n = str(tmp_path / "aff.pkl"), str(tmp_path / "neg.pkl")
    torch.save({'model_acgt': aff_sd}, a)          # same key layout as the reference's pickled modules
    torch.save({'model_nacgt': neg_sd}, n)
    return a, n


def _module_tree(sd, class_names):
    """A torch.nn.Module object graph with the attribute layout of a reference module (parameters / buffers under the
    dotted names of ``sd``), whose classes are called ``clairs.model.<name>``: what ``torch.save({'model_acgt': model})``
    pickles in the reference (clairs/predict.py:513-517).  The classes are bare shells registered under a temporary
    ``clairs.model`` module, so the file refers to ``clairs.model.CvT`` etc. exactly like a real checkpoint."""
    import types
    mod = types.ModuleType("clairs.model")
    pkg = types.ModuleType("clairs")
    pkg.model = mod
    classes = {}
    for name in class_names:
        classes[name] = type(name, (torch.nn.Module,), {"__module__": "clairs.model"})
        setattr(mod, name, classes[name])
    root = classes[class_names[0]]()
    inner = classes[class_names[1]]
    for key, value in sd.items():
        node = root
        parts = key.split(".")
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, inner())
            node = node._modules[part]
        if parts[-1].startswith("running_") or parts[-1] == "num_batches_tracked":
            node.register_buffer(parts[-1], value.clone())
        else:
            node.register_parameter(parts[-1], torch.nn.Parameter(value.clone(), requires_grad=False))
    return root, {"clairs": pkg, "clairs.model": mod}


def _save_pickled_module_checkpoints(tmp_path, n_heads):
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads), 100 + n_heads)
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 200 + n_heads)
    a, n = str(tmp_path / "pileup_affirmative.pkl"), str(tmp_path / "pileup_negational.pkl")
    saved = {k: sys.modules.get(k) for k in ("clairs", "clairs.model")}
    try:
        for path, key, sd, names in ((a, 'model_acgt', aff_sd, ["CvT" if n_heads == 4 else "CvT_Indel", "Transformer"]),
                                     (n, 'model_nacgt', neg_sd, ["BiGRU_NACGT" if n_heads == 4 else "BiGRU_NACGT_Indel", "Transformer"])):
            model, mods = _module_tree(sd, names)
            sys.modules.update(mods)
            torch.save({key: model}, path)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return a, n


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n_heads,pickled_modules", [("snv", 4, False), ("indel", 6, False), ("snv", 4, True), ("indel", 6, True)])
def test_predict_cli_against_reference_predict_file(golden_dir, tmp_path, tag, n_heads, pickled_modules):
    """pickled_modules: the checkpoints are whole pickled ``clairs.model`` modules (the reference's real format) and
    ``clairs.model`` is NOT importable when the sub-command loads them (tests/test_oracle_vs_reference.py checks the same
    loader on checkpoints written by the reference itself)."""
    from clairs_to_b200 import predict as pr
    pdir = os.path.join(golden_dir, "pipeline")
    ck_a, ck_n = (_save_pickled_module_checkpoints if pickled_modules else _save_checkpoints)(tmp_path, n_heads)
    if pickled_modules:
        assert "clairs.model" not in sys.modules
    out = str(tmp_path / ("predict_" + tag))
    pr.main(["--tensor_fn_acgt", os.path.join(pdir, "tensor_can_aff_" + tag),
             "--tensor_fn_nacgt", os.path.join(pdir, "tensor_can_neg_" + tag), "--predict_fn", out,
             "--chkpnt_fn_acgt", ck_a, "--chkpnt_fn_nacgt", ck_n, "--use_gpu", "True",
             "--platform", "ont_r10_dorado_sup_5khz", "--ctg_name", "chr20", "--pileup",
             "--disable_indel_calling", "True" if n_heads == 4 else "False"])
    got = [r.rstrip("\n").split("\t") for r in gzip.open(out, "rt")]
    want = [r.rstrip("\n").split("\t") for r in gzip.open(os.path.join(pdir, "predict_" + tag), "rt")]
    assert len(got) == len(want) == 39
    for g, w in zip(got, want):
        assert len(g) == len(w)
        assert g[:6] == w[:6]                                     # chrom, pos, ref, alt_info, strand-count list reprs
        for a, b in zip(g[6:6 + 2 * n_heads], w[6:6 + 2 * n_heads]):
            assert np.allclose([float(v) for v in a.split()], [float(v) for v in b.split()], atol=1e-3)
            assert all(len(v.split(".")[1]) == 8 for v in a.split())
        assert g[6 + 2 * n_heads:] == w[6 + 2 * n_heads:]


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n_heads,show_ref", [("snv", 4, False), ("snv", 4, True), ("indel", 6, False), ("indel", 6, True),
                                                  ("indel_hand", 6, False), ("indel_hand", 6, True)])
def test_call_variants_cli_identical_vcf(golden_dir, tmp_path, tag, n_heads, show_ref):
    """``indel_hand``: a hand-written predict file whose probabilities drive insertion / deletion calls ('#'-anchored
    insertions, multi-base deletions, competing alleles; tests/golden/make_golden.py:indel_call_golden) -- the seeded-weight
    pipeline golden holds no I/D ALT row."""
    from clairs_to_b200 import call_variants as cv
    pdir = os.path.join(golden_dir, "pipeline")
    out = str(tmp_path / "out" / ("call_%s.vcf" % tag))
    argv = ["--predict_fn", os.path.join(pdir, "predict_" + tag), "--call_fn", out,
            "--ref_fn", os.path.join(pdir, "ref.fa"), "--platform", "ont_r10_dorado_sup_5khz",
            "--likelihood_matrix_data", os.path.join(pdir, "likelihood_%s.txt" % tag.split("_")[0]),
            "--disable_indel_calling", "True" if n_heads == 4 else "False"]
    if show_ref:
        argv.append("--show_ref")
    cv.main(argv)
    ref_vcf = os.path.join(pdir, "call_%s%s.vcf" % (tag, "_showref" if show_ref else ""))
    if not os.path.exists(ref_vcf):
        assert not os.path.exists(out)              # the reference removed its empty VCF, so must we
        return
    assert open(out).read() == open(ref_vcf).read()
    if tag == "indel_hand":
        rows = [r.split("\t") for r in open(out) if r[0] != '#']
        assert sum(len(r[3]) > 1 for r in rows) >= 3 and sum(len(r[4]) > 1 for r in rows) >= 3


def test_vcf_header_matches_reference_fixture(golden_dir):
    """CPU: the stand-alone VCF header equals the reference writer's (first lines of a golden VCF)."""
    from clairs_to_b200.call_variants import vcf_header_text
    want = [l for l in open(os.path.join(golden_dir, "pipeline", "call_snv_showref.vcf")) if l.startswith("##")
            and not l.startswith("##contig")]
    assert vcf_header_text() == "".join(want)


@pytest.mark.gpu
def test_hot_path_in_memory_equals_the_three_sub_commands(golden_dir, tmp_path):
    """clairs_to_b200.hot_path (encoder -> networks -> posterior -> VCF in one process, tensors never leave the GPU) against
    the three drop-in sub-commands run one after the other through their gzip chunk files (each of which is pinned to the
    reference above): identical VCF, and identical intermediate files when asked to write them."""
    from clairs_to_b200 import call_variants as cv, create_tensor_pileup_calling as ct, hot_path, predict as pr
    work = os.path.join(golden_dir, "create_tensor")
    lk = os.path.join(golden_dir, "pipeline", "likelihood_snv.txt")
    ck_a, ck_n = _save_checkpoints(tmp_path, 4)
    common = ["--tumor_bam_fn", os.path.join(work, "tumor.bam"), "--ref_fn", os.path.join(work, "ref.fa"), "--ctg_name", "chr20",
              "--samtools", SHIM, "--candidates_bed_regions", os.path.join(work, "chr20.0_0_9_snv"),
              "--platform", "ont_r10_dorado_sup_5khz"]
    files = {k: str(tmp_path / k) for k in ("aff", "neg", "predict", "vcf_files", "aff2", "neg2", "predict2", "vcf_mem")}
    for name, min_bq in (("aff", 20), ("neg", 0)):
        ct.main(common + ["--min_bq", str(min_bq), "--tensor_can_fn", files[name]])
    pr.main(["--tensor_fn_acgt", files["aff"], "--tensor_fn_nacgt", files["neg"], "--predict_fn", files["predict"],
             "--chkpnt_fn_acgt", ck_a, "--chkpnt_fn_nacgt", ck_n, "--use_gpu", "True", "--platform", "ont_r10_dorado_sup_5khz",
             "--ctg_name", "chr20", "--pileup", "--disable_indel_calling", "True"])
    cv.main(["--predict_fn", files["predict"], "--call_fn", files["vcf_files"], "--ref_fn", os.path.join(work, "ref.fa"),
             "--platform", "ont_r10_dorado_sup_5khz", "--likelihood_matrix_data", lk, "--disable_indel_calling", "True", "--show_ref"])
    hot_path.main(common + ["--min_bq", "20", "--chkpnt_fn_acgt", ck_a, "--chkpnt_fn_nacgt", ck_n, "--likelihood_matrix_data", lk,
                            "--disable_indel_calling", "True", "--show_ref", "--call_fn", files["vcf_mem"],
                            "--tensor_can_fn_acgt", files["aff2"], "--tensor_can_fn_nacgt", files["neg2"], "--predict_fn", files["predict2"]])
    assert os.path.exists(files["vcf_files"]) and open(files["vcf_mem"]).read() == open(files["vcf_files"]).read()
    assert sum(1 for l in open(files["vcf_mem"]) if not l.startswith("#")) >= 3
    for a, b in (("aff", "aff2"), ("neg", "neg2"), ("predict", "predict2")):
        assert gzip.open(files[a], "rb").read() == gzip.open(files[b], "rb").read(), a
