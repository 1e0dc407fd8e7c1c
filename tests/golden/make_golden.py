"""Generate the golden fixtures by running the UNMODIFIED reference (imported read-only from
/root/reference) on seeded synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these files are what pins the
oracle (and through it the CUDA path) to the reference's behaviour.  Nothing here is needed at
test time except the written fixtures.
"""

import argparse
import gzip
import json
import os
import subprocess
import sys
from argparse import Namespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from clairs_to_b200 import synth  # noqa: E402
from oracle import nn_oracle  # noqa: E402


HAND_ROWS = [
    # (bases, BQ, MQ, ref_base, candidate?, chunk_ref_seq)
    ("ACGTacgtNn*#", "IIIIIIIIIIII", "]]]]]]]]]]]]", "A", True, "ACGTACGTAC"),
    ("A+2GTA+2GTa+2gtC-1NC-1Nc-1nG", "I" * 7, "]" * 7, "G", True, "GATTACA"),
    ("A+2GT-1NA", "II", "]]", "A", True, "AC"),                     # second indel overwrites the first
    ("^]A$C<>T", "IIIII", "]]5]]", "C", True, "CCCC"),             # '<' '>' shift the qualities (quirk 6)
    ("A+70" + "G" * 70 + "A+60" + "C" * 60 + "T-60" + "N" * 60 + "T-59" + "N" * 59, "IIII", "]]]]", "T", True, "T" * 60),
    ("*+1A#+1a*-2NN#-2nnN+1T", "I" * 5, "]" * 5, "A", True, "ACG"),
    ("AAAAaaaaCCcc", "+5+5+5+5+5+5", "]!]!]!]!]!]!", "A", True, "A"),   # low MQ / low BQ mix
    ("TTTT", "IIII", "]]", "T", False, "T"),                         # MQ string shorter than reads
    ("", "", "", "A", True, "A"),
    ("*", "*", "*", "C", True, "C"),                                 # samtools' empty-column row
    ("gggGGGg+3acgG+3ACG", "5555555I", "]]]]]]]]", "G", True, "GTT"),
]


def encoder_golden():
    from src.create_tensor_pileup_calling import decode_pileup_bases
    args = Namespace(max_indel_length=60)
    cases = []

    def run(bases, bq, mq, ref, is_cand, chunk_ref, platform):
        cand = {100: 'snv'} if is_cand else {}
        vec, _, _, _, _, alt_info = decode_pileup_bases(
            args=args, pos=100, pileup_bases=bases, reference_base=ref,
            minimum_snp_af_for_candidate=0.05, minimum_indel_af_for_candidate=0.05,
            has_pileup_candidates=True, candidates_type_dict=cand, is_tumor=True,
            mapping_quality=[ord(c) - 33 for c in mq], base_quality=[ord(c) - 33 for c in bq],
            phasing_info=None, chunk_ref_seq=chunk_ref, platform=platform)
        cases.append(dict(bases=bases, bq=bq, mq=mq, ref=ref, candidate=is_cand, chunk_ref=chunk_ref,
                          platform=platform, vec=[int(v) for v in vec], alt_info=alt_info if is_cand else None))

    for row in HAND_ROWS:
        for platform in ("ont", "ont_r10_dorado_sup_5khz", "ilmn"):
            run(*row, platform)
    # seeded synthetic rows rendered to mpileup text (with ^ / $ decorations)
    for seed, platform in ((11, 'ont'), (12, 'ilmn'), (13, 'hifi')):
        stream, aux = synth.synth_stream(4, seed, platform, depth_lo=1, depth_hi=90)
        rows = synth.render_mpileup(stream, aux, decorate_seed=seed)
        for i, text in enumerate(rows):
            cols = text.rstrip("\n").split("\t")
            ref = "ACGT"[int(stream.ref_code[i])]
            run(cols[4], cols[5], cols[6], ref, i % 3 == 0, (ref + "ACGTTGCA" * 8)[:60],
                "ont" if platform == 'ont' and i % 2 else {"ont": "ont_r10_dorado_sup_5khz", "ilmn": "ilmn_ssrs",
                                                           "hifi": "hifi_revio"}[platform])
    with open(os.path.join(HERE, "encoder_golden.json"), "w") as f:
        json.dump(cases, f)
    print("encoder_golden.json: %d cases" % len(cases))


CVT_KW = dict(num_classes=2, s1_emb_dim=16, s1_emb_kernel=3, s1_emb_stride=2, s1_proj_kernel=3, s1_kv_proj_stride=2,
              s1_heads=1, s1_depth=1, s1_mlp_mult=4, s2_emb_dim=64, s2_emb_kernel=3, s2_emb_stride=2, s2_proj_kernel=3,
              s2_kv_proj_stride=2, s2_heads=3, s2_depth=2, s2_mlp_mult=4, s3_emb_dim=128, s3_emb_kernel=3,
              s3_emb_stride=2, s3_proj_kernel=3, s3_kv_proj_stride=2, s3_heads=4, s3_depth=3, s3_mlp_mult=4,
              dropout=0., dropout_fc=0.3, depth=1, width=33, dim=34, apply_softmax=False, model_type="acgt")


def build_reference_models(n_heads, seed_aff, seed_neg):
    from clairs.model import CvT, CvT_Indel, BiGRU_NACGT, BiGRU_NACGT_Indel
    aff = (CvT if n_heads == 4 else CvT_Indel)(**CVT_KW).eval()
    neg = (BiGRU_NACGT if n_heads == 4 else BiGRU_NACGT_Indel)(apply_softmax=False, num_classes=2, channel_size=34,
                                                                 model_type="nacgt").eval()
    aff.load_state_dict(nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads), seed_aff), strict=False)
    neg.load_state_dict(nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), seed_neg), strict=False)
    return aff, neg


def synth_tensor(n, seed, deep=False):
    """Count-like int16 tensors [n,33,34] through the same generator the tests use."""
    (aff, _), (neg, _) = synth.synth_pair(n, seed, 'ont', depth_mean=120 if deep else 50, depth_hi=200)
    return aff, neg


def nn_golden():
    torch.set_num_threads(4)
    out = {}
    rng = np.random.default_rng(5)
    for n_heads in (4, 6):
        aff, neg = build_reference_models(n_heads, 100 + n_heads, 200 + n_heads)
        x = rng.integers(-60, 61, size=(24, 33, 34)).astype(np.float32)
        x[12:] *= np.float32(0.37)                     # rescaled (non-integer) inputs as after predict.py:179-197
        with torch.no_grad():
            la = torch.stack(aff(torch.from_numpy(x)), 1).numpy()
            ln = torch.stack(neg(torch.from_numpy(x)), 1).numpy()
        out["x_%d" % n_heads] = x
        out["aff_logits_%d" % n_heads] = la
        out["neg_logits_%d" % n_heads] = ln
    np.savez_compressed(os.path.join(HERE, "nn_golden.npz"), **out)
    print("nn_golden.npz:", {k: v.shape for k, v in out.items()})


CLASS_DEFAULT_CVT = dict(s1=(32, 1, 1), s2=(64, 3, 2), s3=(128, 6, 10))     # clairs/model.py:153-176 (CvT() defaults)


def nn_golden_default():
    """The CvT CLASS-DEFAULT hyper-parameters (stage-1 width 32, stage 3 with 6 heads and depth 10): SNV checkpoints are
    pickled modules whose dimensions live in the pickle (clairs/predict.py:513-517), so the default-constructed class is
    a shape the drop-in has to take.  Reference: ``CvT()`` / ``CvT_Indel()`` with no arguments."""
    from clairs.model import CvT, CvT_Indel
    torch.set_num_threads(4)
    out = {}
    rng = np.random.default_rng(6)
    for n_heads in (4, 6):
        aff = (CvT if n_heads == 4 else CvT_Indel)().eval()
        sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads, CLASS_DEFAULT_CVT), 300 + n_heads, 0.7)
        missing = aff.load_state_dict(sd, strict=False)
        assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing
        x = rng.integers(-60, 61, size=(24, 33, 34)).astype(np.float32)
        x[rng.random(x.shape) < 0.5] = 0.0
        x[12:] *= np.float32(0.41)
        with torch.no_grad():
            la = torch.stack(aff(torch.from_numpy(x)), 1).numpy()
        out["x_%d" % n_heads] = x
        out["aff_logits_%d" % n_heads] = la
    np.savez_compressed(os.path.join(HERE, "nn_golden_default.npz"), **out)
    print("nn_golden_default.npz:", {k: v.shape for k, v in out.items()}, {k: float(np.abs(v).max()) for k, v in out.items()})


def likelihood_file(path, n_heads, seed):
    rng = np.random.default_rng(seed)
    rows = [rng.uniform(0.05, 0.95, size=(10, 10)) for _ in range(n_heads)]
    edges = [np.sort(rng.uniform(0.02, 0.98, size=10)) for _ in range(2 * n_heads)]
    np.savetxt(path, np.concatenate(rows + [e[None, :] for e in edges], axis=0), fmt="%.6f")


def pipeline_golden():
    """Unmodified reference CLIs `predict` and `call_variants` on a synthetic tensor_can chunk."""
    from oracle import pileup_oracle
    work = os.path.join(HERE, "pipeline")
    os.makedirs(work, exist_ok=True)
    ctg = "chr20"
    for n_heads, tag in ((4, "snv"), (6, "indel")):
        n = 40
        (aff, aff_aux), (neg, neg_aux) = synth.synth_pair(n, 300 + n_heads, 'ont', depth_mean=60, depth_hi=160)
        ref_rng = np.random.default_rng(7)
        files = {}
        for name, stream, aux in (("aff", aff, aff_aux), ("neg", neg, neg_aux)):
            rows = synth.render_mpileup(stream, aux, ctg=ctg, first_pos=1001)
            # every candidate owns 33 consecutive rows; give candidates far-apart coordinates
            text_rows = []
            for c in range(n):
                centre = 5000 + 200 * c
                ref33 = ''.join("ACGT"[int(stream.ref_code[c * 33 + j])] for j in range(33))
                if c == 3:
                    ref33 = ref33[:16] + 'N' + ref33[17:]        # dropped by predict.py:219-220
                window, alt_info = [], None
                for j in range(33):
                    cols = rows[c * 33 + j].rstrip("\n").split("\t")
                    rb = "ACGT"[int(stream.ref_code[c * 33 + j])]
                    vec, ai = pileup_oracle.position_vector(
                        cols[4], [ord(ch) - 33 for ch in cols[6]], [ord(ch) - 33 for ch in cols[5]], rb,
                        is_candidate=(j == 16), chunk_ref_seq=(rb + "ACGTTGCA" * 8)[:60],
                        platform="ont")
                    window.append(vec)
                    if j == 16:
                        alt_info = ai
                flat = " ".join(" ".join("%d" % v for v in vec) for vec in window)
                text_rows.append("%s\t%d\t%s\t%s\t%s\t%s\t%s\n" % (ctg, centre, ref33, flat, alt_info, "unknown", ref33[16]))
            path = os.path.join(work, "tensor_can_%s_%s" % (name, tag))
            with gzip.open(path, "wt") as f:
                f.writelines(text_rows)
            files[name] = path
        aff_m, neg_m = build_reference_models(n_heads, 100 + n_heads, 200 + n_heads)
        ck_a = os.path.join("/tmp", "golden_aff_%s.pkl" % tag)
        ck_n = os.path.join("/tmp", "golden_neg_%s.pkl" % tag)
        torch.save({'model_acgt': aff_m}, ck_a)
        torch.save({'model_nacgt': neg_m}, ck_n)
        predict_fn = os.path.join(work, "predict_%s" % tag)
        env = dict(os.environ, PYTHONPATH=REF)
        subprocess.run([sys.executable, os.path.join(REF, "clairs_to.py"),
                        "predict", "--tensor_fn_acgt", files["aff"], "--tensor_fn_nacgt", files["neg"],
                        "--predict_fn", predict_fn, "--chkpnt_fn_acgt", ck_a, "--chkpnt_fn_nacgt", ck_n,
                        "--use_gpu", "False", "--platform", "ont_r10_dorado_sup_5khz", "--ctg_name", ctg, "--pileup",
                        "--disable_indel_calling", "True" if n_heads == 4 else "False"], check=True, env=env)
        lk = os.path.join(work, "likelihood_%s.txt" % tag)
        likelihood_file(lk, n_heads, 400 + n_heads)
        fai = os.path.join(work, "ref.fa.fai")
        with open(fai, "w") as f:
            f.write("%s\t64444167\t7\t60\t61\n" % ctg)
        fa = os.path.join(work, "ref.fa")
        open(fa, "a").close()
        for show_ref in (False, True):
            vcf = os.path.join(work, "call_%s%s.vcf" % (tag, "_showref" if show_ref else ""))
            cmd = [sys.executable, os.path.join(REF, "clairs_to.py"), "call_variants", "--predict_fn", predict_fn,
                   "--call_fn", vcf, "--ref_fn", fa, "--platform", "ont_r10_dorado_sup_5khz",
                   "--likelihood_matrix_data", lk, "--disable_indel_calling", "True" if n_heads == 4 else "False"]
            if show_ref:
                cmd.append("--show_ref")
            subprocess.run(cmd, check=True, env=env)
            if not os.path.exists(vcf):
                open(vcf + ".absent", "w").close()
        print("pipeline golden (%s) written" % tag)


def create_tensor_golden():
    """Unmodified reference `create_tensor_pileup_calling` end to end, with tests/fake_samtools.py
    answering `samtools faidx` / `samtools mpileup` from fixture files."""
    work = os.path.join(HERE, "create_tensor")
    os.makedirs(work, exist_ok=True)
    rng = np.random.default_rng(77)
    ctg, length = "chr20", 6000
    seq = rng.choice(list("ACGT"), size=length)
    for p in (1190, 1203, 2504):                       # IUPAC / N reference bases inside windows
        seq[p - 1] = rng.choice(list("NRY"))
    seq = "".join(seq)
    fa = os.path.join(work, "ref.fa")
    with open(fa, "w") as f:
        f.write(">%s\n" % ctg)
        for i in range(0, length, 60):
            f.write(seq[i:i + 60] + "\n")
    with open(fa + ".fai", "w") as f:
        f.write("%s\t%d\t%d\t60\t61\n" % (ctg, length, len(ctg) + 2))
    centres = [10, 1200, 1215, 1300, 2500, 2533, 3100, 4000, 4100]
    cand_fn = os.path.join(work, "%s.0_0_9_snv" % ctg)
    with open(cand_fn, "w") as f:
        for x in centres:
            f.write("%s\t%d\t%d\n" % (ctg, max(x - 17, 1), x + 17))
    covered = sorted({p for x in centres for p in range(max(x - 17, 1), x + 18)})
    dropped = {1290, 1291, 3100, 4012}                  # no pileup row: zero rows / a candidate without alt_info
    covered = [p for p in covered if p not in dropped]
    n_c = (len(covered) + 32) // 33
    neg, aux = synth.synth_stream(n_c, 555, 'ont', depth_lo=0, depth_hi=120, depth_mean=45)
    aff, aff_aux = synth.filter_min_bq(neg, 20, aux)
    bam = os.path.join(work, "tumor.bam")
    open(bam, "w").close()
    for stream, a, k in ((neg, aux, 0), (aff, aff_aux, 20)):
        rows = synth.render_mpileup(stream, a, ctg=ctg, first_pos=0, decorate_seed=5)
        with open("%s.minbq%d.mpileup" % (bam, k), "w") as f:
            for i, p in enumerate(covered):
                cols = rows[i].split("\t")
                cols[1] = str(p)
                cols[2] = seq[p - 1]
                f.write("\t".join(cols))
    shim = os.path.join(ROOT, "tests", "fake_samtools.py")
    env = dict(os.environ, PYTHONPATH=REF)
    for k, name in ((20, "aff"), (0, "neg")):
        out = os.path.join(work, "tensor_can_%s" % name)
        subprocess.run([sys.executable, os.path.join(REF, "clairs_to.py"), "create_tensor_pileup_calling",
                        "--tumor_bam_fn", bam, "--ref_fn", fa, "--ctg_name", ctg, "--samtools", shim,
                        "--min_bq", str(k), "--candidates_bed_regions", cand_fn, "--tensor_can_fn", out,
                        "--platform", "ont_r10_dorado_sup_5khz"], check=True, env=env)
        n = sum(1 for _ in gzip.open(out, "rt"))
        print("create_tensor golden (%s): %d rows" % (name, n))


def candidates_golden():
    """Unmodified reference `extract_candidates_calling` (STEP 1) through tests/fake_samtools.py:
    one chunk with SNV + indel selection and an indel BED, one SNV-only chunk of a 2-chunk split, one chunk with a
    confident BED (--bed_fn) and --bed_fn_source set (no indel BED filter)."""
    import shutil
    work = os.path.join(HERE, "candidates")
    shutil.rmtree(work, ignore_errors=True)
    os.makedirs(work)
    ctg, length, first = "chr20", 4000, 41
    rows, reference = synth.scan_rows_text(3900, 2024, ctg=ctg, first_pos=first, depth_mean=35, weird=0.004)
    rng = np.random.default_rng(9)
    seq = list(rng.choice(list("ACGT"), size=length))
    seq[first - 1:first - 1 + len(reference)] = list(reference)
    seq = "".join(seq)
    fa = os.path.join(work, "ref.fa")
    with open(fa, "w") as f:
        f.write(">%s\n" % ctg)
        for i in range(0, length, 60):
            f.write(seq[i:i + 60] + "\n")
    with open(fa + ".fai", "w") as f:
        f.write("%s\t%d\t%d\t60\t61\n" % (ctg, length, len(ctg) + 2))
    bam = os.path.join(work, "tumor.bam")
    open(bam, "w").close()
    with open(bam + ".minbq20.mpileup", "w") as f:
        f.writelines(rows)
    indel_bed = os.path.join(work, "indel_regions.bed")
    with open(indel_bed, "w") as f:
        f.write("# indel calling regions\n%s\t100\t900\n%s\t1500\t1500\n%s\t2000\t3500\nchr21\t0\t5000\n" % (ctg, ctg, ctg))
    conf_bed = os.path.join(work, "confident.bed")
    with open(conf_bed, "w") as f:
        f.write("%s\t500\t1800\n%s\t2200\t3000\n" % (ctg, ctg))
    shim = os.path.join(ROOT, "tests", "fake_samtools.py")
    env = dict(os.environ, PYTHONPATH=REF)
    common = [sys.executable, os.path.join(REF, "clairs_to.py"), "extract_candidates_calling", "--tumor_bam_fn", bam, "--ref_fn", fa,
              "--samtools", shim, "--ctg_name", ctg, "--platform", "ont_r10_dorado_sup_5khz", "--min_coverage", "4", "--min_bq", "20",
              "--output_depth", "True", "--genotyping_mode_vcf_fn", "None", "--hybrid_mode_vcf_fn", "None"]
    cases = {
        "snv_indel": ["--snv_min_af", "0.05", "--indel_min_af", "0.05", "--chunk_id", "1", "--chunk_num", "1", "--bed_fn_source", "None",
                      "--call_indels_only_in_these_regions", indel_bed, "--select_indel_candidates", "True"],
        "snv_only": ["--snv_min_af", "0.08", "--indel_min_af", "0.1", "--chunk_id", "2", "--chunk_num", "2", "--bed_fn_source", "None"],
        "bed": ["--snv_min_af", "0.05", "--indel_min_af", "0.1", "--chunk_id", "1", "--chunk_num", "2", "--bed_fn_source", conf_bed,
                "--bed_fn", conf_bed, "--call_indels_only_in_these_regions", indel_bed, "--select_indel_candidates", "True"],
    }
    for name, extra in cases.items():
        folder = os.path.join(work, name)
        os.makedirs(folder)
        res = subprocess.run(common + extra + ["--candidates_folder", folder], check=True, env=env, stdout=subprocess.PIPE, text=True)
        # the list files hold absolute paths of the build container: keep them relative to the candidates folder
        for fn in os.listdir(folder):
            if fn.startswith(("SNV_CANDIDATES_FILE_", "INDEL_CANDIDATES_FILE_")):
                path = os.path.join(folder, fn)
                with open(path) as f:
                    text = f.read().replace(folder + "/", "<candidates_folder>/")
                with open(path, "w") as f:
                    f.write(text)
        with open(os.path.join(folder, "stdout.txt"), "w") as f:
            f.write(res.stdout)
        with open(os.path.join(folder, "args.json"), "w") as f:
            json.dump([a.replace(work + "/", "") for a in extra], f)
        print("candidates golden (%s):" % name, res.stdout.strip(), sorted(os.listdir(folder)))


def indel_call_golden():
    """Unmodified reference `call_variants` on a HAND-WRITTEN predict file whose probabilities drive insertion and deletion
    calls (the synthetic-weight pipeline golden never produced an I/D ALT row): '#'-anchored insertions,
    multi-base deletions, competing alleles, depth-0 and low-AF rows."""
    work = os.path.join(HERE, "pipeline")
    src = os.path.join(work, "predict_indel")
    rows = [r.rstrip("\n").split("\t") for r in gzip.open(src, "rt")]
    rng = np.random.default_rng(31)
    alt_infos = [                                 # (reference base, alt_info as create_tensor writes it, CT:158-209)
        ("A", "40-IAGT 21 RA 10 XT 3-"),                   # insertion
        ("C", "55-DCAC 31 RC 20-"),                        # deletion of two bases
        ("G", "60-IGT 6 IGTT 16 RG 30 XA 2-"),             # competing insertion alleles
        ("T", "48-DTA 9 DTACGT 20 RT 15-"),                # competing deletion lengths
        ("A", "35-I#A 9 RA 20-"),                          # insertion anchored on a deleted base (CV:360)
        ("A", "30-RA 25 XC 5-"),                           # no indel support at all
        ("G", "0-"),                                       # empty pileup
        ("A", "80-IAACGTACGTAC 55 RA 20-"),                # long insertion
        ("C", "44-DCACGTACGTACGT 38 RC 6-"),               # long deletion
        ("G", "52-IGC 3 DGA 3 RG 40 XT 6-"),               # weak insertion and deletion plus a SNV allele
        ("T", "20-XA 10 ITA 10-"),                         # tie between a SNV and an insertion
    ]
    out_rows = []
    for k in range(60):
        cols = list(rows[k % len(rows)])
        cols[1] = str(7000 + 50 * k)
        cols[2], cols[3] = alt_infos[k % len(alt_infos)]
        probs = []
        hot = int(rng.integers(0, 6)) if k % 7 else 4 + (k // 7) % 2      # heads 4 / 5 = I / D
        for net in range(2):
            for h in range(6):
                p1 = float(rng.uniform(0.75, 0.999)) if h == hot else float(rng.uniform(0.001, 0.3))
                probs.append("%0.8f %0.8f" % (1.0 - p1, p1))
        n_fixed = len(cols) - 12
        out_rows.append("\t".join(cols[:n_fixed] + probs) + "\n")
    predict_fn = os.path.join(work, "predict_indel_hand")
    with gzip.open(predict_fn, "wt") as f:
        f.writelines(out_rows)
    env = dict(os.environ, PYTHONPATH=REF)
    for show_ref in (False, True):
        vcf = os.path.join(work, "call_indel_hand%s.vcf" % ("_showref" if show_ref else ""))
        cmd = [sys.executable, os.path.join(REF, "clairs_to.py"), "call_variants", "--predict_fn", predict_fn, "--call_fn", vcf,
               "--ref_fn", os.path.join(work, "ref.fa"), "--platform", "ont_r10_dorado_sup_5khz",
               "--likelihood_matrix_data", os.path.join(work, "likelihood_indel.txt"), "--disable_indel_calling", "False"]
        if show_ref:
            cmd.append("--show_ref")
        subprocess.run(cmd, check=True, env=env)
        n_indel = sum(1 for r in open(vcf) if r[0] != '#' and (len(r.split("\t")[3]) > 1 or len(r.split("\t")[4]) > 1))
        print("indel call golden:", vcf, "rows with an insertion / deletion ALT:", n_indel)


def hard_filter_golden():
    """Per-site hard filters: the lines the UNMODIFIED reference functions
    ``_haplotype_build_state_and_line`` (src/haplotype_filtering.py:570-703) and ``_postfilter_build_state_and_line``
    (src/postfilter_variants.py:368-446) return for the sites of two synthetic phased chunks
    (clairs_to_b200.synth.hard_filter_chunk), plus Fisher p-values (repr) and sequence entropies of the reference's helpers."""
    import gzip
    import random
    sys.path.insert(0, os.path.join(REF, "src"))
    import src.haplotype_filtering as HF
    import src.postfilter_variants as PV
    from clairs_to_b200 import synth
    out = os.path.join(HERE, "hard_filter")
    os.makedirs(out, exist_ok=True)
    meta = {}
    for name, phased, seed, kw in (("phased_long", True, 11, dict(depth=45, read_len=(150, 1200))),
                                   ("phased_short", True, 12, dict(depth=30, read_len=(60, 300))),
                                   ("unphased_short", False, 13, dict(depth=60, read_len=(60, 260)))):
        rows, ref, lo, sites = synth.hard_filter_chunk(24, seed, with_phasing=phased, **kw)
        with gzip.open(os.path.join(out, name + ".mpileup.gz"), "wt") as f:
            f.writelines(rows)
        chunk_rows = HF._parse_mpileup_to_chunk_dict(rows) if phased else PV._parse_mpileup_postfilter_chunk_dict(rows)
        lines = {}
        for disable in (False, True):
            for max_co in (3, 1):
                got = []
                for pos, rb, ab, af, het, hom in sites:
                    if phased:
                        got.append(HF._haplotype_build_state_and_line("chr20", pos, rb, ab, 100, chunk_rows, ref, lo, het, hom, disable, max_co, af, 50.0))
                    else:
                        got.append(PV._postfilter_build_state_and_line("chr20", pos, rb, ab, 100, chunk_rows, ref, lo, disable, max_co))
                lines["disable=%d,max_co=%d" % (disable, max_co)] = got
        meta[name] = dict(phased=phased, ref=ref, region_lo=lo, sites=sites, lines=lines)
    rnd = random.Random(7)
    tables = [(0, 0, 0, 0), (1, 1, 1, 1), (3, 1, 1, 3), (1, 3, 3, 1), (10, 0, 0, 10), (5, 5, 5, 6), (7, 30, 2, 41), (0, 12, 9, 3)]
    tables += [tuple(rnd.randint(0, 40) for _ in range(4)) for _ in range(150)]
    tables += [(k, k + d, k + d, k) for k in range(1, 30) for d in (0, 1, 2)]                       # mirror tables: `<=` ties
    tables += [tuple(rnd.randint(0, 2500) for _ in range(4)) for _ in range(25)] + [(1800, 1700, 1650, 1900), (4000, 3, 2, 3900)]
    meta["fisher"] = [[list(t), repr(HF.fisher_exact([[t[0], t[1]], [t[2], t[3]]]))] for t in tables]
    seqs = ["".join(rnd.choice("ACGT") for _ in range(33)) for _ in range(20)] + ["A" * 33, "AC" * 16 + "A", "ACGTN" * 6 + "RYK", "ACG" * 5]
    meta["entropy"] = [[q, repr(HF.calculate_sequence_entropy(q, entropy_window=33))] for q in seqs]
    with open(os.path.join(out, "golden.json"), "w") as f:
        json.dump(meta, f)
    print("hard filter golden:", {k: len(v["sites"]) for k, v in meta.items() if isinstance(v, dict)}, len(meta["fisher"]), "tables")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    if a.only in (None, "encoder"):
        encoder_golden()
    if a.only in (None, "nn"):
        nn_golden()
        nn_golden_default()
    if a.only in (None, "pipeline"):
        pipeline_golden()
    if a.only in (None, "create_tensor"):
        create_tensor_golden()
    if a.only in (None, "candidates"):
        candidates_golden()
    if a.only in (None, "indel_call"):
        indel_call_golden()
    if a.only in (None, "hard_filter"):
        hard_filter_golden()
