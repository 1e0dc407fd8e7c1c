"""The per-chunk hot path of ``run_clairs_to`` in ONE process and in memory:

    create_tensor_pileup_calling (AFF stream, --min_bq <platform>)  \\
    create_tensor_pileup_calling (NEG stream, --min_bq 0)            >  predict  >  call_variants  ->  per-chunk VCF
                                                                    /
(run_clairs_to:1228-1308 SNV, 1562-1647 indel: three sub-commands that hand gzip text files to each other).  Here the
pileup tensors go from the encoder kernel to the networks and the posterior kernel without leaving the GPU, and only the
strings a VCF row needs come back to the host.  Same inputs as the sub-commands, same VCF (the test suite checks it is
byte-identical to running the three drop-in sub-commands through their files); the intermediate tensor_can / predict files
are still written, byte-identically, when their paths are given.

    python -m clairs_to_b200.hot_path --tumor_bam_fn T.bam --ref_fn ref.fa --ctg_name chr20 --samtools samtools \\
        --candidates_bed_regions chr20.0_0_9_snv --min_bq 20 --platform ont_r10_dorado_sup_5khz \\
        --chkpnt_fn_acgt pileup_affirmative.pkl --chkpnt_fn_nacgt pileup_negational.pkl \\
        --likelihood_matrix_data likelihood_matrix.txt --disable_indel_calling True --call_fn vcf_output/p_0.vcf
"""

from __future__ import annotations

import os
import shlex
import sys
from argparse import ArgumentParser
from subprocess import PIPE, Popen

import numpy as np

from . import host
from .call_variants import _make_writer, emit_calls, finish_vcf
from .create_tensor_pileup_calling import ZSTD, encode_chunk


class _Rows:
    """Row accessor with the interface of host.PredictFile over in-memory lists."""

    def __init__(self, cols):
        self.cols = cols

    def field(self, k, f):
        return self.cols[f][k]


def _gz_write(path, payload):
    d = os.path.dirname(path)
    if d and not os.path.exists(d):
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as fpo:
        zp = Popen(shlex.split("%s -c" % ZSTD), stdin=PIPE, stdout=fpo, bufsize=8388608)
        zp.stdin.write(payload)
        zp.stdin.close()
        zp.wait()


def run_chunk(args, engine=None, host_threads=0):
    """One chunk: returns the number of candidates predicted.  ``engine``: reuse a loaded Engine across chunks (one process
    per GPU iterating over chunk files) instead of paying the checkpoint load per chunk like the reference does."""
    import torch
    from .engine import Engine
    from .predict import CENTER

    snv_mode = bool(args.disable_indel_calling)
    n_heads = 4 if snv_mode else 6
    own = engine is None
    if own:
        engine = Engine.from_checkpoints(args.chkpnt_fn_acgt, args.chkpnt_fn_nacgt, max_batch=10240)
    if engine.n_heads != n_heads:
        sys.exit("[ERROR] checkpoints carry %d heads but --disable_indel_calling %s expects %d"
                 % (engine.n_heads, args.disable_indel_calling, n_heads))
    if not engine.has_likelihood or own:
        engine.set_likelihood(args.likelihood_matrix_data)
    # the two streams of the chunk (run_clairs_to:1230-1271); Illumina / HiFi use --min-BQ 0 for both, where the reference
    # symlinks the NEG tensor file to the AFF one (ibid. 1248-1252)
    aff = encode_chunk(args, min_bq=args.min_bq, host_threads=host_threads)
    neg = aff if args.min_bq == 0 else encode_chunk(args, min_bq=0, host_threads=host_threads)
    if args.tensor_can_fn_acgt:
        _gz_write(args.tensor_can_fn_acgt, host.format_tensor_can_rows(aff.ctg, aff.pos, aff.ref33, aff.tensor.cpu().numpy(),
                                                                     aff.alt_info, aff.variant_type) if len(aff) else b"")
    if args.tensor_can_fn_nacgt:
        _gz_write(args.tensor_can_fn_nacgt, host.format_tensor_can_rows(neg.ctg, neg.pos, neg.ref33, neg.tensor.cpu().numpy(),
                                                                      neg.alt_info, neg.variant_type) if len(neg) else b"")
    call_dir = os.path.dirname(args.call_fn)
    if call_dir and not os.path.exists(call_dir):
        os.makedirs(call_dir, exist_ok=True)
    writer = _make_writer(args)
    n = 0
    # predict drops rows whose centre reference base is not ACGT (clairs/predict.py:219-220) in each file and then pairs the
    # two files by row index (ibid. 586-588, 613-620)
    keep_a = [k for k in range(len(aff)) if aff.ref33[k][CENTER] in "ACGT"]
    keep_n = [k for k in range(len(neg)) if neg.ref33[k][CENTER] in "ACGT"]
    n = min(len(keep_a), len(keep_n))
    if n:
        ia = torch.as_tensor(keep_a[:n], device=engine.device)
        inn = torch.as_tensor(keep_n[:n], device=engine.device)
        res = engine.predict(aff.tensor[ia], aff.depth[ia], neg.tensor[inn], neg.depth[inn], posterior=True)
        probs = res['probs'].cpu().numpy()
        fwd, rev = res['fwd'].cpu().numpy(), res['rev'].cpu().numpy()
        cols = [[aff.ctg] * n, [str(int(aff.pos[k])) for k in keep_a[:n]], [aff.ref33[k][CENTER].upper() for k in keep_a[:n]],
                [aff.alt_info[k] for k in keep_a[:n]], [str([float(v) for v in fwd[k]]) for k in range(n)],
                [str([float(v) for v in rev[k]]) for k in range(n)]]
        if args.predict_fn:
            rows = []
            for k in range(n):
                fields = [c[k] for c in cols] + [host.format_prob_fields(probs[k])]
                if n_heads == 4:
                    fields.append("")
                rows.append("\t".join(fields) + "\n")
            _gz_write(args.predict_fn, "".join(rows).encode())
        # the posterior kernel already combined the probabilities exactly as call_variants would after re-reading their
        # 8-decimal text (clairs/call_variants.py:154-304), so the calls equal the file-based path bit for bit
        emit_calls(_Rows(cols), res['post'].cpu().numpy(), res['call'].cpu().numpy(), snv_mode, args.show_ref, args.qual, writer)
    writer.close()
    finish_vcf(args.call_fn)
    if own:
        engine.close()
    return n


def build_parser():
    from .predict import str2bool
    p = ArgumentParser(description="create_tensor x2 -> predict -> call_variants of one chunk in one process (B200 engine)")
    for flag in ("--tumor_bam_fn", "--ref_fn", "--ctg_name", "--candidates_bed_regions", "--chkpnt_fn_acgt", "--chkpnt_fn_nacgt",
                 "--likelihood_matrix_data", "--call_fn"):
        p.add_argument(flag, type=str, required=True)
    p.add_argument('--samtools', type=str, default="samtools")
    p.add_argument('--platform', type=str, default="ont")
    p.add_argument('--min_bq', type=int, required=True, help="--min-BQ of the AFF stream (shared/param.py:34); the NEG stream uses 0")
    p.add_argument('--max_depth', type=int, default=None)
    p.add_argument('--max_indel_length', type=int, default=None)
    p.add_argument('--sample_name', type=str, default="SAMPLE")
    p.add_argument('--qual', type=int, default=0)
    p.add_argument('--show_ref', action='store_true')
    p.add_argument('--disable_indel_calling', type=str2bool, default=0)
    p.add_argument('--tensor_can_fn_acgt', type=str, default=None, help="also write the AFF tensor_can chunk file")
    p.add_argument('--tensor_can_fn_nacgt', type=str, default=None, help="also write the NEG tensor_can chunk file")
    p.add_argument('--predict_fn', type=str, default=None, help="also write the predict chunk file")
    return p


def main(argv=None):
    run_chunk(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
