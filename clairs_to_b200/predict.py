"""Drop-in for the reference sub-command ``predict`` (clairs/predict.py), B200 path.

Same flags (clairs/predict.py:740-809), same inputs (two gzip tensor_can chunk files consumed in
lock-step, ibid. 586-620) and same output (gzip predict file, ibid. 114-152).  What changes is where
the arithmetic runs: rows are parsed by the native codec, the depth rescale, AFF + NEG forward,
softmax and strand-count recovery run in one ``cto_predict`` call per batch on the GPU.

Behaviour kept on purpose: rows whose centre reference base is not ACGT are
dropped (ibid. 219-220); the centre index is 16; the SNV row has a trailing empty field, the indel
row does not; ``--call_fn`` is not supported (it is unusable in the reference, ibid. 89-111).
``--use_gpu`` is accepted and ignored: this engine has no CPU path.
"""

from __future__ import annotations

import logging
import os
import shlex
import sys
from argparse import SUPPRESS, ArgumentParser
from subprocess import PIPE, Popen
from time import time

import numpy as np

from . import host
from .pileup_format import CENTER, N_CH, N_POS

logging.basicConfig(format='%(message)s', level=logging.INFO)

ZSTD = 'gzip'                 # shared/param.py:7
PREDICT_BATCH = 16384         # candidates per GPU call (the reference uses 250, shared/param.py:85)


def str2bool(v):
    """shared/utils.py:121-131 (including its accepted spellings)."""
    if v is None or isinstance(v, bool):
        return v
    if v.lower() in ('yes', 'ture', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'flase', 'false', 'f', 'n', '0'):
        return False
    import argparse
    raise argparse.ArgumentTypeError('Boolean value expected.')


def _open_reader(path):
    if path == "PIPE":
        return None, sys.stdin
    proc = Popen(shlex.split("%s -fdc %s" % (ZSTD, path)), stdout=PIPE, bufsize=8388608, universal_newlines=True)
    return proc, proc.stdout


def read_tensor_file(path) -> "host.TensorFile":
    """A whole tensor_can chunk file (<= 10 000 candidates, shared/param.py:21) decompressed by the same ``gzip -fdc``
    as the reference (clairs/predict.py:159) and parsed by ONE native call instead of a Python loop over rows."""
    if path == "PIPE":
        text = sys.stdin.buffer.read()
    else:
        proc = Popen(shlex.split("%s -fdc %s" % (ZSTD, path)), stdout=PIPE, bufsize=8388608)
        text = proc.stdout.read()
        proc.stdout.close()
        if proc.wait() != 0:
            sys.exit("[ERROR] %s -fdc %s failed" % (ZSTD, path))
    return host.TensorFile(text)


def read_tensor_rows(path):
    """Yield (contig, pos, ref33, int16[33,34], alt_info, variant_type, ref_centre) for rows that
    survive the centre-base filter (clairs/predict.py:172-175, 219-220)."""
    proc, fo = _open_reader(path)
    for row in fo:
        cols = row.split("\t")[:7]
        if len(cols) < 7:
            continue
        contig, coord, seq, tensor_text, alt_info, variant_type, ref_center = cols
        if seq[CENTER] not in "ACGT":
            continue
        yield contig, coord, seq, host.parse_tensor_row(tensor_text), alt_info, variant_type, ref_center.strip()
    if proc is not None:
        fo.close()
        proc.wait()


def _batches(it, size):
    batch = []
    for item in it:
        batch.append(item)
        if len(batch) == size:
            yield batch
            batch = []
    if batch:
        yield batch


def format_rows(meta, fwd, rev, probs, n_heads):
    """clairs/predict.py:114-152: one text row per candidate."""
    out = []
    for k, (contig, coord, seq, _, alt_info, _, _) in enumerate(meta):
        fields = [contig, coord, seq[CENTER].upper(), alt_info,
                  str([float(v) for v in fwd[k]]), str([float(v) for v in rev[k]]),
                  host.format_prob_fields(probs[k])]
        if n_heads == 4:
            fields.append("")
        out.append("\t".join(fields) + "\n")
    return out


def predict(args):
    import torch
    from .engine import Engine

    if args.call_fn is not None:
        sys.exit("[ERROR] --call_fn is not supported by the B200 predict (unusable in the reference as well); "
                 "use --predict_fn followed by call_variants")
    if args.flanking is not None and args.flanking != 16:
        sys.exit("[ERROR] --flanking %d: the pileup models are built for 16 flanking bases" % args.flanking)
    if not args.pileup:
        sys.exit("[ERROR] only --pileup tensors are on the B200 path (run_clairs_to always passes --pileup)")
    if not torch.cuda.is_available():
        sys.exit("[ERROR] no CUDA device: the B200 predict has no CPU fallback")
    start = time()
    if args.is_from_tables:
        return                                                   # clairs/predict.py:575: nothing to do
    engine = Engine.from_checkpoints(args.chkpnt_fn_acgt, args.chkpnt_fn_nacgt, max_batch=10240)     # one reference chunk file (<= 10 000 candidates, shared/param.py:21) per network pass
    expect_heads = 4 if args.disable_indel_calling else 6
    if engine.n_heads != expect_heads:
        sys.exit("[ERROR] checkpoints carry %d heads but --disable_indel_calling %s expects %d"
                 % (engine.n_heads, args.disable_indel_calling, expect_heads))

    predict_fn = args.predict_fn
    if predict_fn != "PIPE":
        predict_dir = os.path.dirname(predict_fn)
        if predict_dir and not os.path.exists(predict_dir):
            os.makedirs(predict_dir, exist_ok=True)
        fpo = open(predict_fn, "wb")
        zproc = Popen(shlex.split("%s -c" % ZSTD), stdin=PIPE, stdout=fpo, bufsize=8388608)
        out_file = zproc.stdin
    else:
        fpo, zproc, out_file = None, None, sys.stdout.buffer

    # whole chunk files: one native parse per file, one GPU call and one native format per PREDICT_BATCH rows.
    # The two files are consumed in lock-step by row index, so min(rows) rows are predicted.
    aff_tf = read_tensor_file(args.tensor_fn_acgt)
    neg_tf = read_tensor_file(args.tensor_fn_nacgt)
    n_rows = min(aff_tf.n, neg_tf.n)
    dev = engine.device
    total = 0
    for r0 in range(0, n_rows, PREDICT_BATCH):
        n = min(PREDICT_BATCH, n_rows - r0)
        xa = torch.from_numpy(aff_tf.tensor[r0:r0 + n]).to(dev)
        xn = torch.from_numpy(neg_tf.tensor[r0:r0 + n]).to(dev)
        da = torch.from_numpy(aff_tf.depth[r0:r0 + n]).to(dev)
        dn = torch.from_numpy(neg_tf.depth[r0:r0 + n]).to(dev)
        res = engine.predict(xa, da, xn, dn, posterior=False)
        probs = res['probs'].cpu().numpy()
        fwd, rev = res['fwd'].cpu().numpy(), res['rev'].cpu().numpy()
        out_file.write(host.format_predict_rows(aff_tf, r0, n, fwd, rev, probs, engine.n_heads))
        if total // 20000 != (total + n) // 20000:
            print("Processed %d tensors" % ((total + n) // 20000 * 20000), file=sys.stderr)
        total += n

    logging.info("[INFO] {} total processed positions: {}, time elapsed: {}".format(
        args.ctg_name, total, "%.1fs" % (time() - start)))
    if zproc is not None:
        zproc.stdin.close()
        zproc.wait()
        fpo.close()
    engine.close()


def build_parser():
    parser = ArgumentParser(description="Candidate variants probability prediction using tensors and trained models "
                                        "(B200 engine)")
    parser.add_argument('--platform', type=str, default="ont")
    parser.add_argument('--tensor_fn_acgt', type=str, default="PIPE")
    parser.add_argument('--tensor_fn_nacgt', type=str, default="PIPE")
    parser.add_argument('--chkpnt_fn_acgt', type=str, default=None)
    parser.add_argument('--chkpnt_fn_nacgt', type=str, default=None)
    parser.add_argument('--call_fn', type=str, default=None)
    parser.add_argument('--ref_fn', type=str, default=None)
    parser.add_argument('--ctg_name', type=str, default=None)
    parser.add_argument('--sample_name', type=str, default="SAMPLE")
    parser.add_argument('--samtools', type=str, default="samtools")
    parser.add_argument('--show_ref', action='store_true')
    parser.add_argument('--min_rescale_cov', type=int, default=50)
    parser.add_argument('--disable_indel_calling', type=str2bool, default=0)
    parser.add_argument('--predict_fn', type=str, default="PIPE")
    parser.add_argument('--use_gpu', type=str2bool, default=False, help=SUPPRESS)
    parser.add_argument('--qual', type=int, default=0, help=SUPPRESS)
    parser.add_argument('--pileup', action='store_true', help=SUPPRESS)
    parser.add_argument('--is_from_tables', type=str2bool, default=False, help=SUPPRESS)
    parser.add_argument('--flanking', type=int, default=None, help=SUPPRESS)
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.min_rescale_cov != 50:
        sys.exit("[ERROR] --min_rescale_cov %d: the device rescale is fixed at the reference default 50 "
                 "(shared/param.py:26)" % args.min_rescale_cov)
    predict(args)


if __name__ == "__main__":
    main()
