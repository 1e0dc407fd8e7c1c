"""Drop-in for the reference sub-command ``call_variants`` (clairs/call_variants.py), B200 path.

Same flags (clairs/call_variants.py:870-914), same input (gzip predict file) and the same per-chunk
VCF.  The likelihood-matrix Bayes combine + arg-max for ALL rows of the chunk runs in one fp64 CUDA
call (``cto_posterior_from_probs``, bit-identical to the python-float arithmetic of ibid. 154-304);
the string side (alt ranking, REF/ALT, GT, AD, INFO, row text; ibid. 306-618) stays on the host.

When the module is deployed inside a ClairS-TO checkout the reference's own ``shared.vcf.VcfWriter``
is used (it is out of scope and re-used unchanged); otherwise the
byte-compatible writer below is.
"""

from __future__ import annotations

import logging
import os
import shlex
import sys
from argparse import SUPPRESS, ArgumentParser
from math import e, log
from subprocess import PIPE, Popen
from time import time

import numpy as np

from . import host

logging.basicConfig(format='%(message)s', level=logging.INFO)

ZSTD = 'gzip'                      # shared/param.py:7
CALLER, VERSION = "clairs_to", "0.4.4"   # shared/param.py:2-3 (written into the VCF header)


# ---------------------------------------------------------------------------------------------
# VCF writer (layout of shared/vcf.py:14-54, 100-182)
# ---------------------------------------------------------------------------------------------
_FILTERS = [
    ("PASS", "All filters passed"),
    ("NonSomatic", "Non-somatic variant tagged by panel of normals"),
    ("LowQual", "Low-quality variant"),
    ("LowAltBQ", "Average alt allele base quality <20"),
    ("LowAltMQ", "Average alt allele read mapping quality <20"),
    ("ReadStartEnd", ">30% of the supporting alt alleles are within 100bp of the start or end of a read"),
    ("VariantCluster", "Three or more variants clustered within 200bp"),
    ("NoAncestry", "Variant without an ancestral haplotype support"),
    ("MultiHap", "Alt alleles existed in multiple haplotypes"),
    ("StrandBias", "Strand bias p-value <0.001"),
    ("LowSeqEntropy", "Sequence entropy <0.9"),
    ("Realignment", "For short-read, both the count of supporting alt alleles and AF decreased after realignment"),
    ("RefCall", "Reference call"),
]
_FLAGS = [("Verdict_Germline", "Variant tagged by verdict as Germline"),
          ("Verdict_Somatic", "Variant tagged by verdict as Somatic"),
          ("Verdict_SubclonalSomatic", "Variant tagged by verdict as Subclonal Somatic"),
          ("H", "Variant found only in one haplotype in the phased reads")]


def vcf_header_text():
    lines = ["##fileformat=VCFv4.2", "##source=ClairS-TO", "##%s_version=%s" % (CALLER, VERSION)]
    lines += ['##FILTER=<ID=%s,Description="%s">' % f for f in _FILTERS]
    lines += ['##INFO=<ID=%s,Number=0,Type=Flag,Description="%s">' % f for f in _FLAGS]
    for strand, tag in (("forward", "F"), ("reverse", "R")):
        lines += ['##INFO=<ID=%s%sU,Number=1,Type=Integer,Description="Count of %s in %s strand in the tumor BAM">'
                  % (tag, b, b, strand) for b in "ACGT"]
    lines.append('##INFO=<ID=SB,Number=1,Type=Float,Description="The p-value of Fisher’s exact test on strand bias">')
    lines += ['##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">',
              '##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype quality">',
              '##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read depth">',
              '##FORMAT=<ID=AF,Number=1,Type=Float,Description="Estimated allele frequency">',
              '##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Allelic depths for the ref and alt alleles in the '
              'order listed in the ALT column">']
    lines += ['##FORMAT=<ID=%sU,Number=1,Type=Integer,Description="Count of %s in the tumor BAM">' % (b, b) for b in "ACGT"]
    return "\n".join(lines) + "\n"


class ChunkVcfWriter:
    def __init__(self, vcf_fn, ctg_name=None, ref_fn=None, sample_name="SAMPLE", show_ref_calls=False):
        folder = os.path.dirname(vcf_fn)
        if not os.path.exists(folder):
            print("[INFO] Output VCF folder {} not found, create it".format(folder))
            os.makedirs(folder, exist_ok=True)
        self.fp = open(vcf_fn, 'w')
        self.show_ref_calls = show_ref_calls
        self.ctg_name = ctg_name
        names = None if ctg_name is None else (ctg_name.split(',') if ',' in ctg_name else [ctg_name])
        header = vcf_header_text()
        if ref_fn is not None:
            fai = ref_fn + ".fai"
            if not os.path.exists(fai):
                sys.exit("[ERROR] file %s not found" % fai)
            with open(fai) as f:
                for row in f:
                    cols = row.strip().split("\t")
                    if names is not None and cols[0] not in names:
                        continue
                    header += "##contig=<ID=%s,length=%s>\n" % (cols[0], cols[1])
        header += '#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n' % sample_name
        self.fp.write(header)

    def write_row(self, CHROM, POS, REF, ALT, QUAL, FILTER, INFO, GT, DP, AF, AD, AU, CU, GU, TU):
        if not self.show_ref_calls and GT in ("0/0", "./."):
            return
        gq = int(float(QUAL))
        self.fp.write("%s\t%d\t.\t%s\t%s\t%.4f\t%s\t%s\tGT:GQ:DP:AF:AD:AU:CU:GU:TU\t%s:%d:%d:%.4f:%s:%d:%d:%d:%d\n"
                      % (CHROM, int(POS), REF, ALT, QUAL, FILTER, INFO, GT, gq, DP, AF, AD, AU, CU, GU, TU))

    def close(self):
        self.fp.close()


def _make_writer(args):
    try:                                     # deployed inside a ClairS-TO tree: re-use its writer unchanged
        from shared.vcf import VcfWriter
        return VcfWriter(vcf_fn=args.call_fn, ref_fn=args.ref_fn, ctg_name=args.ctg_name,
                         show_ref_calls=args.show_ref, sample_name=args.sample_name)
    except ImportError:
        return ChunkVcfWriter(args.call_fn, ctg_name=args.ctg_name, ref_fn=args.ref_fn,
                              sample_name=args.sample_name, show_ref_calls=args.show_ref)


# ---------------------------------------------------------------------------------------------
# host-side call logic
# ---------------------------------------------------------------------------------------------
_PHRED = -10 * log(e, 10)


def quality_score_from(p):
    """clairs/call_variants.py:81-88."""
    p = float(p)
    return float(round(max(_PHRED * log(((1.0 - p) + 1e-10) / (p + 1e-10)) + 2.0, 0.0), 4))


def parse_alt_info(text):
    """clairs/call_variants.py:135-149."""
    parts = text.rstrip().split('-')
    depth = int(parts[0])
    seqs = (parts[1] if len(parts) > 1 else '').split(' ')
    alts = dict(zip(seqs[::2], [int(v) for v in seqs[1::2]])) if len(seqs) else {}
    if depth == 0 and len(alts) == 1:
        for k, v in alts.items():
            if k[0] in 'DI':
                depth = int(v)
    return alts, depth


def call_row(chrom, pos, ref, alt_info, fwd_text, rev_text, post, head, snv_mode, show_ref, qual_for_pass, writer):
    """One predict row -> at most one VCF row (clairs/call_variants.py:111-618 after the combine)."""
    alts, depth = parse_alt_info(alt_info)
    pmax = float(max(post))
    is_variant = ("ACGT"[head] != ref) if snv_mode else head >= 4
    is_reference = not is_variant
    alt = ref
    supported = None
    if is_variant:
        if depth <= 0:
            print("low tumor coverage")
            return
        support = {a: c / float(depth) for a, c in alts.items() if a[0] != 'R' and c / float(depth) > 0}
        if not support:
            return
        ranked = [a for a, _ in sorted(support.items(), key=lambda x: x[1], reverse=True)]
        best = ranked[0]
        supported = alts[best]
        observed = [a[1] for a in ranked if a[0] == 'X']
        if best[0] == 'X':
            alt = best[1]
            if snv_mode and "ACGT"[head] not in observed:            # ibid. 350-358
                is_variant, is_reference = False, True
        elif best[0] == 'I':
            alt = best[1:] if best[1] != '#' else ref + best[2:]
        elif best[0] == 'D':
            alt = ref
            ref = ref + best[2:]
    if (not show_ref and is_reference) or (not is_reference and ref == alt):
        return
    if (len(ref) > 1 or len(alt) > 1) and snv_mode:
        return
    if not snv_mode and len(ref) == 1 and len(alt) == 1 and not show_ref:
        return
    ref_num = 0
    for a, c in alts.items():
        if a[0] == 'R':
            ref_num = int(c)
    if is_reference:
        supported = ref_num
        alt = "."
    af = min((supported / depth) if depth != 0 else 0.0, 1.0)
    gt = '0/0' if is_reference else ("0/1" if af < 1.0 else '1/1')
    qual = quality_score_from(pmax)
    if is_reference:
        filt = 'RefCall'
    elif qual_for_pass is None or qual >= float(qual_for_pass):
        filt = 'PASS'
    else:
        filt = 'LowQual'
    fwd, rev = eval(fwd_text), eval(rev_text)                          # list reprs written by predict (ibid. 588-589)
    f, r = [int(v) for v in fwd[:4]], [int(v) for v in rev[:4]]
    tot = [int(x + y) for x, y in zip(fwd, rev)]
    info = "FAU={};FCU={};FGU={};FTU={};RAU={};RCU={};RGU={};RTU={}".format(*(f + r))
    ad = str(supported) if is_reference else "%s,%s" % (ref_num, supported)
    writer.write_row(CHROM=chrom, POS=pos, REF=ref, ALT=alt, QUAL=qual, FILTER=filt, INFO=info, GT=gt, DP=depth,
                     AF=af, AD=ad, AU=tot[0], CU=tot[1], GU=tot[2], TU=tot[3])


def emit_calls(rows, post, call, snv_mode, show_ref, qual, writer):
    """rows.field(k, 0..5) = chrom, pos, ref, alt_info, forward / reverse strand list-reprs of candidate k."""
    for k in range(len(post)):
        chrom, pos = rows.field(k, 0), rows.field(k, 1)
        if call[k] >> 8:
            # the reference indexes a 10x10 matrix with bin 10 here and dies with IndexError
            sys.exit("[ERROR] probability of exactly 1.0 at %s:%s falls outside the likelihood bins" % (chrom, pos))
        call_row(chrom, pos, rows.field(k, 2), rows.field(k, 3), rows.field(k, 4), rows.field(k, 5), post[k],
                 int(call[k] & 0xFF), snv_mode, show_ref, qual, writer)


def finish_vcf(call_fn):
    """clairs/call_variants.py:858-867: a VCF without records is removed."""
    if os.path.exists(call_fn):
        content = open(call_fn).readlines()
        if not len(content):
            os.remove(call_fn)
        for row in content:
            if row[0] != '#':
                return
        logging.info("[INFO] No vcf output in file {}, remove.".format(call_fn))
        os.remove(call_fn)


def call_variants_from_probability(args):
    import ctypes as C
    import torch
    from . import _lib
    from .weights import likelihood_tables

    if not torch.cuda.is_available():
        sys.exit("[ERROR] no CUDA device: the B200 call_variants has no CPU fallback")
    snv_mode = bool(args.disable_indel_calling)
    n_heads = 4 if snv_mode else 6
    if args.call_fn != "PIPE":
        call_dir = os.path.dirname(args.call_fn)
        if call_dir and not os.path.exists(call_dir):
            os.makedirs(call_dir, exist_ok=True)
    writer = _make_writer(args)
    path = args.predict_fn
    logging.info("[INFO] Calling tumor-only somatic variants from {} ...".format(path.split('/')[-1]))
    start = time()
    if path != "PIPE":
        if not os.path.exists(path):
            print("[ERROR] Prediction path not found!")
            return
        proc = Popen(shlex.split("%s -fdc %s" % (ZSTD, path)), stdout=PIPE, bufsize=8388608)
        text = proc.stdout.read()
        proc.stdout.close()
        proc.wait()
    else:
        text = sys.stdin.buffer.read()
    tables = np.ascontiguousarray(likelihood_tables(args.likelihood_matrix_data, n_heads))
    # the whole file by one native parse (clairs/call_variants.py:798-829 loops over rows and float()s every field)
    pf = host.PredictFile(text, n_heads)

    if pf.n:
        lib = _lib.lib()
        dev = torch.device('cuda', torch.cuda.current_device())
        d_pa = torch.from_numpy(pf.p_aff).to(dev)
        d_pn = torch.from_numpy(pf.p_neg).to(dev)
        d_post = torch.empty((pf.n, n_heads), dtype=torch.float64, device=dev)
        d_call = torch.empty((pf.n,), dtype=torch.int32, device=dev)
        _lib.check(lib.cto_posterior_from_probs(C.c_void_p(tables.ctypes.data), n_heads, C.c_void_p(d_pa.data_ptr()),
                                                C.c_void_p(d_pn.data_ptr()), pf.n, C.c_void_p(d_post.data_ptr()),
                                                C.c_void_p(d_call.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "cto_posterior_from_probs")
        emit_calls(pf, d_post.cpu().numpy(), d_call.cpu().numpy(), snv_mode, args.show_ref, args.qual, writer)

    logging.info("[INFO] Total time elapsed: %.2f s" % (time() - start))
    writer.close()
    finish_vcf(args.call_fn)


def build_parser():
    from .predict import str2bool
    parser = ArgumentParser(description="Call variants using trained models and tensors of candidate variants (B200 engine)")
    parser.add_argument('--platform', type=str, default="ont")
    parser.add_argument('--call_fn', type=str, default=None)
    parser.add_argument('--ref_fn', type=str, default=None)
    parser.add_argument('--ctg_name', type=str, default=None)
    parser.add_argument('--sample_name', type=str, default="SAMPLE")
    parser.add_argument('--qual', type=int, default=0)
    parser.add_argument('--samtools', type=str, default="samtools")
    parser.add_argument('--show_ref', action='store_true')
    parser.add_argument('--likelihood_matrix_data', type=str, default=None)
    parser.add_argument('--disable_indel_calling', type=str2bool, default=0)
    parser.add_argument('--predict_fn', type=str, default="PIPE")
    parser.add_argument('--pileup', action='store_true', help=SUPPRESS)
    return parser


def main(argv=None):
    call_variants_from_probability(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
