"""Drop-in for the reference sub-command ``create_tensor_pileup_calling``
(src/create_tensor_pileup_calling.py), B200 path.

Same flags (ibid. 582-673), same external tools (``samtools faidx`` / ``samtools mpileup`` with the
identical command line, ibid. 426-446; ``gzip`` for the chunk file) and a byte-identical tensor_can
chunk file (ibid. 561-569).  What changes: the mpileup text is tokenised by the native host tokenizer
(``cto_tokenize_mpileup``), the 34-channel vectors and the 33-row windows are produced by the CUDA
encoder (``cto_encode_pileup``), and the 1122-int text field is written by the native formatter.

Only the mode ``run_clairs_to`` uses is on this path: ``--candidates_bed_regions`` given (ibid.
347-370).  The VCF / whole-contig / phasing modes of the script are rejected with a message.
"""

from __future__ import annotations

import os
import shlex
import sys
from argparse import SUPPRESS, ArgumentParser
from subprocess import PIPE, Popen

import numpy as np

from . import host
from .pileup_format import N_POS

ZSTD = 'gzip'                     # shared/param.py:7
FLANK = 16                        # shared/param.py:59
EXPAND_REFERENCE = 1000           # shared/param.py:81
MAX_INDEL_LENGTH = 60             # shared/param.py:101
SAMTOOLS_FILTER_FLAG = 2316       # shared/param.py:27


def get_chunk_id(path):
    """src/create_tensor_pileup_calling.py:70-79."""
    try:
        chunk_id, index, bin_size = path.split('.')[-1].split('_')
        return "chunk {}-{}/{}".format(str(int(index) + 1), bin_size, chunk_id)
    except Exception:
        return ""


def _popen(cmd, **kw):
    return Popen(shlex.split(cmd), bufsize=8388608, universal_newlines=True, **kw)


def read_candidates(path, ctg_name):
    """ibid. 347-370: region rows -> {centre position: variant type}, plus the covered range."""
    proc = _popen("%s -fdc %s" % (ZSTD, path), stdout=PIPE)
    cand = {}
    ctg_start, ctg_end = float('inf'), 0
    for row in proc.stdout:
        cols = row.rstrip().split('\t')
        if cols[0] != ctg_name:
            continue
        position, end = int(cols[1]) + 1, int(cols[2]) + 1
        ctg_start, ctg_end = min(position, ctg_start), max(end, ctg_end)
        centre = end - FLANK - 2 if position < 1 else position + (end - position) // 2 - 1
        cand[centre] = cols[3] if len(cols) == 4 else 'unknown'
    proc.stdout.close()
    proc.wait()
    return cand, ctg_start, ctg_end


def reference_sequence_from(samtools, fasta, region):
    """shared/utils.py:148-174."""
    proc = _popen("%s faidx %s %s" % (samtools, fasta, region), stdout=PIPE)
    lines = [row.rstrip() for row in proc.stdout]
    proc.stdout.close()
    proc.wait()
    if proc.returncode != 0:
        return None
    return "".join(lines[1:]).upper()


class EncodedChunk:
    """The candidates of one chunk after the encoder, still in memory: ``pos`` int64 [n], ``ref33`` list of 33-mers,
    ``tensor`` int16 [n,33,34] and ``depth`` int32 [n] on the device, ``alt_info`` / ``variant_type`` lists."""

    def __init__(self, ctg, pos, ref33, tensor, depth, alt_info, variant_type):
        self.ctg, self.pos, self.ref33, self.tensor, self.depth = ctg, pos, ref33, tensor, depth
        self.alt_info, self.variant_type = alt_info, variant_type

    def __len__(self):
        return len(self.pos)


def encode_chunk(args, min_bq=None, host_threads=0):
    """samtools mpileup -> native tokenizer -> window table -> CUDA encoder for the candidates of one chunk
    (src/create_tensor_pileup_calling.py:306-570 up to the row text).  Returns an EncodedChunk (possibly empty)."""
    import torch
    from .engine import PIPELINE_LOW_BQ_CUT, encode_pileup, stream_to_device

    if not args.candidates_bed_regions:
        sys.exit("[ERROR] the B200 create_tensor_pileup_calling needs --candidates_bed_regions "
                 "(the only mode run_clairs_to uses)")
    if getattr(args, 'vcf_fn', None) or getattr(args, 'truth_vcf_fn', None) or getattr(args, 'phase_tumor', None) or \
            getattr(args, 'extend_bed', None) or getattr(args, 'bed_fn', None):
        sys.exit("[ERROR] --vcf_fn/--truth_vcf_fn/--phase_tumor/--extend_bed/--bed_fn are not on the B200 path")
    if getattr(args, 'flanking', None) is not None and args.flanking != FLANK:
        sys.exit("[ERROR] --flanking %d: the pileup models are built for 16 flanking bases" % args.flanking)
    if not torch.cuda.is_available():
        sys.exit("[ERROR] no CUDA device: the B200 encoder has no CPU fallback")
    ctg_name = args.ctg_name
    max_indel_length = MAX_INDEL_LENGTH if getattr(args, 'max_indel_length', None) is None else args.max_indel_length
    min_bq = args.min_bq if min_bq is None else min_bq
    fai = args.ref_fn + ".fai"
    if not os.path.exists(fai):
        sys.exit("[ERROR] file %s not found" % fai)

    cand, ctg_start, ctg_end = read_candidates(args.candidates_bed_regions, ctg_name)
    cand = {p: t for p, t in cand.items() if ctg_start <= p <= ctg_end}
    if not cand:
        ctg_start, ctg_end = 1, 1
    extend_start = max(1, ctg_start - N_POS)
    extend_end = ctg_end + N_POS
    reference_start = max(1, ctg_start - EXPAND_REFERENCE)
    reference_end = ctg_end + EXPAND_REFERENCE
    reference = reference_sequence_from(args.samtools, args.ref_fn, "%s:%d-%d" % (ctg_name, reference_start, reference_end))
    if reference is None or len(reference) == 0:
        sys.exit("[ERROR] Failed to load reference sequence from file ({}).".format(args.ref_fn))

    # the exact command of ibid. 426-446 (no -f: bases are literal letters)
    cmd = "{} mpileup --reverse-del".format(args.samtools) + ' --output-MQ ' + \
          ' -r {}:{}-{}'.format(ctg_name, extend_start, extend_end) + ' --min-MQ 0' + \
          ' --min-BQ {}'.format(min_bq) + ' -l {}'.format(args.candidates_bed_regions) + \
          ' --excl-flags {}'.format(SAMTOOLS_FILTER_FLAG) + \
          (' --max-depth {}'.format(args.max_depth) if getattr(args, 'max_depth', None) is not None else "")
    mp = Popen(shlex.split(cmd + '  ' + args.tumor_bam_fn), stdout=PIPE, stderr=PIPE, bufsize=8388608)
    text, _ = mp.communicate()

    tok = host.tokenize_mpileup(text, reference, reference_start, sorted(cand), max_indel_length, n_threads=host_threads)
    # window assembly (ibid. 513-516, 537-553), vectorised: rows are position sorted, so the row of a position is a
    # binary search; -1 = no pileup row (an all-zero row, ibid. 461)
    table_len = extend_end - extend_start
    row_pos = np.asarray(tok.row_pos, dtype=np.int64)
    cpos = np.asarray(sorted(cand), dtype=np.int64)
    start = cpos - FLANK - extend_start
    ok = (start >= 0) & (start + N_POS < table_len)                      # ibid. 542-543
    if row_pos.size:
        ci = np.searchsorted(row_pos, cpos)
        ok &= (ci < row_pos.size) & (row_pos[np.minimum(ci, row_pos.size - 1)] == cpos)   # no alt_info for it, ibid. 552-553
    else:
        ok &= False
    keep = cpos[ok]
    # a candidate within 16 bp of a contig end has a reference context shorter than 33 bases; the reference still writes
    # such a row (ibid. 561) and its predict drops it for the centre-base test (clairs/predict.py:219-220) -- or crashes on
    # the index.  It cannot become a call either way, so it is dropped here (used to abort the whole chunk).
    ref33 = [reference[int(p) - reference_start - FLANK: int(p) - reference_start + FLANK + 1].upper() for p in keep]
    full = np.array([len(r) == N_POS for r in ref33], dtype=bool)
    if not full.all():
        keep = keep[full]
        ref33 = [r for r in ref33 if len(r) == N_POS]
    if not keep.size:
        return EncodedChunk(ctg_name, keep, [], None, None, [], [])
    want = keep[:, None] - FLANK + np.arange(N_POS, dtype=np.int64)[None, :]
    wi = np.searchsorted(row_pos, want)
    hit = (wi < row_pos.size) & (row_pos[np.minimum(wi, row_pos.size - 1)] == want)
    windows = np.where(hit, wi, -1).astype(np.int32)
    centre_rows = windows[:, FLANK]
    dev = torch.device('cuda', torch.cuda.current_device())
    tok.stream.win_pos = windows.reshape(-1)
    tensor, depth = encode_pileup(stream_to_device(tok.stream, dev), PIPELINE_LOW_BQ_CUT, dev)
    return EncodedChunk(ctg_name, keep, ref33, tensor, depth, [tok.alt_info_of(int(r)) for r in centre_rows],
                        [cand[int(p)] for p in keep])


def create_tensor(args):
    chunk = encode_chunk(args)
    if tensor_out := (args.tensor_can_fn != "PIPE"):
        fpo = open(args.tensor_can_fn, "wb")
        zp = Popen(shlex.split("{} -c".format(args.zstd)), stdin=PIPE, stdout=fpo, bufsize=8388608)
        out = zp.stdin
    else:
        out = sys.stdout.buffer
    if len(chunk):
        out.write(host.format_tensor_can_rows(chunk.ctg, chunk.pos, chunk.ref33, chunk.tensor.cpu().numpy(), chunk.alt_info,
                                              chunk.variant_type))
    if tensor_out:
        zp.stdin.close()
        zp.wait()
        fpo.close()
    print("[INFO] {} {} Tensors generated: {}".format(args.ctg_name, get_chunk_id(args.candidates_bed_regions), len(chunk)))


def build_parser():
    from .predict import str2bool
    p = ArgumentParser(description="Generate tumor pileup tensors for calling (B200 engine)")
    p.add_argument('--platform', type=str, default='ont')
    p.add_argument('--tumor_bam_fn', type=str, default=None)
    p.add_argument('--ref_fn', type=str, default=None)
    p.add_argument('--tensor_can_fn', type=str, default="PIPE")
    p.add_argument('--vcf_fn', type=str, default=None)
    p.add_argument('--snv_min_af', type=float, default=0.05)
    p.add_argument('--ctg_name', type=str, default=None)
    p.add_argument('--ctg_start', type=int, default=None)
    p.add_argument('--ctg_end', type=int, default=None)
    p.add_argument('--bed_fn', type=str, default=None)
    p.add_argument('--samtools', type=str, default="samtools")
    p.add_argument('--min_coverage', type=float, default=4)
    p.add_argument('--min_mq', type=int, default=20)
    p.add_argument('--min_bq', type=int, default=None)
    p.add_argument('--max_depth', type=int, default=None)
    p.add_argument('--extend_bed', nargs='?', action="store", type=str, default=None)
    p.add_argument('--alt_fn', type=str, default=None)
    p.add_argument('--zstd', type=str, default=ZSTD, help=SUPPRESS)
    p.add_argument('--indel_min_af', type=float, default=1.0, help=SUPPRESS)
    p.add_argument('--chunk_num', type=int, default=None, help=SUPPRESS)
    p.add_argument('--chunk_id', type=int, default=None, help=SUPPRESS)
    p.add_argument('--candidates_bed_regions', type=str, default=None, help=SUPPRESS)
    p.add_argument('--phase_tumor', type=str2bool, default=0, help=SUPPRESS)
    p.add_argument('--flanking', type=int, default=None, help=SUPPRESS)
    p.add_argument('--max_indel_length', type=int, default=None, help=SUPPRESS)
    p.add_argument('--truth_vcf_fn', type=str, default=None, help=SUPPRESS)
    return p


def main(argv=None):
    create_tensor(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
