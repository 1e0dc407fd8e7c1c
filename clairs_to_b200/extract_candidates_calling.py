"""Drop-in for the reference sub-command ``extract_candidates_calling`` (src/extract_candidates_calling.py, cited as EC),
B200 path: STEP 1 of the pipeline, SURVEY.md section 8 row f3.

Same flags (EC:505-612), same external tools (``samtools faidx`` / ``samtools mpileup`` with the identical command line,
EC:299-317), same output files, byte for byte: ``<candidates_folder>/bed/<ctg>_<chunk>.bed`` (EC:382-389),
``<ctg>.<chunk>_<idx>_<n>_snv`` / ``_indel`` region files of at most 10 000 candidates (EC:450-488) and the
``SNV_CANDIDATES_FILE_`` / ``INDEL_CANDIDATES_FILE_`` lists, and the same ``[INFO]`` line on stdout.  What changes: the
Python loop over EVERY mpileup row of the chunk (EC:335-377 calling ``decode_pileup_bases`` EC:55-169) is one CUDA scan
of the whole mpileup text (``cto_scan_candidates_host``); the host only keeps the positions the scan flags.

On this path: the mode ``run_clairs_to`` uses (run_clairs_to:1196-1221).  ``--truth_vcf_fn``, ``--hybrid_mode_vcf_fn``,
``--genotyping_mode_vcf_fn``, ``--alt_fn`` and ``--store_tumor_infos`` (training / debugging / genotyping modes) are
rejected with a message.
"""

from __future__ import annotations

import ctypes as C
import gzip
import os
import shlex
import sys
from argparse import SUPPRESS, ArgumentParser
from subprocess import PIPE, Popen

import numpy as np

from . import _lib

FLANK = 16                        # shared/param.py:59
EXPAND_REFERENCE = 1000           # shared/param.py:81
SAMTOOLS_FILTER_FLAG = 2316       # shared/param.py:27
SPLIT_BED_SIZE = 10000            # shared/param.py:21
SNV_MIN_AF = 0.05                 # shared/param.py:22
MIN_COVERAGE = 4                  # shared/param.py:20
MIN_MQ = 20                       # shared/param.py:17
ALTERNATIVE_BASE_NUM = 3          # shared/param.py:29

F_VALID, F_PASS_AF, F_SNV, F_INDEL, F_MALFORMED, F_BAD_REF, F_OVERFLOW = 1, 2, 4, 8, 32, 64, 128


def str2bool(v):
    """shared/utils.py:121-131."""
    if v is None or isinstance(v, bool):
        return v
    if v.lower() in ('yes', 'ture', 'true', 't', 'y', '1'):
        return True
    if v.lower() in ('no', 'flase', 'false', 'f', 'n', '0'):
        return False
    import argparse
    raise argparse.ArgumentTypeError('Boolean value expected.')


def str_none(v):
    """shared/utils.py:112-118."""
    if v is None:
        return None
    if v.upper() == "NONE":
        return None
    if isinstance(v, str):
        return v


def file_path_from(file_name, suffix="", exit_on_not_found=False, sep="", allow_none=False):
    """shared/utils.py:57-73 (the absolute path of an existing file, optionally with a suffix)."""
    if allow_none and file_name is None:
        return None
    if os.path.isfile(file_name + suffix):
        return os.path.abspath(file_name + suffix)
    if len(sep) == 1:
        candidate = sep.join(file_name.split(sep)[:-1]) + suffix
        if os.path.isfile(candidate):
            return os.path.abspath(candidate)
    if exit_on_not_found:
        sys.exit("[ERROR] file %s not found" % (file_name + suffix))
    return None


def _read_maybe_gzip(path):
    """What ``gzip -fdc`` gives (shared/interval_tree.py:44): the decompressed text, or the file itself if it is plain."""
    with open(path, 'rb') as f:
        raw = f.read()
    return gzip.decompress(raw) if raw[:2] == b'\x1f\x8b' else raw


def bed_intervals(bed_file_path, contig_name=None):
    """shared/interval_tree.py:19-75 for one contig: half-open intervals (start == end widened by one) plus the
    (bed_start, bed_end) range that ``return_bed_region=True`` reports."""
    starts, ends = [], []
    bed_start, bed_end = float('inf'), 0
    if bed_file_path is None or bed_file_path == "":
        return np.empty(0, np.int64), np.empty(0, np.int64), None, None
    for row_id, row in enumerate(_read_maybe_gzip(bed_file_path).decode().splitlines()):
        if not row or row[0] == '#':
            continue
        columns = row.strip().split()
        if contig_name is not None and columns[0] != contig_name:
            continue
        ctg_start, ctg_end = int(columns[1]), int(columns[2])
        if ctg_end < ctg_start or ctg_start < 0 or ctg_end < 0:
            sys.exit("[ERROR] Invalid bed input in {}-th row {} {} {}".format(row_id + 1, columns[0], ctg_start, ctg_end))
        bed_start, bed_end = min(ctg_start, bed_start), max(ctg_end, bed_end)
        if ctg_start == ctg_end:
            ctg_end += 1
        starts.append(ctg_start)
        ends.append(ctg_end)
    return np.asarray(starts, np.int64), np.asarray(ends, np.int64), bed_start, bed_end


def positions_in_intervals(pos, starts, ends):
    """``is_region_in(tree, ctg, pos - 1, pos)`` (shared/interval_tree.py:78-88) for an array of 1-based positions: some
    interval [s, e) with s <= pos - 1 < e.  Intervals are merged first, then one binary search per position."""
    pos = np.asarray(pos, np.int64)
    if starts.size == 0:
        return np.zeros(pos.shape, bool)
    order = np.argsort(starts, kind='stable')
    s, e = starts[order], np.maximum.accumulate(ends[order])
    k = np.searchsorted(s, pos - 1, side='right') - 1           # last interval starting at or before pos - 1
    return (k >= 0) & (e[np.maximum(k, 0)] > pos - 1)


def reference_sequence_from(samtools, fasta, region):
    """shared/utils.py:148-174."""
    proc = Popen(shlex.split("%s faidx %s %s" % (samtools, fasta, region)), stdout=PIPE, bufsize=8388608, universal_newlines=True)
    lines = [row.rstrip() for row in proc.stdout]
    proc.stdout.close()
    proc.wait()
    if proc.returncode != 0:
        return None
    return "".join(lines[1:]).upper()


def scan_mpileup(text, reference, reference_start, min_coverage, snv_min_af, indel_min_af, alternative_base_num,
                 select_indel_candidates, device=None):
    """The CUDA scan of a chunk's mpileup text (bytes / a pinned uint8 tensor): (pos int32, depth int32, flags uint8) per
    row.  No CPU fallback."""
    import torch
    if not torch.cuda.is_available():
        sys.exit("[ERROR] no CUDA device: the B200 candidate scan has no CPU fallback")
    lib = _lib.lib()
    if device is not None:
        torch.cuda.set_device(device)
    if isinstance(text, torch.Tensor):
        n_bytes, text_ptr = text.numel(), C.c_void_p(text.data_ptr())
        cap = int((text == 10).sum()) + 1
    else:
        n_bytes = len(text)
        cap = text.count(b'\n') + 1
        text_ptr = C.cast(C.c_char_p(text), C.c_void_p)
    ref = reference.encode() if isinstance(reference, str) else reference
    pos = np.empty(cap, np.int32)
    depth = np.empty(cap, np.int32)
    flags = np.empty(cap, np.uint8)
    n_rows, n_over = C.c_int64(), C.c_int64()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.cto_scan_candidates_host(text_ptr, n_bytes, C.cast(C.c_char_p(ref), C.c_void_p), int(reference_start), len(ref),
                                            float(min_coverage), float(snv_min_af), float(indel_min_af),
                                            -1 if alternative_base_num is None else int(alternative_base_num),
                                            int(bool(select_indel_candidates)), cap, C.c_void_p(pos.ctypes.data),
                                            C.c_void_p(depth.ctypes.data), C.c_void_p(flags.ctypes.data), C.byref(n_rows),
                                            C.byref(n_over), stream), "cto_scan_candidates_host")
    n = n_rows.value
    return pos[:n], depth[:n], flags[:n]


def candidate_positions(pos, flags):
    """The three position sets of EC:355-377 as sorted arrays (a position appearing in several rows counts once)."""
    bad = flags & (F_MALFORMED | F_BAD_REF | F_OVERFLOW)
    if bad.any():
        r = int(np.flatnonzero(bad)[0])
        what = "malformed row" if flags[r] & F_MALFORMED else \
            "position outside the loaded reference" if flags[r] & F_BAD_REF else "more than 8192 distinct indel alleles"
        sys.exit("[ERROR] mpileup row %d (position %d): %s" % (r + 1, int(pos[r]), what))
    every = np.unique(pos[(flags & F_PASS_AF) != 0]).astype(np.int64)
    snv = np.unique(pos[(flags & F_SNV) != 0]).astype(np.int64)
    indel = np.unique(pos[(flags & F_INDEL) != 0]).astype(np.int64)
    return every, snv, indel


def write_region_files(candidates_folder, ctg_name, chunk_id, positions, suffix, list_prefix):
    """EC:450-467 / 469-488: region files of at most ``split_bed_size`` candidates + the list of their paths."""
    n = len(positions)
    region_num = n // SPLIT_BED_SIZE + 1 if n % SPLIT_BED_SIZE else n // SPLIT_BED_SIZE
    paths = []
    for idx in range(region_num):
        part = positions[idx * SPLIT_BED_SIZE: (idx + 1) * SPLIT_BED_SIZE]
        output_path = os.path.join(candidates_folder, '{}.{}_{}_{}_{}'.format(ctg_name, chunk_id, idx, region_num, suffix))
        paths.append(output_path)
        with open(output_path, 'w') as f:
            f.write('\n'.join('\t'.join([ctg_name, str(max(int(x) - FLANK - 1, 1)), str(int(x) + FLANK + 1)]) for x in part) + '\n')
    with open(os.path.join(candidates_folder, '{}_{}_{}'.format(list_prefix, ctg_name, chunk_id)), 'w') as f:
        f.write('\n'.join(paths) + '\n')


def extract_pair_candidates(args):
    """EC:172-503."""
    for flag in ('truth_vcf_fn', 'hybrid_mode_vcf_fn', 'genotyping_mode_vcf_fn', 'alt_fn'):
        if getattr(args, flag, None):
            sys.exit("[ERROR] --%s is not on the B200 path of extract_candidates_calling" % flag)
    if args.store_tumor_infos:
        sys.exit("[ERROR] --store_tumor_infos is not on the B200 path of extract_candidates_calling")
    ctg_start, ctg_end = args.ctg_start, args.ctg_end
    ctg_name = args.ctg_name
    chunk_id = args.chunk_id - 1 if args.chunk_id else None     # EC:181
    chunk_num = args.chunk_num
    candidates_folder = args.candidates_folder
    confident_bed_fn = file_path_from(args.bed_fn, allow_none=True, exit_on_not_found=False)
    default_indel_bed_fn = file_path_from(args.call_indels_only_in_these_regions, allow_none=True, exit_on_not_found=False)
    flanking = FLANK if args.flanking is None else args.flanking
    no_of_positions = 2 * flanking + 1
    select_indel = bool(args.select_indel_candidates)
    fai_fn = file_path_from(args.ref_fn, suffix=".fai", exit_on_not_found=True, sep='.')

    if chunk_id is not None:                                    # EC:236-262
        if confident_bed_fn is not None:
            _, _, bed_start, bed_end = bed_intervals(confident_bed_fn, ctg_name)
            span = bed_end - bed_start
            chunk_size = span // chunk_num + 1 if span % chunk_num else span // chunk_num
            ctg_start = bed_start + 1 + chunk_size * chunk_id
            ctg_end = ctg_start + chunk_size
        else:
            contig_length = 0
            with open(fai_fn, 'r') as fai_fp:
                for row in fai_fp:
                    columns = row.strip().split("\t")
                    if columns[0] != ctg_name:
                        continue
                    contig_length = int(columns[1])
            chunk_size = contig_length // chunk_num + 1 if contig_length % chunk_num else contig_length // chunk_num
            ctg_start = chunk_size * chunk_id
            ctg_end = ctg_start + chunk_size

    reads_region = None
    if ctg_name is not None and ctg_start is not None and ctg_end is not None:     # EC:271-283
        extend_start = max(ctg_start - no_of_positions, 1)
        extend_end = ctg_end + no_of_positions
        reads_region = "{}:{}-{}".format(ctg_name, extend_start, extend_end)
        reference_start = max(ctg_start - EXPAND_REFERENCE, 1)
        ref_region = "{}:{}-{}".format(ctg_name, reference_start, ctg_end + EXPAND_REFERENCE)
    elif ctg_name is not None:
        reads_region = ref_region = ctg_name
        reference_start = 1
    else:
        sys.exit("[ERROR] --ctg_name is required")
    reference_sequence = reference_sequence_from(args.samtools, args.ref_fn, ref_region)
    if reference_sequence is None or len(reference_sequence) == 0:
        sys.exit("[ERROR] Failed to load reference sequence from file ({}).".format(args.ref_fn))

    # the exact command of EC:299-314
    samtools_command = args.samtools + " mpileup --reverse-del" + ' ' + ' -r {}'.format(reads_region) + \
        ' --min-MQ {}'.format(args.min_mq) + ' --min-BQ {}'.format(args.min_bq) + \
        (' -l {}'.format(confident_bed_fn) if confident_bed_fn is not None else "") + \
        ' --excl-flags {} '.format(SAMTOOLS_FILTER_FLAG) + \
        (' --max-depth {} '.format(args.max_depth) if args.max_depth is not None else " ")
    stdin = None if args.tumor_bam_fn != "PIPE" else sys.stdin
    bam = args.tumor_bam_fn if args.tumor_bam_fn != "PIPE" else "-"
    mp = Popen(shlex.split(samtools_command + ' ' + bam), stdin=stdin, stdout=PIPE, stderr=PIPE, bufsize=8388608)
    text, _ = mp.communicate()

    pos, _, flags = scan_mpileup(text, reference_sequence, reference_start, args.min_coverage, args.snv_min_af, args.indel_min_af,
                                 args.alternative_base_num, select_indel)
    every, snv_list, indel_list = candidate_positions(pos, flags)

    os.makedirs(os.path.join(candidates_folder, 'bed'), exist_ok=True)            # EC:382-389
    with open(os.path.join(candidates_folder, "bed", '{}_{}.bed'.format(ctg_name, chunk_id)), 'w') as output_bed:
        output_bed.write(''.join('\t'.join([ctg_name, str(int(p) - 1), str(int(p))]) + '\n' for p in every))

    if select_indel and args.bed_fn_source is None and len(indel_list):            # EC:395-404
        starts, ends, _, _ = bed_intervals(default_indel_bed_fn, ctg_name)
        if starts.size:                                                             # an empty tree passes everything
            indel_list = indel_list[positions_in_intervals(indel_list, starts, ends)]

    if select_indel:                                                                # EC:437-448
        print("[INFO] {} chunk {}/{}: Total SNV candidates found: {}, total Indel candidates found: {}".format(
            ctg_name, chunk_id, chunk_num, len(snv_list), len(indel_list)))
    else:
        print("[INFO] {} chunk {}/{}: Total SNV candidates found: {}".format(ctg_name, chunk_id, chunk_num, len(snv_list)))
    if candidates_folder is not None and len(snv_list):
        write_region_files(candidates_folder, ctg_name, chunk_id, snv_list, 'snv', 'SNV_CANDIDATES_FILE')
    if select_indel and candidates_folder is not None and len(indel_list):
        write_region_files(candidates_folder, ctg_name, chunk_id, indel_list, 'indel', 'INDEL_CANDIDATES_FILE')
    return every, snv_list, indel_list


def build_parser():
    p = ArgumentParser(description="Generate tumor variant candidates for tensor creation in calling (B200 engine)")
    p.add_argument('--platform', type=str, default='ont')
    p.add_argument('--candidates_folder', type=str, default=None)
    p.add_argument('--tumor_bam_fn', type=str, default=None)
    p.add_argument('--ref_fn', type=str, default=None)
    p.add_argument('--snv_min_af', type=float, default=SNV_MIN_AF)
    p.add_argument('--ctg_name', type=str, default=None)
    p.add_argument('--ctg_start', type=int, default=None)
    p.add_argument('--ctg_end', type=int, default=None)
    p.add_argument('--bed_fn_source', type=str_none, default=None)
    p.add_argument('--bed_fn', type=str, default=None)
    p.add_argument('--call_indels_only_in_these_regions', type=str, default=None)
    p.add_argument('--samtools', type=str, default="samtools")
    p.add_argument('--min_coverage', type=float, default=MIN_COVERAGE)
    p.add_argument('--min_mq', type=int, default=MIN_MQ)
    p.add_argument('--min_bq', type=int, default=None)
    p.add_argument('--max_depth', type=int, default=None)
    p.add_argument('--alternative_base_num', type=int, default=ALTERNATIVE_BASE_NUM)
    p.add_argument('--select_indel_candidates', type=str2bool, default=0)
    p.add_argument('--hybrid_mode_vcf_fn', type=str_none, default=None)
    p.add_argument('--genotyping_mode_vcf_fn', type=str_none, default=None)
    p.add_argument('--output_depth', type=str2bool, default=False)
    p.add_argument('--output_alt_info', type=str2bool, default=False)
    p.add_argument('--min_truth_snv_af', type=float, default=None, help=SUPPRESS)
    p.add_argument('--store_tumor_infos', type=str2bool, default=False, help=SUPPRESS)
    p.add_argument('--alt_fn', type=str, default=None, help=SUPPRESS)
    p.add_argument('--indel_min_af', type=float, default=1.0, help=SUPPRESS)
    p.add_argument('--min_truth_indel_af', type=float, default=None, help=SUPPRESS)
    p.add_argument('--truth_vcf_fn', type=str, default=None, help=SUPPRESS)
    p.add_argument('--chunk_num', type=int, default=None, help=SUPPRESS)
    p.add_argument('--chunk_id', type=int, default=None, help=SUPPRESS)
    p.add_argument('--flanking', type=int, default=None, help=SUPPRESS)
    return p


def main(argv=None):
    extract_pair_candidates(build_parser().parse_args(argv))


if __name__ == "__main__":
    main()
