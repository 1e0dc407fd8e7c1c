"""GPU parity of the candidate scan (STEP 1) through the C ABI: ``cto_index_rows``,
``cto_scan_candidates`` and ``cto_scan_candidates_host`` against ``oracle/candidates_oracle.py`` (itself pinned to the
reference by tests/test_oracle_golden.py and tests/test_oracle_vs_reference.py) and, for the sub-command, against the
files the unmodified reference wrote (tests/golden/candidates).  Integer / byte work: every comparison is exact."""

import ctypes as C
import json
import os

import numpy as np
import pytest
import torch

from clairs_to_b200 import _lib, synth
from clairs_to_b200 import extract_candidates_calling as ecc
from oracle import candidates_oracle as co

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "fake_samtools.py")
DEFAULT = dict(min_coverage=4.0, snv_min_af=0.05, indel_min_af=0.05, alternative_base_num=3, select_indel_candidates=True)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def device_text(text: bytes, misalign=0):
    buf = torch.zeros(misalign + len(text) + 32, dtype=torch.uint8, device="cuda")
    buf[misalign:misalign + len(text)] = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda() if text else buf[:0]
    return buf, buf[misalign:]


def index_rows(view, n_bytes, cap):
    row_off = torch.full((cap + 1,), -1, dtype=torch.int64, device="cuda")
    n = C.c_int64()
    _lib.check(_lib.lib().cto_index_rows(_ptr(view), n_bytes, _ptr(row_off), cap, C.byref(n), None), "cto_index_rows")
    return row_off, n.value


def scan_device(text: bytes, reference: str, reference_start, misalign=0, **kw):
    keep, view = device_text(text, misalign)
    cap = text.count(b"\n") + 1
    row_off, n = index_rows(view, len(text), cap)
    ref = torch.frombuffer(bytearray(reference.encode()), dtype=torch.uint8).cuda()
    pos = torch.empty(n, dtype=torch.int32, device="cuda")
    depth = torch.empty(n, dtype=torch.int32, device="cuda")
    flags = torch.empty(n, dtype=torch.uint8, device="cuda")
    over = C.c_int32()
    alt = kw["alternative_base_num"]
    _lib.check(_lib.lib().cto_scan_candidates(_ptr(view), len(text), _ptr(row_off), n, _ptr(ref), reference_start, len(reference),
                                              float(kw["min_coverage"]), float(kw["snv_min_af"]), float(kw["indel_min_af"]),
                                              -1 if alt is None else alt, int(kw["select_indel_candidates"]), _ptr(pos), _ptr(depth),
                                              _ptr(flags), C.byref(over), None), "cto_scan_candidates")
    torch.cuda.synchronize()
    return pos.cpu().numpy(), depth.cpu().numpy(), flags.cpu().numpy(), over.value


def expected(rows, reference, reference_start, **kw):
    """(pos, depth, flags) per row from the oracle (rows whose reference base is not ACGT: flags 0, depth 0)."""
    pos, depth, flags = [], [], []
    for row in rows:
        cols = row.strip().split("\t")
        p = int(cols[1])
        rb = reference[p - reference_start].upper()
        pos.append(p)
        if rb not in "ACGT":
            depth.append(0)
            flags.append(0)
            continue
        d, pass_af, snv, indel = co.site_decision(cols[4], rb, **kw)
        depth.append(d)
        flags.append(1 | (2 if pass_af else 0) | (4 if snv else 0) | (8 if indel else 0))
    return np.array(pos, np.int32), np.array(depth, np.int32), np.array(flags, np.uint8)


@pytest.mark.parametrize("text", [b"", b"a\n", b"a", b"\n\n\n", b"abc\ndef", b"x" * 8191 + b"\n" + b"y" * 40000 + b"\nz",
                                  b"\n".join(b"r%d" % i * (i % 7) for i in range(5000)) + b"\n"])
@pytest.mark.parametrize("misalign", [0, 3])
def test_row_index_matches_numpy(text, misalign):
    keep, view = device_text(text, misalign)
    cap = text.count(b"\n") + 1
    row_off, n = index_rows(view, len(text), cap)
    arr = np.frombuffer(text, np.uint8)
    want = [0] + [int(i) + 1 for i in np.flatnonzero(arr == 10)]
    if len(text) and text[-1:] != b"\n":
        want.append(len(text))
    if not text:
        want = [0]
    assert n == len(want) - 1
    assert row_off[:n + 1].cpu().tolist() == want


@pytest.mark.parametrize("kw", [
    DEFAULT,
    dict(DEFAULT, select_indel_candidates=False),
    dict(min_coverage=0.0, snv_min_af=0.0, indel_min_af=1.0, alternative_base_num=1, select_indel_candidates=True),
    dict(min_coverage=10.0, snv_min_af=0.08, indel_min_af=0.1, alternative_base_num=None, select_indel_candidates=True),
    dict(min_coverage=4.5, snv_min_af=0.2, indel_min_af=0.02, alternative_base_num=5, select_indel_candidates=True),
])
@pytest.mark.parametrize("misalign", [0, 5])
def test_scan_matches_oracle_fuzz(kw, misalign):
    rows, reference = synth.scan_rows_text(2500, 5, first_pos=301, depth_mean=35, weird=0.02)
    text = "".join(rows).encode()
    if misalign:
        text = text[:-1]                                       # and no trailing newline
    pos, depth, flags, _ = scan_device(text, reference, 301, misalign=misalign, **kw)
    wp, wd, wf = expected(rows, reference, 301, **kw)
    assert np.array_equal(pos, wp)
    assert np.array_equal(flags, wf), np.flatnonzero(flags != wf)[:10]
    assert np.array_equal(depth, wd)
    if kw["alternative_base_num"] is None:                     # EC:131-137: "is not None and count >= ..." never holds
        assert not (wf & 14).any()
    else:
        assert (wf & 4).sum() > 30 and (not kw["select_indel_candidates"] or (wf & 8).sum() > 30)


def test_scan_deep_rows_and_allele_table_overflow():
    """Rows far beyond the shared-memory stage (read in place) and rows with more distinct indel alleles than the
    first-pass table (24) -> second launch; same answers as the oracle."""
    rng = np.random.default_rng(3)
    rows, ref = [], []
    for r in range(300):
        rb = "ACGT"[r % 4]
        ref.append(rb)
        parts = []
        n_alleles = (1, 20, 24, 25, 40, 300)[r % 6]
        depth = 40 if r % 50 else 6000
        for k in range(depth):
            a = int(rng.integers(0, n_alleles))
            if rng.random() < 0.6:
                seq = ''.join("ACGT"[(a >> (2 * j)) & 3] for j in range(5))
                parts.append("%s+%d%s" % (rb, len(seq), seq) if k & 1 else "%s+%d%s" % (rb.lower(), len(seq), seq.lower()))
            elif rng.random() < 0.5:
                parts.append("%s-%d%s" % (rb, a + 1, "N" * (a + 1)))
            else:
                parts.append("ACGT"[int(rng.integers(0, 4))])
        rows.append("chr1\t%d\tN\t%d\t%s\t%s\n" % (1 + r, depth, ''.join(parts), "I" * depth))
    reference = ''.join(ref)
    for kw in (DEFAULT, dict(DEFAULT, indel_min_af=0.3, alternative_base_num=8)):
        pos, depth, flags, n_over = scan_device("".join(rows).encode(), reference, 1, **kw)
        wp, wd, wf = expected(rows, reference, 1, **kw)
        assert n_over >= 40
        assert np.array_equal(flags, wf) and np.array_equal(depth, wd) and np.array_equal(pos, wp)


def test_scan_error_flags():
    rows = ["chr1\t5\tN\t3\tAAA\tIII\n", "chr1\t900\tN\t3\tAAA\tIII\n", "chr1\t6\tN\n", "chr1\t7\tN\t1\tC\tI\n"]
    pos, depth, flags, _ = scan_device("".join(rows).encode(), "ACGTACGTAC", 1, **DEFAULT)
    assert flags.tolist() == [1, ecc.F_BAD_REF, ecc.F_MALFORMED, 1] and pos.tolist() == [5, 900, 6, 7]
    with pytest.raises(SystemExit):
        ecc.candidate_positions(pos, flags)


def test_scan_host_pipeline_equals_device_call():
    """cto_scan_candidates_host on > 2 pieces of 32 MB (copy of piece k + 1 under the scan of piece k) against the
    single-call device path on one tile of the same text; pageable and pinned host memory."""
    rows, reference = synth.scan_rows_text(1500, 11, first_pos=1, depth_mean=45, weird=0.01)
    tile = "".join(rows).encode()
    reps = (80 << 20) // len(tile) + 1
    text = tile * reps
    p1, d1, f1, _ = scan_device(tile, reference, 1, **DEFAULT)
    for pinned in (False, True):
        src = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory() if pinned else text
        pos, depth, flags = ecc.scan_mpileup(src, reference, 1, **DEFAULT)
        assert len(pos) == reps * len(rows)
        assert np.array_equal(pos, np.tile(p1, reps)) and np.array_equal(depth, np.tile(d1, reps)) and np.array_equal(flags, np.tile(f1, reps))
    pos, depth, flags = ecc.scan_mpileup(b"", reference, 1, **DEFAULT)
    assert len(pos) == 0


@pytest.mark.parametrize("name", ["snv_indel", "snv_only", "bed"])
def test_extract_candidates_cli_files_identical(golden_dir, tmp_path, capsys, name):
    """The sub-command against the unmodified reference's output files and [INFO] line."""
    work = os.path.join(golden_dir, "candidates")
    argv = json.load(open(os.path.join(work, name, "args.json")))
    argv = [os.path.join(work, a) if a.endswith(".bed") else a for a in argv]
    folder = str(tmp_path / name)
    os.makedirs(folder)
    ecc.main(["--tumor_bam_fn", os.path.join(work, "tumor.bam"), "--ref_fn", os.path.join(work, "ref.fa"), "--samtools", SHIM,
              "--ctg_name", "chr20", "--platform", "ont_r10_dorado_sup_5khz", "--min_coverage", "4", "--min_bq", "20", "--output_depth", "True",
              "--genotyping_mode_vcf_fn", "None", "--hybrid_mode_vcf_fn", "None", "--candidates_folder", folder] + argv)
    assert capsys.readouterr().out == open(os.path.join(work, name, "stdout.txt")).read()
    want_dir = os.path.join(work, name)
    names = sorted(f for f in os.listdir(want_dir) if f not in ("args.json", "stdout.txt", "bed"))
    assert sorted(f for f in os.listdir(folder) if f != "bed") == names
    for fn in names:
        got = open(os.path.join(folder, fn)).read().replace(folder + "/", "<candidates_folder>/")
        assert got == open(os.path.join(want_dir, fn)).read(), fn
    assert os.listdir(os.path.join(folder, "bed")) == os.listdir(os.path.join(want_dir, "bed"))
    for fn in os.listdir(os.path.join(want_dir, "bed")):
        assert open(os.path.join(folder, "bed", fn)).read() == open(os.path.join(want_dir, "bed", fn)).read()


def test_rejected_modes(tmp_path):
    with pytest.raises(SystemExit):
        ecc.main(["--ctg_name", "chr20", "--ref_fn", "x.fa", "--tumor_bam_fn", "x.bam", "--truth_vcf_fn", "t.vcf", "--candidates_folder",
                  str(tmp_path)])
