#!/usr/bin/env python
"""bench.py -- candidate sites/sec through the per-candidate hot path (pileup encoder + AFF + NEG + posterior
[+ QUAL/FILTER]) on N B200s, next to the CPU port of the reference timed on the same box.

    python bench.py --gpus N --steps K --warmup W [--config C]        (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W [--config C]

--config selects the BASELINE.json workload (index into its `configs` list; default 1, the configuration the metric
is quoted on):
    1  100 000 ONT-shape candidate sites PER GPU, SNV AFF+NEG models                       (weak scaling)
    2  ONT 50x COLO829-shape: 2 000 000 SNV + 200 000 indel candidates in total, sharded    (strong scaling)
    3  Illumina ilmn_ssrs: 1 000 000 candidates, ONE pileup stream feeds both networks      (strong scaling)
    4  PacBio Revio hifi_revio: 1 000 000 SNV + 100 000 indel candidates, posterior + QUAL -> FILTER on device (strong)
Synthetic sites (clairs_to_b200/synth.py), seeded random weights, synthetic likelihood tables.  One step = one pass
of the whole hot path over the batch.  Prints ONE JSON line (rank 0).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidate sites/sec (pileup+AFF+NEG)"
UNIT = "candidate sites/s"

# parts: (n_heads, candidates); "total" = sharded over the ranks (strong scaling), otherwise per GPU (weak scaling)
CONFIGS = {
    1: dict(workload="BASELINE configs[1]: synthetic 100k ONT-shape candidate sites, pileup SNV model, AFF+NEG (+posterior)",
            platform="ont", literal="ont_r10_dorado_sup_5khz", parts=[(4, 100000)], total=False, single_stream=False, qual=None),
    2: dict(workload="BASELINE configs[2]: ONT 50x COLO829-shape synthetic, 2M SNV + 200k indel candidates, SNV then indel model pair",
            platform="ont", literal="ont_r10_dorado_sup_5khz", parts=[(4, 2000000), (6, 200000)], total=True, single_stream=False, qual=None),
    3: dict(workload="BASELINE configs[3]: Illumina ilmn_ssrs, short-read pileup shape, 1M candidates, one stream feeds both networks",
            platform="ilmn", literal="ilmn_ssrs", parts=[(4, 1000000)], total=True, single_stream=True, qual=None),
    4: dict(workload="BASELINE configs[4]: PacBio Revio hifi_revio 50x synthetic, 1M SNV + 100k indel candidates, "
                     "AFF+NEG+posterior+QUAL->FILTER (shared/param.py:35-40 thresholds) on device",
            platform="hifi", literal="hifi_revio", parts=[(4, 1000000), (6, 100000)], total=True, single_stream=False,
            qual=(8.0, 8.0, 12.0)),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        """Samples that arrived inside [t0, t1] (the timed region); the sampler is started before the warm-up because
        nvidia-smi needs a few hundred ms to deliver its first line, so if the region was shorter than one sampling
        period the samples taken under the warm-up load are used instead (and `window` says so)."""
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        every = [(t, r) for t, r in self.rows if len(r) >= 9]
        rows = [r for t, r in every if t0 is None or (t0 <= t <= t1 + 0.1)]
        window = "timed region"
        if not rows:
            rows, window = [r for _, r in every], "warm-up + timed region"
        if not rows:
            return None
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=float(rows[0][2]), reasons=sorted(reasons), samples=len(rows),
                    power_w_max=max(float(r[3]) for r in rows), window=window)


TRAFFIC_KERNEL = {"neg_gru2_recurrent": "gru4_kernel<192", "neg_gru1_recurrent": "gru1_fused_kernel",
                  "encoder": "encode_pileup_kernel", "aff_stage1_fused": "aff_stage1_kernel", "aff_layers_fused": "aff_layers_kernel<64", "aff_layers_fused_last": "aff_layers_kernel<128",
                  "neg_proj2_gemm": "gemm_pair_kernel"}


def ncu_traffic(family, candidates_per_launch):
    """DRAM bytes per launch of the kernel behind a profile family, from the committed ncu capture (profiles/
    r2_traffic.json, falling back to r1_traffic.json; written by profiles/summarize.py traffic), scaled linearly from
    the capture's launch size to this run's mean launch size (every byte these kernels move is per candidate); None
    when no capture holds that kernel."""
    key = TRAFFIC_KERNEL.get(family)
    if not key:
        return None
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        path = os.path.join(ROOT, "profiles", fn)
        if not os.path.exists(path):
            continue
        t = json.load(open(path))
        for name, k in t["kernels"].items():
            if key in name:
                per = k.get("candidates_per_launch", t.get("candidates_per_launch"))
                return k["dram_bytes"] * float(candidates_per_launch) / float(per)
    return None


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run on the host cores next to that GPU (NVML's ideal CPU affinity), so that the pinned
    staging buffers are first-touched on the GPU's NUMA node and its H2D copies do not cross sockets.  Returns the number
    of cores bound to, or None when NVML is not available (then the OS placement stands)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def synthetic_likelihood(n_heads, seed=1):
    import numpy as np
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(0.05, 0.95, size=(10 * n_heads, 10)),
                           np.sort(rng.uniform(0.02, 0.98, size=(2 * n_heads, 10)), axis=1)])


def cli_leg(n_chunk, n_heads, likelihood, host_threads):
    """Chunk-file level: one synthetic chunk of `n_chunk` ONT-shape candidates as
    the files run_clairs_to hands to the hot path (BAM stand-in = mpileup text behind the tests' samtools shim, reference
    FASTA, candidate regions, checkpoints, likelihood matrix) -> per-chunk VCF, through (a) the three drop-in sub-commands
    exchanging gzip files like the reference does, (b) clairs_to_b200.hot_path: the same in one process and in memory with
    a persistent engine.  Returns candidates/s of one process for both, after checking that the two VCFs are identical."""
    import shutil
    import tempfile
    import numpy as np
    import torch
    from clairs_to_b200 import call_variants as cv, create_tensor_pileup_calling as ct, hot_path, predict as pr, synth, synth_weights
    from clairs_to_b200.engine import Engine
    work = tempfile.mkdtemp(prefix="cto_cli_")
    try:
        (aff, aa), (neg, na) = synth.synth_pair_tiled(n_chunk, 77, 'ont', base=200000)
        ctg, first = "chr20", 1001
        for s_, a_, k in ((aff, aa, 20), (neg, na, 0)):
            open(os.path.join(work, "tumor.bam.minbq%d.mpileup" % k), "wb").write(synth.render_mpileup_text(s_, a_, ctg, first))
        open(os.path.join(work, "tumor.bam"), "w").close()
        seq = "N" * (first - 1) + "".join("ACGT"[c] for c in neg.ref_code) + "N" * 64
        with open(os.path.join(work, "ref.fa"), "w") as f:
            f.write(">%s\n" % ctg)
            for i in range(0, len(seq), 60):
                f.write(seq[i:i + 60] + "\n")
        open(os.path.join(work, "ref.fa.fai"), "w").write("%s\t%d\t%d\t60\t61\n" % (ctg, len(seq), len(ctg) + 2))
        with open(os.path.join(work, "%s.0_0_1_snv" % ctg), "w") as f:
            for i in range(n_chunk):
                x = first + 33 * i + 16
                f.write("%s\t%d\t%d\n" % (ctg, max(x - 17, 1), x + 17))
        np.savetxt(os.path.join(work, "likelihood.txt"), likelihood)
        aff_sd = synth_weights.synth_state_dict(synth_weights.aff_state_dict_shapes(n_heads), 100 + n_heads)
        neg_sd = synth_weights.synth_state_dict(synth_weights.neg_state_dict_shapes(n_heads), 200 + n_heads)
        ck_a, ck_n = os.path.join(work, "aff.pkl"), os.path.join(work, "neg.pkl")
        torch.save({'model_acgt': aff_sd}, ck_a)
        torch.save({'model_nacgt': neg_sd}, ck_n)
        shim = os.path.join(ROOT, "tests", "fake_samtools.py")
        common = ["--tumor_bam_fn", os.path.join(work, "tumor.bam"), "--ref_fn", os.path.join(work, "ref.fa"), "--ctg_name", ctg,
                  "--samtools", shim, "--candidates_bed_regions", os.path.join(work, "%s.0_0_1_snv" % ctg),
                  "--platform", "ont_r10_dorado_sup_5khz"]
        f = {k: os.path.join(work, k) for k in ("aff", "neg", "predict", "files.vcf", "mem.vcf")}
        quiet = open(os.devnull, "w")
        stdout = sys.stdout

        def files_path():
            for name, k in (("aff", 20), ("neg", 0)):
                ct.main(common + ["--min_bq", str(k), "--tensor_can_fn", f[name]])
            pr.main(["--tensor_fn_acgt", f["aff"], "--tensor_fn_nacgt", f["neg"], "--predict_fn", f["predict"], "--chkpnt_fn_acgt", ck_a,
                     "--chkpnt_fn_nacgt", ck_n, "--use_gpu", "True", "--platform", "ont_r10_dorado_sup_5khz", "--ctg_name", ctg,
                     "--pileup", "--disable_indel_calling", "True"])
            cv.main(["--predict_fn", f["predict"], "--call_fn", f["files.vcf"], "--ref_fn", os.path.join(work, "ref.fa"), "--platform",
                     "ont_r10_dorado_sup_5khz", "--likelihood_matrix_data", os.path.join(work, "likelihood.txt"),
                     "--disable_indel_calling", "True"])

        hp_args = hot_path.build_parser().parse_args(common + ["--min_bq", "20", "--chkpnt_fn_acgt", ck_a, "--chkpnt_fn_nacgt", ck_n,
                                                              "--likelihood_matrix_data", os.path.join(work, "likelihood.txt"),
                                                              "--disable_indel_calling", "True", "--call_fn", f["mem.vcf"]])
        eng = Engine.from_checkpoints(ck_a, ck_n, max_batch=10240)
        eng.set_likelihood(os.path.join(work, "likelihood.txt"))
        sys.stdout = quiet
        try:
            files_path()                                    # warm-up (lazy CUDA module load)
            hot_path.run_chunk(hp_args, engine=eng, host_threads=host_threads)
            t0 = time.perf_counter()
            files_path()
            t1 = time.perf_counter()
            hot_path.run_chunk(hp_args, engine=eng, host_threads=host_threads)
            t2 = time.perf_counter()
        finally:
            sys.stdout = stdout
            eng.close()
        same = open(f["files.vcf"]).read() == open(f["mem.vcf"]).read() if os.path.exists(f["files.vcf"]) else not os.path.exists(f["mem.vcf"])
        rows = sum(1 for l in open(f["mem.vcf"]) if not l.startswith("#")) if os.path.exists(f["mem.vcf"]) else 0
        return dict(candidates=n_chunk, vcf_rows=rows, identical_vcf=bool(same),
                    sub_commands=dict(value=n_chunk / (t1 - t0), unit=UNIT, seconds=t1 - t0,
                                      what="create_tensor x2 -> predict -> call_variants through gzip chunk files, one process, "
                                           "engine loaded per predict call like the reference"),
                    in_memory=dict(value=n_chunk / (t2 - t1), unit=UNIT, seconds=t2 - t1,
                                   what="clairs_to_b200.hot_path.run_chunk: same inputs, same VCF, no intermediate files, persistent engine"),
                    note="both include the python samtools shim filtering %d MB of mpileup text (twice) and the host tokenizer"
                         % (sum(os.path.getsize(os.path.join(work, "tumor.bam.minbq%d.mpileup" % k)) for k in (0, 20)) // 1000000))
    finally:
        shutil.rmtree(work, ignore_errors=True)


def hard_filter_leg(peaks, dev, n_sites=200, replicas=8, steps=5):
    """the per-site hard filters (src/haplotype_filtering.py:344-703) on one phased chunk of `n_sites`
    called variants (the reference's chunk size, HF:188).  Device resident: the chunk's integer arrays in HBM, the site list
    repeated `replicas` times in one launch (CUDA events).  End to end: mpileup text in host memory -> cto_hf_parse -> site
    tables -> copies -> kernel -> result lines.  CPU: the oracle port on a bounded sample of the same sites (one core)."""
    import numpy as np
    import torch
    from clairs_to_b200 import hard_filters as hf
    from clairs_to_b200 import synth
    rows, ref, lo, sites = synth.hard_filter_chunk(n_sites, 2025, depth=50, read_len=(400, 4000))
    text = "".join(rows).encode()
    hf.parse_chunk(text, True, ref, lo)
    t0 = time.perf_counter()
    chunk = hf.parse_chunk(text, True, ref, lo)
    t_parse = time.perf_counter() - t0
    many = sites * replicas
    tables, _ = hf._site_tables(chunk, 1, sites, 100)
    window_entries = int(sum(int(chunk.row_off[h]) - int(chunk.row_off[l]) for l, h in zip(tables["row_lo"], tables["row_hi"])))
    for _ in range(2):
        hf.run_sites(chunk, 1, many)
    # device time of the kernel alone: the host mirror prepares the tables, the launch is bracketed by events inside run_sites' stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    import ctypes as C
    from clairs_to_b200 import _lib
    lib = _lib.lib()
    tab, mul = hf.entropy_table()
    tt, _ = hf._site_tables(chunk, 1, many, 100)
    devc = hf.chunk_to_device(chunk, dev)
    sd = {k: hf._to_device(tt[k], dev) for k in hf.SITE_FIELDS}
    ca = hf.ChunkArrays(chunk.n_rows, chunk.n_entries, *[C.c_void_p(devc[k].data_ptr()) for k in hf.CHUNK_FIELDS])
    sa = hf.SiteArrays(len(many), *[C.c_void_p(sd[k].data_ptr()) for k in hf.SITE_FIELDS[:19]], len(tt["g_row"]),
                       *[C.c_void_p(sd[k].data_ptr()) for k in hf.SITE_FIELDS[19:]])
    flags = torch.empty(len(many), dtype=torch.int32, device=dev)
    pval = torch.empty(len(many), dtype=torch.float64, device=dev)
    scratch = torch.empty(1, dtype=torch.int32, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def launch():
        _lib.check(lib.cto_hard_filter_sites(C.byref(ca), C.byref(sa), 1, 100, 0, 3, tab, mul, hf.SEQUENCE_ENTROPY_THRESHOLD,
                                             C.c_void_p(scratch.data_ptr()), C.c_void_p(flags.data_ptr()), C.c_void_p(pval.data_ptr()), None,
                                             stream), "cto_hard_filter_sites")
    launch()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps):
        launch()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    dev_flags = flags.cpu().numpy().view(np.uint32)[:len(sites)]
    # end to end from the text
    t0 = time.perf_counter()
    for _ in range(steps):
        lines = hf.haplotype_filter_chunk("chr20", sites, text, ref, lo)
    t_e2e = (time.perf_counter() - t0) / steps
    # the reference runs its chunk jobs on a thread pool (src/haplotype_filtering.py:1126-1138); so can the drop-in: the host
    # tokenizer releases the GIL, one chunk per thread
    from concurrent.futures import ThreadPoolExecutor
    n_thr = max(1, min(8, len(os.sched_getaffinity(0))))
    with ThreadPoolExecutor(max_workers=n_thr) as ex:
        list(ex.map(lambda _: hf.haplotype_filter_chunk("chr20", sites, text, ref, lo, n_threads=1), range(n_thr)))
        t0 = time.perf_counter()
        res = list(ex.map(lambda _: hf.haplotype_filter_chunk("chr20", sites, text, ref, lo, n_threads=1), range(2 * n_thr)))
        t_pool = time.perf_counter() - t0
    assert all(r == lines for r in res)
    # CPU port on a sample of the sites (parse included, like the reference's chunk mode)
    from oracle import hard_filter_oracle as ho
    sample = sites[:40]
    t0 = time.perf_counter()
    parsed = ho.parse_chunk(rows, True)
    t_cpu_parse = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = [ho.site_line('haplotype', "chr20", p, rb, ab, 100, parsed, ref, lo, het, hom, False, 3, af) for p, rb, ab, af, het, hom in sample]
    t_cpu = time.perf_counter() - t0
    assert lines[:len(sample)] == want, "hard filter lines differ from the oracle on the bench sample"
    assert (dev_flags == hf.run_sites(chunk, 1, sites)[0]).all()
    bytes_per_launch = 14.0 * window_entries * replicas
    gbs = bytes_per_launch / (ms * 1e-3) / 1e9
    return dict(workload="%d called variants in one phased chunk (%d pileup rows, %.1f MB of mpileup text with QNAME + HP, depth 50), "
                         "haplotype-filter mode, flanking 100" % (len(sites), chunk.n_rows, len(text) / 1e6),
                sites_per_s=len(many) / (ms * 1e-3), ms_per_launch=ms, sites_per_launch=len(many),
                roofline=dict(bound="hbm", kernel="hard_filter_kernel", achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                              algorithmic_bytes_per_launch=bytes_per_launch,
                              note="algorithmic bytes = 14 B (read id, token id, info, qualities) per pileup entry of every site's window, "
                                   "read once; the %d replicas of the site list share one chunk, so the reads come from L2" % replicas),
                e2e=dict(sites_per_s=len(sites) / t_e2e, seconds=t_e2e, host_parse_s=t_parse, host_parse_mb_per_s=len(text) / 1e6 / t_parse,
                         api="hard_filters.haplotype_filter_chunk: mpileup text -> cto_hf_parse_mt (up to 8 host threads) -> site tables -> "
                             "cto_hard_filter_sites -> lines",
                         chunks_on_a_thread_pool=dict(sites_per_s=2 * n_thr * len(sites) / t_pool, threads=n_thr, chunks=2 * n_thr,
                                                      note="one chunk per thread (tokenizer on one thread each), like the reference's ThreadPoolExecutor over chunk jobs")),
                cpu_baseline=dict(sites_per_s=len(sample) / (t_cpu + t_cpu_parse * len(sample) / len(sites)), cores=1, kind="port",
                                  sample="%d sites, oracle/hard_filter_oracle.py (parse time pro rata)" % len(sample)),
                failing=int((~(dev_flags & 1).astype(bool)).sum()),
                parity="lines equal the oracle on the CPU sample")


def scan_leg(raw, peaks, dev, steps=5):
    """the STEP-1 candidate scan (src/extract_candidates_calling.py:55-169, 335-377) over whole-chunk
    mpileup text.  Device resident (text in HBM: row index + scan kernels, CUDA events), end to end from pinned host text
    (cto_scan_candidates_host: 32 MB pieces, copy under compute) and the CPU port on a bounded sample of the same rows."""
    import ctypes as C
    import numpy as np
    import torch
    from clairs_to_b200 import _lib, synth
    from clairs_to_b200 import extract_candidates_calling as ecc
    lib = _lib.lib()
    aff, aff_aux, neg, neg_aux = raw
    text = synth.render_mpileup_text(neg, neg_aux)
    reference = ''.join("ACGT"[c] for c in neg.ref_code)
    n_bytes = len(text)
    host = torch.frombuffer(bytearray(text), dtype=torch.uint8).pin_memory()
    d_text = torch.zeros(n_bytes + 32, dtype=torch.uint8, device=dev)
    d_text[:n_bytes].copy_(host)
    cap = text.count(b"\n") + 1
    row_off = torch.empty(cap + 1, dtype=torch.int64, device=dev)
    ref = torch.frombuffer(bytearray(reference.encode()), dtype=torch.uint8).to(dev)
    pos = torch.empty(cap, dtype=torch.int32, device=dev)
    depth = torch.empty(cap, dtype=torch.int32, device=dev)
    flags = torch.empty(cap, dtype=torch.uint8, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    n = C.c_int64()
    over = C.c_int32()
    kw = dict(min_coverage=4.0, snv_min_af=0.05, indel_min_af=0.05, alt=3, select_indel=1)
    p = lambda t: C.c_void_p(t.data_ptr())

    def index():
        _lib.check(lib.cto_index_rows(p(d_text), n_bytes, p(row_off), cap, C.byref(n), stream), "cto_index_rows")

    def scan():
        _lib.check(lib.cto_scan_candidates(p(d_text), n_bytes, p(row_off), n.value, p(ref), 1001, len(reference), kw["min_coverage"],
                                           kw["snv_min_af"], kw["indel_min_af"], kw["alt"], kw["select_indel"], p(pos), p(depth), p(flags),
                                           C.byref(over), stream), "cto_scan_candidates")

    for _ in range(2):
        index()
        scan()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_idx = t_scan = 0.0
    for _ in range(steps):
        ev[0].record()
        index()
        ev[1].record()
        scan()
        ev[2].record()
        torch.cuda.synchronize()
        t_idx += ev[0].elapsed_time(ev[1])
        t_scan += ev[1].elapsed_time(ev[2])
    t_idx, t_scan = t_idx / steps, t_scan / steps
    n_rows = n.value
    dev_flags = flags[:n_rows].cpu().numpy()
    # end to end from pinned host text
    out = ecc.scan_mpileup(host, reference, 1001, 4.0, 0.05, 0.05, 3, True)
    assert np.array_equal(out[2], dev_flags)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hp, hd, hf = (np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.uint8))
    hn, hov = C.c_int64(), C.c_int64()
    refb = reference.encode()

    def host_call():
        _lib.check(lib.cto_scan_candidates_host(C.c_void_p(host.data_ptr()), n_bytes, C.cast(C.c_char_p(refb), C.c_void_p), 1001, len(refb),
                                                4.0, 0.05, 0.05, 3, 1, cap, C.c_void_p(hp.ctypes.data), C.c_void_p(hd.ctypes.data),
                                                C.c_void_p(hf.ctypes.data), C.byref(hn), C.byref(hov), stream), "cto_scan_candidates_host")

    host_call()
    e0.record()
    for _ in range(steps):
        host_call()
    e1.record()
    torch.cuda.synchronize()
    t_host = e0.elapsed_time(e1) / steps
    # CPU port on a bounded sample of the same rows
    from oracle import candidates_oracle as co
    sample_rows = 12000
    cut = 0
    for _ in range(sample_rows):
        cut = text.index(b"\n", cut) + 1
    rows = text[:cut].decode().splitlines(keepends=True)
    t0 = time.perf_counter()
    sites = co.scan_rows(rows, reference, 1001, min_coverage=4.0, snv_min_af=0.05, indel_min_af=0.05, alternative_base_num=3,
                         select_indel_candidates=True)
    t_cpu = time.perf_counter() - t0
    want = np.array([1 | (2 if v[1] else 0) | (4 if v[2] else 0) | (8 if v[3] else 0) for v in sites.values()], np.uint8)
    assert np.array_equal(want, dev_flags[:sample_rows]), "candidate scan differs from the oracle on the bench sample"
    gbs = n_bytes / (t_scan * 1e-3) / 1e9
    return dict(workload="%d mpileup rows (%.0f MB of text, ONT-shape, --min-BQ 0 stream of the same sites)" % (n_rows, n_bytes / 1e6),
                rows_per_s=n_rows / ((t_idx + t_scan) * 1e-3), ms_row_index=t_idx, ms_scan=t_scan,
                roofline=dict(bound="hbm", kernel="scan_kernel", achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"],
                              algorithmic_bytes_per_launch=n_bytes, note="algorithmic bytes = the mpileup text, read once"),
                row_index_gbs=n_bytes / (t_idx * 1e-3) / 1e9,
                e2e=dict(rows_per_s=n_rows / (t_host * 1e-3), gb_per_s=n_bytes / (t_host * 1e-3) / 1e9, h2d_bytes_per_step=n_bytes,
                         d2h_bytes_per_step=9 * n_rows, api="cto_scan_candidates_host (pinned host text)"),
                cpu_baseline=dict(rows_per_s=sample_rows / t_cpu, cores=1, kind="port", sample="%d rows, oracle/candidates_oracle.py" % sample_rows),
                candidates=dict(pass_af=int((dev_flags & 2).astype(bool).sum()), snv=int((dev_flags & 4).astype(bool).sum()),
                                indel=int((dev_flags & 8).astype(bool).sum())),
                parity="flags equal the oracle on the CPU sample; host call equals the device call")


def cpu_mix(cfg):
    total = float(sum(n for _, n in cfg["parts"]))
    return [(h, n / total) for h, n in cfg["parts"]]


def run_reference(args, cfg, rank, world):
    """--impl reference: the CPU port of the reference's path on all host cores (oracle/cpu_baseline.py)."""
    if rank != 0:
        return
    from oracle import cpu_baseline
    cores = len(os.sched_getaffinity(0))
    per_proc = args.cpu_sample
    times, vals, last = [], [], None
    pool, cores = cpu_baseline.make_pool(cores)          # workers import torch once, not once per step
    try:
        for step in range(args.warmup + args.steps):
            r = cpu_baseline.run(per_proc=per_proc, procs=cores, seed=9000 + 17 * step, pool=pool, platform=cfg["platform"],
                                 mix=cpu_mix(cfg), single_stream=cfg["single_stream"])
            if step >= args.warmup:
                times.append(r["seconds"])
                vals.append(r["value"])
            last = r
    finally:
        pool.close()
        pool.join()
    value = sum(vals) / len(vals)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * sum(times) / len(times), higher_is_better=True, scaling="strong" if cfg["total"] else "weak",
                vs_baseline=None, dtype="f32 (torch CPU)", data="synthetic", impl="reference",
                config=dict(workload=cfg["workload"], platform=cfg["literal"], heads=[h for h, _ in cfg["parts"]],
                            weights="seeded random init", candidates_per_step=last["candidates"],
                            sample="each step is a bounded sample of the workload: %d candidates per host process x %d "
                                   "single-thread processes (the reference's own process model), from mpileup text" % (per_proc, cores)),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=last["sample"],
                                  encoder_share=last["encoder_share"]),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="index into BASELINE.json configs")
    ap.add_argument("--candidates", type=int, default=None, help="override: candidate sites of the first part (per GPU for "
                                                                  "config 1, in total otherwise); other parts scale along")
    ap.add_argument("--max-batch", type=int, default=37888, help="engine chunk (candidates per network pass)")
    ap.add_argument("--cpu-sample", type=int, default=300, help="CPU baseline: candidates per host process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-text", action="store_true", help="skip the mpileup-text end-to-end leg (config 1)")
    ap.add_argument("--no-cli", action="store_true", help="skip the chunk-file level leg (config 1, one GPU)")
    ap.add_argument("--no-scan", action="store_true", help="skip the STEP-1 candidate-scan leg (config 1, one GPU)")
    ap.add_argument("--no-filters", action="store_true", help="skip the per-site hard-filter leg (config 1, one GPU)")
    ap.add_argument("--cli-candidates", type=int, default=2000, help="candidates of the synthetic chunk of the chunk-file leg")
    ap.add_argument("--ncu", action="store_true", help="profiling run under ncu: allow fewer warm-up steps "
                                                       "(numbers printed in this mode are NOT bench values)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.candidates:
        scale = args.candidates / float(cfg["parts"][0][1])
        cfg["parts"] = [(h, max(1, int(round(n * scale)))) for h, n in cfg["parts"]]
    if args.steps is None:
        args.steps = 20 if args.config == 1 else 5

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return
    if args.warmup < 3 and not args.ncu:
        args.warmup = 3

    import ctypes as C
    import numpy as np
    import torch
    import torch.distributed as dist
    from clairs_to_b200 import _lib, dist as cdist, synth, synth_weights
    from clairs_to_b200.engine import PIPELINE_LOW_BQ_CUT, Engine, packed_to_device
    from clairs_to_b200.host import tokenize_mpileup
    from clairs_to_b200.pileup_format import N_POS, PackedStream, pack_stream

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    cut = PIPELINE_LOW_BQ_CUT
    host_threads = max(1, len(os.sched_getaffinity(0)) // world)

    def pin(a):
        a = np.ascontiguousarray(a)
        return torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory()

    # ---- workload: per part an engine + this rank's shard of packed streams (device-resident and pinned host copies) ----
    t_gen = time.time()
    parts = []
    PIECE = 1000000                       # candidates per batch: one batch holds at most 2^31 reads (int32 CSR offsets)
    for pi, (n_heads, n_part) in enumerate(cfg["parts"]):
        n_loc = (cdist.shard_bounds(n_part, world, rank)[1] - cdist.shard_bounds(n_part, world, rank)[0]) if cfg["total"] else n_part
        aff_sd = synth_weights.synth_state_dict(synth_weights.aff_state_dict_shapes(n_heads), 100 + n_heads)
        neg_sd = synth_weights.synth_state_dict(synth_weights.neg_state_dict_shapes(n_heads), 200 + n_heads)
        eng = Engine(aff_sd, neg_sd, max_batch=args.max_batch, device=dev, likelihood=synthetic_likelihood(n_heads))
        if cfg["qual"]:
            eng.set_qual_thresholds(*cfg["qual"])
        for lo in range(0, n_loc, PIECE):
            n_piece = min(PIECE, n_loc - lo)
            # config 1 generates every site (as in round 1; also gives the text leg its indel sequences), the large
            # configs generate 50 000 distinct sites and tile them
            (aff, aff_aux), (neg, neg_aux) = synth.synth_pair_tiled(n_piece, 20241 + 7 * pi + rank + 101 * (lo // PIECE), cfg["platform"],
                                                                    base=200000 if args.config == 1 else 50000)
            if cfg["single_stream"]:
                aff, aff_aux = neg, neg_aux             # NEG is a symlink of AFF when both use --min-BQ 0 (run_clairs_to:1248-1252)
            streams = [aff] if cfg["single_stream"] else [aff, neg]
            host = [pack_stream(s, cut, n_threads=host_threads) for s in streams]
            pinned = [PackedStream(*[pin(a) for a in h.arrays()], h.n_groups, h.low_bq_cut, h.n_reads) for h in host]
            device = [packed_to_device(h, dev) for h in host]
            parts.append(dict(n=n_piece, n_total=(n_part if cfg["total"] else n_part * world) * n_piece // max(n_loc, 1), heads=n_heads, eng=eng,
                              dev=device, pinned=pinned, raw=(aff, aff_aux, neg, neg_aux) if args.config == 1 else None,
                              alg_bytes=sum(h.algorithmic_bytes() for h in host), h2d=sum(h.nbytes() for h in host),
                              reads=sum(h.n_reads for h in host)))
            del aff, neg, host
    t_gen = time.time() - t_gen
    if world > 1:
        for p in parts:
            p["sizes"] = cdist.exchange_sizes(p["n"], dev)        # shard lengths per batch, exchanged once
    lib = parts[0]["eng"].lib
    torch.cuda.synchronize()
    n_local = sum(p["n"] for p in parts)
    n_global = sum(n for _, n in cfg["parts"]) * (1 if cfg["total"] else world)

    enc_ev = []

    def step(timed):
        out = None
        for p in parts:
            eng = p["eng"]
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            enc = [eng.encode(d) for d in p["dev"]]
            ev[1].record()
            if timed:
                enc_ev.append(ev)
            (xa, da), (xn, dn) = enc[0], enc[-1]
            out = eng.predict(xa, da, xn, dn)
            if world > 1:
                # the path's single exchange: per-candidate results to rank 0
                cdist.gather_rows_padded(out['probs'].reshape(p["n"], -1), p["sizes"])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step(False)
    barrier()
    for p in parts:
        _lib.check(lib.cto_engine_profile(p["eng"].handle, int(os.environ.get('CTO_PROFILE_LEVEL', '2'))))   # idempotent per engine
    launches0 = lib.cto_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step(True)
    ev1.record()
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = lib.cto_launch_count() - launches0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    kinds = lib.cto_engine_profile_kinds()
    fam = {}
    engines = []
    for p in parts:
        if not any(p["eng"] is e for e, _ in engines):
            engines.append((p["eng"], sum(q["n"] for q in parts if q["eng"] is p["eng"])))
    for eng_k, n_k in engines:
        p = dict(eng=eng_k, n=n_k, heads=eng_k.n_heads)
        pms = (C.c_double * kinds)(); pcnt = (C.c_int64 * kinds)(); pfl = (C.c_double * kinds)()
        _lib.check(lib.cto_engine_profile_read(p["eng"].handle, pms, pcnt, pfl))
        _lib.check(lib.cto_engine_profile(p["eng"].handle, 0))
        for k in range(kinds):
            if pcnt[k]:
                name = lib.cto_engine_profile_name(k).decode() + ("" if len(engines) == 1 else "[%d heads]" % p["heads"])
                fam[name] = dict(name=name, ms_per_step=pms[k] / args.steps, launches_per_step=pcnt[k] / args.steps,
                                 flop_per_candidate=pfl[k], candidates_per_step=p["n"])
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = n_global * args.steps / (ms_max / 1e3)

    # ---- roofline: the dominant tensor-core kernel family, and the encoder against the HBM peak -----------------
    fam = list(fam.values())
    enc_ms = sum(a.elapsed_time(b) for a, b in enc_ev) / args.steps
    enc_bytes = sum(p["alg_bytes"] for p in parts)
    enc_real = sum(p["h2d"] + 2 * N_POS * 34 * p["n"] * len(p["dev"]) for p in parts)
    n_enc_launches = sum(len(p["dev"]) for p in parts)
    enc = dict(bound="hbm", achieved=enc_bytes / (enc_ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s", traffic=None,
               kernel="encode_pileup_kernel", ms_per_step=enc_ms, launches_per_step=n_enc_launches,
               algorithmic_bytes_per_step=enc_bytes, bytes_moved_per_step=enc_real,
               moved_gbs=enc_real / (enc_ms * 1e-3) / 1e9, peak_source=peaks["source"],
               note="algorithmic bytes = SURVEY.md 8(d): 3 B per read + 5 B per window slot + 2244 B out per candidate-stream; the "
                    "bit-plane packed input moves 1 B per read, so bytes_moved < algorithmic bytes")
    enc["frac"] = enc["achieved"] / enc["peak"]
    enc_tr = ncu_traffic("encoder", n_local)
    if enc_tr is not None:                              # the capture holds one launch per stream; report their mean
        enc["traffic"] = enc_tr
        enc["algorithmic_bytes_per_launch"] = enc_bytes / n_enc_launches
    tensor_fams = [f for f in fam if f["flop_per_candidate"] > 0]
    for f in tensor_fams:                                # algorithmic rate of every contraction family, against the same peak
        f["tflops"] = f["flop_per_candidate"] * f["candidates_per_step"] / (f["ms_per_step"] * 1e-3) / 1e12
        f["frac"] = f["tflops"] / peaks["tensor_sustained"]
    top = max(tensor_fams, key=lambda f: f["ms_per_step"])
    per_launch_ms = top["ms_per_step"] / top["launches_per_step"]
    cand_per_launch = top["candidates_per_step"] / top["launches_per_step"]
    flop_launch = top["flop_per_candidate"] * top["candidates_per_step"] / top["launches_per_step"]
    achieved = flop_launch / (per_launch_ms * 1e-3) / 1e12
    roofline = dict(bound="tensor", achieved=achieved, peak=peaks["tensor_sustained"], unit="TFLOP/s",
                    frac=achieved / peaks["tensor_sustained"], traffic=ncu_traffic(top["name"].split("[")[0], cand_per_launch),
                    kernel=top["name"], ms_per_launch=per_launch_ms, launches_per_step=top["launches_per_step"],
                    algorithmic_flop_per_launch=flop_launch,
                    peak_source=peaks["source"] + ", dense bf16 sustained; the kernel issues 3 bf16 MMAs per algorithmic "
                                                  "multiply-add (bf16x3 split for fp32-grade accuracy), so 1/3 is its ceiling",
                    families=fam, encoder=enc)
    whole = sum(f["flop_per_candidate"] * f["candidates_per_step"] for f in fam if not f["name"].startswith("aff_forward"))
    roofline["whole_pass"] = dict(tflops_per_gpu=whole / (ms_max / args.steps * 1e-3) / 1e12,
                                  note="algorithmic FLOP of every timed kernel family of one rank / step time")

    # ---- end to end through the host-buffer C-ABI call (packed pinned host buffers -> pinned host results) -------
    e2e = None
    if not args.no_e2e:
        outs = []
        for p in parts:
            h = p["heads"]
            outs.append(dict(probs=torch.empty((p["n"], 2 * h, 2), dtype=torch.float32).pin_memory(),
                             post=torch.empty((p["n"], h), dtype=torch.float64).pin_memory(),
                             call=torch.empty((p["n"],), dtype=torch.int32).pin_memory(),
                             qual=torch.empty((p["n"],), dtype=torch.float64).pin_memory(),
                             filter=torch.empty((p["n"],), dtype=torch.int32).pin_memory()))

        def e2e_step():
            for p, o in zip(parts, outs):
                p["eng"].run_sites_host(p["pinned"][0], None if cfg["single_stream"] else p["pinned"][1], out=o)

        for _ in range(2):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k_e2e = max(2, min(args.steps, 5))
        e0.record()
        for _ in range(k_e2e):
            e2e_step()
        e1.record()
        barrier()
        e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        d2h = sum(sum(t.numel() * t.element_size() for t in o.values()) for o in outs)
        e2e = dict(value=n_global * k_e2e / (float(e_ms.item()) / 1e3), unit=UNIT, h2d_bytes_per_step=sum(p["h2d"] for p in parts),
                   d2h_bytes_per_step=d2h, steps=k_e2e,
                   api="cto_run_sites_host (pinned host buffers in the packed 1-byte-per-read layout the tokenizer emits)")

        # ---- the same from mpileup TEXT: tokenizer (multi-threaded host C++) + packer + the call above --------------
        p0 = parts[0]
        if args.config == 1 and not args.no_text and p0["raw"] is not None and p0["raw"][1] is not None:
            aff, aff_aux, neg, neg_aux = p0["raw"]
            texts = [synth.render_mpileup_text(s, a) for s, a in ((aff, aff_aux), (neg, neg_aux))]
            ref = ''.join("ACGT"[c] for c in neg.ref_code)
            cands = np.arange(1001 + 16, 1001 + neg.n_rows, N_POS, dtype=np.int64)
            win = np.arange(neg.n_rows, dtype=np.int32)

            def text_step(timing=None):
                t0 = time.perf_counter()
                packed = []
                for txt in texts:
                    tok = tokenize_mpileup(txt, ref, 1001, cands, 60, n_threads=host_threads)
                    tok.stream.win_pos = win
                    packed.append(pack_stream(tok.stream, cut, n_threads=host_threads))
                t1 = time.perf_counter()
                p0["eng"].run_sites_host(packed[0], packed[1], out=outs[0])
                if timing is not None:
                    timing.append((t1 - t0, time.perf_counter() - t1))

            text_step()
            barrier()
            tt = []
            w0 = time.perf_counter()
            k_txt = 2
            for _ in range(k_txt):
                text_step(tt)
            torch.cuda.synchronize()
            w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(w, op=dist.ReduceOp.MAX)
            text_bytes = sum(len(t) for t in texts)
            host_tok = dict(value=p0["n"] * world * k_txt / float(w.item()), unit=UNIT, steps=k_txt,
                            text_bytes_per_step=text_bytes, host_threads=host_threads,
                            tokenize_pack_s=sum(a for a, _ in tt) / k_txt, gpu_call_s=sum(b for _, b in tt) / k_txt,
                            tokenizer_gb_per_s=text_bytes / 1e9 / (sum(a for a, _ in tt) / k_txt),
                            api="mpileup text (host memory) -> cto_tokenize_mpileup + cto_pack_reads (host threads) -> cto_run_sites_host")
            # ---- and with the DEVICE tokenizer: the host only copies the text (pinned), rows are indexed and tokenized in HBM --------
            pinned = [torch.frombuffer(bytearray(t), dtype=torch.uint8).pin_memory() for t in texts]
            refb = ref.encode()
            out_txt = {k: torch.empty_like(v).pin_memory() for k, v in outs[0].items()}

            def dev_text_step():
                return p0["eng"].run_sites_text(pinned[0], pinned[1], refb, 1001, cands, cut, out=out_txt)

            got = dev_text_step()                                # outs[0]: what the host-tokenizer path just produced for the same sites
            torch.cuda.synchronize()
            for k, v in outs[0].items():
                assert torch.equal(got[k], v), "device-tokenized %s differ from the host-tokenized ones" % k
            barrier()
            k_dev = 5
            w0 = time.perf_counter()
            for _ in range(k_dev):
                dev_text_step()
            torch.cuda.synchronize()
            w = torch.tensor([time.perf_counter() - w0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(w, op=dist.ReduceOp.MAX)
            e2e["text"] = dict(value=p0["n"] * world * k_dev / float(w.item()), unit=UNIT, steps=k_dev, text_bytes_per_step=text_bytes,
                               h2d_bytes_per_step=text_bytes + len(refb) + 8 * len(cands), d2h_bytes_per_step=sum(t.numel() * t.element_size() for t in out_txt.values()),
                               text_gb_per_s=text_bytes / 1e9 * k_dev / float(w.item()),
                               api="Engine.run_sites_text: mpileup text (pinned host memory) -> H2D per engine chunk on a side stream -> "
                                   "cto_index_rows + cto_tokenize_count / _write + cto_window_table (device) -> encoder -> AFF + NEG -> "
                                   "probabilities D2H; bit-identical to the host-tokenizer path",
                               host_tokenizer=host_tok)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        r = cpu_baseline.run(per_proc=args.cpu_sample, platform=cfg["platform"], mix=cpu_mix(cfg), single_stream=cfg["single_stream"])
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"], encoder_share=r["encoder_share"])

    cli = None
    if rank == 0 and world == 1 and args.config == 1 and not args.no_cli:
        cli = cli_leg(args.cli_candidates, 4, synthetic_likelihood(4), host_threads)

    scan = None
    if rank == 0 and world == 1 and args.config == 1 and not args.no_scan and parts[0]["raw"] is not None and parts[0]["raw"][3] is not None:
        launches_scan0 = lib.cto_launch_count()
        scan = scan_leg(parts[0]["raw"], peaks, dev)
        scan["gpu_launches"] = int(lib.cto_launch_count() - launches_scan0)

    filters = None
    if rank == 0 and world == 1 and args.config == 1 and not args.no_filters:
        launches_f0 = lib.cto_launch_count()
        filters = hard_filter_leg(peaks, dev)
        filters["gpu_launches"] = int(lib.cto_launch_count() - launches_f0)

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="strong" if cfg["total"] else "weak",
                    vs_baseline=None, dtype="f32 (tensor-core contractions as bf16x3 split products, f32 accumulate)", data="synthetic",
                    config=dict(workload=cfg["workload"], candidates_per_gpu=n_local, candidates_total=n_global, platform=cfg["literal"],
                                heads=[h for h, _ in cfg["parts"]], engine_chunk=args.max_batch, weights="seeded random init",
                                l2_policy="inputs (%.0f MB per step and GPU) larger than the 126 MB L2" % (sum(p["h2d"] for p in parts) / 1e6),
                                parallelism="candidates sharded x%d, one gather of probabilities" % world,
                                streams="value: AFF and NEG back to back on one stream with the per-kernel events of the roofline on; "
                                        "e2e: AFF on a second stream beside NEG (cto_engine_set_overlap, bit-identical)",
                                host_cores_bound_to_gpu_numa_node=numa, datagen_s=round(t_gen, 1)),
                    roofline=roofline, cpu_baseline=cpu, e2e=e2e, cli=cli, candidate_scan=scan, hard_filters=filters, gpu_launches=int(launches), clocks=clocks)
        print(json.dumps(line))
    for eng_k, _ in engines:
        eng_k.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
