#!/usr/bin/env python
"""bench.py -- candidate sites/sec through the per-candidate hot path (pileup encoder + AFF + NEG
+ posterior) on N B200s, next to the CPU port of the reference timed on the same box.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): 100 000 synthetic ONT-shape candidate sites per GPU, SNV
AFF+NEG models with seeded random weights, synthetic likelihood tables.  One step = one pass of
the whole hot path over the batch.  Prints ONE JSON line (rank 0).
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
WORKLOAD = "BASELINE configs[1]: synthetic 100k ONT-shape candidate sites, pileup SNV model, AFF+NEG (+posterior)"
UNIT = "candidate sites/s"


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


TRAFFIC_KERNEL = {"neg_gru2_recurrent": "gru3_kernel<192", "neg_gru1_recurrent": "gru1_fused_kernel",
                  "encoder": "encode_pileup_kernel", "aff_stage1_fused": "aff_stage1_kernel"}


def ncu_traffic(family, candidates_per_launch):
    """DRAM bytes per launch of the kernel behind a profile family, from the committed ncu capture
    (profiles/r1_traffic.json, written by profiles/summarize.py traffic), scaled linearly from the capture's
    launch size to this run's mean launch size (every byte these kernels move is per candidate); None when the
    capture does not hold that kernel."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    key = TRAFFIC_KERNEL.get(family)
    if not key or not os.path.exists(path):
        return None
    t = json.load(open(path))
    for name, k in t["kernels"].items():
        if key in name:
            return k["dram_bytes"] * float(candidates_per_launch) / float(t["candidates_per_launch"])
    return None


def synthetic_likelihood(n_heads, seed=1):
    import numpy as np
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(0.05, 0.95, size=(10 * n_heads, 10)),
                           np.sort(rng.uniform(0.02, 0.98, size=(2 * n_heads, 10)), axis=1)])


def run_reference(args, rank, world):
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
            r = cpu_baseline.run(per_proc=per_proc, procs=cores, seed=9000 + 17 * step, pool=pool)
            if step >= args.warmup:
                times.append(r["seconds"])
                vals.append(r["value"])
            last = r
    finally:
        pool.close()
        pool.join()
    value = sum(vals) / len(vals)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * sum(times) / len(times), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32 (torch CPU)", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, platform="ont_r10_dorado_sup_5khz", heads=4, weights="seeded random init",
                            candidates_per_step=per_proc * cores,
                            sample="each step is a bounded sample of the workload: %d candidates per host process x %d "
                                   "single-thread processes (the reference's own process model)" % (per_proc, cores)),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=last["sample"]),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--candidates", type=int, default=100000, help="candidate sites per GPU per step")
    ap.add_argument("--max-batch", type=int, default=37888, help="engine chunk (candidates per network pass)")
    ap.add_argument("--cpu-sample", type=int, default=300, help="CPU baseline: candidates per host process")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ncu", action="store_true", help="profiling run under ncu: allow fewer warm-up steps "
                                                       "(numbers printed in this mode are NOT bench values)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3 and not args.ncu:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    from clairs_to_b200 import _lib, dist as cdist, synth
    from clairs_to_b200.engine import PIPELINE_LOW_BQ_CUT, Engine, stream_to_device
    from clairs_to_b200.pileup_format import PileupStream
    from clairs_to_b200 import synth_weights

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    n = args.candidates
    n_heads = 4
    literal = "ont_r10_dorado_sup_5khz"
    cut = PIPELINE_LOW_BQ_CUT

    t_gen = time.time()
    aff, neg = synth.synth_pair_large(n, 20241 + rank, 'ont')
    t_gen = time.time() - t_gen
    aff_sd = synth_weights.synth_state_dict(synth_weights.aff_state_dict_shapes(n_heads), 100 + n_heads)
    neg_sd = synth_weights.synth_state_dict(synth_weights.neg_state_dict_shapes(n_heads), 200 + n_heads)
    eng = Engine(aff_sd, neg_sd, max_batch=args.max_batch, device=dev, likelihood=synthetic_likelihood(n_heads))
    lib = eng.lib
    if os.environ.get("CTO_DEBUG"):
        lib.cto_debug_set(int(os.environ["CTO_DEBUG"]))
    d_aff, d_neg = stream_to_device(aff, dev), stream_to_device(neg, dev)
    torch.cuda.synchronize()

    enc_ev = []

    def step(timed):
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s2 = torch.cuda.Event(enable_timing=True)
        s0.record()
        xa, da = eng.encode(d_aff, cut)
        s1.record()
        xn, dn = eng.encode(d_neg, cut)
        s2.record()
        if timed:
            enc_ev.append((s0, s1, s2))
        out = eng.predict(xa, da, xn, dn)
        if world > 1:
            # the path's single exchange: per-candidate results to rank 0
            cdist.gather_rows(out['probs'].reshape(n, -1), n * world)
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
    _lib.check(lib.cto_engine_profile(eng.handle, int(os.environ.get('CTO_PROFILE_LEVEL', '1'))))
    launches0 = lib.cto_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        out = step(True)
    ev1.record()
    barrier()
    wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    launches = lib.cto_launch_count() - launches0
    clocks = sampler.stop(wall0, wall1) if rank == 0 else None
    kinds = lib.cto_engine_profile_kinds()
    import ctypes as C
    pms = (C.c_double * kinds)(); pcnt = (C.c_int64 * kinds)(); pfl = (C.c_double * kinds)()
    _lib.check(lib.cto_engine_profile_read(eng.handle, pms, pcnt, pfl))
    _lib.check(lib.cto_engine_profile(eng.handle, 0))
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = n * world * args.steps / (ms_max / 1e3)

    # ---- roofline of the dominant kernel family + the encoder -----------------------------------
    fam = []
    for k in range(kinds):
        if pcnt[k]:
            fam.append(dict(name=lib.cto_engine_profile_name(k).decode(), ms_per_step=pms[k] / args.steps,
                            launches_per_step=pcnt[k] / args.steps, flop_per_candidate=pfl[k]))
    enc_aff_ms = sum(a.elapsed_time(b) for a, b, _ in enc_ev) / len(enc_ev)
    enc_neg_ms = sum(b.elapsed_time(c) for _, b, c in enc_ev) / len(enc_ev)
    enc_bytes = aff.algorithmic_bytes() + neg.algorithmic_bytes()
    enc = dict(bound="hbm", achieved=enc_bytes / ((enc_aff_ms + enc_neg_ms) * 1e-3) / 1e9, peak=peaks["hbm"],
               unit="GB/s", traffic=None, kernel="encode_pileup_kernel", ms_per_step=enc_aff_ms + enc_neg_ms,
               launches_per_step=2,
               algorithmic_bytes_per_step=enc_bytes, peak_source=peaks["source"])
    enc["frac"] = enc["achieved"] / enc["peak"]
    enc_tr = ncu_traffic("encoder", n)
    if enc_tr is not None:                              # the capture holds one launch per stream; report their mean
        enc["traffic"] = enc_tr
        enc["algorithmic_bytes_per_launch"] = enc_bytes / 2
    tensor_fams = [f for f in fam if f["flop_per_candidate"] > 0]
    top = max(tensor_fams, key=lambda f: f["ms_per_step"])
    per_launch_ms = top["ms_per_step"] / top["launches_per_step"]
    cand_per_launch = n / top["launches_per_step"]
    achieved = top["flop_per_candidate"] * cand_per_launch / (per_launch_ms * 1e-3) / 1e12
    roofline = dict(bound="tensor", achieved=achieved, peak=peaks["tensor_sustained"], unit="TFLOP/s",
                    frac=achieved / peaks["tensor_sustained"], traffic=ncu_traffic(top["name"], cand_per_launch),
                    kernel=top["name"], ms_per_launch=per_launch_ms, launches_per_step=top["launches_per_step"],
                    algorithmic_flop_per_launch=top["flop_per_candidate"] * cand_per_launch,
                    peak_source=peaks["source"] + ", dense bf16 sustained; the kernel issues 3 bf16 MMAs per algorithmic "
                                                  "multiply-add (bf16x3 split for fp32-grade accuracy), so 1/3 is its ceiling",
                    families=fam, encoder=enc)

    # ---- end to end through the host-buffer C-ABI call ------------------------------------------
    e2e = None
    if not args.no_e2e:
        def pin(s):
            arrs = []
            for a in s.arrays():
                a = np.ascontiguousarray(a)
                t = torch.from_numpy(a.view(np.int32) if a.dtype == np.uint32 else a).pin_memory()
                arrs.append(t)
            return PileupStream(*arrs)
        p_aff, p_neg = pin(aff), pin(neg)
        h = eng.n_heads
        outb = dict(probs=torch.empty((n, 2 * h, 2), dtype=torch.float32).pin_memory(),
                    post=torch.empty((n, h), dtype=torch.float64).pin_memory(),
                    call=torch.empty((n,), dtype=torch.int32).pin_memory())
        for _ in range(2):
            eng.run_sites_host(p_aff, p_neg, cut, out=outb)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k_e2e = max(2, min(args.steps, 5))
        for _ in range(k_e2e):
            eng.run_sites_host(p_aff, p_neg, cut, out=outb)
        e1.record()
        barrier()
        e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e_ms, op=dist.ReduceOp.MAX)
        h2d = aff.nbytes() + neg.nbytes()
        d2h = sum(t.numel() * t.element_size() for t in outb.values())
        e2e = dict(value=n * world * k_e2e / (float(e_ms.item()) / 1e3), unit=UNIT, h2d_bytes_per_step=h2d,
                   d2h_bytes_per_step=d2h, steps=k_e2e, api="cto_run_sites_host (pinned host buffers)")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline
        r = cpu_baseline.run(per_proc=args.cpu_sample)
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind=r["kind"], sample=r["sample"])

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32 (tensor-core contractions as bf16x3 split products, f32 accumulate)", data="synthetic",
                    config=dict(workload=WORKLOAD, candidates_per_gpu=n, platform=literal,
                                heads=n_heads, engine_chunk=args.max_batch, weights="seeded random init",
                                l2_policy="inputs (%.0f MB per step) larger than the 126 MB L2" % ((aff.nbytes() + neg.nbytes()) / 1e6),
                                parallelism="candidates sharded x%d, one gather of probabilities" % world,
                                datagen_s=round(t_gen, 1)),
                    roofline=roofline, cpu_baseline=cpu, e2e=e2e, gpu_launches=int(launches), clocks=clocks)
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
