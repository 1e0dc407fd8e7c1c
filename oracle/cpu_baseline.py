"""CPU baseline for bench.py: the oracle port of the reference's path timed on host cores.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference is Python + torch CPU
and cannot travel to the GPU box, so the baseline is its port (``kind: "port"``): P independent
single-threaded processes, which is exactly the reference's own parallelism model (GNU parallel
-j <threads> over chunk files, one torch thread per process: run_clairs_to:1274-1276,
clairs/predict.py:475).  Each process does, per candidate: the encoder over both mpileup streams
(CPython; production runs it under pypy3, absent here), depth rescale, AFF + NEG forward in
mini-batches of 250 (shared/param.py:85), softmax and the Bayes combine.  Text I/O (gzip chunk
files) is NOT included, which favours the baseline.
"""

from __future__ import annotations

import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(args):
    seed, n, n_heads, platform, single_stream = args
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from clairs_to_b200 import synth
    from oracle import nn_oracle, pileup_oracle, posterior_oracle
    torch.set_num_threads(1)
    literal = "ont"          # create_tensor() always decodes with the callee default (CT:499-511)
    (aff, aff_aux), (neg, neg_aux) = synth.synth_pair(n, seed, platform)
    pairs = ((neg, neg_aux),) if single_stream else ((aff, aff_aux), (neg, neg_aux))   # Illumina: one tensor feeds both nets
    texts = [synth.render_mpileup(s, a) for s, a in pairs]
    refs = ["ACGT"[int(c)] for c in neg.ref_code]
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads), 100 + n_heads)
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 200 + n_heads)
    rng = np.random.default_rng(1)
    lk = np.concatenate([rng.uniform(0.05, 0.95, size=(10 * n_heads, 10)),
                         np.sort(rng.uniform(0.02, 0.98, size=(2 * n_heads, 10)), axis=1)])
    mats, ea, en = posterior_oracle.load_likelihood(lk, n_heads)

    t0 = time.perf_counter()
    xs = []
    for rows in texts:
        vecs, depths = [], []
        for i, text in enumerate(rows):
            _, bases, bq, mq = pileup_oracle.parse_mpileup_row(text)
            centre = (i % 33) == 16
            v, alt = pileup_oracle.position_vector(bases, mq, bq, refs[i], centre, refs[i], literal)
            vecs.append(v)
            if centre:
                depths.append(posterior_oracle.depth_from_alt_info(alt))
        t = np.array(vecs, dtype=np.int16).reshape(n, 33, 34)
        xs.append(np.stack([posterior_oracle.rescale_tensor(a, d) for a, d in zip(t, depths)]))
    t_enc = time.perf_counter() - t0
    posts = []
    for lo in range(0, n, 250):
        pa = nn_oracle.softmax_heads(nn_oracle.aff_forward(xs[0][lo:lo + 250], aff_sd)).numpy()
        pn = nn_oracle.softmax_heads(nn_oracle.neg_forward(xs[-1][lo:lo + 250], neg_sd)).numpy()
        for k in range(pa.shape[0]):
            p8 = [float("{:0.8f}".format(min(v, 0.99999999))) for v in pa[k, :, 1]]
            q8 = [float("{:0.8f}".format(min(v, 0.99999999))) for v in pn[k, :, 1]]
            posts.append(posterior_oracle.posterior(p8, q8, mats, ea, en))
    total = time.perf_counter() - t0
    return total, t_enc, len(posts)


def make_pool(procs=None):
    """A pool of single-thread worker processes (spawned once: every worker imports torch, which costs seconds)."""
    import multiprocessing as mp
    procs = procs or len(os.sched_getaffinity(0))
    return mp.get_context("spawn").Pool(procs), procs


def run(per_proc=300, procs=None, n_heads=4, seed=9000, pool=None, platform='ont', mix=None, single_stream=False):
    """Returns dict(value=candidates/s over all processes, cores, sample, seconds, encoder_share, candidates).  `pool`
    (from make_pool) is reused when given, so that a multi-step run pays the process start-up once.  `mix`: list of
    (n_heads, fraction) -- the model pairs of the workload (SNV 4 heads, indel 6 heads) and their share of candidates;
    every process runs each pair on its share, one after the other."""
    own = pool is None
    if own:
        pool, procs = make_pool(procs)
    else:
        procs = procs or len(os.sched_getaffinity(0))
    mix = mix or [(n_heads, 1.0)]
    slowest, done, enc, tot = 0.0, 0, 0.0, 0.0
    try:
        for heads, frac in mix:
            n = max(1, int(round(per_proc * frac)))
            res = pool.map(_worker, [(seed + i, n, heads, platform, single_stream) for i in range(procs)])
            slowest += max(r[0] for r in res)
            done += sum(r[2] for r in res)
            enc += sum(r[1] for r in res)
            tot += sum(r[0] for r in res)
    finally:
        if own:
            pool.close()
            pool.join()
    return dict(value=done / slowest, unit="candidate sites/s", cores=procs, kind="port", candidates=done,
                sample="%d candidates (%d per process x %d single-thread processes; model pairs %s), %s-shape synthetic, "
                       "mpileup text parse + encoder(%d stream%s, CPython)+rescale+AFF+NEG(torch CPU fp32)+posterior; no gzip I/O"
                       % (done, per_proc, procs, "+".join("%d-head" % h for h, _ in mix), platform, 1 if single_stream else 2,
                          "" if single_stream else "s"),
                seconds=slowest, encoder_share=enc / tot)


if __name__ == "__main__":
    print(run(per_proc=int(sys.argv[1]) if len(sys.argv) > 1 else 100))
