import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from clairs_to_b200 import synth, _lib
from clairs_to_b200.device_tokenizer import tokenize_text_device
n = 100000
(aff, aa), (neg, na) = synth.synth_pair_tiled(n, 20241, "ont", base=200000)
texts = [synth.render_mpileup_text(s, a) for s, a in ((aff, aa), (neg, na))]
ref = ''.join("ACGT"[c] for c in neg.ref_code).encode()
cands = np.arange(1001 + 16, 1001 + neg.n_rows, 33, dtype=np.int64)
dev = torch.device("cuda:0")
bufs = []
for t in texts:
    b = torch.zeros(len(t) + 32, dtype=torch.uint8, device=dev); b[:len(t)] = torch.frombuffer(bytearray(t), dtype=torch.uint8).to(dev); bufs.append(b)
ref_dev = torch.frombuffer(bytearray(ref), dtype=torch.uint8).to(dev)
cand_dev = torch.from_numpy(cands).to(dev)
def stats():
    s = torch.cuda.memory_stats()
    return s.get("num_device_alloc", 0), s.get("num_device_free", 0), s.get("num_alloc_retries", 0), torch.cuda.memory_reserved() >> 20
for it in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = [tokenize_text_device(b, len(t), ref_dev, 1001, 30, cand_dev)[0] for b, t in zip(bufs, texts)]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("iter %2d  %.1f ms  device_alloc/free/retries/reservedMB %s" % (it, dt * 1e3, stats()), flush=True)
    if it == 5: del out
