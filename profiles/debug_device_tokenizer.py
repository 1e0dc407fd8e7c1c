"""Per-call time of the device tokenizer (both streams of the bench workload) over repeated calls, and where the time of one
call goes (wall clock around every step of device_tokenizer.tokenize_text_device, stream synchronised after each)."""
import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = [tokenize_text_device(b, len(t), ref_dev, 1001, 30, cand_dev)[0] for b, t in zip(bufs, texts)]
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("iter %2d  %.1f ms  reserved %d MB" % (it, dt * 1e3, torch.cuda.memory_reserved() >> 20), flush=True)

lib = _lib.lib()
p = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def step(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print("    %-42s %7.3f ms" % (label, (time.perf_counter() - t0) * 1e3)); return r
for rep in range(2):
    print("one call, NEG stream, step by step (rep %d):" % rep)
    text_dev, n_bytes = bufs[1], len(texts[1])
    nr = C.c_int64()
    step("cto_index_rows (count only)", lambda: lib.cto_index_rows(p(text_dev), n_bytes, None, 0, C.byref(nr), st))
    row_off = step("torch.empty row_off", lambda: torch.empty(nr.value + 1, dtype=torch.int64, device=dev))
    step("cto_index_rows (count + offsets)", lambda: lib.cto_index_rows(p(text_dev), n_bytes, p(row_off), nr.value, C.byref(nr), st))
    n_rows = nr.value
    arrs = step("torch.empty x4", lambda: (torch.empty(n_rows, dtype=torch.int32, device=dev), torch.empty(n_rows, dtype=torch.uint8, device=dev),
                                           torch.empty(n_rows + 1, dtype=torch.int32, device=dev), torch.empty(n_rows + 1, dtype=torch.int32, device=dev)))
    ng, ni = C.c_int64(), C.c_int64()
    step("cto_tokenize_count", lambda: lib.cto_tokenize_count(p(text_dev), n_bytes, p(row_off), n_rows, p(ref_dev), 1001, len(ref), p(arrs[0]), p(arrs[1]), p(arrs[2]), p(arrs[3]), C.byref(ng), C.byref(ni), st))
    planes = step("torch.empty planes + zero tail", lambda: torch.empty(((ng.value * 8 + 15) & ~15) + 16, dtype=torch.uint8, device=dev))
    planes[ng.value * 8:].zero_()
    ie = step("torch.empty ind_entry", lambda: torch.empty(max(ni.value, 1), dtype=torch.int32, device=dev))
    step("cto_tokenize_write", lambda: lib.cto_tokenize_write(p(text_dev), n_bytes, p(row_off), n_rows, p(ref_dev), 1001, len(ref), 30, 60, p(arrs[2]), p(arrs[3]), p(planes), p(ie), st))
    win = torch.empty(len(cands) * 33, dtype=torch.int32, device=dev)
    step("cto_window_table", lambda: lib.cto_window_table(p(arrs[0]), n_rows, p(cand_dev), len(cands), p(win), st))
