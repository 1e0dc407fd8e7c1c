"""Where the time of Engine.run_sites_text goes (mpileup text in pinned host memory -> probabilities): per-stage CUDA-event
times for the bench workload (100 000 ONT-shape sites, two streams, 1.02 GB of text), one piece."""
import sys, os, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairs_to_b200 import synth, _lib
from clairs_to_b200 import synth_weights as sw
from clairs_to_b200.engine import Engine, encode_pileup
from clairs_to_b200.device_tokenizer import tokenize_text_device

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
(aff, aa), (neg, na) = synth.synth_pair_tiled(n, 20241, "ont", base=200000)
texts = [synth.render_mpileup_text(s, a) for s, a in ((aff, aa), (neg, na))]
ref = ''.join("ACGT"[c] for c in neg.ref_code).encode()
cands = np.arange(1001 + 16, 1001 + neg.n_rows, 33, dtype=np.int64)
pinned = [torch.frombuffer(bytearray(t), dtype=torch.uint8).pin_memory() for t in texts]
eng = Engine(sw.synth_state_dict(sw.aff_state_dict_shapes(4), 104), sw.synth_state_dict(sw.neg_state_dict_shapes(4), 204), max_batch=37888)
dev = eng.device
lib = _lib.lib()
p = lambda t: C.c_void_p(t.data_ptr())
stream = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)

def timed(label, fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    print("%-34s %8.2f ms (device)  %8.2f ms (wall)" % (label, e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) * 1e3 / reps))
    return out

bufs = [torch.empty(t.numel() + 32, dtype=torch.uint8, device=dev) for t in pinned]
def h2d():
    for b, t in zip(bufs, pinned): b[:t.numel()].copy_(t, non_blocking=True)
timed("H2D of both texts (%.2f GB)" % (sum(t.numel() for t in pinned) / 1e9), h2d)
ref_dev = torch.frombuffer(bytearray(ref), dtype=torch.uint8).to(dev)
cand_dev = torch.from_numpy(cands).to(dev)
for k, name in ((0, "AFF stream"), (1, "NEG stream")):
    nb = pinned[k].numel()
    nrows = C.c_int64()
    def index():
        lib.cto_index_rows(p(bufs[k]), nb, None, 0, C.byref(nrows), stream())
        ro = torch.empty(nrows.value + 1, dtype=torch.int64, device=dev)
        lib.cto_index_rows(p(bufs[k]), nb, p(ro), nrows.value, C.byref(nrows), stream())
        return ro
    ro = timed(name + ": row index (count + write)", index)
    nr = nrows.value
    row_pos = torch.empty(nr, dtype=torch.int32, device=dev); rc = torch.empty(nr, dtype=torch.uint8, device=dev)
    go = torch.empty(nr + 1, dtype=torch.int32, device=dev); io = torch.empty(nr + 1, dtype=torch.int32, device=dev)
    ng, ni = C.c_int64(), C.c_int64()
    timed(name + ": tokenize_count", lambda: lib.cto_tokenize_count(p(bufs[k]), nb, p(ro), nr, p(ref_dev), 1001, len(ref), p(row_pos), p(rc), p(go), p(io), C.byref(ng), C.byref(ni), stream()))
    planes = torch.zeros(((ng.value * 8 + 15) & ~15) + 16, dtype=torch.uint8, device=dev); ie = torch.empty(max(ni.value, 1), dtype=torch.int32, device=dev)
    timed(name + ": tokenize_write", lambda: lib.cto_tokenize_write(p(bufs[k]), nb, p(ro), nr, p(ref_dev), 1001, len(ref), 30, 60, p(go), p(io), p(planes), p(ie), stream()))
packed = [tokenize_text_device(b, t.numel(), ref_dev, 1001, 30, cand_dev)[0] for b, t in zip(bufs, pinned)]
timed("tokenize_text_device x2 (all of the above + window table)", lambda: [tokenize_text_device(b, t.numel(), ref_dev, 1001, 30, cand_dev)[0] for b, t in zip(bufs, pinned)])
timed("run_sites (encode x2 + AFF + NEG)", lambda: eng.run_sites(packed[0], packed[1], 30))
for pieces in (1, 2, 4, 8, None):
    timed("run_sites_text, %s piece(s)" % (pieces if pieces else "one per engine chunk (3)"), lambda: eng.run_sites_text(pinned[0], pinned[1], ref, 1001, cands, 30, pieces=pieces))
