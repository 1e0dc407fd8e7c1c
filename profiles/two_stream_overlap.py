"""Would running the AFF and the NEG forward of a chunk on two streams pay?  Both networks only meet at the posterior.
Times forward_aff + forward_neg of one engine chunk back to back on one stream, and concurrently on two."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairs_to_b200.engine import Engine
from oracle import nn_oracle
aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104)
neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 204)
n = 37888
eng = Engine(aff_sd, neg_sd, max_batch=n)
x = torch.from_numpy(np.random.default_rng(0).integers(-50, 51, size=(n, 33, 34)).astype(np.float32)).cuda()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def seq():
    a = eng.forward_aff(x); b = eng.forward_neg(x); return a, b
def par():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): a = eng.forward_aff(x)
    with torch.cuda.stream(s2): b = eng.forward_neg(x)
    cur.wait_stream(s1); cur.wait_stream(s2)
    return a, b
ra, rb = seq(); torch.cuda.synchronize()
for name, fn in (("one stream", seq), ("two streams", par), ("one stream", seq), ("two streams", par)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): a, b = fn()
    e1.record(); torch.cuda.synchronize()
    print("%-12s %.3f ms per chunk of %d   (AFF equal %s, NEG equal %s)" % (name, e0.elapsed_time(e1) / 5, n, torch.equal(a, ra), torch.equal(b, rb)))
