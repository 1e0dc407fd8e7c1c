"""Run-to-run consistency of the whole NEG forward in engine modes 2 (one-chain GRU) and 1 (two-chain GRU): fresh engines,
alternating modes, bit-exact comparison of the logits AND of the intermediate workspace tensors against the first mode-2
result, to name the first kernel whose output differs."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairs_to_b200.engine import Engine
from oracle import nn_oracle
aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104)
neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 204)
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
names = {0: "xproj2 (proj2 GEMM out)", 1: "o1_hi (GRU1 out)", 2: "o1_mid", 3: "o2_hi (GRU2 out)", 4: "o2_mid"}
bad = {1: 0, 2: 0}; tot = {1: 0, 2: 0}
for r in range(rounds):
    for n in (1000, 4096, 700):
        eng = Engine(aff_sd, neg_sd, max_batch=n)
        bp = (n + 127) // 128 * 128
        x = torch.from_numpy(np.random.default_rng(n).integers(-50, 51, size=(n, 33, 34)).astype(np.float32)).cuda()
        eng.set_tensor_cores(2); ref = eng.forward_neg(x).clone()
        ref_ws = {k: eng.workspace(k) for k in names}
        for mode in (1, 2, 1, 1, 2):
            eng.set_tensor_cores(mode)
            got = eng.forward_neg(x).clone(); torch.cuda.synchronize()
            tot[mode] += 1
            per = (got - ref).abs().reshape(n, -1).max(1).values.cpu().numpy()
            b = np.nonzero(per > 0)[0]
            if b.size:
                bad[mode] += 1
                print("round", r, "n", n, "mode", mode, "bad candidates", b[:8], "count", b.size, "max", per.max())
                for k in names:
                    w = eng.workspace(k)
                    d = (w != ref_ws[k]).nonzero().flatten().cpu().numpy()
                    if d.size == 0:
                        print("     %-26s identical" % names[k]); continue
                    if k == 0:
                        el = np.unique(d // 4); row, col = el // (33 * bp), el % (33 * bp)
                        print("     %-26s %d elements differ: rows(dir,gate,unit) %s t %s cand %s" % (names[k], el.size, np.unique(row)[:10], np.unique(col // bp)[:6], np.unique(col % bp)[:12]))
                    elif k in (1, 2):
                        el = np.unique(d // 2); rowi, u = el // 256, el % 256
                        print("     %-26s %d elements differ: t %s cand %s units %s" % (names[k], el.size, np.unique(rowi // bp)[:6], np.unique(rowi % bp)[:12], np.unique(u)[:10]))
                    else:
                        el = np.unique(d // 2); rowi, u = el // 384, el % 384
                        print("     %-26s %d elements differ: cand %s t %s units %s" % (names[k], el.size, np.unique(rowi // 33)[:12], np.unique(rowi % 33)[:6], np.unique(u)[:10]))
        del eng
print("forwards per mode", tot, "with mismatches", bad)
