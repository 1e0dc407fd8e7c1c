"""Two-chain GRU kernel (gru_tc4.cu) against the one-chain kernel (gru_tc3.cu) on the same input projection, bit for bit:
stress loop that reports WHERE they differ (candidate -> CTA pair / chain / CTA rank / row, time step, unit), then the
per-phase clock64() counters of cluster 0."""
import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clairs_to_b200 import _lib
from clairs_to_b200.engine import Engine
from oracle import nn_oracle
lib = _lib.lib()
lib.cto_debug_timing.argtypes = [C.c_void_p]
aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104)
neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(4), 204)
H = 192
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for n in (1000, 4096, 33334):
    eng = Engine(aff_sd, neg_sd, max_batch=n)
    bp = (n + 127) // 128 * 128
    g = torch.Generator(device='cuda'); g.manual_seed(n)
    xp = (torch.rand((6 * H, 33 * bp), device='cuda', generator=g) - 0.5) * 4
    rh, rm = eng.neg_recurrence(xp, n, two_chains=False)
    torch.cuda.synchronize()
    nbad = 0
    for rep in range(reps):
        if rep % 3 == 0:      # cold-ish: thrash L2 / instruction caches with another kernel family
            junk = torch.randn(64 << 20, device='cuda').mul_(1.0001)
        gh, gm = eng.neg_recurrence(xp, n, two_chains=True)
        torch.cuda.synchronize()
        diff = (gh != rh) | (gm != rm)
        if diff.any():
            nbad += 1
            idx = diff.nonzero().cpu().numpy()
            cand, t, u = idx[:, 0], idx[:, 1], idx[:, 2]
            first_t = {}
            print("n", n, "rep", rep, "mismatching elements", idx.shape[0], "candidates", np.unique(cand)[:16], "count", np.unique(cand).size)
            c0 = np.unique(cand)[0]
            sel = cand == c0
            for d in (0, 1):
                ds = sel & ((u >= H) == bool(d))
                if ds.any():
                    tt = t[ds]; first = tt.min() if d == 0 else tt.max()
                    print("   candidate", c0, "dir", d, "first wrong t", first, "units at that t", np.unique(u[ds & (t == first)] % H)[:24],
                          "pair", c0 // 256, "chain", (c0 // 128) & 1, "cta", (c0 // 64) & 1, "m", c0 % 64)
    print("n", n, "reps", reps, "bad reps", nbad, "status", eng.fused_status())
buf = torch.zeros(128, dtype=torch.int64, device='cuda')
lib.cto_debug_timing(C.c_void_p(buf.data_ptr())); lib.cto_debug_set(1)
for two in (0, 1):
    for _ in range(2): eng.neg_recurrence(xp, n, two_chains=two)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); eng.neg_recurrence(xp, n, two_chains=two); e1.record(); torch.cuda.synchronize()
    print("two_chains", two, "%.3f ms for %d candidates" % (e0.elapsed_time(e1), n))
t = buf.cpu().tolist()[32:]
names = {0: "mma.wait_h_ready", 1: "mma.wait_slot_free", 2: "mma.wait_w_full", 3: "mma.issue+commit",
         8: "gate0.x_stage", 9: "gate0.wait_slot_full", 10: "gate0.tmem_ld", 11: "gate0.math+tmem_st", 12: "gate0.stores", 13: "gate0.deferred_h_tiles",
         16: "gate1.x_stage", 17: "gate1.wait_slot_full", 18: "gate1.tmem_ld", 19: "gate1.math+tmem_st", 20: "gate1.stores", 21: "gate1.deferred_h_tiles"}
print("cycles per step (33 steps, both chains), H=192:")
for i, nm in names.items(): print("    %-28s %9.0f" % (nm, t[i] / 33.0))
