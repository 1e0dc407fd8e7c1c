"""GPU tests of the GRU recurrence kernels (csrc/gru_tc3.cu one chain per CTA pair, csrc/gru_tc4.cu two chains per CTA
pair; torch.nn.GRU as used by clairs/model.py:440-470):

* the two kernels are the same arithmetic in a different schedule, so their bf16 output planes must agree BIT FOR BIT on
  the same input projection, for batch sizes that leave chains / CTAs / rows empty;
* the recurrence against a plain fp32 numpy restatement of the GRU cell (tolerance: the bf16 hi + mid split of the
  output, 2^-16 relative, plus the bf16x3 products);
* run-to-run determinism of the whole NEG forward right after the projection GEMM has written its output.  This is the
  regression test of a race found in round 2: the gate warps handed a projection stage back to the TMA producer before
  their shared-memory loads had returned, and a refill that hit in L2 (the projection had just been written) overwrote
  the stage under them -- a few candidates in groups of 4..32 came out slightly wrong in ~1 % (one chain) to ~25 % (two
  chains) of the forwards, never in an isolated kernel test with a cold projection.
"""

import numpy as np
import pytest
import torch

from oracle import nn_oracle

pytestmark = pytest.mark.gpu
H2 = 192


def _engine(max_batch, n_heads=4):
    from clairs_to_b200.engine import Engine
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads), 104)
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 204)
    return Engine(aff_sd, neg_sd, max_batch=max_batch), neg_sd


def _planes_to_f32(hi, mid):
    f = lambda p: (p.to(torch.int32) << 16).view(torch.float32)
    return f(hi) + f(mid)


@pytest.mark.parametrize("n", [1, 63, 64, 129, 256, 257, 1000])
def test_two_chain_kernel_is_bit_identical_to_one_chain(n):
    eng, _ = _engine(1024)
    try:
        bp = (n + 127) // 128 * 128
        g = torch.Generator(device="cuda"); g.manual_seed(n)
        xp = (torch.rand((6 * H2, 33 * bp), device="cuda", generator=g) - 0.5) * 6
        a_hi, a_mid = eng.neg_recurrence(xp, n, two_chains=False)
        for _ in range(3):
            b_hi, b_mid = eng.neg_recurrence(xp, n, two_chains=True)
            assert eng.fused_status()[0] == 0
            assert torch.equal(a_hi, b_hi) and torch.equal(a_mid, b_mid)
    finally:
        eng.close()


def test_recurrence_against_fp32_gru_cell():
    n = 200
    eng, neg_sd = _engine(256)
    try:
        bp = 256
        rng = np.random.default_rng(5)
        xp = ((rng.random((6 * H2, 33 * bp), dtype=np.float32) - 0.5) * 4).astype(np.float32)
        hi, mid = eng.neg_recurrence(torch.from_numpy(xp).cuda(), n, two_chains=True)
        got = _planes_to_f32(hi, mid).cpu().numpy()                       # [n, 33, 2H]
        # fp32 restatement: xproj already holds W_ih x + b_ih (+ nothing else): h' = (1 - z) n + z h with
        # r = s(x_r + W_hr h + b_hr), z = s(x_z + W_hz h + b_hz), n = tanh(x_n + r (W_hn h + b_hn))   (torch.nn.GRU)
        sig = lambda v: 1.0 / (1.0 + np.exp(-v))
        for d, sfx in ((0, ""), (1, "_reverse")):
            w = neg_sd["lstm_2.weight_hh_l0" + sfx].numpy().astype(np.float64)
            b = neg_sd["lstm_2.bias_hh_l0" + sfx].numpy().astype(np.float64)
            h = np.zeros((n, H2))
            for step in range(33):
                t = 32 - step if d else step
                xs = xp[d * 3 * H2:(d + 1) * 3 * H2, t * bp:t * bp + n].T.astype(np.float64)      # [n, 3H]
                # the engine folds b_hr and b_hz into the projection's bias (they are part of the hook's xproj); the kernel
                # itself adds only b_hn
                hh = h @ w.T
                r = sig(xs[:, :H2] + hh[:, :H2])
                z = sig(xs[:, H2:2 * H2] + hh[:, H2:2 * H2])
                nn_ = np.tanh(xs[:, 2 * H2:] + r * (hh[:, 2 * H2:] + b[2 * H2:]))
                h = (1 - z) * nn_ + z * h
                err = np.abs(got[:, t, d * H2:(d + 1) * H2] - h).max()
                assert err < 2e-4, (d, step, err)
    finally:
        eng.close()


@pytest.mark.parametrize("n", [700, 1000, 4096])
def test_neg_forward_is_deterministic_behind_the_projection_gemm(n):
    rng = np.random.default_rng(n)
    x = torch.from_numpy(rng.integers(-50, 51, size=(n, 33, 34)).astype(np.float32)).cuda()
    for rep in range(4):                                                  # fresh engines: fresh workspace, cold caches
        eng, _ = _engine(n)
        try:
            eng.set_tensor_cores(2)
            ref = eng.forward_neg(x).clone()
            for mode in (1, 2, 1, 1, 2, 1):
                eng.set_tensor_cores(mode)
                got = eng.forward_neg(x)
                assert torch.equal(got, ref), "mode %d, engine %d: %d candidates differ" % (
                    mode, rep, int((got != ref).reshape(n, -1).any(1).sum()))
            assert eng.fused_status()[0] == 0
        finally:
            eng.close()
