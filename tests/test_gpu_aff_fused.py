"""GPU parity tests of the fused CvT transformer kernel (csrc/aff_fused.cu) against the oracle and against the
kernel-per-op tensor-core path, across batch sizes that leave tiles partly empty and across network shapes
(clairs/predict.py:520-553 hyper-parameters, and the CvT class defaults of clairs/model.py:150-184)."""

import numpy as np
import pytest
import torch

from oracle import nn_oracle

pytestmark = pytest.mark.gpu
TOL = 1e-3

CLASS_DEFAULT_CVT = dict(s1=(32, 1, 1), s2=(64, 3, 2), s3=(128, 6, 10))     # clairs/model.py:153-176


def _count_like(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.integers(-50, 51, size=(n, 33, 34)).astype(np.float32)
    x[rng.random(x.shape) < 0.6] = 0.0
    return x


def _engine(aff_sd, n_heads, max_batch):
    from clairs_to_b200.engine import Engine
    neg_sd = nn_oracle.synth_state_dict(nn_oracle.neg_state_dict_shapes(n_heads), 200 + n_heads)
    return Engine(aff_sd, neg_sd, max_batch=max_batch)


@pytest.mark.parametrize("n", [1, 11, 12, 13, 24, 25, 300, 1000])
def test_fused_layers_vs_oracle_and_per_op(n):
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4), 104)
    eng = _engine(aff_sd, 4, 1024)
    try:
        x = _count_like(n, 50 + n)
        want = nn_oracle.aff_forward(x, aff_sd).numpy()
        eng.set_tensor_cores(1)
        fused = eng.forward_aff(torch.from_numpy(x)).cpu().numpy()
        st = eng.fused_status()
        assert st[0] == 0, "fused kernel barrier timeout: %r" % (st.tolist(),)
        eng.set_tensor_cores(2)
        per_op = eng.forward_aff(torch.from_numpy(x)).cpu().numpy()
        e_f, e_p = np.abs(fused - want).max(), np.abs(per_op - want).max()
        print("n=%d: max |logit err| fused %.3g, per-op %.3g (max |logit| %.2f)" % (n, e_f, e_p, np.abs(want).max()))
        assert e_p < TOL
        assert e_f < TOL
    finally:
        eng.close()


@pytest.mark.parametrize("n_heads", [4, 6])
def test_fused_layers_class_default_cvt(n_heads):
    """CvT class-default hyper-parameters (stage 1 width 32, stage 3 with 6 heads and depth 10): the SNV checkpoints are
    pickled modules whose dimensions live in the pickle (clairs/predict.py:513-517), so this shape must run on the
    tensor-core engine too."""
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads, CLASS_DEFAULT_CVT), 300 + n_heads, 0.7)
    eng = _engine(aff_sd, n_heads, 256)
    try:
        x = _count_like(150, 9)
        want = nn_oracle.aff_forward(x, aff_sd).numpy()
        got = eng.forward_aff(torch.from_numpy(x)).cpu().numpy()
        st = eng.fused_status()
        assert st[0] == 0, "fused kernel barrier timeout: %r" % (st.tolist(),)
        err = np.abs(got - want).max()
        print("class-default CvT, %d heads: max |logit err| %.3g (max |logit| %.2f)" % (n_heads, err, np.abs(want).max()))
        assert err < TOL
        eng.set_tensor_cores(0)
        exact = eng.forward_aff(torch.from_numpy(x)).cpu().numpy()
        assert np.abs(exact - want).max() < 5e-5
    finally:
        eng.close()


@pytest.mark.parametrize("n_heads", [4, 6])
def test_class_default_cvt_against_reference_golden(golden_dir, n_heads):
    """The engine on the CvT class-default shape vs logits of the unmodified reference ``CvT()`` / ``CvT_Indel()``
    (tests/golden/nn_golden_default.npz): tensor-core engine within 1e-3, exact fp32 engine within 5e-5."""
    import os
    g = np.load(os.path.join(golden_dir, "nn_golden_default.npz"))
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(n_heads, CLASS_DEFAULT_CVT), 300 + n_heads, 0.7)
    eng = _engine(aff_sd, n_heads, 16)                      # 24 candidates -> two internal chunks
    try:
        x = torch.from_numpy(g["x_%d" % n_heads])
        want = g["aff_logits_%d" % n_heads]
        got = eng.forward_aff(x).cpu().numpy()
        assert eng.fused_status()[0] == 0
        assert np.abs(got - want).max() < TOL
        eng.set_tensor_cores(0)
        assert np.abs(eng.forward_aff(x).cpu().numpy() - want).max() < 5e-5
    finally:
        eng.close()


def test_unsupported_shape_fails_loudly_on_the_tensor_core_engine():
    """No silent CUDA-core fallback: a CvT whose stage-2 width is 16 has no tcgen05 kernel and must raise on the
    tensor-core engine (and run on the exact fp32 engine)."""
    from clairs_to_b200 import _lib
    cfg = dict(s1=(16, 1, 1), s2=(16, 1, 1), s3=(128, 4, 1))
    aff_sd = nn_oracle.synth_state_dict(nn_oracle.aff_state_dict_shapes(4, cfg), 11)
    eng = _engine(aff_sd, 4, 64)
    try:
        x = torch.from_numpy(_count_like(8, 1))
        with pytest.raises(_lib.CtoError):
            eng.forward_aff(x)
        eng.set_tensor_cores(0)
        want = nn_oracle.aff_forward(x.numpy(), aff_sd).numpy()
        assert np.abs(eng.forward_aff(x).cpu().numpy() - want).max() < 5e-5
    finally:
        eng.close()
