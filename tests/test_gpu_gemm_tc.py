"""GPU: the tcgen05 bf16x3 GEMM (csrc/gemm_tc.cu) against an fp64 torch reference and against the
fp32 CUDA-core GEMM it replaces.  Floating point: both paths are held to fp32-rounding level,
2e-5 relative to sum |a||w| (a single bf16 pass would be ~4e-3: the hi/mid split matters)."""

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, K) taken from the two networks plus tails
    (300, 128, 64), (1000, 1152, 256), (256, 16, 64), (130, 64, 16), (515, 128, 12672),
    (2 * 33 * 128, 768, 256), (45, 512, 128), (128, 192, 64), (77, 384, 64), (640, 256, 128),
]


def _act(x, act):
    if act == 1:
        return torch.nn.functional.gelu(x)
    if act == 2:
        return torch.nn.functional.selu(x)
    return x


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_tf32_vs_fp64(m, n, k, act):
    from clairs_to_b200.engine import gemm_nt
    g = torch.Generator(device="cpu").manual_seed(m * 7 + n * 3 + k + act)
    a = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).cuda()
    bias = torch.randn(n, generator=g).cuda()
    res = torch.randn(m, n, generator=g).cuda() if act == 0 else None
    want = _act(a.double() @ w.double().t() + bias.double(), act)
    if res is not None:
        want = want + res.double()
    scale = (a.double().abs() @ w.double().abs().t()).max().item()
    exact = gemm_nt(a, w, bias, res, act, tensor_cores=False)
    assert (exact.double() - want).abs().max().item() < 1e-5 * max(scale, 1.0)
    if n % 64 != 0:
        from clairs_to_b200._lib import CtoError
        with pytest.raises(CtoError):        # narrow outputs stay on the CUDA-core kernel
            gemm_nt(a, w, bias, res, act, tensor_cores=True)
        return
    got = gemm_nt(a, w, bias, res, act, tensor_cores=True)
    err = (got.double() - want).abs().max().item()
    assert err < 2e-5 * max(scale, 1.0), (err, scale)
    assert torch.isfinite(got).all()


def test_gemm_tf32_strided_a_and_inplace_residual():
    from clairs_to_b200.engine import gemm_nt
    g = torch.Generator(device="cpu").manual_seed(1)
    big = torch.randn(500, 96, generator=g).cuda()
    a = big[:, :64]                                   # lda = 96
    w = (torch.randn(64, 64, generator=g) / 8).cuda()
    res = torch.randn(500, 64, generator=g).cuda()
    want = a.double() @ w.double().t() + res.double()
    got = gemm_nt(a, w, None, res, 0, tensor_cores=True)
    assert (got.double() - want).abs().max().item() < 1e-4


MODES = [3, 5, 7, 9, 11]     # cto_gemm_nt mode masks: 1 | 2 pre-split A | 4 split C | 8 bias per row + n-major tiles


@pytest.mark.parametrize("m,n,k", [(300, 128, 64), (1000, 1152, 256), (515, 128, 12672), (768, 33 * 128, 40), (45, 512, 128),
                                   (1152, 640, 256), (130, 64, 16)])
@pytest.mark.parametrize("mode", MODES)
def test_gemm_engine_modes(m, n, k, mode):
    """The operand / result layouts the engine uses between its own kernels: A handed over as bf16 hi/mid planes
    (no converter pass), C emitted as planes, bias per output row (transposed GRU input projections)."""
    from clairs_to_b200.engine import gemm_nt
    g = torch.Generator(device="cpu").manual_seed(m + 3 * n + 7 * k + mode)
    a = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).cuda()
    by_row = bool(mode & 8)
    bias = torch.randn(m if by_row else n, generator=g).cuda()
    act = 1 if mode & 4 else 0
    want = a.double() @ w.double().t() + (bias.double()[:, None] if by_row else bias.double())
    want = _act(want, act)
    scale = (a.double().abs() @ w.double().abs().t()).max().item()
    got = gemm_nt(a, w, bias, None, act, tensor_cores=mode)
    err = (got.double() - want).abs().max().item()
    assert err < 2e-5 * max(scale, 1.0), (err, scale)
    assert torch.isfinite(got).all()


@pytest.mark.parametrize("m,n,k", [(1152, 33 * 1024, 256), (1152, 33 * 1152, 256), (768, 33 * 1024 + 128, 64), (300, 256 * 400, 192)])
def test_gemm_wide_tiles_are_bit_identical(m, n, k):
    """GEMM_WIDE_N (128 x 256 tiles, two 96 KB stages: the transposed GRU input projections) accumulates every output
    element over the same k-steps in the same order as the 128 x 128 walk: bit-identical, incl. a ragged last column tile
    (n = 128 mod 256) and ragged row blocks."""
    from clairs_to_b200.engine import gemm_nt
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).cuda()
    bias = torch.randn(m, generator=g).cuda()
    plain = gemm_nt(a, w, bias, None, 0, tensor_cores=11)
    for _ in range(2):
        wide = gemm_nt(a, w, bias, None, 0, tensor_cores=27)
        assert torch.equal(plain, wide)
    want = a.double() @ w.double().t() + bias.double()[:, None]
    scale = (a.double().abs() @ w.double().abs().t()).max().item()
    assert (wide.double() - want).abs().max().item() < 2e-5 * max(scale, 1.0)


@pytest.mark.parametrize("m,n,k", [(1152, 33 * 1024, 256), (1152, 33 * 1152, 256), (768, 33 * 1024 + 128, 64), (300, 128 * 400, 192),
                                   (1152, 33 * 1024 + 64, 256), (100, 128 * 600, 256)])
def test_gemm_pair_is_bit_identical(m, n, k):
    """csrc/gemm_pair.cu (CTA pairs, cta_group::2 M = 256, A resident, half a column tile per CTA): the transposed GRU
    input projection.  Every output element is accumulated over the same k-steps in the same order as in gemm_tc.cu, so
    the result is bit-identical -- incl. an odd number of row blocks (the last pair's second CTA lies beyond M), ragged
    row blocks and a ragged last column tile."""
    from clairs_to_b200.engine import gemm_nt
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    a = torch.randn(m, k, generator=g).cuda()
    w = (torch.randn(n, k, generator=g) / np.sqrt(k)).cuda()
    bias = torch.randn(m, generator=g).cuda()
    plain = gemm_nt(a, w, bias, None, 0, tensor_cores=11)
    for _ in range(3):
        pair = gemm_nt(a, w, bias, None, 0, tensor_cores=11 | 32)
        assert torch.equal(plain, pair)
    want = a.double() @ w.double().t() + bias.double()[:, None]
    scale = (a.double().abs() @ w.double().abs().t()).max().item()
    assert (pair.double() - want).abs().max().item() < 2e-5 * max(scale, 1.0)
