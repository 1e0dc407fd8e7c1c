"""AFF on a second stream beside NEG (cto_engine_set_overlap, default on) is the same arithmetic in a different schedule: the
outputs of predict are bit-identical with and without it, over several engine chunks and repeated calls."""

import numpy as np
import pytest
import torch

from clairs_to_b200 import synth
from clairs_to_b200.engine import Engine, low_bq_cut_for, stream_to_device
from clairs_to_b200 import synth_weights as sw

pytestmark = pytest.mark.gpu


def test_two_stream_predict_is_bit_identical():
    aff_sd = sw.synth_state_dict(sw.aff_state_dict_shapes(4), 104)
    neg_sd = sw.synth_state_dict(sw.neg_state_dict_shapes(4), 204)
    eng = Engine(aff_sd, neg_sd, max_batch=256)
    n = 1000                                                   # four engine chunks, the last one ragged
    (aff, _), (neg, _) = synth.synth_pair(n, 77, 'ont')
    a, b = stream_to_device(aff, eng.device), stream_to_device(neg, eng.device)
    cut = low_bq_cut_for("ont")
    eng.set_overlap(False)
    want = eng.run_sites(a, b, cut, posterior=False)
    want = {k: want[k].clone() for k in ("logits_aff", "logits_neg", "probs")}
    eng.set_overlap(True)
    for _ in range(5):
        got = eng.run_sites(a, b, cut, posterior=False)
        torch.cuda.synchronize()
        for k, w in want.items():
            assert torch.equal(got[k], w), k
    eng.close()
