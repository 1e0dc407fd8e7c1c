"""Oracle for predict-side glue and the posterior combine (SURVEY.md §8 rows a3.*, a4.*).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  numpy / Python floats, i.e. the
same fp64 arithmetic the reference performs.

References: clairs/predict.py (P:line), clairs/call_variants.py (CV:line).
"""

from __future__ import annotations

import sys
from math import e, log

import numpy as np

MIN_RESCALE_COV = 50       # shared/param.py:26
CENTER = 16                # hard-coded at P:626-627


def rescale_tensor(int_tensor, depth):
    """P:179-197 + P:207: python-double multiply by 50/depth when depth > 50, then float32."""
    t = np.asarray(int_tensor, dtype=np.float64)
    if float(depth) > MIN_RESCALE_COV:
        t = t * (float(MIN_RESCALE_COV) / float(depth))
    return t.astype(np.float32)


def depth_from_alt_info(alt_info: str) -> float:
    return float(alt_info.split('-')[0])           # P:182


def strand_counts(x_unscaled):
    """P:626-642: forward = centre row cols 0:4, reverse = cols 9:13, the negative
    (reference) entry replaced by -(row sum)."""
    x = np.asarray(x_unscaled, dtype=np.float32)
    out = []
    for cols in (slice(0, 4), slice(9, 13)):
        c = x[:, CENTER, cols].copy()
        fixed = c.copy()
        rows, idx = np.where(c < 0)
        fixed[rows, idx] = c[rows].sum(axis=1) * -1
        out.append(np.where(fixed == -0, 0, fixed))
    return out[0], out[1]


def format_predict_row(chrom, pos, ref_base, alt_info, fwd, rev, probs):
    """P:114-152.  probs: iterable of [p0, p1] float32 pairs in file order
    (a c g t [i d] na nc ng nt [ni nd]); fwd/rev: python lists of floats (list repr)."""
    fields = [chrom, str(pos), ref_base, alt_info, str(fwd), str(rev)]
    fields += [' '.join("{:0.8f}".format(v) for v in p) for p in probs]
    if len(fields) == 14:
        fields.append("")      # the SNV format string has a 15th slot for the empty extra string (P:139-152);
                               # the indel one has exactly 18 slots and silently drops it (P:116-137)
    return "\t".join(fields) + "\n"


# ------------------------------------------------------------------------------------------
# likelihood tables + Bayes combine
# ------------------------------------------------------------------------------------------

def load_likelihood(path_or_array, n_heads):
    """CV:655-796.  Returns (matrices [H,10,10], aff_edges [H,11], neg_edges [H,11])."""
    data = np.loadtxt(path_or_array) if isinstance(path_or_array, str) else np.asarray(path_or_array, dtype=np.float64)
    mats = np.stack([data[10 * h:10 * (h + 1)] for h in range(n_heads)])
    base = 10 * n_heads
    aff, neg = [], []
    for h in range(n_heads):
        for dst, row in ((aff, base + 2 * h), (neg, base + 2 * h + 1)):
            interior = data[row:row + 1].flatten()[:-1]
            dst.append(np.concatenate([[0.0], interior, [1.0]]))
    return mats, np.stack(aff), np.stack(neg)


def posterior(p_aff, p_neg, mats, aff_edges, neg_edges):
    """CV:154-224 / 226-304 for one candidate.  p_aff/p_neg: [H] probabilities of the
    positive class.  Raises IndexError like the reference when a probability is exactly 1.0
    (SURVEY.md §9.12)."""
    post = []
    for h in range(len(p_aff)):
        p, n = float(p_aff[h]), float(p_neg[h])
        i = int(np.digitize(p, aff_edges[h])) - 1
        j = int(np.digitize(1 - n, neg_edges[h])) - 1
        w = mats[h][i][j] + sys.float_info.epsilon
        num = p * (1 - n) * w
        post.append(num / (num + (1 - p) * n * (1 - w)))
    return np.array(post)


_PHRED = -10 * log(e, 10)


def quality_score(p):
    """CV:81-88."""
    p = float(p)
    return float(round(max(_PHRED * log(((1.0 - p) + 1e-10) / (p + 1e-10)) + 2.0, 0.0), 4))


def decide(post, ref_base, snv_mode):
    """CV:211-224 / 290-304: argmax (first max wins), variant / reference decision."""
    k = int(np.argmax(post))
    if snv_mode:
        is_variant = "ACGT"[k] != ref_base
    else:
        is_variant = k >= 4
    return k, float(max(post)), is_variant
