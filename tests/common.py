"""Shared helpers of the test-suite."""
import json
import os

import numpy as np
import torch

from oracle import model as om

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
PKG = "imagecompressionlearnedliftingandlearnedtreebasedmodels_b200"


def meta():
    with open(os.path.join(GOLDEN, "meta.json")) as f:
        return json.load(f)


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def product_model(overrides):
    """The product's LiftingBasedDWTNetWrapper for a golden config (constructed on the CPU)."""
    from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200.graphs.models.LiftingBasedDWT_net import \
        LiftingBasedDWTNetWrapper
    cfg = om.default_cfg(**overrides)
    torch.manual_seed(1337)
    return LiftingBasedDWTNetWrapper(cfg), cfg


def keyed_state(model):
    """Synthetic-weights v2 loaded into ``model``; returns the effective state_dict (CPU)."""
    sd = om.keyed_weights(model.state_dict())
    model.load_state_dict(sd, strict=True)
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def rel_err(a, b):
    """||a-b||_inf / ||b||_inf, the tolerance form north_star states."""
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def flip_audit(q_gpu, q_ref, pre_ref, eps=1e-4):
    """Symbol parity.  Quantised tensors are integers (``round(x)``) or mean-shifted integers
    (``round(x - mu) + mu`` for EntropyBottleneck / onlyEZWT outputs, where mu itself is a
    context-CNN output carrying fp32 noise).  The difference is split into an integer number of
    quantisation steps k and a remainder: a sample *matches* when k == 0 and the remainder is
    within 1e-4 relative (exactly 0 for plain rounding); a mismatch must be a single step
    (|k| == 1) and, for plain rounding, the oracle's pre-quantiser value must sit within ``eps``
    of the rounding boundary (frac = 0.5).  Returns (n_mismatch, n_unexplained)."""
    d = q_gpu - q_ref
    k = torch.round(d)
    integer_q = bool((q_ref == torch.round(q_ref)).all())
    rem_tol = 0.0 if integer_q else 1e-4
    rem_ok = (d - k).abs() <= rem_tol * (1 + q_ref.abs())
    mism = (k != 0) | ~rem_ok
    n = int(mism.sum())
    if n == 0:
        return 0, 0
    ok = (k[mism].abs() == 1) & rem_ok[mism]
    if integer_q:
        frac = (pre_ref[mism] - torch.floor(pre_ref[mism]) - 0.5).abs()
        ok = ok & (frac <= eps)
    return n, int((~ok).sum())


def bits_check(a, b, tol_sum=1e-4, frac_outliers=1e-3):
    """Self-information parity.  Per-coefficient bits are discontinuous where ``round(x - mu)``
    sits on a boundary (a 1e-7 change of mu moves one coefficient by a whole step), so the check
    is: the subband's total within ``tol_sum`` relative, and at most ``frac_outliers`` of the
    coefficients (min 2) off by more than 1e-3 relative.  Returns (rel_sum_err, n_outliers, ok)."""
    rs = abs(float(a.double().sum() - b.double().sum())) / max(abs(float(b.double().sum())), 1e-30)
    out = int(((a - b).abs() > 1e-3 * (1 + b.abs())).sum())
    ok = rs <= tol_sum and out <= max(2, int(frac_outliers * b.numel()))
    return rs, out, ok
