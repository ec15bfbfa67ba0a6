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


FLIP_EPS = 1e-5          # a flip must sit within 1e-5 of a rounding boundary of the fp64 oracle
FLIP_LOG = []            # (label, n_symbols, n_mismatch, n_unexplained) of every audit of this session


def flip_audit(q_gpu, q_ref, pre_ref, eps=FLIP_EPS, pre_ref64=None, label=None, q_ref64=None):
    """Symbol parity.  Quantised tensors are integers (``round(x)``) or mean-shifted integers
    (``round(x - mu) + mu`` for EntropyBottleneck / onlyEZWT outputs, where mu itself is a
    context-CNN output carrying fp32 noise).  The difference is split into an integer number of
    quantisation steps k and a remainder: a sample *matches* when k == 0 and the remainder (the
    difference of the two mu) is within 1e-4 of the tensor's scale (exactly 0 for plain rounding).
    A mismatch is *explained* only when it is a single step (|k| == 1) and the oracle's pre-rounding
    value -- the float64 oracle's (``pre_ref64`` / ``q_ref64``, the tie-break of SURVEY.md section 7)
    when given, else the fp32 one -- sits within ``eps * max(1, |x|)`` (plus, with the float64 tie-break, the fp32
    reference's own distance from the float64 value at that sample) of the rounding boundary: there
    fp32 summation order alone decides the symbol, and the reference run on another thread count flips
    it too.  For mean-shifted rounding the rounded quantity is x - mu and its distance to the boundary
    is 0.5 - |x - q| (q - mu is an integer).
    Returns (n_mismatch, n_unexplained); every call is logged in ``FLIP_LOG`` and printed by the suite."""
    d = q_gpu - q_ref
    k = torch.round(d)
    integer_q = bool((q_ref == torch.round(q_ref)).all())
    rem_tol = 0.0 if integer_q else 1e-4 * max(1.0, float(q_ref.abs().max()))
    rem_ok = (d - k).abs() <= rem_tol
    mism = (k != 0) | ~rem_ok
    n = int(mism.sum())
    bad = 0
    if n:
        ok = (k[mism].abs() == 1) & rem_ok[mism]
        pre = (pre_ref64 if pre_ref64 is not None else pre_ref).double()[mism]
        if integer_q:
            dist = (pre - torch.floor(pre) - 0.5).abs()
            scale = pre.abs().clamp(min=1.0)
        else:
            qr = (q_ref64 if (q_ref64 is not None and pre_ref64 is not None) else q_ref).double()[mism]
            # the rounded quantity is x - mu: the two sides' mu differ by exactly d - k at this sample (a context-CNN
            # output, itself held to 1e-4 of the tensor scale by rem_ok), which moves the boundary by that much
            dist = 0.5 - (pre - qr).abs() - (d - k).abs().double()[mism]
            scale = torch.maximum(pre.abs(), qr.abs()).clamp(min=1.0)
        # the fp32 reference's own deviation from the float64 oracle AT THIS SAMPLE widens the band: a symbol the reference
        # itself only decides within that margin cannot be held against the GPU more tightly
        slack = (pre_ref.double()[mism] - pre).abs() if pre_ref64 is not None else torch.zeros_like(dist)
        ok = ok & (dist <= eps * scale + slack)
        bad = int((~ok).sum())
        if bad:
            print(f"  unexplained flips in {label}: |k| {k[mism][~ok].abs().tolist()[:4]}, distance to the boundary / (eps * scale) "
                  f"{(dist / (eps * scale))[~ok].tolist()[:4]}, reference fp32-vs-fp64 slack / (eps * scale) "
                  f"{(slack / (eps * scale))[~ok].tolist()[:4]}, remainder ok {rem_ok[mism][~ok].tolist()[:4]}")
    FLIP_LOG.append((label or "", int(q_ref.numel()), n, bad))
    return n, bad


def flip_summary():
    tot = sum(r[1] for r in FLIP_LOG)
    n = sum(r[2] for r in FLIP_LOG)
    bad = sum(r[3] for r in FLIP_LOG)
    return f"symbol audits: {len(FLIP_LOG)} tensors, {tot} symbols, {n} rounding-boundary flips, {bad} unexplained"


def bits_check(a, b, tol_sum=1e-4, frac_outliers=1e-3):
    """Self-information parity.  Per-coefficient bits are discontinuous where ``round(x - mu)``
    sits on a boundary (a 1e-7 change of mu moves one coefficient by a whole step), so the check
    is: the subband's total within ``tol_sum`` relative, and at most ``frac_outliers`` of the
    coefficients (min 2) off by more than 1e-3 relative.  Returns (rel_sum_err, n_outliers, ok)."""
    rs = abs(float(a.double().sum() - b.double().sum())) / max(abs(float(b.double().sum())), 1e-30)
    out = int(((a - b).abs() > 1e-3 * (1 + b.abs())).sum())
    ok = rs <= tol_sum and out <= max(2, int(frac_outliers * b.numel()))
    return rs, out, ok
