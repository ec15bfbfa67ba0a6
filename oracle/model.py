"""Whole-model functional restatement (rows a13, a14) + the synthetic-weights recipes.

TEST INFRASTRUCTURE.  ``LiftingBasedDWTNet.forward`` / ``Wrapper.forward``
(graphs/models/LiftingBasedDWT_net.py:154-170, 48-62) and ``TrainRDLoss.forward3``
(graphs/losses/rate_dist.py:35-42).
"""
import re
import zlib

import torch

from . import entropy, thirdparty as tp, transform


class Cfg(dict):
    """Attribute-style config carrying the keys of ``liftingDWT.json`` the path reads."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def default_cfg(**kw):
    c = Cfg(clrch=1, netType="LiftingBasedNeuralWaveletv4", entropy_layer="conditioned2ZTsepSubbands",
            autoencoder="SubbandAutoEncoderBerk", dwtlevels=4, num_lifting_perlayer=2, filtersize=5,
            block_property="same", scale=0, linearity_flag=1, depth_scale=2, res_connection_weight=0.1,
            mode="test", imshow_validation=False, postprocess="none", lambda_=11700)
    c.update(kw)
    return c


def net_forward(x, sd, pfx, cfg, training=False):
    """One colour plane: encode -> entropy model -> decode.  Returns
    (xhat, si_xe, si_xo_list, xe_qnt, xo_list_qnt, out_xe, out_xo_list)."""
    out_xe, out_xo = transform.encode(x, sd, pfx + "autoencoder.", cfg)
    fwd = entropy.FORWARD.get(cfg.entropy_layer)
    if fwd is None:
        raise ValueError(cfg.entropy_layer)
    si_xe, si_xo, xe_q, xo_q = fwd(out_xe, out_xo, sd, pfx + "entropymodel.", cfg.dwtlevels, training)
    xhat = transform.decode(xe_q, xo_q, sd, pfx + "autoencoder.", cfg)
    return xhat, si_xe, si_xo, xe_q, xo_q, out_xe, out_xo


def wrapper_forward(x, sd, cfg, training=False, full=False):
    """``LiftingBasedDWTNetWrapper.forward``: ``clrch == 1`` = three independent planes, outputs
    concatenated plane after plane; ``clrch == 3`` = one net on the whole tensor."""
    if cfg.clrch == 3:        # one net over the three colour channels together (:40-42, 50-51)
        o = net_forward(x, sd, "model.", cfg, training)
        return (o[0], o[1], list(o[2]), [o]) if full else (o[0], o[1], list(o[2]))
    outs = [net_forward(x[:, c:c + 1], sd, f"model{c}.", cfg, training) for c in range(3)]
    xhat = torch.cat([o[0] for o in outs], dim=1)
    si_xe = torch.cat([o[1] for o in outs], dim=1)
    si_xo = []
    for o in outs:
        si_xo.extend(o[2])
    if full:
        return xhat, si_xe, si_xo, outs
    return xhat, si_xe, si_xo


def rd_loss(x, xhat, si_xe, si_xo, lambda_):
    """``TrainRDLoss.forward3``: bpp = total bits / (B*H*W) (the *3 undoes numel's C)."""
    mse = torch.mean((x - xhat) ** 2)
    rate1 = torch.sum(si_xe) / torch.numel(x) * 3
    rate2 = 0
    for s in si_xo:
        rate2 = rate2 + torch.sum(s) / torch.numel(x) * 3
    return rate1 + rate2 + lambda_ * mse, mse, rate1, rate2


def preprocess(rgb):
    """Agent-side pre-processing (agents/liftingDWT_agent.py:170-171): RGB->YCbCr, Y - 0.5."""
    y = tp.rgb2ycbcr(rgb).clone()
    y[:, 0:1] = y[:, 0:1] - 0.5
    return y


# ----------------------------------------------------------------------------
# synthetic weights
# ----------------------------------------------------------------------------

_LAST_DOWN = re.compile(r"autoencoder\.(Yl_ae|Yh_ae\.\d+)\.ae_down\.6\.(weight|bias)$")


def amplify_v1(sd, gain=200.0):
    """Synthetic-weights v1: seeded default init, then the last
    ``ae_down`` conv of every subband auto-encoder x200 so symbols are not all 0."""
    out = {}
    for k, v in sd.items():
        out[k] = v * gain if _LAST_DOWN.search(k) else v
    return out


_SIGMA_BIAS = re.compile(r"(csc_xe\.8|csc_list\.\d+\.8|cgp_out_xo_list\.\d+\.6|plc_list\.\d+\.4)\.bias$")
_SIGMA_BIAS_ZT = re.compile(r"dep_\d_list_sigma\.\d+\.8\.bias$")


def keyed_weights(ref_sd, gain=None):
    """Synthetic-weights v2: every floating parameter is drawn from a generator
    seeded by the CRC32 of its *canonical* key, scaled to roughly its default-init
    range, so the same tensors can be rebuilt anywhere from a module's own
    ``state_dict()`` and loaded with ``load_state_dict(strict=True)``.

    ``ref_sd`` must be a live ``state_dict()``: parameters re-registered under several
    names (the shared lifting blocks appear under ``P_blocks``/``U_blocks`` and again
    under every ``waveletForward.N`` / ``waveletInverse.N``) share storage, and the
    canonical key of a tensor is the first key with that storage.  Buffers (masks,
    bounds, tables, filters), GDN and EntropyBottleneck shape parameters are kept as
    constructed.  The 3-tap pre-filters keep their CDF 9/7 values plus a +-0.05
    perturbation (all taps non-zero), ``nh``/``nl`` get +-0.5.  The last ``ae_down``
    conv is scaled by ``gain`` (default 8 for the Berk auto-encoder, 40 for the
    pointwise one: symbol std about 3-10) and the sigma outputs of the context
    networks get a +4 bias so likelihoods are not pinned at the 1e-9 bound (a
    saturated rate would make the bpp check blind to context-CNN errors)."""
    if gain is None:
        gain = 8.0 if any(k.endswith("ae_down.1.beta") for k in ref_sd) else 40.0
    canon_of = {}
    out = {}
    for k, v in ref_sd.items():
        if not torch.is_floating_point(v) or v.numel() == 0:
            out[k] = v.clone()
            continue
        if any(s in k for s in (".mask", ".bound", "pedestal", "target", "scale_bound", "scale_table",
                                "xfm.", "ifm.", "quantiles", ".beta", ".gamma", "_matrix", "_factor")):
            out[k] = v.clone()
            continue
        canon = canon_of.setdefault((v.data_ptr(), tuple(v.shape)), k)
        if canon != k:
            out[k] = out[canon].clone()
            continue
        g = torch.Generator().manual_seed(zlib.crc32(k.encode()))
        u = torch.rand(v.shape, generator=g, dtype=torch.float32) * 2 - 1
        if "preProcessingList" in k or ".convBlock." in k:
            t = v.detach().float() + 0.05 * u
        elif k.endswith(".nh") or k.endswith(".nl"):
            t = 0.5 * u
        else:
            fan_in = v[0].numel() if v.dim() > 1 else max(v.numel(), 1)
            bound = 0.05 if (k.endswith(".bias") or "_bias" in k) else (3.0 / fan_in) ** 0.5
            t = u * bound
            if _LAST_DOWN.search(k):
                t = t * gain
            if _SIGMA_BIAS.search(k):
                t[0::2] += 4.0
            if _SIGMA_BIAS_ZT.search(k):
                t += 4.0
        out[k] = t.to(v.dtype)
    return out
