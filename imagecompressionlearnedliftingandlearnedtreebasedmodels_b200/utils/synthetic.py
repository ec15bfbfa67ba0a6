"""Synthetic weights and inputs for measurement: the authors' checkpoints are not available and
with plain default initialisation every quantised symbol is 0, so a benchmark on it would measure a degenerate
rate (bpp 0.01).  Two recipes, both deterministic and rebuilt from a module's own ``state_dict()`` so no weight
file travels:

* ``amplify_v1``: seeded default init, last ``ae_down`` conv of every subband auto-encoder x ``gain``.
* ``keyed_weights`` (v2): every floating parameter drawn from a generator seeded by the CRC32 of its canonical
  ``state_dict`` key; aliases of the shared lifting blocks resolved by storage; 3-tap pre-filters perturbed; the
  sigma heads of the context networks biased +4 so likelihoods are not pinned at the 1e-9 bound.

``tests/test_abi_and_host.py`` checks that these produce exactly the tensors of the oracle's own copy
(``oracle/model.py``), which is what the golden fixtures were generated with.
"""
import re
import zlib

import torch

_LAST_DOWN = re.compile(r"autoencoder\.(Yl_ae|Yh_ae\.\d+)\.ae_down\.6\.(weight|bias)$")
_SIGMA_HEAD = re.compile(r"(csc_xe\.8|csc_list\.\d+\.8|cgp_out_xo_list\.\d+\.6|plc_list\.\d+\.4)\.bias$")
_SIGMA_HEAD_ZT = re.compile(r"dep_\d_list_sigma\.\d+\.8\.bias$")
_KEEP = (".mask", ".bound", "pedestal", "target", "scale_bound", "scale_table", "xfm.", "ifm.", "quantiles",
         ".beta", ".gamma", "_matrix", "_factor")


def amplify_v1(sd, gain=200.0):
    return {k: (v * gain if _LAST_DOWN.search(k) else v) for k, v in sd.items()}


def keyed_weights(live_sd, gain=None):
    """``live_sd``: a live ``module.state_dict()`` (shared parameters must share storage).  Returns a new state
    dict loadable with ``strict=True``."""
    if gain is None:
        gain = 8.0 if any(k.endswith("ae_down.1.beta") for k in live_sd) else 40.0
    first_key_of = {}
    out = {}
    for key, ten in live_sd.items():
        if not torch.is_floating_point(ten) or ten.numel() == 0 or any(s in key for s in _KEEP):
            out[key] = ten.clone()
            continue
        canon = first_key_of.setdefault((ten.data_ptr(), tuple(ten.shape)), key)
        if canon != key:
            out[key] = out[canon].clone()
            continue
        gen = torch.Generator().manual_seed(zlib.crc32(key.encode()))
        u = torch.rand(ten.shape, generator=gen, dtype=torch.float32) * 2 - 1
        if "preProcessingList" in key or ".convBlock." in key:
            val = ten.detach().float() + 0.05 * u
        elif key.endswith(".nh") or key.endswith(".nl"):
            val = 0.5 * u
        else:
            fan_in = ten[0].numel() if ten.dim() > 1 else max(ten.numel(), 1)
            is_bias = key.endswith(".bias") or "_bias" in key
            val = u * (0.05 if is_bias else (3.0 / fan_in) ** 0.5)
            if _LAST_DOWN.search(key):
                val = val * gain
            if _SIGMA_HEAD.search(key):
                val[0::2] += 4.0
            if _SIGMA_HEAD_ZT.search(key):
                val += 4.0
        out[key] = val.to(ten.dtype)
    return out


def load_keyed_weights(model, gain=None):
    """Draw v2 weights for ``model`` (on any device) and load them in place.  Returns the model."""
    dev = next(model.parameters()).device
    sd = keyed_weights({k: v for k, v in model.state_dict().items()}, gain) if dev.type == "cpu" else None
    if sd is None:
        cpu_sd = model.to("cpu").state_dict()
        sd = keyed_weights(cpu_sd, gain)
        model.load_state_dict(sd, strict=True)
        return model.to(dev)
    model.load_state_dict(sd, strict=True)
    return model


def synthetic_rgb(batch, height, width, seed):
    """Uniform [0,1) RGB batch (the range ``ToTensor`` produces), seeded."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.rand(batch, 3, height, width, generator=g)
