"""``autoencoder.encode/decode`` of both transform types (rows a5, a6 of SURVEY.md 8a).

TEST INFRASTRUCTURE.  ``LiftingBasedNeuralWaveletv4.encode/decode``
(lifting_dwt_nets.py:724-782) and ``DWTPytorchWaveletsLayer.encode/decode``
(:241-277).
"""
from . import lifting, subband_ae, thirdparty as tp


def encode(x, sd, pfx, cfg):
    if cfg.netType == "CDF97":
        yl, yh5 = tp.dwt97_forward(x, cfg.dwtlevels)
        yh = [h.reshape(h.shape[0], h.shape[1] * 3, h.shape[3], h.shape[4]) for h in yh5]
    elif cfg.netType == "LiftingBasedNeuralWaveletv4":
        yl, yh = lifting.transform_forward(x, sd, pfx, cfg)
    else:
        raise ValueError(cfg.netType)
    out_xe = subband_ae.encode(yl, sd, pfx + "Yl_ae.")
    out_xo = [subband_ae.encode(yh[i], sd, f"{pfx}Yh_ae.{i}.") for i in range(cfg.dwtlevels)]
    return out_xe, out_xo


def decode(out_xe, out_xo, sd, pfx, cfg):
    yl = subband_ae.decode(out_xe, sd, pfx + "Yl_ae.")
    yh = [subband_ae.decode(out_xo[i], sd, f"{pfx}Yh_ae.{i}.") for i in range(cfg.dwtlevels)]
    if cfg.netType == "CDF97":
        yh5 = [h.reshape(h.shape[0], h.shape[1] // 3, 3, h.shape[2], h.shape[3]) for h in yh]
        return tp.dwt97_inverse(yl, yh5)
    return lifting.transform_inverse(yl, yh, sd, pfx, cfg)
