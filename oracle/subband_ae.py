"""Functional CPU restatement of the subband auto-encoders ("scaling networks").

TEST INFRASTRUCTURE.  ``SubbandAutoEncoder`` (lifting_dwt_nets.py:82-125):
pointwise grouped 1x1 convs + tanh, hidden 32.  ``SubbandAutoEncoderBerk``
(:126-165): 3x3 convs with GDN / inverse GDN.
"""
import torch
import torch.nn.functional as F

from . import thirdparty as tp


def is_berk(sd, pfx):
    return (pfx + "ae_down.1.beta") in sd


def encode(x, sd, pfx):
    iC = x.shape[1]
    if is_berk(sd, pfx):
        a = x
        for k in (0, 2, 4):
            a = F.conv2d(a, sd[f"{pfx}ae_down.{k}.weight"], sd[f"{pfx}ae_down.{k}.bias"], padding=1)
            a = tp.gdn(a, sd[f"{pfx}ae_down.{k + 1}.beta"], sd[f"{pfx}ae_down.{k + 1}.gamma"], inverse=False)
        return F.conv2d(a, sd[f"{pfx}ae_down.6.weight"], sd[f"{pfx}ae_down.6.bias"], padding=1)
    a = x
    for k in (0, 2, 4):
        a = torch.tanh(F.conv2d(a, sd[f"{pfx}ae_down.{k}.weight"], sd[f"{pfx}ae_down.{k}.bias"], groups=iC))
    return F.conv2d(a, sd[f"{pfx}ae_down.6.weight"], sd[f"{pfx}ae_down.6.bias"], groups=iC)


def decode(y, sd, pfx):
    iC = y.shape[1]
    if is_berk(sd, pfx):
        a = y
        for k in (0, 2, 4):
            a = F.conv_transpose2d(a, sd[f"{pfx}ae_up.{k}.weight"], sd[f"{pfx}ae_up.{k}.bias"], padding=1)
            a = tp.gdn(a, sd[f"{pfx}ae_up.{k + 1}.beta"], sd[f"{pfx}ae_up.{k + 1}.gamma"], inverse=True)
        return F.conv_transpose2d(a, sd[f"{pfx}ae_up.6.weight"], sd[f"{pfx}ae_up.6.bias"], padding=1)
    a = y
    for k in (0, 2, 4):
        a = torch.tanh(F.conv_transpose2d(a, sd[f"{pfx}ae_up.{k}.weight"], sd[f"{pfx}ae_up.{k}.bias"], groups=iC))
    return F.conv_transpose2d(a, sd[f"{pfx}ae_up.6.weight"], sd[f"{pfx}ae_up.6.bias"], groups=iC)
