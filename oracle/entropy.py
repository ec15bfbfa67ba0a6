"""Functional CPU restatement of the four entropy layers' ``forward`` (rows a7-a12).

TEST INFRASTRUCTURE.  Each function takes ``out_xe`` (B,1,h,w), ``out_xo`` (list of
(B,3,h_l,w_l), finest first), a reference-layout ``state_dict`` and the prefix of
the entropy layer (``"model0.entropymodel."``) and returns
``(si_xe, si_xo_list, xe_qnt, xo_list_qnt)`` exactly like the reference modules
(graphs/models/LiftingBasedDWT_net.py).  ``training`` selects noise (two
independent ``uniform_`` draws per subband, in the reference's order) vs rounding.
"""
import torch
import torch.nn.functional as F

from . import thirdparty as tp

LEAK = 0.01  # nn.LeakyReLU() default slope


def conv_mask(kh, kw, mask_type, like):
    """``MaskedConv2d`` mask geometry (graphs/layers/masked_conv2d.py:9-17)."""
    m = torch.ones(kh, kw, dtype=like.dtype, device=like.device)
    b = 1 if mask_type == "B" else 0
    if kw > 1:
        m[kh // 2, kw // 2 + b:] = 0
    elif kw == 1 and mask_type == "A":
        m[kh // 2, kw // 2 + b:] = 0
    if kh > 1:
        m[kh // 2 + 1:] = 0
    return m


def masked_conv(x, sd, pfx, mask_type, groups):
    """``MaskedConv2d.forward``: weight * mask, then conv (masked_conv2d.py:19-21).
    The mask is rebuilt from its geometry, not read from the buffer, so a
    corrupted buffer in a checkpoint would show up as a golden mismatch."""
    w = sd[pfx + "weight"]
    k = w.shape[-1]
    w = w * conv_mask(w.shape[-2], k, mask_type, w)
    return F.conv2d(x, w, sd[pfx + "bias"], padding=k // 2, groups=groups)


def causal_chain(x, sd, pfx, groups):
    """5-layer masked 3x3 chain A,B,B,B,B with LeakyReLU between
    (``csc_xe`` / ``csc_list[L-1]``, LiftingBasedDWT_net.py:298-317)."""
    a = x
    for j, k in enumerate((0, 2, 4, 6, 8)):
        a = masked_conv(a, sd, f"{pfx}{k}.", "A" if j == 0 else "B", groups)
        if j < 4:
            a = F.leaky_relu(a, LEAK)
    return a


def _gauss_bits(x, sigma, mu, training):
    _, lik = tp.gaussian_conditional_forward(x, sigma, mu, training)
    return -torch.log2(lik)


def upsample2(q):
    return q.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)


def cond2zt_forward(out_xe, out_xo, sd, pfx, L, training=False):
    """``DWTConditioned2EntropyLayerZTsepSubbands.forward`` (:322-372)."""
    mode = "noise" if training else "dequantize"
    xe_q = tp.quantize(out_xe, mode)
    ms = causal_chain(xe_q, sd, pfx + "csc_xe.", xe_q.shape[1])      # groups = ses[L-1] (:308-316)
    si_xe = _gauss_bits(out_xe, ms[:, 0::2], ms[:, 1::2], training)
    qs, sis = [], []
    i = L - 1
    q = tp.quantize(out_xo[i], mode)
    ms = causal_chain(q, sd, f"{pfx}csc_list.{i}.", q.shape[1])      # groups = sos[L-1] (:298-306)
    sis.append(_gauss_bits(out_xo[i], ms[:, 0::2], ms[:, 1::2], training))
    qs.append(q)
    con = upsample2(q)
    for i in range(L - 2, -1, -1):
        q = tp.quantize(out_xo[i], mode)
        csc = masked_conv(q, sd, f"{pfx}csc_list.{i}.", "A", q.shape[1])     # groups = sos[i] (:274-277)
        plc = F.conv2d(con, sd[f"{pfx}plc_list.{i}.0.weight"], sd[f"{pfx}plc_list.{i}.0.bias"], padding=1)
        plc = F.leaky_relu(plc, LEAK)
        plc = F.conv2d(plc, sd[f"{pfx}plc_list.{i}.2.weight"], sd[f"{pfx}plc_list.{i}.2.bias"], padding=1)
        p0, p1, p2 = plc.chunk(3, dim=1)
        c0, c1, c2 = csc.chunk(3, dim=1)
        a = torch.cat((p0, c0, p1, c1, p2, c2), dim=1)
        for j, k in enumerate((0, 2, 4, 6)):
            # groups = inn_ch1 = sos[i+1] = channels of the parent level (:280-290); the chunking above is 3-way whatever clrch is
            a = F.conv2d(a, sd[f"{pfx}cgp_out_xo_list.{i}.{k}.weight"], sd[f"{pfx}cgp_out_xo_list.{i}.{k}.bias"], groups=con.shape[1])
            if j < 3:
                a = F.leaky_relu(a, LEAK)
        sis.append(_gauss_bits(out_xo[i], a[:, 0::2], a[:, 1::2], training))
        qs.append(q)
        con = upsample2(q)
    qs.reverse()
    sis.reverse()
    return si_xe, sis, xe_q, qs


def _eb_params(sd, pfx):
    names = [f"_matrix{i}" for i in range(5)] + [f"_bias{i}" for i in range(5)] + \
            [f"_factor{i}" for i in range(4)] + ["quantiles"]
    return {n: sd[pfx + n] for n in names}


def _eb(x, sd, pfx, training):
    y, lik = tp.entropy_bottleneck_forward(_eb_params(sd, pfx), x, training)
    return y, -torch.log2(lik)


def only_ezwt_forward(out_xe, out_xo, sd, pfx, L, training=False):
    """``onlyEZWT.forward`` (:804-840): coarsest level + LL factorized, finer levels
    conditioned on the upsampled parent only."""
    xe_q, si_xe = _eb(out_xe, sd, pfx + "ent_out_xe.", training)
    qs, sis = [], []
    q, si = _eb(out_xo[L - 1], sd, pfx + "ent_out_xo.", training)
    qs.append(q)
    sis.append(si)
    con = upsample2(q)
    for i in range(L - 2, -1, -1):
        a = F.conv2d(con, sd[f"{pfx}plc_list.{i}.0.weight"], sd[f"{pfx}plc_list.{i}.0.bias"], padding=1)
        a = F.leaky_relu(a, LEAK)
        a = F.conv2d(a, sd[f"{pfx}plc_list.{i}.2.weight"], sd[f"{pfx}plc_list.{i}.2.bias"], padding=1)
        a = F.leaky_relu(a, LEAK)
        a = F.conv2d(a, sd[f"{pfx}plc_list.{i}.4.weight"], sd[f"{pfx}plc_list.{i}.4.bias"])
        q, lik = tp.gaussian_conditional_forward(out_xo[i], a[:, 0::2], a[:, 1::2], training)
        sis.append(-torch.log2(lik))
        qs.append(q)
        con = upsample2(q)
    qs.reverse()
    sis.reverse()
    return si_xe, sis, xe_q, qs


def _dep_net(x, sd, pfx):
    """One ZTBlock context CNN: 3x3, 3x3, 1x1, 1x1, 1x1 with LeakyReLU (:618-680)."""
    a = x
    for j, k in enumerate((0, 2, 4, 6, 8)):
        w = sd[f"{pfx}{k}.weight"]
        a = F.conv2d(a, w, sd[f"{pfx}{k}.bias"], padding=w.shape[-1] // 2)
        if j < 4:
            a = F.leaky_relu(a, LEAK)
    return a


def ztblock_forward(out_xe, out_xo, sd, pfx, L, training=False):
    """``DWTConditioned2EntropyLayerZTBlock.forward`` (:691-757): parent + 2x2 polyphase
    block conditioning, phases ee -> eo -> oe -> oo."""
    mode = "noise" if training else "dequantize"
    xe_q, si_xe = _eb(out_xe, sd, pfx + "ent_out_xe.", training)
    qs, sis = [], []
    q, si = _eb(out_xo[L - 1], sd, pfx + "ent_out_xo.", training)
    qs.append(q)
    sis.append(si)
    con = q
    for i in range(0, L - 1):
        lvl = L - i - 2
        si_j, q_j = [], []
        for j in range(3):
            xin = out_xo[lvl][:, j:j + 1]
            B, _, H, W = xin.shape
            mu = torch.empty(B, 1, H, W, dtype=xin.dtype, device=xin.device)
            sg = torch.empty(B, 1, H, W, dtype=xin.dtype, device=xin.device)
            qq = tp.quantize(xin, mode)
            ee, eo, oe = qq[:, :, 0::2, 0::2], qq[:, :, 0::2, 1::2], qq[:, :, 1::2, 0::2]
            d1 = con[:, j:j + 1]
            n = j + i * 3
            mu[:, :, 0::2, 0::2] = _dep_net(d1, sd, f"{pfx}dep_1_list_mu.{n}.")
            sg[:, :, 0::2, 0::2] = _dep_net(d1, sd, f"{pfx}dep_1_list_sigma.{n}.")
            d2 = torch.cat((d1, ee), dim=1)
            mu[:, :, 0::2, 1::2] = _dep_net(d2, sd, f"{pfx}dep_2_list_mu.{n}.")
            sg[:, :, 0::2, 1::2] = _dep_net(d2, sd, f"{pfx}dep_2_list_sigma.{n}.")
            d3 = torch.cat((d1, ee, eo), dim=1)
            mu[:, :, 1::2, 0::2] = _dep_net(d3, sd, f"{pfx}dep_3_list_mu.{n}.")
            sg[:, :, 1::2, 0::2] = _dep_net(d3, sd, f"{pfx}dep_3_list_sigma.{n}.")
            d4 = torch.cat((d1, ee, eo, oe), dim=1)
            mu[:, :, 1::2, 1::2] = _dep_net(d4, sd, f"{pfx}dep_4_list_mu.{n}.")
            sg[:, :, 1::2, 1::2] = _dep_net(d4, sd, f"{pfx}dep_4_list_sigma.{n}.")
            si_j.append(_gauss_bits(xin, sg, mu, training))
            q_j.append(qq)
        sis.append(torch.cat(si_j, dim=1))
        con = torch.cat(q_j, dim=1)
        qs.append(con)
    qs.reverse()
    sis.reverse()
    return si_xe, sis, xe_q, qs


def factorized_forward(out_xe, out_xo, sd, pfx, L, training=False):
    """``DWTFactorizedEntropyLayer.forward`` (:215-231)."""
    qs, sis = [], []
    for i in range(L):
        q, si = _eb(out_xo[i], sd, f"{pfx}ent_out_xo_list.{i}.", training)
        qs.append(q)
        sis.append(si)
    xe_q, si_xe = _eb(out_xe, sd, pfx + "ent_out_xe.", training)
    return si_xe, sis, xe_q, qs


FORWARD = {
    "conditioned2ZTsepSubbands": cond2zt_forward,
    "onlyEZWT": only_ezwt_forward,
    "DWTConditioned2EntropyLayerZTBlock": ztblock_forward,
    "factorized": factorized_forward,
}
