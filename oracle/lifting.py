"""Functional CPU restatement of the learned lifting DWT (rows a1-a5 of SURVEY.md 8a).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Weights come from a
reference-layout ``state_dict``; ``prefix`` is e.g. ``"model0.autoencoder."``.
Arithmetic is plain torch on CPU, op for op what the reference executes, so it
doubles as the CPU baseline ("port") in ``bench.py``.
"""
import torch
import torch.nn.functional as F

LIFTING_COEFF = [-1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971,
                 0.869864451624781, 1.149604398860241]  # lifting_dwt_nets.py:431-432


def p_block(t, sd, pfx, linearity_flag=1, pad=2):
    """``P_block_v2.forward`` (graphs/layers/P_block_v2.py:40-55): conv1 -> tanh ->
    conv2 -> tanh -> conv3 + (pre-tanh conv1 output) -> conv4; every conv
    zero-pads its own input."""
    o1 = F.conv2d(t, sd[pfx + "conv1.weight"], sd[pfx + "conv1.bias"], padding=pad)
    a = torch.tanh(o1) if linearity_flag == 1 else o1
    a = F.conv2d(a, sd[pfx + "conv2.weight"], sd[pfx + "conv2.bias"], padding=pad)
    if linearity_flag == 1:
        a = torch.tanh(a)
    a = F.conv2d(a, sd[pfx + "conv3.weight"], sd[pfx + "conv3.bias"], padding=pad)
    a = a + o1
    return F.conv2d(a, sd[pfx + "conv4.weight"], sd[pfx + "conv4.bias"], padding=pad)


def _prefilter(x, w):
    """``convBlock[k]``: bias-free (3,1) cross-correlation, zero pad (1,0)
    (lifting_dwt_nets.py:805-819)."""
    return F.conv2d(x, w, None, padding=(1, 0))


def _step_blocks(pfx):
    # step -> (pre-filter, CNN): wavelet_forward_v2.py:58-74
    return [(pfx + "convBlock.0.weight", pfx + "P.0."), (pfx + "convBlock.1.weight", pfx + "U.0."),
            (pfx + "convBlock.2.weight", pfx + "P.1."), (pfx + "convBlock.3.weight", pfx + "U.1.")]


def lift_rows_forward(L, H, sd, pfx, cfg):
    """``lifting_forward_row_2_stage_lifting`` (wavelet_forward_v2.py:58-81)."""
    rw = cfg.res_connection_weight
    lin = cfg.linearity_flag
    pad = cfg.filtersize // 2
    blocks = _step_blocks(pfx)
    skip = _prefilter(L, sd[blocks[0][0]])
    H = H + skip + p_block(skip, sd, blocks[0][1], lin, pad) * rw
    skip = _prefilter(H, sd[blocks[1][0]])
    L = L + skip + p_block(skip, sd, blocks[1][1], lin, pad) * rw
    skip = _prefilter(L, sd[blocks[2][0]])
    H = H + skip + p_block(skip, sd, blocks[2][1], lin, pad) * rw
    skip = _prefilter(H, sd[blocks[3][0]])
    L = L + skip + p_block(skip, sd, blocks[3][1], lin, pad) * rw
    if cfg.scale == 1:
        H = H * (LIFTING_COEFF[4] + sd[pfx + "nh"] * 0.1)
        L = L * (LIFTING_COEFF[5] + sd[pfx + "nl"] * 0.1)
    return L, H


def lift_rows_inverse(L, H, sd, pfx, cfg):
    """``lifting_inverse_row_2_stage_lifting`` (wavelet_inverse_v2.py:68-92)."""
    rw = cfg.res_connection_weight
    lin = cfg.linearity_flag
    pad = cfg.filtersize // 2
    blocks = _step_blocks(pfx)
    if cfg.scale == 1:
        H = H / (LIFTING_COEFF[4] + sd[pfx + "nh"] * 0.1)
        L = L / (LIFTING_COEFF[5] + sd[pfx + "nl"] * 0.1)
    skip = _prefilter(H, sd[blocks[3][0]])
    L = L - skip - p_block(skip, sd, blocks[3][1], lin, pad) * rw
    skip = _prefilter(L, sd[blocks[2][0]])
    H = H - skip - p_block(skip, sd, blocks[2][1], lin, pad) * rw
    skip = _prefilter(H, sd[blocks[1][0]])
    L = L - skip - p_block(skip, sd, blocks[1][1], lin, pad) * rw
    skip = _prefilter(L, sd[blocks[0][0]])
    H = H - skip - p_block(skip, sd, blocks[0][1], lin, pad) * rw
    return L, H


def one_level_forward(x, sd, pfx, cfg):
    """``wavelet_forward_v2.one_level_lifting`` (wavelet_forward_v2.py:26-54):
    rows (even/odd rows), then the same on the transposed halves."""
    L = x[:, :, 0::2, :]
    H = x[:, :, 1::2, :]
    L, H = lift_rows_forward(L, H, sd, pfx, cfg)
    Lt = L.transpose(2, 3)
    LL, HL = lift_rows_forward(Lt[:, :, 0::2, :], Lt[:, :, 1::2, :], sd, pfx, cfg)
    Ht = H.transpose(2, 3)
    LH, HH = lift_rows_forward(Ht[:, :, 0::2, :], Ht[:, :, 1::2, :], sd, pfx, cfg)
    return LL.transpose(2, 3), LH.transpose(2, 3), HL.transpose(2, 3), HH.transpose(2, 3)


def _interleave_t(up, bot):
    """``reconstruct_fun`` (wavelet_inverse_v2.py:40-56): interleave along dim 2, transpose."""
    n, c, a, b = up.shape
    out = up.new_empty(n, c, 2 * a, b)
    out[:, :, 0::2, :] = up
    out[:, :, 1::2, :] = bot
    return out.transpose(2, 3)


def one_level_inverse(LL, LH, HL, HH, sd, pfx, cfg):
    """``wavelet_inverse_v2.one_level_lifting`` (wavelet_inverse_v2.py:20-38)."""
    a, b = lift_rows_inverse(LL.transpose(2, 3), HL.transpose(2, 3), sd, pfx, cfg)
    L = _interleave_t(a, b)
    a, b = lift_rows_inverse(LH.transpose(2, 3), HH.transpose(2, 3), sd, pfx, cfg)
    H = _interleave_t(a, b)
    L, H = lift_rows_inverse(L, H, sd, pfx, cfg)
    return _interleave_t(L, H).transpose(2, 3)


def transform_forward(x, sd, pfx, cfg):
    """Lifting levels of ``LiftingBasedNeuralWaveletv4.encode`` before the subband
    auto-encoders (lifting_dwt_nets.py:724-734): returns (LL, [Yh_l (B,3,h,w)])."""
    yh = []
    ll = x
    for lvl in range(cfg.dwtlevels):
        LL, LH, HL, HH = one_level_forward(ll, sd, f"{pfx}waveletForward.{lvl}.", cfg)
        yh.append(torch.cat((LH, HL, HH), dim=1))
        ll = LL
    return ll, yh


def transform_inverse(yl, yh, sd, pfx, cfg):
    """Inverse levels of ``LiftingBasedNeuralWaveletv4.decode``
    (lifting_dwt_nets.py:762-775), coarse to fine."""
    ll = yl
    for lvl in range(cfg.dwtlevels - 1, -1, -1):
        h = yh[lvl]
        ll = one_level_inverse(ll, h[:, 0:1], h[:, 1:2], h[:, 2:3], sd, f"{pfx}waveletInverse.{lvl}.", cfg)
    return ll
