"""Differentiable torch (CUDA) formulations of the fused ops -- used ONLY by the backward pass.

The forward pass of every op runs on the hand-written kernels.  When gradients are required
(training, BASELINE config 4) the op is wrapped in ``_autograd.RecomputeFn``: its backward
re-evaluates the op with the functions below (plain torch ops on the GPU, TF32 off) under
``enable_grad`` and back-propagates through that graph.  Nothing here is reachable from an
inference call and nothing here runs on the CPU in the product.
"""
import torch
import torch.nn.functional as F

LIFTING_COEFF = [-1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971,
                 0.869864451624781, 1.149604398860241]


class exact_math:
    """cuDNN / cuBLAS TF32 off: the recompute must see the same fp32 function as the forward."""

    def __enter__(self):
        self.c, self.m = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False

    def __exit__(self, *a):
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.c, self.m


# ---- learned lifting -------------------------------------------------------------------------
# parameter packing order of one step: pre_w, (w1,b1,w2,b2,w3,b3,w4,b4); a level takes 4 steps + nh + nl
def _cnn(t, p, linear):
    o1 = F.conv2d(t, p[0], p[1], padding=2)
    a = o1 if linear else torch.tanh(o1)
    a = F.conv2d(a, p[2], p[3], padding=2)
    if not linear:
        a = torch.tanh(a)
    a = F.conv2d(a, p[4], p[5], padding=2) + o1
    return F.conv2d(a, p[6], p[7], padding=2)


def _step(src, dst, sp, sign, rw, linear):
    skip = F.conv2d(src, sp[0], None, padding=(1, 0))
    upd = skip + _cnn(skip, sp[1:], linear) * rw
    return dst + upd if sign > 0 else dst - upd


def lift_rows(L, H, steps, nh, nl, sign, rw, linear, scale):
    if sign > 0:
        H = _step(L, H, steps[0], 1, rw, linear)
        L = _step(H, L, steps[1], 1, rw, linear)
        H = _step(L, H, steps[2], 1, rw, linear)
        L = _step(H, L, steps[3], 1, rw, linear)
        if scale:
            H = H * (LIFTING_COEFF[4] + nh * 0.1)
            L = L * (LIFTING_COEFF[5] + nl * 0.1)
        return L, H
    if scale:
        H = H / (LIFTING_COEFF[4] + nh * 0.1)
        L = L / (LIFTING_COEFF[5] + nl * 0.1)
    L = _step(H, L, steps[3], -1, rw, linear)
    H = _step(L, H, steps[2], -1, rw, linear)
    L = _step(H, L, steps[1], -1, rw, linear)
    H = _step(L, H, steps[0], -1, rw, linear)
    return L, H


def _split_steps(params):
    return [params[9 * k:9 * k + 9] for k in range(4)], params[36], params[37]


def lift_level_fwd(x, params, rw, linear, scale):
    steps, nh, nl = _split_steps(params)
    L, H = lift_rows(x[:, :, 0::2, :], x[:, :, 1::2, :], steps, nh, nl, 1, rw, linear, scale)
    Lt, Ht = L.transpose(2, 3), H.transpose(2, 3)
    LL, HL = lift_rows(Lt[:, :, 0::2, :], Lt[:, :, 1::2, :], steps, nh, nl, 1, rw, linear, scale)
    LH, HH = lift_rows(Ht[:, :, 0::2, :], Ht[:, :, 1::2, :], steps, nh, nl, 1, rw, linear, scale)
    return LL.transpose(2, 3).contiguous(), torch.cat((LH.transpose(2, 3), HL.transpose(2, 3), HH.transpose(2, 3)), dim=1)


def _interleave_t(a, b):
    return torch.stack((a, b), dim=3).flatten(2, 3).transpose(2, 3)


def lift_level_inv(ll, yh, params, rw, linear, scale):
    steps, nh, nl = _split_steps(params)
    LH, HL, HH = yh[:, 0:1], yh[:, 1:2], yh[:, 2:3]
    a, b = lift_rows(ll.transpose(2, 3), HL.transpose(2, 3), steps, nh, nl, -1, rw, linear, scale)
    L = _interleave_t(a, b)
    a, b = lift_rows(LH.transpose(2, 3), HH.transpose(2, 3), steps, nh, nl, -1, rw, linear, scale)
    H = _interleave_t(a, b)
    L, H = lift_rows(L, H, steps, nh, nl, -1, rw, linear, scale)
    return _interleave_t(L, H).transpose(2, 3).contiguous()


# ---- CDF 9/7 (periodised, N >= 10) ---------------------------------------------------------------
def _taps(name, like):
    from .compat import BIOR44
    return torch.tensor(BIOR44[name], dtype=like.dtype, device=like.device)


def _analysis(x, dim):
    lo_f, hi_f = _taps("dec_lo", x), _taps("dec_hi", x)
    N = x.shape[dim]
    idx = (2 * torch.arange(N // 2, device=x.device).unsqueeze(1) + 5 - torch.arange(10, device=x.device).unsqueeze(0)) % N
    g = x.index_select(dim, idx.flatten()).unflatten(dim, (N // 2, 10))
    return (g * lo_f.view([-1 if d == dim + 1 else 1 for d in range(g.dim())])).sum(dim + 1), \
           (g * hi_f.view([-1 if d == dim + 1 else 1 for d in range(g.dim())])).sum(dim + 1)


def dwt97_fwd_level(x):
    """x (B,C,h,w), h,w >= 10 -> ll (B,C,h/2,w/2), yh (B,C,3,h/2,w/2)."""
    lo_w, hi_w = _analysis(x, 3)
    ll, lh = _analysis(lo_w, 2)
    hl, hh = _analysis(hi_w, 2)
    return ll, torch.stack((lh, hl, hh), dim=2)


def _synthesis(lo, hi, dim):
    rl, rh = _taps("rec_lo", lo), _taps("rec_hi", lo)
    n2 = lo.shape[dim]
    j = torch.arange(n2, device=lo.device)
    outs = []
    for par in (0, 1):
        acc = 0
        for t in range(5):
            n = (j + 2 - t) % n2
            acc = acc + rl[2 * t + par] * lo.index_select(dim, n) + rh[2 * t + par] * hi.index_select(dim, n)
        outs.append(acc)
    return torch.stack(outs, dim=dim + 1).flatten(dim, dim + 1)


def dwt97_inv_level(ll, yh):
    lo = _synthesis(ll, yh[:, :, 0], 2)
    hi = _synthesis(yh[:, :, 1], yh[:, :, 2], 2)
    return _synthesis(lo, hi, 3)


# ---- conv / rate ---------------------------------------------------------------------------------
def conv2d(x, w, b, groups, lrelu, upsample2):
    if upsample2:
        x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    y = F.conv2d(x, w, b, padding=w.shape[-1] // 2, groups=groups)
    return F.leaky_relu(y, 0.01) if lrelu else y


def _phi(z):
    return 0.5 * torch.erfc(float(-(2 ** -0.5)) * z)


def gauss_bits(x, ms, noise):
    from .compat import _LowerBoundFn
    sg, mu = ms[:, 0::2], ms[:, 1::2]
    y = x + noise if noise is not None else torch.round(x - mu) + mu   # torch.round: zero gradient, as in the reference
    v = torch.abs(y - mu)
    s = _LowerBoundFn.apply(sg, torch.tensor([0.11], device=x.device, dtype=x.dtype))
    p = _phi((0.5 - v) / s) - _phi((-0.5 - v) / s)
    p = _LowerBoundFn.apply(p, torch.tensor([1e-9], device=x.device, dtype=x.dtype))
    return -torch.log2(p), y


def eb_bits(x, eb_params, noise):
    """eb_params: the 15 tensors in registration order (see ops.pack_eb)."""
    from .compat import _LowerBoundFn
    B, C, H, W = x.shape
    M = [eb_params[0], eb_params[3], eb_params[6], eb_params[9], eb_params[12]]
    Bv = [eb_params[1], eb_params[4], eb_params[7], eb_params[10], eb_params[13]]
    Fv = [eb_params[2], eb_params[5], eb_params[8], eb_params[11]]
    q = eb_params[14]
    v = x.permute(1, 0, 2, 3).reshape(C, 1, -1)
    med = q[:, :, 1:2].detach()
    y = v + noise.permute(1, 0, 2, 3).reshape(C, 1, -1) if noise is not None else torch.round(v - med) + med

    def logits(t):
        for i in range(5):
            t = torch.matmul(F.softplus(M[i]), t) + Bv[i]
            if i < 4:
                t = t + torch.tanh(Fv[i]) * torch.tanh(t)
        return t

    lo, up = logits(y - 0.5), logits(y + 0.5)
    sign = -torch.sign(lo + up).detach()
    p = torch.abs(torch.sigmoid(sign * up) - torch.sigmoid(sign * lo))
    p = _LowerBoundFn.apply(p, torch.tensor([1e-9], device=x.device, dtype=x.dtype))
    bits = -torch.log2(p)
    back = lambda t: t.reshape(C, B, H, W).permute(1, 0, 2, 3).contiguous()
    return back(y), back(bits)
