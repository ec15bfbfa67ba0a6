"""Autograd bridge for the fused kernels: forward on the CUDA kernels, backward by recomputing
the op with ``_torch_ref`` (torch CUDA ops) and back-propagating through that graph.

Memory: only the op's inputs are saved (no intermediate activation of the fused kernels ever
exists in HBM, forward or backward-saved); compute: one extra torch forward per op in backward.
"""
import torch

from . import _torch_ref


def needs_grad(tensors):
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in tensors)


class RecomputeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fast_fn, ref_fn, *tensors):
        ctx.ref_fn = ref_fn
        ctx.save_for_backward(*tensors)
        with torch.no_grad():
            out = fast_fn(*[t.detach() if t is not None else None for t in tensors])
        ctx.single = torch.is_tensor(out)
        return out if ctx.single else tuple(out)

    @staticmethod
    def backward(ctx, *grads):
        tensors = ctx.saved_tensors
        ins = [t.detach().requires_grad_(True) if (t.is_floating_point() and ctx.needs_input_grad[i + 2]) else t.detach()
               for i, t in enumerate(tensors)]
        with torch.enable_grad(), _torch_ref.exact_math():
            out = ctx.ref_fn(*ins)
            outs = [out] if torch.is_tensor(out) else list(out)
            pairs = [(o, g) for o, g in zip(outs, grads) if g is not None and o.requires_grad]
            wrt = [t for t in ins if t.requires_grad]
            got = torch.autograd.grad([o for o, _ in pairs], wrt, [g for _, g in pairs], allow_unused=True) if pairs and wrt else []
        it = iter(got)
        res = [next(it) if t.requires_grad else None for t in ins]
        return (None, None, *res)


def run(fast_fn, ref_fn, tensors):
    """``fast_fn(*tensors)`` on the kernels; differentiable via ``ref_fn`` when a gradient is needed."""
    if needs_grad(tensors):
        return RecomputeFn.apply(fast_fn, ref_fn, *tensors)
    return fast_fn(*tensors)


def run_module(fast_fn, module, x):
    """``fast_fn(x)`` on the kernels; differentiable by re-running the torch ``module`` (an nn.Sequential of
    plain torch layers that defines the same function) functionally on the saved parameters."""
    names, params = zip(*module.named_parameters())
    if not needs_grad((x,) + params):
        return fast_fn(x)

    def ref(x, *ps):
        return torch.func.functional_call(module, dict(zip(names, ps)), (x,))

    return RecomputeFn.apply(lambda x, *ps: fast_fn(x), ref, x, *params)
