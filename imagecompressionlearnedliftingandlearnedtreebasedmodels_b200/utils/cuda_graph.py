"""CUDA-graph replay of a module's inference call for one fixed input shape.

Small inputs (a config-5 tile, the deep levels of any transform) are launch-bound: a 5-level codec
forward of one 2048x256 tile is ~1 000 kernel launches of a few microseconds each.  All launches of
this package go through the C ABI on torch's *current* stream, every buffer comes from torch's
caching allocator and no op synchronises with the host, so the whole call can be captured once and
replayed as one graph launch (B200 playbook: "CUDA streams and graphs instead of a tracing
compiler").  No reference counterpart: the reference runs eager PyTorch.
"""
import torch


def _tree_map(fn, obj):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(_tree_map(fn, o) for o in obj)
    raise TypeError(f"GraphedForward: inputs must be tensors or lists / tuples of tensors, got {type(obj).__name__}")


def _tree_leaves(obj, out):
    if torch.is_tensor(obj):
        out.append(obj)
    else:
        for o in obj:
            _tree_leaves(o, out)
    return out


class GraphedForward:
    """``g = GraphedForward(model, *example)``; ``g(*inputs)`` == ``model(*inputs)`` for inputs shaped like ``example``
    (each argument a tensor or a list / tuple of tensors, e.g. an entropy layer's ``(out_xe, out_xo_list)``).

    * inference only (captured under ``torch.no_grad()`` in the module's current train/eval mode);
    * the returned tensors are the graph's static output buffers: they are overwritten by the next
      call -- clone what must survive;
    * the kernels read *packed* copies of the weights (``ll_pack_*`` blobs made during the warm-up
      calls); the graph keeps pointing at those blobs, so build a new ``GraphedForward`` after
      ``load_state_dict`` or an optimiser step.
    """

    def __init__(self, module, *example, warmup=3):
        leaves = _tree_leaves(example, [])
        if not leaves or not all(t.is_cuda for t in leaves):
            raise ValueError("GraphedForward: the example inputs must live on a CUDA device (there is no CPU path)")
        self.module = module
        self.static_in = _tree_map(lambda t: t.detach().clone(), example)
        self._in_leaves = _tree_leaves(self.static_in, [])
        dev = leaves[0].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, int(warmup))):      # weight packing, function attributes, allocator growth
                module(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = module(*self.static_in)

    def __call__(self, *inputs):
        leaves = _tree_leaves(inputs, [])
        if len(leaves) != len(self._in_leaves) or any(a.shape != b.shape or a.dtype != b.dtype
                                                      for a, b in zip(leaves, self._in_leaves)):
            raise ValueError("GraphedForward: inputs differ in count, shape or dtype from the captured example "
                             f"({[tuple(t.shape) for t in self._in_leaves]})")
        for dst, src in zip(self._in_leaves, leaves):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
