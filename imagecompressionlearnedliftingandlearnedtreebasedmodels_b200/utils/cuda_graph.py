"""CUDA-graph replay of a module's inference call for one fixed input shape.

Small inputs (a config-5 tile, the deep levels of any transform) are launch-bound: a 5-level codec
forward of one 2048x256 tile is ~1 000 kernel launches of a few microseconds each.  All launches of
this package go through the C ABI on torch's *current* stream, every buffer comes from torch's
caching allocator and no op synchronises with the host, so the whole call can be captured once and
replayed as one graph launch (B200 playbook: "CUDA streams and graphs instead of a tracing
compiler").  No reference counterpart: the reference runs eager PyTorch.
"""
import torch


class GraphedForward:
    """``g = GraphedForward(model, example)``; ``g(x)`` == ``model(x)`` for inputs of ``example``'s shape.

    * inference only (captured under ``torch.no_grad()`` in the module's current train/eval mode);
    * the returned tensors are the graph's static output buffers: they are overwritten by the next
      call -- clone what must survive;
    * the kernels read *packed* copies of the weights (``ll_pack_*`` blobs made during the warm-up
      calls); the graph keeps pointing at those blobs, so build a new ``GraphedForward`` after
      ``load_state_dict`` or an optimiser step.
    """

    def __init__(self, module, example, warmup=3):
        if not example.is_cuda:
            raise ValueError("GraphedForward: the example input must live on a CUDA device (there is no CPU path)")
        self.module = module
        self.static_in = example.detach().clone()
        dev = example.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, int(warmup))):      # weight packing, function attributes, allocator growth
                module(self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = module(self.static_in)

    def __call__(self, x):
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"GraphedForward: captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, "
                             f"got {tuple(x.shape)} {x.dtype}")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out
