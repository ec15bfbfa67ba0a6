"""Data-parallel plumbing for encode/decode: the hot path shards by image (and tile) with no
exchange step, so all that is needed is a partition of the units over ranks and
two tiny reductions for reporting (max of the device-timed duration, sum of bits / squared
error).  ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is only plumbing here."""
import torch
import torch.distributed as dist


def shard_indices(n_units, rank, world):
    """Units (images / tiles) of rank ``rank``: ``rank::world``, as a list."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_units, world))


def tiles_of(height, width, n_tiles):
    """Cut an image into ``n_tiles`` independent tiles (config 5: the reference has no tiling, so
    tiles are independent images).  Returns (y0, y1, x0, x1) boxes; prefers vertical strips whose
    width stays a multiple of 32 so that five lifting levels divide evenly."""
    if n_tiles <= 0:
        raise ValueError("n_tiles must be positive")
    for nx in range(n_tiles, 0, -1):
        if n_tiles % nx == 0:
            ny = n_tiles // nx
            if width % nx == 0 and height % ny == 0 and (width // nx) % 32 == 0 and (height // ny) % 32 == 0:
                tw, th = width // nx, height // ny
                return [(j * th, (j + 1) * th, i * tw, (i + 1) * tw) for j in range(ny) for i in range(nx)]
    raise ValueError(f"cannot cut {height}x{width} into {n_tiles} tiles with sides divisible by 32")


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over ranks (the slowest rank's device time)."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(values, device=None, group=None):
    """Element-wise sum of a list of python floats over ranks (bits, squared error, pixels)."""
    if not (dist.is_available() and dist.is_initialized()):
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(v) for v in t.tolist()]


def allreduce_gradients(params, bucket_bytes=32 << 20, group=None):
    """Data-parallel training (BASELINE config 4): average the gradients of ``params`` over ranks with
    flat fp32 buckets (one ``all_reduce`` sum per bucket over NCCL / NVLink, or gloo in the CPU tests).
    Parameters without a gradient on this rank (``nh``/``nl`` when ``scale: 0``) contribute zeros, so every
    rank issues the same collectives.  Returns the number of collectives issued."""
    params = [p for p in params if p.requires_grad]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    n_coll, bucket, size = 0, [], 0

    def flush():
        nonlocal n_coll, bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off:off + n].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        n_coll += 1
        bucket, size = [], 0

    seen = set()
    for p in params:
        if id(p) in seen:          # shared lifting blocks are registered under several names
            continue
        seen.add(id(p))
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return n_coll


class GradientBuckets:
    """Bucketed gradient all-reduce overlapped with backward (BASELINE config 4).

    The unique trainable parameters are laid out, in reverse registration order (roughly the order backward
    produces their gradients), in flat fp32 buckets of ``bucket_bytes``; every ``p.grad`` is a *view* into its
    bucket, so autograd accumulates straight into the communication buffer (no gather/scatter copies).  A
    post-accumulate hook per parameter counts its bucket down and, at zero, issues the bucket's all-reduce
    asynchronously (NCCL runs it on its own stream while backward continues on the compute stream).
    ``finish()`` issues the buckets that never completed (parameters without a gradient this step: ``nh``/``nl``
    with ``scale: 0`` -- they contribute zeros, so every rank issues the same collectives), waits for all of them
    and divides by the world size.  Call ``zero()`` instead of ``optimizer.zero_grad()`` (which would drop the views).
    """

    def __init__(self, params, bucket_bytes=16 << 20, group=None):
        seen, uniq = set(), []
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        uniq.reverse()
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.world = dist.get_world_size(group) if self.active else 1
        self.buckets = []           # [flat, [params], pending, handle]
        cur, size = [], 0
        for p in uniq:
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(self._make(cur))
                cur, size = [], 0
        if cur:
            self.buckets.append(self._make(cur))
        self._bucket_of = {}
        self._hooks = []
        for bi, b in enumerate(self.buckets):
            for p in b["params"]:
                self._bucket_of[id(p)] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.collectives = 0          # all-reduces issued since construction
        self.zero()

    @staticmethod
    def _make(ps):
        n = sum(p.numel() for p in ps)
        flat = torch.zeros(n, dtype=torch.float32, device=ps[0].device)
        return {"flat": flat, "params": list(ps), "pending": len(ps), "handle": None}

    def zero(self):
        """Zero every bucket and (re)bind ``p.grad`` to its view."""
        for b in self.buckets:
            b["flat"].zero_()
            off = 0
            for p in b["params"]:
                n = p.numel()
                p.grad = b["flat"][off:off + n].view_as(p)
                off += n
            b["pending"] = len(b["params"])
            b["handle"] = None

    def _launch(self, b):
        if self.active and b["handle"] is None:
            b["handle"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.collectives += 1

    def _on_grad(self, p):
        b = self.buckets[self._bucket_of[id(p)]]
        if p.grad.data_ptr() != b["flat"].data_ptr() + self._offset(b, p) * 4:
            # autograd replaced the view (first accumulation into a None grad): copy into the bucket and rebind
            view = b["flat"][self._offset(b, p):self._offset(b, p) + p.numel()].view_as(p)
            view.copy_(p.grad)
            p.grad = view
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    @staticmethod
    def _offset(b, p):
        off = 0
        for q in b["params"]:
            if q is p:
                return off
            off += q.numel()
        raise KeyError("parameter not in bucket")

    def finish(self):
        """After ``loss.backward()``: flush, wait, average.  Returns the number of collectives of this step."""
        if not self.active:
            return 0
        for b in self.buckets:
            self._launch(b)
        for b in self.buckets:
            b["handle"].wait()
            b["flat"].div_(self.world)
        return len(self.buckets)

    def grad_bytes(self):
        return sum(b["flat"].numel() * 4 for b in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
