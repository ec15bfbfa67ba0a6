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
