"""N>1 host logic on the CPU: world_size-2 gloo run of the sharding helpers bench.py and the
multi-GPU encode path use (no data-path collective exists on this path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from imagecompressionlearnedliftingandlearnedtreebasedmodels_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = parallel.shard_indices(7, rank, world)
        slowest = parallel.max_over_ranks(10.0 + rank)
        total_bits, total_px = parallel.sum_over_ranks([100.0 * (rank + 1), float(len(mine))])
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            out.put((gathered, slowest, total_bits, total_px))
    finally:
        dist.destroy_process_group()


def test_world2_sharding_and_reductions():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    gathered, slowest, bits, px = q.get()
    assert sorted(gathered[0] + gathered[1]) == list(range(7))      # a partition: every unit exactly once
    assert not set(gathered[0]) & set(gathered[1])
    assert slowest == 11.0 and bits == 300.0 and px == 7.0


def test_tiles_config5():
    boxes = parallel.tiles_of(2048, 2048, 8)
    assert len(boxes) == 8
    cover = torch.zeros(2048, 2048, dtype=torch.int32)
    for y0, y1, x0, x1 in boxes:
        assert (y1 - y0) % 32 == 0 and (x1 - x0) % 32 == 0           # 5 lifting levels divide evenly
        cover[y0:y1, x0:x1] += 1
    assert int(cover.min()) == 1 and int(cover.max()) == 1


def _grad_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        shared = torch.nn.Parameter(torch.zeros(5))
        ps = [torch.nn.Parameter(torch.zeros(3, 4)), shared, shared, torch.nn.Parameter(torch.zeros(7)),
              torch.nn.Parameter(torch.zeros(2), requires_grad=False)]
        ps[0].grad = torch.full((3, 4), float(rank + 1))
        shared.grad = torch.arange(5.0) * (rank + 1)
        # ps[3] has no gradient on rank 1 (an unused parameter there): treated as zeros
        if rank == 0:
            ps[3].grad = torch.ones(7)
        n = parallel.allreduce_gradients(ps, bucket_bytes=40)   # tiny buckets: several collectives
        if rank == 0:
            out.put((n, ps[0].grad.tolist(), shared.grad.tolist(), ps[3].grad.tolist(), ps[4].grad))   # plain lists: no shm handles
    finally:
        dist.destroy_process_group()


def test_world2_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    n, g0, gs, g3, g4 = q.get()
    assert n >= 2
    assert torch.allclose(torch.tensor(g0), torch.full((3, 4), 1.5))
    assert torch.allclose(torch.tensor(gs), torch.arange(5.0) * 1.5)
    assert torch.allclose(torch.tensor(g3), torch.full((7,), 0.5))
    assert g4 is None


class _TinyNet(torch.nn.Module):
    """Shared parameter (registered twice, like the lifting blocks) + a parameter that gets no gradient (nh / nl)."""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(6, 5)
        self.b = torch.nn.Linear(5, 4)
        self.again = self.a                      # alias
        self.unused = torch.nn.Parameter(torch.ones(3))

    def forward(self, x):
        return self.b(torch.tanh(self.again(x))).pow(2).sum()


def _bucket_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = _TinyNet()
        bk = parallel.GradientBuckets(net.parameters(), bucket_bytes=64)     # tiny buckets: several collectives
        res = []
        for step in range(2):                    # two steps: zero() must rebind the views and clear the sums
            bk.zero()
            torch.manual_seed(10 * step + rank)
            net(torch.randn(3, 6)).backward()
            n = bk.finish()
            res.append((n, [p.grad.clone().tolist() for p in net.parameters()]))
        if rank == 0:
            out.put(res)
    finally:
        dist.destroy_process_group()


def test_world2_overlapped_gradient_buckets():
    """``GradientBuckets`` (hooks + async bucketed all-reduce) == the mean of the two ranks' plain autograd gradients."""
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = q.get()
    for step, (n, grads) in enumerate(res):
        assert n >= 2
        want = None
        for rank in range(2):
            torch.manual_seed(0)
            net = _TinyNet()
            torch.manual_seed(10 * step + rank)
            net(torch.randn(3, 6)).backward()
            g = [(p.grad if p.grad is not None else torch.zeros_like(p)) for p in net.parameters()]
            want = g if want is None else [a + b for a, b in zip(want, g)]
        for got, w in zip(grads, want):
            assert torch.allclose(torch.tensor(got), w / 2, atol=1e-6)


def test_gradient_buckets_single_process_is_a_no_op():
    net = _TinyNet()
    bk = parallel.GradientBuckets(net.parameters())
    bk.zero()
    net(torch.ones(2, 6)).backward()
    assert bk.finish() == 0 and net.a.weight.grad.abs().sum() > 0 and float(net.unused.grad.abs().sum()) == 0.0
