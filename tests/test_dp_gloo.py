"""World-size-2 gloo test of the data-parallel gradient exchange (CPU): averaged flat-arena gradients equal the
gradient of the concatenated batch, buckets launch during backward, unused parameters reduce as zeros."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(8, 16)
        self.b = torch.nn.Linear(16, 4)
        self.unused = torch.nn.Linear(3, 3)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fusiontransformer_b200.dp import GradSync, shard_indices
    torch.manual_seed(100 + rank)               # deliberately different init per rank
    net = Net()
    sync = GradSync(net, bucket_bytes=256)      # tiny buckets -> several all-reduces
    sync.broadcast_parameters(net)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 8, generator=g)
    y = torch.randn(8, 4, generator=g)
    idx = shard_indices(8, rank, world)
    sync.zero_grad()
    loss = ((net(x[idx]) - y[idx]) ** 2).sum() / 8 * world      # per-rank mean-of-shard scaled so that avg == full mean
    loss.backward()
    launched_in_backward = sum(sync._launched)
    sync.finish()
    # numpy arrays are pickled by value; torch tensors would travel as shared-memory handles that die with the worker
    q.put((rank, [p.grad.clone().numpy() for p in net.parameters()], [p.data.clone().numpy() for p in net.parameters()],
           launched_in_backward, len(sync.buckets), idx))
    dist.barrier()
    dist.destroy_process_group()


def test_gradsync_world2_matches_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [(r, [torch.from_numpy(a) for a in g], [torch.from_numpy(a) for a in w], l, nb, i) for r, g, w, l, nb, i in res]
    (r0, g0, w0, l0, nb0, i0), (r1, g1, w1, l1, nb1, i1) = res
    assert sorted(i0 + i1) == list(range(8))
    for a, b in zip(w0, w1):
        assert torch.equal(a, b)                                 # broadcast made the replicas identical
    for a, b in zip(g0, g1):
        assert torch.allclose(a, b)                              # every rank holds the same averaged gradient
    assert nb0 > 1 and l0 >= 1                                   # at least one bucket was launched during backward
    net = Net()
    with torch.no_grad():
        for p, w in zip(net.parameters(), w0):
            p.copy_(w)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 8, generator=g)
    y = torch.randn(8, 4, generator=g)
    (((net(x) - y) ** 2).sum() / 8).backward()
    for p, got in zip(net.parameters(), g0):
        want = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(got, want, atol=1e-6)


def _worker_optimizer_zero_grad(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fusiontransformer_b200.dp import GradSync, shard_indices
    torch.manual_seed(5)
    net = Net()
    sync = GradSync(net, bucket_bytes=256)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 8, generator=g)
    y = torch.randn(8, 4, generator=g)
    idx = shard_indices(8, rank, world)
    out = []
    for step in range(2):
        opt.zero_grad()                          # set_to_none=True: the arena views are gone
        (((net(x[idx]) - y[idx]) ** 2).sum() / 8 * world).backward()
        sync.finish()
        out.append([None if p.grad is None else p.grad.clone().numpy() for p in net.parameters()])
        opt.step()
    q.put((rank, out, [p.data.clone().numpy() for p in net.parameters()]))
    dist.barrier()
    dist.destroy_process_group()


def test_optimizer_zero_grad_does_not_break_the_exchange():
    """``optimizer.zero_grad()`` sets the gradients to None; the exchange must still reduce what backward
    produced (not a stale arena) and leave every rank with the same averaged gradient and the same weights."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_optimizer_zero_grad, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, g0, w0), (_, g1, w1) = res
    for step in range(2):
        for a, b in zip(g0[step], g1[step]):
            assert (a is None) == (b is None)
            if a is not None:
                assert torch.allclose(torch.from_numpy(a), torch.from_numpy(b))
    for a, b in zip(w0, w1):
        assert torch.allclose(torch.from_numpy(a), torch.from_numpy(b))     # the replicas did not diverge
    # step 0 gradient == gradient of the full batch on one process
    torch.manual_seed(5)
    net = Net()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(8, 8, generator=g)
    y = torch.randn(8, 4, generator=g)
    (((net(x) - y) ** 2).sum() / 8).backward()
    for p, got in zip(net.parameters(), g0[0]):
        if p.grad is not None:
            assert torch.allclose(torch.from_numpy(got), p.grad, atol=1e-6)


def _worker_deferred(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fusiontransformer_b200.dp import GradSync
    torch.manual_seed(7)
    net = Net()
    sync = GradSync(net, bucket_bytes=256)
    sync.deferred = True                       # CUDA-graph mode: hooks launch nothing, finish() exchanges everything
    out = []
    for step in range(3):
        if step == 0:
            sync.zero_grad()                   # the capture step runs the Python body once ...
        else:
            sync.flat.zero_()                  # ... a replay only re-executes the captured memset: no Python zero_grad
        g = torch.Generator().manual_seed(10 * step + rank)
        x = torch.randn(4, 8, generator=g)
        net(x).sum().backward()
        assert sum(sync._launched) == 0 or step > 0
        sync.finish()
        out.append(sync.flat.clone().numpy())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_deferred_exchange_runs_on_every_step_without_python_zero_grad():
    """Regression: under graph replay zero_grad() is not executed in Python, so finish() must re-arm the buckets
    itself -- otherwise only capture steps exchange gradients and ranks that re-capture deadlock the others."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_deferred, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for step in range(3):
        a, b = torch.from_numpy(res[0][step]), torch.from_numpy(res[1][step])
        assert torch.allclose(a, b) and a.abs().sum() > 0        # averaged over ranks on every step, not only the first


def test_shard_indices_wraps_like_distributed_sampler():
    from fusiontransformer_b200.dp import shard_indices
    assert shard_indices(5, 0, 2) == [0, 2, 4] and shard_indices(5, 1, 2) == [1, 3, 0]
    from torch.utils.data.distributed import DistributedSampler
    ds = list(range(11))
    for r in range(4):
        s = DistributedSampler(ds, num_replicas=4, rank=r, shuffle=False)
        assert list(iter(s)) == shard_indices(11, r, 4)
