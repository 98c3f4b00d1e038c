"""N > 1 host logic on CPU: world_size-2 gloo group exercising the flat gradient bucket + mean all-reduce + parameter
broadcast used by bench.py's data-parallel step (no CUDA kernels involved)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from holocron_b200.distributed import GradBucket, OverlappedReducer, broadcast_parameters, shard_batch


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different init per rank on purpose
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.BatchNorm2d(8), torch.nn.Flatten(),
                                torch.nn.Linear(8 * 4 * 4, 5)).to(memory_format=torch.channels_last)
    broadcast_parameters(model)
    bucket = GradBucket(model.parameters())
    # every .grad is a view into the flat buffer, with the parameter's own layout
    for p in model.parameters():
        assert p.grad.shape == p.shape
        assert p.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr()
    torch.manual_seed(7)
    x = torch.randn(8, 3, 4, 4)
    t = torch.randint(0, 5, (8,))
    idx = list(shard_batch(8, rank, world))
    loss = torch.nn.functional.cross_entropy(model(x[idx]), t[idx])
    loss.backward()
    bucket.all_reduce_mean()
    grads = [p.grad.detach().clone() for p in model.parameters()]
    params = [p.detach().clone() for p in model.parameters()]
    bucket.zero_()
    assert all(float(p.grad.abs().sum()) == 0 for p in model.parameters())
    torch.save((params, grads), os.path.join(out, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_bucket_allreduce_matches_single_process(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    (p0, g0), (p1, g1) = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    # replicas hold identical parameters and identical (averaged) gradients
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)
    for a, b in zip(g0, g1):
        assert torch.allclose(a, b, atol=1e-7)
    # the averaged gradient equals the mean of the two shard gradients computed in one process
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.BatchNorm2d(8), torch.nn.Flatten(),
                                torch.nn.Linear(8 * 4 * 4, 5))
    with torch.no_grad():
        for p, v in zip(model.parameters(), p0):
            p.copy_(v)
    torch.manual_seed(7)
    x = torch.randn(8, 3, 4, 4)
    t = torch.randint(0, 5, (8,))
    ref = None
    for r in range(world):
        model.zero_grad()
        idx = list(shard_batch(8, r, world))
        torch.nn.functional.cross_entropy(model(x[idx]), t[idx]).backward()
        gs = [p.grad.clone() / world for p in model.parameters()]
        ref = gs if ref is None else [a + b for a, b in zip(ref, gs)]
    for a, b in zip(g0, ref):
        assert torch.allclose(a, b, atol=1e-6)


def test_shard_batch_covers_everything():
    for n, w in ((256, 8), (10, 3), (5, 8)):
        seen = [i for r in range(w) for i in shard_batch(n, r, w)]
        assert seen == list(range(n))


class _Staged(torch.nn.Module):
    """Toy model with a `features` Sequential of four parameterised stages + head (the shape stage_boundaries expects)."""

    def __init__(self) -> None:
        super().__init__()
        def stage(cin, cout):
            return torch.nn.Sequential(torch.nn.Conv2d(cin, cout, 3, padding=1), torch.nn.BatchNorm2d(cout), torch.nn.ReLU())
        self.features = torch.nn.Sequential(stage(3, 8), stage(8, 8), stage(8, 16), stage(16, 16))
        self.head = torch.nn.Linear(16, 5)

    def forward(self, x):
        return self.head(self.features(x).mean((2, 3)))


def _worker_overlap(rank: int, world: int, port: int, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(3)
    model = _Staged()
    bucket = GradBucket(model.parameters())
    bounds = OverlappedReducer.stage_boundaries(model)
    assert len(bounds) == 2                                   # stages 1 and 2: chunks {head, stage 3}, {stage 2}, {stages 1, 0}
    reducer = OverlappedReducer(bucket, bounds)
    assert reducer.ends == sorted(reducer.ends, reverse=True) and reducer.ends[-1] > 0
    torch.manual_seed(50 + rank)
    x, t = torch.randn(4, 3, 6, 6), torch.randint(0, 5, (4,))
    grads = []
    for _ in range(2):                                        # two steps: the per-step state of the reducer resets
        torch.nn.functional.cross_entropy(model(x), t).backward()
        reducer.finish()
        grads.append(bucket.flat.clone())
        bucket.zero_()
    # reference: plain single all-reduce of the locally computed gradient
    torch.nn.functional.cross_entropy(model(x), t).backward()
    # (the hooks fired again during this backward and reduced their prefixes; finish() completes the tail)
    reducer.finish()
    local = bucket.flat.clone()
    torch.save((grads, local), os.path.join(out, f"ovl{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_reducer_equals_mean_allreduce(tmp_path):
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker_overlap, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    (g0, l0), (g1, l1) = [torch.load(tmp_path / f"ovl{r}.pt") for r in range(world)]
    # chunked reduction gives every rank the same averaged bucket, step after step
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)
    assert torch.allclose(g0[0], g0[1], atol=1e-7) and torch.equal(l0, l1) and torch.allclose(l0, g0[0], atol=1e-7)
    # ... equal to the mean of the two ranks' local gradients (recomputed in this process)
    ref = None
    for rank in range(world):
        torch.manual_seed(3)
        model = _Staged()
        torch.manual_seed(50 + rank)
        x, t = torch.randn(4, 3, 6, 6), torch.randint(0, 5, (4,))
        bucket = GradBucket(model.parameters())
        torch.nn.functional.cross_entropy(model(x), t).backward()
        ref = bucket.flat.clone() / world if ref is None else ref + bucket.flat / world
    assert torch.allclose(g0[0], ref, atol=1e-6)
