"""N > 1 host logic on CPU: world_size-2 gloo group exercising the flat gradient bucket + mean all-reduce + parameter
broadcast used by bench.py's data-parallel step (no CUDA kernels involved)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from holocron_b200.distributed import GradBucket, broadcast_parameters, shard_batch


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
