"""CPU, world_size 2 over gloo: the data-parallel exchange of the training step -- one flat
gradient bucket, one all-reduce, mean over ranks."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from i2pnet_b200.engine import FlatGradBucket
    torch.manual_seed(0)                                   # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    bucket = FlatGradBucket(net.parameters())
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
    g = torch.Generator().manual_seed(100 + rank)          # different shard per rank
    x, y = torch.randn(4, 6, generator=g), torch.randn(4, 3, generator=g)
    bucket.zero()
    torch.nn.functional.mse_loss(net(x), y).backward()     # accumulates INTO the flat buffer
    local = bucket.flat.clone()
    # the step engine's form: drop p.grad, let autograd keep its own tensors, pack them with one cat
    bucket.release()
    assert all(p.grad is None for p in net.parameters())
    torch.nn.functional.mse_loss(net(x), y).backward()
    bucket.gather()
    assert torch.equal(bucket.flat, local)
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in net.parameters())
    bucket.all_reduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    assert torch.allclose(bucket.flat, sum(gathered) / world, atol=1e-7)
    norm = bucket.clip_(1e-3)
    assert float(torch.linalg.vector_norm(bucket.flat)) <= 1e-3 * (1 + 1e-4) and float(norm) > 1e-3
    # the single-process equivalent: the mean gradient over the global batch
    if rank == 0:
        xs, ys = [], []
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            xs.append(torch.randn(4, 6, generator=gr)); ys.append(torch.randn(4, 3, generator=gr))
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        torch.nn.functional.mse_loss(ref(torch.cat(xs)), torch.cat(ys)).backward()
        want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        assert torch.allclose(sum(gathered) / world, want, atol=1e-6)
    ret[rank] = True
    dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
