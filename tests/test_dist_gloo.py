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


def _worker_fused(rank, world, port, ret):
    """The fused-step form of the exchange: aligned flat layout, all-reduce SUM (the averaging happens inside the
    optimiser step), checked against clip_grad_norm_ + torch.optim.Adam on the mean gradient of the global batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from i2pnet_b200.engine import FlatGradBucket
    from oracle.optim_cpu import clip_adam_step
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    bucket = FlatGradBucket(net.parameters(), align=64)
    assert all(o % 64 == 0 for o in bucket.offsets)
    assert bucket.flat.numel() % 64 == 0 and len(bucket.offsets) == 4
    # flat parameter / moment buffers with the bucket's layout (what engine.FlatAdam sets up on the GPU)
    flat_p = torch.zeros_like(bucket.flat)
    for p, o in zip(bucket.params, bucket.offsets):
        view = flat_p[o:o + p.numel()].view_as(p)
        view.copy_(p.detach())
        p.data = view
    m, v, step = torch.zeros_like(flat_p), torch.zeros_like(flat_p), 0
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ref_opt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=1e-4)
    for it in range(3):
        batches = []
        for r in range(world):
            gr = torch.Generator().manual_seed(1000 * it + r)
            batches.append((torch.randn(4, 6, generator=gr) * 30, torch.randn(4, 3, generator=gr) * 30))
        bucket.release()
        torch.nn.functional.mse_loss(net(batches[rank][0]), batches[rank][1]).backward()
        bucket.gather()
        used = torch.zeros_like(bucket.flat, dtype=torch.bool)
        for p, o in zip(bucket.params, bucket.offsets):
            used[o:o + p.numel()] = True
        assert float(bucket.flat[~used].abs().sum()) == 0.0            # the padding carries nothing
        assert bucket.all_reduce_sum() == world
        step = clip_adam_step(flat_p, bucket.flat, m, v, step, max_norm=10.0, world=world)
        ref_opt.zero_grad()
        torch.nn.functional.mse_loss(ref(torch.cat([b[0] for b in batches])), torch.cat([b[1] for b in batches])).backward()
        norm = torch.nn.utils.clip_grad_norm_(ref.parameters(), 10.0)
        assert float(norm) > 10.0                                       # clipping is active
        ref_opt.step()
        for a, b in zip(net.parameters(), ref.parameters()):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-7), (it, float((a - b).abs().max()))
    ret[rank] = True
    dist.destroy_process_group()


def test_fused_step_exchange_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker_fused, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_flat_adam_checkpoint_round_trips_through_torch_adam_format():
    """engine.FlatAdam.state_dict() / load_state_dict() speak torch.optim.Adam's format (the reference trainer's
    `optimizer_state_dict`): moments and step count survive FlatAdam -> torch Adam -> FlatAdam.  Host logic only (CPU)."""
    from i2pnet_b200.engine import FlatAdam, FlatGradBucket
    torch.manual_seed(1)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
    ref = torch.optim.Adam(net.parameters(), lr=2e-3, weight_decay=1e-4)
    for _ in range(3):                                          # a torch Adam with three steps of history
        ref.zero_grad()
        net(torch.randn(4, 6)).square().mean().backward()
        ref.step()
    want = ref.state_dict()
    opt = FlatAdam(FlatGradBucket(net.parameters(), align=64), lr=1e-3)
    opt.load_state_dict(want)
    assert opt.lr == 2e-3 and float(opt._step_count()) == 3.0
    got = opt.state_dict()
    assert got["param_groups"][0]["params"] == want["param_groups"][0]["params"]
    for i, st in want["state"].items():
        assert torch.equal(got["state"][i]["exp_avg"], st["exp_avg"]) and torch.equal(got["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
        assert float(got["state"][i]["step"]) == float(st["step"]) == 3.0
    fresh = torch.optim.Adam(net.parameters(), lr=1e-3)
    fresh.load_state_dict(got)                                   # and torch's own optimiser accepts what FlatAdam wrote
    assert fresh.state_dict()["param_groups"][0]["lr"] == 2e-3
    # the padding of the flat moment buffers stays zero
    used = torch.zeros_like(opt.exp_avg, dtype=torch.bool)
    for p, o in zip(opt.bucket.params, opt.bucket.offsets):
        used[o:o + p.numel()] = True
    assert float(opt.exp_avg[~used].abs().sum()) == 0.0
