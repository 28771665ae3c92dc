"""GPU: the fused clip + Adam step (csrc/optim.cu behind engine.FlatAdam) against the reference trainer's
torch.nn.utils.clip_grad_norm_(parameters, 10) + torch.optim.Adam(lr 1e-3, weight_decay 1e-4)
(train20v2learn_wandb_proj.py:198-205, 481-483), over several steps, with and without clipping active, and with the
gradient arriving as a sum over ranks."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("grad_scale", [1e-3, 30.0], ids=["unclipped", "clipped"])
def test_flat_adam_matches_torch_adam(world, grad_scale):
    from i2pnet_b200.engine import FlatAdam, FlatGradBucket
    dev = torch.device("cuda:0")
    gen = torch.Generator(device=dev).manual_seed(3)
    shapes = [(16, 10, 1, 1), (16,), (33, 7), (1,), (128, 262), (5, 3, 3, 3)]
    mine = [torch.nn.Parameter(torch.randn(s, device=dev, generator=gen)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    bucket = FlatGradBucket(mine, align=64)
    opt = FlatAdam(bucket, lr=1e-3, weight_decay=1e-4, max_norm=10.0)
    assert all(p.data_ptr() % 256 == 0 for p in mine)            # re-seated at the bucket's aligned offsets
    ref_opt = torch.optim.Adam(ref, lr=1e-3, weight_decay=1e-4)
    for step in range(6):
        grads = [torch.randn(s, device=dev, generator=gen) * grad_scale * (1 + step) for s in shapes]
        bucket.release()
        for p, g in zip(mine, grads):
            p.grad = g * world                                   # what the all-reduce (sum) would leave in the bucket
        bucket.gather()
        opt.step(world)
        for p, g in zip(ref, grads):
            p.grad = g.clone()
        norm = torch.nn.utils.clip_grad_norm_(ref, 10.0)
        assert (float(norm) > 10.0) == (grad_scale > 1)
        ref_opt.step()
        for a, b in zip(mine, ref):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (step, float((a - b).abs().max()))
    # the padding between parameters stays zero in every flat buffer
    used = torch.zeros_like(bucket.flat, dtype=torch.bool)
    for o, p in zip(bucket.offsets, bucket.params):
        used[o:o + p.numel()] = True
    assert float(opt.param[~used].abs().max()) == 0.0 and float(opt.exp_avg[~used].abs().max()) == 0.0


def test_train_step_fused_optimizer_tracks_stock_optimizer():
    """Three eager training steps of the whole engine, fused optimiser vs torch.optim.Adam + clip: same loss trajectory."""
    from i2pnet_b200.engine import TrainStep
    from i2pnet_b200.synthetic import make_pairs
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    losses = {}
    for fused in (True, False):
        eng = TrainStep(2, device=dev, seed=0, use_graph=False, fused_optimizer=fused)
        for head in (eng.model.l4_head, eng.model.l3_head):
            head.DP1.p = 0.0
        out = []
        for i in range(3):
            eng.load({k: v.to(dev) for k, v in make_pairs(2, seed=40 + i, occupy_centres=(4, 8)).items()})
            eng.step()
            out.append(float(eng.loss.item()))
        losses[fused] = out
    # Same initial parameters: the first loss agrees to rounding.  Afterwards the trajectories may drift by per cent:
    # Adam's first updates are ~lr * sign(g), so last-bit noise in near-zero gradients (atomics) moves weights by 1e-3;
    # two runs of the SAME engine differ by that much.  The optimiser's arithmetic is pinned by the test above.
    assert abs(losses[True][0] - losses[False][0]) < 1e-4 * abs(losses[False][0]), losses
    for a, b in zip(losses[True][1:], losses[False][1:]):
        assert abs(a - b) < 0.1 * abs(b), losses


def test_host_pipeline_feeds_the_right_batches_and_losses():
    """engine.HostPipeline (double-buffered host feed, losses read one step late): step i runs on batch i, and the
    loss handed out one call later is that step's.  (Loss trajectories of two engines are not comparable bit for bit:
    the scatter / dW atomics make a training step non-deterministic in the last bits.)"""
    from i2pnet_b200.engine import INPUT_KEYS, HostPipeline, TrainStep
    from i2pnet_b200.synthetic import make_pairs
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    host = [{k: v.pin_memory() for k, v in make_pairs(2, seed=60 + i).items()} for i in range(4)]
    eng = TrainStep(2, device=dev, seed=0, use_graph=True)
    eng.load({k: v.to(dev) for k, v in host[0].items()})
    eng.warmup_and_capture(eager_steps=2)
    torch.cuda.synchronize()
    pipe, seen, handed = HostPipeline(eng), [], []
    pipe.submit(host[0])
    for i in range(4):
        pipe.step()
        if i + 1 < 4:
            pipe.submit(host[i + 1])          # overlaps step i; must not disturb the inputs step i is reading
        v = pipe.loss()
        if v is not None:
            handed.append(v)
        torch.cuda.synchronize()
        for k in INPUT_KEYS:
            assert torch.equal(eng.inputs[k].cpu(), host[i][k]), (i, k)
        seen.append(float(eng.loss.item()))
    handed.append(pipe.drain())
    assert handed == seen and all(v == v and v > 0 for v in seen), (handed, seen)
