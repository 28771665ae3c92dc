"""GPU: the step engine's state handling -- learning rate as device state under graph replay, warm-up that leaves the
training state untouched, the trainer's checkpoint keys, and weight gradients of the image branch computed on a side
stream landing in the flat gradient buffer exactly as without the split (train20v2learn_wandb_proj.py:198-221, 255-260,
481-483, 524)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _f32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _batch(b, seed):
    from i2pnet_b200.synthetic import make_pairs
    return {k: v.to(DEV) for k, v in make_pairs(b, seed=seed, occupy_centres=(4, 8)).items()}


def test_learning_rate_is_device_state_of_the_captured_graph():
    """set_lr() / scheduler_step() / load_state_dict() change what REPLAYS of an already captured graph do."""
    from i2pnet_b200.engine import TrainStep
    _f32()
    eng = TrainStep(2, device=DEV, seed=0, use_graph=True)
    eng.load(_batch(2, 70))
    eng.warmup_and_capture(eager_steps=2)
    p0 = eng.opt.param.clone()
    eng.set_lr(0.0)
    eng.step()
    torch.cuda.synchronize()
    assert torch.equal(eng.opt.param, p0), "lr = 0 must leave the parameters where they were"
    assert float(eng.opt._step_count()) == 1.0
    eng.set_lr(1e-3)
    eng.step()
    torch.cuda.synchronize()
    d1 = float((eng.opt.param - p0).abs().max())
    assert 5e-4 < d1 < 3e-3, d1                      # Adam's early updates are ~lr per element
    eng.scheduler_step()
    assert abs(eng.current_lr() - 1e-3 * 0.99) < 1e-12 and eng.epoch == 1
    ck = eng.state_dict()
    assert set(ck) >= {"model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "epoch"}
    # the reference trainer's own objects load what we save ...
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in eng.model.parameters()]
    ref_opt = torch.optim.Adam(ref_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    ref_sched = torch.optim.lr_scheduler.ExponentialLR(ref_opt, 0.99)
    ref_opt.load_state_dict(ck["optimizer_state_dict"])
    ref_sched.load_state_dict(ck["scheduler_state_dict"])
    assert abs(ref_opt.param_groups[0]["lr"] - eng.current_lr()) < 1e-12 and ref_sched.last_epoch == 1
    ref_sched.step()
    # ... and we load what they save, with the captured graph still valid
    eng.load_state_dict({"model_state_dict": eng.model.state_dict(), "optimizer_state_dict": ref_opt.state_dict(),
                         "scheduler_state_dict": ref_sched.state_dict(), "epoch": 2})
    assert abs(eng.current_lr() - 1e-3 * 0.99 ** 2) < 1e-12 and eng.epoch == 2
    assert abs(float(eng.opt._lr_dev) - eng.current_lr()) < 1e-9
    eng.step()
    torch.cuda.synchronize()
    assert float(eng.opt._step_count()) == 3.0 and torch.isfinite(eng.loss).all()


def test_warmup_and_capture_leaves_the_training_state_as_loaded():
    from i2pnet_b200.engine import TrainStep
    _f32()
    eng = TrainStep(2, device=DEV, seed=3, use_graph=True)
    with pytest.raises(RuntimeError):
        eng.warmup_and_capture()                     # no batch loaded: refuses to train on zeros
    before = {k: v.clone() for k, v in eng.model.state_dict().items()}
    eng.load(_batch(2, 71))
    eng.warmup_and_capture(eager_steps=2)
    for k, v in eng.model.state_dict().items():
        assert torch.equal(v, before[k]), k
    assert float(eng.opt._step_count()) == 0.0
    assert float(eng.opt.exp_avg.abs().max()) == 0.0 and float(eng.opt.exp_avg_sq.abs().max()) == 0.0


def test_side_stream_weight_gradients_land_in_the_flat_buffer_under_replay():
    """The image branch's weight gradients are issued on a side stream and accumulate straight into the flat gradient
    buffer (engine.grad_sink): under graph replay they must equal what the single-stream step computes."""
    from i2pnet_b200 import streams
    from i2pnet_b200.engine import TrainStep
    _f32()
    bt = _batch(4, 72)
    grads = {}
    for side in (True, False):
        prev = streams.ENABLED
        streams.ENABLED = side
        try:
            eng = TrainStep(4, device=DEV, seed=0, use_graph=True)
            eng.set_lr(0.0)
            eng.load(bt)
            eng.warmup_and_capture(eager_steps=2)
            for _ in range(3):
                eng.step()
            torch.cuda.synchronize()
            grads[side] = {n: p.grad.detach().clone() for n, p in eng.model.named_parameters() if n.startswith("RGB_net")}
        finally:
            streams.ENABLED = prev
    gmax = max(float(g.norm()) for g in grads[False].values())
    for n, g in grads[False].items():
        assert float((grads[True][n] - g).norm()) <= 2e-3 * max(float(g.norm()), 1e-3 * gmax), n
