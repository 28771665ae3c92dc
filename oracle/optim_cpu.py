"""CPU restatement of the reference trainer's optimiser step on flat buffers (TEST INFRASTRUCTURE ONLY, like the rest
of oracle/): torch.nn.utils.clip_grad_norm_(parameters, max_norm) followed by torch.optim.Adam(lr, betas, eps,
weight_decay).step() (train20v2learn_wandb_proj.py:198-205, 481-483), with the data-parallel averaging folded in the
way csrc/optim.cu does it: `grad_sum` is the SUM of the ranks' gradients."""
import math

import torch


def clip_adam_step(param, grad_sum, exp_avg, exp_avg_sq, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4,
                   max_norm=10.0, world=1):
    """In-place on param / exp_avg / exp_avg_sq (flat f32 tensors); step = number of updates already made.
    -> step + 1"""
    b1, b2 = betas
    inv_world = 1.0 / world
    total_norm = float(torch.linalg.vector_norm(grad_sum.double())) * inv_world         # norm of the mean gradient
    coef = min(max_norm / (total_norm + 1e-6), 1.0) if max_norm > 0 else 1.0            # clip_grad_norm_
    g = grad_sum * (inv_world * coef) + weight_decay * param                            # Adam's L2 weight decay
    exp_avg.mul_(b1).add_(g, alpha=1 - b1)
    exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)
    t = step + 1
    step_size = lr / (1 - b1 ** t)
    denom = exp_avg_sq.sqrt() / math.sqrt(1 - b2 ** t) + eps
    param.addcdiv_(exp_avg, denom, value=-step_size)
    return t
