"""Plain-torch restatement of one mean-teacher batch of ``baseline/main.py`` (TEST ORACLE).

Follows:
* ``main.py:72-78,127``  consistency ramp-up ``2 * sigmoid_rampup(step, len(loader)*n_epoch//2)``
  (``utils/ramps.py:20-27``)
* ``main.py:87-91``      teacher forward on the noisy input (train mode, no grad), student on clean
* ``main.py:95-145``     weak BCE [weak_mask], strong BCE [strong_mask], 2x MSE consistency; teacher
                         BCEs for the meters only
* ``main.py:152-157``    zero_grad / backward / Adam(lr=1e-3, betas=(0.9, 0.999)) / EMA
* ``main.py:45-49``      ``update_ema_variables``: alpha = min(1 - 1/(g+1), 0.999) with g already
                         incremented; parameters only (BN running stats are NOT averaged)
* ``main_simple_CRNN.py:31-82``  is the same body without teacher and consistency terms
  (``teacher_p=None``).
"""
import math

import torch
import torch.nn.functional as F

from . import crnn as ocrnn

MAX_CONSISTENCY_COST = 2.0   # config.py:36
EMA_ALPHA = 0.999            # main.py:157
N_EPOCH = 100                # config.py:44


def sigmoid_rampup(current, rampup_length):
    """utils/ramps.py:20-27."""
    if rampup_length == 0:
        return 1.0
    current = min(max(float(current), 0.0), float(rampup_length))
    phase = 1.0 - current / rampup_length
    return float(math.exp(-5.0 * phase * phase))


def consistency_weight(global_step, steps_per_epoch, n_epoch=N_EPOCH):
    rampup_length = steps_per_epoch * n_epoch // 2
    if global_step < rampup_length:
        return MAX_CONSISTENCY_COST * sigmoid_rampup(global_step, rampup_length)
    return MAX_CONSISTENCY_COST * 1.0


def ema_alpha(global_step_after_increment, alpha=EMA_ALPHA):
    return min(1.0 - 1.0 / (global_step_after_increment + 1), alpha)


def new_adam_state(p):
    return {"step": 0,
            "exp_avg": {k: torch.zeros_like(v) for k, v in p.items()},
            "exp_avg_sq": {k: torch.zeros_like(v) for k, v in p.items()}}


def adam_update(p, grads, state, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.Adam (no amsgrad, no weight decay), single-tensor formulation."""
    state["step"] += 1
    t = state["step"]
    bc1 = 1.0 - beta1 ** t
    bc2 = 1.0 - beta2 ** t
    with torch.no_grad():
        for k in p:
            g = grads[k]
            m = state["exp_avg"][k]
            v = state["exp_avg_sq"][k]
            m.mul_(beta1).add_(g, alpha=1 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
            p[k].addcdiv_(m, denom, value=-lr / bc1)


def mean_teacher_losses(strong, weak, strong_t, weak_t, target, weak_mask, strong_mask, cons_w):
    """Loss terms of main.py:95-145.  Returns (loss, meters dict of python floats)."""
    meters = {}
    loss = None
    target_weak = target.max(-2)[0]
    if weak_mask is not None:
        wl = F.binary_cross_entropy(weak[weak_mask], target_weak[weak_mask])
        meters["weak_class_loss"] = wl.item()
        if weak_t is not None:
            meters["Weak EMA loss"] = F.binary_cross_entropy(weak_t[weak_mask], target_weak[weak_mask]).item()
        loss = wl
    if strong_mask is not None:
        sl = F.binary_cross_entropy(strong[strong_mask], target[strong_mask])
        meters["Strong loss"] = sl.item()
        if strong_t is not None:
            meters["Strong EMA loss"] = F.binary_cross_entropy(strong_t[strong_mask], target[strong_mask]).item()
        loss = sl if loss is None else loss + sl
    if strong_t is not None:
        cs = cons_w * F.mse_loss(strong, strong_t)
        cw = cons_w * F.mse_loss(weak, weak_t)
        meters["Consistency weight"] = cons_w
        meters["Consistency strong"] = cs.item()
        meters["Consistency weak"] = cw.item()
        loss = cs + cw if loss is None else loss + cs + cw
    meters["Loss"] = loss.item()
    return loss, meters


def train_batch(student_p, student_buf, adam_state, x, target, global_step, steps_per_epoch,
                teacher_p=None, teacher_buf=None, x_ema=None,
                weak_mask=None, strong_mask=None, masks_student=None, masks_teacher=None, lr=1e-3):
    """One iteration of the loop body main.py:73-157, mutating all state in place.

    Returns (meters, grads) with grads a {name: tensor} dict of the student gradients."""
    sp = {k: v.detach().requires_grad_(True) for k, v in student_p.items()}
    strong_t = weak_t = None
    if teacher_p is not None:
        with torch.no_grad():
            strong_t, weak_t = ocrnn.crnn_forward(x_ema, teacher_p, teacher_buf, True, masks_teacher)
    strong, weak = ocrnn.crnn_forward(x, sp, student_buf, True, masks_student)
    cons_w = consistency_weight(global_step, steps_per_epoch)
    loss, meters = mean_teacher_losses(strong, weak, strong_t, weak_t, target, weak_mask, strong_mask, cons_w)
    assert not (math.isnan(meters["Loss"]) or meters["Loss"] > 1e5), "Loss explosion"
    assert not meters["Loss"] < 0, "Loss problem, cannot be negative"
    # loss.backward() leaves .grad = None on parameters outside the graph (the attention head when weak_mask is None,
    # main_simple_CRNN.py --no_weak) and Adam skips them: a zero gradient gives the same (unchanged) value
    glist = torch.autograd.grad(loss, list(sp.values()), allow_unused=True)
    grads = {k: (torch.zeros_like(v) if g is None else g) for (k, v), g in zip(sp.items(), glist)}
    adam_update(student_p, grads, adam_state, lr=lr)
    if teacher_p is not None:
        a = ema_alpha(global_step + 1)
        with torch.no_grad():
            for k in teacher_p:
                teacher_p[k].mul_(a).add_(student_p[k], alpha=1 - a)
    meters["strong"] = strong.detach()
    meters["weak"] = weak.detach()
    return meters, grads
