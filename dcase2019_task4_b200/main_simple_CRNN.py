"""``baseline/main_simple_CRNN.py``'s training step (:31-82) on the B200-native kernels.

``train(train_loader, model, optimizer, epoch, weak_mask=None, strong_mask=None)`` keeps the reference signature:
the loader yields ``(batch_input, target)``, the loss is the weak BCE on ``weak_mask`` plus the strong BCE on
``strong_mask`` (either may be None), followed by ``optimizer.zero_grad / loss.backward / optimizer.step``.  It is
the mean-teacher iteration of ``main.train`` without teacher and consistency terms, so one batch is the same two
C-ABI calls: ``dcase_mt_fwd_bwd`` with ``x_teacher == NULL`` (student forward, BCE losses, student backward) and
``dcase_adam_ema_step`` with ``p_ema == NULL`` (Adam only).  Meter names are the reference's (``lr``, ``Weak loss``,
``Strong loss``, ``Loss``); the loss assertions (:66-67) are made on every batch, one batch late (the 32-byte meter
read-back is asynchronous), and on the last batch before returning.
"""
import time

from . import config as cfg
from .main import _engine_for
from .utils.utils import AverageMeterSet


def masks_for(batch_size=cfg.batch_size, no_weak=False):
    """(weak_mask, strong_mask) exactly as main_simple_CRNN.py:186-192: half weak / half synthetic batches, or
    synthetic only."""
    if no_weak:
        return None, slice(batch_size)
    return slice(batch_size // 2), slice(batch_size // 2, batch_size)


def train(train_loader, model, optimizer, epoch, weak_mask=None, strong_mask=None, log=None):
    meters = AverageMeterSet()
    meters.update('lr', optimizer.param_groups[0]['lr'])
    start = time.time()
    engine = None
    dev = model.flat_parameters().device
    n = len(train_loader)
    for i, (batch_input, target) in enumerate(train_loader):
        batch_input = batch_input.to(dev, non_blocking=True)
        target = target.to(dev, non_blocking=True).float()
        if engine is None or engine.B != batch_input.shape[0] or engine.T != batch_input.shape[-2]:
            if engine is not None:
                engine.check_loss()
            engine = _engine_for(model, optimizer, None, weak_mask, strong_mask, batch_input.shape[0],
                                 batch_input.shape[-2])
        engine.step(batch_input, None, target, 0.0, epoch * n + i + 1)
    if engine is not None:
        vals = engine.read_meters()
        engine.check_loss()
        if weak_mask is not None:
            meters.update('Weak loss', vals["weak_class_loss"])
        if strong_mask is not None:
            meters.update('Strong loss', vals["Strong loss"])
        meters.update('Loss', vals["Loss"])
    epoch_time = time.time() - start
    msg = 'Epoch: {}\tTime {:.2f}\t{meters}'.format(epoch, epoch_time, meters=meters)
    (log.info if log is not None else print)(msg)
    return meters
