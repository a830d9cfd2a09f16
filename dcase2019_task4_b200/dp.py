"""Data-parallel plumbing (one process per GPU, torch.distributed): the reference has none.

Batches shard PER STREAM so every rank keeps the ``[weak | unlabeled | synthetic]`` order and the slice masks of
main.py:240-247 stay valid; equal per-rank sub-batch sizes make the mean of the per-rank mean losses equal the
global mean loss for all four loss terms.  The only exchange step is ONE all-reduce (SUM) of the flat 214,356-float
gradient slab; the 1/world_size scale is folded into the fused Adam + EMA kernel (``grad_scale``).  Adam and EMA run
replicated.  BatchNorm statistics stay per replica (the reference's own batch of 24 per device)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise the default process group from torchrun's environment (RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return 0, 1
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def per_rank_batch_sizes(global_batch_sizes, world_size):
    """[48, 144] over 8 ranks -> [6, 18]; every stream must divide evenly so the slice masks are rank-independent."""
    out = []
    for bs in global_batch_sizes:
        if bs % world_size:
            raise ValueError("stream batch size %d is not divisible by world size %d" % (bs, world_size))
        out.append(bs // world_size)
    return out


def allreduce_grads_(flat_grads, group=None):
    """SUM all-reduce of the flat gradient slab in place; returns the scale the optimizer must apply (1/N)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def broadcast_model_(model, src=0, group=None):
    """Identical student / teacher weights and BN statistics on every rank before step 0."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(model.flat_parameters(), src, group=group)
        dist.broadcast(model.flat_bn_running(), src, group=group)
