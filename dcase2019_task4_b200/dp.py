"""Data-parallel plumbing (one process per GPU, torch.distributed): the reference has none.

Batches shard PER STREAM so every rank keeps the ``[weak | unlabeled | synthetic]`` order and the slice masks of
main.py:240-247 stay valid; equal per-rank sub-batch sizes make the mean of the per-rank mean losses equal the
global mean loss for all four loss terms.  The only exchange step is ONE all-reduce (SUM) of the flat 214,356-float
gradient slab; the 1/world_size scale is folded into the fused Adam + EMA kernel (``grad_scale``).  Adam and EMA run
replicated.  BatchNorm statistics stay per replica by default (the reference's own batch of 24 per device);
``SyncBatchNorm`` switches the context to exact-global-batch statistics."""
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


class _RawCudaArray(object):
    """Minimal __cuda_array_interface__ carrier: lets torch alias memory the C library owns (no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class P2PGradExchange(object):
    """The gradient exchange fused with Adam + EMA over NVLink peer memory (csrc/p2p.cu; the default of
    ``MeanTeacherEngine`` for world_size > 1).  ``grads`` is a torch view of the library-owned, IPC-exported slab the
    backward writes into; ``begin_step`` goes before the backward, ``adam_ema_step`` replaces all-reduce + optimizer."""

    def __init__(self, n_floats, group=None):
        import ctypes
        from . import _lib
        self._lib = _lib
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = _lib.lib()
        blob = ctypes.create_string_buffer(L.dcase_p2p_handle_bytes())
        self.handle = ctypes.c_void_p()
        _lib.check(L.dcase_p2p_create(_lib.ctx(), self.world, self.rank, int(n_floats), ctypes.byref(self.handle), blob))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(blob.raw), group=group)
        self._all = ctypes.create_string_buffer(b"".join(gathered))
        _lib.check(L.dcase_p2p_connect(self.handle, self._all))
        dist.barrier(group)                                   # every rank has mapped every slab before anyone signals
        self._raw = _RawCudaArray(L.dcase_p2p_grads(self.handle), n_floats)
        self.grads = torch.as_tensor(self._raw, device=torch.device("cuda", torch.cuda.current_device()))

    def begin_step(self):
        self._lib.check(self._lib.lib().dcase_p2p_begin_step(self.handle, self._lib.stream_ptr()))

    def adam_ema_step(self, p, m, v, p_ema, step_t, lr, beta1, beta2, eps, ema_alpha, scalars=None):
        L, ptr = self._lib.lib(), self._lib.ptr
        self._lib.check(L.dcase_p2p_adam_ema_step(self._lib.ctx(), self.handle, ptr(p), ptr(m), ptr(v), ptr(p_ema),
                                                  lr, beta1, beta2, eps, int(step_t), float(ema_alpha), ptr(scalars),
                                                  self._lib.stream_ptr()))

    def close(self):
        if self.handle:
            self._lib.lib().dcase_p2p_destroy(self.handle)
            self.handle = None


class SyncBatchNorm(object):
    """Exact-global-batch BatchNorm statistics (the reference's models/CNN.py:49 at a global batch of N x 24 on ONE
    device): while an instance is attached, every train-mode forward / backward of this process's context sums the
    per-channel BatchNorm sums over the ranks (csrc/p2p.cu, one single-CTA peer-memory kernel per BatchNorm, replays
    inside the step's CUDA graph).  Every rank must then issue the same sequence of train-mode calls.  ``close()``
    returns to per-replica statistics (the default: the reference's semantics at its own batch of 24 per device)."""

    def __init__(self, group=None):
        import ctypes
        from . import _lib
        self._lib = _lib
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = _lib.lib()
        blob = ctypes.create_string_buffer(L.dcase_syncbn_handle_bytes())
        self.handle = ctypes.c_void_p()
        _lib.check(L.dcase_syncbn_create(_lib.ctx(), self.world, self.rank, ctypes.byref(self.handle), blob))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, bytes(blob.raw), group=group)
        self._all = ctypes.create_string_buffer(b"".join(gathered))
        _lib.check(L.dcase_syncbn_connect(self.handle, self._all))
        dist.barrier(group)                                   # every rank has mapped every mailbox before anyone pushes
        _lib.check(L.dcase_ctx_set_syncbn(_lib.ctx(), self.handle))

    def allreduce_(self, t, slot=15):
        """Sum a small float32 / float64 CUDA tensor over the group in place (tests; slots 0-8 belong to the CRNN path)."""
        self._lib.check(self._lib.lib().dcase_syncbn_allreduce(self.handle, self._lib.ptr(t), t.numel(),
                                                               1 if t.dtype == torch.float64 else 0, int(slot),
                                                               self._lib.stream_ptr()))
        return t

    def close(self):
        if self.handle:
            torch.cuda.synchronize()
            self._lib.check(self._lib.lib().dcase_ctx_set_syncbn(self._lib.ctx(), None))
            self._lib.lib().dcase_syncbn_destroy(self.handle)
            self.handle = None
