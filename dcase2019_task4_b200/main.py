"""Mean-teacher training step of the reference (``baseline/main.py:45-165``) on the B200-native kernels.

``train(train_loader, model, optimizer, epoch, ema_model=None, weak_mask=None, strong_mask=None)`` keeps the
reference signature and semantics (consistency ramp-up, weak / strong BCE on slice masks, 2x MSE consistency,
teacher in train mode on the noisy input, Adam, EMA with alpha = min(1 - 1/(step+1), 0.999), per-batch loss
assertions, meters).  One iteration is: ONE C-ABI call for teacher forward + student forward + losses + student
backward (``dcase_mt_fwd_bwd``), an optional NCCL all-reduce of the flat gradient slab (data parallel), and ONE
fused Adam + EMA launch (``dcase_adam_ema_step``) that updates ``torch.optim.Adam``'s own state tensors in place,
so ``optimizer.state_dict()`` (saved at main.py:339) stays valid.  The reference's >= 9 ``.item()`` host syncs per
batch become one asynchronous 32-byte copy whose assertion is checked one batch later.

``main_simple_CRNN.train`` (main_simple_CRNN.py:31-82) is the same body with ``ema_model=None``.
"""
import ctypes
import os
import time

import numpy as np
import torch

from . import config as cfg
from . import dp
from . import kernels as K
from ._lib import MtArgs, StepScalars
from .utils import ramps
from .utils.utils import AverageMeterSet

METER_NAMES = ["weak_class_loss", "Weak EMA loss", "Strong loss", "Strong EMA loss", "Consistency strong",
               "Consistency weak", "Loss", "Consistency weight"]


def update_ema_variables(model, ema_model, alpha, global_step):
    """main.py:45-49 as one launch over the flat slabs (kept for API parity; the fused step does this itself)."""
    alpha = min(1 - 1 / (global_step + 1), alpha)
    with torch.no_grad():
        ema_model.flat_parameters().mul_(alpha).add_(model.flat_parameters(), alpha=1 - alpha)


def _bounds(mask, B):
    if mask is None:
        return 0, 0
    lo, hi, st = mask.indices(B)
    assert st == 1
    return lo, max(lo, hi)


# mirror of dcase_step_scalars (include/dcase_b200.h)
_SC_RING = 8          # pinned slots for the per-step scalars (see MeanTeacherEngine._upload_scalars)
_GRAPH_CACHE = 16     # captured graphs kept per engine
_SCALARS = np.dtype([("seed", "<u8"), ("step", "<u4"), ("cons_weight", "<f4"), ("ema_alpha", "<f4"), ("lr", "<f4"),
                     ("bc1", "<f4"), ("bc2", "<f4"), ("grad_scale", "<f4"), ("pad", "<f4")])


def bind_flat_adam_state(opt, plist, slices, n, device):
    """Make ``torch.optim.Adam``'s per-parameter ``exp_avg`` / ``exp_avg_sq`` views into two flat slabs (what
    dcase_adam_ema_step updates), so ``optimizer.state_dict()`` (saved at main.py:339) stays valid.  The slabs belong to
    the OPTIMIZER, not to one engine: engines for other batch shapes (a smaller last batch, main_simple_CRNN.py:190)
    keep updating the same moments.  Existing state (a resumed checkpoint) is copied in.  -> (m, v, [step tensors])."""
    if type(opt) is not torch.optim.Adam or len(opt.param_groups) != 1:
        raise NotImplementedError("the fused step implements torch.optim.Adam with one param group (main.py:290)")
    g = opt.param_groups[0]
    if g.get("amsgrad") or g.get("weight_decay", 0) != 0 or g.get("maximize"):
        raise NotImplementedError("amsgrad / weight_decay / maximize are not on the reference's path")
    if [id(p) for p in g["params"]] != [id(p) for p in plist]:
        raise NotImplementedError("the optimizer must hold exactly model.parameters() in order (main.py:290)")
    shared = opt.__dict__.get("_dcase_flat_state")
    bound = shared is not None and shared[0].numel() == n and shared[0].device == device and all(
        "exp_avg" in opt.state[p] and opt.state[p]["exp_avg"].data_ptr() == shared[0][off:off + 1].data_ptr()
        and opt.state[p]["exp_avg_sq"].data_ptr() == shared[1][off:off + 1].data_ptr()
        for p, (off, cnt) in zip(plist, slices))
    if bound:
        m, v = shared
    else:
        m = torch.zeros(n, device=device)
        v = torch.zeros(n, device=device)
        opt.__dict__["_dcase_flat_state"] = (m, v)
    steps = []
    for p, (off, cnt) in zip(plist, slices):
        st = opt.state[p]
        if not bound:
            if "exp_avg" in st:
                m[off:off + cnt].copy_(st["exp_avg"].reshape(-1))
                v[off:off + cnt].copy_(st["exp_avg_sq"].reshape(-1))
            st["exp_avg"] = m[off:off + cnt].view(p.shape)
            st["exp_avg_sq"] = v[off:off + cnt].view(p.shape)
        if "step" not in st:
            st["step"] = torch.tensor(0.0)
        steps.append(st["step"])
    return m, v, steps


class MeanTeacherEngine(object):
    """Device-resident buffers + launch sequence of one mean-teacher iteration for a fixed batch shape."""

    def __init__(self, model, optimizer, ema_model=None, weak_mask=None, strong_mask=None, batch_size=None,
                 frames=cfg.max_frames, process_group=None, use_graph=None):
        self.model, self.ema_model, self.optimizer = model, ema_model, optimizer
        self.B, self.T, self.NC = batch_size, frames, model.nclass
        self.To = frames // cfg.pooling_time_ratio
        self.weak_mask, self.strong_mask = weak_mask, strong_mask
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        dev = model.flat_parameters().device
        if dev.type != "cuda":
            raise RuntimeError("MeanTeacherEngine needs the models on a CUDA device (no CPU fallback)")
        self.dev = dev
        f32 = dict(device=dev, dtype=torch.float32)
        B, To, NC = self.B, self.To, self.NC
        self.strong_s = torch.empty(B, To, NC, **f32)
        self.weak_s = torch.empty(B, NC, **f32)
        self.strong_t = torch.empty(B, To, NC, **f32)
        self.weak_t = torch.empty(B, NC, **f32)
        self.d_strong = torch.empty(B, To, NC, **f32)
        self.d_weak = torch.empty(B, NC, **f32)
        self.meters = torch.zeros(8, **f32)
        # the 32-byte meter read-back is double buffered: the assertions on step i (main.py:147-148) are made AFTER step
        # i + 1 has been enqueued, so the host never leaves the GPU idle between two iterations
        self.meters_host = [torch.zeros(8, dtype=torch.float32).pin_memory() for _ in range(2)]
        self.meters_event = [torch.cuda.Event(), torch.cuda.Event()]
        self._slot = 0
        self._pending = False
        self.ws_s = K.new_workspace(B, frames, NC, dev)
        self.ws_t = K.new_workspace(B, frames, NC, dev) if ema_model is not None else None
        n = model.flat_parameters().numel()
        self.grads = torch.zeros(n, **f32)
        # data parallel: the gradient exchange is FUSED with Adam + EMA into one kernel over NVLink peer memory
        # (csrc/p2p.cu; the backward writes into the library-owned, IPC-exported slab).  DCASE_DP_NCCL=1 selects the
        # plain NCCL all-reduce + optimizer kernel instead (eager launches only: capturing NCCL into the step's graph
        # hung on hardware in rounds 1 and 2 and is not offered)
        self.p2p = None
        if self.world > 1 and os.environ.get("DCASE_DP_NCCL", "0") != "1":
            self.p2p = dp.P2PGradExchange(n, process_group)
            self.grads = self.p2p.grads
        self._bind_adam_state(n)
        self._x = torch.empty(B, frames, 64, **f32)
        self._x_ema = torch.empty(B, frames, 64, **f32) if ema_model is not None else None
        # CUDA-graph replay of step_from_waveforms (single GPU): the ~60 launches of one iteration are captured once
        # per (waveform, target, scaler) buffer set; everything that changes from step to step (Philox seed / step,
        # consistency weight, EMA alpha, lr, Adam bias corrections) lives in a 40-byte device struct the kernels read.
        graph_ok = self.world == 1 or self.p2p is not None
        self.use_graph = (graph_ok and os.environ.get("DCASE_NO_GRAPH", "0") != "1") if use_graph is None else bool(use_graph)
        if self.use_graph and not graph_ok:
            raise NotImplementedError("graph replay under data parallelism needs the fused p2p exchange (the NCCL "
                                      "all-reduce is launched eagerly)")
        K.ctx(dev)                                 # the library context must exist before any stream capture starts
        self._graphs = {}
        self.graph_launches = 0            # kernels launched through graph replays (bench.py's gpu_launches)
        assert K.lib().dcase_sizeof_step_scalars() == _SCALARS.itemsize
        # the per-step scalars go host -> device by an ASYNC copy from pinned memory; graph replay costs the host tens of
        # microseconds per ~ms step, so the host runs ahead: each pinned slot carries an event recorded after its copy
        # and is only rewritten once that copy has been consumed (ring of _SC_RING slots)
        self._sc_ring = [torch.zeros(_SCALARS.itemsize, dtype=torch.uint8).pin_memory() for _ in range(_SC_RING)]
        self._sc_event = [None] * _SC_RING
        self._sc_next = 0
        self._sc_dev = torch.zeros(_SCALARS.itemsize, dtype=torch.uint8, device=dev)

    def _upload_scalars(self, seed, step, cons_weight, global_step_after):
        """Fill the next pinned ring slot with this iteration's dcase_step_scalars and enqueue its copy to the device."""
        g = self.optimizer.param_groups[0]
        t = self._adam_step_count() + 1
        sc = np.zeros(1, dtype=_SCALARS)
        sc["seed"], sc["step"], sc["cons_weight"] = seed, step, cons_weight
        sc["ema_alpha"] = min(1 - 1 / (global_step_after + 1), 0.999)
        sc["lr"] = g["lr"]
        sc["bc1"], sc["bc2"] = 1.0 - g["betas"][0] ** t, 1.0 - g["betas"][1] ** t
        sc["grad_scale"] = 1.0 / self.world
        k = self._sc_next
        self._sc_next = (k + 1) % _SC_RING
        if self._sc_event[k] is not None:
            self._sc_event[k].synchronize()        # the DMA that read this slot _SC_RING steps ago has finished
        else:
            self._sc_event[k] = torch.cuda.Event()
        self._sc_ring[k].copy_(torch.from_numpy(sc.view(np.uint8)))
        self._sc_dev.copy_(self._sc_ring[k], non_blocking=True)
        self._sc_event[k].record()

    def _graph_key_common(self):
        """Every address a captured graph bakes in besides its inputs: the parameter / BN / Adam slabs of both models
        (CRNN._flatten may re-allocate them on a device or dtype change)."""
        model, ema = self.model, self.ema_model
        g = self.optimizer.param_groups[0]
        return (model.flat_parameters().data_ptr(), model.flat_bn_running().data_ptr(),
                ema.flat_parameters().data_ptr() if ema is not None else 0,
                ema.flat_bn_running().data_ptr() if ema is not None else 0,
                self.m.data_ptr(), self.v.data_ptr(), model.forward_flags(), g["betas"], g["eps"])

    @staticmethod
    def _cache_put(cache, key, entry):
        """Bounded graph cache (least recently used out): callers that pass fresh buffers every step must not grow it."""
        cache[key] = entry
        while len(cache) > _GRAPH_CACHE:
            cache.pop(next(iter(cache)))

    @staticmethod
    def _cache_get(cache, key):
        entry = cache.pop(key, None)
        if entry is not None:
            cache[key] = entry                     # re-insert: dicts keep insertion order, so the front is the LRU entry
        return entry

    # torch.optim.Adam bookkeeping -----------------------------------------------------------------------
    def _bind_adam_state(self, n):
        self.m, self.v, self._steps = bind_flat_adam_state(self.optimizer, self.model._param_list,
                                                            self.model._param_slices, n, self.dev)

    def _adam_step_count(self):
        return int(self._steps[0].item()) if self._steps else 0

    # one iteration ------------------------------------------------------------------------------------------
    def step(self, batch_input, ema_batch_input, target, cons_weight, global_step_after, check=True):
        """main.py:84-157 for one batch already on the device.  Returns the meters of the PREVIOUS step (or None); meters are read asynchronously
        (``read_meters``).  ``global_step_after`` is the reference's ``global_step`` after its increment."""
        model, ema = self.model, self.ema_model
        x = batch_input.contiguous()
        xt = ema_batch_input.contiguous() if ema is not None else None
        assert x.shape[0] == self.B and x.shape[-2] == self.T
        flags = model.forward_flags()
        seed, step = model.next_rng()
        a = self._mt_args(x, xt, target.contiguous(), flags, seed, step, cons_weight, None)
        if ema is not None:
            ema._nbt_pending += 1
        model._nbt_pending += 1
        prev_pending, prev_slot = self._pending, self._slot
        with torch.cuda.device(self.dev):
            g = self.optimizer.param_groups[0]
            alpha = min(1 - 1 / (global_step_after + 1), 0.999)
            if self.p2p is not None:
                self.p2p.begin_step()                      # nobody still reads this rank's slab of the previous step
                K.mt_fwd_bwd(a)
                torch._foreach_add_(self._steps, 1.0)
                self.p2p.adam_ema_step(model.flat_parameters(), self.m, self.v,
                                       ema.flat_parameters() if ema is not None else None, self._adam_step_count(),
                                       g["lr"], g["betas"][0], g["betas"][1], g["eps"], alpha)
            else:
                K.mt_fwd_bwd(a)
                grad_scale = dp.allreduce_grads_(self.grads, self.pg) if self.world > 1 else 1.0   # flat slab, SUM
                torch._foreach_add_(self._steps, 1.0)
                K.adam_ema_step(model.flat_parameters(), self.grads, self.m, self.v,
                                ema.flat_parameters() if ema is not None else None, self._adam_step_count(),
                                lr=g["lr"], beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"], ema_alpha=alpha,
                                grad_scale=grad_scale)
            self._enqueue_meter_copy()
        self._pending = True
        if check and prev_pending:
            return self._check_slot(prev_slot)
        return None

    def _mt_args(self, x, xt, target, flags, seed, step, cons_weight, scalars):
        model, ema = self.model, self.ema_model
        wl, wh = _bounds(self.weak_mask, self.B)
        sl, sh = _bounds(self.strong_mask, self.B)
        a = MtArgs()
        a.x_student = x.data_ptr()
        a.x_teacher = xt.data_ptr() if xt is not None else None
        a.target = target.data_ptr()
        a.B, a.T, a.n_class = self.B, self.T, self.NC
        a.weak_lo, a.weak_hi, a.strong_lo, a.strong_hi = wl, wh, sl, sh
        a.params_s = model.flat_parameters().data_ptr()
        a.bn_s = model.flat_bn_running().data_ptr()
        if ema is not None:
            a.params_t = ema.flat_parameters().data_ptr()
            a.bn_t = ema.flat_bn_running().data_ptr()
            a.strong_t, a.weak_t = self.strong_t.data_ptr(), self.weak_t.data_ptr()
            a.ws_t = self.ws_t.data_ptr()
        a.flags, a.seed, a.step, a.cons_weight, a.scalars = flags, seed, step, float(cons_weight), scalars
        a.strong_s, a.weak_s = self.strong_s.data_ptr(), self.weak_s.data_ptr()
        a.meters, a.d_strong, a.d_weak = self.meters.data_ptr(), self.d_strong.data_ptr(), self.d_weak.data_ptr()
        a.ws_s, a.grads = self.ws_s.data_ptr(), self.grads.data_ptr()
        a.after_forward_event = getattr(self, "_fork_handle", None)      # step_pipelined only; NULL otherwise
        moms = getattr(self, "_step_moms", None)                          # step_pipelined only: block-0 input moments of the
        if moms is not None:                                              # current slot, computed beside the previous step
            a.mom_s = moms[0].data_ptr()
            a.mom_t = moms[1].data_ptr() if ema is not None else None
        return a

    def _graph_step(self, wave, target, mean, std, cons_weight, global_step_after, check):
        model, ema = self.model, self.ema_model
        prev_pending, prev_slot = self._pending, self._slot
        target = target.contiguous()
        g = self.optimizer.param_groups[0]
        seed, step = model.next_rng()
        with torch.cuda.device(self.dev):
            self._upload_scalars(seed, step, cons_weight, global_step_after)
            key = (wave.data_ptr(), tuple(wave.shape), wave.dtype, target.data_ptr(), mean.data_ptr(), std.data_ptr()
                   ) + self._graph_key_common()
            entry = self._cache_get(self._graphs, key)
            if entry is None:
                graph = torch.cuda.CUDAGraph()
                l0 = K.launch_count()
                # thread-local capture under DP: NCCL's watchdog thread queries events of its own while we capture
                with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.world > 1 else "global"):
                    scp = self._sc_dev.data_ptr()
                    amp = K.logmel_fwd(wave)
                    if ema is not None:
                        x, x_ema = K.logmel_finish(amp, mean, std, self.T, noisy=True, scalars=self._sc_dev,
                                                   out_clean=self._x, out_noisy=self._x_ema)
                    else:
                        x, x_ema = K.logmel_finish(amp, mean, std, self.T, out_clean=self._x), None
                    if self.p2p is not None:               # nobody still reads this rank's slab of the previous step
                        self.p2p.begin_step()
                    K.mt_fwd_bwd(self._mt_args(x, x_ema, target, model.forward_flags(), 0, 0, 0.0, scp))
                    if self.p2p is not None:               # exchange + Adam + EMA as one kernel over NVLink peer memory
                        self.p2p.adam_ema_step(model.flat_parameters(), self.m, self.v,
                                               ema.flat_parameters() if ema is not None else None, 0, g["lr"],
                                               g["betas"][0], g["betas"][1], g["eps"], 0.0, scalars=self._sc_dev)
                    else:
                        K.adam_ema_step(model.flat_parameters(), self.grads, self.m, self.v,
                                        ema.flat_parameters() if ema is not None else None, 0, lr=g["lr"],
                                        beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"], scalars=self._sc_dev)
                entry = (graph, K.launch_count() - l0)
                self._cache_put(self._graphs, key, entry)
            entry[0].replay()
            self.graph_launches += entry[1]
            self._enqueue_meter_copy()
        torch._foreach_add_(self._steps, 1.0)
        model._nbt_pending += 1
        if ema is not None:
            ema._nbt_pending += 1
        if check and prev_pending:
            self._check_slot(prev_slot)
        self._pending = True

    def step_from_waveforms(self, wave, target, mean, std, cons_weight, global_step_after, check=True):
        """The whole hot path for one batch of raw clips already on the device: wave [B, L] (float32 or int16
        PCM) -> calculate_mel_spec -> noise / dB / pad / z-score (clean + noisy) -> mean-teacher iteration."""
        if self.use_graph:
            return self._graph_step(wave, target, mean, std, cons_weight, global_step_after, check)
        amp = K.logmel_fwd(wave)
        seed, _ = self.model._rng_seed, 0
        if self.ema_model is not None:
            # same (seed, step) as the dropout of this iteration (next_rng in step()); the noise has its own Philox
            # stream id, and the graph path reads the very same pair from the device scalars
            x, x_ema = K.logmel_finish(amp, mean, std, self.T, noisy=True, seed=seed,
                                       step=self.model._rng_step & 0xFFFFFFFF, out_clean=self._x, out_noisy=self._x_ema)
        else:
            x, x_ema = K.logmel_finish(amp, mean, std, self.T, out_clean=self._x), None
        self.step(x, x_ema, target, cons_weight, global_step_after, check=check)

    # ---- features one step ahead -------------------------------------------------------------------------------
    # The log-mel phase of a batch does not depend on the model, and the backward spends ~260 us in kernels that occupy
    # 24-48 CTAs (GRU BPTT, head, loss) while the STFT launches thousands: ``step_pipelined`` runs the iteration on the
    # features already sitting in the current slot and, on a side stream, prepares the features of the NEXT batch into
    # the other slot.  Same kernels, same Philox (seed, step) per batch as ``step_from_waveforms``; only the order of
    # launches differs (tests/test_gpu_api.py::test_pipelined_features_match_plain_steps).
    def _feature_slots(self):
        if getattr(self, "_xp", None) is None:
            f32 = dict(device=self.dev, dtype=torch.float32)
            self._xp = [torch.empty(self.B, self.T, 64, **f32) for _ in range(2)]
            self._xp_ema = ([torch.empty(self.B, self.T, 64, **f32) for _ in range(2)]
                            if self.ema_model is not None else [None, None])
            # block 0's BatchNorm batch statistics follow from 54 moments of its INPUT (include/dcase_b200.h): they are
            # taken right behind the features, beside the previous iteration, instead of at the head of the forward chain
            self._moms = [torch.zeros(2, 56, dtype=torch.float64, device=self.dev) for _ in range(2)]
            self._feat_slot = 0
            self._side = torch.cuda.Stream(self.dev)
            self._fwd_done = torch.cuda.Event()
            self._fwd_done.record()                    # creates the CUDA event; dcase_mt_fwd_bwd re-records it by handle
            self._pgraphs = {}
        return self._xp, self._xp_ema

    def _finish_into(self, amp, mean, std, slot, seed=0, step=0, scalars=None):
        xp, xpe = self._feature_slots()
        if self.ema_model is not None:
            K.logmel_finish(amp, mean, std, self.T, noisy=True, seed=seed, step=step, scalars=scalars,
                            out_clean=xp[slot], out_noisy=xpe[slot])
            K.cnn0_input_moments(xpe[slot], out=self._moms[slot][1])
        else:
            K.logmel_finish(amp, mean, std, self.T, out_clean=xp[slot])
        K.cnn0_input_moments(xp[slot], out=self._moms[slot][0])

    def prime_features(self, wave, mean, std):
        """Features of the FIRST batch of a pipelined run (current slot), with the upcoming iteration's noise."""
        self._feature_slots()
        self._finish_into(K.logmel_fwd(wave), mean, std, self._feat_slot, seed=self.model._rng_seed,
                          step=self.model._rng_step & 0xFFFFFFFF)

    def step_pipelined(self, wave_next, target, mean, std, cons_weight, global_step_after, check=True,
                       wave_ready_event=None):
        """One iteration on the current slot's features (``target`` belongs to THAT batch) while the features of
        ``wave_next`` -- the batch of the next call -- are prepared on a side stream.  ``wave_ready_event``: an event
        the side stream waits for before reading ``wave_next`` (a host-to-device copy still in flight)."""
        xp, xpe = self._feature_slots()
        slot, nxt = self._feat_slot, self._feat_slot ^ 1
        main = torch.cuda.current_stream(self.dev)
        if not self.use_graph:
            next_noise_step = (self.model._rng_step + 1) & 0xFFFFFFFF
            self._fork_handle = self._fwd_done.cuda_event      # recorded by dcase_mt_fwd_bwd after the student forward
            self._step_moms = self._moms[slot]
            try:
                self.step(xp[slot], xpe[slot], target, cons_weight, global_step_after, check=check)
            finally:
                self._fork_handle = None
                self._step_moms = None
            # the side stream starts once this step's forward is done (which also means the previous step -- the last
            # reader of the other slot -- is done) and runs beside the backward
            self._side.wait_event(self._fwd_done)
            if wave_ready_event is not None:
                self._side.wait_event(wave_ready_event)
            with torch.cuda.stream(self._side):
                amp = K.logmel_fwd(wave_next)
                amp.record_stream(self._side)
                self._finish_into(amp, mean, std, nxt, seed=self.model._rng_seed, step=next_noise_step)
            main.wait_stream(self._side)
            self._feat_slot = nxt
            return
        # graph replay: one graph per (next waveform buffer, target buffer, slot)
        model, ema = self.model, self.ema_model
        prev_pending, prev_slot = self._pending, self._slot
        target = target.contiguous()
        g = self.optimizer.param_groups[0]
        seed, step = model.next_rng()
        with torch.cuda.device(self.dev):
            self._upload_scalars(seed, step, cons_weight, global_step_after)
            if wave_ready_event is not None:
                main.wait_event(wave_ready_event)
            key = (wave_next.data_ptr(), tuple(wave_next.shape), wave_next.dtype, target.data_ptr(), mean.data_ptr(),
                   std.data_ptr(), slot, xp[0].data_ptr()) + self._graph_key_common()
            entry = self._cache_get(self._pgraphs, key)
            if entry is None:
                graph = torch.cuda.CUDAGraph()
                l0 = K.launch_count()
                with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.world > 1 else "global"):
                    cap = torch.cuda.current_stream(self.dev)
                    if self.p2p is not None:
                        self.p2p.begin_step()
                    self._fork_handle = self._fwd_done.cuda_event
                    self._step_moms = self._moms[slot]
                    try:
                        K.mt_fwd_bwd(self._mt_args(xp[slot], xpe[slot], target, model.forward_flags(), 0, 0, 0.0,
                                                   self._sc_dev.data_ptr()))
                    finally:
                        self._fork_handle = None
                        self._step_moms = None
                    self._side.wait_event(self._fwd_done)          # graph edge from the end of the student forward
                    with torch.cuda.stream(self._side):
                        amp = K.logmel_fwd(wave_next)
                        self._finish_into(amp, mean, std, nxt, step=1, scalars=self._sc_dev)   # next iteration's noise
                    if self.p2p is not None:
                        self.p2p.adam_ema_step(model.flat_parameters(), self.m, self.v,
                                               ema.flat_parameters() if ema is not None else None, 0, g["lr"],
                                               g["betas"][0], g["betas"][1], g["eps"], 0.0, scalars=self._sc_dev)
                    else:
                        K.adam_ema_step(model.flat_parameters(), self.grads, self.m, self.v,
                                        ema.flat_parameters() if ema is not None else None, 0, lr=g["lr"],
                                        beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"], scalars=self._sc_dev)
                    cap.wait_stream(self._side)
                entry = (graph, K.launch_count() - l0)
                self._cache_put(self._pgraphs, key, entry)
            entry[0].replay()
            self.graph_launches += entry[1]
            self._enqueue_meter_copy()
        torch._foreach_add_(self._steps, 1.0)
        model._nbt_pending += 1
        if ema is not None:
            ema._nbt_pending += 1
        if check and prev_pending:
            self._check_slot(prev_slot)
        self._pending = True
        self._feat_slot = nxt

    def _enqueue_meter_copy(self):
        self._slot ^= 1
        self.meters_host[self._slot].copy_(self.meters, non_blocking=True)
        self.meters_event[self._slot].record()

    def _read_slot(self, slot):
        self.meters_event[slot].synchronize()
        return dict(zip(METER_NAMES, self.meters_host[slot].tolist()))

    @staticmethod
    def _assert_loss(loss):
        assert not (np.isnan(loss) or loss > 1e5), 'Loss explosion: {}'.format(loss)
        assert not loss < 0, 'Loss problem, cannot be negative'

    def _check_slot(self, slot):
        vals = self._read_slot(slot)
        self._assert_loss(vals["Loss"])
        return vals

    def drain(self):
        """Meters of the last enqueued step (asserted like every other batch), or None if nothing is pending."""
        if not self._pending:
            return None
        self._pending = False
        return self._check_slot(self._slot)

    def read_meters(self):
        """Synchronise on the last enqueued step's 32-byte meter copy and return {name: float}."""
        return self._read_slot(self._slot)

    def check_loss(self):
        """The reference's per-batch assertions (main.py:147-148), applied to the last enqueued step."""
        loss = self.read_meters()["Loss"]
        self._pending = False
        self._assert_loss(loss)
        return loss


class HostBatchPrefetcher(object):
    """Double-buffered host -> device staging of (waveform, target) batches on a copy stream, so the H2D copy of
    batch i+1 overlaps the kernels of batch i.  Host tensors should be pinned (``tensor.pin_memory()``)."""

    def __init__(self, device, wave_shape, target_shape, wave_dtype=torch.float32, slots=2):
        self.dev = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.wave = [torch.empty(wave_shape, dtype=wave_dtype, device=self.dev) for _ in range(slots)]
        self.target = [torch.empty(target_shape, dtype=torch.float32, device=self.dev) for _ in range(slots)]
        self.copied = [torch.cuda.Event() for _ in range(slots)]
        self.consumed = [torch.cuda.Event() for _ in range(slots)]
        self.slots, self.head, self.tail, self.used = slots, 0, 0, [False] * slots

    def submit(self, wave_host, target_host):
        """Enqueue the copy of one host batch into the next free slot (waits for that slot's last consumer)."""
        k = self.head % self.slots
        self.head += 1
        with torch.cuda.stream(self.copy_stream):
            if self.used[k]:
                self.copy_stream.wait_event(self.consumed[k])
            self.wave[k].copy_(wave_host, non_blocking=True)
            self.target[k].copy_(target_host, non_blocking=True)
            self.copied[k].record(self.copy_stream)
        self.used[k] = True

    def next(self):
        """Device tensors of the oldest submitted batch; the current stream waits for its copy."""
        k = self.tail % self.slots
        torch.cuda.current_stream(self.dev).wait_event(self.copied[k])
        return self.wave[k], self.target[k]

    def peek(self, offset=1):
        """(wave, target, copied-event) of the batch ``offset`` places behind the one ``next`` returns, WITHOUT making
        any stream wait: the caller hands the event to the stream that reads the buffers (step_pipelined's side
        stream reads the NEXT batch's waveform while the current batch trains)."""
        k = (self.tail + offset) % self.slots
        assert self.tail + offset < self.head, "that batch has not been submitted yet"
        return self.wave[k], self.target[k], self.copied[k]

    def release(self):
        """Call after the kernels consuming the batch returned by ``next`` have been enqueued."""
        k = self.tail % self.slots
        self.tail += 1
        self.consumed[k].record(torch.cuda.current_stream(self.dev))


def _engine_for(model, optimizer, ema_model, weak_mask, strong_mask, B, T):
    key = (id(optimizer), id(ema_model), B, T, repr(weak_mask), repr(strong_mask))
    cache = model.__dict__.setdefault("_mt_engines", {})
    if key not in cache:
        cache[key] = MeanTeacherEngine(model, optimizer, ema_model, weak_mask, strong_mask, B, T)
    return cache[key]


def _update_meters(meters, vals, ema_model, weak_mask, strong_mask):
    """meters.update(...) of main.py:101-145 for one batch (``vals`` from the engine's 32-byte read-back, or None)."""
    if vals is None:
        return
    for name in METER_NAMES:
        if ema_model is None and ("EMA" in name or "Consistency" in name):
            continue
        if weak_mask is None and name in ("weak_class_loss", "Weak EMA loss"):
            continue
        if strong_mask is None and name in ("Strong loss", "Strong EMA loss"):
            continue
        meters.update(name, vals[name])


def train(train_loader, model, optimizer, epoch, ema_model=None, weak_mask=None, strong_mask=None, log=None):
    """One epoch of a Mean Teacher model (or of the plain CRNN when ``ema_model`` is None).

    train_loader yields (teacher input, student input... ) exactly as the reference: ``(batch_input,
    ema_batch_input, target)`` with a teacher, ``(batch_input, target)`` without."""
    meters = AverageMeterSet()
    start = time.time()
    rampup_length = len(train_loader) * cfg.n_epoch // 2
    engine = None
    for i, batch in enumerate(train_loader):
        if ema_model is not None:
            batch_input, ema_batch_input, target = batch
        else:
            (batch_input, target), ema_batch_input = batch, None
        global_step = epoch * len(train_loader) + i
        rampup_value = ramps.sigmoid_rampup(global_step, rampup_length) if global_step < rampup_length else 1.0
        meters.update('lr', optimizer.param_groups[0]['lr'])
        dev = model.flat_parameters().device
        batch_input = batch_input.to(dev, non_blocking=True)
        target = target.to(dev, non_blocking=True).float()
        if ema_batch_input is not None:
            ema_batch_input = ema_batch_input.to(dev, non_blocking=True)
        if engine is None or engine.B != batch_input.shape[0] or engine.T != batch_input.shape[-2]:
            if engine is not None:                     # a smaller last batch: settle the previous engine first
                _update_meters(meters, engine.drain(), ema_model, weak_mask, strong_mask)
            engine = _engine_for(model, optimizer, ema_model, weak_mask, strong_mask, batch_input.shape[0],
                                 batch_input.shape[-2])
        consistency_cost = cfg.max_consistency_cost * rampup_value
        # the reference updates every meter and asserts on the loss every batch (main.py:147-163); here the values of
        # batch i - 1 are read (and asserted) right after batch i has been enqueued, so the GPU never waits on the host
        prev = engine.step(batch_input, ema_batch_input, target, consistency_cost, global_step + 1)
        _update_meters(meters, prev, ema_model, weak_mask, strong_mask)
    if engine is not None:
        _update_meters(meters, engine.drain(), ema_model, weak_mask, strong_mask)
    epoch_time = time.time() - start
    msg = 'Epoch: {}\tTime {:.2f}\t{meters}'.format(epoch, epoch_time, meters=meters)
    (log.info if log is not None else print)(msg)
    return meters


# ---- checkpoint dict (main.py:293-309, :339-356; read back by TestModel.py:26-40) -----------------------------
def build_state(crnn, optimizer, crnn_kwargs, optim_kwargs, pooling_time_ratio, scaler, many_hot_encoder,
                crnn_ema=None):
    """The dict the reference ``torch.save``s: nested model state dicts (CRNN.py:49-53), torch's own optimizer state,
    the Scaler wire format and the encoder.  ``model_ema`` is absent for main_simple_CRNN (main_simple_CRNN.py:203-215)."""
    state = {'model': {"name": crnn.__class__.__name__, 'args': '', "kwargs": crnn_kwargs,
                       'state_dict': crnn.state_dict()}}
    if crnn_ema is not None:
        state['model_ema'] = {"name": crnn_ema.__class__.__name__, 'args': '', "kwargs": crnn_kwargs,
                              'state_dict': crnn_ema.state_dict()}
    state['optimizer'] = {"name": optimizer.__class__.__name__, 'args': '', "kwargs": optim_kwargs,
                          'state_dict': optimizer.state_dict()}
    state["pooling_time_ratio"] = pooling_time_ratio
    state["scaler"] = scaler.state_dict()
    state["many_hot_encoder"] = many_hot_encoder.state_dict()
    return state


def update_state(state, crnn, optimizer, epoch, valid_metric=None, crnn_ema=None):
    """End-of-epoch refresh, main.py:335-339."""
    state['model']['state_dict'] = crnn.state_dict()
    if crnn_ema is not None:
        state['model_ema']['state_dict'] = crnn_ema.state_dict()
    state['optimizer']['state_dict'] = optimizer.state_dict()
    state['epoch'] = epoch
    state['valid_metric'] = valid_metric
    return state


def _to_cpu(obj):
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


def save_checkpoint(state, model_fname):
    """``torch.save(state, model_fname)`` (main.py:342, :352) with every tensor detached from the flat device slabs,
    so the file loads on a machine without a GPU (TestModel.py:75 uses map_location="cpu")."""
    torch.save(_to_cpu(state), model_fname)


def load_checkpoint(model_fname, map_location="cpu"):
    """``torch.load`` of a reference-format checkpoint; the dict holds lists / floats / sed_eval results next to the
    tensors, which torch >= 2.6 refuses unless ``weights_only=False``."""
    return torch.load(model_fname, map_location=map_location, weights_only=False)


def restore_from_state(state):
    """TestModel.py:30-40: (crnn, scaler, many_hot_encoder, pooling_time_ratio) rebuilt from a checkpoint dict --
    ours or one written by the reference (same keys)."""
    from .models.CRNN import CRNN
    from .utils.Scaler import Scaler
    from .utils.utils import ManyHotEncoder
    crnn = CRNN(**state["model"]["kwargs"])
    crnn.load(parameters=state["model"]["state_dict"])
    scaler = Scaler()
    scaler.load_state_dict(state["scaler"])
    return crnn, scaler, ManyHotEncoder.load_state_dict(state["many_hot_encoder"]), state["pooling_time_ratio"]
