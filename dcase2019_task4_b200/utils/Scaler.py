"""``utils.Scaler.Scaler`` of the reference (baseline/utils/Scaler.py): per-mel-bin mean / std of the log-mel
features, ``normalize``, and the ``{"mean_", "mean_of_square_"}`` state-dict / JSON wire format that the
checkpoints store (main.py:306).

``means`` / ``calculate_scaler`` accept what the reference passes at main.py:249-250 (a dataset yielding
``(features, label)``) and keep Scaler.py:34-97's arithmetic (per-sample mean over every axis but the last in
float64, the square taken in the sample's float32, then the mean over samples), but the reduction runs on the GPU
(``dcase_scaler_accumulate`` / ``dcase_scaler_finalize``, csrc/logmel.cu):

* a ``DataLoadDf`` / ``ConcatDataset`` whose transform is the chain of ``get_transforms(frames)`` is read through
  ``get_feature_file_func`` (the cached amplitude mels), batched, and ApplyLog + PadOrTrunc + the reduction are ONE pass over
  each batch -- the log features are never written;
* ``means_from_waveforms`` starts from raw clips (``dcase_logmel_fwd`` first), for runs without a feature cache;
* any other iterable of ``[..., 64]`` samples is stacked and reduced as it is.

There is no CPU fallback: without the CUDA library / an sm_100 GPU ``means`` raises.  ``normalize`` on the training
path is part of the fused finish kernel ((x - mean) / std per mel bin); the method here serves host callers."""
import json

import numpy as np
import torch


def _fused_sources(dataset):
    """The DataLoadDf members of ``dataset`` when every one of them carries the plain chain of
    get_transforms(frames) (no noise, no scaler), else None."""
    from ..DataLoad import Compose, ConcatDataset, DataLoadDf
    members = dataset.datasets if isinstance(dataset, ConcatDataset) else [dataset]
    frames = set()
    for d in members:
        if not isinstance(d, DataLoadDf) or not isinstance(d.transform, Compose) or d.return_indexes:
            return None
        try:
            plan = d.transform._build_plan()
        except NotImplementedError:
            return None
        if plan["noise"] or plan["scaler"] is not None:
            return None
        frames.add(plan["frames"])
    if len(frames) != 1:
        return None
    return members, frames.pop()


class Scaler(object):

    # clips per device launch: 256 x 864 x 64 fp32 = 57 MB staged per batch, still L2-resident (126 MB) for the
    # reduction's second pass; the ~15 us host cost of a launch is amortised 4x better than at 64 clips
    # (profiles/r1_scaler_bench*.json)
    BATCH_CLIPS = 256

    def __init__(self):
        self.mean_ = None
        self.mean_of_square_ = None
        self.std_ = None

    # ---- device reduction --------------------------------------------------------------------------------
    @staticmethod
    def _device():
        from .. import _lib
        _lib.ctx()                       # raises without the library or an sm_100 GPU
        return torch.device("cuda", torch.cuda.current_device())

    def _finish(self, sums, count):
        from .. import kernels as K
        if count == 0:
            raise ValueError("Scaler.means over an empty dataset")
        mean, msq, _, _ = K.scaler_finalize(sums, count)
        self.mean_ = mean.cpu().numpy()
        self.mean_of_square_ = msq.cpu().numpy()
        return self

    def _reduce_batches(self, batches, frames, apply_log):
        """batches: iterable of float32 [B, T, 64] tensors (host or device)."""
        from .. import kernels as K
        dev = self._device()
        sums = torch.zeros(2, 64, dtype=torch.float64, device=dev)
        count = 0
        for b in batches:
            b = b if b.is_cuda else b.pin_memory().to(dev, non_blocking=True)
            K.scaler_accumulate(b, sums, frames=frames, apply_log=apply_log)
            count += b.shape[0]
        return self._finish(sums, count)

    def _amplitude_batches(self, members):
        group, shape = [], None
        for d in members:
            for i in range(len(d)):
                # features only: the labels (pandas work per sample in get_sample) play no part in the statistics
                f = np.ascontiguousarray(d.get_feature_file_func(d.filenames.iloc[i]), dtype=np.float32)
                if f.ndim != 2 or f.shape[1] != 64:
                    raise NotImplementedError("features are [T, 64] amplitude mels, got {}".format(f.shape))
                if group and (f.shape != shape or len(group) == self.BATCH_CLIPS):
                    yield torch.from_numpy(np.stack(group))
                    group = []
                shape = f.shape
                group.append(f)
        if group:
            yield torch.from_numpy(np.stack(group))

    def _finished_batches(self, dataset):
        group, shape = [], None
        for sample in dataset:
            feats = sample[0] if isinstance(sample, (tuple, list)) and len(sample) == 2 else sample
            t = feats.detach() if isinstance(feats, torch.Tensor) else torch.from_numpy(np.asarray(feats))
            if shape is None:
                shape = tuple(t.shape)
            elif tuple(t.shape) != shape:
                raise NotImplementedError("Not possible to add data with different shape in mean calculation yet")
            if t.dim() < 1 or t.shape[-1] != 64:
                raise NotImplementedError("the device reduction is built for 64 mel bins, got {}".format(shape))
            group.append(t.float().reshape(-1, 64))
            if len(group) == self.BATCH_CLIPS:
                yield torch.stack(group)
                group = []
        if group:
            yield torch.stack(group)

    def means(self, dataset):
        """Scaler.py:34-87: sets ``mean_`` and ``mean_of_square_`` (float64 numpy [64])."""
        fused = _fused_sources(dataset)
        if fused is not None:
            members, frames = fused
            return self._reduce_batches(self._amplitude_batches(members), frames, True)
        return self._reduce_batches(self._finished_batches(dataset), None, False)

    def means_from_waveforms(self, wave_batches, frames):
        """Same statistics from raw clips: each item is a [B, L] waveform batch (float or int16 PCM, host or
        device); calculate_mel_spec, ApplyLog, PadOrTrunc(frames) and the reduction all run on the GPU."""
        from .. import kernels as K
        dev = self._device()

        def amps():
            for w in wave_batches:
                w = torch.as_tensor(w)
                w = w if w.dtype == torch.int16 else w.float()
                yield K.logmel_fwd(w.to(dev, non_blocking=True))
        return self._reduce_batches(amps(), frames, True)

    def calculate_scaler_from_waveforms(self, wave_batches, frames):
        self.means_from_waveforms(wave_batches, frames)
        self.std_ = self.std(self.variance(self.mean_, self.mean_of_square_))
        return self.mean_, self.std_

    def variance(self, mean, mean_of_square):
        return mean_of_square - mean ** 2

    def std(self, variance):
        return np.sqrt(variance)

    def calculate_scaler(self, dataset):
        self.means(dataset)
        self.std_ = self.std(self.variance(self.mean_, self.mean_of_square_))
        return self.mean_, self.std_

    def normalize(self, batch):
        if isinstance(batch, torch.Tensor):
            mean = torch.as_tensor(self.mean_, dtype=torch.float32, device=batch.device)
            std = torch.as_tensor(self.std_, dtype=torch.float32, device=batch.device)
            return ((batch - mean) / std).float()
        return (batch - self.mean_) / self.std_

    def device_stats(self, device):
        """(mean, std) as float32 [64] device tensors for dcase_logmel_finish."""
        return (torch.as_tensor(self.mean_, dtype=torch.float32, device=device),
                torch.as_tensor(self.std_, dtype=torch.float32, device=device))

    def state_dict(self):
        if type(self.mean_) is not np.ndarray:
            raise NotImplementedError("Save scaler only implemented for numpy array means_")
        return {"mean_": self.mean_.tolist(), "mean_of_square_": self.mean_of_square_.tolist()}

    def save(self, path):
        with open(path, "w") as f:
            json.dump(self.state_dict(), f)

    def load(self, path):
        with open(path, "r") as f:
            self.load_state_dict(json.load(f))

    def load_state_dict(self, state_dict):
        self.mean_ = np.array(state_dict["mean_"])
        self.mean_of_square_ = np.array(state_dict["mean_of_square_"])
        self.std_ = self.std(self.variance(self.mean_, self.mean_of_square_))
