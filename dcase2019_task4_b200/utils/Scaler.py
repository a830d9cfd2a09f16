"""``utils.Scaler.Scaler`` of the reference (baseline/utils/Scaler.py): per-mel-bin mean / std of the log-mel
features, ``normalize``, and the ``{"mean_", "mean_of_square_"}`` state-dict / JSON wire format that the
checkpoints store (main.py:306).

``calculate_scaler`` accepts what the reference passes (a dataset yielding ``(features, label)``) and reduces in
float64 exactly as Scaler.py:34-87 (mean over frames per sample, then mean over samples).  ``normalize`` on a
CUDA tensor goes through the fused finish kernel's arithmetic contract ((x - mean) / std per mel bin)."""
import json

import numpy as np
import torch


class Scaler(object):

    def __init__(self):
        self.mean_ = None
        self.mean_of_square_ = None
        self.std_ = None

    @staticmethod
    def _reduce_to_last_axis(a):
        a = np.asarray(a)
        while a.ndim != 1:
            a = np.mean(a, axis=0, dtype=np.float64)
        return a

    def means(self, dataset):
        total = None
        total_sq = None
        shape = None
        count = 0
        for sample in dataset:
            feats = sample[0] if isinstance(sample, (tuple, list)) and len(sample) == 2 else sample
            arr = feats.detach().cpu().numpy() if isinstance(feats, torch.Tensor) else np.asarray(feats)
            if shape is None:
                shape = arr.shape
            elif arr.shape != shape:
                raise NotImplementedError("Not possible to add data with different shape in mean calculation yet")
            m = self._reduce_to_last_axis(arr)
            m2 = self._reduce_to_last_axis(arr ** 2)
            total = m if total is None else total + m
            total_sq = m2 if total_sq is None else total_sq + m2
            count += 1
        self.mean_ = total / count
        self.mean_of_square_ = total_sq / count
        return self

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
