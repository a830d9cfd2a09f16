"""Host-side mirror of ``baseline/utils/utils.py`` (boundary objects that ``main.py`` calls around the replaced
hot path): ``ManyHotEncoder`` (:22-172, defines the [108,10] target layout), ``weights_init`` (:205-224),
``to_cuda_if_available`` (:227-239), ``SaveBest`` (:242-283), ``AverageMeterSet`` / ``AverageMeter`` (:337-394),
``create_folder`` (:196-202), ``get_transforms`` (:397-412), ``read_audio`` (:175-193, wav container parsed on the
host, mix-down on the GPU).  ``dcase_util`` is not installed, so the contiguous
region decode (:146-162) is restated with numpy."""
import os

import numpy as np
import pandas as pd
import torch
from torch import nn

from ..DataLoad import AugmentGaussianNoise, ApplyLog, PadOrTrunc, ToTensor, Normalize, Compose


def find_contiguous_regions(activity):
    """dcase_util.data.DecisionEncoder.find_contiguous_regions: [[onset, offset), ...] of a boolean vector."""
    a = np.asarray(activity).astype(bool)
    change = np.flatnonzero(a[1:] != a[:-1]) + 1
    if a.size and a[0]:
        change = np.r_[0, change]
    if a.size and a[-1]:
        change = np.r_[change, a.size]
    return change.reshape(-1, 2)


class ManyHotEncoder:
    """Labels <-> many-hot arrays; strong targets are [n_frames, n_classes], unlabeled clips are all -1."""

    def __init__(self, labels, n_frames=None):
        if isinstance(labels, np.ndarray):
            labels = labels.tolist()
        self.labels = labels
        self.n_frames = n_frames

    def encode_weak(self, labels):
        if isinstance(labels, str) and labels == "empty":
            return np.zeros(len(self.labels)) - 1
        if isinstance(labels, pd.DataFrame):
            labels = [] if labels.empty else (labels["event_label"] if "event_label" in labels.columns else labels)
        y = np.zeros(len(self.labels))
        for label in labels:
            if not pd.isna(label):
                y[self.labels.index(label)] = 1
        return y

    def _mark(self, y, label, onset, offset):
        if not pd.isna(label) and label != "":
            y[int(onset):int(offset), self.labels.index(label)] = 1   # offset excluded

    def encode_strong_df(self, label_df):
        assert self.n_frames is not None, "n_frames need to be specified when using strong encoder"
        if isinstance(label_df, str) and label_df == 'empty':
            return np.zeros((self.n_frames, len(self.labels))) - 1
        y = np.zeros((self.n_frames, len(self.labels)))
        cols = {"onset", "offset", "event_label"}
        if isinstance(label_df, pd.DataFrame):
            if cols.issubset(label_df.columns):
                for _, row in label_df.iterrows():
                    self._mark(y, row["event_label"], row["onset"], row["offset"])
        elif isinstance(label_df, pd.Series) and cols.issubset(label_df.index):
            self._mark(y, label_df["event_label"], label_df["onset"], label_df["offset"])
        elif isinstance(label_df, (pd.Series, list, np.ndarray)):
            for ev in label_df:
                if isinstance(ev, str):          # weak label: present on every frame
                    if ev != "":
                        y[:, self.labels.index(ev)] = 1
                elif len(ev) == 3:               # [label, onset, offset]
                    self._mark(y, ev[0], ev[1], ev[2])
                else:
                    raise NotImplementedError("cannot encode strong, type mismatch: {}".format(type(ev)))
        else:
            raise NotImplementedError("To encode_strong, type is pandas.Dataframe with onset, offset and event_label"
                                      "columns, or it is a list or pandas Series of event labels, "
                                      "type given: {}".format(type(label_df)))
        return y

    def decode_weak(self, labels):
        return [self.labels[i] for i, v in enumerate(labels) if v == 1]

    def decode_strong(self, labels):
        out = []
        for i, column in enumerate(np.asarray(labels).T):
            for onset, offset in find_contiguous_regions(column):
                out.append([self.labels[i], onset, offset])
        return out

    def state_dict(self):
        return {"labels": self.labels, "n_frames": self.n_frames}

    @classmethod
    def load_state_dict(cls, state_dict):
        return cls(state_dict["labels"], state_dict["n_frames"])


def read_wav_frames(path):
    """(interleaved frames [n_frames, n_channels] as int16 PCM or float32, sampling rate) of a RIFF wav file.
    soundfile is not installed; scipy.io.wavfile parses the container (PCM 8/16/24/32 bit, IEEE float).  16-bit PCM
    -- what the DCASE wavs are -- stays int16 (the device applies soundfile's 1 / 32768); other encodings are
    converted to soundfile's float range here."""
    import scipy.io.wavfile
    fs, data = scipy.io.wavfile.read(path)
    if data.ndim == 1:
        data = data[:, None]
    if data.dtype == np.int16:
        pass
    elif data.dtype == np.int32:
        data = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        data = ((data.astype(np.float32) - 128.0) / 128.0).astype(np.float32)
    else:
        data = data.astype(np.float32)
    return np.ascontiguousarray(data), int(fs)


def read_audio_device(path, target_fs=None):
    """read_audio (utils/utils.py:175-193) with the samples left on the GPU: (CUDA float32 mono [n], fs).
    Multi-channel files are mixed down by the mean (dcase_audio_mixdown); 16-bit PCM is scaled by 1 / 32768; a file at
    another rate is resampled like librosa.resample's kaiser_best (dcase_audio_resample)."""
    from .. import kernels as K
    frames, fs = read_wav_frames(path)
    if not torch.cuda.is_available():
        raise RuntimeError("read_audio feeds the GPU feature extraction (dcase_audio_mixdown); no CPU fallback")
    dev_frames = torch.from_numpy(frames).cuda(non_blocking=True)
    if frames.shape[0] == 0:
        return torch.empty(0, device=dev_frames.device), fs if target_fs is None else target_fs
    audio = K.audio_mixdown(dev_frames)
    if target_fs is not None and fs != target_fs:
        # librosa.resample(audio, orig_sr=fs, target_sr=target_fs) (utils/utils.py:190-192) with the 'kaiser_best' filter of
        # the librosa versions the baseline was written for, on the device (dcase_audio_resample)
        audio = K.audio_resample(audio, fs, target_fs)
        fs = target_fs
    return audio, fs


def read_audio(path, target_fs=None):
    """Reference signature: (numpy mono waveform, sampling rate)."""
    audio, fs = read_audio_device(path, target_fs)
    return audio.cpu().numpy(), fs


def create_folder(fd):
    if not os.path.exists(fd):
        os.makedirs(fd)


def weights_init(m):
    """Xavier-uniform(gain sqrt 2) convs, N(1, 0.02) BatchNorm, orthogonal GRU matrices, N(0, 0.01) Linear --
    selected by class-name substring exactly as the reference, so it works with ``model.apply``."""
    classname = m.__class__.__name__
    if classname.find('Conv2d') != -1:
        nn.init.xavier_uniform_(m.weight, gain=np.sqrt(2))
        m.bias.data.fill_(0)
    elif classname.find('BatchNorm') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
    elif classname.find('GRU') != -1:
        for weight in m.parameters():
            if len(weight.size()) > 1:
                nn.init.orthogonal_(weight.data)
    elif classname.find('Linear') != -1:
        m.weight.data.normal_(0, 0.01)
        m.bias.data.zero_()


def to_cuda_if_available(list_args):
    if torch.cuda.is_available():
        for i in range(len(list_args)):
            list_args[i] = list_args[i].cuda()
    return list_args


class SaveBest:
    """Tracks the best value of a metric ('inf': lower is better, 'sup': higher is better)."""

    def __init__(self, val_comp="inf"):
        if val_comp not in ("inf", "sup"):
            raise NotImplementedError("value comparison is only 'inf' or 'sup'")
        self.comp = val_comp
        self.best_val = np.inf if val_comp == "inf" else 0
        self.best_epoch = 0
        self.current_epoch = 0

    def apply(self, value):
        better = value < self.best_val if self.comp == "inf" else value > self.best_val
        decision = self.current_epoch == 0 or better
        if better:
            self.best_epoch = self.current_epoch
            self.best_val = value
        self.current_epoch += 1
        return decision


class AverageMeter:
    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count

    def __format__(self, format):
        return "{self.avg:{format}}".format(self=self, format=format)


class AverageMeterSet:
    def __init__(self):
        self.meters = {}

    def __getitem__(self, key):
        return self.meters[key]

    def update(self, name, value, n=1):
        self.meters.setdefault(name, AverageMeter()).update(value, n)

    def reset(self):
        for meter in self.meters.values():
            meter.reset()

    def values(self, postfix=''):
        return {name + postfix: meter.val for name, meter in self.meters.items()}

    def averages(self, postfix='/avg'):
        return {name + postfix: meter.avg for name, meter in self.meters.items()}

    def sums(self, postfix='/sum'):
        return {name + postfix: meter.sum for name, meter in self.meters.items()}

    def counts(self, postfix='/count'):
        return {name + postfix: meter.count for name, meter in self.meters.items()}

    def __str__(self):
        return "".join("{} {:{f}} \t".format(n, m.val, f=".2E" if m.val < 0.01 else ".4f")
                       for n, m in self.meters.items())


def get_transforms(frames, scaler=None, add_axis_conv=True, augment_type=None):
    """noise -> log -> pad/trunc -> tensor(+channel axis) -> normalise, as one fused GPU chain (DataLoad.Compose)."""
    transf = []
    if augment_type == "noise":
        transf.append(AugmentGaussianNoise(mean=0., std=0.5))
    transf.extend([ApplyLog(), PadOrTrunc(nb_frames=frames), ToTensor(unsqueeze_axis=0 if add_axis_conv else None)])
    if scaler is not None:
        transf.append(Normalize(scaler=scaler))
    return Compose(transf)
