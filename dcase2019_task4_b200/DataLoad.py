"""Host-side mirror of the reference's ``baseline/DataLoad.py``: dataset over a DataFrame, dataset concatenation
with ``cluster_indices``, the fixed-ratio multi-stream batch sampler, and the per-sample transform chain.

What stays host logic (same names, constructor signatures and behaviour): ``DataLoadDf`` (DataLoad.py:25-154),
``ConcatDataset`` (:383-439), ``MultiStreamBatchSampler`` (:539-577), ``grouper`` (:580-585).
What moves to the GPU: the arithmetic of the transform chain ``AugmentGaussianNoise`` (:262-287) -> ``ApplyLog``
(:189-207) -> ``PadOrTrunc`` (:231-259) -> ``ToTensor`` (:290-321) -> ``Normalize`` (:324-350).  The classes keep
their names and constructor arguments; ``Compose`` recognises the chain ``get_transforms`` builds
(utils/utils.py:397-412) and runs it as ONE fused kernel (dcase_logmel_finish), per sample here or per batch via
``Compose.transform_batch``.  ``Sampler.__init__`` no longer takes ``data_source`` (torch >= 2.2), so the
sampler does not forward it.  ClusterRandomSampler / Subset / random_split / GaussianNoise are unused
by both mains and are not rebuilt.
"""
import bisect
import warnings

import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset
from torch.utils.data.sampler import Sampler


class DataLoadDf(Dataset):
    """Dataset over a DataFrame with columns ``filename`` [+ ``event_labels`` | ``onset, offset, event_label``]."""

    def __init__(self, df, get_feature_file_func, encode_function, transform=None, return_indexes=False):
        self.df = df
        self.get_feature_file_func = get_feature_file_func
        self.encode_function = encode_function
        self.transform = transform
        self.return_indexes = return_indexes
        self.filenames = df.filename.drop_duplicates()
        self._strong = {"onset", "offset", "event_label"}.issubset(df.columns)
        self._weak = "event_labels" in df.columns
        self._groups = None

    def set_return_indexes(self, val):
        self.return_indexes = val

    def __len__(self):
        return len(self.filenames)

    def _strong_rows(self, filename):
        # the reference scans the frame per sample (DataLoad.py:100); one groupby gives the same rows
        if self._groups is None:
            self._groups = {k: v for k, v in self.df.groupby("filename", sort=False)}
        return self._groups[filename][["onset", "offset", "event_label"]]

    def get_sample(self, index):
        filename = self.filenames.iloc[index]
        features = self.get_feature_file_func(filename)
        if self._weak:
            label = self.df.iloc[index]["event_labels"]
            if isinstance(label, str):
                label = [] if label == "" else label.split(",")
            elif pd.isna(label):
                label = []
        elif self._strong:
            label = self._strong_rows(filename)
            if label.empty:
                label = []
        else:
            if "filename" not in self.df.columns:
                raise NotImplementedError(
                    "Dataframe to be encoded doesn't have specified columns: columns allowed: 'filename' for "
                    "unlabeled; 'filename', 'event_labels' for weak labels; 'filename' 'onset' 'offset' "
                    "'event_label' for strong labels, yours: {}".format(self.df.columns))
            label = "empty"   # -> all -1 targets for unlabeled clips (utils.py:82-85)
        y = self.encode_function(label) if self.encode_function is not None else label
        return features, y

    def __getitem__(self, index):
        sample = self.get_sample(index)
        if self.transform:
            sample = self.transform(sample)
        if self.return_indexes:
            sample = (sample, index)
        return sample

    def set_transform(self, transform):
        self.transform = transform

    def add_transform(self, transform):
        if type(self.transform) is not Compose:
            raise TypeError("To add transform, the transform should already be a compose of transforms")
        return DataLoadDf(self.df, self.get_feature_file_func, self.encode_function,
                          self.transform.add_transform(transform), self.return_indexes)


# ---- transform chain ---------------------------------------------------------------------------------
class _FusedStage(object):
    """A stage of the fused GPU chain: meaningful inside the Compose built by get_transforms."""

    def __call__(self, sample):
        raise NotImplementedError(
            "%s runs fused on the GPU inside Compose (dcase_logmel_finish); build the chain with "
            "utils.utils.get_transforms" % type(self).__name__)

    def __repr__(self):
        return type(self).__name__ + "()"


class AugmentGaussianNoise(_FusedStage):
    """features + |N(0, 0.25)|; the std is hard-coded to 0.5 ** 2 in the reference whatever ``std`` says."""

    def __init__(self, mean=0, std=0.5):
        self.mean = mean
        self.std = std


class ApplyLog(_FusedStage):
    """librosa.amplitude_to_db(ref=1, amin=1e-5, top_db=80) with the clip-global maximum."""


class PadOrTrunc(_FusedStage):
    def __init__(self, nb_frames):
        self.nb_frames = nb_frames


class ToTensor(_FusedStage):
    def __init__(self, unsqueeze_axis=None):
        self.unsqueeze_axis = unsqueeze_axis


class Normalize(_FusedStage):
    def __init__(self, scaler):
        self.scaler = scaler


class Compose(object):
    """Composes the transform stages; the canonical chain executes as one kernel."""

    def __init__(self, transforms, seed=None):
        self.transforms = transforms
        self._seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
        self._calls = 0
        self._plan = None

    def add_transform(self, transform):
        t = self.transforms.copy()
        t.append(transform)
        return Compose(t)

    def _build_plan(self):
        stages = list(self.transforms)
        plan = {"noise": False, "frames": None, "axis": None, "scaler": None}
        if stages and isinstance(stages[0], AugmentGaussianNoise):
            plan["noise"] = True
            stages = stages[1:]
        kinds = [type(s) for s in stages]
        if kinds[:3] != [ApplyLog, PadOrTrunc, ToTensor] or kinds[3:] not in ([], [Normalize]):
            raise NotImplementedError("only the chain of utils.utils.get_transforms is built, got %r" % (self,))
        plan["frames"] = stages[1].nb_frames
        plan["axis"] = stages[2].unsqueeze_axis
        if len(stages) == 4:
            plan["scaler"] = stages[3].scaler
        return plan

    def transform_batch(self, features, step=None):
        """features: CUDA float32 [B, T, 64] amplitude mels -> [x] or [x, x_noisy], each [B, (1,) frames, 64]."""
        from . import kernels as K
        if self._plan is None:
            self._plan = self._build_plan()
        plan = self._plan
        dev = features.device
        if plan["scaler"] is not None:
            mean, std = plan["scaler"].device_stats(dev)
        else:
            mean = torch.zeros(64, device=dev)
            std = torch.ones(64, device=dev)
        if step is None:
            step = self._calls
            self._calls += 1
        out = K.logmel_finish(features, mean, std, plan["frames"], noisy=plan["noise"], seed=self._seed,
                              step=step & 0xFFFFFFFF)
        outs = list(out) if plan["noise"] else [out]
        if plan["axis"] is not None:
            outs = [o.unsqueeze(plan["axis"] + 1) for o in outs]
        return outs

    def __call__(self, sample):
        features, label = sample
        if not torch.cuda.is_available():
            raise RuntimeError("the transform chain runs on the GPU (dcase_logmel_finish); no CPU fallback")
        f = torch.as_tensor(np.ascontiguousarray(features), dtype=torch.float32).cuda()
        outs = [o[0] for o in self.transform_batch(f[None])]
        lab = torch.as_tensor(np.asarray(label)).float().to(f.device)
        return outs + [lab]

    def __repr__(self):
        return type(self).__name__ + "(" + "".join("\n    {0}".format(t) for t in self.transforms) + "\n)"


class ConcatDataset(Dataset):
    """Concatenation of datasets that remembers which index range belongs to which (``cluster_indices``)."""

    @staticmethod
    def cumsum(sequence):
        return list(np.cumsum([len(e) for e in sequence]).tolist())

    def __init__(self, datasets):
        assert len(datasets) > 0, 'datasets should not be an empty iterable'
        self.datasets = list(datasets)
        self.cumulative_sizes = self.cumsum(self.datasets)

    @property
    def cluster_indices(self):
        starts = [0] + self.cumulative_sizes[:-1]
        return [range(a, b) for a, b in zip(starts, self.cumulative_sizes)]

    def __len__(self):
        return self.cumulative_sizes[-1]

    def __getitem__(self, idx):
        d = bisect.bisect_right(self.cumulative_sizes, idx)
        return self.datasets[d][idx - (self.cumulative_sizes[d - 1] if d else 0)]

    @property
    def cummulative_sizes(self):
        warnings.warn("cummulative_sizes attribute is renamed to cumulative_sizes", DeprecationWarning, stacklevel=2)
        return self.cumulative_sizes

    @property
    def df(self):
        return pd.concat([d.df for d in self.datasets], axis=0, ignore_index=True, sort=False)


class MultiStreamBatchSampler(Sampler):
    """Batches with a fixed number of samples from each stream, in stream order (so main.py's slice masks hold).

    ``rank`` / ``world_size`` shard every stream (rank-strided after the shared permutation) for data-parallel
    training; the defaults reproduce the reference exactly."""

    def __init__(self, data_source, batch_sizes, shuffle=True, rank=0, world_size=1, seed=None):
        self.data_source = data_source
        self.batch_sizes = batch_sizes
        n_streams = len(self.data_source.cluster_indices)
        assert len(batch_sizes) == n_streams, "batch_sizes must be the same length as the number of datasets in " \
                                              "the source {} != {}".format(len(batch_sizes), n_streams)
        self.shuffle = shuffle
        self.rank, self.world_size = rank, world_size
        if world_size > 1 and seed is None:
            # every rank must draw the SAME permutation, or the rank-strided shards overlap / miss clips
            raise ValueError("MultiStreamBatchSampler: world_size > 1 needs a seed shared by all ranks")
        if not 0 <= rank < world_size:
            raise ValueError("rank must be in [0, world_size)")
        self._rng = np.random if seed is None else np.random.RandomState(seed)

    def _stream_indices(self):
        out = []
        n_batches = len(self)
        for ind, bs in zip(self.data_source.cluster_indices, self.batch_sizes):
            ind = self._rng.permutation(ind) if self.shuffle else np.asarray(ind)
            if self.world_size > 1:
                # truncate BEFORE striding so every rank yields exactly len(self) batches (an extra batch on one rank
                # would issue a gradient exchange its peers never join)
                ind = ind[:n_batches * bs * self.world_size][self.rank::self.world_size]
            out.append(ind)
        return out

    def __iter__(self):
        streams = [grouper(ind, bs) for ind, bs in zip(self._stream_indices(), self.batch_sizes)]
        return (sum(parts, ()) for parts in zip(*streams))

    def __len__(self):
        return min(len(ind) // self.world_size // bs
                   for ind, bs in zip(self.data_source.cluster_indices, self.batch_sizes))


def grouper(iterable, n):
    "grouper('ABCDEFG', 3) --> ABC DEF (fixed-length chunks, remainder dropped)"
    return zip(*([iter(iterable)] * n))
