"""Feature-extraction side of ``baseline/DatasetDcase2019Task4.py`` on the GPU.

Kept: ``calculate_mel_spec`` (:197-231), the feature-cache directory naming (:83-88), ``get_feature_file``
(:183-195) and the on-disk format (float32 [T, 64] amplitude-mel ``.npy``), so caches are interchangeable with the
reference's.  Not rebuilt (out of scope): youtube download, tsv bookkeeping.
"""
import os

import numpy as np
import torch

from . import config as cfg
from . import kernels as K


class DatasetDcase2019Task4:

    def __init__(self, local_path="", base_feature_dir="features", recompute_features=False, save_log_feature=True,
                 create_dirs=True):
        if save_log_feature:
            raise NotImplementedError("main.py:201 uses save_log_feature=False; the dB step runs in the transform chain")
        self.local_path = local_path
        self.recompute_features = recompute_features
        self.save_log_feature = save_log_feature
        feature_dir = os.path.join(base_feature_dir, "sr" + str(cfg.sample_rate) + "_win" + str(cfg.n_window)
                                   + "_hop" + str(cfg.hop_length) + "_mels" + str(cfg.n_mels)) + "_nolog"
        self.feature_dir = os.path.join(feature_dir, "features")
        if create_dirs and not os.path.exists(self.feature_dir):
            os.makedirs(self.feature_dir)

    def get_feature_file(self, filename):
        return np.load(os.path.join(self.feature_dir, os.path.splitext(filename)[0] + ".npy"))

    def calculate_mel_spec(self, audio):
        """audio: 1-D numpy / torch waveform at 44.1 kHz (float, or int16 PCM) -> float32 numpy [T, 64]."""
        return self.calculate_mel_spec_batch(torch.as_tensor(np.asarray(audio))[None])[0].cpu().numpy()

    @staticmethod
    def calculate_mel_spec_batch(audio):
        """[B, L] waveforms (host or device; float or int16) -> CUDA float32 [B, 1 + L // 511, 64] amplitude mels."""
        if not torch.cuda.is_available():
            raise RuntimeError("calculate_mel_spec runs on the GPU (dcase_logmel_fwd); no CPU fallback")
        a = torch.as_tensor(audio)
        if a.dtype != torch.int16:
            a = a.float()
        return K.logmel_fwd(a.cuda(non_blocking=True))

    def extract_features_from_files(self, wav_dir, wav_names):
        """The loop body of extract_features_from_meta (:233-270) for a list of wav files: read_audio (mono mix-down
        on the GPU) -> calculate_mel_spec (dcase_logmel_fwd) -> ``<name>.npy`` in the reference's cache format.
        Missing files are reported and skipped, empty ones flagged as corrupted, as the reference does.
        Returns the names whose features exist afterwards."""
        from .utils.utils import read_audio_device
        done = []
        for wav_name in wav_names:
            out_path = os.path.join(self.feature_dir, os.path.splitext(wav_name)[0] + ".npy")
            if self.recompute_features or not os.path.exists(out_path):
                wav_path = os.path.join(wav_dir, wav_name)
                if not os.path.isfile(wav_path):
                    print("File %s is in the tsv file but the feature is not extracted!" % wav_path)
                    continue
                audio, _ = read_audio_device(wav_path, cfg.sample_rate)
                if audio.shape[0] == 0:
                    print("File %s is corrupted!" % wav_path)
                    continue
                np.save(out_path, K.logmel_fwd(audio[None])[0].cpu().numpy())
            done.append(wav_name)
        return done

    def extract_features_to_cache(self, names_and_audio):
        """Write ``<name>.npy`` caches (reference format) for an iterable of (wav_name, waveform)."""
        for name, audio in names_and_audio:
            out = os.path.join(self.feature_dir, os.path.splitext(name)[0] + ".npy")
            if self.recompute_features or not os.path.exists(out):
                np.save(out, self.calculate_mel_spec(audio))
