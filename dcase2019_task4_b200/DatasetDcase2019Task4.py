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

    # ---- tsv bookkeeping around the feature extraction (DatasetDcase2019Task4.py:92-181) --------------------------
    def initialize_and_get_df(self, tsv_path, subpart_data=None, download=True):
        """The DataFrame of one metadata table, with the features of its files extracted into the cache.  The youtube
        download of the reference is not rebuilt: ``download`` is accepted and the audio must already be on disk."""
        meta_name = os.path.join(self.local_path, tsv_path)
        return self.extract_features_from_meta(meta_name, subpart_data)

    @staticmethod
    def get_classes(list_dfs):
        found = set()
        for df in list_dfs:
            if "event_label" in df.columns:
                found.update(df["event_label"].dropna().unique())
            elif "event_labels" in df.columns:
                found.update(df.event_labels.str.split(',', expand=True).unstack().dropna().unique())
        return list(found)

    @staticmethod
    def get_subpart_data(df, subpart_data):
        if subpart_data <= len(df["filename"].unique()):
            keep = df["filename"].drop_duplicates().sample(subpart_data, random_state=10)
            df = df[df["filename"].isin(keep)].reset_index(drop=True)
        return df

    @staticmethod
    def get_df_from_meta(meta_name, subpart_data=None):
        import pandas as pd
        df = pd.read_csv(meta_name, header=0, sep="\t")
        return df if subpart_data is None else DatasetDcase2019Task4.get_subpart_data(df, subpart_data)

    @staticmethod
    def get_audio_dir_path_from_meta(filepath):
        """dataset/metadata/train/weak.tsv -> dataset/audio/train/weak; validation tables share dataset/audio/validation."""
        parts = os.path.splitext(filepath)[0].replace("metadata", "audio").split('/')
        if len(parts) >= 2 and parts[-2] == 'validation':
            parts = parts[:-1]
        return os.path.abspath('/'.join(parts))

    def extract_features_from_meta(self, tsv_audio, subpart_data=None):
        """DatasetDcase2019Task4.py:233-270: features of every file of the table that exists on disk; rows of missing
        files are dropped from the returned DataFrame."""
        df_meta = self.get_df_from_meta(tsv_audio, subpart_data)
        wav_dir = self.get_audio_dir_path_from_meta(tsv_audio)
        names = list(df_meta.filename.unique())
        cached = [n for n in names
                  if os.path.exists(os.path.join(self.feature_dir, os.path.splitext(n)[0] + ".npy"))]
        todo = [n for n in names if n not in set(cached)]
        done = set(cached) | set(self.extract_features_from_files(wav_dir, todo) if todo else [])
        missing = [n for n in names if n not in done and not os.path.isfile(os.path.join(wav_dir, n))]
        if missing:
            df_meta = df_meta[~df_meta.filename.isin(missing)]
        return df_meta.reset_index(drop=True)

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
