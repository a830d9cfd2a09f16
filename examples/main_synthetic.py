#!/usr/bin/env python
"""The flow of ``baseline/main.py:196-360`` on this package, end to end, on synthetic wav files.

Nothing here is new machinery: it strings together the drop-in objects exactly as the reference's ``__main__`` does --
wav files + tsv metadata -> ``read_audio`` / feature cache (``DatasetDcase2019Task4``) -> ``DataLoadDf`` +
``ManyHotEncoder`` -> ``Scaler.calculate_scaler(ConcatDataset(...))`` -> ``get_transforms(..., augment_type="noise")``
-> ``MultiStreamBatchSampler`` + ``DataLoader`` -> ``train(...)`` per epoch (mean teacher) -> ``get_predictions`` +
``compute_strong_metrics`` + ``get_f_measure_by_class`` -> checkpoint dict -> ``restore_from_state``.

There is no audio in the reference tree and no network, so the "dataset" is ``synth.make_clips`` written as 16-bit wav
files with weak / unlabeled / synthetic(strong) tsv tables in the reference's column layout.

    python examples/main_synthetic.py --clips 96 --epochs 2 [--workdir /tmp/dcase_synth]

Needs a B200 (no CPU fallback).  NOTE (round 1): every component used here has its own GPU parity test; this script
itself was written after the round's GPU budget was spent and has not been run on hardware yet.
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np
import pandas as pd
import scipy.io.wavfile
import torch
from torch.utils.data import DataLoader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dcase2019_task4_b200 import config as cfg, synth                                        # noqa: E402
from dcase2019_task4_b200.DataLoad import ConcatDataset, DataLoadDf, MultiStreamBatchSampler  # noqa: E402
from dcase2019_task4_b200.DatasetDcase2019Task4 import DatasetDcase2019Task4                 # noqa: E402
from dcase2019_task4_b200.evaluation_measures import (compute_strong_metrics, get_f_measure_by_class,  # noqa: E402
                                                      get_predictions)
from dcase2019_task4_b200.main import (build_state, load_checkpoint, restore_from_state, save_checkpoint, train,  # noqa: E402
                                       update_state)
from dcase2019_task4_b200.models.CRNN import CRNN                                             # noqa: E402
from dcase2019_task4_b200.utils.Scaler import Scaler                                          # noqa: E402
from dcase2019_task4_b200.utils.utils import ManyHotEncoder, SaveBest, get_transforms, weights_init  # noqa: E402


def write_dataset(audio_dir, n_clips, seed=0):
    """n_clips wav files; the first quarter weak, the next half unlabeled, the last quarter synthetic (strong)."""
    os.makedirs(audio_dir, exist_ok=True)
    waves, events = synth.make_clips(n_clips, seed=seed)
    names = ["clip%04d.wav" % i for i in range(n_clips)]
    for name, w in zip(names, waves):
        scipy.io.wavfile.write(os.path.join(audio_dir, name), cfg.sample_rate,
                               (np.clip(w, -1, 1) * 32767).astype(np.int16))
    n_weak = n_clips // 4
    n_unl = n_clips // 2
    weak_df = pd.DataFrame({"filename": names[:n_weak],
                            "event_labels": [",".join(sorted({synth.CLASSES[c] for c, _, _ in ev}))
                                             for ev in events[:n_weak]]})
    unlabel_df = pd.DataFrame({"filename": names[n_weak:n_weak + n_unl]})
    rows = [(names[i], on, off, synth.CLASSES[c]) for i in range(n_weak + n_unl, n_clips) for c, on, off in events[i]]
    synthetic_df = pd.DataFrame(rows, columns=["filename", "onset", "offset", "event_label"])
    return names, weak_df, unlabel_df, synthetic_df


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=96)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--workdir", default=None)
    args = ap.parse_args()
    work = args.workdir or tempfile.mkdtemp(prefix="dcase_synth_")
    pooling_time_ratio = cfg.pooling_time_ratio
    t0 = time.time()
    names, weak_df, unlabel_df, synthetic_df = write_dataset(os.path.join(work, "audio"), args.clips)

    # ---- DATA (main.py:199-267) ----
    dataset = DatasetDcase2019Task4(work, base_feature_dir=os.path.join(work, "dataset", "features"),
                                    save_log_feature=False)
    dataset.extract_features_from_files(os.path.join(work, "audio"), names)          # read_audio -> mel cache
    print("features cached: %d files, %.1f s" % (len(names), time.time() - t0))
    many_hot_encoder = ManyHotEncoder(cfg.classes, n_frames=cfg.max_frames // pooling_time_ratio)
    transforms = get_transforms(cfg.max_frames)
    train_weak_df = weak_df.sample(frac=0.8, random_state=26)
    valid_weak_df = weak_df.drop(train_weak_df.index).reset_index(drop=True)
    train_weak_df = train_weak_df.reset_index(drop=True)
    filenames_train = synthetic_df.filename.drop_duplicates().sample(frac=0.8, random_state=26)
    train_synth_df = synthetic_df[synthetic_df.filename.isin(filenames_train)].copy()
    valid_synth_df = synthetic_df.drop(train_synth_df.index).reset_index(drop=True)
    train_synth_df.onset = train_synth_df.onset * cfg.sample_rate // cfg.hop_length // pooling_time_ratio
    train_synth_df.offset = train_synth_df.offset * cfg.sample_rate // cfg.hop_length // pooling_time_ratio
    train_weak_data = DataLoadDf(train_weak_df, dataset.get_feature_file, many_hot_encoder.encode_strong_df,
                                 transform=transforms)
    unlabel_data = DataLoadDf(unlabel_df, dataset.get_feature_file, many_hot_encoder.encode_strong_df,
                              transform=transforms)
    train_synth_data = DataLoadDf(train_synth_df, dataset.get_feature_file, many_hot_encoder.encode_strong_df,
                                  transform=transforms)
    list_dataset = [train_weak_data, unlabel_data, train_synth_data]
    batch_sizes = [cfg.batch_size // 4, cfg.batch_size // 2, cfg.batch_size // 4]
    strong_mask = slice(cfg.batch_size // 4 + cfg.batch_size // 2, cfg.batch_size)
    weak_mask = slice(batch_sizes[0])

    scaler = Scaler()
    scaler.calculate_scaler(ConcatDataset(list_dataset))                             # device reduction
    print("scaler mean_[:4] =", scaler.mean_[:4], " std_[:4] =", scaler.std_[:4])

    transforms = get_transforms(cfg.max_frames, scaler, augment_type="noise")
    for d in list_dataset:
        d.set_transform(transforms)
    concat_dataset = ConcatDataset(list_dataset)
    sampler = MultiStreamBatchSampler(concat_dataset, batch_sizes=batch_sizes)
    training_data = DataLoader(concat_dataset, batch_sampler=sampler)
    transforms_valid = get_transforms(cfg.max_frames, scaler=scaler)
    valid_synth_data = DataLoadDf(valid_synth_df, dataset.get_feature_file, many_hot_encoder.encode_strong_df,
                                  transform=transforms_valid)
    valid_weak_data = DataLoadDf(valid_weak_df, dataset.get_feature_file, many_hot_encoder.encode_weak,
                                 transform=transforms_valid)

    # ---- Model (main.py:276-311) ----
    crnn_kwargs = cfg.crnn_kwargs
    crnn = CRNN(**crnn_kwargs)
    crnn_ema = CRNN(**crnn_kwargs)
    crnn.apply(weights_init)
    crnn_ema.apply(weights_init)
    for param in crnn_ema.parameters():
        param.detach_()
    optim_kwargs = {"lr": 0.001, "betas": (0.9, 0.999)}
    optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, crnn.parameters()), **optim_kwargs)
    state = build_state(crnn, optimizer, crnn_kwargs, optim_kwargs, pooling_time_ratio, scaler, many_hot_encoder,
                        crnn_ema=crnn_ema)
    save_best_cb = SaveBest("sup")
    model_dir = os.path.join(work, "model")
    os.makedirs(model_dir, exist_ok=True)

    # ---- Train (main.py:316-356) ----
    for epoch in range(args.epochs):
        crnn, crnn_ema = crnn.train().cuda(), crnn_ema.train().cuda()
        train(training_data, crnn, optimizer, epoch, ema_model=crnn_ema, weak_mask=weak_mask, strong_mask=strong_mask)
        crnn = crnn.eval()
        predictions = get_predictions(crnn, valid_synth_data, many_hot_encoder.decode_strong, pooling_time_ratio)
        valid_events_metric = compute_strong_metrics(predictions, valid_synth_df)
        weak_metric = get_f_measure_by_class(crnn, len(cfg.classes),
                                             DataLoader(valid_weak_data, batch_size=cfg.batch_size))
        print(valid_events_metric)
        print("Weak F1-score macro averaged: {}".format(np.mean(weak_metric)))
        update_state(state, crnn, optimizer, epoch, valid_metric=valid_events_metric.results(), crnn_ema=crnn_ema)
        global_valid = valid_events_metric.results_class_wise_average_metrics()['f_measure']['f_measure']
        global_valid = global_valid + np.mean(weak_metric)
        if save_best_cb.apply(global_valid):
            save_checkpoint(state, os.path.join(model_dir, "baseline_best"))

    # ---- TestModel.py:26-40 on the saved checkpoint ----
    back = load_checkpoint(os.path.join(model_dir, "baseline_best"))
    model, scaler2, encoder2, ptr = restore_from_state(back)
    model = model.eval().cuda()
    valid = DataLoadDf(valid_synth_df, dataset.get_feature_file, encoder2.encode_strong_df,
                       transform=get_transforms(cfg.max_frames, scaler=scaler2))
    again = get_predictions(model, valid, encoder2.decode_strong, ptr)
    print("restored checkpoint of epoch %d: %d predicted events" % (back["epoch"], len(again)))
    print("done in %.1f s (workdir %s)" % (time.time() - t0, work))


if __name__ == "__main__":
    main()
