"""CPU: host-side mirror of the reference interface (no kernels involved)."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from dcase2019_task4_b200 import DataLoad, config as cfg, dp
from dcase2019_task4_b200.DatasetDcase2019Task4 import DatasetDcase2019Task4 as DatasetBook
from dcase2019_task4_b200.models.CRNN import CRNN
from dcase2019_task4_b200.utils import ramps
from dcase2019_task4_b200.utils.Scaler import Scaler
from dcase2019_task4_b200.utils.utils import (AverageMeterSet, ManyHotEncoder, SaveBest, find_contiguous_regions,
                                              get_transforms, weights_init)
from oracle import crnn as ocrnn
from oracle import mel as omel


def _df(n, kind):
    names = ["f%d_%s.wav" % (i, kind) for i in range(n)]
    if kind == "weak":
        return pd.DataFrame({"filename": names, "event_labels": ["Dog,Cat" if i % 2 else "Speech" for i in range(n)]})
    if kind == "unl":
        return pd.DataFrame({"filename": names})
    rows = []
    for i, f in enumerate(names):
        rows.append({"filename": f, "onset": 1.0, "offset": 5.0, "event_label": "Dog"})
        rows.append({"filename": f, "onset": 2.0 + i, "offset": 9.0 + i, "event_label": "Speech"})
    return pd.DataFrame(rows)


def _datasets(enc):
    feats = lambda name: np.full((4, 64), float(len(name)), dtype=np.float32)
    return [DataLoad.DataLoadDf(_df(13, "weak"), feats, enc.encode_strong_df),
            DataLoad.DataLoadDf(_df(31, "unl"), feats, enc.encode_strong_df),
            DataLoad.DataLoadDf(_df(14, "syn"), feats, enc.encode_strong_df)]


def test_dataset_labels_and_target_layout():
    enc = ManyHotEncoder(cfg.classes, n_frames=cfg.max_frames // cfg.pooling_time_ratio)
    weak, unl, syn = _datasets(enc)
    assert len(weak) == 13 and len(unl) == 31 and len(syn) == 14     # strong frame has 2 rows per file
    _, y = weak[1]
    assert y.shape == (108, 10) and y[:, cfg.classes.index("Dog")].all() and y.sum() == 2 * 108
    _, y = unl[0]
    assert (y == -1).all()                                            # utils.py:82-85
    _, y = syn[3]
    assert y[1:5, cfg.classes.index("Dog")].all() and y[5, cfg.classes.index("Dog")] == 0
    assert y[5:12, cfg.classes.index("Speech")].all()


def test_multistream_sampler_matches_reference_semantics():
    enc = ManyHotEncoder(cfg.classes, n_frames=108)
    concat = DataLoad.ConcatDataset(_datasets(enc))
    assert [len(r) for r in concat.cluster_indices] == [13, 31, 14] and len(concat) == 58
    s = DataLoad.MultiStreamBatchSampler(concat, batch_sizes=[2, 4, 2], seed=0)
    assert len(s) == min(13 // 2, 31 // 4, 14 // 2) == 6
    batches = list(s)
    assert len(batches) == 6
    for b in batches:
        assert len(b) == 8
        assert all(i < 13 for i in b[:2]) and all(13 <= i < 44 for i in b[2:6]) and all(44 <= i < 58 for i in b[6:])
    flat = [i for b in batches for i in b]
    assert len(set(flat)) == len(flat)                                # without replacement within an epoch
    with pytest.raises(AssertionError):
        DataLoad.MultiStreamBatchSampler(concat, batch_sizes=[2, 4])


def test_multistream_sampler_rank_sharding():
    enc = ManyHotEncoder(cfg.classes, n_frames=108)
    concat = DataLoad.ConcatDataset(_datasets(enc))
    per_rank = dp.per_rank_batch_sizes([4, 8, 4], 2)
    assert per_rank == [2, 4, 2]
    with pytest.raises(ValueError):
        dp.per_rank_batch_sizes([6, 12, 6], 4)
    seen = []
    for rank in range(2):
        s = DataLoad.MultiStreamBatchSampler(concat, per_rank, rank=rank, world_size=2, seed=5)
        batches = list(s)
        assert len(batches) == len(s) == min(13 // 2 // 2, 31 // 2 // 4, 14 // 2 // 2)
        seen.append({i for b in batches for i in b})
    assert not (seen[0] & seen[1])                                    # ranks see disjoint clips


def test_many_hot_encoder_roundtrip_and_regions():
    enc = ManyHotEncoder(np.array(cfg.classes), n_frames=20)
    y = enc.encode_strong_df([["Dog", 2, 6], ["Cat", 0, 20], ["Dog", 10, 12]])
    dec = enc.decode_strong(y)
    assert ["Dog", 2, 6] in [[a, int(b), int(c)] for a, b, c in dec] and ["Cat", 0, 20] in [[a, int(b), int(c)] for a, b, c in dec]
    assert find_contiguous_regions([0, 1, 1, 0, 1]).tolist() == [[1, 3], [4, 5]]
    assert enc.decode_weak(enc.encode_weak(["Dog", "Speech"])) == ["Dog", "Speech"]
    assert ManyHotEncoder.load_state_dict(enc.state_dict()).labels == cfg.classes


def test_scaler_wire_format_and_no_cpu_fallback(tmp_path):
    """state_dict / save / load / normalize keep the reference's wire format (Scaler.py:99-125); the reduction
    itself (means / calculate_scaler) is a device kernel and raises without a GPU (tests/test_gpu_scaler.py)."""
    rng = np.random.default_rng(0)
    data = [(torch.from_numpy(rng.normal(-20, 8, (1, 50, 64)).astype(np.float32)), None) for _ in range(5)]
    m, m2 = omel.scaler_means([d[0].numpy() for d in data])
    sc = Scaler()
    with pytest.raises(NotImplementedError):
        sc.state_dict()                                   # nothing computed yet (Scaler.py:108-109)
    sc.load_state_dict({"mean_": m.tolist(), "mean_of_square_": m2.tolist()})
    assert np.allclose(sc.std_, omel.scaler_std(m, m2), rtol=0, atol=1e-12)
    sd = sc.state_dict()
    assert set(sd) == {"mean_", "mean_of_square_"} and isinstance(sd["mean_"], list)
    sc.save(tmp_path / "s.json")
    sc2 = Scaler()
    sc2.load(tmp_path / "s.json")
    assert np.array_equal(sc2.std_, sc.std_) and np.array_equal(sc2.mean_, sc.mean_)
    x = data[0][0]
    assert np.allclose(sc.normalize(x).numpy(), (x.numpy() - m) / sc.std_, atol=1e-5)
    assert np.allclose(sc.normalize(x.numpy()), (x.numpy() - m) / sc.std_)
    if not torch.cuda.is_available():
        from dcase2019_task4_b200._lib import DcaseError
        with pytest.raises(DcaseError):
            Scaler().calculate_scaler(data)


def test_transform_chain_structure_and_small_utils():
    t = get_transforms(864, Scaler(), augment_type="noise")
    assert [type(s).__name__ for s in t.transforms] == ["AugmentGaussianNoise", "ApplyLog", "PadOrTrunc", "ToTensor", "Normalize"]
    assert t._build_plan()["frames"] == 864 and t._build_plan()["noise"]
    assert [type(s).__name__ for s in get_transforms(100).transforms] == ["ApplyLog", "PadOrTrunc", "ToTensor"]
    with pytest.raises(NotImplementedError):
        DataLoad.Compose([DataLoad.ToTensor()])._build_plan()
    assert ramps.sigmoid_rampup(0, 10) == pytest.approx(np.exp(-5)) and ramps.sigmoid_rampup(3, 0) == 1.0
    sb = SaveBest("sup")
    assert [sb.apply(v) for v in (0.1, 0.05, 0.3)] == [True, False, True] and sb.best_epoch == 2
    ms = AverageMeterSet()
    ms.update("Loss", 2.0)
    ms.update("Loss", 4.0)
    assert ms["Loss"].avg == 3.0 and "Loss 4.0000" in str(ms)


def test_crnn_module_surface_on_cpu():
    torch.manual_seed(0)
    m = CRNN(**cfg.crnn_kwargs)
    m.apply(weights_init)
    names = [k for k, _ in m.named_parameters()]
    assert names == list(ocrnn.param_shapes(10).keys())
    flat = m.flat_parameters()
    assert flat.numel() == 214356
    assert abs(float(m.cnn.cnn.batchnorm1.weight.mean()) - 1.0) < 0.02 and float(m.dense.bias.abs().max()) == 0.0
    with torch.no_grad():
        m.dense.weight.fill_(3.0)
    off = sum(int(np.prod(s)) for k, s in ocrnn.param_shapes(10).items() if k < "dense" and not k.startswith("dense"))
    assert float(flat[214356 - 2 * 1290: 214356 - 2 * 1290 + 1280].min()) == 3.0   # parameters are views of the slab
    sd = m.state_dict()
    assert set(sd) == {"cnn", "rnn", "dense"} and "dense_softmax" not in sd      # CRNN.py:49-53 quirk kept
    m2 = CRNN(**cfg.crnn_kwargs)
    m2.load(parameters=sd)
    assert torch.equal(m2.rnn.rnn.weight_hh_l1_reverse, m.rnn.rnn.weight_hh_l1_reverse)
    for p in m2.parameters():
        p.detach_()                                                               # main.py:286-287
    with pytest.raises(NotImplementedError):
        CRNN(**dict(cfg.crnn_kwargs, activation="relu"))


# ------------------------------------------------------------------------------------------------------------
# posteriors -> events -> event-based F1 (restated sed_eval definitions, hand-computed cases)
# ------------------------------------------------------------------------------------------------------------
def _events(rows):
    import pandas as pd
    return pd.DataFrame(rows, columns=["filename", "event_label", "onset", "offset"])


def test_postprocess_threshold_median_and_decode_to_seconds():
    import numpy as np
    from dcase2019_task4_b200 import config as cfg
    from dcase2019_task4_b200 import evaluation_measures as em
    from dcase2019_task4_b200.utils.utils import ManyHotEncoder
    T, C = 108, 10
    p = np.full((2, T, C), 0.1, dtype=np.float32)
    p[0, 10:30, 2] = 0.9          # a 20-frame event
    p[0, 50, 2] = 0.99            # a single-frame blip: removed by the 5-frame median filter
    p[0, 70:90, 5] = 0.9
    p[0, 80, 5] = 0.2             # a single-frame hole: filled by the median filter
    p[1, 0:3, 0] = 0.6            # touches the clip start (scipy 'reflect' boundary keeps it)
    act = em.postprocess_posteriors(p)
    assert act.shape == p.shape and act.dtype == bool
    assert act[0, 10:30, 2].all() and not act[0, 40:60, 2].any() and act[0, 70:90, 5].all()
    enc = ManyHotEncoder(["c%d" % i for i in range(C)], n_frames=T)
    ev = enc.decode_strong(act[0])
    assert sorted(ev) == [["c2", 10, 30], ["c5", 70, 90]]
    sec = em.frames_to_seconds(np.array([10.0, 30.0]), pooling_time_ratio=cfg.pooling_time_ratio)
    assert np.allclose(sec, np.array([10.0, 30.0]) * 8 * 511 / 44100)     # evaluation_measures.py:226-227
    assert act[1, 0:3, 0].all() and not act[1, 4:, 0].any()


def test_event_based_f1_collars_and_optimal_matching():
    from dcase2019_task4_b200 import evaluation_measures as em
    ref = _events([["a.wav", "Dog", 1.0, 6.0],        # 5-s event: offset tolerance max(0.2, 0.2 * 5) = 1.0 s
                   ["a.wav", "Speech", 0.0, 1.0],
                   ["a.wav", "Speech", 0.15, 1.15],
                   ["b.wav", "Dog", 2.0, 2.5]])
    est = _events([["a.wav", "Dog", 1.15, 6.9],       # onset +0.15 (<= 0.2), offset +0.9 (<= 1.0): hit
                   ["a.wav", "Speech", 0.1, 1.1],     # X: could pair with either Speech event
                   ["a.wav", "Speech", -0.1, 0.9],    # Y: pairs only with the first one -> optimal matching finds 2 hits
                   ["b.wav", "Dog", 2.25, 2.5],       # onset +0.25: miss
                   ["b.wav", "Cat", 0.0, 1.0]])       # a class absent from the reference
    m = em.event_based_evaluation_df(ref, est)
    r = m.results()
    cw = r["class_wise"]
    assert cw["Dog"]["count"] == {"Nref": 2, "Nsys": 2, "Ntp": 1}
    assert cw["Speech"]["count"] == {"Nref": 2, "Nsys": 2, "Ntp": 2}
    assert cw["Cat"]["count"] == {"Nref": 0, "Nsys": 1, "Ntp": 0}
    assert abs(cw["Dog"]["f_measure"]["f_measure"] - 0.5) < 1e-12 and cw["Speech"]["f_measure"]["f_measure"] == 1.0
    ov = r["overall"]["f_measure"]                    # micro: Ntp 3, Nsys 5, Nref 4
    assert abs(ov["precision"] - 3 / 5) < 1e-12 and abs(ov["recall"] - 3 / 4) < 1e-12
    assert abs(ov["f_measure"] - 2 * 0.6 * 0.75 / 1.35) < 1e-12
    # macro average over the union of labels: Cat (system events, no reference event) scores 0
    assert abs(r["class_wise_average"]["f_measure"]["f_measure"] - 0.5) < 1e-12
    # offset just outside the 20 % tolerance
    est2 = _events([["a.wav", "Dog", 1.0, 7.1]])
    assert em.event_based_evaluation_df(ref[ref.event_label == "Dog"], est2).results()["overall"]["count"]["Ntp"] == 0
    assert "class-wise average" in str(m)


def test_segment_based_metric_and_empty_files():
    import numpy as np
    from dcase2019_task4_b200 import evaluation_measures as em
    ref = _events([["a.wav", "Dog", 0.2, 2.4], ["b.wav", np.nan, np.nan, np.nan]])      # b.wav: no event (tsv convention)
    est = _events([["a.wav", "Dog", 1.1, 3.5]])
    seg = em.segment_based_evaluation_df(ref, est, time_resolution=1.0).results()
    # reference active in segments 0,1,2; system in 1,2,3 -> Ntp 2
    assert seg["class_wise"]["Dog"]["count"] == {"Nref": 3, "Nsys": 3, "Ntp": 2}
    ev = em.compute_strong_metrics(est, ref).results()
    assert ev["overall"]["count"] == {"Nref": 1, "Nsys": 1, "Ntp": 0}
    assert em.get_event_list_current_file(ref, "b.wav") == []


def _random_event_files(draw_rng, n_files, labels):
    files = []
    for _ in range(n_files):
        ref, est = [], []
        for _ in range(draw_rng.integers(0, 6)):
            on = float(np.round(draw_rng.uniform(0, 8), 2))
            ref.append((labels[draw_rng.integers(len(labels))], on, on + float(np.round(draw_rng.uniform(0.1, 4), 2))))
        for lab, on, off in ref:                                   # system events: jittered copies, misses, extras
            u = draw_rng.uniform()
            if u < 0.7:
                est.append((lab, max(0.0, on + float(np.round(draw_rng.normal(0, 0.15), 2))),      # onsets are frame times: >= 0
                            off + float(np.round(draw_rng.normal(0, 0.4), 2))))
            if u > 0.85:
                est.append((lab, on + 0.05, off - 0.05))               # a second candidate for the same reference event
        for _ in range(draw_rng.integers(0, 3)):
            on = float(np.round(draw_rng.uniform(0, 8), 2))
            est.append((labels[draw_rng.integers(len(labels))], on, on + float(np.round(draw_rng.uniform(0.1, 3), 2))))
        files.append((ref, [(l, a, max(b, a + 0.01)) for l, a, b in est]))
    return files


def test_event_and_segment_metrics_match_the_bruteforce_oracle_on_random_event_lists():
    """evaluation_measures.py:124-182 through an independent restatement of sed_eval's published algorithm
    (oracle/sed_metrics.py: exhaustive optimal matching, explicit segment loops): 300 random multi-file cases."""
    from dcase2019_task4_b200 import evaluation_measures as em
    from oracle import sed_metrics as osed
    rng = np.random.default_rng(20191)
    labels = ["Dog", "Speech", "Cat"]
    for case in range(300):
        files = _random_event_files(rng, int(rng.integers(1, 4)), labels)
        rows_r, rows_e = [], []
        for i, (ref, est) in enumerate(files):
            name = "f%d.wav" % i
            rows_r += [[name, l, a, b] for l, a, b in ref] or [[name, np.nan, np.nan, np.nan]]   # tsv convention: no event
            rows_e += [[name, l, a, b] for l, a, b in est]
        ref_df = _events(rows_r)
        est_df = _events(rows_e) if rows_e else _events([["none.wav", np.nan, np.nan, np.nan]])
        used = sorted({l for ref, est in files for l, _, _ in list(ref) + list(est)})
        got_e = em.event_based_evaluation_df(ref_df, est_df).results()
        got_s = em.segment_based_evaluation_df(ref_df, est_df, time_resolution=1.0).results()
        want_e, want_s = osed.event_based(files, used), osed.segment_based(files, used)
        for got, want in ((got_e, want_e), (got_s, want_s)):
            assert got["overall"]["count"] == want["overall"]["count"], (case, got["overall"], want["overall"])
            for l in used:
                assert got["class_wise"][l]["count"] == want["class_wise"][l]["count"], (case, l)
            for k in ("f_measure", "precision", "recall"):
                assert abs(got["overall"]["f_measure"][k] - want["overall"]["f_measure"][k]) < 1e-12
                if used:
                    assert abs(got["class_wise_average"]["f_measure"][k] - want["class_wise_average"]["f_measure"][k]) < 1e-12


def test_checkpoint_dict_matches_reference_layout(tmp_path):
    """main.py:293-309 / :335-356 / TestModel.py:26-40: same keys, nested model state dicts, loadable on CPU."""
    from dcase2019_task4_b200 import main as bmain
    from dcase2019_task4_b200 import main_simple_CRNN as simple
    crnn, ema = CRNN(**cfg.crnn_kwargs), CRNN(**cfg.crnn_kwargs)
    crnn.apply(weights_init)
    optim_kwargs = {"lr": 0.001, "betas": (0.9, 0.999)}
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, crnn.parameters()), **optim_kwargs)
    sc = Scaler()
    sc.load_state_dict({"mean_": list(range(64)), "mean_of_square_": [float(i * i + 4) for i in range(64)]})
    enc = ManyHotEncoder(cfg.classes, n_frames=cfg.max_frames // cfg.pooling_time_ratio)
    state = bmain.build_state(crnn, opt, cfg.crnn_kwargs, optim_kwargs, cfg.pooling_time_ratio, sc, enc, crnn_ema=ema)
    assert list(state) == ["model", "model_ema", "optimizer", "pooling_time_ratio", "scaler", "many_hot_encoder"]
    assert list(state["model"]) == ["name", "args", "kwargs", "state_dict"] and state["model"]["name"] == "CRNN"
    assert set(state["model"]["state_dict"]) == {"cnn", "rnn", "dense"} and state["optimizer"]["name"] == "Adam"
    bmain.update_state(state, crnn, opt, 3, valid_metric={"overall": {"f_measure": {"f_measure": 0.25}}}, crnn_ema=ema)
    bmain.save_checkpoint(state, tmp_path / "baseline_epoch_3")
    back = bmain.load_checkpoint(tmp_path / "baseline_epoch_3")
    assert back["epoch"] == 3 and back["valid_metric"]["overall"]["f_measure"]["f_measure"] == 0.25
    m2, sc2, enc2, ptr = bmain.restore_from_state(back)
    assert ptr == cfg.pooling_time_ratio and enc2.labels == cfg.classes and np.array_equal(sc2.std_, sc.std_)
    for k in ("cnn.cnn.conv1.weight", "rnn.rnn.weight_hh_l1_reverse", "dense.bias"):
        assert torch.equal(dict(m2.named_parameters())[k], dict(crnn.named_parameters())[k])
    simple_state = bmain.build_state(crnn, opt, cfg.crnn_kwargs, optim_kwargs, cfg.pooling_time_ratio, sc, enc)
    assert "model_ema" not in simple_state                       # main_simple_CRNN.py:203-215
    assert simple.masks_for(24) == (slice(12), slice(12, 24)) and simple.masks_for(24, no_weak=True) == (None, slice(24))


def test_wav_container_parsing_and_loud_failures(tmp_path):
    """read_audio's host half (utils/utils.py:175-193): container parsing keeps 16-bit PCM as int16 frames, maps other
    encodings to soundfile's float range; CPU execution fails loudly (mix-down and resampling run on the device)."""
    import scipy.io.wavfile
    from dcase2019_task4_b200.utils.utils import read_audio, read_wav_frames
    rng = np.random.default_rng(0)
    pcm = rng.integers(-32768, 32767, (1000, 2)).astype(np.int16)
    scipy.io.wavfile.write(tmp_path / "a.wav", 44100, pcm)
    scipy.io.wavfile.write(tmp_path / "b.wav", 22050, (pcm[:, 0].astype(np.int32) << 16))
    scipy.io.wavfile.write(tmp_path / "c.wav", 44100, (pcm[:, 0] / 32768.0).astype(np.float32))
    frames, fs = read_wav_frames(tmp_path / "a.wav")
    assert fs == 44100 and frames.dtype == np.int16 and np.array_equal(frames, pcm)
    frames, fs = read_wav_frames(tmp_path / "b.wav")
    assert fs == 22050 and frames.shape == (1000, 1) and np.allclose(frames[:, 0], pcm[:, 0] / 32768.0)
    frames, _ = read_wav_frames(tmp_path / "c.wav")
    assert frames.dtype == np.float32 and np.allclose(frames[:, 0], pcm[:, 0] / 32768.0)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            read_audio(tmp_path / "a.wav", 44100)
        with pytest.raises(RuntimeError):
            read_audio(tmp_path / "b.wav", 44100)


def test_resample_oracle_is_a_band_limited_interpolator():
    """oracle/resample.py restates librosa.resample's kaiser_best (resampy) from its published algorithm -- resampy and
    librosa are absent, so the restatement is checked against what any correct band-limited resampler must do: tones
    below both Nyquist limits come out as the same tones at the new rate (amplitude, frequency AND phase), the output
    length is librosa's ceil(n * ratio), down-sampling removes what is above the new Nyquist, and an independent
    polyphase resampler (scipy.signal.resample_poly) agrees away from the edges."""
    import scipy.signal
    from oracle import resample as oresample
    for sr_in, sr_out in ((16000, 44100), (48000, 44100), (22050, 44100), (44100, 16000)):
        n = sr_in // 2
        t_in, n_out = np.arange(n) / sr_in, int(np.ceil(n * sr_out / sr_in))
        t_out = np.arange(n_out) / sr_out
        f1, f2 = 440.0, 0.35 * min(sr_in, sr_out)
        tone = lambda t: 0.7 * np.sin(2 * np.pi * f1 * t + 0.3) + 0.2 * np.cos(2 * np.pi * f2 * t)
        y = oresample.resample(tone(t_in), sr_in, sr_out)
        assert y.shape == (n_out,)
        core = slice(2000, n_out - 2000)                      # 64 zero crossings of filter support at the edges
        # up-sampling interpolates exactly; down-sampling steps through the filter table in int(ratio * 512) entries (a
        # truncation in resampy's own algorithm: 185 instead of 185.76 at 44.1 -> 16 kHz), a gain error of a few 1e-3
        assert np.abs(y - tone(t_out))[core].max() <= (2e-5 if sr_out > sr_in else 5e-3)
        g = np.gcd(sr_in, sr_out)
        poly = scipy.signal.resample_poly(tone(t_in), sr_out // g, sr_in // g)
        assert np.abs(y[core] - poly[:n_out][core]).max() <= 4e-3    # scipy's default Kaiser(5) filter is much shorter
    # above the new Nyquist: gone (stop band of the kaiser_best filter)
    n = 44100
    high = np.sin(2 * np.pi * 15000.0 * np.arange(n) / 44100)
    assert np.abs(oresample.resample(high, 44100, 16000))[2000:-2000].max() <= 1e-3


def test_weak_f_measure_by_class_hand_computed():
    """evaluation_measures.get_f_measure_by_class (evaluation_measures.py:19-82) on a stub model: counts by hand."""
    from dcase2019_task4_b200 import evaluation_measures as em

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, x):                 # weak posterior = the first three features of frame 0
            return x[:, 0, :, :3].repeat(1, 4, 1), x[:, 0, 0, :3]

    def batch(weak, labels):
        x = torch.zeros(len(weak), 1, 4, 64)
        x[:, 0, 0, :3] = torch.tensor(weak)
        return x, torch.tensor(labels, dtype=torch.float32)

    loader = [batch([[0.9, 0.2, 0.5], [0.8, 0.7, 0.1]], [[1, 0, 0], [0, 1, 0]]),
              batch([[0.1, 0.6, 0.4], [0.6, 0.4, 0.3]], [[1, 1, 0], [1, 0, 0]])]
    # class 0: est 1,1,0,1 vs ref 1,0,1,1 -> tp 2 fp 1 fn 1 -> 4/6;  class 1: est 0,1,1,0 vs ref 0,1,1,0 -> 1.0
    # class 2: est 0 (0.5 is not > 0.5),0,0,0 vs ref 0 -> empty denominator -> 0
    f = em.get_f_measure_by_class(Stub(), 3, loader)
    assert np.allclose(f, [4 / 6, 1.0, 0.0])
    f = em.get_f_measure_by_class(Stub(), 3, loader, thresholds_=[0.85, 0.5, 0.45])
    # class 0 @0.85: est 1,0,0,0 -> tp 1 fn 2 -> 2/4;  class 2 @0.45: est 1,0,0,0 vs ref 0 -> fp 1 -> 0
    assert np.allclose(f, [0.5, 1.0, 0.0])
    # frame-level labels are reduced by the maximum over time
    x, y = loader[0]
    y3 = torch.zeros(2, 5, 3)
    y3[0, 2, 0] = 1
    y3[1, 4, 1] = 1
    assert np.allclose(em.get_f_measure_by_class(Stub(), 3, [(x, y3)]), em.get_f_measure_by_class(Stub(), 3, [(x, y)]))
    tp, fp, fn, tn = em.intermediate_at_measures(np.array([[1, 0], [0, 0]]), np.array([[1, 1], [0, 0]]))
    assert (tp.tolist(), fp.tolist(), fn.tolist(), tn.tolist()) == ([1, 0], [0, 1], [0, 0], [1, 1])


def test_workspace_pool_returns_buffers_to_their_own_shape():
    """Regression: a workspace given back after backward used to land in the first pool with a free slot, whatever
    its shape -- a small buffer could then be handed to a larger batch.  Host bookkeeping only (CPU tensors)."""
    from dcase2019_task4_b200 import kernels as K
    m = CRNN(**cfg.crnn_kwargs)
    big = torch.empty(K.workspace_bytes(2, 64, 10), dtype=torch.uint8)
    small = torch.empty(K.workspace_bytes(1, 64, 10), dtype=torch.uint8)
    assert small.numel() < big.numel()
    m._ws_pool = {(2, 64, "cpu"): [big]}                       # one free slot left in the pool of the larger shape
    m._give_workspace(small, 1, 64)
    assert m._ws_pool[(2, 64, "cpu")] == [big] or all(w is big for w in m._ws_pool[(2, 64, "cpu")])
    assert m._ws_pool[(1, 64, "cpu")][0] is small
    m._give_workspace(small, 2, 64)                            # a buffer of the wrong size is never pooled
    assert all(w is big for w in m._ws_pool[(2, 64, "cpu")])
    assert m._take_workspace(2, 64, torch.device("cpu"), keep=True) is big


def test_flat_adam_state_is_shared_between_engines_and_survives_resume():
    """bind_flat_adam_state (main.py): the flat moment slabs belong to the optimizer -- a second binding (another
    batch shape's engine) gets the SAME slabs, torch's own state_dict sees the fused kernel's updates, and a resumed
    optimizer.load_state_dict() is re-bound with its values copied in."""
    from dcase2019_task4_b200.main import bind_flat_adam_state
    m = CRNN(**cfg.crnn_kwargs)
    opt = torch.optim.Adam(m.parameters(), lr=0.001, betas=(0.9, 0.999))
    n = m.flat_parameters().numel()
    dev = torch.device("cpu")
    m1, v1, steps1 = bind_flat_adam_state(opt, m._param_list, m._param_slices, n, dev)
    m1.fill_(0.25)                                             # "the fused kernel updated the moments"
    v1.fill_(0.5)
    m2, v2, steps2 = bind_flat_adam_state(opt, m._param_list, m._param_slices, n, dev)
    assert m2 is m1 and v2 is v1 and all(a is b for a, b in zip(steps1, steps2)) and len(steps1) == 38
    sd = opt.state_dict()
    assert float(sd["state"][0]["exp_avg"].mean()) == 0.25 and float(sd["state"][37]["exp_avg_sq"].mean()) == 0.5
    # resume: load_state_dict replaces the state tensors -> re-bound, values preserved
    opt2 = torch.optim.Adam(m.parameters(), lr=0.001, betas=(0.9, 0.999))
    opt2.load_state_dict(sd)
    m3, v3, _ = bind_flat_adam_state(opt2, m._param_list, m._param_slices, n, dev)
    assert m3 is not m1 and float(m3.min()) == 0.25 and float(v3.max()) == 0.5
    assert opt2.state[m._param_list[5]]["exp_avg"].data_ptr() == m3[m._param_slices[5][0]:].data_ptr()
    with pytest.raises(NotImplementedError):
        bind_flat_adam_state(torch.optim.Adam(list(m.parameters())[:3]), m._param_list, m._param_slices, n, dev)


def test_audio_tagging_results_hand_computed():
    """evaluation_measures.audio_tagging_results (evaluation_measures.py:259-296): clip-level tags from event tables."""
    from dcase2019_task4_b200 import evaluation_measures as em
    ref = pd.DataFrame([("a.wav", 0.0, 1.0, "Dog"), ("a.wav", 2.0, 3.0, "Cat"), ("b.wav", 0.0, 1.0, "Dog"),
                        ("c.wav", 1.0, 2.0, "Speech")], columns=["filename", "onset", "offset", "event_label"])
    est = pd.DataFrame([("a.wav", 0.1, 0.9, "Dog"), ("b.wav", 0.0, 1.0, "Cat"), ("b.wav", 3.0, 4.0, "Dog"),
                        ("d.wav", 0.0, 1.0, "Speech")], columns=["filename", "onset", "offset", "event_label"])
    res = em.audio_tagging_results(ref, est)
    # Dog: ref a,b  est a,b -> tp 2 -> 1.0;  Cat: ref a  est b -> fp 1 fn 1 -> 0;  Speech: ref c  est d -> fp 1 fn 1 -> 0
    assert res.to_dict() == {"Cat": 0.0, "Dog": 1.0, "Speech": 0.0}
    assert em.audio_tagging_results(ref, est.iloc[0:0]).tolist() == [0.0, 0.0, 0.0]
    weak_ref = pd.DataFrame({"filename": ["a.wav", "b.wav"], "event_labels": ["Dog,Cat", "Dog"]})
    mhe = ManyHotEncoder(["Cat", "Dog"])
    weak_ref2 = pd.DataFrame({"filename": ["a.wav", "b.wav"], "event_labels": ["Dog,Cat", "Dog"],
                              "event_label": [mhe.encode_weak(["Dog", "Cat"]), mhe.encode_weak(["Dog"])]})
    assert sorted(DatasetBook.get_classes([weak_ref])) == ["Cat", "Dog"] and weak_ref2.shape == (2, 3)


def test_dataset_bookkeeping_and_logger(tmp_path, monkeypatch):
    """tsv side of DatasetDcase2019Task4 (DatasetDcase2019Task4.py:92-181) and utils.Logger without side effects."""
    D = DatasetBook
    assert D.get_audio_dir_path_from_meta("/x/dataset/metadata/train/weak.tsv") == "/x/dataset/audio/train/weak"
    assert D.get_audio_dir_path_from_meta("/x/dataset/metadata/validation/eval_dcase2018.tsv") == "/x/dataset/audio/validation"
    tsv = tmp_path / "dataset" / "metadata" / "train" / "synthetic.tsv"
    tsv.parent.mkdir(parents=True)
    pd.DataFrame([("f%d.wav" % (i // 2), 0.5 * i, 0.5 * i + 1, "Dog") for i in range(12)],
                 columns=["filename", "onset", "offset", "event_label"]).to_csv(tsv, sep="\t", index=False)
    df = D.get_df_from_meta(str(tsv))
    assert len(df) == 12 and D.get_classes([df]) == ["Dog"]
    sub = D.get_df_from_meta(str(tsv), 3)
    assert sub.filename.nunique() == 3 and len(sub) == 6
    assert len(D.get_df_from_meta(str(tsv), 100)) == 12            # asking for more files than there are keeps all
    # features already cached -> no device work; rows of files with neither cache nor audio are dropped
    ds = D(str(tmp_path), base_feature_dir=str(tmp_path / "dataset" / "features"), save_log_feature=False)
    for i in range(5):
        np.save(os.path.join(ds.feature_dir, "f%d.npy" % i), np.zeros((3, 64), np.float32))
    out = ds.initialize_and_get_df("dataset/metadata/train/synthetic.tsv", download=False)
    assert sorted(out.filename.unique()) == ["f%d.wav" % i for i in range(5)] and len(out) == 10
    assert ds.get_feature_file("f2.wav").shape == (3, 64)
    monkeypatch.chdir(tmp_path)
    import importlib
    from dcase2019_task4_b200.utils import Logger
    importlib.reload(Logger)
    Logger.LOG.info("hello")
    assert not os.path.exists(tmp_path / "Baseline.log")


def test_properties_of_encoder_regions_and_sampler():
    """Size-independent properties (hypothesis): region decoding inverts frame encoding; every multi-stream batch holds
    exactly batch_sizes[i] distinct indices of stream i in stream order; rank-strided shards of one epoch are disjoint
    and together cover what a single process would draw from the same permutation."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.booleans(), min_size=1, max_size=120))
    def regions_roundtrip(bits):
        a = np.array(bits)
        regions = find_contiguous_regions(a)
        back = np.zeros_like(a)
        for on, off in regions:
            assert off > on
            back[on:off] = True
        assert np.array_equal(back, a)
        assert all(regions[i][1] < regions[i + 1][0] for i in range(len(regions) - 1))     # maximal, separated runs

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.integers(1, 40), min_size=1, max_size=4), st.data())
    def sampler_batches(sizes, data):
        bs = [data.draw(st.integers(1, max(1, n))) for n in sizes]

        class Src:
            cluster_indices = [range(sum(sizes[:i]), sum(sizes[:i + 1])) for i in range(len(sizes))]
        sampler = DataLoad.MultiStreamBatchSampler(Src(), bs, shuffle=True, seed=data.draw(st.integers(0, 1000)))
        batches = list(sampler)
        assert len(batches) == len(sampler) == min(n // b for n, b in zip(sizes, bs))
        seen = set()
        for batch in batches:
            assert len(batch) == sum(bs)
            pos = 0
            for i, b in enumerate(bs):
                part = batch[pos:pos + b]
                assert all(int(v) in Src.cluster_indices[i] for v in part)
                pos += b
            assert not (seen & set(int(v) for v in batch))
            seen |= set(int(v) for v in batch)

    @settings(max_examples=60, deadline=None)
    @given(st.integers(2, 4), st.integers(0, 100), st.integers(1, 60), st.integers(1, 130), st.integers(1, 7),
           st.integers(1, 17))
    def shards_partition(world, seed, n0, n1, b0, b1):
        # stream lengths NOT divisible by the world size (101 clips, batch 17, world 2 gave rank 0 an extra
        # batch and a hung all-reduce): every rank must yield exactly len(sampler) batches, shards disjoint
        class Src:
            cluster_indices = [range(0, n0), range(n0, n0 + n1)]
        drawn, counts = [], []
        for rank in range(world):
            s = DataLoad.MultiStreamBatchSampler(Src(), [b0, b1], shuffle=True, rank=rank, world_size=world, seed=seed)
            batches = list(s)
            counts.append(len(batches))
            assert len(batches) == len(s) == min(n0 // world // b0, n1 // world // b1)
            drawn.append([int(v) for batch in batches for v in batch])
        assert len(set(counts)) == 1
        flat = [v for d in drawn for v in d]
        assert len(flat) == len(set(flat))                      # disjoint across ranks

    def seedless_shards_raise():
        class Src:
            cluster_indices = [range(0, 101)]
        with pytest.raises(ValueError):
            DataLoad.MultiStreamBatchSampler(Src(), [17], rank=0, world_size=2)
        got = [len(list(DataLoad.MultiStreamBatchSampler(Src(), [17], rank=r, world_size=2, seed=3))) for r in (0, 1)]
        assert got == [2, 2]

    seedless_shards_raise()
    regions_roundtrip()
    sampler_batches()
    shards_partition()


def test_flat_slabs_keep_their_address_when_apply_changes_nothing():
    """the reference re-applies ``.cuda()`` to both models every epoch (main.py:316); captured CUDA graphs and
    the engines hold the slab addresses, so a no-op ``_apply`` must not re-allocate them."""
    m = CRNN(**cfg.crnn_kwargs)
    p0, b0 = m.flat_parameters().data_ptr(), m.flat_bn_running().data_ptr()
    m.float()
    m.to("cpu")
    m.train()
    assert m.flat_parameters().data_ptr() == p0 and m.flat_bn_running().data_ptr() == b0
    assert m.cnn.cnn.conv0.weight.data_ptr() == p0
    m.double()                                     # a real change re-flattens (fp32 slab, parameters re-viewed)
    m.float()
    assert m.flat_parameters().dtype == torch.float32
    off = 0
    for p in m.parameters():
        assert p.data_ptr() == m.flat_parameters().data_ptr() + 4 * off
        off += p.numel()
