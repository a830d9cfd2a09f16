"""Run by tests/test_oracle_vs_reference.py in a SUBPROCESS (needs /root/reference).

Pins this package's HOST logic against the reference's own ``baseline/utils/utils.py`` and ``baseline/DataLoad.py``,
imported unmodified.  Those modules import soundfile / librosa / dcase_util at the top (absent here), so empty stub
modules stand in for soundfile / dcase_util, and ``librosa``'s three entry points (stft, feature.melspectrogram,
amplitude_to_db) are served by ``torch.stft`` and ``transformers.audio_utils`` -- independent implementations, not the
oracle -- so the reference's own call sites are what is compared with the oracle's restatement.  ``Sampler.__init__`` is patched to
accept the ``data_source`` argument the reference still passes (removed in torch >= 2.2).
The tsv bookkeeping of ``DatasetDcase2019Task4.py`` (static methods, cache naming) is compared on the metadata tables the
reference ships.  Prints ``REF-HOST-OK <n checks>``.
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/baseline"
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
import torch  # noqa: E402

from oracle import mel as omel  # noqa: E402

from transformers import audio_utils as _au  # noqa: E402   (before the stubs: transformers probes for soundfile)
import importlib.machinery  # noqa: E402

for name in ("soundfile", "dcase_util", "dcase_util.data", "sed_eval"):
    sys.modules[name] = types.ModuleType(name)
    sys.modules[name].__spec__ = importlib.machinery.ModuleSpec(name, None)


class _DecisionEncoder(object):
    """dcase_util.data.DecisionEncoder.find_contiguous_regions as published (written out here, not imported from this
    package, so the reference's decode_strong / get_predictions run over an independent statement of it)."""

    def find_contiguous_regions(self, activity_array):
        activity_array = np.asarray(activity_array).astype(bool)
        change_indices = np.logical_xor(activity_array[1:], activity_array[:-1]).nonzero()[0]
        change_indices += 1
        if activity_array[0]:
            change_indices = np.r_[0, change_indices]
        if activity_array[-1]:
            change_indices = np.r_[change_indices, activity_array.size]
        return change_indices.reshape((-1, 2))


class _ProbabilityEncoder(object):
    """dcase_util.data.ProbabilityEncoder.binarization, 'global_threshold' branch: strictly greater than the threshold."""

    def binarization(self, probabilities, binarization_type="global_threshold", threshold=0.5, time_axis=1):
        assert binarization_type == "global_threshold"
        return np.array(np.asarray(probabilities) > threshold, dtype=int)


sys.modules["dcase_util.data"].DecisionEncoder = _DecisionEncoder
sys.modules["dcase_util.data"].ProbabilityEncoder = _ProbabilityEncoder
# librosa stand-in built on transformers.audio_utils (an independent numpy implementation written upstream to reproduce
# librosa) and torch.stft -- NOT on the oracle -- so that the reference's own call sites (which arguments, which
# transposes and casts: DatasetDcase2019Task4.py:209-231, DataLoad.py:203-207) are what gets compared with the oracle


def _stft(y, n_fft, hop_length, window, center=True, pad_mode="reflect"):
    out = torch.stft(torch.as_tensor(np.asarray(y, dtype=np.float64)), n_fft=n_fft, hop_length=hop_length,
                     window=torch.as_tensor(np.asarray(window, dtype=np.float64)), center=center, pad_mode=pad_mode,
                     return_complex=True)
    return out.numpy()


def _melspectrogram(S, sr, n_mels, fmin, fmax, htk=False, norm=None):
    fb = _au.mel_filter_bank(num_frequency_bins=S.shape[0], num_mel_filters=n_mels, min_frequency=fmin, max_frequency=fmax,
                             sampling_rate=sr, norm=norm, mel_scale="htk" if htk else "slaney")
    return np.dot(fb.T.astype(np.float32), S)            # librosa stores the basis as float32


librosa = types.ModuleType("librosa")
librosa.feature = types.ModuleType("librosa.feature")
librosa.stft = _stft
librosa.feature.melspectrogram = _melspectrogram
librosa.amplitude_to_db = lambda S: _au.amplitude_to_db(np.asarray(S), reference=1.0, min_value=1e-5, db_range=80.0)
sys.modules["librosa"] = librosa
sys.modules["librosa.feature"] = librosa.feature
torch.utils.data.sampler.Sampler.__init__ = lambda self, *a, **k: None

sys.path.insert(1, REF)
import DataLoad as ref_dl  # noqa: E402          (the reference's)
from utils import utils as ref_utils  # noqa: E402
from utils.Scaler import Scaler as RefScaler  # noqa: E402
assert ref_dl.__file__.startswith(REF) and ref_utils.__file__.startswith(REF)

from dcase2019_task4_b200 import DataLoad as our_dl  # noqa: E402
from dcase2019_task4_b200.utils import utils as our_utils  # noqa: E402

checks = 0
CLASSES = ["Alarm_bell_ringing", "Blender", "Cat", "Dishes", "Dog", "Electric_shaver_toothbrush", "Frying",
           "Running_water", "Speech", "Vacuum_cleaner"]

# ---- ManyHotEncoder (utils/utils.py:22-172): the [108, 10] target layout ----
ref_enc, our_enc = ref_utils.ManyHotEncoder(CLASSES, n_frames=108), our_utils.ManyHotEncoder(CLASSES, n_frames=108)
strong_df = pd.DataFrame({"onset": [3.0, 50.0, 0.0], "offset": [9.0, 108.0, 20.5], "event_label": ["Dog", "Cat", "Speech"]})
cases = [strong_df, strong_df.iloc[0], ["Dog", "Speech"], [["Dog", 2, 6], ["Cat", 0, 20]], "empty", [],
         pd.Series(["Blender"]), pd.DataFrame({"onset": [np.nan], "offset": [np.nan], "event_label": [np.nan]})]
for c in cases:
    assert np.array_equal(ref_enc.encode_strong_df(c), our_enc.encode_strong_df(c)), c
    checks += 1
for c in (["Dog", "Cat"], "empty", [], pd.Series(["Speech", np.nan]), strong_df):
    assert np.array_equal(ref_enc.encode_weak(c), our_enc.encode_weak(c)), c
    checks += 1
assert ref_enc.decode_weak(ref_enc.encode_weak(["Dog", "Cat"])) == our_enc.decode_weak(our_enc.encode_weak(["Dog", "Cat"]))
assert ref_enc.state_dict() == our_enc.state_dict()
checks += 2

# ---- DataLoadDf (DataLoad.py:25-154) + ConcatDataset (:383-439) ----
rng = np.random.default_rng(0)
feat = {("f%d.wav" % i): rng.random((7, 64)).astype(np.float32) for i in range(12)}
weak_df = pd.DataFrame({"filename": ["f0.wav", "f1.wav", "f2.wav"], "event_labels": ["Dog,Cat", "Speech", "Dishes"]})
unl_df = pd.DataFrame({"filename": ["f3.wav", "f4.wav", "f5.wav", "f6.wav"]})
syn_df = pd.DataFrame({"filename": ["f7.wav", "f7.wav", "f8.wav", "f9.wav", "f9.wav"], "onset": [1.0, 30.0, 5.0, 0.0, 60.0],
                       "offset": [9.0, 50.0, 80.0, 10.0, 100.0], "event_label": ["Dog", "Cat", "Speech", "Frying", "Dog"]})
sets = {}
for tag, mod, enc in (("ref", ref_dl, ref_enc), ("our", our_dl, our_enc)):
    sets[tag] = [mod.DataLoadDf(df, lambda f: feat[f], enc.encode_strong_df) for df in (weak_df, unl_df, syn_df)]
for a, b in zip(sets["ref"], sets["our"]):
    assert len(a) == len(b) and list(a.filenames) == list(b.filenames)
    for i in range(len(a)):
        (fa, ya), (fb, yb) = a[i], b[i]
        assert np.array_equal(fa, fb) and np.array_equal(ya, yb), (i, ya.sum(), yb.sum())
        checks += 1
cat_ref, cat_our = ref_dl.ConcatDataset(sets["ref"]), our_dl.ConcatDataset(sets["our"])
assert len(cat_ref) == len(cat_our) == 10
assert [list(r) for r in cat_ref.cluster_indices] == [list(r) for r in cat_our.cluster_indices]
for i in range(10):
    assert np.array_equal(cat_ref[i][1], cat_our[i][1])
checks += 3

# ---- MultiStreamBatchSampler (:539-577): same numpy global RNG stream -> same batches ----
for shuffle in (False, True):
    np.random.seed(11)
    ref_batches = [tuple(int(v) for v in b) for b in ref_dl.MultiStreamBatchSampler(cat_ref, [1, 2, 1], shuffle=shuffle)]
    np.random.seed(11)
    our_sampler = our_dl.MultiStreamBatchSampler(cat_our, [1, 2, 1], shuffle=shuffle)
    our_batches = [tuple(int(v) for v in b) for b in our_sampler]
    assert ref_batches == our_batches and len(our_batches) == len(our_sampler) == 2, (ref_batches, our_batches)
    checks += 1

# ---- transform chain (DataLoad.py:189-350 through get_transforms, utils/utils.py:397-412) vs the oracle chain ----
ref_scaler = RefScaler()
ref_scaler.load_state_dict({"mean_": np.linspace(-40, -10, 64).tolist(),
                            "mean_of_square_": (np.linspace(-40, -10, 64) ** 2 + np.linspace(50, 200, 64)).tolist()})
amp = (np.abs(rng.normal(0, 1, (87, 64))) * 30).astype(np.float32)
label = np.zeros((108, 10))
for frames in (96, 64):
    chain = ref_utils.get_transforms(frames, ref_scaler, augment_type="noise")
    np.random.seed(3)
    x, x_noisy, y = chain((amp, label))
    np.random.seed(3)
    noise = np.abs(np.random.normal(0, 0.5 ** 2, amp.shape))
    c, n = omel.transform_chain(amp, ref_scaler.mean_, ref_scaler.std_, noise=noise, frames=frames)
    assert tuple(x.shape) == (1, frames, 64) and x.dtype == torch.float32 and y.dtype == torch.float32
    assert np.abs(x.numpy() - c).max() <= 1e-6 and np.abs(x_noisy.numpy() - n).max() <= 1e-6
    plain = ref_utils.get_transforms(frames)((amp, label))
    # raw dB in float32: 20 log10(max(amin, x)) (stand-in) vs 10 log10(max(amin^2, x^2)) (librosa's form, oracle): 1-2 ulp at ~40 dB
    assert len(plain) == 2 and np.abs(plain[0].numpy() - omel.transform_chain(amp, None, None, frames=frames)[0]).max() <= 2e-5
    ours = our_utils.get_transforms(frames, ref_scaler, augment_type="noise")
    assert [type(t).__name__ for t in ours.transforms] == [type(t).__name__ for t in chain.transforms]
    checks += 3

# ---- small utilities: SaveBest, AverageMeterSet (utils/utils.py:242-394) ----
for comp in ("inf", "sup"):
    a, b = ref_utils.SaveBest(comp), our_utils.SaveBest(comp)
    for v in (0.3, 0.5, 0.1, 0.1, 0.7):
        assert a.apply(v) == b.apply(v)
    assert (a.best_val, a.best_epoch) == (b.best_val, b.best_epoch)
    checks += 1
ma, mb = ref_utils.AverageMeterSet(), our_utils.AverageMeterSet()
for k, v in (("Loss", 2.0), ("Loss", 4.0), ("lr", 0.001)):
    ma.update(k, v)
    mb.update(k, v)
assert str(ma) == str(mb) and ma.averages() == mb.averages() and ma.sums() == mb.sums()
checks += 1

# ---- pure helpers of evaluation_measures.py (:85-122, :183-199); the sed_eval / dcase_util users are restated ----
import evaluation_measures as ref_em  # noqa: E402          (the reference's, with the stubs above)
from dcase2019_task4_b200 import evaluation_measures as our_em  # noqa: E402
assert ref_em.__file__.startswith(REF)
ra = rng.integers(0, 2, (40, 10))
rb = rng.integers(0, 2, (40, 10))
for x, y in zip(ref_em.intermediate_at_measures(ra, rb), our_em.intermediate_at_measures(ra, rb)):
    assert np.array_equal(x, y)
tp, fp, fn, _ = ref_em.intermediate_at_measures(ra, rb)
assert np.array_equal(ref_em.macro_f_measure(tp, fp, fn), our_em.macro_f_measure(tp, fp, fn))
events = pd.DataFrame([("a.wav", 0.0, 1.0, "Dog"), ("a.wav", 2.0, 3.0, "Cat"), ("b.wav", np.nan, np.nan, np.nan)],
                      columns=["filename", "onset", "offset", "event_label"])
assert ref_em.get_event_list_current_file(events, "a.wav") == our_em.get_event_list_current_file(events, "a.wav")
# a file whose single row has no label: the reference returns [{"filename": ...}] (a label-less entry sed_eval ignores),
# ours returns [] -- the same "no events" for the restated metrics
assert ref_em.get_event_list_current_file(events, "b.wav") == [{"filename": "b.wav"}]
assert our_em.get_event_list_current_file(events, "b.wav") == []
checks += 4
try:                                  # pandas >= 2.2 drops the grouping column inside groupby.apply: the reference breaks
    ref_em.audio_tagging_results(events.dropna(), events.dropna())
    ref_tagging = "works"
except KeyError:
    ref_tagging = "KeyError"
print("reference audio_tagging_results under this pandas:", ref_tagging)

# ---- weights_init (utils/utils.py:205-224) applied to THIS package's CRNN through module.apply (main.py:282) ----
from dcase2019_task4_b200 import config as our_cfg  # noqa: E402
from dcase2019_task4_b200.models.CRNN import CRNN as OurCRNN  # noqa: E402
torch.manual_seed(123)
m_ref_init = OurCRNN(**our_cfg.crnn_kwargs)
torch.manual_seed(7)
m_ref_init.apply(ref_utils.weights_init)                 # the reference's function visits our sub-modules by class name
torch.manual_seed(123)
m_our_init = OurCRNN(**our_cfg.crnn_kwargs)
torch.manual_seed(7)
m_our_init.apply(our_utils.weights_init)
for (k, a), (_, b) in zip(m_ref_init.named_parameters(), m_our_init.named_parameters()):
    assert torch.equal(a, b), k                          # same generator consumption, same values
pr = dict(m_ref_init.named_parameters())
assert float(pr["cnn.cnn.conv1.bias"].abs().max()) == 0.0 and abs(float(pr["cnn.cnn.batchnorm1.weight"].mean()) - 1) < 0.02
w = pr["rnn.rnn.weight_hh_l0"].detach()
assert float((w.t() @ w - torch.eye(64)).abs().max()) < 1e-5          # orthogonal init of the GRU matrices
assert 0.008 < float(pr["dense.weight"].std()) < 0.012 and float(pr["dense_softmax.bias"].abs().max()) == 0.0
checks += 2

# ---- get_predictions (evaluation_measures.py:203-231): the reference's clip-by-clip loop vs the batched restatement ----
if not hasattr(pd.DataFrame, "append"):                  # removed in pandas 2: shim for the reference
    pd.DataFrame.append = lambda self, other: pd.concat([self, other])


class _StubModel(torch.nn.Module):                       # posteriors are a fixed function of the input, CPU only
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))

    def forward(self, x):                                # [B, 1, T, 64] -> strong [B, T/8, 10], weak [B, 10]
        s = torch.sigmoid(3.0 * x[:, 0].reshape(x.shape[0], x.shape[2] // 8, 8, 64)[:, :, :, :10].mean(2))
        return s, s.mean(1)


class _ValidSet(list):
    pass


T_valid = 136
valid = _ValidSet((torch.randn(1, T_valid, 64, generator=torch.Generator().manual_seed(40 + i)), None) for i in range(7))
valid.filenames = pd.Series(["clip%d.wav" % i for i in range(7)])
stub = _StubModel().eval()
ref_dec = ref_utils.ManyHotEncoder(CLASSES, n_frames=T_valid // 8)
our_dec = our_utils.ManyHotEncoder(CLASSES, n_frames=T_valid // 8)
with torch.no_grad():
    ref_pred = ref_em.get_predictions(stub, valid, ref_dec.decode_strong, pooling_time_ratio=8)
    our_pred = our_em.get_predictions(stub, valid, our_dec.decode_strong, pooling_time_ratio=8, batch_size=3)
key = ["filename", "event_label", "onset"]
a = ref_pred.sort_values(key).reset_index(drop=True)
b = our_pred.sort_values(key).reset_index(drop=True)
assert len(a) == len(b) and len(a) > 10, (len(a), len(b))
assert list(a.event_label) == list(b.event_label) and list(a.filename) == list(b.filename)
assert np.allclose(a.onset.to_numpy(float), b.onset.to_numpy(float), rtol=0, atol=1e-12)
assert np.allclose(a.offset.to_numpy(float), b.offset.to_numpy(float), rtol=0, atol=1e-12)
for lab in ([1, 0, 0, 1, 1, 0, 1], [0, 0, 0], [1, 1, 1]):
    col = np.zeros((len(lab), 10))
    col[:, 4] = lab
    assert [list(map(int, r[1:])) for r in ref_dec.decode_strong(col)] == \
        [list(map(int, r[1:])) for r in our_dec.decode_strong(col)]
checks += 5

# ---- get_f_measure_by_class (evaluation_measures.py:19-82), global threshold, weak and frame-level labels ----
gq = torch.Generator().manual_seed(77)
loader_weak = [(torch.randn(6, 1, 64, 64, generator=gq), (torch.rand(6, 10, generator=gq) < 0.4).float()) for _ in range(3)]
loader_strong = [(x, (torch.rand(6, 8, 10, generator=gq) < 0.1).float()) for x, _ in loader_weak]
for loader in (loader_weak, loader_strong):
    with torch.no_grad():
        f_ref = ref_em.get_f_measure_by_class(stub, 10, loader)
        f_our = our_em.get_f_measure_by_class(stub, 10, loader)
    assert np.allclose(f_ref, f_our, rtol=0, atol=1e-12) and f_ref.max() > 0
    checks += 1

# ---- tsv bookkeeping of DatasetDcase2019Task4.py (:92-181), static methods only (no audio here) ----
sys.modules["download_data"] = types.ModuleType("download_data")
sys.modules["download_data"].download = lambda *a, **k: None
cwd = os.getcwd()
import DatasetDcase2019Task4 as ref_ds_mod  # noqa: E402   (reference; utils.Logger opens Baseline.log in the CWD = tmp)
from dcase2019_task4_b200.DatasetDcase2019Task4 import DatasetDcase2019Task4 as OurDS  # noqa: E402
RefDS = ref_ds_mod.DatasetDcase2019Task4
assert ref_ds_mod.__file__.startswith(REF)
meta = os.path.join(os.path.dirname(REF), "dataset", "metadata")
for rel in ("train/weak.tsv", "train/synthetic.tsv", "train/unlabel_in_domain.tsv", "validation/validation.tsv",
            "validation/eval_dcase2018.tsv"):
    path = os.path.join(meta, rel)
    assert RefDS.get_audio_dir_path_from_meta(path) == OurDS.get_audio_dir_path_from_meta(path), rel
    a, b = RefDS.get_df_from_meta(path), OurDS.get_df_from_meta(path)
    assert a.equals(b)
    a, b = RefDS.get_df_from_meta(path, 25), OurDS.get_df_from_meta(path, 25)
    assert a.equals(b) and a.filename.nunique() == 25
    checks += 3
dfs = [RefDS.get_df_from_meta(os.path.join(meta, "train/weak.tsv")), RefDS.get_df_from_meta(os.path.join(meta, "train/synthetic.tsv"))]
assert sorted(RefDS.get_classes(dfs)) == sorted(OurDS.get_classes(dfs)) == CLASSES
checks += 1
ref_ds = RefDS(cwd, base_feature_dir=os.path.join(cwd, "features_ref"), save_log_feature=False)
our_ds = OurDS(cwd, base_feature_dir=os.path.join(cwd, "features_our"), save_log_feature=False)
assert os.path.relpath(ref_ds.feature_dir, os.path.join(cwd, "features_ref")) == \
    os.path.relpath(our_ds.feature_dir, os.path.join(cwd, "features_our"))          # cache directory naming
# the reference's own calculate_mel_spec (its window, centring, padding, transposition, float32 cast) over the
# transformers / torch.stft stand-in, against the oracle's restatement
from dcase2019_task4_b200 import synth  # noqa: E402
clips, _ = synth.make_clips(2, seed=9, n_samples=30000)
for wv in clips:
    ref_mel = ref_ds.calculate_mel_spec(wv.astype(np.float64))
    ora_mel = omel.calculate_mel_spec(wv.astype(np.float64))
    assert ref_mel.dtype == np.float32 and ref_mel.shape == ora_mel.shape == (1 + 30000 // 511, 64)
    assert np.abs(ref_mel - ora_mel).max() <= 2e-7 * ora_mel.max() + 1e-6
    checks += 1
np.save(os.path.join(ref_ds.feature_dir, "clip.npy"), amp)
np.save(os.path.join(our_ds.feature_dir, "clip.npy"), amp)
assert np.array_equal(ref_ds.get_feature_file("clip.wav"), our_ds.get_feature_file("clip.wav"))
checks += 2

print("REF-HOST-OK %d" % checks)
