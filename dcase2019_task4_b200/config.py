"""Hyper-parameters of the reference (``baseline/config.py``): every module-level name the reference's code reads
through ``import config as cfg`` exists here with the same value (``tests/test_oracle_vs_reference.py`` compares the two
modules attribute by attribute).

Differences in form only: the reference derives ``classes`` by reading ``validation.tsv`` at import time
(config.py:50-51); that table is not part of this repo, so the resulting sorted list is stated directly
(SURVEY.md section 10).  Quantities the B200 kernels compile in (44.1 kHz, 2048 / 511 STFT, 64 mel bins, 64-channel
CNN, 64-cell 2-layer BiGRU, pooling (2, 4) x 3) are fixed in ``csrc/``: changing them here does not re-shape the
kernels; ``models.CRNN.CRNN`` / ``models.CNN.CNN`` raise for constructor arguments other than ``crnn_kwargs``'.
"""
import math

# ---- where the reference looks for its metadata (relative to ``workspace``) ------------------------------------
workspace = ".."
_META = "dataset/metadata"
weak, unlabel, synthetic = ("%s/train/%s.tsv" % (_META, name) for name in ("weak", "unlabel_in_domain", "synthetic"))
validation, test2018, eval2018 = ("%s/validation/%s.tsv" % (_META, name)
                                  for name in ("validation", "test_dcase2018", "eval_dcase2018"))
eval_desed = "%s/eval/public.tsv" % _META

# ---- feature extraction: what dcase_logmel_fwd implements ------------------------------------------------------
sample_rate, n_window, hop_length, n_mels = 44100, 2048, 511, 64
f_min, f_max = 0., sample_rate / 2                                    # 22050.
max_len_seconds = 10.
max_frames = math.ceil(max_len_seconds * sample_rate / hop_length)    # 864 frames of a 10-s clip

# ---- model: what dcase_crnn_forward / _backward implement -------------------------------------------------------
classes = ["Alarm_bell_ringing", "Blender", "Cat", "Dishes", "Dog", "Electric_shaver_toothbrush", "Frying",
           "Running_water", "Speech", "Vacuum_cleaner"]
_N_BLOCKS = 3
crnn_kwargs = dict(n_in_channel=1, nclass=len(classes), attention=True, n_RNN_cell=64, n_layers_RNN=2,
                   activation="glu", dropout=0.5,
                   kernel_size=_N_BLOCKS * [3], padding=_N_BLOCKS * [1], stride=_N_BLOCKS * [1],
                   nb_filters=_N_BLOCKS * [64], pooling=_N_BLOCKS * [(2, 4)])
pooling_time_ratio = 2 ** _N_BLOCKS                                   # time pooling of the three blocks: 8

# ---- training loop (main.py) -------------------------------------------------------------------------------------
batch_size = 24
n_epoch = 100
num_workers = 12
max_consistency_cost = 2          # consistency weight after the sigmoid ramp-up (main.py:127)
max_learning_rate = 0.001
median_window = 5                 # frames of the posterior median filter (evaluation_measures.py:214)
checkpoint_epochs = 1
save_best = True

# read by nothing on the path (main.py:81 has adjust_learning_rate commented out); kept because the names are public
lr, initial_lr = 0.0001, 0.
beta1_before_rampdown, beta1_after_rampdown = 0.9, 0.5
beta2_during_rampdup, beta2_after_rampup = 0.99, 0.999
weight_decay_during_rampup, weight_decay_after_rampup = 0.99, 0.999
