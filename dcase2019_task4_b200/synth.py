"""Seeded synthetic 10-second clips with strong labels.

There is no audio in the reference tree and no network, so benchmarks and tests use these clips:
0.1*N(0,1) noise floor + 1-3 sine / chirp bursts (random onset, duration >= 250 ms, f in [100, 8000] Hz,
amplitude U(0.05, 0.5)); the burst intervals and class ids are the strong labels; every 8th clip has its
second half zeroed (exercises ``amin`` and ``top_db`` of amplitude_to_db).
"""
import numpy as np

SAMPLE_RATE = 44100
CLIP_SAMPLES = 441000  # 10 s -> exactly 864 frames at hop 511
N_CLASSES = 10
# baseline/config.py:50-51 (sorted unique event_label of validation.tsv)
CLASSES = ["Alarm_bell_ringing", "Blender", "Cat", "Dishes", "Dog", "Electric_shaver_toothbrush", "Frying",
           "Running_water", "Speech", "Vacuum_cleaner"]


def make_clip(rng, n_samples=CLIP_SAMPLES, silent_half=False):
    """Returns (float32 waveform [n_samples], list of (class_id, onset_s, offset_s))."""
    t = np.arange(n_samples, dtype=np.float64) / SAMPLE_RATE
    y = 0.1 * rng.standard_normal(n_samples)
    dur_total = n_samples / SAMPLE_RATE
    events = []
    for _ in range(int(rng.integers(1, 4))):
        dur = float(rng.uniform(min(0.25, dur_total / 2), max(min(0.25, dur_total / 2), dur_total / 2)))
        onset = float(rng.uniform(0.0, dur_total - dur))
        f0 = float(rng.uniform(100.0, 8000.0))
        f1 = f0 if rng.random() < 0.5 else float(rng.uniform(100.0, 8000.0))
        amp = float(rng.uniform(0.05, 0.5))
        cls = int(rng.integers(0, N_CLASSES))
        m = (t >= onset) & (t < onset + dur)
        tt = t[m] - onset
        phase = 2 * np.pi * (f0 * tt + 0.5 * (f1 - f0) / dur * tt * tt)
        y[m] += amp * np.sin(phase)
        events.append((cls, onset, onset + dur))
    if silent_half:
        y[n_samples // 2:] = 0.0
    return y.astype(np.float32), events


def make_clips(n, seed=0, n_samples=CLIP_SAMPLES):
    rng = np.random.default_rng(seed)
    waves = np.empty((n, n_samples), dtype=np.float32)
    events = []
    for i in range(n):
        waves[i], ev = make_clip(rng, n_samples, silent_half=(i % 8 == 7))
        events.append(ev)
    return waves, events


def encode_strong(events, n_frames_out, pooling_time_ratio=8, hop=511):
    """ManyHotEncoder.encode_strong_df (utils/utils.py:69-128) with the frame conversion of
    main.py:227-228: onset * sr // hop // pooling_time_ratio."""
    y = np.zeros((n_frames_out, N_CLASSES), dtype=np.float32)
    for cls, on, off in events:
        a = int(on * SAMPLE_RATE // hop // pooling_time_ratio)
        b = int(off * SAMPLE_RATE // hop // pooling_time_ratio)
        y[a:b, cls] = 1
    return y


def make_targets(events, batch_sizes, n_frames_out=108):
    """Targets for one batch laid out as main.py:240-247: [weak | unlabeled | strong] streams.

    weak clips: event present on all frames (encode_strong_df on a list of labels, utils.py:105-111);
    unlabeled: all -1 (utils.py:82-85); strong: frame-level many-hot."""
    B = sum(batch_sizes)
    assert len(events) == B
    tgt = np.zeros((B, n_frames_out, N_CLASSES), dtype=np.float32)
    n_weak = batch_sizes[0]
    n_unl = batch_sizes[1] if len(batch_sizes) > 1 else 0
    for i in range(B):
        if i < n_weak:
            for cls, _, _ in events[i]:
                tgt[i, :, cls] = 1
        elif i < n_weak + n_unl:
            tgt[i] = -1
        else:
            tgt[i] = encode_strong(events[i], n_frames_out)
    return tgt
