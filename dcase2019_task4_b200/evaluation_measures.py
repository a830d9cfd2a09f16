"""From frame posteriors to the event-based F1 the north star names.

Mirrors the part of the reference's ``baseline/evaluation_measures.py`` that sits directly behind the hot path:

* ``get_predictions``  (evaluation_measures.py:203-231)  posterior -> 0.5 threshold -> median filter (5 frames) ->
  contiguous regions -> onset / offset in seconds.  The reference runs the model clip by clip (batch 1); here the clips
  go through the B200 forward in batches and the post-processing is vectorised over the batch.  ``DataFrame.append``
  (removed in pandas 2) is replaced by one ``pd.concat``.
* ``event_based_evaluation_df`` / ``segment_based_evaluation_df`` / ``compute_strong_metrics``
  (evaluation_measures.py:124-182, :234-246).  The reference delegates to ``sed_eval`` (unpinned in its environment.yml
  and not installed here), so the two metrics are RESTATED from sed_eval's published definitions:
    - event based: a system event is a true positive if its label matches a reference event of the same file, its onset
      lies within ``t_collar`` of the reference onset and its offset within ``max(t_collar, percentage_of_length *
      reference length)`` of the reference offset; reference and system events are paired by a maximum bipartite
      matching (sed_eval's ``event_matching_type='optimal'``), files are evaluated one by one and the counts accumulated;
    - segment based: activity per class on a ``time_resolution`` grid, intermediate statistics accumulated over files.
  PARITY UNPINNED against sed_eval itself (no copy to run, no golden vectors upstream); ``tests/test_host_logic.py``
  pins the definitions on hand-computed cases.

* ``get_f_measure_by_class`` / ``intermediate_at_measures`` / ``macro_f_measure`` (evaluation_measures.py:19-122,
  :183-199): the per-class weak F-measure main.py logs every epoch (dcase_util's global / class threshold binarisation
  restated as ``>``).

Host-side code: the only device work is the batched model forward.
"""
import math

import numpy as np
import pandas as pd
import scipy.ndimage

from . import config as cfg


# ------------------------------------------------------------------------------------------------------------
# posteriors -> events
# ------------------------------------------------------------------------------------------------------------
def postprocess_posteriors(strong, threshold=0.5, median_window=None):
    """[..., T, C] posteriors -> boolean activity of the same shape: global threshold (dcase_util
    ``binarization_type='global_threshold'``, evaluation_measures.py:212-213) then a median filter of
    ``median_window`` frames along time (``scipy.ndimage.median_filter(size=(window, 1))``, :214)."""
    if median_window is None:
        median_window = cfg.median_window
    a = (np.asarray(strong) > threshold).astype(np.float32)
    size = [1] * a.ndim
    size[-2] = median_window
    return scipy.ndimage.median_filter(a, size=tuple(size)) > 0.5


def frames_to_seconds(frames, pooling_time_ratio=1):
    """evaluation_measures.py:226-227."""
    return frames * pooling_time_ratio / (cfg.sample_rate / cfg.hop_length)


def get_predictions(model, valid_dataset, decoder, pooling_time_ratio=1, save_predictions=None, batch_size=24):
    """Same contract as the reference's ``get_predictions``: a DataFrame with ``event_label, onset, offset, filename``
    (times in seconds) for every clip of ``valid_dataset`` (items ``(input [1, T, 64], label)``; ``.filenames``)."""
    import torch

    device = model.flat_parameters().device if hasattr(model, "flat_parameters") else next(model.parameters()).device
    frames = []
    n = len(valid_dataset)
    with torch.no_grad():
        for start in range(0, n, batch_size):
            items = [valid_dataset[i][0] for i in range(start, min(start + batch_size, n))]
            x = torch.stack([torch.as_tensor(it) for it in items]).to(device)
            strong, _ = model(x)
            active = postprocess_posteriors(strong.float().cpu().numpy())
            for k in range(active.shape[0]):
                pred = pd.DataFrame(decoder(active[k]), columns=["event_label", "onset", "offset"])
                pred["filename"] = valid_dataset.filenames.iloc[start + k]
                frames.append(pred)
    prediction_df = (pd.concat(frames, ignore_index=True) if frames
                     else pd.DataFrame(columns=["event_label", "onset", "offset", "filename"]))
    prediction_df["onset"] = frames_to_seconds(prediction_df.onset.astype(float), pooling_time_ratio)
    prediction_df["offset"] = frames_to_seconds(prediction_df.offset.astype(float), pooling_time_ratio)
    if save_predictions is not None:
        prediction_df.to_csv(save_predictions, index=False, sep="\t")
    return prediction_df


# ------------------------------------------------------------------------------------------------------------
# weak (audio tagging) F-measure per class, evaluation_measures.py:19-122, :183-199
# ------------------------------------------------------------------------------------------------------------
def intermediate_at_measures(encoded_ref, encoded_est):
    """(tp, fp, fn, tn) per class of two 0/1 arrays [n, classes] (evaluation_measures.py:85-101)."""
    tp = (encoded_est + encoded_ref == 2).sum(axis=0)
    fp = (encoded_est - encoded_ref == 1).sum(axis=0)
    fn = (encoded_ref - encoded_est == 1).sum(axis=0)
    tn = (encoded_est + encoded_ref == 0).sum(axis=0)
    return tp, fp, fn, tn


def macro_f_measure(tp, fp, fn):
    """2 tp / (2 tp + fp + fn) per class, 0 where the denominator is empty (evaluation_measures.py:183-199)."""
    tp, fp, fn = np.asarray(tp, dtype=float), np.asarray(fp, dtype=float), np.asarray(fn, dtype=float)
    out = np.zeros(tp.shape[-1])
    mask = 2 * tp + fp + fn != 0
    out[mask] = 2 * tp[mask] / (2 * tp + fp + fn)[mask]
    return out


def get_f_measure_by_class(torch_model, nb_tags, dataloader_, thresholds_=None):
    """The "Valid weak metric" of main.py:326-331: F-measure per class of the clip-level (weak) predictions over a
    loader of ``(batch_x, y)``.  Same conventions as the reference: frame-level labels / predictions are reduced by
    the maximum over time, probabilities are binarised with ``> 0.5`` (dcase_util ``global_threshold``) or per-class
    thresholds.  The forward runs on the model's device; the counting is host arithmetic on [batch, classes]."""
    import torch

    device = (torch_model.flat_parameters().device if hasattr(torch_model, "flat_parameters")
              else next(torch_model.parameters()).device)
    tp, tn, fp, fn = (np.zeros(nb_tags) for _ in range(4))
    if thresholds_ is not None:
        assert type(thresholds_) is list
    thresh = 0.5 if thresholds_ is None else np.asarray(thresholds_, dtype=float)[None, :]
    with torch.no_grad():
        for batch_x, y in dataloader_:
            _, pred_weak = torch_model(batch_x.to(device))
            pred_weak = pred_weak.float().cpu().numpy()
            labels = y.cpu().numpy() if isinstance(y, torch.Tensor) else np.asarray(y)
            if pred_weak.ndim == 3:                      # a model predicting only strong outputs
                pred_weak = np.max(pred_weak, axis=1)
            if labels.ndim == 3:
                labels = (np.max(labels, axis=1) > 0.5).astype(int)
            batch_predictions = (pred_weak > thresh).astype(int)
            tp_, fp_, fn_, tn_ = intermediate_at_measures(labels, batch_predictions)
            tp += tp_
            fp += fp_
            fn += fn_
            tn += tn_
    return macro_f_measure(tp, fp, fn)


def format_df(df, mhe):
    """One row per file with the many-hot weak encoding of its event labels (evaluation_measures.py:249-257)."""
    if "onset" in df.columns or "offset" in df.columns:
        rows = [{"filename": fname, "event_label": mhe.encode_weak(g["event_label"].drop_duplicates().dropna().tolist())}
                for fname, g in df.groupby("filename", sort=True)]
        df = pd.DataFrame(rows, columns=["filename", "event_label"])
    return df


def audio_tagging_results(reference, estimated):
    """Per-class F-measure of the clip-level tags implied by two event DataFrames (evaluation_measures.py:259-296):
    files are matched by an outer join, a file missing on one side counts as "no tags" there."""
    from .utils.utils import ManyHotEncoder
    if "event_label" in reference.columns:
        classes = sorted(set(reference.event_label.dropna().unique()) | set(estimated.event_label.dropna().unique()))
        mhe = ManyHotEncoder(classes)
        reference, estimated = format_df(reference, mhe), format_df(estimated, mhe)
    else:
        def split(df):
            return set(df.event_labels.str.split(',', expand=True).unstack().dropna().unique())
        classes = sorted(split(reference) | split(estimated))
        mhe = ManyHotEncoder(classes)
    if estimated.empty:
        return pd.Series(np.zeros(len(classes)), index=mhe.labels)
    matching = reference.merge(estimated, how='outer', on="filename", suffixes=["_ref", "_pred"])

    def tags(val):
        return val if isinstance(val, np.ndarray) else np.zeros(len(classes))
    ref = np.array([tags(v) for v in matching.event_label_ref])
    est = np.array([tags(v) for v in matching.event_label_pred])
    tp, fp, fn, _ = intermediate_at_measures(ref, est)
    return pd.Series(macro_f_measure(tp, fp, fn), index=mhe.labels)


# ------------------------------------------------------------------------------------------------------------
# restated sed_eval metrics
# ------------------------------------------------------------------------------------------------------------
def get_event_list_current_file(df, fname):
    """evaluation_measures.py:86-100: the events of one file as dicts; a file whose only row has no label is empty."""
    rows = df[df["filename"] == fname]
    if len(rows) == 1 and pd.isna(rows["event_label"].iloc[0]):
        return []
    return rows.dropna(subset=["event_label"]).to_dict("records")


def _f_measure(n_tp, n_sys, n_ref):
    """sed_eval.metric.precision / recall / f_measure are regularised with eps, i.e. an empty denominator gives 0 (this
    is also the reference's ``empty_system_output_handling='zero_score'``): nothing is undefined, every class counts
    in the class-wise average."""
    precision = n_tp / n_sys if n_sys > 0 else 0.0
    recall = n_tp / n_ref if n_ref > 0 else 0.0
    f = 2 * precision * recall / (precision + recall) if precision + recall > 0 else 0.0
    return {"f_measure": f, "precision": precision, "recall": recall}


def _max_bipartite_matching(adj, n_right):
    """Size of a maximum matching; adj[i] = right vertices the left vertex i may pair with (augmenting paths)."""
    match_right = [-1] * n_right

    def try_assign(i, seen):
        for j in adj[i]:
            if not seen[j]:
                seen[j] = True
                if match_right[j] < 0 or try_assign(match_right[j], seen):
                    match_right[j] = i
                    return True
        return False

    return sum(1 for i in range(len(adj)) if try_assign(i, [False] * n_right))


class _Metrics(object):
    """Per-class counters + the overall / class-wise-average views sed_eval reports."""

    def __init__(self, event_label_list):
        self.event_label_list = sorted(event_label_list)
        self.class_wise = {label: {"Nref": 0, "Nsys": 0, "Ntp": 0} for label in self.event_label_list}

    def results_class_wise_metrics(self):
        return {label: {"f_measure": _f_measure(c["Ntp"], c["Nsys"], c["Nref"]), "count": dict(c)}
                for label, c in self.class_wise.items()}

    def results_overall_metrics(self):
        n_tp = sum(c["Ntp"] for c in self.class_wise.values())
        n_sys = sum(c["Nsys"] for c in self.class_wise.values())
        n_ref = sum(c["Nref"] for c in self.class_wise.values())
        return {"f_measure": _f_measure(n_tp, n_sys, n_ref), "count": {"Nref": n_ref, "Nsys": n_sys, "Ntp": n_tp}}

    def results_class_wise_average_metrics(self):
        """Macro average over the classes (the union of the labels of both DataFrames)."""
        per_class = [v["f_measure"] for v in self.results_class_wise_metrics().values()]
        out = {}
        for key in ("f_measure", "precision", "recall"):
            vals = [m[key] for m in per_class if not math.isnan(m[key])]
            out[key] = float(np.mean(vals)) if vals else float("nan")
        return {"f_measure": out}

    def results(self):
        return {"overall": self.results_overall_metrics(), "class_wise": self.results_class_wise_metrics(),
                "class_wise_average": self.results_class_wise_average_metrics()}

    def __str__(self):
        r = self.results()
        lines = ["%s" % type(self).__name__,
                 "  overall (micro)      F %.2f %%  P %.2f %%  R %.2f %%" % tuple(
                     100 * r["overall"]["f_measure"][k] for k in ("f_measure", "precision", "recall")),
                 "  class-wise average   F %.2f %%  P %.2f %%  R %.2f %%" % tuple(
                     100 * r["class_wise_average"]["f_measure"][k] for k in ("f_measure", "precision", "recall"))]
        for label, v in r["class_wise"].items():
            lines.append("    %-28s Nref %5d  Nsys %5d  F %.2f %%" % (label, v["count"]["Nref"], v["count"]["Nsys"],
                                                                     100 * v["f_measure"]["f_measure"]))
        return "\n".join(lines)


class EventBasedMetrics(_Metrics):
    """sed_eval.sound_event.EventBasedMetrics restated (onset + offset evaluated, optimal matching)."""

    def __init__(self, event_label_list, t_collar=0.200, percentage_of_length=0.2):
        super(EventBasedMetrics, self).__init__(event_label_list)
        self.t_collar = t_collar
        self.percentage_of_length = percentage_of_length

    def _hit(self, ref, est):
        if abs(ref["onset"] - est["onset"]) > self.t_collar:
            return False
        length = ref["offset"] - ref["onset"]
        return abs(ref["offset"] - est["offset"]) <= max(self.t_collar, self.percentage_of_length * length)

    def evaluate(self, reference_event_list, estimated_event_list):
        for label in self.event_label_list:
            refs = [e for e in reference_event_list if e["event_label"] == label]
            ests = [e for e in estimated_event_list if e["event_label"] == label]
            adj = [[j for j, e in enumerate(ests) if self._hit(r, e)] for r in refs]
            c = self.class_wise[label]
            c["Nref"] += len(refs)
            c["Nsys"] += len(ests)
            c["Ntp"] += _max_bipartite_matching(adj, len(ests))
        return self


class SegmentBasedMetrics(_Metrics):
    """sed_eval.sound_event.SegmentBasedMetrics restated: per class, a segment is active if any event overlaps it."""

    def __init__(self, event_label_list, time_resolution=1.0):
        super(SegmentBasedMetrics, self).__init__(event_label_list)
        self.time_resolution = time_resolution

    def _roll(self, events, label, n_segments):
        active = np.zeros(n_segments, dtype=bool)
        for e in events:
            if e["event_label"] == label:
                lo = int(math.floor(e["onset"] / self.time_resolution))
                hi = int(math.ceil(e["offset"] / self.time_resolution))
                active[lo:max(hi, lo)] = True
        return active

    def evaluate(self, reference_event_list, estimated_event_list):
        ends = [e["offset"] for e in list(reference_event_list) + list(estimated_event_list)]
        n_segments = int(math.ceil(max(ends) / self.time_resolution)) if ends else 0
        for label in self.event_label_list:
            ref = self._roll(reference_event_list, label, n_segments)
            est = self._roll(estimated_event_list, label, n_segments)
            c = self.class_wise[label]
            c["Nref"] += int(ref.sum())
            c["Nsys"] += int(est.sum())
            c["Ntp"] += int((ref & est).sum())
        return self


def _classes(reference, estimated):
    classes = list(reference.event_label.dropna().unique()) + list(estimated.event_label.dropna().unique())
    return sorted(set(classes))


def event_based_evaluation_df(reference, estimated, t_collar=0.200, percentage_of_length=0.2):
    """evaluation_measures.py:124-155: file-by-file event-based metric of two event DataFrames."""
    metric = EventBasedMetrics(_classes(reference, estimated), t_collar=t_collar, percentage_of_length=percentage_of_length)
    for fname in reference["filename"].unique():
        metric.evaluate(get_event_list_current_file(reference, fname), get_event_list_current_file(estimated, fname))
    return metric


def segment_based_evaluation_df(reference, estimated, time_resolution=1.):
    """evaluation_measures.py:158-182."""
    metric = SegmentBasedMetrics(_classes(reference, estimated), time_resolution=time_resolution)
    for fname in reference["filename"].unique():
        metric.evaluate(get_event_list_current_file(reference, fname), get_event_list_current_file(estimated, fname))
    return metric


def compute_strong_metrics(predictions, valid_df, pooling_time_ratio=None, log=None):
    """evaluation_measures.py:234-246: event-based (200 ms collar, 20 % offset) and 1-s segment-based metrics."""
    if pooling_time_ratio is not None:
        predictions = predictions.copy()
        predictions["onset"] = frames_to_seconds(predictions.onset, pooling_time_ratio)
        predictions["offset"] = frames_to_seconds(predictions.offset, pooling_time_ratio)
    metric_event = event_based_evaluation_df(valid_df, predictions, t_collar=0.200, percentage_of_length=0.2)
    metric_segment = segment_based_evaluation_df(valid_df, predictions, time_resolution=1.)
    if log is not None:
        log.info(metric_event)
        log.info(metric_segment)
    return metric_event
