"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): an independent, brute-force restatement of the two metrics the
reference obtains from sed_eval (not installed here; unpinned in the reference's environment.yml:23), for cross-checking
``dcase2019_task4_b200.evaluation_measures`` on small random cases.

Call sites restated: baseline/evaluation_measures.py:124-155 (event_based_evaluation_df: EventBasedMetrics with
t_collar = 0.2 s, percentage_of_length = 0.2, sed_eval's defaults evaluate_onset = evaluate_offset = True and
event_matching_type = 'optimal') and :158-182 (segment_based_evaluation_df: SegmentBasedMetrics, time_resolution = 1 s).

sed_eval's published algorithm (sed_eval/sound_event.py, sed_eval/metric.py, version 0.2.x):
  * event based, per file and per class: a reference / system event pair is a candidate hit when
        |onset_ref - onset_sys| <= t_collar   and   |offset_ref - offset_sys| <= max(t_collar, percentage_of_length * len_ref);
    'optimal' matching counts the size of a MAXIMUM bipartite matching of the candidate pairs as true positives
    (here: exhaustive search over all injective assignments, no augmenting-path code shared with the package);
  * segment based: a class is active in a segment [k r, (k + 1) r) when an event of the class overlaps it
    (floor(onset / r) .. ceil(offset / r)); true positives are segments active in both;
  * precision = Ntp / Nsys, recall = Ntp / Nref, F = 2 P R / (P + R), each 0 when its denominator is empty (sed_eval adds
    eps to the denominators); "overall" sums the counts over classes and files (micro), "class_wise_average" is the plain
    mean of the per-class scores over the label list (macro).
Parity unpinned against sed_eval itself."""
import itertools
import math


def _prf(n_tp, n_sys, n_ref):
    p = n_tp / n_sys if n_sys else 0.0
    r = n_tp / n_ref if n_ref else 0.0
    return {"f_measure": 2 * p * r / (p + r) if p + r > 0 else 0.0, "precision": p, "recall": r}


def _hit(ref, est, t_collar, percentage_of_length):
    if math.fabs(ref[0] - est[0]) > t_collar:
        return False
    return math.fabs(ref[1] - est[1]) <= max(t_collar, percentage_of_length * (ref[1] - ref[0]))


def _max_matching_bruteforce(refs, ests, t_collar, percentage_of_length):
    """Largest number of disjoint hit pairs, by trying every injective map of the smaller side into the larger one."""
    hit = [[_hit(r, e, t_collar, percentage_of_length) for e in ests] for r in refs]
    if len(refs) <= len(ests):
        small, large, get = range(len(refs)), range(len(ests)), (lambda i, j: hit[i][j])
    else:
        small, large, get = range(len(ests)), range(len(refs)), (lambda i, j: hit[j][i])
    best = 0
    for perm in itertools.permutations(large, len(small)):
        best = max(best, sum(1 for i, j in zip(small, perm) if get(i, j)))
    return best


def event_based(files, labels, t_collar=0.2, percentage_of_length=0.2):
    """files: list of (reference events, system events), an event = (label, onset, offset).  Returns the sed_eval-style
    result dict {"overall", "class_wise", "class_wise_average"}."""
    counts = {l: {"Nref": 0, "Nsys": 0, "Ntp": 0} for l in labels}
    for ref, est in files:
        for l in labels:
            r = [(on, off) for lab, on, off in ref if lab == l]
            e = [(on, off) for lab, on, off in est if lab == l]
            counts[l]["Nref"] += len(r)
            counts[l]["Nsys"] += len(e)
            counts[l]["Ntp"] += _max_matching_bruteforce(r, e, t_collar, percentage_of_length)
    return _results(counts)


def segment_based(files, labels, time_resolution=1.0):
    counts = {l: {"Nref": 0, "Nsys": 0, "Ntp": 0} for l in labels}
    for ref, est in files:
        ends = [off for _, _, off in list(ref) + list(est)]
        n_seg = int(math.ceil(max(ends) / time_resolution)) if ends else 0
        for l in labels:
            for k in range(n_seg):
                def active(events):
                    return any(lab == l and int(math.floor(on / time_resolution)) <= k < int(math.ceil(off / time_resolution))
                               for lab, on, off in events)
                a, b = active(ref), active(est)
                counts[l]["Nref"] += a
                counts[l]["Nsys"] += b
                counts[l]["Ntp"] += a and b
    return _results(counts)


def _results(counts):
    cw = {l: {"f_measure": _prf(c["Ntp"], c["Nsys"], c["Nref"]), "count": dict(c)} for l, c in counts.items()}
    tot = {k: sum(c[k] for c in counts.values()) for k in ("Nref", "Nsys", "Ntp")}
    avg = {k: (sum(v["f_measure"][k] for v in cw.values()) / len(cw) if cw else float("nan"))
           for k in ("f_measure", "precision", "recall")}
    return {"overall": {"f_measure": _prf(tot["Ntp"], tot["Nsys"], tot["Nref"]), "count": tot}, "class_wise": cw,
            "class_wise_average": {"f_measure": avg}}
