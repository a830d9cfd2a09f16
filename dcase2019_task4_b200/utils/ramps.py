"""Consistency ramp-up used by the mean-teacher step (baseline/utils/ramps.py:20-27).

Only ``sigmoid_rampup`` is on the reference's path (main.py:75-78); the other ramps of that file are dead code
there and are not rebuilt."""
import math


def sigmoid_rampup(current, rampup_length):
    """exp(-5 (1 - t/T)^2) with t clipped to [0, T]; 1.0 when T == 0."""
    if rampup_length == 0:
        return 1.0
    t = min(max(float(current), 0.0), float(rampup_length))
    phase = 1.0 - t / rampup_length
    return float(math.exp(-5.0 * phase * phase))
