"""Error type of the integrator, mirroring cosmoTransitions/helper_functions.py:105-109.
The integrator itself (rkqs / _rkck / cubicInterpFunction, helper_functions.py:113-236,583-599)
lives on the device: csrc/ctb_device.cuh."""


class IntegrationError(Exception):
    """Used to indicate an integration error, primarily in the device-side rkqs."""
    pass


def monotonicIndices(x):
    """Indices of `x` that leave a strictly increasing subsequence between the first and the last entry (what
    helper_functions.py:57-75 computes point by point; host-side helper of evenlySpacedPhi).  Vectorised: an interior
    point survives iff it lies below the last value and above every earlier survivor, i.e. it is a record high of the
    candidates below x[-1], the first entry being the initial record."""
    import numpy as np
    x = np.asarray(x)
    flipped = x[0] > x[-1]
    y = x[::-1] if flipped else x
    last = len(y) - 1
    inner = y[1:last]
    below_end = inner < y[last]
    record = np.maximum.accumulate(np.concatenate(([y[0]], np.where(below_end, inner, -np.inf))))[:-1]
    idx = np.concatenate(([0], np.nonzero(below_end & (inner > record))[0] + 1, [last]))
    return last - idx if flipped else idx
