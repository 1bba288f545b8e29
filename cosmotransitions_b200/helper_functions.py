"""Error type of the integrator, mirroring cosmoTransitions/helper_functions.py:105-109.
The integrator itself (rkqs / _rkck / cubicInterpFunction, helper_functions.py:113-236,583-599)
lives on the device: csrc/ctb_device.cuh."""


class IntegrationError(Exception):
    """Used to indicate an integration error, primarily in the device-side rkqs."""
    pass


def monotonicIndices(x):
    """
    Returns the indices of `x` such that `x[i]` is purely increasing
    (helper_functions.py:57-75; host-side helper of evenlySpacedPhi).
    """
    import numpy as np
    x = np.array(x)
    is_reversed = x[0] > x[-1]
    if is_reversed:
        x = x[::-1]
    keep = [0]
    for i in range(1, len(x) - 1):
        if x[i] > x[keep[-1]] and x[i] < x[-1]:
            keep.append(i)
    keep.append(len(x) - 1)
    keep = np.array(keep)
    return len(x) - 1 - keep if is_reversed else keep
