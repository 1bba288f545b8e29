"""Error type of the integrator, mirroring cosmoTransitions/helper_functions.py:105-109.
The integrator itself (rkqs / _rkck / cubicInterpFunction, helper_functions.py:113-236,583-599)
lives on the device: csrc/ctb_device.cuh."""


class IntegrationError(Exception):
    """Used to indicate an integration error, primarily in the device-side rkqs."""
    pass
