"""cosmotransitions_b200 -- B200-native batched bounce solver for CosmoTransitions' hot path.

Public surface (mirrors the reference's for this path):
    tunneling1D.SingleFieldInstanton, tunneling1D.PotentialError, helper_functions.IntegrationError
    finiteT.Jb_spline / Jf_spline / Jb / Jf
    generic_potential.model1 (.Vtot, .Vtot_grid, .line_potential)
    batch.find_bounce_batch (the batched entry point), batch.BouncePipeline (streamed host batches),
    batch.find_wall_batch, batch.find_action_batch, multigpu.find_bounce_batch_multi / DistributedBounce
    phases.trace_minima / find_minimum / model1_phase_scan / phase_scan / line_bounces (N-field minima and phases)
    jit.CompiledModel, generic_potential.compiled_model (a new multi-field model as a CUDA string)
    potentials.* (device functors; SampledPotential for plain Python callables)
The CUDA library is loaded lazily; there is no CPU fallback."""
from . import _lib  # noqa: F401
from .batch import find_bounce_batch, BouncePipeline  # noqa: F401

__version__ = "0.2.0"
