"""vsrevolution_b200 -- B200-native batched evaluator for Voxel-based Soft Robots.

Drop-in for the one hot path of ericmedvet/VSREvolution: `Locomotion.apply(robot) -> Outcome` over a whole
population (reference: locomotion/Starter.java:83-94,427-435).  The compute lives in libvsr_eval.so
(hand-written CUDA for sm_100a, C ABI in include/vsr_eval.h); this package is the host-side mirror of the
reference's interface.  There is no CPU fallback.
"""
from . import _ffi  # noqa: F401

__all__ = ["_ffi"]
