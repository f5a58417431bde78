"""art_planner_b200 -- B200-native (sm_100a CUDA) implementation of art_planner's batchable hot path:
pose validity (ODE box-vs-heightfield torso/feet checks), edge validity over interpolated SE(3) states and
edge cost, behind a C ABI (include/artp.h) and a host-side mirror of the reference's plugin interface."""
from . import synth  # noqa: F401
from .capi import ArtpError  # noqa: F401
from .checker import (MotionCostObjective, MotionValidator, PathLengthObjective, SE3FromSE2Sampler,  # noqa: F401
                      StateValidityChecker)

__all__ = ["synth", "ArtpError", "StateValidityChecker", "MotionValidator", "PathLengthObjective", "MotionCostObjective", "SE3FromSE2Sampler"]
