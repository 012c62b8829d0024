"""rbslam: B200-native Rao-Blackwellized particle filter / smoother (host side).

``particleFilter``, ``particleSmoother`` and ``particleSmootherInformationForm``
mirror the reference's MATLAB entry points; all compute runs in librbslam.so
(hand-written CUDA for sm_100a) through the C ABI in include/rbslam.h.
"""
from . import basis, synth, models
from ._capi import RbslamError, UnsupportedModelError, LIB_PATH
from .api import (Context, particleFilter, particleSmoother, particleSmootherInformationForm,
                  plan_migration, ekf_dense, particleFilterLocalization)

__all__ = ["basis", "synth", "models", "Context", "particleFilter", "particleSmoother",
           "particleSmootherInformationForm", "plan_migration", "ekf_dense", "particleFilterLocalization", "RbslamError",
           "UnsupportedModelError", "LIB_PATH"]
