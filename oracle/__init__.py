"""CPU oracle for the Rao-Blackwellized particle filter / smoother hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  The product path (``rbslam`` + ``librbslam.so``)
never routes through anything in this directory.

What it is: a NumPy fp64 restatement of the reference MATLAB algorithm
(manonkok/Rao-Blackwellized-SLAM-smoothing), following the reference loop by
loop so that every function can cite the ``file:line`` it restates.  Random
numbers are *injected* (arrays of uniforms ``U`` and standard normals ``Z``)
because MATLAB's ``randn`` stream cannot be reproduced outside MATLAB.

PARITY UNPINNED: the reference ships no tests, no golden vectors and its stored
result ``.mat`` files are absent (``.MISSING_LARGE_BLOBS``); neither MATLAB nor
GNU Octave exists in the build container, so the reference itself cannot be
executed.  The oracle is therefore pinned only indirectly, by
  * analytical identities the reference itself states (batch reduced-rank GP
    closed form ``tools/gp_scalar_potential_fast.m:190-193``; information form ==
    covariance form ``src/particleSmootherInformationForm.m:34-37``; the direct
    log-density formulas left as comments ``src/particleSmoother.m:228,285``);
  * an independent extended-precision (mpmath / longdouble) re-derivation of the
    Kalman update and log-weight (``tests/test_oracle_identities.py``);
  * the single data fixture the reference ships (``curve-x2.mat``) as input.
Indices are 0-based here; the reference is 1-based.
"""

from .tools import (sample, expq, qLeft, quat2rmat, quat2rmat_batch, qInv, logq,
                    mcross, chol_lower, domain_cartesian_dx, eigenfun,
                    eigenfun_dx, eigenval, JacobianPhi3D)
from .models import DenseMag3D, DenseRadio2D, SparseVisual2D
from .particle_filter import particleFilter
from .particle_smoother import particleSmoother
from .particle_smoother_info import particleSmootherInformationForm
from .streams import Streams, philox_uniforms_normals
from .ekf import ekf_dense
from .localization import particleFilterLocalization

__all__ = [
    "sample", "expq", "qLeft", "quat2rmat", "quat2rmat_batch", "qInv", "logq",
    "mcross", "chol_lower", "domain_cartesian_dx", "eigenfun", "eigenfun_dx",
    "eigenval", "DenseMag3D", "DenseRadio2D", "SparseVisual2D",
    "particleFilter", "particleSmoother", "particleSmootherInformationForm",
    "Streams", "philox_uniforms_normals", "ekf_dense", "particleFilterLocalization",
]
