"""Model registry: the host-side stand-in for the MATLAB function handles.

The reference passes ``dynModel`` / ``measModel`` / ``dynResNorm`` closures into
its engines (src/particleFilter.m:108,124,129; src/particleSmoother.m:179).
CUDA cannot call back into the host, so each supported closure family is a
*descriptor*: a family id plus the constants the closure captured.  The
descriptor's ``dynModel`` / ``measModel`` / ``dynResNorm`` attributes are handle
objects that the drop-in entry points recognise; any other callable raises
``UnsupportedModelError`` (rbslam:unsupportedModel) -- there is no CPU fallback.
"""
import numpy as np

from . import _capi


class ModelHandle:
    """What a MATLAB function handle carries across the gateway: (model, role)."""

    def __init__(self, model, role):
        self.model = model
        self.role = role

    def __call__(self, *a, **k):   # handles are descriptors, not host code
        raise _capi.UnsupportedModelError(
            _capi.EMODEL, "model handles are evaluated on the GPU; they cannot be called on the host")

    def __repr__(self):
        return "<%s.%s>" % (type(self.model).__name__, self.role)


class _Model:
    family = 0
    sparse = False

    def __init__(self):
        self.dynModel = ModelHandle(self, "dynModel")
        self.measModel = ModelHandle(self, "measModel")
        self.dynResNorm = ModelHandle(self, "dynResNorm")


class DenseMag3D(_Model):
    """6-D pose + magnetic-potential map; closures of
    examples/slam-dense-mag/run_dense3D_magfield.m:265-279,301-308,202-203.
    ``NN`` [m x 3] eigenfunction indices, ``L`` [3] half-widths of the domain."""
    family = _capi.MODEL_DENSE_MAG3D
    n, d, nz, nw, n_odo = 7, 3, 6, 6, 7

    def __init__(self, NN, L):
        super().__init__()
        self.NN = np.asfortranarray(np.asarray(NN, dtype=np.int32))
        self.L = np.ascontiguousarray(np.asarray(L, dtype=np.float64).reshape(3))
        if self.NN.ndim != 2 or self.NN.shape[1] != 3:
            raise ValueError("NN must be [m x 3]")
        self.m_basis = self.NN.shape[0]
        self.M = self.m_basis + 3


class DenseRadio2D(_Model):
    """2-D position + heading, scalar field; closures of
    examples/slam-dense-radio/run_dense2D_withHeading.m:75-77,168."""
    family = _capi.MODEL_DENSE_RADIO2D
    n, d, nz, nw, n_odo = 3, 1, 1, 1, 3

    def __init__(self, NN, L):
        super().__init__()
        self.NN = np.asfortranarray(np.asarray(NN, dtype=np.int32))
        self.L = np.ascontiguousarray(np.asarray(L, dtype=np.float64).reshape(2))
        if self.NN.ndim != 2 or self.NN.shape[1] != 2:
            raise ValueError("NN must be [m x 2]")
        self.m_basis = self.NN.shape[0]
        self.M = self.m_basis


class SparseVisual2D(_Model):
    """2-D pose, 1-D pinhole camera, point landmarks; closures of
    examples/slam-sparse-visual/pfslam.m:81-82 and measurement.m:32-84.
    The reference passes dynResNorm=[] for this family (psslam.m:118)."""
    family = _capi.MODEL_SPARSE_VISUAL2D
    n, nz, nw, n_odo = 3, 3, 3, 3
    sparse = True

    def __init__(self, n_landmarks, f=1.5, fp=0.0, fw=1.0):
        super().__init__()
        self.m_basis = int(n_landmarks)
        self.d = self.m_basis
        self.M = 2 * self.m_basis
        self.NN = None
        self.L = None
        self.camera = (float(f), float(fp), float(fw))
        self.dynResNorm = None


def from_problem(pr):
    """Model descriptor for a problem dict made by rbslam.synth."""
    fam = pr["family"]
    if fam == "dense_mag3d":
        return DenseMag3D(pr["NN"], pr["L"])
    if fam == "dense_radio2d":
        return DenseRadio2D(pr["NN"], pr["L"])
    if fam == "sparse_visual2d":
        return SparseVisual2D(pr["n_landmarks"], *pr["camera"])
    raise _capi.UnsupportedModelError(_capi.EMODEL, "unknown family %r" % (fam,))


def resolve(dynModel, measModel, dynResNorm=None):
    """Map the handle arguments of the reference signature to one model descriptor."""
    handles = [h for h in (dynModel, measModel, dynResNorm) if h is not None]
    for h in handles:
        if isinstance(h, _Model):
            continue
        if not isinstance(h, ModelHandle):
            raise _capi.UnsupportedModelError(
                _capi.EMODEL,
                "rbslam:unsupportedModel: %r is not a registered model handle (dense-mag 3D, "
                "dense-radio 2D, sparse-visual 2D); arbitrary host closures cannot run on the GPU "
                "and there is no CPU fallback" % (h,))
    models = {id(h if isinstance(h, _Model) else h.model): (h if isinstance(h, _Model) else h.model)
              for h in handles}
    if len(models) != 1:
        raise _capi.UnsupportedModelError(_capi.EMODEL,
                                          "dynModel/measModel/dynResNorm must come from one model")
    return next(iter(models.values()))
