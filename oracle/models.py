"""Oracle restatement of the three model-closure families the reference passes
into its inference engines as function handles.  Test infrastructure only.

Each model exposes
  dynModel(xn, dx, dt, Q, z)        -> xpred          (z: injected N(0,1) draws)
  measModel(xn[n x N])              -> dy [N x d x M]  (dense families)
  measModel_sparse(xn_i, xl_i)      -> (yhat, dy)      (sparse family)
  dynResNorm(xnk, xni, dx, dt, Q)   -> row vector      (or None => default form)
  nz                                   normals consumed per dynModel call
"""
import numpy as np
from .tools import (expq, qLeft, qInv, logq, quat2rmat, eigenfun, eigenfun_dx)


def _chol(A):
    return np.linalg.cholesky(np.atleast_2d(np.asarray(A, dtype=np.float64)))


class DenseMag3D:
    """6-D pose + magnetic-field potential map.

    examples/slam-dense-mag/run_dense3D_magfield.m: dynModel :301-308,
    measModel :265-279, dynResNorm :202-203.  State xn = [pos(3); quat(4)].
    """
    family = "dense_mag3d"
    n = 7
    d = 3
    nz = 6
    sparse = False

    def __init__(self, NN, L):
        self.NN = np.asarray(NN, dtype=np.float64)
        self.L = np.asarray(L, dtype=np.float64).reshape(3)
        self.M = self.NN.shape[0] + 3

    def dynModel(self, xn, dx, dt, Q, z):
        # :304 position: pos + dx(1:3)' + chol(dt*Q(1:3,1:3),'lower')*randn(3,1)
        xpred_pos = xn[0:3] + dx[0:3] + _chol(dt * Q[0:3, 0:3]) @ z[0:3]
        # :305 dQuat = qLeft(dx(4:7)') * expq(chol(dt*Q(4:6,4:6),'lower')*randn(3,1))
        dQuat = qLeft(dx[3:7]) @ expq(_chol(dt * Q[3:6, 3:6]) @ z[3:6])
        # :306 q+ = qLeft(q) * dQuat  (no renormalisation)
        xpred_quat = qLeft(xn[3:7]) @ dQuat
        return np.concatenate([xpred_pos, xpred_quat])

    def measModel(self, xn):
        xn = np.asarray(xn, dtype=np.float64).reshape(7, -1)
        N = xn.shape[1]
        pos = xn[0:3, :].T
        # :267-272
        dPhix = np.hstack([np.ones((N, 1)), np.zeros((N, 2)), eigenfun_dx(self.NN, pos, 0, self.L)])
        dPhiy = np.hstack([np.zeros((N, 1)), np.ones((N, 1)), np.zeros((N, 1)),
                           eigenfun_dx(self.NN, pos, 1, self.L)])
        dPhiz = np.hstack([np.zeros((N, 2)), np.ones((N, 1)), eigenfun_dx(self.NN, pos, 2, self.L)])
        dy = np.zeros((N, 3, self.M))
        for i in range(N):  # :275-278
            Rnb = quat2rmat(xn[3:7, i])
            dy[i] = Rnb.T @ np.vstack([dPhix[i], dPhiy[i], dPhiz[i]])
        return dy

    def dynResNorm(self, xnk, xni, dx, dt, Q):
        # :202-203
        r_pos = xnk[0:3] - xni[0:3] - dx[0:3]
        qrel = qLeft(qLeft(qInv(dx[3:7])) @ qInv(xni[3:7])) @ xnk[3:7]
        r = np.concatenate([r_pos, logq(qrel)])
        # row-vector / lower-triangular L  ==  solve  x L = r  ==  L' x' = r'
        Lc = _chol(dt * Q)
        return np.linalg.solve(Lc.T, r)


class DenseRadio2D:
    """2-D position + heading, scalar RSS field.

    examples/slam-dense-radio/run_dense2D_withHeading.m: dynModel :75-76 (= :89-90),
    dynResNorm :77 (= :91), measModel :168.
    """
    family = "dense_radio2d"
    n = 3
    d = 1
    nz = 1
    sparse = False

    def __init__(self, NN, L):
        self.NN = np.asarray(NN, dtype=np.float64)
        self.L = np.asarray(L, dtype=np.float64).reshape(2)
        self.M = self.NN.shape[0]

    def dynModel(self, xn, dx, dt, Q, z):
        c, s = np.cos(xn[2]), np.sin(xn[2])
        Rot = np.array([[c, -s], [s, c]])
        pos = xn[0:2] + Rot.T @ dx[0:2]
        th = xn[2] + dx[2] + (_chol(dt * Q) @ np.atleast_1d(z[0:1]))[0]
        return np.array([pos[0], pos[1], th])

    def measModel(self, xn):
        xn = np.asarray(xn, dtype=np.float64).reshape(3, -1)
        Phi = eigenfun(self.NN, xn[0:2, :].T, self.L)  # [N x M]
        return Phi[:, None, :]

    def dynResNorm(self, xnk, xni, dx, dt, Q):
        Lc = _chol(dt * Q)
        return np.atleast_1d((xnk[2] - xni[2] - dx[2]) / Lc[0, 0])


class SparseVisual2D:
    """2-D pose, 1-D pinhole camera, point-landmark map (conditionally linearised).

    examples/slam-sparse-visual/pfslam.m:81-82 (= psslam.m:91-92),
    measurement.m:32-84.  dynResNorm is [] in the reference (psslam.m:118) so the
    engines use their default form.
    """
    family = "sparse_visual2d"
    n = 3
    nz = 3
    sparse = True
    dynResNorm = None

    def __init__(self, n_landmarks, f=1.5, fp=0.0, fw=1.0):
        self.nl = int(n_landmarks)
        self.M = 2 * self.nl
        self.d = self.nl
        self.f, self.fp, self.fw = float(f), float(fp), float(fw)

    def dynModel(self, xn, dx, dt, Q, z):
        # xn + dx' + sqrt(dt*Q)*randn(3,1)   (element-wise sqrt of the matrix)
        return xn + dx + np.sqrt(dt * Q) @ z[0:3]

    def measModel_sparse(self, xn, xl):
        f, fp = self.f, self.fp
        p = xn[0:2]
        th = xn[2]
        c, s = np.cos(th), np.sin(th)
        R = np.array([[c, -s], [s, c]])
        mp = np.asarray(xl, dtype=np.float64).reshape(-1, 2).T  # 2 x nl (column-major reshape)
        K = np.array([[f, fp], [0.0, 1.0]])
        A = K @ np.hstack([R.T, -(R.T @ p).reshape(2, 1)])
        u = A @ np.vstack([mp, np.ones((1, mp.shape[1]))])
        y = u[0] / u[1]
        div = (mp[1] * c - p[1] * c - mp[0] * s + p[0] * s) ** 2
        dym1 = (f * (mp[1] - p[1])) / div
        dym2 = -(f * (mp[0] - p[0])) / div
        dy = np.zeros((self.nl, self.M))
        idx = np.arange(self.nl)
        dy[idx, 2 * idx] = dym1
        dy[idx, 2 * idx + 1] = dym2
        return y, dy
