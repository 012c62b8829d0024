"""Host-side mirror of the reference's three inference entry points.

Same names, positional arguments and outputs as
  src/particleFilter.m:1-3, src/particleSmoother.m:1-2,
  src/particleSmootherInformationForm.m:1-2
(the reference's toolchain, MATLAB, is absent from the build image, so the host
side above the C ABI is Python here; ``matlab/*.m`` + ``mex/rbslam_mex.cpp`` are
the MATLAB-side twins).  All arithmetic happens in librbslam.so on the GPU.

Differences a caller sees:
  * ``dynModel`` / ``measModel`` / ``dynResNorm`` are model handles from
    ``rbslam.models`` (descriptor of one of the reference's closure families);
  * randomness is explicit: ``rng=<int seed>`` selects the device Philox stream,
    ``rng=<object with U, Z, Uend arrays>`` injects pre-drawn numbers (what the
    MATLAB shim does in compat mode so that MATLAB's own rand/randn are used).
"""
import ctypes as C

import numpy as np

from . import _capi
from . import models as _models


def _as_streams(rng):
    if rng is None:
        return _capi.RNG_PHILOX, 0, None
    if isinstance(rng, (int, np.integer)):
        return _capi.RNG_PHILOX, int(rng), None
    if hasattr(rng, "U") and hasattr(rng, "Z"):
        return _capi.RNG_INJECTED, 0, rng
    raise TypeError("rng must be an int seed or an object with U/Z/Uend arrays")


class Context:
    """Owns one rbslam_ctx (one GPU)."""

    def __init__(self, model, N, T, device=0, rng_mode=_capi.RNG_PHILOX, seed=0,
                 information_form=False, keep_history=True, ld=0, kalman_variant=0, rank=0, world=1,
                 devices=None, replicas=False):
        """devices=[0, 1, ...]: ONE filter sharded over several GPUs, driven from this process
        (rbslam_create_group); filter entry points with the device RNG only.
        devices=[...], replicas=True: full replicas for the smoothers, the ancestor weights split
        over the devices (rbslam_create_replicas)."""
        self._lib = _capi.lib()
        self.model = model
        self.N, self.T = int(N), int(T)
        cfg = _capi.Config()
        cfg.struct_size = C.sizeof(_capi.Config)
        cfg.device = device
        cfg.model = model.family
        cfg.N, cfg.T, cfg.m_basis = self.N, self.T, model.m_basis
        self._keep = [model.NN, model.L]
        cfg.NN = _capi.iptr(model.NN)
        cfg.L = _capi.dptr(model.L)
        cam = getattr(model, "camera", (0.0, 0.0, 0.0))
        cfg.cam_f, cfg.cam_fp, cfg.cam_fw = cam
        cfg.rng_mode = rng_mode
        cfg.seed = seed
        cfg.ld = ld
        cfg.information_form = int(information_form)
        cfg.keep_history = int(keep_history)
        cfg.rank, cfg.world = rank, world
        self.rank, self.world = rank, world
        cfg.kalman_variant = kalman_variant
        self._h = C.c_void_p()
        if devices is not None and len(devices) > 1:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            create = self._lib.rbslam_create_replicas if replicas else self._lib.rbslam_create_group
            rc = create(C.byref(self._h), C.byref(cfg), _capi.iptr(dv), dv.shape[0])
        else:
            if devices is not None and len(devices) == 1:
                cfg.device = int(devices[0])
            rc = self._lib.rbslam_create(C.byref(self._h), C.byref(cfg))
        if rc != _capi.OK:
            msg = self._lib.rbslam_last_error(None) or b""
            self._h = C.c_void_p()
            err = _capi.UnsupportedModelError if rc == _capi.EMODEL else _capi.RbslamError
            raise err(rc, msg.decode())
        dims = (C.c_int32 * 7)()
        self._lib.rbslam_dims(self._h, dims)
        self.n, self.d, self.M, self.nz, self.nw, self.n_odo, self.ld = list(dims)
        self._cb = None

    # -- life cycle ---------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.rbslam_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        _capi.check(self._h, rc)

    # -- inputs -------------------------------------------------------------
    def _inputs(self, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, dt, streams, K=1,
                forced_ancestors=None, forced_ak=None):
        keep = []

        def F(a):
            a = _capi.fcol(a)
            keep.append(a)
            return a

        y = np.asarray(y, dtype=np.float64)
        if y.ndim == 1:
            y = y.reshape(-1, 1)
        T = y.shape[0]
        if y.shape[1] != self.d:
            raise ValueError("y must be [N_T x %d]" % self.d)
        inp = _capi.Inputs()
        inp.T = T
        odometry = np.atleast_2d(np.asarray(odometry, dtype=np.float64))
        if T > 1 and (odometry.shape[1] != self.n_odo or odometry.shape[0] < T - 1):
            raise ValueError("odometry must be [>=N_T-1 x %d]" % self.n_odo)
        odo = F(odometry)
        inp.odometry, inp.odo_rows = _capi.dptr(odo), odo.shape[0]
        inp.y = _capi.dptr(F(y))
        x0n = F(np.asarray(x0_nonLin, dtype=np.float64).reshape(-1))
        if x0n.shape[0] != self.n:
            raise ValueError("x0_nonLin must have %d entries" % self.n)
        inp.x0_nonLin = _capi.dptr(x0n)
        x0l = np.asarray(x0_lin, dtype=np.float64)
        if x0l.ndim == 1:
            x0l = x0l.reshape(-1, 1)
        if x0l.shape[0] != self.M or x0l.shape[1] not in (1, self.N):
            raise ValueError("x0_lin must be [%d x 1] or [%d x %d]" % (self.M, self.M, self.N))
        inp.x0_lin, inp.x0_lin_cols = _capi.dptr(F(x0l)), x0l.shape[1]
        P0 = np.asarray(P0_lin, dtype=np.float64)
        if P0.shape != (self.M, self.M):
            raise ValueError("P0_lin must be [%d x %d]" % (self.M, self.M))
        inp.P0_lin = _capi.dptr(F(P0))
        Q = np.asarray(Q, dtype=np.float64)
        if Q.ndim < 2:
            Q = Q.reshape(1, 1)
        if Q.ndim == 2:
            Q = Q[:, :, None]
        if Q.shape[0] != self.nw or Q.shape[1] != self.nw:
            raise ValueError("Q must be [%d x %d (x pages)]" % (self.nw, self.nw))
        inp.Q, inp.Q_pages = _capi.dptr(F(Q)), Q.shape[2]
        R = np.atleast_2d(np.asarray(R, dtype=np.float64))
        if R.shape != (self.d, self.d):
            raise ValueError("R must be [%d x %d]" % (self.d, self.d))
        inp.R = _capi.dptr(F(R))
        dtv = F(np.asarray(dt, dtype=np.float64).reshape(-1))
        inp.dt, inp.dt_len = _capi.dptr(dtv), dtv.shape[0]
        if streams is not None:
            U = np.ascontiguousarray(np.asarray(streams.U, dtype=np.float64))      # [K,T,N]
            Z = np.ascontiguousarray(np.asarray(streams.Z, dtype=np.float64))      # [K,T,N,nz]
            if U.shape[0] < K or U.shape[1:] != (T, self.N):
                raise ValueError("streams.U must be [>=%d, %d, %d]" % (K, T, self.N))
            if Z.shape[0] < K or Z.shape[1:] != (T, self.N, self.nz):
                raise ValueError("streams.Z must be [>=%d, %d, %d, %d]" % (K, T, self.N, self.nz))
            keep += [U, Z]
            inp.U, inp.Z = _capi.dptr(U), _capi.dptr(Z)
            Uend = getattr(streams, "Uend", None)
            if Uend is not None:
                Uend = np.ascontiguousarray(np.asarray(Uend, dtype=np.float64))
                keep.append(Uend)
                inp.Uend = _capi.dptr(Uend)
        if forced_ancestors is not None:
            fa = np.ascontiguousarray(np.asarray(forced_ancestors, dtype=np.int32))
            if fa.ndim == 2:
                fa = fa[None]
            if fa.shape[0] < K or fa.shape[1:] != (T, self.N):
                raise ValueError("forced_ancestors must be [K, T, N]")
            keep.append(fa)
            inp.forced_ancestors = _capi.iptr(fa)
        if forced_ak is not None:
            fk = np.ascontiguousarray(np.asarray(forced_ak, dtype=np.int32).reshape(-1))
            keep.append(fk)
            inp.forced_ak = _capi.iptr(fk)
        return inp, keep, T

    # -- filter -------------------------------------------------------------
    def _filter_outputs(self, T, want_xn_traj=True, taps=False):
        n, N, M = self.n, self.N, self.M
        o = dict(traj_max=np.zeros((n, T), order="F"), traj_mean=np.zeros((n, T), order="F"),
                 xl_max=np.zeros(M), xl_mean=np.zeros(M), P_max=np.zeros((M, M), order="F"),
                 P_mean=np.zeros((M, M), order="F"),
                 traj_sample_iwmax=np.zeros((n, T), order="F"))
        if want_xn_traj:
            o["xn_traj"] = np.zeros((n, N, T), order="F")
        if taps:
            o["logw_hist"] = np.zeros((N, T), order="F")
            o["w_hist"] = np.zeros((N, T), order="F")
            o["ancestors"] = np.zeros((N, T), dtype=np.int32, order="F")
        out = _capi.FilterOutputs()
        for k, v in o.items():
            setattr(out, k, _capi.iptr(v) if v.dtype == np.int32 else _capi.dptr(v))
        return out, o

    def filter_run(self, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, dt, streams=None,
                   forced_ancestors=None, want_xn_traj=True, taps=False):
        inp, keep, T = self._inputs(odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, dt, streams,
                                    forced_ancestors=forced_ancestors)
        out, o = self._filter_outputs(T, want_xn_traj, taps)
        self._run_T = T
        self._ck(self._lib.rbslam_filter_run(self._h, C.byref(inp), C.byref(out)))
        return o

    def filter_begin(self, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, dt, streams=None,
                     forced_ancestors=None):
        inp, keep, T = self._inputs(odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, dt, streams,
                                    forced_ancestors=forced_ancestors)
        self._ck(self._lib.rbslam_filter_begin(self._h, C.byref(inp)))
        self._run_T = T

    def filter_step(self):
        self._ck(self._lib.rbslam_filter_step(self._h))

    def filter_end(self, T=None, want_xn_traj=False):
        out, o = self._filter_outputs(self._run_T if T is None else T, want_xn_traj, False)
        o.pop("traj_sample_iwmax")
        out.traj_sample_iwmax = None
        self._ck(self._lib.rbslam_filter_end(self._h, C.byref(out)))
        return o

    def sync(self):
        self._ck(self._lib.rbslam_sync(self._h))

    def read_particles(self, P=True):
        n, N, M = self.n, self.N, self.M
        xn = np.zeros((n, N), order="F")
        xl = np.zeros((M, N), order="F")
        Pm = np.zeros((M, M, N), order="F") if P else None
        logw, w = np.zeros(N), np.zeros(N)
        ai = np.zeros(N, dtype=np.int32)
        self._ck(self._lib.rbslam_read_particles(self._h, _capi.dptr(xn), _capi.dptr(xl),
                                                 _capi.dptr(Pm), _capi.dptr(logw), _capi.dptr(w),
                                                 _capi.iptr(ai)))
        return dict(xn=xn, xl=xl, P=Pm, logw=logw, w=w, ai=ai)

    def read_trajectories(self, xn_traj=True):
        """traj_max, traj_mean, yhattraj, xn_traj as the reference holds them at the current step
        (NaN / zero beyond it): the trajectory arguments of the filter's makePlots hook."""
        n, N, d, T = self.n, self.N, self.d, self._run_T
        tm, tmean = np.zeros((n, T), order="F"), np.zeros((n, T), order="F")
        yh = np.zeros((d, T), order="F")
        xt = np.zeros((n, N, T), order="F") if xn_traj else None
        self._ck(self._lib.rbslam_read_trajectories(self._h, _capi.dptr(tm), _capi.dptr(tmean), _capi.dptr(yh),
                                                    _capi.dptr(xt)))
        return dict(traj_max=tm, traj_mean=tmean, yhattraj=yh, xn_traj=xt)

    def read_information(self):
        N, M = self.N, self.M
        ivec = np.zeros((M, N), order="F")
        Imat = np.zeros((M, M, N), order="F")
        hld = np.zeros(N)
        self._ck(self._lib.rbslam_read_information(self._h, _capi.dptr(ivec), _capi.dptr(Imat),
                                                   _capi.dptr(hld)))
        return dict(ivec=ivec, Imat=Imat, halfLogDetP=hld)

    def set_step_callback(self, fn):
        """fn(sweep, t) is called after every time step (the makePlots hook)."""
        if fn is None:
            self._cb = _capi.STEP_FN()
        else:
            self._cb = _capi.STEP_FN(lambda user, k, t: fn(k, t))
        self._ck(self._lib.rbslam_step_callback(self._h, self._cb, None))

    # -- smoother -----------------------------------------------------------
    def smoother_run(self, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, dt, N_K, form=0,
                     streams=None, forced_ancestors=None, forced_ak=None, want_AI=False):
        inp, keep, T = self._inputs(odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, dt, streams,
                                    K=N_K, forced_ancestors=forced_ancestors, forced_ak=forced_ak)
        n, N, M = self.n, self.N, self.M
        o = dict(XNK=np.zeros((n, T, N_K), order="F"), XLK=np.zeros((M, N_K), order="F"),
                 PK=np.zeros((M, M, N_K), order="F"), ak=np.zeros(N_K, dtype=np.int32))
        if want_AI:
            o["AI"] = np.full((N, T, N_K), np.nan, order="F")
        out = _capi.SmootherOutputs()
        out.XNK, out.XLK, out.PK = _capi.dptr(o["XNK"]), _capi.dptr(o["XLK"]), _capi.dptr(o["PK"])
        out.ak = _capi.iptr(o["ak"])
        out.AI = _capi.dptr(o.get("AI"))
        self._smoother_out = o
        self._ck(self._lib.rbslam_smoother_run(self._h, C.byref(inp), N_K, form, C.byref(out)))
        return o

    # -- sharding (one process per GPU) --------------------------------------
    def ipc_export(self, which):
        buf = C.create_string_buffer(64)
        self._ck(self._lib.rbslam_ipc_export(self._h, which, buf))
        return buf.raw

    def ipc_import(self, peer, which, handle):
        self._ck(self._lib.rbslam_ipc_import(self._h, peer, which, C.create_string_buffer(handle, 64)))

    def connect_peers(self, all_gather_object):
        """Exchange the CUDA-IPC handles of the shared buffers with every peer.
        ``all_gather_object(obj) -> list`` is any host collective (torch.distributed)."""
        n = self._lib.rbslam_ipc_count()
        mine = [self.ipc_export(w) for w in range(n)]
        everyone = all_gather_object(mine)
        for r, hs in enumerate(everyone):
            if r != self.rank:
                for w, h in enumerate(hs):
                    self.ipc_import(r, w, h)

    # -- timing / counters ---------------------------------------------------
    def event_record(self, slot):
        self._ck(self._lib.rbslam_event_record(self._h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._ck(self._lib.rbslam_event_elapsed(self._h, a, b, C.byref(ms)))
        return float(ms.value)

    def phase_timing(self, enable):
        self._ck(self._lib.rbslam_phase_timing(self._h, int(enable)))

    def phase_times(self):
        ms = np.zeros(8)
        self._ck(self._lib.rbslam_phase_times(self._h, _capi.dptr(ms)))
        return dict(zip(["resample", "propagate", "meas", "kalman", "normalize", "ancestor",
                         "info", "reserved"], ms.tolist()))

    def counters(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self._lib.rbslam_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(kernel_launches=a.value, h2d_bytes=b.value, d2h_bytes=c.value)

    def status_counters(self):
        """Device status word: jitter retries, clamped draws, resampling steps that needed the exact scan."""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._ck(self._lib.rbslam_status_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(used_jitter=a.value, clamped_draws=b.value, exact_scan_runs=c.value)

    # -- kernel-level ops (parity tests) --------------------------------------
    def op_resample(self, w, u):
        w = np.ascontiguousarray(w, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        ai = np.zeros(u.shape[0], dtype=np.int32)
        self._ck(self._lib.rbslam_op_resample(self._h, w.shape[0], _capi.dptr(w), u.shape[0],
                                              _capi.dptr(u), _capi.iptr(ai)))
        return ai

    def op_normalize(self, logw):
        logw = np.ascontiguousarray(logw, dtype=np.float64)
        w = np.zeros_like(logw)
        im = C.c_int32()
        self._ck(self._lib.rbslam_op_normalize(self._h, logw.shape[0], _capi.dptr(logw),
                                               _capi.dptr(w), C.byref(im)))
        return w, int(im.value)

    def op_propagate(self, xn_in, ai, dx, dt, Q, Z=None):
        xn_in = _capi.fcol(xn_in)
        N = xn_in.shape[1]
        ai = np.ascontiguousarray(ai, dtype=np.int32)
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        Q = _capi.fcol(np.atleast_2d(Q))
        Zc = None if Z is None else np.ascontiguousarray(Z, dtype=np.float64)   # [N, nz]
        out = np.zeros_like(xn_in, order="F")
        self._ck(self._lib.rbslam_op_propagate(self._h, N, _capi.dptr(xn_in), _capi.iptr(ai),
                                               _capi.dptr(dx), float(dt), _capi.dptr(Q),
                                               _capi.dptr(Zc), _capi.dptr(out)))
        return out

    def op_meas_jacobian(self, xn, xl=None):
        xn = _capi.fcol(xn)
        N = xn.shape[1]
        xlc = None if xl is None else _capi.fcol(xl)
        dy = np.zeros((N, self.d, self.M), order="F")
        yhat = np.zeros((self.d, N), order="F")
        self._ck(self._lib.rbslam_op_meas_jacobian(self._h, N, _capi.dptr(xn), _capi.dptr(xlc),
                                                   _capi.dptr(dy), _capi.dptr(yhat)))
        return dy, yhat

    def op_jacobian_phi3d(self, x, xl, xu, yl, yu, zl, zu):
        """JacobianPhi3D(x, N_m, xl, xu, yl, yu, zl, zu, Indices) (tools/JacobianPhi3D.m:1) with
        N_m / Indices taken from the context's basis: x [3 x N] -> J [3 x 3 x m x N]."""
        x = _capi.fcol(np.atleast_2d(x))
        if x.shape[0] != 3:
            raise ValueError("x must be 3 x N")
        N = x.shape[1]
        lo = np.array([xl, yl, zl], dtype=np.float64)
        hi = np.array([xu, yu, zu], dtype=np.float64)
        J = np.zeros((3, 3, self.M - 3, N), order="F")
        self._ck(self._lib.rbslam_op_jacobian_phi3d(self._h, N, _capi.dptr(x), _capi.dptr(lo), _capi.dptr(hi),
                                                    _capi.dptr(J)))
        return J

    def op_kalman_update(self, xl, P, y_t, R, jitter=1e-3, H=None, xn=None):
        """xl [M x N], P [M x M x N] (MATLAB layout), H [N x d x M]; returns updated copies."""
        xl = np.array(xl, dtype=np.float64, order="F")
        P = np.array(P, dtype=np.float64, order="F")
        N = xl.shape[1]
        Hc = None if H is None else _capi.fcol(H)
        xnc = None if xn is None else _capi.fcol(xn)
        y_t = np.ascontiguousarray(y_t, dtype=np.float64)
        R = _capi.fcol(np.atleast_2d(R))
        logw = np.zeros(N)
        self._ck(self._lib.rbslam_op_kalman_update(self._h, N, _capi.dptr(xnc), _capi.dptr(Hc),
                                                   _capi.dptr(y_t), _capi.dptr(R), float(jitter),
                                                   _capi.dptr(xl), _capi.dptr(P), _capi.dptr(logw)))
        return xl, P, logw

    def op_ancestor_weights(self, form, A, v, S, r, R=None, q2=None, hld=None, jitter=1e-2):
        """logwMeas [N] of the reference trajectory's ancestor weights (rbslam_op_ancestor_weights):
        form 0: A = P [M x M x N], v = xl [M x N], S = D [ne x M], r = y_future [ne], R;
        form 1: A = Imat, v = ivec, S = ImatAddt [M x M], r = ivecAddt [M], q2, hld [N]."""
        A = _capi.fcol(A)
        v = _capi.fcol(v)
        S = _capi.fcol(np.atleast_2d(S))
        r = np.ascontiguousarray(r, dtype=np.float64)
        N = v.shape[1]
        ne = S.shape[0] if form == 0 else 0
        Rc = None if R is None else _capi.fcol(np.atleast_2d(R))
        q2c = None if q2 is None else np.ascontiguousarray(q2, dtype=np.float64)
        hc = None if hld is None else np.ascontiguousarray(hld, dtype=np.float64)
        out = np.zeros(N)
        self._ck(self._lib.rbslam_op_ancestor_weights(self._h, int(form), N, ne, _capi.dptr(A), _capi.dptr(v),
                                                      _capi.dptr(S), _capi.dptr(r), _capi.dptr(Rc),
                                                      _capi.dptr(q2c), _capi.dptr(hc), float(jitter),
                                                      _capi.dptr(out)))
        return out

    def op_dyn_logweight(self, xnk_t, xn, dx, dt, Q, use_default=False):
        xn = _capi.fcol(xn)
        N = xn.shape[1]
        xk = np.ascontiguousarray(xnk_t, dtype=np.float64)
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        Q = _capi.fcol(np.atleast_2d(Q))
        out = np.zeros(N)
        self._ck(self._lib.rbslam_op_dyn_logweight(self._h, N, _capi.dptr(xk), _capi.dptr(xn),
                                                   _capi.dptr(dx), float(dt), _capi.dptr(Q),
                                                   int(use_default), _capi.dptr(out)))
        return out


def ekf_dense(model, odometry, y, x0, q0, P0, Q, R, dt, LL=None, *, device=0, keep_P=False):
    """EKF baseline of the dense magnetic-field example; mirrors
    ``[xf_traj,qnb_traj,Pf_traj] = ekf_dense(dynModel,measModel,odometry,y,x0,q0,P0,Q,R,dt)``
    (examples/slam-dense-mag/ekf_dense.m:1-2) with ``model`` standing for the two closures
    (run_dense3D_magfield.m:281-316).  LL [2 x 3]: the domain bounds measModel_ekf passes to
    JacobianPhi3D (default [-L; L]).  Returns (xf_traj, qnb_traj, Pf): the last filtered covariance,
    or all of them [ns x ns x T] with keep_P."""
    if model.family != _capi.MODEL_DENSE_MAG3D:
        raise _capi.UnsupportedModelError(_capi.EMODEL, "ekf_dense is defined for the dense magnetic-field model")
    y = np.asarray(y, dtype=np.float64)
    T = y.shape[0]
    with Context(model, 1, max(T, 1), device=device) as ctx:
        ns = ctx.M + 6
        F = _capi.fcol
        odo, yy = F(np.atleast_2d(odometry)), F(y)
        x0c = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(-1))
        q0c = np.ascontiguousarray(np.asarray(q0, dtype=np.float64).reshape(-1))
        P0c, Rc = F(P0), F(np.atleast_2d(R))
        Qc = np.asarray(Q, dtype=np.float64)
        Qc = F(Qc[:, :, None] if Qc.ndim == 2 else Qc)
        dtc = np.ascontiguousarray(np.asarray(dt, dtype=np.float64).reshape(-1))
        if x0c.shape[0] != ns or q0c.shape[0] != 4 or P0c.shape != (ns, ns) or yy.shape[1] != 3:
            raise ValueError("ekf_dense: x0 [%d], q0 [4], P0 [%d x %d], y [T x 3]" % (ns, ns, ns))
        if Qc.shape[:2] != (6, 6) or Rc.shape != (3, 3) or (T > 1 and (odo.shape[1] != 7 or odo.shape[0] < T - 1)):
            raise ValueError("ekf_dense: Q [6 x 6 (x pages)], R [3 x 3], odometry [>= T-1 x 7]")
        LLc = F(np.vstack([-model.L, model.L]) if LL is None else np.asarray(LL, dtype=np.float64).reshape(2, 3))
        xf = np.zeros((ns, T), order="F")
        qn = np.zeros((4, T), order="F")
        Pl = np.zeros((ns, ns), order="F")
        Pt = np.zeros((ns, ns, T), order="F") if keep_P else None
        ctx._ck(ctx._lib.rbslam_ekf_run(ctx._h, T, _capi.dptr(odo), odo.shape[0], _capi.dptr(yy), _capi.dptr(x0c),
                                        _capi.dptr(q0c), _capi.dptr(P0c), _capi.dptr(Qc), Qc.shape[2],
                                        _capi.dptr(Rc), _capi.dptr(dtc), dtc.shape[0], _capi.dptr(LLc),
                                        _capi.dptr(xf), _capi.dptr(qn), _capi.dptr(Pl), _capi.dptr(Pt)))
    return xf, qn, (Pt if keep_P else Pl)


def particleFilterLocalization(dynModel, measModel, odometry, y, x0_nonLin, Q, R, N_P, dt, makePlots=None, *,
                               map_mean, var_rows, sigma2, rng=None, device=0, want_xn_traj=False, taps=False):
    """Localisation-only particle filter against a fixed map; positional signature of
    examples/mag-localization-mapping/particleFilterLocalization.m:1-2 (``R`` is unused there too).
    dynModel / measModel are the handles of a dense-mag model (rbslam.models.DenseMag3D); the constants
    the reference's measModel closure captures are keyword arguments: ``map_mean`` [M] (`foo`),
    ``var_rows`` [N_P x 3] (dVarft(i,:) per particle) and ``sigma2`` (run_localization.m:259-270).
    rng: int seed (device Philox) or an object with U [T, N], Z [T, N, 6].
    Returns (traj_max, traj_mean) [7 x T] (+ a dict of xn_traj / ancestors / w_hist / n_diverged with
    want_xn_traj / taps)."""
    model = _models.resolve(dynModel, measModel)
    if model.family != _capi.MODEL_DENSE_MAG3D:
        raise _capi.UnsupportedModelError(_capi.EMODEL, "particleFilterLocalization needs the dense-mag model")
    if makePlots is not None:
        raise _capi.RbslamError(_capi.EARG, "particleFilterLocalization: the per-step plotting hook is not forwarded")
    y = np.asarray(y, dtype=np.float64)
    T, N = y.shape[0], int(N_P)
    seed = int(rng) if isinstance(rng, (int, np.integer)) else 0
    streams = None if (rng is None or isinstance(rng, (int, np.integer))) else rng
    F = _capi.fcol
    with Context(model, 1, 1, device=device, seed=seed) as ctx:
        odo, yy = F(np.atleast_2d(odometry)), F(y)
        x0 = np.asarray(x0_nonLin, dtype=np.float64)
        x0 = F(x0.reshape(7, -1))
        Qc = np.asarray(Q, dtype=np.float64)
        Qc = F(Qc[:, :, None] if Qc.ndim == 2 else Qc)
        dtc = np.ascontiguousarray(np.asarray(dt, dtype=np.float64).reshape(-1))
        mm = np.ascontiguousarray(np.asarray(map_mean, dtype=np.float64).reshape(-1))
        vr = F(np.asarray(var_rows, dtype=np.float64))
        if yy.shape[1] != 3 or x0.shape[1] not in (1, N) or mm.shape[0] != ctx.M or vr.shape != (N, 3) or Qc.shape[:2] != (6, 6):
            raise ValueError("particleFilterLocalization: y [T x 3], x0 [7 x (1|N)], map_mean [%d], var_rows [N x 3], Q [6 x 6]" % ctx.M)
        if T > 1 and (odo.shape[1] != 7 or odo.shape[0] < T - 1):
            raise ValueError("odometry must be [>= T-1 x 7]")
        U = Z = None
        if streams is not None:
            U = np.ascontiguousarray(np.asarray(streams.U, dtype=np.float64).reshape(T, N))
            Z = np.ascontiguousarray(np.asarray(streams.Z, dtype=np.float64).reshape(T, N, 6))
        tm, tmean = np.zeros((7, T), order="F"), np.zeros((7, T), order="F")
        xt = np.zeros((7, N, T), order="F") if want_xn_traj else None
        anc = np.zeros((N, T), dtype=np.int32, order="F") if taps else None
        wh = np.zeros((N, T), order="F") if taps else None
        nd = C.c_int32()
        ctx._ck(ctx._lib.rbslam_localization_run(
            ctx._h, N, T, _capi.dptr(odo), odo.shape[0], _capi.dptr(yy), _capi.dptr(x0), x0.shape[1], _capi.dptr(Qc),
            Qc.shape[2], _capi.dptr(dtc), dtc.shape[0], _capi.dptr(mm), _capi.dptr(vr), float(sigma2), _capi.dptr(U),
            _capi.dptr(Z), _capi.dptr(tm), _capi.dptr(tmean), _capi.dptr(xt), _capi.iptr(anc), _capi.dptr(wh),
            C.byref(nd)))
    if want_xn_traj or taps:
        return tm, tmean, dict(xn_traj=xt, ancestors=anc, w_hist=wh, n_diverged=int(nd.value))
    return tm, tmean


def plan_migration(ai, old_owner, world):
    """Host-only planner (rbslam_plan_migration): returns (new_owner [N], n_migrate)."""
    ai = np.ascontiguousarray(ai, dtype=np.int32)
    oo = np.ascontiguousarray(old_owner, dtype=np.int32)
    no = np.zeros_like(ai)
    nm = C.c_int32()
    rc = _capi.lib().rbslam_plan_migration(ai.shape[0], int(world), _capi.iptr(ai), _capi.iptr(oo),
                                           _capi.iptr(no), C.byref(nm))
    if rc != _capi.OK:
        raise _capi.RbslamError(rc, "plan_migration: bad arguments")
    return no, int(nm.value)


# ---------------------------------------------------------------------------
# the three drop-in entry points
# ---------------------------------------------------------------------------
def particleFilter(dynModel, measModel, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R, N_P, dt,
                   sparseFeatures=None, makePlots=None, *, rng=None, device=0, **ctx_kw):
    """Rao-Blackwellized particle filter; signature of src/particleFilter.m:1-3.

    Returns (traj_max, traj_mean, xl_max, xl_mean, P_max, P_mean, traj_sample_iwmax, xn_traj).
    """
    model = _models.resolve(dynModel, measModel)
    if sparseFeatures is None or (isinstance(sparseFeatures, (list, tuple)) and not sparseFeatures):
        sparseFeatures = False        # src/particleFilter.m:85
    if bool(sparseFeatures) != bool(model.sparse):
        raise _capi.UnsupportedModelError(_capi.EMODEL, "sparseFeatures does not match the model family")
    mode, seed, streams = _as_streams(rng)
    yy = np.asarray(y, dtype=np.float64)
    T = yy.shape[0]
    ctx_kw.setdefault("kalman_variant", -1)   # filter only: packed symmetric slabs where they apply
    with Context(model, N_P, T, device=device, rng_mode=mode, seed=seed, **ctx_kw) as ctx:
        if makePlots is not None:
            # makePlots(xn,xl(:,iw_max),P(:,:,iw_max),traj_max,yhattraj,xn_traj,traj_mean,xl,P)
            # (src/particleFilter.m:215-217), every step
            def _cb(k, t, ctx=ctx):
                st = ctx.read_particles()
                tr = ctx.read_trajectories()
                im = int(np.argmax(st["w"]))
                makePlots(st["xn"], st["xl"][:, im], st["P"][:, :, im], tr["traj_max"], tr["yhattraj"],
                          tr["xn_traj"], tr["traj_mean"], st["xl"], st["P"])
            ctx.set_step_callback(_cb)
        o = ctx.filter_run(odometry, yy, x0_nonLin, x0_lin, P0_lin, Q, R, dt, streams)
    return (o["traj_max"], o["traj_mean"], o["xl_max"], o["xl_mean"], o["P_max"], o["P_mean"],
            o["traj_sample_iwmax"], o["xn_traj"])


def _smoother(form, dynModel, measModel, dynResNorm, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R,
              N_P, N_K, dt, sparseFeatures, makePlots, rng, device, ctx_kw):
    model = _models.resolve(dynModel, measModel, dynResNorm)
    if sparseFeatures is None:
        sparseFeatures = False
    if bool(sparseFeatures) != bool(model.sparse):
        raise _capi.UnsupportedModelError(_capi.EMODEL, "sparseFeatures does not match the model family")
    if form == 1 and sparseFeatures:
        # src/particleSmootherInformationForm.m:77-80
        print("This code has only been implemented for dense features")
        return None
    mode, seed, streams = _as_streams(rng)
    yy = np.asarray(y, dtype=np.float64)
    T = yy.shape[0]
    with Context(model, N_P, T, device=device, rng_mode=mode, seed=seed,
                 information_form=(form == 1), **ctx_kw) as ctx:
        def _cb(k, t, ctx=ctx):
            if t != T:
                return
            o = ctx._smoother_out       # filled sweep by sweep by the library
            if makePlots is not None:   # makePlots(xnk,xlk,k,XNK,XLK,PK), src/particleSmoother.m:359-361
                makePlots(o["XNK"][:, :, k], o["XLK"][:, k], k, o["XNK"], o["XLK"], o["PK"])   # k 0-based here, 1-based in MATLAB
            print("Particle smoother iteration %i/%i done." % (k + 1, N_K))   # :365
        ctx.set_step_callback(_cb)
        o = ctx.smoother_run(odometry, yy, x0_nonLin, x0_lin, P0_lin, Q, R, dt, N_K, form, streams)
    return o["XNK"], o["XLK"], o["PK"]


def particleSmoother(dynModel, measModel, dynResNorm, odometry, y, x0_nonLin, x0_lin, P0_lin, Q, R,
                     N_P, N_K, dt, sparseFeatures=None, makePlots=None, *, rng=None, device=0,
                     **ctx_kw):
    """Rao-Blackwellized particle smoother; signature of src/particleSmoother.m:1-2."""
    return _smoother(0, dynModel, measModel, dynResNorm, odometry, y, x0_nonLin, x0_lin, P0_lin, Q,
                     R, N_P, N_K, dt, sparseFeatures, makePlots, rng, device, ctx_kw)


def particleSmootherInformationForm(dynModel, measModel, dynResNorm, odometry, y, x0_nonLin, x0_lin,
                                    P0_lin, Q, R, N_P, N_K, dt, sparseFeatures=None, makePlots=None,
                                    *, rng=None, device=0, **ctx_kw):
    """Information-form RBPS; signature of src/particleSmootherInformationForm.m:1-2."""
    return _smoother(1, dynModel, measModel, dynResNorm, odometry, y, x0_nonLin, x0_lin, P0_lin, Q,
                     R, N_P, N_K, dt, sparseFeatures, makePlots, rng, device, ctx_kw)
