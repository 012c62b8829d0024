"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every
symbol include/rbslam.h declares, host-side argument logic, the model registry and
the loud failure of the product path when no GPU / no library is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, PKG


def _build():
    import __graft_entry__ as g
    g.build()


def test_library_builds_loads_and_exports_every_header_symbol():
    _build()
    from rbslam import _capi
    hdr = open(os.path.join(ROOT, "include", "rbslam.h")).read()
    declared = set(re.findall(r"RBSLAM_API\s+[\w\s\*]+?\b(rbslam_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)
    L = _capi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.rbslam_version() == 1


def test_config_struct_abi_and_loud_failure_without_gpu():
    """struct_size is checked first (EARG on mismatch); with a matching struct the
    library goes on to CUDA and -- on a box without a GPU -- fails with ECUDA.  There is
    no CPU fallback to fall into."""
    _build()
    from rbslam import _capi
    import rbslam
    L = _capi.lib()
    cfg = _capi.Config()
    cfg.struct_size = 4
    h = C.c_void_p()
    assert L.rbslam_create(C.byref(h), C.byref(cfg)) == _capi.EARG
    assert b"struct_size" in L.rbslam_last_error(None)
    if L.rbslam_device_count() == 0:
        pr = rbslam.synth.dense_radio_problem("line_3D", m=8, seed=1, m_sim=50)
        gm = rbslam.models.from_problem(pr)
        with pytest.raises(rbslam.RbslamError) as ei:
            rbslam.Context(gm, 4, 4)
        assert ei.value.code == _capi.ECUDA
        with pytest.raises(rbslam.RbslamError):
            rbslam.particleFilter(gm.dynModel, gm.measModel, pr["odometry"], pr["y"], pr["x0_nonLin"],
                                  pr["x0_lin"], pr["P0_lin"], pr["Q"], pr["R"], 4, pr["dt"])


def test_unknown_model_family_is_rejected():
    _build()
    from rbslam import _capi
    L = _capi.lib()
    cfg = _capi.Config()
    cfg.struct_size = C.sizeof(_capi.Config)
    cfg.model = 99
    cfg.N = cfg.T = cfg.m_basis = 4
    cfg.world = 1
    h = C.c_void_p()
    assert L.rbslam_create(C.byref(h), C.byref(cfg)) == _capi.EMODEL


def test_model_registry_rejects_host_closures():
    import rbslam
    pr = rbslam.synth.dense_radio_problem("line_3D", m=8, seed=1, m_sim=50)
    gm = rbslam.models.from_problem(pr)
    assert rbslam.models.resolve(gm.dynModel, gm.measModel, gm.dynResNorm) is gm
    with pytest.raises(rbslam.UnsupportedModelError):
        rbslam.models.resolve(lambda xn, dx, dt, Q: xn, gm.measModel)
    other = rbslam.models.from_problem(pr)
    with pytest.raises(rbslam.UnsupportedModelError):
        rbslam.models.resolve(gm.dynModel, other.measModel)
    with pytest.raises(rbslam.UnsupportedModelError):
        gm.dynModel(np.zeros(3), np.zeros(3), 1.0, np.eye(1))   # handles are descriptors


def test_product_package_never_imports_the_oracle():
    """The product path must not route through oracle/ (checked statically)."""
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".m")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), (dirpath, f)


def test_plan_migration_properties():
    _build()
    import rbslam
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        for N in (8, 64, 1000):
            owner = np.repeat(np.arange(world), -(-N // world))[:N]
            owner = np.sort(owner)
            # concentrated weights: most offspring come from a few ancestors
            w = rng.random(N) ** 8
            ai = rng.choice(N, size=N, p=w / w.sum())
            new, nmig = rbslam.plan_migration(ai, owner, world)
            cap = np.array([N // world + (r < N % world) for r in range(world)])
            assert np.array_equal(np.bincount(new, minlength=world), cap)        # balanced
            stay = new == owner[ai]
            assert nmig == np.count_nonzero(~stay)
            # minimal traffic: a particle only migrates if its ancestor's rank was full
            want = np.bincount(owner[ai], minlength=world)
            assert nmig == np.sum(np.maximum(want - cap, 0))
            new2, _ = rbslam.plan_migration(ai, owner, world)
            assert np.array_equal(new, new2)                                       # deterministic
    # uniform weights keep (almost) everything local
    N, world = 1000, 8
    owner = np.sort(np.arange(N) % world)
    ai = np.sort(rng.integers(0, N, N))
    _, nmig = rbslam.plan_migration(ai, owner, world)
    assert nmig < 0.1 * N


def test_synthetic_problem_shapes():
    import rbslam
    pr = rbslam.synth.dense_mag_problem(N_T=12, m=30, seed=1, m_sim=60)
    assert pr["y"].shape == (12, 3) and pr["odometry"].shape == (12, 7)
    assert pr["P0_lin"].shape == (33, 33) and pr["NN"].dtype == np.int32
    assert np.allclose(np.linalg.norm(pr["odometry"][:-1, 3:7], axis=1), 1, atol=1e-12)
    pr = rbslam.synth.dense_radio_problem("square_3D", m=16, seed=1, m_sim=60)
    assert pr["y"].shape == (48, 1) and pr["Q"].shape == (1, 1, 48)
    pr = rbslam.synth.sparse_visual_problem(N_T=20, n_landmarks=5, N_P=6, seed=1)
    assert pr["y"].shape == (20, 5) and pr["x0_lin"].shape == (10, 6)
    assert np.isnan(pr["y"]).any()


def test_mex_gateway_compiles_against_stub_header():
    """No MATLAB in the image: the gateway is syntax/type-checked against mex/stub/mex.h."""
    import subprocess
    mexdir = os.path.join(PKG, "mex")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(mexdir, "stub"),
                        "-I", os.path.join(ROOT, "include"), os.path.join(mexdir, "rbslam_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the MATLAB shims keep the reference's signatures
    for name, args in [("particleFilter", "dynModel,measModel,odometry,y"),
                       ("particleSmoother", "dynModel,measModel,dynResNorm,odometry,y"),
                       ("particleSmootherInformationForm", "dynModel,measModel,dynResNorm,odometry,y")]:
        src = open(os.path.join(PKG, "matlab", name + ".m")).read()
        assert name + "(" + args in src.replace(" ", "").replace("...\n", "")


def test_plan_shard_invariants():
    """Slot-level plan of the sharded engine: permutation per rank, in-place keepers,
    migrants only into dead slots."""
    _build()
    from rbslam import _capi
    L = _capi.lib()
    rng = np.random.default_rng(5)
    for world in (2, 4, 8):
        N = 64 * world
        Nloc = N // world
        owner = (np.arange(N) // Nloc).astype(np.int32)
        lslot = (np.arange(N) % Nloc).astype(np.int32)
        for step in range(6):
            w = rng.random(N) ** (1 + 4 * (step % 3))
            ai = rng.choice(N, size=N, p=w / w.sum()).astype(np.int32)
            no, nl = np.zeros(N, np.int32), np.zeros(N, np.int32)
            nm = C.c_int32()
            rc = L.rbslam_plan_shard(N, world, _capi.iptr(ai), _capi.iptr(owner), _capi.iptr(lslot),
                                     _capi.iptr(no), _capi.iptr(nl), C.byref(nm))
            assert rc == 0
            n_child = np.bincount(ai, minlength=N)
            for r in range(world):
                mine = np.flatnonzero(no == r)
                assert len(mine) == Nloc and sorted(nl[mine]) == list(range(Nloc))   # permutation
            for i in range(N):
                a = ai[i]
                if owner[a] != no[i]:                       # migrant -> slot of a DEAD old particle
                    old = np.flatnonzero((owner == no[i]) & (lslot == nl[i]))[0]
                    assert n_child[old] == 0
            for a in np.flatnonzero(n_child):
                kids = np.flatnonzero((ai == a) & (no == owner[a]))
                if len(kids):                               # first local child keeps the slab
                    assert nl[kids[0]] == lslot[a]
                    assert np.count_nonzero(nl[kids] == lslot[a]) == 1
            owner, lslot = no.copy(), nl.copy()


def test_runner_patches_only_swap_the_handles():
    """matlab/examples/*.patch: zero-context diffs against the reference's four L2
    runners; every added line hands rbslam model handles to the drop-ins, nothing else changes."""
    import glob
    pdir = os.path.join(PKG, "matlab", "examples")
    files = sorted(glob.glob(os.path.join(pdir, "*.patch")))
    assert [os.path.basename(f) for f in files] == ["pfslam.patch", "psslam.patch", "run_dense2D_withHeading.patch",
                                                    "run_dense3D_magfield.patch"]
    for f in files:
        added = [ln[1:] for ln in open(f) if ln.startswith("+") and not ln.startswith("+++")]
        removed = [ln[1:] for ln in open(f) if ln.startswith("-") and not ln.startswith("---")]
        assert added and all("mdl" in ln for ln in added), f
        assert len(added) <= 5 and len(removed) <= 4, f
        assert any("rbslam_model(" in ln for ln in added), f
    ref = "/root/reference"
    if os.path.isdir(ref):        # in the build container: the patterns still match the reference
        import importlib.util
        spec = importlib.util.spec_from_file_location("patch_runners", os.path.join(pdir, "patch_runners.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for rel in mod.EDITS:
            out = mod.patch_text(rel, open(os.path.join(ref, rel)).read())
            assert "mdl." in out
