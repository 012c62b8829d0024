#!/usr/bin/env python
"""Make GPU-running copies of the reference's four L2 runners.

    python patch_runners.py /path/to/Rao-Blackwellized-SLAM-smoothing [out_dir]

The reference's sources are NOT redistributed here: this script reads the user's own checkout,
rewrites the 2-3 lines per runner that hand MATLAB closures to the inference engines so that they
hand rbslam model handles instead (rbslam_model.m), and writes the patched copies to out_dir
(default: ./patched).  The same edits are committed next to this script as zero-context diffs
(*.patch, `diff -U0`), to be applied with `patch -p1` from the reference's root.

With matlab/ (the drop-in particleFilter.m / particleSmoother.m / ekf_dense.m / JacobianPhi3D.m and
the MEX gateway) ahead of the reference's src/ and tools/ on the MATLAB path, the patched runners
run the three example configurations on the GPU; everything around the calls (data simulation,
RMSE, plotting callbacks) stays the reference's own MATLAB.
"""
import os
import re
import sys

EDITS = {
    # examples/slam-dense-mag/run_dense3D_magfield.m:91,143,207,210,251
    "examples/slam-dense-mag/run_dense3D_magfield.m": [
        (r"(\[eigenval,~,eigenfun_dx,NN\] = domain_cartesian_dx\(nBasisFunctions,d,LL\);)",
         r"\1\nmdl = rbslam_model('denseMag3D', NN, LL); % rbslam: GPU model handles for the closures below"),
        (r"particleFilter\(@dynModel,@measModel,", r"particleFilter(mdl.dynModel,mdl.measModel,"),
        (r"particleSmoother\(@dynModel,@measModel,dynResNorm,", r"particleSmoother(mdl.dynModel,mdl.measModel,mdl.dynResNorm,"),
        (r"ekf_dense\(@dynModel_ekf,@measModel_ekf,", r"ekf_dense(mdl.dynModel_ekf,mdl.measModel_ekf,"),
    ],
    # examples/slam-dense-radio/run_dense2D_withHeading.m:114,176,200
    "examples/slam-dense-radio/run_dense2D_withHeading.m": [
        (r"(\[eigenval,eigenfun,~,NN\] = domain_cartesian_dx\(nBasisFunctions,d,LL\);)",
         r"\1\nmdl = rbslam_model('denseRadio2D', NN, LL); % rbslam: GPU model handles"),
        (r"particleFilter\(dynModel,measModel,", r"particleFilter(mdl.dynModel,mdl.measModel,"),
        (r"particleSmoother\(dynModel,measModel,dynResNorm,", r"particleSmoother(mdl.dynModel,mdl.measModel,mdl.dynResNorm,"),
    ],
    # examples/slam-sparse-visual/pfslam.m:82,108
    "examples/slam-sparse-visual/pfslam.m": [
        (r"(  measModel = @\(xn,xl\) measurement\(\[xn\(1:3\); xl\],f,fp,fw,true\);)",
         r"\1\n  mdl = rbslam_model('sparseVisual2D', size(map,2), f, fp, fw); % rbslam: GPU model handles"),
        (r"particleFilter\(dynModel,measModel,", r"particleFilter(mdl.dynModel,mdl.measModel,"),
    ],
    # examples/slam-sparse-visual/psslam.m:92,118
    "examples/slam-sparse-visual/psslam.m": [
        (r"(  measModel = @\(xn,xl\) measurement\(\[xn\(1:3\); xl\],f,fp,fw,true\);)",
         r"\1\n  mdl = rbslam_model('sparseVisual2D', size(map,2), f, fp, fw); % rbslam: GPU model handles"),
        (r"particleSmoother\(dynModel,measModel,\[\],", r"particleSmoother(mdl.dynModel,mdl.measModel,mdl.dynResNorm,"),
    ],
}


def patch_text(rel, text):
    for pat, rep in EDITS[rel]:
        text, n = re.subn(pat, rep, text)
        if n == 0:
            raise SystemExit("%s: pattern not found: %s" % (rel, pat))
    return text


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    ref = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.getcwd(), "patched")
    for rel in EDITS:
        src = open(os.path.join(ref, rel)).read()
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, "w").write(patch_text(rel, src))
        print("wrote", dst)


if __name__ == "__main__":
    main()
