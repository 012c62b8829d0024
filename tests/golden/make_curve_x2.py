"""Regenerate tests/golden/curve_x2.npz from the reference's only data fixture.

    python tests/golden/make_curve_x2.py            # needs /root/reference (this container only)

The reference ships no tests and no result files; its one fixture is the INPUT file of the
sparse visual-SLAM example, examples/slam-sparse-visual/curve-x2.mat (loaded by
load_data.m:65).  This script converts the arrays that load_data.m uses (path `p`, heading
`th`, landmark `map`, noise-free observations `Yclean`; plus the stored increments `dPos`,
`dTheta` used as known answers for the odometry recipe) to a NumPy archive so that the C3
configuration can be run on the reference's own data on the GPU box, where /root/reference
does not exist.  Nothing is computed: the values are copied bit for bit.
"""
import os
import sys

import numpy as np
import scipy.io as sio

SRC = "/root/reference/examples/slam-sparse-visual/curve-x2.mat"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "curve_x2.npz")


def main():
    if not os.path.exists(SRC):
        sys.exit("reference fixture not found: " + SRC)
    d = sio.loadmat(SRC)
    out = {k: np.ascontiguousarray(d[k], dtype=np.float64) for k in ("p", "th", "map", "Yclean", "dPos", "dTheta")}
    np.savez_compressed(DST, **out)
    print("wrote", DST, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
