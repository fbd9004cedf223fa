"""Generates tests/golden/*.npz from the CPU oracle (run here; the fixtures travel to the GPU box).

PARITY UNPINNED: neither pydensecrf nor any golden vector of it exists in /root/reference or in this
image, so these fixtures pin the ORACLE against regressions (and give the GPU tests a fixed target
that does not depend on the oracle library being rebuilt); they are not outputs of the reference.
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from wsss_analysis_b200 import synthetic as S  # noqa: E402

CASES = {
    # name: (W, H, L, n_iter, gauss sxy, gauss compat, bilateral sxy, srgb, compat, image kind, seed)
    "voc_small": (50, 38, 21, 10, 3, 3, 80, 13, 10, "natural", 0),          # SEC.py:20 at 1/10 scale
    "sec_train_41": (41, 41, 21, 5, 3 / 12, 3, 80 / 12, 13, 10, "natural", 1),  # SEC.py:19
    "adp_morph": (48, 48, 29, 5, 1, 20, 10, 40, 50, "histo", 2),             # SEC.py:24-25
    "adp_func": (48, 48, 5, 5, 3, 40, 10, 4, 25, "histo", 3),                # SEC.py:29-30
    "hsn_voc": (56, 56, 6, 10, 3 / 2, 3, 80 / 2, 13, 10, "natural", 4),      # 03c_hsn/demo.py:159
    "irn_label": (60, 44, 4, 10, 3, 3, 50, 5, 10, "iid", 5),                 # IRN_CRF_CONFIG
    # HSN VOC-M7 (03c_hsn/demo.py:161): sxy 3/12/4 and 80/12/4 -- the largest lattice coordinates in the
    # tree (x / 0.0625 at 224 px); a 224-wide strip keeps the fixture small
    "hsn_voc_m7": (224, 16, 8, 10, 3 / 12 / 4, 3, 80 / 12 / 4, 13, 10, "natural", 6),
}


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])   # optional: regenerate just the named fixtures
    for name, (W, H, L, n, gs, gc, bs, srgb, bc, kind, seed) in CASES.items():
        if only and name not in only:
            continue
        img = getattr(S, kind + "_image")(H, W, seed)
        U = S.random_unary(L, W * H, seed)
        d = O.DenseCRF2D(W, H, L)
        d.setUnaryEnergy(U)
        d.addPairwiseGaussian(sxy=gs, compat=gc)
        d.addPairwiseBilateral(sxy=bs, srgb=srgb, rgbim=img, compat=bc)
        Q = d.inference(n)
        lg, lb = d.lattice(0), d.lattice(1)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            params=np.array([W, H, L, n, gs, gc, bs, srgb, bc], np.float64), img=img, U=U,
            Q=Q.astype(np.float32), labels=Q.argmax(0).astype(np.int32),
            g_M=lg.M, g_keys=lg.keys, g_offsets=lg.offsets, g_bary=lg.bary, g_neigh=lg.neighbours, g_norm=d.norm(0),
            b_M=lb.M, b_keys=lb.keys, b_offsets=lb.offsets, b_bary=lb.bary, b_neigh=lb.neighbours, b_norm=d.norm(1))
        print(name, "M_g", lg.M, "M_b", lb.M, "Q max", float(Q.max()))


if __name__ == "__main__":
    main()
