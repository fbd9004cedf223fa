"""Golden vectors for the two resizes from the REAL cv2.resize (opencv-python-headless is in the
build image; it is the library the reference calls: 03a_sec-dsrg/model.py:686, 03b_irn/step/
eval_sem_seg.py:36).  Unlike the CRF goldens these pin parity against the reference's own dependency.
    python tools/make_golden_resize.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CASES = [  # (src h, src w, dst h, dst w)
    (41, 41, 161, 161), (81, 81, 94, 125), (81, 81, 125, 84), (34, 50, 136, 136), (64, 48, 17, 23),
    (5, 7, 5, 7), (1, 9, 4, 3), (33, 1, 7, 5), (100, 120, 50, 60),
]


def main():
    rng = np.random.default_rng(0)
    out = {"cv2_version": np.array(cv2.__version__)}
    for k, (sh, sw, dh, dw) in enumerate(CASES):
        lab = rng.integers(0, 21, (sh, sw)).astype(np.uint8)
        feat = rng.standard_normal((sh, sw, 3)).astype(np.float32)
        out["case%d_shape" % k] = np.array([sh, sw, dh, dw])
        out["case%d_lab" % k] = lab
        out["case%d_feat" % k] = feat
        out["case%d_nearest" % k] = cv2.resize(lab, (dw, dh), interpolation=cv2.INTER_NEAREST)
        out["case%d_linear" % k] = cv2.resize(feat, (dw, dh)).reshape(dh, dw, 3)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "resize_cv2.npz"), **out)
    print("wrote", len(CASES), "cases with cv2", cv2.__version__)


if __name__ == "__main__":
    main()
