"""pydensecrf.utils-compatible host helpers ([EXT]).

Called by the reference at /root/reference/03c_hsn/utilities.py:11,431 (`unary_from_softmax`) and,
through the missing wrapper `misc/imutils.py`, by 03b_irn/step/cam_to_ir_label.py:35
(`unary_from_labels`).  These are shape/dtype glue in NumPy, exactly like the package they replace;
the arithmetic-heavy path (lattice, filtering, softmax) is CUDA only.
"""
from numbers import Number

import numpy as np


def unary_from_labels(labels, n_labels, gt_prob, zero_unsure=True):
    """Energy -log(gt_prob) at the labelled class, -log((1-gt_prob)/(n_labels-1)) elsewhere.

    With zero_unsure, label 0 means "unsure" (uniform energy) and classes are 1-based."""
    assert 0 < gt_prob < 1, "`gt_prob must be in (0,1)."
    labels = np.asarray(labels).flatten()
    n_energy = -np.log((1.0 - gt_prob) / (n_labels - 1))
    p_energy = -np.log(gt_prob)
    U = np.full((n_labels, len(labels)), n_energy, dtype="float32")
    U[labels - 1 if zero_unsure else labels, np.arange(U.shape[1])] = p_energy
    if zero_unsure:
        U[:, labels == 0] = -np.log(1.0 / n_labels)
    return U


def compute_unary(labels, M, GT_PROB=0.5):
    """Deprecated upstream alias kept for drop-in completeness."""
    return unary_from_labels(labels, M, GT_PROB)


def unary_from_softmax(sm, scale=None, clip=1e-5):
    """-log of class probabilities (first axis = class), flattened to (n_classes, -1) float32."""
    sm = np.asarray(sm)
    num_cls = sm.shape[0]
    if scale is not None:
        assert 0 < scale <= 1, "`scale` needs to be in (0,1]"
        uniform = np.ones(sm.shape) / num_cls
        sm = scale * sm + (1 - scale) * uniform
    if clip is not None:
        sm = np.clip(sm, clip, 1.0)
    return -np.log(sm).reshape([num_cls, -1]).astype(np.float32)


def softmax_to_unary(sm, GT_PROB=1):
    """Deprecated upstream alias."""
    return unary_from_softmax(sm, scale=GT_PROB, clip=None)


def create_pairwise_gaussian(sdims, shape):
    """Position-only features for an n-D grid: (len(shape), prod(shape)) float32, for addPairwiseEnergy."""
    hcord_range = [range(s) for s in shape]
    mesh = np.array(np.meshgrid(*hcord_range, indexing="ij"), dtype=np.float32)
    for i, s in enumerate(sdims):
        mesh[i] /= s
    return mesh.reshape([len(sdims), -1])


def create_pairwise_bilateral(sdims, schan, img, chdim=-1):
    """Position + channel features for an n-D image, for addPairwiseEnergy."""
    if chdim == -1:
        im_feat = img[np.newaxis].astype(np.float32)
    else:
        im_feat = np.rollaxis(img, chdim).astype(np.float32)
    if isinstance(schan, Number):
        im_feat /= schan
    else:
        for i, s in enumerate(schan):
            im_feat[i] /= s
    cord_range = [range(s) for s in im_feat.shape[1:]]
    mesh = np.array(np.meshgrid(*cord_range, indexing="ij"), dtype=np.float32)
    for i, s in enumerate(sdims):
        mesh[i] /= s
    feats = np.concatenate([mesh, im_feat])
    return feats.reshape([feats.shape[0], -1])
