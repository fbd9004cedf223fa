"""ctypes binding of the CPU oracle (oracle/densecrf_oracle.c).

TEST INFRASTRUCTURE -- PARITY UNPINNED (see the header of densecrf_oracle.c and DESIGN.md).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (wsss_analysis_b200) never does.

The class mirrors the pydensecrf surface the reference exercises
(/root/reference/03c_hsn/utilities.py:427-443) so that parity tests read like the call sites.
"""
import ctypes as C
import os
import subprocess
from numbers import Number

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_dcrf.so")

CONST_KERNEL, DIAG_KERNEL, FULL_KERNEL = 0, 1, 2
NO_NORMALIZATION, NORMALIZE_BEFORE, NORMALIZE_AFTER, NORMALIZE_SYMMETRIC = 0, 1, 2, 3
_POTTS, _DIAGONAL, _MATRIX = 0, 1, 2


def build(force=False):
    """Compile the oracle with gcc (building the checker is not using it)."""
    srcs = [os.path.join(_HERE, f) for f in ("densecrf_oracle.c", "expf_ref.c")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_dcrf.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, fp, ip = C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.orc_lattice_create.restype = vp
        L.orc_lattice_create.argtypes = [vp, C.c_int, C.c_int]
        L.orc_lattice_free.argtypes = [vp]
        for f in (L.orc_lattice_M, L.orc_lattice_d, L.orc_lattice_N):
            f.restype = C.c_int
            f.argtypes = [vp]
        L.orc_lattice_export.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_lattice_compute.argtypes = [vp, vp, vp, C.c_int, C.c_int]
        L.orc_crf_create.restype = vp
        L.orc_crf_create.argtypes = [C.c_int, C.c_int]
        L.orc_crf_free.argtypes = [vp]
        L.orc_crf_set_unary.argtypes = [vp, vp]
        L.orc_crf_add_pairwise.restype = C.c_int
        L.orc_crf_add_pairwise.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]
        L.orc_crf_add_gaussian_2d.restype = C.c_int
        L.orc_crf_add_gaussian_2d.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, vp,
                                              C.c_int, C.c_int]
        L.orc_crf_add_bilateral_2d.restype = C.c_int
        L.orc_crf_add_bilateral_2d.argtypes = [vp, C.c_int, C.c_int] + [C.c_float] * 5 + [
            vp, C.c_int, vp, C.c_int, C.c_int]
        L.orc_crf_num_pairwise.restype = C.c_int
        L.orc_crf_num_pairwise.argtypes = [vp]
        L.orc_crf_lattice.restype = vp
        L.orc_crf_lattice.argtypes = [vp, C.c_int]
        L.orc_crf_norm.argtypes = [vp, C.c_int, vp]
        L.orc_crf_inference.argtypes = [vp, C.c_int, vp]
        L.orc_crf_start_inference.argtypes = [vp, vp]
        L.orc_crf_step_inference.argtypes = [vp, vp]
        L.orc_crf_kl_divergence.restype = C.c_double
        L.orc_crf_kl_divergence.argtypes = [vp, vp]
        L.orc_bruteforce_gaussian.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_int]
        L.orc_expf_ref.restype = C.c_float
        L.orc_expf_ref.argtypes = [C.c_float]
        L.orc_expf_host.argtypes = [vp, vp, C.c_int64]
        L.orc_expf_ref_check.restype = C.c_int64
        L.orc_expf_ref_check.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_confusion_accumulate.restype = C.c_int64
        L.orc_confusion_accumulate.argtypes = [vp, vp, C.c_int64, C.c_int, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _compat(compat, L):
    if isinstance(compat, Number):
        return _POTTS, np.array([compat], dtype=np.float32)
    a = np.ascontiguousarray(compat, dtype=np.float32)
    if a.ndim == 1:
        assert a.shape[0] == L
        return _DIAGONAL, a
    assert a.shape == (L, L)
    return _MATRIX, a


class LatticeExport:
    """Integer/float internals of one lattice in the reference numbering (Appendix A.3)."""

    def __init__(self, handle):
        L = lib()
        self.N, self.d, self.M = L.orc_lattice_N(handle), L.orc_lattice_d(handle), L.orc_lattice_M(handle)
        d1 = self.d + 1
        self.keys = np.zeros((self.M, self.d), np.int16)
        self.offsets = np.zeros((self.N, d1), np.int32)
        self.bary = np.zeros((self.N, d1), np.float32)
        self.neighbours = np.zeros((d1, self.M, 2), np.int32)
        self.rank = np.zeros((self.N, d1), np.int16)
        L.orc_lattice_export(handle, _ptr(self.keys), _ptr(self.offsets), _ptr(self.bary),
                             _ptr(self.neighbours), _ptr(self.rank))


class Lattice:
    """Stand-alone permutohedral lattice over (d, N) features."""

    def __init__(self, features_dN):
        f = np.ascontiguousarray(np.asarray(features_dN, np.float32).T)  # N x d pixel-major
        self.N, self.d = f.shape
        self._h = lib().orc_lattice_create(_ptr(f), self.N, self.d)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_lattice_free(self._h)
            self._h = None

    def export(self):
        return LatticeExport(self._h)

    def compute(self, values_LN, reverse=False):
        v = np.ascontiguousarray(np.asarray(values_LN, np.float32).T)  # N x L
        out = np.empty_like(v)
        lib().orc_lattice_compute(self._h, _ptr(out), _ptr(v), v.shape[1], int(reverse))
        return np.ascontiguousarray(out.T)


class DenseCRF:
    def __init__(self, nvar, nlabels):
        self.N, self.L = int(nvar), int(nlabels)
        self._h = lib().orc_crf_create(self.N, self.L)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_crf_free(self._h)
            self._h = None

    def setUnaryEnergy(self, u):
        u = np.asarray(u)
        if u.dtype != np.float32 or not u.flags.c_contiguous:
            raise ValueError("unary must be C-contiguous float32")
        if u.shape != (self.L, self.N):
            raise ValueError("Bad shape for unary energy (Need {}, got {})".format((self.L, self.N), u.shape))
        lib().orc_crf_set_unary(self._h, _ptr(u))

    def addPairwiseEnergy(self, features, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        f = np.ascontiguousarray(features, dtype=np.float32)
        if f.ndim != 2 or f.shape[1] != self.N:
            raise ValueError("Bad shape for pairwise energy (Need (?, {}), got {})".format(self.N, f.shape))
        kind, c = _compat(compat, self.L)
        lib().orc_crf_add_pairwise(self._h, _ptr(f), f.shape[0], kind, _ptr(c), kernel, normalization)

    def inference(self, niter):
        Q = np.empty((self.L, self.N), np.float32)
        lib().orc_crf_inference(self._h, int(niter), _ptr(Q))
        return Q

    def startInference(self):
        Q = np.empty((self.L, self.N), np.float32)
        lib().orc_crf_start_inference(self._h, _ptr(Q))
        return Q, np.empty_like(Q), np.empty_like(Q)

    def stepInference(self, Q, tmp1=None, tmp2=None):
        assert Q.dtype == np.float32 and Q.flags.c_contiguous
        lib().orc_crf_step_inference(self._h, _ptr(Q))

    def klDivergence(self, Q):
        Q = np.ascontiguousarray(Q, np.float32)
        return float(lib().orc_crf_kl_divergence(self._h, _ptr(Q)))

    # --- introspection for bit-exact lattice tests ---
    def lattice(self, k):
        return LatticeExport(lib().orc_crf_lattice(self._h, k))

    def norm(self, k):
        out = np.empty(self.N, np.float32)
        lib().orc_crf_norm(self._h, k, _ptr(out))
        return out


class DenseCRF2D(DenseCRF):
    def __init__(self, w, h, nlabels):
        super().__init__(int(w) * int(h), nlabels)
        self.W, self.H = int(w), int(h)

    def addPairwiseGaussian(self, sxy, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        if isinstance(sxy, Number):
            sxy = (sxy, sxy)
        kind, c = _compat(compat, self.L)
        lib().orc_crf_add_gaussian_2d(self._h, self.W, self.H, sxy[0], sxy[1], kind, _ptr(c), kernel,
                                      normalization)

    def addPairwiseBilateral(self, sxy, srgb, rgbim, compat, kernel=DIAG_KERNEL,
                             normalization=NORMALIZE_SYMMETRIC):
        if isinstance(sxy, Number):
            sxy = (sxy, sxy)
        if isinstance(srgb, Number):
            srgb = (srgb, srgb, srgb)
        im = np.asarray(rgbim)
        if im.dtype != np.uint8 or not im.flags.c_contiguous:
            raise ValueError("rgbim must be C-contiguous uint8")
        if im.shape != (self.H, self.W, 3):
            raise ValueError("Bad shape for pairwise bilateral (Need {}, got {})".format(
                (self.H, self.W, 3), im.shape))
        kind, c = _compat(compat, self.L)
        lib().orc_crf_add_bilateral_2d(self._h, self.W, self.H, sxy[0], sxy[1], srgb[0], srgb[1], srgb[2],
                                       _ptr(im), kind, _ptr(c), kernel, normalization)


# ---- Appendix A.8 utilities, restated independently of the product's copy ----
def unary_from_softmax(sm, scale=None, clip=1e-5):
    num_cls = sm.shape[0]
    if scale is not None:
        assert 0 < scale <= 1
        sm = scale * sm + (1 - scale) * (np.ones(sm.shape) / num_cls)
    if clip is not None:
        sm = np.clip(sm, clip, 1.0)
    return -np.log(sm).reshape([num_cls, -1]).astype(np.float32)


def unary_from_labels(labels, n_labels, gt_prob, zero_unsure=True):
    assert 0 < gt_prob < 1
    labels = labels.flatten()
    n_energy = -np.log((1.0 - gt_prob) / (n_labels - 1))
    p_energy = -np.log(gt_prob)
    U = np.full((n_labels, len(labels)), n_energy, dtype="float32")
    U[labels - 1 if zero_unsure else labels, np.arange(U.shape[1])] = p_energy
    if zero_unsure:
        U[:, labels == 0] = -np.log(1.0 / n_labels)
    return U


def bruteforce_gaussian(features_dN, values_LN):
    f = np.ascontiguousarray(np.asarray(features_dN, np.float32).T)
    v = np.ascontiguousarray(np.asarray(values_LN, np.float32).T)
    out = np.empty_like(v)
    lib().orc_bruteforce_gaussian(_ptr(f), f.shape[0], f.shape[1], _ptr(v), _ptr(out), v.shape[1])
    return np.ascontiguousarray(out.T)


def confusion(gt, pred, C_):
    """(C+1, C) int64; last row collects ignored GT (chainercv drops it)."""
    gt = np.ascontiguousarray(gt, np.int32).ravel()
    pred = np.ascontiguousarray(pred, np.int32).ravel()
    conf = np.zeros((C_ + 1, C_), np.int64)
    lib().orc_confusion_accumulate(_ptr(gt), _ptr(pred), gt.size, C_, _ptr(conf))
    return conf


def expf_host(x):
    """libm expf element-wise (float32) -- the routine the oracle's softmax calls."""
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib().orc_expf_host(_ptr(x), _ptr(y), x.size)
    return y


def expf_ref_mismatches(u_lo=0x80000000, u_hi=0xC2D00000, stride=1):
    """Number of float bit patterns in [u_lo, u_hi] (step `stride`) where the restated glibc expf
    (expf_ref.c) differs from the host libm, and the first such pattern."""
    first = C.c_uint32(0)
    n = lib().orc_expf_ref_check(u_lo, u_hi, stride, C.byref(first))
    return int(n), int(first.value)
