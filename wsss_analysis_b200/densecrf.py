"""pydensecrf.densecrf-compatible classes backed by libdcrf_b200.so (hand-written sm_100a CUDA).

Mirrors the class surface the reference drives at /root/reference/03c_hsn/utilities.py:427-443:

    d = DenseCRF2D(w, h, nlabels)
    d.setUnaryEnergy(U)                                   # (nlabels, w*h) float32 C-contiguous
    d.addPairwiseGaussian(sxy=..., compat=...)
    d.addPairwiseBilateral(sxy=..., srgb=..., rgbim=..., compat=...)
    Q = d.inference(n)                                    # np.array(Q).reshape(nlabels, h, w)

plus the rest of the [EXT] pydensecrf surface (DenseCRF, addPairwiseEnergy, startInference /
stepInference / klDivergence, the kernel / normalisation enums) and a batched form
(`DenseCRFBatch`) that runs many images per launch.  numpy arrays are treated as host buffers,
torch CUDA tensors as device buffers (zero-copy hand-off through data_ptr()); PyTorch is not needed
for the numpy path.  There is no CPU implementation behind these classes.
"""
import ctypes as C
from numbers import Number

import numpy as np

from . import _lib

CONST_KERNEL, DIAG_KERNEL, FULL_KERNEL = 0, 1, 2
NO_NORMALIZATION, NORMALIZE_BEFORE, NORMALIZE_AFTER, NORMALIZE_SYMMETRIC = 0, 1, 2, 3
_POTTS, _DIAGONAL, _MATRIX = 0, 1, 2


def _is_torch(x):
    return type(x).__module__.split(".")[0] == "torch"


def _buffer(x, np_dtype, what):
    """-> (pointer, on_device, device_index, keepalive).  Validation follows the Cython buffer
    checks of pydensecrf: wrong dtype / non-contiguous input is a ValueError, never a silent cast."""
    if _is_torch(x):
        import torch

        want = {np.float32: torch.float32, np.float64: torch.float64, np.uint8: torch.uint8,
                np.int32: torch.int32}[np_dtype]
        if x.dtype != want:
            raise ValueError("Buffer dtype mismatch for %s: expected %s, got %s" % (what, want, x.dtype))
        if not x.is_contiguous():
            raise ValueError("%s is not C-contiguous" % what)
        if x.is_cuda:
            return x.data_ptr(), 1, x.device.index, x
        return x.data_ptr(), 0, None, x
    a = np.asarray(x)
    if a.dtype != np_dtype:
        raise ValueError("Buffer dtype mismatch for %s: expected %s, got %s" % (what, np.dtype(np_dtype), a.dtype))
    if not a.flags.c_contiguous:
        raise ValueError("%s: ndarray is not C-contiguous" % what)
    return a.ctypes.data, 0, None, a


def _compat(compat, L):
    """[EXT] pydensecrf `_labelcomp`: number -> Potts, 1-D -> diagonal, 2-D -> matrix."""
    if isinstance(compat, Number):
        return _POTTS, np.array([compat], dtype=np.float32)
    a = np.ascontiguousarray(compat, dtype=np.float32)
    if a.ndim == 1:
        if a.shape[0] != L:
            raise ValueError("Bad shape for diagonal compatibility (Need (%d,), got %s)" % (L, a.shape))
        return _DIAGONAL, a
    if a.ndim == 2:
        if a.shape != (L, L):
            raise ValueError("Bad shape for matrix compatibility (Need (%d, %d), got %s)" % (L, L, a.shape))
        return _MATRIX, a
    raise ValueError("LabelCompatibility of dimension >2 not meaningful.")


def _pair(v, n, what):
    if isinstance(v, Number):
        return (float(v),) * n
    v = tuple(float(t) for t in v)
    if len(v) != n:
        raise ValueError("%s needs a number or a %d-sequence" % (what, n))
    return v


class _Model(object):
    """Shared implementation over one C handle (a batch of >= 1 images)."""

    def __init__(self, device, stream):
        self._h = None
        self._lib = _lib.load()
        self._device = -1 if device is None else int(device)
        self._user_stream = stream
        self._stream_ptr = None
        if isinstance(stream, str):
            assert stream == "dedicated"
            self._stream_ptr = C.c_void_p(-1).value  # DCRF_STREAM_DEDICATED
        elif stream is not None:
            self._stream_ptr = int(getattr(stream, "cuda_stream", stream))

    # -- handle management --
    def _created(self, handle):
        self._h = handle

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._lib.dcrf_destroy(h)
            except Exception:
                pass

    def close(self):
        self.__del__()

    def _pre_device_input(self, on_device):
        # device inputs were produced on torch's current stream; order them before our stream
        if on_device and self._user_stream is None:
            import torch

            torch.cuda.current_stream().synchronize()

    def set_async_host(self, on=True):
        """Calls only enqueue their copies / kernels on the handle's stream; call synchronize() before
        touching the caller's buffers (host arrays or device tensors) again."""
        _lib.check(self._lib.dcrf_set_option(self._h, 2, 1 if on else 0))
        self._async = bool(on)

    def _sync_unless_async(self):
        if not getattr(self, "_async", False):
            self.synchronize()

    def set_arithmetic(self, mode):
        """Float arithmetic of the iteration kernels: "fma" (fused multiply-add, folded normalisation,
        CUDA expf: rounding-level differences), "reference" (the association of the sequential CPU
        evaluation: bit-identical marginals, ~20 % slower), "strict" (same, also for the tails of splat
        rows longer than 256 entries) or "auto" (default: "reference" for models with a narrow
        appearance kernel, srgb < 8, where mean field amplifies rounding differences; else "fma")."""
        code = {"fma": 0, "reference": 1, "strict": 2, "auto": 3, 0: 0, 1: 1, 2: 2, 3: 3}[mode]
        _lib.check(self._lib.dcrf_set_option(self._h, 1, code))

    def set_persistent(self, mode):
        """One cooperative launch per inference(n) instead of a launch per phase: None / "auto" = by
        problem size (default), False = never, True = whenever the model allows."""
        code = {None: -1, "auto": -1, False: 0, True: 1}[mode]
        _lib.check(self._lib.dcrf_set_option(self._h, 4, code))

    def arithmetic(self):
        """The mode this model resolved to: "fma", "reference" or "strict"."""
        m = C.c_int(0)
        _lib.check(self._lib.dcrf_get_arithmetic(self._h, C.byref(m)))
        return ("fma", "reference", "strict")[m.value]

    def set_exact_arithmetic(self, on=True):
        """Round-1 name: True -> "strict", False -> "fma"."""
        self.set_arithmetic("strict" if on else "fma")

    def synchronize(self):
        _lib.check(self._lib.dcrf_synchronize(self._h))

    # -- set-up --
    def _check_unary(self, u):
        shape = tuple(u.shape)
        if len(shape) != 2 or shape[0] != self._L or shape[1] != self._Ntot:
            raise ValueError("Bad shape for unary energy (Need {}, got {})".format((self._L, self._Ntot), shape))

    def setUnaryEnergy(self, u, f=None):
        if f is not None:
            raise NotImplementedError("feature-dependent unaries (LogisticUnaryEnergy) are not part of the hot path")
        if u is None:
            raise TypeError("Argument 'u' must not be None")
        self._check_unary(u)
        ptr, dev, _, keep = _buffer(u, np.float32, "unary")
        self._pre_device_input(dev)
        _lib.check(self._lib.dcrf_set_unary(self._h, ptr, dev))
        if dev:
            self._sync_unless_async()
        del keep

    def _add_energy(self, features, compat, kernel, normalization):
        shape = tuple(features.shape)
        if len(shape) != 2 or shape[1] != self._Ntot:
            raise ValueError("Bad shape for pairwise energy (Need (?, {}), got {})".format(self._Ntot, shape))
        ptr, dev, _, keep = _buffer(features, np.float32, "features")
        kind, c = _compat(compat, self._L)
        self._pre_device_input(dev)
        _lib.check(self._lib.dcrf_add_pairwise_energy(self._h, ptr, shape[0], dev, kind, c.ctypes.data,
                                                      int(kernel), int(normalization)))
        if dev:
            self._sync_unless_async()
        del keep

    def addPairwiseEnergy(self, features, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        self._add_energy(features, compat, kernel, normalization)

    # -- inference --
    def _q_host(self, call, *args):
        Q = np.empty((self._L, self._Ntot), np.float32)
        _lib.check(call(self._h, *args, Q.ctypes.data, 0))
        return Q

    def inference(self, niter):
        """n mean-field iterations; returns a float32 ndarray (nlabels, n_pixels).

        `niter` may be float-valued: the reference passes np.float64(10.0)
        (/root/reference/03c_hsn/demo.py:159 -> utilities.py:417,442)."""
        return self._q_host(self._lib.dcrf_inference, int(niter))

    def inference_device(self, niter, out=None):
        """Same as inference() but leaves Q on the GPU as a torch tensor (no host round trip)."""
        import torch

        if out is None:
            out = torch.empty((self._L, self._Ntot), dtype=torch.float32, device="cuda:%d" % self._dev_index())
        ptr, dev, _, _ = _buffer(out, np.float32, "out")
        assert dev == 1 and tuple(out.shape) == (self._L, self._Ntot)
        _lib.check(self._lib.dcrf_inference(self._h, int(niter), ptr, 1))
        self._sync_unless_async()
        return out

    def map(self, niter):
        """inference + argmax over labels -> int32 ndarray (n_pixels,)."""
        lab = np.empty(self._Ntot, np.int32)
        _lib.check(self._lib.dcrf_map(self._h, int(niter), lab.ctypes.data, 0))
        return lab

    def map_device(self, niter, out=None):
        import torch

        if out is None:
            out = torch.empty((self._Ntot,), dtype=torch.int32, device="cuda:%d" % self._dev_index())
        _lib.check(self._lib.dcrf_map(self._h, int(niter), out.data_ptr(), 1))
        self._sync_unless_async()
        return out

    def _dev_index(self):
        if self._device >= 0:
            return self._device
        import torch

        return torch.cuda.current_device()

    def startInference(self):
        """[EXT] returns (Q, tmp1, tmp2); Q = softmax(-unary).  tmp1/tmp2 exist for signature parity."""
        _lib.check(self._lib.dcrf_start_inference(self._h))
        Q = self._q_host(self._lib.dcrf_get_q)
        return Q, np.empty_like(Q), np.empty_like(Q)

    def stepInference(self, Q, tmp1=None, tmp2=None):
        """[EXT] one mean-field update of Q in place."""
        self._check_unary(Q)
        ptr, dev, _, keep = _buffer(Q, np.float32, "Q")
        _lib.check(self._lib.dcrf_set_q(self._h, ptr, dev))
        _lib.check(self._lib.dcrf_step_inference(self._h))
        _lib.check(self._lib.dcrf_get_q(self._h, ptr, dev))
        if dev:
            self._sync_unless_async()
        del keep

    def klDivergence(self, Q):
        self._check_unary(Q)
        ptr, dev, _, keep = _buffer(Q, np.float32, "Q")
        _lib.check(self._lib.dcrf_set_q(self._h, ptr, dev))
        kl = C.c_double(0.0)
        _lib.check(self._lib.dcrf_kl_divergence(self._h, C.byref(kl)))
        del keep
        return kl.value

    # -- measurement hooks (bench.py) --
    def profile_enable(self, on=True):
        _lib.check(self._lib.dcrf_profile_enable(self._h, 1 if on else 0))

    def profile_read(self, kernel_class, tag=-1, reset=False):
        """-> (total device ms, launches) of one kernel class since the last reset."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        _lib.check(self._lib.dcrf_profile_read(self._h, int(kernel_class), int(tag), C.byref(ms), C.byref(n),
                                               1 if reset else 0))
        return ms.value, n.value

    # -- introspection (tests) --
    def num_pairwise(self):
        n = C.c_int(0)
        _lib.check(self._lib.dcrf_num_pairwise(self._h, C.byref(n)))
        return n.value

    def lattice_info(self, kernel):
        d, M = C.c_int(0), C.c_int64(0)
        per = np.zeros(self._B, np.int64)
        _lib.check(self._lib.dcrf_lattice_info(self._h, int(kernel), C.byref(d), C.byref(M), per.ctypes.data))
        return d.value, M.value, per

    def lattice_export(self, kernel, image=0, with_norm=True):
        """dict(keys, offsets, bary, neighbours, norm, M, d) in the reference vertex numbering."""
        d, _, per = self.lattice_info(kernel)
        Mb, Nb = int(per[image]), self._sizes[image][0] * self._sizes[image][1]
        out = dict(d=d, M=Mb,
                   keys=np.zeros((Mb, d), np.int16), offsets=np.zeros((Nb, d + 1), np.int32),
                   bary=np.zeros((Nb, d + 1), np.float32), neighbours=np.zeros((d + 1, Mb, 2), np.int32),
                   norm=np.zeros(Nb, np.float32) if with_norm else None)
        _lib.check(self._lib.dcrf_lattice_export(
            self._h, int(kernel), int(image), out["keys"].ctypes.data, out["offsets"].ctypes.data,
            out["bary"].ctypes.data, out["neighbours"].ctypes.data,
            out["norm"].ctypes.data if with_norm else None))
        return out

    def lattice_filter(self, kernel, values):
        v = np.ascontiguousarray(values, np.float32)
        assert v.ndim == 2 and v.shape[1] == self._Ntot
        out = np.empty_like(v)
        _lib.check(self._lib.dcrf_lattice_filter(self._h, int(kernel), v.ctypes.data, out.ctypes.data, v.shape[0]))
        return out


class DenseCRF(_Model):
    """[EXT] `pydensecrf.densecrf.DenseCRF(nvar, nlabels)`."""

    def __init__(self, nvar, nlabels, device=None, stream=None):
        super(DenseCRF, self).__init__(device, stream)
        self._L, self._Ntot, self._B = int(nlabels), int(nvar), 1
        self._sizes = [(int(nvar), 1)]
        h = C.c_void_p()
        _lib.check(self._lib.dcrf_create_nd(int(nvar), int(nlabels), self._device, self._stream_ptr, C.byref(h)))
        self._created(h)


class DenseCRF2D(_Model):
    """`pydensecrf.densecrf.DenseCRF2D(w, h, nlabels)` -- width first
    (/root/reference/03c_hsn/utilities.py:427)."""

    def __init__(self, w, h, nlabels, device=None, stream=None):
        super(DenseCRF2D, self).__init__(device, stream)
        self._W, self._H, self._L = int(w), int(h), int(nlabels)
        self._Ntot, self._B = self._W * self._H, 1
        self._sizes = [(self._W, self._H)]
        hd = C.c_void_p()
        _lib.check(self._lib.dcrf_create(self._W, self._H, self._L, self._device, self._stream_ptr, C.byref(hd)))
        self._created(hd)

    def addPairwiseGaussian(self, sxy, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        sx, sy = _pair(sxy, 2, "sxy")
        kind, c = _compat(compat, self._L)
        _lib.check(self._lib.dcrf_add_pairwise_gaussian(self._h, sx, sy, kind, c.ctypes.data, int(kernel),
                                                        int(normalization)))

    def addPairwiseBilateral(self, sxy, srgb, rgbim, compat, kernel=DIAG_KERNEL,
                             normalization=NORMALIZE_SYMMETRIC):
        sx, sy = _pair(sxy, 2, "sxy")
        sr, sg, sb = _pair(srgb, 3, "srgb")
        if rgbim is None:
            raise TypeError("Argument 'rgbim' must not be None")
        shape = tuple(rgbim.shape)
        if shape != (self._H, self._W, 3):
            raise ValueError("Bad shape for pairwise bilateral (Need {}, got {})".format((self._H, self._W, 3), shape))
        ptr, dev, _, keep = _buffer(rgbim, np.uint8, "rgbim")
        kind, c = _compat(compat, self._L)
        self._pre_device_input(dev)
        _lib.check(self._lib.dcrf_add_pairwise_bilateral(self._h, sx, sy, sr, sg, sb, ptr, dev, kind,
                                                         c.ctypes.data, int(kernel), int(normalization)))
        if dev:
            self._sync_unless_async()
        del keep


class DenseCRFBatch(_Model):
    """Many independent images (shared label count) behind one handle: every kernel launch covers the
    whole batch.  Replaces the serial per-image loops of the reference
    (/root/reference/03c_hsn/utilities.py:424, 03a_sec-dsrg/SEC.py:274).

    Unaries / images / outputs are passed "concatenated": per-image (L, N_b) float32 blocks (or
    (H_b, W_b, 3) uint8 images) laid back to back in one flat buffer, or as a list that is
    concatenated here."""

    def __init__(self, sizes, nlabels, device=None, stream=None):
        super(DenseCRFBatch, self).__init__(device, stream)
        self._sizes = [(int(w), int(h)) for (w, h) in sizes]
        self._B, self._L = len(self._sizes), int(nlabels)
        self._npix = np.array([w * h for (w, h) in self._sizes], np.int64)
        self._Ntot = int(self._npix.sum())
        w = np.array([s[0] for s in self._sizes], np.int32)
        h = np.array([s[1] for s in self._sizes], np.int32)
        hd = C.c_void_p()
        _lib.check(self._lib.dcrf_create_batch(self._B, w.ctypes.data, h.ctypes.data, self._L, self._device,
                                               self._stream_ptr, C.byref(hd)))
        self._created(hd)

    def _flat(self, x, np_dtype, per_image_elems, what):
        if isinstance(x, (list, tuple)):
            parts = [np.ascontiguousarray(a, dtype=np_dtype).ravel() for a in x]
            if len(parts) != self._B:
                raise ValueError("%s: expected %d images, got %d" % (what, self._B, len(parts)))
            for b, a in enumerate(parts):
                if a.size != per_image_elems[b]:
                    raise ValueError("%s: image %d has %d elements, expected %d" % (what, b, a.size, per_image_elems[b]))
            x = np.concatenate(parts) if parts else np.zeros(0, np_dtype)
        n = int(np.prod(tuple(x.shape)))
        if n != int(sum(per_image_elems)):
            raise ValueError("%s: expected %d elements in total, got %d" % (what, int(sum(per_image_elems)), n))
        return x

    def setUnaryEnergy(self, u, f=None):
        ptr, dev, keep = self._input(u, np.float32, self._npix * self._L, "unary")
        self._pre_device_input(dev)
        _lib.check(self._lib.dcrf_set_unary(self._h, ptr, dev))
        if dev:
            self._sync_unless_async()
        del keep

    def addPairwiseEnergy(self, *a, **k):
        raise NotImplementedError("addPairwiseEnergy is single-image only")

    def _input(self, x, np_dtype, per_image_elems, what):
        """Host (ndarray / list of arrays) or device (torch CUDA tensor / list of them) input ->
        (pointer, on_device, keepalive).  Host data is made contiguous and cast; device tensors must
        already have the right dtype and be contiguous (no hidden device copies)."""
        total = int(sum(per_image_elems))
        if isinstance(x, (list, tuple)) and len(x) and _is_torch(x[0]):
            import torch

            if len(x) != self._B:
                raise ValueError("%s: expected %d images, got %d" % (what, self._B, len(x)))
            x = torch.cat([t.reshape(-1) for t in x])
        if _is_torch(x):
            if x.numel() != total:
                raise ValueError("%s: expected %d elements in total, got %d" % (what, total, x.numel()))
            ptr, dev, _, keep = _buffer(x, np_dtype, what)
            return ptr, dev, keep
        x = self._flat(x, np_dtype, per_image_elems, what)
        a = np.ascontiguousarray(x, dtype=np_dtype).reshape(-1)
        return a.ctypes.data, 0, a

    # -- unary construction on the GPU (the NumPy glue of the reference's wrappers) --
    def setUnaryFromSoftmax(self, probs, scale=None, clip=1e-5):
        """`setUnaryEnergy(unary_from_softmax(probs, scale, clip))` without the host-side log:
        probs = per-image (L, ...) class probabilities (list, one concatenated array, or a torch CUDA
        tensor that then never leaves the GPU), float64 or float32 (03c_hsn/utilities.py:431-432)."""
        first = probs[0] if isinstance(probs, (list, tuple)) else probs
        is64 = str(first.dtype).endswith("float64")
        dt = np.float64 if is64 else np.float32
        if scale is not None and not 0 < scale <= 1:
            raise AssertionError("`scale` needs to be in (0,1]")
        ptr, dev, keep = self._input(probs, dt, self._npix * self._L, "probs")
        self._pre_device_input(dev)
        _lib.check(self._lib.dcrf_set_unary_from_probs(
            self._h, ptr, 1 if is64 else 0, 1.0 if scale is None else float(scale),
            0.0 if clip is None else float(clip), 0 if clip is None else 1, dev))
        if dev:
            self._sync_unless_async()
        del keep

    def setUnaryFromLogits(self, feats, use_log=True):
        """Unary of SEC/DSRG's crf_inference (use_log=True): feats = per-image (H, W, L) float32
        feature maps (host arrays or a CUDA tensor); U = -log softmax_L(feat)."""
        if not use_log:
            # [EXT] lib/crf.py is not in the reference tree and no call site passes use_log=False
            # (SEC.py:275, DSRG.py:328, model.py:689,693 all use the default): not guessed here
            raise NotImplementedError("crf_inference(use_log=False) is not exercised by the reference")
        ptr, dev, keep = self._input(feats, np.float32, self._npix * self._L, "feats")
        self._pre_device_input(dev)
        _lib.check(self._lib.dcrf_set_unary_from_logits(self._h, ptr, 1, dev))
        if dev:
            self._sync_unless_async()
        del keep

    def setUnaryFromLabels(self, labels, gt_prob, zero_unsure=True):
        """`setUnaryEnergy(unary_from_labels(labels, L, gt_prob, zero_unsure))` on the GPU; labels:
        int32 (host arrays of any integer type are cast, CUDA tensors must be int32)."""
        ptr, dev, keep = self._input(labels, np.int32, self._npix, "labels")
        self._pre_device_input(dev)
        _lib.check(self._lib.dcrf_set_unary_from_labels(self._h, ptr, float(gt_prob), 1 if zero_unsure else 0, dev))
        del keep

    def addPairwiseGaussian(self, sxy, compat, kernel=DIAG_KERNEL, normalization=NORMALIZE_SYMMETRIC):
        sx, sy = _pair(sxy, 2, "sxy")
        kind, c = _compat(compat, self._L)
        _lib.check(self._lib.dcrf_add_pairwise_gaussian(self._h, sx, sy, kind, c.ctypes.data, int(kernel),
                                                        int(normalization)))

    def addPairwiseBilateral(self, sxy, srgb, rgbim, compat, kernel=DIAG_KERNEL,
                             normalization=NORMALIZE_SYMMETRIC):
        sx, sy = _pair(sxy, 2, "sxy")
        sr, sg, sb = _pair(srgb, 3, "srgb")
        ptr, dev, keep = self._input(rgbim, np.uint8, self._npix * 3, "rgbim")
        kind, c = _compat(compat, self._L)
        self._pre_device_input(dev)
        _lib.check(self._lib.dcrf_add_pairwise_bilateral(self._h, sx, sy, sr, sg, sb, ptr, dev, kind,
                                                         c.ctypes.data, int(kernel), int(normalization)))
        if dev:
            self._sync_unless_async()
        del keep

    def _split(self, flat, per_elem, shape_fn):
        out, o = [], 0
        for b in range(self._B):
            n = int(self._npix[b]) * per_elem
            out.append(flat[o:o + n].reshape(shape_fn(b)))
            o += n
        return out

    def inference(self, niter, out=None):
        """-> list of (L, N_b) float32 arrays (views of one flat host buffer; `out` may supply it,
        e.g. a numpy view of pinned memory)."""
        flat = np.empty(self._Ntot * self._L, np.float32) if out is None else out
        assert flat.dtype == np.float32 and flat.size == self._Ntot * self._L and flat.flags.c_contiguous
        flat = flat.reshape(-1)
        _lib.check(self._lib.dcrf_inference(self._h, int(niter), flat.ctypes.data, 0))
        return self._split(flat, self._L, lambda b: (self._L, int(self._npix[b])))

    def inference_device(self, niter, out=None):
        import torch

        if out is None:
            out = torch.empty((self._Ntot * self._L,), dtype=torch.float32, device="cuda:%d" % self._dev_index())
        _lib.check(self._lib.dcrf_inference(self._h, int(niter), out.data_ptr(), 1))
        self._sync_unless_async()
        return out

    def map(self, niter, out=None, dtype=np.int32):
        """-> list of (H_b, W_b) label maps (views of `out` when given), int32 or uint8 (`dtype`, or
        the dtype of `out`): uint8 is what leaves the GPU cheapest."""
        flat = np.empty(self._Ntot, dtype) if out is None else out
        assert flat.dtype in (np.int32, np.uint8) and flat.size == self._Ntot and flat.flags.c_contiguous
        flat = flat.reshape(-1)
        fn = self._lib.dcrf_map if flat.dtype == np.int32 else self._lib.dcrf_map_u8
        _lib.check(fn(self._h, int(niter), flat.ctypes.data, 0))
        return self._split(flat, 1, lambda b: (self._sizes[b][1], self._sizes[b][0]))

    def map_device(self, niter, out=None, dtype=None):
        """inference + argmax into a CUDA tensor of Ntot labels (int32, or uint8 when `out` / `dtype` say so)."""
        import torch

        if out is None:
            out = torch.empty((self._Ntot,), dtype=dtype or torch.int32, device="cuda:%d" % self._dev_index())
        assert out.is_cuda and out.is_contiguous() and out.numel() == self._Ntot
        assert out.dtype in (torch.int32, torch.uint8)
        fn = self._lib.dcrf_map if out.dtype == torch.int32 else self._lib.dcrf_map_u8
        _lib.check(fn(self._h, int(niter), out.data_ptr(), 1))
        self._sync_unless_async()
        return out

    # -- split form of inference()/map(): iterate now, download later (pipeline.py) --
    def run(self, niter):
        """startInference + niter steps; the running Q stays inside the handle."""
        _lib.check(self._lib.dcrf_run(self._h, int(niter)))

    def marginals(self, out=None):
        """Download the running Q -> list of (L, N_b) float32 arrays (views of `out` when given)."""
        flat = np.empty(self._Ntot * self._L, np.float32) if out is None else out
        assert flat.dtype == np.float32 and flat.size == self._Ntot * self._L and flat.flags.c_contiguous
        flat = flat.reshape(-1)
        _lib.check(self._lib.dcrf_get_q(self._h, flat.ctypes.data, 0))
        return self._split(flat, self._L, lambda b: (self._L, int(self._npix[b])))

    def marginals_hwc(self, out=None, min_prob=None, log=False):
        """Download the running Q as a list of (H_b, W_b, L) float32 arrays -- the layout SEC / DSRG's
        `crf_inference` returns (03a_sec-dsrg/SEC.py:275) -- with no host transpose.  `min_prob`
        applies `ret[ret < min_prob] = min_prob; ret /= ret.sum(-1, keepdims=True)` and `log` the
        final `np.log` of the `crf` closure (SEC.py:277-279) on the GPU."""
        flat = np.empty(self._Ntot * self._L, np.float32) if out is None else out
        assert flat.dtype == np.float32 and flat.size == self._Ntot * self._L and flat.flags.c_contiguous
        flat = flat.reshape(-1)
        _lib.check(self._lib.dcrf_get_q_hwc(self._h, float(min_prob) if min_prob else 0.0, 1 if log else 0,
                                            flat.ctypes.data, 0))
        return self._split(flat, self._L, lambda b: (self._sizes[b][1], self._sizes[b][0], self._L))

    def marginals_hwc_device(self, out=None, min_prob=None, log=False):
        """Same, into a float32 CUDA tensor of Ntot * L elements (images back to back)."""
        import torch

        if out is None:
            out = torch.empty((self._Ntot * self._L,), dtype=torch.float32, device="cuda:%d" % self._dev_index())
        _lib.check(self._lib.dcrf_get_q_hwc(self._h, float(min_prob) if min_prob else 0.0, 1 if log else 0,
                                            out.data_ptr(), 1))
        self._sync_unless_async()
        return out

    def labels(self, out=None, dtype=np.int32):
        """Download argmax of the running Q -> list of (H_b, W_b) int32 / uint8 label maps."""
        flat = np.empty(self._Ntot, dtype) if out is None else out
        assert flat.dtype in (np.int32, np.uint8) and flat.size == self._Ntot and flat.flags.c_contiguous
        flat = flat.reshape(-1)
        fn = self._lib.dcrf_get_labels if flat.dtype == np.int32 else self._lib.dcrf_get_labels_u8
        _lib.check(fn(self._h, flat.ctypes.data, 0))
        return self._split(flat, 1, lambda b: (self._sizes[b][1], self._sizes[b][0]))

    def startInference(self):
        raise NotImplementedError("stepping API is single-image only")

    stepInference = klDivergence = startInference


def trim_memory():
    """Return the device memory cached by the library's per-stream pools to the driver."""
    _lib.check(_lib.load().dcrf_trim_memory())


def launch_count():
    """Kernels launched by libdcrf_b200.so in this process (bench.py reports it as gpu_launches)."""
    return int(_lib.load().dcrf_launch_count())


def copy_count():
    """(host->device bytes, device->host bytes) copied by libdcrf_b200.so in this process so far."""
    a, b = C.c_int64(0), C.c_int64(0)
    _lib.load().dcrf_copy_count(C.byref(a), C.byref(b))
    return a.value, b.value
