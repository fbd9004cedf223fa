"""ctypes loader for libdcrf_b200.so (the C ABI declared in include/dcrf_b200.h).

There is deliberately NO fallback: if the CUDA library is missing the import of the compute path
fails loudly, and if no CUDA device is present every compute call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCRF_B200_LIB") or os.path.join(_HERE, "csrc", "libdcrf_b200.so")  # env override: tuning builds

DCRF_OK, DCRF_EINVAL, DCRF_ECUDA, DCRF_ESTATE, DCRF_ENOMEM = 0, 1, 2, 3, 4

_vp, _i, _f, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_int64

# name -> (restype, argtypes); mirrors include/dcrf_b200.h one to one
SIGNATURES = {
    "dcrf_last_error": (C.c_char_p, []),
    "dcrf_version": (C.c_char_p, []),
    "dcrf_launch_count": (_i64, []),
    "dcrf_copy_count": (None, [C.POINTER(_i64), C.POINTER(_i64)]),
    "dcrf_trim_memory": (_i, []),
    "dcrf_mem_info": (_i, [_i, C.POINTER(_i64), C.POINTER(_i64)]),
    "dcrf_stream_create": (_i, [_i, C.POINTER(_vp)]),
    "dcrf_stream_destroy": (_i, [_vp]),
    "dcrf_create": (_i, [_i, _i, _i, _i, _vp, C.POINTER(_vp)]),
    "dcrf_create_nd": (_i, [_i, _i, _i, _vp, C.POINTER(_vp)]),
    "dcrf_create_batch": (_i, [_i, _vp, _vp, _i, _i, _vp, C.POINTER(_vp)]),
    "dcrf_destroy": (None, [_vp]),
    "dcrf_set_option": (_i, [_vp, _i, _i]),
    "dcrf_get_arithmetic": (_i, [_vp, C.POINTER(_i)]),
    "dcrf_synchronize": (_i, [_vp]),
    "dcrf_set_unary": (_i, [_vp, _vp, _i]),
    "dcrf_set_unary_from_probs": (_i, [_vp, _vp, _i, C.c_double, C.c_double, _i, _i]),
    "dcrf_set_unary_from_logits": (_i, [_vp, _vp, _i, _i]),
    "dcrf_set_unary_from_labels": (_i, [_vp, _vp, _f, _i, _i]),
    "dcrf_add_pairwise_gaussian": (_i, [_vp, _f, _f, _i, _vp, _i, _i]),
    "dcrf_add_pairwise_bilateral": (_i, [_vp, _f, _f, _f, _f, _f, _vp, _i, _i, _vp, _i, _i]),
    "dcrf_add_pairwise_energy": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _i]),
    "dcrf_inference": (_i, [_vp, _i, _vp, _i]),
    "dcrf_map": (_i, [_vp, _i, _vp, _i]),
    "dcrf_map_u8": (_i, [_vp, _i, _vp, _i]),
    "dcrf_run": (_i, [_vp, _i]),
    "dcrf_get_labels": (_i, [_vp, _vp, _i]),
    "dcrf_get_labels_u8": (_i, [_vp, _vp, _i]),
    "dcrf_start_inference": (_i, [_vp]),
    "dcrf_step_inference": (_i, [_vp]),
    "dcrf_get_q": (_i, [_vp, _vp, _i]),
    "dcrf_get_q_hwc": (_i, [_vp, _f, _i, _vp, _i]),
    "dcrf_set_q": (_i, [_vp, _vp, _i]),
    "dcrf_kl_divergence": (_i, [_vp, C.POINTER(C.c_double)]),
    "dcrf_num_pairwise": (_i, [_vp, C.POINTER(_i)]),
    "dcrf_lattice_info": (_i, [_vp, _i, C.POINTER(_i), C.POINTER(_i64), _vp]),
    "dcrf_lattice_export": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "dcrf_lattice_filter": (_i, [_vp, _i, _vp, _vp, _i]),
    "dcrf_expf_ref": (_i, [_vp, _vp, _i64, _i]),
    "dcrf_profile_enable": (_i, [_vp, _i]),
    "dcrf_profile_read": (_i, [_vp, _i, _i, C.POINTER(C.c_double), C.POINTER(_i64), _i]),
    "dcrf_resize_nearest_i32": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _vp]),
    "dcrf_resize_bilinear_f32": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _vp]),
    "dcrf_confusion_accumulate": (_i, [_vp, _vp, _i64, _i, _vp, _vp, _i, _vp]),
    "dcrf_nccl_unique_id": (_i, [_vp]),
    "dcrf_nccl_comm_create": (_i, [_i, _i, _vp, _i, C.POINTER(_vp)]),
    "dcrf_nccl_comm_destroy": (_i, [_vp]),
    "dcrf_confusion_allreduce": (_i, [_vp, _vp, _i64, _i, _vp]),
}

_lib = None


class DenseCRFLibraryMissing(ImportError):
    pass


def load():
    """Load the shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DenseCRFLibraryMissing(
            "libdcrf_b200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C wsss_analysis_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Map a C return code to the Python exception the pydensecrf caller would have seen."""
    if rc == DCRF_OK:
        return
    msg = load().dcrf_last_error().decode("utf-8", "replace")
    if rc == DCRF_EINVAL:
        raise ValueError(msg)
    if rc == DCRF_ENOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)
