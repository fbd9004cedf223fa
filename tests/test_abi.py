"""CPU: the C-ABI library loads, exports every symbol include/dcrf_b200.h declares, refuses to run
without a GPU, and the product never touches the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "dcrf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dcrf_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = _declared_functions()
    for must in ("dcrf_create", "dcrf_set_unary", "dcrf_add_pairwise_gaussian", "dcrf_add_pairwise_bilateral",
                 "dcrf_inference", "dcrf_map", "dcrf_lattice_export", "dcrf_confusion_accumulate", "dcrf_destroy"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from wsss_analysis_b200 import _lib

    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), "libdcrf_b200.so does not export %s" % name
    # and the Python loader binds exactly the declared set
    assert sorted(_lib.SIGNATURES) == _declared_functions()
    assert b"sm_100a" in _lib.load().dcrf_version()


def test_no_cpu_fallback_without_a_gpu():
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    from wsss_analysis_b200 import DenseCRF2D

    with pytest.raises(RuntimeError, match="no CUDA device"):
        DenseCRF2D(8, 8, 3)


def test_missing_library_fails_loudly(monkeypatch):
    from wsss_analysis_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdcrf_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wsss_analysis_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def test_shim_resolves_reference_imports():
    import subprocess
    import sys

    code = ("import pydensecrf.densecrf as dcrf; from pydensecrf.utils import unary_from_softmax, unary_from_labels;"
            "print(dcrf.DenseCRF2D.__module__, dcrf.DIAG_KERNEL, dcrf.NORMALIZE_SYMMETRIC)")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "shim"))
    out = subprocess.check_output([sys.executable, "-c", code], env=env, cwd="/tmp").decode()
    assert out.split() == ["wsss_analysis_b200.densecrf", "1", "3"]


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 (no C++ in the signatures) and a C program
    that references every declared entry point links against the shared library with gcc alone."""
    import re
    import subprocess

    hdr = os.path.join(ROOT, "include", "dcrf_b200.h")
    names = sorted(set(re.findall(r"\b(dcrf_[a-z0-9_]+)\s*\(", open(hdr).read())))
    src = tmp_path / "link_all.c"
    body = "\n".join("    p[%d] = (fn)%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include "dcrf_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\n'
                   'int main(void) {\n    fn p[%d];\n%s\n'
                   '    printf("%%d %%s\\n", (int)(sizeof p / sizeof p[0]), dcrf_version());\n    return p[0] == 0;\n}\n'
                   % (len(names), body))
    exe = tmp_path / "link_all"
    libdir = os.path.join(ROOT, "wsss_analysis_b200", "csrc")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-ldcrf_b200", "-Wl,-rpath," + libdir],
                   check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == len(names) and out[1].startswith("dcrf_b200")
