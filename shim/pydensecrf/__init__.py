"""Import shim: put `<repo>/shim` on PYTHONPATH and the reference's
`import pydensecrf.densecrf as dcrf` / `from pydensecrf.utils import unary_from_softmax`
(/root/reference/03c_hsn/utilities.py:10-11) resolve to the B200 implementation unchanged."""
