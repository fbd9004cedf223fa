import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.abspath(_os.path.join(_os.path.dirname(__file__), "..", "..")))
from wsss_analysis_b200.densecrf import *  # noqa: F401,F403,E402
from wsss_analysis_b200.densecrf import DenseCRF, DenseCRF2D, DenseCRFBatch  # noqa: F401,E402
