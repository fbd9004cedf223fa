import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.abspath(_os.path.join(_os.path.dirname(__file__), "..", "..")))
from wsss_analysis_b200.utils import *  # noqa: F401,F403,E402
from wsss_analysis_b200.utils import (compute_unary, create_pairwise_bilateral, create_pairwise_gaussian,  # noqa: F401,E402
                                      softmax_to_unary, unary_from_labels, unary_from_softmax)
