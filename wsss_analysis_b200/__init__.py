"""wsss_analysis_b200 -- B200-native drop-in for the DenseCRF hot path of lyndonchan/wsss-analysis.

Only what that path needs: the pydensecrf-compatible classes (`densecrf`), its host helpers
(`utils`), the reference's CRF call-site wrappers (`wsss`) and the integer mIoU reduction
(`evaluation`).  The compute lives in csrc/ (hand-written sm_100a CUDA behind the C ABI of
include/dcrf_b200.h).  Importing this package does not touch CUDA; the first compute call does.
"""
from . import densecrf, utils  # noqa: F401
from .densecrf import (CONST_KERNEL, DIAG_KERNEL, FULL_KERNEL, NO_NORMALIZATION, NORMALIZE_AFTER,  # noqa: F401
                       NORMALIZE_BEFORE, NORMALIZE_SYMMETRIC, DenseCRF, DenseCRF2D, DenseCRFBatch)

__version__ = "0.1.0"
