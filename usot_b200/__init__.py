"""usot_b200 -- Blackwell-native (sm_100a) forward path of the USOT Siamese tracker.

Public surface:
    usot_b200.USOT / USOT_          drop-in for lib.models.models.USOT (see lib/models/models.py in this repo)
    usot_b200.Engine                handle on the C-ABI engine
    usot_b200.ops                   prroi_pool2d / xcorr_depthwise / groupdw_xcorr / conv2d_nhwc
    usot_b200.build.build()         in-tree nvcc build of libusot_b200.so
"""
from .models import USOT, USOT_  # noqa: F401
from .engine import Engine, feature_size  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["USOT", "USOT_", "Engine", "feature_size", "ops"]
