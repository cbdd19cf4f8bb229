"""aswstereomatch_b200 -- B200-native (sm_100a CUDA) dense-matching hot path of aswStereoMatch.

The product is the C-ABI shared library (include/asw/asw.h, built from csrc/ into
libasw_b200.so); this package is its host-side mirror of the reference's method entry points.
"""
from .api import (ADAPTIVE_WEIGHT, ADAPTIVE_WEIGHT_8DIRECT, ADAPTIVE_WEIGHT_BILATERAL_GRID, ADAPTIVE_WEIGHT_BLO1,  # noqa: F401
                  ADAPTIVE_WEIGHT_GEODESIC, ADAPTIVE_WEIGHT_GUIDED_FILTER, ADAPTIVE_WEIGHT_GUIDED_FILTER_2,
                  ADAPTIVE_WEIGHT_GUIDED_FILTER_3, ADAPTIVE_WEIGHT_MEDIAN, BM, DISPARITY_LEFT, DISPARITY_RIGHT, NCC, SGBM,
                  AswError, Batch, Context, EXPORTS, Pool, LIB_PATH, load_library, method_candidates, pinned_empty)
