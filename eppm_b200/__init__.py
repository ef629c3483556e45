"""eppm_b200 — B200-native (sm_100a) implementation of EPPM's dense-correspondence hot path.

The product is the C-ABI shared library eppm_b200/libeppm_b200.so (include/eppm.h, include/eppm_legacy_abi.h) built
from eppm_b200/csrc; this package is its thin ctypes binding plus host-side helpers (synthetic pairs, .flo IO)."""
from .api import (EppmContext, EppmError, BaoFlowPatchmatchMultiscaleCuda, default_params, write_flo, read_flo, PLANE_RGBA1, PLANE_RGBA2, PLANE_CENSUS1,
                  PLANE_CENSUS2, PLANE_NNF_FWD, PLANE_NNF_BWD, PLANE_COST_FWD, PLANE_COST_BWD, PLANE_FLOW, PLANE_FLOW_TMP)

__all__ = ["EppmContext", "EppmError", "BaoFlowPatchmatchMultiscaleCuda", "default_params"]
