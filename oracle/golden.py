"""ctypes wrapper of oracle/libeppm_golden.so (the single-threaded CPU oracle, oracle/golden.cpp).
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg — never by eppm_b200."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libeppm_golden.so")
_lib = None

PLANES = {"rgba1": 0, "rgba2": 1, "census1": 2, "census2": 3, "nnf_fwd": 4, "nnf_bwd": 5, "cost_fwd": 6, "cost_bwd": 7, "flow": 8, "scale": 9}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise OSError(f"{_LIB} missing: run `make -C oracle golden`")
        l = C.CDLL(_LIB)
        P, I = C.c_void_p, C.c_int
        l.golden_create.restype = P; l.golden_create.argtypes = [I, I, I, I]
        l.golden_destroy.argtypes = [P]
        l.golden_set_stride.argtypes = [P, I]
        l.golden_set_pf_cost.argtypes = [P, I]
        l.golden_num_levels.argtypes = [P]
        l.golden_level_dims.argtypes = [P, I, C.POINTER(I), C.POINTER(I)]
        l.golden_prepare.argtypes = [P, P, P]
        l.golden_patchmatch.argtypes = [P, I]
        l.golden_patchmatch_scaled.argtypes = [P]
        l.golden_consistency.argtypes = [P]
        l.golden_c2f.argtypes = [P, P]
        l.golden_compute.argtypes = [P, P, P, P]
        l.golden_read_plane.restype = C.c_long; l.golden_read_plane.argtypes = [P, I, I, P]
        l.golden_write_plane.restype = C.c_long; l.golden_write_plane.argtypes = [P, I, I, P]
        l.golden_xorwow.argtypes = [C.c_ulonglong, C.c_ulonglong, I, P]
        l.golden_level_dims_for.argtypes = [I, I, I, C.POINTER(I), C.POINTER(I)]
        l.golden_patch_cost.restype = C.c_float; l.golden_patch_cost.argtypes = [P, I, I, I, I, I]
        l.golden_census_bicubic.argtypes = [P, I, I, I, I, P]
        l.golden_subpix_refine.argtypes = [P, P, P, P, P, P, I, I, I, I]
        l.golden_lr_check_buffered.argtypes = [P, P, P, P, I, I]
        l.golden_flow_to_nnf.argtypes = [P, P, I, I]
        l.golden_flow_cutoff.argtypes = [P, I, I, C.c_float]
        l.golden_eliminate_still.argtypes = [P, P, P, I, I]
        l.golden_image_smoothing.argtypes = [P, P, I, I]
        l.golden_flow_bilateral_upsample.argtypes = [P, P, I, I, P, I, C.c_float]
        _lib = l
    return _lib


def xorwow(seed, subsequence, n):
    out = np.zeros(n, np.uint32)
    lib().golden_xorwow(seed, subsequence, n, out.ctypes.data)
    return out


def census_bicubic(rgba, w_up, h_up):
    """baoCudaCensusTransform_Bicubic on one image: rgba u8 [h,w,4] -> census u8 [h_up,w_up] (oracle/golden_subpix.cpp)."""
    rgba = np.ascontiguousarray(rgba, np.uint8)
    h, w = rgba.shape[:2]
    out = np.zeros((h_up, w_up), np.uint8)
    lib().golden_census_bicubic(rgba.ctypes.data, w, h, w_up, h_up, out.ctypes.data)
    return out


def subpix_refine(rgba1, rgba2, cen1_up, cen2_up, nnf, flow, y0, y1):
    """baoCudaSubpixRefine on rows [y0, y1): returns a refined copy of flow f32 [h,w,2] (oracle/golden_subpix.cpp)."""
    a = [np.ascontiguousarray(x, np.uint8) for x in (rgba1, rgba2, cen1_up, cen2_up)]
    nnf = np.ascontiguousarray(nnf, np.int16)
    out = np.array(flow, np.float32, copy=True, order="C")
    h, w = out.shape[:2]
    lib().golden_subpix_refine(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, nnf.ctypes.data, out.ctypes.data, w, h, y0, y1)
    return out


def _c(a, dt):
    return np.array(a, dt, copy=True, order="C")


def lr_check_buffered(nnf, cost, nnf2, cost2):
    """baoCudaLeftRightCheck_Buffered: returns the four planes after the check (oracle/golden_stages.cpp)."""
    n1, c1, n2, c2 = _c(nnf, np.int16), _c(cost, np.float32), _c(nnf2, np.int16), _c(cost2, np.float32)
    h, w = c1.shape
    lib().golden_lr_check_buffered(n1.ctypes.data, c1.ctypes.data, n2.ctypes.data, c2.ctypes.data, w, h)
    return n1, c1, n2, c2


def flow_to_nnf(flow):
    f = _c(flow, np.float32); h, w = f.shape[:2]
    out = np.zeros((h, w, 2), np.int16)
    lib().golden_flow_to_nnf(f.ctypes.data, out.ctypes.data, w, h)
    return out


def flow_cutoff(flow, m):
    f = _c(flow, np.float32); h, w = f.shape[:2]
    lib().golden_flow_cutoff(f.ctypes.data, w, h, m)
    return f


def eliminate_still(flow, rgba1, rgba2):
    f, a, b = _c(flow, np.float32), _c(rgba1, np.uint8), _c(rgba2, np.uint8); h, w = f.shape[:2]
    lib().golden_eliminate_still(f.ctypes.data, a.ctypes.data, b.ctypes.data, w, h)
    return f


def image_smoothing(rgba):
    a = _c(rgba, np.uint8); h, w = a.shape[:2]
    out = np.zeros((h, w, 3), np.uint8)
    lib().golden_image_smoothing(a.ctypes.data, out.ctypes.data, w, h)
    return out


def flow_bilateral_upsample(out_init, rgba, small, ratio):
    o, a, s = _c(out_init, np.float32), _c(rgba, np.uint8), _c(small, np.float32); h, w = o.shape[:2]
    lib().golden_flow_bilateral_upsample(o.ctypes.data, a.ctypes.data, w, h, s.ctypes.data, s.shape[1], ratio)
    return o


def level_dims_for(h, w, level):
    oh, ow = C.c_int(), C.c_int()
    lib().golden_level_dims_for(h, w, level, C.byref(oh), C.byref(ow))
    return oh.value, ow.value


class Golden:
    def __init__(self, h, w, levels=3, num_iter=10, stride=2):
        self.l = lib()
        self.h, self.w = h, w
        self.ctx = self.l.golden_create(h, w, levels, num_iter)
        self.l.golden_set_stride(self.ctx, stride)
        self.num_levels = levels

    def __del__(self):
        if getattr(self, "ctx", None):
            self.l.golden_destroy(self.ctx)
            self.ctx = None

    def level_dims(self, level):
        h, w = C.c_int(), C.c_int()
        self.l.golden_level_dims(self.ctx, level, C.byref(h), C.byref(w))
        return h.value, w.value

    def prepare(self, img1, img2):
        a = np.ascontiguousarray(img1, np.uint8); b = np.ascontiguousarray(img2, np.uint8)
        self.l.golden_prepare(self.ctx, a.ctypes.data, b.ctypes.data)

    def patchmatch(self, n_steps=-1, plane_fitting=False):
        """plane_fitting: the forward direction is scored with the four-model plane-fitting cost (baoCudaPatchMatch_PlaneFitting)."""
        self.l.golden_set_pf_cost(self.ctx, 1 if plane_fitting else 0)
        self.l.golden_patchmatch(self.ctx, n_steps)
        self.l.golden_set_pf_cost(self.ctx, 0)

    def patchmatch_scaled(self):
        """baoCudaPatchMatch_Scaled (bao_pmflow_kernel.cu:1828-1895), forward direction: planes nnf_fwd, cost_fwd and scale."""
        self.l.golden_patchmatch_scaled(self.ctx)

    def consistency(self):
        self.l.golden_consistency(self.ctx)

    def c2f(self):
        out = np.zeros((self.h, self.w, 2), np.float32)
        self.l.golden_c2f(self.ctx, out.ctypes.data)
        return out

    def compute(self, img1, img2):
        a = np.ascontiguousarray(img1, np.uint8); b = np.ascontiguousarray(img2, np.uint8)
        out = np.zeros((self.h, self.w, 2), np.float32)
        self.l.golden_compute(self.ctx, a.ctypes.data, b.ctypes.data, out.ctypes.data)
        return out

    def _shape(self, which, level):
        h, w = self.level_dims(level)
        if 4 <= which <= 7 or which == 9:
            h, w = self.level_dims(self.num_levels - 1)
        return {0: ((h, w, 4), np.uint8), 1: ((h, w, 4), np.uint8), 2: ((h, w), np.uint8), 3: ((h, w), np.uint8), 4: ((h, w, 2), np.int16),
                5: ((h, w, 2), np.int16), 6: ((h, w), np.float32), 7: ((h, w), np.float32), 8: ((h, w, 2), np.float32), 9: ((h, w), np.float32)}[which]

    def plane(self, name, level=0):
        which = PLANES[name]
        shape, dt = self._shape(which, level)
        out = np.zeros(shape, dt)
        n = self.l.golden_read_plane(self.ctx, which, level, out.ctypes.data)
        if n != out.nbytes:
            raise RuntimeError(f"golden_read_plane({name}, {level}) -> {n}")
        return out

    def set_plane(self, name, arr, level=0):
        arr = np.ascontiguousarray(arr)
        n = self.l.golden_write_plane(self.ctx, PLANES[name], level, arr.ctypes.data)
        if n != arr.nbytes:
            raise RuntimeError(f"golden_write_plane({name}, {level}) -> {n}")

    def patch_cost(self, direction, x1, y1, x2, y2):
        return float(self.l.golden_patch_cost(self.ctx, direction, x1, y1, x2, y2))
