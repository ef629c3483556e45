"""ctypes wrapper around oracle/_ref/libeppm_ref.so — the UNMODIFIED reference compiled for sm_100a (oracle/Makefile).
Test infrastructure only: nothing in the product imports this."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libeppm_ref.so")
REF_DATA = os.path.join(ROOT, "oracle", "_ref", "data")


def available():
    return os.path.exists(REF_LIB)


def pitched(arr, align=512):
    """Host array [h,w(,c)] -> device tensor whose rows are padded to `align` bytes (cudaMallocPitch-like); returns (tensor, pitch)."""
    import torch
    h = arr.shape[0]
    row = arr.reshape(h, -1).view(np.uint8) if arr.dtype != np.uint8 else arr.reshape(h, -1)
    pitch = (row.shape[1] + align - 1) // align * align
    buf = np.zeros((h, pitch), np.uint8)
    buf[:, : row.shape[1]] = row
    return torch.from_numpy(buf).cuda(), pitch


def variant_lib(name):
    return os.path.join(ROOT, "oracle", "_ref", f"libeppm_ref_{name}.so")


class Ref:
    def __init__(self, lib_path=None):
        lib = C.CDLL(lib_path or REF_LIB)
        P, I, S = C.c_void_p, C.c_int, C.c_size_t
        lib.ref_create.restype = P; lib.ref_create.argtypes = [I, I]
        lib.ref_destroy.argtypes = [P]
        lib.ref_set_data.argtypes = [P, P, P]
        lib.ref_compute_flow.argtypes = [P, P]
        lib.ref_time_pair.restype = C.c_float; lib.ref_time_pair.argtypes = [P, P, P, P]
        lib.ref_read_plane.restype = C.c_long; lib.ref_read_plane.argtypes = [P, I, I, P]
        lib.ref_level_dims.argtypes = [P, I, C.POINTER(I), C.POINTER(I)]
        lib.ref_num_levels.argtypes = [P]
        lib.ref_tap_patchmatch.argtypes = [P] * 6 + [I, I, S, S, S, S, I]
        lib.ref_tap_pm_step.argtypes = [I, I] + [P] * 6 + [I, I, S, S, S, S]
        lib.ref_tap_c2f_refine.argtypes = [P] * 5 + [I, I, S, S]
        lib.ref_probe_texture.argtypes = [P, P, I, P]
        if hasattr(lib, "ref_calc_flow_error"):
            lib.ref_calc_flow_error.argtypes = [P, P, I, I, I, I, P, P, P]
            lib.ref_save_flo.argtypes = [C.c_char_p, P, I, I]
        for name, args in {
            "baoCudaLeftRightCheck": [P, P, P, P, I, I, S, S],
            "baoCudaOutlierRemoval": [P, P, I, I, S, S],
            "baoCudaWeightedMedianFilter": [P, P, P, I, I, S, S, S, I, C.c_bool],
            "baoCudaFillHole": [P, P, P, I, I, S, S, S],
            "baoCudaNNF2Flow": [P, P, I, I, S, S],
            "baoCudaFlowSmoothing": [P, P, I, I, S, S],
            "baoCudaCensusTransform": [P, P, P, P, I, I, S, S],
        }.items():
            getattr(lib, name).argtypes = args
            getattr(lib, name).restype = None
        self.lib = lib

    def shim_stats(self):
        """(texture binds issued, texture objects created) by the texture-reference shim since the library was loaded."""
        if not hasattr(self.lib, "ref_shim_stats"):
            return None
        b, c = C.c_ulonglong(), C.c_ulonglong()
        self.lib.ref_shim_stats(C.byref(b), C.byref(c))
        return int(b.value), int(c.value)

    def calc_flow_error(self, flow, gt, border=0, error_thresh=3):
        """bao_calc_flow_error + bao_calc_flow_error_percentage of the reference on interleaved [h,w,2] float32 arrays -> (epe, aae_deg, outlier_frac)."""
        flow = np.ascontiguousarray(flow, np.float32); gt = np.ascontiguousarray(gt, np.float32)
        e, a, o = C.c_float(), C.c_float(), C.c_float()
        self.lib.ref_calc_flow_error(flow.ctypes.data, gt.ctypes.data, flow.shape[0], flow.shape[1], border, error_thresh, C.byref(e), C.byref(a), C.byref(o))
        return e.value, a.value, o.value

    def save_flo(self, path, flow):
        flow = np.ascontiguousarray(flow, np.float32)
        self.lib.ref_save_flo(str(path).encode(), flow.ctypes.data, flow.shape[0], flow.shape[1])

    def create(self, h, w):
        return self.lib.ref_create(h, w)

    def destroy(self, ctx):
        self.lib.ref_destroy(ctx)

    def set_data(self, ctx, img1, img2):
        self.lib.ref_set_data(ctx, img1.ctypes.data, img2.ctypes.data)

    def compute_flow(self, ctx, h, w):
        out = np.zeros((h, w, 2), np.float32)
        self.lib.ref_compute_flow(ctx, out.ctypes.data)
        return out

    def time_pair(self, ctx, img1, img2, h, w):
        out = np.zeros((h, w, 2), np.float32)
        ms = self.lib.ref_time_pair(ctx, img1.ctypes.data, img2.ctypes.data, out.ctypes.data)
        return float(ms), out

    def level_dims(self, ctx, level):
        h, w = C.c_int(), C.c_int()
        self.lib.ref_level_dims(ctx, level, C.byref(h), C.byref(w))
        return h.value, w.value

    def num_levels(self, ctx):
        return self.lib.ref_num_levels(ctx)

    def read_plane(self, ctx, which, level):
        h, w = self.level_dims(ctx, level)
        shape, dt = {0: ((h, w, 4), np.uint8), 1: ((h, w, 4), np.uint8), 2: ((h, w), np.uint8), 3: ((h, w), np.uint8),
                     4: ((h, w, 2), np.int16), 5: ((h, w, 2), np.int16), 6: ((h, w), np.float32), 7: ((h, w), np.float32),
                     8: ((h, w, 2), np.float32)}[which]
        out = np.zeros(shape, dt)
        n = self.lib.ref_read_plane(ctx, which, level, out.ctypes.data)
        assert n == out.nbytes, (n, out.nbytes)
        return out

    def tap_patchmatch(self, img1, img2, cen1, cen2, w, h, n_steps):
        """img*/cen* = (device tensor, pitch) pairs from pitched(); returns host (nnf int16 [h,w,2], cost f32 [h,w])."""
        import torch
        nnf = torch.zeros((h, w, 2), dtype=torch.int16, device="cuda")
        cost = torch.zeros((h, w), dtype=torch.float32, device="cuda")
        rc = self.lib.ref_tap_patchmatch(nnf.data_ptr(), cost.data_ptr(), img1[0].data_ptr(), img2[0].data_ptr(), cen1[0].data_ptr(),
                                         cen2[0].data_ptr(), w, h, img1[1], w * 4, w * 4, cen1[1], n_steps)
        assert rc == 0, rc
        return nnf.cpu().numpy(), cost.cpu().numpy()
