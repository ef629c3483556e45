"""Run the reference build on one synthetic pair of the given size (development tool; use under compute-sanitizer)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
h, w = int(sys.argv[1]), int(sys.argv[2])
ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libeppm_ref.so"))
ref.ref_create.restype = C.c_void_p; ref.ref_create.argtypes = [C.c_int, C.c_int]
ref.ref_time_pair.restype = C.c_float; ref.ref_time_pair.argtypes = [C.c_void_p] * 4
rng = np.random.default_rng(0)
a = rng.integers(0, 256, (h, w, 3), dtype=np.uint8); b = np.roll(a, (3, 5), (0, 1)).copy()
ctx = ref.ref_create(h, w)
fl = np.zeros((h, w, 2), np.float32)
print("ms", ref.ref_time_pair(ctx, a.ctypes.data, b.ctypes.data, fl.ctypes.data), "mean flow", fl.reshape(-1, 2).mean(0))
