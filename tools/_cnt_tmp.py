import os, sys, ctypes as C
sys.path.insert(0, '/root/repo')
os.environ['EPPM_LIB_PATH'] = '/root/repo/build/ab/libeppm_b200_cnt.so'
import numpy as np, torch
import eppm_b200 as E
from eppm_b200 import synth, _lib
lib = _lib.load()
lib.eppm_debug_prop_counts.argtypes = [C.POINTER(C.c_ulonglong * 4), C.c_int]
h, w, n = 1080, 1920, 2
a, b, _, _ = synth.make_batch(h, w, n, first_idx=0, distinct=2)
ctx = E.EppmContext(h, w, n)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
ctx.stage_prepare(da, db, n)
buf = (C.c_ulonglong * 4)()
prev = [0, 0, 0, 0]
for step in range(1, 52):
    ctx.stage_patchmatch_partial(step)   # reruns from scratch each time; counts are cumulative per call
    lib.eppm_debug_prop_counts(C.byref(buf), 1)
    cur = list(buf)
    d = [c - p for c, p in zip(cur, prev)]
    prev = cur
    kind = (step - 1) % 5 if step > 1 else -1
    if d[0] + d[1] > 0:
        print(f"step {step:2d} it {(step-2)//5} pass {(step-2)%5}: thread eval frac {d[0]/(d[0]+d[1]):.3f}  warp-steps with any eval {d[2]/max(d[3],1):.3f}")
