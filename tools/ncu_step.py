"""One small 1080p batch through the device API (profiling target for ncu; not a benchmark)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import eppm_b200 as E
from eppm_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
h, w = 1080, 1920
a, b, _, _ = synth.make_batch(h, w, n, first_idx=0, distinct=2)
ctx = E.EppmContext(h, w, n)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
out = torch.empty((n, h, w, 2), dtype=torch.float32, device="cuda")
for _ in range(reps):
    ctx.compute_batch_device(da, db, n, out)
ctx.synchronize()
print("done", ctx.launch_count())
