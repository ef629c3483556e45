"""Can the PatchMatch stage of one chunk (L1-bound, 46 % of the issue slots) run BESIDE the coarse-to-fine stage of another chunk
(issue-bound, 71 %) on the same SMs?  Two contexts on two streams: A loops the coarse-to-fine stage, B loops PatchMatch; timed alone and
together.  Knobs (environment, read at context creation): EPPM_PM_PAD_KB = residency cap of B's scoring kernels, priorities via
EPPM_STREAM_PRIORITY.  Development probe, run under gpurun:  python tools/overlap_probe.py [pairs] [pad_kb] [prio_b]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import eppm_b200 as E
from eppm_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pad = int(sys.argv[2]) if len(sys.argv) > 2 else 0
prio_b = int(sys.argv[3]) if len(sys.argv) > 3 else 0
h, w = 1080, 1920
a, b, _, _ = synth.make_batch(h, w, n, first_idx=0, distinct=2)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
os.environ["EPPM_STREAM_PRIORITY"] = "0"
A = E.EppmContext(h, w, n)
os.environ["EPPM_STREAM_PRIORITY"] = str(prio_b)
os.environ["EPPM_PM_PAD_KB"] = str(pad)
B = E.EppmContext(h, w, n)
os.environ.pop("EPPM_PM_PAD_KB"); os.environ.pop("EPPM_STREAM_PRIORITY")
out = torch.empty((n, h, w, 2), dtype=torch.float32, device="cuda")
for c in (A, B):
    c.compute_batch_device(da, db, n, out); c.synchronize()
sa = torch.cuda.ExternalStream(A.lib.eppm_stream(A._ctx)); sb = torch.cuda.ExternalStream(B.lib.eppm_stream(B._ctx))
def run_a(k):
    for _ in range(k): A.stage_c2f(out)
def run_b(k):
    for _ in range(k): B.stage_patchmatch()
def timed(fa, fb):
    torch.cuda.synchronize()
    e0, e1a, e1b = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(torch.cuda.current_stream()); sa.wait_event(e0); sb.wait_event(e0)
    if fb: fb()
    if fa: fa()
    e1a.record(sa); e1b.record(sb)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1a), e0.elapsed_time(e1b)
ka, kb = 2, 4   # c2f ~11.5 ms/pair, PatchMatch ~4.7 ms/pair: comparable durations
ta = timed(lambda: run_a(ka), None)[0]
tb = timed(None, lambda: run_b(kb))[1]
t2 = timed(lambda: run_a(ka), lambda: run_b(kb))
res = {"pairs": n, "pad_kb": pad, "prio_b": prio_b, "c2f_alone_ms": ta, "pm_alone_ms": tb, "together_c2f_done_ms": t2[0], "together_pm_done_ms": t2[1],
       "serial_sum_ms": ta + tb, "together_max_ms": max(t2), "gain": (ta + tb) / max(t2)}
print(json.dumps(res))
