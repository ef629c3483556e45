import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import eppm_b200 as E
from eppm_b200 import synth
from refharness import Ref
ref = Ref()
h, w = 270, 480
frames, flows, valids = synth.make_stream(h, w, 4, first_idx=70, scale_to=0.25)
ctx = E.EppmContext(h, w, 3); out_h = ctx.compute_stream_host(frames); ctx.close()
for rep in range(3):
    rc = ref.create(h, w); ds = []
    for t in range(3):
        ref.set_data(rc, frames[t], frames[t + 1]); fr = ref.compute_flow(rc, h, w)
        ds.append(round(synth.epe(out_h[t], flows[t], valids[t]) - synth.epe(fr, flows[t], valids[t]), 4))
    ref.destroy(rc); print("stream delta epe (bar +0.05):", ds)
for (h, w, idx, bound) in [(480, 640, 0, 0.34), (436, 1024, 1, 0.40), (1080, 1920, 1000, 0.36)]:
    a, b, gt, valid = synth.make_pair(h, w, idx)
    ctx = E.EppmContext(h, w, 1); fm = ctx.compute_batch_host(a[None], b[None])[0]; ctx.close()
    out = []
    for rep in range(3):
        rc = ref.create(h, w); ref.set_data(rc, a, b); fr = ref.compute_flow(rc, h, w); ref.destroy(rc)
        d = np.sqrt(((fm.astype(np.float64) - fr.astype(np.float64)) ** 2).sum(-1))
        out.append((round(d.mean(), 4), round(synth.epe(fm, gt, valid) - synth.epe(fr, gt, valid), 4)))
    print((h, w, idx), "flow-vs-flow mean (bound %.2f), delta epe (bar 0.05):" % bound, out)
