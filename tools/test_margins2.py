import os, sys
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import eppm_b200 as E
from eppm_b200 import synth
from refharness import Ref
ref = Ref()
h, w = 436, 1024
p = E.default_params(); p.rng_mode = 1
ctx = E.EppmContext(h, w, 1, params=p)
e_me = []; e_ref = [[], [], []]
for k, idx in enumerate((1, 2, 3)):
    a, b, gt, valid = synth.make_pair(h, w, idx)
    e_me.append(synth.epe(ctx.compute_batch_host(a[None], b[None])[0], gt, valid))
    for rep in range(4):
        rc = ref.create(h, w); ref.set_data(rc, a, b); e_ref[k].append(synth.epe(ref.compute_flow(rc, h, w), gt, valid)); ref.destroy(rc)
print("philox e_me", [round(x, 4) for x in e_me], "mean", round(float(np.mean(e_me)), 4))
print("ref runs", [[round(x, 4) for x in r] for r in e_ref])
print("mean delta range", round(float(np.mean(e_me)) - float(np.mean([max(r) for r in e_ref])), 4), round(float(np.mean(e_me)) - float(np.mean([min(r) for r in e_ref])), 4), "(bar +0.05)")
# in-place mode margin
h, w = 480, 640
a, b, gt, valid = synth.make_pair(h, w, 0)
p = E.default_params(); p.inplace_filters = 1
c2 = E.EppmContext(h, w, 1, params=p); fm = c2.compute_batch_host(a[None], b[None])[0]
rc = ref.create(h, w); ref.set_data(rc, a, b); fr = ref.compute_flow(rc, h, w)
print("inplace delta", round(synth.epe(fm, gt, valid) - synth.epe(fr, gt, valid), 4), "(bar 0.05)")
