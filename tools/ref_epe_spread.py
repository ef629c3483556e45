"""Run-to-run spread of the REFERENCE build's end-point error against ground truth (its three in-place filters race) on the pairs of
test_end_to_end_epe_vs_ground_truth_no_worse, next to this library's (deterministic) value.  Development probe, run under gpurun."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import eppm_b200 as E
from eppm_b200 import synth
from refharness import Ref
ref = Ref()
out = []
for h, w, idx in [(480, 640, 0), (436, 1024, 1), (436, 1024, 2)]:
    a, b, gt, valid = synth.make_pair(h, w, idx)
    ctx = E.EppmContext(h, w, 1)
    e_me = [synth.epe(ctx.compute_batch_host(a[None], b[None])[0], gt, valid) for _ in range(2)]
    ctx.close()
    e_ref = []
    for rep in range(8):
        rc = ref.create(h, w); ref.set_data(rc, a, b)
        e_ref.append(synth.epe(ref.compute_flow(rc, h, w), gt, valid)); ref.destroy(rc)
    rec = {"h": h, "w": w, "idx": idx, "epe_b200": e_me, "epe_reference_runs": e_ref, "ref_min": min(e_ref), "ref_max": max(e_ref), "ref_mean": float(np.mean(e_ref)),
           "delta_vs_ref_mean": e_me[0] - float(np.mean(e_ref))}
    print(json.dumps(rec)); out.append(rec)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_ref_epe_spread.json"), "w"), indent=1)
