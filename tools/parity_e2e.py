"""End-to-end flow of this library against the reference build's flow on the same pairs (run under gpurun).

For each size: the direct flow-vs-flow end-point difference (mean / median / share above 0.5 px), the EPE of both against ground truth, in the
default mode (snapshot semantics for the three filters the reference runs in place) and with inplace_filters = 1 (the reference's update
order and launch geometry).  Also the per-stage mismatch counts of the in-place kernels fed with the reference's own intermediate state
(the stage chain of tools/parity_stages.py).  Writes gpurun_out/parity_e2e.json."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import eppm_b200 as E
from eppm_b200 import synth
from refharness import Ref

ref = Ref()
res = {}
cases = [(480, 640, 0), (436, 1024, 1), (1080, 1920, 1000), (1080, 1920, 1001)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for h, w, idx in cases:
    a, b, gt, va = synth.make_pair(h, w, idx)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    fr = ref.compute_flow(rc, h, w)
    ref.set_data(rc, a, b)
    fr2 = ref.compute_flow(rc, h, w)
    ref.destroy(rc)
    out = {"epe_gt_reference": synth.epe(fr, gt, va), "reference_run_to_run_mean": float(np.sqrt(((fr - fr2) ** 2).sum(-1)).mean())}
    for mode in (0, 1):
        p = E.default_params()
        p.inplace_filters = mode
        ctx = E.EppmContext(h, w, 1, params=p)
        fm = ctx.compute_batch_host(a[None], b[None])[0]
        ctx.close()
        d = np.sqrt(((fm.astype(np.float64) - fr.astype(np.float64)) ** 2).sum(-1))
        out["inplace" if mode else "snapshot"] = {
            "flow_vs_reference_mean_px": float(d.mean()), "median_px": float(np.median(d)), "frac_gt_0p5px": float((d > 0.5).mean()),
            "frac_identical": float((d == 0).mean()), "epe_gt": synth.epe(fm, gt, va), "epe_delta_vs_reference": synth.epe(fm, gt, va) - out["epe_gt_reference"]}
    res[f"{w}x{h}_pair{idx}"] = out
    print(f"{w}x{h} pair {idx}: {json.dumps(out)}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "parity_e2e.json"), "w"), indent=1)
