"""Generates tests/golden/*.npz from the REFERENCE's own CUDA build (oracle/_ref) on a B200 (run under gpurun, then copy
gpurun_out/golden/*.npz into tests/golden/).  The fixtures pin the CPU oracle (oracle/golden.cpp) and libeppm_b200 to the
reference without needing /root/reference or a GPU at test time.  Inputs are regenerated from eppm_b200.synth (seeded)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from eppm_b200 import synth
from refharness import Ref, pitched

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)
ref = Ref()

CASES = [("s128x96", 96, 128, 7, 0.12), ("s161x121", 121, 161, 3, 0.12), ("s256x192", 192, 256, 11, 0.2)]
for name, h, w, idx, scale in CASES:
    a, b, gt, valid = synth.make_pair(h, w, idx, scale_to=scale)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    out = {"h": h, "w": w, "pair_idx": idx, "scale_to": scale}
    dims = []
    for l in range(3):
        dims.append(ref.level_dims(rc, l))
        for which, nm in ((0, "rgba1"), (1, "rgba2"), (2, "census1"), (3, "census2")):
            out[f"{nm}_L{l}"] = ref.read_plane(rc, which, l)
    out["dims"] = np.array(dims, np.int32)
    hc, wc = dims[2]
    i1 = pitched(out["rgba1_L2"]); i2 = pitched(out["rgba2_L2"]); c1 = pitched(out["census1_L2"]); c2 = pitched(out["census2_L2"])
    nnf0, _ = ref.tap_patchmatch(i1, i2, c1, c2, wc, hc, 1)
    out["rand_field"] = nnf0
    nnf1, cost1 = ref.tap_patchmatch(i1, i2, c1, c2, wc, hc, 2)
    out["cost_init_fwd"] = cost1
    nnf3, cost3 = ref.tap_patchmatch(i1, i2, c1, c2, wc, hc, 3)   # after the first row-forward pass
    out["nnf_after_rowfwd"] = nnf3; out["cost_after_rowfwd"] = cost3
    nnfF, costF = ref.tap_patchmatch(i1, i2, c1, c2, wc, hc, 1000)
    nnfB, costB = ref.tap_patchmatch(i2, i1, c2, c1, wc, hc, 1000)
    out["nnf_pm_fwd"] = nnfF; out["cost_pm_fwd"] = costF; out["nnf_pm_bwd"] = nnfB; out["cost_pm_bwd"] = costB
    flow = ref.compute_flow(rc, h, w)
    out["nnf_consistency_fwd"] = ref.read_plane(rc, 4, 2)
    out["nnf_lr_bwd"] = ref.read_plane(rc, 5, 2)
    out["flow_L2"] = ref.read_plane(rc, 8, 2)
    out["flow_L1"] = ref.read_plane(rc, 8, 1)
    out["flow"] = flow
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **out)
    print(name, "saved", {k: getattr(v, "shape", v) for k, v in out.items() if k.startswith("flow") or k == "dims"})
    ref.destroy(rc)
