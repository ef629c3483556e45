"""A/B timing of the EPPM_VARIANT switches on one 1080p batch (device API, EPPM_PROFILE=1 stage events).

    python tools/variant_times.py [n_pairs] [variant ...]      # default: 8 pairs, variants 0 1 2 4 8 15

Every variant must produce the same flow bits as variant 0 (checked with a hash of the output); the per-stage device times
(prepare, patchmatch, consistency, c2f, total) and the level-0 refine / final smoothing kernel times are written to
gpurun_out/variant_times.json.  Not a benchmark: bench.py is."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["EPPM_PROFILE"] = "1"
import numpy as np
import torch

import eppm_b200 as E
from eppm_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
variants = [int(v) for v in sys.argv[2:]] or [0, 1, 2, 4, 8, 15]
h, w = int(os.environ.get("VT_H", 1080)), int(os.environ.get("VT_W", 1920))
_cache = f"/tmp/vt_cache_{h}_{w}_{n}.npz"   # the knobs read from the environment need one process per setting: generate the batch once per box
if os.path.exists(_cache):
    _z = np.load(_cache)
    a, b = _z["a"], _z["b"]
else:
    a, b, _, _ = synth.make_batch(h, w, n, first_idx=0, distinct=min(n, 4))
    np.savez(_cache, a=a, b=b)
da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
out = torch.empty((n, h, w, 2), dtype=torch.float32, device="cuda")
res = {}
ref_hash = None
for v in variants:
    os.environ["EPPM_VARIANT"] = str(v)
    ctx = E.EppmContext(h, w, n)
    best = None
    for rep in range(3):
        ctx.compute_batch_device(da, db, n, out)
        ctx.synchronize()
        st = ctx.last_stage_ms()
        km = [ctx.last_kernel_ms(0), ctx.last_kernel_ms(1)]
        st = [st[k] for k in ("prepare", "patchmatch", "consistency", "c2f", "total")]
        if best is None or st[4] < best[0][4]:
            best = (st, km)
    hsh = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
    if ref_hash is None:
        ref_hash = hsh
    st, km = best
    res[str(v)] = {"stage_ms_per_pair": [round(x / n, 3) for x in st], "refine_l0_ms_per_pair": round(km[0] / n, 3),
                   "smooth_final_ms_per_pair": round(km[1] / n, 3), "hash": hsh, "same_bits_as_first": hsh == ref_hash}
    print(v, res[str(v)], flush=True)
    ctx.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "variant_times.json"), "w"), indent=1)
