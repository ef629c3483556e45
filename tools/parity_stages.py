"""Development probe (run under gpurun): drives the SAME legacy stage functions (baoCuda*) of the reference build and of
libeppm_b200 on identical device buffers, one stage at a time, chaining the reference's outputs, and prints mismatch
statistics per stage."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import eppm_b200 as E
from eppm_b200 import synth, _lib
from refharness import Ref, pitched

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
res = {}
mine = _lib.load()
ref = Ref()
P = lambda t: t.data_ptr()


def cmp(name, a, b, floatmode=False):
    a = a.cpu().numpy(); b = b.cpu().numpy()
    ne = (a.view(np.uint32) != b.view(np.uint32)) if a.dtype == np.float32 else (a != b)
    if ne.ndim == 3:
        ne = ne.any(-1)
    out = {"mismatch_frac": float(ne.mean()), "n_mismatch": int(ne.sum())}
    if a.dtype == np.float32 and ne.any():
        fin = np.isfinite(a) & np.isfinite(b) & (np.abs(a) < 1e9) & (np.abs(b) < 1e9)
        out["max_abs"] = float(np.abs(a - b)[fin].max())
        out["mean_abs"] = float(np.abs(a - b)[fin].mean())
    print(f"{name:46s} {out}")
    res[name] = out
    return ne


def run(h, w, tag, pair_idx=0, imgs=None):
    if imgs is None:
        a, b, gt, valid = synth.make_pair(h, w, pair_idx)
    else:
        a, b = imgs
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    nl = 3
    dims = [ref.level_dims(rc, l) for l in range(nl)]
    img = [[pitched(ref.read_plane(rc, k, l)) for l in range(nl)] for k in (0, 1)]
    cen = [[pitched(ref.read_plane(rc, 2 + k, l)) for l in range(nl)] for k in (0, 1)]
    L = nl - 1
    hc, wc = dims[L]
    # PatchMatch through the legacy entry point of both libraries
    st = {}
    for name, lib in (("ref", ref.lib), ("mine", mine)):
        nnf = [torch.zeros((hc, wc, 2), dtype=torch.int16, device="cuda") for _ in range(2)]
        cost = [torch.zeros((hc, wc), dtype=torch.float32, device="cuda") for _ in range(2)]
        fn = lib.baoCudaPatchMatch
        fn.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_size_t] * 4
        fn(P(nnf[0]), P(cost[0]), P(img[0][L][0]), P(img[1][L][0]), P(cen[0][L][0]), P(cen[1][L][0]), wc, hc, img[0][L][1], wc * 4, wc * 4, cen[0][L][1])
        fn(P(nnf[1]), P(cost[1]), P(img[1][L][0]), P(img[0][L][0]), P(cen[1][L][0]), P(cen[0][L][0]), wc, hc, img[0][L][1], wc * 4, wc * 4, cen[0][L][1])
        torch.cuda.synchronize()
        st[name] = (nnf, cost)
    cmp(f"{tag} legacy PatchMatch nnf fwd", st["mine"][0][0], st["ref"][0][0])
    cmp(f"{tag} legacy PatchMatch nnf bwd", st["mine"][0][1], st["ref"][0][1])
    cmp(f"{tag} legacy PatchMatch cost fwd", st["mine"][1][0], st["ref"][1][0])
    nnf_r, cost_r = st["ref"]

    def both(stage, call, outs):
        """call(lib, bufs) mutates clones of the chained reference state; returns reference outputs for chaining."""
        r = [t.clone() for t in outs]; m = [t.clone() for t in outs]
        call(ref.lib, r); call(mine, m)
        torch.cuda.synchronize()
        for i, (x, y) in enumerate(zip(m, r)):
            cmp(f"{tag} {stage} out{i}", x, y)
        return r

    i1 = img[0][L]
    s = both("LeftRightCheck", lambda lib, t: lib.baoCudaLeftRightCheck(P(t[0]), P(t[1]), P(t[2]), P(t[3]), wc, hc, wc * 4, wc * 4),
             [nnf_r[0], cost_r[0], nnf_r[1], cost_r[1]])
    s2 = both("OutlierRemoval", lambda lib, t: lib.baoCudaOutlierRemoval(P(t[0]), P(t[1]), wc, hc, wc * 4, wc * 4), [s[0], s[1]])
    for it in (1, 2, 3, 5, 20):
        both(f"WMF iters={it}", lambda lib, t: lib.baoCudaWeightedMedianFilter(P(t[0]), P(t[1]), P(i1[0]), wc, hc, i1[1], wc * 4, wc * 4, it, True), [s2[0], s2[1]])
    s3 = both("WMF iters=20 (chain)", lambda lib, t: lib.baoCudaWeightedMedianFilter(P(t[0]), P(t[1]), P(i1[0]), wc, hc, i1[1], wc * 4, wc * 4, 20, True), [s2[0], s2[1]])
    n_occ = int(((s2[0][..., 0] < 0) | (s2[0][..., 1] < 0)).sum())
    n_left = int(((s3[0][..., 0] < 0) | (s3[0][..., 1] < 0)).sum())
    print(f"   occluded before WMF {n_occ} ({n_occ / (wc * hc):.3f}), after {n_left}")
    s4 = both("FillHole", lambda lib, t: lib.baoCudaFillHole(P(t[0]), P(t[1]), P(i1[0]), wc, hc, i1[1], wc * 4, wc * 4), [s3[0], s3[1]])
    flow = torch.zeros((hc, wc, 2), dtype=torch.float32, device="cuda")
    s5 = both("NNF2Flow", lambda lib, t: lib.baoCudaNNF2Flow(P(t[0]), P(t[1]), wc, hc, wc * 4, wc * 8), [flow, s4[0]])
    cur = s5[0]
    # C2F, level by level, refine (with upsample) and smoothing separately
    PtrArr = C.c_void_p * nl
    IntArr = C.c_int * nl
    SzArr = C.c_size_t * nl
    for l in range(L - 1, -1, -1):
        hl, wl = dims[l]
        fine = torch.zeros((hl, wl, 2), dtype=torch.float32, device="cuda")

        def c2f(lib, t):
            flows = [None] * nl
            flows[l] = t[0]; flows[l + 1] = t[1]
            fp = PtrArr(*[P(x) if x is not None else None for x in flows])
            lib.baoCudaBLF_C2F.argtypes = [C.c_void_p] * 11 + [C.c_int]
            lib.baoCudaBLF_C2F(fp, PtrArr(*[P(img[0][k][0]) for k in range(nl)]), PtrArr(*[P(img[1][k][0]) for k in range(nl)]),
                               PtrArr(*[P(cen[0][k][0]) for k in range(nl)]), PtrArr(*[P(cen[1][k][0]) for k in range(nl)]), None, None,
                               IntArr(*[d[0] for d in dims]), IntArr(*[d[1] for d in dims]), SzArr(*[img[0][k][1] for k in range(nl)]),
                               SzArr(*[cen[0][k][1] for k in range(nl)]), l)
        r = both(f"BLF_C2F L{l} (upsample+refine)", c2f, [fine, cur])
        r2 = both(f"FlowSmoothing L{l}", lambda lib, t: lib.baoCudaFlowSmoothing(P(t[0]), P(img[0][l][0]), wl, hl, img[0][l][1], wl * 8), [r[0]])
        cur = r2[0]
    both("FlowSmoothing final", lambda lib, t: lib.baoCudaFlowSmoothing(P(t[0]), P(img[0][0][0]), dims[0][1], dims[0][0], img[0][0][1], dims[0][1] * 8), [cur])
    ref.destroy(rc)


run(480, 640, "vga")
if len(sys.argv) > 1:
    run(436, 1024, "sintel", 1)
json.dump(res, open(os.path.join(OUT, "parity_stages.json"), "w"), indent=1)
