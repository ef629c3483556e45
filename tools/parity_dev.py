"""Development parity probe (run under gpurun): compares libeppm_b200 with the reference build stage by stage and prints
mismatch statistics instead of asserting.  The pytest versions live in tests/test_parity_gpu.py."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import eppm_b200 as E
from eppm_b200 import synth
from refharness import Ref, pitched

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)
res = {}


def stats(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype.kind == "f":
        ne = a.view(np.uint32) != b.view(np.uint32)
    else:
        ne = a != b
    frac = float(ne.mean())
    out = {"mismatch_frac": frac, "n": int(ne.size)}
    if a.dtype.kind == "f" and frac > 0:
        fin = np.isfinite(a) & np.isfinite(b)
        out["max_abs"] = float(np.abs(a[fin] - b[fin]).max()) if fin.any() else None
    print(f"{name:40s} mismatch {frac:.6f} {out.get('max_abs', '')}")
    res[name] = out
    return frac


def run_size(h, w, tag, imgs=None):
    if imgs is None:
        a, b, gt, valid = synth.make_pair(h, w, 0)
    else:
        a, b = imgs
        gt = valid = None
    ref = Ref()
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    ctx = E.EppmContext(h, w, 1)
    d1 = torch.from_numpy(a[None]).cuda(); d2 = torch.from_numpy(b[None]).cuda()
    ctx.stage_prepare(d1, d2, 1)
    nl = ctx.num_levels
    planes = {}
    for l in range(nl):
        for which, nm in ((0, "rgba1"), (1, "rgba2"), (2, "census1"), (3, "census2")):
            r = ref.read_plane(rc, which, l)
            m = ctx.read_plane(which, l)
            planes[(nm, l)] = r
            stats(f"{tag} prepare {nm} L{l}", m, r)
    # ---- PatchMatch taps: reference fed with ITS OWN level planes
    L = nl - 1
    hc, wc = ctx.level_dims(L)
    i1 = pitched(planes[("rgba1", L)]); i2 = pitched(planes[("rgba2", L)])
    c1 = pitched(planes[("census1", L)]); c2 = pitched(planes[("census2", L)])
    for n_steps in (1, 2, 3, 4, 5, 6, 7, 12, 52):
        nf, cf = ref.tap_patchmatch(i1, i2, c1, c2, wc, hc, n_steps)
        nb, cb = ref.tap_patchmatch(i2, i1, c2, c1, wc, hc, n_steps)
        ctx.stage_prepare(d1, d2, 1)
        ctx.stage_patchmatch_partial(n_steps)
        f1 = stats(f"{tag} pm steps={n_steps} nnf fwd", ctx.read_plane(E.PLANE_NNF_FWD), nf)
        stats(f"{tag} pm steps={n_steps} nnf bwd", ctx.read_plane(E.PLANE_NNF_BWD), nb)
        if n_steps >= 2:
            stats(f"{tag} pm steps={n_steps} cost fwd", ctx.read_plane(E.PLANE_COST_FWD), cf)
            stats(f"{tag} pm steps={n_steps} cost bwd", ctx.read_plane(E.PLANE_COST_BWD), cb)
    # ---- full pipeline taps
    t0 = time.time()
    flow_ref = ref.compute_flow(rc, h, w)
    t_ref = time.time() - t0
    ctx.stage_prepare(d1, d2, 1)
    ctx.stage_patchmatch()
    ref_nnf_pm_f, ref_cost_f = ref.tap_patchmatch(i1, i2, c1, c2, wc, hc, 1000)
    ref_nnf_pm_b, ref_cost_b = ref.tap_patchmatch(i2, i1, c2, c1, wc, hc, 1000)
    stats(f"{tag} pm final nnf fwd", ctx.read_plane(E.PLANE_NNF_FWD), ref_nnf_pm_f)
    stats(f"{tag} pm final nnf bwd", ctx.read_plane(E.PLANE_NNF_BWD), ref_nnf_pm_b)
    # inject the reference's PM result so later stages are compared on identical inputs
    ctx.write_plane(E.PLANE_NNF_FWD, ref_nnf_pm_f); ctx.write_plane(E.PLANE_NNF_BWD, ref_nnf_pm_b)
    ctx.write_plane(E.PLANE_COST_FWD, ref_cost_f); ctx.write_plane(E.PLANE_COST_BWD, ref_cost_b)
    ctx.stage_consistency()
    stats(f"{tag} consistency nnf fwd", ctx.read_plane(E.PLANE_NNF_FWD), ref.read_plane(rc, 4, L))
    stats(f"{tag} consistency nnf bwd", ctx.read_plane(E.PLANE_NNF_BWD), ref.read_plane(rc, 5, L))
    stats(f"{tag} flow L{L}", ctx.read_plane(E.PLANE_FLOW, L), ref.read_plane(rc, 8, L))
    # c2f from the reference's coarse flow
    ctx.write_plane(E.PLANE_FLOW, ref.read_plane(rc, 8, L), level=L)
    ctx.stage_c2f(None)
    for l in range(L - 1, -1, -1):
        m = ctx.read_plane(E.PLANE_FLOW, l); r = ref.read_plane(rc, 8, l)
        stats(f"{tag} c2f flow L{l}", m, r)
        d = np.sqrt(((m - r) ** 2).sum(-1))
        res[f"{tag} c2f flow L{l} epe"] = float(d.mean())
        print(f"   mean EPE vs ref at L{l}: {d.mean():.6f}  frac>0.01: {(d > 0.01).mean():.5f}")
    # end to end through the public host API
    t0 = time.time()
    flow = ctx.compute_batch_host(a[None], b[None])[0]
    t_me = time.time() - t0
    d = np.sqrt(((flow - flow_ref) ** 2).sum(-1))
    print(f"{tag} END-TO-END mean EPE vs ref {d.mean():.6f}, frac differing {(d > 0).mean():.5f}, ref {t_ref*1e3:.1f} ms, mine {t_me*1e3:.1f} ms")
    res[f"{tag} e2e"] = {"epe_vs_ref": float(d.mean()), "frac_diff": float((d > 0).mean()), "ref_ms": t_ref * 1e3, "mine_ms": t_me * 1e3}
    if gt is not None:
        res[f"{tag} e2e"]["epe_gt_ref"] = synth.epe(flow_ref, gt, valid)
        res[f"{tag} e2e"]["epe_gt_mine"] = synth.epe(flow, gt, valid)
        print(f"   EPE vs GT (valid px): ref {res[f'{tag} e2e']['epe_gt_ref']:.4f} mine {res[f'{tag} e2e']['epe_gt_mine']:.4f}")
    ref.destroy(rc)
    ctx.close()


sizes = [(480, 640, "vga")]
if len(sys.argv) > 1 and sys.argv[1] == "all":
    sizes += [(436, 1024, "sintel"), (1080, 1920, "fhd")]
for h, w, tag in sizes:
    run_size(h, w, tag)
fr = os.path.join(ROOT, "oracle", "_ref", "data")
if os.path.exists(os.path.join(fr, "frame10.ppm")):
    a = synth.read_ppm(os.path.join(fr, "frame10.ppm")); b = synth.read_ppm(os.path.join(fr, "frame11.ppm"))
    run_size(a.shape[0], a.shape[1], "frame10", (a, b))
json.dump(res, open(os.path.join(OUT, "parity_dev.json"), "w"), indent=1)
