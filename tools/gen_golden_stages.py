"""Generates tests/golden/refstage_*.npz from the REFERENCE's own CUDA build (oracle/_ref) on a B200: inputs and outputs of the stage
functions its host class declares but compute_flow never calls -- baoCudaLeftRightCheck_Buffered, baoCudaFlow2NNF, baoCudaFlowCutoff,
baoEliminateStillRegionFlow, baoCudaImageSmoothing, baoCudaFlowBilteralUpsampling.  Run under gpurun, then copy
gpurun_out/golden/refstage_*.npz into tests/golden/.  They pin the CPU oracle (oracle/golden_stages.cpp) and, on a GPU box without
oracle/_ref, libeppm_b200 itself."""
import ctypes as C
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
L = ref.lib
S, I, V, F = C.c_size_t, C.c_int, C.c_void_p, C.c_float
L.baoCudaLeftRightCheck_Buffered.argtypes = [V] * 6 + [I, I, S, S]
L.baoCudaFlow2NNF.argtypes = [V, V, I, I, S, S]
L.baoCudaFlowCutoff.argtypes = [V, I, I, S, F]
L.baoEliminateStillRegionFlow.argtypes = [V, V, V, I, I, S]
L.baoCudaImageSmoothing.argtypes = [V, V, I, I, S]
L.baoCudaFlowBilteralUpsampling.argtypes = [V, V, I, I, S, V, I, I, F]
L.baoCudaPatchMatch_PlaneFitting.argtypes = [V] * 6 + [I, I, S, S, S, S]
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
P = lambda t: t.data_ptr()

for name, h, w, idx, scale in [("s128x96", 96, 128, 7, 0.12)]:
    a, b, _, _ = synth.make_pair(h, w, idx, scale_to=scale)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    ref.compute_flow(rc, h, w)
    out = {"h": h, "w": w}
    rg = [[ref.read_plane(rc, k, l) for l in range(3)] for k in (0, 1)]
    ce = [[ref.read_plane(rc, 2 + k, l) for l in range(3)] for k in (0, 1)]
    (h1, w1), (h2, w2) = ref.level_dims(rc, 1), ref.level_dims(rc, 2)
    out["rgba1_L1"], out["rgba2_L1"], out["rgba1_L2"] = rg[0][1], rg[1][1], rg[0][2]
    # PatchMatch of both directions at level 1 (64x48) feeds the buffered left-right check
    i1, i2, c1, c2 = pitched(rg[0][1]), pitched(rg[1][1]), pitched(ce[0][1]), pitched(ce[1][1])
    nf, cf = ref.tap_patchmatch(i1, i2, c1, c2, w1, h1, 1000)
    nb, cb = ref.tap_patchmatch(i2, i1, c2, c1, w1, h1, 1000)
    out.update(lr_nnf1=nf, lr_cost1=cf, lr_nnf2=nb, lr_cost2=cb)
    t = [dev(nf), dev(cf), dev(nb), dev(cb)]
    tn, tc = torch.zeros_like(t[0]), torch.zeros_like(t[1])
    L.baoCudaLeftRightCheck_Buffered(P(t[0]), P(t[1]), P(t[2]), P(t[3]), P(tn), P(tc), w1, h1, w1 * 4, w1 * 4)
    torch.cuda.synchronize()
    out.update(lr_out_nnf1=t[0].cpu().numpy(), lr_out_cost1=t[1].cpu().numpy(), lr_out_nnf2=t[2].cpu().numpy(), lr_out_cost2=t[3].cpu().numpy())
    # flow -> NNF and cut-off on a flow with fractional, negative, unknown, out-of-short-range and non-finite entries
    rng = np.random.default_rng(11)
    fl = (rng.random((h1, w1, 2), dtype=np.float32) * 200 - 100).astype(np.float32)
    fl[3:9, 5:40] = 1e10
    fl[10, :] = 40000.0
    fl[11, :] = -40000.0
    fl[12, 0] = np.nan; fl[12, 1] = np.inf; fl[12, 2] = -np.inf
    out["f2n_flow"] = fl
    o = torch.zeros((h1, w1, 2), dtype=torch.int16, device="cuda"); d = dev(fl)
    L.baoCudaFlow2NNF(P(o), P(d), w1, h1, w1 * 4, w1 * 8); torch.cuda.synchronize()
    out["f2n_nnf"] = o.cpu().numpy()
    d = dev(fl); L.baoCudaFlowCutoff(P(d), w1, h1, w1 * 8, 37.5); torch.cuda.synchronize()
    out["cutoff_out"] = d.cpu().numpy()
    # still-region elimination: image 2 := image 1 on the left half
    bmix = rg[1][1].copy(); bmix[:, : w1 // 2] = rg[0][1][:, : w1 // 2]
    out["still_img2"] = bmix
    ia, ib = pitched(rg[0][1]), pitched(bmix)
    d = torch.full((h1, w1, 2), 3.25, dtype=torch.float32, device="cuda")
    L.baoEliminateStillRegionFlow(P(d), P(ia[0]), P(ib[0]), w1, h1, ia[1]); torch.cuda.synchronize()
    out["still_out"] = d.cpu().numpy()
    # image smoothing (level 1) -- alpha is left uninitialised by the reference, only r, g, b are kept
    o = torch.zeros_like(ia[0])
    L.baoCudaImageSmoothing(P(o), P(ia[0]), w1, h1, ia[1]); torch.cuda.synchronize()
    out["smooth_out"] = o.cpu().numpy()[:, : w1 * 4].reshape(h1, w1, 4)[..., :3].copy()
    # joint-bilateral upsampling of a level-2 flow to level 1
    small = (rng.standard_normal((h2, w2, 2)) * 4).astype(np.float32)
    small[2:9, 3:20] = 1e10
    out["up_small"] = small
    o = torch.full((h1, w1, 2), -7.0, dtype=torch.float32, device="cuda"); d = dev(small)
    L.baoCudaFlowBilteralUpsampling(P(o), P(ia[0]), w1, h1, ia[1], P(d), w2, h2, 2.0); torch.cuda.synchronize()
    out["up_out"] = o.cpu().numpy()
    # the whole PatchMatch scored with the plane-fitting cost, forward direction, coarsest level (the oracle regenerates the planes
    # from the seeded pair: they are pinned bit-exact by ref_*.npz)
    out.update(pair_idx=idx, scale_to=scale)
    j1, j2, k1, k2 = pitched(rg[0][2]), pitched(rg[1][2]), pitched(ce[0][2]), pitched(ce[1][2])
    pn = torch.zeros((h2, w2, 2), dtype=torch.int16, device="cuda"); pc = torch.zeros((h2, w2), dtype=torch.float32, device="cuda")
    L.baoCudaPatchMatch_PlaneFitting(P(pn), P(pc), P(j1[0]), P(j2[0]), P(k1[0]), P(k2[0]), w2, h2, j1[1], w2 * 4, w2 * 4, k1[1]); torch.cuda.synchronize()
    out.update(pf_nnf=pn.cpu().numpy(), pf_cost=pc.cpu().numpy())
    np.savez_compressed(os.path.join(OUT, f"refstage_{name}.npz"), **out)
    print(name, "saved", {k: v.shape for k, v in out.items() if hasattr(v, "shape")})
    ref.destroy(rc)
