"""Generates tests/golden/refscaled_*.npz from the REFERENCE's own CUDA build (oracle/_ref) on a B200: baoCudaPatchMatch_Scaled (declared by
its host class, never called; bao_pmflow_kernel.cu:1828-1895) on the coarsest level of a seeded synthetic pair -- targets, scales and the cost
plane.  Run under gpurun, then copy gpurun_out/golden/refscaled_*.npz into tests/golden/.  Pins the CPU oracle (oracle/golden.cpp,
patchmatch_scaled) and, on a GPU box without oracle/_ref, libeppm_b200 itself.  The input planes are regenerated from the seeded pair by
whoever checks against the fixture (they are pinned bit-exact by ref_*.npz)."""
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
S, I, V = C.c_size_t, C.c_int, C.c_void_p
L.baoCudaPatchMatch_Scaled.argtypes = [V] * 7 + [I, I, S, S, S, S, S]
L.baoCudaPatchMatch_Scaled.restype = None
P = lambda t: t.data_ptr()

for name, h, w, idx, scale in [("s128x96", 96, 128, 7, 0.12), ("s192x160", 160, 192, 9, 0.12)]:
    a, b, _, _ = synth.make_pair(h, w, idx, scale_to=scale)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    ref.compute_flow(rc, h, w)
    rg = [ref.read_plane(rc, k, 2) for k in (0, 1)]
    ce = [ref.read_plane(rc, 2 + k, 2) for k in (0, 1)]
    h2, w2 = ref.level_dims(rc, 2)
    j1, j2, k1, k2 = pitched(rg[0]), pitched(rg[1]), pitched(ce[0]), pitched(ce[1])
    outs = []
    for rep in range(2):   # twice: the function must be deterministic for the fixture to mean anything
        nn = torch.zeros((h2, w2, 2), dtype=torch.int16, device="cuda")
        sc = torch.zeros((h2, w2), dtype=torch.float32, device="cuda"); co = torch.zeros((h2, w2), dtype=torch.float32, device="cuda")
        L.baoCudaPatchMatch_Scaled(P(nn), P(sc), P(co), P(j1[0]), P(j2[0]), P(k1[0]), P(k2[0]), w2, h2, j1[1], w2 * 4, w2 * 4, w2 * 4, k1[1])
        torch.cuda.synchronize()
        outs.append((nn.cpu().numpy(), sc.cpu().numpy(), co.cpu().numpy()))
    assert all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(*outs)), "reference not deterministic"
    np.savez_compressed(os.path.join(OUT, f"refscaled_{name}.npz"), h=h, w=w, pair_idx=idx, scale_to=scale, rgba1_L2=rg[0], rgba2_L2=rg[1],
                        sc_nnf=outs[0][0], sc_scale=outs[0][1], sc_cost=outs[0][2])
    print(name, "saved", outs[0][0].shape, "scales", np.unique(outs[0][1]))
    ref.destroy(rc)
