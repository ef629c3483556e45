"""Generates tests/golden/refsub_*.npz from the REFERENCE's own CUDA build (oracle/_ref) on a B200: inputs and outputs of
baoCudaCensusTransform_Bicubic + baoCudaSubpixRefine on one seeded synthetic pair.  Run under gpurun, then copy
gpurun_out/golden/refsub_*.npz into tests/golden/.  The reference library is loaded in this process only for this purpose (its
sub-pixel entry point leaves the image texture reference in linear-filter mode)."""
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
S, I, V = C.c_size_t, C.c_int, C.c_void_p
ref.lib.baoCudaCensusTransform_Bicubic.argtypes = [V, V, I, I, S, V, V, I, I, S]
ref.lib.baoCudaSubpixRefine.argtypes = [V] * 6 + [I, I, S, S, S, S]

for name, h, w, idx, scale in [("s128x96", 96, 128, 7, 0.12)]:
    a, b, _, _ = synth.make_pair(h, w, idx, scale_to=scale)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    ref.compute_flow(rc, h, w)
    rgba1, rgba2 = ref.read_plane(rc, 0, 0), ref.read_plane(rc, 1, 0)          # level-0 planes (pre-blurred) and their census
    cen1, cen2 = ref.read_plane(rc, 2, 0), ref.read_plane(rc, 3, 0)
    i1, i2, c1, c2 = pitched(rgba1), pitched(rgba2), pitched(cen1), pitched(cen2)
    nnf, _ = ref.tap_patchmatch(i1, i2, c1, c2, w, h, 1000)                    # integer targets: the reference's PatchMatch on level 0
    wu, hu = 2 * w, 2 * h
    cp = (wu + 511) // 512 * 512
    u1 = torch.zeros((hu, cp), dtype=torch.uint8, device="cuda"); u2 = torch.zeros_like(u1)
    ref.lib.baoCudaCensusTransform_Bicubic(u1.data_ptr(), u2.data_ptr(), wu, hu, cp, i1[0].data_ptr(), i2[0].data_ptr(), w, h, i1[1])
    dn = torch.from_numpy(nnf).cuda()
    fl = (dn.float() - torch.stack(torch.meshgrid(torch.arange(w), torch.arange(h), indexing="xy"), -1).float().cuda()).contiguous()
    flow_in = fl.cpu().numpy()
    ref.lib.baoCudaSubpixRefine(fl.data_ptr(), dn.data_ptr(), i1[0].data_ptr(), i2[0].data_ptr(), u1.data_ptr(), u2.data_ptr(), w, h, i1[1], cp, w * 4, w * 8)
    torch.cuda.synchronize()
    out = {"h": h, "w": w, "rgba1": rgba1, "rgba2": rgba2, "nnf": nnf, "census1_up": u1[:, :wu].cpu().numpy(), "census2_up": u2[:, :wu].cpu().numpy(),
           "flow_in": flow_in, "flow_out": fl.cpu().numpy()}
    np.savez_compressed(os.path.join(OUT, f"refsub_{name}.npz"), **out)
    print(name, "saved; refined fraction", float((out["flow_out"] != flow_in).any(-1).mean()))
    ref.destroy(rc)
