"""Development probe (run under gpurun): drives oracle/_ref/libeppm_ref.so on the B200 and records
 - texture unorm8->float table and point-sampling texel selection at fractional coordinates,
 - XORWOW known answers through the reference's own random-field kernel,
 - reference wall/device time per pair at 640x480 and 1920x1080,
 - reference run-to-run noise floor (it has data races).
Writes gpurun_out/probe_ref.json.  Test infrastructure only."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libeppm_ref.so"))
ref.ref_create.restype = C.c_void_p
ref.ref_create.argtypes = [C.c_int, C.c_int]
ref.ref_destroy.argtypes = [C.c_void_p]
ref.ref_time_pair.restype = C.c_float
ref.ref_time_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
ref.ref_read_plane.restype = C.c_long
ref.ref_read_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
ref.ref_probe_texture.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]

res = {}

# --- texture probes -------------------------------------------------------------------------
unorm = np.zeros(256, np.float32)
xs = []
for base in (1, 7, 100, 200):
    for k in range(-64, 65):
        xs.append(np.float32(base) + np.float32(k) * np.float32(2.0 ** -12))
    # immediately adjacent floats around the integer
    b = np.float32(base)
    xs += [np.nextafter(b, np.float32(0)), b, np.nextafter(b, np.float32(1e9))]
xs += [np.float32(-0.5), np.float32(-1e-6), np.float32(0.0), np.float32(255.0), np.float32(255.9), np.float32(256.0), np.float32(300.0)]
xs = np.array(xs, np.float32)
tex = np.zeros(len(xs), np.uint8)
rc = ref.ref_probe_texture(unorm.ctypes.data, xs.ctypes.data, len(xs), tex.ctypes.data)
exp_unorm = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
res["tex_probe_rc"] = rc
res["unorm_equals_k_div_255"] = bool(np.array_equal(unorm.view(np.uint32), exp_unorm.view(np.uint32)))
res["unorm_bits"] = unorm.view(np.uint32).tolist()
floor_sel = np.clip(np.floor(xs), 0, 255).astype(np.uint8)
mism = np.nonzero(floor_sel != tex)[0]
res["point_sample_is_floor"] = bool(len(mism) == 0)
res["point_sample_mismatches"] = [(float(xs[i]), int(tex[i]), int(floor_sel[i])) for i in mism[:200]]
print("unorm==k/255:", res["unorm_equals_k_div_255"], "point==floor:", res["point_sample_is_floor"], "n_mism", len(mism))


def textured_pair(h, w, seed, shift=(7, -4)):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, size=(h // 8 + 6, w // 8 + 6, 3)).astype(np.float32)
    img = np.kron(base, np.ones((8, 8, 1), np.float32))[: h + 32, : w + 32]
    fine = rng.integers(-40, 41, size=(h + 32, w + 32, 3)).astype(np.float32)
    img = np.clip(img * 0.7 + 40 + fine, 0, 255).astype(np.uint8)
    a = img[16 : 16 + h, 16 : 16 + w]
    b = img[16 + shift[1] : 16 + shift[1] + h, 16 + shift[0] : 16 + shift[0] + w]
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


# --- XORWOW KAT through the reference's random field kernel ------------------------------------
import torch  # device buffers only

ref.ref_tap_patchmatch.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_size_t] * 4 + [C.c_int]
for (w, h, exp00, exp10) in ((160, 120, (80, 54), (151, 119)), (480, 270, (177, 214), (178, 112))):
    nnf = torch.zeros((h, w, 2), dtype=torch.int16, device="cuda")
    cost = torch.zeros((h, w), dtype=torch.float32, device="cuda")
    pitch4 = ((w * 4 + 511) // 512) * 512
    pitch1 = ((w + 511) // 512) * 512
    img = torch.zeros((h, pitch4), dtype=torch.uint8, device="cuda")
    cen = torch.zeros((h, pitch1), dtype=torch.uint8, device="cuda")
    rc = ref.ref_tap_patchmatch(nnf.data_ptr(), cost.data_ptr(), img.data_ptr(), img.data_ptr(), cen.data_ptr(),
                                cen.data_ptr(), w, h, pitch4, w * 4, w * 4, pitch1, 1)
    n = nnf.cpu().numpy()
    res[f"xorwow_{w}x{h}"] = {"rc": rc, "p00": n[0, 0].tolist(), "p10": n[0, 1].tolist(), "expect00": exp00, "expect10": exp10,
                              "max_x": int(n[..., 0].max()), "max_y": int(n[..., 1].max())}
    print("xorwow", w, h, n[0, 0], n[0, 1], exp00, exp10)

# --- timing + noise floor -------------------------------------------------------------------------
for (h, w, reps) in ((480, 640, 4), (436, 1024, 3), (1080, 1920, 3)):
    a, b = textured_pair(h, w, 1)
    ctx = ref.ref_create(h, w)
    flows = []
    times = []
    for r in range(reps):
        fl = np.zeros((h, w, 2), np.float32)
        t0 = time.time()
        ms = ref.ref_time_pair(ctx, a.ctypes.data, b.ctypes.data, fl.ctypes.data)
        times.append((float(ms), (time.time() - t0) * 1e3))
        flows.append(fl)
    ref.ref_destroy(ctx)
    d01 = np.sqrt(((flows[0] - flows[1]) ** 2).sum(-1))
    d12 = np.sqrt(((flows[1] - flows[2]) ** 2).sum(-1))
    gt = np.array([-7.0, 4.0], np.float32)
    epe = np.sqrt(((flows[-1] - gt) ** 2).sum(-1))
    res[f"ref_{w}x{h}"] = {
        "ms_event_wall": times,
        "noise_mean_epe_run0_vs_run1": float(d01.mean()), "noise_frac_diff_run0_vs_run1": float((d01 > 0).mean()),
        "noise_mean_epe_run1_vs_run2": float(d12.mean()), "noise_frac_diff_run1_vs_run2": float((d12 > 0).mean()),
        "epe_vs_shift_gt_mean": float(epe.mean()), "epe_vs_shift_gt_median": float(np.median(epe)),
    }
    print(w, h, res[f"ref_{w}x{h}"])

json.dump(res, open(os.path.join(OUT, "probe_ref.json"), "w"), indent=1)
print("probe_ref done")
