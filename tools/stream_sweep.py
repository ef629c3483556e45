"""BASELINE config 5: a 1080p synthetic video stream through eppm_compute_stream_device (every frame prepared once and shared by its two
pairs), swept over pyramid depth x PatchMatch iterations x patch sample stride.

    python tools/stream_sweep.py [n_frames=300] [n_frames_other=120]

The clip is 8 chained synthetic frames (eppm_b200.synth.make_stream: frame t+1 = frame t moved by a fresh large-displacement motion, with
ground truth) cycled to the stream length; the default point (depth 3, 10 iterations, stride 2) runs the full stream, the other 26 points a
shorter one.  Per point: pairs/s from CUDA events around the whole stream (frames resident in HBM) and the mean end-point error against
ground truth over the chained pairs.  Where oracle/_ref holds a reference build of the same point (default, d2i5, d4i2, s3, s1) its
pairs/s on a few pairs of the same clip is recorded beside it.  Writes gpurun_out/stream_sweep.json.  Not the headline benchmark: bench.py is."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import eppm_b200 as E
from eppm_b200 import synth
import refharness

H, W = int(os.environ.get("SS_H", 1080)), int(os.environ.get("SS_W", 1920))
n_full = int(sys.argv[1]) if len(sys.argv) > 1 else 300
n_other = int(sys.argv[2]) if len(sys.argv) > 2 else 120
CLIP = 8
CHUNK = 16   # pairs per eppm_compute_stream_device call (17 frames resident per call)

t0 = time.time()
frames, gt, valid = synth.make_stream(H, W, CLIP, first_idx=200)
print(f"clip of {CLIP} frames generated in {time.time() - t0:.1f}s", flush=True)
d_clip = torch.from_numpy(frames).cuda()


def stream_indices(n_frames):
    return [t % CLIP for t in range(n_frames)]


def run_point(depth, iters, stride, n_frames):
    p = E.default_params()
    p.pyr_levels, p.num_iter, p.patch_stride = depth, iters, stride
    ctx = E.EppmContext(H, W, CHUNK + 1, params=p)
    idx = stream_indices(n_frames)
    d_frames = d_clip[torch.tensor(idx, device="cuda")]                       # [n_frames,h,w,3] resident
    d_flow = torch.empty((CHUNK, H, W, 2), dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.lib.eppm_stream(ctx._ctx))
    epe_sum, epe_n = 0.0, 0

    def one_pass(measure_epe):
        nonlocal epe_sum, epe_n
        for s in range(0, n_frames - 1, CHUNK):
            n = min(CHUNK, n_frames - 1 - s)
            ctx.compute_stream_device(d_frames[s:s + n + 1], n, d_flow)
            if measure_epe and s == 0:
                ctx.synchronize()
                fl = d_flow[:n].cpu().numpy()
                for k in range(n):
                    a = idx[s + k]
                    if a + 1 < CLIP and idx[s + k + 1] == a + 1:              # a chained pair with ground truth (not the wrap-around)
                        epe_sum += synth.epe(fl[k], gt[a], valid[a]); epe_n += 1

    one_pass(True)                                                            # warm-up + quality
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    one_pass(False)
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1)
    ctx.close()
    return {"pairs_per_s": round((n_frames - 1) / (ms / 1e3), 2), "ms_per_pair": round(ms / (n_frames - 1), 3), "frames": n_frames,
            "epe_vs_gt_px": round(epe_sum / max(1, epe_n), 4), "epe_pairs": epe_n}


def run_reference(lib_path, n_pairs=6):
    if not os.path.exists(lib_path):
        return None
    ref = refharness.Ref(lib_path)
    rc = ref.create(H, W)
    ms, ep = [], []
    for k in range(n_pairs + 1):
        a = k % (CLIP - 1)
        t, fl = ref.time_pair(rc, frames[a], frames[a + 1], H, W)
        if k:                                                                 # first pair = warm-up
            ms.append(t); ep.append(synth.epe(fl, gt[a], valid[a]))
    ref.destroy(rc)
    return {"pairs_per_s": round(1e3 / float(np.mean(ms)), 3), "ms_per_pair": round(float(np.mean(ms)), 2), "pairs": n_pairs,
            "epe_vs_gt_px": round(float(np.mean(ep)), 4)}


REF_POINTS = {(3, 10, 2): refharness.REF_LIB, (2, 5, 2): refharness.variant_lib("d2i5"), (4, 2, 2): refharness.variant_lib("d4i2"),
              (3, 10, 3): refharness.variant_lib("s3"), (3, 10, 1): refharness.variant_lib("s1")}
res = {"workload": f"{W}x{H} stream, clip of {CLIP} chained synthetic frames cycled; {n_full} frames at the default point, {n_other} elsewhere",
       "points": []}
STRIDES = tuple(int(x) for x in os.environ.get("SS_STRIDES", "1,2,3").split(","))   # SS_STRIDES=1,3 re-measures two columns only
for depth in (2, 3, 4):
    for iters in (2, 5, 10):
        for stride in STRIDES:
            key = (depth, iters, stride)
            r = run_point(depth, iters, stride, n_full if key == (3, 10, 2) else n_other)
            row = {"pyr_depth": depth, "num_iter": iters, "patch_stride": stride, **r}
            if key in REF_POINTS and not os.environ.get("SS_NO_REF"):
                row["reference"] = run_reference(REF_POINTS[key])
            res["points"].append(row)
            print(row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out_path = os.path.join(ROOT, "gpurun_out", "stream_sweep.json")
if len(STRIDES) < 3 and os.path.exists(os.path.join(ROOT, "profiles", "r01_stream_sweep.json")):   # merge into the committed table
    old = json.load(open(os.path.join(ROOT, "profiles", "r01_stream_sweep.json")))
    new = {(p["pyr_depth"], p["num_iter"], p["patch_stride"]): p for p in res["points"]}
    merged = []
    for p in old["points"]:
        q = new.get((p["pyr_depth"], p["num_iter"], p["patch_stride"]), p)
        if "reference" in p and "reference" not in q:
            q["reference"] = p["reference"]   # the reference build did not change
        merged.append(q)
    res["points"] = merged
json.dump(res, open(out_path, "w"), indent=1)
