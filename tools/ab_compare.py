"""Bitwise A/B comparison of two builds of libeppm_b200.so (regression guard while tuning kernels).

    python tools/ab_compare.py build/ab/libeppm_b200_base.so [eppm_b200/libeppm_b200.so]

Each build runs in its own process (EPPM_LIB_PATH) over a fixed set of cases -- synthetic pairs at several sizes, the patch-stride
and depth/iteration variants, the Philox mode -- and writes the flows; the parent compares them bit for bit and writes
gpurun_out/ab_compare.json.  Both builds claim the same bits as the reference, so ANY difference is a bug in one of them."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # (name, h, w, n_pairs, first_idx, params)
    ("1080p", 1080, 1920, 2, 0, {}),
    ("436x1024", 436, 1024, 3, 1, {}),
    ("odd_121x161", 121, 161, 2, 5, {}),
    ("vga_philox", 480, 640, 2, 2, {"rng_mode": 1}),
    ("436x1024_philox", 436, 1024, 1, 1, {"rng_mode": 1}),
    ("stride1", 240, 320, 1, 3, {"patch_stride": 1}),
    ("stride3", 240, 320, 1, 3, {"patch_stride": 3}),
    ("depth2_it5", 240, 320, 1, 4, {"pyr_levels": 2, "num_iter": 5}),
    ("depth4_it2", 480, 640, 1, 4, {"pyr_levels": 4, "num_iter": 2}),
    ("guesses4", 240, 320, 1, 6, {"num_rand_guess": 4}),
]


def child(out_path):
    import numpy as np
    import eppm_b200 as E
    from eppm_b200 import synth
    res = {}
    for name, h, w, n, idx, prm in CASES:
        a, b, _, _ = synth.make_batch(h, w, n, first_idx=idx, distinct=n)
        p = E.default_params()
        for k, v in prm.items():
            setattr(p, k, v)
        ctx = E.EppmContext(h, w, n, params=p)
        flow = ctx.compute_batch_host(a, b)
        res[name] = hashlib.sha256(np.ascontiguousarray(flow).tobytes()).hexdigest()
        np.save(out_path + "." + name + ".npy", flow)
        ctx.close()
    json.dump(res, open(out_path, "w"))


def main():
    libs = sys.argv[1:3]
    if len(libs) < 2:
        libs.append(os.path.join(ROOT, "eppm_b200", "libeppm_b200.so"))
    import numpy as np
    outs = []
    for i, lib in enumerate(libs):
        out = f"/tmp/ab_{i}.json"
        env = dict(os.environ, EPPM_LIB_PATH=os.path.abspath(lib))
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", out], check=True, env=env)
        outs.append(out)
    ha, hb = json.load(open(outs[0])), json.load(open(outs[1]))
    report = {}
    for name, *_ in CASES:
        fa, fb = np.load(outs[0] + "." + name + ".npy"), np.load(outs[1] + "." + name + ".npy")
        diff = fa.view(np.uint32) != fb.view(np.uint32)
        d = np.sqrt(((fa - fb) ** 2).sum(-1))
        report[name] = {"same_bits": ha[name] == hb[name], "floats_differing": int(diff.sum()), "mean_epd": float(d.mean()), "max_epd": float(d.max())}
        print(name, report[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"libs": libs, "cases": report}, open(os.path.join(ROOT, "gpurun_out", "ab_compare.json"), "w"), indent=1)
    sys.exit(0 if all(r["same_bits"] for r in report.values()) else 1)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        main()
