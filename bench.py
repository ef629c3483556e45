#!/usr/bin/env python
"""bench.py — headline benchmark of the EPPM dense-correspondence hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this implementation (libeppm_b200.so)
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference build (oracle/_ref), same metric

Workload (BASELINE.json configs[2]): synthetic 1920x1080 textured pairs with large-displacement ground truth, a batch
of 256 pairs per GPU per step (weak scaling: every rank owns its own batch, no data-path collective).  One step = one
pass of the whole path (prepare -> PatchMatch both directions -> consistency -> coarse-to-fine refine + smoothing)
over that batch.  `value` = pairs/s with the batch already resident in HBM; `e2e` = the same through
eppm_compute_batch_host with pinned HOST buffers (H2D of both frames and D2H of the flow inside the timed region).
Inputs (2 x 256 x 6.2 MB = 3.2 GB per rank) exceed the 126 MB L2, so no L2 flush is needed between steps.

The reference has no CPU path (README.md:19 of linchaobao/EPPM): `--impl reference` times its own CUDA build on the
same GPU through its public class API (set_data + compute_flow per pair, sequentially -- it cannot batch), on a bounded
sample of the same pairs.  `cpu_baseline` is the single-threaded CPU oracle (oracle/golden.cpp, kind "port") on a
bounded crop, reported for context only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1080, 1920
METRIC = "frame_pairs_per_s_1080p"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)", d.get("sm_max_mhz", 1965.0)
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def algorithmic_counts(h, w, levels=3, num_iter=10, guesses=6):
    """patch samples S, canonical FP32 lane-ops and algorithmic HBM bytes per pair."""
    n = [int(h * 0.5 ** i) * int(w * 0.5 ** i) for i in range(levels)]
    S = 2 * n[-1] * (1 + num_iter * (4 + guesses)) * 100 + sum(n[:-1]) * 9 * 4 * 100
    T = 441 * (sum(n[1:-1]) + 2 * n[0]) + 2 * (25 * n[0] + 49 * sum(n[1:])) + 169 * n[-1]
    return {"samples": S, "fp32_ops": 30 * S + 8 * T,
            # refine kernel at level 0: reads both packed planes (16 B/px each) + coarse flow, writes flow
            "refine_l0_bytes": n[0] * (16 + 16 + 8) + n[1] * 8, "refine_l0_samples": n[0] * 3600}


def make_inputs(n_pairs, distinct, rank):
    from eppm_b200 import synth
    cache = os.path.join(ROOT, "build", f"bench_pairs_{H}x{W}_{distinct}_{rank}.npz")
    if os.path.exists(cache):
        z = np.load(cache)
        base = (z["a"], z["b"], z["gt"], z["valid"])
    else:
        t0 = time.time()
        a, b, gt, va = synth.make_batch(H, W, distinct, first_idx=1000 * rank)
        base = (a, b, gt, va)
        try:
            os.makedirs(os.path.dirname(cache), exist_ok=True)
            np.savez(cache, a=a, b=b, gt=gt, valid=va)
        except Exception:
            pass
        log(f"[rank {rank}] generated {distinct} synthetic {W}x{H} pairs in {time.time() - t0:.1f}s")
    return base


def run_b200(args):
    import torch
    import eppm_b200 as E
    rank, world, local = dist_setup(args.gpus)
    os.environ["EPPM_PROFILE"] = "1"
    a, b, gt, va = make_inputs(args.batch, args.distinct, rank)
    d = a.shape[0]
    chunk = min(args.chunk, args.batch)
    if args.stream_priorities:
        os.environ["EPPM_STREAM_PRIORITY"] = str(-5)
    ctx = E.EppmContext(H, W, chunk, device=local)
    # optional second context on its own stream: alternate chunks so that one chunk's PatchMatch (L1-bound) can overlap the
    # other's refine (issue-bound) on the same SMs
    ctxs = [ctx]
    for k in range(1, args.streams):
        if args.stream_priorities:   # later contexts less urgent: their kernels fill the SMs the first context's leave idle
            os.environ["EPPM_STREAM_PRIORITY"] = str(0)
        ctxs.append(E.EppmContext(H, W, chunk, device=local))
    os.environ.pop("EPPM_STREAM_PRIORITY", None)
    # pinned host batch (cycled distinct pairs) and device-resident copy
    idx = [i % d for i in range(args.batch)]
    h_a = torch.empty((args.batch, H, W, 3), dtype=torch.uint8).pin_memory()
    h_b = torch.empty((args.batch, H, W, 3), dtype=torch.uint8).pin_memory()
    for i, j in enumerate(idx):
        h_a[i] = torch.from_numpy(a[j]); h_b[i] = torch.from_numpy(b[j])
    h_flow = torch.empty((args.batch, H, W, 2), dtype=torch.float32).pin_memory()
    d_a = h_a.cuda(); d_b = h_b.cuda()
    d_flows = [torch.empty((chunk, H, W, 2), dtype=torch.float32, device="cuda") for _ in ctxs]
    stream = torch.cuda.ExternalStream(ctx.lib.eppm_stream(ctx._ctx))
    streams = [torch.cuda.ExternalStream(c.lib.eppm_stream(c._ctx)) for c in ctxs]

    def step_resident():
        for k, s in enumerate(range(0, args.batch, chunk)):
            n = min(chunk, args.batch - s)
            ctxs[k % len(ctxs)].compute_batch_device(d_a[s:s + n], d_b[s:s + n], n, d_flows[k % len(ctxs)])
        for k in range(1, len(ctxs)):  # join the extra streams into the timed one
            ev = torch.cuda.Event(); ev.record(streams[k]); stream.wait_event(ev)

    def step_host():
        # ONE call for the whole batch: the library chunks it and overlaps H2D / D2H with compute
        ctx.compute_batch_host(h_a, h_b, out=h_flow)

    def timed(fn, steps):
        barrier(world)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        ctx.synchronize()
        barrier(world)
        return max_over_ranks(e0.elapsed_time(e1), world)

    for _ in range(args.warmup):
        step_resident()
    ctx.synchronize()
    ctx.launch_count(reset=True)
    sampler = ClockSampler(local); sampler.start()
    ms = timed(step_resident, args.steps)
    launches = ctx.launch_count()
    stage = ctx.last_stage_ms()
    k_ms = ctx.last_kernel_ms(0)
    clocks = sampler.stop()
    # end to end through the host-buffer API
    step_host()
    ms_e2e = timed(step_host, max(1, args.steps // 2) if args.steps > 1 else 1)
    e2e_steps = max(1, args.steps // 2) if args.steps > 1 else 1
    # quality of the last batch against ground truth (not timed)
    from eppm_b200 import synth
    fl = h_flow[:d].numpy()
    epe_gt = float(np.mean([synth.epe(fl[i], gt[idx[i]], va[idx[i]]) for i in range(min(d, args.batch))]))

    if rank == 0:
        hbm_peak, peak_src, sm_max = peaks()
        cnt = algorithmic_counts(H, W)
        pairs = args.batch * world
        value = pairs * args.steps / (ms / 1e3)
        e2e_value = pairs * e2e_steps / (ms_e2e / 1e3)
        n_last = min(chunk, args.batch - (args.batch - 1) // chunk * chunk)  # pairs in the last chunk = what last_kernel_ms timed
        k_bytes = cnt["refine_l0_bytes"] * n_last
        achieved = k_bytes / (k_ms / 1e3) / 1e9
        sm_mhz = clocks["sm_mhz"] or sm_max
        alu_peak = 148 * 128 * sm_mhz * 1e6  # FP32 lane-ops/s at the clock measured under load
        t_pair = ms / 1e3 / (args.batch * args.steps)
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"configs[2]: synthetic {W}x{H} large-displacement pairs, batch {args.batch} per GPU, default params (defs.h)",
                       "batch_per_gpu": args.batch, "chunk": chunk, "distinct_pairs": d, "l2_policy": "inputs 3.2 GB per rank >> 126 MB L2, no flush",
                       "rng": "xorwow (reference stream)"},
            "mpix_per_s": round(value * H * W / 1e6, 2),
            "e2e": {"value": round(e2e_value, 3), "unit": "pairs/s", "h2d_bytes_per_step": int(args.batch * H * W * 3 * 2),
                    "d2h_bytes_per_step": int(args.batch * H * W * 2 * 4), "api": "eppm_compute_batch_host (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_c2f_refine_tab (level 0)", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(achieved / hbm_peak, 5),
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from `ncu --set full` (profiles/r01_ncu_refine_tab_l0_summary.txt:
                         # 288.36 + 59.46 MB for a 4-pair launch), scaled to the pairs of the timed launch; algorithmic = 87.1 MB / pair
                         "traffic": int(round((288.360192e6 + 59.462144e6) / 4 * n_last)) if (H, W) == (1080, 1920) else None,
                         "peak_source": peak_src,
                         "note": "the path is FP32-issue bound, not HBM bound (SURVEY.md §8d); see roofline_alu"},
            "roofline_alu": {"bound": "fp32 issue", "kernel_samples_per_s": round(cnt["refine_l0_samples"] * n_last / (k_ms / 1e3), 1),
                             "kernel_ms": round(k_ms, 3), "kernel_pairs": n_last,
                             "canonical_ops_per_sample": 30,
                             "achieved_tlaneops": round(30 * cnt["refine_l0_samples"] * n_last / (k_ms / 1e3) / 1e12, 3),
                             "peak_tlaneops": round(alu_peak / 1e12, 3), "frac": round(30 * cnt["refine_l0_samples"] * n_last / (k_ms / 1e3) / alu_peak, 4),
                             "whole_pair_frac": round(cnt["fp32_ops"] / t_pair / alu_peak, 4), "sm_mhz_used": sm_mhz},
            "stage_ms_last_chunk": {k: round(v, 3) for k, v in stage.items()},
            "epe_vs_gt_px": round(epe_gt, 4),
        }
        line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline(args):
    """Single-threaded CPU oracle on a bounded crop of the same synthetic workload (about 10-30 s), scaled to 1080p pairs/s."""
    if args.no_cpu_baseline:
        return None
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import golden
    from eppm_b200 import synth
    h, w = 192, 320
    a, b, _, _ = synth.make_pair(h, w, 0, scale_to=0.12)
    g = golden.Golden(h, w)
    t0 = time.time()
    g.compute(a, b)
    dt = time.time() - t0
    return {"value": round((h * w / dt) / (H * W), 6), "unit": "pairs/s", "cores": 1, "kind": "port",
            "sample": f"one {w}x{h} synthetic pair through oracle/golden.cpp in {dt:.1f}s, scaled by pixel count to {W}x{H}",
            "host_cores_available": os.cpu_count()}


def run_reference(args):
    """The reference's own CUDA build (oracle/_ref/libeppm_ref.so: unmodified sources + texture/malloc shims), public class API."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refharness
    if not refharness.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libeppm_ref.so not built (needs /root/reference at build time)"}))
        return
    import torch
    torch.cuda.set_device(0)
    a, b, gt, va = make_inputs(args.batch, args.distinct, 0)
    d = a.shape[0]
    sample = args.ref_sample
    ref = refharness.Ref()
    rc = ref.create(H, W)
    pairs_done = 0

    def step():
        nonlocal pairs_done
        ms = 0.0
        for i in range(sample):
            t, _ = ref.time_pair(rc, a[i % d], b[i % d], H, W)
            ms += t
        return ms

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(0); sampler.start()
    tot = 0.0
    for _ in range(args.steps):
        tot += step()
    clocks = sampler.stop()
    value = sample * args.steps / (tot / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(tot / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"configs[2]: synthetic {W}x{H} large-displacement pairs, default params (defs.h)", "batch_per_gpu": args.batch,
                   "sample_pairs_per_step": sample},
        "mpix_per_s": round(value * H * W / 1e6, 2),
        "e2e": {"value": round(value, 3), "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "clocks": clocks,
        "cpu_baseline": {"value": round(value, 3), "unit": "pairs/s", "cores": 0, "kind": "reference",
                         "sample": f"{sample} of the {args.batch} pairs per step through bao_flow_patchmatch_multiscale_cuda::set_data+compute_flow "
                                   "of the reference's own CUDA build on this GPU (the reference has no CPU path); cudaEvent time, sequential pairs; "
                                   "includes its dead weighted-median pass, debug D2H and per-call cudaMalloc (BASELINE.md §2)"},
    }
    print(json.dumps(line), flush=True)
    ref.destroy(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="pairs per GPU per step")
    ap.add_argument("--chunk", type=int, default=16, help="pairs per kernel launch (context max_batch)")
    ap.add_argument("--distinct", type=int, default=4, help="distinct synthetic pairs generated and cycled through the batch")
    ap.add_argument("--ref-sample", type=int, default=8, help="pairs per step for --impl reference")
    ap.add_argument("--streams", type=int, default=1, help="contexts/streams alternating over the chunks of a step")
    ap.add_argument("--stream-priorities", action="store_true", help="with --streams 2: first context urgent, the others at the lowest priority")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
