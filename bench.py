#!/usr/bin/env python
"""bench.py — headline benchmark of the EPPM dense-correspondence hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this implementation (libeppm_b200.so)
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference build (oracle/_ref), same metric
    python bench.py --config 2|3|4|5 ...                      # BASELINE.json configs[1..4]; 3 (= configs[2]) is the headline and the default

Headline workload (BASELINE.json configs[2], `--config 3`): synthetic 1920x1080 textured pairs with large-displacement ground
truth, 256 pairs per step.  On N GPUs the 256 pairs are sharded 256/N per GPU as BASELINE.json states it (`scaling: "strong"`;
`--scaling weak` keeps 256 pairs PER GPU instead); ranks own disjoint shards, there is no collective on the data path.  One step = one
pass of the whole path (prepare -> PatchMatch both directions -> consistency -> coarse-to-fine refine + smoothing) over the
shard.  `value` = pairs/s with the shard already resident in HBM; `e2e` = the same through eppm_compute_batch_host with pinned HOST
buffers (H2D of both frames and D2H of the flow inside the timed region).  Inputs (2 x 6.2 MB per pair, >= 398 MB per rank at N = 8)
exceed the 126 MB L2, so no L2 flush is needed between steps.

Other configs: 2 = 64 pairs of 1024x436 (same code path, smaller frames); 4 = ONE 3840x2160 pair spatially tiled over the N GPUs
(row bands + halo exchange inside the library, csrc/tiled.cu; N = 1 runs it untiled); 5 = a 300-frame 1080p video stream through
eppm_compute_stream_* at the default sweep point (every frame prepared once; tools/stream_sweep.py runs the whole sweep).

The reference has no CPU path (README.md:19 of linchaobao/EPPM): `--impl reference` times its own CUDA build on the same GPU(s)
through its public class API (set_data + compute_flow per pair, sequentially -- it cannot batch), one reference process per GPU on
disjoint shards, on a bounded sample of the same pairs.  `cpu_baseline` is the single-threaded CPU oracle (oracle/golden.cpp, kind
"port") on a bounded crop, reported for context only.  An untimed checker leg runs the reference on the first pairs of the
batch and reports the end-point-error delta between the two implementations (BASELINE.json's "mean EPE delta vs reference").
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

CONFIGS = {
    2: dict(h=436, w=1024, total=64, metric="frame_pairs_per_s_1024x436", name="configs[1]: synthetic 1024x436 large-displacement pairs, batch 64"),
    3: dict(h=1080, w=1920, total=256, metric="frame_pairs_per_s_1080p", name="configs[2]: synthetic 1920x1080 large-displacement pairs, batch 256"),
    4: dict(h=2160, w=3840, total=1, metric="frame_pairs_per_s_2160p_tiled", name="configs[3]: one synthetic 3840x2160 pair, row bands over the GPUs"),
    5: dict(h=1080, w=1920, total=299, metric="stream_frame_pairs_per_s_1080p", name="configs[4]: 300-frame synthetic 1080p stream, default sweep point"),
}


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


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def dist_setup():
    import torch
    rank, world, local = env_rank()
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


def sum_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def peaks():
    """(HBM GB/s, source, max SM MHz) from the driver-written MEASURED_PEAKS.json, else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)", d.get("sm_max_mhz", 1965.0)
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def alu_peak(sm_mhz):
    """FP32 lane-ops/s at the SM clock sampled under load.  Measured on the pool's B200 with tools/peak_alu.cu (profiles/r02_alu_peaks.json):
    3.79 scalar-FFMA warp-instructions per clock per SM (the nominal figure is 4 = 128 lanes)."""
    p = os.path.join(ROOT, "profiles", "r02_alu_peaks.json")
    if os.path.exists(p):
        d = json.load(open(p))
        per_clk = d["ffma"]["warp_instr_per_clk_per_sm"]
        return d.get("sms", 148) * per_clk * 32 * sm_mhz * 1e6, f"measured FFMA issue rate {per_clk} warp-instr/clk/SM (profiles/r02_alu_peaks.json) x sampled SM clock"
    return 148 * 128 * sm_mhz * 1e6, "nominal 128 FP32 lanes per SM x sampled SM clock"


def algorithmic_counts(h, w, levels=3, num_iter=10, guesses=6):
    """patch samples S, canonical FP32 lane-ops and algorithmic HBM bytes per pair."""
    n = [int(h * 0.5 ** i) * int(w * 0.5 ** i) for i in range(levels)]
    S = 2 * n[-1] * (1 + num_iter * (4 + guesses)) * 100 + sum(n[:-1]) * 9 * 4 * 100
    T = 441 * (sum(n[1:-1]) + 2 * n[0]) + 2 * (25 * n[0] + 49 * sum(n[1:])) + 169 * n[-1]
    return {"samples": S, "fp32_ops": 30 * S + 8 * T,
            # refine kernel at level 0: reads both packed planes (16 B/px each) + coarse flow, writes flow
            "refine_l0_bytes": n[0] * (16 + 16 + 8) + n[1] * 8, "refine_l0_samples": n[0] * 3600}


def _gen_pair(job):
    from eppm_b200 import synth
    h, w, idx = job
    return synth.make_pair(h, w, idx)


def make_inputs(h, w, distinct, first_idx, workers=None):
    """`distinct` synthetic pairs (indices first_idx ...), generated in parallel on the host cores BEFORE CUDA is initialised
    (a pair takes ~10 s of numpy/scipy at 1080p).  Returns (a, b, gt, valid) stacked."""
    t0 = time.time()
    jobs = [(h, w, first_idx + i) for i in range(distinct)]
    nw = workers or max(1, min(distinct, (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    if nw > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(nw) as pool:
            res = pool.map(_gen_pair, jobs)
    else:
        res = [_gen_pair(j) for j in jobs]
    log(f"[rank {env_rank()[0]}] generated {distinct} synthetic {w}x{h} pairs with {nw} workers in {time.time() - t0:.1f}s")
    return tuple(np.stack([r[k] for r in res]) for k in range(4))


def shard_plan(args, cfg, world):
    """Pairs per GPU per step and the scaling label."""
    total = args.batch if args.batch else cfg["total"]
    if world == 1 or args.scaling == "weak":
        return total, "weak"
    return max(1, total // world), "strong"


def reference_quality(a, b, gt, va, h, w, n_q):
    """Untimed checker leg: the reference build's flows of the first n_q pairs (oracle/_ref through its public class API)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refharness
    if not refharness.available():
        return None
    ref = refharness.Ref()
    rc = ref.create(h, w)
    flows = []
    for i in range(n_q):
        ref.set_data(rc, a[i], b[i])
        flows.append(ref.compute_flow(rc, h, w))
    ref.destroy(rc)
    return np.stack(flows)


def quality_keys(mine, ref_flows, gt, va):
    from eppm_b200 import synth
    n_q = len(mine)
    out = {"epe_vs_gt_px": round(float(np.mean([synth.epe(mine[i], gt[i], va[i]) for i in range(n_q)])), 4), "quality_pairs": n_q}
    if ref_flows is not None:
        e_ref = float(np.mean([synth.epe(ref_flows[i], gt[i], va[i]) for i in range(n_q)]))
        d = np.sqrt(((mine.astype(np.float64) - ref_flows.astype(np.float64)) ** 2).sum(-1))
        ok = np.isfinite(d) & (np.abs(mine).max(-1) < 1e9) & (np.abs(ref_flows).max(-1) < 1e9)
        out.update({"epe_vs_gt_px_reference": round(e_ref, 4), "epe_delta_vs_reference": round(out["epe_vs_gt_px"] - e_ref, 4),
                    "flow_vs_reference_mean_epe_px": round(float(d[ok].mean()), 4), "flow_vs_reference_median_epe_px": round(float(np.median(d[ok])), 6),
                    "flow_vs_reference_frac_gt_0p5px": round(float((d[ok] > 0.5).mean()), 5)})
    return out


def run_b200(args):
    cfg = CONFIGS[args.config]
    H, W = cfg["h"], cfg["w"]
    rank, world, local = env_rank()
    if args.config == 4:
        return run_tiled(args, cfg)
    if args.config == 5:
        return run_stream(args, cfg)
    batch, scaling = shard_plan(args, cfg, world)
    d = min(args.distinct, batch)
    a, b, gt, va = make_inputs(H, W, d, 1000 * rank)
    import torch
    import eppm_b200 as E
    dist_setup()
    os.environ["EPPM_PROFILE"] = "1"
    chunk = min(args.chunk, batch)
    ctx = E.EppmContext(H, W, chunk, device=local)
    # pinned host batch (cycled distinct pairs) and device-resident copy
    idx = [i % d for i in range(batch)]
    h_a = torch.empty((batch, H, W, 3), dtype=torch.uint8).pin_memory()
    h_b = torch.empty((batch, H, W, 3), dtype=torch.uint8).pin_memory()
    for i, j in enumerate(idx):
        h_a[i] = torch.from_numpy(a[j]); h_b[i] = torch.from_numpy(b[j])
    h_flow = torch.empty((batch, H, W, 2), dtype=torch.float32).pin_memory()
    d_a = h_a.cuda(); d_b = h_b.cuda()
    d_flow = torch.empty((chunk, H, W, 2), dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.lib.eppm_stream(ctx._ctx))

    def step_resident():
        for s in range(0, batch, chunk):
            n = min(chunk, batch - s)
            ctx.compute_batch_device(d_a[s:s + n], d_b[s:s + n], n, d_flow)

    def step_host():
        # ONE call for the whole shard: the library chunks it and overlaps H2D / D2H with compute
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

    with torch.cuda.stream(stream):
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
    e2e_steps = max(1, args.steps // 2) if args.steps > 1 else 1
    ms_e2e = timed(step_host, e2e_steps)

    if rank == 0:
        hbm_peak, peak_src, sm_max = peaks()
        cnt = algorithmic_counts(H, W)
        pairs = batch * world
        value = pairs * args.steps / (ms / 1e3)
        e2e_value = pairs * e2e_steps / (ms_e2e / 1e3)
        n_last = min(chunk, batch - (batch - 1) // chunk * chunk)  # pairs in the last chunk = what last_kernel_ms timed
        k_bytes = cnt["refine_l0_bytes"] * n_last
        achieved = k_bytes / (k_ms / 1e3) / 1e9
        sm_mhz = clocks["sm_mhz"] or sm_max
        apeak, apeak_src = alu_peak(sm_mhz)
        t_pair = ms / 1e3 / (batch * args.steps)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r02_refine_l0_traffic.json")   # dram bytes per pair of the level-0 refine launch, from an ncu --set full capture
        if os.path.exists(tp) and (H, W) == (1080, 1920):
            traffic = int(round(json.load(open(tp))["dram_bytes_per_pair"] * n_last))
        line = {
            "metric": cfg["metric"], "value": round(value, 3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}, default params (defs.h); {batch} pairs per GPU per step",
                       "pairs_per_step_total": pairs, "batch_per_gpu": batch, "chunk": chunk, "distinct_pairs_per_gpu": d,
                       "l2_policy": f"inputs {2 * batch * H * W * 3 / 1e6:.0f} MB per rank > 126 MB L2, no flush", "rng": "xorwow (reference stream)"},
            "mpix_per_s": round(value * H * W / 1e6, 2),
            "e2e": {"value": round(e2e_value, 3), "unit": "pairs/s", "h2d_bytes_per_step": int(batch * H * W * 3 * 2),
                    "d2h_bytes_per_step": int(batch * H * W * 2 * 4), "api": "eppm_compute_batch_host (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_c2f_refine_row (level 0)", "achieved": round(achieved, 2), "peak": hbm_peak, "unit": "GB/s",
                         "frac": round(achieved / hbm_peak, 5), "traffic": traffic, "peak_source": peak_src,
                         "note": "the path is FP32-issue bound, not HBM bound (SURVEY.md §8d); see roofline_alu"},
            "roofline_alu": {"bound": "fp32 issue", "kernel_samples_per_s": round(cnt["refine_l0_samples"] * n_last / (k_ms / 1e3), 1),
                             "kernel_ms": round(k_ms, 3), "kernel_pairs": n_last,
                             "canonical_ops_per_sample": 30,
                             "achieved_tlaneops": round(30 * cnt["refine_l0_samples"] * n_last / (k_ms / 1e3) / 1e12, 3),
                             "peak_tlaneops": round(apeak / 1e12, 3), "peak_source": apeak_src,
                             "frac": round(30 * cnt["refine_l0_samples"] * n_last / (k_ms / 1e3) / apeak, 4),
                             "whole_pair_frac": round(cnt["fp32_ops"] / t_pair / apeak, 4), "sm_mhz_used": sm_mhz},
            "stage_ms_last_chunk": {k: round(v, 3) for k, v in stage.items()},
        }
        # quality of the last batch against ground truth and against the reference build's flows (untimed checker leg)
        n_q = min(args.quality_pairs, d, batch)
        mine = h_flow[:n_q].numpy().copy()
        ref_flows = None if args.no_reference_check else reference_quality(a, b, gt, va, H, W, n_q)
        line.update(quality_keys(mine, ref_flows, gt, va))
        line["cpu_baseline"] = cpu_baseline(args, H, W)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_stream(args, cfg):
    """BASELINE config 5 at its default point: a 300-frame 1080p stream (a clip of chained synthetic frames cycled), every frame
    prepared once (eppm_compute_stream_device / eppm_compute_stream_host).  N > 1: contiguous frame ranges per rank (shard.stream_shard)."""
    H, W = cfg["h"], cfg["w"]
    rank, world, local = env_rank()
    from eppm_b200 import synth, shard
    n_frames = 300
    clip_len = min(args.distinct + 1, 9)
    clip, flows, valids = synth.make_stream(H, W, clip_len, first_idx=500)
    import torch
    import eppm_b200 as E
    dist_setup()
    f0, f1, _, _ = shard.stream_shard(n_frames, rank, world)   # frames [f0, f1): consecutive pairs inside are this rank's
    my_frames = f1 - f0
    # ping-pong through the clip so consecutive frames are always a generated (frame, next frame) pair or its reverse
    period = 2 * (clip_len - 1)
    def clip_index(t):
        m = t % period
        return m if m < clip_len else period - m
    h_fr = torch.empty((my_frames, H, W, 3), dtype=torch.uint8).pin_memory()
    for i in range(my_frames):
        h_fr[i] = torch.from_numpy(clip[clip_index(f0 + i)])
    d_fr = h_fr.cuda()
    chunk = min(args.chunk, my_frames)
    ctx = E.EppmContext(H, W, chunk, device=local)
    P = chunk - 1
    d_flow = torch.empty((P, H, W, 2), dtype=torch.float32, device="cuda")
    h_flow = torch.empty((my_frames - 1, H, W, 2), dtype=torch.float32).pin_memory()
    stream = torch.cuda.ExternalStream(ctx.lib.eppm_stream(ctx._ctx))

    def step_resident():
        for s in range(0, my_frames - 1, P):
            m = min(P, my_frames - 1 - s)
            ctx.compute_stream_device(d_fr[s:s + m + 1], m, d_flow)

    def step_host():
        ctx.compute_stream_host(h_fr, out=h_flow)

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

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step_resident()
    ctx.synchronize()
    ctx.launch_count(reset=True)
    sampler = ClockSampler(local); sampler.start()
    ms = timed(step_resident, args.steps)
    launches = ctx.launch_count()
    clocks = sampler.stop()
    step_host()
    e2e_steps = max(1, args.steps // 2) if args.steps > 1 else 1
    ms_e2e = timed(step_host, e2e_steps)
    pairs_total = sum_over_ranks(my_frames - 1, world)
    if rank == 0:
        value = pairs_total * args.steps / (ms / 1e3)
        e2e_value = pairs_total * e2e_steps / (ms_e2e / 1e3)
        # pair t -> t+1 of the first pass through the clip is a generated pair with ground truth
        n_q = min(clip_len - 1, my_frames - 1, args.quality_pairs)
        from eppm_b200 import synth as S
        fl = h_flow[:n_q].numpy()
        line = {
            "metric": cfg["metric"], "value": round(value, 3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['name']} (pyramid depth 3, 10 iterations, patch stride 2); clip of {clip_len} chained frames played back and forth",
                       "frames": n_frames, "pairs_per_step_total": int(pairs_total), "chunk_frames": chunk,
                       "l2_policy": f"frames {my_frames * H * W * 3 / 1e6:.0f} MB per rank > 126 MB L2, no flush"},
            "mpix_per_s": round(value * H * W / 1e6, 2),
            "e2e": {"value": round(e2e_value, 3), "unit": "pairs/s", "h2d_bytes_per_step": int(my_frames * H * W * 3),
                    "d2h_bytes_per_step": int((my_frames - 1) * H * W * 2 * 4), "api": "eppm_compute_stream_host (C ABI, pinned host buffers)"},
            "gpu_launches": int(launches), "clocks": clocks,
            "epe_vs_gt_px": round(float(np.mean([S.epe(fl[i], flows[i], valids[i]) for i in range(n_q)])), 4), "quality_pairs": n_q,
            "roofline": None, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_tiled(args, cfg):
    """BASELINE config 4: ONE 3840x2160 pair, row bands of the coarsest level over the N GPUs with a halo exchange per column pass
    (eppm_compute_tiled_*, csrc/tiled.cu; bit-identical to the untiled run).  A step = the whole path for that pair; e2e adds the H2D of the pair on every
    rank (every rank builds the full pyramids) and the D2H of the flow on rank 0."""
    H, W = cfg["h"], cfg["w"]
    rank, world, local = env_rank()
    a, b, gt, va = make_inputs(H, W, 1, 4000, workers=1)
    import torch
    import eppm_b200 as E
    from eppm_b200 import tiled
    dist_setup()
    ctx = E.EppmContext(H, W, 1, device=local)
    if world > 1:
        ctx.tiled_init(rank, world)
    h_a = torch.from_numpy(a).pin_memory(); h_b = torch.from_numpy(b).pin_memory()
    h_flow = torch.empty((1, H, W, 2), dtype=torch.float32).pin_memory()
    d_a = h_a.cuda(); d_b = h_b.cuda()
    d_flow = torch.empty((1, H, W, 2), dtype=torch.float32, device="cuda")
    stream = torch.cuda.ExternalStream(ctx.lib.eppm_stream(ctx._ctx))

    def step_resident():
        if world == 1:
            ctx.compute_batch_device(d_a, d_b, 1, d_flow)
        else:
            ctx.compute_tiled_device(d_a, d_b, d_flow)   # the library drives the halo exchange and the band gathers (csrc/tiled.cu)

    def step_host():
        if world == 1:
            ctx.compute_batch_host(h_a, h_b, out=h_flow)
        else:
            ctx.compute_tiled_host(h_a, h_b, out=h_flow if rank == 0 else None, want_flow=rank == 0)

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

    steps = max(args.steps, 5)
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step_resident()
    ctx.synchronize()
    ctx.launch_count(reset=True)
    sampler = ClockSampler(local); sampler.start()
    ms = timed(step_resident, steps)
    launches = ctx.launch_count()
    clocks = sampler.stop()
    step_host()
    ms_e2e = timed(step_host, steps)
    if rank == 0:
        from eppm_b200 import synth as S
        value = steps / (ms / 1e3)
        line = {
            "metric": cfg["metric"], "value": round(value, 3), "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": round(ms / steps, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}, default params (defs.h)", "bands": world,
                       "l2_policy": "packed planes of one 4K pair = 362 MB > 126 MB L2, no flush"},
            "mpix_per_s": round(value * H * W / 1e6, 2),
            "e2e": {"value": round(steps / (ms_e2e / 1e3), 3), "unit": "pairs/s", "h2d_bytes_per_step": int(H * W * 3 * 2),
                    "d2h_bytes_per_step": int(H * W * 2 * 4), "api": "eppm_compute_tiled_host (C ABI, pinned host buffers; NCCL enqueued by the library)" if world > 1 else "eppm_compute_batch_host"},
            "gpu_launches": int(launches), "clocks": clocks,
            "epe_vs_gt_px": round(float(S.epe(h_flow[0].numpy(), gt[0], va[0])), 4),
            "roofline": None, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args, H, W):
    """Single-threaded CPU oracle on a bounded crop of the same synthetic workload (about 10-30 s), scaled to the config's frame size."""
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
    """The reference's own CUDA build (oracle/_ref/libeppm_ref.so: unmodified sources + texture/malloc shims), public class API,
    one process per GPU on its own shard of the pairs (sequential pairs: the reference cannot batch)."""
    cfg = CONFIGS[args.config]
    H, W = cfg["h"], cfg["w"]
    rank, world, local = env_rank()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refharness
    if not refharness.available():
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libeppm_ref.so not built (needs /root/reference at build time)"}))
        return
    if args.config == 4 and world > 1:
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "the reference cannot tile one frame over several GPUs; run --config 4 --gpus 1"}))
        return
    batch, scaling = shard_plan(args, cfg, world)
    sample = min(args.ref_sample, batch)
    d = min(args.distinct, sample) if args.config != 5 else min(args.distinct + 1, 9) - 1
    if args.config == 5:
        from eppm_b200 import synth
        clip, gt, va = synth.make_stream(H, W, d + 1, first_idx=500)
        a, b = clip[:-1], clip[1:]
    else:
        a, b, gt, va = make_inputs(H, W, d, (4000 if args.config == 4 else 1000 * rank), workers=None if d > 1 else 1)
    import torch
    dist_setup()
    ref = refharness.Ref()
    rc = ref.create(H, W)
    keep = {}

    def step():
        ms = 0.0
        for i in range(sample):
            t, fl = ref.time_pair(rc, a[i % d], b[i % d], H, W)
            if i < d and i < args.quality_pairs:
                keep[i] = fl
            ms += t
        return ms

    for _ in range(args.warmup):
        step()
    barrier(world)
    sampler = ClockSampler(local); sampler.start()
    tot = 0.0
    for _ in range(args.steps):
        tot += step()
    clocks = sampler.stop()
    binds = ref.shim_stats() if hasattr(ref, "shim_stats") else None
    tot_max = max_over_ranks(tot, world)
    if rank == 0:
        from eppm_b200 import synth as S
        value = world * sample * args.steps / (tot_max / 1e3)
        n_q = len(keep)
        line = {
            "impl": "reference", "metric": cfg["metric"], "value": round(value, 3), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(tot_max / args.steps, 3), "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}, default params (defs.h); {batch} pairs per GPU per step", "pairs_per_step_total": batch * world,
                       "batch_per_gpu": batch, "sample_pairs_per_step_per_gpu": sample, "distinct_pairs_per_gpu": d},
            "mpix_per_s": round(value * H * W / 1e6, 2),
            "e2e": {"value": round(value, 3), "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "clocks": clocks,
            "epe_vs_gt_px": round(float(np.mean([S.epe(keep[i], gt[i], va[i]) for i in range(n_q)])), 4) if n_q else None, "quality_pairs": n_q,
            "cpu_baseline": {"value": round(value, 3), "unit": "pairs/s", "cores": 0, "kind": "reference",
                             "sample": f"{sample} of the {batch} pairs per GPU per step through bao_flow_patchmatch_multiscale_cuda::set_data+compute_flow "
                                       "of the reference's own CUDA build on the same GPU(s) (the reference has no CPU path); cudaEvent time around the two "
                                       "class calls only, sequential pairs; includes its dead weighted-median pass, debug D2H and per-call cudaMalloc (BASELINE.md §2)"},
            "texture_shim": {"binds_total": binds[0], "texture_objects_created": binds[1],
                             "mode": "uncached (4 blocking calls per bind)" if os.environ.get("REF_SHIM_NOCACHE") else "cached objects, async handle upload"} if binds else None,
        }
        print(json.dumps(line), flush=True)
    ref.destroy(rc)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json config number (1-based): 3 = the headline 1080p batch")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the config's pairs sharded over the GPUs (BASELINE.json), weak = the config's pairs PER GPU")
    ap.add_argument("--batch", type=int, default=0, help="override the config's pairs per step (total under strong scaling, per GPU under weak)")
    ap.add_argument("--chunk", type=int, default=32,
                    help="pairs per kernel launch (context max_batch); 32: the propagation's 800 dependent launches per chunk are latency bound "
                         "(measured at 1080p: 16 -> 61.9, 32 -> 64.8, 64 -> 65.0 pairs/s; 6 GB of planes per 32 pairs)")
    ap.add_argument("--distinct", type=int, default=32, help="distinct synthetic pairs generated per GPU and cycled through its shard")
    ap.add_argument("--quality-pairs", type=int, default=4, help="pairs whose flow is compared with ground truth and with the reference build (untimed)")
    ap.add_argument("--ref-sample", type=int, default=8, help="pairs per GPU per step for --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-check", action="store_true", help="skip the untimed leg that runs the reference build on the first pairs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
