"""One 3840x2160 pair tiled over the GPUs by the LIBRARY (eppm_compute_tiled_device, csrc/tiled.cu) -- run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/tiled_run_c.py [h w]
Checks bit-identity against the untiled single-GPU result (rank 0 computes it on its own GPU) and reports the per-pair time of both, device
events, max over ranks.  Writes gpurun_out/tiled_c_<N>gpu.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import eppm_b200 as E
from eppm_b200 import synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2160, 3840)
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
a, b, gt, va = synth.make_pair(h, w, 4000)
da, db = torch.from_numpy(a[None]).cuda(), torch.from_numpy(b[None]).cuda()
ctx = E.EppmContext(h, w, 1, device=local)
ctx.tiled_init(rank, world)
out = torch.zeros((1, h, w, 2), dtype=torch.float32, device="cuda")
ref = torch.zeros_like(out)
stream = torch.cuda.ExternalStream(ctx.lib.eppm_stream(ctx._ctx))

def timed(fn, reps):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps): fn()
        e1.record(stream)
    ctx.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

for _ in range(3):
    ctx.compute_tiled_device(da, db, out)
ms_tiled = timed(lambda: ctx.compute_tiled_device(da, db, out), 10)
res = {"h": h, "w": w, "gpus": world, "ms_tiled": ms_tiled}
if rank == 0:
    c1 = E.EppmContext(h, w, 1, device=local)
    for _ in range(2):
        c1.compute_batch_device(da, db, 1, ref)
    c1.synchronize()
    s1 = torch.cuda.ExternalStream(c1.lib.eppm_stream(c1._ctx))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s1):
        e0.record(s1)
        for _ in range(5): c1.compute_batch_device(da, db, 1, ref)
        e1.record(s1)
    c1.synchronize()
    res["ms_single_gpu"] = e0.elapsed_time(e1) / 5
    res["speedup"] = res["ms_single_gpu"] / ms_tiled
    res["bit_identical_to_single_gpu"] = bool(torch.equal(out.view(torch.int32), ref.view(torch.int32)))
    res["epe_vs_gt_px"] = synth.epe(out[0].cpu().numpy(), gt, va)
    c1.close()
# every rank must hold the same complete flow
if world > 1:
    chk = out.view(torch.int32).to(torch.int64).sum().reshape(1)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    res["all_ranks_same_checksum"] = bool(all(int(x.item()) == int(lst[0].item()) for x in lst))
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"tiled_c_{world}gpu.json"), "w"), indent=1)
    print(json.dumps(res))
ctx.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
