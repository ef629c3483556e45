"""Spatially tiled run of ONE frame pair over all ranks (BASELINE config 4).  Launch with torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/tiled_run.py H W [reps]
Rank 0 also computes the pair untiled on its own GPU and checks bit-identity, then prints one JSON line with timings."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eppm_b200 as E
from eppm_b200 import synth, tiled

h, w = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# identical inputs on every rank (seeded generator); a cheap large-frame pair: tiled texture + shift
a, b, gt, valid = synth.make_pair(min(h, 540), min(w, 960), 42)
ry, rx = -(-h // a.shape[0]), -(-w // a.shape[1])
a = np.ascontiguousarray(np.tile(a, (ry, rx, 1))[:h, :w]); b = np.ascontiguousarray(np.tile(b, (ry, rx, 1))[:h, :w])
d1, d2 = torch.from_numpy(a[None]).cuda(), torch.from_numpy(b[None]).cuda()
ctx = E.EppmContext(h, w, 1, device=local)
times = []
for r in range(reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    flow = tiled.compute_flow_tiled(ctx, d1, d2, rank, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times.append((time.time() - t0) * 1e3)
out = {"h": h, "w": w, "world": world, "tiled_ms": times}
if rank == 0:
    single = torch.empty((1, h, w, 2), dtype=torch.float32, device="cuda")
    ts = []
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.time()
        ctx.compute_batch_device(d1, d2, 1, single); ctx.synchronize()
        ts.append((time.time() - t0) * 1e3)
    same = torch.equal(single[0].view(torch.int32), flow.view(torch.int32))
    out.update({"single_gpu_ms": ts, "bit_identical_to_single_gpu": bool(same),
                "max_abs_diff": float((single[0] - flow).abs().max().item())})
    print(json.dumps(out), flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
