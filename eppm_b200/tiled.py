"""Spatial tiling of ONE large frame pair across the GPUs of a box (BASELINE config 4) -- the Python REFERENCE of the
schedule.  The product path is in the library (eppm_compute_tiled_device / _host, csrc/tiled.cu: the same schedule with NCCL calls enqueued
by the library); this module keeps the schedule testable on CPU with gloo (tests/test_cpu.py) and serves as its executable description.

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Every rank holds the full frame pair and builds the full
pyramids itself (prepare is ~0.2 ms; targets of the NNF are unbounded, so the images cannot be banded), but owns only a band of
rows: a multiple of the propagation segment length at the coarsest level, so column segments never straddle two bands.

Data that crosses the fabric, per pair:
  * per PatchMatch iteration, ONE boundary row of both NNF planes before each of the two column passes (row y0-1 from the band
    above before the forward pass, row y1 from the band below before the reverse pass); row passes and the random search are
    band-local.  That is the only exchange the propagation needs: a column segment reads the last pixel of the previous segment
    once, at the start of the pass.
  * after PatchMatch, an all-gather of both NNF / cost fields (the left-right check follows arbitrary targets); the consistency
    stage then runs replicated on every rank (it is tiny at the coarsest level).
  * after every coarse-to-fine step, an all-gather of the rows each rank wrote (smoothing needs a 10-row halo, the next level's
    upsample a 1-row halo; gathering whole planes keeps every rank's planes complete).
Kernels are deterministic and band-agnostic in their arithmetic, so the result is bit-identical to the single-GPU run.

The communication helpers below work on CPU tensors with the gloo backend as well (tests/test_cpu.py)."""
import ctypes as C

import numpy as np

from . import api

PLANE_FLOW_TMP = 9


def band_partition(h_coarse, seg_len, world):
    """Row bands [(y0, y1)] of the coarsest level: whole propagation segments, sizes differing by at most one segment
    (mirrors eppm_set_band in csrc/context.cu)."""
    n_seg = (h_coarse + seg_len - 1) // seg_len
    if world > 1 and n_seg // world < 2:
        raise ValueError("fewer than two segments per band")
    base, extra = divmod(n_seg, world)
    out = []
    for b in range(world):
        s0 = b * base + min(b, extra)
        s1 = s0 + base + (1 if b < extra else 0)
        out.append((s0 * seg_len, min(s1 * seg_len, h_coarse)))
    return out


def level_rows(band, h_coarse, h_level, shift):
    """Rows of a finer level (2**shift times the coarsest) that belong to a coarsest-level band; the last band takes the remainder."""
    y0, y1 = band
    return y0 << shift, (h_level if y1 >= h_coarse else min(h_level, y1 << shift))


def exchange_boundary_rows(planes, bands, rank, world, direction):
    """Halo exchange before a column pass.  planes: list of [h, w, ...] tensors (full-size on every rank).
    direction +1 (forward pass): send my last row y1-1 down to rank+1, receive row y0-1 from rank-1.
    direction -1 (reverse pass): send my first row y0 up to rank-1, receive row y1 from rank+1."""
    import torch.distributed as dist
    if world == 1:
        return
    y0, y1 = bands[rank]
    ops = []
    for t in planes:
        if direction > 0:
            if rank + 1 < world:
                ops.append(dist.P2POp(dist.isend, t[y1 - 1], rank + 1))
            if rank > 0:
                ops.append(dist.P2POp(dist.irecv, t[y0 - 1], rank - 1))
        else:
            if rank > 0:
                ops.append(dist.P2POp(dist.isend, t[y0], rank - 1))
            if rank + 1 < world:
                ops.append(dist.P2POp(dist.irecv, t[y1], rank + 1))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def allgather_bands(plane, row_bands, world):
    """Every rank ends up with all rows of `plane` ([h, ...], full-size everywhere): band r is broadcast from rank r."""
    import torch.distributed as dist
    if world == 1:
        return
    for r, (y0, y1) in enumerate(row_bands):
        if y1 > y0:
            dist.broadcast(plane[y0:y1], src=r)


class _DevMem:
    """Minimal __cuda_array_interface__ carrier so torch can wrap a raw device address without copying."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}


def device_tensor(ptr, shape, dtype):
    import torch
    nbytes = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    return torch.as_tensor(_DevMem(ptr, nbytes), device="cuda").view(dtype).view(*shape)


def compute_flow_tiled(ctx, d_img1, d_img2, rank, world):
    """One frame pair, spatially tiled over `world` ranks.  ctx: EppmContext(h, w, max_batch=1) on this rank's GPU;
    d_img1/d_img2: device uint8 [1,h,w,3] (same content on every rank).  Returns a device float32 [h,w,2] flow, complete on every rank."""
    import torch
    lib, c = ctx.lib, ctx._ctx
    L = ctx.num_levels - 1
    hc, wc = ctx.level_dims(L)
    stream = torch.cuda.ExternalStream(lib.eppm_stream(c))

    def chk(rc, what):
        if rc != 0:
            raise api.EppmError(f"{what} failed ({rc}): {lib.eppm_last_error().decode()}")

    chk(lib.eppm_set_band(c, rank, world), "eppm_set_band")
    try:
        seg = int(ctx.params.prop_seg_length)   # the schedule follows the context's parameters (a band is whole propagation segments)
        bands = band_partition(hc, seg, world)
        y0 = C.c_int(); y1 = C.c_int()
        chk(lib.eppm_band_rows(c, L, C.byref(y0), C.byref(y1)), "eppm_band_rows")
        assert (y0.value, y1.value) == bands[rank], ((y0.value, y1.value), bands[rank])
        nnf = [device_tensor(lib.eppm_device_plane(c, api.PLANE_NNF_FWD + d, L), (hc, wc), torch.int32) for d in range(2)]  # short2 viewed as one int32 per pixel (NCCL has no int16)
        cost = [device_tensor(lib.eppm_device_plane(c, api.PLANE_COST_FWD + d, L), (hc, wc), torch.float32) for d in range(2)]
        with torch.cuda.stream(stream):
            ctx.stage_prepare(d_img1, d_img2, 1)
            n_iter = int(ctx.params.num_iter)
            chk(lib.eppm_tiled_pm_steps(c, 0, 1), "pm init")
            for it in range(n_iter):
                s = 1 + 5 * it
                chk(lib.eppm_tiled_pm_steps(c, s, s + 1), "row forward")
                exchange_boundary_rows(nnf, bands, rank, world, +1)
                chk(lib.eppm_tiled_pm_steps(c, s + 1, s + 3), "column forward, row reverse")
                exchange_boundary_rows(nnf, bands, rank, world, -1)
                chk(lib.eppm_tiled_pm_steps(c, s + 3, s + 5), "column reverse, random search")
            for t in nnf + cost:
                allgather_bands(t, bands, world)
            # consistency on the full coarsest field, replicated (band = whole level for this stage)
            chk(lib.eppm_set_band(c, 0, 1), "eppm_set_band")
            ctx.stage_consistency()
            chk(lib.eppm_set_band(c, rank, world), "eppm_set_band")
            h0, w0 = ctx.level_dims(0)
            tmp_ptr = lib.eppm_device_plane(c, PLANE_FLOW_TMP, 0)
            for level in range(L - 1, -1, -1):
                hl, wl = ctx.level_dims(level)
                rows = [level_rows(b, hc, hl, L - level) for b in bands]
                tmp = device_tensor(tmp_ptr, (hl, wl, 2), torch.float32)
                flow_l = device_tensor(lib.eppm_device_plane(c, api.PLANE_FLOW, level), (hl, wl, 2), torch.float32)
                chk(lib.eppm_tiled_c2f_step(c, level, 0), "refine")
                allgather_bands(tmp, rows, world)
                chk(lib.eppm_tiled_c2f_step(c, level, 1), "smooth")
                allgather_bands(flow_l, rows, world)
            rows0 = [level_rows(b, hc, h0, L) for b in bands]
            out = device_tensor(tmp_ptr, (h0, w0, 2), torch.float32)
            chk(lib.eppm_tiled_c2f_step(c, 0, 2), "final smooth")
            allgather_bands(out, rows0, world)
            result = out.clone()
        stream.synchronize()
        return result
    finally:
        lib.eppm_set_band(c, 0, 1)
