"""Host-side mirror of the reference's operator interface for the dense-correspondence path.

`EppmContext` is the batched context API (eppm_create / eppm_compute_batch_host / ...), and
`BaoFlowPatchmatchMultiscaleCuda` mirrors the reference's C++ class
(bao_flow_patchmatch_multiscale_cuda.h:33-45: init, set_data, compute_flow) for callers who think in those terms.
Everything executes in libeppm_b200.so on the GPU; nothing here computes."""
import ctypes as C
import sys

import numpy as np

from . import _lib

PLANE_RGBA1, PLANE_RGBA2, PLANE_CENSUS1, PLANE_CENSUS2 = 0, 1, 2, 3
PLANE_NNF_FWD, PLANE_NNF_BWD, PLANE_COST_FWD, PLANE_COST_BWD, PLANE_FLOW, PLANE_FLOW_TMP = 4, 5, 6, 7, 8, 9


class EppmError(RuntimeError):
    pass


def default_params():
    p = _lib.EppmParams()
    _lib.load().eppm_default_params(C.byref(p))
    return p


def _ptr(a):
    """Raw address of a numpy array, a torch tensor (host or device) or an int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _check_array(a, shape, dtype, what, device=None):
    """Raw pointers cross the C ABI: refuse anything that is not the exact dtype / shape / C-contiguous layout the library reads.
    `device`: None = host memory expected, an int = a tensor on that CUDA device (raw int addresses are the caller's responsibility)."""
    if isinstance(a, int):
        return
    np_dt = np.dtype(dtype)
    if isinstance(a, np.ndarray):
        if device is not None:
            raise EppmError(f"{what}: a numpy array is host memory, a device tensor is required")
        if a.dtype != np_dt or tuple(a.shape) != tuple(shape) or not a.flags["C_CONTIGUOUS"]:
            raise EppmError(f"{what}: need C-contiguous {np_dt.name} {tuple(shape)}, got {a.dtype.name} {tuple(a.shape)}"
                            f"{'' if a.flags['C_CONTIGUOUS'] else ' (not contiguous)'}")
        return
    if hasattr(a, "data_ptr"):   # torch tensor
        name = str(a.dtype).replace("torch.", "")
        if name != np_dt.name or tuple(a.shape) != tuple(shape) or not a.is_contiguous():
            raise EppmError(f"{what}: need contiguous {np_dt.name} {tuple(shape)}, got {name} {tuple(a.shape)}"
                            f"{'' if a.is_contiguous() else ' (not contiguous)'}")
        if device is None and a.is_cuda:
            raise EppmError(f"{what}: host memory required, got a CUDA tensor")
        if device is not None and (not a.is_cuda or a.device.index != device):
            raise EppmError(f"{what}: a tensor on cuda:{device} is required, got {a.device}")
        return
    raise EppmError(f"{what}: unsupported buffer type {type(a).__name__}")


class EppmContext:
    """One context per (device, h, w, params): owns all device memory (eppm_create).

    Stream contract: the library enqueues on the context's own non-blocking stream (eppm_stream).  The device-tensor methods of this
    wrapper order that stream AFTER the caller's current torch stream on entry and make the current torch stream wait for it on exit, so
    tensors produced by torch just before a call are complete when the library reads them and the outputs are complete for any torch
    work queued afterwards on the same stream -- no explicit synchronize() needed.  C callers do the same with eppm_stream() and events
    (INTEGRATION.md)."""

    def __init__(self, h, w, max_batch=1, device=0, params=None):
        self.lib = _lib.load()
        self.h, self.w, self.max_batch, self.device = h, w, max_batch, device
        self.params = params if params is not None else default_params()   # what the context was created with (tiled.py derives its schedule from it)
        self._ctx = C.c_void_p()
        rc = self.lib.eppm_create(C.byref(self._ctx), device, h, w, max_batch, C.byref(params) if params is not None else None)
        if rc != 0:
            raise EppmError(f"eppm_create failed ({rc}): {self.lib.eppm_last_error().decode()}")

    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.eppm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise EppmError(f"{what} failed ({rc}): {self.lib.eppm_last_error().decode()}")

    class _Ordered:
        """with ctx._ordered(): the context stream waits for the current torch stream, and the current torch stream for the context stream."""

        def __init__(self, ctx):
            self.pair = None
            torch = sys.modules.get("torch")
            if torch is None or not torch.cuda.is_available():
                return
            cur = torch.cuda.current_stream(ctx.device)
            own = ctx.lib.eppm_stream(ctx._ctx)
            if own is None or cur.cuda_stream == own:
                return
            self.torch = torch
            self.pair = (cur, torch.cuda.ExternalStream(own, device=ctx.device))

        def __enter__(self):
            if self.pair:
                ev = self.torch.cuda.Event()
                ev.record(self.pair[0])
                self.pair[1].wait_event(ev)
            return self

        def __exit__(self, *exc):
            if self.pair:
                ev = self.torch.cuda.Event()
                ev.record(self.pair[1])
                self.pair[0].wait_event(ev)
            return False

    def _ordered(self):
        return EppmContext._Ordered(self)

    # --- geometry ---------------------------------------------------------------------------------
    @property
    def num_levels(self):
        return self.lib.eppm_num_levels(self._ctx)

    def level_dims(self, level):
        h, w = C.c_int(), C.c_int()
        self._check(self.lib.eppm_level_dims(self._ctx, level, C.byref(h), C.byref(w)), "eppm_level_dims")
        return h.value, w.value

    # --- whole pipeline ------------------------------------------------------------------------------
    def compute_batch_host(self, img1, img2, out=None):
        """img1, img2: uint8 [n,h,w,3] host arrays (numpy or pinned torch); returns float32 [n,h,w,2] (u,v).
        n may exceed max_batch (chunked, copies overlapped with compute inside the library)."""
        n = int(img1.shape[0])
        if out is None:
            out = np.empty((n, self.h, self.w, 2), np.float32)
        _check_array(img1, (n, self.h, self.w, 3), np.uint8, "img1")
        _check_array(img2, (n, self.h, self.w, 3), np.uint8, "img2")
        _check_array(out, (n, self.h, self.w, 2), np.float32, "out")
        self._check(self.lib.eppm_compute_batch_host(self._ctx, _ptr(img1), _ptr(img2), n, _ptr(out)), "eppm_compute_batch_host")
        return out

    def compute_batch_device(self, d_img1, d_img2, n, d_flow):
        """Device-resident tensors / raw device addresses; stream-ordered, returns without synchronising.  Tensors may hold more than n
        pairs (a slice of a larger batch is the usual case): the first n are used."""
        for t, c, dt, what in ((d_img1, 3, np.uint8, "d_img1"), (d_img2, 3, np.uint8, "d_img2"), (d_flow, 2, np.float32, "d_flow")):
            if not isinstance(t, int):
                if t.shape[0] < n:
                    raise EppmError(f"{what}: holds {t.shape[0]} pairs, {n} requested")
                _check_array(t, (t.shape[0], self.h, self.w, c), dt, what, device=self.device)
        with self._ordered():
            self._check(self.lib.eppm_compute_batch_device(self._ctx, _ptr(d_img1), _ptr(d_img2), n, _ptr(d_flow)), "eppm_compute_batch_device")

    def compute_stream_device(self, d_frames, n_pairs, d_flow):
        """Consecutive pairs of a frame list [n_pairs+1,h,w,3] (device); each frame is prepared once."""
        if not isinstance(d_frames, int):
            if d_frames.shape[0] < n_pairs + 1:
                raise EppmError(f"d_frames: holds {d_frames.shape[0]} frames, {n_pairs + 1} needed")
            _check_array(d_frames, (d_frames.shape[0], self.h, self.w, 3), np.uint8, "d_frames", device=self.device)
        if not isinstance(d_flow, int):
            if d_flow.shape[0] < n_pairs:
                raise EppmError(f"d_flow: holds {d_flow.shape[0]} pairs, {n_pairs} needed")
            _check_array(d_flow, (d_flow.shape[0], self.h, self.w, 2), np.float32, "d_flow", device=self.device)
        with self._ordered():
            self._check(self.lib.eppm_compute_stream_device(self._ctx, _ptr(d_frames), n_pairs, _ptr(d_flow)), "eppm_compute_stream_device")

    def compute_stream_host(self, frames, out=None):
        """frames: uint8 [n_frames,h,w,3] host array (numpy or pinned torch); returns float32 [n_frames-1,h,w,2]: the flow of every
        consecutive pair.  n_frames may exceed max_batch (chunked, copies overlapped with compute inside the library)."""
        n = int(frames.shape[0])
        if out is None:
            out = np.empty((n - 1, self.h, self.w, 2), np.float32)
        _check_array(frames, (n, self.h, self.w, 3), np.uint8, "frames")
        _check_array(out, (n - 1, self.h, self.w, 2), np.float32, "out")
        self._check(self.lib.eppm_compute_stream_host(self._ctx, _ptr(frames), n, _ptr(out)), "eppm_compute_stream_host")
        return out

    def synchronize(self):
        self._check(self.lib.eppm_synchronize(self._ctx), "eppm_synchronize")

    # --- staged ---------------------------------------------------------------------------------------
    def stage_prepare(self, d_img1, d_img2, n):
        for t, what in ((d_img1, "d_img1"), (d_img2, "d_img2")):
            if not isinstance(t, int):
                _check_array(t, (t.shape[0], self.h, self.w, 3), np.uint8, what, device=self.device)
        with self._ordered():
            self._check(self.lib.eppm_stage_prepare(self._ctx, _ptr(d_img1), _ptr(d_img2), n), "eppm_stage_prepare")

    def stage_patchmatch(self):
        self._check(self.lib.eppm_stage_patchmatch(self._ctx), "eppm_stage_patchmatch")

    def stage_patchmatch_partial(self, n_steps):
        self._check(self.lib.eppm_stage_patchmatch_partial(self._ctx, n_steps), "eppm_stage_patchmatch_partial")

    def write_plane(self, which, arr, level=0, pair=0):
        h, w = self.level_dims(level if which in (PLANE_FLOW, PLANE_FLOW_TMP) else self.num_levels - 1)
        shape, dt = {PLANE_NNF_FWD: ((h, w, 2), np.int16), PLANE_NNF_BWD: ((h, w, 2), np.int16), PLANE_COST_FWD: ((h, w), np.float32),
                     PLANE_COST_BWD: ((h, w), np.float32), PLANE_FLOW: ((h, w, 2), np.float32)}.get(which, (None, None))
        if shape is None:
            raise EppmError(f"write_plane: plane {which} is read-only")
        arr = np.ascontiguousarray(arr)
        _check_array(arr, shape, dt, "write_plane")
        n = self.lib.eppm_write_plane(self._ctx, which, level, pair, arr.ctypes.data)
        if n != arr.nbytes:
            raise EppmError(f"eppm_write_plane returned {n} for {arr.nbytes} bytes: {self.lib.eppm_last_error().decode()}")

    def stage_consistency(self):
        self._check(self.lib.eppm_stage_consistency(self._ctx), "eppm_stage_consistency")

    def stage_c2f(self, d_flow=None):
        if d_flow is not None and not isinstance(d_flow, int):
            _check_array(d_flow, (d_flow.shape[0], self.h, self.w, 2), np.float32, "d_flow", device=self.device)
        with self._ordered():
            self._check(self.lib.eppm_stage_c2f(self._ctx, _ptr(d_flow)), "eppm_stage_c2f")

    def read_plane(self, which, level=0, pair=0):
        h, w = self.level_dims(level)
        if which in (PLANE_NNF_FWD, PLANE_NNF_BWD, PLANE_COST_FWD, PLANE_COST_BWD):
            h, w = self.level_dims(self.num_levels - 1)
        shape, dt = {
            PLANE_RGBA1: ((h, w, 4), np.uint8), PLANE_RGBA2: ((h, w, 4), np.uint8),
            PLANE_CENSUS1: ((h, w), np.uint8), PLANE_CENSUS2: ((h, w), np.uint8),
            PLANE_NNF_FWD: ((h, w, 2), np.int16), PLANE_NNF_BWD: ((h, w, 2), np.int16),
            PLANE_COST_FWD: ((h, w), np.float32), PLANE_COST_BWD: ((h, w), np.float32),
            PLANE_FLOW: ((h, w, 2), np.float32), PLANE_FLOW_TMP: ((h, w, 2), np.float32),
        }[which]
        out = np.empty(shape, dt)
        n = self.lib.eppm_read_plane(self._ctx, which, level, pair, out.ctypes.data)
        if n != out.nbytes:
            raise EppmError(f"eppm_read_plane returned {n}: {self.lib.eppm_last_error().decode()}")
        return out

    def c2f_step(self, level, kind):
        """One coarse-to-fine step on the context's band: 0 = upsample + refine (-> FLOW_TMP), 1 = smoothing, 2 = final smoothing."""
        self._check(self.lib.eppm_tiled_c2f_step(self._ctx, level, kind), "eppm_tiled_c2f_step")

    # --- one large frame tiled over the GPUs of a box (BASELINE config 4), driven by the library over NCCL ------------------------------
    def tiled_init(self, rank, world):
        """Collective over torch.distributed's default group: rank 0 draws the NCCL id (eppm_tiled_unique_id), everyone joins
        (eppm_tiled_init).  The library's own communicator is separate from torch's."""
        import torch
        import torch.distributed as dist
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            self._check(self.lib.eppm_tiled_unique_id(buf), "eppm_tiled_unique_id")
        if world > 1:
            t = torch.tensor(list(buf), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda(self.device)
            dist.broadcast(t, src=0)
            buf = (C.c_ubyte * 128)(*t.cpu().tolist())
        self._check(self.lib.eppm_tiled_init(self._ctx, rank, world, buf), "eppm_tiled_init")

    def compute_tiled_device(self, d_img1, d_img2, d_flow):
        """Collective: the same full pair [1,h,w,3] on every rank, the full flow [1,h,w,2] on every rank; stream-ordered."""
        _check_array(d_img1, (d_img1.shape[0], self.h, self.w, 3), np.uint8, "d_img1", device=self.device)
        _check_array(d_img2, (d_img2.shape[0], self.h, self.w, 3), np.uint8, "d_img2", device=self.device)
        _check_array(d_flow, (d_flow.shape[0], self.h, self.w, 2), np.float32, "d_flow", device=self.device)
        with self._ordered():
            self._check(self.lib.eppm_compute_tiled_device(self._ctx, _ptr(d_img1), _ptr(d_img2), _ptr(d_flow)), "eppm_compute_tiled_device")

    def compute_tiled_host(self, img1, img2, out=None, want_flow=True):
        """Collective, host buffers [1,h,w,3]; returns the flow [1,h,w,2] (None when want_flow is False)."""
        _check_array(img1, (1, self.h, self.w, 3), np.uint8, "img1")
        _check_array(img2, (1, self.h, self.w, 3), np.uint8, "img2")
        if want_flow and out is None:
            out = np.empty((1, self.h, self.w, 2), np.float32)
        if out is not None:
            _check_array(out, (1, self.h, self.w, 2), np.float32, "out")
        self._check(self.lib.eppm_compute_tiled_host(self._ctx, _ptr(img1), _ptr(img2), _ptr(out) if want_flow else None), "eppm_compute_tiled_host")
        return out if want_flow else None

    def eval_flow(self, d_flow, d_gt, n, border=0, outlier_thresh=3.0):
        """Device-side EPE / AAE / outlier share of n flows against ground truth (bao_calc_flow_error semantics); list of dicts."""
        for t, what in ((d_flow, "d_flow"), (d_gt, "d_gt")):
            if not isinstance(t, int):
                if t.shape[0] < n:
                    raise EppmError(f"{what}: holds {t.shape[0]} pairs, {n} requested")
                _check_array(t, (t.shape[0], self.h, self.w, 2), np.float32, what, device=self.device)
        out = (_lib.EppmFlowError * n)()
        with self._ordered():
            self._check(self.lib.eppm_eval_flow(self._ctx, _ptr(d_flow), _ptr(d_gt), n, border, outlier_thresh, out), "eppm_eval_flow")
        return [dict(epe=o.epe, aae_deg=o.aae_deg, outlier_frac=o.outlier_frac, n_valid=o.n_valid, n_known=o.n_known) for o in out]

    def last_stage_ms(self):
        buf = (C.c_float * 5)()
        self._check(self.lib.eppm_last_stage_ms(self._ctx, C.byref(buf)), "eppm_last_stage_ms")
        return dict(zip(("prepare", "patchmatch", "consistency", "c2f", "total"), list(buf)))

    def last_kernel_ms(self, which=0):
        ms = C.c_float()
        self._check(self.lib.eppm_last_kernel_ms(self._ctx, which, C.byref(ms)), "eppm_last_kernel_ms")
        return float(ms.value)

    def launch_count(self, reset=False):
        return int(self.lib.eppm_launch_count(1 if reset else 0))


def write_flo(path, flow):
    """eppm_write_flo: float32 [h,w,2] -> Middlebury .flo."""
    flow = np.ascontiguousarray(flow, np.float32)
    lib = _lib.load()
    rc = lib.eppm_write_flo(str(path).encode(), flow.ctypes.data, flow.shape[0], flow.shape[1])
    if rc != 0:
        raise EppmError(f"eppm_write_flo failed ({rc}): {lib.eppm_last_error().decode()}")


def read_flo(path):
    """eppm_read_flo: Middlebury .flo -> float32 [h,w,2]."""
    lib = _lib.load()
    h, w = C.c_int(), C.c_int()
    rc = lib.eppm_read_flo(str(path).encode(), None, C.byref(h), C.byref(w), 0)
    if rc != 0:
        raise EppmError(f"eppm_read_flo failed ({rc}): {lib.eppm_last_error().decode()}")
    out = np.empty((h.value, w.value, 2), np.float32)
    rc = lib.eppm_read_flo(str(path).encode(), out.ctypes.data, C.byref(h), C.byref(w), out.size)
    if rc != 0:
        raise EppmError(f"eppm_read_flo failed ({rc}): {lib.eppm_last_error().decode()}")
    return out


class BaoFlowPatchmatchMultiscaleCuda:
    """Python mirror of the reference's class: init(h,w) / init(img1,img2,h,w), set_data(img1,img2) -> True,
    compute_flow() -> (u, v) float32 [h,w] arrays (the reference fills caller arrays disp1_x / disp1_y)."""

    def __init__(self, device=0):
        self._device = device
        self._ctx = None
        self._pair = None

    def init(self, *args):
        if len(args) == 2:
            h, w = args
            imgs = None
        elif len(args) == 4:
            img1, img2, h, w = args
            imgs = (img1, img2)
        else:
            raise TypeError("init(h, w) or init(img1, img2, h, w)")
        if self._ctx is not None:
            self._ctx.close()
        self._ctx = EppmContext(h, w, 1, self._device)
        if imgs is not None:
            self.set_data(*imgs)

    def set_data(self, img1, img2):
        if self._ctx is None:
            raise EppmError("set_data before init")
        a = np.ascontiguousarray(img1, np.uint8).reshape(1, self._ctx.h, self._ctx.w, 3)
        b = np.ascontiguousarray(img2, np.uint8).reshape(1, self._ctx.h, self._ctx.w, 3)
        self._pair = (a, b)
        return True  # the reference always returns true (bao_flow_patchmatch_multiscale_cuda.cpp:159-168)

    def compute_flow(self):
        if self._ctx is None or self._pair is None:
            raise EppmError("compute_flow before init/set_data")
        flow = self._ctx.compute_batch_host(self._pair[0], self._pair[1])[0]
        return np.ascontiguousarray(flow[..., 0]), np.ascontiguousarray(flow[..., 1])
