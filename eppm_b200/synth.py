"""Synthetic textured frame pairs with known large-displacement ground truth.

Frame 1 is a band-limited multi-octave value-noise texture (per-channel decorrelated, stretched to u8 [16,240]) with a few
flat-shaded rectangles / discs that create edges.  Motion is piecewise smooth: an affine background (translation up to
+-40 px at full HD, <= 2 % scale / rotation) and 3-6 foreground regions with independent translations up to +-120 px
(<= 30 px at the coarsest pyramid level = SEARCH_RANGE).  Frame 2 is rendered by backward-warping frame 1 layer by
layer (bicubic), so occlusions are filled by the layer behind; N(0,1) u8 noise is added.  Everything derives from a
Philox stream keyed by (0x4550504D, pair index), so the reference arm and this implementation regenerate identical
inputs.  Host-side test/bench data only -- never part of the GPU path."""
import numpy as np
from scipy import ndimage

KEY = 0x4550504D  # "EPPM"


def _rng(pair_idx):
    return np.random.Generator(np.random.Philox(key=[KEY, int(pair_idx)]))


def _value_noise(rng, h, w, octaves=6):
    img = np.zeros((h, w, 3), np.float32)
    amp_sum = 0.0
    for o in range(octaves):
        cell = max(2, 128 >> o)
        gh, gw = h // cell + 3, w // cell + 3
        grid = rng.random((gh, gw, 3), dtype=np.float32)
        yy = (np.arange(h, dtype=np.float32) / cell)[:, None]
        xx = (np.arange(w, dtype=np.float32) / cell)[None, :]
        coords = [np.broadcast_to(yy, (h, w)), np.broadcast_to(xx, (h, w))]
        amp = 0.5 ** (o * 0.6)
        for c in range(3):
            img[..., c] += amp * ndimage.map_coordinates(grid[..., c], coords, order=1, mode="nearest")
        amp_sum += amp
    img /= amp_sum
    lo, hi = np.percentile(img, 1), np.percentile(img, 99)
    return np.clip((img - lo) / max(hi - lo, 1e-6), 0, 1)


def make_pair(h, w, pair_idx=0, scale_to=None, base=None):
    """Returns (img1 u8 [h,w,3], img2 u8 [h,w,3], flow f32 [h,w,2] (u,v), valid bool [h,w] = not occluded and target inside).
    `base` (u8 or float [h,w,3]) replaces the generated texture as frame 1: make_stream chains pairs into a video that way."""
    rng = _rng(pair_idx)
    s = (w / 1920.0) if scale_to is None else scale_to  # motion magnitudes scale with the frame width
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    if base is None:
        tex = _value_noise(rng, h, w)
        # flat shapes that add edges to the texture itself
        for _ in range(8):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            ry, rx = rng.uniform(0.03, 0.12) * h, rng.uniform(0.03, 0.12) * w
            col = rng.uniform(0.1, 0.9, 3).astype(np.float32)
            m = (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1) if rng.random() < 0.5 else ((np.abs(yy - cy) <= ry) & (np.abs(xx - cx) <= rx))
            tex[m] = 0.75 * col + 0.25 * tex[m]
        img1f = 16.0 + 224.0 * tex
    else:
        img1f = np.asarray(base, np.float32)

    # layers: index 0 = background, later = nearer
    layers = []
    ang = rng.uniform(-0.02, 0.02)
    sc = 1.0 + rng.uniform(-0.02, 0.02)
    A = sc * np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]], np.float64)
    c0 = np.array([w / 2.0, h / 2.0])
    t = rng.uniform(-40, 40, 2) * s
    layers.append((np.ones((h, w), bool), A, t + c0 - A @ c0))
    for _ in range(int(rng.integers(3, 7))):
        cy, cx = rng.uniform(0.15, 0.85) * h, rng.uniform(0.15, 0.85) * w
        ry, rx = rng.uniform(0.06, 0.2) * h, rng.uniform(0.06, 0.2) * w
        if rng.random() < 0.5:
            m = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
        else:
            m = (np.abs(yy - cy) <= ry) & (np.abs(xx - cx) <= rx)
        layers.append((m, np.eye(2), rng.uniform(-120, 120, 2) * s))

    # forward ground truth on frame-1 pixels (top-most layer wins) and layer id
    flow = np.zeros((h, w, 2), np.float32)
    lid1 = np.zeros((h, w), np.int32)
    for k, (m, A, t) in enumerate(layers):
        fx = (A[0, 0] - 1) * xx + A[0, 1] * yy + t[0]
        fy = A[1, 0] * xx + (A[1, 1] - 1) * yy + t[1]
        flow[m, 0] = fx[m]
        flow[m, 1] = fy[m]
        lid1[m] = k
    # frame 2 by backward warping, back to front
    img2f = np.zeros_like(img1f)
    lid2 = np.full((h, w), -1, np.int32)
    for k, (m, A, t) in enumerate(layers):
        Ai = np.linalg.inv(A)
        sx = Ai[0, 0] * (xx - t[0]) + Ai[0, 1] * (yy - t[1])
        sy = Ai[1, 0] * (xx - t[0]) + Ai[1, 1] * (yy - t[1])
        inside = (sx >= 0) & (sx <= w - 1) & (sy >= 0) & (sy <= h - 1)
        mm = ndimage.map_coordinates(m.astype(np.float32), [sy, sx], order=0, mode="nearest") > 0.5
        # a source pixel only belongs to this layer if it is not covered by a nearer layer in frame 1
        src_l = ndimage.map_coordinates(lid1.astype(np.float32), [sy, sx], order=0, mode="nearest").astype(np.int32)
        cover = mm & (inside | (k == 0)) & ((src_l == k) | (k == 0))
        for c in range(3):
            samp = ndimage.map_coordinates(img1f[..., c], [sy, sx], order=3, mode="nearest")
            img2f[..., c][cover] = samp[cover]
        lid2[cover] = k
    img2f += rng.standard_normal(img2f.shape).astype(np.float32)
    img1 = np.clip(np.rint(img1f), 0, 255).astype(np.uint8)
    img2 = np.clip(np.rint(img2f), 0, 255).astype(np.uint8)
    # validity: target inside the frame and the target pixel shows the same layer
    tx = np.rint(xx + flow[..., 0]).astype(np.int64)
    ty = np.rint(yy + flow[..., 1]).astype(np.int64)
    inside = (tx >= 0) & (tx < w) & (ty >= 0) & (ty < h)
    valid = inside.copy()
    valid[inside] = lid2[ty[inside], tx[inside]] == lid1[inside]
    return img1, img2, flow, valid


def make_stream(h, w, n_frames, first_idx=0, scale_to=None):
    """A short video: frame t+1 is frame t moved by a fresh piecewise-smooth motion (BASELINE config 5 cycles such a clip).
    Returns (frames u8 [n,h,w,3], flows f32 [n-1,h,w,2], valid bool [n-1,h,w])."""
    frames, flows, valids = [], [], []
    cur = None
    for t in range(n_frames - 1):
        a, b, fl, va = make_pair(h, w, first_idx + t, scale_to=scale_to, base=cur)
        if t == 0:
            frames.append(a)
        frames.append(b); flows.append(fl); valids.append(va)
        cur = b
    return np.stack(frames), np.stack(flows), np.stack(valids)


def make_batch(h, w, n, first_idx=0, distinct=None):
    """n pairs as contiguous uint8 [n,h,w,3] arrays (+ flows, valid masks).  `distinct` < n cycles a few generated pairs."""
    d = n if distinct is None else min(distinct, n)
    pairs = [make_pair(h, w, first_idx + i) for i in range(d)]
    i1 = np.stack([pairs[i % d][0] for i in range(n)])
    i2 = np.stack([pairs[i % d][1] for i in range(n)])
    fl = np.stack([pairs[i % d][2] for i in range(n)])
    va = np.stack([pairs[i % d][3] for i in range(n)])
    return i1, i2, fl, va


def epe(flow, gt, mask=None):
    """Mean end-point error (bao_calc_flow_error semantics: Euclidean distance per pixel, basic/bao_flow_tools.cpp:64-111)."""
    e = np.sqrt(((flow.astype(np.float64) - gt.astype(np.float64)) ** 2).sum(-1))
    if mask is not None:
        e = e[mask]
    return float(e.mean()) if e.size else 0.0


def write_flo(path, flow):
    """Middlebury .flo: "PIEH", int32 w, int32 h, rows of interleaved (u,v) float32 (3rdparty/middlebury/README.txt:9-24)."""
    h, w = flow.shape[:2]
    with open(path, "wb") as f:
        f.write(b"PIEH")
        np.array([w, h], np.int32).tofile(f)
        np.ascontiguousarray(flow, np.float32).tofile(f)


def read_flo(path):
    with open(path, "rb") as f:
        if f.read(4) != b"PIEH":
            raise ValueError("bad .flo tag")
        w, h = np.fromfile(f, np.int32, 2)
        return np.fromfile(f, np.float32, 2 * w * h).reshape(h, w, 2)


def read_ppm(path):
    """Binary P6 reader tolerant of '#' comment lines (the shipped frame10/11.ppm carry one)."""
    with open(path, "rb") as f:
        data = f.read()
    toks, pos = [], 0
    while len(toks) < 4:
        while data[pos:pos + 1].isspace():
            pos += 1
        if data[pos:pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end:end + 1].isspace():
            end += 1
        toks.append(data[pos:end])
        pos = end
    pos += 1
    if toks[0] != b"P6":
        raise ValueError("not a P6 file")
    w, h = int(toks[1]), int(toks[2])
    return np.frombuffer(data, np.uint8, w * h * 3, pos).reshape(h, w, 3).copy()
