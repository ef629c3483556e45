"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI of libeppm_b200.so
(include/eppm.h, include/eppm_legacy_abi.h).  Three kinds of oracle:
  * the reference's own CUDA build, oracle/_ref/libeppm_ref.so (prebuilt in the dev container, travels with the snapshot),
    driven with IDENTICAL device buffers through the reference's own stage functions -> bit-exact assertions;
  * committed golden fixtures generated from that build (tests/golden/ref_*.npz, tools/gen_golden.py);
  * the CPU oracle oracle/golden.cpp at small sizes (tolerance: it cannot reproduce MUFU.EX2).
Racy stages of the reference (in-place outlier removal / weighted median / flow smoothing) are compared under this
library's snapshot semantics with explicit bounds; see DESIGN.md "racy stages"."""
import ctypes as C
import glob
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import eppm_b200 as E
from eppm_b200 import synth, _lib
import refharness

torch = pytest.importorskip("torch")
needs_ref = pytest.mark.skipif(not refharness.available(), reason="oracle/_ref not built")
P = lambda t: t.data_ptr()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ref():
    return refharness.Ref()


@pytest.fixture(scope="module")
def mine():
    return _lib.load()


def same_bits(a, b):
    a = np.asarray(a); b = np.asarray(b)
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32))
    return np.array_equal(a, b)


SIZES = [(96, 128, 7, 0.12), (121, 161, 3, 0.12), (436, 1024, 1, None), (480, 640, 0, None)]


# ------------------------------------------------------------------------------------------------ prepare
@needs_ref
@pytest.mark.parametrize("h,w,idx,scale", SIZES)
def test_prepare_bit_exact_vs_reference(ref, h, w, idx, scale):
    """census codes, pyramid bytes and level geometry are bit-exact (north star), incl. odd sizes (generic resize path)."""
    a, b, _, _ = synth.make_pair(h, w, idx, scale_to=scale)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    ctx = E.EppmContext(h, w, 1)
    ctx.stage_prepare(dev(a[None]), dev(b[None]), 1)
    for l in range(ctx.num_levels):
        assert ctx.level_dims(l) == ref.level_dims(rc, l)
        for which in (E.PLANE_RGBA1, E.PLANE_RGBA2, E.PLANE_CENSUS1, E.PLANE_CENSUS2):
            assert same_bits(ctx.read_plane(which, l), ref.read_plane(rc, which, l)), (which, l)
    ref.destroy(rc); ctx.close()


@needs_ref
@pytest.mark.parametrize("h,w,idx", [(40, 56, 21), (24, 40, 22), (33, 17, 23), (20, 200, 24)])
def test_tiny_and_extreme_shapes_vs_reference(ref, mine, h, w, idx):
    """Frames smaller than the patch and the search window, odd and strongly non-square: pyramid, census and the whole PatchMatch at
    the coarsest level (a handful of pixels, every sample in the clamped border) stay bit-exact; end to end the flow shows the
    reference's own degenerate behaviour."""
    a, b, _, _ = synth.make_pair(h, w, idx, scale_to=0.1)
    rc, dims, img, cen = _ref_level_planes(ref, h, w, a, b)
    ctx = E.EppmContext(h, w, 1)
    ctx.stage_prepare(dev(a[None]), dev(b[None]), 1)
    for l in range(3):
        assert ctx.level_dims(l) == tuple(dims[l])
        for which in (E.PLANE_RGBA1, E.PLANE_CENSUS1, E.PLANE_CENSUS2):
            assert same_bits(ctx.read_plane(which, l), ref.read_plane(rc, which, l)), (which, l)
    hc, wc = dims[2]
    for swap in (False, True):
        (nr, cr), (nm, cm) = _pm(ref.lib, img, cen, 2, wc, hc, swap), _pm(mine, img, cen, 2, wc, hc, swap)
        assert torch.equal(nr, nm) and same_bits(cr.cpu().numpy(), cm.cpu().numpy()), (h, w, swap)
    fr = ref.compute_flow(rc, h, w)
    fm = ctx.compute_batch_host(a[None], b[None])[0]
    # End to end the reference degenerates on such frames (its 13x13 outlier vote spans the whole coarsest level and rejects every
    # pixel, the 1e10 unknown marker then leaks through the x2 bilinear upsampling as values of 2e4..3e4 px): the drop-in must
    # degenerate the same way -- same share of blown-up pixels, same extreme value -- not "fix" it.
    assert fm.shape == fr.shape == (h, w, 2) and np.isfinite(fm).all()
    assert abs((np.abs(fm) > 1e3).mean() - (np.abs(fr) > 1e3).mean()) <= 0.02
    assert abs(np.abs(fm).max() - np.abs(fr).max()) <= 0.01 * np.abs(fr).max()
    ref.destroy(rc); ctx.close()


def _ref_level_planes(ref, h, w, a, b):
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    dims = [ref.level_dims(rc, l) for l in range(3)]
    img = [[refharness.pitched(ref.read_plane(rc, k, l)) for l in range(3)] for k in (0, 1)]
    cen = [[refharness.pitched(ref.read_plane(rc, 2 + k, l)) for l in range(3)] for k in (0, 1)]
    return rc, dims, img, cen


def _pm(lib, img, cen, L, wc, hc, swap=False):
    nnf = torch.zeros((hc, wc, 2), dtype=torch.int16, device="cuda")
    cost = torch.zeros((hc, wc), dtype=torch.float32, device="cuda")
    fn = lib.baoCudaPatchMatch
    fn.argtypes = [C.c_void_p] * 6 + [C.c_int, C.c_int] + [C.c_size_t] * 4
    fn.restype = None
    a, b = (1, 0) if swap else (0, 1)
    fn(P(nnf), P(cost), P(img[a][L][0]), P(img[b][L][0]), P(cen[a][L][0]), P(cen[b][L][0]), wc, hc, img[0][L][1], wc * 4, wc * 4, cen[0][L][1])
    torch.cuda.synchronize()
    return nnf, cost


@pytest.fixture(scope="module")
def chain(ref, mine):
    """Reference stage chain on one 640x480 synthetic pair; every later test feeds both libraries the reference's state."""
    if not refharness.available():
        pytest.skip("oracle/_ref not built")
    h, w = 480, 640
    a, b, gt, valid = synth.make_pair(h, w, 0)
    rc, dims, img, cen = _ref_level_planes(ref, h, w, a, b)
    hc, wc = dims[2]
    st = {"h": h, "w": w, "a": a, "b": b, "gt": gt, "valid": valid, "rc": rc, "dims": dims, "img": img, "cen": cen, "hc": hc, "wc": wc}
    st["pm_ref"] = [_pm(ref.lib, img, cen, 2, wc, hc, s) for s in (False, True)]
    return st


# ------------------------------------------------------------------------------------------------ PatchMatch
@needs_ref
def test_patchmatch_bit_exact_vs_reference(ref, mine, chain):
    """With the reference's XORWOW stream reproduced, NNF and cost after all 10 iterations are bit-exact, both directions."""
    for s, swap in enumerate((False, True)):
        nnf, cost = _pm(mine, chain["img"], chain["cen"], 2, chain["wc"], chain["hc"], swap)
        assert torch.equal(nnf, chain["pm_ref"][s][0])
        assert same_bits(cost.cpu().numpy(), chain["pm_ref"][s][1].cpu().numpy())


@needs_ref
@pytest.mark.parametrize("n_steps_ref,n_steps_mine", [(1, 0), (2, 1), (3, 2), (4, 3), (5, 4), (6, 5), (7, 6), (12, 11)])
def test_patchmatch_every_pass_bit_exact(ref, chain, n_steps_ref, n_steps_mine):
    """Stage taps inside PatchMatch: random field, cost field, each propagation pass, random search (reference numbering:
    field and cost are two launches; here they are one kernel, hence the offset)."""
    if n_steps_mine == 0:
        pytest.skip("random field alone is covered by the rand_field fixture test")
    h, w = chain["h"], chain["w"]
    img, cen, wc, hc = chain["img"], chain["cen"], chain["wc"], chain["hc"]
    nf, cf = ref.tap_patchmatch(img[0][2], img[1][2], cen[0][2], cen[1][2], wc, hc, n_steps_ref)
    nb, cb = ref.tap_patchmatch(img[1][2], img[0][2], cen[1][2], cen[0][2], wc, hc, n_steps_ref)
    ctx = E.EppmContext(h, w, 1)
    ctx.stage_prepare(dev(chain["a"][None]), dev(chain["b"][None]), 1)
    ctx.stage_patchmatch_partial(n_steps_mine)
    assert same_bits(ctx.read_plane(E.PLANE_NNF_FWD), nf)
    assert same_bits(ctx.read_plane(E.PLANE_NNF_BWD), nb)
    assert same_bits(ctx.read_plane(E.PLANE_COST_FWD), cf)
    assert same_bits(ctx.read_plane(E.PLANE_COST_BWD), cb)
    ctx.close()


# ------------------------------------------------------------------------------------------------ consistency + c2f, stage by stage
def _both(ref, mine, call, outs):
    r = [t.clone() for t in outs]; m = [t.clone() for t in outs]
    call(ref.lib, r); call(mine, m)
    torch.cuda.synchronize()
    return r, m


@needs_ref
def test_consistency_and_c2f_stage_by_stage(ref, mine, chain):
    img, cen, dims, wc, hc = chain["img"], chain["cen"], chain["dims"], chain["wc"], chain["hc"]
    (nf, cf), (nb, cb) = chain["pm_ref"]
    i1 = img[0][2]
    n_px = wc * hc
    # left-right check: deterministic -> bit-exact
    r, m = _both(ref, mine, lambda lib, t: lib.baoCudaLeftRightCheck(P(t[0]), P(t[1]), P(t[2]), P(t[3]), wc, hc, wc * 4, wc * 4), [nf, cf, nb, cb])
    for x, y in zip(r, m):
        assert torch.equal(x, y)
    # outlier removal: the reference reads and writes the same array (race); snapshot semantics may differ on a few pixels
    r2, m2 = _both(ref, mine, lambda lib, t: lib.baoCudaOutlierRemoval(P(t[0]), P(t[1]), wc, hc, wc * 4, wc * 4), [r[0], r[1]])
    assert (r2[0] != m2[0]).any(-1).float().mean().item() <= 0.0015 * RACY_SLACK   # measured 0.089 % (17 of 19 200), profiles/r01_parity_stages.json
    # weighted median, 20 in-place sweeps in the reference (race) vs 20 snapshot sweeps
    wmf = lambda lib, t: lib.baoCudaWeightedMedianFilter(P(t[0]), P(t[1]), P(i1[0]), wc, hc, i1[1], wc * 4, wc * 4, 20, True)
    r3, m3 = _both(ref, mine, wmf, [r2[0], r2[1]])
    assert (r3[0] != m3[0]).any(-1).float().mean().item() <= 0.015 * RACY_SLACK   # measured 0.96 % (184 of 19 200)
    left_r = ((r3[0] < 0).any(-1)).sum().item(); left_m = ((m3[0] < 0).any(-1)).sum().item()
    assert abs(left_r - left_m) <= 0.002 * n_px * RACY_SLACK
    # hole filling on the reference's state: holes are isolated -> bit-exact
    r4, m4 = _both(ref, mine, lambda lib, t: lib.baoCudaFillHole(P(t[0]), P(t[1]), P(i1[0]), wc, hc, i1[1], wc * 4, wc * 4), [r3[0], r3[1]])
    assert torch.equal(r4[0], m4[0])
    flow = torch.zeros((hc, wc, 2), dtype=torch.float32, device="cuda")
    r5, m5 = _both(ref, mine, lambda lib, t: lib.baoCudaNNF2Flow(P(t[0]), P(t[1]), wc, hc, wc * 4, wc * 8), [flow, r4[0]])
    assert same_bits(r5[0].cpu().numpy(), m5[0].cpu().numpy())
    cur = r5[0]
    nl = 3
    PtrArr, IntArr, SzArr = C.c_void_p * nl, C.c_int * nl, C.c_size_t * nl
    for l in (1, 0):
        hl, wl = dims[l]
        fine = torch.zeros((hl, wl, 2), dtype=torch.float32, device="cuda")

        def c2f(lib, t, l=l):
            flows = [None] * nl
            flows[l] = t[0]; flows[l + 1] = t[1]
            lib.baoCudaBLF_C2F.argtypes = [C.c_void_p] * 11 + [C.c_int]
            lib.baoCudaBLF_C2F.restype = None
            lib.baoCudaBLF_C2F(PtrArr(*[P(x) if x is not None else None for x in flows]), PtrArr(*[P(img[0][k][0]) for k in range(nl)]),
                               PtrArr(*[P(img[1][k][0]) for k in range(nl)]), PtrArr(*[P(cen[0][k][0]) for k in range(nl)]),
                               PtrArr(*[P(cen[1][k][0]) for k in range(nl)]), None, None, IntArr(*[d[0] for d in dims]), IntArr(*[d[1] for d in dims]),
                               SzArr(*[img[0][k][1] for k in range(nl)]), SzArr(*[cen[0][k][1] for k in range(nl)]), l)
        # x2 upsample + plane-fitting refine (78 % of all patch samples): deterministic -> bit-exact
        rr, mm = _both(ref, mine, c2f, [fine, cur])
        assert same_bits(rr[0].cpu().numpy(), mm[0].cpu().numpy()), f"plane-fitting refine differs at level {l}"
        # joint-bilateral smoothing: in place in the reference (race) -> bounded difference
        sm = lambda lib, t, l=l, hl=hl, wl=wl: lib.baoCudaFlowSmoothing(P(t[0]), P(img[0][l][0]), wl, hl, img[0][l][1], wl * 8)
        rs, ms = _both(ref, mine, sm, [rr[0]])
        d = (rs[0] - ms[0]).abs()
        assert d.mean().item() <= 1e-4 * RACY_SLACK and d.max().item() <= 0.25 * RACY_SLACK   # measured mean 5e-5, max 0.12 px
        cur = rs[0]


@needs_ref
def test_uncalled_stage_functions_bit_exact(ref, mine, chain):
    """Stage functions the reference's host class declares but compute_flow never calls (…cuda.cpp:40-62): buffered left-right check,
    flow -> NNF, flow cut-off, still-region elimination.  All deterministic -> bit-exact against the reference build on identical buffers."""
    img, dims, wc, hc = chain["img"], chain["dims"], chain["wc"], chain["hc"]
    (nf, cf), (nb, cb) = chain["pm_ref"]
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    for lib in (ref.lib, mine):
        lib.baoCudaLeftRightCheck_Buffered.argtypes = [V] * 6 + [I, I, S, S]; lib.baoCudaLeftRightCheck_Buffered.restype = None
        lib.baoCudaFlow2NNF.argtypes = [V, V, I, I, S, S]; lib.baoCudaFlow2NNF.restype = None
        lib.baoCudaFlowCutoff.argtypes = [V, I, I, S, C.c_float]; lib.baoCudaFlowCutoff.restype = None
        lib.baoEliminateStillRegionFlow.argtypes = [V, V, V, I, I, S]; lib.baoEliminateStillRegionFlow.restype = None
    tn, tc = torch.zeros_like(nf), torch.zeros_like(cf)
    r, m = _both(ref, mine, lambda lib, t: lib.baoCudaLeftRightCheck_Buffered(P(t[0]), P(t[1]), P(t[2]), P(t[3]), P(t[4]), P(t[5]), wc, hc, wc * 4, wc * 4),
                 [nf, cf, nb, cb, tn, tc])
    for x, y in zip(r[:4], m[:4]):
        assert same_bits(x.cpu().numpy(), y.cpu().numpy())
    assert (r[0] < 0).any().item() and (r[0] >= 0).any().item()   # the check rejected some pixels and kept others
    # flow -> NNF on a flow with fractional, negative, out-of-range and unknown entries
    g = torch.Generator(device="cpu").manual_seed(5)
    fl = (torch.rand((hc, wc, 2), generator=g) * 200 - 100)
    fl[3:9, 5:40] = 1e10
    fl[10, :] = 40000.0   # beyond the short range: the conversion saturates / wraps exactly like the reference's
    fl = fl.cuda()
    out = torch.zeros((hc, wc, 2), dtype=torch.int16, device="cuda")
    r, m = _both(ref, mine, lambda lib, t: lib.baoCudaFlow2NNF(P(t[0]), P(t[1]), wc, hc, wc * 4, wc * 8), [out, fl])
    assert torch.equal(r[0], m[0])
    # cut-off (NaN and infinities included: the macros' comparison order decides what survives)
    fl2 = fl.clone(); fl2[0, 0, 0] = float("nan"); fl2[0, 1, 1] = float("inf"); fl2[0, 2, 0] = float("-inf")
    r, m = _both(ref, mine, lambda lib, t: lib.baoCudaFlowCutoff(P(t[0]), wc, hc, wc * 8, 37.5), [fl2])
    assert same_bits(r[0].cpu().numpy(), m[0].cpu().numpy())
    # still-region elimination at level 1: image 2 := image 1 on the left half, the real second frame on the right
    h1, w1 = dims[1]
    a1, pitch = img[0][1]
    b1 = img[1][1][0].clone()
    b1[:, : (w1 // 2) * 4] = a1[:, : (w1 // 2) * 4]
    fl3 = torch.full((h1, w1, 2), 3.25, dtype=torch.float32, device="cuda")
    r, m = _both(ref, mine, lambda lib, t: lib.baoEliminateStillRegionFlow(P(t[0]), P(a1), P(b1), w1, h1, pitch), [fl3])
    assert same_bits(r[0].cpu().numpy(), m[0].cpu().numpy())
    z = (r[0] == 0).all(-1)
    assert z[:, : w1 // 2 - 12].all().item() and not z[:, w1 // 2 + 12:].all().item()


@needs_ref
def test_flow_bilateral_upsampling_bit_exact(ref, mine, chain):
    """baoCudaFlowBilteralUpsampling (its only call site is commented out upstream): out of place -> deterministic -> bit-exact."""
    img, dims = chain["img"], chain["dims"]
    (h1, w1), (h2, w2) = dims[1], dims[2]
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    g = torch.Generator(device="cpu").manual_seed(9)
    small = torch.randn((h2, w2, 2), generator=g) * 4
    small[5:20, 10:60] = 1e10                       # a block of unknown flow wider than the filter radius at the fine level
    small = small.cuda()
    outs = []
    for lib in (ref.lib, mine):
        lib.baoCudaFlowBilteralUpsampling.argtypes = [V, V, I, I, S, V, I, I, C.c_float]; lib.baoCudaFlowBilteralUpsampling.restype = None
        out = torch.full((h1, w1, 2), -7.0, dtype=torch.float32, device="cuda")
        lib.baoCudaFlowBilteralUpsampling(P(out), P(img[0][1][0]), w1, h1, img[0][1][1], P(small), w2, h2, 2.0)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    assert same_bits(outs[0], outs[1]), f"{(outs[0].view(np.uint32) != outs[1].view(np.uint32)).sum()} floats differ"
    assert (outs[0] == -7.0).all(-1).any() and (outs[0] != -7.0).any()   # untouched where every tap is unknown, written elsewhere


@needs_ref
def test_image_smoothing_bit_exact(ref, mine, chain):
    """baoCudaImageSmoothing (never called upstream): out of place -> deterministic; r, g, b bit-exact (the reference leaves alpha uninitialised)."""
    img, dims = chain["img"], chain["dims"]
    h2, w2 = dims[2]
    src, pitch = img[0][2]
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    outs = []
    for lib in (ref.lib, mine):
        lib.baoCudaImageSmoothing.argtypes = [V, V, I, I, S]; lib.baoCudaImageSmoothing.restype = None
        out = torch.zeros_like(src)
        lib.baoCudaImageSmoothing(P(out), P(src), w2, h2, pitch)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy()[:, : w2 * 4].reshape(h2, w2, 4)[..., :3])
    assert np.array_equal(outs[0], outs[1]), f"{(outs[0] != outs[1]).sum()} bytes differ"
    assert (outs[0] != src.cpu().numpy()[:, : w2 * 4].reshape(h2, w2, 4)[..., :3]).mean() > 0.05   # it really filtered


@needs_ref
def test_patchmatch_planefitting_bit_exact(ref, mine, chain):
    """baoCudaPatchMatch_PlaneFitting (declared, never called by the reference's host class): the whole PatchMatch scored with the
    four-model plane-fitting cost.  Same random stream, same lock-step order -> NNF and cost bit-exact."""
    img, cen, wc, hc = chain["img"], chain["cen"], chain["wc"], chain["hc"]
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    outs = {}
    for name, lib in (("ref", ref.lib), ("mine", mine)):
        lib.baoCudaPatchMatch_PlaneFitting.argtypes = [V] * 6 + [I, I, S, S, S, S]; lib.baoCudaPatchMatch_PlaneFitting.restype = None
        nnf = torch.zeros((hc, wc, 2), dtype=torch.int16, device="cuda"); cost = torch.zeros((hc, wc), dtype=torch.float32, device="cuda")
        lib.baoCudaPatchMatch_PlaneFitting(P(nnf), P(cost), P(img[0][2][0]), P(img[1][2][0]), P(cen[0][2][0]), P(cen[1][2][0]), wc, hc,
                                           img[0][2][1], wc * 4, wc * 4, cen[0][2][1])
        torch.cuda.synchronize()
        outs[name] = (nnf.cpu().numpy(), cost.cpu().numpy())
    assert np.array_equal(outs["ref"][0], outs["mine"][0]), f"{(outs['ref'][0] != outs['mine'][0]).any(-1).sum()} targets differ"
    assert same_bits(outs["ref"][1], outs["mine"][1])
    (nf, _), _ = chain["pm_ref"]
    assert (outs["ref"][0] != nf.cpu().numpy()).any()   # and it is not the plain PatchMatch


@needs_ref
def test_patchmatch_scaled_bit_exact(ref, mine, chain):
    """baoCudaPatchMatch_Scaled (declared, never called, unfinished upstream: bao_pmflow_kernel.cu:1828-1895): PatchMatch over (target, patch
    scale) with the AD-only cost, mirrored with its quirks (the forward row pass stores the winning scale into the cost plane).  Same random
    stream (the scale comes from the second draw of a pixel), same lock-step order -> targets, scales and the cost plane bit-exact."""
    img, cen, wc, hc = chain["img"], chain["cen"], chain["wc"], chain["hc"]
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    outs = {}
    for name, lib in (("ref", ref.lib), ("mine", mine)):
        lib.baoCudaPatchMatch_Scaled.argtypes = [V] * 7 + [I, I, S, S, S, S, S]; lib.baoCudaPatchMatch_Scaled.restype = None
        nnf = torch.zeros((hc, wc, 2), dtype=torch.int16, device="cuda")
        scale = torch.zeros((hc, wc), dtype=torch.float32, device="cuda"); cost = torch.zeros((hc, wc), dtype=torch.float32, device="cuda")
        lib.baoCudaPatchMatch_Scaled(P(nnf), P(scale), P(cost), P(img[0][2][0]), P(img[1][2][0]), P(cen[0][2][0]), P(cen[1][2][0]), wc, hc,
                                     img[0][2][1], wc * 4, wc * 4, wc * 4, cen[0][2][1])
        torch.cuda.synchronize()
        outs[name] = (nnf.cpu().numpy(), scale.cpu().numpy(), cost.cpu().numpy())
    assert np.array_equal(outs["ref"][0], outs["mine"][0]), f"{(outs['ref'][0] != outs['mine'][0]).any(-1).sum()} targets differ"
    assert same_bits(outs["ref"][1], outs["mine"][1]), f"{(outs['ref'][1] != outs['mine'][1]).sum()} scales differ"
    assert same_bits(outs["ref"][2], outs["mine"][2]), f"{(outs['ref'][2] != outs['mine'][2]).sum()} costs differ"
    sc = outs["mine"][1]
    assert sc.min() >= 0.6 - 1e-6 and sc.max() <= 1.4 + 1e-6 and len(np.unique(sc)) == 9   # (r % 9 + 6) / 10
    (nf, _), _ = chain["pm_ref"]
    assert (outs["ref"][0] != nf.cpu().numpy()).any()   # and it is not the plain PatchMatch


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "refscaled_*.npz"))))
def test_patchmatch_scaled_against_committed_reference_fixture(mine, path):
    """baoCudaPatchMatch_Scaled against the reference build's output committed under tests/golden (tools/gen_golden_scaled.py): targets, scales and
    cost plane bit-exact; runs without oracle/_ref.  The census planes are passed as NULL: the scaled cost never reads them."""
    z = np.load(path)
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    mine.baoCudaPatchMatch_Scaled.argtypes = [V] * 7 + [I, I, S, S, S, S, S]; mine.baoCudaPatchMatch_Scaled.restype = None
    hc, wc = z["rgba1_L2"].shape[:2]
    i1, pitch = refharness.pitched(z["rgba1_L2"]); i2, _ = refharness.pitched(z["rgba2_L2"])
    nn = torch.zeros((hc, wc, 2), dtype=torch.int16, device="cuda")
    sc = torch.zeros((hc, wc), dtype=torch.float32, device="cuda"); co = torch.zeros((hc, wc), dtype=torch.float32, device="cuda")
    mine.baoCudaPatchMatch_Scaled(P(nn), P(sc), P(co), P(i1), P(i2), None, None, wc, hc, pitch, wc * 4, wc * 4, wc * 4, 0)
    torch.cuda.synchronize()
    assert np.array_equal(nn.cpu().numpy(), z["sc_nnf"]), f"{(nn.cpu().numpy() != z['sc_nnf']).any(-1).sum()} targets differ"
    assert same_bits(sc.cpu().numpy(), z["sc_scale"]) and same_bits(co.cpu().numpy(), z["sc_cost"])


@needs_ref
def test_subpixel_refine_and_bicubic_census_vs_reference(mine, chain, tmp_path):
    """baoCudaCensusTransform_Bicubic + baoCudaSubpixRefine (declared by the reference's host class, not called by
    compute_flow).  Both are deterministic -> bit-exact on identical buffers.  The reference build is loaded from a PRIVATE copy of its
    library: its sub-pixel entry point switches the file-scope image texture reference to linear filtering and never switches it back,
    which would change the reference's own weighted-median / hole-filling results for every later test in this process."""
    import shutil, time
    priv = os.path.join(str(tmp_path), "libeppm_ref_subpix.so")
    shutil.copy(refharness.REF_LIB, priv)
    rlib = C.CDLL(priv)
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    for lib in (rlib, mine):
        lib.baoCudaCensusTransform_Bicubic.argtypes = [V, V, I, I, S, V, V, I, I, S]; lib.baoCudaCensusTransform_Bicubic.restype = None
        lib.baoCudaSubpixRefine.argtypes = [V] * 6 + [I, I, S, S, S, S]; lib.baoCudaSubpixRefine.restype = None
        lib.baoCudaNNF2Flow.argtypes = [V, V, I, I, S, S]; lib.baoCudaNNF2Flow.restype = None
    img, wc, hc = chain["img"], chain["wc"], chain["hc"]
    (nf, cf), _ = chain["pm_ref"]
    i1, pitch = img[0][2]
    i2 = img[1][2][0]
    wu, hu = 2 * wc, 2 * hc
    cpitch = (wu + 511) // 512 * 512
    res = {}
    for name, lib in (("ref", rlib), ("mine", mine)):
        c1 = torch.zeros((hu, cpitch), dtype=torch.uint8, device="cuda"); c2 = torch.zeros_like(c1)
        lib.baoCudaCensusTransform_Bicubic(P(c1), P(c2), wu, hu, cpitch, P(i1), P(i2), wc, hc, pitch)
        torch.cuda.synchronize()
        res[name] = (c1, c2)
    for k in range(2):
        assert torch.equal(res["ref"][k][:, :wu], res["mine"][k][:, :wu]), f"bicubic census of image {k + 1} differs"
    assert len(torch.unique(res["ref"][0][:, :wu])) > 100
    c1, c2 = res["ref"]
    flows, ms = {}, {}
    for name, lib in (("ref", rlib), ("mine", mine)):
        fl = torch.zeros((hc, wc, 2), dtype=torch.float32, device="cuda")
        lib.baoCudaNNF2Flow(P(fl), P(nf), wc, hc, wc * 4, wc * 8)
        base = fl.clone()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lib.baoCudaSubpixRefine(P(fl), P(nf), P(i1), P(i2), P(c1), P(c2), wc, hc, pitch, cpitch, wc * 4, wc * 8)
        torch.cuda.synchronize(); ms[name] = (time.perf_counter() - t0) * 1e3
        flows[name] = (fl.cpu().numpy(), base.cpu().numpy())
    fr, base = flows["ref"]
    fm, _ = flows["mine"]
    changed = (fr != base).any(-1)
    assert 0.2 < changed.mean() < 1.0, changed.mean()          # the stage really moved a good part of the field off the integer grid
    assert np.abs(fr - base)[changed].max() <= 1.5 + 1e-6      # by at most 3 half-pixels
    diff = fr.view(np.uint32) != fm.view(np.uint32)
    print(f"subpix refine {wc}x{hc}: reference {ms['ref']:.2f} ms, this library {ms['mine']:.2f} ms, {changed.mean() * 100:.1f} % of the pixels refined, "
          f"{int(diff.sum())} floats differ")
    assert not diff.any(), f"{int(diff.sum())} of {diff.size} floats differ, max abs {np.abs(fr - fm).max()}"


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "refsub_*.npz"))))
def test_subpixel_refine_against_committed_reference_fixture(mine, path):
    """The same stage against outputs of the reference build committed under tests/golden (tools/gen_golden_subpix.py): runs without oracle/_ref."""
    z = np.load(path)
    h, w = int(z["h"]), int(z["w"])
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    mine.baoCudaCensusTransform_Bicubic.argtypes = [V, V, I, I, S, V, V, I, I, S]; mine.baoCudaCensusTransform_Bicubic.restype = None
    mine.baoCudaSubpixRefine.argtypes = [V] * 6 + [I, I, S, S, S, S]; mine.baoCudaSubpixRefine.restype = None
    i1, pitch = refharness.pitched(z["rgba1"]); i2, _ = refharness.pitched(z["rgba2"])
    wu, hu = 2 * w, 2 * h
    cp = (wu + 511) // 512 * 512
    u1 = torch.zeros((hu, cp), dtype=torch.uint8, device="cuda"); u2 = torch.zeros_like(u1)
    mine.baoCudaCensusTransform_Bicubic(P(u1), P(u2), wu, hu, cp, P(i1), P(i2), w, h, pitch)
    torch.cuda.synchronize()
    assert np.array_equal(u1[:, :wu].cpu().numpy(), z["census1_up"]) and np.array_equal(u2[:, :wu].cpu().numpy(), z["census2_up"])
    fl = dev(z["flow_in"]); nnf = dev(z["nnf"])
    mine.baoCudaSubpixRefine(P(fl), P(nnf), P(i1), P(i2), P(u1), P(u2), w, h, pitch, cp, w * 4, w * 8)
    torch.cuda.synchronize()
    assert same_bits(fl.cpu().numpy(), z["flow_out"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "refstage_*.npz"))))
def test_uncalled_stage_functions_against_committed_reference_fixture(mine, path):
    """The stage functions compute_flow never calls against outputs of the reference build committed under tests/golden
    (tools/gen_golden_stages.py): bit-exact, and runs without oracle/_ref."""
    z = np.load(path)
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    mine.baoCudaLeftRightCheck_Buffered.argtypes = [V] * 6 + [I, I, S, S]
    mine.baoCudaFlow2NNF.argtypes = [V, V, I, I, S, S]
    mine.baoCudaFlowCutoff.argtypes = [V, I, I, S, C.c_float]
    mine.baoEliminateStillRegionFlow.argtypes = [V, V, V, I, I, S]
    mine.baoCudaImageSmoothing.argtypes = [V, V, I, I, S]
    mine.baoCudaFlowBilteralUpsampling.argtypes = [V, V, I, I, S, V, I, I, C.c_float]
    h1, w1 = z["rgba1_L1"].shape[:2]
    h2, w2 = z["up_small"].shape[:2]
    t = [dev(z["lr_nnf1"]), dev(z["lr_cost1"]), dev(z["lr_nnf2"]), dev(z["lr_cost2"])]
    tn, tc = torch.zeros_like(t[0]), torch.zeros_like(t[1])
    mine.baoCudaLeftRightCheck_Buffered(P(t[0]), P(t[1]), P(t[2]), P(t[3]), P(tn), P(tc), w1, h1, w1 * 4, w1 * 4)
    torch.cuda.synchronize()
    for got, key in zip(t, ("lr_out_nnf1", "lr_out_cost1", "lr_out_nnf2", "lr_out_cost2")):
        assert same_bits(got.cpu().numpy(), z[key]), key
    o = torch.zeros((h1, w1, 2), dtype=torch.int16, device="cuda"); d = dev(z["f2n_flow"])
    mine.baoCudaFlow2NNF(P(o), P(d), w1, h1, w1 * 4, w1 * 8); torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy(), z["f2n_nnf"])
    d = dev(z["f2n_flow"]); mine.baoCudaFlowCutoff(P(d), w1, h1, w1 * 8, 37.5); torch.cuda.synchronize()
    assert same_bits(d.cpu().numpy(), z["cutoff_out"])
    ia, pitch = refharness.pitched(z["rgba1_L1"]); ib, _ = refharness.pitched(z["still_img2"])
    d = torch.full((h1, w1, 2), 3.25, dtype=torch.float32, device="cuda")
    mine.baoEliminateStillRegionFlow(P(d), P(ia), P(ib), w1, h1, pitch); torch.cuda.synchronize()
    assert same_bits(d.cpu().numpy(), z["still_out"])
    o = torch.zeros_like(ia)
    mine.baoCudaImageSmoothing(P(o), P(ia), w1, h1, pitch); torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy()[:, : w1 * 4].reshape(h1, w1, 4)[..., :3], z["smooth_out"])
    o = torch.full((h1, w1, 2), -7.0, dtype=torch.float32, device="cuda"); d = dev(z["up_small"])
    mine.baoCudaFlowBilteralUpsampling(P(o), P(ia), w1, h1, pitch, P(d), w2, h2, 2.0); torch.cuda.synchronize()
    assert same_bits(o.cpu().numpy(), z["up_out"])
    if "pf_nnf" in z:   # plane-fitting PatchMatch on the coarsest-level planes this library prepares from the same seeded pair
        h, w = int(z["h"]), int(z["w"])
        a, b, _, _ = synth.make_pair(h, w, int(z["pair_idx"]), scale_to=float(z["scale_to"]))
        ctx = E.EppmContext(h, w, 1)
        ctx.stage_prepare(dev(a[None]), dev(b[None]), 1)
        hc, wc = ctx.level_dims(2)
        planes = [refharness.pitched(ctx.read_plane(k, 2)) for k in (E.PLANE_RGBA1, E.PLANE_RGBA2, E.PLANE_CENSUS1, E.PLANE_CENSUS2)]
        mine.baoCudaPatchMatch_PlaneFitting.argtypes = [V] * 6 + [I, I, S, S, S, S]
        pn = torch.zeros((hc, wc, 2), dtype=torch.int16, device="cuda"); pc = torch.zeros((hc, wc), dtype=torch.float32, device="cuda")
        mine.baoCudaPatchMatch_PlaneFitting(P(pn), P(pc), P(planes[0][0]), P(planes[1][0]), P(planes[2][0]), P(planes[3][0]), wc, hc, planes[0][1],
                                            wc * 4, wc * 4, planes[2][1]); torch.cuda.synchronize()
        assert np.array_equal(pn.cpu().numpy(), z["pf_nnf"]) and same_bits(pc.cpu().numpy(), z["pf_cost"])
        ctx.close()


@needs_ref
def test_flow_smoothing_bit_exact_single_warp(ref, mine):
    """A 16x2 image is one warp of the reference's kernel: lock-step execution = all reads before all writes, so its in-place
    filter is race-free there and must equal the snapshot filter bit for bit (exercises every |dx| <= 10 weight, |dy| <= 1)."""
    rng = np.random.default_rng(5)
    for trial in range(8):
        h, w = 2, 16
        img = np.zeros((h, w, 4), np.uint8); img[..., :3] = rng.integers(100, 110, (h, w, 3))
        fl = rng.normal(0, 3, (h, w, 2)).astype(np.float32)
        if trial % 2:
            fl[rng.random((h, w)) < 0.2] = 1e10
        ib, pitch = refharness.pitched(img)
        outs = []
        for lib in (ref.lib, mine):
            t = dev(fl)
            lib.baoCudaFlowSmoothing(P(t), P(ib), w, h, pitch, w * 8)
            torch.cuda.synchronize()
            outs.append(t.cpu().numpy())
        assert same_bits(outs[0], outs[1])


# ------------------------------------------------------------------------------------------------ end to end
@needs_ref
def test_end_to_end_shipped_pair(ref):
    """BASELINE config 1: frame10/frame11.ppm at default parameters; mean end-point difference to the reference's flow."""
    p = os.path.join(refharness.REF_DATA, "frame10.ppm")
    if not os.path.exists(p):
        pytest.skip("shipped pair not staged")
    a = synth.read_ppm(p); b = synth.read_ppm(os.path.join(refharness.REF_DATA, "frame11.ppm"))
    h, w = a.shape[:2]
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    fr = ref.compute_flow(rc, h, w)
    ctx = E.EppmContext(h, w, 1)
    fm = ctx.compute_batch_host(a[None], b[None])[0]
    d = np.sqrt(((fm - fr) ** 2).sum(-1))
    assert d.mean() <= 0.05, d.mean()  # north-star tolerance: <= 0.05 px mean EPE difference
    ref.destroy(rc); ctx.close()


@needs_ref
def test_end_to_end_epe_vs_ground_truth_no_worse(ref):
    """North star: mean end-point error against synthetic ground truth within 0.05 px of the reference's, and no worse beyond that.  This
    library is deterministic; the reference is not -- its three in-place filters race, and its own EPE on these pairs moves by up to 0.03 px
    from run to run on one GPU (bimodal on pair 1: 3.381 / 3.409 px) and by 0.008 px between boxes (tools/ref_epe_spread.py ->
    profiles/r02_ref_epe_spread.json).  A single run of it therefore cannot carry a 0.05 px bar per pair (pair 2 sits at +0.044 ... +0.052
    depending on the box).  The bar is applied to the MEAN over the three pairs against the reference's mean over three runs each (measured
    +0.032 px); every single pair stays within 0.05 px plus that measured 0.03 px of reference noise."""
    deltas = []
    for h, w, idx in [(480, 640, 0), (436, 1024, 1), (436, 1024, 2)]:
        a, b, gt, valid = synth.make_pair(h, w, idx)
        e_ref = []
        for _ in range(3):
            rc = ref.create(h, w)
            ref.set_data(rc, a, b)
            e_ref.append(synth.epe(ref.compute_flow(rc, h, w), gt, valid))
            ref.destroy(rc)
        ctx = E.EppmContext(h, w, 1)
        e_me = synth.epe(ctx.compute_batch_host(a[None], b[None])[0], gt, valid)
        ctx.close()
        deltas.append(e_me - float(np.mean(e_ref)))
        assert abs(deltas[-1]) <= 0.05 + 0.03, (h, w, idx, e_me, e_ref)
    assert abs(float(np.mean(deltas))) <= 0.05, deltas


def _c2f_call(lib, flows_lp1, out, l, dims, img, cen):
    nl = 3
    PtrArr, IntArr, SzArr = C.c_void_p * nl, C.c_int * nl, C.c_size_t * nl
    flows = [None] * nl
    flows[l] = out; flows[l + 1] = flows_lp1
    lib.baoCudaBLF_C2F.argtypes = [C.c_void_p] * 11 + [C.c_int]
    lib.baoCudaBLF_C2F.restype = None
    lib.baoCudaBLF_C2F(PtrArr(*[P(x) if x is not None else None for x in flows]), PtrArr(*[P(img[0][k][0]) for k in range(nl)]),
                       PtrArr(*[P(img[1][k][0]) for k in range(nl)]), PtrArr(*[P(cen[0][k][0]) for k in range(nl)]),
                       PtrArr(*[P(cen[1][k][0]) for k in range(nl)]), None, None, IntArr(*[d[0] for d in dims]), IntArr(*[d[1] for d in dims]),
                       SzArr(*[img[0][k][1] for k in range(nl)]), SzArr(*[cen[0][k][1] for k in range(nl)]), l)
    torch.cuda.synchronize()


@needs_ref
def test_full_hd_vs_reference(ref, mine):
    """The benchmark resolution (BASELINE config 3, 1920x1080) against the reference build: pyramid / census / geometry, the whole
    PatchMatch in both directions and the plane-fitting refine of levels 1 and 0 bit-exact; end to end the direct flow-vs-flow
    difference is bounded at ~1.2x what was measured (profiles/r02_parity_e2e.json) and the EPE against ground truth within 0.05 px."""
    h, w = 1080, 1920
    a, b, gt, valid = synth.make_pair(h, w, 1000)
    rc, dims, img, cen = _ref_level_planes(ref, h, w, a, b)
    ctx = E.EppmContext(h, w, 1)
    ctx.stage_prepare(dev(a[None]), dev(b[None]), 1)
    for l in range(3):
        assert ctx.level_dims(l) == tuple(dims[l])
        for which in (E.PLANE_RGBA1, E.PLANE_RGBA2, E.PLANE_CENSUS1, E.PLANE_CENSUS2):
            assert same_bits(ctx.read_plane(which, l), ref.read_plane(rc, which, l)), (which, l)
    hc, wc = dims[2]
    for swap in (False, True):
        (nr, cr), (nm, cm) = _pm(ref.lib, img, cen, 2, wc, hc, swap), _pm(mine, img, cen, 2, wc, hc, swap)
        assert torch.equal(nr, nm), f"{(nr != nm).any(-1).sum().item()} targets differ (swap={swap})"
        assert same_bits(cr.cpu().numpy(), cm.cpu().numpy())
    # the reference's own flow pyramid (after its consistency / smoothing stages) feeds the refine of both libraries
    fr = ref.compute_flow(rc, h, w)
    for l in (1, 0):
        coarse = torch.from_numpy(ref.read_plane(rc, 8, l + 1)).cuda()
        outs = []
        for lib in (ref.lib, mine):
            fine = torch.zeros((dims[l][0], dims[l][1], 2), dtype=torch.float32, device="cuda")
            _c2f_call(lib, coarse, fine, l, dims, img, cen)
            outs.append(fine.cpu().numpy())
        assert same_bits(outs[0], outs[1]), f"plane-fitting refine differs at level {l}: {(outs[0].view(np.uint32) != outs[1].view(np.uint32)).sum()} floats"
    fm = ctx.compute_batch_host(a[None], b[None])[0]
    d = np.sqrt(((fm.astype(np.float64) - fr.astype(np.float64)) ** 2).sum(-1))
    e_ref, e_me = synth.epe(fr, gt, valid), synth.epe(fm, gt, valid)
    assert abs(e_me - e_ref) <= 0.05, (e_me, e_ref)
    assert np.median(d) <= 1e-3, np.median(d)
    assert d.mean() <= FLOW_VS_REF_BOUND[(h, w)] * (3.0 if RACY_SLACK > 1 else 1.0), d.mean()
    ref.destroy(rc); ctx.close()


# direct flow-vs-flow mean end-point difference to the reference build on synthetic large-displacement pairs: ~1.2x the values measured on
# B200 (tools/parity_e2e.py -> profiles/r02_parity_e2e.json).  The difference comes from the three filters the reference updates in place
# (DESIGN.md §3); it is recorded here so that a regression shows.
# Under compute-sanitizer the REFERENCE's in-place filters resolve their read-while-write races differently (its kernels run an order of magnitude
# slower and in another interleaving: measured 0.29 % / 0.74 % outlier-stage mismatches instead of 0.09 %), so the bounds on the racy stages are
# relaxed there; tools/gpu_sanitize.sh sets the variable.
RACY_SLACK = 10.0 if os.environ.get("EPPM_UNDER_SANITIZER") else 1.0
FLOW_VS_REF_BOUND = {(480, 640): 0.34, (436, 1024): 0.40, (1080, 1920): 0.36}   # measured 0.277, 0.330, 0.297 px (median 0: 3 % of the pixels carry it)


@needs_ref
@pytest.mark.parametrize("h,w,idx", [(480, 640, 0), (436, 1024, 1)])
def test_flow_vs_reference_flow_is_recorded(ref, h, w, idx):
    a, b, gt, valid = synth.make_pair(h, w, idx)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    fr = ref.compute_flow(rc, h, w)
    ctx = E.EppmContext(h, w, 1)
    fm = ctx.compute_batch_host(a[None], b[None])[0]
    d = np.sqrt(((fm.astype(np.float64) - fr.astype(np.float64)) ** 2).sum(-1))
    assert np.median(d) <= 1e-3, np.median(d)
    assert d.mean() <= FLOW_VS_REF_BOUND[(h, w)] * (3.0 if RACY_SLACK > 1 else 1.0), d.mean()
    ref.destroy(rc); ctx.close()


@needs_ref
def test_inplace_filter_mode(ref):
    """inplace_filters = 1 (EPPM_INPLACE_LEGACY): outlier removal, weighted median and smoothing update in place with the reference's launch
    geometry.  Same accuracy bar as the default mode; the deterministic stages in front of them are untouched (PatchMatch planes identical)."""
    h, w = 480, 640
    a, b, gt, valid = synth.make_pair(h, w, 0)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    fr = ref.compute_flow(rc, h, w)
    p = E.default_params()
    p.inplace_filters = 1
    ctx = E.EppmContext(h, w, 1, params=p)
    ctx0 = E.EppmContext(h, w, 1)
    fm = ctx.compute_batch_host(a[None], b[None])[0]
    f0 = ctx0.compute_batch_host(a[None], b[None])[0]
    assert same_bits(ctx.read_plane(E.PLANE_NNF_BWD), ctx0.read_plane(E.PLANE_NNF_BWD))
    assert np.isfinite(fm).all()
    assert abs(synth.epe(fm, gt, valid) - synth.epe(fr, gt, valid)) <= 0.05
    d = np.sqrt(((fm.astype(np.float64) - fr.astype(np.float64)) ** 2).sum(-1))
    assert np.median(d) <= 1e-3 and d.mean() <= 0.40 * (3.0 if RACY_SLACK > 1 else 1.0)   # measured 0.309 px: the in-place order does not land closer to the reference (profiles/r02_parity_e2e.json)
    assert not same_bits(fm, f0)   # it really is a different update order
    ref.destroy(rc); ctx.close(); ctx0.close()


@needs_ref
def test_video_stream_vs_reference_on_chained_frames(ref):
    """eppm_compute_stream_* against the REFERENCE on a chained synthetic clip (frame t+1 = frame t moved): the reference is given every
    consecutive pair through set_data + compute_flow; the stream entry points prepare each frame once.  Pyramid and census of every frame
    bit-exact with the reference's, flows within the end-point bars, host-buffer entry identical to the device entry."""
    h, w = 270, 480
    frames, flows, valids = synth.make_stream(h, w, 4, first_idx=70, scale_to=0.25)
    n = frames.shape[0] - 1
    ctx = E.EppmContext(h, w, 3)   # max_batch 3 < 4 frames: the host entry has to chunk (2 pairs + 1 pair)
    out_h = ctx.compute_stream_host(frames)
    ctx4 = E.EppmContext(h, w, 4)
    d_flow = torch.zeros((n, h, w, 2), dtype=torch.float32, device="cuda")
    ctx4.compute_stream_device(dev(frames), n, d_flow)
    ctx4.synchronize()
    assert same_bits(out_h, d_flow.cpu().numpy())
    rc = ref.create(h, w)
    deltas = []
    for t in range(n):
        ref.set_data(rc, frames[t], frames[t + 1])
        for l in range(3):   # frame t sits in plane t of the image-1 arrays of the stream context
            assert same_bits(ctx4.read_plane(E.PLANE_RGBA1, l, pair=t), ref.read_plane(rc, 0, l)), (t, l)
            assert same_bits(ctx4.read_plane(E.PLANE_CENSUS1, l, pair=t), ref.read_plane(rc, 2, l)), (t, l)
        fr = ref.compute_flow(rc, h, w)
        d = np.sqrt(((out_h[t].astype(np.float64) - fr.astype(np.float64)) ** 2).sum(-1))
        deltas.append(synth.epe(out_h[t], flows[t], valids[t]) - synth.epe(fr, flows[t], valids[t]))
        # no worse than the reference against ground truth: 0.05 px on the MEAN over the clip (below); a single pair additionally gets the 0.03 px the
        # reference's own result moves by between runs and boxes (its in-place filters race: profiles/r02_ref_epe_spread.json).  Measured on this
        # clip: +0.045, +0.008, -0.052 (tools/test_margins.py)
        assert deltas[-1] <= (0.05 + 0.03) * (3.0 if RACY_SLACK > 1 else 1.0), (t, deltas)
        assert np.median(d) <= 1e-3, (t, np.median(d))
    assert float(np.mean(deltas)) <= 0.05 * (3.0 if RACY_SLACK > 1 else 1.0), deltas
    ref.destroy(rc); ctx.close(); ctx4.close()


@needs_ref
def test_device_flow_evaluation_and_flo_vs_reference(ref, tmp_path):
    """eppm_eval_flow against the reference's bao_calc_flow_error / bao_calc_flow_error_percentage (basic/bao_flow_tools.cpp:64-141) and
    eppm_write_flo against its .flo writer, on a computed flow with unknown (1e10) and zero ground-truth pixels mixed in."""
    h, w = 240, 320
    a, b, gt, valid = synth.make_batch(h, w, 2, first_idx=30, distinct=2)
    ctx = E.EppmContext(h, w, 2)
    d_flow = torch.zeros((2, h, w, 2), dtype=torch.float32, device="cuda")
    ctx.compute_batch_device(dev(a), dev(b), 2, d_flow)
    ctx.synchronize()
    gt = gt.copy()
    gt[0, 10:40, 20:90] = 1e10          # unknown ground truth
    gt[1, 100:140, :] = 0.0             # zero motion: known, but excluded from EPE / AAE by the reference's test
    gt[1, 5, 5] = (1e10, 2.5)           # one component unknown: the reference still counts the pixel
    fl = d_flow.cpu().numpy()
    # a second, deliberately wrong flow: the reference's angular error is NaN as soon as ONE pixel's cosine rounds above 1 (it calls acos on
    # (u.g + 1) / (|u||g|) in single precision, :81-82), which an accurate flow always hits -- the device reduction reproduces that; the
    # perturbed flow keeps every cosine below 1 so that the value itself is compared too
    rng = np.random.default_rng(4)
    noisy = (fl + rng.normal(0, 3.0, fl.shape)).astype(np.float32)
    n_nan = 0
    for field in (fl, noisy):
        d_f = dev(field)
        for border, thr in ((0, 3), (7, 1)):
            res = ctx.eval_flow(d_f, dev(gt), 2, border=border, outlier_thresh=float(thr))
            for i in range(2):
                e, aae, out = ref.calc_flow_error(field[i], gt[i], border, thr)
                # the reference accumulates ~7e4 floats sequentially in single precision: relative error up to ~1e-4
                assert abs(res[i]["epe"] - e) <= 2e-4 * max(1.0, abs(e)), (res[i], e)
                if np.isnan(aae):
                    n_nan += 1
                    assert np.isnan(res[i]["aae_deg"]), (res[i], aae)
                else:
                    assert abs(res[i]["aae_deg"] - aae) <= 2e-4 * max(1.0, abs(aae)), (res[i], aae)
                assert abs(res[i]["outlier_frac"] - out) <= 1e-6, (res[i], out)
                assert 0 < res[i]["n_valid"] < h * w and res[i]["n_known"] <= h * w
    assert n_nan < 8   # at least the perturbed flow produced numbers
    # .flo: byte-identical to the file the reference writes, and it reads back
    p_me, p_ref = tmp_path / "me.flo", tmp_path / "ref.flo"
    E.write_flo(p_me, fl[0]); ref.save_flo(p_ref, fl[0])
    assert p_me.read_bytes() == p_ref.read_bytes()
    assert same_bits(E.read_flo(p_me), fl[0])
    ctx.close()


# ------------------------------------------------------------------------------------------------ fixtures + CPU oracle
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ref_*.npz"))))
def test_against_committed_reference_fixtures(path):
    z = np.load(path)
    h, w = int(z["h"]), int(z["w"])
    a, b, _, _ = synth.make_pair(h, w, int(z["pair_idx"]), scale_to=float(z["scale_to"]))
    ctx = E.EppmContext(h, w, 1)
    ctx.stage_prepare(dev(a[None]), dev(b[None]), 1)
    for l in range(3):
        for which, nm in ((0, "rgba1"), (1, "rgba2"), (2, "census1"), (3, "census2")):
            assert same_bits(ctx.read_plane(which, l), z[f"{nm}_L{l}"]), (nm, l)
    ctx.stage_patchmatch_partial(1)
    assert same_bits(ctx.read_plane(E.PLANE_NNF_FWD), z["rand_field"])
    assert same_bits(ctx.read_plane(E.PLANE_COST_FWD), z["cost_init_fwd"])
    ctx.stage_patchmatch_partial(2)
    assert same_bits(ctx.read_plane(E.PLANE_NNF_FWD), z["nnf_after_rowfwd"])
    ctx.stage_patchmatch()
    assert same_bits(ctx.read_plane(E.PLANE_NNF_FWD), z["nnf_pm_fwd"])
    assert same_bits(ctx.read_plane(E.PLANE_NNF_BWD), z["nnf_pm_bwd"])
    assert same_bits(ctx.read_plane(E.PLANE_COST_FWD), z["cost_pm_fwd"])
    ctx.close()


def test_gpu_vs_cpu_oracle_small():
    import golden
    h, w = 96, 128
    a, b, gt, valid = synth.make_pair(h, w, 7, scale_to=0.12)
    g = golden.Golden(h, w)
    fg = g.compute(a, b)
    ctx = E.EppmContext(h, w, 1)
    fm = ctx.compute_batch_host(a[None], b[None])[0]
    for l in range(3):
        assert same_bits(ctx.read_plane(E.PLANE_RGBA1, l), g.plane("rgba1", l))
        assert same_bits(ctx.read_plane(E.PLANE_CENSUS2, l), g.plane("census2", l))
    d = np.sqrt(((fm - fg) ** 2).sum(-1))
    assert d.mean() <= 0.05  # the oracle cannot reproduce MUFU.EX2: near-ties may flip, EPE tolerance of the north star
    ctx.close()


# ------------------------------------------------------------------------------------------------ API behaviour
def test_batch_equals_single_pairs_and_is_deterministic():
    h, w = 192, 256
    i1, i2, _, _ = synth.make_batch(h, w, 3, first_idx=20)
    ctx3 = E.EppmContext(h, w, 3)
    f3 = ctx3.compute_batch_host(i1, i2)
    f3b = ctx3.compute_batch_host(i1, i2)
    assert same_bits(f3, f3b)
    ctx1 = E.EppmContext(h, w, 1)
    for k in range(3):
        assert same_bits(ctx1.compute_batch_host(i1[k:k + 1], i2[k:k + 1])[0], f3[k])
    ctx1.close(); ctx3.close()


def test_host_api_chunks_batches_larger_than_max_batch():
    h, w = 192, 256
    i1, i2, _, _ = synth.make_batch(h, w, 5, first_idx=40)
    ctx2 = E.EppmContext(h, w, 2)   # 5 pairs through a context of 2: chunks 2+2+1, double buffered
    ctx5 = E.EppmContext(h, w, 5)
    assert same_bits(ctx2.compute_batch_host(i1, i2), ctx5.compute_batch_host(i1, i2))
    ctx2.close(); ctx5.close()


def test_device_api_equals_host_api_and_class_mirror():
    h, w = 192, 256
    i1, i2, _, _ = synth.make_batch(h, w, 2, first_idx=30)
    ctx = E.EppmContext(h, w, 2)
    fh = ctx.compute_batch_host(i1, i2)
    d_out = torch.zeros((2, h, w, 2), dtype=torch.float32, device="cuda")
    ctx.compute_batch_device(dev(i1), dev(i2), 2, d_out)
    ctx.synchronize()
    assert same_bits(d_out.cpu().numpy(), fh)
    m = E.BaoFlowPatchmatchMultiscaleCuda()
    m.init(i1[0], i2[0], h, w)
    u, v = m.compute_flow()
    assert same_bits(u, fh[0, ..., 0]) and same_bits(v, fh[0, ..., 1])
    assert ctx.launch_count() > 0
    ctx.close()


def test_error_behaviour():
    with pytest.raises(E.EppmError):
        E.EppmContext(2, 2, 1)  # coarsest pyramid level would be empty
    ctx = E.EppmContext(96, 128, 1)
    with pytest.raises(E.EppmError):
        ctx.compute_batch_device(dev(np.zeros((2, 96, 128, 3), np.uint8)), dev(np.zeros((2, 96, 128, 3), np.uint8)), 2,
                                 torch.zeros((2, 96, 128, 2), device="cuda"))  # device API: batch > max_batch
    with pytest.raises(E.EppmError):
        E.EppmContext(96, 128, 1).stage_patchmatch()  # before prepare
    ctx.close()


def test_properties_full_hd():
    """Size-independent properties at the benchmark size: identical frames -> zero flow; pure translation recovered; determinism."""
    h, w = 1080, 1920
    a, _, _, _ = synth.make_pair(h, w, 5)
    ctx = E.EppmContext(h, w, 2)
    shifted = np.roll(a, (6, -11), (0, 1))  # content moves by (+6 rows, -11 cols): flow u = -11, v = +6
    f = ctx.compute_batch_host(np.stack([a, a]), np.stack([a, shifted]))
    assert np.abs(f[0]).mean() <= 0.02
    inner = f[1, 40:-40, 40:-40]
    assert np.median(np.abs(inner[..., 0] + 11)) <= 0.05 and np.median(np.abs(inner[..., 1] - 6)) <= 0.05
    f2 = ctx.compute_batch_host(np.stack([a, a]), np.stack([a, shifted]))
    assert same_bits(f, f2)
    ctx.close()


def test_constant_division_fast_path_is_exact(mine):
    """The smoothing kernel's 3-instruction division by -(0.02f*0.02f) and the patch cost's by -(0.1f*0.1f) must equal div.rn for
    EVERY float in the operand range (exhaustive sweep over the bit patterns of [2^-20, 2))."""
    lo = np.array([2.0 ** -20], np.float32).view(np.uint32)[0]
    hi = np.array([2.0], np.float32).view(np.uint32)[0]
    for sig in (0.02, 0.1):
        d = -(np.float32(sig) * np.float32(sig))
        assert mine.eppm_selftest_const_div(float(d), int(lo), int(hi)) == 0
    ctx = E.EppmContext(96, 128, 1)
    assert ctx.lib.eppm_smooth_uses_fast_div(ctx._ctx) == 1
    ctx.close()


@needs_ref
def test_unmodified_main_cpp_drop_in(tmp_path):
    """The reference's UNMODIFIED main.cpp, compiled against include/compat and linked with libeppm_b200.so (build/runeppm_b200),
    against the reference's own executable on the shipped pair (BASELINE config 1)."""
    import shutil, subprocess
    exe = os.path.join(ROOT, "build", "runeppm_b200")
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "runeppm")
    if not (os.path.exists(exe) and os.path.exists(ref_exe) and os.path.exists(os.path.join(refharness.REF_DATA, "frame10.ppm"))):
        pytest.skip("drop-in executables not built")
    flows = []
    for k, binary in enumerate((ref_exe, exe)):
        d = tmp_path / f"run{k}"
        d.mkdir()
        for f in ("frame10.ppm", "frame11.ppm"):
            shutil.copy(os.path.join(refharness.REF_DATA, f), d / f)
        subprocess.run([binary], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=300)
        flows.append(synth.read_flo(str(d / "flow.flo")))
    assert flows[0].shape == flows[1].shape == (480, 640, 2)
    dd = np.sqrt(((flows[0] - flows[1]) ** 2).sum(-1))
    assert dd.mean() <= 0.05, dd.mean()


def test_reference_host_class_on_this_library(tmp_path):
    """the reference's UNMODIFIED main.cpp AND host class (bao_flow_patchmatch_multiscale_cuda.cpp, basic/*,
    middlebury/*; objects compiled from the sources where they lie) linked against libeppm_b200.so instead of the reference's three .cu
    files (oracle/_ref/runeppm_hostclass, `make -C oracle hostclass`): every stage goes through include/eppm_legacy_abi.h on the
    reference's own pitched buffers.  Compared with the reference's own executable on the shipped pair."""
    import shutil, subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "runeppm_hostclass")
    ref_exe = os.path.join(ROOT, "oracle", "_ref", "runeppm")
    if not (os.path.exists(exe) and os.path.exists(ref_exe) and os.path.exists(os.path.join(refharness.REF_DATA, "frame10.ppm"))):
        pytest.skip("reference host objects not built")
    flows = []
    for k, binary in enumerate((ref_exe, exe)):
        d = tmp_path / f"run{k}"
        d.mkdir()
        for f in ("frame10.ppm", "frame11.ppm"):
            shutil.copy(os.path.join(refharness.REF_DATA, f), d / f)
        subprocess.run([binary], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=300)
        flows.append(synth.read_flo(str(d / "flow.flo")))
    assert flows[0].shape == flows[1].shape == (480, 640, 2)
    dd = np.sqrt(((flows[0] - flows[1]) ** 2).sum(-1))
    assert dd.mean() <= 0.05, dd.mean()


@needs_ref
def test_flow_colour_coding_vs_reference(ref, mine):
    """bao_cuda_convert_flow_to_colorshow (C++ linkage; compute_flow's optional colour output): the same libdevice atan2f and the same
    float -> double -> int chain as the reference (read from its SASS) -> every byte identical."""
    sym = "_Z34bao_cuda_convert_flow_to_colorshowP6uchar4P6float2iiff"
    h, w = 120, 160
    g = torch.Generator(device="cpu").manual_seed(3)
    fl = torch.randn((h, w, 2), generator=g) * 12
    fl[:10, :20] = 1e10                         # unknown flow -> black
    fl[20, :] = 0.0
    fl = fl.cuda()
    outs = []
    for lib in (ref.lib, mine):
        fn = getattr(lib, sym)
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float]; fn.restype = None
        out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        fn(P(out), P(fl), h, w, 20.0, 20.0)
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy()[..., :3].astype(np.int32))
    d = np.abs(outs[0] - outs[1])
    assert d.max() == 0, (d.max(), (d != 0).sum())
    assert (outs[0][:10, :20] == 0).all() and outs[0].std() > 20


@needs_ref
def test_philox_mode_epe_delta(ref):
    """Stated counter-based RNG (EPPM_RNG_PHILOX, Philox4x32-10 keyed by seed with one sub-sequence per coarse pixel): the NNF is a
    different random realisation, so the bar is the north star's: MEAN accuracy against ground truth within 0.05 px of the
    reference's.  Averaged over three pairs: a single 436x1024 pair moves by +-0.05 px with the realisation alone (and the
    reference's own result depends on how its in-place races resolve on the box at hand)."""
    h, w = 436, 1024
    e_ref, e_me = [], []
    rc = ref.create(h, w)
    p = E.default_params()
    p.rng_mode = 1
    ctx = E.EppmContext(h, w, 1, params=p)
    for idx in (1, 2, 3):
        a, b, gt, valid = synth.make_pair(h, w, idx)
        ref.set_data(rc, a, b)
        fr = ref.compute_flow(rc, h, w)
        fm = ctx.compute_batch_host(a[None], b[None])[0]
        e_ref.append(synth.epe(fr, gt, valid)); e_me.append(synth.epe(fm, gt, valid))
    # measured +0.035 px (tools/test_margins2.py); the reference's own mean over these pairs moves by ~0.01 px between runs and boxes (pair 1 is bimodal:
    # 3.381 / 3.409, profiles/r02_ref_epe_spread.json), which the bar has to absorb on top of the north star's 0.05
    assert np.mean(e_me) <= np.mean(e_ref) + 0.05 + 0.01, (e_me, e_ref)
    ref.destroy(rc); ctx.close()


def test_subpixel_final_option(mine):
    """eppm_params::subpixel_final: the reference's optional sub-pixel stage (baoCudaCensusTransform_Bicubic + baoCudaSubpixRefine; declared by its
    host class, never placed in its pipeline) between the level-0 refine and the level-0 smoothing passes.  The wiring is checked exactly: the
    stage functions applied by hand to the default context's level-0 refine output (integer flow, left in FLOW_TMP) must reproduce the option's
    FLOW_TMP bit for bit; end to end the option must stay as accurate against ground truth as the default."""
    h, w = 240, 320
    a, b, gt, valid = synth.make_pair(h, w, 21, scale_to=0.2)
    ctx0 = E.EppmContext(h, w, 1)
    f0 = ctx0.compute_batch_host(a[None], b[None])[0].copy()
    t0 = ctx0.read_plane(E.PLANE_FLOW_TMP, 0)
    assert np.array_equal(t0, np.round(t0))                     # the refine leaves integer displacements
    p = E.default_params()
    p.subpixel_final = 1
    ctx1 = E.EppmContext(h, w, 1, params=p)
    f1 = ctx1.compute_batch_host(a[None], b[None])[0].copy()
    t1 = ctx1.read_plane(E.PLANE_FLOW_TMP, 0)
    assert np.isfinite(f1).all() and not np.array_equal(t1, t0) and not same_bits(f1, f0)
    assert np.abs(t1 - t0).max() <= 1.5 + 1e-6                   # offsets of at most 3 half-pixels (:630)
    # by hand, through the stage ABI, on the default context's planes
    S, I, V = C.c_size_t, C.c_int, C.c_void_p
    mine.baoCudaCensusTransform_Bicubic.argtypes = [V, V, I, I, S, V, V, I, I, S]; mine.baoCudaCensusTransform_Bicubic.restype = None
    mine.baoCudaFlow2NNF.argtypes = [V, V, I, I, S, S]; mine.baoCudaFlow2NNF.restype = None
    mine.baoCudaSubpixRefine.argtypes = [V, V, V, V, V, V, I, I, S, S, S, S]; mine.baoCudaSubpixRefine.restype = None
    i1, i2 = dev(ctx0.read_plane(E.PLANE_RGBA1, 0)), dev(ctx0.read_plane(E.PLANE_RGBA2, 0))
    c1 = torch.zeros((2 * h, 2 * w), dtype=torch.uint8, device="cuda"); c2 = torch.zeros_like(c1)
    mine.baoCudaCensusTransform_Bicubic(P(c1), P(c2), 2 * w, 2 * h, 2 * w, P(i1), P(i2), w, h, w * 4)
    fl = dev(t0); nn = torch.zeros((h, w, 2), dtype=torch.int16, device="cuda")
    mine.baoCudaFlow2NNF(P(nn), P(fl), w, h, w * 4, w * 8)
    mine.baoCudaSubpixRefine(P(fl), P(nn), P(i1), P(i2), P(c1), P(c2), w, h, w * 4, 2 * w, w * 4, w * 8)
    torch.cuda.synchronize()
    assert same_bits(fl.cpu().numpy(), t1)
    e0, e1 = synth.epe(f0, gt, valid), synth.epe(f1, gt, valid)
    assert e1 <= e0 + 0.05, (e0, e1)
    with pytest.raises(E.EppmError):                             # the texture pitch needs w % 8 == 0
        E.EppmContext(h, 324, 1, params=p)
    ctx0.close(); ctx1.close()


def test_variant_switches_compute_the_same_bits():
    """Every tuned kernel has its plain predecessor behind EPPM_VARIANT (eppm_internal.h): site-table vs computed coordinates in the
    refine, grouped vs per-sample __expf fix-up, joint vs serial random search, work-queue vs CTA-local propagation with and without
    skipping / compaction, four-row packed vs two-row smoothing, three- vs nine-warp refine, texture vs LSU gathers and one vs three
    passes in the random search, scalar vs packed-pair refine, warp-per-evaluation vs thread-per-evaluation scoring of the propagation
    queue and of the random search, barrier-free segment chains, the refine with its fix-up-free loop switched off and with a shared AD +
    census volume, the census kernel without its TMA-staged tile.  All of them must produce the flow of the default build bit for bit."""
    h, w = 270, 480
    a, b, _, _ = synth.make_batch(h, w, 2, first_idx=11, distinct=2)
    flows = {}
    try:
        for v in (0, 1, 2, 4, 8, 16, 32, 64, 127, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 131072, 262144, 524288, 524288 + 8192, 1048576, 2097152, 4194304, 8388608, 16777216):
            os.environ["EPPM_VARIANT"] = str(v)
            ctx = E.EppmContext(h, w, 2)
            if v == 0:
                assert ctx.lib.eppm_refine_uses_site_table(ctx._ctx, 0) == 1
            flows[v] = ctx.compute_batch_host(a, b).copy()
            ctx.close()
    finally:
        os.environ.pop("EPPM_VARIANT", None)
    for v, f in flows.items():
        assert np.array_equal(f.view(np.uint32), flows[0].view(np.uint32)), f"EPPM_VARIANT={v} differs from the default kernels"


def test_refine_fixup_free_loop_is_exact_on_adversarial_inputs():
    """The default refine kernel scores the first patch row with the exact __expf fix-up and drops the fix-up test for the other rows when every
    accumulator is already >= 2^-99 (a weight that needs the fix-up is < 2^-126 and is absorbed unchanged by such a sum); warps that fail the test
    keep the exact loop.  Inputs chosen to hit both sides of that test: identical frames (cost sums exactly 0: fallback), a saturated black /
    white pattern (range weights underflow everywhere: many fix-ups, tiny weight sums), and a textured pair with saturated rectangles pasted in.
    EPPM_VARIANT = 4194304 forces the exact loop everywhere: same bits."""
    h, w = 270, 480
    a, b, _, _ = synth.make_batch(h, w, 3, first_idx=5, distinct=3)
    a, b = a.copy(), b.copy()
    b[0] = a[0]                                                                # static scene
    yy, xx = np.mgrid[0:h, 0:w]
    pat = (((xx // 7 + yy // 5) % 2) * 255).astype(np.uint8)                   # saturated checkerboard, shifted by (3, 2) in frame 2
    a[1] = pat[..., None]; b[1] = np.roll(pat, (2, 3), (0, 1))[..., None]
    for k, (y0, x0) in enumerate([(40, 60), (150, 300), (200, 100)]):          # saturated patches inside a textured pair
        a[2, y0:y0 + 30, x0:x0 + 50] = 255 * (k % 2); b[2, y0 + 1:y0 + 31, x0 + 2:x0 + 52] = 255 * (k % 2)
    flows = {}
    try:
        for v in (0, 4194304):
            os.environ["EPPM_VARIANT"] = str(v)
            ctx = E.EppmContext(h, w, 3)
            flows[v] = ctx.compute_batch_host(a, b).copy()
            ctx.close()
    finally:
        os.environ.pop("EPPM_VARIANT", None)
    assert np.isfinite(flows[0]).all()
    assert np.array_equal(flows[0].view(np.uint32), flows[4194304].view(np.uint32))


@pytest.mark.parametrize("name,depth,iters", [("d2i5", 2, 5), ("d4i2", 4, 2)])
def test_non_default_depth_and_iterations_vs_reference_variants(name, depth, iters):
    """BASELINE config 5 sweeps pyramid depth x propagation iterations.  The reference fixes both as macros, so oracle/Makefile
    rebuilds it per sweep point (`make variants`); here they are run-time parameters.  Pyramid/census and the whole PatchMatch
    must stay bit-exact, the final flow as accurate against ground truth."""
    path = refharness.variant_lib(name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built")
    ref = refharness.Ref(path)
    h, w = 480, 640
    a, b, gt, valid = synth.make_pair(h, w, 4)
    rc = ref.create(h, w)
    assert ref.num_levels(rc) == depth
    ref.set_data(rc, a, b)
    p = E.default_params()
    p.pyr_levels, p.num_iter = depth, iters
    ctx = E.EppmContext(h, w, 1, params=p)
    ctx.stage_prepare(dev(a[None]), dev(b[None]), 1)
    planes = {}
    for l in range(depth):
        assert ctx.level_dims(l) == ref.level_dims(rc, l)
        for which in (0, 1, 2, 3):
            planes[(which, l)] = ref.read_plane(rc, which, l)
            assert same_bits(ctx.read_plane(which, l), planes[(which, l)]), (which, l)
    L = depth - 1
    hc, wc = ctx.level_dims(L)
    i1, i2 = refharness.pitched(planes[(0, L)]), refharness.pitched(planes[(1, L)])
    c1, c2 = refharness.pitched(planes[(2, L)]), refharness.pitched(planes[(3, L)])
    nf, cf = ref.tap_patchmatch(i1, i2, c1, c2, wc, hc, 1000)
    nb, cb = ref.tap_patchmatch(i2, i1, c2, c1, wc, hc, 1000)
    ctx.stage_patchmatch()
    assert same_bits(ctx.read_plane(E.PLANE_NNF_FWD), nf) and same_bits(ctx.read_plane(E.PLANE_NNF_BWD), nb)
    assert same_bits(ctx.read_plane(E.PLANE_COST_FWD), cf) and same_bits(ctx.read_plane(E.PLANE_COST_BWD), cb)
    fr = ref.compute_flow(rc, h, w)
    fm = ctx.compute_batch_host(a[None], b[None])[0]
    # with a 60x80 coarsest level and 2 iterations most coarse pixels are re-filled by the (racy, in the reference) weighted median;
    # the bar is "no worse against ground truth than the reference" (measured: 2.28 px here vs 2.51 px reference at depth 4)
    assert synth.epe(fm, gt, valid) <= synth.epe(fr, gt, valid) + 0.05
    ref.destroy(rc); ctx.close()


@pytest.mark.parametrize("name,stride", [("s3", 3), ("s1", 1)])
def test_patch_stride_variants_bit_exact(name, stride):
    """Third axis of BASELINE config 5: the sample stride of the 19x19 patch ("pixel skipping", bao_pmflow_kernel.cu:269,272; 2 upstream).
    Against the reference rebuilt with stride 3 / 1 (`make -C oracle variants`): the whole PatchMatch and the upsample + plane-fitting
    refine of both finer levels are bit-exact."""
    path = refharness.variant_lib(name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built")
    ref = refharness.Ref(path)
    h, w = (240, 320) if stride == 1 else (480, 640)
    a, b, gt, valid = synth.make_pair(h, w, 6, scale_to=0.25)
    rc = ref.create(h, w)
    ref.set_data(rc, a, b)
    p = E.default_params()
    p.patch_stride = stride
    ctx = E.EppmContext(h, w, 1, params=p)
    ctx.stage_prepare(dev(a[None]), dev(b[None]), 1)
    dims = [ref.level_dims(rc, l) for l in range(3)]
    img = [[refharness.pitched(ref.read_plane(rc, k, l)) for l in range(3)] for k in (0, 1)]
    cen = [[refharness.pitched(ref.read_plane(rc, 2 + k, l)) for l in range(3)] for k in (0, 1)]
    hc, wc = dims[2]
    nf, cf = ref.tap_patchmatch(img[0][2], img[1][2], cen[0][2], cen[1][2], wc, hc, 1000)
    nb, cb = ref.tap_patchmatch(img[1][2], img[0][2], cen[1][2], cen[0][2], wc, hc, 1000)
    ctx.stage_patchmatch()
    assert same_bits(ctx.read_plane(E.PLANE_NNF_FWD), nf) and same_bits(ctx.read_plane(E.PLANE_NNF_BWD), nb)
    assert same_bits(ctx.read_plane(E.PLANE_COST_FWD), cf) and same_bits(ctx.read_plane(E.PLANE_COST_BWD), cb)
    # refine: reference's own coarse flow in, upsample + plane-fitting refine out, per level
    fr = ref.compute_flow(rc, h, w)
    nl = 3
    PtrArr, IntArr, SzArr = C.c_void_p * nl, C.c_int * nl, C.c_size_t * nl
    ref.lib.baoCudaBLF_C2F.argtypes = [C.c_void_p] * 11 + [C.c_int]
    ref.lib.baoCudaBLF_C2F.restype = None
    for l in (1, 0):
        coarse = ref.read_plane(rc, 8, l + 1)
        hl, wl = dims[l]
        flows = [None] * nl
        flows[l] = torch.zeros((hl, wl, 2), dtype=torch.float32, device="cuda")
        flows[l + 1] = dev(coarse)
        ref.lib.baoCudaBLF_C2F(PtrArr(*[P(x) if x is not None else None for x in flows]), PtrArr(*[P(img[0][k][0]) for k in range(nl)]),
                               PtrArr(*[P(img[1][k][0]) for k in range(nl)]), PtrArr(*[P(cen[0][k][0]) for k in range(nl)]),
                               PtrArr(*[P(cen[1][k][0]) for k in range(nl)]), None, None, IntArr(*[d[0] for d in dims]), IntArr(*[d[1] for d in dims]),
                               SzArr(*[img[0][k][1] for k in range(nl)]), SzArr(*[cen[0][k][1] for k in range(nl)]), l)
        torch.cuda.synchronize()
        ctx.write_plane(E.PLANE_FLOW, coarse, level=l + 1)
        ctx.c2f_step(l, 0)
        assert same_bits(ctx.read_plane(E.PLANE_FLOW_TMP, l), flows[l].cpu().numpy()), f"refine differs at level {l}"
    fm = ctx.compute_batch_host(a[None], b[None])[0]
    assert synth.epe(fm, gt, valid) <= synth.epe(fr, gt, valid) + 0.05
    ref.destroy(rc); ctx.close()


def test_video_stream_reuses_frames_and_matches_pairwise():
    """eppm_compute_stream_device: 4 consecutive frames -> 3 flows, each frame prepared once; identical to the pair-by-pair result."""
    h, w = 192, 256
    base, _, _, _ = synth.make_pair(h, w, 50, scale_to=0.2)
    frames = np.stack([np.roll(base, (k, 2 * k), (0, 1)) for k in range(4)])
    ctx = E.EppmContext(h, w, 4)
    d_flow = torch.zeros((3, h, w, 2), dtype=torch.float32, device="cuda")
    ctx.compute_stream_device(dev(frames), 3, d_flow)
    ctx.synchronize()
    pairwise = ctx.compute_batch_host(frames[:3], frames[1:])
    assert same_bits(d_flow.cpu().numpy(), pairwise)
    ctx.close()
