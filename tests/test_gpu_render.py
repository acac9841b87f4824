"""GPU: the fused DeepVoxels render tail ("next" row, SURVEY 8f rank 1: deepvoxel.py:879-892 around
AccumulativeOcclusionNet.forward :574-587, depth rescale :903-904) through the C-ABI and the Python mirror,
against golden vectors produced by the reference's own code and against the CPU oracle at production size."""
import ctypes

import numpy as np
import pytest

from conftest import assert_grad_close, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

RENDER_CASES = ["render_g16", "render_g12_thr3"]


def _rel(a, b):
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))


def _call(P, R, grid, cam, W1, b1, W2, b2, g_novel=None, g_depth=None, g_fg=None, use_saved=True):
    from gpu_util import DEV, dev, p, stream
    from rgbd_gan_b200 import _lib
    B, F = grid.shape[:2]
    HW = P.H * P.W
    d = [dev(a) for a in (grid, cam.reshape(B, 16), W1, b1, W2.reshape(-1), b2)]
    ws = torch.empty(_lib.load().rgbd_dv_render_workspace_bytes(ctypes.byref(P), B, F), dtype=torch.uint8, device=DEV)
    novel = torch.full((B, F, HW), float("nan"), device=DEV)
    depth = torch.full((B, HW), float("nan"), device=DEV)
    fg = torch.full((B, HW), float("nan"), device=DEV)
    saved = torch.full((_lib.load().rgbd_dv_render_saved_bytes(ctypes.byref(P), B) // 4,), float("nan"), device=DEV) \
        if use_saved else None
    _lib.call("rgbd_dv_render_fwd", ctypes.byref(P), ctypes.byref(R), *[p(t) for t in d], B, F, p(novel), p(depth), p(fg),
              p(saved), p(ws), ws.numel(), stream())
    out = [novel.cpu().numpy(), depth.cpu().numpy(), fg.cpu().numpy()]
    if g_novel is None:
        return out
    gg = torch.full((B, F, P.G ** 3), float("nan"), device=DEV)
    gW1, gb1 = torch.full(W1.shape, float("nan"), device=DEV), torch.full(b1.shape, float("nan"), device=DEV)
    gW2, gb2 = torch.full((W2.size,), float("nan"), device=DEV), torch.full((1,), float("nan"), device=DEV)
    ups = [dev(g_novel), dev(g_depth), None if g_fg is None else dev(g_fg)]     # keep the device copies alive
    _lib.call("rgbd_dv_render_bwd", ctypes.byref(P), ctypes.byref(R), *[p(t) for t in d], B, F, p(saved), p(ups[0]), p(ups[1]),
              p(ups[2]), p(gg), p(gW1), p(gb1), p(gW2), p(gb2), p(ws), ws.numel(), stream())
    torch.cuda.synchronize()
    return out + [t.cpu().numpy() for t in (gg, gW1, gb1, gW2, gb2)]


def _params(G, img, D, voxel_size, near_plane, F, nf, threshold):
    from rgbd_gan_b200._lib import DvParams, DvRenderParams
    P = DvParams(img, img, D, G, 2. * img, 2. * img, img / 2., img / 2., float(np.float32(voxel_size)),
                 float(np.float32(near_plane)))
    R = DvRenderParams(nf, int(np.ceil(np.sqrt(3) * G)), float(threshold),
                       float(np.float32(np.sqrt(2) * np.sqrt(1.0 / (F + 1)))), float(np.float32(np.sqrt(2) * np.sqrt(1.0 / nf))))
    return P, R


@pytest.mark.parametrize("use_saved", [True, False])
@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_against_reference_golden(name, use_saved):
    """use_saved: the backward reads the running sums the forward left (one reverse sweep); otherwise it repeats
    the forward walk first -- same results"""
    g = load_golden(name)
    G, img, F, D = int(g["G"]), int(g["img"]), int(g["F"]), int(g["D"])
    P, R = _params(G, img, D, float(g["voxel_size"]), float(g["near_plane"]), F, int(g["nf"]), float(g["threshold"]))
    novel, depth, fg, gg, gW1, gb1, gW2, gb2 = _call(P, R, g["grid"], g["cam"], g["W1"], g["b1"], g["W2"], g["b2"],
                                                      g["g_novel"], g["g_depth"], g["g_fg"], use_saved=use_saved)
    assert _rel(novel.reshape(g["novel"].shape), g["novel"]) <= 1e-5
    assert _rel(depth.reshape(g["depth"].shape), g["depth"]) <= 1e-5
    assert _rel(fg.reshape(g["fg"].shape), g["fg"]) <= 1e-5
    assert_grad_close(gg.reshape(g["g_grid"].shape), g["g_grid"])
    for got, key in ((gW1, "g_W1"), (gb1, "g_b1"), (gW2, "g_W2"), (gb2, "g_b2")):
        assert _rel(got.reshape(g[key].shape), g[key]) <= 1e-5, key


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_mirror_like_the_generator(name):
    """ProjectionHelper.render_accumulative driven the way deepvoxels_generator.py:287-299 drives the reference"""
    from rgbd_gan_b200.projection import ProjectionHelper
    g = load_golden(name)
    G, img, D = int(g["G"]), int(g["img"]), int(g["D"])
    h = ProjectionHelper(g["intrinsic"], g["intrinsic"], [img, img], [img, img], 0., 1., [G] * 3,
                         float(g["voxel_size"]), g["near_plane"], D, verbose=False)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0").requires_grad_(True)
    grid, W1, b1, W2, b2 = t(g["grid"]), t(g["W1"]), t(g["b1"]), t(g["W2"]), t(g["b2"])
    novel, depth, fg = h.render_accumulative(grid, g["cam"], W1, b1, W2, b2, float(g["threshold"]), True)
    assert tuple(novel.shape) == g["novel"].shape and tuple(depth.shape) == g["depth"].shape
    dev = lambda a: torch.from_numpy(a).to("cuda:0")
    loss = (novel * dev(g["g_novel"])).sum() + (depth * dev(g["g_depth"])).sum() + (fg * dev(g["g_fg"])).sum()
    loss.backward()
    assert _rel(novel.detach().cpu().numpy(), g["novel"]) <= 1e-5
    assert _rel(depth.detach().cpu().numpy(), g["depth"]) <= 1e-5
    assert_grad_close(grid.grad.cpu().numpy(), g["g_grid"])
    for got, key in ((W1, "g_W1"), (b1, "g_b1"), (W2, "g_W2"), (b2, "g_b2")):
        assert _rel(got.grad.cpu().numpy(), g[key]) <= 1e-5, key
    # without the foreground weight (the default return of DeepVoxels.forward)
    out = h.render_accumulative(grid.detach(), g["cam"], W1.detach(), b1.detach(), W2.detach(), b2.detach(), float(g["threshold"]))
    assert len(out) == 2


@pytest.mark.parametrize("G,B", [(32, 2)])
def test_render_full_size_against_oracle_and_properties(G, B, oracle_mod):
    """production geometry (deepvoxels_generator.py:229-253: G=32, F=32, 64x64x56, nf 4, threshold 4)"""
    from oracle import numpy_port as poses
    img, F, nf, thr = 64, 32, 4, 4.0
    D = int(np.ceil(np.sqrt(3) * G))
    vs, near = (1. / G) * 1.1 * 0.5, np.sqrt(3) / 4
    K = np.array([[128., 0, 32., 0], [0, 128., 32., 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    np.random.seed(5)
    thetas = poses.sample_camera_prior(2 * B, poses.CAR_RANGES, True)[:B]
    cam = poses.get_camera_matries(thetas)
    rng = np.random.default_rng(5)
    grid = rng.normal(size=(B, F, G, G, G)).astype(np.float32)
    W1 = rng.normal(size=(nf, F + 1)).astype(np.float32)
    b1 = rng.normal(scale=0.3, size=(nf,)).astype(np.float32)
    W2 = rng.normal(size=(1, nf)).astype(np.float32)
    b2 = np.array([-0.5], np.float32)                     # some rays saturate, others do not
    P, R = _params(G, img, D, vs, near, F, nf, thr)
    P0 = oracle_mod.dv_params(img, img, D, G, K, vs, near)
    args = (P0, grid, cam, W1, b1, W2, b2, thr, R.inv_c1, R.inv_c2, D)
    ref_novel, ref_depth, ref_fg = oracle_mod.dv_render_fwd(*args)
    g_novel = rng.normal(size=ref_novel.shape).astype(np.float32)
    g_depth = rng.normal(size=(B, img, img)).astype(np.float32)
    g_fg = rng.normal(size=(B, img, img)).astype(np.float32)
    novel, depth, fg, gg, gW1, gb1, gW2, gb2 = _call(P, R, grid, cam, W1, b1, W2, b2, g_novel, g_depth, g_fg)
    sat = float((ref_fg > 0.999999).mean())
    assert 0.02 < sat < 0.98, sat                          # the test exercises both saturated and open rays
    assert _rel(novel.reshape(ref_novel.shape), ref_novel) <= 1e-5
    assert _rel(depth.reshape(ref_depth.shape), ref_depth) <= 1e-5
    assert _rel(fg.reshape(ref_fg.shape), ref_fg) <= 1e-5
    assert float(fg.min()) >= 0.0 and float(fg.max()) <= 1.0 + 1e-6      # a convex combination along every ray
    rg, rW1, rb1, rW2, rb2 = oracle_mod.dv_render_bwd(*args, g_novel, g_depth, g_fg)
    assert_grad_close(gg.reshape(rg.shape), rg)
    for got, want in ((gW1, rW1), (gb1, rb1), (gW2.reshape(rW2.shape), rW2), (gb2, rb2)):
        assert _rel(got, want) <= 2e-5
    # without g_fg == with zeros
    out2 = _call(P, R, grid, cam, W1, b1, W2, b2, g_novel, g_depth, None)
    out3 = _call(P, R, grid, cam, W1, b1, W2, b2, g_novel, g_depth, np.zeros_like(g_fg), use_saved=False)
    np.testing.assert_allclose(out2[4], out3[4], rtol=1e-6, atol=1e-9)


def test_render_rejects_unsupported_shapes():
    from gpu_util import DEV, p, stream
    from rgbd_gan_b200 import _lib
    P, R = _params(16, 32, 28, 0.03, 0.43, 32, 4, 4.0)
    R.nf = 8
    x = torch.zeros(64, device=DEV)
    rc = _lib.load().rgbd_dv_render_fwd(ctypes.byref(P), ctypes.byref(R), p(x), p(x), p(x), p(x), p(x), p(x), 1, 32,
                                        p(x), p(x), None, None, p(x), 0, stream())
    assert rc == -4
