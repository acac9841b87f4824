"""GPU parity tests of the DeepVoxels projection path through the C-ABI:
lin_ind / voxel_coords / frustum values BIT-EXACT vs the reference's golden vectors and the C oracle;
the lift (backward) within 1e-5 of max-norm."""
import ctypes

import numpy as np
import pytest

from conftest import DV_CASES, assert_frustum_close, assert_grad_close, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _params(g):
    from rgbd_gan_b200._lib import DvParams
    K = g["intrinsic"]
    img = int(g["img"])
    return DvParams(img, img, int(g["D"]), int(g["G"]), float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]),
                    float(np.float32(g["voxel_size"])), float(np.float32(g["near_plane"])))


def _proj_idcs(P, cam):
    from gpu_util import DEV, dev, p, stream
    from rgbd_gan_b200 import _lib
    n = P.W * P.H * P.D
    lin = torch.empty(n, dtype=torch.int32, device=DEV)
    vc = torch.empty((3, n), device=DEV)
    ws = torch.empty(_lib.load().rgbd_dv_workspace_bytes(ctypes.byref(P)), dtype=torch.uint8, device=DEV)
    M = ctypes.c_int(-1)
    d_cam = dev(cam.reshape(16))
    _lib.call("rgbd_dv_compute_proj_idcs", ctypes.byref(P), p(d_cam), p(lin), p(vc), ctypes.byref(M),
              p(ws), ws.numel(), stream())
    return M.value, lin, vc


@pytest.mark.parametrize("name", DV_CASES)
def test_projection_against_reference_golden(name, monkeypatch):
    from gpu_util import DEV, dev, p, stream
    from rgbd_gan_b200 import _lib
    g = load_golden(name)
    P = _params(g)
    F, n = int(g["F"]), P.W * P.H * P.D
    G3 = P.G ** 3
    ns = g["cam"].shape[0]
    for i in range(ns):
        M, lin, vc = _proj_idcs(P, g["cam"][i])
        assert M == g["lin_ind_%d" % i].size
        np.testing.assert_array_equal(lin[:M].cpu().numpy(), g["lin_ind_%d" % i])          # bit-exact, ordered
        np.testing.assert_array_equal(vc[:, :M].cpu().numpy(), g["voxel_coords_%d" % i])   # bit-exact
        grid = dev(g["grid"][i])
        out = torch.full((F, n), float("nan"), device=DEV)
        _lib.call("rgbd_dv_trilinear_fwd", p(grid), p(lin), p(vc), n, M, F, ctypes.byref(P), p(out), stream())
        np.testing.assert_array_equal(out.cpu().numpy().reshape(g["frustum_%d" % i][0].shape), g["frustum_%d" % i][0])
        gout = dev(g["g_out"][i])
        ggrid = torch.full((F, G3), float("nan"), device=DEV)
        _lib.call("rgbd_dv_trilinear_bwd", p(gout), p(lin), p(vc), n, M, F, ctypes.byref(P), p(ggrid), stream())
        assert_grad_close(ggrid.cpu().numpy().reshape(g["g_grid_%d" % i][0].shape), g["g_grid_%d" % i][0])
    # fused batch path: channels-last fast path (with workspace) and planar fallback (without)
    grid, cam, gout = dev(g["grid"]), dev(g["cam"].reshape(ns, 16)), dev(g["g_out"])
    nbytes = _lib.load().rgbd_dv_project_workspace_bytes(ctypes.byref(P), ns, F)
    assert nbytes >= G3 * F * 4
    for ws, exact in ((torch.empty(nbytes, dtype=torch.uint8, device=DEV), "1"),
                      (torch.empty(nbytes, dtype=torch.uint8, device=DEV), "0"), (None, "0")):
        monkeypatch.setenv("RGBD_B200_DV_EXACT", exact)
        out = torch.full((ns, F, n), float("nan"), device=DEV)
        _lib.call("rgbd_dv_project_fwd", ctypes.byref(P), p(grid), p(cam), ns, F, p(out), p(ws),
                  0 if ws is None else ws.numel(), stream())
        ggrid = torch.full((ns, F, G3), float("nan"), device=DEV)
        _lib.call("rgbd_dv_project_bwd", ctypes.byref(P), p(gout), p(cam), ns, F, p(ggrid), p(ws),
                  0 if ws is None else ws.numel(), stream())
        out, ggrid = out.cpu().numpy(), ggrid.cpu().numpy()
        for i in range(ns):
            # (the planar fallback without workspace always evaluates the exact chain)
            assert_frustum_close(out[i].reshape(g["frustum_%d" % i][0].shape), g["frustum_%d" % i][0],
                                 exact == "1" or ws is None)
            assert_grad_close(ggrid[i].reshape(g["g_grid_%d" % i][0].shape), g["g_grid_%d" % i][0])


@pytest.mark.parametrize("exact", ["0", "1"])
@pytest.mark.parametrize("G,F,B", [(32, 32, 4), (64, 32, 2)])
def test_full_size_against_oracle(G, F, B, exact, oracle_mod, monkeypatch):
    """production geometry (deepvoxels_generator.py:229-253: G=32, F=32, 64x64x56) and BASELINE's 64^3"""
    from gpu_util import DEV, dev, p, stream
    from rgbd_gan_b200 import _lib
    from oracle import numpy_port as poses
    monkeypatch.setenv("RGBD_B200_DV_EXACT", exact)
    img = 64
    D = int(np.ceil(np.sqrt(3) * G))
    vs = (1. / G) * 1.1 * 0.5
    K = np.array([[128., 0, 32., 0], [0, 128., 32., 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    np.random.seed(3)
    thetas = poses.sample_camera_prior(2 * B, poses.CAR_RANGES, True)[:B]
    cam = poses.get_camera_matries(thetas)
    rng = np.random.default_rng(3)
    grid = rng.normal(size=(B, F, G, G, G)).astype(np.float32)
    P0 = oracle_mod.dv_params(img, img, D, G, K, vs, np.sqrt(3) / 4)
    ref = oracle_mod.dv_project_fwd(P0, grid, cam)
    from rgbd_gan_b200._lib import DvParams
    P = DvParams(img, img, D, G, 128., 128., 32., 32., float(np.float32(vs)), float(np.float32(np.sqrt(3) / 4)))
    n = img * img * D
    out = torch.empty((B, F, n), device=DEV)
    ws = torch.empty(_lib.load().rgbd_dv_project_workspace_bytes(ctypes.byref(P), B, F), dtype=torch.uint8, device=DEV)
    d_grid, d_cam = dev(grid), dev(cam.reshape(B, 16))       # referenced until the end: raw pointers go to the library
    _lib.call("rgbd_dv_project_fwd", ctypes.byref(P), p(d_grid), p(d_cam), B, F, p(out), p(ws), ws.numel(), stream())
    assert_frustum_close(out.cpu().numpy().reshape(ref.shape), ref, exact == "1")
    g_out = rng.normal(size=ref.shape).astype(np.float32)
    ref_g = oracle_mod.dv_project_bwd(P0, g_out, cam)
    gg = torch.empty((B, F, G ** 3), device=DEV)
    d_gout = dev(g_out)
    _lib.call("rgbd_dv_project_bwd", ctypes.byref(P), p(d_gout), p(d_cam), B, F, p(gg), p(ws), ws.numel(), stream())
    assert_grad_close(gg.cpu().numpy().reshape(ref_g.shape), ref_g)
    # adjointness <frustum(grid), g_out> == <grid, lift(g_out)> : a size-independent property of the pair
    lhs = float((ref.astype(np.float64) * g_out).sum())
    rhs = float((grid.astype(np.float64) * gg.cpu().numpy().reshape(grid.shape)).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0) + 1e-3


def test_empty_frustum_returns_zero_count():
    g = load_golden("dv_g16_f3")
    P = _params(g)
    cam = g["cam"][0].copy()
    cam[:3, 3] += 100.0
    M, _, _ = _proj_idcs(P, cam)
    assert M == 0


def test_grid2world_argument_bit_exact():
    """compute_proj_idcs(cam2world, grid2world) (projection.py:48,53-54,83-84) on the device in the reference's order"""
    from rgbd_gan_b200.projection import ProjectionHelper
    g = load_golden("dv_g16_f3")
    G, img, D = int(g["G"]), int(g["img"]), int(g["D"])
    h = ProjectionHelper(g["intrinsic"], g["intrinsic"], [img, img], [img, img], 0., 1., [G] * 3, float(g["voxel_size"]),
                         float(g["near_plane"]), D, verbose=False)
    lin, vc = h.compute_proj_idcs(g["cam"][0], g["grid2world"])
    np.testing.assert_array_equal(lin.cpu().numpy(), g["lin_ind_g2w_0"])
    np.testing.assert_array_equal(vc.cpu().numpy(), g["voxel_coords_g2w_0"])
    lin0, vc0 = h.compute_proj_idcs(g["cam"][0])                 # the plain call is untouched
    np.testing.assert_array_equal(lin0.cpu().numpy(), g["lin_ind_0"])
    np.testing.assert_array_equal(vc0.cpu().numpy(), g["voxel_coords_0"])
