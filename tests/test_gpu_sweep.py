"""GPU parity tests of the persistent row-sweep kernel (rgbd_gan_b200/csrc/sweep.cuh), through the C-ABI.

The sweep is the default C == 4 path (RGBD_B200_SWEEP=1); =2 is its debug variant (same compute loop, REDs into a
global accumulator + stage-out kernel) and =0 the three-kernel chain of round 1.  All three must reproduce the C
oracle (oracle/rgbd_oracle.c, itself pinned to the reference's golden vectors): loss parts 1e-5, gradients 1e-5 of
the max-norm.  RGBD_B200_SWEEP_CTAS forces chunk boundaries inside pairs (scatter contributions that cross a
boundary go through the per-CTA list + fix-up kernel) and tiny chunks (most of the scatter goes through the list);
wild poses / depths push taps outside the shared-memory window (exact slow path through L2)."""
import numpy as np
import pytest

from conftest import assert_grad_close, case_options, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _problem(B, S, depth, seed, ranges=None, wild=False):
    from oracle import numpy_port as poses
    from oracle import numpy_port as npp
    kw = {} if ranges is None else dict(ranges=ranges)
    x, cam = poses.synthetic_batch(B, S, depth=depth, seed=seed, **kw)
    if wild:
        rng = np.random.default_rng(seed + 100)
        x[:, -1] = rng.uniform(0.05, 4.0, size=x[:, -1].shape).astype(np.float32)      # huge parallax
        x[::3, -1, ::7, ::5] = 0.0                                                      # q2 <= 1e-4 cases
    port = npp.LossFuncRotateNP(lambda_geometric=3)
    port.init_params(S)
    return x, cam, port


def _run(monkeypatch, mode, ctas, x, cam, port, B, gy, **kw):
    from gpu_util import Consistency
    monkeypatch.setenv("RGBD_B200_SWEEP", mode)
    if ctas:
        monkeypatch.setenv("RGBD_B200_SWEEP_CTAS", str(ctas))
    else:
        monkeypatch.delenv("RGBD_B200_SWEEP_CTAS", raising=False)
    drv = Consistency(x, cam, B, port.K, port.inv_K, **kw)
    parts, gi, gr = drv.fwd_bwd(gy=gy)
    return drv, parts, gi, gr


@pytest.mark.parametrize("B,S,depth,ctas", [
    (4, 128, "rough", 0),        # cfg0: 64 blocks -> 64 CTAs of one block each (everything crosses a boundary)
    (4, 128, "rough", 5),        # uneven chunks, boundaries inside pairs
    (32, 128, "rough", 0),       # cfg1 on a full machine
    (6, 64, "smooth", 7),
    (3, 64, "rough", 1),         # one CTA sweeps everything: no list traffic at all
    (37, 128, "smooth", 0),
])
@pytest.mark.parametrize("mode", ["1", "2"])
def test_sweep_against_oracle(B, S, depth, ctas, mode, oracle_mod, monkeypatch):
    x, cam, port = _problem(B, S, depth, seed=B + S)
    drv, parts, gi, gr = _run(monkeypatch, mode, ctas, x, cam, port, B, 2.0, lam=3.0, occ=True)
    M, c, Mi, ci = drv.host_poses
    ref_parts = oracle_mod.consistency_fwd(x[:B], x[B:], M, c, Mi, ci, norm=1, occlusion=True)
    ref_gi, ref_gr = oracle_mod.consistency_bwd(x[:B], x[B:], M, c, Mi, ci, norm=1, occlusion=True,
                                                lambda_geometric=3, gy=2.0)
    np.testing.assert_allclose(parts[:4], ref_parts, rtol=1e-5)
    assert_grad_close(gi, ref_gi)
    assert_grad_close(gr, ref_gr)
    # loss-only and gradient-only entry points run the same kernel with other template flags
    parts2, _, _ = drv.fwd(want_zp=False, want_masks=False)
    np.testing.assert_allclose(parts2[:5], parts[:5], rtol=1e-6)
    gi2, gr2 = drv.bwd(gy=1.0, gy_dev=2.0)
    assert_grad_close(gi2, ref_gi)
    assert_grad_close(gr2, ref_gr)
    # repeated calls on the same workspace (ring rows are re-zeroed ahead of the sweep, lists restart at 0)
    parts3, gi3, gr3 = drv.fwd_bwd(gy=2.0)
    np.testing.assert_array_equal(parts3, parts)
    assert np.abs(gi3 - gi).max() <= 1e-6 * np.abs(gi).max()
    assert np.abs(gr3 - gr).max() <= 1e-6 * np.abs(gr).max()


@pytest.mark.parametrize("S,ctas", [(64, 0), (128, 3), (128, 0)])
@pytest.mark.parametrize("norm", ["l1", "l2"])
def test_sweep_wild_poses_leave_the_window(S, ctas, norm, oracle_mod, monkeypatch):
    """yaw up to pi between twins is not what the reference samples, but the kernel must stay exact: taps far outside the
    +-16-row window take the L2 gather and the list scatter"""
    B = 5
    x, cam, port = _problem(B, S, "rough", seed=3, ranges=(1.2, 3.1415, 0.8, 0.3, 0.3, 0.3), wild=True)
    drv, parts, gi, gr = _run(monkeypatch, "1", ctas, x, cam, port, B, 0.7, norm=norm, lam=3.0, occ=True)
    M, c, Mi, ci = drv.host_poses
    n = 1 if norm == "l1" else 2
    ref_parts = oracle_mod.consistency_fwd(x[:B], x[B:], M, c, Mi, ci, norm=n, occlusion=True)
    ref_gi, ref_gr = oracle_mod.consistency_bwd(x[:B], x[B:], M, c, Mi, ci, norm=n, occlusion=True,
                                                lambda_geometric=3, gy=0.7)
    np.testing.assert_allclose(parts[:4], ref_parts, rtol=1e-5)
    assert_grad_close(gi, ref_gi)
    assert_grad_close(gr, ref_gr)


@pytest.mark.parametrize("name", ["loss_cfg0_l1_occ", "loss_s64_l1_noocc", "loss_car_l1_occ", "loss_dv_maxdepth",
                                  "loss_dv_mindepth"])
@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_all_paths_reproduce_the_reference_golden(name, mode, monkeypatch):
    """incl. the depth-range masks of the DeepVoxels updater (updater_deepvoxels.py:176-190), which the sweep
    evaluates itself instead of dropping to the generic kernel"""
    from gpu_util import Consistency
    from oracle import numpy_port as npp
    monkeypatch.setenv("RGBD_B200_SWEEP", mode)
    g = load_golden(name)
    o = case_options(g)
    port = npp.LossFuncRotateNP(K=None if o["K"] is None else o["K"].copy(), norm=o["norm"], lambda_geometric=o["lam"])
    port.init_params(o["S"])
    drv = Consistency(g["x"], g["cam"], o["B"], port.K, port.inv_K, norm=o["norm"], lam=o["lam"], occ=o["occ"],
                      max_depth=o["max_depth"], min_depth=o["min_depth"])
    parts, gi, gr = drv.fwd_bwd(gy=o["gy"])
    assert abs(parts[4] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert_grad_close(gi, g["g_img"])
    assert_grad_close(gr, g["g_img_rot"])


@pytest.mark.parametrize("name", ["hinge_ffhq", "hinge_car"])
@pytest.mark.parametrize("ctas", [0, 3])
def test_sweep_depth_hinge(name, ctas, monkeypatch):
    from gpu_util import Consistency
    from oracle import numpy_port as npp
    monkeypatch.setenv("RGBD_B200_SWEEP", "1")
    if ctas:
        monkeypatch.setenv("RGBD_B200_SWEEP_CTAS", str(ctas))
    g = load_golden(name)
    B, S, gy = int(g["B"]), int(g["S"]), float(g["gy"])
    port = npp.LossFuncRotateNP(lambda_geometric=3)
    port.init_params(S)
    drv = Consistency(g["x"], g["cam"], B, port.K, port.inv_K, lam=3.0, occ=True)
    drv.opts.hinge_depth_min, drv.opts.hinge_lambda = float(g["depth_min"]), float(g["lambda_depth"])
    parts, gi, gr = drv.fwd_bwd(gy=gy)
    assert abs(parts[5] - float(g["hinge"])) <= 1e-5 * max(abs(float(g["hinge"])), 1e-12)
    assert abs(parts[6] - float(g["total"])) <= 1e-5 * abs(float(g["total"]))
    assert_grad_close(gi, g["g_img"])
    assert_grad_close(gr, g["g_img_rot"])


@pytest.mark.parametrize("e_lo,e_hi", [(-20, 20), (-70, 70), (-126, 126)])
def test_div2_matches_ieee_division(e_lo, e_hi):
    """the shared-reciprocal division that feeds the truncated pixel indices (both kernels' variants) against __fdiv_rn on
    1e8 pseudo-random (a0, a1, b in [1e-4, 1e4]) per exponent range, incl. the 2^-60 / 2^60 fallbacks: bit-equal"""
    import ctypes
    from rgbd_gan_b200 import _lib
    counts = torch.zeros(3, dtype=torch.int64, device="cuda:0")
    _lib.call("rgbd_debug_div2", ctypes.c_ulonglong(100_000_000), 12345 + e_hi, e_lo, e_hi, ctypes.c_void_p(counts.data_ptr()),
              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    c = counts.cpu().numpy()
    assert c[0] == 0 and c[1] == 0, c
    if e_hi >= 70:
        assert c[2] > 0            # the fallback was exercised


def test_feature_space_shape_c257_against_oracle(oracle_mod):
    """SURVEY 8(f) rank 3 at its production shape (updater.py:345-354): C = 256 features + 1 depth at 32x32, norm l2 --
    8 lane iterations and an odd row stride in the warp-per-pixel kernels"""
    from gpu_util import Consistency
    from oracle import numpy_port as npp
    B, C, S = 6, 257, 32
    rng = np.random.default_rng(4)
    _, cam = npp.synthetic_batch(B, S, depth="rough", seed=4)
    x = rng.uniform(-1, 1, size=(2 * B, C, S, S)).astype(np.float32)
    x[:, -1] = rng.uniform(0.7, 1.5, size=(2 * B, S, S))
    port = npp.LossFuncRotateNP(norm="l2", lambda_geometric=3)
    port.init_params(S)
    drv = Consistency(x, cam, B, port.K, port.inv_K, norm="l2", lam=3.0, occ=True)
    M, c, Mi, ci = drv.host_poses
    ref_parts, d = oracle_mod.consistency_fwd(x[:B], x[B:], M, c, Mi, ci, norm=2, occlusion=True, debug=True)
    ref_gi, ref_gr = oracle_mod.consistency_bwd(x[:B], x[B:], M, c, Mi, ci, norm=2, occlusion=True, lambda_geometric=3, gy=2.0)
    parts, zp, masks = drv.fwd()
    np.testing.assert_array_equal(zp, d["new_zp"])
    np.testing.assert_array_equal(masks[0], d["mask"])
    np.testing.assert_array_equal(masks[1], d["occ"])
    np.testing.assert_allclose(parts[:4], ref_parts, rtol=1e-5)
    parts2, gi, gr = drv.fwd_bwd(gy=2.0)
    np.testing.assert_allclose(parts2[:4], ref_parts, rtol=1e-5)
    assert_grad_close(gi, ref_gi)
    assert_grad_close(gr, ref_gr)


@pytest.mark.parametrize("mode", ["0", "1"])
def test_configs2_car_poses_full_size(mode, oracle_mod, monkeypatch):
    """BASELINE.json configs[2] (dcgan_shapenet_car.yml): 64 pairs at 128x128, car pose ranges (yaw +-pi), lambda_geometric 1,
    occlusion on, through both C == 4 paths"""
    from oracle import numpy_port as npp
    monkeypatch.setenv("RGBD_B200_SWEEP", mode)
    B, S = 64, 128
    x, cam = npp.synthetic_batch(B, S, depth="rough", ranges=npp.CAR_RANGES, seed=21)
    port = npp.LossFuncRotateNP(lambda_geometric=1)
    port.init_params(S)
    from gpu_util import Consistency
    drv = Consistency(x, cam, B, port.K, port.inv_K, lam=1.0, occ=True)
    M, c, Mi, ci = drv.host_poses
    ref_parts = oracle_mod.consistency_fwd(x[:B], x[B:], M, c, Mi, ci, norm=1, occlusion=True)
    ref_gi, ref_gr = oracle_mod.consistency_bwd(x[:B], x[B:], M, c, Mi, ci, norm=1, occlusion=True, lambda_geometric=1, gy=2.0)
    parts, gi, gr = drv.fwd_bwd(gy=2.0)
    np.testing.assert_allclose(parts[:4], ref_parts, rtol=1e-5)
    lam = np.float32(1.0)
    assert parts[4] == np.float32((parts[0] + parts[1]) + (parts[2] * lam + parts[3] * lam))
    assert_grad_close(gi, ref_gi)
    assert_grad_close(gr, ref_gr)
