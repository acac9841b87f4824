"""GPU parity tests of the consistency-loss path, through the C-ABI:
CUDA vs the golden vectors of the unmodified reference and vs the C oracle on the same inputs.
Bit-exact: new_zp, in-bounds mask, occlusion mask.  1e-5 (relative to max-norm): loss parts, gradients."""
import os

import numpy as np
import pytest

from conftest import LOSS_CASES, assert_grad_close, case_options, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _port(o, S):
    from oracle import numpy_port as npp
    port = npp.LossFuncRotateNP(K=None if o["K"] is None else o["K"].copy(), norm=o["norm"], lambda_geometric=o["lam"])
    port.init_params(S)
    return port


def _driver(g, o, **kw):
    from gpu_util import Consistency
    port = _port(o, o["S"])
    args = dict(norm=o["norm"], lam=o["lam"], occ=o["occ"], max_depth=o["max_depth"], min_depth=o["min_depth"])
    args.update(kw)
    return Consistency(g["x"], g["cam"], o["B"], port.K, port.inv_K, **args)


@pytest.mark.parametrize("name", LOSS_CASES)
def test_forward_against_reference_golden(name):
    g = load_golden(name)
    o = case_options(g)
    B, S = o["B"], o["S"]
    N = B * S * S
    drv = _driver(g, o)
    parts, zp, masks = drv.fwd()
    np.testing.assert_array_equal(zp, g["new_zp_cat"])                                  # bit-exact geometry
    np.testing.assert_array_equal(masks[0][:N].astype(bool), g["not_out"])              # bit-exact masks
    np.testing.assert_array_equal(masks[0][N:].astype(bool), g["not_out_rot"])
    if o["occ"]:
        q2 = g["new_zp_cat"].reshape(-1, 3)[:, 2]
        np.testing.assert_array_equal(masks[1][:N].astype(bool), g["warped"][:, -1] > q2[:N])
        np.testing.assert_array_equal(masks[1][N:].astype(bool), g["warped_rot"][:, -1] > q2[N:])
    else:
        assert masks[1].all()
    assert abs(parts[4] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    lam = np.float32(o["lam"])
    assert parts[4] == np.float32((parts[0] + parts[1]) + (parts[2] * lam + parts[3] * lam))
    assert parts[5] == 0 and parts[6] == parts[4] and parts[7] == 0        # depth hinge off


@pytest.mark.parametrize("name", LOSS_CASES)
def test_backward_against_reference_golden(name):
    g = load_golden(name)
    o = case_options(g)
    drv = _driver(g, o)
    gi, gr = drv.bwd(gy=o["gy"])
    assert_grad_close(gi, g["g_img"])
    assert_grad_close(gr, g["g_img_rot"])
    # upstream gradient handed over as a device scalar (no host read in a FunctionNode.backward)
    gi2, gr2 = drv.bwd(gy=1.0, gy_dev=o["gy"])
    assert_grad_close(gi2, g["g_img"])
    assert_grad_close(gr2, g["g_img_rot"])
    # one-pass forward+backward
    parts, gi3, gr3 = drv.fwd_bwd(gy=o["gy"])
    assert abs(parts[4] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert_grad_close(gi3, g["g_img"])
    assert_grad_close(gr3, g["g_img_rot"])


@pytest.mark.parametrize("name", ["loss_cfg0_l1_occ", "loss_s32_l2_feat", "loss_edge_wild"])
def test_against_c_oracle_incl_parts_and_new_zp_grad(name, oracle_mod):
    g = load_golden(name)
    o = case_options(g)
    B, S = o["B"], o["S"]
    drv = _driver(g, o)
    M, c, Mi, ci = drv.host_poses
    norm = 1 if o["norm"] == "l1" else 2
    kw = dict(norm=norm, occlusion=o["occ"], max_depth=o["max_depth"], min_depth=o["min_depth"])
    ref_parts, d = oracle_mod.consistency_fwd(g["x"][:B], g["x"][B:], M, c, Mi, ci, debug=True, **kw)
    parts, zp, masks = drv.fwd()
    np.testing.assert_allclose(parts[:4], ref_parts, rtol=1e-5)
    np.testing.assert_array_equal(zp, d["new_zp"])
    np.testing.assert_array_equal(masks[0], d["mask"])
    np.testing.assert_array_equal(masks[1], d["occ"])
    rng = np.random.default_rng(3)
    gzp = (rng.normal(size=zp.shape) * 1e-6).astype(np.float32)
    ref_gi, ref_gr = oracle_mod.consistency_bwd(g["x"][:B], g["x"][B:], M, c, Mi, ci, lambda_geometric=o["lam"],
                                                gy=0.7, g_new_zp=gzp, **kw)
    gi, gr = drv.bwd(gy=0.7, g_new_zp=gzp)
    assert_grad_close(gi, ref_gi)
    assert_grad_close(gr, ref_gr)


def test_sharded_denominators_and_chunking(oracle_mod, monkeypatch):
    """n_pairs_global != B (a shard) and a chunk budget that forces several chunks"""
    g = load_golden("loss_s64_l1_noocc")
    o = case_options(g)
    B = o["B"]
    monkeypatch.setenv("RGBD_B200_CHUNK_MB", "1")          # 8 * 64 KiB per pair -> 2 pairs per chunk
    drv = _driver(g, o, n_pairs_global=4 * B)
    M, c, Mi, ci = drv.host_poses
    ref_parts = oracle_mod.consistency_fwd(g["x"][:B], g["x"][B:], M, c, Mi, ci, norm=1, occlusion=o["occ"],
                                           n_pairs_global=4 * B)
    ref_gi, ref_gr = oracle_mod.consistency_bwd(g["x"][:B], g["x"][B:], M, c, Mi, ci, norm=1, occlusion=o["occ"],
                                                lambda_geometric=o["lam"], n_pairs_global=4 * B, gy=2.0)
    parts, gi, gr = drv.fwd_bwd(gy=2.0)
    np.testing.assert_allclose(parts[:4], ref_parts, rtol=1e-5)
    np.testing.assert_allclose(parts[:4] * 4, [float(x) for x in _full_parts(g, o, oracle_mod)], rtol=1e-5)
    assert_grad_close(gi, ref_gi)
    assert_grad_close(gr, ref_gr)
    assert_grad_close(gi * 4, g["g_img"])


def _full_parts(g, o, oracle_mod):
    B = o["B"]
    port = _port(o, o["S"])
    M, c, Mi, ci = port.pose_algebra(g["cam"][:B], g["cam"][B:])
    return oracle_mod.consistency_fwd(g["x"][:B], g["x"][B:], M, c, Mi, -ci, norm=1, occlusion=o["occ"])


@pytest.mark.parametrize("name", ["loss_cfg0_l1_occ", "loss_s64_l1_noocc", "loss_edge_wild", "loss_s32_l2_noocc", "loss_car_l1_occ"])
@pytest.mark.parametrize("band", ["0", "1"])
@pytest.mark.parametrize("band_tr", ["0", "8"])
def test_band_and_staged_kernels_agree_with_reference(name, band, band_tr, monkeypatch):
    """C=4 has two main kernels: the staged fast kernel (default) and the opt-in shared-memory band kernel
    (no staging copy, RGBD_B200_BAND=1).  Both must reproduce the reference; a small band (TR=8) forces the
    out-of-band global fallback of the gather to be exercised."""
    monkeypatch.setenv("RGBD_B200_BAND", band)
    monkeypatch.setenv("RGBD_B200_BAND_TR", band_tr)
    g = load_golden(name)
    o = case_options(g)
    drv = _driver(g, o)
    parts, gi, gr = drv.fwd_bwd(gy=o["gy"])
    assert abs(parts[4] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert_grad_close(gi, g["g_img"])
    assert_grad_close(gr, g["g_img_rot"])
    parts2, _, _ = drv.fwd(want_zp=False, want_masks=False)
    np.testing.assert_allclose(parts2[:5], parts[:5], rtol=1e-6)
    gi2, gr2 = drv.bwd(gy=o["gy"])
    assert_grad_close(gi2, g["g_img"])
    assert_grad_close(gr2, g["g_img_rot"])


@pytest.mark.parametrize("name", ["loss_cfg0_l1_occ", "loss_s32_l2_noocc", "hinge_ffhq"])
@pytest.mark.parametrize("lags", [("2", "3"), ("1", "1"), ("40", "40")])
def test_pipeline_kernel_agrees_with_reference(name, lags, monkeypatch):
    """Opt-in single-launch pipeline kernel (RGBD_B200_MEGA=1): stage-in, main and stage-out tickets of one
    persistent launch ordered by per-pair dependency counters in the head of the workspace.  Same results as the
    three-kernel chain for any lag (1 = dependency waits are exercised, 40 = fully sequential phases), the
    control block is back in its rest state afterwards (repeated calls on one workspace), no wait timed out."""
    import ctypes
    from gpu_util import Consistency
    from oracle import numpy_port as npp
    from rgbd_gan_b200 import _lib
    monkeypatch.setenv("RGBD_B200_MEGA", "1")
    monkeypatch.setenv("RGBD_B200_MEGA_LAG_MAIN", lags[0])
    monkeypatch.setenv("RGBD_B200_MEGA_LAG_SO", lags[1])
    g = load_golden(name)
    hinge = name.startswith("hinge")
    if hinge:
        port = npp.LossFuncRotateNP(lambda_geometric=3)
        port.init_params(int(g["S"]))
        drv = Consistency(g["x"], g["cam"], int(g["B"]), port.K, port.inv_K, lam=3.0, occ=True)
        drv.opts.hinge_depth_min, drv.opts.hinge_lambda = float(g["depth_min"]), float(g["lambda_depth"])
        gy, want, slot = float(g["gy"]), float(g["total"]), 6
    else:
        o = case_options(g)
        drv = _driver(g, o)
        gy, want, slot = o["gy"], float(g["loss"]), 4
    for _ in range(2):
        parts, gi, gr = drv.fwd_bwd(gy=gy)
        assert abs(parts[slot] - want) <= 1e-5 * abs(want)
        assert_grad_close(gi, g["g_img"])
        assert_grad_close(gr, g["g_img_rot"])
    parts2, zp, _ = drv.fwd()
    if not hinge:
        np.testing.assert_array_equal(zp, g["new_zp_cat"])
    np.testing.assert_allclose(parts2[:7], parts[:7], rtol=1e-6, atol=1e-12)
    gi2, gr2 = drv.bwd(gy=1.0, gy_dev=gy)
    assert_grad_close(gi2, g["g_img"])
    assert_grad_close(gr2, g["g_img_rot"])
    status = ctypes.c_int(-1)
    _lib.call("rgbd_consistency_status", ctypes.c_void_p(drv.ws.data_ptr()), None, ctypes.byref(status))
    assert status.value == 0


@pytest.mark.parametrize("name", ["loss_s32_l2_feat", "loss_edge_c2"])
@pytest.mark.parametrize("wide", ["1", "0"])
def test_many_channel_kernels_agree_with_reference(name, wide, monkeypatch):
    """C != 4 (the feature-space loss of updater.py:345-354): the warp-per-pixel kernels (default) and the simple
    thread-per-pixel variant (RGBD_B200_WIDE=0) both reproduce the reference, incl. new_zp, masks and the two-pass path"""
    monkeypatch.setenv("RGBD_B200_WIDE", wide)
    g = load_golden(name)
    o = case_options(g)
    drv = _driver(g, o)
    parts, gi, gr = drv.fwd_bwd(gy=o["gy"])
    assert abs(parts[4] - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert_grad_close(gi, g["g_img"])
    assert_grad_close(gr, g["g_img_rot"])
    parts2, zp, masks = drv.fwd()
    np.testing.assert_array_equal(zp, g["new_zp_cat"])
    np.testing.assert_array_equal(masks[0].astype(bool), np.concatenate([g["not_out"], g["not_out_rot"]]))
    np.testing.assert_allclose(parts2[:5], parts[:5], rtol=1e-6)
    gi2, gr2 = drv.bwd(gy=1.0, gy_dev=o["gy"])
    assert_grad_close(gi2, g["g_img"])
    assert_grad_close(gr2, g["g_img_rot"])


def test_growing_sizes_reuse():
    g = load_golden("loss_growing")
    from gpu_util import Consistency
    B = int(g["B"])
    for S in (32, 64):
        drv = Consistency(g["x%d" % S], g["cam"], B, g["K_%d" % S], g["inv_K_%d" % S], lam=3.0, occ=True)
        parts, zp, _ = drv.fwd()
        np.testing.assert_array_equal(zp, g["new_zp_cat_%d" % S])
        assert abs(parts[4] - float(g["loss_%d" % S])) <= 1e-5 * abs(float(g["loss_%d" % S]))
        gi, gr = drv.bwd(gy=1.0)
        assert_grad_close(gi, g["g_img_%d" % S])
        assert_grad_close(gr, g["g_img_rot_%d" % S])


@pytest.mark.parametrize("S,B,depth", [(128, 256, "rough"), (128, 64, "smooth"), (256, 32, "rough")])
def test_full_size_against_oracle_and_properties(S, B, depth, oracle_mod):
    """BASELINE.json full sizes (batch 256 at 128^2, 256^2): direct comparison with the OpenMP C
    oracle plus size-independent properties: linearity in gy, batch-permutation equivariance,
    fused == two-pass, run-to-run reproducibility bound of the atomics."""
    from gpu_util import Consistency
    from oracle import numpy_port as poses
    from oracle import numpy_port as npp
    x, cam = poses.synthetic_batch(B, S, depth=depth, seed=11)
    port = npp.LossFuncRotateNP(lambda_geometric=3)
    port.init_params(S)
    drv = Consistency(x, cam, B, port.K, port.inv_K, lam=3.0, occ=True)
    M, c, Mi, ci = drv.host_poses
    ref_parts, d = oracle_mod.consistency_fwd(x[:B], x[B:], M, c, Mi, ci, norm=1, occlusion=True, debug=True)
    parts, zp, masks = drv.fwd()
    np.testing.assert_array_equal(zp, d["new_zp"])
    np.testing.assert_array_equal(masks[0], d["mask"])
    np.testing.assert_array_equal(masks[1], d["occ"])
    np.testing.assert_allclose(parts[:4], ref_parts, rtol=1e-5)
    ref_gi, ref_gr = oracle_mod.consistency_bwd(x[:B], x[B:], M, c, Mi, ci, norm=1, occlusion=True,
                                                lambda_geometric=3, gy=2.0)
    gi, gr = drv.bwd(gy=2.0)
    assert_grad_close(gi, ref_gi)
    assert_grad_close(gr, ref_gr)
    # linearity in the upstream gradient (x2 is exact in fp32 up to atomic ordering)
    gi1, gr1 = drv.bwd(gy=1.0)
    assert np.abs(gi - 2 * gi1).max() <= 1e-6 * np.abs(gi).max()
    # run-to-run nondeterminism of the fp32 RED scatter stays far below the tolerance
    gi_b, gr_b = drv.bwd(gy=2.0)
    assert np.abs(gi - gi_b).max() <= 1e-6 * np.abs(gi).max()
    assert np.abs(gr - gr_b).max() <= 1e-6 * np.abs(gr).max()
    # fused one-pass == two-pass
    parts_f, gi_f, gr_f = drv.fwd_bwd(gy=2.0)
    np.testing.assert_allclose(parts_f[:5], parts[:5], rtol=1e-6)      # (fwd with outputs and the fused call run different kernels)
    assert np.abs(gi - gi_f).max() <= 1e-6 * np.abs(gi).max()
    # batch permutation equivariance
    perm = np.random.default_rng(0).permutation(B)
    xp_ = np.concatenate([x[:B][perm], x[B:][perm]])
    camp = np.concatenate([cam[:B][perm], cam[B:][perm]])
    drv_p = Consistency(xp_, camp, B, port.K, port.inv_K, lam=3.0, occ=True)
    parts_p, gi_p, gr_p = drv_p.fwd_bwd(gy=2.0)
    np.testing.assert_allclose(parts_p[:5], parts[:5], rtol=1e-6)
    assert np.abs(gi_p - gi[perm]).max() <= 1e-6 * np.abs(gi).max()
    assert np.abs(gr_p - gr[perm]).max() <= 1e-6 * np.abs(gr).max()


def test_identity_pose_closed_form():
    """Zero rotation for both cameras and power-of-two depths make every operation of the recipe
    exact: K R K^-1 = I, (z*x)/z = x, so u0 = row, v0 = col, the sampled value is img_rot[b,:,row,col]
    and the in-bounds mask is exactly the interior (strict `< H-1`).  The loss then has a closed
    form that torch evaluates at full 128^2 size."""
    from gpu_util import Consistency, DEV
    from oracle import numpy_port as poses
    from oracle import numpy_port as npp
    B, S = 16, 128
    x, _ = poses.synthetic_batch(B, S, depth="rough", seed=5)
    rng = np.random.default_rng(5)
    x[:, -1] = rng.choice(np.array([0.5, 1.0, 2.0], np.float32), size=(2 * B, S, S))
    cam = poses.get_camera_matries(np.zeros((2 * B, 6), np.float32))
    port = npp.LossFuncRotateNP(lambda_geometric=3)
    port.init_params(S)
    drv = Consistency(x, cam, B, port.K, port.inv_K, lam=3.0, occ=False)
    parts, zp, masks = drv.fwd()
    inner = np.zeros((S, S), bool)
    inner[:S - 1, :S - 1] = True
    np.testing.assert_array_equal(masks[0].reshape(2 * B, S, S).astype(bool), np.broadcast_to(inner, (2 * B, S, S)))
    a, b = torch.from_numpy(x[:B]).to(DEV).double(), torch.from_numpy(x[B:]).to(DEV).double()
    sel = torch.from_numpy(inner).to(DEV)[None, None]
    N = B * S * S
    rgb = ((b[:, :3] - a[:, :3]).abs() * sel).sum().item() / (N * 3)
    dep = ((b[:, 3:] - a[:, 3:]).abs() * sel).sum().item() / N
    np.testing.assert_allclose(parts[:4], [rgb, rgb, dep, dep], rtol=2e-6)
    # gradient w.r.t. own colour is -sign(diff)/(3N) on the interior, plus +sign(diff)/(3N) scattered back
    # from the other direction onto the same pixel: closed form = 2 * sign(a - b) / (3N) * gy
    gi, gr = drv.bwd(gy=1.0)
    ref = (2.0 * torch.sign(a[:, :3] - b[:, :3]) * sel / (3 * N)).cpu().numpy()
    assert np.abs(gi[:, :3] - ref).max() <= 1e-6 * np.abs(ref).max()
    assert np.abs(gr[:, :3] + ref).max() <= 1e-6 * np.abs(ref).max()


def test_argument_validation_on_device():
    from gpu_util import Consistency, p, stream
    from rgbd_gan_b200 import _lib
    g = load_golden("loss_edge_c2")
    o = case_options(g)
    drv = _driver(g, o)
    parts = torch.zeros(8, device="cuda:0")
    lib = _lib.load()
    rc = lib.rgbd_consistency_fwd(*drv._common(), p(parts), None, None, p(drv.ws), 16, stream())
    assert rc == -3 and b"workspace" in lib.rgbd_last_error()
    mis = torch.zeros(drv.img.numel() + 1, device="cuda:0")[1:]
    args = drv._common()
    args[0] = p(mis)
    rc = lib.rgbd_consistency_fwd(*args, p(parts), None, None, p(drv.ws), drv.ws.numel(), stream())
    assert rc == -2


@pytest.mark.parametrize("name", ["hinge_ffhq", "hinge_car"])
def test_depth_hinge_fused_next_row(name, oracle_mod):
    """SURVEY 8(f) rank 2, fused into the staging kernels: loss_rotate += mean(relu(depth_min - depth)^2) * lambda
    (updater.py:357-359); against the reference's golden total / gradients and the C oracle."""
    from gpu_util import Consistency
    from oracle import numpy_port as npp
    g = load_golden(name)
    B, S, gy = int(g["B"]), int(g["S"]), float(g["gy"])
    dmin, lam = float(g["depth_min"]), float(g["lambda_depth"])
    port = npp.LossFuncRotateNP(lambda_geometric=3)
    port.init_params(S)
    drv = Consistency(g["x"], g["cam"], B, port.K, port.inv_K, lam=3.0, occ=True)
    drv.opts.hinge_depth_min, drv.opts.hinge_lambda = dmin, lam
    parts, gi, gr = drv.fwd_bwd(gy=gy)
    assert abs(parts[4] - float(g["loss_rotate"])) <= 1e-5 * abs(float(g["loss_rotate"]))
    assert abs(parts[5] - float(g["hinge"])) <= 1e-5 * max(abs(float(g["hinge"])), 1e-12)
    assert abs(parts[6] - float(g["total"])) <= 1e-5 * abs(float(g["total"]))
    assert_grad_close(gi, g["g_img"])
    assert_grad_close(gr, g["g_img_rot"])
    parts2, _, _ = drv.fwd(want_zp=False, want_masks=False)
    np.testing.assert_allclose(parts2[:7], parts[:7], rtol=1e-6, atol=1e-12)
    gi2, gr2 = drv.bwd(gy=1.0, gy_dev=gy)
    assert_grad_close(gi2, g["g_img"])
    assert_grad_close(gr2, g["g_img_rot"])
    # sharded denominators + several chunks
    drv.opts.n_pairs_global = 2 * B
    parts3, gi3, _ = drv.fwd_bwd(gy=gy)
    np.testing.assert_allclose(parts3[5] * 2, parts[5], rtol=1e-6, atol=1e-12)
    assert_grad_close(gi3 * 2, g["g_img"])
