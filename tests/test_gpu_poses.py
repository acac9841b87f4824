"""the device pose pipeline (csrc/poses.cu) against the reference's host path.

Bit-exact: thetas for replayed np.random draws, cam2world when the caller supplies NumPy's cos / sin, the pose constants
M, c, Mi, ci from the golden cam2world matrices (and, through them, new_zp and the loss).  Tolerance (1 ulp of the
rotation entries) only where the kernel evaluates cos / sin itself."""
import glob
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, LOSS_CASES, case_options, load_golden
from oracle import numpy_port as npp
from rgbd_gan_b200.host_math import intrinsics_for_size, pose_algebra

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cfg(ranges, uniform):
    return types.SimpleNamespace(x_rotate=ranges[0], y_rotate=ranges[1], z_rotate=ranges[2], x_translate=ranges[3],
                                 y_translate=ranges[4], z_translate=ranges[5], uniform_distribution=uniform)


def _replay(seed, B):
    """the raw draws of CameraParamPrior.sample, in the reference's order (train_rgbd.py:200-202)"""
    np.random.seed(seed)
    u = np.random.uniform(-1, 1, size=(B, 6))
    e = np.random.uniform(0, 0.5, size=(B, 6))
    s = np.random.choice(2, size=(B, 3))
    return np.concatenate([u, e, s.astype(np.float64)], axis=1)


@pytest.mark.parametrize("ranges", [npp.FFHQ_RANGES, npp.CAR_RANGES, (0.5, 3.1415, 0.2, 0.1, 0.05, 0.3)])
@pytest.mark.parametrize("uniform", [False, True])
def test_sample_replays_the_reference_bit_for_bit(ranges, uniform):
    from rgbd_gan_b200.pose_pipeline import CameraParamPrior
    B = 257
    np.random.seed(11)
    want = npp.sample_camera_prior(2 * B, ranges, uniform)
    got = CameraParamPrior(_cfg(ranges, uniform), DEV).sample(2 * B, draws=_replay(11, B)).cpu().numpy()
    assert got.dtype == np.float32 and got.shape == (2 * B, 6)
    np.testing.assert_array_equal(got, want)


def test_sample_philox_distribution():
    """own generator: the reference's distribution (ranges, |theta2 - theta| <= 0.5 * limit, sign rule), new numbers
    every call, reproducible for a seed"""
    from rgbd_gan_b200.pose_pipeline import CameraParamPrior
    B = 4096
    pr = CameraParamPrior(_cfg(npp.CAR_RANGES, False), DEV, seed=5)
    a, b = pr.sample(2 * B).cpu().numpy(), pr.sample(2 * B).cpu().numpy()
    again = CameraParamPrior(_cfg(npp.CAR_RANGES, False), DEV, seed=5).sample(2 * B).cpu().numpy()
    np.testing.assert_array_equal(a, again)
    assert not np.array_equal(a, b)
    rng = np.asarray(npp.CAR_RANGES, np.float32)
    t1, t2 = a[:B], a[B:]
    # first views inside the range; the non-uniform prior does not reflect the second view back (train_rgbd.py:205-212)
    assert (np.abs(t1) <= rng[None] * (1 + 1e-6)).all() and (np.abs(t2) <= rng[None] * 1.5 + 1e-6).all() and (a[:, 2:] == 0).all()
    uni = CameraParamPrior(_cfg(npp.CAR_RANGES, True), DEV, seed=5).sample(2 * B).cpu().numpy()
    assert (np.abs(uni) <= rng[None] * (1 + 1e-6)).all()             # the uniform prior reflects into the range
    for k in (0, 1):
        u = t1[:, k] / rng[k]
        assert abs(u.mean()) < 0.05 and abs(u.std() - 1 / np.sqrt(3)) < 0.02          # U(-1, 1)
        lim = min(1.0, 1.0 / (rng[k] + 1e-8))
        d = (t2[:, k] - t1[:, k]) / rng[k]
        assert (np.abs(d) <= 0.5 * lim + 1e-6).all() and np.abs(d).max() > 0.45 * lim
    # x rotation (range != 3.1415): theta2 moves towards 0; y rotation (== 3.1415): either direction
    assert (np.abs(t2[:, 0]) <= np.abs(t1[:, 0]) + 1e-7).mean() > 0.8
    away = (np.abs(t2[:, 1]) > np.abs(t1[:, 1])).mean()
    assert 0.3 < away < 0.7


def _golden_files():
    return [f for f in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))) if "thetas" in np.load(f).files]


@pytest.mark.parametrize("path", _golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def test_camera_matrices_against_every_golden(path):
    from rgbd_gan_b200.pose_pipeline import get_camera_matries
    g = np.load(path)
    th, cam = g["thetas"], g["cam"]
    cs = np.concatenate([np.cos(th[:, :3]), np.sin(th[:, :3])], axis=1).astype(np.float32)
    t = torch.from_numpy(th).to(DEV)
    got = get_camera_matries(t, cos_sin=torch.from_numpy(cs).to(DEV)).cpu().numpy()
    if np.array_equal(npp.get_camera_matries(th), cam):      # this host's NumPy cos / sin are the golden host's
        np.testing.assert_array_equal(got, cam)
    else:
        np.testing.assert_allclose(got, cam, rtol=0, atol=5e-7)
    own = get_camera_matries(t).cpu().numpy()                # cos / sin on the device: one rounding of the exact value
    np.testing.assert_allclose(own, cam, rtol=0, atol=5e-7)         # three factors, each entry within 1 ulp of 1
    assert np.array_equal(own[:, 3], cam[:, 3])                     # bottom row (0, 0, 0, 1) exactly


def test_camera_matrices_order_argument():
    from rgbd_gan_b200.pose_pipeline import get_camera_matries
    th = np.random.default_rng(2).uniform(-1, 1, size=(9, 6)).astype(np.float32)
    for order in ((0, 1, 2), (2, 1, 0), (1, 0, 2)):
        cs = np.concatenate([np.cos(th[:, :3]), np.sin(th[:, :3])], axis=1).astype(np.float32)
        got = get_camera_matries(torch.from_numpy(th).to(DEV), order=order, cos_sin=torch.from_numpy(cs).to(DEV)).cpu().numpy()
        np.testing.assert_allclose(got, npp.get_camera_matries(th, order), rtol=0, atol=1e-6)


@pytest.mark.parametrize("case", LOSS_CASES)
def test_pose_algebra_bit_exact_on_golden(case):
    """M, c, Mi, ci from the device kernel == the reference's NumPy matmul sequence on the golden cam2world matrices,
    checked (i) against this host's NumPy when its BLAS orders are the golden host's, (ii) always through the kernels:
    new_zp from device poses == golden new_zp bit for bit"""
    from rgbd_gan_b200.loss_functions import LossFuncRotate, unpack_poses
    from rgbd_gan_b200.pose_pipeline import pose_algebra_device
    g = load_golden(case)
    o = case_options(g)
    B, S = o["B"], o["S"]
    K, inv_K = intrinsics_for_size(o["K"], S, first=True)
    np.testing.assert_array_equal(K, g["K"])
    cam = torch.from_numpy(g["cam"]).to(DEV)
    poses = pose_algebra_device(K, inv_K, cam[:B], cam[B:])
    got = [a.cpu().numpy() for a in unpack_poses(poses, B)]
    host = pose_algebra(K, inv_K, g["cam"][:B], g["cam"][B:])
    for a, b in zip(got, host):
        np.testing.assert_allclose(a, b, rtol=3e-7, atol=1e-7)
    x = torch.from_numpy(g["x"]).to(DEV)
    f = LossFuncRotate(None, K=o["K"], norm=o["norm"], lambda_geometric=o["lam"])
    loss, zp = f(x[:B], cam[:B], x[B:], cam[B:], occlusion_aware=o["occ"], max_depth=o["max_depth"], min_depth=o["min_depth"])
    np.testing.assert_array_equal(zp.cpu().numpy(), g["new_zp_cat"])          # device thetas -> device pose kernel
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    # the fused call: poses handed over directly
    loss2, zp2 = f(x[:B], None, x[B:], None, occlusion_aware=o["occ"], max_depth=o["max_depth"], min_depth=o["min_depth"],
                   poses=poses)
    assert torch.equal(zp, zp2) and loss2.item() == loss.item()


def test_pipeline_one_launch_matches_the_stages():
    from rgbd_gan_b200 import _lib
    from rgbd_gan_b200.pose_pipeline import CameraParamPrior, PosePipeline, get_camera_matries, pose_algebra_device
    B, S = 64, 128
    K, inv_K = intrinsics_for_size(None, S, first=True)
    draws = _replay(3, B)
    pr = CameraParamPrior(_cfg(npp.FFHQ_RANGES, False), DEV)
    n0 = _lib.load().rgbd_launch_count()
    thetas, cam, poses = PosePipeline(pr, K, inv_K).step(2 * B, draws=draws)
    assert _lib.load().rgbd_launch_count() - n0 == 1
    np.random.seed(3)
    want_th = npp.sample_camera_prior(2 * B, npp.FFHQ_RANGES, False)
    np.testing.assert_array_equal(thetas.cpu().numpy(), want_th)
    assert torch.equal(cam, get_camera_matries(thetas))
    assert torch.equal(poses, pose_algebra_device(K, inv_K, cam[:B], cam[B:]))
    # against the reference's host chain: cos / sin differ by <= 1 ulp, everything downstream is continuous in them
    M, c, Mi, ci = pose_algebra(K, inv_K, *np.split(npp.get_camera_matries(want_th), 2))
    got = poses.cpu().numpy()
    np.testing.assert_allclose(got[:9 * B].reshape(B, 3, 3), M, rtol=0, atol=2e-4)     # entries up to 2 S = 256
    np.testing.assert_allclose(got[9 * B:12 * B].reshape(B, 3, 1), c, rtol=0, atol=2e-4)


def test_bad_arguments_raise():
    from rgbd_gan_b200 import _lib
    from rgbd_gan_b200.pose_pipeline import get_camera_matries
    with pytest.raises(TypeError):
        get_camera_matries(torch.zeros(4, 6))
    with pytest.raises(_lib.RgbdB200Error):
        get_camera_matries(torch.zeros(4, 6, device=DEV), order=(0, 1, 3))
