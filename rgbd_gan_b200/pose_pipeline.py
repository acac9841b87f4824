"""The pose pipeline on the device: torch glue over rgbd_pose_* (csrc/poses.cu).

Mirrors the reference's host-side producers of the camera poses -- `CameraParamPrior` (train_rgbd.py:192-217) and
`get_camera_matries` (updater.py:45-60) -- with CUDA tensors in and out, and adds the pose algebra of
LossFuncRotate.__call__ (common/loss_functions.py:85-91, :174, :181) as a device call, so that a step needs neither
np.random / NumPy work on the host nor an upload.  `LossFuncRotate` uses `pose_algebra_device` by itself whenever the
cam2world matrices it is given are CUDA tensors.  No CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import PosePrior


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f9(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32)[:3, :3]).reshape(9)
    return (ctypes.c_float * 9)(*a.tolist())


def _order(order):
    return None if order is None else (ctypes.c_int * 3)(*[int(o) for o in order])


def _dev_rows(t, cols, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (rgbd_gan_b200 has no CPU path)" % what)
    t = t.to(torch.float32).contiguous()
    if t.dim() < 2 or tuple(t.shape[1:]) != cols:
        raise ValueError("%s must be (N,%s)" % (what, ",".join(str(c) for c in cols)))
    return t


class CameraParamPrior:
    """train_rgbd.py:192-217 on the device.  `config` carries x_rotate, y_rotate, z_rotate, x_translate, y_translate,
    z_translate and uniform_distribution like the reference's; `sample(batch_size)` returns the (batch_size, 6) float32
    thetas (first half: first views, second half: rotated views) as a CUDA tensor.

    draws=None: Philox4x32-10 keyed by (seed, call counter, pair) -- the reference's distribution, not its np.random
    stream.  draws=(B,15) float64 [uniform(-1,1) x6 | uniform(0,0.5) x6 | choice(2) x3] per pair replays a host stream:
    the thetas are then bit-identical to the reference's for the same draws."""

    def __init__(self, config, device="cuda", seed=0):
        rng = [config.x_rotate, config.y_rotate, config.z_rotate, config.x_translate, config.y_translate, config.z_translate]
        self.camera_param_range = np.array(rng, dtype=np.float64)
        self.rotation_range = self.camera_param_range[:3]
        self.uniform = bool(config.uniform_distribution)
        self.device = torch.device(device)
        self.seed, self.step = int(seed), 0
        self.c_prior = PosePrior((ctypes.c_double * 6)(*self.camera_param_range.tolist()), int(self.uniform))

    def _draws(self, draws, B):
        if draws is None:
            return None
        d = torch.as_tensor(np.ascontiguousarray(draws, dtype=np.float64) if not isinstance(draws, torch.Tensor) else draws)
        d = d.to(self.device, torch.float64).contiguous()
        if tuple(d.shape) != (B, 15):
            raise ValueError("draws must be (batch_size // 2, 15)")
        return d

    def sample(self, batch_size, draws=None):
        B = batch_size // 2
        thetas = torch.empty((2 * B, 6), dtype=torch.float32, device=self.device)
        d = self._draws(draws, B)
        with torch.cuda.device(self.device):
            _lib.call("rgbd_pose_sample", ctypes.byref(self.c_prior), B, _ptr(d), self.seed, self.step, _ptr(thetas),
                      _stream(self.device))
        self.step += 1
        return thetas


def get_camera_matries(thetas, order=(0, 1, 2), cos_sin=None):
    """updater.py:45-60: (N,6) thetas -> (N,4,4) cam2world, CUDA tensors.  cos_sin (N,6) = [cos | sin] of the three
    angles as the caller computed them makes the result bit-identical to the reference's; without it they are evaluated
    in double and rounded once."""
    thetas = _dev_rows(thetas, (6,), "thetas")
    cs = None if cos_sin is None else _dev_rows(cos_sin, (6,), "cos_sin")
    cam = torch.empty((thetas.shape[0], 4, 4), dtype=torch.float32, device=thetas.device)
    with torch.cuda.device(thetas.device):
        _lib.call("rgbd_pose_camera_matrices", _ptr(thetas), _ptr(cs), thetas.shape[0], _order(order), _ptr(cam),
                  _stream(thetas.device))
    return cam


def pose_algebra_device(K, inv_K, theta, theta_rot, out=None):
    """common/loss_functions.py:85-91 + the constants of :174 / :181 from CUDA cam2world matrices (B,4,4).
    Returns the packed (24*B,) tensor [M | c | Mi | ci] that the loss entry points take; no host synchronisation."""
    theta, theta_rot = _dev_rows(theta, (4, 4), "theta"), _dev_rows(theta_rot, (4, 4), "theta_rot")
    B = theta.shape[0]
    if theta_rot.shape[0] != B:
        raise ValueError("theta and theta_rot must hold the same number of poses")
    poses = out if out is not None else torch.empty(24 * B, dtype=torch.float32, device=theta.device)
    base = poses.data_ptr()
    ptrs = [ctypes.c_void_p(base + 4 * off) for off in (0, 9 * B, 12 * B, 21 * B)]
    with torch.cuda.device(theta.device):
        _lib.call("rgbd_pose_algebra", _ptr(theta), _ptr(theta_rot), B, _f9(K), _f9(inv_K), *ptrs, _stream(theta.device))
    return poses


def unpack_poses(poses, B):
    """packed (24*B,) -> M (B,3,3), c (B,3,1), Mi (B,3,3), ci (B,3,1) views"""
    from .loss_functions import unpack_poses as _u
    return _u(poses, B)


class PosePipeline:
    """sample -> cam2world -> pose constants in ONE launch (rgbd_pose_pipeline).  `step(batch_size)` returns
    (thetas (2B,6), cam2world (2B,4,4), poses (24B,) packed); the packed poses go straight into
    LossFuncRotate.__call__(..., poses=...)."""

    def __init__(self, prior, K, inv_K, order=(0, 1, 2)):
        self.prior, self.order = prior, order
        self.K, self.inv_K = np.array(K, dtype=np.float32), np.array(inv_K, dtype=np.float32)

    def step(self, batch_size, draws=None):
        pr, B = self.prior, batch_size // 2
        dev = pr.device
        thetas = torch.empty((2 * B, 6), dtype=torch.float32, device=dev)
        cam = torch.empty((2 * B, 4, 4), dtype=torch.float32, device=dev)
        poses = torch.empty(24 * B, dtype=torch.float32, device=dev)
        base = poses.data_ptr()
        ptrs = [ctypes.c_void_p(base + 4 * off) for off in (0, 9 * B, 12 * B, 21 * B)]
        d = pr._draws(draws, B)
        with torch.cuda.device(dev):
            _lib.call("rgbd_pose_pipeline", ctypes.byref(pr.c_prior), B, _ptr(d), pr.seed, pr.step, _order(self.order),
                      _f9(self.K), _f9(self.inv_K), _ptr(thetas), _ptr(cam), *ptrs, _stream(dev))
        pr.step += 1
        return thetas, cam, poses


__all__ = ["CameraParamPrior", "get_camera_matries", "pose_algebra_device", "unpack_poses", "PosePipeline"]
