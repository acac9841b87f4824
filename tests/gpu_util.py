"""helpers for the -m gpu tests: call the C-ABI with torch-owned device memory"""
import ctypes

import numpy as np
import torch

from rgbd_gan_b200 import _lib
from rgbd_gan_b200.loss_functions import pose_algebra

DEV = "cuda:0"


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV).to(dtype).contiguous()


def p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Consistency:
    """thin driver of rgbd_consistency_* for a fixed problem"""

    def __init__(self, x, cam, B, K, inv_K, norm="l1", lam=3.0, occ=False, max_depth=None, min_depth=None,
                 n_pairs_global=0):
        self.B, self.C, self.H, self.W = B, x.shape[1], x.shape[2], x.shape[3]
        self.img, self.img_rot = dev(x[:B]), dev(x[B:])
        M, c, Mi, ci = pose_algebra(K, inv_K, cam[:B], cam[B:])
        self.host_poses = (M, c, Mi, ci)
        self.M, self.c, self.Mi, self.ci = dev(M), dev(c), dev(Mi), dev(ci)
        self.opts = _lib.LossOpts(1 if norm == "l1" else 2, int(occ),
                                  float("nan") if max_depth is None else max_depth,
                                  float("nan") if min_depth is None else min_depth, lam, n_pairs_global)
        nbytes = _lib.load().rgbd_consistency_workspace_bytes(self.B, self.C, self.H, self.W)
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)

    def _common(self):
        return [p(self.img), p(self.img_rot), p(self.M), p(self.c), p(self.Mi), p(self.ci), self.B, self.C, self.H,
                self.W, ctypes.byref(self.opts)]

    def fwd(self, want_zp=True, want_masks=True):
        N2 = 2 * self.B * self.H * self.W
        parts = torch.full((8,), float("nan"), device=DEV)
        zp = torch.empty((2 * self.B, self.H * self.W, 3), device=DEV) if want_zp else None
        masks = torch.empty((2, N2), dtype=torch.uint8, device=DEV) if want_masks else None
        _lib.call("rgbd_consistency_fwd", *self._common(), p(parts), p(zp), p(masks), p(self.ws), self.ws.numel(),
                  stream())
        torch.cuda.synchronize()
        return parts.cpu().numpy(), None if zp is None else zp.cpu().numpy(), None if masks is None else masks.cpu().numpy()

    def bwd(self, gy=1.0, gy_dev=None, g_new_zp=None):
        g_img = torch.full_like(self.img, float("nan"))
        g_rot = torch.full_like(self.img_rot, float("nan"))
        gyd = None if gy_dev is None else torch.tensor([gy_dev], dtype=torch.float32, device=DEV)
        gz = None if g_new_zp is None else dev(g_new_zp)
        _lib.call("rgbd_consistency_bwd", *self._common(), ctypes.c_float(gy), p(gyd), p(gz), p(g_img), p(g_rot),
                  p(self.ws), self.ws.numel(), stream())
        torch.cuda.synchronize()
        return g_img.cpu().numpy(), g_rot.cpu().numpy()

    def fwd_bwd(self, gy=1.0, want_zp=False):
        parts = torch.full((8,), float("nan"), device=DEV)
        g_img = torch.full_like(self.img, float("nan"))
        g_rot = torch.full_like(self.img_rot, float("nan"))
        zp = torch.empty((2 * self.B, self.H * self.W, 3), device=DEV) if want_zp else None
        _lib.call("rgbd_consistency_fwd_bwd", *self._common(), ctypes.c_float(gy), p(parts), p(zp), p(g_img), p(g_rot),
                  p(self.ws), self.ws.numel(), stream())
        torch.cuda.synchronize()
        return parts.cpu().numpy(), g_img.cpu().numpy(), g_rot.cpu().numpy()
