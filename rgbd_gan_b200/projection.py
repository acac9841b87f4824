"""Host-side mirror of deepvoxel/projection.py (ProjectionHelper) and
deepvoxel/deepvoxel.py:388-433 (interpolate_trilinear, MakeSlice) over librgbdgan_b200.so.

`ProjectionHelper.project()` is the B200-first entry: the whole batch's
compute_proj_idcs + interpolate_trilinear in one launch, with no index lists, no stream
compaction and no host sync (what deepvoxels_generator.py:287-299 -> deepvoxel.py:879-884
does with two per-sample Python loops).  The reference-shaped functions are kept for
drop-in use and produce identical values.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import DvParams
from .loss_functions import _as_numpy, _dev_f32, _ptr, _stream


def _cam_to_device(cam2world, device):
    if isinstance(cam2world, torch.Tensor) and cam2world.is_cuda:
        return cam2world.to(torch.float32).contiguous()
    a = np.ascontiguousarray(_as_numpy(cam2world), dtype=np.float32)
    return torch.from_numpy(a).to(device)


def _project_ws(params, B, F, device):
    nbytes = _lib.load().rgbd_dv_project_workspace_bytes(ctypes.byref(params), B, F)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


class _ProjectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid, cam, params):
        B, F = grid.shape[:2]
        out = torch.empty((B, F, params.D, params.H, params.W), dtype=torch.float32, device=grid.device)
        ws = _project_ws(params, B, F, grid.device)          # channels-last staging copy of a chunk of grids
        _lib.call("rgbd_dv_project_fwd", ctypes.byref(params), _ptr(grid), _ptr(cam), B, F, _ptr(out), _ptr(ws),
                  ws.numel(), _stream())
        ctx.cam, ctx.params, ctx.gshape = cam, params, grid.shape
        return out

    @staticmethod
    def backward(ctx, g_out):
        B, F = ctx.gshape[:2]
        g_out = g_out.to(torch.float32).contiguous()
        g_grid = torch.empty(ctx.gshape, dtype=torch.float32, device=g_out.device)
        ws = _project_ws(ctx.params, B, F, g_out.device)
        _lib.call("rgbd_dv_project_bwd", ctypes.byref(ctx.params), _ptr(g_out), _ptr(ctx.cam), B, F, _ptr(g_grid),
                  _ptr(ws), ws.numel(), _stream())
        return g_grid, None, None


class ProjectionHelper:
    """deepvoxel/projection.py:5-39 -- same constructor arguments and attributes."""

    def __init__(self, lifting_intrinsic, projection_intrinsic, projection_image_dims, lifting_image_dims,
                 depth_min, depth_max, grid_dims, voxel_size, near_plane, frustrum_depth, device=None,
                 verbose=True):
        self.grid_dims = grid_dims
        self.projection_intrinsic = projection_intrinsic
        self.lifting_intrinsic = lifting_intrinsic
        self.depth_min = depth_min
        self.depth_max = depth_max
        self.projection_image_dims = projection_image_dims
        self.lifting_image_dims = lifting_image_dims
        self.voxel_size = voxel_size
        self.device = device
        self.near_plane = near_plane
        self.frustrum_depth = frustrum_depth
        if not (grid_dims[0] == grid_dims[1] == grid_dims[2]):
            raise ValueError("cubic grids only (the reference clamps every axis with a different dim, "
                             "deepvoxel.py:390,406-408)")
        if verbose:                                     # the reference prints this banner (:31-39)
            print("\n" + "*" * 100)
            print("Lifting intrinsic is %s" % self.lifting_intrinsic)
            print("Projection intrinsic is %s" % self.projection_intrinsic)
            print("Lifting image dims is ", self.lifting_image_dims)
            print("Projection image dims is ", self.projection_image_dims)
            print("voxel size is %s" % self.voxel_size)
            print("*" * 100 + "\n")
        self._ws = None

    def params(self):
        K = self.projection_intrinsic
        return DvParams(int(self.projection_image_dims[0]), int(self.projection_image_dims[1]),
                        int(self.frustrum_depth), int(self.grid_dims[2]), float(K[0][0]), float(K[1][1]),
                        float(K[0][2]), float(K[1][2]), float(np.float32(self.voxel_size)),
                        float(np.float32(self.near_plane)))

    def compute_proj_idcs(self, cam2world, grid2world=None):
        """projection.py:48-105 -> (lin_ind_frustrum int32 (M,), voxel_coords fp32 (3,M)) or None."""
        dev = torch.device("cuda", torch.cuda.current_device()) if self.device is None else torch.device(self.device)
        if isinstance(cam2world, torch.Tensor) and cam2world.is_cuda:
            dev = cam2world.device
        cam = _cam_to_device(cam2world, dev).reshape(16)
        P = self.params()
        n = P.W * P.H * P.D
        lin = torch.empty(n, dtype=torch.int32, device=dev)
        vc = torch.empty((3, n), dtype=torch.float32, device=dev)
        if self._ws is None or self._ws.device != dev:
            self._ws = torch.empty(_lib.load().rgbd_dv_workspace_bytes(ctypes.byref(P)), dtype=torch.uint8, device=dev)
        M = ctypes.c_int(0)
        with torch.cuda.device(dev):
            if grid2world is not None:
                # :53-54 world2grid = xp.linalg.inv(grid2world); :83-84 a SECOND product per element, evaluated on the device
                # in the reference's order (folding the two matrices on the host changes the last bits of the coordinates)
                w2g = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(_as_numpy(grid2world)), dtype=np.float32)).to(dev)
                _lib.call("rgbd_dv_compute_proj_idcs_g2w", ctypes.byref(P), _ptr(cam), _ptr(w2g.reshape(16)), _ptr(lin), _ptr(vc),
                          ctypes.byref(M), _ptr(self._ws), self._ws.numel(), _stream())
            else:
                _lib.call("rgbd_dv_compute_proj_idcs", ctypes.byref(P), _ptr(cam), _ptr(lin), _ptr(vc), ctypes.byref(M),
                          _ptr(self._ws), self._ws.numel(), _stream())
        if M.value == 0:
            print('error: nothing in frustum bounds')   # :98-100
            return None
        return lin[:M.value], vc[:, :M.value]

    def project(self, grid, cam2world):
        """Fused batch path: grid (B,F,G,G,G), cam2world (B,4,4) -> frustum (B,F,D,H,W); differentiable in grid."""
        grid = _dev_f32(grid, "grid")
        cam = _cam_to_device(cam2world, grid.device).reshape(grid.shape[0], 16)
        return _ProjectFn.apply(grid, cam, self.params())

    def render_accumulative(self, grid, cam2world, W1, b1, W2, b2, accmulative_threshold=4,
                            return_foreground_weight=False):
        """"Next" row: what DeepVoxels.forward does per sample with `occlusion_type:
        accumulative` (deepvoxel.py:879-892 + AccumulativeOcclusionNet.forward :574-587 + depth rescale :903-904),
        for the whole batch in one fused pass that never materialises the (B,F,D,H,W) view volume.
        grid (B,F,G,G,G); cam2world (B,4,4); W1 (nf,F+1), b1 (nf,), W2 (1,nf), b2 (1,): the `.c.W` / `.c.b` of
        the two EqualizedConv3d layers of `occlusion_net.occlusion` (1x1x1 kernels squeezed; input channel 0 of
        W1 is the depth coordinate, :575).  Returns (novel_views (B,F,H,W), depth_maps (B,1,H,W)[, foreground
        weight (B,1,H,W)]); differentiable in grid and the four parameter arrays."""
        grid = _dev_f32(grid, "grid")
        B, F = grid.shape[:2]
        cam = _cam_to_device(cam2world, grid.device).reshape(B, 16)
        W1, b1, W2, b2 = (_dev_f32(t, n).contiguous() for t, n in ((W1, "W1"), (b1, "b1"), (W2, "W2"), (b2, "b2")))
        nf = W1.shape[0]
        if W1.shape[1] != F + 1 or W2.numel() != nf or b1.numel() != nf or b2.numel() != 1:
            raise ValueError("W1 must be (nf, F+1), b1 (nf,), W2 (1, nf), b2 (1,)")
        G = int(self.grid_dims[-1])
        rp = _lib.DvRenderParams(int(nf), int(np.ceil(np.sqrt(3) * G)), float(accmulative_threshold),
                                 float(np.float32(np.sqrt(2) * np.sqrt(1.0 / (F + 1)))),       # pggan.py:31 (ksize 1)
                                 float(np.float32(np.sqrt(2) * np.sqrt(1.0 / nf))))
        novel, depth, fg = _RenderFn.apply(grid, W1, b1, W2, b2, cam, self.params(), rp)
        return (novel, depth, fg) if return_foreground_weight else (novel, depth)


class _RenderFn(torch.autograd.Function):
    """fused projection + accumulative render tail (rgbd_dv_render_fwd / _bwd)"""

    @staticmethod
    def forward(ctx, grid, W1, b1, W2, b2, cam, params, rparams):
        B, F = grid.shape[:2]
        dev = grid.device
        novel = torch.empty((B, F, params.H, params.W), dtype=torch.float32, device=dev)
        depth = torch.empty((B, 1, params.H, params.W), dtype=torch.float32, device=dev)
        fg = torch.empty((B, 1, params.H, params.W), dtype=torch.float32, device=dev)
        ws = torch.empty(_lib.load().rgbd_dv_render_workspace_bytes(ctypes.byref(params), B, F), dtype=torch.uint8, device=dev)
        need_grad = any(ctx.needs_input_grad[:5])
        saved = torch.empty(_lib.load().rgbd_dv_render_saved_bytes(ctypes.byref(params), B) // 4, dtype=torch.float32,
                            device=dev) if need_grad else None         # running sums of every ray, for backward
        _lib.call("rgbd_dv_render_fwd", ctypes.byref(params), ctypes.byref(rparams), _ptr(grid), _ptr(cam), _ptr(W1),
                  _ptr(b1), _ptr(W2), _ptr(b2), B, F, _ptr(novel), _ptr(depth), _ptr(fg),
                  None if saved is None else _ptr(saved), _ptr(ws), ws.numel(), _stream())
        ctx.save_for_backward(grid, W1, b1, W2, b2, cam)
        ctx.params, ctx.rparams, ctx.ws, ctx.saved = params, rparams, ws, saved
        return novel, depth, fg

    @staticmethod
    def backward(ctx, g_novel, g_depth, g_fg):
        grid, W1, b1, W2, b2, cam = ctx.saved_tensors
        B, F = grid.shape[:2]
        z = lambda g, like: torch.zeros_like(like) if g is None else g.to(torch.float32).contiguous()
        g_novel = z(g_novel, torch.empty((B, F, ctx.params.H, ctx.params.W), device=grid.device))
        g_depth = z(g_depth, torch.empty((B, 1, ctx.params.H, ctx.params.W), device=grid.device))
        g_fg = None if g_fg is None else g_fg.to(torch.float32).contiguous()
        g_grid = torch.empty_like(grid)
        gW1, gb1, gW2, gb2 = torch.empty_like(W1), torch.empty_like(b1), torch.empty_like(W2), torch.empty_like(b2)
        _lib.call("rgbd_dv_render_bwd", ctypes.byref(ctx.params), ctypes.byref(ctx.rparams), _ptr(grid), _ptr(cam),
                  _ptr(W1), _ptr(b1), _ptr(W2), _ptr(b2), B, F, None if ctx.saved is None else _ptr(ctx.saved),
                  _ptr(g_novel), _ptr(g_depth),
                  None if g_fg is None else _ptr(g_fg), _ptr(g_grid), _ptr(gW1), _ptr(gb1), _ptr(gW2), _ptr(gb2),
                  _ptr(ctx.ws), ctx.ws.numel(), _stream())
        return g_grid, gW1, gb1, gW2, gb2, None, None, None


class _TrilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, grid, lin_ind, voxel_coords, params):
        b, F = grid.shape[:2]
        M = lin_ind.numel()
        ld = voxel_coords.stride(0) if M > 0 else 0
        n = params.W * params.H * params.D
        out = torch.empty((b, F, n), dtype=torch.float32, device=grid.device)
        G3 = params.G ** 3
        for i in range(b):
            _lib.call("rgbd_dv_trilinear_fwd", ctypes.c_void_p(grid.data_ptr() + 4 * i * F * G3), _ptr(lin_ind),
                      _ptr(voxel_coords), ld, M, F, ctypes.byref(params),
                      ctypes.c_void_p(out.data_ptr() + 4 * i * F * n), _stream())
        ctx.save_for_backward(lin_ind, voxel_coords)
        ctx.params, ctx.gshape, ctx.ld = params, grid.shape, ld
        return out

    @staticmethod
    def backward(ctx, g_out):
        lin_ind, voxel_coords = ctx.saved_tensors
        params = ctx.params
        b, F = ctx.gshape[:2]
        n = params.W * params.H * params.D
        G3 = params.G ** 3
        g_out = g_out.to(torch.float32).contiguous()
        g_grid = torch.empty(ctx.gshape, dtype=torch.float32, device=g_out.device)
        for i in range(b):
            _lib.call("rgbd_dv_trilinear_bwd", ctypes.c_void_p(g_out.data_ptr() + 4 * i * F * n), _ptr(lin_ind),
                      _ptr(voxel_coords), ctx.ld, lin_ind.numel(), F, ctypes.byref(params),
                      ctypes.c_void_p(g_grid.data_ptr() + 4 * i * F * G3), _stream())
        return g_grid, None, None, None


def interpolate_trilinear(grid, lin_ind_frustrum, voxel_coords, img_shape, frustrum_depth):
    """deepvoxel/deepvoxel.py:388-428: grid (b,F,G,G,G) -> (b,F,frustrum_depth,img_shape[0],img_shape[1])."""
    grid = _dev_f32(grid, "grid")
    batch, num_feats, height, width, depth = grid.shape
    if not (height == width == depth):
        raise ValueError("cubic grids only")
    lin = lin_ind_frustrum.to(torch.int32).contiguous()
    vc = voxel_coords.to(torch.float32)
    if vc.stride(1) != 1:
        vc = vc.contiguous()
    params = DvParams(int(img_shape[1]), int(img_shape[0]), int(frustrum_depth), int(depth), 1.0, 1.0, 0.0, 0.0,
                      1.0, 0.0)
    out = _TrilinearFn.apply(grid, lin, vc, params)
    return out.reshape(batch, num_feats, frustrum_depth, img_shape[0], img_shape[1])


class MakeSlice:
    """deepvoxel/deepvoxel.py:431-433"""

    def __getitem__(self, item):
        return item


__all__ = ["ProjectionHelper", "interpolate_trilinear", "MakeSlice"]
