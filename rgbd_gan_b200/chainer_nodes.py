"""Chainer FunctionNodes over CuPy arrays for a deployment that has Chainer >= 7 and CuPy >= 7
(the reference's own stack, README.md:19-27).  They call the SAME C-ABI entry points, in the same
order and with the same arguments, as the torch glue in loss_functions.py / projection.py that the
GPU tests exercise; only the array container (cupy.ndarray: `arr.data.ptr`, `cupy.cuda.get_current_
stream().ptr`) and the autograd hook (chainer.FunctionNode) differ.

Neither package can be installed in the build image, so this module is import-guarded and is NOT
covered by the -m gpu tests; tests/test_chainer_nodes.py checks its node logic against the Chainer-v7
shim with a recording fake of the library (argument order, shapes, retained state).

Drop-in use in the reference:
    from rgbd_gan_b200.chainer_nodes import LossFuncRotate      # instead of common.loss_functions
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import LossOpts
from .loss_functions import pose_algebra

try:                                            # pragma: no cover - not installable in the build image
    import chainer
    from chainer import FunctionNode, Variable
except ImportError:                             # the module stays importable for documentation
    chainer = None
    Variable = None

    class FunctionNode(object):                 # minimal base so the class bodies below are defined
        def apply(self, inputs):
            raise RuntimeError("chainer_nodes needs Chainer (>= 7) and CuPy")

try:                                            # pragma: no cover
    import cupy
except ImportError:
    cupy = None
HAVE_CHAINER = chainer is not None and cupy is not None and not getattr(chainer, "_is_shim", False)


def _ptr(a):
    return ctypes.c_void_p(0 if a is None else int(a.data.ptr))


def _stream(xp):
    return ctypes.c_void_p(int(xp.cuda.get_current_stream().ptr))


class ConsistencyLoss(FunctionNode):
    """LossFuncRotate.__call__ body (common/loss_functions.py:93-146) as one node.
    inputs: (img, img_rot) cupy float32 (B,C,H,W); outputs: (loss 0-dim, new_zp_cat (2B,HW,3))."""

    def __init__(self, M, c, Mi, ci, opts, workspace, grad_scale=None, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib
        self.poses = tuple(self.xp.asarray(a, dtype="float32") for a in (M, c, Mi, ci))   # one small H2D copy each
        self.opts, self.ws, self.grad_scale = opts, workspace, grad_scale
        self.stash = None

    def check_type_forward(self, in_types):
        if chainer is None or getattr(chainer, "_is_shim", False):
            return
        chainer.utils.type_check.expect(in_types.size() == 2, in_types[0].dtype == np.float32,
                                        in_types[0].ndim == 4, in_types[0].shape == in_types[1].shape)

    def forward(self, inputs):
        xp = self.xp
        img, img_rot = (xp.ascontiguousarray(a) for a in inputs)
        self.retain_inputs((0, 1))
        B, C, H, W = img.shape
        parts = xp.empty(8, dtype="float32")
        new_zp = xp.empty((2 * B, H * W, 3), dtype="float32")
        pp = [_ptr(a) for a in self.poses]
        if self.grad_scale is not None:
            g_img, g_rot = xp.empty_like(img), xp.empty_like(img_rot)
            self.lib.call("rgbd_consistency_fwd_bwd", _ptr(img), _ptr(img_rot), *pp, B, C, H, W,
                          ctypes.byref(self.opts), ctypes.c_float(self.grad_scale), _ptr(parts), _ptr(new_zp),
                          _ptr(g_img), _ptr(g_rot), _ptr(self.ws), self.ws.size, _stream(xp))
            self.stash = (g_img, g_rot)
        else:
            self.lib.call("rgbd_consistency_fwd", _ptr(img), _ptr(img_rot), *pp, B, C, H, W,
                          ctypes.byref(self.opts), _ptr(parts), _ptr(new_zp), None, _ptr(self.ws), self.ws.size,
                          _stream(xp))
        self.loss_parts = parts
        # parts[4] = (p0+p1) + (p2*l + p3*l), :141-144; parts[6] adds the fused depth hinge (== parts[4] when off)
        return parts[6].reshape(()), new_zp

    def backward(self, target_input_indexes, grad_outputs):
        xp = self.xp
        img, img_rot = (v.array for v in self.get_retained_inputs())
        gy, g_zp = grad_outputs
        B, C, H, W = img.shape
        gy_dev = xp.zeros((), dtype="float32") if gy is None else xp.ascontiguousarray(gy.array, dtype="float32")
        gz = None if g_zp is None else xp.ascontiguousarray(g_zp.array, dtype="float32")
        if self.stash is not None and gz is None:
            g_img, g_rot = self.stash
            self.lib.call("rgbd_consistency_rescale", _ptr(g_img), _ptr(g_rot), g_img.size, _ptr(gy_dev),
                          ctypes.c_float(self.grad_scale), _stream(xp))
        else:
            g_img, g_rot = xp.empty_like(img), xp.empty_like(img_rot)
            self.lib.call("rgbd_consistency_bwd", _ptr(img), _ptr(img_rot), *[_ptr(a) for a in self.poses], B, C, H, W,
                          ctypes.byref(self.opts), ctypes.c_float(1.0), _ptr(gy_dev), _ptr(gz), _ptr(g_img), _ptr(g_rot),
                          _ptr(self.ws), self.ws.size, _stream(xp))
        return _as_var(g_img), _as_var(g_rot)


def _as_var(a):
    return a if Variable is None else Variable(a)


class LossFuncRotate:
    """Drop-in for common/loss_functions.py:31-146 (same constructor and call signature)."""

    def __init__(self, xp, K=None, norm="l1", lambda_geometric=3, grad_scale=None, lib=None):
        self.xp = xp
        self.size = None
        self.K = K
        self.norm = norm
        self.lambda_geometric = lambda_geometric
        self.grad_scale = grad_scale
        self._lib = lib if lib is not None else _lib
        self._ws = None

    def init_params(self, xp, size=4):
        """:39-61 (host NumPy; K, inv_K, p are constants of the kernels)"""
        if self.size is None:
            if self.K is not None:
                K = self.K.get() if hasattr(self.K, "get") else self.K
                self.K = np.array(np.asarray(K)[:3, :3], "float32")
                self.K[:2] *= size / self.K[0, 2] / 2
                self.size = size
            else:
                self.size = size
                self.K = np.array([[size * 2, 0, size / 2], [0, size * 2, size / 2], [0, 0, 1]], dtype="float32")
        else:
            self.size = size
            self.K[:2] *= size / self.K[0, 2] / 2
        self.inv_K = np.linalg.inv(self.K).astype("float32")
        self.p = np.asarray(list(np.meshgrid(np.arange(size), np.arange(size))) + [np.ones((size, size))],
                            dtype="float32").reshape(3, -1)

    def __call__(self, img, theta, img_rot, theta_rot, occlusion_aware=False, debug=False, max_depth=None,
                 min_depth=None, depth_hinge=None):
        """reference signature plus depth_hinge=(depth_min, lambda_depth): updater.py:357-359 fused in"""
        if debug:
            raise NotImplementedError("debug=True: use rgbd_gan_b200.loss_functions (warp/bilinear kernels)")
        xp = self.xp
        if self.size != img.shape[-1]:
            self.init_params(xp, size=img.shape[-1])
        if hasattr(theta, "array"):                         # :82-84
            theta, theta_rot = theta.array, theta_rot.array
        to_host = (lambda a: a.get()) if hasattr(theta, "get") else np.asarray
        M, c, Mi, ci = pose_algebra(self.K, self.inv_K, to_host(theta), to_host(theta_rot))
        B, C, H, W = img.shape
        nbytes = self._lib.load().rgbd_consistency_workspace_bytes(B, C, H, W)
        if self._ws is None or self._ws.size < nbytes:
            self._ws = xp.empty(nbytes, dtype="uint8")
        opts = LossOpts(_lib.NORM_L1 if self.norm == "l1" else _lib.NORM_L2, int(bool(occlusion_aware)),
                        float("nan") if max_depth is None else float(max_depth),
                        float("nan") if min_depth is None else float(min_depth), float(self.lambda_geometric), B, None,
                        0, 0, float("nan") if depth_hinge is None else float(depth_hinge[0]),
                        0.0 if depth_hinge is None else float(depth_hinge[1]))
        node = ConsistencyLoss(M, c, Mi, ci, opts, self._ws, self.grad_scale, xp=xp, lib=self._lib)
        loss, new_zp = node.apply((img, img_rot))
        return loss, new_zp


class RenderAccumulative(FunctionNode):
    """"Next" row: the per-sample loop of DeepVoxels.forward with `occlusion_type: accumulative`
    (deepvoxel.py:879-892 + AccumulativeOcclusionNet.forward :574-587 + depth rescale :903-904) as one node.
    inputs: (deepvoxels (B,F,G,G,G), W1 (nf,F+1), b1 (nf,), W2 (1,nf), b2 (1,)) -- the `.c.W` (1x1x1 kernels squeezed)
    and `.c.b` of the two EqualizedConv3d of `occlusion_net.occlusion`; outputs: (novel_views (B,F,H,W),
    depth_maps (B,1,H,W), foreground_weight (B,1,H,W))."""

    def __init__(self, cam2world, dv_params, accmulative_threshold=4, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib
        self.cam = self.xp.asarray(np.asarray(cam2world, dtype="float32").reshape(-1, 16), dtype="float32")
        self.P, self.threshold = dv_params, float(accmulative_threshold)

    def _rparams(self, F, nf):
        return _lib.DvRenderParams(int(nf), int(np.ceil(np.sqrt(3) * self.P.G)), self.threshold,
                                   float(np.float32(np.sqrt(2) * np.sqrt(1.0 / (F + 1)))),
                                   float(np.float32(np.sqrt(2) * np.sqrt(1.0 / nf))))

    def forward(self, inputs):
        xp = self.xp
        grid, W1, b1, W2, b2 = (xp.ascontiguousarray(a, dtype="float32") for a in inputs)
        B, F = grid.shape[:2]
        P = self.P
        self.R = self._rparams(F, W1.shape[0])
        self.ws = xp.empty(self.lib.load().rgbd_dv_render_workspace_bytes(ctypes.byref(P), B, F), dtype="uint8")
        novel, depth, fg = xp.empty((B, F, P.H, P.W), "float32"), xp.empty((B, 1, P.H, P.W), "float32"), \
            xp.empty((B, 1, P.H, P.W), "float32")
        self.saved = xp.empty((B * P.H * P.W * (P.D + 1),), "float32")    # running sums of every ray, for backward
        self.lib.call("rgbd_dv_render_fwd", ctypes.byref(P), ctypes.byref(self.R), _ptr(grid), _ptr(self.cam), _ptr(W1),
                      _ptr(b1), _ptr(W2), _ptr(b2), B, F, _ptr(novel), _ptr(depth), _ptr(fg), _ptr(self.saved),
                      _ptr(self.ws), int(self.ws.size), _stream(xp))
        self.retain_inputs((0, 1, 2, 3, 4))
        return novel, depth, fg

    def backward(self, target_input_indexes, grad_outputs):
        xp = self.xp
        grid, W1, b1, W2, b2 = (xp.ascontiguousarray(v.array, dtype="float32") for v in self.get_retained_inputs())
        B, F = grid.shape[:2]
        P = self.P
        arr = lambda g, shape: xp.zeros(shape, "float32") if g is None else xp.ascontiguousarray(
            g.array if hasattr(g, "array") else g, dtype="float32")
        g_novel, g_depth = arr(grad_outputs[0], (B, F, P.H, P.W)), arr(grad_outputs[1], (B, 1, P.H, P.W))
        g_fg = None if grad_outputs[2] is None else arr(grad_outputs[2], (B, 1, P.H, P.W))
        outs = [xp.empty_like(a) for a in (grid, W1, b1, W2, b2)]
        self.lib.call("rgbd_dv_render_bwd", ctypes.byref(P), ctypes.byref(self.R), _ptr(grid), _ptr(self.cam), _ptr(W1),
                      _ptr(b1), _ptr(W2), _ptr(b2), B, F, _ptr(self.saved), _ptr(g_novel), _ptr(g_depth), _ptr(g_fg),
                      *[_ptr(o) for o in outs], _ptr(self.ws), int(self.ws.size), _stream(xp))
        return tuple(_as_var(outs[i]) for i in target_input_indexes)
