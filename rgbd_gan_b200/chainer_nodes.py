"""The reference's call surface as Chainer FunctionNodes over CuPy arrays (Chainer >= 7, CuPy >= 7: the reference's own
stack, README.md:19-27).  Every node calls the SAME C-ABI entry points (include/rgbdgan_b200.h), in the same order and
with the same arguments, as the torch glue in loss_functions.py / projection.py that the GPU tests and bench.py drive;
only the array container and the autograd hook differ.  This module imports neither torch nor (at import time) CuPy:

    from rgbd_gan_b200.chainer_nodes import LossFuncRotate, warp, inv_warp, bilinear          # common/loss_functions.py
    from rgbd_gan_b200.chainer_nodes import ProjectionHelper, interpolate_trilinear           # deepvoxel/projection.py,
                                                                                              # deepvoxel/deepvoxel.py:388

Containers: anything `xp` hands out that exposes a device address -- cupy.ndarray (`a.data.ptr`), an object with
`__cuda_array_interface__`, or a torch tensor -- see host_math.device_pointer.  `xp` is the array module the reference
passes around (`cupy`); only `xp.empty / zeros / asarray / ascontiguousarray / empty_like` and
`xp.cuda.get_current_stream().ptr` are used.

Neither Chainer nor CuPy can be installed in the build image.  tests/test_chainer_nodes.py drives every symbol below
through the Chainer-v7 shim with a stand-in library that executes each C-ABI call with the CPU oracle on the pointers it
is handed (argument order, shapes, retained state, gradients), and tests/test_gpu_chainer_surface.py drives the same
nodes on the GPU with a cupy-like array module over torch memory and the real library.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import DvParams, LossOpts
from .host_math import (as_numpy, device_pointer, grid_dims, intrinsics_for_size, pixel_grid, pose_algebra,
                        warp_constants)

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
HAVE_CHAINER = chainer is not None and cupy is not None


def _ptr(a):
    return ctypes.c_void_p(device_pointer(a))


def _stream(xp):
    return ctypes.c_void_p(int(xp.cuda.get_current_stream().ptr))


def _as_var(a):
    return a if Variable is None else Variable(a)


def _arr(v):
    """Variable | array -> array"""
    return v.array if hasattr(v, "array") and not isinstance(v, np.ndarray) else v


def _f32c(xp, a):
    return xp.ascontiguousarray(_arr(a), dtype="float32")


def _select(grads, target_input_indexes):
    """FunctionNode.backward returns one gradient per REQUESTED input"""
    return tuple(None if grads[i] is None else _as_var(grads[i]) for i in target_input_indexes)


# ----------------------------------------------------------------------------------------------- consistency loss
class ConsistencyLoss(FunctionNode):
    """LossFuncRotate.__call__ body (common/loss_functions.py:93-146) as one node.
    inputs: (img, img_rot) float32 (B,C,H,W); outputs: (loss 0-dim, new_zp_cat (2B,HW,3))."""

    def __init__(self, M, c, Mi, ci, opts, workspace, grad_scale=None, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib
        self.poses = tuple(self.xp.asarray(a, dtype="float32") for a in (M, c, Mi, ci))   # one small H2D copy each
        self.opts, self.ws, self.grad_scale = opts, workspace, grad_scale
        self.stash = None

    def check_type_forward(self, in_types):
        tc = getattr(getattr(chainer, "utils", None), "type_check", None)
        if tc is None:
            return
        tc.expect(in_types.size() == 2, in_types[0].dtype == np.float32, in_types[0].ndim == 4,
                  in_types[0].shape == in_types[1].shape)

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
                          _ptr(g_img), _ptr(g_rot), _ptr(self.ws), int(self.ws.size), _stream(xp))
            self.stash = (g_img, g_rot)
        else:
            self.lib.call("rgbd_consistency_fwd", _ptr(img), _ptr(img_rot), *pp, B, C, H, W,
                          ctypes.byref(self.opts), _ptr(parts), _ptr(new_zp), None, _ptr(self.ws), int(self.ws.size),
                          _stream(xp))
        self.loss_parts = parts
        # parts[4] = (p0+p1) + (p2*l + p3*l), :141-144; parts[6] adds the fused depth hinge (== parts[4] when off)
        return parts[6].reshape(()), new_zp

    def backward(self, target_input_indexes, grad_outputs):
        xp = self.xp
        img, img_rot = (_arr(v) for v in self.get_retained_inputs())
        gy, g_zp = grad_outputs
        B, C, H, W = img.shape
        gy_dev = xp.zeros((), dtype="float32") if gy is None else _f32c(xp, gy)
        gz = None if g_zp is None else _f32c(xp, g_zp)
        if self.stash is not None and gz is None:
            g_img, g_rot = self.stash
            self.stash = None
            self.lib.call("rgbd_consistency_rescale", _ptr(g_img), _ptr(g_rot), int(g_img.size), _ptr(gy_dev),
                          ctypes.c_float(self.grad_scale), _stream(xp))
        else:
            g_img, g_rot = xp.empty_like(img), xp.empty_like(img_rot)
            self.lib.call("rgbd_consistency_bwd", _ptr(img), _ptr(img_rot), *[_ptr(a) for a in self.poses], B, C, H, W,
                          ctypes.byref(self.opts), ctypes.c_float(1.0), _ptr(gy_dev), _ptr(gz), _ptr(g_img), _ptr(g_rot),
                          _ptr(self.ws), int(self.ws.size), _stream(xp))
        return _select((g_img, g_rot), target_input_indexes)


# ------------------------------------------------------------------------------------ pose pipeline on the device
def _f9(a):
    a = np.ascontiguousarray(np.asarray(as_numpy(a), dtype=np.float32)[:3, :3]).reshape(9)
    return (ctypes.c_float * 9)(*a.tolist())


def pose_algebra_device(K, inv_K, theta, theta_rot, xp=None, lib=None):
    """common/loss_functions.py:85-91 + the constants of :174 / :181 from DEVICE cam2world matrices (B,4,4):
    rgbd_pose_algebra, no host synchronisation.  Returns device arrays M (B,3,3), c (B,3,1), Mi (B,3,3), ci (B,3,1)."""
    xp = _xp_of(theta, xp)
    lib = lib if lib is not None else _lib
    theta, theta_rot = _f32c(xp, theta), _f32c(xp, theta_rot)
    B = int(theta.shape[0])
    M, c = xp.empty((B, 3, 3), dtype="float32"), xp.empty((B, 3, 1), dtype="float32")
    Mi, ci = xp.empty((B, 3, 3), dtype="float32"), xp.empty((B, 3, 1), dtype="float32")
    lib.call("rgbd_pose_algebra", _ptr(theta), _ptr(theta_rot), B, _f9(K), _f9(inv_K), _ptr(M), _ptr(c), _ptr(Mi), _ptr(ci),
             _stream(xp))
    return M, c, Mi, ci


def get_camera_matries(thetas, order=(0, 1, 2), cos_sin=None, xp=None, lib=None):
    """updater.py:45-60 on the device: (N,6) thetas -> (N,4,4) cam2world.  cos_sin (N,6) = [cos | sin] of the angles as
    the caller's array library computed them (updater.py:316-317 computes exactly these for the generator) makes the
    result bit-identical to the reference's NumPy path."""
    xp = _xp_of(thetas, xp)
    lib = lib if lib is not None else _lib
    thetas = _f32c(xp, thetas)
    cs = None if cos_sin is None else _f32c(xp, cos_sin)
    n = int(thetas.shape[0])
    cam = xp.empty((n, 4, 4), dtype="float32")
    lib.call("rgbd_pose_camera_matrices", _ptr(thetas), _ptr(cs), n, (ctypes.c_int * 3)(*[int(o) for o in order]), _ptr(cam),
             _stream(xp))
    return cam


class CameraParamPrior:
    """train_rgbd.py:192-217 on the device: sample(batch_size) -> (batch_size, 6) float32 device array.  draws=None:
    Philox keyed by (seed, call counter, pair); draws (B,15) float64 replays a host np.random stream bit for bit."""

    def __init__(self, config, xp=None, seed=0, lib=None):
        rng = [config.x_rotate, config.y_rotate, config.z_rotate, config.x_translate, config.y_translate, config.z_translate]
        self.camera_param_range = np.array(rng, dtype=np.float64)
        self.rotation_range = self.camera_param_range[:3]
        self.uniform = bool(config.uniform_distribution)
        self.xp = xp if xp is not None else cupy
        self.seed, self.step = int(seed), 0
        self._lib = lib if lib is not None else _lib
        self.c_prior = _lib.PosePrior((ctypes.c_double * 6)(*self.camera_param_range.tolist()), int(self.uniform))

    def sample(self, batch_size, draws=None):
        xp, B = self.xp, batch_size // 2
        thetas = xp.empty((2 * B, 6), dtype="float32")
        d = None if draws is None else xp.ascontiguousarray(xp.asarray(draws), dtype="float64")
        self._lib.call("rgbd_pose_sample", ctypes.byref(self.c_prior), B, _ptr(d), self.seed, self.step, _ptr(thetas),
                       _stream(xp))
        self.step += 1
        return thetas


# ------------------------------------------------------------------------------------ warp / inv_warp / bilinear
class Warp(FunctionNode):
    """warp / inv_warp (common/loss_functions.py:171-182) with the constant factors M = K R K^-1, cv folded on the
    host: input z (B,1,HW) -> new_zp (B,HW,3) = (M (z p) - cv)^T"""

    def __init__(self, M, cv, H, W, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib
        self.M = self.xp.asarray(M, dtype="float32")
        self.cv = self.xp.asarray(cv, dtype="float32")
        self.H, self.W = H, W

    def forward(self, inputs):
        xp = self.xp
        z = xp.ascontiguousarray(inputs[0], dtype="float32")
        self.zshape = z.shape
        B = z.shape[0]
        out = xp.empty((B, self.H * self.W, 3), dtype="float32")
        self.lib.call("rgbd_warp_fwd", _ptr(z), _ptr(self.M), _ptr(self.cv), B, self.H, self.W, _ptr(out), _stream(xp))
        return out,

    def backward(self, target_input_indexes, grad_outputs):
        xp = self.xp
        g = _f32c(xp, grad_outputs[0])
        B = g.shape[0]
        gz = xp.empty((B, self.H * self.W), dtype="float32")
        self.lib.call("rgbd_warp_bwd", _ptr(g), _ptr(self.M), B, self.H, self.W, _ptr(gz), _stream(xp))
        return _as_var(gz.reshape(self.zshape)),


def _xp_of(a, xp):
    if xp is not None:
        return xp
    if cupy is None:
        raise RuntimeError("pass xp= (the CuPy-like array module) when CuPy is not importable")
    return cupy


def warp(K, inv_K, R, t, z, p, xp=None, lib=None):
    """common/loss_functions.py:171-175: (K R K^-1)(z p) - (K R) t as (B,HW,3); differentiable in z"""
    H, W = grid_dims(p, z.shape[-1])
    M, cv = warp_constants(K, inv_K, R, t, inverse=False)
    return Warp(M, cv, H, W, xp=_xp_of(z, xp), lib=lib).apply((z,))[0]


def inv_warp(K, inv_K, inv_R, t, z, p, xp=None, lib=None):
    """common/loss_functions.py:178-182: (K R^T K^-1)(z p) + K t as (B,HW,3); differentiable in z"""
    H, W = grid_dims(p, z.shape[-1])
    M, cv = warp_constants(K, inv_K, inv_R, t, inverse=True)
    return Warp(M, cv, H, W, xp=_xp_of(z, xp), lib=lib).apply((z,))[0]


class Bilinear(FunctionNode):
    """bilinear(img, zp) (common/loss_functions.py:185-228): inputs (img (B,C,H,W), zp (B,HW,3)) ->
    warped (B*HW, C); the in-bounds mask (:215-216) is kept as `self.mask` (non-differentiable)"""

    def __init__(self, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib

    def forward(self, inputs):
        xp = self.xp
        img, zp = (xp.ascontiguousarray(a, dtype="float32") for a in inputs)
        self.retain_inputs((0, 1))
        B, C, H, W = img.shape
        warped = xp.empty((B * H * W, C), dtype="float32")
        mask = xp.empty((B * H * W,), dtype="uint8")
        self.lib.call("rgbd_bilinear_fwd", _ptr(img), _ptr(zp), B, C, H, W, _ptr(warped), _ptr(mask), _stream(xp))
        self.mask = mask
        return warped,

    def backward(self, target_input_indexes, grad_outputs):
        xp = self.xp
        img, zp = (xp.ascontiguousarray(_arr(v), dtype="float32") for v in self.get_retained_inputs())
        B, C, H, W = img.shape
        g = _f32c(xp, grad_outputs[0])
        g_img, g_zp = xp.empty_like(img), xp.empty_like(zp)
        self.lib.call("rgbd_bilinear_bwd", _ptr(img), _ptr(zp), _ptr(g), B, C, H, W, _ptr(g_img), _ptr(g_zp), _stream(xp))
        return _select((g_img, g_zp), target_input_indexes)


def bilinear(img, zp, xp=None, lib=None):
    """common/loss_functions.py:185-228 -> (warped (B*HW, C) Variable, not_getting_out (B*HW,) bool array)"""
    node = Bilinear(xp=_xp_of(img, xp), lib=lib)
    warped, = node.apply((img, zp))
    return warped, node.mask.astype("bool")


class LossFuncRotate:
    """Drop-in for common/loss_functions.py:31-168: same constructor, attributes and call signature.
    Extra keyword arguments (not in the reference): grad_scale (the constant upstream gradient lambda_rotate of
    updater.py:363-365: loss AND gradients then come out of one pass), peer_comm / n_pairs_global / defer_loss (pairs
    sharded over the GPUs of one box, see distributed.PeerComm), lib (test hook: the C-ABI binding to call)."""

    def __init__(self, xp, K=None, norm="l1", lambda_geometric=3, grad_scale=None, peer_comm=None, n_pairs_global=None,
                 defer_loss=False, lib=None):
        self.xp = xp
        self.size = None
        self.K = K
        self.norm = norm
        self.lambda_geometric = lambda_geometric
        self.inv_K = None
        self.p = None
        self.grad_scale = grad_scale
        self.peer_comm, self.n_pairs_global = peer_comm, n_pairs_global
        self.defer_loss = (2 if defer_loss == "lazy" else int(bool(defer_loss))) if peer_comm is not None else 0
        self._lib = lib if lib is not None else _lib
        self._ws = None

    def init_params(self, xp, size=4):
        """:39-61 (host NumPy; K, inv_K, p are constants of the kernels)"""
        self.K, self.inv_K = intrinsics_for_size(self.K, size, first=self.size is None)   # in place later (quirk Q9)
        self.size = size
        self.p = pixel_grid(size)

    def _opts(self, occlusion_aware, max_depth, min_depth, B, depth_hinge):
        world = 1 if self.peer_comm is None else self.peer_comm.world
        return LossOpts(_lib.NORM_L1 if self.norm == "l1" else _lib.NORM_L2, int(bool(occlusion_aware)),
                        float("nan") if max_depth is None else float(max_depth),
                        float("nan") if min_depth is None else float(min_depth), float(self.lambda_geometric),
                        int(self.n_pairs_global) if self.n_pairs_global else int(B) * world,
                        None if self.peer_comm is None else self.peer_comm.handle, int(self.defer_loss), 0,
                        float("nan") if depth_hinge is None else float(depth_hinge[0]),
                        0.0 if depth_hinge is None else float(depth_hinge[1]))

    def __call__(self, img, theta, img_rot, theta_rot, occlusion_aware=False, debug=False, max_depth=None,
                 min_depth=None, depth_hinge=None):
        """reference signature plus depth_hinge=(depth_min, lambda_depth): updater.py:357-359 fused in"""
        xp = self.xp
        if self.size != img.shape[-1]:
            self.init_params(xp, size=img.shape[-1])
        if hasattr(theta, "array") and not isinstance(theta, np.ndarray):                   # :82-84
            theta, theta_rot = theta.array, theta_rot.array
        B, C, H, W = img.shape
        if not isinstance(theta, np.ndarray) and not isinstance(theta, (list, tuple)) and not debug:
            # device arrays (the reference's production case, updater.py:315): one small kernel, no .get() sync
            M, c, Mi, ci = pose_algebra_device(self.K, self.inv_K, theta, theta_rot, xp=xp, lib=self._lib)
        else:
            M, c, Mi, ci = pose_algebra(self.K, self.inv_K, theta, theta_rot)
        if debug:                                           # :100-102 -- the six intermediates
            z = _arr(img)[:, -1:].reshape(B, 1, -1)
            z_rot = _arr(img_rot)[:, -1:].reshape(B, 1, -1)
            new_zp, = Warp(M, c, H, W, xp=xp, lib=self._lib).apply((z,))
            new_zp_rot, = Warp(Mi, ci, H, W, xp=xp, lib=self._lib).apply((z_rot,))
            warped, not_out = bilinear(img_rot, new_zp, xp=xp, lib=self._lib)
            warped_rot, not_out_rot = bilinear(img, new_zp_rot, xp=xp, lib=self._lib)
            return warped, not_out, new_zp, warped_rot, not_out_rot, new_zp_rot
        nbytes = self._lib.load().rgbd_consistency_workspace_bytes(B, C, H, W)
        if self._ws is None or self._ws.size < nbytes:
            self._ws = xp.empty(nbytes, dtype="uint8")
        opts = self._opts(occlusion_aware, max_depth, min_depth, B, depth_hinge)
        node = ConsistencyLoss(M, c, Mi, ci, opts, self._ws, self.grad_scale, xp=xp, lib=self._lib)
        loss, new_zp = node.apply((img, img_rot))
        self.last_loss_parts = node.loss_parts
        return loss, new_zp

    # -- :148-158 (no caller in the reference): plain array-library math on Chainer's own functions
    def calc_real_pos(self, img, theta):
        import chainer.functions as F
        xp = self.xp
        theta = as_numpy(theta)
        if theta.ndim == 1:
            assert False, "only rotation matrices are supported for theta"
        R = xp.asarray(np.ascontiguousarray(theta[:, :3, :3], dtype="float32"))
        t = xp.asarray(np.ascontiguousarray(theta[:, :3, -1:], dtype="float32"))
        arr = _arr(img)
        z = arr[:, -1:].reshape(arr.shape[0], 1, -1)
        rgb = arr[:, :3].reshape(arr.shape[0], 3, -1)
        inv_K, p = xp.asarray(self.inv_K), xp.asarray(self.p)
        real_pos = F.matmul(xp.matmul(R, inv_K), z * p) + t
        return F.concat([rgb, real_pos], axis=1)

    # -- :160-168 (only with `use_occupancy_net_loss`, absent from every shipped config)
    def occupancy_net_loss(self, occupancy_net, depth, theta, z):
        import chainer.functions as F
        xp = self.xp
        theta = as_numpy(theta)
        R = xp.asarray(np.ascontiguousarray(theta[:, :3, :3], dtype="float32"))
        t = xp.asarray(np.ascontiguousarray(theta[:, :3, -1:], dtype="float32"))
        depth = F.reshape(depth, (depth.shape[0], 1, -1))
        eps = xp.asarray(np.random.normal(size=depth.shape).astype("float32") * 0.05)
        inv_K, p = xp.asarray(self.inv_K), xp.asarray(self.p)
        real_pos = F.matmul(xp.matmul(R, inv_K), (depth + eps) * p) + t
        label = (eps > 0).reshape(-1, 1).astype("int32")
        occupancy_field = occupancy_net(z, real_pos + eps)
        return F.sigmoid_cross_entropy(occupancy_field, label)


class DepthHead(FunctionNode):
    """"next" row: the generators' depth head (net.py:294-299, :756-761) as one node:
    h (B,C,H,W) -> concat([h[:, :-1], 1 / (softplus(h[:, -1:]) + 1e-4)])"""

    def __init__(self, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib

    def forward(self, inputs):
        xp = self.xp
        h = xp.ascontiguousarray(inputs[0], dtype="float32")
        self.retain_inputs((0,))
        B, C, H, W = h.shape
        out = xp.empty_like(h)
        self.lib.call("rgbd_depth_head_fwd", _ptr(h), B, C, H, W, _ptr(out), _stream(xp))
        return out,

    def backward(self, target_input_indexes, grad_outputs):
        xp = self.xp
        h = xp.ascontiguousarray(_arr(self.get_retained_inputs()[0]), dtype="float32")
        B, C, H, W = h.shape
        g = _f32c(xp, grad_outputs[0])
        g_h = xp.empty_like(h)
        self.lib.call("rgbd_depth_head_bwd", _ptr(h), _ptr(g), B, C, H, W, _ptr(g_h), _stream(xp))
        return _as_var(g_h),


def depth_head(h, xp=None, lib=None):
    return DepthHead(xp=_xp_of(h, xp), lib=lib).apply((h,))[0]


# ------------------------------------------------------------------------------------------------- DeepVoxels
class Project(FunctionNode):
    """the generator's per-sample loops (deepvoxels_generator.py:287-299 -> deepvoxel.py:879-884: compute_proj_idcs +
    interpolate_trilinear) for the whole batch: input grid (B,F,G,G,G) -> frustum (B,F,D,H,W)"""

    def __init__(self, cam2world, dv_params, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib
        self.cam = self.xp.asarray(np.asarray(as_numpy(cam2world), dtype="float32").reshape(-1, 16), dtype="float32")
        self.P = dv_params

    def _ws(self, B, F):
        n = self.lib.load().rgbd_dv_project_workspace_bytes(ctypes.byref(self.P), B, F)
        return self.xp.empty(max(int(n), 16), dtype="uint8")

    def forward(self, inputs):
        xp, P = self.xp, self.P
        grid = xp.ascontiguousarray(inputs[0], dtype="float32")
        self.gshape = grid.shape
        B, F = grid.shape[:2]
        out = xp.empty((B, F, P.D, P.H, P.W), dtype="float32")
        ws = self._ws(B, F)
        self.lib.call("rgbd_dv_project_fwd", ctypes.byref(P), _ptr(grid), _ptr(self.cam), B, F, _ptr(out), _ptr(ws),
                      int(ws.size), _stream(xp))
        return out,

    def backward(self, target_input_indexes, grad_outputs):
        xp, P = self.xp, self.P
        B, F = self.gshape[:2]
        g = _f32c(xp, grad_outputs[0])
        g_grid = xp.empty(self.gshape, dtype="float32")
        ws = self._ws(B, F)
        self.lib.call("rgbd_dv_project_bwd", ctypes.byref(P), _ptr(g), _ptr(self.cam), B, F, _ptr(g_grid), _ptr(ws),
                      int(ws.size), _stream(xp))
        return _as_var(g_grid),


class Trilinear(FunctionNode):
    """interpolate_trilinear (deepvoxel/deepvoxel.py:388-428) on an index list from compute_proj_idcs:
    input grid (b,F,G,G,G) -> (b,F,D*H*W); lin_ind (M,) int32 and voxel_coords (3,M) are closed over"""

    def __init__(self, lin_ind, voxel_coords, dv_params, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib
        self.lin = self.xp.ascontiguousarray(_arr(lin_ind), dtype="int32")
        self.vc = self.xp.ascontiguousarray(_arr(voxel_coords), dtype="float32")
        self.P = dv_params

    def _run(self, name, src, dst, b, F, n_src, n_dst):
        xp, P = self.xp, self.P
        M = int(self.lin.size)
        for i in range(b):
            self.lib.call(name, ctypes.c_void_p(device_pointer(src) + 4 * i * F * n_src), _ptr(self.lin), _ptr(self.vc),
                          M, M, F, ctypes.byref(P), ctypes.c_void_p(device_pointer(dst) + 4 * i * F * n_dst), _stream(xp))

    def forward(self, inputs):
        xp, P = self.xp, self.P
        grid = xp.ascontiguousarray(inputs[0], dtype="float32")
        self.gshape = grid.shape
        b, F = grid.shape[:2]
        n = P.W * P.H * P.D
        out = xp.empty((b, F, n), dtype="float32")
        self._run("rgbd_dv_trilinear_fwd", grid, out, b, F, P.G ** 3, n)
        return out,

    def backward(self, target_input_indexes, grad_outputs):
        xp, P = self.xp, self.P
        b, F = self.gshape[:2]
        g = _f32c(xp, grad_outputs[0])
        g_grid = xp.empty(self.gshape, dtype="float32")
        self._run("rgbd_dv_trilinear_bwd", g, g_grid, b, F, P.W * P.H * P.D, P.G ** 3)
        return _as_var(g_grid),


def interpolate_trilinear(grid, lin_ind_frustrum, voxel_coords, img_shape, frustrum_depth, xp=None, lib=None):
    """deepvoxel/deepvoxel.py:388-428: grid (b,F,G,G,G) -> (b,F,frustrum_depth,img_shape[0],img_shape[1])"""
    batch, num_feats, height, width, depth = grid.shape
    if not (height == width == depth):
        raise ValueError("cubic grids only")
    P = DvParams(int(img_shape[1]), int(img_shape[0]), int(frustrum_depth), int(depth), 1.0, 1.0, 0.0, 0.0, 1.0, 0.0)
    out, = Trilinear(lin_ind_frustrum, voxel_coords, P, xp=_xp_of(grid, xp), lib=lib).apply((grid,))
    if hasattr(out, "reshape"):
        return out.reshape(batch, num_feats, frustrum_depth, img_shape[0], img_shape[1])
    return out


class ProjectionHelper:
    """deepvoxel/projection.py:5-105 -- same constructor arguments and attributes; `xp` (keyword) is the array module"""

    def __init__(self, lifting_intrinsic, projection_intrinsic, projection_image_dims, lifting_image_dims,
                 depth_min, depth_max, grid_dims, voxel_size, near_plane, frustrum_depth, device=None, verbose=True,
                 xp=None, lib=None):
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
        self.xp = xp if xp is not None else cupy
        self._lib = lib if lib is not None else _lib
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
        K = as_numpy(self.projection_intrinsic)
        return DvParams(int(self.projection_image_dims[0]), int(self.projection_image_dims[1]),
                        int(self.frustrum_depth), int(self.grid_dims[2]), float(K[0][0]), float(K[1][1]),
                        float(K[0][2]), float(K[1][2]), float(np.float32(self.voxel_size)),
                        float(np.float32(self.near_plane)))

    def compute_proj_idcs(self, cam2world, grid2world=None):
        """projection.py:48-105 -> (lin_ind_frustrum int32 (M,), voxel_coords fp32 (3,M)) or None"""
        xp = self.xp
        cam = np.asarray(as_numpy(cam2world), dtype="float32")
        cam = xp.asarray(np.ascontiguousarray(cam.reshape(16)), dtype="float32")
        P = self.params()
        n = P.W * P.H * P.D
        lin = xp.empty((n,), dtype="int32")
        vc = xp.empty((3, n), dtype="float32")
        nbytes = self._lib.load().rgbd_dv_workspace_bytes(ctypes.byref(P))
        if self._ws is None or self._ws.size < nbytes:
            self._ws = xp.empty(int(nbytes), dtype="uint8")
        M = ctypes.c_int(0)
        if grid2world is not None:
            # :53-54 world2grid = xp.linalg.inv(grid2world); :83-84 the second product per element runs on the device in the
            # reference's order (rgbd_dv_compute_proj_idcs_g2w)
            w2g = xp.asarray(np.ascontiguousarray(np.linalg.inv(as_numpy(grid2world)), dtype="float32").reshape(16), dtype="float32")
            self._lib.call("rgbd_dv_compute_proj_idcs_g2w", ctypes.byref(P), _ptr(cam), _ptr(w2g), _ptr(lin), _ptr(vc),
                           ctypes.byref(M), _ptr(self._ws), int(self._ws.size), _stream(xp))
        else:
            self._lib.call("rgbd_dv_compute_proj_idcs", ctypes.byref(P), _ptr(cam), _ptr(lin), _ptr(vc), ctypes.byref(M),
                           _ptr(self._ws), int(self._ws.size), _stream(xp))
        if M.value == 0:
            print('error: nothing in frustum bounds')   # :98-100
            return None
        return lin[:M.value], vc[:, :M.value]

    def project(self, grid, cam2world):
        """fused batch path: grid (B,F,G,G,G), cam2world (B,4,4) -> frustum (B,F,D,H,W); differentiable in grid"""
        return Project(cam2world, self.params(), xp=self.xp, lib=self._lib).apply((grid,))[0]

    def render_accumulative(self, grid, cam2world, W1, b1, W2, b2, accmulative_threshold=4):
        """"next" row: projection + AccumulativeOcclusionNet + collapse + depth map in one fused pass"""
        node = RenderAccumulative(cam2world, self.params(), accmulative_threshold, xp=self.xp, lib=self._lib)
        return node.apply((grid, W1, b1, W2, b2))


class RenderAccumulative(FunctionNode):
    """"Next" row: the per-sample loop of DeepVoxels.forward with `occlusion_type: accumulative`
    (deepvoxel.py:879-892 + AccumulativeOcclusionNet.forward :574-587 + depth rescale :903-904) as one node.
    inputs: (deepvoxels (B,F,G,G,G), W1 (nf,F+1), b1 (nf,), W2 (1,nf), b2 (1,)) -- the `.c.W` (1x1x1 kernels squeezed)
    and `.c.b` of the two EqualizedConv3d of `occlusion_net.occlusion`; outputs: (novel_views (B,F,H,W),
    depth_maps (B,1,H,W), foreground_weight (B,1,H,W))."""

    def __init__(self, cam2world, dv_params, accmulative_threshold=4, xp=None, lib=None):
        self.xp = xp if xp is not None else cupy
        self.lib = lib if lib is not None else _lib
        self.cam = self.xp.asarray(np.asarray(as_numpy(cam2world), dtype="float32").reshape(-1, 16), dtype="float32")
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
        grid, W1, b1, W2, b2 = (xp.ascontiguousarray(_arr(v), dtype="float32") for v in self.get_retained_inputs())
        B, F = grid.shape[:2]
        P = self.P
        arr = lambda g, shape: xp.zeros(shape, "float32") if g is None else _f32c(xp, g)
        g_novel, g_depth = arr(grad_outputs[0], (B, F, P.H, P.W)), arr(grad_outputs[1], (B, 1, P.H, P.W))
        g_fg = None if grad_outputs[2] is None else arr(grad_outputs[2], (B, 1, P.H, P.W))
        outs = [xp.empty_like(a) for a in (grid, W1, b1, W2, b2)]
        self.lib.call("rgbd_dv_render_bwd", ctypes.byref(P), ctypes.byref(self.R), _ptr(grid), _ptr(self.cam), _ptr(W1),
                      _ptr(b1), _ptr(W2), _ptr(b2), B, F, _ptr(self.saved), _ptr(g_novel), _ptr(g_depth), _ptr(g_fg),
                      *[_ptr(o) for o in outs], _ptr(self.ws), int(self.ws.size), _stream(xp))
        return _select(outs, target_input_indexes)


__all__ = ["LossFuncRotate", "warp", "inv_warp", "bilinear", "ProjectionHelper", "interpolate_trilinear",
           "ConsistencyLoss", "Warp", "Bilinear", "Project", "Trilinear", "RenderAccumulative", "DepthHead", "depth_head",
           "HAVE_CHAINER"]
