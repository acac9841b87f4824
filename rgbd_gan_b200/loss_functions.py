"""Host-side mirror of the reference's common/loss_functions.py call surface.

Same names, argument meaning and return values as the reference
(`LossFuncRotate`, `warp`, `inv_warp`, `bilinear`; reference lines cited per function),
but every array operation of the reference's Chainer graph is replaced by calls into
librgbdgan_b200.so (hand-written sm_100a kernels) through ctypes.  Device memory, streams
and autograd hooks come from torch in this image (the reference's Chainer/CuPy are not
installable here); `chainer_nodes.py` holds the same glue as Chainer FunctionNodes for a
deployment that has Chainer + CuPy.  There is no CPU fallback: inputs must be CUDA tensors.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import DvParams, LossOpts  # noqa: F401
from .host_math import as_numpy as _as_numpy
from .host_math import combine_loss_parts, grid_dims as _grid_dims, intrinsics_for_size, pixel_grid, pose_algebra, warp_constants


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_f32(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError("%s must be a CUDA tensor (rgbd_gan_b200 has no CPU path)" % what)
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (what, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


class _PoseUploader:
    """Packs M|c|Mi|ci (24 floats per pair) into pinned memory and copies it asynchronously."""

    def __init__(self):
        self._slots = {}

    def upload(self, M, c, Mi, ci, device):
        B = M.shape[0]
        key = (B, str(device))
        if key not in self._slots:
            self._slots[key] = dict(bufs=[torch.empty(24 * B, dtype=torch.float32).pin_memory() for _ in range(2)],
                                    events=[None, None], k=0)
        s = self._slots[key]
        k = s["k"]
        s["k"] = 1 - k
        if s["events"][k] is not None:
            s["events"][k].synchronize()          # the copy that last used this slot has finished
        host = s["bufs"][k]
        hv = host.numpy()
        hv[0:9 * B] = M.reshape(-1)
        hv[9 * B:12 * B] = c.reshape(-1)
        hv[12 * B:21 * B] = Mi.reshape(-1)
        hv[21 * B:24 * B] = ci.reshape(-1)
        dev = host.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        s["events"][k] = ev
        return dev


def unpack_poses(poses, B):
    return (poses[:9 * B].view(B, 3, 3), poses[9 * B:12 * B].view(B, 3, 1), poses[12 * B:21 * B].view(B, 3, 3),
            poses[21 * B:24 * B].view(B, 3, 1))


def _pose_ptrs(dev, B):
    base = dev.data_ptr()
    return [ctypes.c_void_p(base + 4 * off) for off in (0, 9 * B, 12 * B, 21 * B)]


class _ConsistencyFn(torch.autograd.Function):
    """LossFuncRotate.__call__ body (common/loss_functions.py:93-146) as one autograd node."""

    @staticmethod
    def forward(ctx, img, img_rot, owner, poses, opts, want_zp):
        B, C, H, W = img.shape
        dev = img.device
        ws = owner._workspace(B, C, H, W, dev)
        parts = torch.empty(8, dtype=torch.float32, device=dev)
        new_zp = torch.empty((2 * B, H * W, 3), dtype=torch.float32, device=dev) if want_zp else None
        pp = _pose_ptrs(poses, B)
        need_grad = img.requires_grad or img_rot.requires_grad
        ctx.owner, ctx.opts, ctx.poses, ctx.shape = owner, opts, poses, (B, C, H, W)
        # one pass for loss + gradients whenever a gradient will be asked for; the expected upstream gradient
        # defaults to 1 and backward rescales on the device if a different one arrives
        ctx.fused = need_grad and owner.fuse_backward
        ctx.expected_gy = 1.0 if owner.grad_scale is None else owner.grad_scale
        ctx.set_materialize_grads(False)
        if ctx.fused:
            g_img, g_img_rot = torch.empty_like(img), torch.empty_like(img_rot)
            _lib.call("rgbd_consistency_fwd_bwd", _ptr(img), _ptr(img_rot), *pp, B, C, H, W, ctypes.byref(opts),
                      ctypes.c_float(ctx.expected_gy), _ptr(parts), _ptr(new_zp), _ptr(g_img), _ptr(g_img_rot),
                      _ptr(ws), ws.numel(), _stream())
            ctx.stash = (g_img, g_img_rot)
        else:
            _lib.call("rgbd_consistency_fwd", _ptr(img), _ptr(img_rot), *pp, B, C, H, W, ctypes.byref(opts),
                      _ptr(parts), _ptr(new_zp), None, _ptr(ws), ws.numel(), _stream())
        ctx.save_for_backward(img, img_rot)
        owner.last_loss_parts = parts
        k = 6 if opts.hinge_lambda > 0 and opts.hinge_depth_min == opts.hinge_depth_min else 4
        if owner.process_group is not None:
            loss = owner._allreduce_combine(parts, k == 6)
        elif owner.defer_loss:
            loss = parts[k]                   # becomes valid after owner.peer_comm.wait(); do not touch it before
        else:
            loss = parts[k].clone()
        if new_zp is None:
            return loss, None
        return loss, new_zp

    @staticmethod
    def backward(ctx, g_loss, g_new_zp=None):
        img, img_rot = ctx.saved_tensors
        B, C, H, W = ctx.shape
        owner = ctx.owner
        if g_loss is None:
            g_loss = torch.zeros((), dtype=torch.float32, device=img.device)
        g_loss = g_loss.to(torch.float32).contiguous()
        if ctx.fused and g_new_zp is None and ctx.stash is not None:
            g_img, g_img_rot = ctx.stash
            ctx.stash = None                  # the stash is handed out once; a second backward recomputes
            _lib.call("rgbd_consistency_rescale", _ptr(g_img), _ptr(g_img_rot), g_img.numel(), _ptr(g_loss),
                      ctypes.c_float(ctx.expected_gy), _stream())
            return g_img, g_img_rot, None, None, None, None
        ws = owner._workspace(B, C, H, W, img.device)
        g_img, g_img_rot = torch.empty_like(img), torch.empty_like(img_rot)
        gz = None if g_new_zp is None else g_new_zp.to(torch.float32).contiguous()
        _lib.call("rgbd_consistency_bwd", _ptr(img), _ptr(img_rot), *_pose_ptrs(ctx.poses, B), B, C, H, W,
                  ctypes.byref(ctx.opts), ctypes.c_float(1.0), _ptr(g_loss), _ptr(gz), _ptr(g_img), _ptr(g_img_rot),
                  _ptr(ws), ws.numel(), _stream())
        return g_img, g_img_rot, None, None, None, None


class LossFuncRotate:
    """Mirror of the reference class (common/loss_functions.py:31-168).

    Extra keyword-only arguments (not in the reference):
      fuse_backward   -- (default True) when an input requires grad, forward produces the loss AND both
                         image gradients in one pass (rgbd_consistency_fwd_bwd); backward then only
                         compares the upstream gradient that arrives with the expected one on the device
                         and rescales the stashed gradients if it differs.  False: forward computes the
                         loss only and backward recomputes (rgbd_consistency_fwd / _bwd).
      grad_scale      -- the expected upstream gradient of the fused path (default 1).  In the
                         reference's loop it is the constant lambda_rotate (updater.py:363-365); giving
                         it here makes the rescale pass a no-op.
      return_new_zp   -- materialise the second return value (:146).  No reference caller uses
                         it (updater.py:340 drops it); False returns None in its place.
      process_group   -- torch.distributed group over which the PAIRS are sharded; the four
                         loss means are all-reduced (4 floats, NCCL), gradients need no communication.
      n_pairs_global  -- total number of pairs over all shards (default: local pairs x world size,
                         i.e. equal shards).
      defer_loss      -- with peer_comm: the loss exchange is not joined into the current stream; the
                         returned loss is valid after `peer_comm.wait()` (gradients are unaffected).  The
                         exchange then overlaps the rest of the step, as a training loop that only logs
                         the loss (updater.py:361, chainer.report) allows.  defer_loss="lazy": every call only
                         publishes its parts to the peers from the kernel that finishes the loss (no wait, no
                         side stream); `peer_comm.wait()` sums the LATEST call's parts of all ranks -- the GPUs
                         are not coupled step by step (at most 7 calls apart).
      peer_comm       -- rgbd_gan_b200.distributed.PeerComm: same sharding, but the 4 floats are
                         exchanged inside the loss kernel over NVLink peer memory (no NCCL launch).
    """

    def __init__(self, xp=None, K=None, norm="l1", lambda_geometric=3, *, grad_scale=None, return_new_zp=True,
                 process_group=None, peer_comm=None, n_pairs_global=None, defer_loss=False, fuse_backward=True):
        self.xp = xp
        self.size = None
        self.K = K
        self.norm = norm
        self.lambda_geometric = lambda_geometric
        self.inv_K = None
        self.p = None
        self.grad_scale = None if grad_scale is None else float(grad_scale)
        self.fuse_backward = bool(fuse_backward)
        self.return_new_zp = return_new_zp
        self.process_group = process_group
        self.peer_comm = peer_comm
        self.n_pairs_global = n_pairs_global
        # False / True (all-reduce on a side stream, joined by peer_comm.wait()) / "lazy" (publish only, summed by wait())
        self.defer_loss = (2 if defer_loss == "lazy" else int(bool(defer_loss))) if peer_comm is not None else 0
        if peer_comm is not None and process_group is not None:
            raise ValueError("give either process_group (NCCL all-reduce) or peer_comm (fused), not both")
        self.last_loss_parts = None
        self._uploader = _PoseUploader()
        self._ws = {}

    # -- :39-61
    def init_params(self, xp=None, size=4):
        self.K, self.inv_K = intrinsics_for_size(self.K, size, first=self.size is None)   # in place after the first call (Q9)
        self.size = size
        self.p = pixel_grid(size)

    def _workspace(self, B, C, H, W, device):
        key = (B, C, H, W, str(device))
        ws = self._ws.get(key)
        if ws is None:
            nbytes = _lib.load().rgbd_consistency_workspace_bytes(B, C, H, W)
            if self._ws and self.peer_comm is not None and self.defer_loss == 1:
                # a deferred exchange of the previous call may still read the old workspace's partial sums on the
                # comm's side stream: order it before the old buffer can be reused by the allocator
                self.peer_comm.wait()
            ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
            self._ws = {key: ws}                       # keep one size only
        return ws

    def _opts(self, occlusion_aware, max_depth, min_depth, B, depth_hinge=None):
        world = 1
        if self.process_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(self.process_group)
        peer = None
        if self.peer_comm is not None:
            world, peer = self.peer_comm.world, self.peer_comm.handle
        return LossOpts(_lib.NORM_L1 if self.norm == "l1" else _lib.NORM_L2, int(bool(occlusion_aware)),
                        float("nan") if max_depth is None else float(max_depth),
                        float("nan") if min_depth is None else float(min_depth),
                        float(self.lambda_geometric),
                        int(self.n_pairs_global) if self.n_pairs_global else int(B) * world, peer,
                        int(self.defer_loss), 0,
                        float("nan") if depth_hinge is None else float(depth_hinge[0]),
                        0.0 if depth_hinge is None else float(depth_hinge[1]))

    def _allreduce_combine(self, parts, with_hinge=False):
        """sum the per-shard means over the group, then combine as :141-144 (fp32) [+ the depth hinge term]"""
        import torch.distributed as dist
        p = parts[:6].clone()
        dist.all_reduce(p, op=dist.ReduceOp.SUM, group=self.process_group)
        loss = combine_loss_parts(p, self.lambda_geometric)
        return loss + p[5] if with_hinge else loss

    # -- :63-146
    def __call__(self, img, theta, img_rot, theta_rot, occlusion_aware=False, debug=False, max_depth=None,
                 min_depth=None, depth_hinge=None, poses=None):
        """Reference signature (:63-64) plus `depth_hinge=(depth_min, lambda_depth)`: the term the updaters add
        right after this call, loss_rotate += mean(relu(depth_min - x_fake[:, -1]) ** 2) * lambda_depth
        (updater.py:357-359), evaluated inside the same kernels; the returned loss then includes it.
        `poses`: the packed pose constants of pose_pipeline.PosePipeline.step / pose_algebra_device (theta and
        theta_rot are then ignored).  Host (NumPy) thetas take the reference's NumPy matmul sequence on the host;
        CUDA-tensor thetas take the device pose kernel (bit-identical on the golden vectors, no synchronisation)."""
        img = _dev_f32(img, "img")
        img_rot = _dev_f32(img_rot, "img_rot")
        if img.shape != img_rot.shape or img.dim() != 4:
            raise ValueError("img and img_rot must both be (B,C,H,W)")
        if img.shape[-1] != img.shape[-2]:
            raise ValueError("the reference's intrinsics assume square images (loss_functions.py:48-61)")
        if self.size != img.shape[-1]:
            self.init_params(self.xp, size=img.shape[-1])
        B, C, H, W = img.shape
        if poses is None and isinstance(theta, torch.Tensor) and theta.is_cuda and not debug:
            # cam2world matrices already on the device (the reference's production case: xp.array(...) thetas,
            # updater.py:315): R, t and the warp constants come from one small kernel, no D2H sync, no upload
            from .pose_pipeline import pose_algebra_device
            poses = pose_algebra_device(self.K, self.inv_K, theta.detach(), theta_rot.detach())
        if poses is None:
            M, c, Mi, ci = pose_algebra(self.K, self.inv_K, theta, theta_rot)
            if debug:
                return self._debug(img, img_rot, M, c, Mi, ci)
            with torch.cuda.device(img.device):
                poses = self._uploader.upload(M, c, Mi, ci, img.device)
        elif debug:
            M, c, Mi, ci = (a.cpu().numpy() for a in unpack_poses(poses, B))
            return self._debug(img, img_rot, M, c, Mi, ci)
        elif poses.numel() != 24 * B or not poses.is_cuda:
            raise ValueError("poses must be the packed (24*B,) CUDA tensor of pose_pipeline.pose_algebra_device")
        opts = self._opts(occlusion_aware, max_depth, min_depth, B, depth_hinge)
        with torch.cuda.device(img.device):
            return _ConsistencyFn.apply(img, img_rot, self, poses, opts, bool(self.return_new_zp))

    def _debug(self, img, img_rot, M, c, Mi, ci):
        """:100-102 -- warped images / masks for eyeballing"""
        B, C, H, W = img.shape
        dev = img.device
        z = img[:, -1:].reshape(B, 1, -1)
        z_rot = img_rot[:, -1:].reshape(B, 1, -1)
        with torch.cuda.device(dev):
            new_zp = _WarpFn.apply(z, torch.from_numpy(M).to(dev), torch.from_numpy(c).to(dev), H, W)
            new_zp_rot = _WarpFn.apply(z_rot, torch.from_numpy(Mi).to(dev), torch.from_numpy(ci).to(dev), H, W)
        warped, not_out = bilinear(img_rot, new_zp)
        warped_rot, not_out_rot = bilinear(img, new_zp_rot)
        return warped, not_out, new_zp, warped_rot, not_out_rot, new_zp_rot

    # -- :148-158 (no caller in the reference; plain array-library math)
    def calc_real_pos(self, img, theta):
        theta = _as_numpy(theta)
        if theta.ndim == 1:
            assert False, "only rotation matrices are supported for theta"
        dev = img.device
        R = torch.from_numpy(np.ascontiguousarray(theta[:, :3, :3], dtype=np.float32)).to(dev)
        t = torch.from_numpy(np.ascontiguousarray(theta[:, :3, -1:], dtype=np.float32)).to(dev)
        z = img[:, -1:].detach().reshape(img.shape[0], 1, -1)
        rgb = img[:, :3].detach().reshape(img.shape[0], 3, -1)
        inv_K, p = torch.from_numpy(self.inv_K).to(dev), torch.from_numpy(self.p).to(dev)
        real_pos = torch.matmul(torch.matmul(R, inv_K), z * p) + t
        return torch.cat([rgb, real_pos], dim=1)

    # -- :160-168 (only with `use_occupancy_net_loss`, absent from every shipped config)
    def occupancy_net_loss(self, occupancy_net, depth, theta, z):
        theta = _as_numpy(theta)
        dev = depth.device
        R = torch.from_numpy(np.ascontiguousarray(theta[:, :3, :3], dtype=np.float32)).to(dev)
        t = torch.from_numpy(np.ascontiguousarray(theta[:, :3, -1:], dtype=np.float32)).to(dev)
        depth = depth.reshape(depth.shape[0], 1, -1)
        eps = torch.randn(depth.shape, device=dev) * 0.05
        inv_K, p = torch.from_numpy(self.inv_K).to(dev), torch.from_numpy(self.p).to(dev)
        real_pos = torch.matmul(torch.matmul(R, inv_K), (depth + eps) * p) + t
        label = (eps > 0).reshape(-1, 1).to(torch.float32)
        occupancy_field = occupancy_net(z, real_pos + eps)
        return torch.nn.functional.binary_cross_entropy_with_logits(occupancy_field, label)


# ------------------------------------------------------------------------------ free functions
class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, M, cv, H, W):
        B = z.shape[0]
        zc = _dev_f32(z, "z").reshape(B, H * W)
        M = M.contiguous()
        cv = cv.contiguous()
        out = torch.empty((B, H * W, 3), dtype=torch.float32, device=z.device)
        _lib.call("rgbd_warp_fwd", _ptr(zc), _ptr(M), _ptr(cv), B, H, W, _ptr(out), _stream())
        ctx.M, ctx.dims, ctx.zshape = M, (B, H, W), z.shape
        return out

    @staticmethod
    def backward(ctx, g):
        B, H, W = ctx.dims
        g = g.to(torch.float32).contiguous()
        gz = torch.empty((B, H * W), dtype=torch.float32, device=g.device)
        _lib.call("rgbd_warp_bwd", _ptr(g), _ptr(ctx.M), B, H, W, _ptr(gz), _stream())
        return gz.reshape(ctx.zshape), None, None, None, None


def warp(K, inv_K, R, t, z, p):
    """common/loss_functions.py:171-175: (K R K^-1)(z p) - (K R) t, returned as (B,HW,3). Differentiable in z."""
    H, W = _grid_dims(p, z.shape[-1])
    M, cv = warp_constants(K, inv_K, R, t, inverse=False)
    with torch.cuda.device(z.device):
        return _WarpFn.apply(z, torch.from_numpy(M).to(z.device), torch.from_numpy(cv).to(z.device), H, W)


def inv_warp(K, inv_K, inv_R, t, z, p):
    """common/loss_functions.py:178-182: (K R^T K^-1)(z p) + K t, returned as (B,HW,3). Differentiable in z."""
    H, W = _grid_dims(p, z.shape[-1])
    M, cv = warp_constants(K, inv_K, inv_R, t, inverse=True)
    with torch.cuda.device(z.device):
        return _WarpFn.apply(z, torch.from_numpy(M).to(z.device), torch.from_numpy(cv).to(z.device), H, W)


class _BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, zp):
        B, C, H, W = img.shape
        warped = torch.empty((B * H * W, C), dtype=torch.float32, device=img.device)
        mask = torch.empty(B * H * W, dtype=torch.uint8, device=img.device)
        _lib.call("rgbd_bilinear_fwd", _ptr(img), _ptr(zp), B, C, H, W, _ptr(warped), _ptr(mask), _stream())
        ctx.save_for_backward(img, zp)
        mask = mask.bool()
        ctx.mark_non_differentiable(mask)
        return warped, mask

    @staticmethod
    def backward(ctx, g_warped, _g_mask=None):
        img, zp = ctx.saved_tensors
        B, C, H, W = img.shape
        g_warped = g_warped.to(torch.float32).contiguous()
        g_img = torch.empty_like(img)
        g_zp = torch.empty_like(zp)
        _lib.call("rgbd_bilinear_bwd", _ptr(img), _ptr(zp), _ptr(g_warped), B, C, H, W, _ptr(g_img), _ptr(g_zp),
                  _stream())
        return g_img, g_zp


def bilinear(img, zp):
    """common/loss_functions.py:185-228 -> (warped (B*HW, C), not_getting_out (B*HW,) bool)."""
    img = _dev_f32(img, "img")
    zp = _dev_f32(zp, "zp")
    b, hw, _ = zp.shape
    if img.shape[0] != b or img.shape[2] * img.shape[3] != hw:
        raise ValueError("zp must be (B, H*W, 3) matching img (B,C,H,W)")
    with torch.cuda.device(img.device):
        return _BilinearFn.apply(img, zp)


class _DepthHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h):
        B, C, H, W = h.shape
        out = torch.empty_like(h)
        _lib.call("rgbd_depth_head_fwd", _ptr(h), B, C, H, W, _ptr(out), _stream())
        ctx.save_for_backward(h)
        return out

    @staticmethod
    def backward(ctx, g_out):
        h, = ctx.saved_tensors
        B, C, H, W = h.shape
        g_out = g_out.to(torch.float32).contiguous()
        g_h = torch.empty_like(h)
        _lib.call("rgbd_depth_head_bwd", _ptr(h), _ptr(g_out), B, C, H, W, _ptr(g_h), _stream())
        return g_h


def depth_head(h):
    """"next" row: the generators' depth head (net.py:294-299, :756-761) as one kernel --
    F.concat([h[:, :-1], 1 / (F.softplus(h[:, -1:]) + 1e-4)]); differentiable in h"""
    h = _dev_f32(h, "h")
    with torch.cuda.device(h.device):
        return _DepthHeadFn.apply(h)


__all__ = ["LossFuncRotate", "warp", "inv_warp", "bilinear", "pose_algebra", "combine_loss_parts", "depth_head", "unpack_poses"]
