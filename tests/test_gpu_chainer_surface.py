"""GPU: the Chainer / CuPy call surface (rgbd_gan_b200/chainer_nodes.py) driven with CuPy-STYLE arrays on the real
library.  CuPy and Chainer cannot be installed here, so:
  * `TorchXP` below is a minimal cupy-like array module whose arrays are thin wrappers over torch CUDA memory exposing
    exactly what CuPy arrays expose to the glue (`.data.ptr`, `__cuda_array_interface__`, shape / size / dtype, slicing,
    `get()`), plus `xp.cuda.get_current_stream().ptr`;
  * `chainer` is the Chainer-v7 shim of tests/golden (FunctionNode.apply / Variable / reverse walk).
Every symbol of SURVEY.md 8(b) goes through the same C-ABI entry points as the torch glue and is compared with the
reference's golden vectors: bit-exact geometry / masks / indices / sampled values, 1e-5 on losses and gradients."""
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN, assert_grad_close, case_options, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

sys.path.insert(0, GOLDEN)
import chainer_shim  # noqa: E402

_DT = {"float32": "float32", "float64": "float64", "uint8": "uint8", "int32": "int32", "bool": "bool"}


def _tdt(dtype):
    return getattr(torch, _DT[np.dtype(dtype).name if not isinstance(dtype, str) else dtype])


class _Mem:
    def __init__(self, ptr):
        self.ptr = ptr


class TArr:
    """what a cupy.ndarray shows to the glue, over a torch CUDA tensor"""

    def __init__(self, t):
        self.t = t

    shape = property(lambda s: tuple(s.t.shape))
    size = property(lambda s: s.t.numel())
    ndim = property(lambda s: s.t.dim())
    dtype = property(lambda s: np.dtype(str(s.t.dtype).replace("torch.", "")))
    data = property(lambda s: _Mem(s.t.data_ptr()))

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.t.data_ptr(), False), "version": 2}

    def __len__(self):
        return self.t.shape[0]

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return TArr(self.t.reshape(tuple(shape)))

    def __getitem__(self, k):
        return TArr(self.t[k])

    def astype(self, dtype, copy=True):
        return TArr(self.t.to(_tdt(dtype)))

    def __add__(self, o):
        return TArr(self.t + (o.t if isinstance(o, TArr) else o))

    def get(self):
        return self.t.detach().cpu().numpy()


class TorchXP(types.ModuleType):
    def __init__(self):
        super().__init__("torch_backed_cupy")
        self.cuda = types.SimpleNamespace(
            get_current_stream=lambda: types.SimpleNamespace(ptr=torch.cuda.current_stream().cuda_stream))

    @staticmethod
    def _t(a, dtype=None):
        if isinstance(a, TArr):
            t = a.t
        elif isinstance(a, torch.Tensor):
            t = a
        else:
            t = torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")
        return t if dtype is None else t.to(_tdt(dtype))

    def asarray(self, a, dtype=None):
        return TArr(self._t(a, dtype))

    def ascontiguousarray(self, a, dtype=None):
        return TArr(self._t(a, dtype).contiguous())

    def empty(self, shape, dtype="float32"):
        t = torch.empty(shape, dtype=_tdt(dtype), device="cuda:0")
        if t.is_floating_point():
            t.fill_(float("nan"))
        return TArr(t)

    def zeros(self, shape, dtype="float32"):
        return TArr(torch.zeros(shape, dtype=_tdt(dtype), device="cuda:0"))

    def empty_like(self, a):
        return self.empty(a.shape, a.dtype.name)


@pytest.fixture()
def nodes():
    chainer_shim.install()
    for m in [k for k in sys.modules if k.startswith("rgbd_gan_b200.chainer_nodes")]:
        del sys.modules[m]
    import rgbd_gan_b200.chainer_nodes as cn       # picks up the shim as `chainer`; the library is the real one
    return cn


def _V(xp, a, **kw):
    return chainer_shim.Variable(xp.asarray(np.ascontiguousarray(a)), **kw)


def _backward(xp, out, g):
    out.grad = xp.asarray(np.asarray(g, dtype=np.float32).reshape(out.shape))
    out.backward()


@pytest.mark.parametrize("name", ["loss_cfg0_l1_occ", "loss_dv_maxdepth", "loss_s32_l2_feat"])
@pytest.mark.parametrize("grad_scale", [None, 2.0])
def test_loss_node_on_cupy_style_arrays(nodes, name, grad_scale):
    g = load_golden(name)
    o = case_options(g)
    B = o["B"]
    xp = TorchXP()
    f = nodes.LossFuncRotate(xp, K=None if o["K"] is None else o["K"].copy(), norm=o["norm"], lambda_geometric=o["lam"],
                             grad_scale=grad_scale)
    img, img_rot = _V(xp, g["x"][:B]), _V(xp, g["x"][B:])
    kw = dict(occlusion_aware=o["occ"])
    if o["max_depth"] is not None:
        kw["max_depth"] = o["max_depth"]
    if o["min_depth"] is not None:
        kw["min_depth"] = o["min_depth"]
    loss, zp = f(img, g["cam"][:B], img_rot, g["cam"][B:], **kw)
    _backward(xp, loss, o["gy"])
    torch.cuda.synchronize()
    assert abs(float(loss.array.get()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    np.testing.assert_array_equal(zp.array.get(), g["new_zp_cat"])
    assert_grad_close(img.grad.get(), g["g_img"])
    assert_grad_close(img_rot.grad.get(), g["g_img_rot"])
    np.testing.assert_array_equal(f.K, g["K"])


def test_debug_tuple_and_free_functions_on_cupy_style_arrays(nodes, oracle_mod):
    g = load_golden("loss_cfg0_l1_occ")
    o = case_options(g)
    B, S = o["B"], o["S"]
    xp = TorchXP()
    f = nodes.LossFuncRotate(xp, lambda_geometric=o["lam"])
    img, img_rot = _V(xp, g["x"][:B]), _V(xp, g["x"][B:])
    warped, not_out, new_zp, warped_rot, not_out_rot, new_zp_rot = f(img, g["cam"][:B], img_rot, g["cam"][B:], debug=True)
    np.testing.assert_array_equal(warped.array.get(), g["warped"])
    np.testing.assert_array_equal(warped_rot.array.get(), g["warped_rot"])
    np.testing.assert_array_equal(not_out.get(), g["not_out"])
    np.testing.assert_array_equal(not_out_rot.get(), g["not_out_rot"])
    np.testing.assert_array_equal(np.concatenate([new_zp.array.get(), new_zp_rot.array.get()]), g["new_zp_cat"])
    th, thr = g["cam"][:B], g["cam"][B:]
    R = np.matmul(thr[:, :3, :3].transpose(0, 2, 1), th[:, :3, :3]).astype("float32")
    t = np.matmul(th[:, :3, :3].transpose(0, 2, 1), thr[:, :3, -1:] - th[:, :3, -1:]).astype("float32")
    z = _V(xp, g["x"][:B, -1:].reshape(B, 1, -1))
    zp = nodes.warp(f.K, f.inv_K, R, t, z, f.p, xp=xp)
    zpr = nodes.inv_warp(f.K, f.inv_K, R.transpose(0, 2, 1), t, _V(xp, g["x"][B:, -1:].reshape(B, 1, -1)), f.p, xp=xp)
    np.testing.assert_array_equal(np.concatenate([zp.array.get(), zpr.array.get()]), g["new_zp_cat"])
    w, m = nodes.bilinear(img_rot, zp, xp=xp)
    np.testing.assert_array_equal(w.array.get(), g["warped"])
    np.testing.assert_array_equal(m.get(), g["not_out"])
    gw = np.random.default_rng(0).normal(size=w.shape).astype(np.float32)
    _backward(xp, w, gw)
    torch.cuda.synchronize()
    ref_gi, ref_gzp = oracle_mod.bilinear_bwd(g["x"][B:], g["new_zp_cat"][:B], gw)
    assert_grad_close(img_rot.grad.get(), ref_gi)
    from rgbd_gan_b200.host_math import warp_constants
    M, _ = warp_constants(f.K, f.inv_K, R, t, False)
    assert_grad_close(z.grad.get().reshape(B, 1, -1), oracle_mod.warp_bwd(ref_gzp, M, S, S))


def test_projection_surface_on_cupy_style_arrays(nodes):
    g = load_golden("dv_g16_f3")
    G, img, F, D = int(g["G"]), int(g["img"]), int(g["F"]), int(g["D"])
    xp = TorchXP()
    h = nodes.ProjectionHelper(g["intrinsic"], g["intrinsic"], [img, img], [img, img], 0., 1., [G] * 3,
                               float(g["voxel_size"]), float(g["near_plane"]), D, verbose=False, xp=xp)
    for i in range(g["cam"].shape[0]):
        lin, vc = h.compute_proj_idcs(g["cam"][i])
        np.testing.assert_array_equal(lin.get(), g["lin_ind_%d" % i])
        np.testing.assert_array_equal(vc.get(), g["voxel_coords_%d" % i])
        grid = _V(xp, g["grid"][i:i + 1])
        out = nodes.interpolate_trilinear(grid, lin, vc, [img, img], D, xp=xp)
        assert out.shape == (1, F, D, img, img)
        np.testing.assert_array_equal(out.array.get(), g["frustum_%d" % i])
        _backward(xp, out, g["g_out"][i:i + 1])
        torch.cuda.synchronize()
        assert_grad_close(grid.grad.get(), g["g_grid_%d" % i])
    far = g["cam"][0].copy()
    far[:3, 3] += 100.0
    assert h.compute_proj_idcs(far) is None
    os.environ["RGBD_B200_DV_EXACT"] = "1"                     # the reference's ((v*wx)*wy)*wz order: bit-exact frustum
    try:
        grid = _V(xp, g["grid"])
        fr = h.project(grid, g["cam"])
        _backward(xp, fr, g["g_out"])
        torch.cuda.synchronize()
    finally:
        del os.environ["RGBD_B200_DV_EXACT"]
    for i in range(g["cam"].shape[0]):
        np.testing.assert_array_equal(fr.array.get()[i], g["frustum_%d" % i][0])
        assert_grad_close(grid.grad.get()[i], g["g_grid_%d" % i][0])


def test_render_node_on_cupy_style_arrays(nodes):
    from rgbd_gan_b200._lib import DvParams
    g = load_golden("render_g12_thr3")
    G, img, D = int(g["G"]), int(g["img"]), int(g["D"])
    P = DvParams(img, img, D, G, 2. * img, 2. * img, img / 2., img / 2., float(np.float32(g["voxel_size"])),
                 float(np.float32(g["near_plane"])))
    xp = TorchXP()
    node = nodes.RenderAccumulative(g["cam"], P, float(g["threshold"]), xp=xp)
    ins = [_V(xp, g[k]) for k in ("grid", "W1", "b1", "W2", "b2")]
    novel, depth, fg = node.apply(tuple(ins))
    rel = lambda a, b: float(np.abs(np.asarray(a) - b).max() / np.abs(b).max())
    assert rel(novel.array.get(), g["novel"]) <= 1e-5 and rel(depth.array.get(), g["depth"]) <= 1e-5
    novel.grad, depth.grad, fg.grad = (xp.asarray(np.ascontiguousarray(g[k]).reshape(v.shape)) for k, v in
                                       (("g_novel", novel), ("g_depth", depth), ("g_fg", fg)))
    grads = node.backward((0, 1, 2, 3, 4), tuple(chainer_shim.Variable(v.grad, requires_grad=False) for v in (novel, depth, fg)))
    torch.cuda.synchronize()
    for gv, key in zip(grads, ("g_grid", "g_W1", "g_b1", "g_W2", "g_b2")):
        assert rel(gv.array.get().reshape(g[key].shape), g[key]) <= 1e-5, key


def test_pose_pipeline_on_cupy_style_arrays(nodes):
    """SURVEY 8f rank 4 on the Chainer / CuPy surface: device cam2world matrices (what the reference's updater holds after
    xp.array(get_camera_matries(thetas)), updater.py:315) go through rgbd_pose_algebra -- no .get() -- and give the
    golden new_zp bit for bit; get_camera_matries / CameraParamPrior work on cupy-style arrays"""
    from oracle import numpy_port as npp
    g = load_golden("loss_cfg0_l1_occ")
    o = case_options(g)
    B = o["B"]
    xp = TorchXP()
    f = nodes.LossFuncRotate(xp, lambda_geometric=o["lam"], grad_scale=1.0)
    img, img_rot = _V(xp, g["x"][:B]), _V(xp, g["x"][B:])
    cam = xp.asarray(g["cam"])
    loss, zp = f(img, cam[:B], img_rot, cam[B:], occlusion_aware=o["occ"])
    torch.cuda.synchronize()
    np.testing.assert_array_equal(zp.array.get(), g["new_zp_cat"])
    assert abs(float(loss.array.get()) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    th = g["thetas"]
    cs = np.concatenate([np.cos(th[:, :3]), np.sin(th[:, :3])], axis=1).astype(np.float32)
    got = nodes.get_camera_matries(xp.asarray(th), cos_sin=xp.asarray(cs), xp=xp).get()
    np.testing.assert_allclose(got, g["cam"], rtol=0, atol=5e-7)
    cfg = types.SimpleNamespace(x_rotate=0.3054, y_rotate=1.0472, z_rotate=0, x_translate=0, y_translate=0, z_translate=0,
                                uniform_distribution=False)
    np.random.seed(4)
    u, e, s = np.random.uniform(-1, 1, (B, 6)), np.random.uniform(0, 0.5, (B, 6)), np.random.choice(2, (B, 3))
    np.random.seed(4)
    want = npp.sample_camera_prior(2 * B, npp.FFHQ_RANGES, False)
    got = nodes.CameraParamPrior(cfg, xp=xp).sample(2 * B, draws=np.concatenate([u, e, s.astype(np.float64)], 1)).get()
    np.testing.assert_array_equal(got, want)
