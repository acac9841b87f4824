"""CPU: the Chainer FunctionNode glue (rgbd_gan_b200/chainer_nodes.py) driven through the Chainer-v7
shim the golden generator uses, with the C-ABI replaced by a stand-in that executes each entry point
with the CPU oracle on the very pointers it is handed.  This checks what cannot be checked on the
GPU box (no Chainer/CuPy there either): argument order and shapes of every C-ABI call the nodes make,
retained state, the device-scalar upstream gradient, the fused grad_scale path, and that
`LossFuncRotate(...)(img, theta, img_rot, theta_rot)` + `loss.backward()` reproduce the reference."""
import ctypes
import os
import sys
import types

import numpy as np
import pytest

from conftest import GOLDEN, assert_grad_close, case_options, load_golden

sys.path.insert(0, GOLDEN)
import chainer_shim  # noqa: E402


class _Ptr:
    def __init__(self, a):
        self.ptr = a.ctypes.data


class HostArr(np.ndarray):
    """numpy array that looks like a cupy array to the glue (`.data.ptr`)"""

    @property
    def data(self):
        return _Ptr(self)


def _wrap(a):
    return np.ascontiguousarray(a).view(HostArr)


class FakeXP(types.ModuleType):
    def __init__(self):
        super().__init__("fake_cupy")
        self.cuda = types.SimpleNamespace(get_current_stream=lambda: types.SimpleNamespace(ptr=0))

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def asarray(a, dtype=None):
        return _wrap(np.asarray(a, dtype=dtype))

    @staticmethod
    def ascontiguousarray(a, dtype=None):
        return _wrap(np.ascontiguousarray(a, dtype=dtype))

    @staticmethod
    def empty(shape, dtype="float32"):
        return _wrap(np.full(shape, np.nan if np.dtype(dtype).kind == "f" else 0, dtype=dtype))

    @staticmethod
    def zeros(shape, dtype="float32"):
        return _wrap(np.zeros(shape, dtype=dtype))

    @staticmethod
    def empty_like(a):
        return _wrap(np.full(a.shape, np.nan, dtype=a.dtype))


def _view(p, shape, ctype=ctypes.c_float):
    p = p.value if isinstance(p, ctypes.c_void_p) else p
    if not p:
        return None
    n = int(np.prod(shape))
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctype)), shape=(n,)).reshape(shape)


class OracleBackedLib:
    """same call surface as rgbd_gan_b200._lib, every entry point executed by the CPU oracle"""

    def __init__(self, oracle):
        self.oracle, self.calls = oracle, []

    def load(self):
        return types.SimpleNamespace(rgbd_consistency_workspace_bytes=lambda B, C, H, W: 1024,
                                     rgbd_dv_render_workspace_bytes=lambda P, B, F: 1024,
                                     rgbd_dv_workspace_bytes=lambda P: 1024,
                                     rgbd_dv_project_workspace_bytes=lambda P, B, F: 1024)

    def _dvp(self, P):
        K = np.array([[P.fx, 0, P.cx, 0], [0, P.fy, P.cy, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
        return self.oracle.dv_params(P.W, P.H, P.D, P.G, K, P.voxel_size, P.near_plane)

    def _common(self, a):
        img, img_rot, M, c, Mi, ci, B, C, H, W, opts = a[:11]
        o = opts._obj
        sh = (B, C, H, W)
        kw = dict(norm=o.norm, occlusion=bool(o.occlusion_aware),
                  max_depth=None if np.isnan(o.max_depth) else o.max_depth,
                  min_depth=None if np.isnan(o.min_depth) else o.min_depth,
                  n_pairs_global=o.n_pairs_global or B)
        arrs = (_view(img, sh), _view(img_rot, sh), _view(M, (B, 9)), _view(c, (B, 3)), _view(Mi, (B, 9)), _view(ci, (B, 3)))
        return arrs, (B, C, H, W), o, kw

    @staticmethod
    def _hinge(o):
        return (o.hinge_depth_min == o.hinge_depth_min) and o.hinge_lambda > 0

    def _fwd(self, arrs, dims, o, kw, parts, new_zp):
        B, C, H, W = dims
        p, d = self.oracle.consistency_fwd(*arrs, debug=True, **kw)
        out = _view(parts, (8,))
        out[:4] = p
        out[4] = self.oracle.combine_loss(p, o.lambda_geometric)
        out[5:] = 0
        if self._hinge(o):
            out[5] = self.oracle.depth_hinge(arrs[0], arrs[1], o.hinge_depth_min, o.hinge_lambda,
                                             n_pairs_global=kw["n_pairs_global"])
        out[6] = out[4] + out[5]
        z = _view(new_zp, (2 * B, H * W, 3))
        if z is not None:
            z[...] = d["new_zp"]

    def call(self, name, *a):
        self.calls.append(name)
        if name == "rgbd_consistency_fwd":
            arrs, dims, o, kw = self._common(a)
            assert len(a) == 17
            self._fwd(arrs, dims, o, kw, a[11], a[12])
        elif name == "rgbd_consistency_fwd_bwd":
            arrs, dims, o, kw = self._common(a)
            assert len(a) == 19
            gy = a[11].value
            self._fwd(arrs, dims, o, kw, a[12], a[13])
            gi, gr = self.oracle.consistency_bwd(*arrs, lambda_geometric=o.lambda_geometric, gy=gy, **kw)
            if self._hinge(o):
                self.oracle.depth_hinge(arrs[0], arrs[1], o.hinge_depth_min, o.hinge_lambda,
                                        n_pairs_global=kw["n_pairs_global"], gy=gy, g_img=gi, g_img_rot=gr)
            _view(a[14], dims)[...] = gi
            _view(a[15], dims)[...] = gr
        elif name == "rgbd_consistency_bwd":
            arrs, dims, o, kw = self._common(a)
            assert len(a) == 19
            B, C, H, W = dims
            gy = a[11].value * (float(_view(a[12], (1,))[0]) if a[12].value else 1.0)
            gz = _view(a[13], (2 * B, H * W, 3))
            gi, gr = self.oracle.consistency_bwd(*arrs, lambda_geometric=o.lambda_geometric, gy=gy, g_new_zp=gz, **kw)
            if self._hinge(o):
                self.oracle.depth_hinge(arrs[0], arrs[1], o.hinge_depth_min, o.hinge_lambda,
                                        n_pairs_global=kw["n_pairs_global"], gy=gy, g_img=gi, g_img_rot=gr)
            _view(a[14], dims)[...] = gi
            _view(a[15], dims)[...] = gr
        elif name == "rgbd_consistency_rescale":
            g0, g1, n, gy_dev, expected, _ = a
            r = float(_view(gy_dev, (1,))[0]) / expected.value
            if r != 1.0:
                _view(g0, (n,))[...] *= r
                _view(g1, (n,))[...] *= r
        elif name in ("rgbd_dv_render_fwd", "rgbd_dv_render_bwd"):
            P, R = a[0]._obj, a[1]._obj
            B, F = a[8], a[9]
            G, HW = P.G, P.H * P.W
            K = np.array([[P.fx, 0, P.cx, 0], [0, P.fy, P.cy, 0], [0, 0, 1, 0], [0, 0, 0, 1.0]])
            P0 = self.oracle.dv_params(P.W, P.H, P.D, G, K, P.voxel_size, P.near_plane)
            grid, cam = _view(a[2], (B, F, G, G, G)), _view(a[3], (B, 4, 4))
            W1, b1, W2, b2 = _view(a[4], (R.nf, F + 1)), _view(a[5], (R.nf,)), _view(a[6], (1, R.nf)), _view(a[7], (1,))
            args = (P0, grid, cam, W1, b1, W2, b2, R.threshold, R.inv_c1, R.inv_c2, R.depth_steps)
            if name == "rgbd_dv_render_fwd":
                assert len(a) == 17 and a[13].value
                novel, depth, fg = self.oracle.dv_render_fwd(*args)
                _view(a[10], (B, F, HW))[...] = novel.reshape(B, F, HW)
                _view(a[11], (B, HW))[...] = depth.reshape(B, HW)
                _view(a[12], (B, HW))[...] = fg.reshape(B, HW)
            else:
                assert len(a) == 22 and a[10].value
                g_fg = _view(a[13], (B, P.H, P.W))
                outs = self.oracle.dv_render_bwd(*args, _view(a[11], (B, F, P.H, P.W)), _view(a[12], (B, P.H, P.W)), g_fg)
                for ptr, o in zip(a[14:19], outs):
                    _view(ptr, o.shape)[...] = o
        elif name == "rgbd_warp_fwd":
            z, M, cv, B, H, W, out, _ = a
            _view(out, (B, H * W, 3))[...] = self.oracle.warp_fwd(_view(z, (B, 1, H * W)), _view(M, (B, 9)), _view(cv, (B, 3)), H, W)
        elif name == "rgbd_warp_bwd":
            g, M, B, H, W, gz, _ = a
            _view(gz, (B, 1, H * W))[...] = self.oracle.warp_bwd(_view(g, (B, H * W, 3)), _view(M, (B, 9)), H, W)
        elif name == "rgbd_bilinear_fwd":
            img, zp, B, C, H, W, warped, mask, _ = a
            w, m = self.oracle.bilinear_fwd(_view(img, (B, C, H, W)), _view(zp, (B, H * W, 3)))
            _view(warped, (B * H * W, C))[...] = w
            _view(mask, (B * H * W,), ctypes.c_uint8)[...] = m
        elif name == "rgbd_bilinear_bwd":
            img, zp, g, B, C, H, W, g_img, g_zp, _ = a
            gi, gz = self.oracle.bilinear_bwd(_view(img, (B, C, H, W)), _view(zp, (B, H * W, 3)), _view(g, (B * H * W, C)))
            _view(g_img, (B, C, H, W))[...] = gi
            _view(g_zp, (B, H * W, 3))[...] = gz
        elif name == "rgbd_dv_compute_proj_idcs":
            P, cam, lin, vc, M, ws, wsb, _ = a
            P = P._obj
            n = P.W * P.H * P.D
            r = self.oracle.dv_compute_proj_idcs(self._dvp(P), _view(cam, (4, 4)))
            M._obj.value = 0 if r is None else r[0].size
            if r is not None:
                _view(lin, (n,), ctypes.c_int32)[:r[0].size] = r[0]
                _view(vc, (3, n))[:, :r[0].size] = r[1]
        elif name in ("rgbd_dv_trilinear_fwd", "rgbd_dv_trilinear_bwd"):
            src, lin, vc, ld, M, F, P, dst, _ = a
            P = P._obj
            n, G = P.W * P.H * P.D, P.G
            li = _view(lin, (M,), ctypes.c_int32)
            v = _view(vc, (3, ld))[:, :M]
            if name.endswith("fwd"):
                _view(dst, (F, n))[...] = self.oracle.dv_trilinear_fwd(_view(src, (F, G, G, G)), li, v, P).reshape(F, n)
            else:
                _view(dst, (F, G, G, G))[...] = self.oracle.dv_trilinear_bwd(_view(src, (F, n)), li, v, P)
        elif name in ("rgbd_dv_project_fwd", "rgbd_dv_project_bwd"):
            P, src, cam, B, F, dst, ws, wsb, _ = a
            P0 = self._dvp(P._obj)
            G, sh = P0.G, (B, F, P0.D, P0.H, P0.W)
            if name.endswith("fwd"):
                _view(dst, sh)[...] = self.oracle.dv_project_fwd(P0, _view(src, (B, F, G, G, G)), _view(cam, (B, 4, 4)))
            else:
                _view(dst, (B, F, G, G, G))[...] = self.oracle.dv_project_bwd(P0, _view(src, sh), _view(cam, (B, 4, 4)))
        elif name == "rgbd_dv_compute_proj_idcs_g2w":
            P, cam, w2g, lin, vc, M, ws, wsb, _ = a
            P = P._obj
            n = P.W * P.H * P.D
            g2w = np.linalg.inv(_view(w2g, (4, 4)).astype(np.float64)).astype(np.float32)
            r = self.oracle.dv_compute_proj_idcs(self._dvp(P), _view(cam, (4, 4)), np.linalg.inv(np.linalg.inv(g2w)))
            M._obj.value = 0 if r is None else r[0].size
            if r is not None:
                _view(lin, (n,), ctypes.c_int32)[:r[0].size] = r[0]
                _view(vc, (3, n))[:, :r[0].size] = r[1]
        elif name == "rgbd_pose_algebra":
            from rgbd_gan_b200.host_math import pose_algebra
            th, thr, B, K, iK, M, c, Mi, ci, _ = a
            Kn, iKn = np.array(list(K), np.float32).reshape(3, 3), np.array(list(iK), np.float32).reshape(3, 3)
            outs = pose_algebra(Kn, iKn, _view(th, (B, 4, 4)), _view(thr, (B, 4, 4)))
            for ptr, o in zip((M, c, Mi, ci), outs):
                _view(ptr, o.shape)[...] = o
        elif name == "rgbd_pose_camera_matrices":
            from oracle import numpy_port as npp_
            th, cs, n, order, cam, _ = a
            _view(cam, (n, 4, 4))[...] = npp_.get_camera_matries(_view(th, (n, 6)), tuple(order))
        elif name == "rgbd_pose_sample":
            prior, B, draws, seed, step, out, _ = a
            pr = prior._obj
            d = _view(draws, (B, 15), ctypes.c_double)
            rng_, uni = np.array(list(pr.camera_param_range)), bool(pr.uniform_distribution)
            u, e, sign = d[:, :6].copy(), d[:, 6:12].copy(), d[:, 12:] * 2 - 1
            lim = np.clip(1 / (rng_[:3] + 1e-8), 0, 1)
            e[:, :3] = e[:, :3] * (sign if uni else (sign * (rng_[:3] == 3.1415) + np.abs(sign) * (rng_[:3] != 3.1415))) * lim
            t2 = -e * np.sign(u) + u
            if uni:
                t2 = t2 * (-1 <= t2) * (t2 <= 1) + (-2 - t2) * (t2 < -1) + (2 - t2) * (t2 > 1)
            _view(out, (2 * B, 6))[...] = (np.concatenate([u, t2], axis=0) * rng_[None]).astype("float32")
        elif name == "rgbd_depth_head_fwd":
            from oracle import numpy_port as npp_
            h, B, C, H, W, out, _ = a
            _view(out, (B, C, H, W))[...] = npp_.depth_head_fwd(_view(h, (B, C, H, W)))
        elif name == "rgbd_depth_head_bwd":
            from oracle import numpy_port as npp_
            h, g, B, C, H, W, gh, _ = a
            _view(gh, (B, C, H, W))[...] = npp_.depth_head_bwd(_view(h, (B, C, H, W)), _view(g, (B, C, H, W)))
        else:
            raise AssertionError("unexpected C-ABI call " + name)


@pytest.fixture()
def nodes():
    chainer_shim.install()
    for m in [k for k in sys.modules if k.startswith("rgbd_gan_b200.chainer_nodes")]:
        del sys.modules[m]
    import rgbd_gan_b200.chainer_nodes as cn       # picks up the shim as `chainer`
    assert cn.FunctionNode is chainer_shim.FunctionNode and not cn.HAVE_CHAINER
    return cn


@pytest.mark.parametrize("name", ["loss_cfg0_l1_occ", "loss_s32_l2_feat", "loss_dv_maxdepth", "loss_edge_wild"])
@pytest.mark.parametrize("grad_scale", [None, 2.0, 0.5])
def test_chainer_node_reproduces_reference(nodes, oracle_mod, name, grad_scale):
    g = load_golden(name)
    o = case_options(g)
    B = o["B"]
    lib = OracleBackedLib(oracle_mod)
    f = nodes.LossFuncRotate(FakeXP(), K=None if o["K"] is None else o["K"].copy(), norm=o["norm"],
                             lambda_geometric=o["lam"], grad_scale=grad_scale, lib=lib)
    V = chainer_shim.Variable
    img, img_rot = V(_wrap(g["x"][:B])), V(_wrap(g["x"][B:]))
    kw = dict(occlusion_aware=o["occ"])
    if o["max_depth"] is not None:
        kw["max_depth"] = o["max_depth"]
    if o["min_depth"] is not None:
        kw["min_depth"] = o["min_depth"]
    loss, zp = f(img, g["cam"][:B], img_rot, g["cam"][B:], **kw)
    assert loss.shape == () and zp.shape == g["new_zp_cat"].shape
    (loss * o["gy"]).backward()                      # updater.py:363-365,387
    assert abs(float(loss.array) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    np.testing.assert_array_equal(np.asarray(zp.array), g["new_zp_cat"])
    assert_grad_close(np.asarray(img.grad), g["g_img"])
    assert_grad_close(np.asarray(img_rot.grad), g["g_img_rot"])
    np.testing.assert_array_equal(f.K, g["K"])
    expect = ["rgbd_consistency_fwd", "rgbd_consistency_bwd"] if grad_scale is None else \
             ["rgbd_consistency_fwd_bwd", "rgbd_consistency_rescale"]
    assert lib.calls == expect


@pytest.mark.parametrize("grad_scale", [None, 2.0])
def test_chainer_node_with_fused_depth_hinge(nodes, oracle_mod, grad_scale):
    g = load_golden("hinge_ffhq")
    B = int(g["B"])
    lib = OracleBackedLib(oracle_mod)
    f = nodes.LossFuncRotate(FakeXP(), lambda_geometric=3, grad_scale=grad_scale, lib=lib)
    V = chainer_shim.Variable
    img, img_rot = V(_wrap(g["x"][:B])), V(_wrap(g["x"][B:]))
    loss, _ = f(img, g["cam"][:B], img_rot, g["cam"][B:], True,
                depth_hinge=(float(g["depth_min"]), float(g["lambda_depth"])))
    (loss * float(g["gy"])).backward()
    assert abs(float(loss.array) - float(g["total"])) <= 1e-5 * abs(float(g["total"]))
    assert_grad_close(np.asarray(img.grad), g["g_img"])
    assert_grad_close(np.asarray(img_rot.grad), g["g_img_rot"])


@pytest.mark.parametrize("name", ["render_g12_thr3"])
def test_chainer_render_node(nodes, oracle_mod, name):
    """next row: RenderAccumulative.apply((deepvoxels, W1, b1, W2, b2)) + backward through the shim"""
    from rgbd_gan_b200._lib import DvParams
    g = load_golden(name)
    G, img, D = int(g["G"]), int(g["img"]), int(g["D"])
    P = DvParams(img, img, D, G, 2. * img, 2. * img, img / 2., img / 2., float(np.float32(g["voxel_size"])),
                 float(np.float32(g["near_plane"])))
    lib = OracleBackedLib(oracle_mod)
    node = nodes.RenderAccumulative(g["cam"], P, float(g["threshold"]), xp=FakeXP(), lib=lib)
    V = chainer_shim.Variable
    ins = [V(_wrap(g[k])) for k in ("grid", "W1", "b1", "W2", "b2")]
    novel, depth, fg = node.apply(tuple(ins))
    F_ = chainer_shim.functions
    loss = F_.sum(novel * g["g_novel"]) + F_.sum(depth * g["g_depth"]) + F_.sum(fg * g["g_fg"])
    loss.backward()
    rel = lambda a, b: float(np.abs(np.asarray(a) - b).max() / np.abs(b).max())
    assert rel(novel.array, g["novel"]) <= 1e-5 and rel(depth.array, g["depth"]) <= 1e-5
    for v, key in zip(ins, ("g_grid", "g_W1", "g_b1", "g_W2", "g_b2")):
        assert rel(np.asarray(v.grad).reshape(g[key].shape), g[key]) <= 1e-5, key
    assert lib.calls == ["rgbd_dv_render_fwd", "rgbd_dv_render_bwd"]


# ------------------------------------------------------------------ the rest of the reference's call surface
def test_chainer_module_imports_without_torch():
    """the Chainer + CuPy deployment has no torch: importing the nodes must not pull it in"""
    import subprocess
    code = ("import sys, importlib.abc\n"
            "class Block(importlib.abc.MetaPathFinder):\n"
            "    def find_spec(self, name, path, target=None):\n"
            "        if name == 'torch' or name.startswith('torch.'): raise ImportError('torch is blocked')\n"
            "sys.meta_path.insert(0, Block())\n"
            "import rgbd_gan_b200.chainer_nodes as cn\n"
            "assert 'torch' not in sys.modules\n"
            "print(sorted(cn.__all__))\n")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0, out.stderr
    assert "LossFuncRotate" in out.stdout and "interpolate_trilinear" in out.stdout


def test_chainer_debug_tuple_and_free_functions(nodes, oracle_mod):
    """debug=True (:100-102) and the free functions warp / inv_warp / bilinear (:171-228) against the golden intermediates,
    gradients of bilinear through the shim's reverse walk"""
    g = load_golden("loss_cfg0_l1_occ")
    o = case_options(g)
    B, S = o["B"], o["S"]
    N = B * S * S
    lib = OracleBackedLib(oracle_mod)
    xp = FakeXP()
    f = nodes.LossFuncRotate(xp, lambda_geometric=o["lam"], lib=lib)
    V = chainer_shim.Variable
    img, img_rot = V(_wrap(g["x"][:B])), V(_wrap(g["x"][B:]))
    warped, not_out, new_zp, warped_rot, not_out_rot, new_zp_rot = f(img, g["cam"][:B], img_rot, g["cam"][B:], debug=True)
    np.testing.assert_array_equal(np.asarray(new_zp.array), g["new_zp_cat"][:B])
    np.testing.assert_array_equal(np.asarray(new_zp_rot.array), g["new_zp_cat"][B:])
    np.testing.assert_array_equal(np.asarray(not_out), g["not_out"])
    np.testing.assert_array_equal(np.asarray(not_out_rot), g["not_out_rot"])
    np.testing.assert_array_equal(np.asarray(warped.array), g["warped"])
    np.testing.assert_array_equal(np.asarray(warped_rot.array), g["warped_rot"])
    assert lib.calls == ["rgbd_warp_fwd", "rgbd_warp_fwd", "rgbd_bilinear_fwd", "rgbd_bilinear_fwd"]
    # free functions with the reference's argument lists
    from oracle import numpy_port as npp
    port = npp.LossFuncRotateNP(lambda_geometric=o["lam"])
    port.init_params(S)
    th, thr = g["cam"][:B], g["cam"][B:]
    R = np.matmul(thr[:, :3, :3].transpose(0, 2, 1), th[:, :3, :3]).astype("float32")
    t = np.matmul(th[:, :3, :3].transpose(0, 2, 1), thr[:, :3, -1:] - th[:, :3, -1:]).astype("float32")
    z = V(_wrap(g["x"][:B, -1:].reshape(B, 1, -1)))
    zp = nodes.warp(port.K, port.inv_K, R, t, z, port.p, xp=xp, lib=lib)
    np.testing.assert_array_equal(np.asarray(zp.array), g["new_zp_cat"][:B])
    z_rot = V(_wrap(g["x"][B:, -1:].reshape(B, 1, -1)))
    zpr = nodes.inv_warp(port.K, port.inv_K, R.transpose(0, 2, 1), t, z_rot, port.p, xp=xp, lib=lib)
    np.testing.assert_array_equal(np.asarray(zpr.array), g["new_zp_cat"][B:])
    w, m = nodes.bilinear(img_rot, zp, xp=xp, lib=lib)
    np.testing.assert_array_equal(np.asarray(m), g["not_out"])
    # gradient flows into the sampled image and, through new_zp, into the depth of the source image
    gw = np.random.default_rng(0).normal(size=w.shape).astype(np.float32)
    F_ = chainer_shim.functions
    F_.sum(w * gw).backward()
    ref_gi, ref_gzp = oracle_mod.bilinear_bwd(g["x"][B:], g["new_zp_cat"][:B], gw)
    assert_grad_close(np.asarray(img_rot.grad), ref_gi)
    M, _ = __import__("rgbd_gan_b200.host_math", fromlist=["x"]).warp_constants(port.K, port.inv_K, R, t, False)
    assert_grad_close(np.asarray(z.grad).reshape(B, 1, -1), oracle_mod.warp_bwd(ref_gzp, M, S, S))


@pytest.mark.parametrize("name", ["dv_g16_f3"])
def test_chainer_projection_surface(nodes, oracle_mod, name):
    """ProjectionHelper.compute_proj_idcs (projection.py:48-105), interpolate_trilinear (deepvoxel.py:388-428) and the
    fused batch projection, values and gradients against the reference's golden vectors"""
    g = load_golden(name)
    G, img, F, D = int(g["G"]), int(g["img"]), int(g["F"]), int(g["D"])
    lib = OracleBackedLib(oracle_mod)
    xp = FakeXP()
    h = nodes.ProjectionHelper(g["intrinsic"], g["intrinsic"], [img, img], [img, img], 0., 1., [G] * 3,
                               float(g["voxel_size"]), float(g["near_plane"]), D, verbose=False, xp=xp, lib=lib)
    V = chainer_shim.Variable
    F_ = chainer_shim.functions
    for i in range(g["cam"].shape[0]):
        lin, vc = h.compute_proj_idcs(g["cam"][i])
        np.testing.assert_array_equal(np.asarray(lin), g["lin_ind_%d" % i])
        np.testing.assert_array_equal(np.asarray(vc), g["voxel_coords_%d" % i])
        grid = V(_wrap(g["grid"][i:i + 1]))
        out = nodes.interpolate_trilinear(grid, lin, vc, [img, img], D, xp=xp, lib=lib)
        assert out.shape == (1, F, D, img, img)
        np.testing.assert_array_equal(np.asarray(out.array), g["frustum_%d" % i])
        F_.sum(out * g["g_out"][i:i + 1]).backward()
        assert_grad_close(np.asarray(grid.grad), g["g_grid_%d" % i])
    far = g["cam"][0].copy()
    far[:3, 3] += 100.0
    assert h.compute_proj_idcs(far) is None                       # "error: nothing in frustum bounds" (:98-100)
    grid = V(_wrap(g["grid"]))
    fr = h.project(grid, g["cam"])
    F_.sum(fr * g["g_out"]).backward()
    for i in range(g["cam"].shape[0]):
        np.testing.assert_array_equal(np.asarray(fr.array)[i], g["frustum_%d" % i][0])
        assert_grad_close(np.asarray(grid.grad)[i], g["g_grid_%d" % i][0])
    assert lib.calls.count("rgbd_dv_project_fwd") == 1 and lib.calls.count("rgbd_dv_project_bwd") == 1


def test_chainer_node_returns_only_requested_gradients(nodes, oracle_mod):
    """FunctionNode.backward contract: one gradient per target_input_index (only img_rot requires grad here)"""
    g = load_golden("loss_cfg0_l1_occ")
    o = case_options(g)
    B = o["B"]
    lib = OracleBackedLib(oracle_mod)
    f = nodes.LossFuncRotate(FakeXP(), lambda_geometric=o["lam"], lib=lib)
    V = chainer_shim.Variable
    img, img_rot = V(_wrap(g["x"][:B]), requires_grad=False), V(_wrap(g["x"][B:]))
    loss, _ = f(img, g["cam"][:B], img_rot, g["cam"][B:], occlusion_aware=o["occ"])
    (loss * o["gy"]).backward()
    assert img.grad is None
    assert_grad_close(np.asarray(img_rot.grad), g["g_img_rot"])


def test_chainer_depth_head_node(nodes, oracle_mod):
    """next row: the generators' depth head (net.py:294-299) as a node, against the reference expression's golden"""
    g = load_golden("depth_head_s32")
    lib = OracleBackedLib(oracle_mod)
    V = chainer_shim.Variable
    h = V(_wrap(g["h"]))
    out = nodes.depth_head(h, xp=FakeXP(), lib=lib)
    chainer_shim.functions.sum(out * g["g_out"]).backward()
    np.testing.assert_array_equal(np.asarray(out.array), g["out"])
    np.testing.assert_allclose(np.asarray(h.grad), g["g_h"], rtol=1e-6, atol=1e-12)
    assert lib.calls == ["rgbd_depth_head_fwd", "rgbd_depth_head_bwd"]


class DevArr:
    """a device array as the glue sees it when it is NOT a numpy array: `.data.ptr`, shape, dtype, slicing, `get()`"""

    def __init__(self, a):
        self.a = np.ascontiguousarray(a)

    data = property(lambda s: _Ptr(s.a))
    shape = property(lambda s: s.a.shape)
    dtype = property(lambda s: s.a.dtype)
    ndim = property(lambda s: s.a.ndim)

    def __array__(self, dtype=None, copy=None):
        return self.a if dtype is None else self.a.astype(dtype)

    def __getitem__(self, k):
        return DevArr(self.a[k])

    def get(self):
        return self.a


def test_device_thetas_take_the_pose_kernel(nodes, oracle_mod):
    """SURVEY 8f rank 4 on the Chainer surface: thetas that are DEVICE arrays go through rgbd_pose_algebra (no .get()); host
    thetas keep the NumPy path; get_camera_matries / CameraParamPrior marshal their arguments as the header says"""
    g = load_golden("loss_cfg0_l1_occ")
    o = case_options(g)
    B = o["B"]
    lib = OracleBackedLib(oracle_mod)
    f = nodes.LossFuncRotate(FakeXP(), lambda_geometric=o["lam"], lib=lib)
    img, img_rot = chainer_shim.Variable(_wrap(g["x"][:B])), chainer_shim.Variable(_wrap(g["x"][B:]))
    cam = DevArr(g["cam"])
    loss, zp = f(img, cam[:B], img_rot, cam[B:], occlusion_aware=o["occ"])
    assert lib.calls[0] == "rgbd_pose_algebra" and "rgbd_consistency_fwd" in lib.calls
    np.testing.assert_array_equal(np.asarray(zp.array), g["new_zp_cat"])
    assert abs(float(loss.array) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    lib.calls.clear()
    f(img, g["cam"][:B], img_rot, g["cam"][B:], occlusion_aware=o["occ"])              # host thetas: no pose kernel
    assert "rgbd_pose_algebra" not in lib.calls
    got = nodes.get_camera_matries(DevArr(g["thetas"]), xp=FakeXP(), lib=lib)
    np.testing.assert_array_equal(np.asarray(got), g["cam"])
    cfg = types.SimpleNamespace(x_rotate=0.3054, y_rotate=1.0472, z_rotate=0, x_translate=0, y_translate=0, z_translate=0,
                                uniform_distribution=False)
    from oracle import numpy_port as npp
    np.random.seed(9)
    u, e, s = np.random.uniform(-1, 1, (B, 6)), np.random.uniform(0, 0.5, (B, 6)), np.random.choice(2, (B, 3))
    np.random.seed(9)
    want = npp.sample_camera_prior(2 * B, npp.FFHQ_RANGES, False)
    pr = nodes.CameraParamPrior(cfg, xp=FakeXP(), lib=lib)
    got = pr.sample(2 * B, draws=np.concatenate([u, e, s.astype(np.float64)], 1))
    np.testing.assert_array_equal(np.asarray(got), want)
    assert pr.step == 1 and lib.calls[-1] == "rgbd_pose_sample"
