#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

How the reference is executed here
  * `common/loss_functions.py`, `deepvoxel/projection.py`, `deepvoxel/deepvoxel.py`
    are imported from /root/reference as they are.  Their `chainer` dependency is
    served by tests/golden/chainer_shim.py (a restatement of the Chainer-v7 ops
    those files call; Chainer/CuPy are not installable here).
  * the pure-NumPy helpers `get_camera_matries` / `update_camera_matrices`
    (updater.py:26-60) and `CameraParamPrior` (train_rgbd.py:192-217) are lifted
    out of their files with `ast` and executed as they are (their modules import
    cupy/PIL/chainer.training at top level, which is irrelevant to the helpers).
  * NumPy-2 accommodations that do not change reference-era results:
      - `xp.meshgrid` returns a list (loss_functions.py:59-61 concatenates it);
      - `near_plane` is handed to ProjectionHelper as np.float32: under the
        NumPy-1.x / CuPy value-based casting the reference ran on,
        `coords[2] += near_plane` (projection.py:74) is an fp32 add; NumPy >= 2
        (NEP 50) would do it in fp64.

Outputs: one .npz per case (inputs + every output and intermediate needed by the
parity tests).  Sizes are kept small; cfg0 (B=4 at 128^2) is stored at B=2.
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RGBDGAN_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)

import chainer_shim  # noqa: E402

chainer = chainer_shim.install()
xp = chainer_shim.xp_compat
Variable = chainer_shim.Variable


def _lift(path, names, extra=None):
    """exec the named top-level defs of a reference file, unmodified."""
    src = open(path).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(keep) == len(names), (path, names)
    ns = {"np": np}
    ns.update(extra or {})
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return ns


def load_reference():
    sys.path.insert(0, REF)
    # network layers imported by deepvoxel/deepvoxel.py at module level: out of scope
    for name in ("common.networks", "common.networks.component", "common.networks.component.pggan"):
        sys.modules[name] = chainer_shim._Stub(name)
    import common.loss_functions as lf
    import deepvoxel.projection as pj
    import deepvoxel.deepvoxel as dv
    cam = _lift(os.path.join(REF, "updater.py"), ["update_camera_matrices", "get_camera_matries"])
    prior = _lift(os.path.join(REF, "train_rgbd.py"), ["CameraParamPrior"])
    return lf, pj, dv, cam["get_camera_matries"], prior["CameraParamPrior"]


class Cfg:
    def __init__(self, x_rotate, y_rotate, z_rotate=0, x_t=0, y_t=0, z_t=0, uniform=None):
        self.x_rotate, self.y_rotate, self.z_rotate = x_rotate, y_rotate, z_rotate
        self.x_translate, self.y_translate, self.z_translate = x_t, y_t, z_t
        self.uniform_distribution = uniform


FFHQ = dict(x_rotate=0.3054, y_rotate=1.0472)           # configs/ffhq_stylegan_occlusion.yml:37-43
CAR = dict(x_rotate=0.3054, y_rotate=3.1415)            # configs/dcgan_shapenet_car.yml:38-44
DV_CAR = dict(x_rotate=0.3054, y_rotate=3.1415, uniform=True)  # deepvoxels_shapenet_car.yml:38-55


def make_images(B, C, S, depth):
    """(2B,C,S,S) fp32: RGB/features U(-1,1); last channel = depth."""
    x = np.random.uniform(-1, 1, size=(2 * B, C, S, S)).astype("float32")
    if depth == "rough":
        x[:, -1] = np.random.uniform(0.7, 1.5, size=(2 * B, S, S))
    elif depth == "smooth":
        col = np.arange(S, dtype="float32")[None, None, :]
        row = np.arange(S, dtype="float32")[None, :, None]
        ph = np.random.uniform(0, 6.28, size=(2 * B, 1, 1)).astype("float32")
        x[:, -1] = 1 + 0.1 * np.sin(col / 20 * (128 / S) + ph) + 0.05 * np.cos(row / 13 * (128 / S))
    elif depth == "wide":          # exercises max_depth / min_depth = 3
        x[:, -1] = np.random.uniform(0.7, 5.0, size=(2 * B, S, S))
    elif depth == "wild":          # non-positive / tiny depths: clip + `zp2 > 1e-4` paths
        x[:, -1] = np.random.uniform(-0.5, 2.0, size=(2 * B, S, S))
    else:
        raise ValueError(depth)
    return x


def run_loss(lf, loss_obj, x, cam, B, gy, **kw):
    img, img_rot = Variable(x[:B].copy()), Variable(x[B:].copy())
    loss, zp = loss_obj(img, cam[:B], img_rot, cam[B:], **kw)
    (loss * gy).backward()
    dbg = loss_obj(Variable(x[:B].copy()), cam[:B], Variable(x[B:].copy()), cam[B:], debug=True, **kw)
    warped, not_out, new_zp, warped_rot, not_out_rot, new_zp_rot = dbg
    return dict(
        loss=loss.array, new_zp_cat=np.ascontiguousarray(zp.array),
        g_img=img.grad, g_img_rot=img_rot.grad,
        warped=warped.array, not_out=not_out, warped_rot=warped_rot.array, not_out_rot=not_out_rot,
        K=np.array(loss_obj.K), inv_K=np.array(loss_obj.inv_K), p=np.array(loss_obj.p))


def consistency_case(lf, get_cam, Prior, name, seed, B, C, S, depth, pose, norm="l1", lam=3,
                     occ=False, K=None, max_depth=None, min_depth=None, gy=2.0, scale_pose=1.0,
                     translate=None):
    np.random.seed(seed)
    cfg = Cfg(**pose)
    if translate:
        cfg.x_translate, cfg.y_translate, cfg.z_translate = translate
    thetas = Prior(cfg).sample(2 * B) * np.float32(scale_pose)
    cam = get_cam(thetas)
    x = make_images(B, C, S, depth)
    obj = lf.LossFuncRotate(xp, K=None if K is None else K.copy(), norm=norm, lambda_geometric=lam)
    kw = dict(occlusion_aware=occ)
    if max_depth is not None:
        kw["max_depth"] = max_depth
    if min_depth is not None:
        kw["min_depth"] = min_depth
    out = run_loss(lf, obj, x, cam, B, gy, **kw)
    meta = dict(B=B, C=C, S=S, norm=norm, lambda_geometric=lam, occlusion_aware=occ, gy=gy,
                max_depth=-1.0 if max_depth is None else max_depth,
                min_depth=-1.0 if min_depth is None else min_depth, has_K=K is not None)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, thetas=thetas, cam=cam,
                        K_in=np.zeros((0,)) if K is None else K, **meta, **out)
    print("%-22s loss=%.9g  in=%.3f/%.3f" % (name, float(out["loss"]), out["not_out"].mean(),
                                             out["not_out_rot"].mean()))


def hinge_case(lf, get_cam, Prior, name, seed, B, S, depth, pose, depth_min, lambda_depth, gy=2.0):
    """the generator-step expression around the loss, as updater.py:340-365 evaluates it:
    loss_rotate = LossFuncRotate(...) ; loss_rotate += mean(relu(depth_min - x_fake[:, -1]) ** 2) * lambda_depth ;
    loss_gen += loss_rotate * lambda_rotate ; loss_gen.backward()"""
    F = chainer.functions
    np.random.seed(seed)
    thetas = Prior(Cfg(**pose)).sample(2 * B)
    cam = get_cam(thetas)
    x = make_images(B, 4, S, depth)
    obj = lf.LossFuncRotate(xp, lambda_geometric=3)
    img, img_rot = Variable(x[:B].copy()), Variable(x[B:].copy())
    loss_rotate, _ = obj(img, cam[:B], img_rot, cam[B:], True)
    x_fake = F.concat([img, img_rot], axis=0)
    hinge = F.mean(F.relu(depth_min - x_fake[:, -1]) ** 2) * lambda_depth            # updater.py:357-359
    total = loss_rotate + hinge
    (total * gy).backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, thetas=thetas, cam=cam, B=B, S=S, gy=gy,
                        depth_min=depth_min, lambda_depth=lambda_depth, loss_rotate=loss_rotate.array,
                        hinge=hinge.array, total=total.array, g_img=img.grad, g_img_rot=img_rot.grad)
    print("%-22s loss_rotate=%.9g hinge=%.9g" % (name, float(loss_rotate.array), float(hinge.array)))


def depth_head_case(name, seed, B, S):
    """next row: the generators' depth head, net.py:294-299 / :756-761 --
        depth = 1 / (F.softplus(h[:, -1:]) + 1e-4);  h = F.concat([h[:, :3], depth])
    the reference's expression evaluated over the shim (the surrounding Generator.forward needs the conv layers)"""
    import chainer
    import chainer.functions as F
    rng = np.random.default_rng(seed)
    h_in = (rng.normal(size=(B, 4, S, S)) * 2.0).astype(np.float32)
    h_in[0, -1, 0, :8] = [-30.0, -12.0, -1e-3, 0.0, 1e-3, 9.0, 17.0, 40.0]         # both tails of softplus
    h = chainer.Variable(h_in.copy())
    depth = 1 / (F.softplus(h[:, -1:]) + 1e-4)
    out = F.concat([h[:, :3], depth])
    g_out = rng.normal(size=out.shape).astype(np.float32)
    F.sum(out * g_out).backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), h=h_in, out=out.array, g_out=g_out, g_h=h.grad)
    print("%-22s depth range [%.4g, %.4g]" % (name, float(out.array[:, -1].min()), float(out.array[:, -1].max())))


def growing_case(lf, get_cam, Prior, name, seed):
    """Q9: one LossFuncRotate instance reused across sizes 32 -> 64 (K mutated in place)."""
    np.random.seed(seed)
    obj = lf.LossFuncRotate(xp, lambda_geometric=3)
    B = 2
    thetas = Prior(Cfg(**FFHQ)).sample(2 * B)
    cam = get_cam(thetas)
    x32 = make_images(B, 4, 32, "smooth")
    o32 = run_loss(lf, obj, x32, cam, B, 1.0, occlusion_aware=True)
    x64 = make_images(B, 4, 64, "smooth")
    o64 = run_loss(lf, obj, x64, cam, B, 1.0, occlusion_aware=True)
    d = dict(thetas=thetas, cam=cam, x32=x32, x64=x64, B=B)
    d.update({k + "_32": v for k, v in o32.items()})
    d.update({k + "_64": v for k, v in o64.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print("%-22s loss32=%.9g loss64=%.9g" % (name, float(o32["loss"]), float(o64["loss"])))


def dv_intrinsic(img):
    return np.array([[img * 2., 0.0, img / 2., 0.0],
                     [0.0, img * 2., img / 2., 0.0],
                     [0.0, 0.0, 1.0, 0.0],
                     [0.0, 0.0, 0.0, 1.0]])       # deepvoxels_generator.py:233-236


def dv_helper(pj, G, img, scale=0.5):
    """ProjectionHelper built exactly as deepvoxels_generator.py:229-253 does (G=32, img=64 there)."""
    near_plane = np.float32(np.sqrt(3) / 4)      # fp32-pinned, see module docstring / Q6
    voxel_size = (1. / G) * 1.1 * scale
    D = int(np.ceil(np.sqrt(3) * G))
    intr = dv_intrinsic(img)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        h = pj.ProjectionHelper(projection_intrinsic=intr, lifting_intrinsic=intr, depth_min=0.,
                                depth_max=G * voxel_size + near_plane,
                                projection_image_dims=[img, img], lifting_image_dims=[img, img],
                                grid_dims=3 * [G], voxel_size=voxel_size, device=None,
                                frustrum_depth=D, near_plane=near_plane)
    return h, D, voxel_size, near_plane


def dv_case(pj, dv, get_cam, Prior, name, seed, G, img, F, nsamp, g2w=False):
    np.random.seed(seed)
    h, D, voxel_size, near_plane = dv_helper(pj, G, img)
    thetas = Prior(Cfg(**DV_CAR)).sample(2 * ((nsamp + 1) // 2))[:nsamp]
    cam = get_cam(thetas)
    grid = np.random.normal(size=(nsamp, F, G, G, G)).astype("float32")
    g_out = np.random.normal(size=(nsamp, F, D, img, img)).astype("float32")
    d = dict(G=G, img=img, F=F, D=D, voxel_size=voxel_size, near_plane=near_plane, thetas=thetas, cam=cam,
             grid=grid, g_out=g_out, intrinsic=dv_intrinsic(img))
    for i in range(nsamp):
        lin_ind, vc = h.compute_proj_idcs(cam[i])
        gv = Variable(grid[i:i + 1].copy())
        out = dv.interpolate_trilinear(gv, lin_ind, vc, [img, img], D)
        out.grad = g_out[i:i + 1]
        out.backward()
        d["lin_ind_%d" % i] = lin_ind
        d["voxel_coords_%d" % i] = vc
        d["frustum_%d" % i] = out.array
        d["g_grid_%d" % i] = gv.grad
        print("%-22s sample %d: M=%d of %d" % (name, i, lin_ind.size, D * img * img))
    if g2w:
        # the optional second argument of compute_proj_idcs (projection.py:48,53-54,83-84; no caller in the reference):
        # a rigid motion + scale of the grid, drawn from its own generator so that the rest of the file is unchanged
        rs = np.random.RandomState(seed + 1000)
        ang = rs.uniform(-0.4, 0.4, size=3)
        Rg = np.eye(3)
        for ax, a in enumerate(ang):
            c_, s_ = np.cos(a), np.sin(a)
            R1 = np.eye(3)
            i0, i1 = (ax + 1) % 3, (ax + 2) % 3
            R1[i0, i0], R1[i0, i1], R1[i1, i0], R1[i1, i1] = c_, -s_, s_, c_
            Rg = R1 @ Rg
        g2w_m = np.eye(4)
        g2w_m[:3, :3] = Rg * 1.07
        g2w_m[:3, 3] = rs.uniform(-0.05, 0.05, size=3)
        g2w_m = g2w_m.astype("float32")
        lin_ind, vc = h.compute_proj_idcs(cam[0], g2w_m)
        d["grid2world"] = g2w_m
        d["lin_ind_g2w_0"] = lin_ind
        d["voxel_coords_g2w_0"] = vc
        print("%-22s grid2world: M=%d" % (name, lin_ind.size))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)


def render_case(pj, dv, get_cam, Prior, name, seed, G, img, F, nsamp, nf=4, threshold=4):
    """Next row: the per-sample render tail of DeepVoxels.forward with the shipped
    `occlusion_type: accumulative` (deepvoxels_shapenet_car.yml:34):
        can_view_vol = interpolate_trilinear(...)                        deepvoxel.py:880-884   (reference code, unmodified)
        weights, depth_map = AccumulativeOcclusionNet.forward(vol)       deepvoxel.py:574-587   (reference code, unmodified)
        collapsed = F.sum(weights * vol, axis=2); fg = F.sum(weights, 2) deepvoxel.py:888,892   (restated here)
        depth = (depth + 0.5) * ceil(sqrt(3) G) * voxel_size + near      deepvoxel.py:903-904   (restated here)
    `self.occlusion` (deepvoxel.py:560-567) is a chainer.Sequential of two Conv3dSame(kernel_size=1) =
    EqualizedConv3d (pggan.py:27-38: conv(inv_c * x), inv_c = sqrt(2) * sqrt(1 / in_ch)) around leaky_relu,
    `x - threshold`, sigmoid; the 1x1x1 convolutions are restated as channel matmuls over the shim."""
    F_ = chainer.functions
    np.random.seed(seed)
    h, D, voxel_size, near_plane = dv_helper(pj, G, img)
    thetas = Prior(Cfg(**DV_CAR)).sample(2 * ((nsamp + 1) // 2))[:nsamp]
    cam = get_cam(thetas)
    grid = np.random.normal(size=(nsamp, F, G, G, G)).astype("float32")
    W1 = Variable(np.random.normal(size=(nf, F + 1)).astype("float32"))             # initialW = Normal(1.0)
    b1 = Variable(np.random.normal(scale=0.3, size=(nf,)).astype("float32"))
    W2 = Variable(np.random.normal(size=(1, nf)).astype("float32"))
    b2 = Variable(np.random.normal(scale=0.3, size=(1,)).astype("float32"))
    inv_c1 = np.sqrt(2) * np.sqrt(1.0 / (F + 1))                                    # pggan.py:31 (ksize = 1)
    inv_c2 = np.sqrt(2) * np.sqrt(1.0 / nf)

    def conv1x1(x, W, b, inv_c):
        bsz, c = x.shape[0], x.shape[1]
        sp = x.shape[2:]
        xs = F_.reshape(inv_c * x, (bsz, c, -1))                                     # pggan.py:38  self.c(self.inv_c * x)
        y = F_.matmul(W, xs)                                                         # 1x1x1 convolution
        y = y + F_.reshape(b, (1, -1, 1))
        return F_.reshape(y, (bsz, W.shape[0]) + tuple(sp))

    def occlusion(x):                                                                # deepvoxel.py:560-567
        hdn = F_.leaky_relu(conv1x1(x, W1, b1, inv_c1))
        o = conv1x1(hdn, W2, b2, inv_c2)
        return F_.sigmoid(o - threshold)

    net = object.__new__(dv.AccumulativeOcclusionNet)                                # forward() is the reference's own
    depth_coords = np.arange(-D // 2, D // 2)[None, None, :, None, None] / D         # deepvoxel.py:568-571
    net.depth_coords = np.tile(depth_coords, (1, 1, 1, img, img)).astype("float32")
    net.occlusion = occlusion
    net.xp = np
    voxels = Variable(grid.copy())
    novel, depths, fgs = [], [], []
    for i in range(nsamp):                                                           # deepvoxel.py:879-892
        lin_ind, vc = h.compute_proj_idcs(cam[i])
        vol = dv.interpolate_trilinear(voxels[None, i], lin_ind, vc, [img, img], D)
        w, depth_map = dv.AccumulativeOcclusionNet.forward(net, vol)
        novel.append(F_.reshape(F_.sum(w * vol, axis=2), (1, -1, img, img)))
        depths.append(depth_map)
        fgs.append(F_.sum(w, axis=2))
    novel = F_.concat(novel, axis=0)
    depth = F_.concat(depths, axis=0)
    fg = F_.concat(fgs, axis=0)
    depth = ((depth + 0.5) * int(np.ceil(np.sqrt(3) * G)) * voxel_size + near_plane)  # deepvoxel.py:903-904
    g_novel = np.random.normal(size=novel.shape).astype("float32")
    g_depth = np.random.normal(size=depth.shape).astype("float32")
    g_fg = np.random.normal(size=fg.shape).astype("float32")
    loss = F_.sum(novel * g_novel) + F_.sum(depth * g_depth) + F_.sum(fg * g_fg)
    loss.backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), G=G, img=img, F=F, D=D, nf=nf, threshold=threshold,
                        voxel_size=voxel_size, near_plane=near_plane, thetas=thetas, cam=cam, grid=grid,
                        intrinsic=dv_intrinsic(img), W1=W1.array, b1=b1.array, W2=W2.array, b2=b2.array,
                        inv_c1=inv_c1, inv_c2=inv_c2, novel=novel.array, depth=depth.array, fg=fg.array,
                        g_novel=g_novel, g_depth=g_depth, g_fg=g_fg, g_grid=voxels.grad, g_W1=W1.grad, g_b1=b1.grad,
                        g_W2=W2.grad, g_b2=b2.grad)
    print("%-22s novel |max| %.4g  depth range [%.4g, %.4g]  fg range [%.3g, %.3g]" % (
        name, np.abs(novel.array).max(), depth.array.min(), depth.array.max(), fg.array.min(), fg.array.max()))


def main():
    lf, pj, dv, get_cam, Prior = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "render":     # only the render-tail cases (leaves the other files untouched)
        render_case(pj, dv, get_cam, Prior, "render_g16", 20, G=16, img=32, F=32, nsamp=2)
        render_case(pj, dv, get_cam, Prior, "render_g12_thr3", 21, G=12, img=24, F=32, nsamp=3, threshold=3)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "depthhead":  # only the depth-head case
        depth_head_case("depth_head_s32", 30, B=3, S=32)
        return
    c = lambda *a, **k: consistency_case(lf, get_cam, Prior, *a, **k)
    # cfg0 / cfg1 shape (BASELINE.json configs[0], [1]) at B=2
    c("loss_cfg0_l1_occ", 0, B=2, C=4, S=128, depth="rough", pose=FFHQ, occ=True, lam=3)
    # cfg2 shape: car poses (full-circle yaw), lambda_geometric 1
    c("loss_car_l1_occ", 1, B=2, C=4, S=64, depth="smooth", pose=CAR, occ=True, lam=1)
    c("loss_s64_l1_noocc", 2, B=3, C=4, S=64, depth="rough", pose=FFHQ, occ=False, lam=3)
    # feature-space variant: norm l2, C != 4 (updater.py:240,345-354)
    c("loss_s32_l2_feat", 3, B=2, C=6, S=32, depth="rough", pose=FFHQ, norm="l2", occ=True, lam=3)
    c("loss_s32_l2_noocc", 4, B=2, C=4, S=32, depth="smooth", pose=CAR, norm="l2", occ=False, lam=3)
    # DeepVoxels updater variant: K = projection intrinsic (4x4), fore/background depth masks
    K = dv_intrinsic(64)
    c("loss_dv_maxdepth", 5, B=2, C=4, S=64, depth="wide", pose=DV_CAR, K=K, max_depth=3, lam=3)
    c("loss_dv_mindepth", 6, B=2, C=4, S=64, depth="wide", pose=DV_CAR, K=K, min_depth=3, lam=3)
    # edge cases: non-positive depths, big rotations + translations (out-of-bounds, behind camera)
    c("loss_edge_wild", 7, B=2, C=4, S=32, depth="wild", pose=CAR, occ=True, lam=3, scale_pose=1.0,
      translate=(0.3, 0.2, 0.4))
    c("loss_edge_c2", 8, B=1, C=2, S=16, depth="rough", pose=FFHQ, occ=True, lam=3)
    growing_case(lf, get_cam, Prior, "loss_growing", 9)
    # next row: the depth hinge the updaters add to the loss, yml values of the configs
    hinge_case(lf, get_cam, Prior, "hinge_ffhq", 12, B=2, S=64, depth="rough", pose=FFHQ, depth_min=1.0, lambda_depth=10)
    hinge_case(lf, get_cam, Prior, "hinge_car", 13, B=3, S=32, depth="smooth", pose=CAR, depth_min=0.6, lambda_depth=10)
    # DeepVoxels projection: scaled-down geometry with several features, and the
    # production geometry (deepvoxels_generator.py:229-253) with one feature
    dv_case(pj, dv, get_cam, Prior, "dv_g16_f3", 10, G=16, img=32, F=3, nsamp=2, g2w=True)
    dv_case(pj, dv, get_cam, Prior, "dv_g32_f1", 11, G=32, img=64, F=1, nsamp=1)
    # next row: DeepVoxels render tail with the accumulative occlusion module
    render_case(pj, dv, get_cam, Prior, "render_g16", 20, G=16, img=32, F=32, nsamp=2)
    render_case(pj, dv, get_cam, Prior, "render_g12_thr3", 21, G=12, img=24, F=32, nsamp=3, threshold=3)
    depth_head_case("depth_head_s32", 30, B=3, S=32)


if __name__ == "__main__":
    main()
