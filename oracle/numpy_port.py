"""Op-by-op NumPy port of the reference's Chainer CPU path (forward AND backward).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This file is the stand-in for "the
reference's NumPy-backed Chainer CPU path" that bench.py times as `cpu_baseline`
(kind "port") and as `--impl reference`: Chainer itself cannot be installed in this
image.  It executes the same sequence of array operations Chainer would run for
common/loss_functions.py:63-146,171-228 and deepvoxel/{projection.py:48-105,
deepvoxel.py:388-428}: one NumPy call per FunctionNode in forward, and the matching
backward rule per node in reverse (GetItem backward = zeros + np.add.at, etc.), with
the dtype casts Chainer applies (`Variable <op> ndarray` casts the ndarray to the
Variable's dtype first).  What it leaves out is Chainer's own
Python overhead per node (graph bookkeeping, type checks), so it is, if anything,
faster than the real thing.

Pinned against tests/golden (tests/test_oracle_golden.py) together with the C restatement.
"""
import numpy as np

f32 = np.float32


# ------------------------------------------------------------------ host-side generators
def update_camera_matrices(mat, axis1, axis2, theta):
    """updater.py:26-42"""
    rot = np.zeros_like(mat)
    rot[:, range(4), range(4)] = 1
    rot[:, axis1, axis1] = np.cos(theta)
    rot[:, axis1, axis2] = -np.sin(theta)
    rot[:, axis2, axis1] = np.sin(theta)
    rot[:, axis2, axis2] = np.cos(theta)
    return np.matmul(rot, mat)


def get_camera_matries(thetas, order=(0, 1, 2)):
    """updater.py:45-60: 6-DoF -> cam2world (B,4,4) fp32"""
    mat = np.zeros((len(thetas), 4, 4), dtype="float32")
    mat[:, range(4), range(4)] = [1, 1, -1, 1]
    mat[:, 2, 3] = 1
    for i in order:
        mat = update_camera_matrices(mat, (i + 1) % 3, (i + 2) % 3, thetas[:, i])
    mat[:, :3, 3] = mat[:, :3, 3] + thetas[:, 3:]
    return mat


def sample_camera_prior(batch_size, ranges, uniform=False):
    """train_rgbd.py:192-217 (CameraParamPrior.sample); uses the global np.random like the reference."""
    ranges = np.asarray(ranges, dtype=np.float64)
    rot = ranges[:3]
    h = batch_size // 2
    thetas = np.random.uniform(-1, 1, size=(h, 6))
    eps = np.random.uniform(0, 0.5, size=(h, 6))
    sign = np.random.choice(2, size=(h, 3)) * 2 - 1
    lim = np.clip(1 / (rot + 1e-8), 0, 1)
    if uniform:
        eps[:, :3] = eps[:, :3] * sign * lim
    else:
        eps[:, :3] = eps[:, :3] * (sign * (rot == 3.1415) + np.abs(sign) * (rot != 3.1415)) * lim
    thetas2 = -eps * np.sign(thetas) + thetas
    if uniform:
        thetas2 = thetas2 * (-1 <= thetas2) * (thetas2 <= 1) + (-2 - thetas2) * (thetas2 < -1) + \
                  (2 - thetas2) * (thetas2 > 1)
    thetas = np.concatenate([thetas, thetas2], axis=0) * ranges[None]
    return thetas.astype("float32")


FFHQ_RANGES = (0.3054, 1.0472, 0, 0, 0, 0)          # configs/ffhq_stylegan_occlusion.yml:37-43
CAR_RANGES = (0.3054, 3.1415, 0, 0, 0, 0)           # configs/dcgan_shapenet_car.yml:38-44


def synthetic_batch(B, S, C=4, depth="rough", ranges=FFHQ_RANGES, uniform=None, seed=0):
    """The synthetic workload of SURVEY.md 8(d) for the parity tests: 2B images (B pairs) + their cam2world matrices,
    poses replayed through the reference's own generators above.  RGB ~ U(-1,1); depth "rough" ~ U(0.7,1.5) or
    "smooth" = 1 + 0.1 sin(col/20).  Returns x (2B,C,S,S) fp32 and cam (2B,4,4) fp32; img = x[:B], img_rot = x[B:]."""
    np.random.seed(seed)
    cam = get_camera_matries(sample_camera_prior(2 * B, ranges, bool(uniform)))
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, size=(2 * B, C, S, S)).astype("float32")
    if depth == "rough":
        x[:, -1] = rng.uniform(0.7, 1.5, size=(2 * B, S, S))
    elif depth == "smooth":
        col = np.arange(S, dtype="float32")[None, None, :]
        x[:, -1] = 1 + 0.1 * np.sin(col / 20 * (128.0 / S))
    else:
        raise ValueError(depth)
    return x, cam


# ------------------------------------------------------------------------ consistency loss
class LossFuncRotateNP:
    """common/loss_functions.py:31-146 with a hand-written reverse pass."""

    def __init__(self, K=None, norm="l1", lambda_geometric=3):
        self.size = None
        self.K = K
        self.norm = norm
        self.lambda_geometric = lambda_geometric

    def init_params(self, size):
        """:39-61"""
        if self.size is None:
            if self.K is not None:
                self.K = np.array(self.K[:3, :3], "float32")
                self.K[:2] *= size / self.K[0, 2] / 2
            else:
                self.K = np.array([[size * 2, 0, size / 2], [0, size * 2, size / 2], [0, 0, 1]], dtype="float32")
            self.size = size
        else:
            self.size = size
            self.K[:2] *= size / self.K[0, 2] / 2
        self.inv_K = np.linalg.inv(self.K).astype("float32")
        self.p = np.asarray(list(np.meshgrid(np.arange(size), np.arange(size))) + [np.ones((size, size))],
                            dtype="float32").reshape(3, -1)

    def pose_algebra(self, theta, theta_rot):
        """:85-91 + the constant factors of warp/inv_warp (:174,:181).
        Returns M (B,3,3), c (B,3,1) [subtracted], Mi (B,3,3), ci (B,3,1) [added]."""
        K, inv_K = self.K, self.inv_K
        R1, R2 = theta[:, :3, :3], theta_rot[:, :3, :3]
        t1, t2 = theta[:, :3, -1:], theta_rot[:, :3, -1:]
        R = np.matmul(R2.transpose(0, 2, 1), R1).astype("float32")
        inv_R = R.transpose(0, 2, 1)
        t = np.matmul(R1.transpose(0, 2, 1), t2 - t1).astype("float32")
        M = np.matmul(np.matmul(K, R), inv_K)
        c = np.matmul(np.matmul(K, R), t)
        Mi = np.matmul(np.matmul(K, inv_R), inv_K)
        ci = np.matmul(K, t)
        return M, c, Mi, ci

    # -- bilinear (:185-228), forward with tape
    @staticmethod
    def _bilinear_fwd(img, zp):
        b, hw, _ = zp.shape
        _, C, h, w = img.shape
        zpf = zp.reshape(-1, 3)
        q0, q1, q2 = zpf[:, 0], zpf[:, 1], zpf[:, 2]
        zc0 = np.clip(q2, 1e-4, 10000).astype(f32)
        zc1 = np.clip(q2, 1e-4, 10000).astype(f32)
        uu = q0 / zc0
        vv = q1 / zc1
        v, u = uu, vv
        u0 = u.astype("int32"); u1 = u0 + 1
        v0 = v.astype("int32"); v1 = v0 + 1
        a = u1.astype(f32) - u; bb = u - u0.astype(f32)
        cc = v1.astype(f32) - v; d = v - v0.astype(f32)
        w1 = a * cc; w2 = bb * cc; w3 = a * d; w4 = bb * d
        img_coord = np.arange(b * hw) // hw
        m = (u >= 0) * (u < h - 1) * (v >= 0) * (v < w - 1) * (q2 > 1e-4)
        u0 = u0 * m; u1 = u0 * m; v0 = v0 * m; v1 = v1 * m        # Q1: u1 := u0
        mf = m.astype(f32)
        w1m = w1 * mf; w2m = w2 * mf; w3m = w3 * mf; w4m = w4 * mf
        g1 = img[img_coord, :, u0, v0]; g2 = img[img_coord, :, u1, v0]
        g3 = img[img_coord, :, u0, v1]; g4 = img[img_coord, :, u1, v1]
        warped = w1m[:, None] * g1 + w2m[:, None] * g2 + w3m[:, None] * g3 + w4m[:, None] * g4
        tape = dict(q0=q0, q1=q1, q2=q2, zc0=zc0, zc1=zc1, a=a, bb=bb, cc=cc, d=d, mf=mf,
                    w=(w1m, w2m, w3m, w4m), g=(g1, g2, g3, g4),
                    idx=((img_coord, slice(None), u0, v0), (img_coord, slice(None), u1, v0),
                         (img_coord, slice(None), u0, v1), (img_coord, slice(None), u1, v1)),
                    shape=img.shape, zshape=zp.shape)
        return warped, m, tape

    @staticmethod
    def _bilinear_bwd(tape, g_warped):
        """returns g_img, g_zp (b,hw,3)"""
        g_img = None
        gw = []
        for k in range(4):
            gk = g_warped * tape["w"][k][:, None]                 # Mul backward (gathered operand)
            gw.append((g_warped * tape["g"][k]).sum(axis=1))     # sum_to (N,1) -> weight grad
            gx = np.zeros(tape["shape"], f32)
            np.add.at(gx, tape["idx"][k], gk)                    # GetItemGrad
            g_img = gx if g_img is None else g_img + gx
        mf, a, bb, cc, d = tape["mf"], tape["a"], tape["bb"], tape["cc"], tape["d"]
        gw = [g * mf for g in gw]                                # through `w * not_getting_out`
        g_a = gw[0] * cc + gw[2] * d
        g_bb = gw[1] * cc + gw[3] * d
        g_cc = gw[0] * a + gw[1] * bb
        g_d = gw[2] * a + gw[3] * bb
        g_u = (-g_a) + g_bb
        g_v = (-g_cc) + g_d
        # v = q0 / zc0 ; u = q1 / zc1   (DivGrad)
        gq0 = g_v / tape["zc0"]
        gzc0 = -gq0 * tape["q0"] / tape["zc0"]
        gq1 = g_u / tape["zc1"]
        gzc1 = -gq1 * tape["q1"] / tape["zc1"]
        q2 = tape["q2"]
        cond = (1e-4 <= q2) & (q2 <= 10000)
        gq2 = gzc0 * cond + gzc1 * cond
        g_zp = np.stack([gq0, gq1, gq2.astype(f32)], axis=1).reshape(tape["zshape"])
        return g_img, g_zp

    def forward(self, img, theta, img_rot, theta_rot, occlusion_aware=False, max_depth=None, min_depth=None):
        """:63-146.  Returns (loss fp32 0-dim, new_zp_cat (2B,HW,3)); keeps the tape for backward()."""
        if self.size != img.shape[-1]:
            self.init_params(img.shape[-1])
        B, C = img.shape[:2]
        z = img[:, -1:].reshape(B, 1, -1)
        z_rot = img_rot[:, -1:].reshape(B, 1, -1)
        M, c, Mi, ci = self.pose_algebra(theta, theta_rot)
        new_zp = (np.matmul(M, z * self.p) - c).transpose(0, 2, 1)
        new_zp_rot = (np.matmul(Mi, z_rot * self.p) + ci).transpose(0, 2, 1)
        warped, not_out, tp = self._bilinear_fwd(img_rot, new_zp)
        warped_rot, not_out_rot, tp_rot = self._bilinear_fwd(img, new_zp_rot)

        def target(im, zp_, no):
            return np.concatenate([im[:, :-1].transpose(0, 2, 3, 1).reshape(-1, C - 1),
                                   zp_[:, :, 2].reshape(-1, 1)], axis=1) * no[:, None].astype(f32)

        tgt, tgt_rot = target(img, new_zp, not_out), target(img_rot, new_zp_rot, not_out_rot)
        masks, masks_rot = [], []                                 # masks applied after bilinear, in order
        wm, wm_rot = warped, warped_rot
        if occlusion_aware:
            o = (warped[:, -1:] > new_zp[:, :, 2].reshape(-1, 1)).astype(f32)
            o_rot = (warped_rot[:, -1:] > new_zp_rot[:, :, 2].reshape(-1, 1)).astype(f32)
            wm, wm_rot, tgt, tgt_rot = wm * o, wm_rot * o_rot, tgt * o, tgt_rot * o_rot
            masks.append(o); masks_rot.append(o_rot)
        if max_depth is not None:
            s = (z.transpose(0, 2, 1).reshape(-1, 1) < max_depth).astype(f32)
            s_rot = (z_rot.transpose(0, 2, 1).reshape(-1, 1) < max_depth).astype(f32)
            wm, tgt, wm_rot, tgt_rot = wm * s, tgt * s, wm_rot * s_rot, tgt_rot * s_rot
            masks.append(s); masks_rot.append(s_rot)
        if min_depth is not None:
            s = (z.transpose(0, 2, 1).reshape(-1, 1) > min_depth).astype(f32)
            s_rot = (z_rot.transpose(0, 2, 1).reshape(-1, 1) > min_depth).astype(f32)
            wm, tgt, wm_rot, tgt_rot = wm * s, tgt * s, wm_rot * s_rot, tgt_rot * s_rot
            masks.append(s); masks_rot.append(s_rot)

        def crit(x0, x1):
            diff = x0 - x1
            r = diff.ravel()
            if self.norm == "l1":
                return np.array(abs(r).sum() / r.size, dtype=f32), diff
            return np.array(r.dot(r) / r.size, dtype=f32), diff

        l_rgb, d_rgb = crit(wm[:, :-1], tgt[:, :-1])
        l_rgb_r, d_rgb_r = crit(wm_rot[:, :-1], tgt_rot[:, :-1])
        l_d, d_d = crit(wm[:, -1], tgt[:, -1])
        l_d_r, d_d_r = crit(wm_rot[:, -1], tgt_rot[:, -1])
        lam = f32(self.lambda_geometric)
        loss = (l_rgb + l_rgb_r) + (l_d * lam + l_d_r * lam)
        self.parts = (l_rgb, l_rgb_r, l_d, l_d_r)
        self._tape = dict(B=B, C=C, shape=img.shape, M=M, Mi=Mi, tp=tp, tp_rot=tp_rot, masks=masks,
                          masks_rot=masks_rot, not_out=not_out, not_out_rot=not_out_rot,
                          diffs=(d_rgb, d_rgb_r, d_d, d_d_r))
        self.debug = dict(warped=warped, not_out=not_out, warped_rot=warped_rot, not_out_rot=not_out_rot)
        return loss.astype(f32), np.concatenate([new_zp, new_zp_rot], axis=0)

    def backward(self, gy=1.0):
        """reverse pass of forward(); returns (g_img, g_img_rot)"""
        T = self._tape
        B, C, shape = T["B"], T["C"], T["shape"]
        H, W = shape[2], shape[3]
        gy = f32(gy)
        lam = f32(self.lambda_geometric)

        def crit_bwd(g, diff):
            if self.norm == "l1":
                coeff = g * f32(1. / diff.size)
                return coeff * np.sign(diff)
            return g * diff * f32(2. / diff.size)

        def direction(diff_rgb, diff_d, masks, not_out, tp, Mm):
            g_w = np.empty((diff_d.size, C), f32)
            g_w[:, :-1] = crit_bwd(gy, diff_rgb)
            g_w[:, -1] = crit_bwd(lam * gy, diff_d)
            g_t = -g_w
            for mk in reversed(masks):
                g_w = g_w * mk
                g_t = g_t * mk
            g_t = g_t * not_out[:, None].astype(f32)
            g_oth, g_zp = self._bilinear_bwd(tp, g_w)
            g_src = np.zeros(shape, f32)
            g_src[:, :-1] += g_t[:, :-1].reshape(B, H, W, C - 1).transpose(0, 3, 1, 2)
            g_zp = g_zp.copy()
            g_zp[:, :, 2] += g_t[:, -1].reshape(B, -1)
            g_P = np.matmul(Mm.transpose(0, 2, 1), g_zp.transpose(0, 2, 1))     # MatMul backward
            g_z = (g_P * self.p).sum(axis=1, keepdims=True)                       # z*p backward (sum_to)
            g_src[:, -1:] += g_z.reshape(B, 1, H, W)
            return g_src, g_oth

        d_rgb, d_rgb_r, d_d, d_d_r = T["diffs"]
        gs1, go1 = direction(d_rgb, d_d, T["masks"], T["not_out"], T["tp"], T["M"])
        gs2, go2 = direction(d_rgb_r, d_d_r, T["masks_rot"], T["not_out_rot"], T["tp_rot"], T["Mi"])
        return gs1 + go2, gs2 + go1


def depth_hinge(x_fake, depth_min, lambda_depth, gy=None):
    """updater.py:357-359: F.mean(F.relu(depth_min - x_fake[:, -1]) ** 2) * lambda_depth, and (gy given) its
    gradient w.r.t. x_fake, node by node"""
    d = x_fake[:, -1]
    h = np.maximum(f32(depth_min) - d, 0, dtype=f32)
    sq = h ** f32(2)
    val = (sq.mean(dtype=f32) * f32(lambda_depth)).astype(f32)
    if gy is None:
        return val
    g = np.zeros_like(x_fake)
    g_mean = np.broadcast_to((f32(lambda_depth) * f32(gy)) * f32(1.0 / sq.size), sq.shape)
    g[:, -1] = -((f32(2) * h * g_mean) * (h > 0))
    return val, g


# ---------------------------------------------------------------------- DeepVoxels projection
def depth_head_fwd(h):
    """net.py:294-299 / :756-761 op by op (Chainer's softplus forward_cpu, AddConstant, DivFromConstant, Concat)"""
    h = np.asarray(h, f32)
    x = h[:, -1:]
    sp = (np.fmax(x, 0) + np.log1p(np.exp(-np.fabs(x)))).astype(f32)
    return np.concatenate([h[:, :-1], (f32(1.0) / (sp + f32(1e-4))).astype(f32)], axis=1)


def depth_head_bwd(h, g_out):
    h, g_out = np.asarray(h, f32), np.asarray(g_out, f32)
    x = h[:, -1:]
    sp = (np.fmax(x, 0) + np.log1p(np.exp(-np.fabs(x)))).astype(f32)
    v = sp + f32(1e-4)
    g_v = (-f32(1.0) * g_out[:, -1:] / (v ** 2)).astype(f32)                  # DivFromConstant.backward
    g_x = ((1 - 1 / (1 + np.exp(x))) * g_v).astype(f32)                       # SoftplusGrad.forward_cpu
    return np.concatenate([g_out[:, :-1], g_x], axis=1)


class ProjectionHelperNP:
    """deepvoxel/projection.py:5-105 with fp32-pinned scalar semantics (Q6)."""

    def __init__(self, projection_intrinsic, projection_image_dims, grid_dims, voxel_size, near_plane,
                 frustrum_depth):
        self.projection_intrinsic = projection_intrinsic
        self.projection_image_dims = projection_image_dims
        self.grid_dims = grid_dims
        self.voxel_size = voxel_size
        self.near_plane = f32(near_plane)
        self.frustrum_depth = frustrum_depth

    def compute_proj_idcs(self, cam2world, grid2world=None):
        dims, K = self.projection_image_dims, self.projection_intrinsic
        if grid2world is not None:
            world2grid = np.linalg.inv(grid2world)
        n = dims[0] * dims[1] * int(self.frustrum_depth)
        lin = np.arange(0, n).astype("int32")
        coords = np.zeros((4, n), dtype="float32")
        coords[2] = lin // (dims[0] * dims[1])
        tmp = lin - (coords[2] * dims[0] * dims[1]).astype("int32")
        coords[1] = tmp / dims[0]
        coords[0] = tmp % dims[0]
        coords[3].fill(1)
        coords[2] *= f32(self.voxel_size)
        coords[2] += self.near_plane
        coords[0] = (coords[0] - f32(K[0][2])) / f32(K[0][0])
        coords[1] = (coords[1] - f32(K[1][2])) / f32(K[1][1])
        coords[:2] *= coords[2]
        grid_coords = np.dot(cam2world, coords)
        if grid2world is not None:
            grid_coords = np.dot(world2grid, grid_coords)
        vc = grid_coords[:3, :] / f32(self.voxel_size)
        vc = vc + f32(self.grid_dims[2] / 2)
        mask = np.all(vc >= 0, axis=0)
        mask = mask * (vc[0] < self.grid_dims[0]) * (vc[1] < self.grid_dims[1]) * (vc[2] < self.grid_dims[2])
        if not mask.any():
            return None
        return lin[mask], vc[:, mask]


def _trilinear_terms(grid_shape, voxel_coords):
    _, _, height, width, depth = grid_shape
    xi, yi, zi = voxel_coords[2, :], voxel_coords[1, :], voxel_coords[0, :]
    x0, y0, z0 = xi.astype("int32"), yi.astype("int32"), zi.astype("int32")
    x1 = np.clip(x0 + 1, 0, width - 1); y1 = np.clip(y0 + 1, 0, height - 1); z1 = np.clip(z0 + 1, 0, depth - 1)
    x = xi - x0.astype(np.float64); y = yi - y0.astype(np.float64); z = zi - z0.astype(np.float64)   # fp64 (Q7)
    X = ((1 - x).astype(f32), x.astype(f32)); Y = ((1 - y).astype(f32), y.astype(f32))
    Z = ((1 - z).astype(f32), z.astype(f32))
    xs, ys, zs = (x0, x1), (y0, y1), (z0, z1)
    order = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (0, 1, 1), (1, 1, 0), (1, 1, 1)]
    return [((xs[i], ys[j], zs[k]), (X[i], Y[j], Z[k])) for i, j, k in order]


def interpolate_trilinear_fwd(grid, lin_ind, voxel_coords, img_shape, frustrum_depth):
    """deepvoxel/deepvoxel.py:388-428 forward"""
    batch, F = grid.shape[:2]
    added = None
    for (ix, iy, iz), (wx, wy, wz) in _trilinear_terms(grid.shape, voxel_coords):
        term = grid[:, :, ix, iy, iz] * wx * wy * wz
        added = term if added is None else added + term
    out = np.zeros((batch, F, img_shape[0] * img_shape[1] * frustrum_depth), dtype="float32")
    np.add.at(out, (slice(None), slice(None), lin_ind), added)
    return out.reshape(batch, F, frustrum_depth, img_shape[0], img_shape[1])


def interpolate_trilinear_bwd(grid_shape, lin_ind, voxel_coords, g_out):
    """autograd of the above: g_out (b,F,D,H,W) -> g_grid"""
    b, F = grid_shape[:2]
    g_added = g_out.reshape(b, F, -1)[:, :, lin_ind]
    g_grid = None
    for (ix, iy, iz), (wx, wy, wz) in _trilinear_terms(grid_shape, voxel_coords):
        gx = np.zeros(grid_shape, f32)
        np.add.at(gx, (slice(None), slice(None), ix, iy, iz), g_added * wz * wy * wx)
        g_grid = gx if g_grid is None else g_grid + gx
    return g_grid
