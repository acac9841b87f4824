"""Synthetic inputs for bench.py and the profiling scripts: random RGB-D pairs + twin camera poses with the
statistics SURVEY.md 8(d) prescribes.  Not part of the package and not a checker -- it only feeds the benchmark.

The pose distribution follows what the reference's trainer samples (train_rgbd.py:192-217: a pose uniform in the yml
ranges and a twin displaced by up to half a normalised unit per axis, rotation displacement capped at 0.5 rad) and the
cam2world convention of updater.py:45-60 (camera on the unit sphere looking at the origin), written here in closed
form with NumPy's Generator API.  The bit-faithful replay of the reference's own helpers (global np.random stream,
float32 matmul chain) lives with the test oracle (oracle/numpy_port.py) and is what the parity tests use.
"""
import numpy as np

# yml pose ranges: [x_rotate, y_rotate, z_rotate, x_translate, y_translate, z_translate]
FFHQ_RANGES = (0.3054, 1.0472, 0, 0, 0, 0)          # configs/ffhq_stylegan_occlusion.yml:37-43
CAR_RANGES = (0.3054, 3.1415, 0, 0, 0, 0)           # configs/dcgan_shapenet_car.yml:38-44


def cam2world(thetas):
    """(n,6) [rx, ry, rz, tx, ty, tz] -> (n,4,4) fp32: R = Rz Ry Rx, camera axes R diag(1,1,-1), position R e_z + t"""
    th = np.asarray(thetas, np.float64)
    cx, sx, cy, sy, cz, sz = (f(th[:, i]) for i in range(3) for f in (np.cos, np.sin))
    one, zero = np.ones_like(cx), np.zeros_like(cx)
    Rx = np.stack([one, zero, zero, zero, cx, -sx, zero, sx, cx], -1).reshape(-1, 3, 3)
    Ry = np.stack([cy, zero, sy, zero, one, zero, -sy, zero, cy], -1).reshape(-1, 3, 3)
    Rz = np.stack([cz, -sz, zero, sz, cz, zero, zero, zero, one], -1).reshape(-1, 3, 3)
    R = Rz @ Ry @ Rx
    out = np.zeros((len(th), 4, 4))
    out[:, :3, :3] = R * np.array([1.0, 1.0, -1.0])
    out[:, :3, 3] = R[:, :, 2] + th[:, 3:]
    out[:, 3, 3] = 1.0
    return out.astype(np.float32)


def sample_pose_pairs(n_pairs, ranges=FFHQ_RANGES, uniform=False, rng=None):
    """(2*n_pairs, 6) fp32: rows [0, n) are the poses, rows [n, 2n) their perturbed twins"""
    rng = rng or np.random.default_rng(0)
    ranges = np.asarray(ranges, np.float64)
    base = rng.uniform(-1.0, 1.0, size=(n_pairs, 6))
    step = rng.uniform(0.0, 0.5, size=(n_pairs, 6))
    flip = rng.integers(0, 2, size=(n_pairs, 3)) * 2.0 - 1.0
    cap = np.minimum(1.0 / (ranges[:3] + 1e-8), 1.0)                 # at most 0.5 rad between twins
    full_circle = ranges[:3] == 3.1415
    direction = flip if uniform else np.where(full_circle, flip, 1.0)
    step[:, :3] *= direction * cap
    twin = base - step * np.sign(base)
    if uniform:                                                      # reflect back into [-1, 1]
        twin = np.where(twin < -1, -2 - twin, np.where(twin > 1, 2 - twin, twin))
    return (np.concatenate([base, twin]) * ranges).astype(np.float32)


def synthetic_batch(B, S, C=4, depth="rough", ranges=FFHQ_RANGES, uniform=False, seed=0):
    """2B images (B pairs) + their cam2world matrices.  RGB ~ U(-1,1) (the data range of train_rgbd.py:308); depth
    "rough" ~ U(0.7,1.5) (worst-case scatter) or "smooth" = 1 + 0.1 sin(col/20) (generator-like: depth ~ 1).
    Returns x (2B,C,S,S) fp32 and cam (2B,4,4) fp32; img = x[:B], img_rot = x[B:]."""
    rng = np.random.default_rng(seed)
    cam = cam2world(sample_pose_pairs(B, ranges, uniform, rng))
    x = rng.uniform(-1, 1, size=(2 * B, C, S, S)).astype(np.float32)
    if depth == "rough":
        x[:, -1] = rng.uniform(0.7, 1.5, size=(2 * B, S, S))
    elif depth == "smooth":
        col = np.arange(S, dtype=np.float32)[None, None, :]
        x[:, -1] = 1 + 0.1 * np.sin(col / 20 * (128.0 / S))
    else:
        raise ValueError(depth)
    return x, cam
