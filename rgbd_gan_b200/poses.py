"""Host-side input generators of the hot path: the camera-pose prior and the 6-DoF ->
cam2world conversion.  Restated from train_rgbd.py:192-217 (CameraParamPrior) and
updater.py:26-60 (update_camera_matrices / get_camera_matries; duplicated at
updater_deepvoxels.py:29-63).  NumPy on the host, exactly as in the reference.
"""
import numpy as np

# yml pose ranges: [x_rotate, y_rotate, z_rotate, x_translate, y_translate, z_translate]
FFHQ_RANGES = (0.3054, 1.0472, 0, 0, 0, 0)          # configs/ffhq_stylegan_occlusion.yml:37-43
CAR_RANGES = (0.3054, 3.1415, 0, 0, 0, 0)           # configs/dcgan_shapenet_car.yml:38-44


def update_camera_matrices(mat, axis1, axis2, theta):
    """left-multiply a rotation in the (axis1, axis2) plane (updater.py:26-42)"""
    rot = np.zeros_like(mat)
    rot[:, range(4), range(4)] = 1
    rot[:, axis1, axis1] = np.cos(theta)
    rot[:, axis1, axis2] = -np.sin(theta)
    rot[:, axis2, axis1] = np.sin(theta)
    rot[:, axis2, axis2] = np.cos(theta)
    return np.matmul(rot, mat)


def get_camera_matries(thetas, order=(0, 1, 2)):
    """thetas (B,6) [x,y,z rotation, x,y,z translation] -> cam2world (B,4,4) fp32 (updater.py:45-60).
    The camera starts at z=1 looking at the origin (diag(1,1,-1,1), mat[2,3]=1)."""
    mat = np.zeros((len(thetas), 4, 4), dtype="float32")
    mat[:, range(4), range(4)] = [1, 1, -1, 1]
    mat[:, 2, 3] = 1
    for i in order:
        mat = update_camera_matrices(mat, (i + 1) % 3, (i + 2) % 3, thetas[:, i])
    mat[:, :3, 3] = mat[:, :3, 3] + thetas[:, 3:]
    return mat


class CameraParamPrior:
    """train_rgbd.py:192-217.  `config` needs x/y/z_rotate, x/y/z_translate, uniform_distribution."""

    def __init__(self, config):
        self.rotation_range = np.array([config.x_rotate, config.y_rotate, config.z_rotate])
        self.camera_param_range = np.array([config.x_rotate, config.y_rotate, config.z_rotate,
                                            config.x_translate, config.y_translate, config.z_translate])
        self.uniform = config.uniform_distribution

    @classmethod
    def from_ranges(cls, ranges, uniform=None):
        class _C:
            pass
        c = _C()
        (c.x_rotate, c.y_rotate, c.z_rotate, c.x_translate, c.y_translate, c.z_translate) = ranges
        c.uniform_distribution = uniform
        return cls(c)

    def sample(self, batch_size):
        """first half U(-1,1)^6, second half = perturbed twins; uses the global np.random like the reference"""
        h = batch_size // 2
        thetas = np.random.uniform(-1, 1, size=(h, 6))
        eps = np.random.uniform(0, 0.5, size=(h, 6))
        sign = np.random.choice(2, size=(h, 3)) * 2 - 1
        limit = np.clip(1 / (self.rotation_range + 1e-8), 0, 1)      # limit angle difference
        if self.uniform:
            eps[:, :3] = eps[:, :3] * sign * limit
        else:
            eps[:, :3] = eps[:, :3] * (sign * (self.rotation_range == 3.1415) +
                                       np.abs(sign) * (self.rotation_range != 3.1415)) * limit
        thetas2 = -eps * np.sign(thetas) + thetas
        if self.uniform:
            thetas2 = thetas2 * (-1 <= thetas2) * (thetas2 <= 1) + (-2 - thetas2) * (thetas2 < -1) + \
                      (2 - thetas2) * (thetas2 > 1)
        thetas = np.concatenate([thetas, thetas2], axis=0)
        thetas = thetas * self.camera_param_range[None]
        return thetas.astype("float32")


def synthetic_batch(B, S, C=4, depth="rough", ranges=FFHQ_RANGES, uniform=None, seed=0):
    """The synthetic workload of SURVEY.md 8(d): 2B images (B pairs) + their cam2world matrices.
    RGB ~ U(-1,1) (the data range of train_rgbd.py:308); depth "rough" ~ U(0.7,1.5) (worst-case
    scatter) or "smooth" = 1 + 0.1 sin(col/20) (generator-like: depth ~ 1, net.py:211-214,296).
    Returns x (2B,C,S,S) fp32 and cam (2B,4,4) fp32; img = x[:B], img_rot = x[B:]."""
    np.random.seed(seed)
    thetas = CameraParamPrior.from_ranges(ranges, uniform).sample(2 * B)
    cam = get_camera_matries(thetas)
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
