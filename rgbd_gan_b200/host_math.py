"""Host-side NumPy pieces shared by the torch glue (loss_functions.py, projection.py) and the Chainer / CuPy glue
(chainer_nodes.py).  No torch, no chainer, no cupy import: a deployment on the reference's own stack (Chainer + CuPy,
README.md:19-27) must be able to import the Chainer nodes without torch being installed.

The reference evaluates these quantities with raw `xp.matmul` / `xp.linalg.inv` on constants (no gradient flows through
them); doing it on the host with the same NumPy call sequence makes the matrices the kernels receive identical to the
reference's by construction.
"""
import numpy as np


def as_numpy(a):
    """theta / K style inputs -> ndarray: ndarray, list, chainer.Variable (`.array`), cupy.ndarray (`.get()`),
    torch.Tensor (`.detach().cpu().numpy()`), duck-typed so that none of those packages is imported here."""
    if hasattr(a, "array") and not isinstance(a, np.ndarray):
        a = a.array
    if isinstance(a, np.ndarray):
        return a
    if hasattr(a, "get") and hasattr(a, "dtype") and not isinstance(a, dict):      # cupy.ndarray
        return np.asarray(a.get())
    if hasattr(a, "detach") and hasattr(a, "cpu"):                                # torch.Tensor
        return a.detach().cpu().numpy()
    return np.asarray(a)


def intrinsics_for_size(K, size, first):
    """common/loss_functions.py:39-56.  first=True: K is the constructor argument (None -> the default pinhole of :48-50;
    a given (4,4) / (3,3) matrix is cut to 3x3 and rescaled, :43-44); first=False: K is the cached matrix, rescaled IN
    PLACE as the reference does on a size change (:52-54, quirk Q9).  Returns (K, inv_K) float32."""
    if first:
        if K is not None:
            K = np.array(as_numpy(K)[:3, :3], "float32")
            K[:2] *= size / K[0, 2] / 2
        else:
            K = np.array([[size * 2, 0, size / 2],
                          [0, size * 2, size / 2],
                          [0, 0, 1]], dtype="float32")
    else:
        K[:2] *= size / K[0, 2] / 2
    return K, np.linalg.inv(K).astype("float32")


def pixel_grid(size):
    """:59-61 -> p (3, size*size): p[0] = column, p[1] = row, p[2] = 1 (row-major)"""
    return np.asarray(list(np.meshgrid(np.arange(size), np.arange(size))) + [np.ones((size, size))],
                      dtype="float32").reshape(3, -1)


def pose_algebra(K, inv_K, theta, theta_rot):
    """common/loss_functions.py:85-91 and the constant factors of warp (:174) / inv_warp (:181),
    evaluated with the same NumPy matmul sequence on the host.
    Returns float32 arrays M (B,3,3), c (B,3,1), Mi (B,3,3), ci (B,3,1) with the convention of
    include/rgbdgan_b200.h: new_zp = M (z p) - c ; new_zp_rot = Mi (z_rot p) - ci  (ci = -(K t))."""
    theta, theta_rot = as_numpy(theta), as_numpy(theta_rot)
    R1, R2 = theta[:, :3, :3], theta_rot[:, :3, :3]
    t1, t2 = theta[:, :3, -1:], theta_rot[:, :3, -1:]
    R = np.matmul(R2.transpose(0, 2, 1), R1).astype("float32")
    inv_R = R.transpose(0, 2, 1)
    t = np.matmul(R1.transpose(0, 2, 1), t2 - t1).astype("float32")
    M = np.matmul(np.matmul(K, R), inv_K)
    c = np.matmul(np.matmul(K, R), t)
    Mi = np.matmul(np.matmul(K, inv_R), inv_K)
    ci = -np.matmul(K, t)
    return (np.ascontiguousarray(M, dtype=np.float32), np.ascontiguousarray(c, dtype=np.float32),
            np.ascontiguousarray(Mi, dtype=np.float32), np.ascontiguousarray(ci, dtype=np.float32))


def warp_constants(K, inv_K, R, t, inverse):
    """the constant factors of warp (:174: K R K^-1, (K R) t) or inv_warp (:181: K R^T K^-1, -(K t)) as (M, cv)
    float32 with new_zp = M (z p) - cv"""
    K, inv_K, R, t = (as_numpy(a).astype("float32") for a in (K, inv_K, R, t))
    M = np.ascontiguousarray(np.matmul(np.matmul(K, R), inv_K), dtype=np.float32)
    cv = -np.matmul(K, t) if inverse else np.matmul(np.matmul(K, R), t)
    return M, np.ascontiguousarray(cv, dtype=np.float32)


def grid_dims(p, hw):
    """(H, W) of the pixel grid `p` (3, H*W) the reference passes to warp / inv_warp"""
    p = as_numpy(p)
    if p.shape != (3, hw):
        raise ValueError("p must be (3, H*W)")
    W, H = int(p[0].max()) + 1, int(p[1].max()) + 1
    if W * H != hw:
        raise ValueError("p is not a full pixel grid")
    return H, W


def combine_loss_parts(parts, lambda_geometric):
    """loss = (rgb + rgb_rot) + (depth*lambda + depth_rot*lambda), fp32 (:141-144).
    `parts`: tensor/array of the four means (summed over shards)."""
    lam = float(lambda_geometric)
    return (parts[0] + parts[1]) + (parts[2] * lam + parts[3] * lam)


def device_pointer(a):
    """raw device address of a CUDA array, whatever the container: cupy (`a.data.ptr`), torch (`a.data_ptr()`), or any
    object that implements `__cuda_array_interface__` (Numba, a custom allocator); None -> 0"""
    if a is None:
        return 0
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    d = getattr(a, "data", None)
    if d is not None and hasattr(d, "ptr"):
        return int(d.ptr)
    cai = getattr(a, "__cuda_array_interface__", None)
    if cai is not None:
        return int(cai["data"][0])
    raise TypeError("not a CUDA array: %r" % type(a))
