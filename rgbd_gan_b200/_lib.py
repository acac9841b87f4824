"""ctypes binding of librgbdgan_b200.so (include/rgbdgan_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises, and every
entry point turns a non-zero return code into RgbdB200Error.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RGBD_B200_LIB") or os.path.join(_HERE, "lib", "librgbdgan_b200.so")   # env: tuning variants

c_void = ctypes.c_void_p
c_int = ctypes.c_int
c_float = ctypes.c_float
c_size = ctypes.c_size_t

NORM_L1, NORM_L2 = 1, 2


class RgbdB200Error(RuntimeError):
    pass


class LossOpts(ctypes.Structure):
    """rgbd_loss_opts"""
    _fields_ = [("norm", c_int), ("occlusion_aware", c_int), ("max_depth", c_float), ("min_depth", c_float),
                ("lambda_geometric", c_float), ("n_pairs_global", ctypes.c_longlong), ("peer_comm", c_void),
                ("defer_loss", c_int), ("reserved", c_int), ("hinge_depth_min", c_float), ("hinge_lambda", c_float)]

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        if len(args) < 10 and "hinge_depth_min" not in kw:
            self.hinge_depth_min = float("nan")          # depth hinge off unless asked for


class DvParams(ctypes.Structure):
    """rgbd_dv_params"""
    _fields_ = [("W", c_int), ("H", c_int), ("D", c_int), ("G", c_int),
                ("fx", c_float), ("fy", c_float), ("cx", c_float), ("cy", c_float),
                ("voxel_size", c_float), ("near_plane", c_float)]


class PosePrior(ctypes.Structure):
    """rgbd_pose_prior"""
    _fields_ = [("camera_param_range", ctypes.c_double * 6), ("uniform_distribution", c_int)]


class DvRenderParams(ctypes.Structure):
    """rgbd_dv_render_params"""
    _fields_ = [("nf", c_int), ("depth_steps", c_int), ("threshold", c_float), ("inv_c1", c_float), ("inv_c2", c_float)]


# name -> (restype, argtypes); must list every symbol the header declares (tests check this)
SIGNATURES = {
    "rgbd_version": (c_int, []),
    "rgbd_last_error": (ctypes.c_char_p, []),
    "rgbd_launch_count": (ctypes.c_ulonglong, []),
    "rgbd_profile_hook": (None, [c_void, c_void]),
    "rgbd_peer_comm_create": (c_int, [c_int, c_int, ctypes.POINTER(c_void), ctypes.c_char_p]),
    "rgbd_peer_comm_connect": (c_int, [c_void, ctypes.c_char_p]),
    "rgbd_peer_comm_destroy": (c_int, [c_void]),
    "rgbd_peer_comm_wait": (c_int, [c_void, c_void]),
    "rgbd_peer_comm_status": (c_int, [c_void, c_void, ctypes.POINTER(c_int)]),
    "rgbd_debug_peer_comm_loopback": (c_int, [c_void]),
    "rgbd_consistency_workspace_bytes": (c_size, [c_int, c_int, c_int, c_int]),
    "rgbd_consistency_uses_sweep": (c_int, [c_int, c_int, c_int, c_int]),
    "rgbd_consistency_status": (c_int, [c_void, c_void, ctypes.POINTER(c_int)]),
    "rgbd_depth_head_fwd": (c_int, [c_void, c_int, c_int, c_int, c_int, c_void, c_void]),
    "rgbd_depth_head_bwd": (c_int, [c_void, c_void, c_int, c_int, c_int, c_int, c_void, c_void]),
    "rgbd_pose_sample": (c_int, [ctypes.POINTER(PosePrior), c_int, c_void, ctypes.c_ulonglong, ctypes.c_ulonglong, c_void,
                                 c_void]),
    "rgbd_pose_camera_matrices": (c_int, [c_void, c_void, c_int, ctypes.POINTER(c_int), c_void, c_void]),
    "rgbd_pose_algebra": (c_int, [c_void, c_void, c_int, ctypes.POINTER(c_float), ctypes.POINTER(c_float)] + [c_void] * 5),
    "rgbd_pose_pipeline": (c_int, [ctypes.POINTER(PosePrior), c_int, c_void, ctypes.c_ulonglong, ctypes.c_ulonglong,
                                   ctypes.POINTER(c_int), ctypes.POINTER(c_float), ctypes.POINTER(c_float)] + [c_void] * 7),
    "rgbd_debug_div2": (c_int, [ctypes.c_ulonglong, ctypes.c_uint, c_int, c_int, c_void, c_void]),
    "rgbd_debug_mega_schedule": (c_int, [c_int] * 7 + [ctypes.POINTER(c_int), c_int, ctypes.POINTER(c_int)]),
    "rgbd_consistency_fwd": (c_int, [c_void] * 6 + [c_int] * 4 + [ctypes.POINTER(LossOpts), c_void, c_void, c_void,
                                                                 c_void, c_size, c_void]),
    "rgbd_consistency_bwd": (c_int, [c_void] * 6 + [c_int] * 4 + [ctypes.POINTER(LossOpts), c_float, c_void, c_void,
                                                                 c_void, c_void, c_void, c_size, c_void]),
    "rgbd_consistency_fwd_bwd": (c_int, [c_void] * 6 + [c_int] * 4 + [ctypes.POINTER(LossOpts), c_float, c_void,
                                                                     c_void, c_void, c_void, c_void, c_size, c_void]),
    "rgbd_consistency_rescale": (c_int, [c_void, c_void, c_size, c_void, c_float, c_void]),
    "rgbd_warp_fwd": (c_int, [c_void, c_void, c_void, c_int, c_int, c_int, c_void, c_void]),
    "rgbd_warp_bwd": (c_int, [c_void, c_void, c_int, c_int, c_int, c_void, c_void]),
    "rgbd_bilinear_fwd": (c_int, [c_void, c_void, c_int, c_int, c_int, c_int, c_void, c_void, c_void]),
    "rgbd_bilinear_bwd": (c_int, [c_void, c_void, c_void, c_int, c_int, c_int, c_int, c_void, c_void, c_void]),
    "rgbd_dv_workspace_bytes": (c_size, [ctypes.POINTER(DvParams)]),
    "rgbd_dv_compute_proj_idcs": (c_int, [ctypes.POINTER(DvParams), c_void, c_void, c_void, ctypes.POINTER(c_int),
                                          c_void, c_size, c_void]),
    "rgbd_dv_compute_proj_idcs_g2w": (c_int, [ctypes.POINTER(DvParams), c_void, c_void, c_void, c_void, ctypes.POINTER(c_int),
                                              c_void, c_size, c_void]),
    "rgbd_dv_trilinear_fwd": (c_int, [c_void, c_void, c_void, c_int, c_int, c_int, ctypes.POINTER(DvParams), c_void,
                                      c_void]),
    "rgbd_dv_trilinear_bwd": (c_int, [c_void, c_void, c_void, c_int, c_int, c_int, ctypes.POINTER(DvParams), c_void,
                                      c_void]),
    "rgbd_dv_project_workspace_bytes": (c_size, [ctypes.POINTER(DvParams), c_int, c_int]),
    "rgbd_dv_render_workspace_bytes": (c_size, [ctypes.POINTER(DvParams), c_int, c_int]),
    "rgbd_dv_render_saved_bytes": (c_size, [ctypes.POINTER(DvParams), c_int]),
    "rgbd_dv_render_fwd": (c_int, [ctypes.POINTER(DvParams), ctypes.POINTER(DvRenderParams)] + [c_void] * 6 + [c_int, c_int] +
                           [c_void] * 5 + [c_size, c_void]),
    "rgbd_dv_render_bwd": (c_int, [ctypes.POINTER(DvParams), ctypes.POINTER(DvRenderParams)] + [c_void] * 6 + [c_int, c_int] +
                           [c_void] * 10 + [c_size, c_void]),
    "rgbd_dv_project_fwd": (c_int, [ctypes.POINTER(DvParams), c_void, c_void, c_int, c_int, c_void, c_void, c_size,
                                    c_void]),
    "rgbd_dv_project_bwd": (c_int, [ctypes.POINTER(DvParams), c_void, c_void, c_int, c_int, c_void, c_void, c_size,
                                    c_void]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises RgbdB200Error if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RgbdB200Error(
            "%s not found: build it with `python -m rgbd_gan_b200.build` (nvcc, sm_100a). "
            "rgbd_gan_b200 has no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().rgbd_last_error()
        raise RgbdB200Error("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else ""))


def call(name, *args):
    check(getattr(load(), name)(*args), name)
