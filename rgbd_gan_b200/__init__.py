"""rgbd_gan_b200 -- B200-native (sm_100a) implementation of RGBD-GAN's 3D-consistency hot path.

Public surface mirrors the reference modules it replaces:
  rgbd_gan_b200.loss_functions  <- common/loss_functions.py  (LossFuncRotate, warp, inv_warp, bilinear)
  rgbd_gan_b200.projection      <- deepvoxel/projection.py + deepvoxel/deepvoxel.py:388-433
All compute goes through include/rgbdgan_b200.h (librgbdgan_b200.so); there is no CPU fallback.
"""
__version__ = "0.1.0"


def __getattr__(name):
    # lazy: importing the package must not require torch/CUDA (e.g. for `python -m rgbd_gan_b200.build`)
    if name in ("LossFuncRotate", "warp", "inv_warp", "bilinear"):
        from . import loss_functions
        return getattr(loss_functions, name)
    if name in ("ProjectionHelper", "interpolate_trilinear", "MakeSlice"):
        from . import projection
        return getattr(projection, name)
    raise AttributeError(name)
