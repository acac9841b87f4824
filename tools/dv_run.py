"""run the DeepVoxels fused fwd/bwd a few times (for ncu): python tools/dv_run.py [G] [B] [planar]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rgbd_gan_b200 import _lib
from tools import synthetic as poses
G = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
planar = len(sys.argv) > 3
F, img = 32, 64
D = int(np.ceil(np.sqrt(3) * G)); vs = (1. / G) * 1.1 * 0.5
P = _lib.DvParams(img, img, D, G, 128., 128., 32., 32., float(np.float32(vs)), float(np.float32(np.sqrt(3) / 4)))
np.random.seed(3)
th = poses.sample_pose_pairs(B, poses.CAR_RANGES, True)[:B]
cam = torch.from_numpy(poses.cam2world(th).reshape(B, 16)).cuda()
n = img * img * D
lib = _lib.load()
grid = torch.randn((B, F, G, G, G), device="cuda"); fr = torch.empty((B, F, n), device="cuda"); gg = torch.empty((B, F, G ** 3), device="cuda")
ws = torch.empty(lib.rgbd_dv_project_workspace_bytes(ctypes.byref(P), B, F), dtype=torch.uint8, device="cuda")
wp = None if planar else ctypes.c_void_p(ws.data_ptr())
for k in range(3):
    _lib.call("rgbd_dv_project_fwd", ctypes.byref(P), ctypes.c_void_p(grid.data_ptr()), ctypes.c_void_p(cam.data_ptr()), B, F, ctypes.c_void_p(fr.data_ptr()), wp, ws.numel(), None)
    _lib.call("rgbd_dv_project_bwd", ctypes.byref(P), ctypes.c_void_p(fr.data_ptr()), ctypes.c_void_p(cam.data_ptr()), B, F, ctypes.c_void_p(gg.data_ptr()), wp, ws.numel(), None)
torch.cuda.synchronize()
print("ok")
