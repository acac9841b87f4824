"""device-resident throughput of the public Python API (host overhead check)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tools import synthetic as poses
from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra
B, S = 32, 128
x, cam = poses.synthetic_batch(B, S, seed=0)
xd = torch.from_numpy(x).cuda()
f = LossFuncRotate(None, lambda_geometric=3, grad_scale=2.0, return_new_zp=False)
def step():
    img = xd[:B].detach().requires_grad_(True); img_rot = xd[B:].detach().requires_grad_(True)
    loss, _ = f(img, cam[:B], img_rot, cam[B:], occlusion_aware=True)
    (loss * 2.0).backward()
for _ in range(20): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 300
for _ in range(n): step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
print("API fwd+bwd per call: %.1f us -> %.0f pairs/s" % (dt * 1e6, B / dt))
f.init_params(None, size=S)
t0 = time.perf_counter()
for _ in range(300): pose_algebra(f.K, f.inv_K, cam[:B], cam[B:])
print("pose_algebra (numpy): %.1f us" % ((time.perf_counter() - t0) / 300 * 1e6))
