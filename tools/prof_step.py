"""A few fwd+bwd steps of the headline workload (32 pairs @128x128) for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:k_consistency -s 6 -c 1 -o gpurun_out/prof python tools/prof_step.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rgbd_gan_b200 import _lib
from tools import synthetic as poses
from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra

B = int(os.environ.get("PAIRS", "32")); S = int(os.environ.get("SIZE", "128")); C = 4
steps = int(os.environ.get("STEPS", "8")); depth = os.environ.get("DEPTH", "rough")
dev = torch.device("cuda", 0)
lib = _lib.load()
hf = LossFuncRotate(None, lambda_geometric=3.0)
hf.init_params(None, size=S)
sets = []
for s in range(6):
    x, cam = poses.synthetic_batch(B, S, depth=depth, seed=s)
    M, c, Mi, ci = pose_algebra(hf.K, hf.inv_K, cam[:B], cam[B:])
    pv = torch.from_numpy(np.concatenate([M.reshape(-1), c.reshape(-1), Mi.reshape(-1), ci.reshape(-1)])).to(dev)
    xt = torch.from_numpy(x).to(dev)
    sets.append((xt[:B].contiguous(), xt[B:].contiguous(), pv))
g0, g1 = torch.empty((B, C, S, S), device=dev), torch.empty((B, C, S, S), device=dev)
parts = torch.zeros(8, device=dev)
ws = torch.empty(lib.rgbd_consistency_workspace_bytes(B, C, S, S), dtype=torch.uint8, device=dev)
opts = _lib.LossOpts(_lib.NORM_L1, 1, float("nan"), float("nan"), 3.0, B, None)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for k in range(steps):
    im, ir, pv = sets[k % len(sets)]
    base = pv.data_ptr()
    pp = [ctypes.c_void_p(base + 4 * o) for o in (0, 9 * B, 12 * B, 21 * B)]
    _lib.call("rgbd_consistency_fwd_bwd", ctypes.c_void_p(im.data_ptr()), ctypes.c_void_p(ir.data_ptr()), *pp, B, C, S, S,
              ctypes.byref(opts), ctypes.c_float(2.0), ctypes.c_void_p(parts.data_ptr()), None, ctypes.c_void_p(g0.data_ptr()),
              ctypes.c_void_p(g1.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(), st)
torch.cuda.synchronize()
print("loss parts", parts.cpu().numpy())
