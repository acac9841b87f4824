"""a few steps of the many-channel (C = 257) loss for ncu captures"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"]
import bench, torch
from rgbd_gan_b200 import _lib
torch.cuda.set_device(0)
ctx = dict(dev=torch.device("cuda", 0), lib=_lib.load(), hbm_peak=bench.peaks()[0])
print(bench.feature_consistency_bench(None, ctx))
