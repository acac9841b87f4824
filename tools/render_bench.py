"""time the fused render kernels (bench.render_bench) with the library selected by RGBD_B200_LIB"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"]
import bench, torch
from rgbd_gan_b200 import _lib
torch.cuda.set_device(0)
ctx = dict(dev=torch.device("cuda", 0), lib=_lib.load(), hbm_peak=bench.peaks()[0])
for r in bench.render_bench(ctx):
    print(os.environ.get("RGBD_B200_LIB", "default")[-10:], "G", r["G"], "fwd_ms %.3f bwd_ms %.3f sat %.2f" % (r["fwd_ms"], r["bwd_ms"], r["saturated_rays"]))
