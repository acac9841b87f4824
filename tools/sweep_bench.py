"""batch-256 sweep of bench.py alone (cfg4), e.g. with RGBD_B200_STREAMS=1/2"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"]
import bench, torch
from rgbd_gan_b200 import _lib
torch.cuda.set_device(0)
ctx = dict(dev=torch.device("cuda", 0), lib=_lib.load(), hbm_peak=bench.peaks()[0])
a = argparse.Namespace(depth=os.environ.get("DEPTH", "rough"), size=128, pairs=32)
for r in bench.sweep(a, ctx, sizes=((128, 64), (128, 128), (128, 256), (256, 256))):
    print("streams", os.environ.get("RGBD_B200_STREAMS", "default"), r["pairs"], r["size"], "%.0f pairs/s  %.1f us  frac %.3f" % (r["pairs_per_s"], r["ms_per_step"] * 1e3, r["frac_of_hbm_peak"]))
