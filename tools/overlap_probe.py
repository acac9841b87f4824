"""Do the staging kernels (HBM-bound) and the main kernel (issue / L1-bound) overlap when they co-run?

Two independent batches on two streams (CUDA-graph replays of rgbd_consistency_fwd_bwd, separate workspaces)
against the same two batches back to back on one stream.  If the aggregate rate rises well above the
single-stream rate, fusing the three phases into one software-pipelined launch pays.
"""
import ctypes
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rgbd_gan_b200 import _lib
from tools import synthetic as poses
from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra

dev = torch.device("cuda", 0)
lib = _lib.load()
S, C = 128, 4
out = []
NS = int(os.environ.get("NSTREAMS", "2"))
for B in (4, 8, 16, 32):
    HW = S * S
    hf = LossFuncRotate(None, lambda_geometric=3.0)
    hf.init_params(None, size=S)
    n_sets = max(4, (3 * 126 * 2 ** 20) // (4 * B * C * HW * 4) + 1)
    sets = []
    for s in range(n_sets):
        x, cam = poses.synthetic_batch(B, S, depth="rough", seed=s % 4)
        M, c, Mi, ci = pose_algebra(hf.K, hf.inv_K, cam[:B], cam[B:])
        pv = torch.from_numpy(np.concatenate([M.reshape(-1), c.reshape(-1), Mi.reshape(-1), ci.reshape(-1)])).to(dev)
        xt = torch.from_numpy(x).to(dev)
        sets.append(dict(img=xt[:B].contiguous(), rot=xt[B:].contiguous(), pv=pv, g0=torch.empty((B, C, S, S), device=dev),
                         g1=torch.empty((B, C, S, S), device=dev), parts=torch.zeros(8, device=dev)))
    opts = _lib.LossOpts(_lib.NORM_L1, 1, float("nan"), float("nan"), 3.0, B, None)
    streams = [torch.cuda.Stream(device=dev) for _ in range(NS)]
    wss = [torch.empty(lib.rgbd_consistency_workspace_bytes(B, C, S, S), dtype=torch.uint8, device=dev) for _ in range(NS)]

    def step(e, ws, st):
        base = e["pv"].data_ptr()
        pp = [ctypes.c_void_p(base + 4 * o) for o in (0, 9 * B, 12 * B, 21 * B)]
        _lib.call("rgbd_consistency_fwd_bwd", ctypes.c_void_p(e["img"].data_ptr()), ctypes.c_void_p(e["rot"].data_ptr()), *pp,
                  B, C, S, S, ctypes.byref(opts), ctypes.c_float(2.0), ctypes.c_void_p(e["parts"].data_ptr()), None,
                  ctypes.c_void_p(e["g0"].data_ptr()), ctypes.c_void_p(e["g1"].data_ptr()), ctypes.c_void_p(ws.data_ptr()),
                  ws.numel(), ctypes.c_void_p(st.cuda_stream))

    graphs = [[] for _ in range(NS)]
    for si in range(NS):
        with torch.cuda.stream(streams[si]):
            step(sets[0], wss[si], streams[si])
            torch.cuda.synchronize()
            for e in sets:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=streams[si]):
                    step(e, wss[si], streams[si])
                graphs[si].append(g)
    torch.cuda.synchronize()
    N = 400

    def run(two):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(streams[0])
        for s_ in streams[1:]:
            s_.wait_event(e0)
        for k in range(N):
            si = (k % NS) if two else 0
            with torch.cuda.stream(streams[si]):
                graphs[si][k % n_sets].replay()
        for s_ in streams[1:]:
            streams[0].wait_stream(s_)
        e1.record(streams[0])
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / N

    for _ in range(2):
        run(False); run(True)
    one, two = run(False), run(True)
    r = {"streams": NS, "pairs_per_call": B, "one_stream_us_per_call": one * 1e3, "two_streams_us_per_call": two * 1e3,
         "one_stream_pairs_s": B / (one * 1e-3), "two_streams_pairs_s": B / (two * 1e-3)}
    print(json.dumps(r), flush=True)
    out.append(r)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/overlap_probe_%d.json" % NS, "w"), indent=1)
