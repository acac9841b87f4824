"""One B-pair call split into N sub-batches on N streams with a fork at the start and a join at the end of EVERY
call (what a multi-stream chunk schedule inside rgbd_consistency_fwd_bwd would do), against the unsplit call."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rgbd_gan_b200 import _lib
from tools import synthetic as poses
from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra

dev = torch.device("cuda", 0)
lib = _lib.load()
S, C = 128, 4
HW = S * S
hf = LossFuncRotate(None, lambda_geometric=3.0)
hf.init_params(None, size=S)
out = []
for Btot, N in ((32, 1), (32, 2), (32, 3), (32, 4), (256, 1), (256, 2), (256, 3)):
    subs = [Btot // N + (1 if i < Btot % N else 0) for i in range(N)]
    n_sets = max(3, (3 * 126 * 2 ** 20) // (4 * Btot * C * HW * 4) + 1)
    streams = [torch.cuda.Stream(device=dev) for _ in range(N)]
    data = []
    for s in range(n_sets):
        per = []
        for i, B in enumerate(subs):
            x, cam = poses.synthetic_batch(B, S, depth="rough", seed=(s * 7 + i) % 5)
            M, c, Mi, ci = pose_algebra(hf.K, hf.inv_K, cam[:B], cam[B:])
            pv = torch.from_numpy(np.concatenate([M.reshape(-1), c.reshape(-1), Mi.reshape(-1), ci.reshape(-1)])).to(dev)
            xt = torch.from_numpy(x).to(dev)
            per.append(dict(B=B, img=xt[:B].contiguous(), rot=xt[B:].contiguous(), pv=pv, g0=torch.empty((B, C, S, S), device=dev),
                            g1=torch.empty((B, C, S, S), device=dev), parts=torch.zeros(8, device=dev)))
        data.append(per)
    wss = [torch.empty(lib.rgbd_consistency_workspace_bytes(B, C, S, S), dtype=torch.uint8, device=dev) for B in subs]
    optss = [_lib.LossOpts(_lib.NORM_L1, 1, float("nan"), float("nan"), 3.0, Btot, None) for _ in subs]

    def sub_call(e, i):
        B = e["B"]
        base = e["pv"].data_ptr()
        pp = [ctypes.c_void_p(base + 4 * o) for o in (0, 9 * B, 12 * B, 21 * B)]
        _lib.call("rgbd_consistency_fwd_bwd", ctypes.c_void_p(e["img"].data_ptr()), ctypes.c_void_p(e["rot"].data_ptr()), *pp,
                  B, C, S, S, ctypes.byref(optss[i]), ctypes.c_float(2.0), ctypes.c_void_p(e["parts"].data_ptr()), None,
                  ctypes.c_void_p(e["g0"].data_ptr()), ctypes.c_void_p(e["g1"].data_ptr()), ctypes.c_void_p(wss[i].data_ptr()),
                  wss[i].numel(), ctypes.c_void_p(streams[i].cuda_stream))

    def call(per):
        ev = torch.cuda.Event()
        ev.record(streams[0])
        for i in range(1, N):
            streams[i].wait_event(ev)
        for i in range(N):
            sub_call(per[i], i)
        for i in range(1, N):
            e2 = torch.cuda.Event()
            e2.record(streams[i])
            streams[0].wait_event(e2)

    graphs = []
    with torch.cuda.stream(streams[0]):
        call(data[0])
        torch.cuda.synchronize()
        for per in data:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=streams[0]):
                call(per)
            graphs.append(g)
    torch.cuda.synchronize()
    res = {}
    for mode in ("graph", "direct"):
        reps = 300 if Btot <= 32 else 40
        with torch.cuda.stream(streams[0]):
            for k in range(20):
                graphs[k % n_sets].replay() if mode == "graph" else call(data[k % n_sets])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(streams[0])
            for k in range(reps):
                graphs[k % n_sets].replay() if mode == "graph" else call(data[k % n_sets])
            e1.record(streams[0])
            torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / reps * 1e3
    r = {"pairs": Btot, "streams": N, "us_per_call_graph": res["graph"], "us_per_call_direct": res["direct"],
         "pairs_per_s_best": Btot / (min(res.values()) * 1e-6)}
    print(json.dumps(r), flush=True)
    out.append(r)
    del data, wss, graphs
    torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/split_probe.json", "w"), indent=1)
