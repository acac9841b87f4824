#!/usr/bin/env python3
"""bench.py -- throughput of the 3D-consistency hot path (fwd + bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--pairs B] [--size S] [--depth rough|smooth] [--no-graph] [--no-sweep]

One "step" = one pass of the hot path over one batch of B synthetic RGB-D pairs: the consistency
loss (both warp directions, occlusion mask) AND the gradients w.r.t. both 4-channel images.
Default workload = BASELINE.json configs[1] (ffhq_stylegan_occlusion.yml: batch 32 at 128x128,
L1, occlusion on, lambda_geometric 3, poses from the yml's CameraParamPrior ranges).

Prints ONE JSON line (rank 0): metric/value (device-resident inputs, CUDA-event timed, max over
ranks), e2e (public API, host buffers, H2D + D2H inside the timed region), roofline of the dominant
kernel, cpu_baseline (NumPy port of the Chainer CPU path on this box's host cores) and clocks.
`--impl reference` times that CPU port alone on the same workload definition.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "warped RGB-D pairs/sec fwd+bwd at 128^2"
UNIT = "pairs/s"
L2_BYTES = 126 * 2 ** 20
LAMBDA_ROTATE = 2.0          # upstream gradient of the loss: updater.py:363 (lambda_rotate default 2)
LAMBDA_GEO = 3.0             # updater.py:238


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    """one fwd+bwd of the NumPy port (the Chainer CPU path's array work) on `pairs` pairs"""
    pairs, S, depth, seed, reps = args
    from oracle import numpy_port as npp
    from rgbd_gan_b200 import poses
    x, cam = poses.synthetic_batch(pairs, S, depth=depth, seed=seed)
    f = npp.LossFuncRotateNP(lambda_geometric=LAMBDA_GEO)
    t0 = time.perf_counter()
    for _ in range(reps):
        f.forward(x[:pairs], cam[:pairs], x[pairs:], cam[pairs:], occlusion_aware=True)
        f.backward(LAMBDA_ROTATE)
    return time.perf_counter() - t0


def cpu_port_throughput(S, depth, steps, warmup, sample_pairs=4, procs=None):
    """pairs/s of the NumPy port with one process per host core (NumPy's elementwise kernels are
    single-threaded, so data-parallel processes are how this path can use every core)."""
    import multiprocessing as mp
    procs = procs or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        if warmup:
            pool.map(_cpu_worker, [(sample_pairs, S, depth, 100 + i, warmup) for i in range(procs)])
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(sample_pairs, S, depth, i, steps) for i in range(procs)])
        dt = time.perf_counter() - t0
    return procs * sample_pairs * steps / dt, dt / steps, procs


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    S = a.size
    steps, warmup = max(1, min(a.steps, 600)), max(0, min(a.warmup, 5))     # one step = procs x 4 pairs, ~70 ms
    sample = 4
    value, s_per_step, procs = cpu_port_throughput(S, a.depth, steps, warmup, sample_pairs=sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, a.pairs),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": "oracle/numpy_port.py (op-by-op NumPy port of the Chainer CPU path; Chainer is not "
                                   "installable here), %d processes x %d pairs per step, %d steps, numpy %s"
                                   % (procs, sample, steps, np.__version__)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(a, pairs):
    return {"workload": "configs[1] ffhq_stylegan_occlusion.yml consistency loss fwd+bwd, occlusion mask on, L1, "
                        "lambda_geometric 3", "pairs_per_gpu": pairs, "size": a.size, "channels": 4,
            "depth": a.depth, "pose_ranges": "x 0.3054 / y 1.0472 rad (yml)", "upstream_grad": LAMBDA_ROTATE}


# ------------------------------------------------------------------------------------ clock sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:           # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"GpuIdle": 0x1, "ApplicationsClocksSetting": 0x2, "sw_power_cap": 0x4, "hw_slowdown": 0x8,
                 "SyncBoost": 0x10, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "DisplayClockSetting": 0x100}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:        # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit and k != "GpuIdle":
                        self.reasons.add(k)
            except Exception:            # noqa: BLE001
                pass
            self._stop_evt.wait(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from rgbd_gan_b200 import _lib, poses
    from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, S, C = a.pairs, a.size, 4
    HW = S * S
    hbm_peak, peak_src = peaks()

    # ---- synthetic input pool, larger than L2, rotated between steps (no L2 reuse across steps)
    bytes_per_set = 4 * B * C * HW * 4                       # 2 images in + 2 gradients out
    pool_n = max(2, -(-3 * L2_BYTES // bytes_per_set))       # >= 3 x L2
    pool_n = min(pool_n, 64)
    sets = []
    host_f = LossFuncRotate(None, lambda_geometric=LAMBDA_GEO)
    host_f.init_params(None, size=S)
    for s in range(min(pool_n, 8)):                          # 8 distinct contents are enough; more are copies
        x, cam = poses.synthetic_batch(B, S, depth=a.depth, seed=1000 * rank + s)
        M, c, Mi, ci = pose_algebra(host_f.K, host_f.inv_K, cam[:B], cam[B:])
        sets.append((x, cam, np.concatenate([M.reshape(-1), c.reshape(-1), Mi.reshape(-1), ci.reshape(-1)])))
    pool = []
    for s in range(pool_n):
        x, cam, pv = sets[s % len(sets)]
        xt = torch.from_numpy(x).to(dev)
        pool.append(dict(img=xt[:B].contiguous(), img_rot=xt[B:].contiguous(), poses=torch.from_numpy(pv).to(dev),
                         g_img=torch.empty((B, C, S, S), device=dev), g_rot=torch.empty((B, C, S, S), device=dev),
                         parts=torch.zeros(8, device=dev), red=torch.zeros(4, device=dev)))
    ws = torch.empty(lib.rgbd_consistency_workspace_bytes(B, C, S, S), dtype=torch.uint8, device=dev)
    peer = None
    if world > 1 and a.collective == "peer":
        from rgbd_gan_b200.distributed import PeerComm
        peer = PeerComm()                 # fused 16-byte all-reduce inside the loss kernel (NVLink peer memory)
    # N > 1: the loss exchange is deferred (joined once at the end of the timed region, inside it), so it
    # overlaps the stage-out of its step and the stage-in of the next; gradients are never deferred
    opts = _lib.LossOpts(_lib.NORM_L1, 1, float("nan"), float("nan"), LAMBDA_GEO, B * world,
                         peer.handle if peer is not None else None, 1 if peer is not None else 0, 0)
    stream = torch.cuda.Stream(device=dev)
    sp = ctypes.c_void_p(stream.cuda_stream)

    def ptrs(e):
        base = e["poses"].data_ptr()
        return [ctypes.c_void_p(e["img"].data_ptr()), ctypes.c_void_p(e["img_rot"].data_ptr())] + \
               [ctypes.c_void_p(base + 4 * o) for o in (0, 9 * B, 12 * B, 21 * B)]

    def step_fused(e):
        _lib.call("rgbd_consistency_fwd_bwd", *ptrs(e), B, C, S, S, ctypes.byref(opts), ctypes.c_float(LAMBDA_ROTATE),
                  ctypes.c_void_p(e["parts"].data_ptr()), None, ctypes.c_void_p(e["g_img"].data_ptr()),
                  ctypes.c_void_p(e["g_rot"].data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(), sp)

    def step_two_pass(e):
        _lib.call("rgbd_consistency_fwd", *ptrs(e), B, C, S, S, ctypes.byref(opts),
                  ctypes.c_void_p(e["parts"].data_ptr()), None, None, ctypes.c_void_p(ws.data_ptr()), ws.numel(), sp)
        _lib.call("rgbd_consistency_bwd", *ptrs(e), B, C, S, S, ctypes.byref(opts), ctypes.c_float(LAMBDA_ROTATE), None,
                  None, ctypes.c_void_p(e["g_img"].data_ptr()), ctypes.c_void_p(e["g_rot"].data_ptr()),
                  ctypes.c_void_p(ws.data_ptr()), ws.numel(), sp)

    comm = torch.cuda.Stream(device=dev) if world > 1 else None

    def allreduce(e):
        """the only collective of the path: all-reduce(sum) of the 4 loss means.  The gradients do not
        depend on it (denominators are global by construction), so it runs on a side stream and overlaps
        the next step's kernels; the timed region ends with both streams drained."""
        if world == 1 or peer is not None:
            return
        ev = torch.cuda.Event()
        ev.record(stream)
        comm.wait_event(ev)
        with torch.cuda.stream(comm):
            e["red"].copy_(e["parts"][:4], non_blocking=True)
            dist.all_reduce(e["red"])
            e["comm_done"] = torch.cuda.Event()
            e["comm_done"].record(comm)

    def before_reuse(e):
        if world > 1 and e.get("comm_done") is not None:
            stream.wait_event(e["comm_done"])          # slot's previous all-reduce has consumed `parts`

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step, steps, warmup, graphs=None, sampler=None):
        """EXACTLY `steps` steps between barrier+sync, CUDA events on the launch stream, max over ranks"""
        with torch.cuda.stream(stream):
            for k in range(warmup):
                before_reuse(pool[k % pool_n])
                (graphs[k % pool_n].replay() if graphs else step(pool[k % pool_n]))
                allreduce(pool[k % pool_n])
            if comm is not None:
                stream.wait_stream(comm)
            barrier()
            if sampler:
                sampler.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for k in range(steps):
                before_reuse(pool[k % pool_n])
                (graphs[k % pool_n].replay() if graphs else step(pool[k % pool_n]))
                allreduce(pool[k % pool_n])
            if comm is not None:
                stream.wait_stream(comm)               # the last all-reduces are inside the timed region
            if peer is not None:
                _lib.check(lib.rgbd_peer_comm_wait(peer.handle, sp), "rgbd_peer_comm_wait")   # join the last exchange
            e1.record(stream)
            barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def capture(step):
        gs = []
        with torch.cuda.stream(stream):
            for e in pool:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    step(e)
                gs.append(g)
        return gs

    # ---- launches per step (counted by the library) and clock pre-warm
    with torch.cuda.stream(stream):
        n0 = lib.rgbd_launch_count()
        step_fused(pool[0])
        launches_fused = lib.rgbd_launch_count() - n0
        n0 = lib.rgbd_launch_count()
        step_two_pass(pool[0])
        launches_two = lib.rgbd_launch_count() - n0
        torch.cuda.synchronize(dev)
        for k in range(4000):                        # bring clocks up before anything is timed (fixed count:
            step_fused(pool[k % pool_n])             # every rank must make the same sequence of loss calls)
        torch.cuda.synchronize(dev)

    # direct launches (PDL chain) are the default: measured 27.9 us/step vs 28.9 us/step for CUDA-graph replays
    # of the same calls (profiles/r01_tuning.md); --graph switches to graph replays (not with a deferred exchange)
    use_graph = a.graph and peer is None
    graphs_fused = capture(step_fused) if use_graph else None
    graphs_two = capture(step_two_pass) if use_graph else None

    sampler = ClockSampler(local)
    ms = timed(step_fused, a.steps, a.warmup, graphs_fused, sampler)
    clocks = sampler.stop()
    ms_two = timed(step_two_pass, a.steps, a.warmup, graphs_two)
    value = world * B * a.steps / (ms * 1e-3)
    value_two = world * B * a.steps / (ms_two * 1e-3)

    # ---- roofline of the dominant kernel: k_consistency<4,LOSS,GRAD> timed live with CUDA events
    # recorded by the library around that launch (rgbd_profile_hook), same steps, same pool rotation
    kern_ms = []
    with torch.cuda.stream(stream):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        for e0, e1 in evs:               # torch creates the cudaEvent lazily: record once to get a handle
            e0.record(stream); e1.record(stream)
        for k in range(a.steps):
            lib.rgbd_profile_hook(ctypes.c_void_p(evs[k][0].cuda_event), ctypes.c_void_p(evs[k][1].cuda_event))
            step_fused(pool[k % pool_n])
        torch.cuda.synchronize(dev)
        kern_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    kern_avg_ms = sum(kern_ms) / len(kern_ms)
    with torch.cuda.stream(stream):      # same for the loss-only main kernel of the two-pass forward
        for k in range(a.steps):
            lib.rgbd_profile_hook(ctypes.c_void_p(evs[k][0].cuda_event), ctypes.c_void_p(evs[k][1].cuda_event))
            e = pool[k % pool_n]
            _lib.call("rgbd_consistency_fwd", *ptrs(e), B, C, S, S, ctypes.byref(opts),
                      ctypes.c_void_p(e["parts"].data_ptr()), None, None, ctypes.c_void_p(ws.data_ptr()), ws.numel(), sp)
        torch.cuda.synchronize(dev)
        kern_fwd_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / len(evs)
    alg_kernel = 16 * C * HW * B                     # reads both images once + writes both gradients once
    alg_step = 24 * C * HW * B                       # two-pass fwd+bwd definition, per pair 24*C*HW
    achieved = alg_kernel / (kern_avg_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_consistency<4,LOSS,GRAD> (project+gather+residual+scatter)",
                "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "peak_source": peak_src, "kernel_ms": kern_avg_ms, "kernel_ms_loss_only_variant": kern_fwd_ms,
                "kernel_share_of_step": kern_avg_ms * a.steps / ms,
                "algorithmic_bytes_per_launch": alg_kernel,
                "step": {"algorithmic_bytes": alg_step, "achieved": world * alg_step * a.steps / (ms * 1e-3) / 1e9 / world,
                         "frac": alg_step * a.steps / (ms * 1e-3) / 1e9 / hbm_peak,
                         "note": "whole fwd+bwd step per GPU, 24*C*H*W bytes per pair (SURVEY 8d)"}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("k_consistency_bytes_per_launch")
        except Exception:            # noqa: BLE001
            pass

    # ---- e2e: public API (LossFuncRotate mirror + autograd), host buffers, H2D/D2H inside the timed region
    f = LossFuncRotate(None, lambda_geometric=LAMBDA_GEO, grad_scale=LAMBDA_ROTATE, return_new_zp=False,
                       process_group=dist.group.WORLD if (world > 1 and peer is None) else None, peer_comm=peer)
    host_sets = []
    for s in range(len(sets)):
        x, cam, _ = sets[s]
        host_sets.append((torch.from_numpy(x).pin_memory(), cam))
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()

    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = [torch.empty((2 * B, C, S, S), device=dev) for _ in range(2)]     # double-buffered device inputs
    in_ready = [None, None]

    def h2d(k):
        """H2D of step k's inputs (pinned -> device) on the copy stream: overlaps step k-1's kernels"""
        xh, _ = host_sets[k % len(host_sets)]
        with torch.cuda.stream(copy_stream):
            dev_in[k % 2].copy_(xh, non_blocking=True)
            in_ready[k % 2] = torch.cuda.Event()
            in_ready[k % 2].record(copy_stream)

    def e2e_step(k, last):
        _, cam = host_sets[k % len(host_sets)]
        torch.cuda.current_stream().wait_event(in_ready[k % 2])
        if not last:
            h2d(k + 1)                                                   # next step's copy is in flight during this step
        xd = dev_in[k % 2]
        img = xd[:B].detach().requires_grad_(True)
        img_rot = xd[B:].detach().requires_grad_(True)
        loss, _ = f(img, cam[:B], img_rot, cam[B:], occlusion_aware=True)
        (loss * LAMBDA_ROTATE).backward()
        loss_host.copy_(loss.detach(), non_blocking=False)               # D2H read of the step's result (blocks)
        return img.grad

    e2e_steps = min(a.steps, 100)
    nw = max(3, min(a.warmup, 5))
    h2d(0)
    for k in range(nw):
        e2e_step(k, False)
    torch.cuda.synchronize(dev)
    barrier()
    t0 = time.perf_counter()
    h2d(nw)                      # (the copy issued by the last warm-up step is repeated inside the timed region)
    for k in range(nw, nw + e2e_steps):
        e2e_step(k, k == nw + e2e_steps - 1)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": world * B * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * B * C * HW * 4 + 24 * B * 4,
           "d2h_bytes_per_step": 4, "steps": e2e_steps,
           "pipeline": "step k+1's H2D (copy stream, double-buffered) overlaps step k's kernels; loss read back every step",
           "api": "rgbd_gan_b200.loss_functions.LossFuncRotate(grad_scale=lambda_rotate, return_new_zp=False) + backward()"}

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(a, B), l2="input pool of %d sets x %.0f MB rotated between steps (> L2)"
                           % (pool_n, bytes_per_set / 2 ** 20), cuda_graph=bool(use_graph),
                           path="rgbd_consistency_fwd_bwd (one-pass fwd+bwd, upstream grad = lambda_rotate)",
                           parallelism="dp%d (pairs sharded; 4-float loss all-reduce: %s)" % (
                               world, "none" if world == 1 else ("own finalize kernel exchanges them over NVLink peer memory on a side stream, joined at the end of the timed region"
                                                                 if peer is not None else "NCCL on a side stream"))),
            "two_pass": {"value": value_two, "ms_per_step": ms_two / a.steps,
                         "path": "rgbd_consistency_fwd + rgbd_consistency_bwd (recompute)"},
            "e2e": e2e, "gpu_launches": int(launches_fused * a.steps),
            "launches_per_step": {"fwd_bwd": int(launches_fused), "two_pass": int(launches_two)},
            "roofline": roofline, "clocks": clocks,
        }
    if peer is not None:
        peer.close()
    return line, dict(dev=dev, world=world, rank=rank, lib=lib, hbm_peak=hbm_peak)


def sweep(a, ctx_, sizes=((128, 256), (256, 256)), graph_like=False, hinge=None):
    """cfg4 of BASELINE.json: batch 256 at 128^2 and 256^2 (single GPU): pairs/s and step-level roofline.
    Also used for the generator-like smooth-depth variant of the headline workload."""
    import torch
    from rgbd_gan_b200 import _lib, poses
    from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra
    dev, lib = ctx_["dev"], ctx_["lib"]
    out = []
    for S, B in sizes:
        C, HW = 4, S * S
        x, cam = poses.synthetic_batch(B, S, depth=a.depth, seed=7)
        hf = LossFuncRotate(None, lambda_geometric=LAMBDA_GEO)
        hf.init_params(None, size=S)
        M, c, Mi, ci = pose_algebra(hf.K, hf.inv_K, cam[:B], cam[B:])
        pv = torch.from_numpy(np.concatenate([M.reshape(-1), c.reshape(-1), Mi.reshape(-1), ci.reshape(-1)])).to(dev)
        xt = torch.from_numpy(x).to(dev)
        n_sets = max(1, min(16, -(-3 * L2_BYTES // (4 * B * C * HW * 4))))      # rotate > 3 x L2 of in+out data
        imgs = [(xt[:B].clone(), xt[B:].clone()) for _ in range(n_sets)]
        g_img, g_rot = torch.empty((B, C, S, S), device=dev), torch.empty((B, C, S, S), device=dev)
        parts = torch.zeros(8, device=dev)
        ws = torch.empty(lib.rgbd_consistency_workspace_bytes(B, C, S, S), dtype=torch.uint8, device=dev)
        opts = _lib.LossOpts(_lib.NORM_L1, 1, float("nan"), float("nan"), LAMBDA_GEO, B, None)
        if hinge is not None:
            opts.hinge_depth_min, opts.hinge_lambda = hinge
        base = pv.data_ptr()
        pp = [ctypes.c_void_p(base + 4 * o) for o in (0, 9 * B, 12 * B, 21 * B)]

        def step(k):
            im, ir = imgs[k % n_sets]
            _lib.call("rgbd_consistency_fwd_bwd", ctypes.c_void_p(im.data_ptr()), ctypes.c_void_p(ir.data_ptr()), *pp,
                      B, C, S, S, ctypes.byref(opts), ctypes.c_float(LAMBDA_ROTATE), ctypes.c_void_p(parts.data_ptr()),
                      None, ctypes.c_void_p(g_img.data_ptr()), ctypes.c_void_p(g_rot.data_ptr()),
                      ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        for k in range(3):
            step(k)
        torch.cuda.synchronize(dev)
        n = 20 if B >= 128 else 200
        for k in range(n_sets):
            step(k)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n):
            step(k)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / n
        gbs = 24 * C * HW * B / (ms * 1e-3) / 1e9
        out.append({"pairs": B, "size": S, "depth": a.depth, "pairs_per_s": B / (ms * 1e-3), "ms_per_step": ms,
                    "step_algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / ctx_["hbm_peak"]})
        del imgs, g_img, g_rot, ws, xt
        torch.cuda.empty_cache()
    return out


def deepvoxels_bench(ctx_):
    """cfg3 of BASELINE.json: DeepVoxels projection sampling fwd (frustum gather) + bwd (lift scatter).
    Production geometry (deepvoxels_generator.py:229-253: G=32, F=32, 64x64x56, yml batch 10 -> run 16) and
    BASELINE's 64^3 volume.  Algorithmic bytes per sample fwd+bwd = 2*(F*G^3 + F*D*H*W)*4."""
    import torch
    from rgbd_gan_b200 import _lib, poses
    dev, lib = ctx_["dev"], ctx_["lib"]
    out = []
    for G, B in ((32, 16), (64, 16)):
        F, img = 32, 64
        D = int(np.ceil(np.sqrt(3) * G))
        vs = (1. / G) * 1.1 * 0.5
        P = _lib.DvParams(img, img, D, G, 128., 128., 32., 32., float(np.float32(vs)), float(np.float32(np.sqrt(3) / 4)))
        np.random.seed(3)
        thetas = poses.CameraParamPrior.from_ranges(poses.CAR_RANGES, True).sample(2 * B)[:B]
        cam = torch.from_numpy(poses.get_camera_matries(thetas).reshape(B, 16)).to(dev)
        n = img * img * D
        n_sets = 3 if G == 32 else 2                     # rotate buffers: > L2 in total
        grids = [torch.randn((B, F, G, G, G), device=dev) for _ in range(n_sets)]
        frs = [torch.empty((B, F, n), device=dev) for _ in range(n_sets)]
        ggs = [torch.empty((B, F, G ** 3), device=dev) for _ in range(n_sets)]
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        dws = torch.empty(lib.rgbd_dv_project_workspace_bytes(ctypes.byref(P), B, F), dtype=torch.uint8, device=dev)
        wsp = ctypes.c_void_p(dws.data_ptr())

        def fwd(k):
            _lib.call("rgbd_dv_project_fwd", ctypes.byref(P), ctypes.c_void_p(grids[k % n_sets].data_ptr()),
                      ctypes.c_void_p(cam.data_ptr()), B, F, ctypes.c_void_p(frs[k % n_sets].data_ptr()), wsp,
                      dws.numel(), st)

        def bwd(k):
            _lib.call("rgbd_dv_project_bwd", ctypes.byref(P), ctypes.c_void_p(frs[k % n_sets].data_ptr()),
                      ctypes.c_void_p(cam.data_ptr()), B, F, ctypes.c_void_p(ggs[k % n_sets].data_ptr()), wsp,
                      dws.numel(), st)

        def timeit(fn, reps=10):
            for k in range(3):
                fn(k)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(reps):
                fn(k)
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / reps

        ms_f, ms_b = timeit(fwd), timeit(bwd)
        alg = 2 * (F * G ** 3 + F * n) * 4 * B
        gbs = alg / ((ms_f + ms_b) * 1e-3) / 1e9
        out.append({"G": G, "F": F, "frustum": [D, img, img], "batch": B, "fwd_ms": ms_f, "bwd_ms": ms_b,
                    "samples_per_s": B / ((ms_f + ms_b) * 1e-3), "algorithmic_GBps": gbs,
                    "frac_of_hbm_peak": gbs / ctx_["hbm_peak"]})
        del grids, frs, ggs
        torch.cuda.empty_cache()
    return out


def render_bench(ctx_):
    """next row: fused projection + accumulative render tail (rgbd_dv_render_fwd/_bwd), which
    replaces project + occlusion MLP + cumsum/clip/diff + collapse + depth map without materialising the (B,F,D,H,W)
    view volume.  Algorithmic bytes per sample fwd+bwd = 2*(F*G^3 + (F+2)*H*W)*4."""
    import torch
    from rgbd_gan_b200 import _lib, poses
    dev, lib = ctx_["dev"], ctx_["lib"]
    out = []
    for G, B in ((32, 16), (64, 16)):
        F, img, nf = 32, 64, 4
        D = int(np.ceil(np.sqrt(3) * G))
        vs = (1. / G) * 1.1 * 0.5
        P = _lib.DvParams(img, img, D, G, 128., 128., 32., 32., float(np.float32(vs)), float(np.float32(np.sqrt(3) / 4)))
        R = _lib.DvRenderParams(nf, D, 4.0, float(np.float32(np.sqrt(2) * np.sqrt(1.0 / (F + 1)))),
                                float(np.float32(np.sqrt(2) * np.sqrt(1.0 / nf))))
        np.random.seed(3)
        thetas = poses.CameraParamPrior.from_ranges(poses.CAR_RANGES, True).sample(2 * B)[:B]
        cam = torch.from_numpy(poses.get_camera_matries(thetas).reshape(B, 16)).to(dev)
        rng = np.random.default_rng(3)
        W1 = torch.from_numpy(rng.normal(size=(nf, F + 1)).astype(np.float32)).to(dev)
        b1 = torch.from_numpy(rng.normal(scale=0.3, size=(nf,)).astype(np.float32)).to(dev)
        W2 = torch.from_numpy(rng.normal(size=(nf,)).astype(np.float32)).to(dev)
        b2 = torch.tensor([-0.5], device=dev)
        n_sets = 4 if G == 32 else 2
        grids = [torch.randn((B, F, G, G, G), device=dev) for _ in range(n_sets)]
        ggs = [torch.empty((B, F, G ** 3), device=dev) for _ in range(n_sets)]
        novel, depth, fg = torch.empty((B, F, img * img), device=dev), torch.empty((B, img * img), device=dev), torch.empty((B, img * img), device=dev)
        g_novel, g_depth = torch.randn_like(novel), torch.randn_like(depth)
        gW1, gb1, gW2, gb2 = torch.empty_like(W1), torch.empty_like(b1), torch.empty_like(W2), torch.empty_like(b2)
        ws = torch.empty(lib.rgbd_dv_render_workspace_bytes(ctypes.byref(P), B, F), dtype=torch.uint8, device=dev)
        saved = torch.empty(lib.rgbd_dv_render_saved_bytes(ctypes.byref(P), B), dtype=torch.uint8, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        pw = [ctypes.c_void_p(t.data_ptr()) for t in (W1, b1, W2, b2)]

        def fwd(k):
            _lib.call("rgbd_dv_render_fwd", ctypes.byref(P), ctypes.byref(R), ctypes.c_void_p(grids[k % n_sets].data_ptr()),
                      ctypes.c_void_p(cam.data_ptr()), *pw, B, F, ctypes.c_void_p(novel.data_ptr()),
                      ctypes.c_void_p(depth.data_ptr()), ctypes.c_void_p(fg.data_ptr()), ctypes.c_void_p(saved.data_ptr()),
                      ctypes.c_void_p(ws.data_ptr()), ws.numel(), st)

        def bwd(k):
            _lib.call("rgbd_dv_render_bwd", ctypes.byref(P), ctypes.byref(R), ctypes.c_void_p(grids[k % n_sets].data_ptr()),
                      ctypes.c_void_p(cam.data_ptr()), *pw, B, F, ctypes.c_void_p(saved.data_ptr()),
                      ctypes.c_void_p(g_novel.data_ptr()),
                      ctypes.c_void_p(g_depth.data_ptr()), None, ctypes.c_void_p(ggs[k % n_sets].data_ptr()),
                      *[ctypes.c_void_p(t.data_ptr()) for t in (gW1, gb1, gW2, gb2)], ctypes.c_void_p(ws.data_ptr()),
                      ws.numel(), st)

        def timeit(fn, reps=10):
            for k in range(3):
                fn(k)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(reps):
                fn(k)
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / reps

        # (bwd(k) reads the running sums left by fwd(k): time the pair, then the forward alone)
        ms_f = timeit(fwd)
        ms_fb = timeit(lambda k: (fwd(k), bwd(k)))
        ms_b = ms_fb - ms_f
        alg = 2 * (F * G ** 3 + (F + 2) * img * img) * 4 * B
        gbs = alg / ((ms_f + ms_b) * 1e-3) / 1e9
        row = {"G": G, "F": F, "frustum": [D, img, img], "batch": B, "fwd_ms": ms_f, "bwd_ms": ms_b,
               "samples_per_s": B / ((ms_f + ms_b) * 1e-3), "algorithmic_GBps": gbs,
               "frac_of_hbm_peak": gbs / ctx_["hbm_peak"], "saturated_rays": float((fg > 0.999999).float().mean().item())}
        # context: the same math as the reference's op chain (deepvoxel.py:574-587,888,903-904), UNFUSED, on this GPU:
        # our projection kernel for the view volume, then library (PyTorch) elementwise / einsum / cumsum ops + autograd
        try:
            from rgbd_gan_b200.projection import _ProjectFn
            dc = torch.from_numpy((np.arange(-D // 2, D // 2) / D).astype(np.float32)).to(dev)

            def unfused(k, backward):
                g = grids[k % n_sets].detach().requires_grad_(backward)
                vol = _ProjectFn.apply(g, cam, P)
                x = torch.cat([dc.view(1, 1, D, 1, 1).expand(B, 1, D, img, img), vol], 1) * R.inv_c1
                a = torch.einsum("jc,bcdhw->bjdhw", W1, x) + b1.view(1, -1, 1, 1, 1)
                h = torch.nn.functional.leaky_relu(a, 0.2)
                sg = torch.sigmoid(torch.einsum("j,bjdhw->bdhw", W2, R.inv_c2 * h) + b2 - R.threshold)
                cl = torch.clamp(torch.cumsum(sg, 1), 0, 1)
                w = torch.diff(torch.cat([torch.zeros_like(cl[:, :1]), cl], 1), dim=1)
                dm = ((dc.view(1, D, 1, 1) * w).sum(1) + 0.5) * D * P.voxel_size + P.near_plane
                nv = (w[:, None] * vol).sum(2)
                if backward:
                    ((nv * g_novel.view_as(nv)).sum() + (dm * g_depth.view_as(dm)).sum()).backward()

            row["unfused_torch_fwd_ms"] = timeit(lambda k: unfused(k, False), reps=3)
            row["unfused_torch_fwd_bwd_ms"] = timeit(lambda k: unfused(k, True), reps=3)
            row["fused_speedup_fwd_bwd"] = row["unfused_torch_fwd_bwd_ms"] / (ms_f + ms_b)
        except Exception as e:           # noqa: BLE001  (context only; never fails the bench)
            row["unfused_torch_error"] = repr(e)[:200]
        out.append(row)
        del grids, ggs
        torch.cuda.empty_cache()
    return out


def feature_consistency_bench(a, ctx_):
    """the feature-space consistency loss of updater.py:345-354 (norm l2, C = 256 features +
    1 depth at 32x32, yml batch 32 -> 16 pairs); runs through the generic-C kernels (not tuned)."""
    import torch
    from rgbd_gan_b200 import _lib, poses
    from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra
    dev, lib = ctx_["dev"], ctx_["lib"]
    B, C, S = 16, 257, 32
    HW = S * S
    rng = np.random.default_rng(0)
    _, cam = poses.synthetic_batch(B, S, depth="rough", seed=3)
    x = rng.uniform(-1, 1, size=(2 * B, C, S, S)).astype(np.float32)
    x[:, -1] = rng.uniform(0.7, 1.5, size=(2 * B, S, S))
    hf = LossFuncRotate(None, lambda_geometric=LAMBDA_GEO)
    hf.init_params(None, size=S)
    M, c, Mi, ci = pose_algebra(hf.K, hf.inv_K, cam[:B], cam[B:])
    pv = torch.from_numpy(np.concatenate([M.reshape(-1), c.reshape(-1), Mi.reshape(-1), ci.reshape(-1)])).to(dev)
    xt = torch.from_numpy(x).to(dev)
    img, rot = xt[:B].contiguous(), xt[B:].contiguous()
    g0, g1, parts = torch.empty_like(img), torch.empty_like(rot), torch.zeros(8, device=dev)
    ws = torch.empty(lib.rgbd_consistency_workspace_bytes(B, C, S, S), dtype=torch.uint8, device=dev)
    opts = _lib.LossOpts(_lib.NORM_L2, 1, float("nan"), float("nan"), LAMBDA_GEO, B, None)
    pp = [ctypes.c_void_p(pv.data_ptr() + 4 * o) for o in (0, 9 * B, 12 * B, 21 * B)]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def step():
        _lib.call("rgbd_consistency_fwd_bwd", ctypes.c_void_p(img.data_ptr()), ctypes.c_void_p(rot.data_ptr()), *pp, B, C, S, S,
                  ctypes.byref(opts), ctypes.c_float(LAMBDA_ROTATE), ctypes.c_void_p(parts.data_ptr()), None,
                  ctypes.c_void_p(g0.data_ptr()), ctypes.c_void_p(g1.data_ptr()), ctypes.c_void_p(ws.data_ptr()), ws.numel(), st)
    for _ in range(5):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 100
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / n
    gbs = 24 * C * HW * B / (ms * 1e-3) / 1e9
    return {"pairs": B, "channels": C, "size": S, "norm": "l2", "ms_per_step": ms, "pairs_per_s": B / (ms * 1e-3),
            "step_algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / ctx_["hbm_peak"],
            "note": "L2-resident (34 MB of inputs, re-used every step): context only"}


def cpu_baseline_leg(a):
    """rank 0, N=1 only: bounded sample (~10-20 s) of the same workload on the host cores"""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    t0 = time.perf_counter()
    v1 = 4 * 6 / _cpu_worker((4, a.size, a.depth, 0, 6))                       # single process, 1 core
    value, _, procs = cpu_port_throughput(a.size, a.depth, steps=100, warmup=1, sample_pairs=4)
    return {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "single_core_value": v1,
            "sample": "oracle/numpy_port.py fwd+bwd (NumPy port of the Chainer CPU path), %d processes x 4 pairs x 100 "
                      "steps at %dx%d, occlusion on; wall %.1f s" % (procs, a.size, a.size, time.perf_counter() - t0)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--depth", default="rough", choices=["rough", "smooth"])
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="(default; kept for compatibility)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.impl == "reference":
        run_reference(a)
        return
    # stdout carries exactly ONE line (the JSON record): while the benchmark runs, file descriptor 1 points at stderr,
    # so banners that libraries write to stdout (e.g. "NCCL version ..." at communicator creation) end up there
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(record):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(record), flush=True)
        os.dup2(2, 1)

    line, ctx_ = run_ours(a)
    if ctx_["world"] > 1:
        import torch.distributed as dist
        if line is not None:
            emit(line)
        dist.barrier()
        dist.destroy_process_group()
        return
    if not a.no_sweep:
        line["sweep"] = sweep(a, ctx_)
        if a.depth == "rough":
            a2 = argparse.Namespace(**vars(a))
            a2.depth = "smooth"
            line["smooth_depth"] = sweep(a2, ctx_, sizes=((a.size, a.pairs),), graph_like=False)[0]
        # next row: the updaters' depth hinge (yml: depth_min 1.0, lambda_depth 10) fused in
        line["with_depth_hinge"] = sweep(a, ctx_, sizes=((a.size, a.pairs),), hinge=(1.0, 10.0))[0]
        line["deepvoxels"] = deepvoxels_bench(ctx_)
        line["deepvoxels_render_fused"] = render_bench(ctx_)
        line["feature_consistency_c257"] = feature_consistency_bench(a, ctx_)
    if not a.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg(a)
    emit(line)


if __name__ == "__main__":
    main()
