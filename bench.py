#!/usr/bin/env python3
"""bench.py -- throughput of the 3D-consistency hot path (fwd + bwd) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--pairs B] [--size S] [--depth rough|smooth] [--no-graph] [--no-sweep]

One "step" = one pass of the hot path over one batch of B synthetic RGB-D pairs: the consistency
loss (both warp directions, occlusion mask) AND the gradients w.r.t. both 4-channel images.
Default workload = BASELINE.json configs[4] at the metric's size: 256 pairs per GPU at 128x128 (L1,
occlusion on, lambda_geometric 3, poses from ffhq_stylegan_occlusion.yml's CameraParamPrior ranges);
configs[1] (32 pairs), configs[4] at 256x256 and configs[2] (car poses, 64 pairs in total) are
timed the same way at every N and reported under "extra".

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
    from tools import synthetic
    x, cam = synthetic.synthetic_batch(pairs, S, depth=depth, seed=seed)
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
        "run": {"sample_pairs_per_step": procs * sample,
                "note": "each step is a bounded sample of the workload: %d processes x %d pairs of the same "
                        "distribution (the NumPy port's cost per pair does not depend on the batch size)" % (procs, sample)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": "oracle/numpy_port.py (op-by-op NumPy port of the Chainer CPU path; Chainer is not "
                                   "installable here), %d processes x %d pairs per step, %d steps, numpy %s"
                                   % (procs, sample, steps, np.__version__)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


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
FFHQ, CAR = "ffhq", "car"


class Harness:
    """process-wide state of the GPU arm: device, stream, library, the loss collective (peer mailbox or NCCL)"""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        from rgbd_gan_b200 import _lib
        self.torch, self.dist, self._lib = torch, dist, _lib
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = _lib.load()
        self.hbm_peak, self.peak_src = peaks()
        self.peer = None
        self.exchange = getattr(a, "exchange", "lazy")
        if self.world > 1 and a.collective == "peer":
            from rgbd_gan_b200.distributed import PeerComm
            self.peer = PeerComm()            # fused 20-byte all-reduce inside the finalize kernel (NVLink peer memory)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.sp = ctypes.c_void_p(self.stream.cuda_stream)
        self.comm = torch.cuda.Stream(device=self.dev) if (self.world > 1 and self.peer is None) else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, v):
        if self.world > 1:
            t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            v = float(t.item())
        return v


class Workload:
    """B pairs per GPU at SxS: a pool of input sets larger than L2 (rotated between steps), outputs, workspace, options"""

    def __init__(self, h, B, S, depth="rough", ranges=FFHQ, lam=LAMBDA_GEO, n_global=None, hinge=None, K=None,
                 max_depth=None, min_depth=None, label=""):
        torch, _lib = h.torch, h._lib
        from tools import synthetic as poses
        from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra
        self.h, self.B, self.S, self.C, self.depth, self.lam, self.label = h, B, S, 4, depth, lam, label
        C, HW = 4, S * S
        self.bytes_per_set = 4 * B * C * HW * 4                  # 2 images in + 2 gradients out
        self.pool_n = min(max(2, -(-3 * L2_BYTES // self.bytes_per_set)), 64)       # >= 3 x L2 in rotation
        rg = poses.FFHQ_RANGES if ranges == FFHQ else poses.CAR_RANGES
        hf = LossFuncRotate(None, K=K, lambda_geometric=lam)
        hf.init_params(None, size=S)
        self.host_sets, self.pool = [], []
        for s in range(min(self.pool_n, 4)):                     # a few distinct contents; further sets are copies
            x, cam = poses.synthetic_batch(B, S, depth=depth, ranges=rg, seed=1000 * h.rank + s)
            M, c, Mi, ci = pose_algebra(hf.K, hf.inv_K, cam[:B], cam[B:])
            self.host_sets.append((x, cam, np.concatenate([M.reshape(-1), c.reshape(-1), Mi.reshape(-1), ci.reshape(-1)])))
        for s in range(self.pool_n):
            x, cam, pv = self.host_sets[s % len(self.host_sets)]
            xt = torch.from_numpy(x).to(h.dev)
            self.pool.append(dict(img=xt[:B].contiguous(), img_rot=xt[B:].contiguous(), poses=torch.from_numpy(pv).to(h.dev),
                                  g_img=torch.empty((B, C, S, S), device=h.dev), g_rot=torch.empty((B, C, S, S), device=h.dev),
                                  parts=torch.zeros(8, device=h.dev), red=torch.zeros(4, device=h.dev)))
            del xt
        self.ws = torch.empty(h.lib.rgbd_consistency_workspace_bytes(B, C, S, S), dtype=torch.uint8, device=h.dev)
        nan = float("nan")
        ng = n_global if n_global is not None else B * h.world
        # N > 1, --exchange: "lazy" (default) = every call PUBLISHES its loss parts into the peers' mailboxes from its own
        # finishing kernel and nobody waits; the sum is formed when the loss is read (rgbd_peer_comm_wait: at the end of
        # the timed region, inside it) -- a training loop only logs the loss (updater.py:361).  "deferred" = all-reduce
        # every call on a side stream, lag bounded to one call; "joined" = every call waits for its own all-reduce.
        # Gradients are never deferred
        defer = {"lazy": 2, "deferred": 1, "joined": 0}[h.exchange] if h.peer is not None else 0
        self.opts = _lib.LossOpts(_lib.NORM_L1, 1, nan if max_depth is None else max_depth,
                                  nan if min_depth is None else min_depth, lam, ng,
                                  h.peer.handle if h.peer is not None else None, defer, 0)
        self.opts_local = _lib.LossOpts(_lib.NORM_L1, 1, nan if max_depth is None else max_depth,
                                        nan if min_depth is None else min_depth, lam, ng, None, 0, 0)
        if hinge is not None:
            for o in (self.opts, self.opts_local):
                o.hinge_depth_min, o.hinge_lambda = hinge
        self.sweep_path = bool(h.lib.rgbd_consistency_uses_sweep(B, C, S, S))

    def ptrs(self, e):
        B = self.B
        base = e["poses"].data_ptr()
        return [ctypes.c_void_p(e["img"].data_ptr()), ctypes.c_void_p(e["img_rot"].data_ptr())] + \
               [ctypes.c_void_p(base + 4 * o) for o in (0, 9 * B, 12 * B, 21 * B)]

    def step_fused(self, e, opts=None, parts=None):
        h = self.h
        h._lib.call("rgbd_consistency_fwd_bwd", *self.ptrs(e), self.B, self.C, self.S, self.S,
                    ctypes.byref(opts or self.opts), ctypes.c_float(LAMBDA_ROTATE),
                    ctypes.c_void_p((parts if parts is not None else e["parts"]).data_ptr()), None,
                    ctypes.c_void_p(e["g_img"].data_ptr()), ctypes.c_void_p(e["g_rot"].data_ptr()),
                    ctypes.c_void_p(self.ws.data_ptr()), self.ws.numel(), h.sp)

    def step_two_pass(self, e):
        h = self.h
        h._lib.call("rgbd_consistency_fwd", *self.ptrs(e), self.B, self.C, self.S, self.S, ctypes.byref(self.opts),
                    ctypes.c_void_p(e["parts"].data_ptr()), None, None, ctypes.c_void_p(self.ws.data_ptr()), self.ws.numel(), h.sp)
        h._lib.call("rgbd_consistency_bwd", *self.ptrs(e), self.B, self.C, self.S, self.S, ctypes.byref(self.opts),
                    ctypes.c_float(LAMBDA_ROTATE), None, None, ctypes.c_void_p(e["g_img"].data_ptr()),
                    ctypes.c_void_p(e["g_rot"].data_ptr()), ctypes.c_void_p(self.ws.data_ptr()), self.ws.numel(), h.sp)

    def free(self):
        self.pool, self.ws = [], None
        self.h.torch.cuda.empty_cache()

    def join_exchange(self):
        h = self.h
        if h.peer is not None:
            h._lib.check(h.lib.rgbd_peer_comm_wait(h.peer.handle, h.sp), "rgbd_peer_comm_wait")

    def _nccl_allreduce(self, e):
        """--collective nccl: the only collective of the path, all-reduce(sum) of the 4 loss means, on a side stream"""
        h = self.h
        if h.comm is None:
            return
        torch = h.torch
        ev = torch.cuda.Event()
        ev.record(h.stream)
        h.comm.wait_event(ev)
        with torch.cuda.stream(h.comm):
            e["red"].copy_(e["parts"][:4], non_blocking=True)
            h.dist.all_reduce(e["red"])
            e["comm_done"] = torch.cuda.Event()
            e["comm_done"].record(h.comm)

    def timed(self, steps, warmup, step=None, sampler=None):
        """EXACTLY `steps` steps, CUDA events on the launch stream, max over ranks.  Sequence: `warmup` untimed steps,
        barrier + device sync, ALIGN_STEPS untimed steps the last of which has its loss exchange joined (at N > 1 every
        GPU leaves it within microseconds of the others, so the ranks enter the timed region aligned and the NCCL
        barrier's exit skew stays outside), e0, `steps` steps, join of the last exchange, e1, barrier + device sync."""
        h, torch = self.h, self.h.torch
        step = step or self.step_fused
        pool, n = self.pool, self.pool_n

        def one(k):
            e = pool[k % n]
            if h.comm is not None and e.get("comm_done") is not None:
                h.stream.wait_event(e["comm_done"])              # slot's previous all-reduce has consumed `parts`
            step(e)
            self._nccl_allreduce(e)
        with torch.cuda.stream(h.stream):
            for k in range(warmup):
                one(k)
            if h.comm is not None:
                h.stream.wait_stream(h.comm)
            self.join_exchange()
            h.barrier()
            # alignment steps (untimed, stream-ordered right in front of e0 with no host gap): the barrier, the device
            # sync and the start of the clock sampler leave the GPU idle for milliseconds; a 20-step timed region right
            # after that measured 4 % slower than the steady state (126.7 vs 121.9 us).  The last one is joined
            for k in range(ALIGN_STEPS):
                one(warmup + k)
            if h.comm is not None:
                h.stream.wait_stream(h.comm)
            self.join_exchange()
            if sampler:
                sampler.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(h.stream)
            for k in range(steps):
                one(warmup + ALIGN_STEPS + k)
            if h.comm is not None:
                h.stream.wait_stream(h.comm)                     # the last all-reduces are inside the timed region
            self.join_exchange()                                 # ... and so is the join of the last deferred exchange
            e1.record(h.stream)
            h.barrier()
        return h.max_over_ranks(e0.elapsed_time(e1))

    def record(self, steps, warmup, step=None, sampler=None):
        ms = self.timed(steps, warmup, step, sampler)
        h = self.h
        per_gpu = 24 * self.C * self.S * self.S * self.B * steps / (ms * 1e-3) / 1e9      # 24*C*H*W per pair
        return {"pairs_per_gpu": self.B, "size": self.S, "depth": self.depth, "workload": self.label,
                "value": h.world * self.B * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps,
                "step_algorithmic_GBps_per_gpu": per_gpu, "frac_of_hbm_peak": per_gpu / h.hbm_peak,
                "path": "row sweep (k_consistency_sweep + k_sweep_fixup)" if self.sweep_path
                        else "three-kernel chain (stage-in, k_consistency_fast, stage-out)"}

    def parity(self):
        """N > 1: every rank also evaluates its shard WITHOUT the collective; rank 0 checks that the exchanged loss parts
        are bit-for-bit the rank-ordered fp32 sum of the local parts (the exchange adds the mailbox slots in rank order)"""
        h, torch = self.h, self.h.torch
        if h.world == 1:
            return None
        e = self.pool[0]
        loc = torch.zeros(8, device=h.dev)
        with torch.cuda.stream(h.stream):
            self.step_fused(e)
            self.join_exchange()
            if h.comm is not None:
                self._nccl_allreduce(e)
                h.stream.wait_stream(h.comm)
            self.step_fused(e, opts=self.opts_local, parts=loc)
        torch.cuda.synchronize(h.dev)
        got = (e["red"] if h.comm is not None else e["parts"][:4]).clone()
        allloc = [torch.zeros(8, device=h.dev) for _ in range(h.world)]
        h.dist.all_gather(allloc, loc)
        if h.rank != 0:
            return None
        locs = np.stack([t.cpu().numpy() for t in allloc]).astype(np.float32)
        want = np.zeros(4, np.float32)
        for r in range(h.world):
            want = (want + locs[r, :4]).astype(np.float32)
        gotn = got.cpu().numpy().astype(np.float32)
        exact = bool(np.array_equal(gotn, want))
        close = bool(np.allclose(gotn, want, rtol=1e-6, atol=0))
        return {"ok": exact if h.peer is not None else close, "bit_exact": exact, "exchanged": [float(v) for v in gotn],
                "sum_of_local_parts": [float(v) for v in want], "ranks": h.world,
                "collective": "peer mailbox (rank-ordered fp32 sum)" if h.peer is not None else "NCCL all-reduce"}


def kernel_roofline(wl, steps, step_ms_total):
    """the dominant kernel timed live with CUDA events recorded by the library around that launch (rgbd_profile_hook)"""
    h, torch = wl.h, wl.h.torch
    B, C, HW = wl.B, wl.C, wl.S * wl.S
    n = min(steps, 200)
    with torch.cuda.stream(h.stream):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        for e0, e1 in evs:               # torch creates the cudaEvent lazily: record once to get a handle
            e0.record(h.stream); e1.record(h.stream)
        for k in range(n):
            h.lib.rgbd_profile_hook(ctypes.c_void_p(evs[k][0].cuda_event), ctypes.c_void_p(evs[k][1].cuda_event))
            wl.step_fused(wl.pool[k % wl.pool_n])
        wl.join_exchange()
        torch.cuda.synchronize(h.dev)
        kern_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / n
        for k in range(n):               # same for the loss-only kernel of the two-pass forward
            h.lib.rgbd_profile_hook(ctypes.c_void_p(evs[k][0].cuda_event), ctypes.c_void_p(evs[k][1].cuda_event))
            e = wl.pool[k % wl.pool_n]
            h._lib.call("rgbd_consistency_fwd", *wl.ptrs(e), B, C, wl.S, wl.S, ctypes.byref(wl.opts),
                        ctypes.c_void_p(e["parts"].data_ptr()), None, None, ctypes.c_void_p(wl.ws.data_ptr()), wl.ws.numel(), h.sp)
        wl.join_exchange()
        torch.cuda.synchronize(h.dev)
        kern_fwd_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs) / n
    alg_kernel = 16 * C * HW * B                     # reads both images once + writes both gradients once
    alg_step = 24 * C * HW * B                       # two-pass fwd+bwd definition, per pair 24*C*HW
    achieved = alg_kernel / (kern_ms * 1e-3) / 1e9
    name = ("k_consistency_sweep<128,L1,LOSS,GRAD,RING> (persistent TMA-fed row sweep: unproject+transform+reproject, "
            "LDS gather, occlusion-masked residual, ring scatter + write-back)") if wl.sweep_path else \
           "k_consistency_fast<LOSS,GRAD> (project+gather+residual+scatter on the staged NHWC copy)"
    r = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": h.hbm_peak, "unit": "GB/s",
         "frac": achieved / h.hbm_peak, "traffic": None, "peak_source": h.peak_src, "kernel_ms": kern_ms,
         "kernel_ms_loss_only_variant": kern_fwd_ms, "kernel_share_of_step": kern_ms * steps / step_ms_total,
         "algorithmic_bytes_per_launch": alg_kernel,
         "step": {"algorithmic_bytes": alg_step, "achieved": alg_step * steps / (step_ms_total * 1e-3) / 1e9,
                  "frac": alg_step * steps / (step_ms_total * 1e-3) / 1e9 / h.hbm_peak,
                  "note": "whole fwd+bwd step per GPU, 24*C*H*W bytes per pair (SURVEY 8d)"}}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            t = json.load(open(tr))
            key = "sweep_256pairs_128" if wl.sweep_path else "k_consistency_bytes_per_launch"
            ent = t.get(key)
            if isinstance(ent, dict):
                r["traffic"] = ent["dram_bytes_per_launch"] * B // ent["pairs"]
                r["traffic_source"] = "%s (static: one ncu --set full capture, dram__bytes_read+write per launch, scaled " \
                                      "from %d pairs; not measured live)" % (ent["source"], ent["pairs"])
            elif ent is not None:
                r["traffic"] = ent
                r["traffic_source"] = "profiles/traffic.json (static, round-1 ncu capture of the chain's main kernel)"
        except Exception:            # noqa: BLE001
            pass
    return r


def e2e_leg(h, wl, steps, warmup, with_grads):
    """public API (LossFuncRotate mirror + autograd) with HOST buffers: every step copies its inputs pinned-host -> device
    and reads the loss (and optionally both gradient tensors) back; wall clock, max over ranks"""
    torch, dist = h.torch, h.dist
    from rgbd_gan_b200.loss_functions import LossFuncRotate
    B, C, S = wl.B, wl.C, wl.S
    f = LossFuncRotate(None, lambda_geometric=wl.lam, grad_scale=LAMBDA_ROTATE, return_new_zp=False,
                       process_group=dist.group.WORLD if (h.world > 1 and h.peer is None) else None, peer_comm=h.peer)
    host_sets = [(torch.from_numpy(x).pin_memory(), cam) for x, cam, _ in wl.host_sets[:2]]
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    g_host = [torch.empty((B, C, S, S)).pin_memory() for _ in range(2)] if with_grads else None
    copy_stream = torch.cuda.Stream(device=h.dev)
    dev_in = [torch.empty((2 * B, C, S, S), device=h.dev) for _ in range(2)]     # double-buffered device inputs
    in_ready = [None, None]

    def h2d(k):
        """H2D of step k's inputs (pinned -> device) on the copy stream: overlaps step k-1's kernels"""
        xh, _ = host_sets[k % len(host_sets)]
        with torch.cuda.stream(copy_stream):
            dev_in[k % 2].copy_(xh, non_blocking=True)
            in_ready[k % 2] = torch.cuda.Event()
            in_ready[k % 2].record(copy_stream)

    def step(k, last):
        _, cam = host_sets[k % len(host_sets)]
        torch.cuda.current_stream().wait_event(in_ready[k % 2])
        if not last:
            h2d(k + 1)                                                   # next step's copy is in flight during this step
        xd = dev_in[k % 2]
        img = xd[:B].detach().requires_grad_(True)
        img_rot = xd[B:].detach().requires_grad_(True)
        loss, _ = f(img, cam[:B], img_rot, cam[B:], occlusion_aware=True)
        (loss * LAMBDA_ROTATE).backward()
        if with_grads:
            g_host[0].copy_(img.grad, non_blocking=True)
            g_host[1].copy_(img_rot.grad, non_blocking=True)
        loss_host.copy_(loss.detach(), non_blocking=False)               # D2H read of the step's result (blocks)

    nw = max(3, min(warmup, 5))
    h2d(0)
    for k in range(nw):
        step(k, False)
    torch.cuda.synchronize(h.dev)
    h.barrier()
    t0 = time.perf_counter()
    h2d(nw)                      # (the copy issued by the last warm-up step is repeated inside the timed region)
    for k in range(nw, nw + steps):
        step(k, k == nw + steps - 1)
    torch.cuda.synchronize(h.dev)
    dt = h.max_over_ranks(time.perf_counter() - t0)
    d2h = 4 + (2 * B * C * S * S * 4 if with_grads else 0)
    return {"value": h.world * B * steps / dt, "unit": UNIT, "h2d_bytes_per_step": 2 * B * C * S * S * 4 + 24 * B * 4,
            "d2h_bytes_per_step": d2h, "steps": steps,
            "pipeline": "step k+1's H2D (copy stream, double-buffered) overlaps step k's kernels; loss%s read back every step"
                        % (" and both gradient tensors" if with_grads else ""),
            "api": "rgbd_gan_b200.loss_functions.LossFuncRotate(grad_scale=lambda_rotate, return_new_zp=False) + backward()"}


def workload_config(a, pairs):
    return {"workload": "configs[4] throughput sweep: %d RGB-D pairs per GPU at %dx%d, consistency loss fwd+bwd (both warp "
                        "directions, occlusion mask on, L1, lambda_geometric 3, ffhq_stylegan_occlusion.yml pose ranges)"
                        % (pairs, a.size, a.size), "pairs_per_gpu": pairs, "size": a.size, "channels": 4,
            "depth": a.depth, "pose_ranges": "x 0.3054 / y 1.0472 rad (yml)", "upstream_grad": LAMBDA_ROTATE}


ALIGN_STEPS = 48                # untimed steps between the barrier and e0 of every timed region (see Workload.timed)
PREWARM_STEPS = 1500            # fixed count (every rank makes the same sequence of loss calls): ~0.2 s at 256 pairs


def run_ours(a):
    h = Harness(a)
    torch = h.torch
    B, S = a.pairs, a.size
    wl = Workload(h, B, S, depth=a.depth, label="headline")
    with torch.cuda.stream(h.stream):
        n0 = h.lib.rgbd_launch_count()
        wl.step_fused(wl.pool[0])
        launches_fused = h.lib.rgbd_launch_count() - n0
        n0 = h.lib.rgbd_launch_count()
        wl.step_two_pass(wl.pool[0])
        launches_two = h.lib.rgbd_launch_count() - n0
        wl.join_exchange()
        torch.cuda.synchronize(h.dev)
        for k in range(PREWARM_STEPS):               # bring clocks up before anything is timed
            wl.step_fused(wl.pool[k % wl.pool_n])
        wl.join_exchange()
        torch.cuda.synchronize(h.dev)
    sampler = ClockSampler(h.local)
    ms = wl.timed(a.steps, a.warmup, sampler=sampler)
    clocks = sampler.stop()
    value = h.world * B * a.steps / (ms * 1e-3)
    two_steps = max(3, min(a.steps, 200))
    ms_two = wl.timed(two_steps, a.warmup, step=wl.step_two_pass)
    roofline = kernel_roofline(wl, a.steps, ms)
    parity = wl.parity()
    e2e_steps = max(3, min(a.steps, 50))
    e2e = e2e_leg(h, wl, e2e_steps, a.warmup, with_grads=False)
    e2e["with_gradients_d2h"] = e2e_leg(h, wl, e2e_steps, a.warmup, with_grads=True)
    wl.free()

    # ---- the other BASELINE.json configurations, same timing method, every N
    k_extra = max(3, min(a.steps, 200))
    extra = {}
    if not a.no_extra:
        for name, kw in (("cfg1_32_pairs_per_gpu", dict(B=32, S=128, label="configs[1] ffhq_stylegan_occlusion.yml: 32 pairs per GPU")),
                         ("cfg4_256_pairs_at_256", dict(B=256, S=256, label="configs[4]: 256 pairs per GPU at 256x256")),
                         ("cfg2_car_64_pairs_total", dict(B=max(1, 64 // h.world), S=128, ranges=CAR, lam=1.0, n_global=max(1, 64 // h.world) * h.world,
                                                          label="configs[2] dcgan_shapenet_car.yml: 64 pairs in total (strong scaling), car pose "
                                                                "ranges (yaw +-pi), lambda_geometric 1"))):
            w2 = Workload(h, kw.pop("B"), kw.pop("S"), depth=a.depth, **kw)
            extra[name] = w2.record(k_extra, 3)
            if name.startswith("cfg2"):
                extra[name]["scaling"] = "strong"
            w2.free()

    line = None
    if h.rank == 0:
        coll = "none" if h.world == 1 else ({
            "lazy": "the kernel that finishes the loss publishes its 5 values into every peer's mailbox over NVLink peer memory "
                    "each call (no wait, at most 7 calls ahead of the slowest rank); summed in rank order when the loss is read: "
                    "once, at the end of the timed region, inside it",
            "deferred": "own finalize kernel all-reduces them over NVLink peer memory on a side stream; deferred by one call, the "
                        "last one joined inside the timed region",
            "joined": "own finalize kernel all-reduces them over NVLink peer memory; every call joined"}[h.exchange]
            if h.peer is not None else "NCCL on a side stream")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": h.world, "steps": a.steps, "warmup": a.warmup,
            "prewarm_steps": PREWARM_STEPS + 3, "align_steps": ALIGN_STEPS, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # `config` names the workload only (identical in both arms); how THIS arm ran it is under `run`
            "config": workload_config(a, B),
            "run": dict(l2="input pool of %d sets x %.0f MB rotated between steps (> L2)"
                        % (wl.pool_n, wl.bytes_per_set / 2 ** 20), cuda_graph=False,
                        path="rgbd_consistency_fwd_bwd (one-pass fwd+bwd, upstream grad = lambda_rotate): "
                             + ("row sweep" if wl.sweep_path else "three-kernel chain"),
                        parallelism="dp%d (pairs sharded; 4-float loss all-reduce: %s)" % (h.world, coll)),
            "two_pass": {"value": h.world * B * two_steps / (ms_two * 1e-3), "ms_per_step": ms_two / two_steps,
                         "path": "rgbd_consistency_fwd + rgbd_consistency_bwd (recompute)"},
            "e2e": e2e, "gpu_launches": int(launches_fused * a.steps),
            "launches_per_step": {"fwd_bwd": int(launches_fused), "two_pass": int(launches_two)},
            "roofline": roofline, "clocks": clocks, "extra": extra,
        }
        if parity is not None:
            line["parity"] = parity
    return line, dict(dev=h.dev, world=h.world, rank=h.rank, lib=h.lib, hbm_peak=h.hbm_peak, harness=h)


def sweep(a, ctx_, sizes, hinge=None, depth=None, **kw):
    """single-GPU sub-records (smooth depth, depth hinge, the DeepVoxels updater's variant): same Workload / timing"""
    h = ctx_["harness"]
    out = []
    for S, B in sizes:
        w = Workload(h, B, S, depth=depth or a.depth, hinge=hinge, **kw)
        out.append(w.record(20 if B >= 128 else 200, 3))
        w.free()
    return out


def deepvoxels_bench(ctx_):
    """cfg3 of BASELINE.json: DeepVoxels projection sampling fwd (frustum gather) + bwd (lift scatter).
    Production geometry (deepvoxels_generator.py:229-253: G=32, F=32, 64x64x56, yml batch 10 -> run 16) and
    BASELINE's 64^3 volume.  Algorithmic bytes per sample fwd+bwd = 2*(F*G^3 + F*D*H*W)*4."""
    import torch
    from rgbd_gan_b200 import _lib
    from tools import synthetic as poses
    dev, lib = ctx_["dev"], ctx_["lib"]
    out = []
    for G, B in ((32, 16), (64, 16)):
        F, img = 32, 64
        D = int(np.ceil(np.sqrt(3) * G))
        vs = (1. / G) * 1.1 * 0.5
        P = _lib.DvParams(img, img, D, G, 128., 128., 32., 32., float(np.float32(vs)), float(np.float32(np.sqrt(3) / 4)))
        np.random.seed(3)
        thetas = poses.sample_pose_pairs(B, poses.CAR_RANGES, True, np.random.default_rng(3))[:B]
        cam = torch.from_numpy(poses.cam2world(thetas).reshape(B, 16)).to(dev)
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
    from rgbd_gan_b200 import _lib
    from tools import synthetic as poses
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
        thetas = poses.sample_pose_pairs(B, poses.CAR_RANGES, True, np.random.default_rng(3))[:B]
        cam = torch.from_numpy(poses.cam2world(thetas).reshape(B, 16)).to(dev)
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


def pose_pipeline_bench(a, ctx_):
    """camera parameters -> cam2world -> warp constants for `pairs` pairs.  Device: ONE launch of
    rgbd_pose_pipeline (sampling included), CUDA events over 200 calls.  Host: what a step pays when the thetas live on the
    host -- the reference's NumPy chain (sampler, get_camera_matries, the pose algebra of LossFuncRotate.__call__) plus the
    pinned upload of the 24 floats per pair, wall clock."""
    import torch
    from types import SimpleNamespace
    from tools import synthetic as poses
    from rgbd_gan_b200.host_math import intrinsics_for_size, pose_algebra
    from rgbd_gan_b200.loss_functions import _PoseUploader
    from rgbd_gan_b200.pose_pipeline import CameraParamPrior, PosePipeline
    dev, B = ctx_["dev"], a.pairs
    K, inv_K = intrinsics_for_size(None, a.size, first=True)
    r = poses.FFHQ_RANGES
    cfg = SimpleNamespace(x_rotate=r[0], y_rotate=r[1], z_rotate=r[2], x_translate=r[3], y_translate=r[4], z_translate=r[5],
                          uniform_distribution=False)
    pipe = PosePipeline(CameraParamPrior(cfg, dev, seed=1), K, inv_K)
    for _ in range(10):
        pipe.step(2 * B)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        pipe.step(2 * B)
    e1.record()
    torch.cuda.synchronize(dev)
    dev_us = e0.elapsed_time(e1) * 1e3 / n
    up, rng = _PoseUploader(), np.random.default_rng(0)
    t0 = time.perf_counter()
    m = 50
    for _ in range(m):
        cam = poses.cam2world(poses.sample_pose_pairs(B, r, False, rng))
        M, c, Mi, ci = pose_algebra(K, inv_K, cam[:B], cam[B:])
        up.upload(M, c, Mi, ci, dev)
    torch.cuda.synchronize(dev)
    host_us = (time.perf_counter() - t0) * 1e6 / m
    return {"pairs": B, "device_one_launch_us": dev_us, "host_numpy_chain_plus_upload_us": host_us,
            "note": "device: rgbd_pose_pipeline (Philox sampling + get_camera_matries + pose algebra; the figure is the issue rate of "
                    "the Python glue -- three output allocations + one ctypes call -- the kernel itself runs a few us); host: NumPy sampler + get_camera_matries + pose_algebra + pinned H2D of 24 floats per "
                    "pair; bit-exactness of the device path against the host path: tests/test_gpu_poses.py"}


def feature_consistency_bench(a, ctx_):
    """the feature-space consistency loss of updater.py:345-354 (norm l2, C = 256 features +
    1 depth at 32x32, yml batch 32 -> 16 pairs); runs through the generic-C kernels (not tuned)."""
    import torch
    from rgbd_gan_b200 import _lib
    from tools import synthetic as poses
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
    ap.add_argument("--pairs", type=int, default=256)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--depth", default="rough", choices=["rough", "smooth"])
    ap.add_argument("--graph", action="store_true", help="(ignored: direct PDL-chained launches are the measured default)")
    ap.add_argument("--no-graph", action="store_true", help="(default; kept for compatibility)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"])
    ap.add_argument("--exchange", default="lazy", choices=["lazy", "deferred", "joined"],
                    help="N > 1 with --collective peer: when the exchanged loss parts are summed (see Workload)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the single-GPU sub-records (DeepVoxels, hinge, ...)")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configurations")
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
        h = ctx_["harness"]
        if h.peer is not None:
            h.peer.close()
        dist.barrier()
        dist.destroy_process_group()
        return
    if not a.no_sweep:
        if a.depth == "rough":
            line["smooth_depth"] = sweep(a, ctx_, sizes=((a.size, a.pairs),), depth="smooth", label="generator-like smooth depth")[0]
        # next row: the updaters' depth hinge (yml: depth_min 1.0, lambda_depth 10) fused in
        line["with_depth_hinge"] = sweep(a, ctx_, sizes=((a.size, a.pairs),), hinge=(1.0, 10.0), label="depth hinge fused")[0]
        # the DeepVoxels updater's call (updater_deepvoxels.py:176-190): 64x64, K = projection intrinsic, depth-range masks
        Kdv = np.array([[128., 0, 32., 0], [0, 128., 32., 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
        line["dv_updater_64_depth_masks"] = sweep(a, ctx_, sizes=((64, 1024),), K=Kdv, max_depth=3.0, min_depth=0.2, ranges=CAR,
                                                  label="updater_deepvoxels.py:176-190: 64x64, K given, max/min depth masks")[0]
        line["dv_updater_64_no_masks"] = sweep(a, ctx_, sizes=((64, 1024),), K=Kdv, ranges=CAR,
                                               label="same without the depth-range masks")[0]
        line["deepvoxels"] = deepvoxels_bench(ctx_)
        line["deepvoxels_render_fused"] = render_bench(ctx_)
        line["feature_consistency_c257"] = feature_consistency_bench(a, ctx_)
        line["pose_pipeline"] = pose_pipeline_bench(a, ctx_)
    if not a.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg(a)
    emit(line)


if __name__ == "__main__":
    main()
