"""CPU: host-side logic of the product package (no compute calls: there is no GPU here) and the
C-ABI library's exported symbols."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import LOSS_CASES, ROOT, case_options, load_golden
from oracle import numpy_port as npp


def test_library_builds_and_exports_header_symbols():
    from rgbd_gan_b200 import _lib, build
    path = build.build()
    assert os.path.exists(path)
    hdr = open(os.path.join(ROOT, "include", "rgbdgan_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rgbd_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 17
    lib = ctypes.CDLL(path)                       # loads without a GPU (the CUDA runtime is linked statically)
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib.rgbd_version.restype = ctypes.c_int
    assert lib.rgbd_version() == 100
    # pure host queries work without a device
    lib.rgbd_consistency_workspace_bytes.restype = ctypes.c_size_t
    assert lib.rgbd_consistency_workspace_bytes(4, 4, 128, 128) > 2 * 2 * 4 * 4 * 128 * 128 * 4
    assert lib.rgbd_consistency_workspace_bytes(0, 4, 128, 128) == 0


def test_argument_errors_are_reported_not_raised():
    from rgbd_gan_b200 import _lib
    lib = _lib.load()
    opts = _lib.LossOpts(1, 1, float("nan"), float("nan"), 3.0, 0)
    rc = lib.rgbd_consistency_fwd(None, None, None, None, None, None, 1, 4, 8, 8, ctypes.byref(opts), None, None, None,
                                  None, 0, None)
    assert rc == -1 and b"null" in lib.rgbd_last_error()
    with pytest.raises(_lib.RgbdB200Error):
        _lib.check(rc, "rgbd_consistency_fwd")


def test_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    from rgbd_gan_b200.loss_functions import LossFuncRotate, bilinear
    x = torch.zeros(1, 4, 8, 8)
    cam = np.tile(np.eye(4, dtype=np.float32), (1, 1, 1))
    with pytest.raises(TypeError):
        LossFuncRotate(None)(x, cam, x, cam)
    with pytest.raises(TypeError):
        bilinear(x, torch.zeros(1, 64, 3))


@pytest.mark.parametrize("name", LOSS_CASES)
def test_pose_algebra_matches_reference_intermediates(name):
    """product host code == oracle port; K/inv_K/p == what the reference computed"""
    from rgbd_gan_b200.loss_functions import LossFuncRotate, pose_algebra
    g = load_golden(name)
    o = case_options(g)
    B, S = o["B"], o["S"]
    obj = LossFuncRotate(None, K=None if o["K"] is None else o["K"].copy(), norm=o["norm"], lambda_geometric=o["lam"])
    obj.init_params(None, size=S)
    np.testing.assert_array_equal(obj.K, g["K"])
    np.testing.assert_array_equal(obj.inv_K, g["inv_K"])
    np.testing.assert_array_equal(obj.p, g["p"])
    port = npp.LossFuncRotateNP(K=None if o["K"] is None else o["K"].copy())
    port.init_params(S)
    M, c, Mi, ci = pose_algebra(obj.K, obj.inv_K, g["cam"][:B], g["cam"][B:])
    M2, c2, Mi2, ci2 = port.pose_algebra(g["cam"][:B], g["cam"][B:])
    np.testing.assert_array_equal(M, M2)
    np.testing.assert_array_equal(c, c2)
    np.testing.assert_array_equal(Mi, Mi2)
    np.testing.assert_array_equal(ci, -ci2)


def test_growing_state_product_host():
    from rgbd_gan_b200.loss_functions import LossFuncRotate
    g = load_golden("loss_growing")
    obj = LossFuncRotate(None)
    for S in (32, 64):
        obj.init_params(None, size=S)
        np.testing.assert_array_equal(obj.K, g["K_%d" % S])
        np.testing.assert_array_equal(obj.inv_K, g["inv_K_%d" % S])
        assert obj.size == S


@pytest.mark.parametrize("name,seed,ranges,uniform", [
    ("loss_cfg0_l1_occ", 0, (0.3054, 1.0472, 0, 0, 0, 0), None),
    ("loss_car_l1_occ", 1, (0.3054, 3.1415, 0, 0, 0, 0), None),
    ("loss_dv_maxdepth", 5, (0.3054, 3.1415, 0, 0, 0, 0), True),
    ("loss_edge_wild", 7, (0.3054, 3.1415, 0, 0.3, 0.2, 0.4), None),
])
def test_pose_generators_replay_reference(name, seed, ranges, uniform):
    """CameraParamPrior.sample + get_camera_matries, replayed with the reference's seed, reproduce
    the thetas / cam2world matrices the reference's own helpers produced (make_golden.py)."""
    g = load_golden(name)
    np.random.seed(seed)
    np.testing.assert_array_equal(npp.sample_camera_prior(g["thetas"].shape[0], ranges, bool(uniform)), g["thetas"])
    np.testing.assert_array_equal(npp.get_camera_matries(g["thetas"]), g["cam"])


def test_shard_range():
    from rgbd_gan_b200.distributed import shard_range
    for n in (1, 7, 16, 64, 256):
        for w in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 4, 4)


def test_dv_params_follow_generator_constants():
    """deepvoxels_generator.py:229-253 constants -> rgbd_dv_params"""
    from rgbd_gan_b200.projection import ProjectionHelper
    G = 32
    intr = np.array([[128., 0, 32., 0], [0, 128., 32., 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    vs = (1. / G) * 1.1 * 0.5
    h = ProjectionHelper(intr, intr, [64, 64], [64, 64], 0., G * vs + np.sqrt(3) / 4, [G] * 3, vs, np.sqrt(3) / 4,
                         int(np.ceil(np.sqrt(3) * G)), verbose=False)
    P = h.params()
    assert (P.W, P.H, P.D, P.G) == (64, 64, 56, 32)
    assert P.fx == 128.0 and P.cx == 32.0
    assert P.voxel_size == np.float32(vs) and P.near_plane == np.float32(np.sqrt(3) / 4)
    with pytest.raises(ValueError):
        ProjectionHelper(intr, intr, [64, 64], [64, 64], 0., 1., [32, 32, 16], vs, 0.4, 56, verbose=False)


@pytest.mark.parametrize("Bc,S,grad,fold,lags", [(1, 16, 1, 1, (1, 1)), (3, 32, 1, 1, (2, 3)), (5, 128, 1, 1, (2, 3)),
                                                  (4, 64, 0, 1, (2, 3)), (4, 64, 1, 0, (1, 4)), (2, 128, 1, 1, (40, 40)),
                                                  (7, 24, 1, 1, (3, 1))])
def test_pipeline_ticket_schedule_is_complete_and_ordered(Bc, S, grad, fold, lags):
    """the ticket order of the opt-in single-launch pipeline kernel (rgbd_debug_mega_schedule, host code only):
    every stage-in / main / stage-out tile of every pair appears exactly once, there is one finalize ticket iff the
    loss is folded in, and every ticket depends only on tickets with SMALLER numbers -- main(p) on all stage-in(p),
    stage-out(p) on all main(p), finalize on all main -- which is what makes the in-order hand-out deadlock-free"""
    import ctypes
    from rgbd_gan_b200 import _lib
    lib = _lib.load()
    cap = 1 << 20
    buf = (ctypes.c_int * (3 * cap))()
    total = ctypes.c_int(0)
    rc = lib.rgbd_debug_mega_schedule(Bc, S, S, grad, fold, lags[0], lags[1], buf, cap, ctypes.byref(total))
    assert rc == 0 and 0 < total.value <= cap
    t = np.ctypeslib.as_array(buf)[:3 * total.value].reshape(-1, 3)
    HW = S * S
    TS, TM = -(-HW // 1024), -(-HW // 512)
    seen = {}
    for n, (role, pair, idx) in enumerate(t):
        if role == 0:
            continue                                          # padding ticket (tile index beyond the image)
        assert (role, pair, idx) not in seen
        seen[(role, pair, idx)] = n
    for p in range(Bc):
        si = [seen[(1, p, i)] for i in range(2 * TS)]
        mn = [seen[(2, p, i)] for i in range(2 * TM)]
        assert max(si) < min(mn)
        if grad:
            so = [seen[(3, p, i)] for i in range(2 * TS)]
            assert max(mn) < min(so)
    n_fin = sum(1 for k in seen if k[0] == 4)
    assert n_fin == (1 if fold else 0)
    if fold:
        assert all(v < seen[(4, 0, 0)] for k, v in seen.items() if k[0] == 2)
    assert len(seen) == Bc * (2 * TS * (2 if grad else 1) + 2 * TM) + n_fin
    assert lib.rgbd_debug_mega_schedule(Bc, S, S, grad, fold, 0, 1, buf, cap, ctypes.byref(total)) == -1   # lag 0 is refused


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU stand-in of the reference's path, runnable without a GPU): exactly one JSON
    line on stdout with the keys the driver's contract names"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]


# ---- the rounding orders csrc/poses.cu hard-codes, restated in NumPy (fp32 products are exact in fp64, so
# float32(float64(a) * float64(b) + float64(c)) is a single-rounding fma except for rare double roundings) and checked against
# this host's NumPy / BLAS on every golden case.  Pins the claim "the device pose pipeline reproduces the reference's CPU
# path product by product" on the CPU side too; skipped on a host whose BLAS uses other kernels (the GPU tests compare
# the kernel with the golden vectors directly and do not depend on the host).
def _fma(a, b, c):
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def _chain(A, Bm):
    n, k = A.shape
    out = np.zeros((n, Bm.shape[1]), np.float32)
    for i in range(n):
        for j in range(Bm.shape[1]):
            s = np.float32(A[i, 0] * Bm[0, j])
            for l in range(1, k):
                s = _fma(A[i, l], Bm[l, j], s)
            out[i, j] = s
    return out


def _matvec_blas(A, v):
    out = np.zeros((3, 1), np.float32)
    for i in range(2):
        s = np.float32(A[i, 0] * v[0, 0])
        s = np.float32(s + np.float32(A[i, 1] * v[1, 0]))
        out[i, 0] = np.float32(s + np.float32(A[i, 2] * v[2, 0]))
    out[2, 0] = _fma(A[2, 2], v[2, 0], _fma(A[2, 0], v[0, 0], np.float32(A[2, 1] * v[1, 0])))
    return out


def _device_order_pose_algebra(K, inv_K, th, thr):
    R1, R2 = th[:3, :3], thr[:3, :3]
    R = _chain(np.ascontiguousarray(R2.T), R1)
    dt = (thr[:3, 3:] - th[:3, 3:]).astype(np.float32)
    t = _chain(np.ascontiguousarray(R1.T), dt)
    KR = _chain(K, R)
    return _chain(KR, inv_K), _matvec_blas(KR, t), _chain(_chain(K, np.ascontiguousarray(R.T)), inv_K), -_matvec_blas(K, t)


def _host_blas_matches_golden_host():
    rng = np.random.default_rng(0)
    A, v = rng.normal(size=(5, 3, 3)).astype(np.float32), rng.normal(size=(5, 3, 1)).astype(np.float32)
    Bm = rng.normal(size=(5, 3, 3)).astype(np.float32)
    return all(np.array_equal(np.matmul(A, v)[b], _matvec_blas(A[b], v[b])) and
               np.array_equal(np.matmul(A, Bm)[b], _chain(A[b], Bm[b])) for b in range(5))


@pytest.mark.parametrize("name", LOSS_CASES)
def test_device_pose_orders_restated_in_numpy(name):
    from rgbd_gan_b200.host_math import intrinsics_for_size, pose_algebra
    if not _host_blas_matches_golden_host():
        pytest.skip("this host's BLAS evaluates small products in another order than the golden host's")
    g = load_golden(name)
    o = case_options(g)
    B = o["B"]
    K, inv_K = intrinsics_for_size(o["K"], o["S"], first=True)
    M, c, Mi, ci = pose_algebra(K, inv_K, g["cam"][:B], g["cam"][B:])
    for b in range(B):
        m, cc, mi, cci = _device_order_pose_algebra(K, inv_K, g["cam"][b], g["cam"][B + b])
        np.testing.assert_array_equal(m, M[b])
        np.testing.assert_array_equal(cc, c[b])
        np.testing.assert_array_equal(mi, Mi[b])
        np.testing.assert_array_equal(cci, ci[b])


def test_device_camera_matrix_order_restated_in_numpy():
    """get_camera_matries (updater.py:45-60): three 4x4 products as fma chains reproduce every golden cam2world matrix when
    fed NumPy's own cos / sin"""
    import glob
    import os
    from conftest import GOLDEN
    if not _host_blas_matches_golden_host():
        pytest.skip("other BLAS order on this host")
    for path in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        g = np.load(path)
        if "thetas" not in g.files:
            continue
        th = g["thetas"]
        if not np.array_equal(npp.get_camera_matries(th), g["cam"]):      # this host's cos / sin differ from the golden host's
            continue
        for b in range(min(len(th), 2)):
            mat = np.zeros((4, 4), np.float32)
            mat[range(4), range(4)] = [1, 1, -1, 1]
            mat[2, 3] = 1
            for i in (0, 1, 2):
                a1, a2 = (i + 1) % 3, (i + 2) % 3
                rot = np.eye(4, dtype=np.float32)
                cs, sn = np.cos(th[b:b + 1, i])[0], np.sin(th[b:b + 1, i])[0]
                rot[a1, a1], rot[a1, a2], rot[a2, a1], rot[a2, a2] = cs, -sn, sn, cs
                mat = _chain(rot, mat)
            mat[:3, 3] = mat[:3, 3] + th[b, 3:]
            np.testing.assert_array_equal(mat, g["cam"][b])
