import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """a bare `pytest` on a box without a GPU skips the gpu-marked tests instead of failing them"""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:            # noqa: BLE001
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


def load_golden(name):
    return np.load(golden_path(name))


LOSS_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "loss_*.npz"))
                    if "growing" not in p)
DV_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "dv_*.npz")))


def case_options(g):
    """decode the option fields make_golden.py stores with every loss case"""
    return dict(
        B=int(g["B"]), C=int(g["C"]), S=int(g["S"]), norm=str(g["norm"]), lam=float(g["lambda_geometric"]),
        occ=bool(g["occlusion_aware"]), gy=float(g["gy"]),
        max_depth=None if float(g["max_depth"]) < 0 else float(g["max_depth"]),
        min_depth=None if float(g["min_depth"]) < 0 else float(g["min_depth"]),
        K=g["K_in"] if bool(g["has_K"]) else None)


def rel_max(a, b):
    """max-norm error relative to the max-norm of the reference tensor (SURVEY.md 8(c) tolerance)"""
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() /
                 max(float(np.abs(b).max()), 1e-30))


def assert_grad_close(got, ref, tol=1e-5, elem_rtol=5e-4):
    """Gradient tolerance of the north star: |a-b| <= 1e-5 * max|ref| for every element (error relative to the tensor's
    max-norm; measured worst case over every golden case and a 64-pair full-size run, both C == 4 paths: 6.8e-7,
    profiles/r02_grad_tolerance.json).

    SURVEY 8(c) adds an element-wise check for entries larger than 1e-3 * max|ref|.  Many entries are exactly 0 (masked /
    occluded pixels) and the depth-channel entries are sums with heavy cancellation (M^T gq . (x, y, 1) with x, y up to
    127), so two fp32 evaluations of the SAME reference graph that differ only in summation order already disagree at the
    1e-4 level on such entries.  The bound is set from data, not by argument: tools/grad_tolerance.py measures the
    element-wise relative error of the CUDA paths against the golden vectors / the C oracle (worst entry 2.1e-4, 99.9th
    percentile <= 3e-5, median ~1e-7); `elem_rtol` is about 2x the measured worst case."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= tol * scale, (np.abs(got - ref).max(), scale)
    big = np.abs(ref) > 1e-3 * scale
    if big.any():
        assert (np.abs(got - ref)[big] / np.abs(ref)[big]).max() <= elem_rtol


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def assert_frustum_close(got, ref, exact):
    """fused DeepVoxels projection: `exact` (RGBD_B200_DV_EXACT=1) is bit-equal to the reference's fp32 chain
    ((v*wx)*wy)*wz; the default folds the corner weights once per element (8 FMAs per feature): identical
    zero pattern (the in-grid mask is computed the same way), values within 1e-5 of the tensor's max-norm."""
    import numpy as np
    if exact:
        np.testing.assert_array_equal(got, ref)
        return
    assert got.shape == ref.shape
    assert np.array_equal(got == 0, ref == 0) or np.abs(got[(got == 0) != (ref == 0)]).max() < 1e-30
    scale = float(np.abs(ref).max())
    assert float(np.abs(got - ref).max()) <= 1e-5 * scale, float(np.abs(got - ref).max()) / scale
