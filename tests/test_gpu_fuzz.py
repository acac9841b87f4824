"""A seeded slice of the randomised parity sweep (tools/fuzz_parity.py: random shapes, channel counts, depth statistics,
pose ranges, options and execution paths against the C oracle; masks and new_zp bit-exact, losses / gradients 1e-5).  The
long runs are logged under profiles/ (r02_fuzz_parity.json)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_cases_against_oracle(seed, oracle_mod, monkeypatch):
    import fuzz_parity
    monkeypatch.delenv("RGBD_B200_SWEEP", raising=False)
    monkeypatch.delenv("RGBD_B200_SWEEP_CTAS", raising=False)
    saved = {k: os.environ.get(k) for k in ("RGBD_B200_SWEEP", "RGBD_B200_SWEEP_CTAS")}
    rng = np.random.default_rng(seed)
    try:
        for _ in range(8):
            k = fuzz_parity.draw_case(rng)
            fuzz_parity.run_case(k, oracle_mod)
    finally:
        for name, val in saved.items():
            if val is None:
                os.environ.pop(name, None)
            else:
                os.environ[name] = val


@pytest.mark.parametrize("seed", [1, 2])
def test_random_deepvoxels_cases_against_oracle(seed, oracle_mod, monkeypatch):
    """tools/fuzz_parity_dv.py: random grid / image sizes (incl. widths that are not powers of two), feature counts, poses,
    grid2world, exact / folded mode: compute_proj_idcs bit-exact, fused projection and lift against the C oracle"""
    import fuzz_parity_dv
    saved = os.environ.get("RGBD_B200_DV_EXACT")
    rng = np.random.default_rng(seed)
    try:
        for _ in range(10):
            fuzz_parity_dv.run_case(fuzz_parity_dv.draw_case(rng), oracle_mod)
    finally:
        if saved is None:
            os.environ.pop("RGBD_B200_DV_EXACT", None)
        else:
            os.environ["RGBD_B200_DV_EXACT"] = saved
