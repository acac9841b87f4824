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
