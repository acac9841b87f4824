"""Randomised parity sweep of the DeepVoxels projection entry points against the C oracle.

    python tools/fuzz_parity_dv.py [n_cases] [seed] [out.json]

Random grid size, image size (incl. widths that are not powers of two: quirk Q5's true division), feature count (F % 4 != 0
takes the planar kernels), batch, voxel scale, intrinsics, poses (car ranges, translated cameras, partly or wholly outside
the grid) and the exact / folded-weight mode.  Checks through the C-ABI: compute_proj_idcs lin_ind / voxel_coords BIT-EXACT
(also with a random grid2world), fused project forward (bit-exact with RGBD_B200_DV_EXACT=1, else same zero pattern and
1e-5), lift (gradient) 1e-5, and the adjointness <project(grid), g> == <grid, lift(g)>.  Test infrastructure."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def draw_case(rng):
    G = int(rng.choice([8, 12, 16, 24, 32]))
    img = int(rng.choice([16, 20, 24, 32, 48, 64]))
    return dict(G=G, img=img, F=int(rng.choice([1, 3, 4, 8, 12, 32])), B=int(rng.integers(1, 4)),
                scale=float(rng.choice([0.5, 0.35, 0.8])), focal=float(rng.choice([2.0, 1.5, 2.7])),
                shift=float(rng.choice([0.0, 0.0, 0.2, 1.5])), exact=bool(rng.integers(0, 2)),
                g2w=bool(rng.random() < 0.3), seed=int(rng.integers(0, 1 << 30)))


def run_case(k, oracle):
    import torch
    from conftest import assert_frustum_close, assert_grad_close
    from gpu_util import DEV, dev, p, stream
    from oracle import numpy_port as npp
    from rgbd_gan_b200 import _lib
    from rgbd_gan_b200._lib import DvParams
    G, img, F, B = k["G"], k["img"], k["F"], k["B"]
    D = int(np.ceil(np.sqrt(3) * G))
    vs = (1. / G) * 1.1 * k["scale"]
    near = np.sqrt(3) / 4
    K = np.array([[img * k["focal"], 0, img / 2., 0], [0, img * k["focal"], img / 2., 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    np.random.seed(k["seed"] % (1 << 31))
    cam = npp.get_camera_matries(npp.sample_camera_prior(2 * B, npp.CAR_RANGES, True)[:B])
    rng = np.random.default_rng(k["seed"])
    cam[:, :3, 3] += (k["shift"] * rng.uniform(-1, 1, size=(B, 3))).astype(np.float32)
    grid = rng.normal(size=(B, F, G, G, G)).astype(np.float32)
    P0 = oracle.dv_params(img, img, D, G, K, vs, near)
    P = DvParams(img, img, D, G, float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2]), float(np.float32(vs)),
                 float(np.float32(near)))
    n = img * img * D
    lib = _lib.load()
    ws_i = torch.empty(lib.rgbd_dv_workspace_bytes(ctypes.byref(P)), dtype=torch.uint8, device=DEV)
    g2w = None
    if k["g2w"]:
        g2w = np.eye(4, dtype=np.float32)
        g2w[:3, :3] += rng.uniform(-0.1, 0.1, size=(3, 3)).astype(np.float32)
        g2w[:3, 3] = rng.uniform(-0.05, 0.05, size=3).astype(np.float32)
    kept = 0
    # device copies stay referenced until the end of the case: a temporary handed to the library as a raw pointer would go
    # back to torch's caching allocator at once and could be overwritten by the next temporary
    d_grid, d_cam = dev(grid), dev(cam.reshape(B, 16))
    for b in range(B):                                            # compute_proj_idcs, sample by sample like the reference
        lin = torch.full((n,), -1, dtype=torch.int32, device=DEV)
        vc = torch.full((3, n), float("nan"), device=DEV)
        M = ctypes.c_int(-1)
        if g2w is None:
            _lib.call("rgbd_dv_compute_proj_idcs", ctypes.byref(P), ctypes.c_void_p(d_cam[b].data_ptr()), p(lin), p(vc),
                      ctypes.byref(M), p(ws_i), ws_i.numel(), stream())
        else:
            d_w2g = dev(np.ascontiguousarray(np.linalg.inv(g2w), dtype=np.float32).reshape(16))
            _lib.call("rgbd_dv_compute_proj_idcs_g2w", ctypes.byref(P), ctypes.c_void_p(d_cam[b].data_ptr()), p(d_w2g),
                      p(lin), p(vc), ctypes.byref(M), p(ws_i), ws_i.numel(), stream())
        ref = oracle.dv_compute_proj_idcs(P0, cam[b], g2w)
        if ref is None:
            assert M.value == 0
            continue
        assert M.value == ref[0].size
        kept += M.value
        np.testing.assert_array_equal(lin[:M.value].cpu().numpy(), ref[0])
        np.testing.assert_array_equal(vc[:, :M.value].cpu().numpy(), ref[1])
    os.environ["RGBD_B200_DV_EXACT"] = "1" if k["exact"] else "0"
    ref_f = oracle.dv_project_fwd(P0, grid, cam)
    out = torch.full((B, F, n), float("nan"), device=DEV)
    nbytes = lib.rgbd_dv_project_workspace_bytes(ctypes.byref(P), B, F)
    ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=DEV)
    _lib.call("rgbd_dv_project_fwd", ctypes.byref(P), p(d_grid), p(d_cam), B, F, p(out), p(ws), int(nbytes), stream())
    got = out.cpu().numpy().reshape(ref_f.shape)
    planar = (F % 4) != 0                                         # the planar kernels always evaluate the exact chain
    if np.abs(ref_f).max() > 0:
        assert_frustum_close(got, ref_f, k["exact"] or planar)
    else:
        assert not got.any()
    g_out = rng.normal(size=ref_f.shape).astype(np.float32)
    ref_g = oracle.dv_project_bwd(P0, g_out, cam)
    gg = torch.full((B, F, G ** 3), float("nan"), device=DEV)
    d_gout = dev(g_out)
    _lib.call("rgbd_dv_project_bwd", ctypes.byref(P), p(d_gout), p(d_cam), B, F, p(gg), p(ws), int(nbytes), stream())
    ggn = gg.cpu().numpy().reshape(ref_g.shape)
    if np.abs(ref_g).max() > 0:
        assert_grad_close(ggn, ref_g)
    else:
        assert not ggn.any()
    lhs = float((got.astype(np.float64) * g_out).sum())
    rhs = float((grid.astype(np.float64) * ggn.reshape(grid.shape)).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0) + 1e-3
    return dict(kept_fraction=kept / float(B * n))


def main():
    import oracle
    oracle.build()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = sys.argv[3] if len(sys.argv) > 3 else None
    rng = np.random.default_rng(seed)
    res, fails, t0 = [], 0, time.time()
    for _ in range(n):
        k = draw_case(rng)
        try:
            res.append(dict(case=k, ok=True, **run_case(k, oracle)))
        except AssertionError as e:
            fails += 1
            res.append(dict(case=k, ok=False, error=str(e)[:400]))
            print("FAIL", k, str(e)[:300], flush=True)
    summary = dict(cases=n, seed=seed, failed=fails, seconds=round(time.time() - t0, 1),
                   empty_frusta=sum(1 for r in res if r["ok"] and r["kept_fraction"] == 0.0),
                   with_grid2world=sum(1 for r in res if r["case"]["g2w"]),
                   checks="compute_proj_idcs (+ grid2world) bit-exact; fused project fwd bit-exact (exact mode / planar) or same "
                          "zero pattern + 1e-5; lift 1e-5 of max-norm; adjointness")
    print(json.dumps(summary))
    if out:
        json.dump(dict(summary=summary, failures=[r for r in res if not r["ok"]]), open(out, "w"), indent=1)
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
