"""Measures the element-wise relative gradient error of the CUDA paths against the reference's golden vectors and the
C oracle, over entries with |ref| > 1e-3 * max|ref| (SURVEY 8c's element-wise criterion), so the bound in
tests/conftest.py::assert_grad_close is set from data instead of by argument.  Run on the GPU box:
    python tools/grad_tolerance.py > gpurun_out/grad_tolerance.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def elem_err(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max()
    big = np.abs(ref) > 1e-3 * scale
    rel = np.abs(got - ref)[big] / np.abs(ref)[big]
    return {"max_norm_rel": float(np.abs(got - ref).max() / scale), "elem_rel_max": float(rel.max()),
            "elem_rel_p999": float(np.quantile(rel, 0.999)), "elem_rel_median": float(np.median(rel)), "n": int(big.sum())}


def main():
    import oracle
    from conftest import LOSS_CASES, case_options, load_golden
    from gpu_util import Consistency
    from oracle import numpy_port as npp
    oracle.build()
    out = {}
    for mode in ("0", "1"):
        os.environ["RGBD_B200_SWEEP"] = mode
        for name in LOSS_CASES:
            g = load_golden(name)
            o = case_options(g)
            port = npp.LossFuncRotateNP(K=None if o["K"] is None else o["K"].copy(), norm=o["norm"], lambda_geometric=o["lam"])
            port.init_params(o["S"])
            drv = Consistency(g["x"], g["cam"], o["B"], port.K, port.inv_K, norm=o["norm"], lam=o["lam"], occ=o["occ"],
                              max_depth=o["max_depth"], min_depth=o["min_depth"])
            _, gi, gr = drv.fwd_bwd(gy=o["gy"])
            out["%s/mode%s/golden" % (name, mode)] = {"g_img": elem_err(gi, g["g_img"]), "g_img_rot": elem_err(gr, g["g_img_rot"])}
        # full size against the C oracle
        B, S = 64, 128
        x, cam = npp.synthetic_batch(B, S, depth="rough", seed=11)
        port = npp.LossFuncRotateNP(lambda_geometric=3)
        port.init_params(S)
        drv = Consistency(x, cam, B, port.K, port.inv_K, lam=3.0, occ=True)
        M, c, Mi, ci = drv.host_poses
        ref_gi, ref_gr = oracle.consistency_bwd(x[:B], x[B:], M, c, Mi, ci, norm=1, occlusion=True, lambda_geometric=3, gy=2.0)
        _, gi, gr = drv.fwd_bwd(gy=2.0)
        out["full_64x128/mode%s/c_oracle" % mode] = {"g_img": elem_err(gi, ref_gi), "g_img_rot": elem_err(gr, ref_gr)}
    worst = max(max(v["g_img"]["elem_rel_max"], v["g_img_rot"]["elem_rel_max"]) for v in out.values())
    worst_norm = max(max(v["g_img"]["max_norm_rel"], v["g_img_rot"]["max_norm_rel"]) for v in out.values())
    print(json.dumps({"worst_elem_rel": worst, "worst_max_norm_rel": worst_norm, "cases": out}, indent=1))


if __name__ == "__main__":
    main()
