"""CPU, world_size 2 over gloo: the sharding contract of the multi-GPU path.

Each rank evaluates ITS shard of pairs with the global pair count in the denominators (here
through the oracle, since there is no GPU), the four loss means are all-reduced, and the
result must equal the single-process evaluation; shard gradients must equal the matching
slices of the global gradients with no communication."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from conftest import ROOT, case_options, load_golden  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import numpy_port as npp
    from rgbd_gan_b200.distributed import allreduce_loss, max_over_ranks, shard_range
    g = load_golden(name)
    o = case_options(g)
    B, S = o["B"], o["S"]
    x, cam = g["x"], g["cam"]
    port_ = npp.LossFuncRotateNP(lambda_geometric=o["lam"])
    port_.init_params(S)
    M, c, Mi, ci = port_.pose_algebra(cam[:B], cam[B:])
    lo, hi = shard_range(B, rank, world)
    sl = slice(lo, hi)
    kw = dict(norm=1, occlusion=o["occ"], n_pairs_global=B)
    parts = oracle.consistency_fwd(x[:B][sl], x[B:][sl], M[sl], c[sl], Mi[sl], -ci[sl], **kw)
    loss, gparts = allreduce_loss(torch.tensor(parts, dtype=torch.float32), o["lam"])
    gi, gr = oracle.consistency_bwd(x[:B][sl], x[B:][sl], M[sl], c[sl], Mi[sl], -ci[sl], lambda_geometric=o["lam"],
                                    gy=o["gy"], **kw)
    slowest = max_over_ranks(float(rank + 1), torch.device("cpu"))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=loss.numpy(), gi=gi, gr=gr, lo=lo, hi=hi, slowest=slowest)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["loss_s64_l1_noocc", "loss_cfg0_l1_occ"])
def test_sharded_loss_and_grads_equal_single_process(name, tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(name)
    o = case_options(g)
    for r in range(world):
        d = np.load(tmp_path / ("rank%d.npz" % r))
        assert abs(float(d["loss"]) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
        lo, hi = int(d["lo"]), int(d["hi"])
        sc = np.abs(g["g_img"]).max()
        assert np.abs(d["gi"] - g["g_img"][lo:hi]).max() <= 1e-5 * sc
        assert np.abs(d["gr"] - g["g_img_rot"][lo:hi]).max() <= 1e-5 * np.abs(g["g_img_rot"]).max()
        assert float(d["slowest"]) == float(world)
