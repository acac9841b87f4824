"""Multi-GPU (needs >= 2 GPUs; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
pairs sharded over one process per GPU; the global loss from the fused peer-memory all-reduce and from
the NCCL path both equal the single-process reference value, gradients equal the slices of the global
gradients, and all ranks hold bit-identical loss values."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, case_options, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from rgbd_gan_b200.distributed import PeerComm, shard_range
    from rgbd_gan_b200.loss_functions import LossFuncRotate
    g = load_golden(name)
    o = case_options(g)
    B = o["B"]
    lo, hi = shard_range(B, rank, world)
    x = torch.from_numpy(g["x"]).cuda()
    res = {}
    comm = PeerComm()
    for mode in ("peer", "nccl", "peer_fused", "peer_deferred", "peer_lazy"):
        kw = dict(peer_comm=comm) if mode.startswith("peer") else dict(process_group=dist.group.WORLD)
        if mode in ("peer_fused", "peer_deferred", "peer_lazy"):
            kw["grad_scale"] = o["gy"]
        if mode == "peer_deferred":
            kw["defer_loss"] = True
        if mode == "peer_lazy":
            kw["defer_loss"] = "lazy"             # publish only; summed by comm.wait()
        f = LossFuncRotate(None, norm=o["norm"], lambda_geometric=o["lam"], n_pairs_global=B, **kw)
        for rep in range(20 if mode == "peer_lazy" else 3):   # several calls: epochs advance, mailbox slots are reused
            img = x[:B][lo:hi].clone().requires_grad_(True)
            img_rot = x[B:][lo:hi].clone().requires_grad_(True)
            loss, _ = f(img, g["cam"][:B][lo:hi], img_rot, g["cam"][B:][lo:hi], occlusion_aware=o["occ"])
            (loss * o["gy"]).backward()
        if mode in ("peer_deferred", "peer_lazy"):
            comm.wait()                           # only now may the loss be read
        if mode == "peer_lazy":
            res["lazy_status"] = np.int32(comm.status())
        res[mode + "_loss"] = loss.detach().cpu().numpy()
        res[mode + "_gi"] = img.grad.cpu().numpy()
        res[mode + "_gr"] = img_rot.grad.cpu().numpy()
    comm.close()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), lo=lo, hi=hi, **res)
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["loss_s64_l1_noocc", "loss_cfg0_l1_occ"])
def test_sharded_pairs_two_gpus(name, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), name, str(tmp_path)), nprocs=world, join=True)
    g = load_golden(name)
    outs = [np.load(tmp_path / ("rank%d.npz" % r)) for r in range(world)]
    assert all(int(d["lazy_status"]) == 0 for d in outs)
    for mode in ("peer", "nccl", "peer_fused", "peer_deferred", "peer_lazy"):
        for d in outs:
            assert abs(float(d[mode + "_loss"]) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
            lo, hi = int(d["lo"]), int(d["hi"])
            assert np.abs(d[mode + "_gi"] - g["g_img"][lo:hi]).max() <= 1e-5 * np.abs(g["g_img"]).max()
            assert np.abs(d[mode + "_gr"] - g["g_img_rot"][lo:hi]).max() <= 1e-5 * np.abs(g["g_img_rot"]).max()
        assert outs[0][mode + "_loss"].tobytes() == outs[1][mode + "_loss"].tobytes()    # same bits on every rank
