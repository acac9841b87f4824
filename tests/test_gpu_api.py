"""GPU: the Python mirror of the reference call surface (rgbd_gan_b200.loss_functions / projection),
driven the way updater.py / deepvoxels_generator.py drive the reference, checked against the golden
vectors of the unmodified reference."""
import numpy as np
import pytest

from conftest import DV_CASES, LOSS_CASES, assert_frustum_close, assert_grad_close, case_options, load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
DEV = "cuda:0"


def _run(name, **ctor):
    from rgbd_gan_b200.loss_functions import LossFuncRotate
    g = load_golden(name)
    o = case_options(g)
    B = o["B"]
    x = torch.from_numpy(g["x"]).to(DEV)
    img = x[:B].clone().requires_grad_(True)
    img_rot = x[B:].clone().requires_grad_(True)
    f = LossFuncRotate(None, K=None if o["K"] is None else o["K"].copy(), norm=o["norm"], lambda_geometric=o["lam"],
                       **ctor)
    kw = dict(occlusion_aware=o["occ"])
    if o["max_depth"] is not None:
        kw["max_depth"] = o["max_depth"]
    if o["min_depth"] is not None:
        kw["min_depth"] = o["min_depth"]
    loss, zp = f(img, g["cam"][:B], img_rot, g["cam"][B:], **kw)
    (loss * o["gy"]).backward()            # loss_gen += loss_rotate * lambda_rotate; loss_gen.backward()
    return g, o, f, loss, zp, img.grad, img_rot.grad


@pytest.mark.parametrize("name", LOSS_CASES)
@pytest.mark.parametrize("fuse", [True, False])
def test_loss_func_rotate_like_updater(name, fuse):
    g, o, f, loss, zp, gi, gr = _run(name, fuse_backward=fuse)
    assert loss.dim() == 0
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    np.testing.assert_array_equal(zp.detach().cpu().numpy(), g["new_zp_cat"])
    assert_grad_close(gi.cpu().numpy(), g["g_img"])
    assert_grad_close(gr.cpu().numpy(), g["g_img_rot"])
    np.testing.assert_array_equal(f.K, g["K"])
    np.testing.assert_array_equal(f.inv_K, g["inv_K"])


@pytest.mark.parametrize("name", ["loss_cfg0_l1_occ", "loss_s32_l2_feat", "loss_dv_mindepth"])
@pytest.mark.parametrize("scale", [2.0, 0.5])
def test_fused_grad_scale_path(name, scale):
    """grad_scale = the expected upstream gradient: exact when it matches (2.0 = golden gy), rescaled when not"""
    g, o, f, loss, zp, gi, gr = _run(name, grad_scale=scale, return_new_zp=False)
    assert zp is None
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    assert_grad_close(gi.cpu().numpy(), g["g_img"])
    assert_grad_close(gr.cpu().numpy(), g["g_img_rot"])


def test_debug_tuple_and_free_functions():
    from rgbd_gan_b200.loss_functions import LossFuncRotate, bilinear, inv_warp, warp
    g = load_golden("loss_cfg0_l1_occ")
    o = case_options(g)
    B, S = o["B"], o["S"]
    x = torch.from_numpy(g["x"]).to(DEV)
    f = LossFuncRotate(None, lambda_geometric=o["lam"])
    out = f(x[:B], g["cam"][:B], x[B:], g["cam"][B:], debug=True)
    warped, not_out, new_zp, warped_rot, not_out_rot, new_zp_rot = out
    np.testing.assert_array_equal(warped.cpu().numpy(), g["warped"])
    np.testing.assert_array_equal(warped_rot.cpu().numpy(), g["warped_rot"])
    np.testing.assert_array_equal(not_out.cpu().numpy(), g["not_out"])
    np.testing.assert_array_equal(not_out_rot.cpu().numpy(), g["not_out_rot"])
    np.testing.assert_array_equal(torch.cat([new_zp, new_zp_rot]).cpu().numpy(), g["new_zp_cat"])
    # free functions with the reference's argument lists (loss_functions.py:89-94)
    cam = g["cam"]
    R1, R2 = cam[:B, :3, :3], cam[B:, :3, :3]
    t1, t2 = cam[:B, :3, -1:], cam[B:, :3, -1:]
    R = np.matmul(R2.transpose(0, 2, 1), R1).astype("float32")
    t = np.matmul(R1.transpose(0, 2, 1), t2 - t1).astype("float32")
    z = x[:B, -1:].reshape(B, 1, -1).clone().requires_grad_(True)
    z_rot = x[B:, -1:].reshape(B, 1, -1)
    zp = warp(f.K, f.inv_K, R, t, z, f.p)
    zp_rot = inv_warp(f.K, f.inv_K, R.transpose(0, 2, 1), t, z_rot, f.p)
    np.testing.assert_array_equal(torch.cat([zp, zp_rot]).detach().cpu().numpy(), g["new_zp_cat"])
    img_rot = x[B:].clone().requires_grad_(True)
    w, m = bilinear(img_rot, zp)
    np.testing.assert_array_equal(w.detach().cpu().numpy(), g["warped"])
    # composing the free functions and autograd gives the fused gradients (no occlusion variant)
    from oracle import numpy_port as npp
    gw = torch.from_numpy(np.random.default_rng(0).normal(size=tuple(w.shape)).astype(np.float32)).to(DEV)
    (w * gw).sum().backward()
    _, _, tape = npp.LossFuncRotateNP._bilinear_fwd(g["x"][B:], zp.detach().cpu().numpy())
    gi_ref, gzp_ref = npp.LossFuncRotateNP._bilinear_bwd(tape, gw.cpu().numpy())
    assert_grad_close(img_rot.grad.cpu().numpy(), gi_ref)
    port = npp.LossFuncRotateNP()
    port.init_params(S)
    M = port.pose_algebra(cam[:B], cam[B:])[0]
    gzp_ref[:, :, 1] = 0
    gz_ref = (np.matmul(M.transpose(0, 2, 1), gzp_ref.transpose(0, 2, 1)) * port.p).sum(axis=1, keepdims=True)
    assert_grad_close(z.grad.cpu().numpy(), gz_ref, tol=2e-5)


def test_growing_instance_state():
    from rgbd_gan_b200.loss_functions import LossFuncRotate
    g = load_golden("loss_growing")
    B = int(g["B"])
    f = LossFuncRotate(None)
    for S in (32, 64):
        x = torch.from_numpy(g["x%d" % S]).to(DEV)
        img, img_rot = x[:B].clone().requires_grad_(True), x[B:].clone().requires_grad_(True)
        loss, zp = f(img, g["cam"][:B], img_rot, g["cam"][B:], occlusion_aware=True)
        loss.backward()
        assert abs(loss.item() - float(g["loss_%d" % S])) <= 1e-5 * float(g["loss_%d" % S])
        np.testing.assert_array_equal(f.K, g["K_%d" % S])
        assert_grad_close(img.grad.cpu().numpy(), g["g_img_%d" % S])


@pytest.mark.parametrize("name", DV_CASES)
def test_projection_helper_like_generator(name):
    from rgbd_gan_b200.projection import ProjectionHelper, interpolate_trilinear
    g = load_golden(name)
    G, img, F, D = int(g["G"]), int(g["img"]), int(g["F"]), int(g["D"])
    h = ProjectionHelper(g["intrinsic"], g["intrinsic"], [img, img], [img, img], 0., 1., [G] * 3,
                         float(g["voxel_size"]), g["near_plane"], D, verbose=False)
    ns = g["cam"].shape[0]
    for i in range(ns):
        lin, vc = h.compute_proj_idcs(g["cam"][i])
        np.testing.assert_array_equal(lin.cpu().numpy(), g["lin_ind_%d" % i])
        np.testing.assert_array_equal(vc.cpu().numpy(), g["voxel_coords_%d" % i])
        grid = torch.from_numpy(g["grid"][i:i + 1]).to(DEV).requires_grad_(True)
        out = interpolate_trilinear(grid, lin, vc, [img, img], D)
        assert tuple(out.shape) == (1, F, D, img, img)
        np.testing.assert_array_equal(out.detach().cpu().numpy(), g["frustum_%d" % i])
        out.backward(torch.from_numpy(g["g_out"][i:i + 1]).to(DEV))
        assert_grad_close(grid.grad.cpu().numpy(), g["g_grid_%d" % i])
    grid = torch.from_numpy(g["grid"]).to(DEV).requires_grad_(True)
    fr = h.project(grid, g["cam"])
    fr.backward(torch.from_numpy(g["g_out"]).to(DEV))
    for i in range(ns):
        assert_frustum_close(fr[i].detach().cpu().numpy(), g["frustum_%d" % i][0], False)
        assert_grad_close(grid.grad[i].cpu().numpy(), g["g_grid_%d" % i][0])
    cam = g["cam"][0].copy()
    cam[:3, 3] += 100.0
    assert h.compute_proj_idcs(cam) is None


def test_calc_real_pos_matches_reference_formula():
    """LossFuncRotate.calc_real_pos (loss_functions.py:148-158): real_pos = (R K^-1)(z p) + t, concatenated
    with the RGB planes; no caller in the reference, plain array-library math here."""
    from rgbd_gan_b200.loss_functions import LossFuncRotate
    g = load_golden("loss_s64_l1_noocc")
    o = case_options(g)
    B, S = o["B"], o["S"]
    x = torch.from_numpy(g["x"]).to(DEV)
    f = LossFuncRotate(None)
    f.init_params(None, size=S)
    out = f.calc_real_pos(x[:B], g["cam"][:B]).cpu().numpy()
    R, t = g["cam"][:B, :3, :3], g["cam"][:B, :3, -1:]
    z = g["x"][:B, -1:].reshape(B, 1, -1)
    ref = np.matmul(np.matmul(R, f.inv_K), z * f.p) + t
    assert out.shape == (B, 6, S * S)
    np.testing.assert_array_equal(out[:, :3], g["x"][:B, :3].reshape(B, 3, -1))
    np.testing.assert_allclose(out[:, 3:], ref, rtol=1e-5, atol=1e-5)


def test_interpolate_trilinear_batch_and_errors():
    from rgbd_gan_b200.projection import ProjectionHelper, interpolate_trilinear
    g = load_golden("dv_g16_f3")
    G, img, F, D = int(g["G"]), int(g["img"]), int(g["F"]), int(g["D"])
    h = ProjectionHelper(g["intrinsic"], g["intrinsic"], [img, img], [img, img], 0., 1., [G] * 3,
                         float(g["voxel_size"]), g["near_plane"], D, verbose=False)
    lin, vc = h.compute_proj_idcs(g["cam"][0])
    grid = torch.from_numpy(g["grid"]).to(DEV)            # batch of 2 grids through the same index list
    out = interpolate_trilinear(grid, lin, vc, [img, img], D)
    np.testing.assert_array_equal(out[0].cpu().numpy(), g["frustum_0"][0])
    with pytest.raises(TypeError):
        interpolate_trilinear(grid.cpu(), lin, vc, [img, img], D)
    with pytest.raises(ValueError):
        interpolate_trilinear(grid[:, :, :, :, :8].contiguous(), lin, vc, [img, img], D)


@pytest.mark.parametrize("fuse", [True, False])
def test_depth_hinge_through_python_mirror(fuse):
    from rgbd_gan_b200.loss_functions import LossFuncRotate
    g = load_golden("hinge_ffhq")
    B, gy = int(g["B"]), float(g["gy"])
    x = torch.from_numpy(g["x"]).to(DEV)
    img, img_rot = x[:B].clone().requires_grad_(True), x[B:].clone().requires_grad_(True)
    f = LossFuncRotate(None, lambda_geometric=3, fuse_backward=fuse)
    loss, _ = f(img, g["cam"][:B], img_rot, g["cam"][B:], True,
                depth_hinge=(float(g["depth_min"]), float(g["lambda_depth"])))      # updater.py:340-359 in one call
    (loss * gy).backward()
    assert abs(loss.item() - float(g["total"])) <= 1e-5 * abs(float(g["total"]))
    assert_grad_close(img.grad.cpu().numpy(), g["g_img"])
    assert_grad_close(img_rot.grad.cpu().numpy(), g["g_img_rot"])


def test_peer_exchange_wait_is_bounded(monkeypatch):
    """a rank whose peers never show up (dead process, different call sequence) must get an error, not a hung GPU:
    world-2 comm on ONE GPU whose second mailbox is a local buffer nobody writes (rgbd_debug_peer_comm_loopback)"""
    import ctypes
    import time
    from gpu_util import Consistency
    from oracle import numpy_port as npp
    from rgbd_gan_b200 import _lib
    monkeypatch.setenv("RGBD_B200_PEER_TIMEOUT_MS", "50")
    lib = _lib.load()
    h = ctypes.c_void_p()
    ipc = ctypes.create_string_buffer(64)
    _lib.check(lib.rgbd_peer_comm_create(0, 2, ctypes.byref(h), ipc), "create")
    _lib.check(lib.rgbd_debug_peer_comm_loopback(h), "loopback")
    g = load_golden("loss_s64_l1_noocc")
    o = case_options(g)
    port = npp.LossFuncRotateNP(lambda_geometric=o["lam"])
    port.init_params(o["S"])
    drv = Consistency(g["x"], g["cam"], o["B"], port.K, port.inv_K, lam=o["lam"], occ=o["occ"], n_pairs_global=2 * o["B"])
    drv.opts.peer_comm = h.value
    t0 = time.perf_counter()
    parts, gi, gr = drv.fwd_bwd(gy=o["gy"])                       # returns: the wait gives up after 50 ms
    status = ctypes.c_int(-1)
    _lib.check(lib.rgbd_peer_comm_status(h, None, ctypes.byref(status)), "status")
    assert status.value == 1 and time.perf_counter() - t0 < 5.0
    assert_grad_close(gi * 2, g["g_img"])                          # gradients never depend on the exchange
    drv.opts.peer_comm = None
    torch.cuda.synchronize()
    lib.rgbd_peer_comm_destroy(h)


def test_publish_only_exchange_on_one_gpu(monkeypatch):
    """defer_loss == 2 (publish only) with a world-2 comm on ONE GPU whose peer mailbox is a local buffer nobody writes:
    the loss call itself never waits and leaves the shard's own parts; rgbd_peer_comm_wait (the collect) gives up after
    the time limit and raises the sticky flag; so does the flow control once the rank is 7 calls ahead of its silent peer"""
    import ctypes
    import time
    from gpu_util import Consistency
    from oracle import numpy_port as npp
    from rgbd_gan_b200 import _lib
    monkeypatch.setenv("RGBD_B200_PEER_TIMEOUT_MS", "20")
    lib = _lib.load()
    h = ctypes.c_void_p()
    ipc = ctypes.create_string_buffer(64)
    _lib.check(lib.rgbd_peer_comm_create(0, 2, ctypes.byref(h), ipc), "create")
    _lib.check(lib.rgbd_debug_peer_comm_loopback(h), "loopback")
    g = load_golden("loss_s64_l1_noocc")
    o = case_options(g)
    port = npp.LossFuncRotateNP(lambda_geometric=o["lam"])
    port.init_params(o["S"])
    drv = Consistency(g["x"], g["cam"], o["B"], port.K, port.inv_K, lam=o["lam"], occ=o["occ"])
    local, _, _ = drv.fwd_bwd(gy=o["gy"])                          # reference: the same shard without a comm
    drv.opts.peer_comm, drv.opts.defer_loss = h.value, 2
    status = ctypes.c_int(-1)
    t0 = time.perf_counter()
    for _ in range(6):                                             # fewer than the ring depth: nothing waits
        parts, gi, gr = drv.fwd_bwd(gy=o["gy"])
    _lib.check(lib.rgbd_peer_comm_status(h, None, ctypes.byref(status)), "status")
    assert status.value == 0 and time.perf_counter() - t0 < 2.0
    np.testing.assert_array_equal(parts[:7], local[:7])            # the shard's own values until the collect
    assert_grad_close(gi, g["g_img"])                              # gradients never depend on the exchange
    _lib.check(lib.rgbd_peer_comm_wait(h, None), "wait")           # collect: the peer never published
    _lib.check(lib.rgbd_peer_comm_status(h, None, ctypes.byref(status)), "status")
    assert status.value == 1
    for _ in range(4):                                             # now more than 7 ahead of the silent peer: bounded
        drv.fwd_bwd(gy=o["gy"])
    assert time.perf_counter() - t0 < 5.0
    drv.opts.peer_comm = None
    torch.cuda.synchronize()
    lib.rgbd_peer_comm_destroy(h)


def test_sharded_step_is_cuda_graph_capturable():
    """round-1 verdict: a sharded step must be capturable.  With the publish-only exchange (defer_loss == 2) a step is
    the same launch chain as on one GPU (no side stream, no events; the epoch lives in device memory): capture one
    fwd+bwd call of a world-2 comm (loopback peer) into a CUDA graph, replay it, compare with the direct call"""
    import ctypes
    from gpu_util import Consistency, p, stream
    from oracle import numpy_port as npp
    from rgbd_gan_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    ipc = ctypes.create_string_buffer(64)
    _lib.check(lib.rgbd_peer_comm_create(0, 2, ctypes.byref(h), ipc), "create")
    _lib.check(lib.rgbd_debug_peer_comm_loopback(h), "loopback")
    x, cam = npp.synthetic_batch(64, 128, depth="rough", seed=3)         # 64 pairs at 128x128: the row-sweep path
    port = npp.LossFuncRotateNP(lambda_geometric=3.0)
    port.init_params(128)
    drv = Consistency(x, cam, 64, port.K, port.inv_K, lam=3.0, occ=True)
    assert lib.rgbd_consistency_uses_sweep(64, 4, 128, 128) == 1
    ref_parts, ref_gi, ref_gr = drv.fwd_bwd(gy=2.0)
    drv.opts.peer_comm, drv.opts.defer_loss = h.value, 2
    parts = torch.zeros(8, device=DEV)
    g_img, g_rot = torch.zeros_like(drv.img), torch.zeros_like(drv.img_rot)
    s = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        def call():
            _lib.call("rgbd_consistency_fwd_bwd", *drv._common(), ctypes.c_float(2.0), p(parts), None, p(g_img), p(g_rot),
                      p(drv.ws), drv.ws.numel(), stream())
        call()                                                           # warm-up outside the capture (attributes, modules)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=s):
            call()
    for _ in range(3):
        g_img.zero_(); g_rot.zero_(); parts.zero_()
        graph.replay()
        torch.cuda.synchronize()
        np.testing.assert_array_equal(parts.cpu().numpy()[:7], ref_parts[:7])
        assert_grad_close(g_img.cpu().numpy(), ref_gi)
        assert_grad_close(g_rot.cpu().numpy(), ref_gr)
    status = ctypes.c_int(-1)
    _lib.check(lib.rgbd_peer_comm_status(h, None, ctypes.byref(status)), "status")
    assert status.value == 0                                             # 5 publishes: below the flow-control depth
    drv.opts.peer_comm = None
    lib.rgbd_peer_comm_destroy(h)


def test_occupancy_net_loss_and_calc_real_pos_smoke():
    """LossFuncRotate.calc_real_pos (:148-158) against the NumPy expression and occupancy_net_loss (:160-168) end to
    end with a stand-in occupancy network: finite scalar, gradient reaches the network and the depth"""
    from rgbd_gan_b200.loss_functions import LossFuncRotate
    g = load_golden("loss_s64_l1_noocc")
    o = case_options(g)
    B, S = o["B"], o["S"]
    x = torch.from_numpy(g["x"]).to(DEV)
    f = LossFuncRotate(None, lambda_geometric=o["lam"])
    f.init_params(None, size=S)
    theta = g["cam"][:B]
    out = f.calc_real_pos(x[:B], theta)
    z = g["x"][:B, -1:].reshape(B, 1, -1)
    ref = np.matmul(np.matmul(theta[:, :3, :3], f.inv_K), z * f.p) + theta[:, :3, -1:]
    assert out.shape == (B, 6, S * S)
    np.testing.assert_allclose(out[:, 3:].cpu().numpy(), ref, rtol=1e-5, atol=1e-5)
    np.testing.assert_array_equal(out[:, :3].cpu().numpy(), g["x"][:B, :3].reshape(B, 3, -1))

    class Occ(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.randn(3, device=DEV) * 0.1)

        def forward(self, zlat, pos):                                  # pos (B,3,HW) -> logits (B*HW,1)
            return (pos * self.w[None, :, None]).sum(1).reshape(-1, 1) + zlat.mean()

    net = Occ()
    depth = x[:B, -1:].clone().requires_grad_(True)
    loss = f.occupancy_net_loss(net, depth, theta, torch.randn(B, 8, device=DEV))
    assert loss.dim() == 0 and torch.isfinite(loss)
    loss.backward()
    assert torch.isfinite(net.w.grad).all() and float(net.w.grad.abs().sum()) > 0
    assert depth.grad is not None and torch.isfinite(depth.grad).all()


def test_depth_head_next_row():
    """depth = 1 / (softplus(h) + 1e-4) (net.py:294-299) as one kernel, values and gradient against
    the reference expression's golden vectors, in place and with an odd plane size (scalar path)"""
    import ctypes
    from rgbd_gan_b200 import _lib
    from rgbd_gan_b200.loss_functions import depth_head
    from oracle import numpy_port as npp
    g = load_golden("depth_head_s32")
    h = torch.from_numpy(g["h"]).to(DEV).requires_grad_(True)
    out = depth_head(h)
    (out * torch.from_numpy(g["g_out"]).to(DEV)).sum().backward()
    np.testing.assert_array_equal(out[:, :-1].detach().cpu().numpy(), g["out"][:, :-1])
    np.testing.assert_allclose(out[:, -1].detach().cpu().numpy(), g["out"][:, -1], rtol=1e-5)
    np.testing.assert_array_equal(h.grad[:, :-1].cpu().numpy(), g["g_h"][:, :-1])
    np.testing.assert_allclose(h.grad[:, -1].cpu().numpy(), g["g_h"][:, -1], rtol=1e-5, atol=1e-5 * np.abs(g["g_h"][:, -1]).max() * 1e-3)
    # in place + odd plane size (HW % 4 != 0 -> scalar path)
    x = (np.random.default_rng(1).normal(size=(2, 4, 5, 7)) * 3).astype(np.float32)
    t = torch.from_numpy(x).to(DEV)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.call("rgbd_depth_head_fwd", ctypes.c_void_p(t.data_ptr()), 2, 4, 5, 7, ctypes.c_void_p(t.data_ptr()), st)
    np.testing.assert_allclose(t.cpu().numpy(), npp.depth_head_fwd(x), rtol=1e-5)
