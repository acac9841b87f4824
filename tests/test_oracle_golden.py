"""CPU: both oracle restatements against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  Indices, masks, new_zp, warped values,
lin_ind, voxel_coords and the frustum are BIT-EXACT; loss and gradients within 1e-5."""
import numpy as np
import pytest

from conftest import DV_CASES, LOSS_CASES, assert_grad_close, case_options, load_golden
from oracle import numpy_port as npp


def _blas_is_fma_chain():
    """on this host NumPy's K=3 matmul is fma(a2,b2,fma(a1,b1,rn(a0*b0))).
    The golden vectors were generated on such a host; elsewhere compare new_zp with a tolerance."""
    rng = np.random.default_rng(0)
    a = rng.normal(size=(4, 3, 3)).astype(np.float32)
    b = rng.normal(size=(4, 3, 4096)).astype(np.float32)
    got = np.matmul(a, b)
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    t = (a[:, :, None, 0] * b[:, None, 0, :]).astype(np.float64)            # rn(a0*b0) in fp32
    t = (a64[:, :, None, 1] * b64[:, None, 1, :] + t).astype(np.float32).astype(np.float64)
    t = (a64[:, :, None, 2] * b64[:, None, 2, :] + t).astype(np.float32)
    return bool((got == t).mean() > 0.9999)


@pytest.mark.parametrize("name", LOSS_CASES)
def test_c_oracle_matches_reference(name, oracle_mod):
    g = load_golden(name)
    o = case_options(g)
    B, S = o["B"], o["S"]
    x, cam = g["x"], g["cam"]
    port = npp.LossFuncRotateNP(K=None if o["K"] is None else o["K"].copy(), norm=o["norm"], lambda_geometric=o["lam"])
    port.init_params(S)
    np.testing.assert_array_equal(port.K, g["K"])
    np.testing.assert_array_equal(port.inv_K, g["inv_K"])
    np.testing.assert_array_equal(port.p, g["p"])
    M, c, Mi, ci = port.pose_algebra(cam[:B], cam[B:])
    norm = 1 if o["norm"] == "l1" else 2
    parts, d = oracle_mod.consistency_fwd(x[:B], x[B:], M, c, Mi, -ci, norm=norm, occlusion=o["occ"],
                                          max_depth=o["max_depth"], min_depth=o["min_depth"], debug=True)
    N = B * S * S
    # bit-exact geometry, masks and sampled values
    np.testing.assert_array_equal(d["new_zp"], g["new_zp_cat"])
    np.testing.assert_array_equal(d["mask"][:N].astype(bool), g["not_out"])
    np.testing.assert_array_equal(d["mask"][N:].astype(bool), g["not_out_rot"])
    np.testing.assert_array_equal(d["warped"][:N], g["warped"])
    np.testing.assert_array_equal(d["warped"][N:], g["warped_rot"])
    if o["occ"]:
        zp = g["new_zp_cat"]
        np.testing.assert_array_equal(d["occ"][:N].astype(bool), g["warped"][:, -1] > zp[:B].reshape(-1, 3)[:, 2])
        np.testing.assert_array_equal(d["occ"][N:].astype(bool), g["warped_rot"][:, -1] > zp[B:].reshape(-1, 3)[:, 2])
    loss = oracle_mod.combine_loss(parts, o["lam"])
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    gi, gr = oracle_mod.consistency_bwd(x[:B], x[B:], M, c, Mi, -ci, norm=norm, occlusion=o["occ"],
                                        max_depth=o["max_depth"], min_depth=o["min_depth"],
                                        lambda_geometric=o["lam"], gy=o["gy"])
    assert_grad_close(gi, g["g_img"])
    assert_grad_close(gr, g["g_img_rot"])


@pytest.mark.parametrize("name", LOSS_CASES)
def test_numpy_port_matches_reference(name):
    g = load_golden(name)
    o = case_options(g)
    B = o["B"]
    x, cam = g["x"], g["cam"]
    port = npp.LossFuncRotateNP(K=None if o["K"] is None else o["K"].copy(), norm=o["norm"], lambda_geometric=o["lam"])
    loss, zp = port.forward(x[:B], cam[:B], x[B:], cam[B:], occlusion_aware=o["occ"], max_depth=o["max_depth"],
                            min_depth=o["min_depth"])
    if _blas_is_fma_chain():
        np.testing.assert_array_equal(zp, g["new_zp_cat"])
        np.testing.assert_array_equal(port.debug["warped"], g["warped"])
        np.testing.assert_array_equal(port.debug["not_out"], g["not_out"])
        np.testing.assert_array_equal(port.debug["not_out_rot"], g["not_out_rot"])
        assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
        gi, gr = port.backward(o["gy"])
        assert_grad_close(gi, g["g_img"])
        assert_grad_close(gr, g["g_img_rot"])
    else:   # another BLAS kernel: geometry may differ in the last bit, which can flip an index
        np.testing.assert_allclose(zp, g["new_zp_cat"], rtol=2e-6, atol=1e-5)


def test_growing_state_q9(oracle_mod):
    """one instance reused 32 -> 64: K is rescaled in place (loss_functions.py:52-54)"""
    g = load_golden("loss_growing")
    B = int(g["B"])
    port = npp.LossFuncRotateNP(lambda_geometric=3)
    for S in (32, 64):
        x = g["x%d" % S]
        loss, zp = port.forward(x[:B], g["cam"][:B], x[B:], g["cam"][B:], occlusion_aware=True)
        np.testing.assert_array_equal(port.K, g["K_%d" % S])
        np.testing.assert_array_equal(port.inv_K, g["inv_K_%d" % S])
        assert abs(float(loss) - float(g["loss_%d" % S])) <= 1e-5 * abs(float(g["loss_%d" % S]))
        M, c, Mi, ci = port.pose_algebra(g["cam"][:B], g["cam"][B:])
        parts, d = oracle_mod.consistency_fwd(x[:B], x[B:], M, c, Mi, -ci, norm=1, occlusion=True, debug=True)
        np.testing.assert_array_equal(d["new_zp"], g["new_zp_cat_%d" % S])
        gi, gr = oracle_mod.consistency_bwd(x[:B], x[B:], M, c, Mi, -ci, norm=1, occlusion=True, lambda_geometric=3, gy=1.0)
        assert_grad_close(gi, g["g_img_%d" % S])
        assert_grad_close(gr, g["g_img_rot_%d" % S])


def test_standalone_warp_bilinear(oracle_mod):
    """warp / inv_warp / bilinear as separate functions reproduce the fused intermediates"""
    g = load_golden("loss_s64_l1_noocc")
    o = case_options(g)
    B, S = o["B"], o["S"]
    x, cam = g["x"], g["cam"]
    port = npp.LossFuncRotateNP(lambda_geometric=o["lam"])
    port.init_params(S)
    M, c, Mi, ci = port.pose_algebra(cam[:B], cam[B:])
    zp = oracle_mod.warp_fwd(x[:B, -1].reshape(B, -1), M, c, S, S)
    zp_rot = oracle_mod.warp_fwd(x[B:, -1].reshape(B, -1), Mi, -ci, S, S)
    np.testing.assert_array_equal(np.concatenate([zp, zp_rot]), g["new_zp_cat"])
    warped, mask = oracle_mod.bilinear_fwd(x[B:], zp)
    np.testing.assert_array_equal(warped, g["warped"])
    np.testing.assert_array_equal(mask, g["not_out"])
    # bilinear backward against the NumPy port's node-by-node reverse pass
    rng = np.random.default_rng(1)
    gw = rng.normal(size=warped.shape).astype(np.float32)
    _, _, tape = npp.LossFuncRotateNP._bilinear_fwd(x[B:], zp)
    gi_ref, gzp_ref = npp.LossFuncRotateNP._bilinear_bwd(tape, gw)
    gi, gzp = oracle_mod.bilinear_bwd(x[B:], zp, gw)
    assert_grad_close(gi, gi_ref)
    sc = np.abs(gzp_ref).max()
    assert np.abs(gzp[:, :, [0, 2]] - gzp_ref[:, :, [0, 2]]).max() <= 1e-5 * sc
    assert np.abs(gzp_ref[:, :, 1]).max() <= 1e-4 * sc        # row-coordinate gradient is rounding noise (Q2)
    gz = oracle_mod.warp_bwd(gzp_ref, M, S, S)
    gP = np.matmul(M.transpose(0, 2, 1), gzp_ref.transpose(0, 2, 1))
    np.testing.assert_allclose(gz, (gP * port.p).sum(axis=1, keepdims=True), rtol=1e-5, atol=1e-5 * np.abs(gz).max())


@pytest.mark.parametrize("name", DV_CASES)
def test_deepvoxels_oracle_matches_reference(name, oracle_mod):
    g = load_golden(name)
    G, img, F, D = int(g["G"]), int(g["img"]), int(g["F"]), int(g["D"])
    P = oracle_mod.dv_params(img, img, D, G, g["intrinsic"], float(g["voxel_size"]), float(g["near_plane"]))
    helper = npp.ProjectionHelperNP(g["intrinsic"], [img, img], [G] * 3, float(g["voxel_size"]), g["near_plane"], D)
    ns = g["cam"].shape[0]
    for i in range(ns):
        lin, vc = oracle_mod.dv_compute_proj_idcs(P, g["cam"][i])
        np.testing.assert_array_equal(lin, g["lin_ind_%d" % i])
        np.testing.assert_array_equal(vc, g["voxel_coords_%d" % i])
        lin2, vc2 = helper.compute_proj_idcs(g["cam"][i])
        np.testing.assert_array_equal(lin2, lin)
        np.testing.assert_allclose(vc2, vc, rtol=1e-6, atol=1e-5)
        out = oracle_mod.dv_trilinear_fwd(g["grid"][i], lin, vc, P)
        np.testing.assert_array_equal(out, g["frustum_%d" % i][0])
        gg = oracle_mod.dv_trilinear_bwd(g["g_out"][i], lin, vc, P)
        assert_grad_close(gg, g["g_grid_%d" % i][0])
        out2 = npp.interpolate_trilinear_fwd(g["grid"][i:i + 1], lin, vc, [img, img], D)
        np.testing.assert_array_equal(out2[0], out)
        gg2 = npp.interpolate_trilinear_bwd(g["grid"][i:i + 1].shape, lin, vc, g["g_out"][i:i + 1])
        assert_grad_close(gg2[0], g["g_grid_%d" % i][0])
    fused = oracle_mod.dv_project_fwd(P, g["grid"], g["cam"])
    gfused = oracle_mod.dv_project_bwd(P, g["g_out"], g["cam"])
    for i in range(ns):
        np.testing.assert_array_equal(fused[i], g["frustum_%d" % i][0])
        assert_grad_close(gfused[i], g["g_grid_%d" % i][0])


def test_deepvoxels_grid2world_argument(oracle_mod):
    """the optional second argument of compute_proj_idcs (projection.py:48,53-54,83-84): two K = 4 products per element;
    both restatements reproduce the reference's lin_ind / voxel_coords bit for bit, the folded single product does not"""
    from oracle import numpy_port as npp
    g = load_golden("dv_g16_f3")
    G, img, D = int(g["G"]), int(g["img"]), int(g["D"])
    P = oracle_mod.dv_params(img, img, D, G, g["intrinsic"], float(g["voxel_size"]), float(g["near_plane"]))
    lin, vc = oracle_mod.dv_compute_proj_idcs(P, g["cam"][0], g["grid2world"])
    np.testing.assert_array_equal(lin, g["lin_ind_g2w_0"])
    np.testing.assert_array_equal(vc, g["voxel_coords_g2w_0"])
    h = npp.ProjectionHelperNP(g["intrinsic"], [img, img], [G] * 3, float(g["voxel_size"]), float(g["near_plane"]), D)
    lin2, vc2 = h.compute_proj_idcs(g["cam"][0], g["grid2world"])
    np.testing.assert_array_equal(lin2, g["lin_ind_g2w_0"])
    np.testing.assert_array_equal(vc2, g["voxel_coords_g2w_0"])
    folded = np.dot(np.linalg.inv(g["grid2world"]), g["cam"][0]).astype("float32")
    _, vc3 = oracle_mod.dv_compute_proj_idcs(P, folded)
    assert vc3.shape != vc.shape or not np.array_equal(vc3, vc)


def test_dv_empty_frustum(oracle_mod):
    """camera far outside the grid: the reference prints an error and returns None (projection.py:98-100)"""
    g = load_golden("dv_g16_f3")
    G, img, D = int(g["G"]), int(g["img"]), int(g["D"])
    P = oracle_mod.dv_params(img, img, D, G, g["intrinsic"], float(g["voxel_size"]), float(g["near_plane"]))
    cam = g["cam"][0].copy()
    cam[:3, 3] += 100.0
    assert oracle_mod.dv_compute_proj_idcs(P, cam) is None
    helper = npp.ProjectionHelperNP(g["intrinsic"], [img, img], [G] * 3, float(g["voxel_size"]), g["near_plane"], D)
    assert helper.compute_proj_idcs(cam) is None
    fr = oracle_mod.dv_project_fwd(P, g["grid"][:1], cam[None])
    assert not fr.any()


@pytest.mark.parametrize("name", ["hinge_ffhq", "hinge_car"])
def test_depth_hinge_next_row(name, oracle_mod):
    """loss_rotate += mean(relu(depth_min - depth)^2) * lambda_depth (updater.py:357-359)"""
    g = load_golden(name)
    B, S, gy = int(g["B"]), int(g["S"]), float(g["gy"])
    dmin, lam = float(g["depth_min"]), float(g["lambda_depth"])
    x, cam = g["x"], g["cam"]
    port = npp.LossFuncRotateNP(lambda_geometric=3)
    port.init_params(S)
    M, c, Mi, ci = port.pose_algebra(cam[:B], cam[B:])
    parts = oracle_mod.consistency_fwd(x[:B], x[B:], M, c, Mi, -ci, norm=1, occlusion=True)
    gi, gr = oracle_mod.consistency_bwd(x[:B], x[B:], M, c, Mi, -ci, norm=1, occlusion=True, lambda_geometric=3, gy=gy)
    hinge = oracle_mod.depth_hinge(x[:B], x[B:], dmin, lam, gy=gy, g_img=gi, g_img_rot=gr)
    assert abs(hinge - float(g["hinge"])) <= 1e-6 * max(abs(float(g["hinge"])), 1e-12)
    total = np.float32(oracle_mod.combine_loss(parts, 3)) + np.float32(hinge)
    assert abs(float(total) - float(g["total"])) <= 1e-5 * abs(float(g["total"]))
    assert_grad_close(gi, g["g_img"])
    assert_grad_close(gr, g["g_img_rot"])
    val, gx = npp.depth_hinge(x, dmin, lam, gy=gy)
    assert abs(float(val) - float(g["hinge"])) <= 1e-6 * max(abs(float(g["hinge"])), 1e-12)
    loss, _ = port.forward(x[:B], cam[:B], x[B:], cam[B:], occlusion_aware=True)
    gi2, gr2 = port.backward(gy)
    assert_grad_close(gi2 + gx[:B], g["g_img"])
    assert_grad_close(gr2 + gx[B:], g["g_img_rot"])


RENDER_CASES = ["render_g16", "render_g12_thr3"]


@pytest.mark.parametrize("name", RENDER_CASES)
def test_render_tail_oracle_against_reference_golden(name, oracle_mod):
    """next row: the closed-form NumPy restatement of the accumulative render tail against vectors
    produced by the reference's own interpolate_trilinear + AccumulativeOcclusionNet.forward over the shim"""
    g = load_golden(name)
    G, img, D = int(g["G"]), int(g["img"]), int(g["D"])
    P = oracle_mod.dv_params(img, img, D, G, g["intrinsic"], float(g["voxel_size"]), float(g["near_plane"]))
    args = (P, g["grid"], g["cam"], g["W1"], g["b1"], g["W2"], g["b2"], float(g["threshold"]), float(g["inv_c1"]),
            float(g["inv_c2"]), D)
    novel, depth, fg = oracle_mod.dv_render_fwd(*args)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    assert rel(novel, g["novel"]) <= 1e-5
    assert rel(depth.reshape(g["depth"].shape), g["depth"]) <= 1e-6
    assert rel(fg.reshape(g["fg"].shape), g["fg"]) <= 1e-5
    gg, gW1, gb1, gW2, gb2 = oracle_mod.dv_render_bwd(*args, g["g_novel"], g["g_depth"], g["g_fg"])
    for got, key in ((gg, "g_grid"), (gW1, "g_W1"), (gb1, "g_b1"), (gW2, "g_W2"), (gb2, "g_b2")):
        assert rel(got.reshape(g[key].shape), g[key]) <= 1e-5, key


def test_depth_head_port_matches_reference_expression():
    """next row: net.py:294-299 evaluated over the shim vs the op-by-op NumPy port"""
    g = load_golden("depth_head_s32")
    np.testing.assert_array_equal(npp.depth_head_fwd(g["h"]), g["out"])
    np.testing.assert_allclose(npp.depth_head_bwd(g["h"], g["g_out"]), g["g_h"], rtol=1e-6, atol=1e-12)
