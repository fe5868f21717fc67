"""Pin oracle/chore_oracle.py to the golden vectors that oracle/make_golden.py produced by
running the REAL reference (CPU tests; no GPU).  When /root/reference is present (build
container) the oracle is additionally compared with the live reference modules."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import chore_oracle as O
from oracle import ref_shim
from oracle.make_golden import checksum

TOL = 2e-5   # oracle vs golden: same torch CPU ops, possibly another CPU ISA / thread split


def T(x):
    return torch.from_numpy(np.asarray(x))


@pytest.fixture(scope="module")
def sd():
    return O.make_state_dict(0, "unit")


def test_state_dict_census(sd):
    assert len(sd) == 561
    assert sd["image_filter.conv1.weight"].shape == (64, 5, 7, 7)
    assert sd["df.0.weight"].shape == (128, 323, 1) and sd["part_predictor.6.weight"].shape == (14, 128, 1)
    n_enc = sum(v.numel() for k, v in sd.items() if k.startswith("image_filter") and ".downsample.0." not in k)
    assert abs(n_enc - 17.95e6) < 0.05e6


def test_query_vs_golden(sd):
    g = load_golden("query.npz")
    feat, tmpx = O.synth_features(int(g["seed"]), B=2)
    np.testing.assert_allclose(checksum(feat), g["feat_ck"], rtol=1e-9)
    np.testing.assert_allclose(checksum(tmpx), g["tmpx_ck"], rtol=1e-9)
    np.testing.assert_allclose(checksum(sd["df.0.weight"]), g["w_ck"], rtol=1e-9)
    pts, cc = T(g["points"]), T(g["crop_center"])
    df, pca, parts, centers, in_img = O.query(sd, feat, tmpx, pts, cc)
    proj = O.project_points(pts, cc)
    assert torch.equal(proj, T(g["proj"])) or np.array_equal(proj.numpy(), g["proj"], equal_nan=True)
    for name, got in (("df", df), ("pca", pca), ("parts", parts), ("centers", centers)):
        assert rel_err(got, g[name]) < TOL, name
    assert torch.equal(parts.argmax(1), T(g["parts"]).argmax(1))
    assert torch.equal(df[:, 0] == O.OUT_DIST, ~in_img)
    ga = O.query_grad_points(sd, feat, tmpx, pts, cc, T(g["g_df"]), T(g["g_pca"]), T(g["g_parts"]), T(g["g_centers"]))
    assert rel_err(ga, g["grad_all"]) < 5e-5


def test_query_numpy_restatement(sd):
    """Second, torch-free restatement agrees with the torch one (guards ATen-semantics misreadings)."""
    feat, tmpx = O.synth_features(5, B=1, hw=16)
    pts = O.synth_points("init_box", 6, 1, 48)
    pts[0, :24] = O.synth_points("frustum", 7, 1, 24)[0]
    cc = torch.tensor([[1008., 995.]])
    a = O.query(sd, feat, tmpx, pts, cc)
    b = O.query_numpy(sd, feat.numpy(), tmpx.numpy(), pts.numpy(), cc.numpy())
    for x, y in zip(a[:4], b[:4]):
        assert rel_err(T(y), x) < 1e-4
    assert np.array_equal(a[4].numpy(), b[4])


def test_encoder_vs_golden(sd):
    g = load_golden("encoder_128.npz")
    img = O.synth_images(int(g["seed"]), B=2, size=128)
    np.testing.assert_allclose(checksum(img), g["img_ck"], rtol=1e-9)
    with torch.no_grad():
        outs, tmpx, normx = O.hg_filter(sd, img)
    assert rel_err(outs[-1], g["feat"]) < 1e-4
    assert rel_err(tmpx, g["tmpx"]) < TOL
    assert rel_err(normx[:, :, ::4, ::4], g["normx"]) < TOL
    g2 = load_golden("encoder_128_refinit.npz")
    sd2 = O.make_state_dict(int(g2["weights_seed"]), "ref_init")
    with torch.no_grad():
        f2, t2 = O.encode(sd2, O.synth_images(int(g2["seed"]), B=1, size=128))
    assert rel_err(f2, g2["feat"]) < 1e-4 and rel_err(t2, g2["tmpx"]) < TOL


def test_approx_surface_vs_golden(sd):
    g = load_golden("approx_surface.npz")
    feat, tmpx = O.synth_features(int(g["seed"]), B=1)
    cc = T(g["crop_center"])
    for k, name in enumerate(("human", "object")):
        s0 = O.synth_points("frustum", 32, 1, 512)
        samples, preds = O.approx_surface(sd, feat, tmpx, s0, cc, 10, k)
        # ten chained gradient steps amplify rounding; compare in absolute metres
        assert (samples - T(g[f"samples_{name}"])).abs().max() < 1e-3
        assert (samples - T(g[f"samples_{name}"])).abs().median() < 1e-5


def test_lbs_vs_golden():
    g = load_golden("lbs.npz")
    buf = O.make_smplh_buffers(int(g["buffers_seed"]))
    np.testing.assert_allclose(checksum(buf["posedirs"]), g["posedirs_ck"], rtol=1e-9)
    pose, betas, trans, offs = (T(g[k]).clone().requires_grad_(True) for k in ("pose", "betas", "trans", "offsets"))
    verts, jtr, v_posed, naked = O.lbs_forward(buf, pose, betas, trans, offs)
    assert rel_err(verts, g["verts"]) < TOL and rel_err(jtr, g["jtr"]) < TOL
    assert rel_err(v_posed[:, ::10], g["v_posed_s"]) < TOL and rel_err(naked[:, ::10], g["naked_s"]) < TOL
    ((T(g["g_verts"]) * verts).sum() + (T(g["g_jtr"]) * jtr).sum()).backward()
    assert rel_err(pose.grad, g["grad_pose"]) < 1e-4
    assert rel_err(betas.grad, g["grad_betas"]) < 1e-4
    assert rel_err(trans.grad, g["grad_trans"]) < 1e-4


def test_rigid_and_fit_vs_golden(sd):
    g = load_golden("rigid.npz")
    R = O.project_so3(T(g["rot"]))
    assert rel_err(R, g["R"]) < TOL
    assert rel_err(O.transform_obj_verts(T(g["obj"]), R, T(g["t"]), T(g["s"])), g["moved"]) < TOL
    assert rel_err(O.project_so3(T(g["bad"])), g["R_bad"]) < TOL
    eye = torch.eye(3)
    assert (torch.bmm(R, R.transpose(1, 2)) - eye).abs().max() < 1e-5 and (torch.det(R) - 1).abs().max() < 1e-5

    f = load_golden("fit_object_only.npz")
    feat, tmpx = O.synth_features(int(f["seed"]), B=2)
    rot, t, s = (T(f[k]).clone().requires_grad_(True) for k in ("rot", "t", "s"))
    Rn = O.decopose_axis(rot, T(f["noise"]))
    losses = O.object_only_losses(sd, feat, tmpx, T(f["crop_center"]), T(f["obj"]), Rn, t, s, T(f["smpl_center"]))
    for k in ("object", "scale", "ocent"):
        assert rel_err(losses[k], f[f"loss_{k}"]) < TOL, k
    total = O.sum_dict(losses, float(f["it"]))
    assert rel_err(total, f["total"]) < TOL
    total.backward()
    assert rel_err(rot.grad, f["grad_rot"]) < 2e-3      # through torch.svd backward: ill-conditioned
    assert rel_err(t.grad, f["grad_t"]) < 1e-4 and rel_err(s.grad, f["grad_s"]) < 1e-4


def test_create_grid_semantics():
    c = O.create_grid((4, 3, 2), [-3, -0.9, 0.2], [3, 1.8, 4.0])
    assert c.shape == (3, 24)
    np.testing.assert_allclose(c[:, 0], [-3, -0.9, 0.2])
    np.testing.assert_allclose(c[:, 1], [-3, -0.9, 0.2 + 3.8 / 2])      # z fastest (np.mgrid order)
    np.testing.assert_allclose(c[:, -1], [-3 + 6 * 3 / 4, -0.9 + 2.7 * 2 / 3, 0.2 + 3.8 / 2])


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_oracle_equals_live_reference(sd):
    net, _ = ref_shim.load_chore()
    net.load_state_dict(sd)
    img = O.synth_images(3, B=1, size=128)
    with torch.no_grad():
        net.filter(img)
        f, t = O.encode(sd, img)
    assert torch.equal(f, net.im_feat_list[-1]) and torch.equal(t, net.tmpx)
    pts, cc = O.synth_points("init_box", 4, 1, 256), torch.tensor([[1008., 995.]])
    feat, tmpx = O.synth_features(8, B=1, hw=32)
    net.im_feat_list, net.tmpx = [feat], tmpx
    with torch.no_grad():
        net.query(pts, crop_center=cc)
    for a, b in zip(net.get_preds(), O.query(sd, feat, tmpx, pts, cc)[:4]):
        assert torch.equal(a, b)


def test_fit_smpl_full_vs_golden(sd):
    """forward_smpl(phase='kpts') with every term + sum_dict + backward to the split SMPL parameters:
    the oracle restatement against the reference's own run (tests/golden/fit_smpl_full.npz)."""
    from conftest import golden_smpl_assets
    g = load_golden("fit_smpl_full.npz")
    regs, pri = golden_smpl_assets(g)
    regs_t = [T(r.toarray()).float() for r in regs]
    pri_t = {k: T(v).float() for k, v in pri.items()}
    buf = O.make_smplh_buffers(int(g["buffers_seed"]))
    feat, tmpx = O.synth_features(int(g["feat_seed"]), B=2)
    pose, betas, trans = (T(g[k]).clone().requires_grad_(True) for k in ("pose", "betas", "trans"))
    losses = O.smpl_full_losses(sd, feat, tmpx, T(g["crop_center"]), buf, pose, betas, trans, T(g["part_labels"]),
                                T(g["pose_init"]), regs_t, pri_t, body_kpts=T(g["body_kpts"]))
    assert list(losses) == [str(k) for k in g["loss_order"]]
    for k, v in losses.items():
        assert rel_err(v, g[f"loss_{k}"]) < TOL, k
    total = O.sum_dict(losses, float(g["decay"]))
    assert rel_err(total, g["total"]) < TOL
    total.backward()
    gp = pose.grad
    for name, got in (("trans", trans.grad), ("global_pose", gp[:, :3]), ("body_pose", gp[:, 3:66]), ("hand_pose", gp[:, 66:]),
                      ("top_betas", betas.grad[:, :2]), ("other_betas", betas.grad[:, 2:])):
        assert rel_err(got, g[f"grad_{name}"]) < 2e-4, name
    # landmarks = sparse regressors applied to the posed vertices (wrapper_pytorch.py:176-190)
    verts = O.lbs_forward(buf, T(g["pose"]), T(g["betas"]), T(g["trans"]))[0]
    for r, name in zip(regs_t, ("J", "face", "hands")):
        assert rel_err(torch.matmul(r, verts), g[name]) < TOL, name


def test_lbs_numpy_restatement():
    """Second, torch-free LBS (float64) agrees with the torch restatement and with the reference's own run."""
    g = load_golden("lbs.npz")
    buf = O.make_smplh_buffers(int(g["buffers_seed"]))
    v, j = O.lbs_numpy(buf, g["pose"], g["betas"], g["trans"], g["offsets"])
    assert rel_err(T(v), g["verts"]) < 2e-5 and rel_err(T(j), g["jtr"]) < 2e-5
    vt, jt, _, _ = O.lbs_forward(buf, T(g["pose"]), T(g["betas"]), T(g["trans"]), T(g["offsets"]))
    assert rel_err(T(v), vt) < 2e-5 and rel_err(T(j), jt) < 2e-5
