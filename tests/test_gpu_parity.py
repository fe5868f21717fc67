"""GPU parity tests: the CUDA path (through the C ABI via chore_b200) against the oracle and the
golden vectors generated from the live reference (tests/golden, oracle/make_golden.py).

Bars (BASELINE.json north_star): distance / centre / PCA fields within 1e-4 relative
(conftest.rel_err: max |a-b| / (|b| + rms(b))); in-image masks and texel indices bit-exact;
part-label argmax exact wherever the reference's own top-2 logit margin exceeds the fp32
summation-order noise (the golden vectors contain margins down to 8e-6).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden, pure_rel_err, rel_err, report
from oracle import chore_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4
CHAINED_TOL = 5e-2          # chained parameter gradients (see the module docstring); measured values go to the parity report
ENC_TOL = 1e-4      # encoder output after ~60 GroupNorm + conv layers: measured 1.1e-5 .. 3.4e-5 (gpurun_out/parity_report.jsonl)
E2E_TOL = 1e-4      # encoder + query end to end: measured <= 3.2e-5
DEV = "cuda:0"


def T(x, dev=DEV):
    return torch.from_numpy(np.asarray(x)).to(dev)


@pytest.fixture(scope="module")
def sd():
    return O.make_state_dict(0, "unit")


@pytest.fixture(scope="module")
def net(sd):
    import chore_b200
    n = chore_b200.CHORE(device=DEV)
    n.load_state_dict(sd)
    return n.eval()


def set_maps(net, feat, tmpx):
    net.im_feat_list, net.tmpx = [feat.to(DEV)], tmpx.to(DEV)


def argmax_agrees(got, ref, margin=5e-4):
    """argmax must agree wherever the reference's top-2 margin exceeds `margin`."""
    got, ref = got.cpu(), ref.cpu()
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > margin
    same = got.argmax(1) == ref.argmax(1)
    assert bool(same[clear].all()), f"{int((~same & clear).sum())} clear-margin labels differ"
    return int((~same).sum()), int((~clear).sum())


def grad_close(got, want, tol=2e-4, frac=0.9995, worst=5e-2):
    """Gradients to the points: a hidden unit whose pre-activation is within rounding of zero can
    flip its ReLU mask under a different fp32 summation order, which changes that ONE point's
    gradient discretely (61 M hidden units are evaluated for 20 k points).  So: all but a
    vanishing fraction of the points within `tol`, and no point off by more than `worst`."""
    got, want = got.detach().double().cpu().reshape(-1, 3), want.detach().double().cpu().reshape(-1, 3)
    scale = want.abs().amax(-1) + want.pow(2).mean().sqrt() + 1e-30
    err = (got - want).abs().amax(-1) / scale
    assert (err < tol).double().mean().item() >= frac, ((err < tol).double().mean().item(), err.max().item())
    assert err.max().item() < worst, err.max().item()


# ------------------------------------------------------------------------------------------------
# point query
# ------------------------------------------------------------------------------------------------
def test_query_fwd_vs_golden(net):
    g = load_golden("query.npz")
    feat, tmpx = O.synth_features(int(g["seed"]), B=2)
    set_maps(net, feat, tmpx)
    pts, cc = T(g["points"]), T(g["crop_center"])
    outs, in_img = net.handle.query_fwd(*net._maps(), pts, cc, 15, want_in_img=True)
    df, pca, parts, centers = outs
    for name, got in (("df", df), ("pca", pca.view(2, 3, 3, -1)), ("parts", parts), ("centers", centers)):
        assert rel_err(got, g[name]) < TOL, (name, rel_err(got, g[name]))
    # bit-exact image-plane mask: the projection reproduces the reference's fp32 op order
    proj = torch.from_numpy(g["proj"])
    ref_in = (proj[:, 0] >= -1) & (proj[:, 0] <= 1) & (proj[:, 1] >= -1) & (proj[:, 1] <= 1)
    assert torch.equal(in_img.cpu().bool(), ref_in)
    assert torch.equal(df[:, 0].cpu() == O.OUT_DIST, ~ref_in)
    n_diff, n_unclear = argmax_agrees(parts, torch.from_numpy(g["parts"]))
    assert n_diff <= n_unclear


def test_query_module_interface_and_grad_vs_golden(net):
    g = load_golden("query.npz")
    feat, tmpx = O.synth_features(int(g["seed"]), B=2)
    set_maps(net, feat, tmpx)
    cc = T(g["crop_center"])
    p = T(g["points"]).clone().requires_grad_(True)
    net.query(p, crop_center=cc)
    df, pca, parts, centers = net.get_preds()
    assert pca.shape == (2, 3, 3, 2048) and df.shape == (2, 2, 2048)
    loss = (T(g["g_df"]) * df).sum() + (T(g["g_pca"]) * pca).sum() + (T(g["g_parts"]) * parts).sum() + \
        (T(g["g_centers"]) * centers).sum()
    loss.backward()
    grad_close(p.grad, T(g["grad_all"]))
    # the approx_surface pattern (recon/generator.py:63-70): clamp(df_h, max=2).sum().backward()
    p2 = T(g["points"]).clone().requires_grad_(True)
    net.query(p2, crop_center=cc)
    torch.clamp(net.get_preds()[0][:, 0, :], max=2.0).sum().backward()
    grad_close(p2.grad, T(g["grad_dfh"]))


@pytest.mark.parametrize("B,N,kind", [(1, 1, "frustum"), (1, 31, "init_box"), (3, 33, "frustum"), (2, 100, "init_box"),
                                      (1, 2049, "frustum"), (2, 20000, "init_box")])
def test_query_ragged_sizes_vs_oracle(net, sd, B, N, kind):
    feat, tmpx = O.synth_features(100 + N, B=B, hw=32)
    set_maps(net, feat, tmpx)
    cc = torch.tensor([[1008., 995.]]).repeat(B, 1) + torch.arange(B).float().view(B, 1) * 7
    pts = O.synth_points(kind, N, B, N, cc)
    with torch.no_grad():
        ref = O.query(sd, feat, tmpx, pts, cc)
    outs, in_img = net.handle.query_fwd(*net._maps(), pts.to(DEV), cc.to(DEV), 15, want_in_img=True)
    for name, got, want in zip(("df", "pca", "parts", "centers"), outs, ref[:4]):
        assert rel_err(got, want.reshape(got.shape)) < TOL, (name, rel_err(got, want.reshape(got.shape)))
    assert torch.equal(in_img.cpu().bool(), ref[4])
    g_df = torch.randn(B, 2, N, generator=torch.Generator().manual_seed(N))
    g_c = torch.randn(B, 6, N, generator=torch.Generator().manual_seed(N + 1))
    want_g = O.query_grad_points(sd, feat, tmpx, pts, cc, g_df=g_df, g_centers=g_c)
    got_g = net.handle.query_bwd(*net._maps(), pts.to(DEV), cc.to(DEV), [g_df.to(DEV), None, None, g_c.to(DEV)])
    grad_close(got_g, want_g)


def test_query_empty_and_head_mask(net):
    feat, tmpx = O.synth_features(7, B=1, hw=16)
    set_maps(net, feat, tmpx)
    cc = torch.tensor([[1008., 995.]], device=DEV)
    outs, _ = net.handle.query_fwd(*net._maps(), torch.zeros(1, 0, 3, device=DEV), cc, 15)
    assert outs[0].shape == (1, 2, 0)
    pts = O.synth_points("frustum", 3, 1, 257).to(DEV)
    full, _ = net.handle.query_fwd(*net._maps(), pts, cc, 15)
    only_df, _ = net.handle.query_fwd(*net._maps(), pts, cc, 1)
    assert only_df[1] is None and torch.equal(only_df[0], full[0])


def test_query_grid_vs_oracle(net, sd):
    """model/sdf.py:4-48 semantics: coordinates generated in-kernel == create_grid + query."""
    feat, tmpx = O.synth_features(9, B=2, hw=32)
    set_maps(net, feat, tmpx)
    res, bmin, bmax = (12, 10, 9), [-3.0, -0.9, 0.2], [3.0, 1.8, 4.0]
    coords = torch.from_numpy(O.create_grid(res, bmin, bmax).T.astype(np.float32)).unsqueeze(0)   # (1, XYZ, 3)
    cc = torch.tensor([[1008., 995.], [990., 1010.]])
    with torch.no_grad():
        ref = O.query(sd, feat[1:2], tmpx[1:2], coords, cc[1:2])
    outs = net.query_grid(res, bmin, bmax, cc, batch_index=1, head_mask=15, chunk=500)
    for got, want in zip(outs, ref[:4]):
        assert rel_err(got, want.reshape(got.shape)) < TOL
    # chunking does not change a single bit
    outs2 = net.query_grid(res, bmin, bmax, cc, batch_index=1, head_mask=15, chunk=1 << 20)
    for a, b in zip(outs, outs2):
        assert torch.equal(a, b)


def test_query_permutation_equivariance_large(net):
    """Size-independent property at production size: permuting 1M points permutes the outputs."""
    feat, tmpx = O.synth_features(11, B=1)
    set_maps(net, feat, tmpx)
    N = 1 << 20
    cc = torch.tensor([[1008., 995.]], device=DEV)
    pts = O.synth_points("init_box", 5, 1, N).to(DEV)
    perm = torch.randperm(N, generator=torch.Generator().manual_seed(1)).to(DEV)
    a, ia = net.handle.query_fwd(*net._maps(), pts, cc, 15, want_in_img=True)
    b, ib = net.handle.query_fwd(*net._maps(), pts[:, perm].contiguous(), cc, 15, want_in_img=True)
    for x, y in zip(a, b):
        assert torch.equal(x[:, :, perm], y)
    assert torch.equal(ia[:, perm], ib)
    frac_in = ia.float().mean().item()
    assert 0.1 < frac_in < 0.5        # the init box puts ~24 % of the points in the image


def test_approx_surface_vs_golden(net, sd):
    """Generator.approx_surface (recon/generator.py:50-79).

    The 10-step chain is chaotic on the white-noise test features: the ORACLE ITSELF, restarted
    from points perturbed by 1e-7 relative, ends with a median deviation of 7e-7 m (human field)
    / 4e-3 m (object field) and only 87 % / 47 % of the samples within 1 mm (measured on CPU).
    So the golden end points are compared loosely, and the trajectory is checked tightly step by
    step with teacher forcing: every CUDA step starts from the oracle's samples of that step."""
    import chore_b200
    g = load_golden("approx_surface.npz")
    feat, tmpx = O.synth_features(int(g["seed"]), B=1)
    set_maps(net, feat, tmpx)
    gen = chore_b200.Generator(net, threshold=2.0, filter_val=0.004, device=DEV)
    cc = T(g["crop_center"])
    s0 = O.synth_points("frustum", 32, 1, 512).to(DEV).requires_grad_(True)
    samples, preds = gen.approx_surface(net, s0, 10, {"crop_center": cc}, "human")
    d = (samples.detach() - T(g["samples_human"])).abs().amax(-1)
    assert d.median() < 1e-5 and (d < 1e-3).float().mean() > 0.8, (d.median().item(), (d < 1e-3).float().mean().item())
    assert preds[0].shape == (1, 2, 512) and samples.requires_grad
    for k, name in enumerate(("human", "object")):
        cur = O.synth_points("frustum", 32, 1, 512)
        for step in range(10):
            nxt, _ = O.approx_surface(sd, feat, tmpx, cur, cc.cpu(), 1, k)
            got, _ = gen.approx_surface(net, cur.to(DEV).requires_grad_(True), 1, {"crop_center": cc}, name)
            d = (got.detach().cpu() - nxt).abs().amax(-1)
            # a point whose gradient is ~0 has an ill-defined direction: allow a sliver of outliers
            assert d.median() < 1e-5 and (d < 1e-4).float().mean() > 0.99, (name, step, d.median().item(), (d < 1e-4).float().mean().item())
            cur = nxt
    if int(g["seed"]) == 31:        # the teacher-forced chain reproduces the golden end points
        assert (cur - torch.from_numpy(g["samples_object"])).abs().max() < 1e-3



# ------------------------------------------------------------------------------------------------
# encoder
# ------------------------------------------------------------------------------------------------
def test_encoder_vs_golden_128(net):
    g = load_golden("encoder_128.npz")
    img = O.synth_images(int(g["seed"]), B=2, size=128).to(DEV)
    net.filter(img)
    feat, tmpx, normx = net.get_im_feat(), net.tmpx, net.normx
    assert feat.shape == (2, 256, 32, 32) and tmpx.shape == (2, 64, 64, 64)
    report("encoder_128", tmpx=rel_err(tmpx, g["tmpx"]), normx=rel_err(normx[:, :, ::4, ::4], g["normx"]), feat=rel_err(feat, g["feat"]),
           feat_pure=pure_rel_err(feat, g["feat"]))
    assert rel_err(tmpx, g["tmpx"]) < TOL, rel_err(tmpx, g["tmpx"])
    assert rel_err(normx[:, :, ::4, ::4], g["normx"]) < TOL, rel_err(normx[:, :, ::4, ::4], g["normx"])
    assert rel_err(feat, g["feat"]) < ENC_TOL, rel_err(feat, g["feat"])     # ~60 GN+conv layers deep


def test_encoder_vs_golden_512(net):
    g = load_golden("encoder_512.npz")
    img = O.synth_images(int(g["seed"]), B=1, size=512).to(DEV)
    net.filter(img)
    feat, tmpx = net.get_im_feat(), net.tmpx
    assert feat.shape == (1, 256, 128, 128) and tmpx.shape == (1, 64, 256, 256)
    report("encoder_512", tmpx=rel_err(tmpx[:, :, ::8, ::8], g["tmpx_s8"]), feat=rel_err(feat[:, :, ::8, ::8], g["feat_s8"]),
           feat_pure=pure_rel_err(feat[:, :, ::8, ::8], g["feat_s8"]))
    assert rel_err(tmpx[:, :, ::8, ::8], g["tmpx_s8"]) < TOL
    assert rel_err(feat[:, :, ::8, ::8], g["feat_s8"]) < ENC_TOL, rel_err(feat[:, :, ::8, ::8], g["feat_s8"])
    from oracle.make_golden import checksum
    ck = checksum(feat.cpu().contiguous())
    assert abs(ck[0] - g["feat_ck"][0]) < 1e-3 * abs(g["feat_ck"][2]) * feat.numel() ** 0.5


def test_encoder_batched_equals_per_image(net):
    """B > 1 at the production size: a CTA of the conv kernels then walks work items of DIFFERENT images (128 tiles per image
    at 128^2 on 148 SMs), so per-image GroupNorm statistics, table switches and the helpers' statistics slots are all
    exercised; the result must be the per-image result (regression: helper slots of the next image's item were flushed into
    the previous image's sums)."""
    img = O.synth_images(3, B=3, size=512).to(DEV)
    fb, sb, nb = [o.clone() for o in net.handle.encode(img)]
    for i in range(3):
        f1, s1, n1 = net.handle.encode(img[i:i + 1].clone())
        torch.cuda.synchronize()
        for name, a, b in (("feat", fb[i:i + 1], f1), ("skip", sb[i:i + 1], s1), ("normx", nb[i:i + 1], n1)):
            assert rel_err(a, b) < 5e-5, (i, name, rel_err(a, b))      # summation order of the fp64 statistics atomics only


def test_encoder_refinit_weights():
    import chore_b200
    g = load_golden("encoder_128_refinit.npz")
    n = chore_b200.CHORE(device=DEV)
    n.load_state_dict(O.make_state_dict(int(g["weights_seed"]), "ref_init"))
    n.filter(O.synth_images(int(g["seed"]), B=1, size=128).to(DEV))
    report("encoder_refinit", tmpx=rel_err(n.tmpx, g["tmpx"]), feat=rel_err(n.get_im_feat(), g["feat"]))
    assert rel_err(n.tmpx, g["tmpx"]) < TOL
    assert rel_err(n.get_im_feat(), g["feat"]) < ENC_TOL


def test_encode_then_query_end_to_end(net, sd):
    """filter() -> query() on the encoder's own channels-last maps == oracle encode + query."""
    img = O.synth_images(77, B=1, size=128)
    net.filter(img.to(DEV))
    with torch.no_grad():
        f, t = O.encode(sd, img)
    cc = torch.tensor([[1008., 995.]])
    pts = O.synth_points("frustum", 78, 1, 3000)
    with torch.no_grad():
        ref = O.query(sd, f, t, pts, cc)
    net.query(pts.to(DEV), crop_center=cc.to(DEV))
    errs = [rel_err(got, want) for got, want in zip(net.get_preds(), ref[:4])]
    report("encode_then_query", df=errs[0], pca=errs[1], parts=errs[2], centers=errs[3])
    for e in errs:
        assert e < E2E_TOL, errs


def test_query_grid_256_planes_vs_golden(net):
    """BASELINE config 2 at its real size: three x-planes (3 x 65 536 points) of the 256^3 grid, coordinates generated in the
    kernel, against the reference's create_grid (model/sdf.py:4-27) + CHORE.query run."""
    g = load_golden("query_grid256.npz")
    feat, tmpx = O.synth_features(int(g["seed"]), B=1)
    set_maps(net, feat, tmpx)
    res = [int(r) for r in g["res"]]
    total, plane = res[0] * res[1] * res[2], res[1] * res[2]
    outs = [torch.empty(c, total, device=DEV) for c in (2, 9, 14, 6)]
    cc = T(g["crop_center"])
    for ix in g["planes"]:
        net.handle.query_grid(*net._maps(), cc, 0, res, list(g["bmin"]), list(g["bmax"]), int(ix) * plane, plane, 15, outs)
    sel = torch.cat([torch.arange(int(ix) * plane, (int(ix) + 1) * plane) for ix in g["planes"]]).to(DEV)
    df, pca, parts, centers = [o[:, sel] for o in outs]
    st = int(g["stride"])
    errs = {"df": rel_err(df, g["df"][0]), "pca": rel_err(pca[:, ::st], g["pca_s"].reshape(9, -1)),
            "parts": rel_err(parts[:, ::st], g["parts_s"][0]), "centers": rel_err(centers[:, ::st], g["centers_s"][0])}
    inimg = (torch.from_numpy(g["df"][0][0]) != O.OUT_DIST)[::st]
    pe = ((parts[:, ::st].cpu().double() - torch.from_numpy(g["parts_s"][0]).double()).abs().amax(0))
    scale = torch.from_numpy(g["parts_s"][0]).double().pow(2).mean().sqrt()
    report("grid256_planes", **errs, df_pure=pure_rel_err(df, g["df"][0]), parts_abs_err_in_image=float(pe[inimg].max() / scale),
           parts_abs_err_out_of_image=float(pe[~inimg].max() / scale), frac_in_image=float(inimg.float().mean()),
           parts_rms=float(scale), parts_absmax=float(torch.from_numpy(g["parts_s"][0]).abs().max()))
    for k, e in errs.items():
        assert e < TOL, (k, e)
    # OUT_DIST mask and part labels: bit-exact where the reference's top-2 margin is clear
    assert torch.equal(df.cpu() == O.OUT_DIST, torch.from_numpy(g["df"][0]) == O.OUT_DIST)
    clear = torch.from_numpy(g["top2_margin"][0].astype(np.float32)) > 5e-4
    same = parts.argmax(0).cpu() == torch.from_numpy(g["parts_argmax"][0]).long()
    assert bool(same[clear].all()), int((~same & clear).sum())


def test_query_batch32_vs_golden(net):
    """BASELINE config 4 batch shape: 32 images x 1 024 points in one launch, per-image crop centres."""
    g = load_golden("query_b32.npz")
    B, N = 32, 1024
    feat, tmpx = O.synth_features(int(g["seed"]), B=B, hw=64)
    set_maps(net, feat, tmpx)
    cc = torch.from_numpy(g["crop_center"])
    pts = torch.cat([O.synth_points("init_box", 97, B, N // 2), O.synth_points("frustum", 98, B, N // 2, cc)], 1)
    from oracle.make_golden import checksum
    assert np.allclose(checksum(pts), g["points_ck"], rtol=1e-6)
    outs, _ = net.handle.query_fwd(*net._maps(), pts.to(DEV), cc.to(DEV), 15)
    df, pca, parts, centers = outs
    st = int(g["stride"])
    errs = {"df": rel_err(df, g["df"]), "pca": rel_err(pca[..., ::st], g["pca_s"].reshape(B, 9, -1)),
            "parts": rel_err(parts[..., ::st], g["parts_s"]), "centers": rel_err(centers[..., ::st], g["centers_s"])}
    report("query_b32", **errs)
    for k, e in errs.items():
        assert e < TOL, (k, e)
    assert (parts.argmax(1).cpu() != torch.from_numpy(g["parts_argmax"]).long()).float().mean() < 1e-3


def test_example_frame_end_to_end_vs_golden(net):
    """The reference's shipped demo frame (tests/golden/example_frame = example/000000117377): TestData crop -> CHORE.filter ->
    CHORE.query against the reference's own TestData + network run -- realistic (non white-noise) input."""
    import os
    import chore_b200
    from conftest import GOLDEN, golden_smpl_assets
    g = load_golden("example_frame.npz")
    regs, _ = golden_smpl_assets(load_golden("fit_smpl_full.npz"))
    ds = chore_b200.TestData([os.path.join(GOLDEN, "example_frame", "k1.color.jpg")], image_size=(512, 512), crop_size=1200,
                             body25_reg=regs[0], write_crop_info=False)
    item = ds.get_item(0)
    assert np.array_equal(item["images"][:, ::4, ::4], g["images_own_s4"])
    net.filter(torch.from_numpy(item["images"]).unsqueeze(0).to(DEV))
    feat, tmpx = net.get_im_feat(), net.tmpx
    e_feat, e_tmpx = rel_err(feat[:, :, ::8, ::8], g["feat_s8"]), rel_err(tmpx[:, :, ::8, ::8], g["tmpx_s8"])
    cc = torch.from_numpy(item["crop_center"]).float().unsqueeze(0).to(DEV)
    net.query(T(g["points"]), crop_center=cc)
    errs = {k: rel_err(got, g[k]) for k, got in zip(("df", "pca", "parts", "centers"), net.get_preds())}
    report("example_frame", feat=e_feat, tmpx=e_tmpx, feat_pure=pure_rel_err(feat[:, :, ::8, ::8], g["feat_s8"]), **errs)
    assert e_tmpx < TOL and e_feat < ENC_TOL, (e_feat, e_tmpx)
    for k, e in errs.items():
        assert e < E2E_TOL, (k, e)


# ------------------------------------------------------------------------------------------------
# SMPL-H LBS, rigid transform, SO(3), fit step
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def smpl_layer():
    import chore_b200
    return chore_b200.SMPLHLayer(O.make_smplh_buffers(0), device=DEV)


def test_lbs_vs_golden(smpl_layer):
    g = load_golden("lbs.npz")
    pose, betas, trans, offs = (T(g[k]).clone().requires_grad_(True) for k in ("pose", "betas", "trans", "offsets"))
    verts, jtr, v_posed, naked = smpl_layer(pose, th_betas=betas, th_trans=trans, th_offsets=offs)
    assert rel_err(verts, g["verts"]) < TOL and rel_err(jtr, g["jtr"]) < TOL
    assert rel_err(v_posed[:, ::10], g["v_posed_s"]) < TOL and rel_err(naked[:, ::10], g["naked_s"]) < TOL
    ((T(g["g_verts"]) * verts).sum() + (T(g["g_jtr"]) * jtr).sum()).backward()
    assert rel_err(pose.grad, g["grad_pose"]) < 2e-4, rel_err(pose.grad, g["grad_pose"])
    assert rel_err(betas.grad, g["grad_betas"]) < 2e-4, rel_err(betas.grad, g["grad_betas"])
    assert rel_err(trans.grad, g["grad_trans"]) < 2e-4
    assert rel_err(offs.grad[:, ::10], g["grad_offsets_s"]) < 2e-4


def test_lbs_wrappers(smpl_layer):
    import chore_b200
    g = load_golden("lbs.npz")
    w = chore_b200.SMPLPyTorchWrapperBatch(smpl_layer, 2, betas=T(g["betas"]), pose=T(g["pose"]), trans=T(g["trans"]),
                                           offsets=T(g["offsets"]), device=DEV)
    verts = w()[0]
    assert rel_err(verts, g["verts"]) < TOL
    split = chore_b200.SMPLPyTorchWrapperBatchSplitParams.from_smpl(w)
    v2, j2, _, _ = split()
    assert torch.equal(v2, verts)
    (v2.sum() + j2.sum()).backward()
    assert split.body_pose.grad is not None and split.top_betas.grad.abs().sum() > 0


def test_rigid_and_so3_vs_golden():
    import chore_b200
    g = load_golden("rigid.npz")
    fit = chore_b200.ReconFitterBehave(device=DEV)
    R = fit.project_so3(T(g["rot"]))
    assert rel_err(R, g["R"]) < 2e-5
    assert rel_err(fit.project_so3(T(g["bad"])), g["R_bad"]) < 2e-5
    moved = fit.transform_obj_verts(T(g["obj"]), T(g["R"]), T(g["t"]), T(g["s"]))
    assert rel_err(moved, g["moved"]) < 2e-5


def test_so3_backward_vs_torch_svd():
    """the analytic adjoint of the projection == autograd through torch.svd (CPU, float64)."""
    import chore_b200
    gen = torch.Generator().manual_seed(5)
    m = torch.eye(3).unsqueeze(0) + 0.3 * torch.randn(16, 3, 3, generator=gen)
    gout = torch.randn(16, 3, 3, generator=gen)
    md = m.double().requires_grad_(True)
    (O.project_so3(md) * gout.double()).sum().backward()
    mg = m.to(DEV).requires_grad_(True)
    (chore_b200.ReconFitterBase.project_so3(mg) * gout.to(DEV)).sum().backward()
    assert rel_err(mg.grad, md.grad.float()) < 1e-4, rel_err(mg.grad, md.grad.float())


def test_fit_object_only_step_vs_golden(net, sd):
    """forward_step(phase='object only') (recon/recon_fit_behave.py:165-198): loss terms against the
    golden of the live reference, gradients stage by stage.

    Why stage by stage: the reference projects to SO(3) with an fp32 LAPACK SVD whose R differs from
    the fp64 in-kernel projection by ~1e-5, i.e. the object points move by ~2e-6 m.  On white-noise
    feature maps the per-point gradient is discontinuous across texel borders (bilinear), so a few
    points per thousand get a different gradient, and the parameter gradients -- sums of 3000
    per-point terms of both signs -- inherit that.  With identical points the per-point gradients
    agree to 1e-6 (first block); the chain is then checked with the oracle's R."""
    import chore_b200
    f = load_golden("fit_object_only.npz")
    feat, tmpx = O.synth_features(int(f["seed"]), B=2)
    set_maps(net, feat, tmpx)
    fit = chore_b200.ReconFitterBehave(device=DEV)
    rot, t, s = (T(f[k]).clone().requires_grad_(True) for k in ("rot", "t", "s"))
    data = {"objects": T(f["obj"]), "query_dict": {"crop_center": T(f["crop_center"])}, "smpl_center": T(f["smpl_center"])}
    losses = fit.forward_step(net, None, data, rot, t, s, "object only", noise=T(f["noise"]))
    for k in ("object", "scale", "ocent"):
        assert rel_err(losses[k], f[f"loss_{k}"]) < TOL, k
    total = fit.sum_dict(losses, fit.get_loss_weights(), int(f["it"]))
    assert rel_err(total, f["total"]) < TOL
    total.backward()
    for name, got, tol in (("t", t.grad, 5e-2), ("s", s.grad, 5e-2), ("rot", rot.grad, 1e-1)):
        assert rel_err(got, f[f"grad_{name}"]) < tol, (name, rel_err(got, f[f"grad_{name}"]))

    def losses_of(query, obj, sc, s_):
        cen = query(obj)[3]
        df = query(obj)[0]            # queried twice in the reference
        return {"object": torch.clamp(df[:, 1:2, :], max=0.8).mean(), "scale": torch.mean((s_ - 1.0) ** 2),
                "ocent": torch.nn.functional.mse_loss(torch.mean(obj, 1), sc + torch.mean(cen[:, 3:, :], -1),
                                                      reduction="none").sum(-1).mean()}

    def q_cuda(obj):
        net.query(obj, **data["query_dict"])
        return net.get_preds()

    # oracle chain (CPU), keeping R and the per-point gradient
    ro, to, so = (torch.from_numpy(f[k]).clone().requires_grad_(True) for k in ("rot", "t", "s"))
    R_o = O.decopose_axis(ro, torch.from_numpy(f["noise"]))
    R_o.retain_grad()
    obj_o = O.transform_obj_verts(torch.from_numpy(f["obj"]), R_o, to, so)
    obj_o.retain_grad()
    cc_o, sc_o = torch.from_numpy(f["crop_center"]), torch.from_numpy(f["smpl_center"])
    O.sum_dict(losses_of(lambda o: O.query(sd, feat, tmpx, o, cc_o), obj_o, sc_o, so), float(f["it"])).backward()
    # (1) identical points: per-point gradient of the whole loss
    p_c = obj_o.detach().contiguous().to(DEV).requires_grad_(True)
    fit.sum_dict(losses_of(q_cuda, p_c, data["smpl_center"], T(f["s"])), fit.get_loss_weights(), int(f["it"])).backward()
    grad_close(p_c.grad, obj_o.grad, tol=1e-4, frac=0.999, worst=0.5)
    # (2) identical R: rigid transform + queries + losses, gradients to (R, t, s)
    R_c = R_o.detach().to(DEV).requires_grad_(True)
    t2, s2 = (T(f[k]).clone().requires_grad_(True) for k in ("t", "s"))
    obj_c = fit.transform_obj_verts(data["objects"], R_c, t2, s2)
    assert rel_err(obj_c, obj_o) < 1e-6
    fit.sum_dict(losses_of(q_cuda, obj_c, data["smpl_center"], s2), fit.get_loss_weights(), int(f["it"])).backward()
    one_point = 2.0 * obj_o.grad.abs().amax().item() * 1.5       # two texel-border crossings
    for name, got, want in (("R", R_c.grad, R_o.grad), ("t", t2.grad, to.grad), ("s", s2.grad, so.grad)):
        bound = 2e-4 * (want.abs() + want.pow(2).mean().sqrt()) + one_point
        assert bool(((got.cpu() - want).abs() <= bound).all()), (name, (got.cpu() - want).abs().max().item(), one_point)


def test_smpl_fit_step_vs_oracle(net, sd, smpl_layer):
    """LBS -> query -> df_h + part loss -> backward to (pose, betas, trans): one SMPL-phase step
    (recon/recon_fit_behave.py:293-337 field terms).  Losses end to end; gradients stage by stage
    (see test_fit_object_only_step_vs_golden for why)."""
    import chore_b200
    feat, tmpx = O.synth_features(61, B=1)
    set_maps(net, feat, tmpx)
    buf = O.make_smplh_buffers(0)
    gen = torch.Generator().manual_seed(62)
    pose0, betas0 = 0.1 * torch.randn(1, 156, generator=gen), 0.5 * torch.randn(1, 10, generator=gen)
    trans0 = torch.tensor([[0.0, 0.1, 2.2]])
    labels = torch.randint(14, (1, 6890), generator=gen)
    cc = torch.tensor([[1008., 995.]])
    # oracle
    p, b, t = (x.clone().requires_grad_(True) for x in (pose0, betas0, trans0))
    verts = O.lbs_forward(buf, p, b, t)[0]
    verts.retain_grad()
    lo = O.smpl_losses(sd, feat, tmpx, cc, verts, labels)
    tot_o = O.sum_dict(lo, 1.0)
    tot_o.backward()
    # CUDA, end to end through the wrappers
    w = chore_b200.SMPLPyTorchWrapperBatch(smpl_layer, 1, betas=betas0, pose=pose0, trans=trans0, device=DEV)
    fit = chore_b200.ReconFitterBehave(device=DEV)
    data = {"net": net, "query_dict": {"crop_center": cc.to(DEV)}, "part_labels": labels.to(DEV)}
    lc = fit.forward_smpl(w, data)
    tot_c = fit.sum_dict(lc, fit.get_loss_weights(), 1)
    tot_c.backward()
    assert rel_err(w()[0], verts) < 1e-5
    assert rel_err(lc["df_h"], lo["df_h"]) < TOL, (lc["df_h"].item(), lo["df_h"].item())
    assert rel_err(lc["part"], lo["part"]) < TOL, (lc["part"].item(), lo["part"].item())
    assert rel_err(tot_c, tot_o) < TOL
    chained = {n: rel_err(got, want) for n, got, want in (("trans", w.trans.grad, t.grad), ("betas", w.betas.grad, b.grad), ("pose", w.pose.grad, p.grad))}
    report("smpl_fit_step_chained_gradients_vs_oracle", **chained)
    for n, e in chained.items():
        assert e < CHAINED_TOL, (n, e)
    # stage 1: identical vertices -> per-vertex gradient of the field losses
    v_c = verts.detach().contiguous().to(DEV).requires_grad_(True)
    net.query(v_c, crop_center=cc.to(DEV))
    df_c, _, parts_c, _ = net.get_preds()
    l1 = {"df_h": torch.clamp(df_c[:, 0:1, :], max=0.1).mean(),
          "part": torch.nn.functional.cross_entropy(parts_c, labels.to(DEV), reduction="none").sum(-1).mean()}
    fit.sum_dict(l1, fit.get_loss_weights(), 1).backward()
    grad_close(v_c.grad, verts.grad, tol=1e-4, frac=0.999, worst=0.5)
    # stage 2: identical vertex gradients -> LBS adjoint
    g_pose, g_betas, g_trans, _ = smpl_layer.handle.lbs_bwd(pose0.to(DEV), betas0.to(DEV), trans0.to(DEV), None,
                                                            verts.grad.contiguous().to(DEV), None, False)
    assert rel_err(g_trans, t.grad) < 2e-4 and rel_err(g_betas, b.grad) < 2e-4, (rel_err(g_trans, t.grad), rel_err(g_betas, b.grad))
    assert rel_err(g_pose, p.grad) < 2e-4, rel_err(g_pose, p.grad)


def test_fused_fit_steps_match_autograd_path(net, smpl_layer):
    """FusedFitSteps (explicit adjoint kernels, no autograd graph, CUDA-graph replay) == the autograd-Function
    path (forward_smpl / forward_step + backward): same losses and the same parameter gradients."""
    import chore_b200
    feat, tmpx = O.synth_features(71, B=1)
    set_maps(net, feat, tmpx)
    gen = torch.Generator().manual_seed(72)
    mk = lambda: chore_b200.SMPLPyTorchWrapperBatchSplitParams.from_smpl(chore_b200.SMPLPyTorchWrapperBatch(
        smpl_layer, 1, betas=0.3 * torch.randn(1, 10, generator=torch.Generator().manual_seed(1)),
        pose=0.1 * torch.randn(1, 156, generator=torch.Generator().manual_seed(2)), trans=torch.tensor([[0.0, 0.1, 2.2]]), device=DEV))
    labels = torch.randint(14, (1, 6890), generator=gen).to(DEV)
    obj = (0.2 * torch.randn(1, 5000, 3, generator=gen)).to(DEV)
    noise = torch.rand(1, 3, 3, generator=gen).to(DEV)
    data = {"net": net, "query_dict": {"crop_center": torch.tensor([[1008., 995.]], device=DEV)}, "part_labels": labels,
            "objects": obj, "smpl_center": torch.tensor([[0.0, 0.1, 2.2]], device=DEV)}
    mkobj = lambda: ((torch.eye(3).unsqueeze(0) + 0.05 * torch.randn(1, 3, 3, generator=torch.Generator().manual_seed(3))).to(DEV).requires_grad_(True),
                     torch.tensor([[0.2, 0.1, 2.3]], device=DEV, requires_grad=True), torch.full((1,), 1.05, device=DEV, requires_grad=True))
    fit = chore_b200.ReconFitterBehave(device=DEV)
    w = fit.get_loss_weights()
    # autograd path
    sa = mk(); Ra, ta, sca = mkobj()
    la = fit.sum_dict(fit.forward_smpl(sa, data), w, 1); la.backward()
    lo = fit.sum_dict(fit.forward_step(net, sa, data, Ra, ta, sca, "object only", noise=noise), w, 1); lo.backward()
    # fused path, learning rate 0 so that the parameters stay put
    sf = mk(); Rf, tf, scf = mkobj()
    fused = chore_b200.FusedFitSteps(net, sf, data, Rf, tf, scf, lr_smpl=0.0, lr_obj=0.0)
    lf = fused.smpl_step(); lfo = fused.object_step(noise)
    assert rel_err(lf, la) < 1e-5 and rel_err(lfo, lo) < 1e-5, (lf.item(), la.item(), lfo.item(), lo.item())
    for name in ("trans", "global_pose", "body_pose", "top_betas", "other_betas"):
        assert rel_err(getattr(sf, name).grad, getattr(sa, name).grad) < 1e-4, name
    for a, b, name in ((Rf, Ra, "R"), (tf, ta, "t"), (scf, sca, "s")):
        assert rel_err(a.grad, b.grad) < 1e-4, (name, rel_err(a.grad, b.grad))
    # CUDA-graph replay: the loss decreases over 30 replayed Adam steps
    sg = mk(); Rg, tg, scg = mkobj()
    # (plain Adam here: with the reference's accumulate-over-10-steps gradients the loss of a white-noise field need not go
    # down monotonically; that loop shape is pinned against the autograd path in test_fused_fit_loop_matches_reference_loop_semantics)
    fused_g = chore_b200.FusedFitSteps(net, sg, data, Rg, tg, scg, lr_smpl=0.006, lr_obj=0.006, accumulate=False)
    g_smpl, g_obj = fused_g.graphed()
    first = (float(g_smpl()), float(g_obj()))
    for i in range(30):
        g_smpl(); g_obj()
    last = (float(g_smpl()), float(g_obj()))
    assert last[0] < first[0] and last[1] < first[1], (first, last)


def test_fused_adam_matches_torch_adam_with_accumulation():
    """chore_adam_step against torch.optim.Adam over 12 steps in the reference's loop shape: zero_grad() once per outer
    iteration, gradients ADD UP over the inner steps (recon/recon_fit_behave.py:135-152,244-273), every gradient scaled by
    1 / (1 + decay) with a decay that changes per outer iteration."""
    import chore_b200
    gen = torch.Generator().manual_seed(5)
    shapes = [(2, 3), (2, 63), (2,), (2, 3, 3)]
    p_ref = [torch.randn(*sh, generator=gen).to(DEV).requires_grad_(True) for sh in shapes]
    p_fus = [p.detach().clone().requires_grad_(True) for p in p_ref]
    opt = torch.optim.Adam(p_ref, lr=0.006)
    fused = chore_b200.FusedAdam(p_fus, lr=0.006, accumulate=True)
    kdev = torch.ones(1, device=DEV)
    for outer in range(3):
        opt.zero_grad(); fused.zero_grad()
        k = 1.0 / (1.0 + outer / 3)
        kdev.fill_(k)
        for inner in range(4):
            grads = [torch.randn(*sh, generator=gen).to(DEV) for sh in shapes]
            for p, g in zip(p_ref, grads):
                p.grad = (p.grad if p.grad is not None else torch.zeros_like(p)) + k * g
            opt.step()
            fused.step(grads, kdev)
            for a, b in zip(p_fus, p_ref):
                assert rel_err(a, b) < 1e-5, (outer, inner, rel_err(a, b))
                assert rel_err(a.grad, b.grad) < 1e-5


def test_fused_fit_loop_matches_reference_loop_semantics(net, smpl_layer):
    """FusedFitSteps driven like optimize_smpl / optimize_smpl_object drive their step (zero_grad per outer iteration,
    accumulating .grad, decay = it / 3 in phase 'kpts', torch.optim.Adam) == the autograd path run in that same loop:
    the parameters after 2 outer x 3 inner steps agree, eagerly and replayed from CUDA graphs (whose warm-up must not
    leave extra updates behind)."""
    import chore_b200
    g = load_golden("fit_smpl_full.npz")
    fit, w, data, _ = _smpl_full_setup(net, smpl_layer, g)
    gen = torch.Generator().manual_seed(19)
    obj = (0.2 * torch.randn(2, 3000, 3, generator=gen)).to(DEV)
    noise = torch.rand(2, 3, 3, generator=gen).to(DEV)
    data.update({"objects": obj, "smpl_center": torch.tensor([[0.0, 0.1, 2.2], [0.0, 0.0, 2.2]], device=DEV)})
    mkobj = lambda: ((torch.eye(3).repeat(2, 1, 1) + 0.05 * torch.randn(2, 3, 3, generator=torch.Generator().manual_seed(3))).to(DEV).requires_grad_(True),
                     torch.tensor([[0.2, 0.1, 2.3], [0.1, 0.0, 2.2]], device=DEV, requires_grad=True), torch.ones(2, device=DEV, requires_grad=True))
    names = ("trans", "global_pose", "body_pose", "top_betas", "other_betas")
    wd = fit.get_loss_weights()
    outer_iters, inner = (3, 6), 3          # 'kpts' iterations it = 3 and 6: decay 1 and 2

    # (1) learning rate 0: the parameters stay put, so `.grad` after the loop is exactly what the reference loop accumulates
    #     (3 x the decayed gradient of the last outer iteration) -- no chaotic Adam trajectory in the comparison
    sa = fit.split_smpl(w); Ra, ta, sca = mkobj()
    opt_s = torch.optim.Adam([getattr(sa, n) for n in names], 0.0)
    opt_o = torch.optim.Adam([ta, Ra, sca], lr=0.0)
    for it in outer_iters:
        opt_s.zero_grad(); opt_o.zero_grad()
        for _ in range(inner):
            fit.sum_dict(fit.forward_smpl(sa, data, "kpts"), wd, it / 3).backward(); opt_s.step()
            fit.sum_dict(fit.forward_step(net, sa, data, Ra, ta, sca, "object only", noise=noise), wd, it / 3).backward(); opt_o.step()

    def run(graphed, lr, outer=outer_iters, inner=inner):
        sf = fit.split_smpl(w); Rf, tf, scf = mkobj()
        fused = chore_b200.FusedFitSteps(net, sf, data, Rf, tf, scf, lr_smpl=lr, lr_obj=lr, fitter=fit, phase="kpts")
        s_step = fused.graphed()[0] if graphed else fused.smpl_step
        o_step = lambda: fused.object_step(noise)      # (a captured object step draws its own noise: keep it pinned)
        for it in outer:
            fused.zero_grad(); fused.set_decay(it / 3)
            for _ in range(inner):
                s_step(); o_step()
        return fused, sf, (Rf, tf, scf)

    for graphed in (False, True):
        fused, sf, (Rf, tf, scf) = run(graphed, 0.0)
        for n in names:
            assert rel_err(getattr(sf, n).grad, getattr(sa, n).grad) < 2e-4, (graphed, n, rel_err(getattr(sf, n).grad, getattr(sa, n).grad))
        for a, b, n in ((Rf, Ra, "R"), (tf, ta, "t"), (scf, sca, "s")):
            assert rel_err(a.grad, b.grad) < 2e-4, (graphed, n, rel_err(a.grad, b.grad))
        # the capture's warm-up steps were undone: exactly 6 updates were applied
        assert int(fused.opt_smpl.step_count) == len(outer_iters) * inner and int(fused.opt_obj.step_count) == len(outer_iters) * inner
    # (2) learning rate 0.006, ONE step: the replayed step lands where the eager step lands, i.e. the warm-up left nothing behind in
    #     parameters, moments or accumulators (3 stray updates would show as ~0.02).  Only one step is compared: the adjoint kernels
    #     use fp32 atomics, and on white-noise features the Adam trajectory amplifies that 5e-10 run-to-run noise by ~30x per step
    #     (measured: two EAGER runs of the same 6 steps differ by 2e-3).
    _, se, _ = run(False, 0.006, outer=(3,), inner=1)
    _, sg2, _ = run(True, 0.006, outer=(3,), inner=1)
    for n in names:
        assert (getattr(se, n) - getattr(sg2, n)).abs().max() < 1e-6, n
        assert (getattr(se, n) - getattr(sa, n)).abs().max() > 1e-3, n        # ... and the parameters did move (one Adam step = lr)
    # the 'global' phase of optimize_smpl: only top_betas and trans move, lr 0.02
    sg = fit.split_smpl(w); Rg, tg, scg = mkobj()
    fused = chore_b200.FusedFitSteps(net, sg, data, Rg, tg, scg, fitter=fit, phase="global")
    before = {n: getattr(sg, n).detach().clone() for n in names + ("hand_pose",)}
    fused.zero_grad(); fused.smpl_step()
    assert not torch.equal(sg.trans, before["trans"]) and not torch.equal(sg.top_betas, before["top_betas"])
    for n in ("global_pose", "body_pose", "hand_pose", "other_betas"):
        assert torch.equal(getattr(sg, n), before[n]), n
    # the full pose gradient (incl. the hand pose, which carries the hand prior) matches autograd
    sh = fit.split_smpl(w)
    fit.sum_dict(fit.forward_smpl(sh, data, "kpts"), wd, 1).backward()
    fh = chore_b200.FusedFitSteps(net, fit.split_smpl(w), data, Rg, tg, scg, lr_smpl=0.0, fitter=fit, phase="kpts")
    fh.set_decay(1); fh.smpl_step()
    assert rel_err(0.5 * fh._g_pose_full[:, 66:], sh.hand_pose.grad) < 1e-4, rel_err(0.5 * fh._g_pose_full[:, 66:], sh.hand_pose.grad)


def test_fused_iteration_two_streams_equals_sequential(net, smpl_layer):
    """FusedFitSteps.iteration() (object step forked onto a side stream, joined) and its captured form give the same first update as
    the two steps run one after the other: the steps touch disjoint parameters and scratch."""
    import chore_b200
    g = load_golden("fit_smpl_full.npz")
    fit, w, data, _ = _smpl_full_setup(net, smpl_layer, g)
    gen = torch.Generator().manual_seed(29)
    data.update({"objects": (0.2 * torch.randn(2, 4000, 3, generator=gen)).to(DEV),
                 "smpl_center": torch.tensor([[0.0, 0.1, 2.2], [0.0, 0.0, 2.2]], device=DEV)})
    noise = torch.rand(2, 3, 3, generator=gen).to(DEV)
    mkobj = lambda: ((torch.eye(3).repeat(2, 1, 1) + 0.05 * torch.randn(2, 3, 3, generator=torch.Generator().manual_seed(3))).to(DEV).requires_grad_(True),
                     torch.tensor([[0.2, 0.1, 2.3], [0.1, 0.0, 2.2]], device=DEV, requires_grad=True), torch.ones(2, device=DEV, requires_grad=True))

    def run(mode):
        sf = fit.split_smpl(w); R, t, s = mkobj()
        fused = chore_b200.FusedFitSteps(net, sf, data, R, t, s, fitter=fit, phase="kpts")
        fused.zero_grad()
        if mode == "seq":
            fused.smpl_step(); fused.object_step(noise)
        elif mode == "fork":
            fused.iteration(noise)
        else:
            step = fused.graphed_iteration()         # draws its own noise on the device: compare the SMPL side + counters only
            step()
        torch.cuda.synchronize()
        return fused, sf, (R, t, s)

    f0, s0, o0 = run("seq")
    f1, s1, o1 = run("fork")
    f2, s2, o2 = run("graph")
    for n in ("trans", "global_pose", "body_pose", "top_betas", "other_betas"):
        assert (getattr(s0, n) - getattr(s1, n)).abs().max() < 1e-6 and (getattr(s0, n) - getattr(s2, n)).abs().max() < 1e-6, n
    for a, b in zip(o0, o1):
        assert (a - b).abs().max() < 1e-6
    assert int(f2.opt_smpl.step_count) == 1 and int(f2.opt_obj.step_count) == 1 and not torch.equal(o2[1], mkobj()[1])


def test_generator_device_kernels(net):
    """csrc/generator.cu against torch: ordered compaction == boolean-mask indexing, resampling == the reference's formula for
    given draws, Philox draws reproducible and in range, finalize == mean over the kept prefix."""
    h = net.handle
    g = torch.Generator().manual_seed(11)
    B, N, cap = 3, 5000, 6000
    df = torch.rand(B, 2, N, generator=g).to(DEV) * 0.02
    df[2] = 1.0                                   # image 2: no hit at all
    df[2, 1, 77] = 0.0                            # ... except one (the reference's "<= 1 hit" fallback)
    surf, samples = torch.randn(B, N, 3, generator=g).to(DEV), torch.randn(B, N, 3, generator=g).to(DEV)
    pca, parts, cen = torch.randn(B, 9, N, generator=g).to(DEV), torch.randn(B, 14, N, generator=g).to(DEV), torch.randn(B, 6, N, generator=g).to(DEV)
    out = (torch.zeros(B, cap, 3, device=DEV), torch.zeros(B, cap, dtype=torch.int32, device=DEV), torch.zeros(B, cap, 9, device=DEV),
           torch.zeros(B, cap, 6, device=DEV), torch.zeros(B, dtype=torch.int32, device=DEV))
    packed, cnt = torch.zeros(B, N, 3, device=DEV), torch.zeros(B, dtype=torch.int32, device=DEV)
    total = torch.zeros(1, dtype=torch.int32, device=DEV)
    for rep in range(2):                          # two appends: the second lands behind the first
        h.gen_compact(df, 1, 2.0, 0.01, samples, packed, cnt, surf=surf, preds=(pca, parts, cen), out=out)
        h.gen_total(cnt, total)
    mask = torch.clamp(df[:, 1], max=2.0) < 0.01
    for b in range(B):
        m = mask[b]
        k = int(m.sum())
        assert int(cnt[b]) == k and int(out[4][b]) == min(cap, 2 * k)
        assert torch.equal(packed[b, :k], samples[b, m])
        for rep in range(2):
            lo, hi = rep * k, min(cap, (rep + 1) * k)
            assert torch.equal(out[0][b, lo:hi], surf[b, m][:hi - lo])
            assert torch.equal(out[1][b, lo:hi].long(), parts[b][:, m].argmax(0)[:hi - lo])
            assert torch.equal(out[2][b, lo:hi], pca[b][:, m].t()[:hi - lo]) and torch.equal(out[3][b, lo:hi], cen[b][:, m].t()[:hi - lo])
    assert int(total) == 2 * int(mask.sum(1).min())
    pm, cm = h.gen_finalize(out[2], out[3], total)
    n = int(total)
    assert rel_err(pm, out[2][:, :n].mean(1)) < 1e-5 and rel_err(cm, out[3][:, :n].mean(1)) < 1e-5
    # resampling with given draws: hit images pick packed[floor(u * count)], the starved image restarts from the initial samples
    S = 4000
    init = torch.randn(B, 300, 3, generator=g).to(DEV)
    u, nrm = torch.rand(B, S, generator=g).to(DEV), torch.randn(B, S, 3, generator=g).to(DEV)
    new = h.gen_resample(packed, cnt, init, S, 2.0 / 3, 0.5, 0, 0, u, nrm)
    for b in range(B):
        k = int(cnt[b])
        want = packed[b][(u[b] * k).long().clamp(max=k - 1)] + (2.0 / 3) * nrm[b] if k > 1 else init[b][(u[b] * 300).long().clamp(max=299)] + 0.5 * nrm[b]
        assert torch.equal(new[b], want), b
    # Philox: reproducible for (seed, offset), different otherwise, noise of the right scale, every sample near a hit
    a1 = h.gen_resample(packed, cnt, init, S, 0.0, 0.0, 5, 3)
    a2 = h.gen_resample(packed, cnt, init, S, 0.0, 0.0, 5, 3)
    a3 = h.gen_resample(packed, cnt, init, S, 0.0, 0.0, 5, 4)
    assert torch.equal(a1, a2) and not torch.equal(a1, a3)
    k0 = int(cnt[0])
    same = (a1[0][:, None, :] == packed[0, :k0][None]).all(-1)      # (S, hits): exact row matches
    assert bool(same.any(1).all())                                 # sigma = 0: every new sample IS one of the hits
    assert same.float().argmax(1).unique().numel() > 0.5 * min(k0, S)              # ... spread over the hits
    noisy = h.gen_resample(packed, cnt, init, S, 2.0 / 3, 0.5, 5, 3)
    assert abs(float((noisy[0] - a1[0]).std()) - 2.0 / 3) < 0.03 and abs(float((noisy[2] - a1[2]).std()) - 0.5) < 0.03


def test_generator_device_loop_equals_reference_order_loop(net):
    """Generator(rng='device') == Generator(rng='reference') when the device loop is fed the reference loop's own draws (torch CPU
    generator, randint then randn per image and outer iteration, recon/generator.py:166-176): same points, labels, means."""
    import chore_b200
    feat, tmpx = O.synth_features(83, B=2)
    set_maps(net, feat, tmpx)
    cc = torch.tensor([[1008., 995.], [1000., 990.]], device=DEV)
    S = 2000
    host = chore_b200.Generator(net, threshold=2.0, filter_val=10.0, device=DEV)
    devg = chore_b200.Generator(net, threshold=2.0, filter_val=10.0, device=DEV, rng="device")
    torch.manual_seed(3)
    init = host.init_samples(3000, batch_size=2)
    torch.manual_seed(4)
    want = host.gen_pc_batch(net, "object", init, 5000, {"crop_center": cc}, num_steps=2, sample_num=S)

    def draws(it, iter_count):                    # what the reference loop draws, in its order
        us, ns = [], []
        for k in iter_count.tolist():
            n = k if k > 1 else init.shape[1]
            idx = torch.randint(n, (S,))
            us.append((idx.double() + 0.5) / n)
            ns.append(torch.randn(1, S, 3)[0])
        return torch.stack(us).float().to(DEV), torch.stack(ns).to(DEV)

    torch.manual_seed(4)
    got = devg._gen_pc_batch_device(net, "object", init, 5000, {"crop_center": cc}, 2, 100, S, randoms=draws)
    assert got["points"].shape == want["points"].shape and got["parts"].dtype == torch.int64
    assert torch.equal(got["points"].cpu(), want["points"]) and torch.equal(got["parts"].cpu(), want["parts"])
    assert rel_err(got["pca_axis"], want["pca_axis"]) < 1e-5 and rel_err(got["centers"], want["centers"]) < 1e-5
    # and the stand-alone device loop (Philox draws) runs, is reproducible for a seed and fills the request
    torch.manual_seed(3)
    a = devg.gen_pc_batch(net, "human", init, 5000, {"crop_center": cc}, num_steps=2, sample_num=S)
    b = devg.gen_pc_batch(net, "human", init, 5000, {"crop_center": cc}, num_steps=2, sample_num=S)
    assert a["points"].shape[1] >= 5000 and torch.equal(a["points"], b["points"]) and a["centers"].shape == (2, 6)


def test_generator_gen_pc_batch_mechanics(net):
    """Generator.gen_pc_batch (recon/generator.py:123-217) on-device: projection steps, surface filter, resampling and
    the final argmax / mean reductions.  With random weights the field is not a distance field, so the filter value is
    opened up; this checks the loop mechanics, shapes and that every returned point passed the filter."""
    import chore_b200
    feat, tmpx = O.synth_features(81, B=2)
    set_maps(net, feat, tmpx)
    gen = chore_b200.Generator(net, threshold=2.0, filter_val=10.0, device=DEV)
    cc = torch.tensor([[1008., 995.], [1000., 990.]], device=DEV)
    torch.manual_seed(0)
    init = gen.init_samples(3000, batch_size=2)
    init[1] = init[0]                       # the reference only rescales batch element 0 (kept quirk): reuse it
    out = gen.gen_pc_batch(net, "human", init, 2500, {"crop_center": cc}, num_steps=3, sample_num=2000)
    assert out["points"].shape[0] == 2 and out["points"].shape[2] == 3 and out["points"].shape[1] >= 2500
    n = out["points"].shape[1]
    assert out["parts"].shape == (2, n) and out["parts"].dtype == torch.int64
    assert out["pca_axis"].shape == (2, 3, 3) and out["centers"].shape == (2, 6)
    assert int(out["parts"].min()) >= 0 and int(out["parts"].max()) < 14
    assert torch.isfinite(out["points"]).all()


# ------------------------------------------------------------------------------------------------
# full SMPL phase: landmark regressors, priors, smplz, 2-D keypoints, optimisation loops
# ------------------------------------------------------------------------------------------------
def _smpl_full_setup(net, smpl_layer, g):
    import chore_b200
    from conftest import golden_smpl_assets
    regs, pri = golden_smpl_assets(g)
    feat, tmpx = O.synth_features(int(g["feat_seed"]), B=2)
    set_maps(net, feat, tmpx)
    bp = chore_b200.fitter.MahalanobisPrior(pri["body_mean"], pri["body_prec"], device=DEV)
    hp = chore_b200.fitter.HandPrior(pri["hand_mean"], pri["lh_prec"], pri["rh_prec"], device=DEV)
    fit = chore_b200.ReconFitterBehave(device=DEV, priors=(bp, hp), strict=True)
    w = chore_b200.SMPLPyTorchWrapperBatch(smpl_layer, 2, betas=g["betas"], pose=g["pose"], trans=g["trans"], device=DEV,
                                           regressors=regs)
    data = {"net": net, "part_labels": T(g["part_labels"]), "pose_init": T(g["pose_init"]), "body_kpts": T(g["body_kpts"]),
            "query_dict": {"crop_center": T(g["crop_center"])}}
    return fit, w, data, regs


def test_landmarks_vs_golden(net, smpl_layer):
    """get_landmarks (lib_smpl/wrapper_pytorch.py:176-190) on the CSR kernel: values against the reference's run,
    adjoint against the dense transpose."""
    g = load_golden("fit_smpl_full.npz")
    fit, w, data, regs = _smpl_full_setup(net, smpl_layer, g)
    J, face, hands = w.get_landmarks()
    assert J.shape == (2, 25, 3) and face.shape == (2, 70, 3) and hands.shape == (2, 42, 3)
    for got, name in ((J, "J"), (face, "face"), (hands, "hands")):
        assert rel_err(got, g[name]) < 1e-5, name
    gen = torch.Generator().manual_seed(5)
    g_out = torch.randn(2, 137, 3, generator=gen)
    dense = torch.from_numpy(np.vstack([r.toarray() for r in regs])).double()
    want = torch.einsum("lv,blc->bvc", dense, g_out.double())
    h = w.regressors.handle
    got = h.landmarks_bwd(g_out.to(DEV))
    assert rel_err(got, want) < 1e-5
    acc = torch.ones(2, 6890, 3, device=DEV)
    h.landmarks_bwd(g_out.to(DEV), acc)
    assert rel_err(acc, want + 1.0) < 1e-5
    # autograd reaches the SMPL parameters through the landmark kernel
    w.zero_grad()
    w.get_landmarks()[0][:, 8, 2].sum().backward()
    assert w.trans.grad is not None and rel_err(w.trans.grad[:, 2], torch.ones(2)) < 1e-5


def test_fit_smpl_full_step_vs_golden(net, smpl_layer):
    """forward_smpl(phase='kpts') with every term of recon/recon_fit_behave.py:293-337 against the reference's own run:
    each loss, the decayed sum, and the gradients to the split parameters (chained through query + LBS adjoints)."""
    import chore_b200
    g = load_golden("fit_smpl_full.npz")
    fit, w, data, _ = _smpl_full_setup(net, smpl_layer, g)
    split = fit.split_smpl(w)
    losses = fit.forward_smpl(split, data, "kpts")
    assert list(losses) == [str(k) for k in g["loss_order"]]
    for k, v in losses.items():
        assert rel_err(v, g[f"loss_{k}"]) < TOL, (k, float(v), float(g[f"loss_{k}"]))
    total = fit.sum_dict(losses, fit.get_loss_weights(), float(g["decay"]))
    assert rel_err(total, g["total"]) < TOL
    total.backward()
    # chained gradients: sums of thousands of per-vertex terms with ReLU / clamp gates (see the module docstring)
    chained = {name: rel_err(getattr(split, name).grad, g[f"grad_{name}"]) for name in ("trans", "global_pose", "body_pose", "hand_pose", "top_betas", "other_betas")}
    report("fit_smpl_full_chained_gradients_vs_reference", **chained)
    for name, e in chained.items():
        assert e < CHAINED_TOL, (name, e)
    # the fused (autograd-free) step computes the same loss and the same gradients as the autograd path
    split2 = fit.split_smpl(w)
    R, t, s = torch.eye(3, device=DEV).repeat(2, 1, 1), torch.zeros(2, 3, device=DEV), torch.ones(2, device=DEV)
    fused = chore_b200.FusedFitSteps(net, split2, data, R, t, s, lr_smpl=0.0, decay=float(g["decay"]), fitter=fit, phase="kpts")
    lf = fused.smpl_step()
    assert rel_err(lf, total) < 1e-5, (float(lf), float(total))
    for name in ("trans", "global_pose", "body_pose", "top_betas", "other_betas"):
        assert rel_err(getattr(split2, name).grad, getattr(split, name).grad) < 1e-4, name


def test_optimize_loops_run_and_descend(net, smpl_layer):
    """optimize_smpl / optimize_smpl_object (recon/recon_fit_behave.py:90-163,224-291) with a shortened schedule:
    same phases, optimisers and decay; the weighted loss goes down and the parameters are copied back."""
    g = load_golden("fit_smpl_full.npz")
    fit, w, data, _ = _smpl_full_setup(net, smpl_layer, g)
    logs = []
    pose_before = w.pose.detach().clone()
    smpl, scale = fit.optimize_smpl(w, data, iter_for_betas=2, iter_for_pose=2, iter_for_kpts=1, steps_per_iter=3, max_iter=1,
                                    log=logs.append)
    assert smpl is w and scale.shape == (2,) and torch.isfinite(scale).all()
    # 6 outer x 3 inner steps unless the reference's early-stop rule fires inside the last ('kpts') iteration
    assert 5 * 3 < len(logs) <= 6 * 3 and "j2d" in logs[-1] and "j2d" not in logs[0]
    assert not torch.equal(w.pose.detach(), pose_before)
    first = float(logs[6].split("df_h: ")[1].split(",")[0])       # first 'smpl all pose' step
    last = float(logs[11].split("df_h: ")[1].split(",")[0])
    assert last <= first + 0.05 * abs(first), (first, last)      # synthetic fields can be negative
    gen = torch.Generator().manual_seed(9)
    data.update({"smpl": w, "objects": (0.2 * torch.randn(2, 3000, 3, generator=gen)).to(DEV),
                 "obj_R": (torch.eye(3).repeat(2, 1, 1) + 0.05 * torch.randn(2, 3, 3, generator=gen)).to(DEV).requires_grad_(True),
                 "obj_t": torch.tensor([[0.2, 0.1, 2.3], [0.1, 0.0, 2.2]], device=DEV, requires_grad=True),
                 "obj_s": torch.ones(2, device=DEV, requires_grad=True)})
    logs = []
    t0 = data["obj_t"].detach().clone()
    fit.part_labels = torch.randint(14, (6890,), generator=gen).to(torch.int32)
    out = fit.optimize_smpl_object(net, data, obj_iter=2, joint_iter=1, steps_per_iter=3, max_iter=1, log=logs.append)
    # 2 'object only' iterations, no silhouette data (the 'sil' phase is skipped), then joint_iter + max_iter = 2 'joint' iterations
    # unless the reference's early-stop rule fires
    assert out[0] is w and 6 < len(logs) <= 12 and "ocent" in logs[0] and logs[0].startswith("object only") and logs[6].startswith("joint")
    assert data["smpl_center"].shape == (2, 3) and not torch.equal(data["obj_t"].detach(), t0)
    # with an object template and the network-input masks the 'sil' phase runs in between (recon_fit_behave.py:129-136)
    import types
    from test_silhouette_gpu import uv_sphere
    v, f = uv_sphere(12, 16, 0.3)
    fit.scan = types.SimpleNamespace(v=v, f=f)
    yy, xx = torch.meshgrid(torch.arange(512.0), torch.arange(512.0), indexing="ij")
    images = torch.zeros(2, 5, 512, 512)
    images[:, 3] = ((xx - 230) ** 2 + (yy - 250) ** 2 < 60 ** 2).float()
    images[:, 4] = ((xx - 300) ** 2 + (yy - 260) ** 2 < 70 ** 2).float()
    data.pop("silhouette", None)
    data["images"] = images.to(DEV)
    logs = []
    fit.optimize_smpl_object(net, data, obj_iter=1, joint_iter=1, steps_per_iter=2, max_iter=1, sil_iter=1, log=logs.append)
    assert logs[0].startswith("object only") and logs[2].startswith("sil") and "mask" in logs[2] and "trans" in logs[2]
    assert logs[4].startswith("joint") and "rot_init" in data and data["silhouette"].image_ref.shape == (2, 256, 256)


def test_contact_loss_vs_oracle():
    """csrc/contact.cu against the restatement of compute_contact_loss + pytorch3d's default chamfer_distance
    (recon/recon_fit_base.py:553-608): value and gradients to both point sets; an image without contacts on one side pulls all
    points of that side, an image without any contact is skipped, no contact at all adds no term."""
    import chore_b200
    from chore_b200 import _lib
    g = torch.Generator().manual_seed(23)
    B, Nh, No = 4, 6890, 3000
    verts = (0.3 * torch.randn(B, Nh, 3, generator=g) + torch.tensor([0.0, 0.0, 2.2])).requires_grad_(True)
    obj = (0.3 * torch.randn(B, No, 3, generator=g) + torch.tensor([0.1, 0.0, 2.2])).requires_grad_(True)
    df_h, df_o = torch.rand(B, Nh, generator=g), torch.rand(B, No, generator=g)       # 8 % below 0.08
    df_o[1] = 1.0                                   # image 1: no contact points on the object -> all object points are pulled
    df_h[2] = 1.0                                   # image 2: none on the human
    df_h[3] = 1.0; df_o[3] = 1.0                    # image 3: no contact at all -> skipped
    part_o = torch.randn(B, 14, No, generator=g)
    part_o[0, 5] = -100.0                           # part 5 never wins on the object of image 0: that pair does not exist
    labels = torch.randint(14, (Nh,), generator=g)
    want = O.contact_loss(df_h, df_o, obj, verts, part_o, labels)
    want.backward()
    h = _lib.get_handle(torch.device(DEV))
    loss, pairs, g_s, g_o = h.contact_loss(verts.detach().to(DEV), obj.detach().to(DEV), df_h.to(DEV), df_o.to(DEV), part_o.to(DEV),
                                           labels.to(torch.int32).to(DEV))
    assert int(pairs) == 13 + 14 + 14 and rel_err(loss, want.detach()) < 1e-5, (int(pairs), float(loss), float(want))
    assert rel_err(g_s, verts.grad) < 1e-4 and rel_err(g_o, obj.grad) < 1e-4, (rel_err(g_s, verts.grad), rel_err(g_o, obj.grad))
    assert float(g_s[3].abs().max()) == 0.0 and float(g_o[3].abs().max()) == 0.0
    # through the fitter: autograd reaches the object pose; nothing is added when no contact exists
    fit = chore_b200.ReconFitterBehave(device=DEV)
    fit.part_labels = labels.to(torch.int32)
    o2 = obj.detach().to(DEV).requires_grad_(True)
    ld = {}
    fit.compute_contact_loss(df_h.to(DEV), df_o.to(DEV), o2, verts.detach().to(DEV), ld, part_o=part_o.to(DEV))
    ld["contact"].backward()
    assert rel_err(o2.grad, obj.grad) < 1e-4
    ld = {}
    fit.compute_contact_loss(torch.ones(B, Nh, device=DEV), torch.ones(B, No, device=DEV), o2, verts.detach().to(DEV), ld, part_o=part_o.to(DEV))
    assert "contact" not in ld and O.contact_loss(torch.ones(B, Nh), torch.ones(B, No), obj, verts, part_o, labels) is None


def test_head_mask_skips_heads_without_changing_results(net):
    """A head evaluated alone (the kernel skips the other heads' MMAs, epilogues and weight panels) gives bit-identical
    values to the same head evaluated with all four; `heads()` scopes the mask for the reference-style callers."""
    import chore_b200
    from chore_b200 import _lib
    from chore_b200.net import heads
    feat, tmpx = O.synth_features(91, B=2)
    set_maps(net, feat, tmpx)
    f, s = net._maps()
    pts = torch.cat([O.synth_points("frustum", 92, 2, 700), O.synth_points("init_box", 93, 2, 333)], 1).to(DEV)
    cc = torch.tensor([[1008., 995.], [990., 1001.]], device=DEV)
    full, _ = net.handle.query_fwd(f, s, pts, cc, 15)
    for mask in range(1, 15):          # every subset of the heads (3-head masks run on one issuer warp: see query_tc.cu)
        part, _ = net.handle.query_fwd(f, s, pts, cc, mask)
        for h in range(4):
            if mask & (1 << h):
                assert torch.equal(part[h], full[h]), (mask, h)
            else:
                assert part[h] is None
    with pytest.raises(chore_b200.ChoreError):
        net.handle.query_fwd(f, s, pts, cc, 0)
    with heads(net, _lib.HEAD_DF):
        p = pts.clone().requires_grad_(True)
        net.query(p, crop_center=cc)
        df, pca, parts, centers = net.get_preds()
        assert torch.equal(df, full[0]) and parts.numel() == 0 and centers.numel() == 0
        torch.clamp(df[:, 0], max=2.0).sum().backward()
    assert net.head_mask == 15
    p2 = pts.clone().requires_grad_(True)
    net.query(p2, crop_center=cc)
    torch.clamp(net.get_preds()[0][:, 0], max=2.0).sum().backward()
    assert torch.equal(p.grad, p2.grad)


def test_query_bwd_concurrent_heads_equal_sequential(net):
    """chore_query_bwd_ws: with scratch for k heads the heads run in one launch (own gX buffer each, summed in head order);
    with the base scratch they run one after the other, accumulating.  Same summation order => identical bits."""
    from chore_b200 import _lib
    feat, tmpx = O.synth_features(95, B=1)
    set_maps(net, feat, tmpx)
    f, s = net._maps()
    N = 1500
    pts = O.synth_points("frustum", 96, 1, N).to(DEV)
    cc = torch.tensor([[1008., 995.]], device=DEV)
    gen = torch.Generator().manual_seed(97)
    grads = [torch.randn(1, c, N, generator=gen).to(DEV) for c in (2, 9, 14, 6)]
    h = net.handle
    base = int(h.lib.chore_query_bwd_workspace_bytes(1, N))
    assert base == N * 384 * 4

    def run(ws_bytes, gs):
        ws = torch.empty(ws_bytes // 4, device=DEV)
        out = torch.empty(1, N, 3, device=DEV)
        rc = h.lib.chore_query_bwd_ws(h.h, f.data_ptr(), s.data_ptr(), f.shape[1], f.shape[2], pts.data_ptr(), cc.data_ptr(), 1, N,
                                      *[None if g is None else g.data_ptr() for g in gs], out.data_ptr(), ws.data_ptr(), ws_bytes,
                                      torch.cuda.current_stream().cuda_stream)
        assert rc == 0, h.lib.chore_last_error()
        torch.cuda.synchronize()
        return out

    for gs in (grads, [grads[0], None, grads[2], None], [None, None, None, grads[3]]):
        k = sum(g is not None for g in gs)
        seq, con = run(base, gs), run(base * k, gs)
        assert torch.isfinite(seq).all() and seq.abs().max() > 0
        assert torch.equal(seq, con), (k, (seq - con).abs().max().item())


def test_partial_head_masks_on_many_tiles_are_bit_identical(net):
    """Regression: with more tiles than SMs (a CTA walks several tiles) and a mask that leaves one epilogue group without
    output-layer work (single heads, {0,2}, {1,3}) that group used to skip a completion of the accumulator barrier, test the
    next tile's layer-1 phase one phase early and read a stale accumulator (nondeterministic rows in the second tile of
    a CTA).  HEAD_DF alone is what Generator.approx_surface and the fitter use on 20-30 k points = 157-235 tiles."""
    feat, tmpx = O.synth_features(5, B=2)
    set_maps(net, feat, tmpx)
    f, s = net._maps()
    cc = torch.tensor([[1008., 995.], [990., 1001.]], device=DEV)
    pts = torch.cat([O.synth_points("frustum", 7, 2, 5000), O.synth_points("init_box", 8, 2, 30001)], 1).to(DEV)
    full, _ = net.handle.query_fwd(f, s, pts, cc, 15)
    full = [x.clone() for x in full]
    for mask in (1, 2, 4, 8, 5, 10, 3, 12, 7, 14):
        for rep in range(2):
            got, _ = net.handle.query_fwd(f, s, pts, cc, mask)
            torch.cuda.synchronize()
            for h in range(4):
                if (mask >> h) & 1:
                    assert torch.equal(got[h], full[h]), (mask, h, rep, (got[h] - full[h]).abs().max().item())
                else:
                    assert got[h] is None


def test_projected_map_grid_query_matches_per_point_kernel(net, sd, monkeypatch):
    """query_g.cu (opt-in, CHORE_B200_QUERY_PRE=1): layer 1 applied once per pixel (W1 . feat, W1 . skip as 1x1 convolutions on the
    encoder's tensor-core kernel), then bilinear taps of the projected maps.  Same function as model/chore.py:107-167 up to
    the order of fp32 additions: compared with the oracle and with the per-point kernel for every kind of head mask, ragged
    ranges and both issuer layouts."""
    feat, tmpx = O.synth_features(5, B=2)
    set_maps(net, feat, tmpx)
    f, s = net._maps()
    cc = torch.tensor([[1008., 995.], [990., 1001.]], device=DEV)
    pmin, pmax = (-3.0, -0.9, 0.2), (3.0, 1.8, 4.0)

    def run(flag, res, start, count, mask, b=1):
        monkeypatch.setenv("CHORE_B200_QUERY_PRE", flag)
        total = res[0] * res[1] * res[2]
        o = [torch.zeros(c, total, device=DEV) if (mask >> i) & 1 else None for i, c in enumerate((2, 9, 14, 6))]
        net.handle.query_grid(f, s, cc, b, list(res), list(pmin), list(pmax), start, count, mask, o)
        torch.cuda.synchronize()
        return o

    worst = 0.0
    for issuers in ("2", "1"):
        monkeypatch.setenv("CHORE_B200_QUERY_PRE_ISSUERS", issuers)
        for res, start, cnt, mask in [((48, 40, 64), 0, 48 * 40 * 64, 15), ((48, 40, 64), 100, 48 * 40 * 64 - 137, 15),
                                      ((40, 36, 50), 0, 72000, 1), ((40, 36, 50), 7, 71991, 13), ((40, 36, 50), 0, 72000, 10),
                                      ((40, 36, 50), 0, 72000, 14), ((24, 20, 28), 0, 13440, 8), ((5, 4, 3), 0, 60, 15)]:
            ref = run("0", res, start, cnt, mask)
            got = run("1", res, start, cnt, mask)
            for a, b in zip(got, ref):
                if a is not None:
                    e = rel_err(a, b)
                    worst = max(worst, e)
                    assert e < 2e-5, (issuers, res, start, cnt, mask, e)
                    assert torch.equal(a[:, :start], b[:, :start]) and torch.equal(a[:, start + cnt:], b[:, start + cnt:])   # untouched outside the range
    # and against the oracle itself
    res = (12, 10, 9)
    coords = torch.from_numpy(O.create_grid(res, list(pmin), list(pmax)).T.astype(np.float32)).unsqueeze(0)
    with torch.no_grad():
        ref = O.query(sd, feat[1:2], tmpx[1:2], coords, cc[1:2].cpu())
    got = run("1", res, 0, res[0] * res[1] * res[2], 15)
    for a, b in zip(got, ref[:4]):
        assert rel_err(a, b.reshape(a.shape)) < TOL
    report("projected_map_grid_query", worst_vs_per_point_kernel=worst)


def test_cta_pair_kernel_is_bit_identical(net, monkeypatch):
    """query_tc2_kernel (tcgen05 cta_group::2: two CTAs of a cluster share every weight panel, layer 1 as N = 256 MMAs; opt-in
    with CHORE_B200_QUERY_2CTA=1) must reproduce query_tc_kernel bit for bit: same operands, same accumulation order.  Covers an
    odd tile count (dead second tile of the last pair), B > 1, out-of-image points and the dense-grid entry point."""
    feat, tmpx = O.synth_features(101, B=2)
    set_maps(net, feat, tmpx)
    f, s = net._maps()
    cc = torch.tensor([[1008., 995.], [990., 1001.]], device=DEV)
    for N in (1, 129, 1000, 20000):
        pts = torch.cat([O.synth_points("frustum", 102 + N, 2, N - N // 3), O.synth_points("init_box", 103, 2, N // 3)], 1).to(DEV)
        monkeypatch.setenv("CHORE_B200_QUERY_2CTA", "0")
        ref, m0 = net.handle.query_fwd(f, s, pts, cc, 15, want_in_img=True)
        monkeypatch.setenv("CHORE_B200_QUERY_2CTA", "1")
        got, m1 = net.handle.query_fwd(f, s, pts, cc, 15, want_in_img=True)
        torch.cuda.synchronize()
        assert torch.equal(m0, m1)
        for h in range(4):
            assert torch.equal(got[h], ref[h]), (N, h, (got[h] - ref[h]).abs().max().item())
    res, pmin, pmax = (24, 20, 28), (-3.0, -0.9, 0.2), (3.0, 1.8, 4.0)
    total = res[0] * res[1] * res[2]
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("CHORE_B200_QUERY_2CTA", flag)
        o = [torch.zeros(c, total, device=DEV) for c in (2, 9, 14, 6)]
        net.handle.query_grid(f, s, cc, 1, res, pmin, pmax, 100, total - 100, 15, o)
        torch.cuda.synchronize()
        outs[flag] = o
    for a, b in zip(outs["0"], outs["1"]):
        assert torch.equal(a, b)


def test_fused_surface_projection_matches_autograd_path(net):
    """Generator.approx_surface (recon/generator.py:50-79): the autograd-free kernel sequence (query -> clamp gradient ->
    query adjoint -> normalised step) against the reference's autograd formulation on the same kernels.  One step is compared
    tightly (identical inputs); F.normalize's norm may differ in the last bit, so a few steps are compared with a loose bound
    on all but a vanishing fraction of the points (the projection is discontinuous across texel / ReLU boundaries)."""
    import chore_b200
    feat, tmpx = O.synth_features(111, B=2)
    set_maps(net, feat, tmpx)
    cc = torch.tensor([[1008., 995.], [1000., 990.]], device=DEV)
    q = {"crop_center": cc}
    pts = O.synth_points("frustum", 112, 2, 3000).to(DEV)
    fused, plain = chore_b200.Generator(net, threshold=2.0, device=DEV, fused=True), chore_b200.Generator(net, threshold=2.0, device=DEV, fused=False)
    for df_type in ("human", "object"):
        a, pa = fused.approx_surface(net, pts.clone().requires_grad_(True), 1, q, df_type)
        b, pb = plain.approx_surface(net, pts.clone().requires_grad_(True), 1, q, df_type)
        assert a.requires_grad and a.shape == b.shape
        assert rel_err(a, b) < 1e-5, rel_err(a, b)
        for x, y in zip(pa, pb):
            assert torch.equal(x, y)                       # the returned preds are the last query's, all heads
        a3, _ = fused.approx_surface(net, pts.clone().requires_grad_(True), 3, q, df_type)
        b3, _ = plain.approx_surface(net, pts.clone().requires_grad_(True), 3, q, df_type)
        close = ((a3 - b3).norm(dim=-1) < 1e-3).float().mean().item()
        assert close > 0.97, close
