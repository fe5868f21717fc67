"""Generate tests/golden/*.npz by running the REAL reference (/root/reference, imported
read-only through oracle/ref_shim.py) on seeded synthetic inputs.  TEST INFRASTRUCTURE ONLY.

Run in the build container (the GPU box has no /root/reference):

    python -m oracle.make_golden            # writes tests/golden/*.npz

The reference ships no golden vectors for this path, so these files
are what pins oracle/chore_oracle.py (tests/test_oracle.py) and, through it, the CUDA path.
Large inputs (weights, feature maps, SMPL-H-shaped buffers) are not stored: they are
re-generated from the recorded seeds by the same torch CPU generator calls used here
(`oracle.chore_oracle.make_state_dict`, `make_smplh_buffers`, `synth_*` below), and a checksum
of each regenerated input is stored so a generator drift is detected rather than silently
compared against.

Reference entry points exercised (paths relative to /root/reference):
  model/chore.py:87-96,107-167   CHORE.filter / CHORE.query / get_preds
  model/HGFilters.py:144-185     HGFilter.forward
  recon/generator.py:50-79       Generator.approx_surface
  lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:72-175   SMPL_Layer.forward
  recon/recon_fit_base.py:167-188,367-384   project_so3 / transform_obj_verts / decopose_axis
  recon/recon_fit_behave.py:165-222          forward_step(phase='object only')
  recon/recon_fit_behave.py:293-337          forward_smpl(phase='kpts'), with compute_prior_loss / smplz_loss /
                                             compute_kpts_loss (recon_fit_base.py:230-231,522-535,653-676),
                                             get_landmarks (lib_smpl/wrapper_pytorch.py:176-190)
  recon/generator.py:123-217,275-282         Generator.gen_pc_batch / compose_outdict / init_samples (closed-form field)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

from . import chore_oracle as O
from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ------------------------------------------------------------------------------------------
# seeded synthetic inputs (shared with tests / bench through oracle.chore_oracle re-exports)
# ------------------------------------------------------------------------------------------
def checksum(t: torch.Tensor) -> np.ndarray:
    t = t.detach().double().reshape(-1)
    idx = torch.arange(t.numel(), dtype=torch.float64)
    return np.array([t.sum().item(), (t * torch.cos(idx)).sum().item(), t.abs().max().item()])


def save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **conv)
    print(f"wrote {path}: {os.path.getsize(path) / 1e3:.1f} kB")


def gold_query(net):
    """CHORE.query fwd + gradient to the points on synthetic feature maps, B=2."""
    seed = 11
    sd = O.make_state_dict(0, "unit")
    net.load_state_dict(sd)
    feat, tmpx = O.synth_features(seed, B=2)
    pts = torch.cat([O.synth_points("init_box", seed, 2, 1024), O.synth_points("frustum", seed + 1, 2, 1024)], 1)
    # a few hand-made edge cases: exactly on the image border, behind/at the camera plane, far outside
    cc = torch.tensor([[1008., 995.], [900., 1100.]])
    edge = torch.tensor([[0.0, 0.0, 2.2], [5.0, 0.0, 2.2], [0.0, -4.0, 2.2], [0.3, 0.2, 1e-3],
                         [0.1, 0.1, -1.0], [-0.0102, 0.5288, 2.2], [1e3, 1e3, 2.2], [0.6, -0.4, 2.0]])
    pts[:, :8] = edge
    net.im_feat_list, net.tmpx = [feat], tmpx
    p = pts.clone().requires_grad_(True)
    net.query(p, crop_center=cc)
    df, pca, parts, centers = net.get_preds()
    g = torch.Generator().manual_seed(seed + 2)
    gd = {k: torch.randn(v.shape, generator=g) for k, v in
          (("df", df), ("pca", pca), ("parts", parts), ("centers", centers))}
    (gd["df"] * df).sum().add((gd["pca"] * pca).sum()).add((gd["parts"] * parts).sum()).add(
        (gd["centers"] * centers).sum()).backward()
    grad_all = p.grad.clone()
    # the approx_surface pattern: clamp(df_h, max=2).sum().backward()
    p2 = pts.clone().requires_grad_(True)
    net.query(p2, crop_center=cc)
    torch.clamp(net.get_preds()[0][:, 0, :], max=2.0).sum().backward()
    xyz = net.camera.project_points(pts, cc) if hasattr(net, "camera") else O.project_points(pts, cc)
    save("query.npz", seed=seed, weights_seed=0, points=pts, crop_center=cc, df=df, pca=pca, parts=parts,
         centers=centers, g_df=gd["df"], g_pca=gd["pca"], g_parts=gd["parts"], g_centers=gd["centers"],
         grad_all=grad_all, grad_dfh=p2.grad, proj=O.project_points(pts, cc),
         feat_ck=checksum(feat), tmpx_ck=checksum(tmpx), w_ck=checksum(sd["df.0.weight"]))


def gold_encoder(net):
    """HGFilter through CHORE.filter: a 2x5x128x128 batch stored in full, and the 1x5x512x512
    production shape stored on a stride-8 pixel lattice."""
    sd = O.make_state_dict(0, "unit")
    net.load_state_dict(sd)
    img = O.synth_images(21, B=2, size=128)
    with torch.no_grad():
        net.filter(img)
        feat, tmpx, normx = net.im_feat_list[-1].clone(), net.tmpx.clone(), net.normx.clone()
    save("encoder_128.npz", seed=21, weights_seed=0, feat=feat, tmpx=tmpx, normx=normx[:, :, ::4, ::4],
         img_ck=checksum(img))
    img = O.synth_images(22, B=1, size=512)
    with torch.no_grad():
        net.filter(img)
        feat, tmpx = net.im_feat_list[-1], net.tmpx
    save("encoder_512.npz", seed=22, weights_seed=0, feat_s8=feat[:, :, ::8, ::8], tmpx_s8=tmpx[:, :, ::8, ::8],
         feat_ck=checksum(feat), tmpx_ck=checksum(tmpx), img_ck=checksum(img))
    # reference-faithful init (N(0, 0.02), zero bias): magnitudes only, as a second weight family
    sd2 = O.make_state_dict(1, "ref_init")
    net.load_state_dict(sd2)
    img = O.synth_images(23, B=1, size=128)
    with torch.no_grad():
        net.filter(img)
    save("encoder_128_refinit.npz", seed=23, weights_seed=1, feat=net.im_feat_list[-1], tmpx=net.tmpx)


def gold_approx_surface(net):
    """Generator.approx_surface, 10 steps, both distance fields (recon/generator.py:50-79)."""
    Generator = ref_shim.load_generator_class()
    sd = O.make_state_dict(0, "unit")
    net.load_state_dict(sd)
    feat, tmpx = O.synth_features(31, B=1)
    net.im_feat_list, net.tmpx = [feat], tmpx
    cc = torch.tensor([[1008., 995.]])
    fake = types.SimpleNamespace(threshold=2.0)
    out = {}
    for name in ("human", "object"):
        s = O.synth_points("frustum", 32, 1, 512).requires_grad_(True)
        samples, preds = Generator.approx_surface(fake, net, s, 10, {"crop_center": cc}, name)
        out[f"samples_{name}"] = samples.detach()
        out[f"df_{name}"] = preds[0].detach()
    save("approx_surface.npz", seed=31, weights_seed=0, crop_center=cc, **out)


def gold_lbs():
    """SMPL_Layer.forward on synthetic SMPL-H-shaped buffers + gradients to pose/betas/trans."""
    buf = O.make_smplh_buffers(0)
    layer = ref_shim.load_smpl_layer(buf)
    g = torch.Generator().manual_seed(41)
    B = 2
    pose = (0.2 * torch.randn(B, 156, generator=g)).requires_grad_(True)
    pose.data[1, 6:9] = 0.0                       # an exactly-zero joint rotation (the 1e-8 branch)
    betas = torch.randn(B, 10, generator=g).requires_grad_(True)
    trans = (torch.tensor([[0.0, 0.0, 2.2]]) + 0.1 * torch.randn(B, 3, generator=g)).requires_grad_(True)
    offsets = (0.003 * torch.randn(B, 6890, 3, generator=g)).requires_grad_(True)
    verts, jtr, v_posed, naked = layer(pose, th_betas=betas, th_trans=trans, th_offsets=offsets)
    g_verts = torch.randn(verts.shape, generator=g)
    g_jtr = torch.randn(jtr.shape, generator=g)
    ((g_verts * verts).sum() + (g_jtr * jtr).sum()).backward()
    save("lbs.npz", seed=41, buffers_seed=0, pose=pose, betas=betas, trans=trans, offsets=offsets,
         verts=verts, jtr=jtr, v_posed_s=v_posed[:, ::10], naked_s=naked[:, ::10],
         g_verts=g_verts, g_jtr=g_jtr, grad_pose=pose.grad, grad_betas=betas.grad, grad_trans=trans.grad,
         grad_offsets_s=offsets.grad[:, ::10], posedirs_ck=checksum(buf["posedirs"]),
         weights_ck=checksum(buf["weights"]))


def gold_rigid_and_fit(net):
    """transform_obj_verts / project_so3 / decopose_axis and one 'object only' forward_step with
    backward to (R, t, s) (recon/recon_fit_behave.py:165-198)."""
    Fitter = ref_shim.load_fitter_class()
    fit = object.__new__(Fitter)
    fit.debug = False
    fit.obj_scale = 1.0
    g = torch.Generator().manual_seed(51)
    B, No = 2, 3000
    obj = 0.2 * torch.randn(B, No, 3, generator=g)
    rot = (torch.eye(3).unsqueeze(0) + 0.2 * torch.randn(B, 3, 3, generator=g)).requires_grad_(True)
    t = (torch.tensor([[0.2, 0.1, 2.3]]) + 0.05 * torch.randn(B, 3, generator=g)).requires_grad_(True)
    s = (1.0 + 0.05 * torch.randn(B, generator=g)).requires_grad_(True)
    R0 = Fitter.project_so3(rot.detach())
    moved = fit.transform_obj_verts(obj, R0, t.detach(), s.detach())
    bad = torch.eye(3).unsqueeze(0).repeat(2, 1, 1)
    bad[1, 2, 2] = -1.0                            # det < 0 input
    save("rigid.npz", seed=51, obj=obj, rot=rot, t=t, s=s, R=R0, moved=moved, bad=bad, R_bad=Fitter.project_so3(bad))

    sd = O.make_state_dict(0, "unit")
    net.load_state_dict(sd)
    feat, tmpx = O.synth_features(52, B=B)
    net.im_feat_list, net.tmpx = [feat], tmpx
    cc = torch.tensor([[1008., 995.], [1000., 980.]])
    smpl_center = torch.tensor([[0.0, 0.1, 2.2], [0.05, 0.0, 2.25]])
    data = {"objects": obj, "query_dict": {"crop_center": cc}, "smpl_center": smpl_center}
    torch.manual_seed(53)
    noise = torch.rand(B, 3, 3)                    # what decopose_axis will draw (recon_fit_base.py:384)
    torch.manual_seed(53)
    smpl = lambda: (torch.zeros(B, 6890, 3), None, None, None)
    losses = fit.forward_step(net, smpl, data, rot, t, s, "object only")
    import recon.recon_fit_behave  # noqa: F401  (already imported by load_fitter_class)
    wts = fit.get_loss_weights()
    total = fit.sum_dict(losses, wts, 3)
    total.backward()
    save("fit_object_only.npz", seed=52, weights_seed=0, obj=obj, rot=rot, t=t, s=s, crop_center=cc,
         smpl_center=smpl_center, noise=noise, it=3, loss_object=losses["object"], loss_scale=losses["scale"],
         loss_ocent=losses["ocent"], total=total, grad_rot=rot.grad, grad_t=t.grad, grad_s=s.grad)


def load_reference_assets():
    """Landmark regressors and priors of the reference (assets/*.pkl) as plain arrays: the inputs of
    fit_smpl_full.npz (stored in it, because /root/reference does not exist on the GPU box)."""
    import pickle as pkl
    import scipy.sparse as sp
    root = os.path.join(ref_shim.REF_ROOT, "assets")
    regs = []
    for n in ("body25_regressor.pkl", "face_regressor.pkl", "hand_regressor.pkl"):
        m = sp.csr_matrix(pkl.load(open(os.path.join(root, n), "rb"), encoding="latin1").T)
        m.sum_duplicates(); m.sort_indices()
        regs.append(m)
    rd = lambda n: pkl.load(open(os.path.join(root, "priors", n), "rb"), encoding="latin1")
    body, lh, rh = rd("body_prior.pkl"), rd("lh_prior.pkl"), rd("rh_prior.pkl")
    pri = {"body_mean": np.asarray(body["mean"], np.float32), "body_prec": np.asarray(body["precision"], np.float32),
           "hand_mean": np.concatenate([lh["mean"], rh["mean"]]).astype(np.float32),
           "lh_prec": np.asarray(lh["precision"], np.float32), "rh_prec": np.asarray(rh["precision"], np.float32)}
    return regs, pri


def gold_fit_smpl_full(net):
    """ReconFitterBehave.forward_smpl, phase 'kpts' (every term: df_h, pose / hand priors, part CE, smplz, pinit,
    j2d), sum_dict with decay it/3 and backward to the split SMPL parameters -- recon/recon_fit_behave.py:265-271,
    293-337.  The reference's own modules run unmodified; only the SMPL-H pickle is replaced by synthetic buffers
    and `.cuda()` by a no-op (oracle/ref_shim.py)."""
    import types as _t
    Fitter = ref_shim.load_fitter_class()
    from model.camera import KinectColorCamera
    fit = object.__new__(Fitter)
    fit.debug, fit.z_0, fit.obj_scale = False, 2.2, 1.0
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        fit.camera = KinectColorCamera(1200)
    fit.net_in_size = 512
    B = 2
    buf = O.make_smplh_buffers(0)
    g = torch.Generator().manual_seed(61)
    pose = 0.15 * torch.randn(B, 156, generator=g)
    betas = 0.5 * torch.randn(B, 10, generator=g)
    trans = torch.tensor([[0.0, 0.0, 2.2]]) + 0.05 * torch.randn(B, 3, generator=g)
    smpl = ref_shim.load_split_smpl(buf, B, pose, betas, trans)
    sd = O.make_state_dict(0, "unit")
    net.load_state_dict(sd)
    feat, tmpx = O.synth_features(62, B=B)
    net.im_feat_list, net.tmpx = [feat], tmpx
    cc = torch.tensor([[1008., 995.], [1000., 980.]])
    part_labels = torch.randint(14, (B, 6890), generator=g)
    pose_init = pose[:, 3:72] + 0.05 * torch.randn(B, 69, generator=g)
    kpts = torch.cat([512 * torch.rand(B, 25, 2, generator=g), torch.rand(B, 25, 1, generator=g)], -1)
    data = {"net": net, "part_labels": part_labels, "pose_init": pose_init, "body_kpts": kpts,
            "query_dict": {"crop_center": cc}}
    it = 25
    with ref_shim.cpu_priors():
        losses = fit.forward_smpl(smpl, data, "kpts")
    total = fit.sum_dict(losses, fit.get_loss_weights(), it / 3)
    total.backward()
    J, face, hands = smpl.get_landmarks()
    regs, pri = load_reference_assets()
    save("fit_smpl_full.npz", seed=61, weights_seed=0, buffers_seed=0, feat_seed=62, pose=pose, betas=betas, trans=trans,
         crop_center=cc, part_labels=part_labels, pose_init=pose_init, body_kpts=kpts, decay=it / 3,
         **{f"loss_{k}": v for k, v in losses.items()}, total=total, loss_order=np.array(list(losses)),
         grad_trans=smpl.trans.grad, grad_global_pose=smpl.global_pose.grad, grad_body_pose=smpl.body_pose.grad,
         grad_hand_pose=smpl.hand_pose.grad, grad_top_betas=smpl.top_betas.grad, grad_other_betas=smpl.other_betas.grad,
         J=J, face=face, hands=hands,
         **{f"reg{i}_{a}": getattr(m, a) for i, m in enumerate(regs) for a in ("indptr", "indices", "data")},
         **{f"prior_{k}": v for k, v in pri.items()})


def gold_gen_pc_batch():
    """Generator.gen_pc_batch + compose_outdict (recon/generator.py:123-217) of the REFERENCE, driven by the
    closed-form field of oracle/analytic_field.py on the CPU: surface filter, CPU-generator index resampling,
    min-count truncation, argmax / mean reductions."""
    from .analytic_field import AnalyticField
    Generator = ref_shim.load_generator_class()
    gen = object.__new__(Generator)
    gen.threshold, gen.filter_val, gen.device = 2.0, 0.004, "cpu"
    out = {}
    for df_type, seed in (("human", 71), ("object", 72)):
        torch.manual_seed(seed)
        init = gen.init_samples(3000, batch_size=2)
        init[1] = init[0].flip(0)                        # the reference rescales batch element 0 only (kept quirk)
        res = gen.gen_pc_batch(AnalyticField(), df_type, init, 25000, {"crop_center": torch.tensor([[1008., 995.]] * 2)}, 10, mute=True)
        out[f"{df_type}_init"] = init
        for k, v in res.items():
            out[f"{df_type}_{k}"] = v
    save("gen_pc_batch.npz", seed_human=71, seed_object=72, num_points=25000, num_steps=10, **out)


def strided(t, step):
    """every `step`-th point (last axis) of a (B, C, N) / (B, 3, 3, N) prediction"""
    return t[..., ::step]


def gold_query_grid256(net):
    """BASELINE config 2 at its real size: three x-planes (ix = 40, 128, 215; 3 x 65 536 points) of the 256^3 grid
    create_grid(256, 256, 256, pmin, pmax) (model/sdf.py:4-27, bounds recon/generator.py:45-48) through CHORE.query.
    df is stored for every point, the other heads on every 8th point + the part argmax of every point."""
    install_sdf = ref_shim.install
    install_sdf()
    with ref_shim.ref_cwd():
        from model.sdf import create_grid
    sd = O.make_state_dict(0, "unit")
    net.load_state_dict(sd)
    feat, tmpx = O.synth_features(91, B=1)
    net.im_feat_list, net.tmpx = [feat], tmpx
    cc = torch.tensor([[1008., 995.]])
    res = (256, 256, 256)
    bmin, bmax = np.array([-3.0, -0.9, 0.2]), np.array([3.0, 1.8, 4.0])
    coords, _ = create_grid(*res, bmin, bmax)
    planes = (40, 128, 215)
    pts = torch.from_numpy(np.concatenate([coords[:, ix].reshape(3, -1) for ix in planes], 1).T.copy()).float().unsqueeze(0)
    with torch.no_grad():
        net.query(pts, crop_center=cc)
        df, pca, parts, centers = net.get_preds()
    save("query_grid256.npz", seed=91, weights_seed=0, crop_center=cc, res=np.array(res), bmin=bmin, bmax=bmax,
         planes=np.array(planes), df=df, parts_argmax=parts.argmax(1).to(torch.uint8), stride=8,
         pca_s=strided(pca, 8), parts_s=strided(parts, 8), centers_s=strided(centers, 8),
         top2_margin=(parts.topk(2, dim=1).values[:, 0] - parts.topk(2, dim=1).values[:, 1]).half(),
         feat_ck=checksum(feat), tmpx_ck=checksum(tmpx))


def gold_query_b32(net):
    """BASELINE config 4 batch shape: 32 images x 1 024 points in ONE CHORE.query call (per-image crop centres)."""
    sd = O.make_state_dict(0, "unit")
    net.load_state_dict(sd)
    B, N = 32, 1024
    feat, tmpx = O.synth_features(95, B=B, hw=64)        # 256 x 256 inputs: 32 full-size maps would be 2 GB of fixtures to regenerate
    net.im_feat_list, net.tmpx = [feat], tmpx
    g = torch.Generator().manual_seed(96)
    cc = torch.tensor([[1008., 995.]]) + 40 * torch.randn(B, 2, generator=g)
    pts = torch.cat([O.synth_points("init_box", 97, B, N // 2), O.synth_points("frustum", 98, B, N // 2, cc)], 1)
    with torch.no_grad():
        net.query(pts, crop_center=cc)
        df, pca, parts, centers = net.get_preds()
    save("query_b32.npz", seed=95, weights_seed=0, crop_center=cc, points_ck=checksum(pts), df=df,
         parts_argmax=parts.argmax(1).to(torch.uint8), stride=8, pca_s=strided(pca, 8), parts_s=strided(parts, 8),
         centers_s=strided(centers, 8), feat_ck=checksum(feat))


def gold_testdata_example(net):
    """The shipped demo frame (example/000000117377/k1.*) through the reference's TestData.prepare_image_crop
    (data/test_data.py:59-125) and then through CHORE.filter: the realistic (non white-noise) input of the suite.
    The frame itself is committed under tests/golden/example_frame/ (data fixture of the reference, 425 kB)."""
    import shutil, tempfile
    TestData = ref_shim.load_testdata_class()
    tmp = tempfile.mkdtemp()
    src = os.path.join(ref_shim.REF_ROOT, "example", "000000117377")
    for f in os.listdir(src):
        shutil.copy(os.path.join(src, f), tmp)
    rgb = os.path.join(tmp, "k1.color.jpg")
    out = {}
    for mean_center in (False, True):
        ds = TestData([rgb], 1, 0, image_size=(512, 512), crop_size=1200, use_mean_center=mean_center)
        item = ds.get_item(0)
        tag = "mc" if mean_center else "own"
        out[f"images_{tag}_s4"] = item["images"][:, ::4, ::4]
        out[f"images_{tag}_ck"] = checksum(torch.from_numpy(item["images"]))
        out[f"crop_center_{tag}"] = item["crop_center"]
        out[f"old_crop_center_{tag}"] = np.asarray(item["old_crop_center"])
        out[f"resize_scale_{tag}"] = item["resize_scale"]
        out[f"crop_scale_{tag}"] = item["crop_scale"]
        if not mean_center:
            import pickle as pkl
            info = pkl.load(open(rgb.replace(".color.jpg", ".crop_info.pkl"), "rb"))
            out.update({f"crop_info_{k}": np.asarray(v) for k, v in info.items()})
            images = torch.from_numpy(item["images"]).unsqueeze(0)
    sd = O.make_state_dict(0, "unit")
    net.load_state_dict(sd)
    with torch.no_grad():
        net.filter(images)
        feat, tmpx = net.im_feat_list[-1], net.tmpx
    # a fitting-style query on the real-image features: SMPL-like points around the person at z0
    cc = torch.from_numpy(out["crop_center_own"]).float().unsqueeze(0)
    pts = O.synth_points("frustum", 99, 1, 2048, cc)
    with torch.no_grad():
        net.query(pts, crop_center=cc)
        df, pca, parts, centers = net.get_preds()
    save("example_frame.npz", weights_seed=0, feat_s8=feat[:, :, ::8, ::8], tmpx_s8=tmpx[:, :, ::8, ::8], feat_ck=checksum(feat),
         tmpx_ck=checksum(tmpx), points=pts, df=df, pca=pca, parts=parts, centers=centers, **out)
    shutil.rmtree(tmp)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    if not ref_shim.available():
        sys.exit("reference tree not available: goldens can only be generated in the build container")
    net, _ = ref_shim.load_chore()
    with ref_shim.ref_cwd():
        gold_query(net)
        gold_encoder(net)
        gold_approx_surface(net)
        gold_lbs()
        gold_rigid_and_fit(net)
        gold_fit_smpl_full(net)
        gold_gen_pc_batch()
        gold_query_grid256(net)
        gold_query_b32(net)
        gold_testdata_example(net)


if __name__ == "__main__":
    if len(sys.argv) > 1:          # python -m oracle.make_golden gold_query_b32 ...: regenerate selected files only
        torch.manual_seed(0)
        torch.set_num_threads(max(1, os.cpu_count() or 1))
        net, _ = ref_shim.load_chore()
        with ref_shim.ref_cwd():
            for name in sys.argv[1:]:
                fn = globals()[name]
                fn() if name in ("gold_lbs", "gold_gen_pc_batch") else fn(net)
    else:
        main()
