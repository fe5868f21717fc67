"""Host-side mirrors of the fitter's non-field terms, on CPU tensors (no kernel involved), against the reference's own run
(tests/golden/fit_smpl_full.npz) and the oracle: Mahalanobis / hand priors incl. the reference's (1,45) hand-prior shape,
smplz, the 2-D keypoint projection loss, loss weights + sum_dict, and the asset loaders."""
import os

import numpy as np
import pytest
import torch

from conftest import golden_smpl_assets, load_golden, rel_err
from oracle import chore_oracle as O
from oracle import ref_shim

T = lambda x: torch.from_numpy(np.asarray(x))


@pytest.fixture(scope="module")
def gold():
    return load_golden("fit_smpl_full.npz")


@pytest.fixture(scope="module")
def fitter(gold):
    import chore_b200
    from chore_b200.fitter import HandPrior, MahalanobisPrior
    _, pri = golden_smpl_assets(gold)
    bp = MahalanobisPrior(pri["body_mean"], pri["body_prec"], device="cpu")
    hp = HandPrior(pri["hand_mean"], pri["lh_prec"], pri["rh_prec"], device="cpu")
    return chore_b200.ReconFitterBehave(device="cpu", priors=(bp, hp))


def test_priors_match_reference_run(gold, fitter):
    pose = T(gold["pose"])
    bp, hp = fitter.priors
    assert hp(pose).shape == (1, 45)                     # the reference's shape: summed over batch rows and both hands
    assert rel_err(torch.mean(bp(pose[:, :72])), gold["loss_pose"]) < 1e-5
    assert rel_err(torch.mean(hp(pose)), gold["loss_hand"]) < 1e-5
    _, pri = golden_smpl_assets(gold)
    assert torch.allclose(bp(pose[:, :72]), O.mahalanobis(pose[:, :72], T(pri["body_mean"]), T(pri["body_prec"])), rtol=1e-6)
    assert torch.allclose(hp(pose), O.hand_prior(pose, T(pri["hand_mean"]), T(pri["lh_prec"]), T(pri["rh_prec"])), rtol=1e-6)
    smpl = type("S", (), {"pose": pose, "betas": T(gold["betas"])})()
    d = {}
    fitter.compute_prior_loss(d, smpl, nobeta=False)
    assert set(d) == {"beta", "pose", "hand"} and rel_err(d["beta"], torch.mean(T(gold["betas"]) ** 2)) < 1e-6


def test_landmark_terms_match_reference_run(gold, fitter):
    J, cc, kp = T(gold["J"]), T(gold["crop_center"]), T(gold["body_kpts"])
    d = {}
    fitter.smplz_loss(J, d)
    assert rel_err(d["smplz"], gold["loss_smplz"]) < 1e-5
    assert rel_err(fitter.projection_loss(J, kp, cc), gold["loss_j2d"]) < 1e-5
    assert torch.allclose(fitter.project_points(J, cc), O.project_to_input_image(J, cc), rtol=1e-6, atol=1e-4)
    # project_points without a crop centre: raw Kinect pixels scaled to the network input
    raw = fitter.project_points(J)
    assert raw.shape == (2, 25, 2) and torch.isfinite(raw).all()


def test_loss_weights_and_sum_dict_match_reference_run(gold, fitter):
    losses = {str(k): T(gold[f"loss_{k}"]) for k in gold["loss_order"]}
    w = fitter.get_loss_weights()
    total = fitter.sum_dict(losses, w, float(gold["decay"]))
    assert rel_err(total, gold["total"]) < 1e-6
    assert set(w) == set(O.LOSS_WEIGHTS) and all(abs(float(w[k](torch.tensor(1.0), 0)) - O.LOSS_WEIGHTS[k]) < 1e-6 * O.LOSS_WEIGHTS[k] for k in w)
    line = fitter.get_loss_str("3-1", losses, w, float(gold["decay"]))
    assert line.startswith("Iter: 3-1, df_h: ") and line.count(",") == len(losses)


@pytest.mark.skipif(not ref_shim.available(), reason="reference assets not present")
def test_asset_loaders_read_the_reference_files(gold):
    from chore_b200.fitter import load_priors
    from chore_b200.smpl import _to_csr, load_regressors
    root = os.path.join(ref_shim.REF_ROOT, "assets")
    regs = load_regressors(root)
    want, pri = golden_smpl_assets(gold)
    for r, w in zip(regs, want):
        m = _to_csr(r)
        assert m.shape == w.shape and (abs(m - w)).max() < 1e-7
    bp, hp = load_priors(root, device="cpu")
    assert rel_err(bp.mean.reshape(-1), pri["body_mean"]) < 1e-6 and rel_err(bp.prec, pri["body_prec"]) < 1e-6
    assert rel_err(hp.mean.reshape(-1), pri["hand_mean"]) < 1e-6 and rel_err(hp.lhand_prec[0], pri["lh_prec"]) < 1e-6


def test_copy_smpl_params_follows_the_reference():
    import chore_b200
    ns = lambda **k: type("S", (), k)()
    P = lambda *s: torch.nn.Parameter(torch.randn(*s))
    split = ns(global_pose=P(2, 3), body_pose=P(2, 63), hand_pose=P(2, 90), top_betas=P(2, 2), other_betas=P(2, 8), trans=P(2, 3))
    smpl = ns(pose=P(2, 156), betas=P(2, 10), trans=P(2, 3))
    b_old = smpl.betas.data.clone()
    chore_b200.ReconFitterBase.copy_smpl_params(split, smpl)
    assert torch.equal(smpl.pose.data, torch.cat([split.global_pose, split.body_pose, split.hand_pose], 1).data)
    assert torch.equal(smpl.betas.data[:, :2], split.top_betas.data) and torch.equal(smpl.betas.data[:, 2:], b_old[:, 2:])
    assert torch.equal(smpl.trans.data, split.trans.data)
