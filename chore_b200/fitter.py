"""Host-side mirror of the fitting step of the reference on the CUDA kernels.

Mirrors, with the same names and argument meaning (paths relative to /root/reference):
  ReconFitterBase.project_so3 / decopose_axis / transform_obj_verts / transform_object
      recon/recon_fit_base.py:167-188,361-384
  ReconFitterBase.sum_dict, compute_obj_loss, compute_df_h_loss, compute_smpl_center_pred
      recon/recon_fit_base.py:351-359,513-551
  ReconFitterBase.compute_prior_loss, smplz_loss, compute_kpts_loss, project_points, projection_loss,
      split_smpl, copy_smpl_params, get_smpl_bbox, get_smpl_height, get_loss_str
      recon/recon_fit_base.py:230-231,345-349,522-535,653-700
  th_Mahalanobis / Prior / HandPrior   lib_smpl/th_smpl_prior.py:19-60, lib_smpl/th_hand_prior.py:20-78
  ReconFitterBehave.get_loss_weights, forward_step ('object only'), forward_smpl (all terms),
      optimize_smpl, optimize_smpl_object ('object only' phase)
      recon/recon_fit_behave.py:90-163,165-222,224-358
The joint-phase contact term (pytorch3d Chamfer in the reference) is csrc/contact.cu; the silhouette phase lives in
silhouette.py; the interpenetration term (torch-mesh-isect, not vendored by the reference) is a plug-in hook.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib
from .net import heads


class _RigidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, R, t, s, handle):
        args = [x.detach().contiguous().float() for x in (verts, R, t, s)]
        ctx.handle = handle
        ctx.save_for_backward(*args)
        return handle.rigid_fwd(*args)

    @staticmethod
    def backward(ctx, g_out):
        verts, R, t, s = ctx.saved_tensors
        g_R, g_t, g_s, g_v = ctx.handle.rigid_bwd(verts, R, t, s, g_out, ctx.needs_input_grad[0])
        return g_v, g_R, g_t, g_s, None


class _So3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mats, handle):
        m = mats.detach().contiguous().float()
        ctx.handle = handle
        ctx.save_for_backward(m)
        return handle.project_so3(m)

    @staticmethod
    def backward(ctx, g_out):
        (m,) = ctx.saved_tensors
        return ctx.handle.project_so3_bwd(m, g_out), None


class _ContactFn(torch.autograd.Function):
    """chore_contact_loss with its closed-form gradients to the object points and the SMPL vertices."""

    @staticmethod
    def forward(ctx, obj, smpl_verts, df_hum_o, df_obj_h, part_o, part_labels, handle):
        args = [x.detach().contiguous().float() for x in (smpl_verts, obj, df_hum_o, df_obj_h, part_o)]
        loss, pairs, g_s, g_o = handle.contact_loss(*args, part_labels.to(torch.int32).contiguous())
        ctx.save_for_backward(g_s, g_o)
        ctx.mark_non_differentiable(pairs)
        return loss[0], pairs

    @staticmethod
    def backward(ctx, g_loss, _g_pairs):
        g_s, g_o = ctx.saved_tensors
        return g_o * g_loss, g_s * g_loss, None, None, None, None, None


class MahalanobisPrior:
    """th_Mahalanobis (lib_smpl/th_smpl_prior.py:25-44): ||(pose[:, prefix:end] - mean) @ prec||^2 per batch row."""

    def __init__(self, mean, prec, prefix: int = 3, end: int = 66, device="cuda:0"):
        self.mean = torch.as_tensor(mean, dtype=torch.float32).reshape(1, -1).to(device)
        self.prec = torch.as_tensor(prec, dtype=torch.float32).to(device)
        self.prefix, self.end = prefix, end

    def __call__(self, pose, prior_weight: float = 1.0):
        temp = pose[:, self.prefix:self.end] - self.mean
        temp2 = torch.matmul(temp, self.prec) * prior_weight
        return (temp2 * temp2).sum(dim=1)


class HandPrior:
    """HandPrior(type='grab') (lib_smpl/th_hand_prior.py:47-78): left / right hand Mahalanobis terms on
    pose[:, 66:] (2 x 45)."""
    HAND_POSE_NUM = 45

    def __init__(self, mean, lhand_prec, rhand_prec, prefix: int = 66, device="cuda:0"):
        self.prefix = prefix
        self.mean = torch.as_tensor(mean, dtype=torch.float32).reshape(1, -1).to(device)
        self.lhand_prec = torch.as_tensor(lhand_prec, dtype=torch.float32).unsqueeze(0).to(device)
        self.rhand_prec = torch.as_tensor(rhand_prec, dtype=torch.float32).unsqueeze(0).to(device)

    def __call__(self, full_pose):
        """Shapes as in the reference (:69-78): the precisions are (1,45,45), so the products are (1,B,45),
        the concatenation along dim 1 is (1,2B,45) and the result is (1,45) -- summed over batch and hands,
        NOT one value per batch row; torch.mean() of it is total / 45.  Kept for parity."""
        temp = full_pose[:, self.prefix:] - self.mean
        lhand = torch.matmul(temp[:, :self.HAND_POSE_NUM], self.lhand_prec)
        rhand = torch.matmul(temp[:, self.HAND_POSE_NUM:], self.rhand_prec)
        temp2 = torch.cat([lhand, rhand], 1)
        return (temp2 * temp2).sum(dim=1)


def load_priors(assets_root: str, device="cuda:0") -> Tuple[MahalanobisPrior, HandPrior]:
    """get_prior() + HandPrior('grab') of the reference (lib_smpl/th_smpl_prior.py:19-52,
    lib_smpl/th_hand_prior.py:20-45), read ONCE from `assets_root`/priors/*.pkl -- the reference
    un-pickles all three files on every optimisation step (recon/recon_fit_base.py:530,534)."""
    import pickle as pkl
    from os.path import join
    import numpy as np
    rd = lambda n: pkl.load(open(join(assets_root, "priors", n), "rb"), encoding="latin1")
    body, lh, rh = rd("body_prior.pkl"), rd("lh_prior.pkl"), rd("rh_prior.pkl")
    body_prior = MahalanobisPrior(np.asarray(body["mean"]).astype("float32"), np.asarray(body["precision"]).astype("float32"), device=device)
    hand_prior = HandPrior(np.concatenate([lh["mean"], rh["mean"]], 0), lh["precision"], rh["precision"], device=device)
    return body_prior, hand_prior


class GraphedStep:
    """One optimisation step (zero_grad -> losses -> backward -> optimizer.step) captured in a CUDA graph
    and replayed: no Python, no autograd bookkeeping and no launch gaps on the hot loop.  This is the
    B200-side replacement for the per-step Python of optimize_smpl / optimize_smpl_object
    (recon/recon_fit_behave.py:90-163,224-291), which additionally syncs the device every step for its
    tqdm strings (.item(), :154-157).

    `step_fn()` must run the whole step on the current stream with static tensors (use
    `torch.optim.Adam(..., capturable=True)` and device-side RNG) and return the loss tensor(s); the
    returned tensors are refreshed in place by every `__call__`."""

    def __init__(self, step_fn: Callable, warmup: int = 3, snapshot: Optional[Callable] = None,
                 restore: Optional[Callable] = None):
        """snapshot() / restore(state): the warm-up runs `step_fn` for real (lazy allocations must happen outside the
        capture); when the step mutates optimiser state, pass these so the warm-up steps are undone and a fixed
        schedule of N replays applies exactly N updates."""
        self.step_fn = step_fn
        state = snapshot() if snapshot is not None else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):          # lazy allocations / attribute setup happen outside the capture
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if restore is not None:
            restore(state)
            torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        inner = None
        try:
            with torch.cuda.graph(self.graph):
                try:
                    self.out = step_fn()
                except Exception as e:      # capture_end() would otherwise mask the real error
                    inner = e
        except Exception as outer:
            raise (inner or outer)
        if inner is not None:
            raise inner

    def __call__(self):
        self.graph.replay()
        return self.out


def backward_to(loss: torch.Tensor, params) -> None:
    """`loss.backward()` for a captured step: gradients are taken with torch.autograd.grad and stored in
    `.grad`, so no AccumulateGrad node runs.  Those nodes are bound to the stream on which a parameter was
    first used (often the legacy default stream), and syncing with that stream is illegal while capturing."""
    params = list(params)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for p, g in zip(params, grads):
        p.grad = g


class ReconFitterBase:
    """The kernel-backed subset of recon/recon_fit_base.py:ReconFitterBase."""

    def __init__(self, device="cuda:0", obj_scale: float = 1.0, debug: bool = False, priors=None,
                 net_in_size: int = 512, crop_size: float = 1200.0, strict: bool = False, scan=None, part_labels=None):
        """priors: (MahalanobisPrior, HandPrior) from load_priors(), or None.  net_in_size / crop_size:
        args.net_img_size[0] / KinectColorCamera(args.loadSize).crop_size (recon_fit_base.py:74-75).
        strict: raise when forward_smpl lacks the priors / landmark regressors / data_dict entries its
        non-field terms need (the reference always has them); otherwise those terms are left out.
        scan: the object template mesh (recon_fit_base.py:107-125 loads it from the BEHAVE object folder); part_labels:
        load_part_labels(assets_root)."""
        self.device = device
        self.obj_scale = obj_scale
        self.debug = debug
        self.z_0 = 2.2
        self.priors = priors
        self.net_in_size = net_in_size
        self.crop_size = float(crop_size)
        if self.crop_size != 1200.0:
            # the query kernels project with KinectColorCamera(loadSize = 1200) (csrc/query_tc_shared.cuh); a different
            # crop here would make the keypoint term and the field query disagree silently
            raise ValueError(f"chore_b200 is built for loadSize / crop_size 1200, got {crop_size}")
        self.strict = strict
        self.scan = scan                    # centred object template (.v (V,3), .f (F,3)): silhouette phase, save_outputs
        self.part_labels = part_labels      # (6890,) SMPL part index per vertex: contact term (load_part_labels)
        self.collision_fn = None            # optional interpenetration term (see compute_collision_loss)
        # KinectColorCamera pixel intrinsics (model/camera.py:26-40)
        self.fx_px, self.fy_px = 979.7844 / 2048. * 2048, 979.840 / 2048. * 2048
        self.cx_px, self.cy_px = 1018.952 / 2048. * 2048, 779.486 / 2048. * 2048

    # ---- SO(3) / rigid helpers ----------------------------------------------------------------
    @staticmethod
    def project_so3(mat: torch.Tensor) -> torch.Tensor:
        """R = U diag(1,1,det(U V^T)) V^T (recon_fit_base.py:167-188), closed form in-kernel instead
        of torch.svd; differentiable."""
        return _So3Fn.apply(mat, _lib.get_handle(mat.device))

    @staticmethod
    def decopose_axis(rot: torch.Tensor, no_rand: bool = False, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """recon_fit_base.py:373-384.  `noise` (B,3,3) in [0,1) replaces the reference's
        torch.rand draw when given (reproducible runs)."""
        if no_rand:
            return ReconFitterBase.project_so3(rot)
        if noise is None:
            if torch.cuda.is_current_stream_capturing():
                noise = torch.rand(rot.shape[0], 3, 3, device=rot.device)    # graph-safe device RNG
            else:
                noise = torch.rand(rot.shape[0], 3, 3)            # CPU generator, like the reference
        return ReconFitterBase.project_so3(rot + 1e-4 * noise.to(rot.device))

    def transform_obj_verts(self, verts, obj_R, obj_t, obj_s):
        """(verts @ R + t) * s, scale after rotation and translation (recon_fit_base.py:367-371)."""
        return _RigidFn.apply(verts, obj_R, obj_t, obj_s, _lib.get_handle(verts.device))

    def transform_object(self, object_init, rot, obj_t, obj_s):
        return self.transform_obj_verts(object_init, self.decopose_axis(rot), obj_t, obj_s)

    # ---- losses -------------------------------------------------------------------------------
    @staticmethod
    def sum_dict(loss_dict: Dict[str, torch.Tensor], weight_dict: Dict[str, Callable], it) -> torch.Tensor:
        return torch.stack([weight_dict[k](v, it) for k, v in loss_dict.items()]).sum()

    def compute_obj_loss(self, data_dict, loss_dict, model, obj_s, object):
        with heads(model, _lib.HEAD_DF):                    # only preds[0] is read below and by the callers
            model.query(object, **data_dict["query_dict"])
        preds = model.get_preds()
        loss_dict["object"] = torch.clamp(preds[0][:, 1:2, :], max=0.8).mean()
        loss_dict["scale"] = torch.mean((obj_s - self.obj_scale) ** 2)
        return preds

    def compute_df_h_loss(self, data_dict, loss_dict, model, smpl_verts):
        # centers_pred is only read by the reference's debug visualisation (recon_fit_behave.py:321-331)
        with heads(model, _lib.HEAD_DF | _lib.HEAD_PARTS | (_lib.HEAD_CENTERS if self.debug else 0)):
            model.query(smpl_verts, **data_dict["query_dict"])
        df_pred, _, parts_pred, centers_pred = model.get_preds()
        loss_dict["df_h"] = torch.clamp(df_pred[:, 0:1, :], max=0.1).mean()
        return df_pred, parts_pred, centers_pred

    def compute_contact_loss(self, df_hum_o, df_obj_h, object, smpl_verts, loss_dict, part_o=None):
        """recon_fit_base.py:553-608: pull the contact points (cross distance field < 0.08 m) of matching SMPL parts together
        (part-wise Chamfer distance, csrc/contact.cu).  Needs `self.part_labels` ((6890,) part index per SMPL vertex,
        load_part_labels :277-287).  Adds nothing when no contact is found, like the reference (one host sync for that test)."""
        if getattr(self, "part_labels", None) is None:
            raise RuntimeError("compute_contact_loss needs fitter.part_labels (6890 part indices, assets/smpl_parts_dense.pkl)")
        loss, pairs = _ContactFn.apply(object, smpl_verts, df_hum_o, df_obj_h, part_o, self.part_labels.to(object.device),
                                       _lib.get_handle(object.device))
        if int(pairs) == 0:
            if self.debug:
                print("no contact")
            return
        loss_dict["contact"] = loss

    def compute_collision_loss(self, smpl_verts, smpl_faces, obj_R, obj_t, obj_s):
        """recon_fit_base.py:610-642 is torch-mesh-isect's BVH + DistanceFieldPenetrationLoss (third-party CUDA, source not under
        the reference tree, no pinned version): not restated here.  Plug it in with `fitter.collision_fn = callable(smpl_verts,
        smpl_faces, obj_R, obj_t, obj_s) -> scalar`; without one the 'collide' term is left out (its weight is 3^2 / (1 + decay)
        against 30^2 for contact)."""
        fn = getattr(self, "collision_fn", None)
        return None if fn is None else fn(smpl_verts, smpl_faces, obj_R, obj_t, obj_s)

    @staticmethod
    def load_part_labels(assets_root: str, device="cpu") -> torch.Tensor:
        """(6890,) part index per SMPL vertex from smpl_parts_dense.pkl (recon_fit_base.py:277-287; dict order = label)."""
        import pickle as pkl
        from os.path import join
        d = pkl.load(open(join(assets_root, "smpl_parts_dense.pkl"), "rb"), encoding="latin1")
        labels = torch.zeros(6890, dtype=torch.int32)
        for n, k in enumerate(d):
            labels[torch.as_tensor(d[k]).long()] = n
        return labels.to(device)

    def compute_prior_loss(self, loss_dict, smpl, nobeta: bool = False):
        """recon_fit_base.py:522-535 with the priors loaded once (self.priors)."""
        if self.priors is None:
            raise RuntimeError("no SMPL priors: construct the fitter with priors=load_priors(assets_root)")
        prior, hand_prior = self.priors
        if not nobeta:
            loss_dict["beta"] = torch.mean(smpl.betas ** 2)
        loss_dict["pose"] = torch.mean(prior(smpl.pose[:, :72]))
        loss_dict["hand"] = torch.mean(hand_prior(smpl.pose))

    def smplz_loss(self, J, loss_dict):
        """recon_fit_base.py:230-231: the depth of body25 joint 8 (mid hip) stays at z_0."""
        loss_dict["smplz"] = torch.mean((J[:, 8, 2] - self.z_0) ** 2)

    def project_points(self, joints3d, crop_center=None):
        """recon_fit_base.py:661-670 + KinectColorCamera.project_screen (model/camera.py:51-71)."""
        x, y, z = joints3d[:, :, 0:1], joints3d[:, :, 1:2], joints3d[:, :, 2:3]
        px = self.fx_px * x / z + self.cx_px
        py = self.fy_px * y / z + self.cy_px
        if crop_center is not None:
            px = self.crop_size / 2 + px - crop_center[:, 0].unsqueeze(1).unsqueeze(1)
            py = self.crop_size / 2 + py - crop_center[:, 1].unsqueeze(1).unsqueeze(1)
        return torch.cat([px, py], -1) * self.net_in_size / self.crop_size

    def projection_loss(self, joints3d, joints2d, crop_center):
        """recon_fit_base.py:672-676: confidence-weighted squared pixel distance."""
        joints_proj = self.project_points(joints3d, crop_center)
        loss = F.mse_loss(joints_proj[:, :, :2], joints2d[:, :, :2], reduction="none")
        return torch.mean(torch.sum(loss, dim=-1) * joints2d[:, :, 2])

    def compute_kpts_loss(self, data_dict, loss_dict, smpl, J=None):
        """recon_fit_base.py:653-659.  `J` (body25 joints of this step) may be passed in to avoid the
        reference's third LBS + regressor pass of the step."""
        if J is None:
            J = smpl.get_landmarks()[0]
        loss_dict["j2d"] = self.projection_loss(J, data_dict["body_kpts"], data_dict["query_dict"]["crop_center"])

    @staticmethod
    def split_smpl(smpl):
        from .smpl import SMPLPyTorchWrapperBatchSplitParams
        return SMPLPyTorchWrapperBatchSplitParams.from_smpl(smpl)

    @staticmethod
    def copy_smpl_params(split_smpl, smpl):
        """recon_fit_base.py:682-690 (other_betas are not copied there either)."""
        smpl.pose.data[:, :3] = split_smpl.global_pose.data
        smpl.pose.data[:, 3:66] = split_smpl.body_pose.data
        smpl.pose.data[:, 66:] = split_smpl.hand_pose.data
        smpl.betas.data[:, :2] = split_smpl.top_betas.data
        smpl.trans.data = split_smpl.trans.data
        return smpl

    @staticmethod
    def get_smpl_bbox(smpl):
        with torch.no_grad():
            verts = smpl()[0]
        return torch.min(verts, 1)[0], torch.max(verts, 1)[0]

    def get_smpl_height(self, smpl):
        bmin, bmax = self.get_smpl_bbox(smpl)
        return bmax[:, 1] - bmin[:, 1]

    @staticmethod
    def get_loss_str(it, loss_dict, weight_dict, weight_decay) -> str:
        """recon_fit_base.py:345-349; one device->host copy for the whole line instead of one .item() per term."""
        keys = list(loss_dict)
        vals = torch.stack([weight_dict[k](loss_dict[k], weight_decay).mean().detach() for k in keys]).tolist()
        return "Iter: {}".format(it) + "".join(", {}: {:0.4f}".format(k, v) for k, v in zip(keys, vals))

    def compute_smpl_center_pred(self, data_dict, model, smpl):
        with torch.no_grad():
            smpl_verts = smpl()[0]
            with heads(model, _lib.HEAD_CENTERS):
                model.query(smpl_verts, **data_dict["query_dict"])
            return torch.mean(model.get_preds()[3][:, :3], -1)


class ReconFitterBehave(ReconFitterBase):
    """The kernel-backed subset of recon/recon_fit_behave.py:ReconFitterBehave."""

    def get_loss_weights(self):
        w = {"beta": 1.0, "pose": 1e-5, "hand": 1e-5, "j2d": 0.3 ** 2, "object": 30.0 ** 2, "part": 0.05 ** 2,
             "contact": 30.0 ** 2, "scale": 10.0 ** 2, "df_h": 30.0 ** 2, "smplz": 30 ** 2, "mask": 0.003 ** 2,
             "ocent": 15 ** 2, "collide": 3 ** 2, "pinit": 5 ** 2, "rot": 10.0 ** 2, "trans": 10.0 ** 2}
        return {k: (lambda cst, it, c=c: c * cst / (1 + it)) for k, c in w.items()}

    def forward_step(self, model, smpl, data_dict, obj_R, obj_t, obj_s, phase, noise=None):
        """recon_fit_behave.py:165-222, all three phases.  'object only': object / scale / ocent (the reference's double query of
        the object points, :179 + recon_fit_base.py:515, is kept so results and cost are comparable).  'sil': the occlusion-aware
        silhouette term of data_dict['silhouette'] (SilLossROI) + scale / trans regularisers.  'joint': 'object only' + the
        contact term + (when a collision_fn is plugged in) the interpenetration term."""
        if phase not in ("object only", "sil", "joint"):
            raise ValueError(f"unknown phase {phase!r}")
        loss_dict = {}
        R = self.decopose_axis(obj_R, noise=noise)
        object = self.transform_obj_verts(data_dict["objects"], R, obj_t, obj_s)
        need = _lib.HEAD_CENTERS | ((_lib.HEAD_DF | _lib.HEAD_PARTS) if phase == "joint" else 0)
        with heads(model, need):
            model.query(object, **data_dict["query_dict"])
        preds = model.get_preds()
        df_pred, part_o, centers_pred_o = preds[0], preds[2], preds[3]
        obj_center_pred = data_dict["smpl_center"] + torch.mean(centers_pred_o[:, 3:, :], -1)
        if phase == "sil":
            obj_losses, image, edges, image_ref, edt_ref = data_dict["silhouette"](R, obj_t, obj_s)
            loss_dict["mask"] = obj_losses["mask"]
            data_dict["image_ref"], data_dict["edt_ref"] = image_ref, edt_ref
            loss_dict["scale"] = torch.mean((obj_s - self.obj_scale) ** 2)
            loss_dict["trans"] = torch.mean((obj_t - data_dict["trans_init"]) ** 2)
            return loss_dict
        self.compute_obj_loss(data_dict, loss_dict, model, obj_s, object)
        obj_center_act = torch.mean(object, 1)
        loss_dict["ocent"] = F.mse_loss(obj_center_act, obj_center_pred, reduction="none").sum(-1).mean()
        if phase == "joint":
            smpl_verts = smpl()[0]
            df_obj_h = df_pred[:, 0, :]
            with heads(model, _lib.HEAD_DF):
                model.query(smpl_verts, **data_dict["query_dict"])
            df_hum_o = model.get_preds()[0][:, 1, :]
            self.compute_contact_loss(df_hum_o, df_obj_h, object, smpl_verts, loss_dict, part_o=part_o)
            pen = self.compute_collision_loss(smpl_verts, smpl.faces, R, obj_t, obj_s)
            if pen is not None:
                loss_dict["collide"] = pen
        return loss_dict

    def forward_smpl(self, smpl, data_dict, phase="global"):
        """recon_fit_behave.py:293-337, every term: df_h, priors (pose, hand), part cross-entropy, fixed depth of
        the mid hip (smplz), initial-pose regulariser (pinit) and, in phase 'kpts', the 2-D keypoint term (j2d).
        The LBS runs ONCE per step: the landmark regressors are applied to the vertices of that same forward
        (the reference runs a second and, in 'kpts', a third LBS inside get_landmarks, :305,315 and
        recon_fit_base.py:649 -- same values).  Terms whose assets are absent are skipped unless self.strict."""
        loss_dict = {}
        model = data_dict["net"]
        smpl_verts, jtr, _, _ = smpl()
        _, parts_pred, _ = self.compute_df_h_loss(data_dict, loss_dict, model, smpl_verts)
        if self.priors is not None:
            self.compute_prior_loss(loss_dict, smpl, nobeta=True)
        elif self.strict:
            raise RuntimeError("forward_smpl: no SMPL priors (strict)")
        loss_dict["part"] = F.cross_entropy(parts_pred, data_dict["part_labels"], reduction="none").sum(-1).mean()
        J = None
        if getattr(smpl, "regressors", None) is not None:
            J, _, _ = smpl.get_landmarks(smpl_verts)
            self.smplz_loss(J, loss_dict)
        elif self.strict:
            raise RuntimeError("forward_smpl: the SMPL wrapper has no landmark regressors (strict)")
        if "pose_init" in data_dict:
            loss_dict["pinit"] = torch.mean(torch.sum((smpl.pose[:, 3:72] - data_dict["pose_init"]) ** 2, -1))
        elif self.strict:
            raise KeyError("pose_init")
        if phase == "kpts":
            if J is None:
                raise RuntimeError("phase 'kpts' needs landmark regressors on the SMPL wrapper")
            self.compute_kpts_loss(data_dict, loss_dict, smpl, J)
        return loss_dict

    # ---- optimisation loops ---------------------------------------------------------------------
    def optimize_smpl(self, smpl, data_dict, iter_for_betas=10, iter_for_pose=10, iter_for_kpts=5, steps_per_iter=10,
                      max_iter=150, log: Optional[Callable[[str], None]] = None):
        """recon_fit_behave.py:224-291: phase 'global' (top betas + translation, lr 0.02), then all poses
        (lr 0.006), then 'kpts' until convergence.  Same schedule, decay and early-stop rule; the convergence
        test is the only device->host sync of a step (the reference also syncs for its tqdm string)."""
        smpl_split = self.split_smpl(smpl)
        opt = torch.optim.Adam([smpl_split.top_betas, smpl_split.trans], lr=0.02)
        height_init = self.get_smpl_height(smpl)
        weight_dict = self.get_loss_weights()
        prev_loss = 300.0
        phase = "global"
        for it in range(iter_for_betas + iter_for_kpts + iter_for_pose + max_iter):
            opt.zero_grad()        # once per outer iteration, like the reference (:244): grads accumulate inside
            if it == iter_for_betas:
                phase = "smpl all pose"
                opt = torch.optim.Adam([smpl_split.trans, smpl_split.global_pose, smpl_split.body_pose,
                                        smpl_split.top_betas, smpl_split.other_betas], 0.006, betas=(0.9, 0.999))
            elif it == iter_for_betas + iter_for_pose:
                phase = "kpts"
            for i in range(steps_per_iter):
                loss_dict = self.forward_smpl(smpl_split, data_dict, phase)
                decay = 1 if phase != "kpts" else it / 3
                loss = self.sum_dict(loss_dict, weight_dict, decay)
                loss.backward()
                opt.step()
                if log is not None:
                    log(f"{phase}: " + self.get_loss_str(f"{it}-{i}", loss_dict, weight_dict, decay))
                # early stop (:275-282); evaluated (one host sync) only inside the window where it can fire
                if it > 0.25 * max_iter + iter_for_betas + iter_for_pose:
                    lv, pv = float(loss.detach()), float(prev_loss)
                    if abs(pv - lv) / pv < pv * 0.001:
                        scale = self.get_smpl_height(smpl_split) / height_init
                        return self.copy_smpl_params(smpl_split, smpl), scale
                prev_loss = loss.detach()
        scale = self.get_smpl_height(smpl_split) / height_init
        return self.copy_smpl_params(smpl_split, smpl), scale

    def optimize_smpl_object(self, model, data_dict, obj_iter=20, joint_iter=10, steps_per_iter=10,
                             log: Optional[Callable[[str], None]] = None, sil_iter: int = 50, max_iter: int = 100):
        """recon_fit_behave.py:90-163, the reference's schedule: `obj_iter` outer iterations 'object only' (Adam on obj_t, obj_R,
        obj_s, lr 0.006, decay 1), then `sil_iter` (50) 'sil' iterations (a fresh Adam on obj_R, obj_s, obj_t; decay it - obj_iter + 1),
        then 'joint' (a fresh Adam on obj_t, obj_s, lr 0.002; decay (it - obj_iter + 1) / 5) for joint_iter + max_iter iterations
        with the reference's early-stop rule.  zero_grad() once per outer iteration (gradients accumulate over the inner steps).
        The 'sil' phase runs when data_dict['silhouette'] holds a silhouette loss (chore_b200.silhouette.SilLossROI, built from
        data_dict['images'] when a `scan` mesh was given to the fitter) and is skipped otherwise."""
        smpl = data_dict["smpl"]
        smpl_split = self.split_smpl(smpl)
        data_dict["smpl"] = smpl_split
        obj_R, obj_t, obj_s = data_dict["obj_R"], data_dict["obj_t"], data_dict["obj_s"]
        if "silhouette" not in data_dict and getattr(self, "scan", None) is not None and "images" in data_dict:
            from .silhouette import SilLossROI
            images = data_dict["images"]
            data_dict["silhouette"] = SilLossROI(images[:, 3, :, :], images[:, 4, :, :], self.scan, data_dict["query_dict"]["crop_center"],
                                                 device=obj_t.device)
        has_sil = "silhouette" in data_dict
        iter_for_sil = sil_iter if has_sil else 0
        opt = torch.optim.Adam([obj_t, obj_R, obj_s], lr=0.006)
        weight_dict = self.get_loss_weights()
        data_dict["smpl_center"] = self.compute_smpl_center_pred(data_dict, model, smpl)
        prev_loss = 300.0
        phase = "object only"
        for it in range(joint_iter + obj_iter + max_iter + iter_for_sil):
            opt.zero_grad()
            if it == obj_iter and has_sil:
                phase = "sil"
                opt = torch.optim.Adam([obj_R, obj_s, obj_t], lr=0.006)
                data_dict["rot_init"] = self.decopose_axis(obj_R).detach().clone()
                data_dict["trans_init"] = obj_t.detach().clone()
            if it == obj_iter + iter_for_sil:
                phase = "joint"
                opt = torch.optim.Adam([obj_t, obj_s], lr=0.002)
            for i in range(steps_per_iter):
                loss_dict = self.forward_step(model, smpl_split, data_dict, obj_R, obj_t, obj_s, phase)
                decay = 1 if phase == "object only" else (it - obj_iter + 1 if phase == "sil" else (it - obj_iter + 1) / 5)
                loss = self.sum_dict(loss_dict, weight_dict, decay)
                loss.backward()
                opt.step()
                if log is not None:
                    log(f"{phase} " + self.get_loss_str(f"{it}-{i}", loss_dict, weight_dict, decay))
                if phase == "joint" and it > 0.25 * max_iter:       # early stop (:158-160): one host sync, only where it can fire
                    lv, pv = float(loss.detach()), float(prev_loss)
                    if abs(pv - lv) / pv < pv * 0.0001:
                        return smpl, data_dict["obj_R"], data_dict["obj_t"]
                prev_loss = loss.detach()
        return smpl, data_dict["obj_R"], data_dict["obj_t"]


class FusedAdam:
    """torch.optim.Adam(params, lr, betas, eps) (no weight decay, no amsgrad) as ONE kernel launch per step for up to
    8 small tensors (chore_adam_step); the step counter lives on the device, so a captured step replays correctly.
    `grads[i]` is the tensor the gradient of params[i] is read from at step() time: a (rows, cols) view that may be a
    column slice of a wider, persistent buffer (its storage must not move between steps -- true inside a CUDA graph
    and for buffers the caller keeps).

    accumulate=True keeps one accumulator per parameter and gives the update the SUM of the gradients passed to step()
    since the last zero_grad() -- what `.grad` holds in the reference loops, which call optimizer.zero_grad() once per
    outer iteration and loss.backward() in each of the 10 inner steps (recon/recon_fit_behave.py:135-152,244-273).
    `gscale` (device scalar, optional) multiplies every gradient: the 1 / (1 + decay) of get_loss_weights."""

    def __init__(self, params, lr=0.006, betas=(0.9, 0.999), eps=1e-8, handle=None, accumulate: bool = False):
        self.params = [p for p in params]
        assert 0 < len(self.params) <= 8
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        dev = self.params[0].device
        self.handle = handle or _lib.get_handle(dev)
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        # `.grad` of every parameter IS its accumulator (as in torch, where backward() adds into .grad)
        self.accumulate = accumulate
        self.grad_acc = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        for p, a in zip(self.params, self.grad_acc):
            p.grad = a
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)

    def zero_grad(self) -> None:
        torch._foreach_zero_(self.grad_acc)

    def state(self):
        return [t.detach().clone() for t in (*self.params, *self.exp_avg, *self.exp_avg_sq, *self.grad_acc, self.step_count)]

    def load_state(self, state) -> None:
        with torch.no_grad():
            for t, sv in zip((*self.params, *self.exp_avg, *self.exp_avg_sq, *self.grad_acc, self.step_count), state):
                t.copy_(sv)

    def step(self, grads, gscale: Optional[torch.Tensor] = None, loss: Optional[torch.Tensor] = None) -> None:
        if not self.accumulate:
            self.zero_grad()
        ent = (_lib.AdamEntry * len(self.params))()
        for i, (e, p, g, m, v) in enumerate(zip(ent, self.params, grads, self.exp_avg, self.exp_avg_sq)):
            p2 = p.detach().view(p.shape[0], -1) if p.dim() > 1 else p.detach().view(1, -1)
            g2 = g.view(g.shape[0], -1) if g.dim() > 1 else g.view(1, -1)
            assert p2.is_contiguous() and g2.stride(-1) == 1 and g2.shape == p2.shape and g2.dtype == torch.float32
            e.param, e.grad, e.exp_avg, e.exp_avg_sq = p2.data_ptr(), g2.data_ptr(), m.data_ptr(), v.data_ptr()
            e.grad_acc = self.grad_acc[i].data_ptr()
            e.rows, e.cols, e.grad_ld = p2.shape[0], p2.shape[1], (g2.stride(0) if g2.shape[0] > 1 else p2.shape[1])
        self.handle.adam_step(ent, len(self.params), self.lr, self.betas[0], self.betas[1], self.eps, self.step_count, gscale, loss)


class FusedFitSteps:
    """The two inner optimisation steps of the fitting loop with NO autograd graph and no eager elementwise ops:
      SMPL phase (recon/recon_fit_behave.py:293-337):   LBS -> landmarks -> field query (df, parts) -> loss kernels
          (df_h + part CE, smplz + j2d, priors + pinit) -> query adjoint -> landmark adjoint -> LBS adjoint -> Adam
      'object only' phase (recon/recon_fit_behave.py:165-198):   SO(3) -> rigid -> field query (df, centers) -> loss
          kernels (object, scale, ocent) -> query adjoint -> rigid adjoint -> SO(3) adjoint -> Adam
    Every stage is one or two launches of libchore_b200.so (csrc/fit_loss.cu holds the loss / Adam kernels), all
    stream-ordered, hence capturable in a CUDA graph (`graphed()`).

    The reference queries the object points twice per step (recon_fit_behave.py:179 + recon_fit_base.py:515) with
    identical inputs, so one forward + one adjoint launch (heads df and centers) gives the same losses and gradients.
    Loss weights / decay are those of get_loss_weights(); Adam is torch.optim.Adam's update rule (FusedAdam)."""

    W = {"object": 30.0 ** 2, "part": 0.05 ** 2, "scale": 10.0 ** 2, "df_h": 30.0 ** 2, "ocent": 15 ** 2, "pinit": 5 ** 2,
         "pose": 1e-5, "hand": 1e-5, "smplz": 30 ** 2, "j2d": 0.3 ** 2}

    def __init__(self, net, smpl, data_dict, obj_R, obj_t, obj_s, lr_smpl=0.006, lr_obj=0.006, obj_scale=1.0, decay=1.0,
                 fitter: Optional["ReconFitterBase"] = None, phase: str = "smpl all pose", accumulate: bool = True):
        """fitter: supplies the priors / camera constants for the pose-prior, hand-prior, smplz and (phase 'kpts')
        2-D keypoint terms of forward_smpl; without it only the field terms (+ pinit) are used.  The landmark terms
        need regressors on `smpl`.

        Reference semantics of the loop around the step (recon/recon_fit_behave.py:90-163,224-291):
          * accumulate=True: gradients add up between zero_grad() calls (the reference zeroes once per OUTER iteration,
            so inner step i sees the sum of the gradients of inner steps 0..i); call zero_grad() where it does.
          * set_decay(decay): every loss weight is divided by (1 + decay) (1 in the early phases, it / 3 in 'kpts');
            the factor lives in a device scalar, so captured graphs follow the schedule.
          * set_phase('global' | 'smpl all pose' | 'kpts'): 'global' optimises top_betas + trans with lr 0.02, the
            switch to 'smpl all pose' creates a fresh Adam (lr 0.006) like the reference, 'kpts' keeps it and adds j2d."""
        self.net, self.smpl, self.data = net, smpl, data_dict
        self.fitter = fitter
        self.R, self.t, self.s = obj_R, obj_t, obj_s
        self.obj_scale = obj_scale
        self.accumulate = accumulate
        self.h_net = net.handle
        self.h_lbs = smpl.smpl.handle
        self.h_aux = _lib.get_handle(obj_t.device)
        dev = obj_t.device
        self.kdev = torch.ones(1, device=dev)             # 1 / (1 + decay), read by the Adam kernel
        self.lr_smpl = lr_smpl
        self.phase = None
        self.opt_smpl = None
        self.set_phase(phase)
        self.set_decay(decay)
        self.opt_obj = FusedAdam([obj_t, obj_R, obj_s], lr_obj, handle=self.h_aux, accumulate=accumulate)
        self.cc = data_dict["query_dict"]["crop_center"].detach().float().contiguous()
        nmax = max(smpl.offsets.shape[1], data_dict["objects"].shape[1]) if "objects" in data_dict else smpl.offsets.shape[1]
        self.ws = self.h_aux.fit_workspace(obj_t.shape[0], nmax, dev)
        self.ws_obj = self.h_aux.fit_workspace(obj_t.shape[0], nmax, dev)      # own scratch: the two steps may run concurrently
        self.loss_smpl = torch.zeros(1, device=dev)
        self.loss_obj = torch.zeros(1, device=dev)
        self.priors_dev = None
        if fitter is not None and fitter.priors is not None:
            bp, hp = fitter.priors
            assert bp.prefix == 3 and bp.end == 66 and hp.prefix == 66, "priors are expected on pose[3:66] and pose[66:]"
            self.priors_dev = tuple(t.to(dev).float().contiguous() for t in
                                    (bp.mean.reshape(-1), bp.prec, hp.mean.reshape(-1), hp.lhand_prec[0], hp.rhand_prec[0]))
        self.pose_init = data_dict["pose_init"].detach().float().contiguous() if "pose_init" in data_dict else None

    # ---- schedule hooks (the Python around the step in optimize_smpl / optimize_smpl_object) -------------------
    def set_phase(self, phase: str) -> None:
        sm = self.smpl
        assert phase in ("global", "smpl all pose", "kpts"), phase
        if phase == "global":
            names, lr = ("top_betas", "trans"), 0.02                     # recon_fit_behave.py:233
        else:
            names, lr = ("trans", "global_pose", "body_pose", "top_betas", "other_betas"), self.lr_smpl   # :250-253
        fresh = self.opt_smpl is None or names != self.smpl_names        # 'kpts' keeps the optimiser of 'smpl all pose'
        self.phase, self.smpl_names = phase, names
        if fresh:
            self.smpl_params = [getattr(sm, n) for n in names]
            self.opt_smpl = FusedAdam(self.smpl_params, lr, handle=self.h_aux, accumulate=self.accumulate)

    def set_decay(self, decay: float) -> None:
        self.decay = float(decay)
        self.kdev.fill_(1.0 / (1.0 + self.decay))

    def zero_grad(self) -> None:
        """optimizer.zero_grad() of the reference loops: once per outer iteration."""
        self.opt_smpl.zero_grad()
        self.opt_obj.zero_grad()

    @torch.no_grad()
    def smpl_step(self):
        sm, d = self.smpl, self.data
        feat, skip = self.net._maps()
        pose = torch.cat([sm.global_pose, sm.body_pose, sm.hand_pose], 1)
        betas = torch.cat([sm.top_betas, sm.other_betas], 1)
        trans, off = sm.trans.detach(), sm.offsets.detach()
        loss = self.loss_smpl.zero_()
        verts, _, _, _ = self.h_lbs.lbs_fwd(pose, betas, trans, off, want_posed=False)
        (df, _, parts, _), _ = self.h_net.query_fwd(feat, skip, verts, self.cc, _lib.HEAD_DF | _lib.HEAD_PARTS)
        B, _, N = df.shape
        k = 1.0          # the 1 / (1 + decay) factor is applied to the summed gradient and the loss by the Adam kernel (kdev)
        # L = w_dfh * mean(min(df_h, 0.1)) + w_part * mean_b sum_n CE(parts, labels)
        g_df, g_parts = self.h_aux.fit_smpl_field_grads(df, parts, d["part_labels"], self.W["df_h"] * k / (B * N), self.W["part"] * k / B,
                                                        loss, self.ws)
        g_verts = self.h_net.query_bwd(feat, skip, verts, self.cc, [g_df, None, g_parts, None])
        fit = self.fitter
        if getattr(sm, "regressors", None) is not None and fit is not None:
            # smplz: 30^2 mean_b (J8.z - z0)^2; j2d ('kpts'): 0.3^2 mean_{b,j} conf ||proj(J) - kpts||^2 (recon_fit_base.py:230,653-676)
            hl = sm.regressors.handle
            lm = hl.landmarks_fwd(verts)
            nJ = sm.regressors.sizes[0]
            kp = d["body_kpts"] if self.phase == "kpts" else None
            cam = (fit.fx_px, fit.fy_px, fit.cx_px, fit.cy_px, fit.crop_size / 2, fit.net_in_size / fit.crop_size)
            g_lm = self.h_aux.fit_landmark_grads(lm, kp, self.cc, nJ, fit.z_0, self.W["smplz"] * k / B, self.W["j2d"] * k / (B * nJ), cam, loss,
                                                 self.ws)
            hl.landmarks_bwd(g_lm, g_verts)          # g_verts += R^T g_lm
        g_pose, g_betas, g_trans, _ = self.h_lbs.lbs_bwd(pose, betas, trans, off, g_verts, None, False)
        if self.priors_dev is not None or self.pose_init is not None:
            # Mahalanobis priors (recon_fit_base.py:522-535; the hand term is total / 45, see HandPrior.__call__) and
            # 5^2 * mean_b sum (pose[3:72] - pose_init)^2 (recon_fit_behave.py:317-319)
            self.h_aux.fit_pose_prior_grads(pose, self.pose_init, self.priors_dev, self.W["pose"] * k / B, self.W["hand"] * k / 45.0,
                                            self.W["pinit"] * k / B, g_pose, loss, self.ws)
        by_name = {"trans": g_trans, "global_pose": g_pose[:, :3], "body_pose": g_pose[:, 3:66], "hand_pose": g_pose[:, 66:],
                   "top_betas": g_betas[:, :2], "other_betas": g_betas[:, 2:]}
        self._g_smpl = tuple(by_name[n] for n in self.smpl_names)
        self._g_pose_full = g_pose                                    # (B,156) incl. the hand-pose gradient (not optimised here)
        self.opt_smpl.step(self._g_smpl, self.kdev, loss)
        return loss

    @torch.no_grad()
    def object_step(self, noise: Optional[torch.Tensor] = None):
        d = self.data
        feat, skip = self.net._maps()
        obj0 = d["objects"]
        B, N, _ = obj0.shape
        if noise is None:
            noise = torch.rand(B, 3, 3, device=obj0.device)
        loss = self.loss_obj.zero_()
        rot_in = torch.add(self.R, noise, alpha=1e-4)                        # decopose_axis (recon_fit_base.py:373-384)
        Rm = self.h_aux.project_so3(rot_in)
        t, s = self.t.detach(), self.s.detach()
        obj = self.h_aux.rigid_fwd(obj0, Rm, t, s)
        (df, _, _, cen), _ = self.h_net.query_fwd(feat, skip, obj, self.cc, _lib.HEAD_DF | _lib.HEAD_CENTERS)
        k = 1.0          # see smpl_step
        wc = self.W["ocent"] * k / B
        g_df, g_cen, dvec = self.h_aux.fit_obj_field_grads(obj, df, cen, d["smpl_center"], s, self.obj_scale, self.W["object"] * k / (B * N), wc,
                                                           self.W["scale"] * k / B, loss, self.ws_obj)
        g_obj = self.h_net.query_bwd(feat, skip, obj, self.cc, [g_df, None, None, g_cen])
        self.h_aux.add_rowvec(g_obj, dvec, 2.0 * wc / N)
        g_R, g_t, g_s, _ = self.h_aux.rigid_bwd(obj0, Rm, t, s, g_obj, False)
        g_s.add_(s - self.obj_scale, alpha=self.W["scale"] * k * 2.0 / B)
        g_rot = self.h_aux.project_so3_bwd(rot_in, g_R)
        self._g_obj = (g_t, g_rot, g_s)
        self.opt_obj.step(self._g_obj, self.kdev, loss)
        return loss

    def iteration(self, noise: Optional[torch.Tensor] = None):
        """One fit iteration = SMPL step + 'object only' step.  The two touch disjoint parameters and scratch (the object step
        reads the SMPL centre computed once before the loop, recon_fit_behave.py:112), so the object step is forked onto a side
        stream and joined at the end: their many single-block kernels (kinematic chain, SO(3), Adam, reductions) overlap with the
        other step's field queries.  Capturable: `graphed_iteration()`."""
        cur = torch.cuda.current_stream()
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.t.device)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            lo = self.object_step(noise)
        ls = self.smpl_step()
        cur.wait_stream(self._side)
        return ls, lo

    def graphed_iteration(self):
        """`iteration()` captured in ONE CUDA graph (fork / join inside); warm-up undone like `graphed()`."""
        snap = lambda: (self.opt_smpl.state(), self.opt_obj.state(), self.loss_smpl.clone(), self.loss_obj.clone())
        rest = lambda st: (self.opt_smpl.load_state(st[0]), self.opt_obj.load_state(st[1]), self.loss_smpl.copy_(st[2]), self.loss_obj.copy_(st[3]))
        return GraphedStep(self.iteration, snapshot=snap, restore=rest)

    def graphed(self):
        """(smpl_step, object_step) of the CURRENT phase captured in CUDA graphs; each call replays one optimisation
        step.  The warm-up steps the capture needs are undone (parameters, Adam moments, accumulators, step counters), so
        N replays apply exactly N updates.  zero_grad() / set_decay() act on device buffers the graphs read; after
        set_phase('global' <-> others) capture again (the parameter set changes)."""
        snap_s = lambda: (self.opt_smpl.state(), self.loss_smpl.clone())
        rest_s = lambda st: (self.opt_smpl.load_state(st[0]), self.loss_smpl.copy_(st[1]))
        snap_o = lambda: (self.opt_obj.state(), self.loss_obj.clone())
        rest_o = lambda st: (self.opt_obj.load_state(st[0]), self.loss_obj.copy_(st[1]))
        return (GraphedStep(self.smpl_step, snapshot=snap_s, restore=rest_s),
                GraphedStep(self.object_step, snapshot=snap_o, restore=rest_o))
