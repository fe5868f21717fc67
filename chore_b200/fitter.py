"""Host-side mirror of the fitting step of the reference on the CUDA kernels.

Mirrors, with the same names and argument meaning (paths relative to /root/reference):
  ReconFitterBase.project_so3 / decopose_axis / transform_obj_verts / transform_object
      recon/recon_fit_base.py:167-188,361-384
  ReconFitterBase.sum_dict, compute_obj_loss, compute_df_h_loss, compute_smpl_center_pred
      recon/recon_fit_base.py:351-359,513-551
  ReconFitterBehave.get_loss_weights, forward_step ('object only'), forward_smpl (field terms)
      recon/recon_fit_behave.py:165-222,293-358
Out of scope: the silhouette phase (neural_renderer + detectron2), the
joint-phase contact / collision terms (pytorch3d, torch-mesh-isect), keypoint and prior terms.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

from . import _lib


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


class ReconFitterBase:
    """The kernel-backed subset of recon/recon_fit_base.py:ReconFitterBase."""

    def __init__(self, device="cuda:0", obj_scale: float = 1.0, debug: bool = False):
        self.device = device
        self.obj_scale = obj_scale
        self.debug = debug
        self.z_0 = 2.2

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
        model.query(object, **data_dict["query_dict"])
        preds = model.get_preds()
        loss_dict["object"] = torch.clamp(preds[0][:, 1:2, :], max=0.8).mean()
        loss_dict["scale"] = torch.mean((obj_s - self.obj_scale) ** 2)
        return preds

    def compute_df_h_loss(self, data_dict, loss_dict, model, smpl_verts):
        model.query(smpl_verts, **data_dict["query_dict"])
        df_pred, _, parts_pred, centers_pred = model.get_preds()
        loss_dict["df_h"] = torch.clamp(df_pred[:, 0:1, :], max=0.1).mean()
        return df_pred, parts_pred, centers_pred

    def compute_smpl_center_pred(self, data_dict, model, smpl):
        with torch.no_grad():
            smpl_verts = smpl()[0]
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
        """recon_fit_behave.py:165-222 for phase 'object only' (the reference's double query of the
        object points, :179 + recon_fit_base.py:515, is kept so results and cost are comparable)."""
        if phase != "object only":
            raise NotImplementedError(f"phase {phase!r} needs the silhouette renderer / contact terms (out of scope)")
        loss_dict = {}
        R = self.decopose_axis(obj_R, noise=noise)
        object = self.transform_obj_verts(data_dict["objects"], R, obj_t, obj_s)
        model.query(object, **data_dict["query_dict"])
        centers_pred_o = model.get_preds()[3]
        obj_center_pred = data_dict["smpl_center"] + torch.mean(centers_pred_o[:, 3:, :], -1)
        self.compute_obj_loss(data_dict, loss_dict, model, obj_s, object)
        obj_center_act = torch.mean(object, 1)
        loss_dict["ocent"] = F.mse_loss(obj_center_act, obj_center_pred, reduction="none").sum(-1).mean()
        return loss_dict

    def forward_smpl(self, smpl, data_dict, phase="global"):
        """Field terms of recon_fit_behave.py:293-337: df_h, part cross-entropy, fixed depth and
        initial-pose regularisers (priors / 2D keypoints need external assets and are skipped)."""
        loss_dict = {}
        model = data_dict["net"]
        smpl_verts, jtr, _, _ = smpl()
        _, parts_pred, _ = self.compute_df_h_loss(data_dict, loss_dict, model, smpl_verts)
        loss_dict["part"] = F.cross_entropy(parts_pred, data_dict["part_labels"], reduction="none").sum(-1).mean()
        if "pose_init" in data_dict:
            loss_dict["pinit"] = torch.mean(torch.sum((smpl.pose[:, 3:72] - data_dict["pose_init"]) ** 2, -1))
        return loss_dict
