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


class GraphedStep:
    """One optimisation step (zero_grad -> losses -> backward -> optimizer.step) captured in a CUDA graph
    and replayed: no Python, no autograd bookkeeping and no launch gaps on the hot loop.  This is the
    B200-side replacement for the per-step Python of optimize_smpl / optimize_smpl_object
    (recon/recon_fit_behave.py:90-163,224-291), which additionally syncs the device every step for its
    tqdm strings (.item(), :154-157).

    `step_fn()` must run the whole step on the current stream with static tensors (use
    `torch.optim.Adam(..., capturable=True)` and device-side RNG) and return the loss tensor(s); the
    returned tensors are refreshed in place by every `__call__`."""

    def __init__(self, step_fn: Callable, warmup: int = 3):
        self.step_fn = step_fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):          # lazy allocations / attribute setup happen outside the capture
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
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


class FusedFitSteps:
    """The two inner optimisation steps of the fitting loop with NO autograd graph: explicit forward kernels,
    closed-form loss gradients (a handful of elementwise torch ops under no_grad) and explicit adjoint kernels,
    so a step is  LBS -> field query -> dL/d(preds) -> query adjoint -> LBS adjoint -> Adam  (SMPL phase,
    recon/recon_fit_behave.py:293-337 field terms) or  SO(3) -> rigid -> query -> dL/d(preds) -> query adjoint
    -> rigid adjoint -> SO(3) adjoint -> Adam  ('object only' phase, recon/recon_fit_behave.py:165-198).
    Both are plain stream-ordered launches, hence capturable in a CUDA graph (`graphed()`).

    The reference queries the object points twice per step (recon_fit_behave.py:179 + recon_fit_base.py:515);
    the two queries have identical inputs, so one forward + one adjoint launch (heads df and centers) gives the
    same losses and gradients.  Loss weights / decay are those of get_loss_weights()."""

    W = {"object": 30.0 ** 2, "part": 0.05 ** 2, "scale": 10.0 ** 2, "df_h": 30.0 ** 2, "ocent": 15 ** 2, "pinit": 5 ** 2}

    def __init__(self, net, smpl, data_dict, obj_R, obj_t, obj_s, lr_smpl=0.006, lr_obj=0.006, obj_scale=1.0, decay=1.0):
        self.net, self.smpl, self.data = net, smpl, data_dict
        self.R, self.t, self.s = obj_R, obj_t, obj_s
        self.obj_scale, self.decay = obj_scale, decay
        self.smpl_params = [smpl.trans, smpl.global_pose, smpl.body_pose, smpl.top_betas, smpl.other_betas]
        self.opt_smpl = torch.optim.Adam(self.smpl_params, lr_smpl, capturable=True)
        self.opt_obj = torch.optim.Adam([obj_t, obj_R, obj_s], lr_obj, capturable=True)
        self.h_net = net.handle
        self.h_lbs = smpl.smpl.handle
        self.h_aux = _lib.get_handle(obj_t.device)
        self.cc = data_dict["query_dict"]["crop_center"].detach().float().contiguous()

    @torch.no_grad()
    def smpl_step(self):
        sm, d = self.smpl, self.data
        feat, skip = self.net._maps()
        pose = torch.cat([sm.global_pose, sm.body_pose, sm.hand_pose], 1)
        betas = torch.cat([sm.top_betas, sm.other_betas], 1)
        trans, off = sm.trans.detach(), sm.offsets.detach()
        verts, _, _, _ = self.h_lbs.lbs_fwd(pose, betas, trans, off, want_posed=False)
        (df, _, parts, _), _ = self.h_net.query_fwd(feat, skip, verts, self.cc, _lib.HEAD_DF | _lib.HEAD_PARTS)
        B, _, N = df.shape
        k = 1.0 / (1.0 + self.decay)
        # L = w_dfh * mean(min(df_h, 0.1)) + w_part * mean_b sum_n CE(parts, labels)
        dfh = df[:, 0]
        logp = torch.log_softmax(parts, 1)
        labels = d["part_labels"]
        loss = self.W["df_h"] * k * torch.clamp(dfh, max=0.1).mean() - self.W["part"] * k * logp.gather(1, labels.unsqueeze(1)).sum() / B
        g_df = torch.zeros_like(df)
        g_df[:, 0] = (dfh <= 0.1).float() * (self.W["df_h"] * k / (B * N))
        g_parts = logp.exp()
        g_parts.scatter_add_(1, labels.unsqueeze(1), torch.full_like(labels, -1, dtype=g_parts.dtype).unsqueeze(1))
        g_parts *= self.W["part"] * k / B
        g_verts = self.h_net.query_bwd(feat, skip, verts, self.cc, [g_df, None, g_parts, None])
        g_pose, g_betas, g_trans, _ = self.h_lbs.lbs_bwd(pose, betas, trans, off, g_verts, None, False)
        if "pose_init" in d:     # 5^2 * mean_b sum (pose[3:72] - pose_init)^2
            diff = pose[:, 3:72] - d["pose_init"]
            loss = loss + self.W["pinit"] * k * (diff ** 2).sum(-1).mean()
            g_pose[:, 3:72] += self.W["pinit"] * k * 2.0 / B * diff
        sm.trans.grad, sm.global_pose.grad, sm.body_pose.grad = g_trans, g_pose[:, :3].contiguous(), g_pose[:, 3:66].contiguous()
        sm.top_betas.grad, sm.other_betas.grad = g_betas[:, :2].contiguous(), g_betas[:, 2:].contiguous()
        self.opt_smpl.step()
        return loss

    @torch.no_grad()
    def object_step(self, noise: Optional[torch.Tensor] = None):
        d = self.data
        feat, skip = self.net._maps()
        obj0 = d["objects"]
        B, N, _ = obj0.shape
        if noise is None:
            noise = torch.rand(B, 3, 3, device=obj0.device)
        rot_in = (self.R + 1e-4 * noise).contiguous()                        # decopose_axis (recon_fit_base.py:373-384)
        Rm = self.h_aux.project_so3(rot_in)
        t, s = self.t.detach(), self.s.detach()
        obj = self.h_aux.rigid_fwd(obj0, Rm, t, s)
        (df, _, _, cen), _ = self.h_net.query_fwd(feat, skip, obj, self.cc, _lib.HEAD_DF | _lib.HEAD_CENTERS)
        k = 1.0 / (1.0 + self.decay)
        dfo = df[:, 1]
        dvec = obj.mean(1) - d["smpl_center"] - cen[:, 3:].mean(-1)          # (B,3)
        loss = (self.W["object"] * k * torch.clamp(dfo, max=0.8).mean() + self.W["scale"] * k * ((s - self.obj_scale) ** 2).mean()
                + self.W["ocent"] * k * (dvec ** 2).sum(-1).mean())
        g_df = torch.zeros_like(df)
        g_df[:, 1] = (dfo <= 0.8).float() * (self.W["object"] * k / (B * N))
        g_cen = torch.zeros_like(cen)
        coef = self.W["ocent"] * k * 2.0 / (B * N)
        g_cen[:, 3:] = (-coef * dvec).unsqueeze(-1)
        g_obj = self.h_net.query_bwd(feat, skip, obj, self.cc, [g_df, None, None, g_cen])
        g_obj += (coef * dvec).unsqueeze(1)
        g_R, g_t, g_s, _ = self.h_aux.rigid_bwd(obj0, Rm, t, s, g_obj, False)
        g_s += self.W["scale"] * k * 2.0 / B * (s - self.obj_scale)
        self.R.grad, self.t.grad, self.s.grad = self.h_aux.project_so3_bwd(rot_in, g_R), g_t, g_s
        self.opt_obj.step()
        return loss

    def graphed(self):
        """(smpl_step, object_step) captured in CUDA graphs; each call replays one optimisation step."""
        return GraphedStep(self.smpl_step), GraphedStep(self.object_step)
