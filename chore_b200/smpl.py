"""SMPL-H body model on libchore_b200.so: mirrors `SMPL_Layer`
(lib_smpl/smplpytorch/smplpytorch/pytorch/smpl_layer.py:21-175) and the parameter wrappers
`SMPLPyTorchWrapperBatch` / `SMPLPyTorchWrapperBatchSplitParams` (lib_smpl/wrapper_pytorch.py:23-218).

The licensed SMPL-H pickle is not read here: the layer is built from the registered-buffer
tensors themselves (`v_template`, `shapedirs`, `posedirs`, `J_regressor`, `weights`, `faces`,
`parents`), i.e. `SMPLHLayer.from_reference_layer(smpl_layer)` for a loaded reference layer, or
a dict of those arrays.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib

TOP_BETA_NUM, GLOBAL_POSE_NUM, BODY_POSE_NUM = 2, 3, 63       # lib_smpl/const.py
SMPLH_POSE_PRAMS_NUM, SMPL_POSE_PRAMS_NUM = 156, 72


class _LbsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, betas, trans, offsets, handle):
        args = [t.detach().contiguous().float() for t in (pose, betas, trans)]
        off = None if offsets is None else offsets.detach().contiguous().float()
        verts, jtr, v_posed, naked = handle.lbs_fwd(*args, off)
        ctx.handle = handle
        ctx.has_off = off is not None
        ctx.save_for_backward(*args, *(() if off is None else (off,)))
        ctx.mark_non_differentiable(v_posed, naked)
        return verts, jtr, v_posed, naked

    @staticmethod
    def backward(ctx, g_verts, g_jtr, _gp, _gn):
        saved = ctx.saved_tensors
        pose, betas, trans = saved[:3]
        off = saved[3] if ctx.has_off else None
        if g_verts is None:
            g_verts = torch.zeros(pose.shape[0], ctx.handle.lbs_shape[0], 3, device=pose.device)
        want_off = ctx.has_off and ctx.needs_input_grad[3]
        g_pose, g_betas, g_trans, g_off = ctx.handle.lbs_bwd(pose, betas, trans, off, g_verts, g_jtr, want_off)
        return g_pose, g_betas, g_trans, g_off, None


class _LandmarkFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, handle):
        ctx.handle = handle
        return handle.landmarks_fwd(verts.detach().contiguous().float())

    @staticmethod
    def backward(ctx, g_out):
        return ctx.handle.landmarks_bwd(g_out), None


def load_regressors(assets_root: str):
    """lib_smpl/body_landmark.py:16-28 (batch_size=None branch): the body25 / face / hand regressor
    pickles (scipy sparse, stored (V,L)) transposed to (L,V)."""
    import pickle as pkl
    from os.path import join
    out = []
    for name in ("body25_regressor.pkl", "face_regressor.pkl", "hand_regressor.pkl"):
        with open(join(assets_root, name), "rb") as f:
            out.append(pkl.load(f, encoding="latin1").T)
    return tuple(out)


def _to_csr(reg):
    """One regressor (scipy sparse, torch sparse COO (optionally batched like the reference's stacked
    tensors), or dense (L,V)) -> (rowptr, col, val, L, V) with duplicate entries summed."""
    import numpy as np
    import scipy.sparse as sp
    if sp.issparse(reg):
        m = sp.csr_matrix(reg)
    elif torch.is_tensor(reg) and reg.is_sparse:
        r = reg.coalesce()
        idx, val = r.indices().cpu().numpy(), r.values().cpu().numpy()
        if idx.shape[0] == 3:                       # (B,L,V) stack of identical matrices: take batch 0
            keep = idx[0] == 0
            idx, val = idx[1:, keep], val[keep]
        m = sp.csr_matrix((val, (idx[0], idx[1])), shape=tuple(reg.shape[-2:]))
    else:
        m = sp.csr_matrix(np.asarray(torch.as_tensor(reg).detach().cpu().numpy()))
    m.sum_duplicates()
    m.sort_indices()
    return m


class LandmarkRegressors:
    """The three landmark regressors stacked into one CSR matrix on the device (chore_landmarks_*)."""

    def __init__(self, regressors, handle):
        import scipy.sparse as sp
        mats = [_to_csr(r) for r in regressors]
        self.sizes = [m.shape[0] for m in mats]
        full = sp.vstack(mats).tocsr()
        full.sort_indices()
        self.handle = handle
        handle.landmarks_load(torch.from_numpy(full.indptr.astype("int32")), torch.from_numpy(full.indices.astype("int32")),
                              torch.from_numpy(full.data.astype("float32")), full.shape[0], full.shape[1])

    def __call__(self, verts):
        return torch.split(_LandmarkFn.apply(verts, self.handle), self.sizes, dim=1)


class SMPLHLayer(nn.Module):
    """SMPL_Layer.forward on the LBS kernels.  `forward(pose, th_betas, th_trans, th_offsets)`
    -> (verts, jtr, v_posed, naked) like smpl_layer.py:72-175 (hands=True, scale 1)."""

    def __init__(self, buffers: Dict[str, torch.Tensor], device="cuda:0"):
        super().__init__()
        self._device = torch.device(device)
        self.handle = _lib.Handle(self._device.index or 0)
        self.handle.lbs_load_model(buffers["v_template"].reshape(-1, 3), buffers["shapedirs"], buffers["posedirs"],
                                   buffers["J_regressor"], buffers["weights"], buffers["parents"])
        self.register_buffer("th_faces", torch.as_tensor(buffers["faces"]).long())
        self.kintree_parents = [int(p) for p in buffers["parents"]]
        self.num_joints = len(self.kintree_parents)
        self.num_verts = int(buffers["weights"].shape[0])

    @classmethod
    def from_reference_layer(cls, layer, device="cuda:0"):
        """Build from a loaded reference SMPL_Layer (reads its registered buffers, smpl_layer.py:49-70)."""
        par = list(layer.kintree_parents)
        par[0] = -1
        return cls({"v_template": layer.th_v_template[0], "shapedirs": layer.th_shapedirs,
                    "posedirs": layer.th_posedirs, "J_regressor": layer.th_J_regressor,
                    "weights": layer.th_weights, "faces": layer.th_faces, "parents": torch.tensor(par)}, device)

    def forward(self, th_pose_axisang, th_betas=None, th_trans=None, th_offsets=None, scale=1.0):
        B = th_pose_axisang.shape[0]
        dev = self._device
        if th_betas is None:
            th_betas = torch.zeros(B, self.handle.lbs_shape[2], device=dev)
        if th_trans is None:
            th_trans = torch.zeros(B, 3, device=dev)
        assert scale == 1.0, "scale != 1 is not used on the CHORE path"
        return _LbsFn.apply(th_pose_axisang, th_betas, th_trans, th_offsets, self.handle)


def _make_regressors(regressors, model: "SMPLHLayer"):
    if regressors is None or isinstance(regressors, LandmarkRegressors):
        return regressors
    return LandmarkRegressors(regressors, model.handle)


class SMPLPyTorchWrapperBatch(nn.Module):
    """lib_smpl/wrapper_pytorch.py:23-90 with the LBS kernels underneath.  `model` is an
    SMPLHLayer (the reference passes a model_root and loads the pickle itself)."""

    def __init__(self, model: SMPLHLayer, batch_sz: int, betas=None, pose=None, trans=None, offsets=None,
                 gender="male", num_betas=10, hands=True, device="cuda:0", regressors=None):
        super().__init__()
        self.model_root = model
        self.hands, self.device, self.gender = hands, device, gender
        npose = SMPLH_POSE_PRAMS_NUM if hands else SMPL_POSE_PRAMS_NUM
        nv = model.num_verts
        mk = lambda t, shape: nn.Parameter(torch.zeros(*shape) if t is None else torch.as_tensor(t).float().clone())
        self.betas = mk(betas, (batch_sz, num_betas))
        self.pose = mk(pose, (batch_sz, npose))
        self.trans = mk(trans, (batch_sz, 3))
        self.offsets = mk(offsets, (batch_sz, nv, 3))
        assert self.pose.shape[1] == npose, f"pose shape {tuple(self.pose.shape)} does not match hands={hands}"
        self.smpl = model
        self.faces = model.th_faces.clone()
        # landmark regressors (lib_smpl/body_landmark.py:16-28): 3 x (L,V) sparse/dense matrices, an
        # already-built LandmarkRegressors, or None
        self.regressors = _make_regressors(regressors, model)
        self.to(device)

    def forward(self):
        return self.smpl(self.pose, th_betas=self.betas, th_trans=self.trans, th_offsets=self.offsets)

    def get_landmarks(self, verts=None):
        """body25 / face / hand landmarks = sparse regressors applied to the posed vertices
        (wrapper_pytorch.py:78-90) in one CSR kernel.  The reference re-runs the LBS here; pass the
        `verts` of a forward() of the same parameters to skip that second LBS (same values)."""
        assert self.regressors is not None, "no landmark regressors were given"
        return self.regressors(self.forward()[0] if verts is None else verts)


class SMPLPyTorchWrapperBatchSplitParams(nn.Module):
    """lib_smpl/wrapper_pytorch.py:93-218: the same model with independently optimisable blocks."""

    def __init__(self, model: SMPLHLayer, batch_sz: int, top_betas=None, other_betas=None, global_pose=None,
                 body_pose=None, hand_pose=None, trans=None, offsets=None, faces=None, gender="male", hands=True,
                 num_betas=10, device="cuda:0", regressors=None):
        super().__init__()
        self.model_root = model
        nv = model.num_verts
        hand_num = 90 if hands else 6
        mk = lambda t, shape: nn.Parameter(torch.zeros(*shape) if t is None else torch.as_tensor(t).float().clone())
        self.top_betas = mk(top_betas, (batch_sz, TOP_BETA_NUM))
        self.other_betas = mk(other_betas, (batch_sz, num_betas - TOP_BETA_NUM))
        self.global_pose = mk(global_pose, (batch_sz, GLOBAL_POSE_NUM))
        self.body_pose = mk(body_pose, (batch_sz, BODY_POSE_NUM))
        self.hand_pose = mk(hand_pose, (batch_sz, hand_num))
        self.trans = mk(trans, (batch_sz, 3))
        self.offsets = mk(offsets, (batch_sz, nv, 3))
        with torch.no_grad():   # values only: building an autograd graph here would bind the parameters' grad
            # accumulators to the construction-time (legacy default) stream and break CUDA-graph capture
            self.betas = torch.cat([self.top_betas, self.other_betas], 1)
            self.pose = torch.cat([self.global_pose, self.body_pose, self.hand_pose], 1)
        self.faces, self.gender, self.hands, self.device = faces, gender, hands, device
        self.smpl = model
        self.regressors = _make_regressors(regressors, model)
        self.to(device)

    def forward(self):
        self.betas = torch.cat([self.top_betas, self.other_betas], 1)
        self.pose = torch.cat([self.global_pose, self.body_pose, self.hand_pose], 1)
        return self.smpl(self.pose, th_betas=self.betas, th_trans=self.trans, th_offsets=self.offsets)

    def get_landmarks(self, verts=None):
        """wrapper_pytorch.py:176-190; see SMPLPyTorchWrapperBatch.get_landmarks."""
        assert self.regressors is not None, "no landmark regressors were given"
        return self.regressors(self.forward()[0] if verts is None else verts)

    @staticmethod
    def from_smpl(smpl: SMPLPyTorchWrapperBatch):
        B = smpl.pose.shape[0]
        p, b = smpl.pose.data, smpl.betas.data
        return SMPLPyTorchWrapperBatchSplitParams(
            smpl.model_root, B, trans=smpl.trans.data, top_betas=b[:, :TOP_BETA_NUM], other_betas=b[:, TOP_BETA_NUM:],
            global_pose=p[:, :GLOBAL_POSE_NUM], body_pose=p[:, GLOBAL_POSE_NUM:GLOBAL_POSE_NUM + BODY_POSE_NUM],
            hand_pose=p[:, GLOBAL_POSE_NUM + BODY_POSE_NUM:], offsets=smpl.offsets.data, faces=smpl.faces,
            gender=smpl.gender, hands=smpl.hands, num_betas=b.shape[1], device=smpl.device, regressors=smpl.regressors)
