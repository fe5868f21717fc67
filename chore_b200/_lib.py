"""ctypes binding of libchore_b200.so (include/chore_b200.h).

The product path has no CPU or eager fallback: if the library is missing or the device is not
sm_100 every entry point raises.  torch is used only for device memory and streams
(`tensor.data_ptr()`, `torch.cuda.current_stream()`).
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Iterable, Optional, Tuple

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libchore_b200.so")

HEAD_DF, HEAD_PCA, HEAD_PARTS, HEAD_CENTERS, HEAD_ALL = 1, 2, 4, 8, 15
HEAD_OUT = (2, 9, 14, 6)   # kernel head order: df, pca, parts, centers


class ChoreError(RuntimeError):
    pass


class AdamEntry(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("grad_acc", C.c_void_p), ("rows", C.c_int), ("cols", C.c_int), ("grad_ld", C.c_int)]


class TensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int), ("shape", C.c_int64 * 4),
                ("on_device", C.c_int)]


_P, _I, _U32, _I64, _F = C.c_void_p, C.c_int, C.c_uint32, C.c_int64, C.c_float
# every symbol include/chore_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "chore_create": (_I, [_I, C.POINTER(_P)]),
    "chore_destroy": (None, [_P]),
    "chore_last_error": (C.c_char_p, []),
    "chore_abi_version": (_I, []),
    "chore_launch_count": (C.c_uint64, []),
    "chore_load_weights": (_I, [_P, C.POINTER(TensorDesc), _I]),
    "chore_encode": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P]),
    "chore_query_fwd": (_I, [_P, _P, _P, _I, _I, _P, _P, _I, _I, _U32, _P, _P, _P, _P, _P, _P]),
    "chore_query_bwd": (_I, [_P, _P, _P, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "chore_query_bwd_workspace_bytes": (C.c_size_t, [_I, _I]),
    "chore_query_bwd_ws": (_I, [_P, _P, _P, _I, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "chore_query_grid": (_I, [_P, _P, _P, _I, _I, _P, _I, C.POINTER(_I), C.POINTER(C.c_double), C.POINTER(C.c_double),
                              _I64, _I64, _U32, _P, _P, _P, _P, _P]),
    "chore_lbs_load_model": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I]),
    "chore_lbs_fwd": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "chore_lbs_bwd": (_I, [_P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P]),
    "chore_landmarks_load": (_I, [_P, _P, _P, _P, _I, _I, _I]),
    "chore_landmarks_fwd": (_I, [_P, _P, _I, _P, _P]),
    "chore_landmarks_bwd": (_I, [_P, _P, _I, _P, _I, _P]),
    "chore_fit_workspace_floats": (C.c_size_t, [_I, _I]),
    "chore_fit_smpl_field_grads": (_I, [_P, _P, _P, _P, _I, _I, _F, _F, _P, _P, _P, _P, _P]),
    "chore_fit_landmark_grads": (_I, [_P, _P, _P, _P, _I, _I, _I, _F, _F, _F, C.POINTER(C.c_float), _P, _P, _P, _P]),
    "chore_fit_pose_prior_grads": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _F, _P, _P, _P, _P]),
    "chore_fit_obj_field_grads": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _F, _F, _F, _P, _P, _P, _P, _P, _P]),
    "chore_add_rowvec": (_I, [_P, _P, _P, _I, _I, _F, _P]),
    "chore_surface_clamp_grad": (_I, [_P, _P, _I, _F, _I, _I, _P, _P]),
    "chore_surface_step": (_I, [_P, _P, _P, _P, _I, _F, _I, _I, _P, _P]),
    "chore_silhouette_workspace_bytes": (C.c_size_t, [_I, _I]),
    "chore_silhouette_fwd": (_I, [_P, _P, _I, _I, _I, _F, _F, _P, _P, _P, C.c_size_t, _P]),
    "chore_silhouette_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P]),
    "chore_contact_workspace_bytes": (C.c_size_t, [_I, _I, _I]),
    "chore_contact_loss": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "chore_gen_compact": (_I, [_P, _P, _I, _F, _F, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    "chore_gen_resample": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, C.c_uint64, C.c_uint64, _P, _P, _P, _P]),
    "chore_gen_total": (_I, [_P, _P, _I, _P, _P]),
    "chore_gen_finalize": (_I, [_P, _P, _P, _I, _I, _P, _P, _P, _P]),
    "chore_adam_step": (_I, [_P, C.POINTER(AdamEntry), _I, _F, _F, _F, _F, _P, _P, _P, _P]),
    "chore_rigid_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P]),
    "chore_rigid_bwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "chore_project_so3": (_I, [_P, _P, _I, _P, _P]),
    "chore_project_so3_bwd": (_I, [_P, _P, _P, _I, _P, _P]),
}

_lib = None
_lock = threading.RLock()
_handles: Dict[int, "Handle"] = {}


def load_library() -> C.CDLL:
    """dlopen the in-tree library and type every entry point.  Raises if it is not built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ChoreError(f"{LIB_PATH} is missing: build it with `python -m chore_b200.build` "
                             "(there is no CPU / PyTorch fallback for the CHORE hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)     # AttributeError here = header / library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
        return lib


def launch_count() -> int:
    return int(load_library().chore_launch_count())


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def check_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ChoreError(f"expected a contiguous fp32 CUDA tensor, got {t.dtype} {t.device} "
                             f"contiguous={t.is_contiguous()} shape={tuple(t.shape)}")


class Handle:
    """One chore_handle per CUDA device."""

    def __init__(self, device: int):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise ChoreError("no CUDA device: the CHORE hot path runs on sm_100a only (no CPU fallback)")
        self.device = device
        h = _P()
        self._check(self.lib.chore_create(device, C.byref(h)))
        self.h = h
        self._keep = []          # host tensors that must outlive a call

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise ChoreError(f"libchore_b200 error {rc}: {self.lib.chore_last_error().decode()}")

    def close(self) -> None:
        if self.h:
            self.lib.chore_destroy(self.h)
            self.h = None

    # ---- weights -------------------------------------------------------------------------
    def load_weights(self, state_dict: Dict[str, torch.Tensor]) -> None:
        items = [(k, v.detach().float().contiguous()) for k, v in state_dict.items() if torch.is_tensor(v) and v.dim() >= 1]
        arr = (TensorDesc * len(items))()
        for d, (k, v) in zip(arr, items):
            d.name = k.encode()
            d.data = v.data_ptr()
            d.ndim = min(v.dim(), 4) if v.dim() != 3 else 3
            for i, s in enumerate(v.shape[:4]):
                d.shape[i] = s
            d.on_device = int(v.is_cuda)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_load_weights(self.h, arr, len(items)))

    # ---- encoder -------------------------------------------------------------------------
    def encode(self, images: torch.Tensor, want_normx: bool = True) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
        check_cuda(images)
        B, Cin, H, W = images.shape
        if Cin != 5:
            raise ChoreError(f"images must have 5 channels (RGB + 2 masks), got {Cin}")
        feat = torch.empty(B, H // 4, W // 4, 256, device=images.device)
        skip = torch.empty(B, H // 2, W // 2, 64, device=images.device)
        normx = torch.empty(B, H // 4, W // 4, 128, device=images.device) if want_normx else None
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_encode(self.h, images.data_ptr(), B, H, W, feat.data_ptr(), skip.data_ptr(),
                                              _ptr(normx), _stream()))
        return feat, skip, normx

    # ---- point query ---------------------------------------------------------------------
    def query_fwd(self, feat: torch.Tensor, skip: torch.Tensor, points: torch.Tensor, crop_center: torch.Tensor,
                  head_mask: int = HEAD_ALL, want_in_img: bool = False):
        check_cuda(feat, skip, points, crop_center)
        B, N = points.shape[0], points.shape[1]
        fh, fw = feat.shape[1], feat.shape[2]
        if feat.shape[0] != B or feat.shape[3] != 256 or tuple(skip.shape) != (B, 2 * fh, 2 * fw, 64):
            raise ChoreError(f"feature maps {tuple(feat.shape)} / {tuple(skip.shape)} do not match B={B}")
        if tuple(crop_center.shape) != (B, 2) or points.shape[2] != 3:
            raise ChoreError("points must be (B,N,3) and crop_center (B,2)")
        outs = [torch.empty(B, c, N, device=points.device) if head_mask & (1 << i) else None
                for i, c in enumerate(HEAD_OUT)]
        in_img = torch.empty(B, N, dtype=torch.uint8, device=points.device) if want_in_img else None
        if N == 0:                      # empty query: nothing to launch (data_ptr() of an empty tensor is NULL)
            return outs, in_img
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_query_fwd(self.h, feat.data_ptr(), skip.data_ptr(), fh, fw, points.data_ptr(),
                                                 crop_center.data_ptr(), B, N, head_mask, *[_ptr(o) for o in outs],
                                                 _ptr(in_img), _stream()))
        return outs, in_img

    def query_bwd(self, feat, skip, points, crop_center, grads: Iterable[Optional[torch.Tensor]]) -> torch.Tensor:
        grads = [None if g is None else g.contiguous() for g in grads]
        check_cuda(feat, skip, points, crop_center, *grads)
        B, N = points.shape[0], points.shape[1]
        g_points = torch.empty_like(points)
        if N == 0:
            return g_points
        # scratch comes from torch's allocator: stable under CUDA-graph capture (graph-private pool)
        nbytes = int(self.lib.chore_query_bwd_workspace_bytes(B, N))
        # k x the base size lets the k heads with a gradient run concurrently in ONE launch (latency-bound small queries)
        k = sum(g is not None for g in grads)
        if k > 1 and nbytes * k <= (512 << 20):
            nbytes *= k
        ws = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=points.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_query_bwd_ws(self.h, feat.data_ptr(), skip.data_ptr(), feat.shape[1], feat.shape[2],
                                                    points.data_ptr(), crop_center.data_ptr(), B, N,
                                                    *[_ptr(g) for g in grads], g_points.data_ptr(), ws.data_ptr(), nbytes,
                                                    _stream()))
        return g_points

    def query_grid(self, feat, skip, crop_center, b: int, res, b_min, b_max, start: int, count: int,
                   head_mask: int, outs) -> None:
        """outs: per-head (nout, total) tensors of image b (or None); fills columns [start, start+count)."""
        check_cuda(feat, skip, crop_center, *outs)
        res_c = (_I * 3)(*[int(r) for r in res])
        mn = (C.c_double * 3)(*[float(x) for x in b_min])      # float64 like create_grid's numpy arithmetic
        mx = (C.c_double * 3)(*[float(x) for x in b_max])
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_query_grid(self.h, feat.data_ptr(), skip.data_ptr(), feat.shape[1], feat.shape[2],
                                                  crop_center.data_ptr(), b, res_c, mn, mx, start, count, head_mask,
                                                  *[_ptr(o) for o in outs], _stream()))

    # ---- SMPL-H ----------------------------------------------------------------------------
    def lbs_load_model(self, v_template, shapedirs, posedirs, J_regressor, weights, parents) -> None:
        ts = [t.detach().float().contiguous().cpu() for t in (v_template, shapedirs, posedirs, J_regressor, weights)]
        par = torch.as_tensor(parents).to(torch.int32).contiguous().cpu()
        V, J, nb = ts[4].shape[0], ts[4].shape[1], ts[1].shape[-1]
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_lbs_load_model(self.h, *[t.data_ptr() for t in ts], par.data_ptr(), V, J, nb, 0))
        self.lbs_shape = (V, J, nb)

    def lbs_fwd(self, pose, betas, trans, offsets, want_posed: bool = True):
        check_cuda(pose, betas, trans, offsets)
        V, J, _ = self.lbs_shape
        B = pose.shape[0]
        verts = torch.empty(B, V, 3, device=pose.device)
        jtr = torch.empty(B, J, 3, device=pose.device)
        v_posed = torch.empty(B, V, 3, device=pose.device) if want_posed else None
        naked = torch.empty(B, V, 3, device=pose.device) if want_posed else None
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_lbs_fwd(self.h, pose.data_ptr(), betas.data_ptr(), trans.data_ptr(), _ptr(offsets),
                                               B, verts.data_ptr(), jtr.data_ptr(), _ptr(v_posed), _ptr(naked), _stream()))
        return verts, jtr, v_posed, naked

    def lbs_bwd(self, pose, betas, trans, offsets, g_verts, g_jtr, want_offsets: bool):
        g_verts = g_verts.contiguous()
        g_jtr = None if g_jtr is None else g_jtr.contiguous()
        check_cuda(pose, betas, trans, offsets, g_verts, g_jtr)
        B = pose.shape[0]
        g_pose, g_betas, g_trans = torch.empty_like(pose), torch.empty_like(betas), torch.empty_like(trans)
        g_off = torch.empty_like(g_verts) if want_offsets else None
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_lbs_bwd(self.h, pose.data_ptr(), betas.data_ptr(), trans.data_ptr(), _ptr(offsets),
                                               B, g_verts.data_ptr(), _ptr(g_jtr), g_pose.data_ptr(), g_betas.data_ptr(),
                                               g_trans.data_ptr(), _ptr(g_off), _stream()))
        return g_pose, g_betas, g_trans, g_off

    # ---- landmark regressors -------------------------------------------------------------------
    def landmarks_load(self, rowptr, col, val, L: int, V: int) -> None:
        """CSR (host) of the stacked (L,V) regressor matrix."""
        rowptr = torch.as_tensor(rowptr).to(torch.int32).contiguous().cpu()
        col = torch.as_tensor(col).to(torch.int32).contiguous().cpu()
        val = torch.as_tensor(val).to(torch.float32).contiguous().cpu()
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_landmarks_load(self.h, rowptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                                      int(L), int(V), int(val.numel())))
        self.lmk_shape = (int(L), int(V))

    def landmarks_fwd(self, verts):
        check_cuda(verts)
        L, V = self.lmk_shape
        if verts.shape[1] != V or verts.shape[2] != 3:
            raise ChoreError(f"verts must be (B,{V},3), got {tuple(verts.shape)}")
        out = torch.empty(verts.shape[0], L, 3, device=verts.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_landmarks_fwd(self.h, verts.data_ptr(), verts.shape[0], out.data_ptr(), _stream()))
        return out

    def landmarks_bwd(self, g_out, g_verts: Optional[torch.Tensor] = None):
        """g_out (B,L,3) -> g_verts (B,V,3); accumulates into `g_verts` when one is given."""
        g_out = g_out.contiguous()
        check_cuda(g_out, g_verts)
        L, V = self.lmk_shape
        acc = g_verts is not None
        if g_verts is None:
            g_verts = torch.empty(g_out.shape[0], V, 3, device=g_out.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_landmarks_bwd(self.h, g_out.data_ptr(), g_out.shape[0], g_verts.data_ptr(), int(acc),
                                                     _stream()))
        return g_verts

    # ---- fit-step losses / optimiser (csrc/fit_loss.cu) ------------------------------------------
    def fit_workspace(self, B: int, N: int, device) -> torch.Tensor:
        return torch.empty(int(self.lib.chore_fit_workspace_floats(B, N)), device=device)

    def fit_smpl_field_grads(self, df, parts, labels, wd: float, wp: float, loss, ws):
        check_cuda(df, parts, loss, ws)
        if labels.dtype != torch.int64 or not labels.is_contiguous():
            raise ChoreError("labels must be a contiguous int64 tensor")
        B, _, N = df.shape
        g_df, g_parts = torch.empty_like(df), torch.empty_like(parts)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_fit_smpl_field_grads(self.h, df.data_ptr(), parts.data_ptr(), labels.data_ptr(), B, N, wd, wp,
                                                            g_df.data_ptr(), g_parts.data_ptr(), loss.data_ptr(), ws.data_ptr(), _stream()))
        return g_df, g_parts

    def fit_landmark_grads(self, lm, kpts, crop_center, n_joints: int, z0: float, cz: float, cj: float, cam, loss, ws):
        check_cuda(lm, kpts, crop_center, loss, ws)
        g_lm = torch.empty_like(lm)
        cam_c = (C.c_float * 6)(*[float(x) for x in cam])
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_fit_landmark_grads(self.h, lm.data_ptr(), _ptr(kpts), crop_center.data_ptr(), lm.shape[0], lm.shape[1],
                                                          n_joints, z0, cz, cj, cam_c, g_lm.data_ptr(), loss.data_ptr(), ws.data_ptr(),
                                                          _stream()))
        return g_lm

    def fit_pose_prior_grads(self, pose, pose_init, priors, cb: float, ch: float, cp: float, g_pose, loss, ws) -> None:
        """priors: (body_mean, body_prec, hand_mean, lhand_prec, rhand_prec) or None; adds into g_pose."""
        pr = priors if priors is not None else (None,) * 5
        check_cuda(pose, pose_init, g_pose, loss, ws, *pr)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_fit_pose_prior_grads(self.h, pose.data_ptr(), _ptr(pose_init), *[_ptr(t) for t in pr], pose.shape[0],
                                                            pose.shape[1], cb, ch, cp, g_pose.data_ptr(), loss.data_ptr(), ws.data_ptr(),
                                                            _stream()))

    def fit_obj_field_grads(self, obj, df, centers, smpl_center, s, s0: float, wo: float, wc: float, wsc: float, loss, ws):
        check_cuda(obj, df, centers, smpl_center, s, loss, ws)
        B, N = obj.shape[0], obj.shape[1]
        g_df, g_cen = torch.empty_like(df), torch.empty_like(centers)
        dvec = torch.empty(B, 3, device=obj.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_fit_obj_field_grads(self.h, obj.data_ptr(), df.data_ptr(), centers.data_ptr(), smpl_center.data_ptr(),
                                                           s.data_ptr(), B, N, s0, wo, wc, wsc, g_df.data_ptr(), g_cen.data_ptr(),
                                                           dvec.data_ptr(), loss.data_ptr(), ws.data_ptr(), _stream()))
        return g_df, g_cen, dvec

    def add_rowvec(self, x, v, alpha: float) -> None:
        check_cuda(x, v)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_add_rowvec(self.h, x.data_ptr(), v.data_ptr(), x.shape[0], x.shape[1], alpha, _stream()))

    def surface_clamp_grad(self, df, df_idx: int, threshold: float):
        check_cuda(df)
        g_df = torch.empty_like(df)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_surface_clamp_grad(self.h, df.data_ptr(), df_idx, threshold, df.shape[0], df.shape[2], g_df.data_ptr(),
                                                          _stream()))
        return g_df

    def surface_step(self, points, g_points, df, df_idx: int, threshold: float):
        check_cuda(points, g_points, df)
        out = torch.empty_like(points)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_surface_step(self.h, points.data_ptr(), g_points.data_ptr(), df.data_ptr(), df_idx, threshold,
                                                    points.shape[0], points.shape[1], out.data_ptr(), _stream()))
        return out

    # ---- silhouette rasteriser (csrc/silhouette.cu) ----------------------------------------------------
    def silhouette_fwd(self, faces, image_size: int, near: float = 0.1, far: float = 100.0):
        """faces (B,F,3,3) -> (alpha (B,S,S), face_index int32 (B,S,S)), rows not flipped."""
        check_cuda(faces)
        B, F = faces.shape[0], faces.shape[1]
        alpha = torch.empty(B, image_size, image_size, device=faces.device)
        index = torch.empty(B, image_size, image_size, dtype=torch.int32, device=faces.device)
        nbytes = int(self.lib.chore_silhouette_workspace_bytes(B, F))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=faces.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_silhouette_fwd(self.h, faces.data_ptr(), B, F, image_size, near, far, alpha.data_ptr(),
                                                      index.data_ptr(), ws.data_ptr(), nbytes, _stream()))
        return alpha, index

    def silhouette_bwd(self, faces, face_index, alpha, g_alpha, eps: float = 1e-4):
        check_cuda(faces, alpha, g_alpha)
        B, F, S = faces.shape[0], faces.shape[1], alpha.shape[1]
        g_faces = torch.empty_like(faces)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_silhouette_bwd(self.h, faces.data_ptr(), face_index.data_ptr(), alpha.data_ptr(), g_alpha.data_ptr(),
                                                      B, F, S, eps, g_faces.data_ptr(), _stream()))
        return g_faces

    # ---- joint-phase contact term (csrc/contact.cu) ----------------------------------------------------
    def contact_loss(self, smpl_verts, obj, df_hum_o, df_obj_h, part_o, part_labels, thresh: float = 0.08, want_grads: bool = True):
        """-> (loss (1,), n_pairs (1,) int32, g_smpl, g_obj)"""
        check_cuda(smpl_verts, obj, df_hum_o, df_obj_h, part_o)
        B, Nh, No = smpl_verts.shape[0], smpl_verts.shape[1], obj.shape[1]
        dev = obj.device
        loss, pairs = torch.zeros(1, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
        g_s = torch.empty_like(smpl_verts) if want_grads else None
        g_o = torch.empty_like(obj) if want_grads else None
        nbytes = int(self.lib.chore_contact_workspace_bytes(B, Nh, No))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_contact_loss(self.h, smpl_verts.data_ptr(), obj.data_ptr(), df_hum_o.data_ptr(), df_obj_h.data_ptr(),
                                                    part_o.data_ptr(), part_labels.data_ptr(), B, Nh, No, thresh, loss.data_ptr(),
                                                    pairs.data_ptr(), _ptr(g_s), _ptr(g_o), ws.data_ptr(), nbytes, _stream()))
        return loss, pairs, g_s, g_o

    # ---- generator bookkeeping (csrc/generator.cu) --------------------------------------------------
    def gen_compact(self, df, df_idx, threshold, filter_val, samples, packed, iter_count, surf=None, preds=None, out=None) -> None:
        """out = (points (B,cap,3), labels int32 (B,cap), pca (B,cap,9), centers (B,cap,6), count int32 (B,)) or None."""
        B, N = samples.shape[0], samples.shape[1]
        append = out is not None
        cap = out[0].shape[1] if append else 0
        pca, parts, centers = preds if append else (None, None, None)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_gen_compact(self.h, df.data_ptr(), df_idx, threshold, filter_val, _ptr(surf), samples.data_ptr(),
                                                   _ptr(pca), _ptr(parts), _ptr(centers), B, N, cap, int(append),
                                                   *([_ptr(o) for o in out] if append else [None] * 5), packed.data_ptr(),
                                                   iter_count.data_ptr(), _stream()))

    def gen_resample(self, packed, iter_count, samples_init, sample_num, sigma_hit, sigma_miss, seed, offset, uniforms=None,
                     normals=None):
        B, N = packed.shape[0], packed.shape[1]
        out = torch.empty(B, sample_num, 3, device=packed.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_gen_resample(self.h, packed.data_ptr(), iter_count.data_ptr(), samples_init.data_ptr(), B, N,
                                                    samples_init.shape[1], sample_num, sigma_hit, sigma_miss, seed, offset,
                                                    _ptr(uniforms), _ptr(normals), out.data_ptr(), _stream()))
        return out

    def gen_total(self, iter_count, samples_count) -> None:
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_gen_total(self.h, iter_count.data_ptr(), iter_count.shape[0], samples_count.data_ptr(), _stream()))

    def gen_finalize(self, out_pca, out_centers, samples_count):
        B, cap = out_pca.shape[0], out_pca.shape[1]
        pca_mean, cen_mean = torch.empty(B, 9, device=out_pca.device), torch.empty(B, 6, device=out_pca.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_gen_finalize(self.h, out_pca.data_ptr(), out_centers.data_ptr(), B, cap, samples_count.data_ptr(),
                                                    pca_mean.data_ptr(), cen_mean.data_ptr(), _stream()))
        return pca_mean, cen_mean

    def adam_step(self, entries, n: int, lr: float, beta1: float, beta2: float, eps: float, step, gscale=None, loss=None) -> None:
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_adam_step(self.h, entries, n, lr, beta1, beta2, eps, step.data_ptr(), _ptr(gscale), _ptr(loss),
                                                 _stream()))

    # ---- rigid object ------------------------------------------------------------------------
    def rigid_fwd(self, verts, R, t, s):
        check_cuda(verts, R, t, s)
        out = torch.empty_like(verts)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_rigid_fwd(self.h, verts.data_ptr(), R.data_ptr(), t.data_ptr(), s.data_ptr(),
                                                 verts.shape[0], verts.shape[1], out.data_ptr(), _stream()))
        return out

    def rigid_bwd(self, verts, R, t, s, g_out, want_verts: bool):
        g_out = g_out.contiguous()
        check_cuda(verts, R, t, s, g_out)
        g_R, g_t, g_s = torch.empty_like(R), torch.empty_like(t), torch.empty_like(s)
        g_v = torch.empty_like(verts) if want_verts else None
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_rigid_bwd(self.h, verts.data_ptr(), R.data_ptr(), t.data_ptr(), s.data_ptr(),
                                                 verts.shape[0], verts.shape[1], g_out.data_ptr(), g_R.data_ptr(),
                                                 g_t.data_ptr(), g_s.data_ptr(), _ptr(g_v), _stream()))
        return g_R, g_t, g_s, g_v

    def project_so3(self, mats):
        check_cuda(mats)
        out = torch.empty_like(mats)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_project_so3(self.h, mats.data_ptr(), mats.shape[0], out.data_ptr(), _stream()))
        return out

    def project_so3_bwd(self, mats, g_out):
        g_out = g_out.contiguous()
        check_cuda(mats, g_out)
        g = torch.empty_like(mats)
        with torch.cuda.device(self.device):
            self._check(self.lib.chore_project_so3_bwd(self.h, mats.data_ptr(), g_out.data_ptr(), mats.shape[0],
                                                       g.data_ptr(), _stream()))
        return g


def get_handle(device=None) -> Handle:
    """The process-wide handle of a CUDA device (created on first use)."""
    load_library()
    if not torch.cuda.is_available():
        raise ChoreError("no CUDA device: the CHORE hot path runs on sm_100a only (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    with _lock:
        if idx not in _handles:
            _handles[idx] = Handle(idx)
        return _handles[idx]
