"""Silhouette phase of the object fit on the sm_100a rasteriser (csrc/silhouette.cu).

Mirrors, with the same names and argument meaning (paths relative to /root/reference):
  SilLossROI                       recon/obj_pose_roi.py:20-177   occlusion-aware silhouette term rendered in the object's ROI
  make_bbox_square / bbox_xy_to_wh / bbox_wh_to_xy   recon/bbox.py:26-74 (PHOSA helpers; detectron2's BoxMode.convert restated)
  mask2bbox                        recon/opt_utils.py:105-114
  projection                       external/neural_renderer/neural_renderer/projection.py:6-43
  vertices_to_faces                external/neural_renderer/neural_renderer/vertices_to_faces.py:4-22
  rasterize_silhouettes            external/neural_renderer/neural_renderer/rasterize.py:15-213,391-414 (alpha only)
  Renderer.render_silhouettes      external/neural_renderer/neural_renderer/renderer.py:119-152 (camera_mode='projection')

Third-party pieces of the reference that are not vendored and are restated from their documented behaviour:
  detectron2 BitMasks.crop_and_resize = ROIAlign((S, S), 1.0, 0, aligned=True) on the float mask, thresholded at 0.5
  (detectron2/structures/masks.py; detectron2's ROIAlign is torchvision.ops.roi_align) -- used once per image at set-up.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib

DEFAULT_NEAR, DEFAULT_FAR, DEFAULT_EPS = 0.1, 100.0, 1e-4          # neural_renderer/rasterize.py:8-13


# ---------------------------------------------------------------------------------------------
# bounding-box helpers (recon/bbox.py, recon/opt_utils.py)
# ---------------------------------------------------------------------------------------------
def mask2bbox(mask: np.ndarray) -> np.ndarray:
    """uint8 mask -> xyxy box of all contours above 127."""
    import cv2
    _, binary = cv2.threshold(mask, 127, 255, cv2.THRESH_BINARY)
    contours, _ = cv2.findContours(binary, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    lo, hi = np.array([50000, 50000]), np.array([-100, -100])
    for c in contours:
        x, y, w, h = cv2.boundingRect(c)
        lo, hi = np.minimum(lo, [x, y]), np.maximum(hi, [x + w, y + h])
    return np.concatenate([lo, hi])


def bbox_xy_to_wh(bbox):
    b = np.asarray(bbox, dtype=float).reshape(-1, 4).copy()
    b[:, 2:] -= b[:, :2]
    return b.reshape(np.shape(bbox))


def bbox_wh_to_xy(bbox):
    b = np.asarray(bbox, dtype=float).reshape(-1, 4).copy()
    b[:, 2:] += b[:, :2]
    return b.reshape(np.shape(bbox))


def make_bbox_square(bbox, bbox_expansion: float = 0.0):
    """xywh box(es) -> square xywh box(es) around the same centre, side = max(w, h) * (1 + expansion)."""
    b = np.array(bbox, dtype=float)
    shape = b.shape
    b = b.reshape(-1, 4)
    center = np.stack((b[:, 0] + b[:, 2] / 2, b[:, 1] + b[:, 3] / 2), axis=1)
    side = np.maximum(b[:, 2], b[:, 3])[:, None] * (1 + bbox_expansion)
    return np.hstack((center - side / 2, side, side)).reshape(shape)


def crop_and_resize_masks(masks: torch.Tensor, boxes_xyxy: torch.Tensor, size: int) -> torch.Tensor:
    """detectron2 BitMasks(masks).crop_and_resize(boxes, size): box i cropped from mask i, bilinear ROIAlign (aligned=True,
    adaptive sampling), >= 0.5 -> bool (size, size) per box."""
    from torchvision.ops import roi_align
    m = (masks != 0).to(torch.float32).cpu()[:, None]                 # BitMasks stores bool(mask)
    rois = torch.cat([torch.arange(len(boxes_xyxy), dtype=torch.float32)[:, None], boxes_xyxy.float().cpu()], 1)
    out = roi_align(m, rois, (size, size), spatial_scale=1.0, sampling_ratio=0, aligned=True)[:, 0]
    return out >= 0.5


# ---------------------------------------------------------------------------------------------
# neural_renderer pieces
# ---------------------------------------------------------------------------------------------
def projection(vertices, K, R, t, dist_coeffs, orig_size, eps: float = 1e-9):
    """K [R|t] projection to normalised image coordinates, v flipped, z kept (projection.py:6-43)."""
    vertices = torch.matmul(vertices, R.transpose(2, 1)) + t
    x, y, z = vertices[:, :, 0], vertices[:, :, 1], vertices[:, :, 2]
    x_, y_ = x / (z + eps), y / (z + eps)
    k1, k2, p1, p2, k3 = [dist_coeffs[:, None, i] for i in range(5)]
    r = torch.sqrt(x_ ** 2 + y_ ** 2)
    x__ = x_ * (1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)) + 2 * p1 * x_ * y_ + p2 * (r ** 2 + 2 * x_ ** 2)
    y__ = y_ * (1 + k1 * (r ** 2) + k2 * (r ** 4) + k3 * (r ** 6)) + p1 * (r ** 2 + 2 * y_ ** 2) + 2 * p2 * x_ * y_
    v = torch.matmul(torch.stack([x__, y__, torch.ones_like(z)], dim=-1), K.transpose(1, 2))
    u, vv = v[:, :, 0], orig_size - v[:, :, 1]
    u = 2 * (u - orig_size / 2.0) / orig_size
    vv = 2 * (vv - orig_size / 2.0) / orig_size
    return torch.stack([u, vv, z], dim=-1)


def vertices_to_faces(vertices, faces):
    """(B,V,3), (B,F,3) int -> (B,F,3,3)"""
    bs, nv = vertices.shape[:2]
    idx = faces.long() + (torch.arange(bs, device=vertices.device) * nv)[:, None, None]
    return vertices.reshape(bs * nv, 3)[idx]


class _RasterizeSilhouette(torch.autograd.Function):
    @staticmethod
    def forward(ctx, faces, image_size, near, far, eps, handle):
        f = faces.detach().contiguous().float()
        alpha, index = handle.silhouette_fwd(f, image_size, near, far)
        ctx.handle, ctx.eps = handle, eps
        ctx.save_for_backward(f, index, alpha)
        return alpha.clone()

    @staticmethod
    def backward(ctx, g_alpha):
        f, index, alpha = ctx.saved_tensors
        return ctx.handle.silhouette_bwd(f, index, alpha, g_alpha.contiguous().float(), ctx.eps), None, None, None, None, None


def rasterize_silhouettes(faces, image_size: int = 256, anti_aliasing: bool = True, near: float = DEFAULT_NEAR, far: float = DEFAULT_FAR,
                          eps: float = DEFAULT_EPS):
    """faces (B,F,3,3) in normalised image coordinates -> alpha (B,S,S), rows flipped like the reference (rasterize.py:318-322),
    2x supersampled + average-pooled when anti_aliasing."""
    size = image_size * 2 if anti_aliasing else image_size
    alpha = _RasterizeSilhouette.apply(faces, size, near, far, eps, _lib.get_handle(faces.device))
    alpha = alpha.flip(1)
    if anti_aliasing:
        alpha = torch.nn.functional.avg_pool2d(alpha[:, None], kernel_size=(2, 2))[:, 0]
    return alpha


class Renderer(nn.Module):
    """neural_renderer.Renderer restricted to what SilLossROI uses: camera_mode='projection', mode='silhouettes'."""

    def __init__(self, image_size=256, anti_aliasing=True, fill_back=True, K=None, R=None, t=None, dist_coeffs=None, orig_size=1024,
                 near=DEFAULT_NEAR, far=DEFAULT_FAR):
        super().__init__()
        self.image_size, self.anti_aliasing, self.fill_back = image_size, anti_aliasing, fill_back
        self.K, self.R, self.t, self.orig_size, self.near, self.far = K, R, t, orig_size, near, far
        self.dist_coeffs = dist_coeffs if dist_coeffs is not None else torch.zeros(1, 5, device=K.device)

    def forward(self, vertices, faces, mode="silhouettes"):
        if mode != "silhouettes":
            raise NotImplementedError("only the silhouette mode is on the fitting path")
        return self.render_silhouettes(vertices, faces)

    def render_silhouettes(self, vertices, faces):
        if self.fill_back:                       # both windings: renderer.py:121-123
            faces = torch.cat((faces, faces.flip(-1)), dim=1)
        vertices = projection(vertices, self.K, self.R, self.t, self.dist_coeffs, self.orig_size)
        return rasterize_silhouettes(vertices_to_faces(vertices, faces), self.image_size, self.anti_aliasing, self.near, self.far)


class SilLossROI(nn.Module):
    """Occlusion-aware silhouette loss rendered only in the object's region of interest (recon/obj_pose_roi.py:20-177)."""

    def __init__(self, person_masks, obj_masks, temp_mesh, crop_centers, rend_size=256, kernel_size=7, bbox_expansion=0.3,
                 device="cuda:0"):
        """person_masks / obj_masks: (B,H,W) network-input masks in [0,1]; temp_mesh: centred template with .v (V,3) and .f (F,3);
        crop_centers (B,2): crop centres of the network input in the 2048-px image."""
        super().__init__()
        self.net_input_size = 512
        self.temp_mesh = temp_mesh
        B = person_masks.shape[0]
        obj_bboxes = self.masks2bboxes(obj_masks).astype(float)                          # xyxy
        squares = make_bbox_square(bbox_xy_to_wh(obj_bboxes), bbox_expansion)            # xywh
        squares_xyxy = torch.as_tensor(bbox_wh_to_xy(squares), dtype=torch.float32)
        obj_crop = crop_and_resize_masks(obj_masks, squares_xyxy, rend_size)
        ps_crop = crop_and_resize_masks(person_masks, squares_xyxy, rend_size)
        scale = 1200 / 512.0                                                             # crop size / network input size
        Ks, keep, refs = [], [], []
        for ps, ob, bbox, cc in zip(ps_crop, obj_crop, squares, crop_centers):
            keep.append(self.cvt_masks(ps, ob).float())
            refs.append((ob > 0).float())
            Ks.append(self.compute_K_roi(self.to_original_bbox(bbox, scale, np.asarray(cc.detach().cpu(), dtype=float))))
        self.register_buffer("image_ref", torch.stack(refs, 0).to(device))
        self.register_buffer("keep_mask", torch.stack(keep, 0).to(device))
        self.pool = nn.MaxPool2d(kernel_size=kernel_size, stride=1, padding=kernel_size // 2)
        self.prepare_dist_trans(refs)
        self.prepare_render(temp_mesh, B, torch.cat(Ks, 0).to(device), rend_size, device)
        self.to(device)

    def prepare_render(self, temp_mesh, batch_size, cam_Ks, rend_size, device):
        verts = torch.as_tensor(np.asarray(temp_mesh.v), dtype=torch.float32)
        faces = torch.as_tensor(np.asarray(temp_mesh.f).astype(np.int64))
        self.register_buffer("vertices", verts.repeat(batch_size, 1, 1).to(device))
        self.register_buffer("faces", faces.repeat(batch_size, 1, 1).to(device))
        self.renderer = Renderer(image_size=rend_size, K=cam_Ks, R=torch.eye(3, device=device).unsqueeze(0),
                                 t=torch.zeros(1, 3, device=device), orig_size=1, anti_aliasing=False)

    def prepare_dist_trans(self, image_refs, power=0.25):
        from scipy.ndimage import distance_transform_edt
        edges = []
        for ref in image_refs:
            e = self.compute_edges(ref.unsqueeze(0)).cpu().numpy()
            edges.append(distance_transform_edt(1 - (e > 0)) ** (power * 2))
        self.register_buffer("edt_ref_edge", torch.from_numpy(np.concatenate(edges, 0)).float())

    def compute_edges(self, silhouette):
        return self.pool(silhouette) - silhouette

    @staticmethod
    def to_original_bbox(bbox_square, scale, trans, crop_size=1200):
        b = np.array(bbox_square, dtype=float)
        b *= scale
        b[:2] += trans - crop_size / 2.0
        return b

    @staticmethod
    def compute_K_roi(bbox_square, kinect_width=2048):
        """Kinect intrinsics re-expressed for the square ROI (unit image size)."""
        x, y, b, w = bbox_square
        assert b == w, "the given bbox is not square!"
        fx, fy, cx, cy = 979.7844 / kinect_width, 979.840 / kinect_width, 1018.952 / kinect_width, 779.486 / kinect_width
        return torch.tensor([[[fx * kinect_width / b, 0, (cx * kinect_width - x) / b],
                              [0, fy * kinect_width / b, (cy * kinect_width - y) / b], [0, 0, 1]]], dtype=torch.float32)

    @staticmethod
    def cvt_masks(person_mask, obj_mask):
        """1 = object or background (counts), 0 = occluded by the person and not object (ignored)."""
        fore, ps = obj_mask > 0.5, person_mask > 0.5
        inv = -ps.clone().float()
        inv[fore] = 1.0
        return inv >= 0

    @staticmethod
    def masks2bboxes(masks):
        return np.stack([mask2bbox((m.detach().cpu().numpy() * 255).astype(np.uint8)) for m in masks], 0)

    def apply_transformation(self, R, obj_t, obj_s):
        return obj_s.view(-1, 1, 1) * (torch.bmm(self.vertices, R) + obj_t.unsqueeze(1))

    def forward(self, R, obj_t, obj_s):
        """-> ({'mask': L2 silhouette loss}, masked render, its edges, reference mask, edge distance transform)"""
        verts = self.apply_transformation(R, obj_t, obj_s)
        image = self.keep_mask * self.renderer(verts, self.faces, mode="silhouettes")
        loss = {"mask": torch.sum((image - self.image_ref) ** 2, dim=(1, 2)).mean()}
        return loss, image, self.compute_edges(image), self.image_ref, self.edt_ref_edge
