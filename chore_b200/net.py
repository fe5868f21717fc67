"""Host-side mirror of the reference network interface on top of libchore_b200.so.

`CHORE` presents exactly what the reference's callers use:
`filter(images)`, `query(points, crop_center=...)`, `get_preds()`, `get_im_feat()`, `.eval()`,
`.to(device)`, `.load_state_dict(checkpoint['model_state_dict'])`, and the attributes `preds`,
`im_feat_list`, `tmpx`, `normx`, `OUT_DIST` -- so `recon/generator.py` and the fitters can hold it
instead of `model.chore.CHORE` (model/chore.py:11-174, model/BasePIFuNet.py:42-70).

The predictions are autograd-connected to `points` through `_QueryFn`, whose backward is the
hand-written gradient kernel (no autograd graph through the MLP / grid_sample / projection).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import HEAD_ALL, ChoreError


def _to_nhwc(t: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) in any memory format -> contiguous (B,H,W,C).  Free for tensors made by `filter`."""
    return t.permute(0, 2, 3, 1).contiguous()


class _QueryFn(torch.autograd.Function):
    """CHORE.query as one op: forward = chore_query_fwd, backward = chore_query_bwd."""

    @staticmethod
    def forward(ctx, points, crop_center, feat, skip, handle, head_mask):
        pts = points.detach().contiguous().float()
        outs, _ = handle.query_fwd(feat, skip, pts, crop_center, head_mask)
        ctx.handle, ctx.head_mask = handle, head_mask
        ctx.save_for_backward(pts, crop_center, feat, skip)
        df, pca, parts, centers = outs
        B, N = pts.shape[0], pts.shape[1]
        # heads that were not requested come back as empty placeholders
        z = lambda c: pts.new_zeros(B, c, 0)
        return (df if df is not None else z(2), pca if pca is not None else z(9),
                parts if parts is not None else z(14), centers if centers is not None else z(6))

    @staticmethod
    def backward(ctx, g_df, g_pca, g_parts, g_centers):
        pts, crop_center, feat, skip = ctx.saved_tensors
        grads = []
        for i, g in enumerate((g_df, g_pca, g_parts, g_centers)):
            grads.append(g if (g is not None and ctx.head_mask & (1 << i) and g.numel() > 0) else None)
        if all(g is None for g in grads):
            return torch.zeros_like(pts), None, None, None, None, None
        g_points = ctx.handle.query_bwd(feat, skip, pts, crop_center, grads)
        return g_points, None, None, None, None, None


class heads:
    """`with heads(model, HEAD_DF | HEAD_PARTS): model.query(...)` -- evaluate only the decoder heads a caller
    reads (the kernel skips the others entirely; their entries of get_preds() are empty placeholders).  A no-op
    for models without a `head_mask` attribute, e.g. the reference's own CHORE."""

    def __init__(self, model, mask: int):
        self.model, self.mask = model, mask

    def __enter__(self):
        self.old = getattr(self.model, "head_mask", None)
        if self.old is not None:
            self.model.head_mask = self.mask
        return self.model

    def __exit__(self, *exc):
        if self.old is not None:
            self.model.head_mask = self.old
        return False


class CHORE(nn.Module):
    """Drop-in for model.chore.CHORE at inference / fitting time (chore-release configuration)."""

    OUT_DIST = 5.0     # model/chore.py:65
    Z0 = 2.2

    def __init__(self, opt=None, device="cuda:0", rank: int = -1, **_):
        super().__init__()
        self.name = "chore_b200"
        self.opt = opt
        if opt is not None:
            # the kernels are compiled for the chore-release configuration (config/chore-release.json)
            expect = {"z_feat": "xyz", "projection_mode": "perspective", "skip_hourglass": True, "num_stack": 5,
                      "hourglass_dim": 256, "num_hourglass": 2, "norm": "group", "hg_down": "ave_pool", "loadSize": 1200}
            for k, v in expect.items():
                got = getattr(opt, k, v)
                assert got == v, f"chore_b200 is built for {k}={v!r}, the config asks for {got!r}"
        self._device = torch.device(device if rank < 0 else f"cuda:{rank}")
        self._handle: Optional[_lib.Handle] = None
        self._sd: Dict[str, torch.Tensor] = {}
        self._anchor = nn.Parameter(torch.zeros(1), requires_grad=False)   # gives .parameters() a device
        self.im_feat_list: List[torch.Tensor] = []
        self.tmpx: Optional[torch.Tensor] = None
        self.normx: Optional[torch.Tensor] = None
        self.preds = None
        self.intermediate_preds_list = []
        self.head_mask = HEAD_ALL
        self._nhwc_cache = {}

    # ---- nn.Module plumbing -------------------------------------------------------------------
    @property
    def handle(self) -> _lib.Handle:
        if self._handle is None:
            idx = self._device.index if self._device.index is not None else 0
            self._handle = _lib.Handle(idx)
            if self._sd:
                self._handle.load_weights(self._sd)
        return self._handle

    def to(self, device=None, *a, **k):
        if device is not None and torch.device(device).type == "cuda" and torch.device(device) != self._device:
            self._device = torch.device(device)
            if self._handle is not None:
                self._handle.close()
                self._handle = None
        return super().to(device, *a, **k)

    def state_dict(self, *a, **k):
        return dict(self._sd)

    def load_state_dict(self, state_dict, strict: bool = True):
        """Accepts checkpoint['model_state_dict'] of the reference (561 tensors; an optional DDP
        'module.' prefix is stripped like recon/generator.py:255-262 does)."""
        sd = {(k[7:] if k.startswith("module.") else k): v.detach().float().cpu() for k, v in state_dict.items()}
        if strict:
            need = ["image_filter.conv1.weight", "df.0.weight", "df.6.bias", "part_predictor.6.weight",
                    "pca_predictor.6.weight", "center_predictor.6.weight", "image_filter.l4.weight"]
            missing = [k for k in need if k not in sd]
            if missing:
                raise RuntimeError(f"Error(s) in loading state_dict for CHORE: missing keys {missing}")
        self._sd = sd
        if self._handle is not None:
            self._handle.load_weights(sd)
        return torch.nn.modules.module._IncompatibleKeys([], [])

    # ---- the reference interface ------------------------------------------------------------
    def filter(self, images: torch.Tensor) -> None:
        """CHORE.filter (model/chore.py:87-96), eval mode: keeps the last stack output."""
        images = images.to(self._device, torch.float32).contiguous()
        feat, skip, normx = self.handle.encode(images)
        # NCHW *views* of channels-last memory: same values/shapes as the reference tensors
        self.im_feat_list = [feat.permute(0, 3, 1, 2)]
        self.tmpx = skip.permute(0, 3, 1, 2)
        self.normx = normx.permute(0, 3, 1, 2)
        self._nhwc_cache = {}

    def _maps(self) -> Tuple[torch.Tensor, torch.Tensor]:
        assert self.im_feat_list and self.tmpx is not None, "call filter(images) before query()"
        out = []
        for name, t in (("feat", self.im_feat_list[-1]), ("skip", self.tmpx)):
            # keyed on the tensor OBJECT (kept alive here, so its storage cannot be recycled for
            # another map with the same address) and its version counter (in-place edits)
            hit = self._nhwc_cache.get(name)
            if hit is None or hit[0] is not t or hit[1] != t._version:
                hit = (t, t._version, _to_nhwc(t.detach().to(self._device, torch.float32)))
                self._nhwc_cache[name] = hit
            out.append(hit[2])
        return out[0], out[1]

    def project_points(self, points: torch.Tensor, offsets: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError("projection is fused into the query kernel; use query()/get_preds()")

    def query(self, points: torch.Tensor, crop_center: Optional[torch.Tensor] = None, **kwargs) -> None:
        """CHORE.query (model/chore.py:107-154).  points (B,N,3) camera space, crop_center (B,2)."""
        assert crop_center is not None, "crop_center (B,2) is required"
        assert points.dim() == 3 and points.shape[-1] == 3, "points must be (B,N,3)"
        self.points, self.crop_center = points, crop_center
        feat, skip = self._maps()
        cc = crop_center.detach().to(self._device, torch.float32).contiguous()
        pts = points if points.is_cuda else points.to(self._device)
        df, pca, parts, centers = _QueryFn.apply(pts, cc, feat, skip, self.handle, self.head_mask)
        B, N = pts.shape[0], pts.shape[1]
        self.preds = (df, pca.view(B, 3, 3, -1), parts, centers)     # model/chore.py:156-167
        self.intermediate_preds_list = [self.preds]

    def get_preds(self):
        return self.preds

    def get_im_feat(self):
        return self.im_feat_list[-1]

    def forward(self, *a, **k):
        raise NotImplementedError("training forward (model/chore.py:176-242) is out of scope; use filter()/query()")

    # ---- dense grid (model/sdf.py:4-48 semantics on top of query) ------------------------------
    @torch.no_grad()
    def query_grid(self, res, b_min, b_max, crop_center: torch.Tensor, batch_index: int = 0,
                   head_mask: int = _lib.HEAD_DF, chunk: int = 1 << 22):
        """Evaluates the field on create_grid(res, b_min, b_max) of image `batch_index` without
        materialising the coordinates.  Returns per-head (n_out, X*Y*Z) tensors (None if not asked)."""
        feat, skip = self._maps()
        cc = crop_center.detach().to(self._device, torch.float32).contiguous()
        total = int(res[0]) * int(res[1]) * int(res[2])
        outs = [torch.empty(c, total, device=self._device) if head_mask & (1 << i) else None
                for i, c in enumerate(_lib.HEAD_OUT)]
        for start in range(0, total, chunk):
            self.handle.query_grid(feat, skip, cc, batch_index, res, b_min, b_max, start, min(chunk, total - start),
                                   head_mask, outs)
        return outs
