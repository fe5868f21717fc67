"""A closed-form stand-in for the network with the CHORE query interface (`query(points, crop_center=...)`,
`get_preds()`): human / object distance fields of two spheres, smooth part logits, PCA and centre fields.
TEST INFRASTRUCTURE ONLY: it lets the point-cloud generation loop (recon/generator.py:50-79,123-217) run
deterministically on the CPU in BOTH the reference's `Generator` and `chore_b200.Generator`, so the host logic of
the latter (surface filter, index resampling with the CPU generator, truncation, argmax / mean reductions) is
pinned bit-for-bit against the former -- the "query-point indices" parity set of SURVEY.md section 8f-1.
With the real network on white-noise features the 10-step projection is chaotic and only stage-wise parity is
meaningful."""
from __future__ import annotations

import torch


class AnalyticField:
    OUT_DIST = 5.0

    def __init__(self):
        g = torch.Generator().manual_seed(123)
        self.c_h = torch.tensor([0.0, 0.1, 2.2])
        self.r_h = 0.45
        self.c_o = torch.tensor([0.6, -0.2, 2.35])
        self.r_o = 0.3
        self.w_parts = torch.randn(3, 14, generator=g)
        self.w_pca = torch.randn(3, 9, generator=g)
        self.preds = None

    # nn.Module-ish plumbing the generators touch
    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    def parameters(self):
        return []

    def query(self, points, crop_center=None, **_kw):
        assert crop_center is not None
        B, N, _ = points.shape
        d_h = ((points - self.c_h).norm(dim=-1) - self.r_h).abs()
        d_o = ((points - self.c_o).norm(dim=-1) - self.r_o).abs()
        df = torch.stack([d_h, d_o], 1)                                       # (B,2,N)
        rel = points - self.c_h
        parts = torch.matmul(rel, self.w_parts).transpose(1, 2)               # (B,14,N)
        pca = torch.tanh(torch.matmul(rel, self.w_pca)).transpose(1, 2).reshape(B, 3, 3, N)
        centers = torch.cat([self.c_h.view(1, 3, 1) + 0.01 * rel.transpose(1, 2),
                             (self.c_o - self.c_h).view(1, 3, 1) + 0.02 * torch.sin(rel.transpose(1, 2))], 1)   # (B,6,N)
        self.preds = (df, pca, parts, centers)

    def get_preds(self):
        return self.preds
