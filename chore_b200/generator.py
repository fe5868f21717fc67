"""Neural point-cloud generator on the query kernels: mirrors recon/generator.py:Generator
(approx_surface :50-79, gen_pc_batch :123-188, compose_outdict :190-217, init_samples :275-282)
and the dense-grid evaluation of model/sdf.py:4-48 (create_grid / eval_grid)."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .net import heads


class Generator:
    def __init__(self, model, threshold: float = 2.0, filter_val: float = 0.004, sparse_thres: float = 0.03,
                 device="cuda:0", fused: bool = True):
        """fused: run approx_surface through the autograd-free kernel sequence when the model is a chore_b200.CHORE
        (same values up to the last bit of F.normalize's norm); False keeps the reference's autograd formulation."""
        self.fused = fused
        self.model = model
        self.threshold = threshold
        self.filter_val = filter_val
        self.sparse_thres = sparse_thres
        self.device = device
        self.pmin = np.array([-3.0, -0.9, 0.2])      # recon/generator.py:45-48
        self.pmax = np.array([3.0, 1.8, 4.0])

    # ---- recon/generator.py:50-79 ---------------------------------------------------------------
    def approx_surface(self, model, samples, num_steps, query_input, df_type):
        """num_steps x { query; t = clamp(df_k, max=thr); t.sum().backward(); p <- p - normalize(grad) * t }"""
        df_idx = 0 if df_type == "human" else 1
        if self.fused and hasattr(model, "handle") and samples.is_cuda and num_steps > 0:
            return self._approx_surface_fused(model, samples, num_steps, query_input, df_idx)
        preds = None
        for step in range(num_steps):
            # only the distance head drives the projection; the other three are needed from the LAST query only
            # (the reference returns that query's preds), so the earlier steps evaluate one head of four
            with heads(model, getattr(model, "head_mask", 15) if step == num_steps - 1 else _lib.HEAD_DF):
                model.query(samples, **query_input)
            preds = model.get_preds()
            df_target = torch.clamp(preds[0][:, df_idx, :], max=self.threshold)
            df_target.sum().backward()
            gradient = samples.grad.detach()
            samples = samples.detach() - F.normalize(gradient, dim=2) * df_target.detach().unsqueeze(-1)
            samples = samples.detach()
            samples.requires_grad = True
        return samples, preds

    @torch.no_grad()
    def _approx_surface_fused(self, model, samples, num_steps, query_input, df_idx):
        """The same projection loop without autograd: per step  query (df head only; all heads on the last step, whose
        preds are returned) -> d clamp(df).sum() / d df -> query adjoint -> p - normalize(grad) * t , i.e. two query
        launches and two elementwise kernels instead of an autograd round trip and ~10 eager ops."""
        h = model.handle
        feat, skip = model._maps()
        cc = query_input["crop_center"].detach().to(samples.device, torch.float32).contiguous()
        p = samples.detach().float().contiguous()
        B, N = p.shape[0], p.shape[1]
        outs = None
        for step in range(num_steps):
            last = step == num_steps - 1
            mask = getattr(model, "head_mask", _lib.HEAD_ALL) | _lib.HEAD_DF if last else _lib.HEAD_DF
            outs, _ = h.query_fwd(feat, skip, p, cc, mask)
            g_df = h.surface_clamp_grad(outs[0], df_idx, float(self.threshold))
            g_p = h.query_bwd(feat, skip, p, cc, [g_df, None, None, None])
            p = h.surface_step(p, g_p, outs[0], df_idx, float(self.threshold))
        z = lambda c: p.new_zeros(B, c, 0)
        df, pca, parts, centers = [o if o is not None else z(c) for o, c in zip(outs, _lib.HEAD_OUT)]
        preds = (df, pca.view(B, 3, 3, -1), parts, centers)
        model.preds = preds
        p.requires_grad = True
        return p, preds

    def init_samples(self, sample_num, batch_size=1):
        """recon/generator.py:275-282 (only batch element 0 is rescaled there; kept)."""
        samples = torch.rand(batch_size, sample_num, 3).float().to(self.device)
        samples[0, :, 0] = samples[0, :, 0] * 6 - 3
        samples[0, :, 1] = samples[0, :, 1] * 5 - 2.5
        samples[0, :, 2] = (samples[0, :, 2] - 0.5) * 0.5 + 2.2
        return samples

    # ---- recon/generator.py:123-217 -------------------------------------------------------------
    def gen_pc_batch(self, model, df_type, samples_init, num_points, query_input, num_steps=10, max_iter=100,
                     sample_num=20000) -> Dict[str, torch.Tensor]:
        df_idx = 0 if df_type == "human" else 1
        B = samples_init.shape[0]
        names = ["points", "pca_axis", "parts", "centers"]
        out = {n: [[] for _ in range(B)] for n in names}
        it, count = 0, 0
        samples = samples_init.clone().to(self.device)
        samples.requires_grad = True
        while count < num_points:
            surf, preds = self.approx_surface(model, samples, num_steps, query_input, df_type)
            df_t = torch.clamp(preds[0][:, df_idx, :], max=self.threshold).detach()
            mask = df_t < self.filter_val
            if it > 0:
                counts = []
                for i in range(B):
                    m = mask[i]
                    out["points"][i].append(surf[i, m].detach())
                    out["pca_axis"][i].append(preds[1][i][..., m].detach())
                    out["parts"][i].append(preds[2][i][:, m].detach())
                    out["centers"][i].append(preds[3][i][:, m].detach())
                    counts.append(int(m.sum()))
                count += min(counts)
            new = []
            for i in range(B):
                s_i = samples[i, mask[i], :].detach().unsqueeze(0)
                if s_i.shape[1] > 1:
                    idx = torch.randint(s_i.shape[1], (sample_num,)).to(self.device)
                    s_i = s_i[:, idx] + (self.threshold / 3) * torch.randn(1, sample_num, 3).to(self.device)
                else:
                    idx = torch.randint(samples_init.shape[1], (sample_num,))
                    s_i = samples_init[i:i + 1, idx].to(self.device) + 0.5 * torch.randn(1, sample_num, 3).to(self.device)
                new.append(s_i)
            samples = torch.cat(new, 0).detach()
            samples.requires_grad = True
            it += 1
            if it == max_iter:
                raise RuntimeError("point generation failed after 100 iterations")
        res = {}
        for n in names:
            comb = []
            for i in range(B):
                if n == "points":
                    comb.append(torch.cat(out[n][i], 0)[:count])
                    continue
                o = torch.cat(out[n][i], -1)[..., :count]
                comb.append(torch.argmax(o, 0) if n == "parts" else torch.mean(o, -1))
            res[n] = torch.stack(comb, 0)
        return res

    # ---- model/sdf.py:4-48 semantics --------------------------------------------------------------
    @torch.no_grad()
    def eval_grid_chunk(self, resolution: Sequence[int], crop_center: torch.Tensor, batch_index: int, start: int, count: int,
                        outs, head_mask: int = _lib.HEAD_DF) -> None:
        """One batch_eval chunk (model/sdf.py:30-41): grid points [start, start + count) of create_grid(resolution, pmin,
        pmax) into the preallocated per-head (n_out, X*Y*Z) tensors `outs` -- lets a caller overlap the copy of one chunk
        with the evaluation of the next."""
        feat, skip = self.model._maps()
        cc = crop_center.detach().to(feat.device, torch.float32).contiguous()
        self.model.handle.query_grid(feat, skip, cc, batch_index, resolution, self.pmin, self.pmax, start, count, head_mask, outs)

    @torch.no_grad()
    def eval_grid(self, resolution: Sequence[int], crop_center: torch.Tensor, batch_index: int = 0,
                  head_mask: int = _lib.HEAD_DF, chunk: int = 1 << 22):
        """Field on create_grid(resolution, pmin, pmax): per-head (n_out, X, Y, Z) tensors."""
        outs = self.model.query_grid(resolution, self.pmin, self.pmax, crop_center, batch_index, head_mask, chunk)
        return [None if o is None else o.view(o.shape[0], *[int(r) for r in resolution]) for o in outs]
