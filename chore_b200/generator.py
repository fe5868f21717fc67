"""Neural point-cloud generator on the query kernels: mirrors recon/generator.py:Generator
(approx_surface :50-79, gen_pc_batch :123-188, compose_outdict :190-217, init_samples :275-282)
and the dense-grid evaluation of model/sdf.py:4-48 (create_grid / eval_grid)."""
from __future__ import annotations

import os
from glob import glob
from os.path import isfile
from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib
from .net import heads


class Generator:
    """Drop-in for recon/generator.py:Generator (same constructor arguments, defaults, methods and result dict).

    Extensions, all keyword-only:
      fused      approx_surface as the autograd-free kernel sequence (same values up to the last bit of F.normalize's norm);
                 False keeps the reference's autograd formulation on the same kernels.
      rng        "reference": gen_pc_batch draws indices / noise from torch's CPU generator in the reference's order (bit-exact
                 against the reference loop, tests/test_generator_cpu.py) and keeps its per-image Python bookkeeping;
                 "device": the whole loop stays on the GPU (csrc/generator.cu: ordered stream compaction, Philox resampling,
                 min-count and final reductions), ONE host sync per outer iteration (the loop condition).  Seed contract:
                 Philox4x32-10, seed = `seed`, subsequence = image * sample_num + sample, offset = outer iteration.
      experiments_root   where experiments/<exp_name>/checkpoints lives (the reference derives it from its own file location).
    exp_name=None skips checkpoint discovery (weights already loaded into `model`)."""

    def __init__(self, model, exp_name=None, threshold=1.0, checkpoint=None, device="cuda", multi_gpus=True, sparse_thres=0.05,
                 filter_val=0.03, *, fused: bool = True, rng: str = "reference", seed: int = 0, experiments_root=None):
        assert rng in ("reference", "device")
        self.fused, self.rng, self.seed = fused, rng, int(seed)
        self.sparse_thres = sparse_thres
        self.filter_val = filter_val
        self.sample_num = 100000
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.model = model.to(self.device) if hasattr(model, "to") else model
        if hasattr(self.model, "eval"):
            self.model.eval()
        self.threshold = threshold
        self.multi_gpus = multi_gpus
        if exp_name is not None:
            root = experiments_root or os.path.join(os.getcwd(), "experiments")
            self.exp_path = os.path.join(root, exp_name) + os.sep
            self.checkpoint_path = os.path.join(self.exp_path, "checkpoints") + os.sep
            assert os.path.exists(self.checkpoint_path), f"{self.checkpoint_path} does not exist!"
            self.load_checkpoint(checkpoint)
        if hasattr(self.model, "parameters"):
            for param in self.model.parameters():
                param.requires_grad = False
        self.pmin = np.array([-3.0, -0.9, 0.2])      # recon/generator.py:45-48
        self.pmax = np.array([3.0, 1.8, 4.0])

    # ---- checkpoints (recon/generator.py:219-267) ------------------------------------------------------
    def get_val_min_ck(self):
        files = glob(self.exp_path + "val_min=*")
        if len(files) == 0:
            return None
        log = np.load(files[0])
        return log[2] if isfile(self.checkpoint_path + str(log[2])) else None

    def find_best_checkpoint(self, checkpoints):
        """val_min=<epoch>.npy names the best checkpoint; otherwise the latest `checkpoint_{h}h:{m}m:{s}s_{secs}.tar`."""
        best = self.get_val_min_ck()
        if best is not None:
            return self.checkpoint_path + best
        secs = np.sort(np.array([os.path.splitext(os.path.basename(p))[0].split("_")[-1] for p in checkpoints], dtype=float))[-1]
        h, m, sec = int(secs / 3600), int((secs / 60) % 60), int(secs % 60)
        return self.checkpoint_path + "checkpoint_{}h:{}m:{}s_{}.tar".format(h, m, sec, secs)

    def load_checkpoint(self, checkpoint):
        if checkpoint is None:
            checkpoints = glob(self.checkpoint_path + "/*")
            if len(checkpoints) == 0:
                print("No checkpoints found at {}".format(self.checkpoint_path))
                return 0, 0
            path = self.find_best_checkpoint(checkpoints)
        else:
            path = self.checkpoint_path + "{}".format(checkpoint)
        ck = torch.load(path, map_location="cpu", weights_only=False)
        print("Loaded checkpoint from: {}".format(path))
        sd = ck["model_state_dict"]
        if self.multi_gpus:                          # DistributedDataParallel prefix
            sd = {k.replace("module.", ""): v for k, v in sd.items()}
        self.model.load_state_dict(sd)
        return ck["epoch"], ck["training_time"]

    # ---- small pieces of the reference interface -----------------------------------------------------------
    def prep_query_input(self, batch):
        return {"crop_center": batch.get("crop_center").to(self.device)}

    def filter(self, data):
        "encode image features"
        self.model.filter(data["images"].to(self.device))

    def get_grid_samples(self, sample_num, batch_size=1):
        return self.init_samples(sample_num, batch_size)

    def generate_pclouds_batch(self, data, num_steps=10, num_points=50000, mute=False):
        """recon/generator.py:102-121: encode, then the neural point clouds of the human and the object field."""
        self.filter(data)
        batch_size = data.get("images").shape[0]
        samples = self.get_grid_samples(30000, batch_size=batch_size)
        return {t: self.gen_pc_batch(self.model, t, samples, num_points, data, num_steps, mute=mute) for t in ("human", "object")}

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
    def gen_pc_batch(self, model, df_type, samples_init, num_points, batch, num_steps=10, max_iter=100, mute=False,
                     sample_num=20000) -> Dict[str, torch.Tensor]:
        """recon/generator.py:123-188.  `batch`: the loader batch (or any dict with 'crop_center')."""
        query_input = self.prep_query_input(batch)
        if self.rng == "device" and hasattr(model, "handle") and self.device.type == "cuda":
            return self._gen_pc_batch_device(model, df_type, samples_init, num_points, query_input, num_steps, max_iter, sample_num)
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
                    out["points"][i].append(surf[i, m].detach().cpu())            # parse_preds (:88-100) collects on the host
                    out["pca_axis"][i].append(preds[1][i][..., m].detach().cpu())
                    out["parts"][i].append(preds[2][i][:, m].detach().cpu())
                    out["centers"][i].append(preds[3][i][:, m].detach().cpu())
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

    @torch.no_grad()
    def _gen_pc_batch_device(self, model, df_type, samples_init, num_points, query_input, num_steps, max_iter, sample_num,
                             randoms=None):
        """gen_pc_batch with every stage on the GPU.  `randoms`: optional callable (it, iter_count) -> (uniforms (B,sample_num),
        normals (B,sample_num,3)) replacing the Philox draws (tests replay the reference's CPU draws through it)."""
        h = model.handle
        df_idx = 0 if df_type == "human" else 1
        dev = self.device
        init = samples_init.detach().to(dev, torch.float32).contiguous()
        B, n_init = init.shape[0], init.shape[1]
        cap = int(num_points) + max(n_init, sample_num)
        out = (torch.empty(B, cap, 3, device=dev), torch.empty(B, cap, dtype=torch.int32, device=dev),
               torch.empty(B, cap, 9, device=dev), torch.empty(B, cap, 6, device=dev), torch.zeros(B, dtype=torch.int32, device=dev))
        iter_count = torch.zeros(B, dtype=torch.int32, device=dev)
        total = torch.zeros(1, dtype=torch.int32, device=dev)
        samples, it, count = init, 0, 0
        while count < num_points:
            surf, preds = self.approx_surface(model, samples, num_steps, query_input, df_type)
            n = samples.shape[1]
            packed = torch.empty(B, n, 3, device=dev)
            append = it > 0
            h.gen_compact(preds[0], df_idx, float(self.threshold), float(self.filter_val), samples.detach().contiguous(), packed, iter_count,
                          surf=surf.detach().contiguous(), preds=(preds[1].reshape(B, 9, -1), preds[2], preds[3]) if append else None,
                          out=out if append else None)
            if append:
                h.gen_total(iter_count, total)
            u, nrm = randoms(it, iter_count) if randoms is not None else (None, None)
            samples = h.gen_resample(packed, iter_count, init, sample_num, float(self.threshold) / 3, 0.5, self.seed + (df_idx << 32), it, u, nrm)
            if append:
                count = int(total.item())            # the only host sync of the outer iteration: the loop condition
            it += 1
            if it == max_iter:
                raise RuntimeError("point generation failed after 100 iterations")
        pca_mean, cen_mean = h.gen_finalize(out[2], out[3], total)
        return {"points": out[0][:, :count].clone(), "pca_axis": pca_mean.view(B, 3, 3), "parts": out[1][:, :count].long(),
                "centers": cen_mean}

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
