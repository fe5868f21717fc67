"""CPU emulation: error of the 3-term fp16 split with and without scaling each layer's weights by a power of two
before the split (exact; undone in fp32 after the accumulation).  Motivation: with small weights (|w| ~ 0.02) the `lo`
halves fall into the fp16 subnormal range.  Run: python tools/split_scale_emul.py"""
import os
import sys

sys.path.insert(0, os.getcwd())
import math

import torch

from oracle import chore_oracle as O


def q16(x):
    return x.clamp(-65504, 65504).to(torch.float16).to(torch.float64)


def mm3(A, W, scale):
    Ws = W * scale
    Ah, Wh = q16(A), q16(Ws)
    return (Ah @ Wh + q16(A - Ah) @ Wh + Ah @ q16(Ws - Wh)) / scale


def mlp(sd, head, X, scaled):
    x = X
    for li in (0, 2, 4, 6):
        W = sd[f"{head}.{li}.weight"][:, :, 0].double().t()
        s = 2.0 ** math.floor(math.log2(256.0 / W.abs().max().item())) if scaled else 1.0
        x = mm3(x, W, s) + sd[f"{head}.{li}.bias"].double()
        if li != 6:
            x = torch.relu(x).float().double()
    return x


def mlp64(sd, head, X):
    x = X
    for li in (0, 2, 4, 6):
        x = x @ sd[f"{head}.{li}.weight"][:, :, 0].double().t() + sd[f"{head}.{li}.bias"].double()
        if li != 6:
            x = torch.relu(x).float().double()
    return x


def rel(a, b):
    return ((a - b).abs() / (b.abs() + b.pow(2).mean().sqrt() + 1e-30)).max().item()


for kind in ("unit", "ref_init"):
    sd = O.make_state_dict(0, kind)
    if kind == "unit":
        feat, tmpx = O.synth_features(11, B=1)
    else:
        with torch.no_grad():
            feat, tmpx = O.encode(sd, O.synth_images(23, B=1, size=128))
    pts = O.synth_points("frustum", 5, 1, 20000)
    cc = torch.tensor([[1008., 995.]])
    xy = O.project_points(pts, cc)[:, :2, :]
    z_feat = torch.cat([pts[:, :, 0:2].transpose(1, 2), (pts[:, :, 2:3] - 2.2).transpose(1, 2)], 1)
    X = torch.cat([O.index(feat, xy), z_feat, O.index(tmpx, xy)], 1)[0].t().double()
    for head in ("df", "part_predictor", "pca_predictor", "center_predictor"):
        ref = mlp64(sd, head, X)
        print(f"{kind:9s} {head:17s} 3x fp16 split: {rel(mlp(sd, head, X, False), ref):.2e}   with per-layer 2^k weight scaling: {rel(mlp(sd, head, X, True), ref):.2e}")
