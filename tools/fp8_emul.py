"""Emulate the MLP with different operand splits (CPU, fp64 accumulation of exactly rounded operands)."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from oracle import chore_oracle as O
torch.manual_seed(0)

def q16(x): return x.clamp(-65504, 65504).to(torch.float16).to(torch.float64)
def q8(x, fmt, scale):
    dt = torch.float8_e4m3fn if fmt == "e4m3" else torch.float8_e5m2
    lim = 448.0 if fmt == "e4m3" else 57344.0
    return (x * scale).clamp(-lim, lim).to(torch.float32).to(dt).to(torch.float64) / scale

def split_mm(A, W, mode, cfg):
    """A (n,K) fp32-valued float64, W (K,M). returns A@W under the emulated arithmetic"""
    if mode == "fp64": return A @ W
    Ah, Wh = q16(A), q16(W)
    Al, Wl = A - Ah, W - Wh
    if mode == "2term": return Ah @ Wh
    if mode == "3x16": return Ah @ Wh + q16(Al) @ Wh + Ah @ q16(Wl)
    if mode == "fp8":
        fa_lo, sa_lo, fw_hi, fa_hi, sa_hi, fw_lo = cfg
        # term2: q8(Al * 2^s) x q8(Wh * 2^-s); term3: q8(Ah * 2^-t) x q8(Wl * 2^t)
        t2 = q8(Al, fa_lo, 2.0 ** sa_lo) @ q8(Wh, fw_hi, 2.0 ** -sa_lo)
        t3 = q8(Ah, fa_hi, 2.0 ** -sa_hi) @ q8(Wl, fw_lo, 2.0 ** sa_hi)
        return Ah @ Wh + t2 + t3
    raise ValueError

def mlp(sd, head, X, mode, cfg=None):
    x = X
    for li in (0, 2, 4, 6):
        W = sd[f"{head}.{li}.weight"][:, :, 0].double().t()
        b = sd[f"{head}.{li}.bias"].double()
        x = split_mm(x, W, mode, cfg) + b
        if li != 6:
            x = torch.relu(x).float().double()      # activations are fp32 between layers
    return x

def rel(a, b):
    return ((a - b).abs() / (b.abs() + b.pow(2).mean().sqrt() + 1e-30)).max().item()

for kind in ("unit", "ref_init"):
    sd = O.make_state_dict(0, kind)
    if kind == "unit":
        feat, tmpx = O.synth_features(11, B=1)
    else:
        img = O.synth_images(23, B=1, size=128)
        with torch.no_grad(): feat, tmpx = O.encode(sd, img)
    N = 20000
    pts = O.synth_points("frustum", 5, 1, N)
    cc = torch.tensor([[1008., 995.]])
    xyz = O.project_points(pts, cc); xy = xyz[:, :2, :]
    z_feat = torch.cat([pts[:, :, 0:2].transpose(1, 2), (pts[:, :, 2:3] - 2.2).transpose(1, 2)], 1)
    local = torch.cat([O.index(feat, xy), z_feat, O.index(tmpx, xy)], 1)[0].t().double()   # (N,323)
    print(kind, "feature rms", local.pow(2).mean().sqrt().item(), "max", local.abs().max().item())
    cfgs = {
        "A_lo e4m3*2^8 x W_hi e5m2 | A_hi e5m2*2^-4 x W_lo e4m3": ("e4m3", 8, "e5m2", "e5m2", 4, "e4m3"),
        "all e5m2 (s=6, t=6)": ("e5m2", 6, "e5m2", "e5m2", 6, "e5m2"),
        "A_lo e4m3*2^9 x W_hi e5m2 | A_hi e4m3*2^-2 x W_lo e5m2*2^2": ("e4m3", 9, "e5m2", "e4m3", 2, "e5m2"),
        "A_lo e4m3*2^10 x W_hi e5m2 | A_hi e4m3*2^-3 x W_lo e5m2*2^3": ("e4m3", 10, "e5m2", "e4m3", 3, "e5m2"),
    }
    for head in ("df", "part_predictor", "pca_predictor", "center_predictor"):
        ref = mlp(sd, head, local, "fp64")
        f32 = torch.from_numpy(np.ascontiguousarray(O._mlp(sd, head, local.float().t().unsqueeze(0))[0].t().numpy())).double()
        print(f"  {head:18s} fp32-torch {rel(f32, ref):.2e}  2term {rel(mlp(sd, head, local, '2term'), ref):.2e}  3x16 {rel(mlp(sd, head, local, '3x16'), ref):.2e}")
        for name, cfg in cfgs.items():
            print(f"      fp8[{name}] {rel(mlp(sd, head, local, 'fp8', cfg), ref):.2e}")
