"""CPU restatement of the CHORE hot path (encoder, point query, SMPL-H LBS, rigid
object transform, fit-step losses).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The reference is Python/PyTorch, so the oracle is a *functional* fp32 restatement on
torch CPU ops driven directly by a ``state_dict`` (no nn.Module), op for op in the
reference's order so that (a) it reproduces the reference's CPU numbers to rounding and
(b) timing it is a fair stand-in for the reference CPU path (bench.py cpu_baseline
kind "port").  A second, independent restatement of the point query in plain numpy
(`query_numpy`) guards against a shared misunderstanding of the ATen semantics.

Pinned by tests/golden/*.npz, which oracle/make_golden.py generates by importing the
*real* reference from /root/reference in the build container (tests/test_oracle.py).
The reference itself ships no golden vectors for this path.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------
# constants of the chore-release configuration (config/chore-release.json)
# ----------------------------------------------------------------------------------------
Z0 = 2.2                      # "z_0"
CROP_SIZE = 1200              # "loadSize"
OUT_DIST = 5.0                # model/chore.py:65
FX_PX, FY_PX = 979.7844 / 2048. * 2048, 979.840 / 2048. * 2048      # model/camera.py:26-38
CX_PX, CY_PX = 1018.952 / 2048. * 2048, 779.486 / 2048. * 2048
HEADS = (("df", 2), ("pca_predictor", 9), ("part_predictor", 14), ("center_predictor", 6))
NUM_PARTS = 14
SMPLH_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19,
                 20, 22, 23, 20, 25, 26, 20, 28, 29, 20, 31, 32, 20, 34, 35,
                 21, 37, 38, 21, 40, 41, 21, 43, 44, 21, 46, 47, 21, 49, 50]
assert len(SMPLH_PARENTS) == 52


# ----------------------------------------------------------------------------------------
# weights: key/shape census of CHORE(chore-release).state_dict() and seeded synthesis
# ----------------------------------------------------------------------------------------
def _convblock_spec(prefix: str, cin: int, cout: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Registration order of ConvBlock parameters (model/net_util.py:347-371)."""
    s = [(f"{prefix}.conv1.weight", (cout // 2, cin, 3, 3)),
         (f"{prefix}.conv2.weight", (cout // 4, cout // 2, 3, 3)),
         (f"{prefix}.conv3.weight", (cout // 4, cout // 4, 3, 3))]
    for name, c in (("bn1", cin), ("bn2", cout // 2), ("bn3", cout // 4), ("bn4", cin)):
        s += [(f"{prefix}.{name}.weight", (c,)), (f"{prefix}.{name}.bias", (c,))]
    if cin != cout:
        s += [(f"{prefix}.downsample.0.weight", (cin,)), (f"{prefix}.downsample.0.bias", (cin,)),
              (f"{prefix}.downsample.2.weight", (cout, cin, 1, 1))]
    return s


def _hourglass_spec(prefix: str, level: int, c: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """HourGlass._generate_network registration order (model/HGFilters.py:14-24)."""
    s = _convblock_spec(f"{prefix}.b1_{level}", c, c) + _convblock_spec(f"{prefix}.b2_{level}", c, c)
    if level > 1:
        s += _hourglass_spec(prefix, level - 1, c)
    else:
        s += _convblock_spec(f"{prefix}.b2_plus_{level}", c, c)
    s += _convblock_spec(f"{prefix}.b3_{level}", c, c)
    return s


def weight_spec(num_stack: int = 5, depth: int = 2, in_ch: int = 5, hidden: int = 128
                ) -> List[Tuple[str, Tuple[int, ...]]]:
    """Ordered (key, shape) list equal to CHORE(chore-release).state_dict()
    (model/HGFilters.py:100-142, model/chore.py:44-55,74-85)."""
    p = "image_filter"
    s = [(f"{p}.conv1.weight", (64, in_ch, 7, 7)), (f"{p}.conv1.bias", (64,)),
         (f"{p}.bn1.weight", (64,)), (f"{p}.bn1.bias", (64,))]
    s += _convblock_spec(f"{p}.conv2", 64, 128)
    s += _convblock_spec(f"{p}.conv3", 128, 128)
    s += _convblock_spec(f"{p}.conv4", 128, 256)
    for i in range(num_stack):
        s += _hourglass_spec(f"{p}.m{i}", depth, 256)
        s += _convblock_spec(f"{p}.top_m_{i}", 256, 256)
        s += [(f"{p}.conv_last{i}.weight", (256, 256, 1, 1)), (f"{p}.conv_last{i}.bias", (256,)),
              (f"{p}.bn_end{i}.weight", (256,)), (f"{p}.bn_end{i}.bias", (256,)),
              (f"{p}.l{i}.weight", (256, 256, 1, 1)), (f"{p}.l{i}.bias", (256,))]
        if i < num_stack - 1:
            s += [(f"{p}.bl{i}.weight", (256, 256, 1, 1)), (f"{p}.bl{i}.bias", (256,)),
                  (f"{p}.al{i}.weight", (256, 256, 1, 1)), (f"{p}.al{i}.bias", (256,))]
    feat = 256 + 3 + 64
    for head, out in HEADS_IN_STATE_ORDER:
        dims = [(hidden, feat), (hidden, hidden), (hidden, hidden), (out, hidden)]
        for li, (o, i_) in zip((0, 2, 4, 6), dims):
            s += [(f"{head}.{li}.weight", (o, i_, 1)), (f"{head}.{li}.bias", (o,))]
    return s


# registration order in CHORE.__init__ (model/chore.py:48-55): df, part, pca, center
HEADS_IN_STATE_ORDER = (("df", 2), ("part_predictor", 14), ("pca_predictor", 9), ("center_predictor", 6))


def make_state_dict(seed: int = 0, kind: str = "unit", **spec_kw) -> Dict[str, Tensor]:
    """Seeded synthetic weights.

    kind="unit":  conv weights ~ N(0, 2/fan_in), biases ~ N(0, 0.1), GroupNorm affine
                  ~ N(1, 0.1) / N(0, 0.1): activations and logits are O(1) (the
                  parity-discriminating set).
    kind="ref_init": the reference's init_weights: conv N(0, 0.02), zero bias, GN affine
                  untouched (1, 0) (model/net_util.py:218-251 -- only BatchNorm2d is matched).
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for key, shape in weight_spec(**spec_kw):
        is_norm = (".bn" in key) or (".downsample.0." in key)
        if key.endswith(".downsample.0.weight") or key.endswith(".downsample.0.bias"):
            # nn.Sequential(self.bn4, ...) aliases bn4 (model/net_util.py:363-369)
            sd[key] = sd[key.replace(".downsample.0.", ".bn4.")]
            continue
        if kind == "unit":
            if is_norm:
                base = 1.0 if key.endswith("weight") else 0.0
                t = base + 0.1 * torch.randn(shape, generator=g)
            elif key.endswith("bias"):
                t = 0.1 * torch.randn(shape, generator=g)
            else:
                fan_in = int(np.prod(shape[1:]))
                t = math.sqrt(2.0 / fan_in) * torch.randn(shape, generator=g)
        elif kind == "ref_init":
            if is_norm:
                t = torch.ones(shape) if key.endswith("weight") else torch.zeros(shape)
            elif key.endswith("bias"):
                t = torch.zeros(shape)
            else:
                t = 0.02 * torch.randn(shape, generator=g)
        else:
            raise ValueError(kind)
        sd[key] = t.float().contiguous()
    return sd


# ----------------------------------------------------------------------------------------
# seeded synthetic inputs; torch CPU generator => same bits everywhere
# ----------------------------------------------------------------------------------------
def synth_images(seed: int, B: int = 1, size: int = 512) -> Tensor:
    """(B,5,size,size) in [0,1]: RGB*mask + person mask + object mask stand-in (data/base_data.py:179-192)."""
    return torch.rand(B, 5, size, size, generator=torch.Generator().manual_seed(seed))


def synth_features(seed: int, B: int = 1, hw: int = 128) -> Tuple[Tensor, Tensor]:
    """Stand-ins for (last hourglass output (B,256,hw,hw), post-ReLU stem tmpx (B,64,2hw,2hw))."""
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(B, 256, hw, hw, generator=g)
    tmpx = torch.relu(torch.randn(B, 64, 2 * hw, 2 * hw, generator=g))
    return feat, tmpx


def synth_points(kind: str, seed: int, B: int, N: int, crop_center: Tensor | None = None) -> Tensor:
    """"init_box": Generator.init_samples box x[-3,3] y[-2.5,2.5] z[1.95,2.45] (recon/generator.py:275-282,
    applied to every batch element); "frustum": uniform in the crop's normalised image square,
    back-projected at z in [1.95,2.45] (all points inside the image)."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(B, N, 3, generator=g)
    z = (u[..., 2] - 0.5) * 0.5 + Z0
    if kind == "init_box":
        return torch.stack([u[..., 0] * 6 - 3, u[..., 1] * 5 - 2.5, z], -1).contiguous()
    if kind == "frustum":
        cc = torch.tensor([[1008., 995.]]).repeat(B, 1) if crop_center is None else crop_center
        nx, ny = u[..., 0] * 1.98 - 0.99, u[..., 1] * 1.98 - 0.99
        px = (nx + 1) * CROP_SIZE / 2 - CROP_SIZE / 2 + cc[:, 0:1]
        py = (ny + 1) * CROP_SIZE / 2 - CROP_SIZE / 2 + cc[:, 1:2]
        return torch.stack([(px - CX_PX) * z / FX_PX, (py - CY_PX) * z / FY_PX, z], -1).float().contiguous()
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------
# encoder: stacked hourglass (model/HGFilters.py, model/net_util.py:346-396)
# ----------------------------------------------------------------------------------------
def _gn_relu(sd, key: str, x: Tensor) -> Tensor:
    return F.relu(F.group_norm(x, 32, sd[f"{key}.weight"], sd[f"{key}.bias"], 1e-5))


def conv_block(sd, p: str, x: Tensor) -> Tensor:
    """ConvBlock.forward (model/net_util.py:374-396)."""
    o1 = F.conv2d(_gn_relu(sd, f"{p}.bn1", x), sd[f"{p}.conv1.weight"], padding=1)
    o2 = F.conv2d(_gn_relu(sd, f"{p}.bn2", o1), sd[f"{p}.conv2.weight"], padding=1)
    o3 = F.conv2d(_gn_relu(sd, f"{p}.bn3", o2), sd[f"{p}.conv3.weight"], padding=1)
    out = torch.cat((o1, o2, o3), 1)
    if f"{p}.downsample.2.weight" in sd:
        res = F.conv2d(_gn_relu(sd, f"{p}.bn4", x), sd[f"{p}.downsample.2.weight"])
    else:
        res = x
    return out + res


def hourglass(sd, p: str, level: int, x: Tensor) -> Tensor:
    """HourGlass._forward (model/HGFilters.py:26-50)."""
    up1 = conv_block(sd, f"{p}.b1_{level}", x)
    low = conv_block(sd, f"{p}.b2_{level}", F.avg_pool2d(x, 2, stride=2))
    if level > 1:
        low = hourglass(sd, p, level - 1, low)
    else:
        low = conv_block(sd, f"{p}.b2_plus_{level}", low)
    low = conv_block(sd, f"{p}.b3_{level}", low)
    up2 = F.interpolate(low, scale_factor=2, mode="bicubic", align_corners=True)
    return up1 + up2


def hg_filter(sd, images: Tensor, num_stack: int = 5, depth: int = 2
              ) -> Tuple[List[Tensor], Tensor, Tensor]:
    """HGFilter.forward (model/HGFilters.py:144-185): returns (outputs[num_stack], tmpx, normx)."""
    p = "image_filter"
    x = F.conv2d(images, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], stride=2, padding=3)
    x = _gn_relu(sd, f"{p}.bn1", x)
    tmpx = x
    x = F.avg_pool2d(conv_block(sd, f"{p}.conv2", x), 2, stride=2)
    normx = x
    x = conv_block(sd, f"{p}.conv3", x)
    x = conv_block(sd, f"{p}.conv4", x)
    previous = x
    outputs = []
    for i in range(num_stack):
        ll = hourglass(sd, f"{p}.m{i}", depth, previous)
        ll = conv_block(sd, f"{p}.top_m_{i}", ll)
        ll = F.conv2d(ll, sd[f"{p}.conv_last{i}.weight"], sd[f"{p}.conv_last{i}.bias"])
        ll = _gn_relu(sd, f"{p}.bn_end{i}", ll)
        out = F.conv2d(ll, sd[f"{p}.l{i}.weight"], sd[f"{p}.l{i}.bias"])
        outputs.append(out)
        if i < num_stack - 1:
            ll = F.conv2d(ll, sd[f"{p}.bl{i}.weight"], sd[f"{p}.bl{i}.bias"])
            back = F.conv2d(out, sd[f"{p}.al{i}.weight"], sd[f"{p}.al{i}.bias"])
            previous = previous + ll + back
    return outputs, tmpx.detach(), normx


def encode(sd, images: Tensor) -> Tuple[Tensor, Tensor]:
    """CHORE.filter in eval mode (model/chore.py:87-96): keeps the last stack output + tmpx."""
    outs, tmpx, _ = hg_filter(sd, images)
    return outs[-1], tmpx


# ----------------------------------------------------------------------------------------
# point query (model/camera.py:44-88, model/geometry.py:4-14, model/chore.py:107-167)
# ----------------------------------------------------------------------------------------
def project_points(points: Tensor, crop_center: Tensor) -> Tensor:
    """KinectColorCamera.project_points: (B,N,3),(B,2) -> (B,3,N); fp32 op order of
    model/camera.py:64-65,75-78 kept exactly (python-double constants times fp32 tensors)."""
    x, y, z = points[:, :, 0:1], points[:, :, 1:2], points[:, :, 2:3]
    px = FX_PX * x / z + CX_PX
    py = FY_PX * y / z + CY_PX
    px = CROP_SIZE / 2 + px - crop_center[:, 0].unsqueeze(1).unsqueeze(1)
    py = CROP_SIZE / 2 + py - crop_center[:, 1].unsqueeze(1).unsqueeze(1)
    nx = 2 * px / CROP_SIZE - 1
    ny = 2 * py / CROP_SIZE - 1
    return torch.cat([nx, ny, z], -1).transpose(1, 2)


def index(feat: Tensor, uv: Tensor) -> Tensor:
    """model/geometry.py:4-14: bilinear, zeros padding, align_corners=True."""
    grid = uv.transpose(1, 2).unsqueeze(2)
    return F.grid_sample(feat, grid, mode="bilinear", padding_mode="zeros", align_corners=True)[:, :, :, 0]


def _mlp(sd, head: str, x: Tensor) -> Tensor:
    """make_decoder (model/chore.py:74-85): conv1d(k=1) x4 with ReLU between."""
    for li in (0, 2, 4):
        x = F.relu(F.conv1d(x, sd[f"{head}.{li}.weight"], sd[f"{head}.{li}.bias"]))
    return F.conv1d(x, sd[f"{head}.6.weight"], sd[f"{head}.6.bias"])


def query(sd, feat: Tensor, tmpx: Tensor, points: Tensor, crop_center: Tensor):
    """CHORE.query + decode (model/chore.py:107-167).

    Returns (df (B,2,N), pca (B,3,3,N), parts (B,14,N), centers (B,6,N), in_img (B,N) bool).
    Channel order of the 323-vector: [feat 256 | x, y, z-2.2 | tmpx 64] (chore.py:139-143).
    """
    xyz = project_points(points, crop_center)
    xy = xyz[:, :2, :]
    rela_z = (points[:, :, 2:3] - Z0).transpose(1, 2)
    z_feat = torch.cat([points[:, :, 0:2].transpose(1, 2), rela_z], 1)
    in_img = (xy[:, 0] >= -1.0) & (xy[:, 0] <= 1.0) & (xy[:, 1] >= -1.0) & (xy[:, 1] <= 1.0)
    local = torch.cat([index(feat, xy), z_feat, index(tmpx, xy)], 1)
    df = _mlp(sd, "df", local)
    pca = _mlp(sd, "pca_predictor", local).view(df.shape[0], 3, 3, -1)
    parts = _mlp(sd, "part_predictor", local)
    centers = _mlp(sd, "center_predictor", local)
    df_t = df.transpose(1, 2)            # view: the masked write lands in df (chore.py:147-150)
    df_t[~in_img] = OUT_DIST
    return df, pca, parts, centers, in_img


def query_numpy(sd, feat: np.ndarray, tmpx: np.ndarray, points: np.ndarray, crop_center: np.ndarray):
    """Independent plain-numpy restatement of the query for small N (no torch ops):
    projection (camera.py:64-78), grid_sample semantics (geometry.py:12; ATen
    grid_sampler_2d bilinear/zeros/align_corners=True), 4-head MLP (chore.py:74-85,156-167)."""
    f32 = np.float32
    B, N, _ = points.shape
    w = {k: v.numpy() for k, v in sd.items() if not k.startswith("image_filter")}
    out = {h: np.zeros((B, o, N), f32) for h, o in HEADS}
    in_img = np.zeros((B, N), bool)

    def sample(fm, gx, gy):
        C, H, W = fm.shape
        ix = (gx + f32(1)) / f32(2) * f32(W - 1)
        iy = (gy + f32(1)) / f32(2) * f32(H - 1)
        x0, y0 = int(np.floor(ix)), int(np.floor(iy))
        acc = np.zeros(C, f32)
        for yy, xx in ((y0, x0), (y0, x0 + 1), (y0 + 1, x0), (y0 + 1, x0 + 1)):
            wgt = (f32(1) - abs(ix - f32(xx))) * (f32(1) - abs(iy - f32(yy)))
            if 0 <= xx < W and 0 <= yy < H:
                acc += fm[:, yy, xx] * f32(wgt)
        return acc

    for b in range(B):
        for n in range(N):
            x, y, z = (f32(v) for v in points[b, n])
            px = f32(FX_PX) * x / z + f32(CX_PX)
            py = f32(FY_PX) * y / z + f32(CY_PX)
            px = f32(CROP_SIZE / 2) + px - f32(crop_center[b, 0])
            py = f32(CROP_SIZE / 2) + py - f32(crop_center[b, 1])
            nx = f32(2) * px / f32(CROP_SIZE) - f32(1)
            ny = f32(2) * py / f32(CROP_SIZE) - f32(1)
            inside = (-1.0 <= nx <= 1.0) and (-1.0 <= ny <= 1.0)
            in_img[b, n] = inside
            if np.isfinite(nx) and np.isfinite(ny):
                fa, fb = sample(feat[b], nx, ny), sample(tmpx[b], nx, ny)
            else:
                fa, fb = np.zeros(feat.shape[1], f32), np.zeros(tmpx.shape[1], f32)
            v = np.concatenate([fa, np.array([x, y, z - f32(Z0)], f32), fb]).astype(f32)
            for h, _ in HEADS:
                a = v
                for li in (0, 2, 4):
                    a = np.maximum(w[f"{h}.{li}.weight"][:, :, 0] @ a + w[f"{h}.{li}.bias"], 0).astype(f32)
                out[h][b, :, n] = w[f"{h}.6.weight"][:, :, 0] @ a + w[f"{h}.6.bias"]
            if not inside:
                out["df"][b, :, n] = OUT_DIST
    return out["df"], out["pca_predictor"].reshape(B, 3, 3, N), out["part_predictor"], out["center_predictor"], in_img


def query_grad_points(sd, feat, tmpx, points, crop_center, g_df=None, g_pca=None, g_parts=None, g_centers=None):
    """d(sum of <g_k, head_k>)/d points through torch autograd: the gradient the reference's
    callers obtain with loss.backward() (recon/generator.py:63-70, recon_fit_behave.py:149-152)."""
    pts = points.detach().clone().requires_grad_(True)
    df, pca, parts, centers, _ = query(sd, feat, tmpx, pts, crop_center)
    tot = 0.0
    for g, o in ((g_df, df), (g_pca, pca), (g_parts, parts), (g_centers, centers)):
        if g is not None:
            tot = tot + (g.reshape(o.shape) * o).sum()
    tot.backward()
    return pts.grad.detach()


# ----------------------------------------------------------------------------------------
# neural surface projection (recon/generator.py:50-79)
# ----------------------------------------------------------------------------------------
def approx_surface(sd, feat, tmpx, samples: Tensor, crop_center: Tensor, num_steps: int,
                   df_idx: int, threshold: float = 2.0):
    """Generator.approx_surface: num_steps x { query; t = clamp(df_k, max=thr);
    p <- p - normalize(dt.sum()/dp) * t }.  Returns (samples, last preds)."""
    preds = None
    samples = samples.detach().clone().requires_grad_(True)
    for _ in range(num_steps):
        preds = query(sd, feat, tmpx, samples, crop_center)
        tgt = torch.clamp(preds[0][:, df_idx, :], max=threshold)
        tgt.sum().backward()
        grad = samples.grad.detach()
        samples = (samples.detach() - F.normalize(grad, dim=2) * tgt.detach().unsqueeze(-1))
        samples = samples.detach().requires_grad_(True)
    return samples.detach(), tuple(p.detach() for p in preds)


def create_grid(res: Sequence[int], b_min, b_max) -> np.ndarray:
    """model/sdf.py:4-27 coordinates only: (3, X*Y*Z) float64, x-major (np.mgrid order),
    coord = b_min + (b_max-b_min)/res * idx."""
    rx, ry, rz = res
    idx = np.mgrid[:rx, :ry, :rz].reshape(3, -1).astype(np.float64)
    length = np.asarray(b_max, np.float64) - np.asarray(b_min, np.float64)
    m = np.diag(length / np.array([rx, ry, rz], np.float64))
    return m @ idx + np.asarray(b_min, np.float64)[:, None]


# ----------------------------------------------------------------------------------------
# SMPL-H linear blend skinning (lib_smpl/smplpytorch/smplpytorch/pytorch/*.py)
# ----------------------------------------------------------------------------------------
def make_smplh_buffers(seed: int = 0, n_verts: int = 6890, n_joints: int = 52, n_betas: int = 10
                       ) -> Dict[str, Tensor]:
    """Synthetic SMPL-H-shaped model buffers (the licensed pickle is absent; shapes from
    smpl_layer.py:49-64).  Skinning weights: 4 non-zeros per vertex summing to 1;
    joint regressor: 32 non-zeros per joint summing to 1 (stored dense like th_J_regressor)."""
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(n_verts, 3, generator=g) * torch.tensor([0.25, 0.45, 0.12])
    shapedirs = 0.02 * torch.randn(n_verts, 3, n_betas, generator=g)
    posedirs = 0.005 * torch.randn(n_verts, 3, (n_joints - 1) * 9, generator=g)
    jreg = torch.zeros(n_joints, n_verts)
    for j in range(n_joints):
        idx = torch.randint(n_verts, (32,), generator=g)
        wj = torch.rand(32, generator=g) + 0.05
        jreg[j].index_add_(0, idx, wj / wj.sum())
    weights = torch.zeros(n_verts, n_joints)
    idx = torch.randint(n_joints, (n_verts, 4), generator=g)
    wv = torch.rand(n_verts, 4, generator=g) + 0.05
    wv = wv / wv.sum(1, keepdim=True)
    weights.scatter_add_(1, idx, wv)
    faces = torch.randint(n_verts, (13776, 3), generator=g)
    return {"v_template": v.unsqueeze(0).contiguous(), "shapedirs": shapedirs, "posedirs": posedirs,
            "J_regressor": jreg, "weights": weights, "faces": faces,
            "parents": torch.tensor(SMPLH_PARENTS, dtype=torch.int64)}


def batch_rodrigues(axisang: Tensor) -> Tensor:
    """rodrigues_layer.py:13-52: theta = ||r + 1e-8||, quaternion (cos t/2, sin t/2 * r/theta),
    re-normalised, to a flattened 3x3."""
    ang = torch.norm(axisang + 1e-8, p=2, dim=1).unsqueeze(-1)
    n = axisang / ang
    half = ang * 0.5
    q = torch.cat([torch.cos(half), torch.sin(half) * n], 1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], 1)


def _with_zeros(m34: Tensor) -> Tensor:
    pad = m34.new_tensor([0., 0., 0., 1.]).view(1, 1, 4).repeat(m34.shape[0], 1, 1)
    return torch.cat([m34, pad], 1)


def lbs_forward(model: Dict[str, Tensor], pose: Tensor, betas: Tensor, trans: Tensor,
                offsets: Tensor | None = None):
    """SMPL_Layer.forward (smpl_layer.py:72-175) for hands=True, scale=1, betas given.
    Returns (verts (B,V,3), jtr (B,J,3), v_posed, naked).  Same loop structure as the
    reference (per-joint Rodrigues, sequential chain), so CPU timing is representative."""
    B = pose.shape[0]
    J = model["weights"].shape[1]
    parents = [int(p) for p in model["parents"]]
    rots = torch.cat([batch_rodrigues(pose[:, 3 * j:3 * j + 3]) for j in range(J)], 1)
    root = rots[:, :9].view(B, 3, 3)
    rest = rots[:, 9:]
    pose_map = rest - torch.eye(3, dtype=rest.dtype).view(1, 9).repeat(B, J - 1)
    v_shaped = model["v_template"] + torch.matmul(model["shapedirs"], betas.transpose(1, 0)).permute(2, 0, 1)
    jnt = torch.matmul(model["J_regressor"], v_shaped)
    naked = v_shaped + torch.matmul(model["posedirs"], pose_map.transpose(0, 1)).permute(2, 0, 1)
    v_posed = naked + offsets if offsets is not None else naked
    chain = [_with_zeros(torch.cat([root, jnt[:, 0, :].reshape(B, 3, 1)], 2))]
    for i in range(1, J):
        r = rest[:, (i - 1) * 9:i * 9].reshape(B, 3, 3)
        rel = _with_zeros(torch.cat([r, (jnt[:, i, :] - jnt[:, parents[i], :]).reshape(B, 3, 1)], 2))
        chain.append(torch.matmul(chain[parents[i]], rel))
    A = jnt.new_zeros((B, 4, 4, J))
    for i in range(J):
        jh = torch.cat([jnt[:, i], jnt.new_zeros(B, 1)], 1)
        t = torch.bmm(chain[i], jh.unsqueeze(2))
        A[:, :, :, i] = chain[i] - torch.cat([t.new_zeros(B, 4, 3), t], 2)
    T = torch.matmul(A, model["weights"].transpose(0, 1))
    vh = torch.cat([v_posed.transpose(2, 1), T.new_ones((B, 1, v_posed.shape[1]))], 1)
    verts = (T * vh.unsqueeze(1)).sum(2).transpose(2, 1)[:, :, :3]
    jtr = torch.stack(chain, dim=1)[:, :, :3, 3]
    return verts + trans.unsqueeze(1), jtr + trans.unsqueeze(1), v_posed, naked


# ----------------------------------------------------------------------------------------
# rigid object transform + SO(3) projection (recon/recon_fit_base.py:167-188,361-384)
# ----------------------------------------------------------------------------------------
def project_so3(mat: Tensor) -> Tensor:
    """R = U diag(1,1,det(U V^T)) V^T (recon_fit_base.py:167-188)."""
    u, _, v = torch.svd(mat)
    vt = v.transpose(1, 2)
    det = torch.det(torch.matmul(u, vt)).view(-1, 1, 1)
    vt = torch.cat((vt[:, :2, :], vt[:, -1:, :] * det), 1)
    return torch.matmul(u, vt)


def decopose_axis(rot: Tensor, noise: Tensor | None = None) -> Tensor:
    """recon_fit_base.py:373-384 with the 1e-4*rand perturbation passed in explicitly
    (None == no_rand=True) so that runs are reproducible."""
    return project_so3(rot if noise is None else rot + 1e-4 * noise)


def transform_obj_verts(verts: Tensor, R: Tensor, t: Tensor, s: Tensor) -> Tensor:
    """(verts @ R + t) * s, row-vector convention, scale last (recon_fit_base.py:367-371)."""
    return (torch.bmm(verts, R) + t.unsqueeze(1)) * s.unsqueeze(1).unsqueeze(1)


# ----------------------------------------------------------------------------------------
# fit-step losses (recon/recon_fit_behave.py:165-222,293-358; recon_fit_base.py:351-359,513-551)
# ----------------------------------------------------------------------------------------
LOSS_WEIGHTS = {"beta": 1.0, "pose": 1e-5, "hand": 1e-5, "j2d": 0.3 ** 2, "object": 30.0 ** 2,
                "part": 0.05 ** 2, "contact": 30.0 ** 2, "scale": 10.0 ** 2, "df_h": 30.0 ** 2,
                "smplz": 30 ** 2, "mask": 0.003 ** 2, "ocent": 15 ** 2, "collide": 3 ** 2,
                "pinit": 5 ** 2, "rot": 10.0 ** 2, "trans": 10.0 ** 2}   # recon_fit_behave.py:339-358


def sum_dict(loss_dict: Dict[str, Tensor], decay: float) -> Tensor:
    """recon_fit_base.py:351-359 with get_loss_weights: w_k * l_k / (1 + decay), summed."""
    return torch.stack([LOSS_WEIGHTS[k] * v / (1 + decay) for k, v in loss_dict.items()]).sum()


def object_only_losses(sd, feat, tmpx, crop_center, obj_pts: Tensor, R: Tensor, t: Tensor, s: Tensor,
                       smpl_center: Tensor, obj_scale: float = 1.0) -> Dict[str, Tensor]:
    """forward_step(phase='object only') (recon_fit_behave.py:165-198 + recon_fit_base.py:513-520).
    R is the already SO(3)-projected rotation."""
    obj = transform_obj_verts(obj_pts, R, t, s)
    _, _, _, centers, _ = query(sd, feat, tmpx, obj, crop_center)
    center_pred = smpl_center + torch.mean(centers[:, 3:, :], -1)
    df, _, _, _, _ = query(sd, feat, tmpx, obj, crop_center)      # queried twice in the reference
    return {"object": torch.clamp(df[:, 1:2, :], max=0.8).mean(),
            "scale": torch.mean((s - obj_scale) ** 2),
            "ocent": F.mse_loss(torch.mean(obj, 1), center_pred, reduction="none").sum(-1).mean()}


def smpl_losses(sd, feat, tmpx, crop_center, verts: Tensor, part_labels: Tensor) -> Dict[str, Tensor]:
    """Field terms of forward_smpl (recon_fit_behave.py:293-313 + recon_fit_base.py:537-542):
    df_h = mean(min(df_h, 0.1)); part = CE(parts, labels).sum(-1).mean()."""
    df, _, parts, _, _ = query(sd, feat, tmpx, verts, crop_center)
    return {"df_h": torch.clamp(df[:, 0:1, :], max=0.1).mean(),
            "part": F.cross_entropy(parts, part_labels, reduction="none").sum(-1).mean()}


# ----------------------------------------------------------------------------------------
# joint-phase contact term (recon/recon_fit_base.py:553-608)
# ----------------------------------------------------------------------------------------
def chamfer_distance_lists(xs: Sequence[Tensor], ys: Sequence[Tensor]) -> Tensor:
    """pytorch3d.loss.chamfer_distance(Pointclouds(xs), Pointclouds(ys)) with its default arguments (point_reduction='mean',
    batch_reduction='mean', norm=2): per pair mean_p min_q |x_p - y_q|^2 + mean_q min_p |x_p - y_q|^2, averaged over the pairs.
    pytorch3d is not vendored by the reference (requirements.txt:22, unpinned git HEAD): this restates the documented default
    (pytorch3d/loss/chamfer.py, releases 0.4 - 0.7) -- parity unpinned for this term."""
    total = 0.0
    for x, y in zip(xs, ys):
        d = ((x[:, None, :] - y[None, :, :]) ** 2).sum(-1)
        total = total + d.min(1).values.mean() + d.min(0).values.mean()
    return total / len(xs)


def contact_loss(df_hum_o: Tensor, df_obj_h: Tensor, obj: Tensor, smpl_verts: Tensor, part_o: Tensor, part_labels: Tensor,
                 thresh: float = 0.08):
    """ReconFitterBase.compute_contact_loss (recon_fit_base.py:553-608): None when no contact is found anywhere."""
    pts_h, pts_o = [], []
    po_all = torch.argmax(part_o, 1)
    for hum, ob, mh, mo, po in zip(smpl_verts, obj, df_hum_o < thresh, df_obj_h < thresh, po_all):
        ch, co = int(mh.sum()), int(mo.sum())
        if ch + co == 0:
            continue
        obj_v, label_o = (ob[mo], po[mo]) if co > 0 else (ob, po)
        hum_v, label_h = (hum[mh], part_labels[mh]) if ch > 0 else (hum, part_labels)
        for i in range(14):
            if not bool((label_h == i).any()) or not bool((label_o == i).any()):
                continue
            pts_h.append(hum_v[label_h == i])
            pts_o.append(obj_v[label_o == i])
    if not pts_o:
        return None
    return chamfer_distance_lists(pts_h, pts_o)


# ----------------------------------------------------------------------------------------
# full SMPL-phase step (recon/recon_fit_behave.py:293-337 with recon_fit_base.py:230-231,522-542,653-676)
# ----------------------------------------------------------------------------------------
def mahalanobis(pose: Tensor, mean: Tensor, prec: Tensor, prefix: int = 3, end: int = 66) -> Tensor:
    """th_Mahalanobis.__call__ (lib_smpl/th_smpl_prior.py:32-39)."""
    t = torch.matmul(pose[:, prefix:end] - mean.view(1, -1), prec)
    return (t * t).sum(dim=1)


def hand_prior(pose: Tensor, mean: Tensor, lprec: Tensor, rprec: Tensor, prefix: int = 66) -> Tensor:
    """HandPrior.__call__ (lib_smpl/th_hand_prior.py:69-78), including its shapes: the precisions carry a
    leading 1, so the result is (1,45), summed over batch rows and both hands."""
    temp = pose[:, prefix:] - mean.view(1, -1)
    lh = torch.matmul(temp[:, :45], lprec.unsqueeze(0))
    rh = torch.matmul(temp[:, 45:], rprec.unsqueeze(0))
    t2 = torch.cat([lh, rh], 1)
    return (t2 * t2).sum(dim=1)


def project_to_input_image(joints3d: Tensor, crop_center: Tensor, net_in_size: int = 512) -> Tensor:
    """ReconFitterBase.project_points (recon_fit_base.py:661-670) over KinectColorCamera.project_screen
    (model/camera.py:51-71)."""
    x, y, z = joints3d[:, :, 0:1], joints3d[:, :, 1:2], joints3d[:, :, 2:3]
    px = FX_PX * x / z + CX_PX
    py = FY_PX * y / z + CY_PX
    px = CROP_SIZE / 2 + px - crop_center[:, 0].unsqueeze(1).unsqueeze(1)
    py = CROP_SIZE / 2 + py - crop_center[:, 1].unsqueeze(1).unsqueeze(1)
    return torch.cat([px, py], -1) * net_in_size / CROP_SIZE


def smpl_full_losses(sd, feat, tmpx, crop_center, model: Dict[str, Tensor], pose: Tensor, betas: Tensor, trans: Tensor,
                     part_labels: Tensor, pose_init: Tensor, regressors: Sequence[Tensor], priors: Dict[str, Tensor],
                     body_kpts: Tensor | None = None, offsets: Tensor | None = None) -> Dict[str, Tensor]:
    """forward_smpl, every term, in the reference's insertion order (sum_dict only sums, but keep it).
    regressors: dense (L,V) body25 / face / hand; priors: body_mean, body_prec, hand_mean, lh_prec, rh_prec.
    body_kpts given == phase 'kpts'."""
    verts = lbs_forward(model, pose, betas, trans, offsets)[0]
    df, _, parts, _, _ = query(sd, feat, tmpx, verts, crop_center)
    out = {"df_h": torch.clamp(df[:, 0:1, :], max=0.1).mean(),
           "pose": torch.mean(mahalanobis(pose[:, :72], priors["body_mean"], priors["body_prec"])),
           "hand": torch.mean(hand_prior(pose, priors["hand_mean"], priors["lh_prec"], priors["rh_prec"])),
           "part": F.cross_entropy(parts, part_labels, reduction="none").sum(-1).mean()}
    J = torch.matmul(regressors[0], verts)
    out["smplz"] = torch.mean((J[:, 8, 2] - Z0) ** 2)
    out["pinit"] = torch.mean(torch.sum((pose[:, 3:72] - pose_init) ** 2, -1))
    if body_kpts is not None:
        proj = project_to_input_image(J, crop_center)
        l2 = F.mse_loss(proj[:, :, :2], body_kpts[:, :, :2], reduction="none")
        out["j2d"] = torch.mean(torch.sum(l2, dim=-1) * body_kpts[:, :, 2])
    return out


# ----------------------------------------------------------------------------------------
# independent plain-numpy restatement of the LBS (float64), for small batches: guards the torch restatement above
# against misread tensor semantics (smpl_layer.py:72-175, tensutils.py:6-53, rodrigues_layer.py:13-52)
# ----------------------------------------------------------------------------------------
def lbs_numpy(model: Dict[str, Tensor], pose: np.ndarray, betas: np.ndarray, trans: np.ndarray,
              offsets: np.ndarray | None = None):
    vt = model["v_template"].double().numpy().reshape(-1, 3)
    sdirs = model["shapedirs"].double().numpy()            # (V,3,nb)
    pdirs = model["posedirs"].double().numpy()             # (V,3,(J-1)*9)
    jreg = model["J_regressor"].double().numpy()           # (J,V)
    wts = model["weights"].double().numpy()                # (V,J)
    parents = [int(p) for p in model["parents"]]
    J = wts.shape[1]
    out_v, out_j = [], []
    for b in range(pose.shape[0]):
        R = []
        for j in range(J):
            r = pose[b, 3 * j:3 * j + 3].astype(np.float64)
            ang = np.linalg.norm(r + 1e-8)
            n = r / ang
            q = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * n])
            q = q / np.linalg.norm(q)
            w, x, y, z = q
            R.append(np.array([[w * w + x * x - y * y - z * z, 2 * x * y - 2 * w * z, 2 * w * y + 2 * x * z],
                               [2 * w * z + 2 * x * y, w * w - x * x + y * y - z * z, 2 * y * z - 2 * w * x],
                               [2 * x * z - 2 * w * y, 2 * w * x + 2 * y * z, w * w - x * x - y * y + z * z]]))
        v_shaped = vt + sdirs @ betas[b].astype(np.float64)
        jnt = jreg @ v_shaped
        pmap = np.concatenate([(R[j] - np.eye(3)).reshape(-1) for j in range(1, J)])
        v_posed = v_shaped + pdirs @ pmap
        if offsets is not None:
            v_posed = v_posed + offsets[b]
        G = [None] * J
        for j in range(J):
            T = np.eye(4)
            T[:3, :3] = R[j]
            T[:3, 3] = jnt[j] - (jnt[parents[j]] if j > 0 else 0.0)
            G[j] = T if j == 0 else G[parents[j]] @ T
        A = []
        for j in range(J):
            M = G[j].copy()
            M[:3, 3] -= G[j][:3, :3] @ jnt[j]
            A.append(M)
        A = np.stack(A)                                      # (J,4,4)
        Tv = np.einsum("vj,jab->vab", wts, A)
        vh = np.concatenate([v_posed, np.ones((v_posed.shape[0], 1))], 1)
        verts = np.einsum("vab,vb->va", Tv, vh)[:, :3] + trans[b]
        out_v.append(verts)
        out_j.append(np.stack([g[:3, 3] for g in G]) + trans[b])
    return np.stack(out_v), np.stack(out_j)
