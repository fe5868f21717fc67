"""Batch / tiling consistency at production sizes (B200, through the C ABI).

The golden vectors are small (the oracle has to finish in seconds); the two worst bugs of round 2 only showed with more
work items than SMs AND more than one image (a CTA walking items of different images, an epilogue group idling across
tiles).  These tests have no oracle: they check that a batched / multi-tile call equals the per-image calls, which must hold
bit for bit for the query kernels (deterministic, no atomics) and to summation order for the rest."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import rel_err
from oracle import chore_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def net():
    import chore_b200
    n = chore_b200.CHORE(device=DEV)
    n.load_state_dict(O.make_state_dict(0, "unit"))
    return n.eval()


def _maps(net, B, seed=21):
    feat, tmpx = O.synth_features(seed, B=B)
    net.im_feat_list, net.tmpx = [feat.to(DEV)], tmpx.to(DEV)
    return net._maps()


@pytest.mark.parametrize("mask", [15, 1, 6])
def test_query_forward_batched_equals_per_image(net, mask):
    """4 images x 40 003 points = 1 252 tiles on 148 SMs: every CTA walks ~8 tiles of several images (ragged last tiles)."""
    B, N = 4, 40003
    f, s = _maps(net, B)
    cc = torch.tensor([[1008., 995.], [990., 1001.], [1100., 900.], [950., 1020.]], device=DEV)
    pts = torch.cat([O.synth_points("frustum", 31, B, N - N // 3), O.synth_points("init_box", 32, B, N // 3)], 1).to(DEV)
    full, m = net.handle.query_fwd(f, s, pts, cc, mask, want_in_img=True)
    full = [None if x is None else x.clone() for x in full]
    for b in range(B):
        one, m1 = net.handle.query_fwd(f[b:b + 1].contiguous(), s[b:b + 1].contiguous(), pts[b:b + 1].contiguous(), cc[b:b + 1].contiguous(),
                                       mask, want_in_img=True)
        torch.cuda.synchronize()
        assert torch.equal(m[b:b + 1], m1)
        for h in range(4):
            if (mask >> h) & 1:
                assert torch.equal(full[h][b:b + 1], one[h]), (b, h, (full[h][b:b + 1] - one[h]).abs().max().item())


def test_query_backward_batched_equals_per_image(net):
    """Gradient to the points: 3 images x 30 001 points, two heads with a gradient (705 work items per head)."""
    B, N = 3, 30001
    f, s = _maps(net, B, seed=22)
    cc = torch.tensor([[1008., 995.], [990., 1001.], [1100., 900.]], device=DEV)
    pts = torch.cat([O.synth_points("frustum", 33, B, N - N // 4), O.synth_points("init_box", 34, B, N // 4)], 1).to(DEV)
    g = torch.Generator().manual_seed(5)
    g_df = torch.randn(B, 2, N, generator=g).to(DEV)
    g_parts = torch.randn(B, 14, N, generator=g).to(DEV)
    full = net.handle.query_bwd(f, s, pts, cc, [g_df, None, g_parts, None]).clone()
    again = net.handle.query_bwd(f, s, pts, cc, [g_df, None, g_parts, None])
    torch.cuda.synchronize()
    assert torch.equal(full, again)                                   # deterministic
    for b in range(B):
        one = net.handle.query_bwd(f[b:b + 1].contiguous(), s[b:b + 1].contiguous(), pts[b:b + 1].contiguous(), cc[b:b + 1].contiguous(),
                                   [g_df[b:b + 1].contiguous(), None, g_parts[b:b + 1].contiguous(), None])
        torch.cuda.synchronize()
        assert rel_err(full[b:b + 1], one) < 1e-6, (b, rel_err(full[b:b + 1], one))
    only_df = net.handle.query_bwd(f, s, pts, cc, [g_df, None, None, None])
    only_parts = net.handle.query_bwd(f, s, pts, cc, [None, None, g_parts, None])
    torch.cuda.synchronize()
    assert rel_err(only_df + only_parts, full) < 1e-5                 # linear in the upstream gradients


def test_query_grid_chunking_is_bit_identical_at_production_chunk_sizes(net):
    """128 x 128 x 256 grid (4.19 M points) in one call, in four 2^20-point chunks and in ragged chunks."""
    f, s = _maps(net, 2, seed=23)
    cc = torch.tensor([[1008., 995.], [990., 1001.]], device=DEV)
    res, pmin, pmax = [128, 128, 256], [-3.0, -0.9, 0.2], [3.0, 1.8, 4.0]
    total = res[0] * res[1] * res[2]

    def run(chunks):
        o = [torch.zeros(c, total, device=DEV) for c in (2, 9, 14, 6)]
        for start, count in chunks:
            net.handle.query_grid(f, s, cc, 1, res, pmin, pmax, start, count, 15, o)
        torch.cuda.synchronize()
        return o

    one = run([(0, total)])
    four = run([(i * (total // 4), total // 4) for i in range(4)])
    ragged = run([(0, 1000003), (1000003, 2000001), (3000004, total - 3000004)])
    for a, b, c in zip(one, four, ragged):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_lbs_batched_equals_per_item():
    """SMPL-H LBS forward + backward for 8 bodies in one call against 8 single calls."""
    import chore_b200
    layer = chore_b200.SMPLHLayer(O.make_smplh_buffers(0), device=DEV)
    B = 8
    g = torch.Generator().manual_seed(9)
    pose = (0.3 * torch.randn(B, 156, generator=g)).to(DEV)
    betas = torch.randn(B, 10, generator=g).to(DEV)
    trans = torch.randn(B, 3, generator=g).to(DEV)
    offs = (0.01 * torch.randn(B, 6890, 3, generator=g)).to(DEV)
    gv = torch.randn(B, 6890, 3, generator=g).to(DEV)
    gj = torch.randn(B, 52, 3, generator=g).to(DEV)

    def run(sl):
        p, b, t, o = (x[sl].clone().requires_grad_(True) for x in (pose, betas, trans, offs))
        v, j, _, _ = layer(p, th_betas=b, th_trans=t, th_offsets=o)
        ((gv[sl] * v).sum() + (gj[sl, :j.shape[1]] * j).sum()).backward()
        return v.detach(), j.detach(), p.grad, b.grad, t.grad, o.grad

    full = run(slice(0, B))
    for i in range(B):
        one = run(slice(i, i + 1))
        for name, a, b in zip(("verts", "jtr", "g_pose", "g_betas", "g_trans", "g_offsets"), full, one):
            assert rel_err(a[i:i + 1], b) < 1e-5, (i, name, rel_err(a[i:i + 1], b))


_ENC_SHAPES = [(3, 512, 512), (2, 384, 512), (5, 256, 256), (2, 128, 192), (1, 64, 64), (9, 128, 128)]
_ENC_CHILD = r"""
import sys, torch
sys.path.insert(0, sys.argv[1])
import chore_b200
from oracle import chore_oracle as O
net = chore_b200.CHORE(device="cuda:0")
net.load_state_dict(O.make_state_dict(0, "unit"))
out = {}
for (B, H, W) in eval(sys.argv[3]):
    g = torch.Generator().manual_seed(B * 1000 + H + W)
    img = torch.rand(B, 5, H, W, generator=g)
    img[:, 3:] = (img[:, 3:] > 0.5).float()
    out[(B, H, W)] = [t.cpu() for t in net.handle.encode(img.to("cuda:0"))]
torch.cuda.synchronize()
torch.save(out, sys.argv[2])
"""


def test_encoder_matches_round1_encoder_across_shapes(tmp_path):
    """Two independent implementations of the hourglass encoder -- conv_hx.cu / encoder_hx.cu (halo tiles, GroupNorm folded into
    the convolutions, cluster K split) and the round-1 conv_tc.cu / encoder.cu (TMA im2col, separate statistics kernels) -- on
    batched, non-square and small inputs.  The implementation is chosen per process (CHORE_B200_ENCODER), hence the children."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for mode in ("tc1", "hx"):
        path = str(tmp_path / f"enc_{mode}.pt")
        env = dict(os.environ, CHORE_B200_ENCODER=mode)
        subprocess.run([sys.executable, "-c", _ENC_CHILD, root, path, repr(_ENC_SHAPES)], check=True, env=env, timeout=300)
        outs[mode] = torch.load(path)
    for k in _ENC_SHAPES:
        for name, a, b in zip(("feat", "skip", "normx"), outs["hx"][k], outs["tc1"][k]):
            assert rel_err(a, b) < 5e-5, (k, name, rel_err(a, b))


def test_encode_with_ever_new_buffers_uses_the_generic_graph_and_agrees(net):
    """chore_encode caches address-specific CUDA graphs; after four per shape it falls back to one generic graph on staging
    buffers + copies.  Ten encodes of the same image into buffers that are all kept alive (ten distinct address sets) must agree."""
    img0 = O.synth_images(4, B=1, size=256)
    keep, first = [], None
    for i in range(10):
        img = img0.to(DEV).clone()
        out = net.handle.encode(img)
        torch.cuda.synchronize()
        keep.append((img, out))
        if first is None:
            first = [o.clone() for o in out]
        else:
            for name, a, b in zip(("feat", "skip", "normx"), out, first):
                assert rel_err(a, b) < 5e-5, (i, name, rel_err(a, b))
    assert len({k[1][0].data_ptr() for k in keep}) == 10
