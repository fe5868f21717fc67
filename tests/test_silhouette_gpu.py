"""Silhouette rasteriser (csrc/silhouette.cu, chore_b200/silhouette.py) against the REFERENCE's own CUDA kernels
(external/neural_renderer/.../rasterize_cuda_kernel.cu compiled unmodified into oracle/_ref/nmr_rasterize_ref.so by
oracle/build_ref.py): face-index / alpha maps, the NMR gradient with the reference's maps as input (stage-wise), and the whole
SilLossROI term end to end with the reference kernels substituted for ours."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref
    m = build_ref.load_built()
    if m is None:
        pytest.skip("oracle/_ref/nmr_rasterize_ref.so not built (python -m oracle.build_ref in the build container)")
    return m


def ref_forward(ref, faces, S, near=0.1, far=100.0):
    """The buffers of RasterizeFunction.forward (rasterize.py:52-91) for return_alpha only."""
    B, F = faces.shape[:2]
    bs, buf = 4, 512
    face_list = torch.zeros(B, (S - 1) // bs + 1, (S - 1) // bs + 1, buf, dtype=torch.int32, device=DEV)
    index = torch.full((B, S, S), -1, dtype=torch.int32, device=DEV)
    weight = torch.zeros(B, S, S, 3, device=DEV)
    depth = torch.full((B, S, S), far, device=DEV)
    one = torch.zeros(1, device=DEV)
    ref.forward_face_index_map(faces.clone(), index, weight, depth, one, torch.zeros_like(faces), torch.zeros(1, dtype=torch.int32, device=DEV),
                               face_list, S, bs, near, far, 0, 1, 0, 0)
    return index, (index >= 0).float(), int(face_list[..., 0].max())


def ref_backward(ref, faces, index, alpha, g_alpha, S, eps=1e-4):
    one = torch.zeros(1, device=DEV)
    return ref.backward_pixel_map(faces, index, one, alpha, one, g_alpha.contiguous(), torch.zeros_like(faces), S, eps, 0, 1)


def random_faces(seed, B, F, size=0.12):
    """Small random triangles (both windings) in normalised coordinates with depths in [1, 3]."""
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(B, F, 1, 2, generator=g) * 2.2 - 1.1                       # some stick out of the image
    xy = c + size * torch.randn(B, F, 3, 2, generator=g)
    z = 1.0 + 2.0 * torch.rand(B, F, 3, 1, generator=g)
    return torch.cat([xy, z], -1).to(DEV).contiguous()


def uv_sphere(n_lat=24, n_lon=32, r=0.35):
    th = np.linspace(0, np.pi, n_lat + 1)[1:-1]
    ph = np.linspace(0, 2 * np.pi, n_lon, endpoint=False)
    v = [[0, r, 0]] + [[r * np.sin(t) * np.cos(p), r * np.cos(t), r * np.sin(t) * np.sin(p)] for t in th for p in ph] + [[0, -r, 0]]
    f = []
    ring = lambda i, j: 1 + i * n_lon + (j % n_lon)
    for j in range(n_lon):
        f.append([0, ring(0, j + 1), ring(0, j)])
        f.append([len(v) - 1, ring(n_lat - 2, j), ring(n_lat - 2, j + 1)])
    for i in range(n_lat - 2):
        for j in range(n_lon):
            f.append([ring(i, j), ring(i, j + 1), ring(i + 1, j)])
            f.append([ring(i, j + 1), ring(i + 1, j + 1), ring(i + 1, j)])
    return np.array(v, np.float32), np.array(f, np.int64)


@pytest.mark.parametrize("S,F", [(64, 200), (256, 1500)])
def test_forward_and_gradient_vs_reference_kernels(ref, S, F):
    from chore_b200 import _lib
    h = _lib.get_handle(torch.device(DEV))
    faces = random_faces(S + F, 2, F)
    r_index, r_alpha, fill = ref_forward(ref, faces, S)
    assert fill <= 512, "test mesh overflows the reference's 512-face block lists (it would silently drop faces)"
    alpha, index = h.silhouette_fwd(faces, S)
    assert (alpha != r_alpha).float().mean() < 2e-4, (alpha != r_alpha).float().mean().item()      # razor-edge pixels only
    same = index == r_index
    assert same.float().mean() > 0.995, same.float().mean().item()          # equal-depth ties are broken in list order by the reference
    # the NMR gradient, stage-wise: the reference's maps in, random upstream gradient
    g_alpha = torch.randn(2, S, S, generator=torch.Generator().manual_seed(1)).to(DEV)
    want = ref_backward(ref, faces, r_index, r_alpha, g_alpha, S)
    got = h.silhouette_bwd(faces, r_index, r_alpha, g_alpha)
    assert float(want.abs().max()) > 0 and rel_err(got, want) < 1e-4, rel_err(got, want)
    assert float(got[..., 2].abs().max()) == 0.0                              # no gradient to the depths
    # and on our own maps: only the faces touching a differing pixel may differ
    got2 = h.silhouette_bwd(faces, index, alpha, g_alpha)
    close = ((got2 - want).abs().amax((2, 3)) <= 1e-4 * (want.abs().amax((2, 3)) + want.abs().mean()))
    assert close.float().mean() > 0.99, close.float().mean().item()


def test_sil_loss_roi_vs_reference_rasteriser(ref):
    """SilLossROI end to end (ROI crop, K_roi, projection, rasteriser, occlusion mask, L2) with a sphere template: loss and the
    gradients to (R, t, s) against the same module with the reference's kernels in place of ours."""
    import types
    import chore_b200.silhouette as SL
    v, f = uv_sphere()
    mesh = types.SimpleNamespace(v=v, f=f)
    B = 2
    yy, xx = torch.meshgrid(torch.arange(512.0), torch.arange(512.0), indexing="ij")
    obj_masks = torch.stack([((xx - 300) ** 2 + (yy - 260) ** 2 < 70 ** 2).float(), ((xx - 200) ** 2 + (yy - 300) ** 2 < 55 ** 2).float()])
    ps_masks = torch.stack([((xx - 240) ** 2 / 2 + (yy - 250) ** 2 < 60 ** 2).float(), ((xx - 260) ** 2 + (yy - 250) ** 2 / 3 < 50 ** 2).float()])
    cc = torch.tensor([[1008.0, 995.0], [1000.0, 990.0]])
    sil = SL.SilLossROI(ps_masks, obj_masks, mesh, cc, device=DEV)
    assert sil.image_ref.shape == (B, 256, 256) and sil.keep_mask.shape == (B, 256, 256) and sil.edt_ref_edge.shape == (B, 256, 256)

    def run(rasteriser):
        R = (torch.eye(3).repeat(B, 1, 1) + 0.05 * torch.randn(B, 3, 3, generator=torch.Generator().manual_seed(2))).to(DEV)
        t = torch.tensor([[0.25, 0.05, 2.3], [-0.1, 0.15, 2.1]], device=DEV)
        s = torch.tensor([0.7, 1.4], device=DEV)
        # gradient to the transformed vertices: the chain to (R, t, s) behind it is plain torch in both arms, and the scalar
        # gradients are residuals of +-1e4-sized per-vertex terms (NMR divides by pixel distances down to eps = 1e-4)
        verts = sil.apply_transformation(R, t, s).detach().requires_grad_(True)
        old = SL.rasterize_silhouettes
        SL.rasterize_silhouettes = rasteriser or old
        try:
            image = sil.keep_mask * sil.renderer(verts, sil.faces, mode="silhouettes")
        finally:
            SL.rasterize_silhouettes = old
        loss = torch.sum((image - sil.image_ref) ** 2, dim=(1, 2)).mean()
        loss.backward()
        return loss.detach(), image.detach(), verts.grad

    class RefRaster(torch.autograd.Function):
        @staticmethod
        def forward(ctx, faces, size):
            faces = faces.detach().contiguous()
            idx, alpha, _ = ref_forward(ref, faces, size)
            ctx.save_for_backward(faces, idx, alpha)
            ctx.size = size
            return alpha.clone()

        @staticmethod
        def backward(ctx, g):
            faces, idx, alpha = ctx.saved_tensors
            return ref_backward(ref, faces, idx, alpha, g, ctx.size), None

    ref_raster = lambda faces, image_size, anti_aliasing, near, far: RefRaster.apply(faces, image_size).flip(1)
    l0, img0, gv0 = run(ref_raster)
    l1, img1, gv1 = run(None)
    assert float(img0.sum()) > 500 and (img0 != img1).float().mean() < 2e-4
    assert rel_err(l1, l0) < 2e-3, (float(l1), float(l0))
    assert float(gv0.abs().max()) > 0
    # per vertex: equal except around the handful of razor-edge pixels where the two alpha maps differ
    close = (gv1 - gv0).abs().amax(-1) <= 1e-3 * (gv0.abs().amax(-1) + gv0.abs().mean())
    assert close.float().mean() > 0.98, close.float().mean().item()
    # and the module's own forward() returns the reference's 5-tuple and is differentiable down to (R, t, s)
    R = torch.eye(3, device=DEV).repeat(B, 1, 1).requires_grad_(True)
    t = torch.tensor([[0.25, 0.05, 2.3], [-0.1, 0.15, 2.1]], device=DEV, requires_grad=True)
    sc = torch.ones(B, device=DEV, requires_grad=True)
    loss, image, edges, image_ref, edt = sil(R, t, sc)
    loss["mask"].backward()
    assert image.shape == edges.shape == image_ref.shape == edt.shape == (B, 256, 256)
    assert all(x.grad is not None and torch.isfinite(x.grad).all() for x in (R, t, sc)) and float(t.grad.abs().max()) > 0
