"""Host-side pieces of the silhouette phase (chore_b200/silhouette.py): PHOSA bbox helpers, the ROI intrinsics and the occlusion
mask convention of recon/obj_pose_roi.py / recon/bbox.py.  CPU only."""
import numpy as np
import torch

from chore_b200 import silhouette as SL


def test_bbox_helpers():
    xyxy = np.array([[10.0, 20.0, 110.0, 70.0]])
    wh = SL.bbox_xy_to_wh(xyxy)
    assert np.array_equal(wh, [[10, 20, 100, 50]]) and np.array_equal(SL.bbox_wh_to_xy(wh), xyxy)
    sq = SL.make_bbox_square(wh, 0.3)
    assert np.allclose(sq, [[60 - 65, 45 - 65, 130, 130]])                   # same centre, side = max(w, h) * 1.3
    m = np.zeros((64, 64), np.uint8)
    m[10:20, 30:50] = 255
    assert np.array_equal(SL.mask2bbox(m), [30, 10, 50, 20])


def test_roi_camera_maps_bbox_to_unit_square():
    """A point that projects to the bbox corner / centre in the 2048-px Kinect image lands at 0 / 0.5 of the ROI image."""
    bbox = SL.SilLossROI.to_original_bbox(np.array([100.0, 150.0, 200.0, 200.0]), 1200 / 512.0, np.array([1008.0, 995.0]))
    assert np.allclose(bbox, [100 * 1200 / 512 + 1008 - 600, 150 * 1200 / 512 + 995 - 600, 200 * 1200 / 512, 200 * 1200 / 512])
    K = SL.SilLossROI.compute_K_roi(bbox)[0]
    z = 2.0
    for frac in (0.0, 0.5, 1.0):
        px, py = bbox[0] + frac * bbox[2], bbox[1] + frac * bbox[2]
        X = torch.tensor([(px - 1018.952) * z / 979.7844, (py - 779.486) * z / 979.840, z], dtype=torch.float32)
        uv = K @ (X / z)
        assert torch.allclose(uv[:2], torch.tensor([frac, frac]), atol=1e-5)


def test_occlusion_mask_convention_and_projection():
    ps = torch.tensor([[1.0, 1.0, 0.0, 0.0]])
    ob = torch.tensor([[1.0, 0.0, 1.0, 0.0]])
    assert SL.SilLossROI.cvt_masks(ps, ob).tolist() == [[True, False, True, True]]      # person-only pixels are ignored
    # projection: orig_size 1, identity extrinsics: u = 2 (fx x/z + cx) - 1, v flipped
    K = torch.tensor([[[2.0, 0, 0.5], [0, 2.0, 0.5], [0, 0, 1]]])
    v = SL.projection(torch.tensor([[[0.1, -0.2, 2.0]]]), K, torch.eye(3)[None], torch.zeros(1, 3), torch.zeros(1, 5), 1)
    assert torch.allclose(v, torch.tensor([[[2 * (2 * 0.05 + 0.5) - 1, 2 * (1 - (2 * -0.1 + 0.5)) - 1, 2.0]]]), atol=1e-6)
    faces = SL.vertices_to_faces(torch.arange(24.0).view(2, 4, 3), torch.tensor([[[0, 1, 2]], [[1, 2, 3]]]))
    assert faces.shape == (2, 1, 3, 3) and faces[1, 0, 0].tolist() == [15.0, 16.0, 17.0]


def test_crop_and_resize_identity_box():
    m = torch.zeros(1, 32, 32)
    m[0, 8:24, 4:20] = 1
    out = SL.crop_and_resize_masks(m, torch.tensor([[0.0, 0.0, 32.0, 32.0]]), 32)
    assert out.shape == (1, 32, 32) and torch.equal(out[0], m[0].bool())
    half = SL.crop_and_resize_masks(m, torch.tensor([[4.0, 8.0, 20.0, 24.0]]), 8)
    assert bool(half.all())
