"""Input / output contract of the hot path: what goes INTO `CHORE.filter` and what the fitter writes OUT.

Mirrors, with the same names and argument meaning (paths relative to /root/reference):
  data/test_data.py:17-57      TestData.__init__ / get_item           -> `TestData`
  data/test_data.py:59-125     TestData.prepare_image_crop             (images (5,512,512), crop_center, scales, crop_info.pkl)
  data/test_data.py:127-231    change_crop_center / pad_image / load_j2d / fullbody_crop / persp_proj / get_bbox
  data/base_data.py:71-192     load_masks / masks2bbox / load_rgb / crop / resize / compose_images
  recon/recon_fit_base.py:240-275   get_output_paths / save_outputs    -> `get_output_paths`, `save_outputs`
  recon/opt_utils.py:74-102    save_smplfits / save_smpl_params        -> `save_smplfits`
  recon/recon_fit_base.py:289-321   load_kpts / scale_body_kpts        -> `load_kpts`, `scale_body_kpts`

This is CPU image IO (cv2 / numpy), not a kernel path; it is here so that `demo.py`-style drivers can feed the
kernels the exact tensors the reference feeds its network and find the reference's output files afterwards.  The
reference needs psbody.mesh for two things only -- reading the vertices of `kX.mocap.ply` and writing `.ply` meshes;
both are done here with a small binary-PLY reader / writer.
"""
from __future__ import annotations

import json
import os
import pickle as pkl
from os.path import isfile, join
from typing import List, Optional, Sequence, Tuple

import numpy as np

# KinectColorCamera pixel intrinsics (model/camera.py:26-40)
FX_PX, FY_PX = 979.7844 / 2048.0 * 2048, 979.840 / 2048.0 * 2048
CX_PX, CY_PX = 1018.952 / 2048.0 * 2048, 779.486 / 2048.0 * 2048


# ---------------------------------------------------------------------------------------------
# PLY (vertices + triangle faces), binary little endian or ascii
# ---------------------------------------------------------------------------------------------
def read_ply(path: str) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Vertices (V,3) float64 and faces (F,3) int32 (None if the file has none) of a PLY mesh."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elems, cur = None, [], None
        while True:
            tok = f.readline().split()
            if not tok:
                continue
            if tok[0] == b"format":
                fmt = tok[1].decode()
            elif tok[0] == b"element":
                cur = {"name": tok[1].decode(), "count": int(tok[2]), "props": []}
                elems.append(cur)
            elif tok[0] == b"property":
                cur["props"].append([t.decode() for t in tok[1:]])
            elif tok[0] == b"end_header":
                break
        npt = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4", "float": "f4",
               "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4",
               "float32": "f4", "float64": "f8"}
        verts, faces = None, None
        for e in elems:
            if e["name"] == "vertex":
                names = [p[-1] for p in e["props"]]
                if fmt == "ascii":
                    rows = np.array([f.readline().split() for _ in range(e["count"])], dtype=np.float64)
                    verts = np.stack([rows[:, names.index(k)] for k in "xyz"], 1)
                else:
                    dt = np.dtype([(p[-1], "<" + npt[p[0]]) for p in e["props"]])
                    rec = np.frombuffer(f.read(dt.itemsize * e["count"]), dtype=dt)
                    verts = np.stack([rec[k].astype(np.float64) for k in "xyz"], 1)
            elif e["name"] == "face":
                if fmt == "ascii":
                    faces = np.array([f.readline().split()[1:4] for _ in range(e["count"])], dtype=np.int32)
                else:
                    cnt_t, idx_t = npt[e["props"][0][1]], npt[e["props"][0][2]]
                    dt = np.dtype([("n", "<" + cnt_t), ("v", "<" + idx_t, (3,))])
                    rec = np.frombuffer(f.read(dt.itemsize * e["count"]), dtype=dt)
                    if not (rec["n"] == 3).all():
                        raise ValueError(f"{path}: only triangle meshes are supported")
                    faces = rec["v"].astype(np.int32)
        return verts, faces


def write_ply(path: str, verts: np.ndarray, faces: Optional[np.ndarray] = None) -> None:
    """Binary little-endian PLY with float vertices and (optionally) triangle faces: what psbody's Mesh.write_ply
    produces for the reconstruction outputs (recon/recon_fit_base.py:267-268, recon/opt_utils.py:88-90)."""
    verts = np.asarray(verts, dtype="<f4").reshape(-1, 3)
    head = ["ply", "format binary_little_endian 1.0", f"element vertex {len(verts)}", "property float x",
            "property float y", "property float z"]
    if faces is not None:
        faces = np.asarray(faces).reshape(-1, 3)
        head += [f"element face {len(faces)}", "property list uchar int vertex_indices"]
    head.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode())
        f.write(verts.tobytes())
        if faces is not None:
            rec = np.empty(len(faces), dtype=np.dtype([("n", "u1"), ("v", "<i4", (3,))]))
            rec["n"], rec["v"] = 3, faces
            f.write(rec.tobytes())


# ---------------------------------------------------------------------------------------------
# landmark regressor on mocap vertices (lib_smpl/body_landmark.py:16-28,61-65)
# ---------------------------------------------------------------------------------------------
class BodyLandmarks:
    def __init__(self, assets_root: Optional[str] = None, body25_reg=None):
        """assets_root: directory with body25_regressor.pkl (the reference's assets/); or pass the (25, 6890) sparse
        regressor directly."""
        if body25_reg is None:
            body25_reg = pkl.load(open(join(assets_root, "body25_regressor.pkl"), "rb"), encoding="latin1").T
        self.body25_reg = body25_reg

    def get_body_kpts(self, verts: np.ndarray) -> np.ndarray:
        """(6890,3) SMPL vertices -> (25,3) body joints."""
        return self.body25_reg.dot(np.asarray(verts))


# ---------------------------------------------------------------------------------------------
# image helpers (data/base_data.py:71-192)
# ---------------------------------------------------------------------------------------------
def masks2bbox(masks: Sequence[np.ndarray], thres: int = 127) -> Tuple[np.ndarray, np.ndarray]:
    """xyxy bounding box of the union of the masks (contours of the thresholded, uint8-wrapping sum, as the reference)."""
    import cv2
    comb = np.zeros_like(masks[0])
    for m in masks:
        comb += m                                  # uint8 arithmetic wraps exactly like the reference's `+=`
    comb = np.clip(comb, 0, 255)
    _, binary = cv2.threshold(comb, thres, 255, cv2.THRESH_BINARY)
    contours, _ = cv2.findContours(binary, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    lo, hi = np.array([50000, 50000]), np.array([-100, -100])
    for c in contours:
        x, y, w, h = cv2.boundingRect(c)
        lo, hi = np.minimum(lo, [x, y]), np.maximum(hi, [x + w, y + h])
    return lo, hi


def crop(img: np.ndarray, center: np.ndarray, crop_size: np.ndarray) -> np.ndarray:
    """Square crop around `center`, zero padded outside the image (note the reference's `w - 1` / `h - 1` limits)."""
    h, w = img.shape[:2]
    tl = np.round(center - crop_size / 2).astype(int)
    br = np.round(center + crop_size / 2).astype(int)
    inner = img[max(0, tl[1]):min(h - 1, br[1]), max(0, tl[0]):min(w - 1, br[0])]
    pad_y = (max(0, -tl[1]), max(0, br[1] - h + 1))
    pad_x = (max(0, -tl[0]), max(0, br[0] - w + 1))
    if img.ndim not in (2, 3):
        raise NotImplementedError
    return np.pad(inner, [pad_y, pad_x] + ([(0, 0)] if img.ndim == 3 else []))


def resize(img: np.ndarray, img_size: Tuple[int, int], mode=None) -> np.ndarray:
    import cv2
    h, w = img.shape[:2]
    assert 1.0 * w / h == 1.0 * img_size[0] / img_size[1], f"image aspect ratio not matching: {img.shape} vs net input {img_size}"
    return cv2.resize(img, img_size, interpolation=cv2.INTER_LINEAR if mode is None else mode)


def compose_images(obj_mask: np.ndarray, person_mask: np.ndarray, rgb: np.ndarray) -> np.ndarray:
    """RGBM3: background-masked RGB + person mask + object mask, (H,W,5)."""
    keep = (person_mask > 0.5) | (obj_mask > 0.5)
    return np.dstack((rgb * keep[..., None], person_mask, obj_mask))


def project_screen(points: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """KinectColorCamera.project_screen for (N,3) points (model/camera.py:51-66)."""
    x, y, z = points[:, 0:1], points[:, 1:2], points[:, 2:3]
    return FX_PX * x / z + CX_PX, FY_PX * y / z + CY_PX


class TestData:
    """Test-time loader: crops and scales the patch so that the person appears as if at z_0 (data/test_data.py:17-231).
    `assets_root` replaces the reference's PATHS.yml:SMPL_ASSETS_ROOT."""
    __test__ = False          # not a pytest class

    def __init__(self, data_paths, batch_size=1, num_workers=0, dtype=np.float32, image_size=(512, 512), input_type="RGBM3",
                 crop_size=1200, use_mean_center=False, assets_root: Optional[str] = None, write_crop_info: bool = True,
                 body25_reg=None, **kwargs):
        assert input_type == "RGBM3"
        self.data_paths, self.batch_size, self.num_workers, self.dtype = list(data_paths), batch_size, num_workers, dtype
        self.img_size = tuple(image_size)                                 # width, height
        self.CROP_SIZE = np.array([crop_size, crop_size])
        self.mean_crop_center = np.array([1008.0, 995.0])                 # BEHAVE training-set mean (test_data.py:32)
        self.use_mean_center = use_mean_center
        self.depth = kwargs.get("z_0", 2.2)
        self.landmark = BodyLandmarks(assets_root, body25_reg) if (assets_root is not None or body25_reg is not None) else None
        self.write_crop_info = write_crop_info
        self.aug_blur = 0.0

    def __len__(self):
        return len(self.data_paths)

    def __getitem__(self, idx):
        return self.get_item(idx)

    def get_loader(self, shuffle=False, rank=-1, world_size=-1):
        from torch.utils.data import DataLoader
        return DataLoader(self, batch_size=self.batch_size, num_workers=self.num_workers, shuffle=shuffle, drop_last=False)

    # ---- per-file loaders -------------------------------------------------------------------------
    @staticmethod
    def load_masks(rgb_file: str, flip: bool = False):
        import cv2
        assert not flip
        def first(*cands):
            for c in cands:
                if isfile(c):
                    return c
            return cands[-1]
        pm = first(rgb_file.replace(".color.jpg", ".person_mask.jpg"), rgb_file.replace(".color.jpg", ".person_mask.png"))
        om = first(rgb_file.replace(".color.jpg", ".obj_rend_mask.jpg"), rgb_file.replace(".color.jpg", ".obj_mask.jpg"),
                   rgb_file.replace(".color.jpg", ".obj_mask.png"))
        return cv2.imread(pm, cv2.IMREAD_GRAYSCALE), cv2.imread(om, cv2.IMREAD_GRAYSCALE)

    @staticmethod
    def load_rgb(rgb_file: str, flip: bool = False) -> np.ndarray:
        from PIL import Image
        assert not flip
        return np.array(Image.open(rgb_file))

    @staticmethod
    def load_j2d(rgb_file: str) -> np.ndarray:
        data = json.load(open(rgb_file.replace(".color.jpg", ".color.json")))
        return np.array(data["body_joints"]).reshape((-1, 3))

    @staticmethod
    def load_mocap_verts(rgb_file: str) -> np.ndarray:
        return read_ply(rgb_file.replace(".color.jpg", ".mocap.ply"))[0]

    # ---- geometry of the crop ---------------------------------------------------------------------
    @staticmethod
    def get_bbox(j2d: np.ndarray, exp: float = 1.1):
        lo, hi = np.min(j2d, 0), np.max(j2d, 0)
        return lo, (hi - lo) * exp

    def persp_proj(self, points: np.ndarray) -> np.ndarray:
        px, py = project_screen(points)
        return np.concatenate([px, py, np.ones_like(px)], 1)

    def fullbody_crop(self, pts: np.ndarray, rgb_file: str) -> float:
        """Scale of the crop square such that, after resizing, the detected 2-D joints have the extent the mocap body
        would have at depth z_0 (test_data.py:172-211)."""
        if self.landmark is None:
            raise RuntimeError("TestData needs assets_root (body25_regressor.pkl) to compute the crop scale")
        verts = self.load_mocap_verts(rgb_file)
        verts = verts - np.mean(verts, 0) + np.array([0, 0, self.depth])
        proj = self.persp_proj(self.landmark.get_body_kpts(verts))
        valid = pts[:, 2] > 0.3
        _, ext = self.get_bbox(pts[valid][:, :2])
        _, ext_mocap = self.get_bbox(proj[valid][:, :2])
        use_width = ext[0] >= ext[1] and ext_mocap[0] >= ext_mocap[1]
        return ext[0] / ext_mocap[0] if use_width else ext[1] / ext_mocap[1]

    def change_crop_center(self, crop_center):
        return self.mean_crop_center.copy() if self.use_mean_center else crop_center

    def pad_image(self, img, crop_center):
        """Shift the image so that `crop_center` lands on the mean crop centre (only with use_mean_center)."""
        if not self.use_mean_center:
            return img
        h, w = img.shape[:2]
        tl = (self.mean_crop_center - crop_center).astype(int)
        br = np.array([w, h]) + tl
        kw, kh = 2048, 1536
        size = np.maximum(np.array([kw, kh]), br).astype(int)
        out = np.zeros((size[1], size[0], 3) if img.ndim == 3 else (size[1], size[0]))
        d0 = np.maximum(np.zeros(2), tl).astype(int)
        d1 = np.minimum(np.array([kw, kh]), br).astype(int)
        sx0, sy0 = max(0, -tl[0]), max(0, -tl[1])
        sx1, sy1 = min(w, w - (br[0] - kw)), min(h, h - (br[1] - kh))
        out[d0[1]:d1[1], d0[0]:d1[0]] = img[sy0:sy1, sx0:sx1]
        return out

    # ---- the contract -----------------------------------------------------------------------------
    def prepare_image_crop(self, rgb_file: str, flip: bool = False):
        """-> images (5,H,W) in [0,1], crop_center (2,), resize_scale, crop_scale, old_center (test_data.py:59-125)."""
        import cv2
        assert not flip, "for evaluation, do not do flip!"
        person_mask, obj_mask = self.load_masks(rgb_file)
        lo, hi = masks2bbox([person_mask, obj_mask])
        extent = hi - lo
        assert extent[0] <= self.CROP_SIZE[0] and extent[1] <= self.CROP_SIZE[1], f"crop too small for {rgb_file} with bbox {extent}"
        crop_center = (lo + hi) // 2
        rgb = self.load_rgb(rgb_file)
        rh, rw = rgb.shape[:2]
        if rw > rh:                                     # to the 2048 x 1536 Kinect pixel space
            resize_scale = 2048 / rw
            newsize = (2048, int(rh * resize_scale))
        else:
            resize_scale = 1536 / rh
            newsize = (int(rw * resize_scale), 1536)
        crop_center = np.round(resize_scale * crop_center)
        rgb, person_mask, obj_mask = (cv2.resize(x, newsize) for x in (rgb, person_mask, obj_mask))
        kpts = self.load_j2d(rgb_file)
        if np.sum(kpts[:, 2]) == 0:
            raise ValueError(f"no valid person keypoints in image {rgb_file}")
        kpts = kpts.copy()
        kpts[:, :2] *= resize_scale
        scale = self.fullbody_crop(kpts, rgb_file)
        crop_size = scale * self.CROP_SIZE
        rgb, person_mask, obj_mask = (self.pad_image(x, crop_center) for x in (rgb, person_mask, obj_mask))
        old_center = crop_center.copy()
        crop_center = self.change_crop_center(crop_center)
        rgb, person_mask, obj_mask = (resize(crop(x, crop_center, crop_size), self.img_size) / 255.0
                                      for x in (rgb, person_mask, obj_mask))
        images = compose_images(obj_mask, person_mask, rgb)
        info_file = rgb_file.replace(".color.jpg", ".crop_info.pkl")
        if self.write_crop_info and not isfile(info_file):
            pkl.dump({"rgb_newsize": np.array(newsize), "resize_scale": resize_scale, "crop_center": old_center,
                      "crop_scale": scale, "crop_size": crop_size}, open(info_file, "wb"))
        return images.transpose((2, 0, 1)).astype(self.dtype), crop_center, resize_scale, scale, old_center

    def get_item(self, idx):
        rgb_file = self.data_paths[idx]
        images, center, resize_scale, scale, old_center = self.prepare_image_crop(rgb_file, False)
        return {"images": images.astype(self.dtype), "path": rgb_file, "resize_scale": resize_scale, "crop_scale": scale,
                "crop_center": center.astype(self.dtype), "old_crop_center": old_center}


# ---------------------------------------------------------------------------------------------
# keypoints into network-input pixels (recon/recon_fit_base.py:294-321)
# ---------------------------------------------------------------------------------------------
def load_kpts(json_paths: Sequence[str], tol: float = 0.3, device="cpu"):
    import torch
    out = []
    for p in json_paths:
        j = np.array(json.load(open(p))["body_joints"]).reshape((-1, 3))
        j[:, 2][j[:, 2] < tol] = 0
        out.append(j)
    return torch.tensor(np.stack(out, 0), dtype=torch.float32).to(device)


def scale_body_kpts(kpts, resize_scale, crop_scale, crop_center, crop_size: float = 1200.0, net_in_size: int = 512):
    """(B,25,3) keypoints of the original image -> pixels of the network input (crop of crop_scale * crop_size around
    crop_center in the 2048-px image, resized to net_in_size)."""
    import torch
    pxy = kpts[:, :, :2] * resize_scale.unsqueeze(1).unsqueeze(1)
    org = crop_scale * crop_size
    pxy = pxy - crop_center.unsqueeze(1) + org.unsqueeze(1).unsqueeze(1) / 2
    pxy = pxy * net_in_size / org.unsqueeze(1).unsqueeze(1)
    return torch.cat([pxy, kpts[:, :, 2:3]], -1)


# ---------------------------------------------------------------------------------------------
# outputs (recon/recon_fit_base.py:240-275, recon/opt_utils.py:74-102)
# ---------------------------------------------------------------------------------------------
def get_output_paths(outpath: str, image_paths: Sequence[str], save_name: str, test_id: int = 1) -> Tuple[List[str], List[str]]:
    """ROOT/SEQ/frame/kX.color.jpg -> outpath/SEQ/frame/save_name/kX.{smpl,object}.ply (directories are created)."""
    smpl_files, obj_files = [], []
    for x in image_paths:
        parts = x.split(os.sep)
        folder = join(outpath, parts[-3], parts[-2], save_name)
        os.makedirs(folder, exist_ok=True)
        smpl_files.append(join(folder, f"k{test_id}.smpl.ply"))
        obj_files.append(join(folder, f"k{test_id}.object.ply"))
    return smpl_files, obj_files


def is_done(outpath: str, image_paths: Sequence[str], save_name: str, test_id: int = 1) -> bool:
    smpl_files, obj_files = get_output_paths(outpath, image_paths, save_name, test_id)
    return all(isfile(a) and isfile(b) for a, b in zip(smpl_files, obj_files))


def save_smplfits(save_paths: Sequence[str], scores, smpl, save_mesh: bool = True, ext: str = ".ply"):
    """kX.smpl.ply (posed vertices + faces) and kX.smpl.pkl {'pose','betas','trans','score'} per batch element."""
    verts = smpl()[0].detach().cpu().numpy()
    faces = smpl.faces.detach().cpu().numpy()
    poses, betas, trans = (getattr(smpl, k).detach().cpu().numpy() for k in ("pose", "betas", "trans"))
    for i, path in enumerate(save_paths):
        if save_mesh:
            write_ply(path, verts[i], faces)
        pkl.dump({"pose": poses[i], "betas": betas[i], "trans": trans[i], "score": scores[i]}, open(path.replace(ext, ".pkl"), "wb"))
    return poses, betas, trans, scores


def save_outputs(fitter, smpl, obj_R, obj_t, traindata_paths, save_name, test_id, obj_s, outpath: str, scan_verts, scan_faces):
    """SMPL meshes / parameters and the transformed object template + {'rot','trans','scale'} pickles
    (recon/recon_fit_base.py:258-275).  `scan_verts` (V,3) / `scan_faces` (F,3): the object template the fitter samples."""
    import torch
    smpl_files, obj_files = get_output_paths(outpath, traindata_paths, save_name, test_id)
    save_smplfits(smpl_files, np.zeros(len(obj_t)), smpl)
    B = len(obj_files)
    template = torch.as_tensor(np.asarray(scan_verts), dtype=torch.float32).repeat(B, 1, 1).to(obj_t.device)
    with torch.no_grad():
        moved = fitter.transform_object(template, obj_R, obj_t, obj_s).cpu().numpy()
        rot = fitter.decopose_axis(obj_R, no_rand=True).cpu().numpy()
    for v, path, r, s, t in zip(moved, obj_files, rot, obj_s.detach().cpu().numpy(), obj_t.detach().cpu().numpy()):
        write_ply(path, v, scan_faces)
        pkl.dump({"rot": r, "trans": t, "scale": s}, open(path.replace(".ply", ".pkl"), "wb"))
    return smpl_files, obj_files
