"""Import the REAL reference (read-only, /root/reference) with import-time stubs for the
third-party packages this image lacks.  Usable only in the build container -- the GPU box
has no /root/reference -- and only by oracle/make_golden.py and tests that are skipped
when the reference tree is absent.  TEST INFRASTRUCTURE ONLY.

Nothing here copies or modifies reference sources; it only arranges sys.path / cwd / stubs
(SURVEY.md section 8c).
"""
from __future__ import annotations

import contextlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("CHORE_REFERENCE_ROOT", "/root/reference")
_STUB_ROOTS = ("skimage", "chumpy", "psbody", "pytorch3d", "mesh_intersection", "trimesh",
               "neural_renderer", "detectron2", "igl", "open3d", "sklearn_stub_never")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


class _AnyStub(types.ModuleType):
    """Module whose every attribute is a dummy class (enough for `from x import Y`)."""
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__call__": lambda self, *a, **k: None})
        setattr(self, name, cls)
        return cls


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _AnyStub(spec.name)

    def exec_module(self, module):
        pass


_installed = False


def install():
    """Put the reference on sys.path behind the stub finder (idempotent)."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    sys.meta_path.append(_StubFinder())          # appended: real packages win when present
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, os.path.join(REF_ROOT, "lib_smpl", "smplpytorch"))
    _installed = True


@contextlib.contextmanager
def ref_cwd():
    """The reference reads config/ and PATHS.yml relative to CWD (config_loader.py:10,
    wrapper_pytorch.py:16)."""
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


def load_chore():
    """-> (CHORE instance on CPU in eval mode, args Namespace)."""
    install()
    with ref_cwd(), contextlib.redirect_stdout(open(os.devnull, "w")):
        from config.config_loader import load_configs
        from model.chore import CHORE
        args = load_configs("chore-release")
        net = CHORE(args)
    net.eval()
    for p in net.parameters():
        p.requires_grad_(False)
    return net, args


def load_smpl_layer(buffers):
    """SMPL_Layer with synthetic SMPL-H-shaped buffers (the licensed pickle is absent):
    bypass __init__ (smpl_layer.py:19-70) and register what forward() touches."""
    install()
    import torch
    with ref_cwd():
        from smplpytorch.pytorch.smpl_layer import SMPL_Layer
    layer = SMPL_Layer.__new__(SMPL_Layer)
    torch.nn.Module.__init__(layer)
    layer.hands = True
    layer.gender = "male"
    layer.center_idx = 0
    layer.register_buffer("th_betas", torch.zeros(1, buffers["shapedirs"].shape[-1]))
    layer.register_buffer("th_shapedirs", buffers["shapedirs"])
    layer.register_buffer("th_posedirs", buffers["posedirs"])
    layer.register_buffer("th_v_template", buffers["v_template"])
    layer.register_buffer("th_J_regressor", buffers["J_regressor"])
    layer.register_buffer("th_weights", buffers["weights"])
    layer.register_buffer("th_faces", buffers["faces"])
    layer.kintree_parents = [int(p) if p >= 0 else 4294967295 for p in buffers["parents"]]
    layer.num_joints = len(layer.kintree_parents)
    return layer


def load_fitter_class():
    install()
    with ref_cwd(), contextlib.redirect_stdout(open(os.devnull, "w")):
        from recon.recon_fit_behave import ReconFitterBehave
    return ReconFitterBehave


def load_testdata_class():
    """The reference's TestData (data/test_data.py).  psbody.mesh is stubbed, so `load_mocap_mesh` (the only Mesh use on
    this path, test_data.py:213-217) is replaced by a numpy read of the binary PLY vertex block."""
    install()
    import numpy as np
    with ref_cwd(), contextlib.redirect_stdout(open(os.devnull, "w")):
        from data.test_data import TestData

    def load_mocap_mesh(self, rgb_file):
        raw = open(rgb_file.replace(".color.jpg", ".mocap.ply"), "rb").read()
        head, body = raw.split(b"end_header\n", 1)
        n = int([l for l in head.decode().splitlines() if l.startswith("element vertex")][0].split()[-1])
        assert b"binary_little_endian" in head and b"property float x" in head
        return types.SimpleNamespace(v=np.frombuffer(body[:n * 12], dtype="<f4").reshape(n, 3).astype(np.float64))

    TestData.load_mocap_mesh = load_mocap_mesh
    return TestData


def load_generator_class():
    install()
    with ref_cwd(), contextlib.redirect_stdout(open(os.devnull, "w")):
        from recon.generator import Generator
    return Generator


def load_split_smpl(buffers, batch_sz, pose, betas, trans, assets_root=None):
    """The reference's SMPLPyTorchWrapperBatchSplitParams (lib_smpl/wrapper_pytorch.py:93-190) around a
    synthetic-buffer SMPL_Layer: __init__ would read the licensed SMPL-H pickle, so the module is built with
    __new__ and given exactly the attributes __init__ sets (:108-160); forward / get_landmarks are the
    reference's own code.  The landmark regressors are the reference's real assets."""
    install()
    import torch
    import torch.nn as nn
    with ref_cwd():
        from lib_smpl.wrapper_pytorch import SMPLPyTorchWrapperBatchSplitParams
        from lib_smpl.body_landmark import load_regressors
        regs = load_regressors(assets_root or os.path.join(REF_ROOT, "assets"), batch_size=batch_sz)
    m = SMPLPyTorchWrapperBatchSplitParams.__new__(SMPLPyTorchWrapperBatchSplitParams)
    nn.Module.__init__(m)
    m.top_betas = nn.Parameter(betas[:, :2].clone())
    m.other_betas = nn.Parameter(betas[:, 2:].clone())
    m.global_pose = nn.Parameter(pose[:, :3].clone())
    m.body_pose = nn.Parameter(pose[:, 3:66].clone())
    m.hand_pose = nn.Parameter(pose[:, 66:].clone())
    m.trans = nn.Parameter(trans.clone())
    m.offsets = nn.Parameter(torch.zeros(batch_sz, buffers["weights"].shape[0], 3))
    m.betas = torch.cat([m.top_betas, m.other_betas], 1)
    m.pose = torch.cat([m.global_pose, m.body_pose, m.hand_pose], 1)
    m.faces, m.gender = None, "male"
    m.smpl = load_smpl_layer(buffers)
    m.body25_reg_torch, m.face_reg_torch, m.hand_reg_torch = regs
    return m


@contextlib.contextmanager
def cpu_priors():
    """compute_prior_loss (recon/recon_fit_base.py:522-535) moves the prior tensors to the GPU
    unconditionally (th_smpl_prior.py:27-28 `.cuda()`, th_hand_prior.py:50 device='cuda:0').  In this
    GPU-less container run the same code with `.cuda()` a no-op and HandPrior defaulting to the CPU."""
    install()
    import functools
    import torch
    with ref_cwd():
        import recon.recon_fit_base as rfb
    old_cuda, old_hp = torch.Tensor.cuda, rfb.HandPrior
    torch.Tensor.cuda = lambda self, *a, **k: self
    rfb.HandPrior = functools.partial(old_hp, device="cpu")
    try:
        yield
    finally:
        torch.Tensor.cuda, rfb.HandPrior = old_cuda, old_hp
