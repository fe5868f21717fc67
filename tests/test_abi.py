"""CPU checks of the drop-in boundary: libchore_b200.so builds/loads, exports every symbol that
include/chore_b200.h declares, and the product path fails loudly without a GPU (no compute)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "chore_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chore_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from chore_b200 import build
    return build.build()


def test_header_symbols_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/chore_b200.h but not exported"


def test_binding_covers_header(lib_path):
    from chore_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load_library()
    assert lib.chore_abi_version() == 3
    assert lib.chore_launch_count() >= 0


def test_sass_is_sm100a(lib_path):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_product_path_fails_loudly_without_gpu(lib_path):
    import chore_b200
    with pytest.raises(chore_b200.ChoreError):
        chore_b200.get_handle("cuda:0")
    net = chore_b200.CHORE()
    with pytest.raises(Exception):
        net.filter(torch.zeros(1, 5, 64, 64))


def test_null_arguments_are_rejected(lib_path):
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    from chore_b200 import _lib
    lib = _lib.load_library()
    assert lib.chore_query_fwd(None, None, None, 0, 0, None, None, 0, 0, 15, None, None, None, None, None, None) == 1
    assert b"null handle" in lib.chore_last_error()
    assert lib.chore_encode(None, None, 1, 512, 512, None, None, None, None) == 1
    assert lib.chore_lbs_fwd(None, None, None, None, None, 1, None, None, None, None, None) == 1
