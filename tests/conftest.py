"""pytest configuration: `gpu` marker (needs a real B200 + the built libchore_b200.so) and
shared helpers.  `-m "not gpu"` must pass on a CPU-only box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def rel_err(a, b):
    """max |a-b| / (|b| + rms(b)): the relative-error measure used for the 1e-4 field bar."""
    import torch
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    scale = b.abs() + b.pow(2).mean().sqrt() + 1e-30
    return ((a - b).abs() / scale).max().item()


def golden_smpl_assets(g):
    """Landmark regressors (scipy CSR, (L,V)) and priors stored in fit_smpl_full.npz."""
    import scipy.sparse as sp
    regs = [sp.csr_matrix((g[f"reg{i}_data"], g[f"reg{i}_indices"], g[f"reg{i}_indptr"]), shape=(L, 6890))
            for i, L in enumerate((25, 70, 42))]
    pri = {k[len("prior_"):]: g[k] for k in g if k.startswith("prior_")}
    return regs, pri


def pure_rel_err(a, b, floor_frac=0.05):
    """max |a-b| / |b| over the elements with |b| > floor_frac * rms(b): the plain relative error north_star quotes,
    without rel_err's rms term in the denominator (elements near zero are excluded instead)."""
    import torch
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    keep = b.abs() > floor_frac * b.pow(2).mean().sqrt()
    return ((a - b).abs()[keep] / b.abs()[keep]).max().item() if bool(keep.any()) else 0.0


_REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def report(test, **values):
    """Append measured errors to gpurun_out/parity_report.jsonl (scratch; read after a GPU run)."""
    import json
    try:
        os.makedirs(os.path.dirname(_REPORT), exist_ok=True)
        with open(_REPORT, "a") as f:
            f.write(json.dumps({"test": test, **{k: float(v) for k, v in values.items()}}) + "\n")
    except OSError:
        pass
