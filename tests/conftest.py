import ctypes
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def sha16(t) -> str:
    a = t.detach().cpu().contiguous().numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]


def params_sha(p) -> str:
    h = hashlib.sha256()
    for k in sorted(p):
        h.update(k.encode())
        h.update(p[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def rel_err(a, b) -> float:
    """max|a-b| / max|b| -- the tolerance metric of SURVEY.md 8c."""
    a = a.detach().double().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(a).double()
    b = b.detach().double().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(b).double()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


@pytest.fixture(scope="session")
def c_oracle():
    """oracle/librotate_oracle.so (plain-C restatement), built on demand with oracle/Makefile."""
    path = os.path.join(ROOT, "oracle", "librotate_oracle.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(path)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.orc_rotate_coords.argtypes = [vp, vp, ci, ci]
    lib.orc_rotate_fwd.argtypes = [vp, vp, vp, ci, ci, ci]
    lib.orc_rotate_bwd.argtypes = [vp, vp, vp, ci, ci, ci]
    for f in (lib.orc_rotate_coords, lib.orc_rotate_fwd, lib.orc_rotate_bwd):
        f.restype = None
    return lib


def np_ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture
def hg_option():
    """Set tuning options of libhologan_b200.so (hg_set_option) for one test; previous values are restored."""
    from lightning_gan_zoo_b200 import _lib
    saved = {}

    def set_(name, value):
        old = _lib.set_option(name, value)
        saved.setdefault(name, old)

    yield set_
    for name, old in saved.items():
        _lib.set_option(name, old)
