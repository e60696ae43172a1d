"""The C-ABI library loads on a GPU-less host and exports every symbol include/hologan_b200.h declares.
No compute calls here (argument validation only)."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from lightning_gan_zoo_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hologan_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    syms = declared_symbols()
    assert "hg_rotate_fwd" in syms and "hg_adain_act_bwd" in syms
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(re.findall(r"\bT (hg_[a-z0-9_]+)", out))
    for s in syms:
        assert s in exported, f"{s} declared in the header but not exported"
        assert hasattr(lib, s)
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert exported == set(syms), f"exported but undeclared: {exported - set(syms)}"


def test_abi_version():
    assert _lib.load().hg_abi_version() == 2


def test_argument_validation_without_gpu():
    lib = _lib.load()
    n = ctypes.c_void_p(0)
    assert lib.hg_rotate_fwd(n, n, n, n, n, 1, 1, 16, 0, 0, 0, 0, n) == -1
    assert b"null pointer" in lib.hg_last_error()
    p = ctypes.c_void_p(16)
    assert lib.hg_rotate_fwd(p, p, p, n, n, 1, 1, 17, 0, 0, 0, 0, n) == -2      # size not 8/16/32
    assert lib.hg_rotate_fwd(p, p, p, n, n, 1, 1, 16, 0, 0, 7, 0, n) == -1      # bad dtype
    assert lib.hg_rotate_bwd(p, p, p, n, 0, 0, 1, 16, 0, 0, 0, 0, n) == -1      # batch 0
    assert lib.hg_adain_act_fwd(p, p, p, p, p, p, 1, 1, 6, 6, 1, 1e-8, 0.0, 0, 0, n) == -2   # N % 4 != 0
    with pytest.raises(_lib.HologanB200Error):
        _lib.call("hg_rotate_bwd", n, n, n, n, 0, 1, 1, 16, 0, 0, 0, 0, n)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.HologanB200Error, match="no CPU / PyTorch fallback"):
        _lib.load()
