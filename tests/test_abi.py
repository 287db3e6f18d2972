"""CPU tier: the sm_100a library builds, loads and exports exactly what include/lcsim_b200.h declares.
No compute entry point is called here (no GPU)."""
import ctypes as C
import os
import re

import pytest

from lcsim_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "lcsim_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lcsim_b200_[a-z_0-9]+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _declared() == sorted(native.EXPORTS)


def test_cuda_library_builds_and_exports_every_symbol():
    path = native.build()
    lib = C.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), name
    assert native.bind(path).lcsim_b200_abi_version() == native.ABI_VERSION


def test_create_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(native.NativeLibraryError):
        native.TorchCudaBackend(0)
    lib = native.bind(native.build())
    d = native.BatchDesc()
    d.abi_version, d.S, d.A, d.T, d.E = native.ABI_VERSION, 1, 32, 8, 32
    h = C.c_void_p()
    assert lib.lcsim_b200_create(C.byref(d), C.byref(h)) == -2  # LCSIM_B200_ERR_CUDA
    assert not h


def test_product_never_imports_oracle_or_hostemu():
    pkg = os.path.join(ROOT, "lcsim_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn
            assert "hostemu" not in src, fn
