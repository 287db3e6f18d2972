"""Host emulation of the CUDA library — TEST INFRASTRUCTURE ONLY.

``lcsim_b200/csrc/lcsim_b200.cu`` compiled by g++ with ``-DLCSIM_HOSTEMU``: device memory becomes
host memory and a kernel launch becomes a loop over (scenario, phase, thread).  It lets the CPU test
tier (``-m "not gpu"``) drive the very same kernel text and C ABI against the oracle and the golden
vectors on a machine without a GPU.  The package (``lcsim_b200/``) never imports this module and has
no way to reach it; the product path is the sm_100a build and fails loudly without it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from lcsim_b200 import native

_HERE = os.path.dirname(os.path.abspath(__file__))
EMU_PATH = os.path.join(_HERE, "liblcsim_hostemu.so")


def build(force: bool = False) -> str:
    srcs = [native.SRC] + native.HDRS
    if not force and os.path.exists(EMU_PATH) and all(os.path.getmtime(EMU_PATH) >= os.path.getmtime(s) for s in srcs):
        return EMU_PATH
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-DLCSIM_HOSTEMU", "-x", "c++",
           native.SRC, "-o", EMU_PATH + ".tmp", "-lpthread", "-Wno-unknown-pragmas"]
    subprocess.check_call(cmd)
    os.replace(EMU_PATH + ".tmp", EMU_PATH)
    return EMU_PATH


class NumpyEmuBackend:
    name = "hostemu"
    device_index = 0

    def __init__(self):
        self.lib = native.bind(build())

    def view(self, ptr, shape, typestr):
        dt = np.dtype(typestr)
        n = int(np.prod(shape))
        buf = (C.c_char * (n * dt.itemsize)).from_address(int(ptr))
        return np.frombuffer(buf, dtype=dt).reshape(shape)

    def stream(self):
        return None

    def to_device(self, arr, dtype):
        return np.ascontiguousarray(arr, dtype=dtype)

    def empty(self, shape, dtype):
        return np.zeros(tuple(shape), dtype=dtype)

    def ptr(self, x):
        return x.ctypes.data

    def to_numpy(self, x):
        return np.asarray(x)

    def sync(self):
        pass
