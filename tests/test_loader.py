"""Native scenario loader / packer (lcsim_b200_load_dirs): bit-identical to lcsim_b200/packer.py — which
the golden tests pin to the real reference — on synthetic WOMD-shaped scenarios written as real map.pb / agents.pb,
and on the two bundled scenarios of the reference when its tree is present (build container)."""
import os
import time

import numpy as np
import pytest

from lcsim_b200 import native, packer, synth

FIELDS = [n for n, _ in native.STATIC_FIELDS] + ["agent_id"]
BUNDLED = "/root/reference/examples/data/waymo"


def _lib():
    """the loader is host code: the test tier without a GPU reaches it through the host-emulation build of the same source"""
    try:
        import torch

        if torch.cuda.is_available():
            return native.load()
    except Exception:
        pass
    from hostemu import NumpyEmuBackend

    return NumpyEmuBackend().lib


def _compare(dirs, **kw):
    want = packer.pack_dirs(dirs, **kw)
    scs = packer.scenarios_from_dirs(dirs)
    got, centres, junction = native.load_dirs(_lib(), dirs, **kw)
    assert (got.S, got.A, got.T, got.E) == (want.S, want.A, want.T, want.E)
    for f in FIELDS:
        a, b = getattr(got, f), getattr(want, f)
        assert a.dtype == b.dtype and a.shape == b.shape, f
        np.testing.assert_array_equal(a.view(np.uint8), b.view(np.uint8), err_msg=f)  # bit for bit (NaN-safe)
    for s, sc in enumerate(scs):
        c = packer.polyline_centres(sc.lanes)
        assert centres[s].shape == c.shape
        np.testing.assert_array_equal(centres[s][:, :2], c[:, :2], err_msg=f"polyline centres of scenario {s}")
        # heading = float32 atan2 of the mean direction: numpy (SIMD loops), torch (the reference, Sleef) and libm agree
        # to one float32 ulp, not to the bit — the same tolerance tests/test_obs.py gives the feature that uses it
        np.testing.assert_allclose(centres[s][:, 2], c[:, 2], rtol=0, atol=2.5e-7)
        assert junction[s] == sc.junction_id
    return got


def test_loader_matches_packer_on_synthetic_pb(tmp_path):
    dirs = []
    for seed, n in ((0, 128), (1, 64), (3, 96), (11, 17)):
        d = str(tmp_path / f"s{seed}")
        synth.write_scenario(d, seed, n)
        dirs.append(d)
    _compare(dirs)
    _compare(dirs[1:3], start=0.3, interval=0.05, total_steps=40)


@pytest.mark.skipif(not os.path.isdir(BUNDLED), reason="the reference's bundled scenarios are only in the build container")
def test_loader_matches_packer_on_bundled_womd():
    dirs = [os.path.join(BUNDLED, d) for d in sorted(os.listdir(BUNDLED))]
    got = _compare(dirs)
    assert got.n_agents.tolist() == [45, 61] or got.n_agents.tolist() == [61, 45]


def test_loader_errors(tmp_path):
    with pytest.raises(native.NativeLibraryError, match="map.pb"):
        native.load_dirs(_lib(), [str(tmp_path / "nothing")])
    d = str(tmp_path / "bad")
    synth.write_scenario(d, 5, 8)
    with open(os.path.join(d, "agents.pb"), "wb") as f:
        f.write(b"\x0a\xff\xff\xff")  # truncated length-delimited field
    with pytest.raises(native.NativeLibraryError, match="agents.pb"):
        native.load_dirs(_lib(), [d])


def test_loader_is_fast(tmp_path):
    """64 scenarios: the native loader beats the Python packer by a wide margin (init of a 1024-scenario batch is
    reported by bench.py as init_s)"""
    dirs = []
    for seed in range(16):
        d = str(tmp_path / f"s{seed}")
        synth.write_scenario(d, 100 + seed, 128)
        dirs.append(d)
    dirs = dirs * 4
    t0 = time.perf_counter()
    native.load_dirs(_lib(), dirs)
    t_native = time.perf_counter() - t0
    t0 = time.perf_counter()
    packer.pack_dirs(dirs[:8])
    t_python = (time.perf_counter() - t0) * 8
    assert t_native < 0.5 * t_python, (t_native, t_python)
