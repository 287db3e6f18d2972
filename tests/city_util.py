"""Loading of the LaneIDM / MOSS-shaped golden cases (tests/golden_city, oracle/gen_golden_city.py)."""
import dataclasses
import glob
import os

import numpy as np

from lcsim_b200 import city

HERE = os.path.dirname(os.path.abspath(__file__))


def case_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden_city", "*.npz")))


class CityCase:
    def __init__(self, name):
        z = np.load(os.path.join(HERE, "golden_city", name + ".npz"))
        kw = {}
        for f in dataclasses.fields(city.CityStatic):
            v = z["st_" + f.name]
            kw[f.name] = v.item() if v.ndim == 0 else v
        self.static = city.CityStatic(**kw)
        self.ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
        self.n_ticks = len(self.ref["n_active"]) - 1


def check_against_ref(ref, k, get, tol=1e-9, lead=True):
    """get(name) -> numpy array of the implementation under test (agent-indexed unless noted), after tick k."""
    n = int(ref["n_active"][k])
    assert int(get("n_active")) == n, f"tick {k}: n_active"
    np.testing.assert_array_equal(get("active")[:n], ref["active"][k, :n], err_msg=f"tick {k}: active order")
    np.testing.assert_array_equal(get("status"), ref["status"][k], err_msg=f"tick {k}: status")
    np.testing.assert_array_equal(get("lane_pbid"), ref["lane"][k], err_msg=f"tick {k}: lane id")
    for nm in ("s", "v", "x", "y", "h"):
        np.testing.assert_allclose(get(nm), ref[nm][k], rtol=0, atol=tol, err_msg=f"tick {k}: {nm}")
    act = ref["active"][k, :n]
    if lead:
        np.testing.assert_array_equal(get("lead_valid")[act], ref["lead_valid"][k][act], err_msg=f"tick {k}: lead_valid")
        m = act[ref["lead_valid"][k][act] > 0]
        np.testing.assert_allclose(get("lead_dis")[m], ref["lead_dis"][k][m], rtol=0, atol=tol, err_msg=f"tick {k}: lead_dis")
        np.testing.assert_allclose(get("lead_v")[m], ref["lead_v"][k][m], rtol=0, atol=tol, err_msg=f"tick {k}: lead_v")
    np.testing.assert_array_equal(get("coll_count")[act], ref["coll"][k][act], err_msg=f"tick {k}: collision counts")
    np.testing.assert_array_equal(get("offroad")[act], ref["offroad"][k][act], err_msg=f"tick {k}: offroad")
