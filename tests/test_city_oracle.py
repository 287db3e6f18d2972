"""The LaneIDM restatement (oracle/city_oracle.py) against golden vectors of the REAL reference engine on
MOSS-shaped cities (oracle/gen_golden_city.py): every tick, integers bit-exact, floats 1e-9."""
import numpy as np
import pytest

from city_util import CityCase, case_names, check_against_ref
from oracle.city_oracle import CityOracle


@pytest.mark.parametrize("name", case_names())
def test_city_oracle_matches_reference(name):
    c = CityCase(name)
    o = CityOracle(c.static)

    def get(nm):
        if nm == "n_active":
            return len(o.active)
        if nm == "active":
            return np.asarray(o.active + [-1] * (o.N - len(o.active)))
        if nm == "lane_pbid":
            return c.static.lane_pbid[o.lane]
        if nm in ("x", "y"):
            return o.xy[:, 0 if nm == "x" else 1]
        return getattr(o, nm)

    check_against_ref(c.ref, 0, get)
    for k in range(c.n_ticks):
        o.step()
        check_against_ref(c.ref, k + 1, get)
        np.testing.assert_array_equal(o.is_lc, c.ref["is_lc"][k + 1], err_msg=f"tick {k + 1}: is_lc")
    assert o.lc_required == name.endswith("_lc")
    np.testing.assert_array_equal(o.is_lc, c.ref["is_lc"][-1])
    np.testing.assert_array_equal(o.ring_valid, c.ref["ring_valid"])
    np.testing.assert_allclose(o.ring, c.ref["ring"], rtol=0, atol=1e-9)
