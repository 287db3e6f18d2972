"""CPU tier: the kernel text of lcsim_b200/csrc (compiled as plain C++ by tests/hostemu) driven through the
real C ABI against the reference golden vectors.  The GPU tier (test_gpu_parity.py) runs the same checks on
the sm_100a build."""
import pytest

from golden_util import Case, case_names
from hostemu import NumpyEmuBackend
from lcsim_b200.batch import BatchEngine
from parity_util import run_case_against_golden


@pytest.fixture(scope="module")
def emu():
    return NumpyEmuBackend()


@pytest.mark.parametrize("name", case_names())
def test_hostemu_matches_reference_golden(name, emu):
    c = Case(name)
    be = BatchEngine(c.static, reactive=c.reactive, no_static=c.no_static, allow_new_obj=c.allow_new_obj,
                     with_metrics=True, backend=emu)
    run_case_against_golden(c, be)
    be.close()
