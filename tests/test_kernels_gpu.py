"""GPU parity of every libcidb200 kernel against plain PyTorch fp32 references (tests/kernel_checks.py)."""
import pytest

from tests.kernel_checks import CHECKS, run


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CHECKS))
def test_kernel(name):
    r = run(name)
    assert r["ok"], r
