"""GPU parity of every kernel of librestoragen.so against a PyTorch fp32 reference of the same op (see kernel_cases.py)."""
import pytest

import kernel_cases as kc


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(kc.CASES))
def test_kernel_case(name):
    err, tol = kc.CASES[name]()
    assert err <= tol, f"{name}: error {err:.3e} exceeds tolerance {tol:.1e}"
