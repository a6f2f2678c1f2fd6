"""-m gpu: bandwidth-bound entry points (through the C ABI) vs plain PyTorch fp32."""
import pytest

pytestmark = pytest.mark.gpu


def _names():
    import ops_cases as C
    return list(C.CASES.keys())


@pytest.mark.parametrize("name", _names())
def test_ops_case(name):
    import ops_cases as C
    err, tol = C.CASES[name]()
    assert err <= tol, f"{name}: err {err:.3e} > {tol:.1e}"
