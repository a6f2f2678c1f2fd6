"""-m gpu: tcgen05 implicit-GEMM entry points (through the C ABI) vs plain PyTorch fp32."""
import pytest

pytestmark = pytest.mark.gpu


def _names():
    import igemm_cases as C
    return list(C.CASES.keys())


@pytest.mark.parametrize("name", _names())
def test_igemm_case(name):
    import igemm_cases as C
    err, tol = C.CASES[name]()
    assert err <= tol, f"{name}: rel err {err:.3e} > {tol:.1e}"
