"""-m gpu: uninitialised / out-of-bounds read detector (NaN-poisoned and NaN-guarded buffers), see uninit_cases.py."""
import pytest

pytestmark = pytest.mark.gpu


def _names():
    import uninit_cases as C
    return list(C.CASES.keys())


@pytest.mark.parametrize("name", _names())
def test_uninit_case(name):
    import uninit_cases as C
    details, ok = C.CASES[name]()
    assert ok, f"{name}: {details}"
