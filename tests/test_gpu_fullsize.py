"""-m gpu: properties at BASELINE.json's full sizes + a 512x512 forward against the oracle."""
import pytest

pytestmark = pytest.mark.gpu


def _names():
    import fullsize_cases as C
    return list(C.CASES.keys())


@pytest.mark.parametrize("name", _names())
def test_fullsize_case(name):
    import fullsize_cases as C
    res, ok = C.CASES[name]()
    assert ok, f"{name}: {res}"
