"""-m gpu: drop-in networks, VGG loss and train_step (CUDA path through the C ABI) vs the CPU oracle."""
import pytest

pytestmark = pytest.mark.gpu


def _names():
    import net_cases as C
    return list(C.CASES.keys())


@pytest.mark.parametrize("name", _names())
def test_net_case(name):
    import net_cases as C
    details, ok = C.CASES[name]()
    assert ok, f"{name}: {details}"
