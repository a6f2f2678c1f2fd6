"""-m gpu: layer-wise teacher-forced parity of the generator (forward + backward) vs the bf16-storage-
emulating oracle; see tests/layerwise_cases.py. Margins are appended to gpurun_out/parity.jsonl."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _names():
    import layerwise_cases as C
    return list(C.CASES.keys())


@pytest.mark.parametrize("name", _names())
def test_layerwise_case(name):
    import layerwise_cases as C
    details, ok = C.CASES[name]()
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, "ok": bool(ok), **details}) + "\n")
    except OSError:
        pass
    assert ok, f"{name}: {details}"
