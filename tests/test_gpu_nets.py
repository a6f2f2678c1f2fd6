"""-m gpu: drop-in networks, VGG loss and train_step (CUDA path through the C ABI) vs the CPU oracle
(plain fp32 and bf16-storage-emulating). The measured margins of every case are appended to
gpurun_out/parity.jsonl (copied to profiles/parity_r2.json for the record)."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _names():
    import net_cases as C
    return list(C.CASES.keys())


def _fmt(v):
    if isinstance(v, float):
        return float(f"{v:.6g}")
    if isinstance(v, (tuple, list)):
        return [_fmt(x) for x in v]
    if isinstance(v, dict):
        return {k: _fmt(x) for k, x in v.items()}
    return v


@pytest.mark.parametrize("name", _names())
def test_net_case(name):
    import net_cases as C
    details, ok = C.CASES[name]()
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "parity.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, "ok": bool(ok), **_fmt(details)}) + "\n")
    except OSError:
        pass
    assert ok, f"{name}: {details}"
