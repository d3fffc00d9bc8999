import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import model

    model.build_lib()
    return model


class ParityBounds:
    """Statistical parity assertions with bounds tied to measurements (every bound <= 3x the measured value).

    tests/parity_bounds.json holds, per named statistic, the value measured on a B200 and the limit derived from it by
    tests/tools/update_bounds.py (max: 3 x measured; spearman: 1 - 3 (1 - measured); min_int: the measured minimum).  check()
    asserts against that limit -- or against `default` for a statistic that has not been measured yet -- and appends what it
    measured to gpurun_out/parity_measured.jsonl so the next update can tighten it."""

    def __init__(self):
        import json
        self.path = os.path.join(ROOT, "tests", "parity_bounds.json")
        self.table = json.load(open(self.path)) if os.path.exists(self.path) else {}
        self.out = os.path.join(ROOT, "gpurun_out", "parity_measured.jsonl")

    def check(self, name: str, value: float, default: float, kind: str = "max"):
        import json
        value = float(value)
        limit = self.table.get(name, {}).get("limit", default)
        try:
            os.makedirs(os.path.dirname(self.out), exist_ok=True)
            with open(self.out, "a") as f:
                f.write(json.dumps({"name": name, "kind": kind, "value": value, "limit": limit}) + "\n")
        except OSError:
            pass
        if kind == "max":
            assert value <= limit, f"{name}: measured {value} > limit {limit}"
        else:  # "spearman" and "min_int": larger is better
            assert value >= limit, f"{name}: measured {value} < limit {limit}"


@pytest.fixture(scope="session")
def bounds():
    return ParityBounds()
