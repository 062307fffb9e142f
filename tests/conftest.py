import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library is built in-tree (git-ignored); build it when a fresh checkout runs the tests before build()."""
    import importlib

    b = importlib.import_module("disentangled-subject-to-vid_b200._build")
    if not os.path.exists(b.LIB_PATH):
        b.build()
    return b.LIB_PATH


class ParityLog:
    """Every GPU parity test reports its measured errors here.  `check(key, err, ref=..., default=...)` appends one JSON line to
    gpurun_out/r02_parity.jsonl (so a GPU run leaves a record that tools/make_parity_table.py turns into profiles/r02_parity.md and
    tests/parity_bounds.json) and asserts err <= bound, where bound = the committed measured value x 1.5 when the key has one
    (tests/parity_bounds.json), else the looser `default` used before the first measurement."""

    def __init__(self):
        import json
        self.path = os.environ.get("S2V_PARITY_LOG", os.path.join(ROOT, "gpurun_out", "r02_parity.jsonl"))
        bp = os.path.join(ROOT, "tests", "parity_bounds.json")
        self.bounds = json.load(open(bp)) if os.path.exists(bp) else {}

    def check(self, key, err, *, default, ref=None, metric="rel_fro", **extra):
        import json
        bound = float(self.bounds.get(key, {}).get("bound", default))
        row = {"key": key, "metric": metric, "product": float(err), "reference_bf16": None if ref is None else float(ref),
               "bound": bound, **extra}
        try:
            os.makedirs(os.path.dirname(self.path), exist_ok=True)
            with open(self.path, "a") as f:
                f.write(json.dumps(row) + "\n")
        except OSError:
            pass
        print(f"parity {key}: product {err:.3e}" + (f"  reference-bf16 {ref:.3e}" if ref is not None else "") + f"  bound {bound:.3e}")
        assert err <= bound, (key, err, bound)


@pytest.fixture(scope="session")
def parity():
    return ParityLog()
