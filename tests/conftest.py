import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names(pinned_only=True):
    with open(os.path.join(GOLDEN_DIR, "MANIFEST.json")) as fh:
        m = json.load(fh)
    return list(m["pinned"]) + ([] if pinned_only else list(m["unpinned"]))


def golden_tolerance(name):
    """kcal/mol within which the fixture reproduced the published OPTIMOL total when it was generated: 1e-5, or
    1e-4 for the molecules listed under "tolerance" (tests/golden/make_golden.py explains the two levels)."""
    with open(os.path.join(GOLDEN_DIR, "MANIFEST.json")) as fh:
        return float(json.load(fh).get("tolerance", {}).get(name, 1e-5))


def load_golden(name):
    from jmolecularenergy_b200.system import MolecularSystem
    return MolecularSystem.load(os.path.join(GOLDEN_DIR, f"{name}.json"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o
