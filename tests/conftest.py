"""pytest configuration: the `gpu` marker and common helpers.

`-m "not gpu"` : oracle vs golden vectors, host logic, CPU emulation of the FFT sweeps, C-ABI
                 load/symbol checks (no compute calls).
`-m gpu`       : parity tests proper -- the CUDA path, called through the C ABI, against the
                 oracle and the committed golden fixtures.
"""
import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
