"""pytest configuration: registers the ``gpu`` marker and makes the repo importable.

`python -m pytest tests -m "not gpu"` runs the CPU suite (oracle pinning, host logic, C-ABI symbol
checks); `-m gpu` runs the parity tests proper on a B200 through libacgpu's C ABI.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")
