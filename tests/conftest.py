import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def parity_log(request):
    """Appends one line per comparison to gpurun_out/parity_report.txt (scratch; the copy under profiles/ is
    committed by hand after a GPU run) so the measured rel-L2 of every parity test can be read, not just pass/fail."""
    path = os.path.join(ROOT, "gpurun_out", "parity_report.txt")

    def log(**values):
        try:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "a") as f:
                f.write(request.node.name + " " + " ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}"
                                                           for k, v in values.items()) + "\n")
        except OSError:
            pass
    return log
