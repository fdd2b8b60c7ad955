import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
PKG_DIR = REPO / "knn-for-homology_b200"
for p in (str(REPO), str(PKG_DIR)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def expected():
    import numpy as np

    return dict(np.load(GOLDEN / "expected.npz"))


def pytest_terminal_summary(terminalreporter):
    """Totals of every oracle parity check of the run: positions compared, tie-excused positions, worst relative
    distance error (BASELINE.json's bar: ids identical except at ties, distances within 1e-5 relative)."""
    try:
        from oracle.parity import RUN_LOG
    except Exception:
        return
    if RUN_LOG:
        terminalreporter.write_line(
            "oracle parity: %d checks, %d positions, %d tie-excused (%d at the k-boundary), max_rel_err_D = %.3g"
            % (len(RUN_LOG), sum(s["positions"] for s in RUN_LOG), sum(s["excused"] for s in RUN_LOG),
               sum(s["boundary_swaps"] for s in RUN_LOG), max(s["max_rel_err_D"] for s in RUN_LOG)))
