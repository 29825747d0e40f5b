import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# ---- measured parity errors: tests call record_err(); the numbers are printed in the terminal summary
# (also with -q) and written to gpurun_out/parity_errors.json so that they can be kept under profiles/
_ERRS = {}


def record_err(name: str, err: float, tol: float, note: str = "") -> None:
    _ERRS[name] = {"max_err": float(err), "tol": float(tol), "note": note}


@pytest.fixture
def record():
    return record_err


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    if not _ERRS:
        return
    terminalreporter.write_line("measured parity errors (max |got - ref| vs the stated tolerance):")
    for k in sorted(_ERRS):
        v = _ERRS[k]
        terminalreporter.write_line(f"  {k}: {v['max_err']:.4g} (tol {v['tol']:.4g}) {v['note']}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        import json
        with open(os.path.join(out_dir, "parity_errors.json"), "w") as f:
            json.dump(_ERRS, f, indent=1, sort_keys=True)
    except OSError:
        pass
