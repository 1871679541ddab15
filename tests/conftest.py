import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, "junction-tree_b200")):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The shared library is a build product (git-ignored).  It is rebuilt whenever the kernel
    or ABI sources differ from the ones it was built from (content hash, `csrc/.build_hash`), so
    the suite never runs against a stale binary; nvcc cross-compiles without a GPU."""
    import __graft_entry__ as entry
    entry.ensure_built(quiet=True)


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        have_cuda = False
    if have_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_sessionfinish(session, exitstatus):
    """Max per-entry relative error of every labelled comparison (helpers.assert_close) ->
    gpurun_out/parity_report.json, when that scratch directory exists (GPU box runs)."""
    import json
    try:
        import helpers
    except Exception:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    if helpers.PARITY_REPORT and os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_report.json"), "w") as fh:
            json.dump(dict(sorted(helpers.PARITY_REPORT.items())), fh, indent=1)
