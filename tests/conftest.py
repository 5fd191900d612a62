import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Likelihood parity as measured: worst relative error per comparison and how many elements needed the
    absolute floor of tests/_common.py (north_star asks for 1e-5 relative; the floor covers few-ulp erfc
    differences through upper - lower).  Also written to gpurun_out/lik_error_report.json when that exists."""
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from _common import LIK_ATOL, LIK_RTOL, LIK_STATS
    except Exception:
        return
    if not LIK_STATS:
        return
    tests = {}
    for test, what, size, rel, floor_n, excess in LIK_STATS:
        t = tests.setdefault(test, [0, 0.0, 0, -1.0])
        t[0] += size
        t[1] = max(t[1], rel)
        t[2] += floor_n
        t[3] = max(t[3], excess)
    total = sum(t[0] for t in tests.values())
    floor_total = sum(t[2] for t in tests.values())
    tr = terminalreporter
    tr.write_sep("-", "likelihood parity (rtol %.0e, abs floor %.0e)" % (LIK_RTOL, LIK_ATOL))
    tr.write_line(f"{total} likelihoods compared in {len(tests)} tests; {floor_total} "
                  f"({100.0 * floor_total / max(total, 1):.4f} %) exceeded the purely relative bound and used the floor; "
                  f"worst relative error {max(t[1] for t in tests.values()):.3e}, "
                  f"worst excess over rtol*|ref| {max(t[3] for t in tests.values()):.3e}")
    worst = sorted(tests.items(), key=lambda kv: -kv[1][1])[:8]
    for name, (size, rel, floor_n, excess) in worst:
        tr.write_line(f"  {name[-70:]:70s} n={size:9d} max_rel={rel:.3e} floor_used={floor_n:7d} max_excess={excess:.2e}")
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        import json
        with open(os.path.join(out_dir, "lik_error_report.json"), "w") as f:
            json.dump({"rtol": LIK_RTOL, "atol": LIK_ATOL, "compared": total, "floor_used": floor_total,
                       "tests": {k: {"n": v[0], "max_rel": v[1], "floor_used": v[2], "max_excess_over_rel": v[3]}
                                 for k, v in tests.items()}}, f, indent=1)
