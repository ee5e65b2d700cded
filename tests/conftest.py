import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def ctx():
    """One pandrs_b200 device context for the whole GPU session."""
    import pandrs_b200 as pb
    c = pb.Context(device=0)
    yield c
    c.close()


@pytest.fixture
def fctx(ctx):
    """The shared context as the context of the frame mirror (pandrs_b200.frame) for ONE test: whatever was there before comes back."""
    import pandrs_b200.frame as fr
    prev = fr._CTX
    fr.set_context(ctx)
    yield ctx
    fr.set_context(prev)


def pytest_sessionfinish(session, exitstatus):
    """The measured f64 errors of the CUDA path (tests/_util.py: compare_groupby) - printed and kept, so that the 1e-12
    claim is a number on record: per op [max |gpu - reference| / scale, max |gpu - exact| / |exact|]."""
    _util = sys.modules.get("_util")
    if _util is None or not getattr(_util, "MEASURED", None):
        return
    import json
    names = {0: "sum", 1: "mean", 5: "std", 6: "var"}
    rec = {names.get(k, str(k)): {"max_err_vs_reference": v[0], "max_rel_err_vs_exact": v[1]} for k, v in sorted(_util.MEASURED.items())}
    print("\n[parity] measured f64 errors of the CUDA path:", json.dumps(rec))
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_f64_errors.json"), "w") as f:
            json.dump(rec, f, indent=1)
    except OSError:
        pass
