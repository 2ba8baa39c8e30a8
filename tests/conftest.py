import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def bpt():
    import bifrost3d_b200 as b
    ctx = b.Bpt(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def bpt8():
    """A context that builds the compressed eight-wide nodes whatever the size of the scene (the default takes them from
    131 072 triangles on; BPT_CW is read by bpt_create)."""
    import os
    import bifrost3d_b200 as b
    before = os.environ.get("BPT_CW")
    os.environ["BPT_CW"] = "1"
    try:
        ctx = b.Bpt(0)
    finally:
        if before is None:
            del os.environ["BPT_CW"]
        else:
            os.environ["BPT_CW"] = before
    yield ctx
    ctx.close()


@pytest.fixture(params=["node_format_by_size", "eight_wide_nodes"])
def tracer(request):
    """Both node formats for the tests whose scenes are below the size at which the build switches to the eight-wide nodes."""
    return request.getfixturevalue("bpt" if request.param == "node_format_by_size" else "bpt8")


@pytest.fixture(scope="session")
def ref():
    from tests import oracle_lib
    return oracle_lib.load()
