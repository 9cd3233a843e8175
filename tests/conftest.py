import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import vstab_loader  # noqa: E402

vstab_loader.load()

REFERENCE_DIR = "/root/reference"
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir(os.path.join(REFERENCE_DIR, "nodes"))
    skip_ref = pytest.mark.skip(reason="/root/reference not present (GPU box); golden fixtures cover it")
    for item in items:
        if "reference" in item.keywords and not have_ref:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def handle():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vstab_b200 import _native

    return _native.get_handle(torch.device("cuda", 0))


@pytest.fixture(scope="session")
def reference_nodes():
    """The reference's node modules imported under ComfyUI stubs (build container only)."""
    from tests import ref_import

    return ref_import.load_reference()
