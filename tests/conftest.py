import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    from oracle_lib import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref not built (needs /root/reference; build container only)")
    return Reference()


@pytest.fixture(scope="session")
def ghlib():
    """The product library. Host-side entry points work without a GPU; device ones return GH_ERR_CUDA."""
    import golden_huffman_b200 as gh
    if not os.path.exists(gh.DEFAULT_LIB):
        import importlib.util
        spec = importlib.util.spec_from_file_location("gh_build", os.path.join(ROOT, "golden-huffman_b200", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return gh.GhLib()


@pytest.fixture(scope="session")
def codec(ghlib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import golden_huffman_b200 as gh
    return gh.Codec(ghlib)
