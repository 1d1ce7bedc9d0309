import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The libraries are built in-tree by __graft_entry__.build() and are git-ignored: build them
    # when a fresh checkout runs the tests first (needs nvcc; never silently skipped).
    needed = [os.path.join(ROOT, "llm_instance_gateway_b200", "liblig.so"),
              os.path.join(ROOT, "llm_instance_gateway_b200", "liblig_host.so"),
              os.path.join(ROOT, "oracle", "liblig_oracle.so")]
    if not all(os.path.exists(p) for p in needed):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "go_filter_test_vectors.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.load()
    return binding
