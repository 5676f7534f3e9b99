"""pytest plumbing: registers the `gpu` marker, loads the hyphenated package directory as
module `sdrjfm_b200`, and builds the CPU checkers (test infrastructure) on demand."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_package():
    if "sdrjfm_b200" in sys.modules:
        return sys.modules["sdrjfm_b200"]
    pkg_dir = os.path.join(ROOT, "sdr-j-fm_b200")
    spec = importlib.util.spec_from_file_location(
        "sdrjfm_b200", os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["sdrjfm_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


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
def pkg():
    mod = load_package()
    mod.build()            # incremental (make); nvcc cross-compiles without a GPU
    return mod


@pytest.fixture(scope="session")
def signals(pkg):
    import importlib
    return importlib.import_module("sdrjfm_b200.signals")


@pytest.fixture(scope="session")
def chainlib():
    from oracle import chainlib as cl
    cl.build(("oracle", "ref"))
    return cl


@pytest.fixture(scope="session")
def ref_available(chainlib):
    return chainlib.available("ref")
