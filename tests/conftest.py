import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_pkg():
    """The package directory is named `modecouplingtheory.jl_b200` (a dot in the name), so import it by path."""
    if "mct_b200" in sys.modules:
        return sys.modules["mct_b200"]
    pkg_dir = os.path.join(ROOT, "modecouplingtheory.jl_b200")
    spec = importlib.util.spec_from_file_location("mct_b200", os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mct_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()
