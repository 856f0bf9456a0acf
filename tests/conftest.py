import importlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
HOSTSIM_SO = os.path.join(ROOT, "tests", "hostsim", "_build", "libwmbus_hostsim.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module("rtl-wmbus_b200")


@pytest.fixture(scope="session")
def orc_mod():
    import orc
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def golden_lines():
    return json.load(open(os.path.join(GOLDEN, "golden_lines.json")))


@pytest.fixture(scope="session")
def hostsim_lib(pkg):
    """CPU simulation of the device path -- TEST INFRASTRUCTURE, only for `not gpu` tests of the
    host logic (tests/hostsim/hostsim_cuda.h)."""
    src = os.path.join(ROOT, "rtl-wmbus_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src))
    if not os.path.exists(HOSTSIM_SO) or os.path.getmtime(HOSTSIM_SO) < newest:
        subprocess.run([os.path.join(ROOT, "tests", "hostsim", "build.sh")], check=True, capture_output=True)
    return pkg.load_library(HOSTSIM_SO)


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    """The product library on a real GPU.  No fallback: missing library or device is an error."""
    assert has_gpu(), "gpu tests need a CUDA device"
    return pkg.load_library()


def load_fixture(name):
    return np.fromfile(os.path.join(GOLDEN, name), np.uint8)
