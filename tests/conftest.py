"""pytest configuration: markers, library fixtures.

* tests marked ``gpu`` call the product library (scbonita_b200/libscbonita_b200.so) on a real B200;
* everything else runs on CPU: the oracle against golden vectors and the reference build, host logic,
  the C-ABI symbol table, and the *same kernel sources* under the host-thread emulator (tests/emu).
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emu_lib():
    """Builds (if stale) and returns the emulator build of the kernels + host API."""
    emu_dir = os.path.join(ROOT, "tests", "emu")
    subprocess.check_call(["make", "-C", emu_dir, "all"], stdout=subprocess.DEVNULL)
    return os.path.join(emu_dir, "_build", "libscbonita_emu.so")


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail loudly when it is missing or no device is usable."""
    from scbonita_b200 import _lib
    lib = _lib.load()
    assert lib.scb_device_count() >= 1, "no CUDA device: GPU tests cannot run (there is no CPU fallback)"
    return _lib.DEFAULT_PATH


@pytest.fixture(scope="session", autouse=True)
def _oracle_built():
    import oracle
    oracle.build(ref=os.path.exists(oracle.REF_SRC), quiet=True)
