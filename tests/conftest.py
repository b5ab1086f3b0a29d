"""pytest configuration: markers, import paths, shared fixtures.

`-m "not gpu"`: oracle vs reference/golden vectors, host logic, ABI and symbol checks (no CUDA compute).
`-m gpu`      : parity of the CUDA path (through the C ABI) against the CPU oracle.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """Build missing native pieces once per session (CPU-only machines can build everything, nvcc cross-compiles)."""
    core = os.path.join(ROOT, "jurassic-gpu_b200", "lib", "libjurassic_b200.so")
    orc = os.path.join(ROOT, "oracle", "libjr_oracle.so")
    if not (os.path.exists(core) and os.path.exists(orc)):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def jr():
    _ensure_built()
    return importlib.import_module("jurassic-gpu_b200")


@pytest.fixture(scope="session")
def refdrv(jr):
    import refdrv as r
    return r


@pytest.fixture(scope="session")
def oracle(refdrv):
    return refdrv.Oracle()


@pytest.fixture(scope="session")
def gpu_ctx_factory(jr):
    made = []

    def make(device=0):
        c = jr.Context(device)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
