import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "qsimulator.jl_b200", "host")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_sessionstart(session):
    """The tests call the CUDA library through its C ABI (problem assembly and evaluation work without a GPU):
    build it when the checkout has no library yet.  nvcc cross-compiles for sm_100a on a CPU-only box."""
    lib = os.path.join(ROOT, "qsimulator.jl_b200", "lib", "libqsim_b200.so")
    if not os.path.exists(lib):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
