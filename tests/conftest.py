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


def pytest_collection_modifyitems(config, items):
    """`gpu` tests skip (instead of failing) on a box without a usable CUDA device."""
    import pytest
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    from qsim_b200 import _ffi
    if not _ffi.device_available():   # only the library's own QSIM_ERR_NO_DEVICE skips; any other failure raises
        skip = pytest.mark.skip(reason="no usable CUDA device (qsim_create returned QSIM_ERR_NO_DEVICE)")
        for it in gpu_items:
            it.add_marker(skip)
