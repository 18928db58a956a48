"""The C ABI consumed from plain C (tests/c_driver/abi_driver.c): the call sequence of INTEGRATION.md with no
Python between the caller and the library.  CPU: problem assembly, evaluation and the no-device refusal;
-m gpu: a Rabi sweep through qsim_unitary_propagator checked against the rotating-wave answer."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "qsimulator.jl_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "abi_driver")
    cmd = ["gcc", "-std=c99", "-D_GNU_SOURCE", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_driver", "abi_driver.c"), "-o", exe,
           "-L", LIBDIR, "-lqsim_b200", "-lm", f"-Wl,-rpath,{LIBDIR}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe


def _run(exe, mode):
    r = subprocess.run([exe, mode], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASSED" in r.stdout and "FAIL" not in r.stdout.replace("FAILED", ""), r.stdout + r.stderr
    return r.stdout


def test_c_driver_host(tmp_path):
    out = _run(_build(tmp_path), "host")
    assert "H(t) = H0" in out


@pytest.mark.gpu
def test_c_driver_gpu(tmp_path):
    out = _run(_build(tmp_path), "gpu")
    assert "Rabi population" in out
