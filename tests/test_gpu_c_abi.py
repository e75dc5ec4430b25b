"""The C-ABI driven from plain C (no Python, no torch in the process): tests/c_abi_driver.c loads a
network, replicates the handle on every visible GPU and checks multi-GPU == single-GPU bitwise."""
import os
import subprocess

import pytest

from conftest import IC_GARROD, NET_A, ROOT

pytestmark = pytest.mark.gpu


def test_c_driver_runs_all_gpus_through_the_c_abi(rb, tmp_path):
    exe = str(tmp_path / "c_abi_driver")
    libdir = os.path.join(ROOT, "rac-2d_b200")
    subprocess.check_call(["gcc", "-O1", "-w", os.path.join(ROOT, "tests", "c_abi_driver.c"), "-I" + os.path.join(ROOT, "include"),
                           "-L" + libdir, "-lracg", "-lm", "-Wl,-rpath," + libdir, "-o", exe])
    out = subprocess.run([exe, NET_A, IC_GARROD, "192"], capture_output=True, text=True, timeout=600)
    print(out.stdout, out.stderr)
    assert out.returncode == 0 and "C_ABI_DRIVER OK" in out.stdout, out.stdout + out.stderr
