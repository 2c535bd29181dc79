import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture()
def abi_check_exe(tmp_path):
    """gcc -std=c99 over tests/c_abi/abi_check.c: a client of the C ABI that knows only include/sasvqa.h."""
    import subprocess
    lib_dir = os.path.join(ROOT, "sas-vqa_b200")
    cuda_lib = "/usr/local/cuda/lib64"
    exe = str(tmp_path / "abi_check")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "abi_check.c"), "-L", lib_dir, "-lsasvqa_b200", "-L", cuda_lib, "-lcudart",
           f"-Wl,-rpath,{lib_dir}", f"-Wl,-rpath,{cuda_lib}", "-o", exe]
    done = subprocess.run(cmd, capture_output=True, text=True)
    assert done.returncode == 0, done.stderr
    return exe
