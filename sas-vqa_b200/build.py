"""Builds libsasvqa_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# development A/B builds: SASVQA_LIB_SUFFIX=_x SASVQA_DEFINES="-DFOO=1" -> libsasvqa_b200_x.so
SUFFIX = os.environ.get("SASVQA_LIB_SUFFIX", "")
LIB = os.path.join(HERE, f"libsasvqa_b200{SUFFIX}.so")
SOURCES = ["capi.cu", "encoder.cu", "gemm_tcgen05.cu", "attention.cu", "attention_tcgen05.cu", "attention_git_tcgen05.cu", "elementwise.cu", "select.cu", "resize.cu", "nvdec.cu", "scorer.cu", "git_decoder.cu"]
# test-only check kernels (CUDA-core GEMM, mma.sync encoder attention): a separate library the product never loads
TEST_LIB = os.path.join(HERE, "libsasvqa_b200_test.so")
TEST_SOURCES = ["check/capi_check.cu", "check/gemm_simt.cu", "check/attention_mma_check.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(lib: str) -> bool:
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(root, f) for root, _, files in os.walk(CSRC) for f in files]
    deps.append(os.path.join(HERE, "..", "include", "sasvqa.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(nvcc: str, sources, lib: str, verbose: bool) -> None:
    objs, procs = [], []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    defines = os.environ.get("SASVQA_DEFINES", "").split()
    for src in sources:
        obj = os.path.join(HERE, "build", os.path.basename(src).replace(".cu", f"{SUFFIX}.o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"---- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", lib, *objs, "-lcudart", "-ldl"])


def build(force: bool = False, verbose: bool = False) -> str:
    """Builds the product library and (unless this is a suffixed A/B build) the test-only check library."""
    nvcc = _nvcc()
    if force or _stale(LIB):
        _compile(nvcc, SOURCES, LIB, verbose)
    if not SUFFIX and (force or _stale(TEST_LIB)):
        _compile(nvcc, TEST_SOURCES, TEST_LIB, verbose)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
