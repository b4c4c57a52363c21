"""Plain-C conformance of the C ABI (tests/abi_c/conformance.c): builds an engine, searches, shards, streams, provokes
every error status -- without Python in the loop.  CPU suite: compiles and links against libfacgpu.so and checks the
no-device behaviour (FAC_CUDA_ERROR, no CPU fallback) plus the pure host shard planner; GPU suite: the full program,
including the > 4 GiB cases (FAC_HAYSTACK_TOO_LARGE, absolute u64 stream offsets past 2^32)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "abi_c")
CSRC = os.path.join(ROOT, "fuzzy-aho-corasick-rs_b200", "csrc")
OUT = os.path.join(SRC, "_build")
CUDA_LIB = "/usr/local/cuda/lib64"


def build():
    import __graft_entry__ as ge
    ge.build_gpu_library()
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "conformance")
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", os.path.join(SRC, "conformance.c"), os.path.join(SRC, "facio.c"),
                           "-I", os.path.join(ROOT, "include"), "-L", CSRC, "-lfacgpu", "-L", CUDA_LIB, "-lcudart", "-lm",
                           "-Wl,-rpath," + CSRC, "-Wl,-rpath," + CUDA_LIB, "-o", exe])
    subprocess.check_call(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", "-shared", "-fPIC", os.path.join(SRC, "facio.c"),
                           "-I", os.path.join(ROOT, "include"), "-o", os.path.join(OUT, "libfacio.so")])
    return exe


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_c_program_builds_and_fails_loudly_without_a_device():
    exe = build()
    if _has_gpu():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([exe, "--no-device"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout and "FAIL" not in out.stdout


@pytest.mark.gpu
def test_c_conformance_on_device():
    exe = build()
    out = subprocess.run([exe, "--big"], capture_output=True, text=True, timeout=1200)
    print(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout and "FAIL " not in out.stdout
    assert out.stdout.count("OK  ") >= 18
