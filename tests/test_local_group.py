"""The in-process collective behind `pingpongpro --gpus N` (csrc/local_group.cpp), without a GPU: its two CUDA calls
are replaced by host copies (tests/native/local_group_test.cpp) and the barrier / reduction logic runs under
ThreadSanitizer with up to eight rank threads -- sums and maxima of both widths, abort, mismatching sizes."""
import os
import shutil
import subprocess

import pytest

from pingpongpro_b200 import _build

HERE = os.path.dirname(os.path.abspath(__file__))


def _cuda_include():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), "include")
    return inc if os.path.exists(os.path.join(inc, "cuda_runtime.h")) else None


def _compiler_with(sanitizer, tmp_path):
    """A g++ that can link -fsanitize=<sanitizer> (some toolchain wrappers ship without the runtime libraries)."""
    probe = tmp_path / "probe.cpp"
    probe.write_text("int main() { return 0; }\n")
    for cxx in dict.fromkeys([_build.CXX, "/usr/bin/g++", "g++"]):
        if shutil.which(cxx) and subprocess.run([cxx, f"-fsanitize={sanitizer}", str(probe), "-o", str(tmp_path / "probe")], capture_output=True).returncode == 0:
            return cxx
    return None


@pytest.mark.parametrize("sanitizer", ["thread", "address,undefined"])
def test_local_group_under_sanitizers(tmp_path, sanitizer):
    inc = _cuda_include()
    if inc is None:
        pytest.skip("cuda_runtime.h not found")
    cxx = _compiler_with(sanitizer, tmp_path)
    flags = [f"-fsanitize={sanitizer}", "-fno-omit-frame-pointer"] if cxx else []  # no sanitizer runtime here: the plain logic test still runs
    exe = str(tmp_path / "local_group_test")
    subprocess.check_call([cxx or _build.CXX, "-O1", "-g", "-std=gnu++17", *flags, "-I", inc, "-I", _build.CSRC,
                           os.path.join(HERE, "native", "local_group_test.cpp"), os.path.join(_build.CSRC, "local_group.cpp"), "-lpthread", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "local_group_test: ok" in r.stdout, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr and "AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr
