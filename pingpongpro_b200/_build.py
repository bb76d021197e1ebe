"""In-tree build of every native artefact (explicit g++ / nvcc command lines, sm_100a only).

    python -m pingpongpro_b200._build [target ...]

Targets: host (libppp_host.so), synth (libppp_synth.so + bin/ppp_synth), cuda (libppp_b200.so),
cli (bin/pingpongpro), oracle (oracle/_build + oracle/_ref when /root/reference exists).
Artefacts are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib")
BIN = os.path.join(PKG, "bin")
INC = os.path.join(ROOT, "include")
HOST = os.path.join(PKG, "host")
TOOLS = os.path.join(PKG, "tools")
CSRC = os.path.join(PKG, "csrc")

CXX = os.environ.get("CXX", "g++")
NVCC = os.environ.get("NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")
CXXFLAGS = ["-O3", "-std=gnu++17", "-fPIC", "-Wall", "-fvisibility=hidden", "-I", INC, "-I", HOST, "-I", TOOLS]
# -fmad=false: the bit-exact float contract of SURVEY.md App. A forbids FMA contraction.
NVCCFLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "-I", INC, "-I", CSRC, "-I", HOST,
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def _run(cmd: list[str]) -> None:
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def _headers() -> list[str]:
    out = []
    for d in (INC, HOST, TOOLS, CSRC):
        if os.path.isdir(d):
            out += [os.path.join(d, f) for f in os.listdir(d) if f.endswith((".h", ".cuh"))]
    return out


def build_host() -> str:
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libppp_host.so")
    src = [os.path.join(HOST, f) for f in ("aln_reader.cpp", "fast_inflate.cpp", "frontend.cpp", "annotation.cpp")]
    if _newer(out, src + _headers()):
        _run([CXX, *CXXFLAGS, "-shared", "-o", out, *src, "-lz", "-lpthread"])
    return out


def build_synth() -> str:
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(BIN, exist_ok=True)
    out = os.path.join(LIB, "libppp_synth.so")
    src = [os.path.join(TOOLS, "synth_lib.cpp")]
    if _newer(out, src + _headers()):
        _run([CXX, *CXXFLAGS, "-shared", "-o", out, *src, "-lpthread"])
    tool = os.path.join(BIN, "ppp_synth")
    tsrc = [os.path.join(TOOLS, "ppp_synth.cpp")]
    if _newer(tool, tsrc + _headers()):
        _run([CXX, *CXXFLAGS, "-o", tool, *tsrc, "-lz", "-lpthread"])
    return out


def build_cuda() -> str:
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libppp_b200.so")
    src = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))
    if _newer(out, src + _headers()):
        _run([NVCC, *NVCCFLAGS, "-shared", "-o", out, *src, "-cudart", "static"])
    return out


def build_cli() -> str:
    os.makedirs(BIN, exist_ok=True)
    out = os.path.join(BIN, "pingpongpro")
    src = [os.path.join(HOST, f) for f in ("cli_main.cpp", "cli_args.cpp", "annotation.cpp", "writers.cpp", "pipeline.cpp", "aln_reader.cpp", "fast_inflate.cpp", "frontend.cpp")]
    src = [s for s in src if os.path.exists(s)]
    if not os.path.exists(os.path.join(HOST, "cli_main.cpp")):
        return ""
    lib = build_cuda()
    if _newer(out, src + _headers() + [lib]):
        _run([CXX, "-O3", "-std=gnu++17", "-Wall", "-I", INC, "-I", HOST, "-o", out, *src,
              "-L", LIB, "-lppp_b200", "-Wl,-rpath,$ORIGIN/../lib", "-lz", "-lpthread"])
    return out


def build_oracle() -> None:
    odir = os.path.join(ROOT, "oracle")
    _run(["make", "-s", "-C", odir, "port"])
    ref_src = os.environ.get("PPP_REFERENCE_SOURCE", "/root/reference/pingpongpro.cpp")
    harness = os.path.join(odir, "_ref", "ppp_harness")
    deps = [os.path.join(odir, "harness.cpp"), os.path.join(odir, "sha256.h"), os.path.join(odir, "seqan_shim", "seqan", "shim_impl.h"),
            os.path.join(HOST, "aln_reader.cpp"), os.path.join(HOST, "aln_reader.h"), os.path.join(HOST, "fast_inflate.cpp"), os.path.join(TOOLS, "synth_core.h")]
    if os.path.exists(ref_src) and _newer(harness, deps):
        _run(["make", "-s", "-C", odir, "ref", f"REF={ref_src}"])


TARGETS = {"host": build_host, "synth": build_synth, "cuda": build_cuda, "cli": build_cli, "oracle": build_oracle}


def build_all() -> None:
    for name in ("host", "synth", "cuda", "cli", "oracle"):
        TARGETS[name]()


if __name__ == "__main__":
    names = sys.argv[1:] or list(TARGETS)
    for n in names:
        TARGETS[n]()
