"""Build recipe for the product library (CUDA kernels + C ABI), sm_100a only.

    python -m hmmufotu_b200.build        # -> hmmufotu_b200/lib/libhmmufotu_b200.so
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libhmmufotu_b200.so")
SOURCES = ["hu_db.cu", "hu_viterbi.cu", "hu_seed.cu", "hu_seedmm.cu", "hu_estimate.cu", "hu_place.cu", "hu_api.cu"]
HEADERS = ["hu_internal.cuh", "hu_batch.cuh", "hu_mm.cuh", "hu_pipe.cuh", os.path.join("..", "..", "include", "hmmufotu_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + (["-DHU_PROF"] if os.environ.get("HU_PROF") else [])
# --fmad=false where results must be bit-identical to the reference's separate multiplies and adds (Viterbi
# costs) or to the host evaluation of the same code (model matrices, tree evaluation); the SEP kernels are
# checked to 1e-9 relative and may contract
NO_FMAD = {"hu_viterbi.cu", "hu_db.cu", "hu_api.cu", "hu_seed.cu", "hu_seedmm.cu"}


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (["--fmad=false"] if src in NO_FMAD else ["--fmad=true"]) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose:
            sys.stderr.write(out)
        with open(os.path.join(LIBDIR, src + ".ptxas.log"), "w") as f:
            f.write(out)
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    return LIB


HOST = os.path.join(HERE, "host")
BINDIR = os.path.join(HERE, "bin")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"  # dynamic libstdc++ (see oracle/Makefile)


def build_host(force: bool = False) -> str:
    """The C++ host layer: hmmufotu (the reference's command line over the C ABI) and its CPU test driver."""
    lib = build()
    os.makedirs(BINDIR, exist_ok=True)
    srcs = [os.path.join(HOST, f) for f in ("hu_host.cpp", "hu_build.cpp", "hu_facade.cpp", "hu_host.h")]
    out = []
    for name, main in (("hmmufotu", "hmmufotu.cpp"), ("hmmufotu-build", "hmmufotu-build.cpp"), ("hmmufotu-sim", "hmmufotu-sim.cpp"), ("hmmufotu-inspect", "hmmufotu-inspect.cpp"),
                       ("hu_host_test", "hu_host_test.cpp")):
        exe = os.path.join(BINDIR, name)
        deps = srcs + [os.path.join(HOST, main), lib]
        if force or _stale(exe, deps):
            subprocess.check_call([CXX, "-O2", "-g", "-std=c++11", "-pthread", "-Wall", "-Wextra", os.path.join(HOST, "hu_host.cpp"),
                                   os.path.join(HOST, "hu_build.cpp"), os.path.join(HOST, "hu_facade.cpp"), os.path.join(HOST, main), "-L" + LIBDIR, "-lhmmufotu_b200", "-Wl,-rpath," + LIBDIR,
                                   "-Wl,-rpath,$ORIGIN/../lib", "-o", exe])
        out.append(exe)
    return out[0]


if __name__ == "__main__":
    build_host(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
