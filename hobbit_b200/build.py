"""In-tree build of the CUDA library (sm_100a only).  `python -m hobbit_b200.build` or __graft_entry__.build()."""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libhobbit_b200.so")
SOURCES = ["ntt.cu", "encode.cu", "merkle.cu", "commit.cu", "sumcheck.cu", "open.cu", "trace.cu", "dist.cu", "ubench.cu"]
HEADERS = ["common.cuh", "field.cuh", "blake3.cuh", "reduce.cuh", os.path.join("..", "..", "include", "hobbit_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    if not any(_newer(d, obj) for d in deps):
        return obj, ""
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, p.stdout, p.stderr))
    return obj, p.stderr


def build_library(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    objs = [r[0] for r in res]
    log = "".join(r[1] for r in res)
    if force or any(_newer(o, LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    # host-side C++ mirror of the reference API (plain g++, links against the C ABI)
    host_srcs = [os.path.join(HERE, "host", f) for f in ("hobbit_host.cpp", "hobbit_open.cpp", "hobbit_circuit.cpp")]
    host_lib = os.path.join(HERE, "libhobbit_host.so")
    if force or any(_newer(f, host_lib) for f in host_srcs) or _newer(os.path.join(HERE, "host", "hobbit_host.hpp"), host_lib) or _newer(LIB, host_lib):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", host_lib] + host_srcs + ["-L" + HERE, "-lhobbit_b200", "-lpthread", "-Wl,-rpath,$ORIGIN"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode:
            raise RuntimeError("host library build failed:\n%s\n%s" % (p.stdout, p.stderr))
    # reference-free example / bench binaries (tools/mlp_prove.cpp: the circuit proofs; tools/pc_prove.cpp: test_PC / test_Elastic_PC)
    for tool in ("mlp_prove", "pc_prove"):
        tool_src = os.path.join(HERE, "..", "tools", tool + ".cpp")
        tool_bin = os.path.join(HERE, tool)
        if os.path.exists(tool_src) and (force or _newer(tool_src, tool_bin) or _newer(host_lib, tool_bin)):
            cmd = ["g++", "-O2", "-std=c++17", "-o", tool_bin, tool_src, "-L" + HERE, "-lhobbit_host", "-lhobbit_b200", "-Wl,-rpath,$ORIGIN"]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode:
                raise RuntimeError("%s build failed:\n%s\n%s" % (tool, p.stdout, p.stderr))
    return LIB, log


if __name__ == "__main__":
    lib, log = build_library(force="--force" in sys.argv, verbose="-v" in sys.argv)
    if log:
        print(log)
    print(lib)
