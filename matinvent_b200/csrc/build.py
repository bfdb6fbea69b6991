"""In-tree build of libmatinvent_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m matinvent_b200.csrc.build [--force]

The .so lands in matinvent_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmatinvent_b200.so")
SOURCES = ["mi_gemm.cu", "mi_ops.cu", "mi_graph.cu", "mi_tc.cu", "mi_node.cu", "mi_edge.cu", "mi_pipeline.cu"]
HEADERS = ["mi_common.cuh", "mi_tc_common.cuh", os.path.join("..", "..", "include", "matinvent_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isfile(c) or c == "nvcc"):
            return c
    return "nvcc"


def _sources():
    return [s for s in SOURCES if os.path.isfile(os.path.join(HERE, s))]


def _digest():
    h = hashlib.sha256()
    for f in _sources() + HEADERS:
        with open(os.path.join(HERE, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.sha256")
    dig = _digest()
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(stamp) and open(stamp).read().strip() == dig:
        return LIB_PATH
    objs = []
    procs = []
    for s in _sources():
        o = os.path.join(LIB_DIR, s.replace(".cu", ".o"))
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(HERE, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs
    subprocess.check_call(cmd)
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
