"""In-tree build of the CUDA library (sm_100a only) with nvcc; no JIT cache, no torch extension.

``python -m clairs_to_b200.build`` (or ``__graft_entry__.build()``) produces
``clairs_to_b200/libcto_b200.so`` next to this file, so the binary travels with the tree.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libcto_b200.so")
SOURCES = ["api.cu", "encoder.cu", "candidates.cu", "tokenize_dev.cu", "hard_filter.cu", "hard_filter_host.cpp", "nn_kernels.cu", "engine.cu", "gemm_tc.cu", "gemm_pair.cu", "gru_tc3.cu", "gru_tc4.cu", "gru_in_tc.cu", "aff_stage1.cu", "aff_fused.cu", "host_codec.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--use_fast_math=false",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = find_nvcc()
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "clairs_to_b200.h"))
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(obj_dir, src + ".o")
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + ["-c", path, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), file=sys.stderr)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
        if p.returncode != 0:
            sys.stderr.write("[clairs_to_b200.build] nvcc failed on %s\n" % src)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB_PATH, objs):
        cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
