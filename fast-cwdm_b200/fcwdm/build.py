"""Build libfcwdm.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python fast-cwdm_b200/fcwdm/build.py [--force]

The shared object lands next to this file (git-ignored, but it travels to the GPU box with the snapshot).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libfcwdm.so")

SOURCES = ["core.cu", "haar.cu", "diffusion.cu", "norm.cu", "linear.cu", "conv3d.cu", "conv3d_pair.cu", "conv3d_chain.cu", "conv3d_wgrad.cu", "train.cu", "resample.cu", "preprocess.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr"]
if os.environ.get("FCWDM_CONV_TRACE") == "1":          # development build with in-kernel clock stamps (tools/conv_trace.py)
    NVCC_FLAGS.append("-DFCWDM_CONV_TRACE")


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, INCLUDE):
        for f in sorted(os.listdir(root)):
            with open(os.path.join(root, f), "rb") as fh:
                h.update(f.encode() + b"\0" + fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    logs = {}

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        p = subprocess.run(cmd, capture_output=True, text=True)
        logs[src] = p.stdout + p.stderr
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{logs[src]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    with open(os.path.join(OBJ, "ptxas.log"), "w") as fh:
        for src in SOURCES:
            fh.write(f"==== {src}\n{logs[src]}\n")
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-lcudart", "-Xlinker", "--no-as-needed"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("link failed:\n" + p.stdout + p.stderr)
    with open(stamp, "w") as fh:
        fh.write(digest)
    if verbose:
        print(open(os.path.join(OBJ, "ptxas.log")).read())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
