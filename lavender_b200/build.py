"""Builds lavender_b200/liblavender_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

    python -m lavender_b200.build [--force]

nvcc cross-compiles on a GPU-less box; the resulting .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "liblavender_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", os.path.join(os.path.dirname(HERE), "include")]
# --use_fast_math only where the instruction count matters (tensor-core kernels' epilogues / softmax, the optimizer pass);
# the HBM-bound row kernels (LayerNorm statistics, embeddings, cross-entropy log-sum-exp) use IEEE division / sqrt / exp.
FAST_MATH = {"gemm.cu", "attention_fwd.cu", "attention_flash.cu", "attention_bwd.cu", "optim.cu"}


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "lavender_b200.h"))
    hs.append(os.path.abspath(__file__))   # flag changes rebuild everything
    return max(os.path.getmtime(h) for h in hs)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    hm = _headers_mtime()
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hm):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        fm = ["--use_fast_math"] if os.path.basename(s) in FAST_MATH else []
        r = subprocess.run([NVCC] + FLAGS + fm + ["-c", s, "-o", o], capture_output=True, text=True)
        with open(o + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{r.stdout}\n{r.stderr}")
        return s, r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for s, log in ex.map(compile_one, jobs):
                if verbose:
                    print(f"== {os.path.basename(s)}\n{log}")
    if jobs or not os.path.exists(LIB) or force:
        r = subprocess.run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
