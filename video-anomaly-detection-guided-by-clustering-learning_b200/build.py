"""Build libvadc.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m videoad_b200.build        (or: python <this file>)

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so
that it travels to the GPU box with the repo snapshot (it is git-ignored)."""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvadc.so")
STAMP = os.path.join(HERE, "build", "libvadc.stamp")
SOURCES = ["abi.cu", "cluster.cu", "cluster_tc.cu", "cluster_fwd_ws.cu", "cluster_bwd_fused.cu", "cluster_bwd_tc2.cu", "tc_gemm.cu", "space_cluster.cu", "memory.cu", "reduce.cu", "collective.cu", "decoder_entry.cu", "encoder_tail.cu", "umma_test.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "-Xptxas", "-v", "-cudart", "static",
]
NVCC_FLAGS = [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvadc.so cannot be built (no CPU fallback exists)")


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu into objects (in parallel) and link libvadc.so.
    Skips the work when sources are unchanged since the last build."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, logs = [], []
    for src, obj, p in procs:
        out, _ = p.communicate()
        logs.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(logs))
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(obj)
    with open(os.path.join(objdir, "ptxas.log"), "w") as fh:
        fh.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-o", LIB, *objs, "-ldl", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libvadc.so failed")
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
