"""Build libsgcn_b200.so (sm_100a only) in-tree with nvcc.

The library has no torch / Python dependency: it is a plain C-ABI shared object (include/sgcn_b200.h)
linked against the CUDA runtime.  ``python -m stochastic_gcn_b200.build`` rebuilds it.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsgcn_b200.so")
SOURCES = ["api.cu", "rows.cu", "aggregate.cu", "sampler.cu", "exchange.cu", "step.cu", "precompute.cu", "dense.cu", "gemm.cu"]
HEADERS = ["common.cuh", "scan.cuh", "mt19937.cuh", "exchange.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",                 # aggregate kernels use explicit fmaf; the sampler uses __f*_rn
    "-Xcompiler", "-fPIC,-O2,-Wall",
    "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
    "-shared", "-cudart", "shared",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libsgcn_b200.so")
    return exe


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(ROOT, "include", "sgcn_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ into stochastic_gcn_b200/libsgcn_b200.so."""
    if not force and not _stale():
        return LIB
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f not in ("-shared",)] + \
              ["-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append("== %s ==\n%s" % (src, out))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        objs.append(obj)
    with open(os.path.join(build_dir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    cmd = [_nvcc(), "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", LIB] + objs
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("link failed:\n" + out.stdout + out.stderr)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
