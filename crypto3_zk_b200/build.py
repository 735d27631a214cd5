"""Builds crypto3_zk_b200/libzkb200.so (CUDA, sm_100a only) in-tree with nvcc.

    python -m crypto3_zk_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU, so this also runs in the GPU-less build container; the resulting
.so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libzkb200.so")

CU_SOURCES = ["zkb_ctx.cu", "zkb_ntt.cu", "zkb_msm.cu", "zkb_hash.cu", "zkb_poly.cu", "zkb_scan.cu", "zkb_multi.cu", "zkb_plonk.cu", "zkb_points.cu"]
CXX_SOURCES = ["zkb_msm_host.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _run(cmd, verbose, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(log[-1])
    if r.returncode != 0:
        raise RuntimeError("build step failed: " + " ".join(cmd))


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    jobs, objs, log = [], [], []
    for s in CU_SOURCES:
        o = os.path.join(OBJ, s + ".o")
        objs.append(o)
        jobs.append([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, s), "-o", o])
    for s in CXX_SOURCES:
        o = os.path.join(OBJ, s + ".o")
        objs.append(o)
        jobs.append(["g++", "-O2", "-std=c++17", "-fPIC", "-c", os.path.join(CSRC, s), "-o", o])
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        list(ex.map(lambda c: _run(c, verbose, log), jobs))
    _run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lpthread"], verbose, log)
    with open(os.path.join(OBJ, "build.log"), "w") as f:
        f.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
