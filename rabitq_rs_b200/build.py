"""Builds librbq.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Run as `python rabitq_rs_b200/build.py` or through __graft_entry__.build() (loaded by path, because
importing the package requires the library to exist).  nvcc cross-compiles
without a GPU.  Flags that matter for parity: -fmad=false (no implicit mul+add contraction; the
kernels spell out every fma the reference uses), IEEE division and square root (nvcc defaults).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "librbq.so")
SOURCES = ["format.cc", "nccl_loader.cc", "query_prep.cu", "coarse.cu", "coarse_tc.cu", "scan.cu", "scan_tail.cu", "tail_tc.cu", "resolve.cu", "fetch.cu", "build.cu", "kmeans.cu", "bruteforce.cu", "exact_merge.cu", "api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC,-O2,-fno-fast-math",
          "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    headers = [os.path.join(CSRC, "rbq_internal.h"), os.path.join(CSRC, "scan_common.cuh"), os.path.join(CSRC, "rotate.cuh"), os.path.join(CSRC, "nccl_loader.h"), os.path.join(os.path.dirname(HERE), "include", "rbq.h"),
               os.path.abspath(__file__)]
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, os.path.splitext(s)[0] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC, "-x", "cu"] + ARCH + COMMON + list(extra) + ["-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([NVCC] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs +
            ["-ccbin", COMMON[-1], "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
