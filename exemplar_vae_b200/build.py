"""Build recipe for the exvae_b200 CUDA library: nvcc -> csrc/libexvae_b200.so (sm_100a only).

The library is built IN-TREE so that it travels with a repository snapshot; nothing is
JIT-compiled at import time.  `python -m exemplar_vae_b200.build` (or `__graft_entry__.build()`)
rebuilds it; nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libexvae_b200.so")
SOURCES = ["api.cu", "prior_lse.cu", "pairdist_knn.cu", "knn_fused.cu", "gemm.cu", "gemm_tc.cu", "prior_lse_tc.cu", "prior_fused.cu", "prior_bwd_tc.cu", "elementwise.cu", "fused_small.cu", "mc_coll.cu", "vamp.cu", "conv.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the exvae_b200 CUDA library cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    extra = os.environ.get("NVCC_EXTRA", "").split()       # e.g. -DEXVAE_PF_TRACE (debug trace of the fused K1 forward)
    headers = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "gemm_tc.cuh"), os.path.join(CSRC, "tc_common.cuh"), os.path.join(CSRC, "prior_lse_tc.cuh"),
               os.path.join(CSRC, "..", "..", "include", "exvae_b200.h")]
    objs, jobs = [], []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    if jobs or force or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
