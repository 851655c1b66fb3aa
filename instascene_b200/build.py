"""Builds instascene_b200/libisr.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libisr.so")
OBJ = os.path.join(HERE, "_obj")
SOURCES = ["isr_api.cu", "isr_preprocess.cu", "isr_binning.cu", "isr_blend_fwd.cu", "isr_blend_bwd.cu",
           "isr_contrastive.cu", "isr_sampler.cu", "isr_knn.cu", "isr_auxmaps.cu", "isr_tracker.cu", "isr_photometric.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newer(src: str, dst: str) -> bool:
    return not os.path.exists(dst) or os.path.getmtime(src) > os.path.getmtime(dst)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, "isr_common.cuh"), os.path.join(HERE, "..", "include", "isr.h")]
    hdr_time = max(os.path.getmtime(h) for h in headers)
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s.replace(".cu", ".o"))
        if force or _newer(src, obj) or os.path.getmtime(obj) < hdr_time:
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run(["nvcc", *NVCC_FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        log = os.path.join(OBJ, os.path.basename(src) + ".log")
        with open(log, "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return src

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for done in ex.map(compile_one, jobs):
                if verbose:
                    print("compiled", os.path.basename(done))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(OUT):
        r = subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs,
                            "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
