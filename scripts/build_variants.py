#!/usr/bin/env python
"""Tuning: build libust.so variants with different streaming-kernel geometry into build_variants/ (git-ignored; they
travel to the GPU box with gpurun). scripts/gpu_variants.sh times each of them (UST_LIB=...)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "k8s-operator-libs_b200", "csrc")
OUT = os.path.join(ROOT, "build_variants")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "-Xcompiler", "-ffp-contract=off", "--fmad=false"]
VARIANTS = {
    "w8": ["-DUST_CONSUMER_WARPS=8"],
    "rep1": ["-DUST_HOT_REP=1"],
    "t5120w20s3": ["-DUST_TILE_NODES=5120", "-DUST_STAGES=3", "-DUST_CONSUMER_WARPS=20"],
    "t5120w10s3": ["-DUST_TILE_NODES=5120", "-DUST_STAGES=3", "-DUST_CONSUMER_WARPS=10"],
    "t3072w12s4": ["-DUST_TILE_NODES=3072", "-DUST_STAGES=4", "-DUST_CONSUMER_WARPS=12"],
    "t3072w12s5": ["-DUST_TILE_NODES=3072", "-DUST_STAGES=5", "-DUST_CONSUMER_WARPS=12"],
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        os.remove(os.path.join(OUT, f))
    want = sys.argv[1:] or list(VARIANTS)
    procs = []
    for name in want:
        out = os.path.join(OUT, name + ".so")
        cmd = ["/usr/local/cuda/bin/nvcc"] + FLAGS + VARIANTS[name] + ["-shared", "-o", out] + \
              [os.path.join(CSRC, s) for s in ("ust_stream.cu", "ust_kernels.cu", "ust_api.cu")] + ["-ldl"]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"variant {name} failed to build")
        print("built", name)


if __name__ == "__main__":
    main()
