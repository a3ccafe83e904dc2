"""Build libmlvae_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m ml_vae_b200.build [--force]

nvcc cross-compiles without a GPU, so this runs in the CPU build container; the
resulting .so is git-ignored but travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libmlvae_b200.so")
STAMP = LIB + ".stamp"
SOURCES = ["latent_loss.cu", "fbank.cu", "norm.cu", "gemm_chain.cu", "dense_bwd.cu", "pcm.cu", "lstm.cu", "lstm_pack.cu", "dropout.cu", "gemm.cu", "optim.cu", "mlp_chain.cu", "dp_optim.cu", "md_decode.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default"]


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _digest() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + [os.path.join("..", "..", "include", "mlvae_b200.h")]
    for f in files:
        p = os.path.normpath(os.path.join(CSRC, f))
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB          # GPU box without toolkit: use the shipped build
        raise RuntimeError("nvcc not found and no prebuilt libmlvae_b200.so")
    objs = []
    build_dir = os.path.join(PKG, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(build_dir, os.path.basename(src) + ".o")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
               "-Xcompiler", "-fPIC", "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("nvcc link failed")
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
