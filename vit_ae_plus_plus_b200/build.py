"""Builds libvitae_b200.so (sm_100a only) in-tree with nvcc.  `python -m vit_ae_plus_plus_b200.build`."""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libvitae_b200.so")
BUILD_DIR = os.path.join(PKG_DIR, "csrc", "build")
SOURCES = ["api.cu", "gemm_tcgen05.cu", "attention.cu", "attention_tc.cu", "layernorm.cu", "token_ops.cu", "loss.cu", "edge_loss.cu", "ingest.cu", "predictor.cu", "dp_probe.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-warn-spills"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libvitae_b200.so)")


def _digest() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(PKG_DIR), "include")):
        for name in sorted(os.listdir(root)):
            path = os.path.join(root, name)
            if os.path.isfile(path) and name.endswith((".cu", ".cuh", ".h")):
                h.update(name.encode())
                with open(path, "rb") as f:
                    h.update(f.read())
    return h.hexdigest()


def _build_variant(flags, variant: str, verbose: bool) -> str:
    out_dir = os.path.join(BUILD_DIR, variant)
    os.makedirs(out_dir, exist_ok=True)
    out = LIB_PATH.replace(".so", f"_{variant}.so")
    nvcc = _nvcc()
    objs = []
    for src in SOURCES:
        obj = os.path.join(out_dir, src.replace(".cu", ".o"))
        r = subprocess.run([nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(obj)
    r = subprocess.run([nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {out}")
    return out


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    stamp = os.path.join(BUILD_DIR, "digest.txt")
    flags = list(NVCC_FLAGS)
    if os.environ.get("VITAE_TRACE") == "1":      # debug build: per-phase clock stamps in the GEMM (tools/gemm_trace.py)
        flags.append("-DVITAE_GEMM_TRACE")
    if os.environ.get("VITAE_EPI_WARPS"):          # experiment: 4 or 8 epilogue warps in the GEMM
        flags.append("-DVITAE_EPI_WARPS=" + os.environ["VITAE_EPI_WARPS"])
    if os.environ.get("VITAE_EPI_NOGELU") == "1":  # experiment only: GELU code compiled out of the GEMM epilogue
        flags.append("-DVITAE_EPI_NOGELU")
    if os.environ.get("VITAE_PDL_EARLY") == "1":   # experiment: GEMM / attention let their dependents launch at their start
        flags.append("-DVITAE_PDL_EARLY")
    if os.environ.get("VITAE_ATTN_TAIL") == "1":   # experiment: attention loops bounded by the valid part of ragged tail tiles
        flags.append("-DVITAE_ATTN_TAIL")
    if os.environ.get("VITAE_ATTN_TRACE") == "1":  # debug build: per-phase stamps in the attention forward (tools/attn_trace.py)
        flags.append("-DVITAE_ATTN_TRACE")
    variant = os.environ.get("VITAE_BUILD_VARIANT")     # experiment builds go to libvitae_b200_<variant>.so (VITAE_LIB selects)
    if variant:
        return _build_variant(flags, variant, verbose)
    digest = _digest() + "".join(f for f in flags if f.startswith("-DVITAE"))
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(BUILD_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr.strip():
            print(r.stderr.strip(), file=sys.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print(f"built {LIB_PATH}")
    return LIB_PATH


if __name__ == "__main__":
    build(force="--force" in sys.argv)
