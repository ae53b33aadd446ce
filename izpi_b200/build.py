"""Builds izpi_b200/libizpi_cuda.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension
machinery: the library is a plain C-ABI shared object, see include/izpi_cuda.h)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libizpi_cuda.so")
BUILD = os.path.join(HERE, "build")

CU = ["device/context.cu", "device/trace.cu", "device/render.cu", "device/displace.cu", "device/bvh_build.cu"]
CPP = ["host/error.cpp", "host/bvh4_builder.cpp", "host/host_scene.cpp", "host/proto_scene.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# --fmad=false: Go/amd64 never fuses multiply-add; bit-exact parity with the reference depends on it.
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--fmad=false", "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread",
              "-Xptxas", "-v", "--expt-relaxed-constexpr"]


def _sources():
    out = []
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files]
    out += [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    if os.environ.get("IZPI_LIB_PATH") and out is None:
        return os.environ["IZPI_LIB_PATH"]  # a prebuilt experiment variant is in use: leave it alone
    if out is None and not force and not needs_build():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    log = []
    extra = os.environ.get("IZPI_NVCC_DEFS", "").split()  # e.g. "-DIZPI_G4_MIN_BLOCKS=7" for occupancy experiments
    for src in CU + CPP:
        obj = os.path.join(BUILD, src.replace("/", "_") + ".o")
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, *extra, "-x", "cu" if src.endswith(".cu") else "c++", "-c", os.path.join(CSRC, src), "-o", obj]
        if not src.endswith(".cu"):
            cmd = [nvcc, "-O2", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread", "-x", "c++", "-c",
                   os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on " + src)
        objs.append(obj)
    r = subprocess.run([nvcc, *ARCH, "-shared", "-o", out or LIB, *objs, "-lcudart", "-Xcompiler", "-pthread"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(BUILD, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return out or LIB


if __name__ == "__main__":
    outs = [a[len("--out="):] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=outs[0] if outs else None))
