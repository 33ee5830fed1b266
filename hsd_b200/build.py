"""In-tree build of the C-ABI library (nvcc, sm_100a only).

`python hsd_b200/build.py` or `__graft_entry__.build()` compiles every
`csrc/*.cu` into `hsd_b200/libhsd_b200.so`.  The .so is git-ignored but travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# HSD_B200_LIB: load / build an alternative library file (kernel-variant experiments: a second build with
# other -D flags next to the shipped one); HSD_B200_NVCC_FLAGS: extra nvcc flags for that build
LIB = os.environ.get("HSD_B200_LIB") or os.path.join(HERE, "libhsd_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas=-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the HSD kernels cannot be built")
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def host_sources():
    """Host-only C++ (the e2e path's multi-threaded mirror): g++, no device code."""
    return sorted(glob.glob(os.path.join(CSRC, "*.cpp")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + host_sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(HERE), "include", "hsd_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "build" if not os.environ.get("HSD_B200_LIB") else
                          "build_" + os.path.basename(LIB).replace(".", "_"))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get("HSD_B200_NVCC_FLAGS", "").split(), "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    gxx = shutil.which("g++") or "g++"
    for src in host_sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-4] + "_host.o")
        cmd = [gxx, "-O3", "-std=c++17", "-fPIC", "-pthread", "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"compiler failed on {src}")
    # the arch flag also on the link step: nvcc's device-link stub otherwise targets its default arch (sm_52)
    cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-Xcompiler", "-pthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        print(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
