"""Build libp3d_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python pseudo-3d-interpolation_b200/build.py [--force] [--verbose]

Each translation unit is compiled to an object file in parallel and only when its sources
changed; the objects are linked into ``csrc/libp3d_b200.so`` (cudart linked statically so
the library has no dependency on a particular libcudart.so at load time).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libp3d_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("P3D_NVCC_FLAGS", "").split()


def _stamp(paths):
    h = hashlib.sha1()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, headers, force, verbose):
    obj = src[:-3] + ".o"
    stamp_file = obj + ".stamp"
    stamp = _stamp([src] + headers)
    if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, "", False
    cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, r.stderr, True


def build(force=False, verbose=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                     glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h")))
    objs, rebuilt, logs = [], False, []
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for obj, log, did in ex.map(lambda s: _compile(s, headers, force, verbose), srcs):
            objs.append(obj); rebuilt |= did; logs.append(log)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for log in logs:
            sys.stderr.write(log)
    with open(os.path.join(CSRC, "ptxas_info.log"), "a" if not force else "w") as f:
        for log in logs:
            f.write(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
