"""
Build quantum_css_codes_b200/libqcss.so in-tree with nvcc for sm_100a.

    python -m quantum_css_codes_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects are rebuilt when a source or header is newer.
"""

import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libqcss.so")
SOURCES = ["api.cu", "small_dispatch.cu", "small_named_steane.cu", "small_named_qrm15.cu",
           "small_named_golay23.cu", "small_generic16.cu", "small_generic32.cu", "tiled_kernels.cu",
           "gf2_kernels.cu", "gf2_m4r.cu", "gf2_m4r2.cu", "gf2_derive.cu", "gf2_normalize.cu", "ec_kernels.cu", "sample_tiles.cu", "table_kernels.cu", "hist_kernels.cu", "dense_kernels.cu", "format_kernels.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _newest_header():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h", ".inc"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "qcss.h"))
    return max(os.path.getmtime(p) for p in deps)


def _compile(src, verbose, extra):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [nvcc(), *ARCH, *FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    return obj, res.stderr


def build(force=False, verbose=False, extra=()):
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = _newest_header()
    todo, objs = [], []
    for src in SOURCES:
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        src_time = max(os.path.getmtime(os.path.join(CSRC, src)), hdr_time)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < src_time:
            todo.append(src)
    logs = []
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(todo)) as pool:
            for obj, log in pool.map(lambda s: _compile(s, verbose, list(extra)), todo):
                logs.append(log)
    if todo or not os.path.exists(LIB):
        cmd = [nvcc(), *ARCH, "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB, "\n".join(logs)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    lib, log = build(force=args.force, verbose=args.verbose)
    if args.verbose:
        sys.stderr.write(log)
    print(lib)
