"""Builds csrc/*.cu into the in-tree C-ABI library ``libplaid_b200.so`` (sm_100a only).

nvcc cross-compiles without a GPU, so this runs in the authoring container; the built .so is
git-ignored but travels to the GPU box with the repo snapshot.  ``python -m
reranking_multimodal_retrievers_b200.build [--force] [--verbose]``.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
BUILD_DIR = os.path.join(PKG_DIR, "csrc", "build")
LIB_PATH = os.path.join(PKG_DIR, "libplaid_b200.so")
HEADER = os.path.join(os.path.dirname(PKG_DIR), "include", "plaid_b200.h")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    deps = [HEADER] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return max(os.path.getmtime(d) for d in deps)


def _compile(src, obj, verbose):
    cmd = [NVCC, *NVCC_FLAGS, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {os.path.basename(src)}:\n{log}")
    if verbose:
        print(log)
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD_DIR, exist_ok=True)
    srcs = sources()
    dep_m = _deps_mtime()
    jobs, objs = [], []
    for src in srcs:
        obj = os.path.join(BUILD_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), dep_m)
        if stale:
            jobs.append((src, obj))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda a: _compile(a[0], a[1], verbose), jobs))
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [NVCC, "-shared", "-o", LIB_PATH, *objs, "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def build_variant(name: str, defines: list[str], files=("maxsim.cu",)) -> str:
    """Development aid for A/B timing: rebuilds `files` with extra -D flags and links them with the
    regular objects into csrc/build/variants/libplaid_b200.<name>.so (select it with PLAID_B200_LIB)."""
    build_library()
    vdir = os.path.join(BUILD_DIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for src in sources():
        base = os.path.basename(src)
        obj = os.path.join(BUILD_DIR, base[:-3] + ".o")
        if base in files:
            obj = os.path.join(vdir, f"{base[:-3]}.{name}.o")
            cmd = [NVCC, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", src, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            with open(obj + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
            if res.returncode != 0:
                raise RuntimeError(res.stdout + res.stderr)
        objs.append(obj)
    out = os.path.join(vdir, f"libplaid_b200.{name}.so")
    subprocess.check_call([NVCC, "-shared", "-o", out, *objs, "-cudart", "static"])
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        rest = sys.argv[i + 2:]
        files = ("maxsim.cu",)
        if "--files" in rest:                      # --variant NAME [--files a.cu,b.cu] DEFINE[=VALUE]...
            j = rest.index("--files")
            files = tuple(rest[j + 1].split(","))
            rest = rest[:j] + rest[j + 2:]
        print(build_variant(sys.argv[i + 1], [d[2:] if d.startswith("-D") else d for d in rest], files))
        sys.exit(0)
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
