"""Build recipe for oracle/_ref: the REFERENCE's own CPU kernels, compiled where they lie.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.

The four pthread/C++ operators of the reference's PLAID search path are compiled
directly with g++ from their sources under /root/reference (no reference build
system is run, no reference source is copied into this repo); only the resulting
pybind modules land in oracle/_ref/ (git-ignored, but shipped to the GPU box by
gpurun like our own .so files):

  filter_pids_cpp           <- third_party/ColBERT/colbert/search/filter_pids.cpp
  decompress_residuals_cpp  <- third_party/ColBERT/colbert/search/decompress_residuals.cpp
  segmented_lookup_cpp      <- third_party/ColBERT/colbert/search/segmented_lookup.cpp
  segmented_maxsim_cpp      <- third_party/ColBERT/colbert/modeling/segmented_maxsim.cpp

Flags follow what the reference passes to torch.utils.cpp_extension.load
(extra_cflags=["-O3"], index_storage.py:36-57, colbert.py:49-59).

Usage:  python oracle/build_ref.py            (no-op if /root/reference is absent)
"""
import os
import subprocess
import sys
import sysconfig

REF_ROOT = os.environ.get("PLAID_REFERENCE_ROOT", "/root/reference")
CB = os.path.join(REF_ROOT, "third_party", "ColBERT", "colbert")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

SOURCES = {
    "filter_pids_cpp": os.path.join(CB, "search", "filter_pids.cpp"),
    "decompress_residuals_cpp": os.path.join(CB, "search", "decompress_residuals.cpp"),
    "segmented_lookup_cpp": os.path.join(CB, "search", "segmented_lookup.cpp"),
    "segmented_maxsim_cpp": os.path.join(CB, "modeling", "segmented_maxsim.cpp"),
}


def ref_so_path(name):
    return os.path.join(OUT, name + ".so")


def have_reference():
    return all(os.path.exists(p) for p in SOURCES.values())


def build(verbose=False, force=False):
    """Compile the reference operators into oracle/_ref/.  Returns list of built paths."""
    if not have_reference():
        return []
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    incs = []
    for p in ce.include_paths():
        incs += ["-isystem", p]
    incs += ["-isystem", sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    built = []
    for name, src in SOURCES.items():
        out = ref_so_path(name)
        if (not force) and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
            built.append(out)
            continue
        cmd = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread",
               f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H",
               f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *incs, src, "-o", out,
               f"-L{libdir}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python",
               f"-Wl,-rpath,{libdir}"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        built.append(out)
    return built


def load(name):
    """Import a built reference operator module from oracle/_ref (torch must import first)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols)
    path = ref_so_path(name)
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    paths = build(verbose=True, force="--force" in sys.argv)
    print("built:" if paths else "reference absent; nothing built", *paths)
