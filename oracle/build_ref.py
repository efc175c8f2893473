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

The reference's two CUDA operators of the GPU branch (ResidualCodec.decompress_residuals / packbits,
CB/indexing/codecs/residual.py:104-130) are compiled the same way with nvcc for sm_100a -- `--gpu` below; about four
minutes of nvcc over the torch headers, so __graft_entry__.build() leaves them to this script:

  decompress_residuals_gpu_cpp <- third_party/ColBERT/colbert/indexing/codecs/decompress_residuals.{cpp,cu}
  packbits_gpu_cpp             <- third_party/ColBERT/colbert/indexing/codecs/packbits.{cpp,cu}

They run only on the GPU box (tests/test_gpu_ops.py::test_gpu_form_codec_operators_equal_the_reference_cuda_kernels),
where they pin both the oracle's restatement of the GPU-form decompression and our operators.

Usage:  python oracle/build_ref.py [--gpu] [--force]     (no-op if /root/reference is absent)
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

CODECS = os.path.join(CB, "indexing", "codecs")
GPU_SOURCES = {
    "decompress_residuals_gpu_cpp": [os.path.join(CODECS, "decompress_residuals.cpp"), os.path.join(CODECS, "decompress_residuals.cu")],
    "packbits_gpu_cpp": [os.path.join(CODECS, "packbits.cpp"), os.path.join(CODECS, "packbits.cu")],
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


def have_gpu_reference():
    return all(os.path.exists(p) for srcs in GPU_SOURCES.values() for p in srcs)


def build_gpu(verbose=False, force=False):
    """Compile the reference's CUDA operators (sm_100a) into oracle/_ref/.  Flags follow torch.utils.cpp_extension.load
    as the reference calls it (residual.py:104-129: extra_cuda_cflags=["-O3"], plus torch's own half-operator defines)."""
    if not have_gpu_reference():
        return []
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT, exist_ok=True)
    incs = []
    for p in ce.include_paths():
        incs += ["-isystem", p]
    incs += ["-isystem", sysconfig.get_paths()["include"], "-isystem", "/usr/local/cuda/include"]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    jobs, built = [], []
    for name, srcs in GPU_SOURCES.items():
        out = ref_so_path(name)
        if (not force) and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
            built.append(out)
            continue
        cmd = ["nvcc", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
               f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
               "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_BFLOAT16_CONVERSIONS__",
               "-D__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr", *incs, *srcs, "-o", out,
               f"-L{libdir}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
               "-Xlinker", f"-rpath={libdir}"]
        if verbose:
            print(" ".join(cmd))
        jobs.append((out, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)))
    for out, job in jobs:                     # the two modules compile side by side
        err = job.communicate()[1]
        if job.returncode != 0:
            raise RuntimeError(f"nvcc failed for {out}:\n{err[-4000:]}")
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
    if "--gpu" in sys.argv:
        paths += build_gpu(verbose=True, force="--force" in sys.argv)
    print("built:" if paths else "reference absent; nothing built", *paths)
