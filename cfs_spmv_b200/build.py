"""Builds the native libraries IN-TREE (they travel to the GPU box with the
snapshot; *.so is git-ignored):

  cfs_spmv_b200/lib/libcfs_cuda.so  -- CUDA kernels + the C ABI (include/cfs_cuda.h)
  cfs_spmv_b200/lib/libsparse.so    -- host C++ drop-in library (MMF loader,
                                       allocator, runtime, CSRMatrix/SpDMV
                                       instantiations) over the C ABI

nvcc cross-compiles for sm_100a without a GPU.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOSTSRC = os.path.join(HERE, "host")
LIBDIR = os.path.join(HERE, "lib")
INCLUDE = os.path.join(ROOT, "include")
OBJDIR = os.path.join(ROOT, "build", "obj")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

CUDA_SOURCES = ["cfs_cuda.cu", "preproc.cu", "windows.cu", "compress.cu", "tiles6.cu", "valindex.cu", "hubs.cu",
                "refmeta.cu", "mmf_ingest.cu", "cg.cu", "csr_path.cu", "multi.cu", "det.cu", "hyb.cu",
                "spmv.cu", "gen.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
    "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", HOST_CXX,
    "-I" + INCLUDE, "-I" + CSRC,
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _headers():
    hs = []
    for d in (CSRC, INCLUDE, HOSTSRC):
        for base, _, files in os.walk(d):
            hs += [os.path.join(base, f) for f in files
                   if f.endswith((".h", ".hpp", ".cuh", ".tpp"))]
    return hs


def build_cuda(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = _headers()
    objs = []
    procs = []
    for src in CUDA_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) \
                + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                stderr=subprocess.STDOUT,
                                                text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== nvcc %s ==\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    lib = os.path.join(LIBDIR, "libcfs_cuda.so")
    if force or procs or _newer(lib, objs):
        subprocess.check_call([NVCC, "-shared", "-o", lib] + objs
                              + ["-gencode", "arch=compute_100a,code=sm_100a",
                                 "-ccbin", HOST_CXX, "-cudart", "static"])
    return lib


def build_host(force=False):
    """libsparse.so: the C++ API of include/cfs.hpp over the C ABI."""
    os.makedirs(LIBDIR, exist_ok=True)
    srcs = [os.path.join(HOSTSRC, f) for f in sorted(os.listdir(HOSTSRC))
            if f.endswith(".cpp")]
    lib = os.path.join(LIBDIR, "libsparse.so")
    if force or _newer(lib, srcs + _headers()):
        subprocess.check_call(
            [HOST_CXX, "-std=c++11", "-O2", "-fPIC", "-shared", "-Wall",
             "-I" + INCLUDE]
            + srcs + ["-o", lib, "-L" + LIBDIR, "-lcfs_cuda",
                      "-Wl,-rpath,$ORIGIN"])
    return lib


def build_all(verbose=False, force=False):
    out = [build_cuda(verbose=verbose, force=force)]
    if os.path.isdir(HOSTSRC):
        out.append(build_host(force=force))
    return out


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv, force="-f" in sys.argv))
