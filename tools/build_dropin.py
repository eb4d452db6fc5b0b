"""Builds the drop-in consumers into build/dropin/ (git-ignored, travels to the
GPU box):

  test_spmv_mmf, bench_spmv_mmf_dp, bench_spmv_mmf_sp
      the reference's UNMODIFIED test/test_spmv_mmf.cpp and
      bench/bench_spmv_mmf.cpp, compiled where they lie under /root/reference
      against THIS repo's include/ and libsparse.so (only where the reference
      tree exists; nothing is copied);
  api_consumer, load_timer
      this repo's own consumers of the same API (tests/cpp/*.cpp), always
      built; load_timer is also linked against the unmodified reference by
      oracle/ref_build/Makefile.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "build", "dropin")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
LIBDIR = os.path.join(ROOT, "cfs_spmv_b200", "lib")
COMMON = ["-std=c++11", "-O2", "-fopenmp", "-w", "-I" + os.path.join(ROOT, "include"),
          "-L" + LIBDIR, "-lsparse", "-lcfs_cuda",
          "-Wl,-rpath,$ORIGIN/../../cfs_spmv_b200/lib"]


def cc(src, out, defs=()):
    cmd = [CXX, src, "-o", os.path.join(OUT, out)] + list(defs) + COMMON
    subprocess.check_call(cmd)


def main():
    os.makedirs(OUT, exist_ok=True)
    cc(os.path.join(ROOT, "tests", "cpp", "api_consumer.cpp"), "api_consumer",
       ["-DCFS_ENABLE_DP"])
    cc(os.path.join(ROOT, "tests", "cpp", "load_timer.cpp"), "load_timer",
       ["-DCFS_ENABLE_DP"])
    if os.path.isdir(REF):
        cc(os.path.join(REF, "test", "test_spmv_mmf.cpp"), "test_spmv_mmf",
           ["-DCFS_ENABLE_DP"])
        cc(os.path.join(REF, "bench", "bench_spmv_mmf.cpp"), "bench_spmv_mmf_dp",
           ["-DCFS_ENABLE_DP"])      # --enable-dp
        cc(os.path.join(REF, "bench", "bench_spmv_mmf.cpp"), "bench_spmv_mmf_sp")
    print("drop-in consumers in", OUT, sorted(os.listdir(OUT)))


if __name__ == "__main__":
    sys.exit(main())
