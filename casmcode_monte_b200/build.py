"""In-tree build of the CUDA library (nvcc, sm_100a) and, later, the pybind11
host module.  Called by __graft_entry__.build(); the .so files stay next to the
package so they travel to the GPU box with the snapshot."""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcasm_monte_b200.so")
# /usr/bin/g++ links libstdc++ dynamically (the image's $CXX wrapper does not)
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", HOST_CXX,
    "-Xcompiler", "-fPIC",
    "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_cuda_library(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, "cmg_capi.cu")]
    deps = srcs + [os.path.join(CSRC, "cmg_device.cuh"), os.path.join(ROOT, "include", "casm_monte_gpu.h")]
    if not force and not _newer(LIB, deps):
        return LIB
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", LIB]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


def build_host_module(force=False):
    """pybind11 module mirroring the libcasm.monte API subset (host C++ above the C ABI)."""
    src = os.path.join(CSRC, "host", "monte_module.cpp")
    if not os.path.exists(src):
        return None
    import pybind11

    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(HERE, "_monte_b200" + ext)
    hdrs = [os.path.join(CSRC, "host", f) for f in os.listdir(os.path.join(CSRC, "host"))]
    inc = os.path.join(ROOT, "include")
    hdrs += [os.path.join(inc, "casm_monte_gpu.h")]
    hdrs += [os.path.join(inc, "casm_monte_b200", f) for f in os.listdir(os.path.join(inc, "casm_monte_b200"))]
    if not force and not _newer(out, hdrs + [LIB]):
        return out
    cmd = [
        HOST_CXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
        "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
        "-I" + os.path.join(ROOT, "include"), src, "-o", out,
        "-L" + HERE, "-lcasm_monte_b200", "-Wl,-rpath,$ORIGIN",
    ]
    subprocess.check_call(cmd)
    return out


def build_oracle():
    """The CPU oracle is test infrastructure; building it is not using it."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j2"], stdout=subprocess.DEVNULL)


def build_all(force=False, verbose=False):
    build_cuda_library(force, verbose)
    build_host_module(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
