"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  librcb200.so        CUDA kernels + extern "C" layer (include/rcb200.h).  Needs only nvcc; built anywhere.
  librecast_b200.so   C++ host shim re-exporting the Recast.h voxel-stage functions on top of librcb200.
                      Needs the reference's *headers* (Recast.h, RecastAlloc.h, RecastAssert.h), which are
                      used unmodified from the reference tree and never copied into this repository, so it is
                      built only where that tree exists; elsewhere the prebuilt file is used.
"""
import glob
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_DIR = os.path.join(PKG, "lib")
CSRC = os.path.join(PKG, "csrc")
RCB_LIB = os.path.join(LIB_DIR, "librcb200.so")
SHIM_LIB = os.path.join(LIB_DIR, "librecast_b200.so")
REFERENCE = os.environ.get("RCB_REFERENCE", "/root/reference")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # the clipper must reproduce the reference's non-FMA IEEE arithmetic bit for bit
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc_path():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_rcb(force=False, verbose=False):
    srcs = [os.path.join(CSRC, "rcb_voxel.cu")] + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
           [os.path.join(ROOT, "include", "rcb200.h")]
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and not _newer(RCB_LIB, srcs):
        return RCB_LIB
    # linked under a temporary name and renamed, so that a repository snapshot never sees a half-written library
    tmp = RCB_LIB + ".tmp%d" % os.getpid()
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", tmp, srcs[0]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    try:
        subprocess.check_call(cmd)
        os.replace(tmp, RCB_LIB)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return RCB_LIB


def build_shim(force=False):
    """Returns the path of the shim library, or None when neither the reference headers nor a prebuilt
    library are available."""
    src = os.path.join(CSRC, "recast_shim.cpp")
    inc = os.path.join(REFERENCE, "Recast", "Include")
    if not os.path.exists(src):
        return None
    if not os.path.isdir(inc):
        return SHIM_LIB if os.path.exists(SHIM_LIB) else None
    if not force and not _newer(SHIM_LIB, [src, os.path.join(ROOT, "include", "rcb200.h")]):
        return SHIM_LIB
    build_rcb()
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fno-rtti", "-fno-exceptions",
           "-DNDEBUG", "-DRC_DISABLE_ASSERTS", "-I", inc, "-I", os.path.join(ROOT, "include"),
           "-o", SHIM_LIB, src, "-L", LIB_DIR, "-lrcb200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return SHIM_LIB


def build_all(force=False):
    build_rcb(force=force)
    build_shim(force=force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv)
    print(RCB_LIB)
