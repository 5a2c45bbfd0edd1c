"""In-tree build of the native libraries (no build system beyond nvcc/g++ command lines).

  libmmmfe_b200.so   CUDA engine behind include/mmmfe_b200.h (sm_100a only)
  libmmmfe_host.so   C++ host front-end behind include/mmmfe_host.h (built when its sources exist)
  DarcyVT            drop-in executable
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB_ENGINE = os.path.join(HERE, "libmmmfe_b200.so")
LIB_HOST = os.path.join(HERE, "libmmmfe_host.so")
EXE = os.path.join(ROOT, "DarcyVT")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + res.stdout)
        raise RuntimeError("native build failed")
    return res.stdout


def build_engine(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("kernels.cu", "wide.cu", "sweep.cu", "engine.cu", "nd_tree.cpp", "sweep_plan.cpp", "exchange_plan.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("kernels.cuh", "wide.cuh", "dev_common.cuh", "nd_tree.h", "sweep.cuh", "sweep_plan.h", "exchange_plan.h", "common.h")] + \
        [os.path.join(ROOT, "include", "mmmfe_b200.h")]
    if not force and not _newer(LIB_ENGINE, deps):
        return LIB_ENGINE
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    tmp = LIB_ENGINE + ".tmp"  # built aside and renamed: a snapshot of the tree never sees a half-written library
    cmd = [nvcc] + NVCC_FLAGS + ["-x", "cu", "-shared", "-o", tmp] + srcs
    if os.path.exists("/usr/include/nccl.h"):
        cmd += ["-DMMMFE_WITH_NCCL", "-ldl"]  # NCCL is dlopen'ed at run time
    if verbose:
        cmd += ["-Xptxas", "-v"]
    out = _run(cmd)
    os.replace(tmp, LIB_ENGINE)
    if verbose:
        print(out)
    return LIB_ENGINE


DATA_OBJ = os.path.join(HERE, "data_device.o")
DATA_DIAG = ["-diag-suppress", "20012,20014,940,20011,20015,177"]


def build_data_device(data_header, force=False):
    """host/data_device.cu: the problem-data header (inc/data.h interface) compiled for the GPU.  A header that does
    not compile for the device (helpers that are not members, exotic includes) is not an error: the object is then
    built with MMMFE_NO_DEVICE_DATA and the front-end keeps its host loops.  Returns True when the device path is in."""
    src = os.path.join(HOST, "data_device.cu")
    deps = [src, os.path.join(HOST, "data_device.h"), data_header]
    for d, _, fs in os.walk(os.path.join(HOST, "shim")):
        deps += [os.path.join(d, f) for f in fs]
    marker = DATA_OBJ + ".mode"
    key = f"{data_header}"
    if not force and not _newer(DATA_OBJ, deps) and os.path.exists(marker):
        mode, old_key = (open(marker).read().split("\n") + [""])[:2]
        if old_key == key:
            return mode == "device"
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    base = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
            "-Xcompiler", "-fPIC", "-I", os.path.join(HOST, "shim"), "-I", HOST, "-I", os.path.join(ROOT, "include"),
            f'-DMMMFE_DATA_HEADER="{data_header}"'] + DATA_DIAG + ["-c", src, "-o", DATA_OBJ + ".tmp"]
    res = subprocess.run(base, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    mode = "device"
    if res.returncode != 0:
        sys.stderr.write("[build] the data header does not compile for the device; keeping the host loops:\n" +
                         "\n".join(res.stdout.splitlines()[:12]) + "\n")
        _run(base[:-4] + ["-DMMMFE_NO_DEVICE_DATA", "-c", src, "-o", DATA_OBJ + ".tmp"])
        mode = "host"
    os.replace(DATA_OBJ + ".tmp", DATA_OBJ)
    with open(marker, "w") as f:
        f.write(mode + "\n" + key)
    return mode == "device"


def _cuda_lib_dir():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    root = os.path.dirname(os.path.dirname(os.path.realpath(nvcc)))
    for d in ("lib64", "targets/x86_64-linux/lib"):
        if os.path.exists(os.path.join(root, d, "libcudart_static.a")):
            return os.path.join(root, d)
    return os.path.join(root, "lib64")


def build_host(force=False):
    srcs = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp") and f != "darcy_main.cpp"] \
        if os.path.isdir(HOST) else []
    if not srcs:
        return None
    hdrs = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".h")]
    for d, _, fs in os.walk(os.path.join(HOST, "shim")):
        hdrs += [os.path.join(d, f) for f in fs]
    deps = srcs + hdrs + [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    data_header = os.environ.get("MMMFE_DATA_HEADER", os.path.join(HOST, "data_default.h"))
    common = ["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-fopenmp", "-I", os.path.join(ROOT, "include"),
              "-I", os.path.join(HOST, "shim"), "-I", HOST, f'-DMMMFE_DATA_HEADER="{data_header}"']
    build_data_device(data_header, force=force)
    if force or _newer(LIB_HOST, deps + [LIB_ENGINE, DATA_OBJ]):
        build_engine()
        _run(common + ["-shared", "-o", LIB_HOST] + srcs + [DATA_OBJ] +
             ["-L", HERE, "-lmmmfe_b200", "-L", _cuda_lib_dir(), "-lcudart_static", "-lrt", "-lpthread", "-ldl", "-lz",
              f"-Wl,-rpath,{HERE}", "-Wl,-rpath,$ORIGIN"])
    main = os.path.join(HOST, "darcy_main.cpp")
    if os.path.exists(main) and (force or _newer(EXE, [main, LIB_HOST])):
        _run(common + ["-o", EXE, main, "-L", HERE, "-lmmmfe_host", "-lmmmfe_b200",
                       f"-Wl,-rpath,{HERE}", "-Wl,-rpath,$ORIGIN/mmmfe-st-dd_b200"])
    return LIB_HOST


def build_all(force=False, verbose=False):
    build_engine(force=force, verbose=verbose)
    build_host(force=force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built:", LIB_ENGINE)
