"""In-tree build of the edl_b200 native extension (``edl_b200/_C*.so``).

Kernels are torch-free ``.cu`` files compiled straight with nvcc for sm_100a; only the binding
``.cpp`` files see torch headers.  Objects are cached by mtime under ``build/`` so a rebuild after
touching one kernel takes seconds.  The resulting ``.so`` lives in the package directory so that it
travels to the GPU box with the repo snapshot.

Usage:  python -m edl_b200.build_ext [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(ROOT, "build", "edl_b200")
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")
TARGET = os.path.join(HERE, "_C" + EXT_SUFFIX)

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _torch_paths():
    import torch
    from torch.utils import cpp_extension as ce

    inc = ce.include_paths()
    lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    return inc, lib, abi


def _newer(src: str, obj: str, deps) -> bool:
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src, *deps])


def _run(cmd, verbose):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    if verbose:
        sys.stdout.write(r.stdout)
    return r.stdout


def sources():
    cu = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    cpp = sorted(f for f in os.listdir(CSRC) if f.endswith(".cpp"))
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return cu, cpp, hdr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA/C++ source for sm_100a and link ``_C``; returns the .so path."""
    os.makedirs(BUILD, exist_ok=True)
    cu, cpp, hdr = sources()
    inc, torch_lib, abi = _torch_paths()
    py_inc = sysconfig.get_paths()["include"]
    jobs = []
    objs = []
    ptxas_log = []
    for f in cu:
        src = os.path.join(CSRC, f)
        obj = os.path.join(BUILD, f[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, hdr):
            jobs.append(([NVCC, *NVCC_FLAGS, "-I", CSRC, "-c", src, "-o", obj], f))
    for f in cpp:
        src = os.path.join(CSRC, f)
        obj = os.path.join(BUILD, f[:-4] + ".cpp.o")
        objs.append(obj)
        if force or _newer(src, obj, hdr):
            cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-DTORCH_EXTENSION_NAME=_C",
                   "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % abi,
                   "-I", CSRC, "-I", py_inc, "-I", "/usr/local/cuda/include"]
            for i in inc:
                cmd += ["-isystem", i]
            cmd += ["-c", src, "-o", obj]
            jobs.append((cmd, f))

    def _do(job):
        cmd, name = job
        out = _run(cmd, False)
        return name, out

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for name, out in ex.map(_do, jobs):
                ptxas_log.append("== %s ==\n%s" % (name, out))
                if verbose:
                    print("compiled", name)
    relink = force or bool(jobs) or not os.path.exists(TARGET)
    if relink:
        cmd = ["g++", "-shared", "-o", TARGET, *objs, "-L", torch_lib, "-lc10", "-ltorch_cpu",
               "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
               "-L", "/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread",
               "-Wl,-rpath," + torch_lib]
        _run(cmd, verbose)
    if ptxas_log:
        with open(os.path.join(BUILD, "ptxas.log"), "a") as fh:
            fh.write("\n".join(ptxas_log))
    return TARGET


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv)
    print(path)
