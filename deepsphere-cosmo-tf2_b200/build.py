"""Build libdeepsphere_b200.so (sm_100a only) in-tree with nvcc.

Usage: python build.py [--force] [-v] [--variant NAME -DMACRO=VALUE ...]
The shared library is a plain C-ABI library (include/deepsphere_b200.h); it links only
against the CUDA runtime (static) — no torch, no Python.

`--variant NAME -D...` builds lib/libdeepsphere_b200_NAME.so from the same sources with extra preprocessor
definitions (kernel experiment switches such as -DC2_FENCE_BY_ISSUER=1): load it with DEEPSPHERE_LIB=<path> for an
A/B run against the default build on one GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libdeepsphere_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--use_fast_math=false",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "-cudart", "static",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def deps():
    d = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    d.append(os.path.join(HERE, "..", "include", "deepsphere_b200.h"))
    d.append(os.path.abspath(__file__))
    return d


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(p) <= t for p in deps())


def build(force=False, verbose=False, variant=None, defines=()):
    lib = LIB if variant is None else os.path.join(LIBDIR, f"libdeepsphere_b200_{variant}.so")
    if variant is None and not force and up_to_date():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    procs = []
    tag = "" if variant is None else f".{variant}"
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + tag + ".o")
        objs.append(obj)
        cmd = [NVCC, *[f for f in FLAGS if not f.startswith("--use_fast_math")], *defines, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    cmd = [NVCC, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-lcuda"]
    subprocess.run(cmd, check=True)
    return lib


if __name__ == "__main__":
    name = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=name,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
