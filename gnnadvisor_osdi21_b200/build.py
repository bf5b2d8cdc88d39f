"""Build libgnna_b200.so (the C-ABI library, include/gnna_b200.h) in-tree with nvcc for sm_100a.

    python -m gnnadvisor_osdi21_b200.build [--force] [--verbose]

The library has no torch dependency: CUDA runtime (static) + cuBLAS only.  nvcc cross-compiles
without a GPU, so this runs in the authoring container; the .so travels to the GPU box.
"""
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libgnna_b200.so")
SOURCES = ["aggregate.cu", "build_part.cu", "ops.cu"]
HEADERS = [os.path.join(CSRC, "common.h"), os.path.join(PKG, "..", "include", "gnna_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-fvisibility=hidden",
    "--threads", "4",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    env = dict(os.environ)
    env.pop("CC", None)   # this image exports a CC wrapper nvcc must not pick up
    env.pop("CXX", None)
    procs = []
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + ["-Xptxas", "-v" if verbose else "-warn-spills",
                                       "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [_nvcc(), "-shared", "-o", LIB] + objs + [
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lcublas", "-Xlinker", "-rpath,/usr/local/cuda/lib64",
    ]
    subprocess.run(link, check=True, env=env)
    for o in objs:
        os.remove(o)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
