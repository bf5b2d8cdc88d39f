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
SOURCES = ["aggregate.cu", "build_part.cu", "ops.cu", "reorder.cu", "halo.cu", "fused_gemm.cu", "aggregate_staged.cu", "aggregate_runs.cu", "probe.cu", "aggregate_small.cu", "graphgen.cu", "gemm_tf32x3.cu", "dataset.cu"]
HEADERS = [os.path.join(CSRC, "common.h"), os.path.join(CSRC, "gather.cuh"), os.path.join(PKG, "..", "include", "gnna_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-fvisibility=hidden,-fopenmp",
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


def build_lib(force=False, verbose=False, defines=(), out=None):
    """defines/out: build a tuning variant (e.g. defines=["GNNA_MIN_CTAS=3"]) into another file;
    select it at run time with GNNA_B200_LIB=<path> (tools/sweep_dims.py does)."""
    if out is None and not force and not needs_build():
        return LIB
    target = out or LIB
    tag = os.path.basename(target).replace(".so", "")
    objs = []
    env = dict(os.environ)
    env.pop("CC", None)   # this image exports a CC wrapper nvcc must not pick up
    env.pop("CXX", None)
    procs = []
    for s in SOURCES:
        obj = os.path.join(CSRC, tag + "_" + s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + ["-Xptxas", "-v" if verbose else "-warn-spills",
                                       "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [_nvcc(), "-shared", "-o", target] + objs + [
        "-gencode", "arch=compute_100a,code=sm_100a",
        "-lcublas", "-lgomp", "-Xlinker", "-rpath,/usr/local/cuda/lib64",
    ]
    subprocess.run(link, check=True, env=env)
    for o in objs:
        os.remove(o)
    return target


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv, defines=defs, out=outs[0] if outs else None))
