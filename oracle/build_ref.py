"""Compile the reference's own extension from /root/reference into oracle/_ref/.  CHECKER ONLY.

Recipe (no reference build system is run, no reference source is committed):
  * GNNAdvisor/GNNConv/GNNAdvisor.cpp is compiled where it lies;
  * GNNAdvisor/GNNConv/GNNAdvisor_kernel.cu needs a 5-line mechanical patch for torch >= 2
    (`AT_DISPATCH_FLOATING_TYPES(x.type(), ...)` -> `x.scalar_type()`); the
    patched text is produced by the regex below into a scratch file under oracle/_ref/_src/
    (git-ignored) that is deleted again after the build;
  * output: oracle/_ref/GNNAdvisor_ref.so, python module name `GNNAdvisor_ref`, built for
    sm_100a.  It travels to the GPU box with the snapshot (git-ignored, not gpurun-ignored).

Uses: (1) build_part of the reference runs on CPU here -> pins the oracle's build_part
bit-exactly (oracle/make_golden_build_part.py); (2) on the GPU box the reference CUDA kernels
generate golden vectors for the float ops (oracle/make_golden_refgpu.py) and are timed as the
"reference kernels recompiled for sm_100a" line next to ours (bench.py "ref_gpu").
/root/reference does not exist on the GPU box: there only the prebuilt .so is used.
"""
import importlib.util
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = "/root/reference/GNNAdvisor/GNNConv"
OUT_DIR = os.path.join(HERE, "_ref")
SO = os.path.join(OUT_DIR, "GNNAdvisor_ref.so")


def build(verbose=False):
    if os.path.exists(SO):
        return SO
    if not os.path.isdir(REF_DIR):
        raise FileNotFoundError("reference sources not present (GPU box?) and no prebuilt oracle/_ref")
    from torch.utils.cpp_extension import load
    src_tmp = os.path.join(OUT_DIR, "_src")
    os.makedirs(src_tmp, exist_ok=True)
    try:
        cu = open(os.path.join(REF_DIR, "GNNAdvisor_kernel.cu")).read()
        cu, n = re.subn(r"(AT_DISPATCH_FLOATING_TYPES\(\s*\w+)\.type\(\)", r"\1.scalar_type()", cu)
        assert n == 5, "expected exactly 5 dispatch sites, found %d" % n
        patched = os.path.join(src_tmp, "kernel_patched.cu")
        with open(patched, "w") as f:
            f.write(cu)
        os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
        load(name="GNNAdvisor_ref",
             sources=[os.path.join(REF_DIR, "GNNAdvisor.cpp"), patched],
             build_directory=OUT_DIR,
             extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3"],
             is_python_module=False, verbose=verbose)
    finally:
        shutil.rmtree(src_tmp, ignore_errors=True)
        for junk in ("build.ninja", ".ninja_deps", ".ninja_log"):
            p = os.path.join(OUT_DIR, junk)
            if os.path.exists(p):
                os.remove(p)
        for f in os.listdir(OUT_DIR):
            if f.endswith(".o"):
                os.remove(os.path.join(OUT_DIR, f))
    return SO


def load_ref():
    """Import the compiled reference as module `GNNAdvisor_ref` (or None if it is not built)."""
    if not os.path.exists(SO):
        return None
    import torch  # noqa: F401  (the extension links against libtorch)
    spec = importlib.util.spec_from_file_location("GNNAdvisor_ref", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
