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


REF_PY_DIR = "/root/reference/GNNAdvisor"
PY_ZIP = os.path.join(OUT_DIR, "ref_py.zip")    # the reference's Python layer, byte for byte, as ONE git-ignored artefact
REF_PY_FILES = ("gnn_conv.py", "param.py", "dataset.py", "GNNA_main.py", "unitest.py")
_REFMOD = ('"""`import GNNAdvisor` -> the reference\'s own extension compiled for sm_100a (oracle/build_ref.py). CHECKER ONLY."""\n'
           "import sys\n"
           "sys.path.insert(0, %r)\n"
           "import build_ref as _b\n"
           "_m = _b.load_ref()\n"
           "SAG, forward, backward, forward_gin, backward_gin, build_part = (\n"
           "    _m.SAG, _m.forward, _m.backward, _m.forward_gin, _m.backward_gin, _m.build_part)\n")


def stage_py():
    """Pack the reference's own gnn_conv.py / param.py / dataset.py / GNNA_main.py / unitest.py UNCHANGED into
    oracle/_ref/ref_py.zip -- a build artefact next to the compiled extension, git-ignored like it, never part of the
    repository -- so that the GPU box, which has no /root/reference, can run the reference's scripts against this runtime
    (tests/test_reference_gpu.py) and against the reference's own kernels (the R-GPU epoch baseline of bench.py)."""
    import zipfile
    if not os.path.isdir(REF_PY_DIR):
        return os.path.exists(PY_ZIP)
    os.makedirs(OUT_DIR, exist_ok=True)
    with zipfile.ZipFile(PY_ZIP, "w", zipfile.ZIP_DEFLATED) as z:
        for f in REF_PY_FILES:
            z.write(os.path.join(REF_PY_DIR, f), f)
    return True


def unpack_py(dest):
    """Extract the staged reference scripts into `dest`/py (a scratch directory of the caller) and write `dest`/refmod/
    GNNAdvisor.py, a loader that makes `import GNNAdvisor` resolve to GNNAdvisor_ref.so.  Returns (py_dir, refmod_dir) or None."""
    import zipfile
    if not os.path.exists(PY_ZIP):
        return None
    py_dir, refmod = os.path.join(dest, "py"), os.path.join(dest, "refmod")
    os.makedirs(py_dir, exist_ok=True)
    os.makedirs(refmod, exist_ok=True)
    with zipfile.ZipFile(PY_ZIP) as z:
        z.extractall(py_dir)
    with open(os.path.join(refmod, "GNNAdvisor.py"), "w") as f:
        f.write(_REFMOD % HERE)
    return py_dir, refmod


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
    print("reference python layer staged:", stage_py())
