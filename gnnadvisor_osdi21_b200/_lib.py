"""ctypes binding of libgnna_b200.so (C ABI: include/gnna_b200.h).

There is NO fallback: if the library is missing or a call fails, this raises.  Nothing in this
package computes an aggregation on the CPU or through torch ops.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# GNNA_B200_LIB selects another build of the same library (tuning variants, tools/sweep_dims.py)
LIB_PATH = os.environ.get("GNNA_B200_LIB") or os.path.join(_PKG, "libgnna_b200.so")

c_i32p = ctypes.c_void_p
c_f32p = ctypes.c_void_p
i64 = ctypes.c_int64
i32 = ctypes.c_int


class LaunchInfo(ctypes.Structure):
    _fields_ = [("vec_width", i32), ("lanes_per_row", i32), ("chunks_per_lane", i32),
                ("warps_per_block", i32), ("groups_per_warp", i32), ("grid_x", i64),
                ("grid_y", i32), ("kernels", i32)]


# name -> (restype, argtypes); the single place that mirrors include/gnna_b200.h
_GRAPH = [c_i32p, c_i32p]                       # row_ptr, col_idx
_PARTS = [c_i32p, c_i32p]                       # part_ptr, part2node
_TUNE = [i32, i32, i32, ctypes.c_void_p]        # part_size, dim_worker, warp_per_block, stream
SIGNATURES = {
    "gnna_abi_version": (i32, []),
    "gnna_last_error": (ctypes.c_char_p, []),
    "gnna_count_parts_host": (i64, [i32, c_i32p, i64]),
    "gnna_build_part_host": (i32, [i32, c_i32p, i64, c_i32p, c_i32p, i64, i32]),
    "gnna_build_part_device": (i32, [i32, c_i32p, i64, c_i32p, c_i32p, ctypes.POINTER(i64),
                                     ctypes.c_void_p, i64, ctypes.c_void_p]),
    "gnna_build_part_workspace_bytes": (i64, [i64]),
    "gnna_degrees": (i32, [c_i32p, i64, c_f32p, ctypes.c_void_p]),
    "gnna_sag_f32": (i32, [c_f32p, c_f32p] + _GRAPH + _PARTS + [i64, i32, i64] + _TUNE),
    "gnna_gcn_aggregate_f32": (i32, [c_f32p, c_f32p] + _GRAPH + [c_f32p] + _PARTS + [i64, i32, i64] + _TUNE),
    "gnna_gin_aggregate_f32": (i32, [c_f32p, c_f32p] + _GRAPH + [ctypes.c_float] + _PARTS + [i64, i32, i64] + _TUNE),
    "gnna_aggregate_f32_ex": (i32, [i32, c_f32p, i64, c_f32p, i64] + _GRAPH + [c_f32p, ctypes.c_float] + _PARTS
                              + [i32, i64] + _TUNE),
    "gnna_aggregate_part_f32_ex": (i32, [i32, i32, c_f32p, i64, c_f32p, i64] + _GRAPH + [c_f32p, ctypes.c_float] + _PARTS
                                   + [i32, i64] + _TUNE),
    "gnna_aggregate_gated_f32": (i32, [i32, c_f32p, i64, c_f32p, i64] + _GRAPH + [c_f32p, ctypes.c_float] + _PARTS
                                 + [i32, i64, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int32), i32, ctypes.c_void_p] + _TUNE),
    "gnna_prescale_rows_f32": (i32, [c_f32p, c_f32p, c_f32p, i64, i32, ctypes.c_void_p]),
    "gnna_sgemm_f32": (i32, [i32, i32, i64, i64, i64, c_f32p, c_f32p, c_f32p, ctypes.c_void_p]),
    "gnna_aggregate_bf16": (i32, [i32, ctypes.c_void_p, c_f32p] + _GRAPH + [c_f32p, ctypes.c_float] + _PARTS
                            + [i64, i32, i64] + _TUNE),
    "gnna_scale_rows_bf16": (i32, [c_f32p, ctypes.c_void_p, c_f32p, i64, i32, i32, ctypes.c_void_p]),
    "gnna_aggregate_bf16_ex": (i32, [i32, ctypes.c_void_p, i32, c_f32p] + _GRAPH + [c_f32p, ctypes.c_float] + _PARTS
                               + [i64, i32, i64] + _TUNE),
    "gnna_forward_mixed": (i32, [c_f32p, c_f32p, c_f32p, ctypes.c_void_p, c_f32p] + _GRAPH + [c_f32p] + _PARTS
                           + [i64, i32, i32, i64] + _TUNE),
    "gnna_backward_mixed": (i32, [c_f32p, c_f32p, c_f32p, ctypes.c_void_p, c_f32p, c_f32p, c_f32p] + _GRAPH + [c_f32p] + _PARTS
                            + [i64, i32, i32, i64] + _TUNE),
    "gnna_forward_gin_mixed": (i32, [c_f32p, c_f32p, ctypes.c_float, ctypes.c_void_p, c_f32p, c_f32p] + _GRAPH + _PARTS
                               + [i64, i32, i32, i64] + _TUNE),
    "gnna_backward_gin_mixed": (i32, [c_f32p, c_f32p, c_f32p, ctypes.c_float, c_f32p, ctypes.c_void_p, c_f32p, c_f32p]
                                + _GRAPH + _PARTS + [i64, i32, i32, i64] + _TUNE),
    "gnna_forward_f32": (i32, [c_f32p] * 4 + _GRAPH + [c_f32p] + _PARTS + [i64, i32, i32, i64] + _TUNE),
    "gnna_backward_f32": (i32, [c_f32p] * 6 + _GRAPH + [c_f32p] + _PARTS + [i64, i32, i32, i64] + _TUNE),
    "gnna_forward_gin_f32": (i32, [c_f32p, c_f32p, ctypes.c_float, c_f32p, c_f32p] + _GRAPH + _PARTS
                             + [i64, i32, i32, i64] + _TUNE),
    "gnna_backward_gin_f32": (i32, [c_f32p, c_f32p, c_f32p, ctypes.c_float, c_f32p, c_f32p, c_f32p] + _GRAPH + _PARTS
                              + [i64, i32, i32, i64] + _TUNE),
    "gnna_aggregate_gemm_fused_bf16": (i32, [i32, ctypes.c_void_p, i32, c_f32p, ctypes.c_float, c_f32p, c_f32p] + _GRAPH
                                       + [c_f32p] + _PARTS + [i64, i32, i32, i64] + _TUNE),
    "gnna_ipc_alloc": (i32, [i64, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_ubyte * 64)]),
    "gnna_ipc_open": (i32, [ctypes.POINTER(ctypes.c_ubyte * 64), ctypes.POINTER(ctypes.c_void_p)]),
    "gnna_ipc_close": (i32, [ctypes.c_void_p]),
    "gnna_ipc_free": (i32, [ctypes.c_void_p]),
    "gnna_halo_push_f32": (i32, [c_f32p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_void_p),
                                 ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p,
                                 i32, i32, i32, ctypes.c_void_p]),
    "gnna_halo_push_ce": (i32, [c_f32p, i64, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_void_p),
                                ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p,
                                i32, i32, i32, ctypes.c_uint32, ctypes.c_void_p]),
    "gnna_halo_begin_step": (i32, [ctypes.c_void_p, ctypes.c_void_p]),
    "gnna_halo_wait": (i32, [ctypes.c_void_p, i32, i32, ctypes.c_uint32, ctypes.c_void_p]),
    "gnna_halo_ack": (i32, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, i32, i32, ctypes.c_void_p]),
    "gnna_rabbit_reorder_host": (i32, [c_i32p, c_i32p, i64, i64, c_i32p]),
    "gnna_rabbit_reorder_host_ex": (i32, [c_i32p, c_i32p, i64, i64, c_i32p, i64]),
    "gnna_edge_text_scan": (i32, [ctypes.c_char_p, ctypes.POINTER(i64)]),
    "gnna_edge_text_parse": (i32, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, i64, ctypes.POINTER(i64), ctypes.POINTER(i64)]),
    "gnna_csr_from_edges_host": (i32, [ctypes.c_void_p, ctypes.c_void_p, i64, i64, c_i32p, c_i32p, ctypes.POINTER(i64)]),
    "gnna_query_launch": (i32, [i32, i32, i64, i32, i32, ctypes.POINTER(LaunchInfo)]),
    "gnna_stream_pairs": (i32, [i64, i64, ctypes.c_uint64, i64, i32, i32, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64,
                                ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "gnna_probe_l2_read": (i32, [ctypes.c_void_p, i64, i32, i32, i32, i32, ctypes.c_void_p, ctypes.POINTER(i64), ctypes.c_void_p]),
    "gnna_launch_count": (i64, [i32]),
    "gnna_set_gcn_exact": (i32, [i32]),
    "gnna_set_tc_gemm": (i32, [i32]),
    "gnna_set_small_parts": (i64, [i64]),
    "gnna_set_staged": (i32, [i32]),
    "gnna_set_runs": (i32, [i32]),
    "gnna_query_runs": (i32, [i32, i32, i64, i64]),
}

_lib = None


def load():
    """Load the C-ABI library; raises if it has not been built (python -m gnnadvisor_osdi21_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libgnna_b200.so is missing (%s). Build it with `python -m gnnadvisor_osdi21_b200.build`; "
            "there is no CPU or torch fallback for this path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().gnna_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count(reset=False):
    return int(load().gnna_launch_count(1 if reset else 0))


def set_gcn_exact(on):
    """True: per-edge rounding of the reference (bit-identical single-group rows); False: pre-scaled (default)."""
    return bool(load().gnna_set_gcn_exact(1 if on else 0))


def set_tc_gemm(mode):
    """Which tall-skinny products of a layer (X*W, X^T*G) run on the tensor cores (csrc/gemm_tf32x3.cu: tcgen05 kind::tf32,
    3xTF32 split): 0 none (cuBLAS SGEMM everywhere), 1 both on the lockstep kernel, 2 both on the warp-specialised kernel,
    3 (default) X^T*G on the warp-specialised kernel and X*W on cuBLAS -- the measured winners.  Also GNNA_TC_GEMM in the
    environment.  Returns the previous mode."""
    return int(load().gnna_set_tc_gemm(int(mode)))


def set_small_parts(limit):
    """Group tables of at most `limit` groups take the single-launch row-owned kernel (csrc/aggregate_small.cu); 0: never.
    Returns the previous limit."""
    return int(load().gnna_set_small_parts(int(limit)))


def set_staged(on):
    """True: TMA-staged persistent aggregation kernel where it applies (csrc/aggregate_staged.cu)."""
    return bool(load().gnna_set_staged(1 if on else 0))


def set_runs(run):
    """run > 0: run-based software-pipelined aggregation kernel (csrc/aggregate_runs.cu), `run` groups per sub-warp;
    0: never; -1 (default): the library chooses.  Returns the previous setting."""
    return int(load().gnna_set_runs(int(run)))


def query_runs(elem_bytes, row_elems, num_nodes, num_parts):
    """Run length the library would use for this shape (0: the one-group-per-sub-warp kernel).  Host-only."""
    return int(load().gnna_query_runs(int(elem_bytes), int(row_elems), int(num_nodes), int(num_parts)))
