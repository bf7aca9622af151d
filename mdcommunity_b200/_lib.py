"""ctypes binding of libmdb200.so (the C ABI declared in include/mdcommunity_b200.h).

There is no CPU fallback: if the library cannot be built/loaded the import of the compute path fails loudly.
"""
import ctypes as C
import os

from . import build as _build

MAX_LAYERS = 4
EMBED = 64
HIDDEN = 32
AUX = 4
GATE = 128
BP_ITERS = 3
TILE_M = 128
GCONST = 16

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_vp = C.c_void_p


class MdGraphs(C.Structure):
    _fields_ = [
        ("n_graphs", C.c_int32), ("n_layers", C.c_int32), ("stride", C.c_int32), ("n_total", C.c_int32),
        ("max_graph_nnz", C.c_int32), ("reserved0", C.c_int32),
        ("rowptr", _vp * MAX_LAYERS), ("col", _vp * MAX_LAYERS), ("src", _vp * MAX_LAYERS),
        ("n_nodes", _vp), ("n_edges", _vp), ("max_rank", _vp),
        ("node_w", _vp * MAX_LAYERS), ("node_cost", _vp),
    ]


class MdState(C.Structure):
    _fields_ = [
        ("label", _vp), ("deg", _vp * MAX_LAYERS), ("act_list", _vp), ("n_act", _vp), ("aptr", _vp * MAX_LAYERS), ("acol", _vp * MAX_LAYERS), ("asrc", _vp * MAX_LAYERS), ("n_removed", _vp),
        ("cov_edges", _vp), ("alive_edges", _vp), ("max_deg", _vp), ("twohop", _vp), ("rank", _vp),
        ("done", _vp), ("n_steps", _vp), ("score", _vp), ("aux", _vp), ("best", _vp),
        ("trace_action", _vp), ("trace_rank", _vp), ("trace_maxcc", _vp), ("n_done", _vp),
        ("max_steps", C.c_int32), ("cost_mode", C.c_int32),
        ("parent", _vp), ("comp_size", _vp), ("flags", _vp), ("cur_action", _vp), ("tail_n", _vp), ("lg_scratch", _vp),
    ]


class MdWeights(C.Structure):
    _fields_ = [(n, _vp) for n in (
        "w_n2l", "conv1", "conv2", "conv3", "w13", "w23", "trans", "bias", "logis_w", "logis_b", "cross", "h1", "h2",
        "wl1", "wl2")]


class MdWork(C.Structure):
    _fields_ = [("h", (_vp * MAX_LAYERS) * 2), ("psum", _vp), ("gconst", _vp), ("q", _vp), ("fused_ws", _vp)]


_LIB = None

_SIGS = {
    "md_env_reset": ([C.POINTER(MdGraphs), C.POINTER(MdState), _vp], C.c_int),
    "md_env_set_state": ([C.POINTER(MdGraphs), C.POINTER(MdState), _vp, _vp], C.c_int),
    "md_env_step": ([C.POINTER(MdGraphs), C.POINTER(MdState), _vp, C.c_int32, C.c_int32, _vp], C.c_int),
    "md_env_small_max_stride": ([], C.c_int32),
    "md_debug_force_large_env": ([C.c_int32], None),
    "md_debug_set_flags": ([C.c_int32], None),
    "md_debug_set_fused_pack": ([C.c_int32], None),
    "md_debug_read_env": ([_vp, C.c_int32], C.c_int),
    "md_debug_read_env_ctas": ([_vp], C.c_int),
    "md_debug_read_fused": ([_vp, C.c_int32], C.c_int),
    "md_debug_read_fused_ctas": ([_vp], C.c_int),
    "md_debug_read_fused_gather": ([_vp], C.c_int),
    "md_env_uses_small_path": ([C.POINTER(MdGraphs)], C.c_int32),
    "md_embed_init": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWeights), C.POINTER(MdWork), _vp], C.c_int),
    "md_embed_iter": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWeights), C.POINTER(MdWork), C.c_int32, _vp], C.c_int),
    "md_embed_fused": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWeights), C.POINTER(MdWork), _vp], C.c_int),
    "md_fused_max_stride": ([], C.c_int32),
    "md_fused_applies": ([C.POINTER(MdGraphs)], C.c_int32),
    "md_fused_ws_bytes": ([C.POINTER(MdGraphs)], C.c_int64),
    "md_graph_embed": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWeights), C.POINTER(MdWork), _vp], C.c_int),
    "md_qhead": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWeights), C.POINTER(MdWork), _vp], C.c_int),
    "md_rollout_steps": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWeights), C.POINTER(MdWork), C.c_int32, _vp], C.c_int),
    "md_tail_applies": ([C.POINTER(MdGraphs)], C.c_int32),
    "md_tail_max_nodes": ([], C.c_int32),
    "md_tail_max_entries": ([], C.c_int32),
    "md_tail_handover": ([C.POINTER(MdGraphs), C.POINTER(MdState), _vp], C.c_int),
    "md_rollout_tail": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWeights), _vp], C.c_int),
    "md_tail_join": ([_vp], C.c_int),
    "md_debug_read_tail": ([_vp], C.c_int),
    "md_heuristic_select": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.c_int32, C.c_int32, _vp], C.c_int),
    "md_rollout_graph_create": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWeights), C.POINTER(MdWork), C.c_int32, _vp,
                                 C.POINTER(C.c_void_p)], C.c_int),
    "md_rollout_graph_launch": ([_vp, _vp], C.c_int),
    "md_rollout_graph_destroy": ([_vp], C.c_int),
    "md_select_topk": ([C.POINTER(MdGraphs), C.POINTER(MdState), C.POINTER(MdWork), _vp, C.c_int32, _vp], C.c_int),
    "md_spmm_csr_f32": ([_vp, _vp, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp], C.c_int),
    "md_gemm_f32": ([_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, C.c_int64, _vp], C.c_int),
    "md_norm_fwd_f32": ([_vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp], C.c_int),
    "md_norm_bwd_f32": ([_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, _vp], C.c_int),
    "md_tanh_bias_fwd_f32": ([_vp, _vp, _vp, C.c_int32, C.c_int32, _vp], C.c_int),
    "md_tanh_bwd_f32": ([_vp, _vp, _vp, C.c_int32, C.c_int32, _vp], C.c_int),
    "md_colsum_f32": ([_vp, _vp, C.c_int32, C.c_int32, _vp, C.c_int64, _vp], C.c_int),
    "md_fuse2_fwd_f32": ([_vp] * 7 + [C.c_int32, _vp], C.c_int),
    "md_fuse2_bwd_f32": ([_vp] * 10 + [C.c_int32, _vp], C.c_int),
    "md_rowscale_f32": ([_vp, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp], C.c_int),
    "md_rowdot_f32": ([_vp, _vp, _vp, C.c_int32, C.c_int32, _vp], C.c_int),
    "md_seg_reduce_f32": ([_vp, _vp, _vp, _vp, C.c_int32, _vp], C.c_int),
    "md_head_mix_fwd_f32": ([_vp] * 8 + [C.c_int32, _vp], C.c_int),
    "md_head_mix_bwd_f32": ([_vp] * 14 + [C.c_int32, C.c_int32, _vp], C.c_int),
    "md_axpby_f32": ([_vp, _vp, C.c_float, C.c_float, _vp, C.c_int64, _vp], C.c_int),
    "md_relu_mask_f32": ([_vp, _vp, C.c_int64, _vp], C.c_int),
    "md_dot_f32": ([_vp, _vp, _vp, C.c_int64, _vp], C.c_int),
    "md_adam_step_f32": ([_vp, _vp, _vp, _vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32, _vp], C.c_int),
    "md_psum_floats": ([C.c_int32, C.c_int32], C.c_int64),
    "md_abi_sizes": ([_vp], None),
    "md_version": ([], C.c_char_p),
}


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """Load (building first if the sources changed) the shared library; raises if unavailable."""
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if not os.path.exists(path) or os.environ.get("MDB200_REBUILD"):
            path = _build.build()
        else:
            try:
                path = _build.build()  # no-op when the stamp matches
            except FileNotFoundError:
                # a box without nvcc uses the prebuilt .so that travelled with the snapshot -- but only if it was built
                # from exactly these sources (a compile ERROR is never swallowed: it propagates)
                if not _build.up_to_date():
                    raise RuntimeError("libmdb200.so is stale (sources changed since it was built) and nvcc is not "
                                       "available to rebuild it")
        h = C.CDLL(path)
        for name, (args, res) in _SIGS.items():
            fn = getattr(h, name)  # AttributeError if the symbol is missing: fail loudly
            fn.argtypes = args
            fn.restype = res
        sizes = (C.c_int32 * 4)()
        h.md_abi_sizes(sizes)
        mine = [C.sizeof(MdGraphs), C.sizeof(MdState), C.sizeof(MdWeights), C.sizeof(MdWork)]
        if list(sizes) != mine:
            raise RuntimeError("ctypes struct mirrors %s do not match the compiled ABI %s" % (mine, list(sizes)))
        _LIB = h
    return _LIB


def check(rc, what):
    if rc != 0:
        if rc > 0:
            raise RuntimeError("%s: CUDA error %d" % (what, rc))
        raise RuntimeError("%s: %s" % (what, {-1: "invalid argument", -2: "size not supported", -3: "unsupported layer count"}.get(rc, rc)))
