"""ctypes binding of libmatinvent_b200.so (include/matinvent_b200.h).

There is NO fallback: if the shared library is missing or a call fails, the product path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmatinvent_b200.so")

c_fp = C.c_void_p   # device pointers travel as integers (tensor.data_ptr())


class Epilogue(C.Structure):
    _fields_ = [("bias", c_fp),
                ("g1", c_fp), ("g1_idx", c_fp), ("g1_ld", C.c_int),
                ("g2", c_fp), ("g2_idx", c_fp), ("g2_ld", C.c_int),
                ("g3", c_fp), ("g3_idx", c_fp), ("g3_ld", C.c_int),
                ("z_out", c_fp), ("z_ld", C.c_int),
                ("z_in", c_fp), ("zin_ld", C.c_int),
                ("resid", c_fp), ("resid_ld", C.c_int),
                ("act", C.c_int), ("alpha", C.c_float), ("beta", C.c_float), ("splitk", C.c_int),
                ("amax_out", c_fp), ("a_amax", c_fp), ("col_scale", c_fp),
                ("scat_out", c_fp), ("scat_ld", C.c_int), ("scat_idx", c_fp), ("scat_w", c_fp), ("scat_amax", c_fp)]


i, f, d, ll, u64, p = C.c_int, C.c_float, C.c_double, C.c_longlong, C.c_ulonglong, c_fp

# name -> argtypes (restype is int unless noted); mirrors include/matinvent_b200.h one to one
PROTOTYPES = {
    "mi_version": [],
    "mi_device_info": [C.POINTER(i), C.POINTER(i), C.POINTER(i)],
    "mi_sgemm": [i, i, i, i, i, p, i, p, i, p, i, C.POINTER(Epilogue), p],
    "mi_f16_split": [p, p, p, ll, f, f, p],
    "mi_f16_split_rows": [p, i, i, i, p, p, p, p],
    "mi_transpose_amax": [p, i, i, i, p, i, p, p],
    "mi_transpose_split": [p, i, i, i, p, p, p, i, p, p],
    "mi_tc_gemm": [i, i, i, p, i, p, p, i, p, i, C.POINTER(Epilogue), i, p],
    "mi_tc_gemm_presplit": [i, i, i, p, p, i, p, p, i, p, i, C.POINTER(Epilogue), i, p],
    "mi_node_chain": [i, i, i, p, i, p, i, p, p, p, p, p, p, i, p, p, i, p, p, p, p, p, p, i, p, i, p, p, f, p, p, p, i, p, p, i, p, p],
    "mi_edge_block1": [i, i, i, p, p, i, p, p, i, p, f, p, p, i, p, p, p, p, p, p, i, p, p],
    "mi_edge_block2": [i, i, i, p, p, i, p, p, p, i, p, p, p, i, p, p, p, p],
    "mi_fc_edges": [p, p, i, i, i, p, p, p, p, p, p, p, p],
    "mi_edge_fourier": [p, p, p, p, i, i, p, p, i, p, p, f, f, p],
    "mi_segment_reduce": [p, i, p, p, p, i, i, i, i, i, p, p],
    "mi_gather_rows_dsilu": [p, i, p, p, p, i, p, i, i, i, p, p],
    "mi_row_amax": [p, i, i, i, p, p],
    "mi_colsum": [p, i, i, i, p, i, p],
    "mi_layernorm_fwd": [p, i, p, p, p, i, p, p, i, i, f, p, p],
    "mi_layernorm_fwd_split": [p, i, p, p, p, i, p, p, i, p, p, i, i, p, p, i, i, f, p],
    "mi_layernorm_bwd": [p, i, p, i, p, p, p, p, i, i, p, p, i, i, p],
    "mi_lattice_ip": [p, p, i, p],
    "mi_lattice_linear": [p, p, p, p, i, i, i, i, ll, ll, ll, p],
    "mi_output_heads": [p, i, p, i, i, p, p, f, p, p, p, p, i, p, p, p, i, p, p],
    "mi_bmm3": [p, p, p, i, i, p],
    "mi_time_embed": [p, p, i, i, p, p],
    "mi_lattice_params_to_matrix": [p, p, p, i, p],
    "mi_lattice_matrix_to_params": [p, p, p, i, p],
    "mi_argmax_rows": [p, i, i, i, i, p, p],
    "mi_reverse_corrector": [p, p, p, p, i, p, p, i, p],
    "mi_reverse_predictor": [p, p, p, p, i, p, p, p, i, p, p, p, i, p, p, i, p],
    "mi_sampler_step_begin": [p, p, p, i, i, p],
    "mi_sampler_step_end": [p, p],
    "mi_add_noise": [p, p, p, p, p, p, i, i, i, p, p, i, p, p, p, p, p],
    "mi_rl_loss": [p, p, p, p, p, p, p, p, p, p, i, i, f, f, f, p, p, f, p, p, p, p, p, p, p],
    "mi_adam_step": [p, p, p, p, ll, d, d, d, d, i, f, i, p],
    "mi_philox_normal": [p, ll, u64, u64, p, i, p],
    "mi_philox_uniform": [p, ll, u64, u64, p, i, p],
    "mi_radius_graph_pbc": [p, p, p, i, i, i, i, i, p, p, p, p, p],
    "mi_compact_edges": [p, i, i, p, p, p, p, p, p, p, p, i, p],
    "mi_build_dst_csr": [p, p, i, i, p, p, p, p],
    "mi_replay_select": [p, p, i, i, d, p, p, p],
    "mi_composition_key": [p, p, i, p, p],
    "mi_weighted_field_sum": [i, i, p, p, p, p, C.POINTER(f), p, p],
    "mi_validity_prefilter": [p, p, p, p, i, f, f, f, f, p, p, p],
    "mi_composition_reward": [p, p, i, p, p, i, C.POINTER(i), C.POINTER(i), C.POINTER(d), C.POINTER(d), C.POINTER(d),
                              C.POINTER(d), i, p, p, p, p],
}

_lib = None


class MatInventLibError(RuntimeError):
    pass


def load(path=None):
    """Load the shared library (once).  Raises if it has not been built: no CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("MATINVENT_B200_LIB", LIB_PATH)
    if not os.path.isfile(path):
        raise MatInventLibError(
            "libmatinvent_b200.so not found at %s — build it with `python -m matinvent_b200.csrc.build` "
            "(or __graft_entry__.build()); matinvent_b200 has no CPU fallback." % path)
    lib = C.CDLL(path)
    lib.mi_last_error.restype = C.c_char_p
    lib.mi_last_error.argtypes = []
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)     # AttributeError here = header/library mismatch
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = lib
    return lib


# kernels launched per C call (default 1); `launches` is the running total the bench reports
KERNELS_PER_CALL = {"mi_layernorm_bwd": 2, "mi_compact_edges": 2, "mi_build_dst_csr": 4, "mi_colsum": 1}
launches = 0


def check(rc, what):
    global launches
    launches += KERNELS_PER_CALL.get(what, 1)
    if rc != 0:
        msg = _lib.mi_last_error().decode("utf-8", "replace") if _lib is not None else ""
        raise MatInventLibError("%s failed (rc=%d): %s" % (what, rc, msg))
