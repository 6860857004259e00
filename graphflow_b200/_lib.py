"""ctypes binding of the C-ABI (include/ccn_b200.h).  No compute happens in Python and there is no fallback:
if graphflow_b200/libccn_b200.so is missing or a call fails, a CCNError is raised."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CCN_B200_LIB") or os.path.join(HERE, "libccn_b200.so")  # the override is for A/B builds (profiles/)

ADJ_POSITIVE_PART = 0  # RisiContraction_18 semantics (RisiContraction_18.h:90)
PATH_AUTO, PATH_GENERIC = 0, 1  # ccn_ctx_set_kernel_path
MIX_AUTO, MIX_SIMT, MIX_TENSOR = 0, 1, 2  # ccn_ctx_set_mix_path
ADJ_RAW = 1  # RisiContraction_18_thread / _50 semantics (RisiContraction_18_thread.h:70-72)

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_void_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/ccn_b200.h one to one (tests/test_abi_cpu.py checks the symbol list
# against the header).
SIGNATURES = {
    "ccn_abi_version": (ctypes.c_int, []),
    "ccn_status_string": (ctypes.c_char_p, [ctypes.c_int]),
    "ccn_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "ccn_ctx_create": (ctypes.c_int, [c_void_pp, ctypes.c_int]),
    "ccn_ctx_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "ccn_ctx_set_workspace_limit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    "ccn_ctx_kernel_launches": (ctypes.c_int64, [ctypes.c_void_p]),
    "ccn_ctx_set_kernel_path": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "ccn_ctx_set_mix_path": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "ccn_ctx_fused_error_flag": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]),
    "ccn_ctx_set_frozen": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "ccn_ctx_set_phase_trace": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "ccn_num_kernels": (ctypes.c_int, []),
    "ccn_kernel_name": (ctypes.c_char_p, [ctypes.c_int]),
    "ccn_ctx_set_kernel_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "ccn_ctx_get_kernel_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                                 ctypes.POINTER(ctypes.c_int64)]),
    "ccn_contract18_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                              ctypes.c_int, ctypes.c_void_p]),
    "ccn_contract18_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int, ctypes.c_float, ctypes.c_void_p]),
    "ccn_contract50_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                              ctypes.c_int, ctypes.c_void_p]),
    "ccn_contract50_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int, ctypes.c_float, ctypes.c_void_p]),
    "ccn_contract_family_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p,
                                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                   ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                                   ctypes.c_void_p]),
    "ccn_contract_family_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p,
                                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                    ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                                    ctypes.c_void_p]),
    "ccn_contract18_forward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int]),
    "ccn_contract18_backward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                                    ctypes.c_float]),
    "ccn_contract18_forward_backward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                            ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int]),
    "ccn_promote_forward": (ctypes.c_int, [ctypes.c_void_p] * 7 + [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                               ctypes.c_void_p]),
    "ccn_promote_backward": (ctypes.c_int, [ctypes.c_void_p] * 7 + [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                                ctypes.c_void_p]),
    "ccn_tensor_mul_forward": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] * 4 + [ctypes.c_int64, ctypes.c_void_p]),
    "ccn_tensor_mul_backward": (ctypes.c_int, [ctypes.c_void_p] * 6 + [ctypes.c_int] * 4 + [ctypes.c_int64, ctypes.c_float,
                                                                                          ctypes.c_void_p]),
    "ccn_custom_matmul_tensor_forward": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                                                ctypes.c_void_p]),
    "ccn_custom_matmul_tensor_backward": (ctypes.c_int, [ctypes.c_void_p] * 6 + [ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                                                 ctypes.c_float, ctypes.c_void_p]),
    "ccn_mix_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_float, ctypes.c_void_p]),
    "ccn_mix_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                        ctypes.c_float, ctypes.c_void_p]),
    "ccn_device_alloc": (ctypes.c_int, [ctypes.c_void_p, c_void_pp, ctypes.c_size_t]),
    "ccn_device_free": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "ccn_h2d": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "ccn_d2h": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "ccn_memset_zero": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "ccn_gather_contract18_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    "ccn_gather_contract18_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    "ccn_level_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]),
    "ccn_level_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]),
    "ccn_graph_tables_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "ccn_graph_tables_destroy": (None, [ctypes.c_void_p]),
    "ccn_graph_tables_feature_width": (ctypes.c_int, [ctypes.c_void_p]),
    "ccn_graph_tables_features": (ctypes.POINTER(ctypes.c_double), [ctypes.c_void_p]),
    "ccn_graph_tables_rank": (ctypes.POINTER(ctypes.c_int32), [ctypes.c_void_p]),
    "ccn_graph_tables_field": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.POINTER(ctypes.c_int32))]),
    "ccn_graph_tables_vertex": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.POINTER(ctypes.c_float)),
                                               ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)), ctypes.POINTER(ctypes.POINTER(ctypes.c_int32)),
                                               ctypes.POINTER(ctypes.POINTER(ctypes.c_int32))]),
    "ccn_allreduce_grads": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "ccn_adam_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                     ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]),
    "ccn_momentum_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_void_p]),
    "ccn_stream_synchronize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "ccn_stream_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]),
    "ccn_gather_level_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]),
    "ccn_gather_level_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p]),
    "ccn_gather_levels_forward_backward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]),
    "ccn_gather_levels_readout_forward_backward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]),
    "ccn_gather_level_forward_backward_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float]),
    "ccn_level_features_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "ccn_level_features_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "ccn_host_register": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "ccn_host_unregister": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "ccn_readout_forward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "ccn_readout_backward": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "ccn_stream_destroy": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
}


class CCNError(RuntimeError):
    pass


_lib = None


def load():
    """dlopen the C-ABI library and attach the prototypes.  Raises CCNError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CCNError("%s not found: build it with `python -m graphflow_b200.build` (nvcc, sm_100a). "
                       "There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header / library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
