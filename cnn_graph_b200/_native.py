"""ctypes binding of ``libcnn_graph_b200.so`` (the C ABI declared in
``include/cnn_graph_b200.h``).  There is NO fallback: if the library is missing or an
entry point fails, a ``NativeError`` is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcnn_graph_b200.so')

ABI_VERSION = 1

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_size_t = ctypes.c_size_t
c_float = ctypes.c_float


class NativeError(RuntimeError):
    pass


# name -> (restype, argtypes); mirrors include/cnn_graph_b200.h one to one
_SIGNATURES = {
    'cg_abi_version': (c_int, []),
    'cg_last_error': (ctypes.c_char_p, []),
    'cg_graph_create': (c_int, [ctypes.POINTER(c_void_p), c_int, c_i64, c_void_p, c_void_p, c_void_p]),
    'cg_graph_destroy': (c_int, [c_void_p]),
    'cg_graph_info': (c_int, [c_void_p, ctypes.POINTER(c_i64)]),
    'cg_cheb_basis': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p]),
    'cg_cheb_step': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_float, c_void_p]),
    'cg_cheb_step_tile_rows': (c_int, [c_void_p, c_int, c_i64]),
    'cg_cheb_step_tiles': (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_i64, c_float, c_void_p, c_int, c_void_p]),
    'cg_halo_pull': (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_int, c_void_p]),
    'cg_cheb_filter_fwd_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    'cg_cheb_filter_bwd_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int]),
    'cg_cheb_filter_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                   c_void_p, c_size_t, c_int, c_void_p]),
    'cg_cheb_filter_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    'cg_cheb_filter_stack_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    'cg_cheb_filter_stack_planes': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int]),
    'cg_cheb_filter_fwd_ex': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_void_p, c_size_t, c_int, c_void_p]),
    'cg_cheb_filter_bwd_ex': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                      c_int, c_int, c_int, c_void_p, c_size_t, c_int, c_void_p]),
    'cg_bias_act_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'cg_bias_act_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                c_void_p]),
    'cg_pool_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'cg_pool_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'cg_bias_act_pool_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_void_p]),
    'cg_bias_act_pool_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_void_p]),
    'cg_gemm_f32_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'cg_gemm_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    'cg_cheb_dw_pooled_supported': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    'cg_cheb_first_layer_fwd_supported': (c_int, [c_void_p, c_int, c_int, c_int, c_int]),
    'cg_cheb_first_layer_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_void_p]),
    'cg_cheb_dw_pooled_workspace_bytes': (c_size_t, [c_void_p, c_int, c_int, c_int]),
    'cg_cheb_dw_pooled': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_void_p, c_size_t, c_void_p]),
    'cg_cheb_contract_workspace_bytes': (c_size_t, [c_i64, c_int, c_int, c_int]),
    'cg_cheb_contract': (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                                 c_void_p]),
    'cg_cheb_contract_dw': (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p, c_size_t,
                                    c_void_p]),
    'cg_bmm_f32': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                           c_i64, c_i64, c_i64, c_void_p]),
    'cg_perm_data': (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p]),
    'cg_lstm_gates_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int,
                                  c_void_p]),
    'cg_lstm_gates_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p]),
    'cg_softmax_xent': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'cg_sgd_momentum': (c_int, [c_void_p, c_int, c_i64, c_float, c_float, c_void_p]),
    'cg_sgd_momentum_dev': (c_int, [c_void_p, c_int, c_i64, c_float, c_void_p, c_float, c_void_p]),
    'cg_adam': (c_int, [c_void_p, c_int, c_i64, c_float, c_float, c_float, c_float, c_void_p, c_void_p]),
    'cg_lstm_gates2_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int,
                                   c_void_p]),
    'cg_lstm_gates2_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p]),
    'cg_csr_densify': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'cg_set_precision': (c_int, [c_int]),
    'cg_get_precision': (c_int, []),
    'cg_launch_count': (c_i64, []),
    'cg_profile_enable': (c_int, [c_int]),
    'cg_profile_reset': (c_int, []),
    'cg_profile_query': (c_int, [c_int, ctypes.c_char_p, c_int, ctypes.POINTER(ctypes.c_double),
                                 ctypes.POINTER(c_i64)]),
    'cg_debug_umma_gemm': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'cg_debug_umma_gemm_m': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'cg_debug_umma_gemm_ts': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'cg_debug_fused_trace': (c_int, [c_void_p]),
    'cg_debug_clenshaw_trace': (c_int, [c_void_p]),
    'cg_debug_fused_plan_info': (c_int, [c_void_p]),
    'cg_debug_gemm_stream': (c_int, [c_int]),
    'cg_host_metis_one_level': (c_int, [c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p,
                                        c_void_p, ctypes.POINTER(c_i64)]),
    'cg_host_metis_one_level_f64': (c_int, [c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_void_p,
                                            c_void_p, ctypes.POINTER(c_i64)]),
    'cg_host_perm_level': (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib():
    """Load (once) and return the native library; raise loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            'native library %s is missing -- build it with `python -m cnn_graph_b200.build` '
            '(there is no CPU / PyTorch fallback for the hot path)' % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError as exc:
            raise NativeError('native library does not export %s' % name) from exc
        fn.restype = restype
        fn.argtypes = argtypes
    got = handle.cg_abi_version()
    if got != ABI_VERSION:
        raise NativeError('ABI mismatch: library reports %d, binding expects %d' % (got, ABI_VERSION))
    _lib = handle
    return _lib


def check(rc, what):
    """Turn a non-zero status of the C ABI into a Python exception."""
    if rc != 0:
        msg = lib().cg_last_error()
        raise NativeError('%s failed (status %d): %s' % (what, rc, msg.decode() if msg else '?'))


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, or None."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return t.data_ptr()
    return t.ctypes.data
