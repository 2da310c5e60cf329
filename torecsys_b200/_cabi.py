"""ctypes binding of libtorecsys_b200.so (include/torecsys_b200.h).

The library is the product; this module only loads it and declares prototypes.  There is NO fallback: if the
shared library is missing (`python -m torecsys_b200.build` not run) or a tensor is not on a CUDA device, the
callers raise -- they never route to torch/CPU code (see DESIGN.md "no CPU fallback").
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

LIB_NAME = 'libtorecsys_b200.so'
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

TRS_OK = 0
TRS_ERR_INVALID_ARGUMENT = -1
TRS_ERR_UNSUPPORTED = -2
TRS_ERR_CUDA = -3
TRS_STATUS_WORDS = 2
TRS_LAUNCH_OVERLAP_PREVIOUS = 1

ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH = 0, 1, 2, 3
OPN_MAT, OPN_VEC, OPN_NUM = 0, 1, 2

_P = c_void_p            # device / host data pointer
_PP = POINTER(c_void_p)  # host array of device pointers
_IP = POINTER(c_int)     # host int array

# trs_forward_fn: int fn(void* ctx, int lane, const void* idx_dev, int idx_bits, int64_t batch, float* logits_dev,
#                         int32_t* status_dev, void* stream)
FORWARD_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p)

# name -> (restype, argtypes): exactly the declarations of include/torecsys_b200.h
PROTOTYPES = {
    'trs_version': (c_char_p, []),
    'trs_last_error': (c_char_p, []),
    'trs_device_arch': (c_int, []),
    'trs_embedding_gather': (c_int, [_P, c_int64, c_int, _P, c_int, _P, c_int64, c_int, _P, _P, _P]),
    'trs_embedding_gather_field_aware': (c_int, [_P, c_int64, c_int, _P, c_int, _P, c_int64, c_int, _P, _P, _P]),
    'trs_index_concat': (c_int, [_PP, _IP, c_int, c_int, c_int64, _P, _P]),
    'trs_fm_forward': (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    'trs_embedding_grad': (c_int, [_P, _P, c_int, _P, c_int64, c_int, c_int64, c_int, c_int64, _P, _P]),
    'trs_fm_backward': (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P]),
    'trs_embedding_rows': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P]),
    'trs_embedding_grad_segments': (c_int, [_P, _P, _P, c_int64, c_int, _P, _P]),
    'trs_ffm_backward': (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P]),
    'trs_ipn_backward': (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P]),
    'trs_cross_backward': (c_int, [_P, _P, _P, _P, c_int, c_int64, c_int, _P, _P, _P, _P]),
    'trs_ffm_forward': (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    'trs_ipn_forward': (c_int, [_P, c_int64, c_int, c_int, _P, _P]),
    'trs_bilinear_forward': (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int, _P, _P]),
    'trs_bilinear_forward_strided': (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int, c_int64, _P, _P]),
    'trs_bilinear_backward': (c_int, [_P, _P, _P, c_int, c_int64, c_int, c_int, _P, _P, _P, _P]),
    'trs_afm_forward': (c_int, [_P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P]),
    'trs_afm_backward_supported': (c_int, [c_int, c_int]),
    'trs_afm_backward': (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    'trs_opn_forward': (c_int, [_P, _P, c_int, c_int64, c_int, c_int, _P, _P]),
    'trs_senet_workspace_bytes': (c_int64, [c_int64, c_int]),
    'trs_senet_forward': (c_int, [_P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, c_int, _P, _P, c_int64, _P]),
    'trs_cross_forward': (c_int, [_P, _P, _P, c_int, c_int64, c_int, _P, _P]),
    'trs_cross_forward_tc5': (c_int, [_P, _P, _P, c_int, c_int64, c_int, _P, _P]),
    'trs_cin_workspace_bytes': (c_int64, [c_int64, c_int, c_int, _IP, c_int, c_int]),
    'trs_cin_forward': (c_int, [_P, _PP, _PP, _PP, _IP, c_int, c_int, c_int, _P, _P, c_int, c_int64, c_int, c_int,
                                _P, _P, c_int64, _P]),
    'trs_mlp_forward': (c_int, [_P, c_int64, _IP, c_int, _PP, _PP, c_int, _P, _P]),
    'trs_fm_model_forward': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _P, _P, _P, _P]),
    'trs_deepfm_forward': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _IP, c_int, _PP, _PP,
                                   c_int, _P, _P, _P]),
    'trs_fm_pack_table': (c_int, [_P, _P, c_int64, c_int, _P, _P]),
    'trs_fm_model_forward_packed': (c_int, [_P, c_int, _P, c_int64, c_int, _P, c_int64, _P, _P, _P, _P]),
    'trs_deepfm_forward_packed': (c_int, [_P, c_int, _P, c_int64, c_int, _P, c_int64, _IP, c_int, _PP, _PP, c_int,
                                          _P, _P, _P]),
    'trs_deepfm_forward_packed_ex': (c_int, [_P, c_int, _P, c_int64, c_int, _P, c_int64, _IP, c_int, _PP, _PP, c_int,
                                             _P, _P, ctypes.c_uint, _P]),
    'trs_deepfm_packed_wide_supported': (c_int, [c_int, _IP, c_int, c_int64]),
    'trs_deepfm_tc_workspace_bytes': (c_int64, [c_int, c_int]),
    'trs_deepfm_tc_supported': (c_int, [c_int, c_int, _IP, c_int, c_int, c_int64, c_int]),
    'trs_deepfm_forward_tc_sharded': (c_int, [_P, c_int, _P, c_int64, c_int, _PP, c_int, c_int64, _IP, c_int, _PP, _PP,
                                              c_int, _P, c_int, _P, _P, ctypes.c_uint, _P]),
    'trs_debug_tc5_trace': (c_int, [_P]),
    'trs_deepfm_tc_prepare': (c_int, [c_int, _P, c_int, _P, _P]),
    'trs_deepfm_forward_tc': (c_int, [_P, c_int, _P, c_int64, c_int, _P, c_int64, _IP, c_int, _PP, _PP, c_int, _P, c_int,
                                      _P, _P, ctypes.c_uint, _P]),
    'trs_dcn_forward': (c_int, [_P, c_int, _P, c_int64, c_int, _P, c_int64, c_int, _P, _P, c_int, _IP, c_int, _PP,
                                _PP, c_int, _P, _P, _P, _P, _P]),
    'trs_xdeepfm_workspace_bytes': (c_int64, [c_int64, c_int, c_int, _IP, c_int, c_int]),
    'trs_xdeepfm_forward': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _PP, _PP, _PP, _IP, c_int,
                                    c_int, c_int, _P, _P, _IP, c_int, _PP, _PP, c_int, _P, _P, _P, c_int64, _P, _P]),
    'trs_ffm_model_forward': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _P, _P, _P, _P]),
    'trs_ffm_interleaved_pitch': (c_int64, [c_int, c_int]),
    'trs_ffm_pack_tables': (c_int, [_P, _P, c_int64, c_int, c_int, _P, _P]),
    'trs_ffm_model_forward_interleaved': (c_int, [_P, c_int, _P, c_int64, c_int, _P, c_int64, c_int, _P, _P, _P, _P]),
    'trs_ffm_model_forward_pairs': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _P, _P, c_int,
                                            c_int64, c_int64, _P, _P, _P]),
    'trs_senet_backward_supported': (c_int, [c_int, c_int]),
    'trs_senet_backward': (c_int, [_P, _P, _P, _P, _P, c_int, _P, c_int64, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    'trs_mlp_backward_supported': (c_int, [_IP, c_int]),
    'trs_mlp_backward': (c_int, [_P, c_int64, _IP, c_int, _PP, _PP, c_int, _P, _P, _PP, _PP, _P]),
    'trs_ffm_shard_plan': (c_int, [c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, _IP, _IP, _IP, _IP]),
    'trs_ffm_shard_pack': (c_int, [_P, c_int, c_int, c_int64, c_int, _P, _P]),
    'trs_ffm_shard_resolve': (c_int, [_P, c_int, _P, c_int64, c_int, c_int64, _P, _P, _P, _P, _P, _P]),
    'trs_ffm_shard_blocks': (c_int, [_P, c_int64, c_int, c_int, _PP, c_int, c_int, _P, c_int, _P, c_int, _P, c_int64,
                                     c_int64, _P, _P]),
    'trs_nfm_forward': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _IP, c_int, _PP, _PP, c_int, _P,
                                _P, _P, _P]),
    'trs_fnn_forward': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _IP, c_int, _PP, _PP, c_int, _P,
                                _P, _P, _P]),
    'trs_pnn_inner_forward': (c_int, [_P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _IP, c_int, _PP, _PP,
                                      c_int, _P, _P, _P, _P]),
    'trs_session_create': (c_int, [c_int64, c_int, c_int, POINTER(c_void_p)]),
    'trs_session_destroy': (c_int, [c_void_p]),
    'trs_session_depth': (c_int, []),
    'trs_session_set_index_narrowing': (c_int, [c_void_p, c_int]),
    'trs_host_narrow_indices': (c_int, [_P, _P, c_int64, c_int]),
    'trs_host_narrow_pool_ns': (c_int64, [_P, _P, c_int64, c_int, c_int]),
    'trs_session_submit_deepfm': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _IP,
                                          c_int, _PP, _PP, c_int, _P, POINTER(c_int64)]),
    'trs_session_submit_deepfm_packed': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, c_int64, _IP,
                                                 c_int, _PP, _PP, c_int, _P, POINTER(c_int64)]),
    'trs_session_set_producer_stream': (c_int, [c_void_p, _P, c_int]),
    'trs_session_lanes': (c_int, []),
    'trs_session_submit_fn': (c_int, [c_void_p, _P, c_int, c_int64, c_int, FORWARD_FN, c_void_p, c_int64, _P,
                                      POINTER(c_int64)]),
    'trs_session_submit_deepfm_tc': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, c_int64, _IP, c_int, _PP, _PP,
                                             c_int, _P, c_int, _P, POINTER(c_int64)]),
    'trs_session_submit_fm': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, _P, _P, c_int64, c_int, _P, _P,
                                      POINTER(c_int64)]),
    'trs_session_submit_dcn': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, c_int64, c_int, _P, _P, c_int, _IP,
                                       c_int, _PP, _PP, c_int, _P, _P, _P, POINTER(c_int64)]),
    'trs_session_submit_xdeepfm': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _PP, _PP, _PP,
                                           _IP, c_int, c_int, c_int, _P, _P, _IP, c_int, _PP, _PP, c_int, _P, _P, c_int64,
                                           _P, POINTER(c_int64)]),
    'trs_session_submit_ffm': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, _P, _P, c_int64, c_int, _P, _P,
                                       POINTER(c_int64)]),
    'trs_session_wait': (c_int, [c_void_p, c_int64, POINTER(c_int64)]),
    'trs_session_deepfm_forward_host': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, _P, c_int64, c_int, _IP,
                                                c_int, _PP, _PP, c_int, _P, POINTER(c_int64)]),
    'trs_session_deepfm_forward_host_packed': (c_int, [c_void_p, _P, c_int, _P, c_int64, c_int, _P, c_int64, _IP,
                                                       c_int, _PP, _PP, c_int, _P, POINTER(c_int64)]),
}

_lib = None


class LibraryNotBuiltError(RuntimeError):
    pass


def load():
    """Loads the shared library (once).  Raises LibraryNotBuiltError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryNotBuiltError(
            f'{LIB_PATH} is missing: build it with `python -m torecsys_b200.build` (nvcc, sm_100a). '
            'torecsys_b200 has no CPU or PyTorch fallback.')
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def last_error() -> str:
    return load().trs_last_error().decode()


def check(rc: int, what: str):
    """Maps a TRS_* return code to the Python exception the reference user would see."""
    if rc == TRS_OK:
        return
    msg = f'{what}: {last_error()}'
    if rc == TRS_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if rc == TRS_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def ptr_array(ptrs):
    """Host array of device pointers (ctypes keeps it alive as long as the returned object lives)."""
    arr = (c_void_p * max(len(ptrs), 1))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def int_array(vals):
    arr = (c_int * max(len(vals), 1))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr
