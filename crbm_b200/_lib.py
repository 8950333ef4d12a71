"""ctypes binding of libcrbm_b200.so — the C ABI declared in include/crbm_b200.h.

PyTorch is used by the callers only for device memory, streams and torch.distributed; every argument that
crosses this boundary is a raw device pointer (tensor.data_ptr()), an int/float, or a cudaStream_t handle.
There is no CPU fallback: if the library is missing, `load()` raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcrbm_b200.so')

c_void_p, c_int, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
c_u64, c_size_t = ctypes.c_uint64, ctypes.c_size_t

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, -1, -2, -3
FANTASY_LEN = 200
FANTASY_WORDS = 208
MAX_MOTIF_LEN = 33
MAX_POOLING = 16
STRAND_FWD, STRAND_RC, STRAND_SUM = 0, 1, 2
OUT_ACTIVITY, OUT_PROBABILITY = 0, 1
PHASE_NEGATIVE, PHASE_POSITIVE, PHASE_REDUCE, PHASE_APPLY = 1, 2, 4, 8


class TrainArgs(ctypes.Structure):
    """struct crbm_train_args (include/crbm_b200.h)."""
    _fields_ = [
        ('W', c_void_p), ('b', c_void_p), ('c', c_void_p),
        ('vW', c_void_p), ('vb', c_void_p), ('vc', c_void_p),
        ('K', c_int), ('M', c_int), ('pooling', c_int), ('doublestranded', c_int), ('cd_k', c_int),
        ('learning_rate', c_float), ('momentum', c_float), ('rho', c_float), ('lambda_rate', c_float),
        ('data', c_void_p), ('data_row_words', c_int), ('n_data', c_int), ('L', c_int),
        ('batchsize', c_int),
        ('fantasy_bits', c_void_p), ('v_model', c_void_p), ('v_row_words', c_int),
        ('seed', c_u64), ('step', c_u64),
        ('u_v', c_void_p), ('u_h', c_void_p), ('u_hp', c_void_p),
        ('step_dev', c_void_p), ('batch_index_dev', c_void_p), ('n_batches', c_int),
        ('stats', c_void_p), ('workspace', c_void_p), ('workspace_bytes', c_size_t),
        ('c_trace', c_void_p), ('c_trace_slots', c_int),
    ]


MAX_PEERS = 8


class P2PArgs(ctypes.Structure):
    """struct crbm_p2p_args."""
    _fields_ = [('n_ranks', c_int), ('rank', c_int), ('parity', c_int), ('sequence', ctypes.c_ulonglong),
                ('peer_stats', c_void_p * MAX_PEERS), ('peer_flags', c_void_p * MAX_PEERS), ('stats_stride', c_size_t),
                ('error_flag', c_void_p), ('ticket', c_void_p), ('peer_ll', c_void_p * MAX_PEERS), ('ll_parity_stride', c_size_t)]


class StatsLayout(ctypes.Structure):
    """struct crbm_stats_layout."""
    _fields_ = [(n, c_int) for n in (
        'vh_data', 'vh_model', 'vh_q', 'h_data', 'h_model', 'h_q', 'h_fwd', 'v_data', 'v_model',
        'n_hidden_data', 'n_hidden_model', 'n_vis_data', 'n_vis_model', 'total')]


# name -> (restype, argtypes); every symbol include/crbm_b200.h declares
SIGNATURES = {
    'crbm_abi_version': (c_int, []),
    'crbm_last_error': (ctypes.c_char_p, []),
    'crbm_packed_row_words': (c_int, [c_int]),
    'crbm_hidden_words': (c_int, [c_int]),
    'crbm_fantasy_words': (c_int, []),
    'crbm_pack_onehot': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'crbm_pack_onehot_accumulate': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'crbm_pack_codes': (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'crbm_unpack_onehot': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'crbm_random_bases': (c_int, [c_void_p, c_int, c_int, c_int, c_u64, c_void_p]),
    'crbm_v2h_dense': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p]),
    'crbm_v2h_dense_strided': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, ctypes.c_longlong, c_void_p]),
    'crbm_scan_reduce': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    'crbm_position_sum': (c_int, [c_void_p, c_int, c_int, c_int, ctypes.c_longlong, c_void_p, c_void_p]),
    'crbm_pool_probability': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'crbm_sample_hidden': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_u64, c_u64, c_void_p, c_void_p]),
    'crbm_h2v_dense': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    'crbm_visible_softmax_sample': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_u64, c_u64, c_void_p, c_void_p]),
    'crbm_free_energy': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_u64, c_u64, c_void_p]),
    'crbm_train_stats_floats': (c_size_t, [c_int, c_int]),
    'crbm_train_workspace_bytes': (c_size_t, [c_int, c_int]),
    'crbm_train_workspace_bytes_for': (c_size_t, [c_int, c_int, c_int]),
    'crbm_train_accumulate': (c_int, [ctypes.POINTER(TrainArgs), c_void_p]),
    'crbm_train_apply': (c_int, [ctypes.POINTER(TrainArgs), c_void_p]),
    'crbm_train_step': (c_int, [ctypes.POINTER(TrainArgs), c_void_p]),
    'crbm_train_phases': (c_int, [ctypes.POINTER(TrainArgs), c_int, c_void_p]),
    'crbm_train_step_hostfed': (c_int, [ctypes.POINTER(TrainArgs), ctypes.POINTER(P2PArgs), c_void_p, c_void_p, c_void_p]),
    'crbm_train_step_staged': (c_int, [ctypes.POINTER(TrainArgs), ctypes.POINTER(P2PArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'crbm_train_step_onehot': (c_int, [ctypes.POINTER(TrainArgs), ctypes.POINTER(P2PArgs), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'crbm_train_step_p2p': (c_int, [ctypes.POINTER(TrainArgs), ctypes.POINTER(P2PArgs), c_void_p]),
    'crbm_p2p_ll_elems': (c_size_t, [c_int, c_int, c_int]),
    'crbm_train_accumulate_p2p': (c_int, [ctypes.POINTER(TrainArgs), ctypes.POINTER(P2PArgs), c_void_p]),
    'crbm_train_apply_p2p': (c_int, [ctypes.POINTER(TrainArgs), ctypes.POINTER(P2PArgs), c_void_p]),
    'crbm_train_stats_layout': (c_int, [c_int, c_int, ctypes.POINTER(StatsLayout)]),
    'crbm_hidden_bits_to_dense': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'crbm_hidden_dense_to_bits': (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    'crbm_chain_layout': (c_int, [c_int, c_int, c_int]),
    'crbm_chain_words': (c_size_t, [c_int, c_int, c_int]),
    'crbm_chain_to_dense': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'crbm_chain_from_dense': (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    'crbm_fasta_parse_block_host': (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_longlong),
                                            ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(c_int)]),
    'crbm_fasta_pack_block_host': (c_int, [c_void_p, c_size_t, c_void_p, c_int, c_size_t, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_longlong),
                                           ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(c_int), c_int]),
    'crbm_pack_onehot_host': (c_int, [c_void_p, ctypes.c_longlong, c_int, ctypes.c_longlong, ctypes.c_longlong, c_void_p, c_int,
                                      ctypes.POINTER(ctypes.c_longlong), c_int]),
    'crbm_v2h_dense_kernel_name': (ctypes.c_char_p, [c_int, c_int, c_int, c_int, ctypes.c_longlong]),
    'crbm_launch_count': (ctypes.c_ulonglong, []),
}

_lib = None


class CrbmError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach the prototypes. Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CrbmError("crbm_b200: %s is missing — build it with `python -m crbm_b200.build` "
                            "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, what=''):
    if rc != OK:
        msg = load().crbm_last_error().decode('utf-8', 'replace')
        if rc == ERR_UNSUPPORTED:
            raise NotImplementedError("crbm_b200 %s: %s" % (what, msg))
        if rc == ERR_ARG:
            raise ValueError("crbm_b200 %s: %s" % (what, msg))
        raise CrbmError("crbm_b200 %s failed (%d): %s" % (what, rc, msg))


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_handle():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
