"""ctypes binding of ``libdcase_b200.so`` (the C ABI declared in ``include/dcase_b200.h``).

There is no CPU fallback: if the library is missing, or no sm_100 GPU is present when a context is
requested, the product path raises.  ``__graft_entry__.build()`` (or ``make -C dcase2019_task4_b200/csrc``)
produces the library in-tree.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdcase_b200.so")

FLAG_BN_BATCH_STATS = 1
FLAG_DROPOUT = 2

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_f = ctypes.c_float
c_u64 = ctypes.c_uint64
c_u32 = ctypes.c_uint32
c_sz = ctypes.c_size_t


class StepScalars(ctypes.Structure):
    """Mirror of ``dcase_step_scalars`` (include/dcase_b200.h)."""
    _fields_ = [("seed", c_u64), ("step", c_u32), ("cons_weight", c_f), ("ema_alpha", c_f), ("lr", c_f),
                ("bias_corr1", c_f), ("bias_corr2", c_f), ("grad_scale", c_f), ("pad_", c_f)]


class MtArgs(ctypes.Structure):
    """Mirror of ``dcase_mt_args`` (include/dcase_b200.h)."""
    _fields_ = [("x_student", c_p), ("x_teacher", c_p), ("target", c_p),
                ("B", c_i), ("T", c_i), ("n_class", c_i),
                ("weak_lo", c_i), ("weak_hi", c_i), ("strong_lo", c_i), ("strong_hi", c_i),
                ("params_s", c_p), ("params_t", c_p), ("bn_s", c_p), ("bn_t", c_p),
                ("flags", c_i), ("seed", c_u64), ("step", c_u32), ("cons_weight", c_f), ("scalars", c_p),
                ("strong_s", c_p), ("weak_s", c_p), ("strong_t", c_p), ("weak_t", c_p), ("meters", c_p),
                ("d_strong", c_p), ("d_weak", c_p), ("ws_s", c_p), ("ws_t", c_p), ("grads", c_p),
                ("after_forward_event", c_p), ("mom_s", c_p), ("mom_t", c_p)]


# name -> (restype, argtypes); every symbol include/dcase_b200.h declares
SIGNATURES = {
    "dcase_version": (c_i, []),
    "dcase_last_error": (ctypes.c_char_p, []),
    "dcase_ctx_create": (c_i, [ctypes.POINTER(c_p), c_i]),
    "dcase_ctx_destroy": (c_i, [c_p]),
    "dcase_launch_count": (ctypes.c_ulonglong, []),
    "dcase_profile_begin": (c_i, []),
    "dcase_profile_end": (c_i, [ctypes.c_char_p, c_sz]),
    "dcase_profile_timeline_begin": (c_i, []),
    "dcase_profile_timeline_end": (c_i, [ctypes.c_char_p, c_sz]),
    "dcase_selftest_umma": (c_i, [c_p, c_i, c_p, c_p, c_p, c_p]),
    "dcase_logmel_num_frames": (c_i, [c_i]),
    "dcase_mel_filterbank": (c_i, [c_p, c_p]),
    "dcase_logmel_fwd": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p]),
    "dcase_logmel_fwd_pcm16": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p]),
    "dcase_audio_mixdown": (c_i, [c_p, c_p, c_i, ctypes.c_longlong, c_i, c_p, c_p]),
    "dcase_audio_resample_len": (ctypes.c_longlong, [ctypes.c_longlong, c_i, c_i]),
    "dcase_audio_resample": (c_i, [c_p, c_p, ctypes.c_longlong, c_i, c_i, c_p, c_p]),
    "dcase_audio_resample_clock": (c_i, [ctypes.c_longlong, c_i, c_i, c_p]),
    "dcase_logmel_finish": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_u64, c_u32, c_p, c_p, c_p, c_p, c_p]),
    "dcase_scaler_accumulate": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "dcase_scaler_finalize": (c_i, [c_p, c_p, ctypes.c_longlong, c_p, c_p, c_p, c_p, c_p]),
    "dcase_crnn_param_count": (c_sz, [c_i]),
    "dcase_crnn_param_offset": (ctypes.c_longlong, [c_i, ctypes.c_char_p]),
    "dcase_crnn_workspace_bytes": (c_sz, [c_i, c_i, c_i]),
    "dcase_crnn_ws_tensor": (c_i, [c_i, c_i, c_i, ctypes.c_char_p, ctypes.POINTER(c_sz), ctypes.POINTER(c_sz)]),
    "dcase_cnn0_input_moments": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p]),
    "dcase_crnn_forward": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_p, c_i, c_u64, c_u32, c_i, c_p, c_p, c_p, c_p, c_p]),
    "dcase_bigru_workspace_bytes": (c_sz, [c_i, c_i]),
    "dcase_bigru_forward": (c_i, [c_p, c_p, c_i, c_i, c_p, c_p, c_p, c_p]),
    "dcase_bigru_param_count_h": (c_sz, [c_i, c_i]),
    "dcase_bigru_workspace_bytes_h": (c_sz, [c_i, c_i, c_i]),
    "dcase_bigru_forward_h": (c_i, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "dcase_crnn_backward": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p, c_i, c_u64, c_u32, c_i, c_p, c_p, c_p, c_p, c_p,
                                  c_p, c_p]),
    "dcase_mt_loss": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_p, c_p,
                            c_p, c_p]),
    "dcase_adam_ema_step": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_sz, c_f, c_f, c_f, c_f, c_i, c_f, c_f, c_p, c_p]),
    "dcase_p2p_handle_bytes": (c_i, []),
    "dcase_p2p_create": (c_i, [c_p, c_i, c_i, c_sz, ctypes.POINTER(c_p), c_p]),
    "dcase_p2p_connect": (c_i, [c_p, c_p]),
    "dcase_p2p_grads": (c_p, [c_p]),
    "dcase_p2p_begin_step": (c_i, [c_p, c_p]),
    "dcase_p2p_adam_ema_step": (c_i, [c_p, c_p, c_p, c_p, c_p, c_p, c_f, c_f, c_f, c_f, c_i, c_f, c_p, c_p]),
    "dcase_p2p_destroy": (c_i, [c_p]),
    "dcase_syncbn_handle_bytes": (c_i, []),
    "dcase_syncbn_create": (c_i, [c_p, c_i, c_i, ctypes.POINTER(c_p), c_p]),
    "dcase_syncbn_connect": (c_i, [c_p, c_p]),
    "dcase_ctx_set_syncbn": (c_i, [c_p, c_p]),
    "dcase_syncbn_allreduce": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p]),
    "dcase_syncbn_destroy": (c_i, [c_p]),
    "dcase_mt_fwd_bwd": (c_i, [c_p, ctypes.POINTER(MtArgs), c_p]),
    "dcase_sizeof_mt_args": (c_sz, []),
    "dcase_sizeof_step_scalars": (c_sz, []),
    "dcase_selftest_umma_shift": (c_i, [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_p]),
    "dcase_bench_umma": (c_i, [c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p]),
}

_lib = None
_lock = threading.RLock()     # re-entrant: ctx() loads the library (lib()) while holding it
_ctx = {}


class DcaseError(RuntimeError):
    pass


def lib():
    """The loaded shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise DcaseError(
                        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "or `make -C dcase2019_task4_b200/csrc` (there is no CPU fallback)")
                handle = ctypes.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                if handle.dcase_sizeof_mt_args() != ctypes.sizeof(MtArgs) or \
                        handle.dcase_sizeof_step_scalars() != ctypes.sizeof(StepScalars):
                    raise DcaseError("ctypes struct mirrors do not match include/dcase_b200.h as compiled")
                _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise DcaseError(f"dcase_b200 error {rc}: {lib().dcase_last_error().decode(errors='replace')}")


def ctx(device=None):
    """Per-device context handle (constant tables on the GPU); raises without an sm_100 GPU."""
    import torch
    if not torch.cuda.is_available():
        raise DcaseError("dcase_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.cuda.current_device()
    device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    handle = lib()                       # load (and lock) before taking the context lock
    if idx not in _ctx:
        with _lock:
            if idx not in _ctx:
                h = c_p()
                with torch.cuda.device(idx):
                    check(handle.dcase_ctx_create(ctypes.byref(h), idx))
                _ctx[idx] = h
    return _ctx[idx]


def stream_ptr():
    import torch
    return c_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a contiguous tensor (None -> NULL)."""
    if t is None:
        return c_p(0)
    assert t.is_contiguous(), "dcase_b200 kernels need contiguous tensors"
    return c_p(t.data_ptr())
