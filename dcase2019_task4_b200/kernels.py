"""Thin torch-tensor wrappers over the C ABI (one function per entry point of include/dcase_b200.h).

Tensors are plumbing only: device memory, streams and (elsewhere) torch.distributed.  All arithmetic
happens in the sm_100a kernels of ``csrc/``.
"""
import ctypes

import torch

from . import _lib
from ._lib import FLAG_BN_BATCH_STATS, FLAG_DROPOUT, check, ctx, lib, ptr, stream_ptr

N_MELS = 64
HOP = 511


def _f32(t):
    assert t.is_cuda and t.dtype == torch.float32, "expected a CUDA float32 tensor"
    return t.contiguous()


def num_frames(n_samples):
    return lib().dcase_logmel_num_frames(int(n_samples))


def mel_filterbank(device=None):
    """Dense float32 [64, 1025] Slaney filterbank used by the kernel (host tensor)."""
    out = torch.empty(64, 1025, dtype=torch.float32)
    check(lib().dcase_mel_filterbank(ctx(device), ctypes.c_void_p(out.data_ptr())))
    return out


def logmel_fwd(wave):
    """wave [B, L] float32 or int16 (CUDA) -> amplitude mel [B, T, 64] (calculate_mel_spec)."""
    assert wave.is_cuda and wave.dim() == 2
    wave = wave.contiguous()
    B, L = wave.shape
    T = num_frames(L)
    out = torch.empty(B, T, N_MELS, device=wave.device, dtype=torch.float32)
    with torch.cuda.device(wave.device):
        if wave.dtype == torch.int16:
            check(lib().dcase_logmel_fwd_pcm16(ctx(wave.device), ptr(wave), B, L, ptr(out), stream_ptr()))
        else:
            check(lib().dcase_logmel_fwd(ctx(wave.device), ptr(_f32(wave)), B, L, ptr(out), stream_ptr()))
    return out


def audio_mixdown(frames):
    """read_audio's mix-down: interleaved [n_frames, n_channels] (CUDA float32 or int16 PCM) -> mono float32 [n_frames]."""
    assert frames.is_cuda and frames.dim() == 2 and frames.dtype in (torch.float32, torch.int16)
    frames = frames.contiguous()
    n, ch = frames.shape
    out = torch.empty(n, device=frames.device, dtype=torch.float32)
    with torch.cuda.device(frames.device):
        check(lib().dcase_audio_mixdown(ctx(frames.device), ptr(frames), int(frames.dtype == torch.int16), n, ch,
                                        ptr(out), stream_ptr()))
    return out


def audio_resample(mono, sr_in, sr_out):
    """read_audio's librosa.resample(audio, orig_sr, target_sr) (kaiser_best): CUDA float32 mono [n] -> [ceil(n * ratio)]."""
    assert mono.is_cuda and mono.dim() == 1
    mono = _f32(mono)
    n_out = lib().dcase_audio_resample_len(mono.numel(), int(sr_in), int(sr_out))
    out = torch.empty(n_out, device=mono.device, dtype=torch.float32)
    with torch.cuda.device(mono.device):
        check(lib().dcase_audio_resample(ctx(mono.device), ptr(mono), mono.numel(), int(sr_in), int(sr_out), ptr(out),
                                         stream_ptr()))
    return out


def logmel_finish(mel_amp, mean, std, frames, noisy=False, noise=None, seed=0, step=0, scalars=None,
                  out_clean=None, out_noisy=None):
    """get_transforms(frames, scaler, augment_type='noise' if noisy) on a batch of amplitude mels.

    Returns clean [B, frames, 64] (and noisy [B, frames, 64] when ``noisy`` or ``noise`` is given)."""
    mel_amp = _f32(mel_amp)
    B, T_in, _ = mel_amp.shape
    dev = mel_amp.device
    mean = _f32(mean)
    std = _f32(std)
    want_noisy = noisy or noise is not None
    clean = out_clean if out_clean is not None else torch.empty(B, frames, N_MELS, device=dev, dtype=torch.float32)
    nz = None
    if want_noisy:
        nz = out_noisy if out_noisy is not None else torch.empty(B, frames, N_MELS, device=dev, dtype=torch.float32)
    scratch = torch.empty(2 * max(B, 1), device=dev, dtype=torch.float32)
    if noise is not None:
        noise = _f32(noise)
        assert noise.shape == mel_amp.shape
    with torch.cuda.device(dev):
        check(lib().dcase_logmel_finish(ctx(dev), ptr(mel_amp), B, T_in, frames, ptr(mean), ptr(std), ptr(noise),
                                        seed, step, ptr(scalars), ptr(scratch), ptr(clean), ptr(nz), stream_ptr()))
    return (clean, nz) if want_noisy else clean


def scaler_accumulate(feats, sums, frames=None, apply_log=True):
    """Scaler.means over one batch: feats [B, T_in, 64] (amplitude mels with apply_log, finished features
    otherwise) are reduced into ``sums`` (CUDA float64 [2, 64], accumulated in place)."""
    feats = _f32(feats)
    B, T_in, F = feats.shape
    assert F == N_MELS and sums.is_cuda and sums.dtype == torch.float64 and sums.numel() == 2 * N_MELS
    assert sums.is_contiguous()
    frames = T_in if frames is None else int(frames)
    dev = feats.device
    scratch = torch.empty(max(B, 1), device=dev, dtype=torch.float32) if apply_log else None
    with torch.cuda.device(dev):
        check(lib().dcase_scaler_accumulate(ctx(dev), ptr(feats), B, T_in, frames, int(bool(apply_log)), ptr(scratch),
                                            ptr(sums), stream_ptr()))
    return sums


def scaler_finalize(sums, n_samples):
    """-> (mean_ f64 [64], mean_of_square_ f64 [64], mean f32 [64], std f32 [64]) device tensors."""
    dev = sums.device
    mean = torch.empty(N_MELS, device=dev, dtype=torch.float64)
    msq = torch.empty(N_MELS, device=dev, dtype=torch.float64)
    mean32 = torch.empty(N_MELS, device=dev, dtype=torch.float32)
    std32 = torch.empty(N_MELS, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib().dcase_scaler_finalize(ctx(dev), ptr(sums), int(n_samples), ptr(mean), ptr(msq), ptr(mean32),
                                          ptr(std32), stream_ptr()))
    return mean, msq, mean32, std32


def param_count(n_class=10):
    return lib().dcase_crnn_param_count(n_class)


def param_offset(name, n_class=10):
    off = lib().dcase_crnn_param_offset(n_class, name.encode())
    if off < 0:
        raise KeyError(name)
    return off


def workspace_bytes(B, T, n_class=10):
    return lib().dcase_crnn_workspace_bytes(B, T, n_class)


def new_workspace(B, T, n_class, device):
    return torch.empty(workspace_bytes(B, T, n_class), dtype=torch.uint8, device=device)


def ws_tensor(ws, B, T, n_class, name):
    """float32 view of a named intermediate inside a workspace (tests / debugging)."""
    off = ctypes.c_size_t()
    n = ctypes.c_size_t()
    check(lib().dcase_crnn_ws_tensor(B, T, n_class, name.encode(), ctypes.byref(off), ctypes.byref(n)))
    return ws[off.value: off.value + 4 * n.value].view(torch.float32)


def crnn_forward(x, params, bn_running, flags, ws, n_class=10, seed=0, step=0, model_id=0, scalars=None,
                 strong=None, weak=None):
    """x [B, T, 64] (or [B, 1, T, 64]) -> strong [B, T/8, n_class], weak [B, n_class]."""
    x = _f32(x)
    if x.dim() == 4:
        assert x.shape[1] == 1
        x = x[:, 0]
    B, T, F = x.shape
    assert F == N_MELS
    dev = x.device
    if strong is None:
        strong = torch.empty(B, T // 8, n_class, device=dev, dtype=torch.float32)
    if weak is None:
        weak = torch.empty(B, n_class, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib().dcase_crnn_forward(ctx(dev), ptr(x), B, T, n_class, ptr(params), ptr(bn_running), flags, seed, step,
                                       model_id, ptr(scalars), ptr(strong), ptr(weak), ptr(ws), stream_ptr()))
    return strong, weak


def cnn0_input_moments(x, out=None):
    """The 54 input moments behind block 0's BatchNorm batch statistics (float64 [56], see include/dcase_b200.h)."""
    x = _f32(x)
    B, T = x.shape[0], x.shape[-2]
    if out is None:
        out = torch.empty(56, dtype=torch.float64, device=x.device)
    check(lib().dcase_cnn0_input_moments(ctx(), ptr(x), B, T, ptr(out), stream_ptr()))
    return out


def crnn_backward(x, params, flags, ws, d_strong, d_weak, weak, n_class=10, seed=0, step=0, model_id=0, scalars=None,
                  grads=None):
    x = _f32(x)
    if x.dim() == 4:
        x = x[:, 0]
    B, T, _ = x.shape
    dev = x.device
    if grads is None:
        grads = torch.empty(param_count(n_class), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib().dcase_crnn_backward(ctx(dev), ptr(x), B, T, n_class, ptr(params), flags, seed, step, model_id,
                                        ptr(scalars), ptr(_f32(d_strong)), ptr(_f32(d_weak)), ptr(_f32(weak)), ptr(ws),
                                        ptr(grads), stream_ptr()))
    return grads


GRU_PARAM_COUNT = 124416      # the 16 rnn.rnn.* tensors of the 2-layer bidirectional GRU(64, 64)


def bigru_forward(x, rnn_params, out=None, ws=None):
    """BidirectionalGRU.forward: x [B, To, 64] -> [B, To, 128]; rnn_params is the GRU block of the flat slab
    (``flat[param_offset('rnn.rnn.weight_ih_l0'):][:GRU_PARAM_COUNT]``)."""
    x = _f32(x)
    B, To, F = x.shape
    assert F == N_MELS and rnn_params.numel() >= GRU_PARAM_COUNT
    dev = x.device
    if out is None:
        out = torch.empty(B, To, 128, device=dev, dtype=torch.float32)
    if ws is None:
        ws = torch.empty(lib().dcase_bigru_workspace_bytes(B, To), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().dcase_bigru_forward(ctx(dev), ptr(x), B, To, ptr(_f32(rnn_params)), ptr(out), ptr(ws), stream_ptr()))
    return out


def bigru_forward_h(x, rnn_params, hidden, out=None, ws=None):
    """EXPERIMENTAL cluster BiGRU for hidden 128 / 256 (csrc/gru_cluster.cu; not yet run on hardware):
    x [B, To, n_in] -> [B, To, 2 * hidden]; rnn_params = nn.GRU's 16 tensors flattened in named_parameters() order."""
    x = _f32(x)
    B, To, n_in = x.shape
    assert rnn_params.numel() == lib().dcase_bigru_param_count_h(n_in, hidden)
    dev = x.device
    if out is None:
        out = torch.empty(B, To, 2 * hidden, device=dev, dtype=torch.float32)
    if ws is None:
        ws = torch.empty(lib().dcase_bigru_workspace_bytes_h(B, To, hidden), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib().dcase_bigru_forward_h(ctx(dev), ptr(x), B, To, n_in, hidden, ptr(_f32(rnn_params)), ptr(out), ptr(ws),
                                          stream_ptr()))
    return out


def _slice_bounds(mask, B):
    if mask is None:
        return 0, 0
    if isinstance(mask, slice):
        lo, hi, st = mask.indices(B)
        assert st == 1, "masks are contiguous slices (main.py:240-247)"
        return lo, max(hi, lo)
    raise TypeError("weak_mask / strong_mask must be slices or None (main.py:240-247)")


def mt_loss(strong_s, weak_s, strong_t, weak_t, target, weak_mask, strong_mask, cons_weight=0.0, scalars=None):
    """Loss terms of main.py:95-145.  Returns (meters[8] device tensor, d_strong, d_weak)."""
    strong_s = _f32(strong_s)
    B, To, NC = strong_s.shape
    dev = strong_s.device
    wl, wh = _slice_bounds(weak_mask, B)
    sl, sh = _slice_bounds(strong_mask, B)
    meters = torch.empty(8, device=dev, dtype=torch.float32)
    d_strong = torch.empty_like(strong_s)
    d_weak = torch.empty(B, NC, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(lib().dcase_mt_loss(ctx(dev), ptr(strong_s), ptr(_f32(weak_s)),
                                  ptr(None if strong_t is None else _f32(strong_t)),
                                  ptr(None if weak_t is None else _f32(weak_t)), ptr(_f32(target)), B, To, NC,
                                  wl, wh, sl, sh, float(cons_weight), ptr(scalars), ptr(meters), ptr(d_strong),
                                  ptr(d_weak), stream_ptr()))
    return meters, d_strong, d_weak


def adam_ema_step(p, g, m, v, p_ema, step_t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, ema_alpha=0.999,
                  grad_scale=1.0, scalars=None):
    """In-place Adam on flat slabs followed by the teacher EMA (p_ema may be None)."""
    dev = p.device
    n = p.numel()
    with torch.cuda.device(dev):
        check(lib().dcase_adam_ema_step(ctx(dev), ptr(p), ptr(g), ptr(m), ptr(v), ptr(p_ema), n, lr, beta1, beta2, eps,
                                        int(step_t), float(ema_alpha), float(grad_scale), ptr(scalars), stream_ptr()))


def launch_count():
    return int(lib().dcase_launch_count())


def profile_begin():
    check(lib().dcase_profile_begin())


def profile_end():
    """-> {kernel name: (launch count, total ms)} measured with CUDA events on the launching stream."""
    buf = ctypes.create_string_buffer(1 << 16)
    check(lib().dcase_profile_end(buf, len(buf)))
    out = {}
    for line in buf.value.decode().splitlines():
        name, cnt, ms = line.split(",")
        out[name] = (int(cnt), float(ms))
    return out


def mt_fwd_bwd(args):
    """One call for teacher fwd + student fwd + losses + student bwd (``args`` is a filled _lib.MtArgs)."""
    dev = torch.cuda.current_device()
    check(lib().dcase_mt_fwd_bwd(ctx(dev), ctypes.byref(args), stream_ptr()))


__all__ = ["FLAG_BN_BATCH_STATS", "FLAG_DROPOUT", "logmel_fwd", "logmel_finish", "crnn_forward", "crnn_backward",
           "mt_loss", "adam_ema_step", "mt_fwd_bwd", "param_count", "param_offset", "new_workspace", "ws_tensor",
           "mel_filterbank", "num_frames", "workspace_bytes", "scaler_accumulate", "scaler_finalize", "bigru_forward", "bigru_forward_h", "audio_mixdown", "audio_resample"]
